"""Generates tests/golden/deform_tvm_lowband.npz: DeformableConvolution outputs of TVM's independent pure-python reference
(`tvm/topi/testing/deformable_conv2d_python.py`, vendored with tilelang in this image; loaded from its file with a stub for its one
tvm import) on seeded inputs whose taps reach the LOWER border band h in (-1, 0) / w in (-1, 0) and beyond, but never the
upper band (every tap position stays <= H - 1.01, W - 1.01).

Why: the MXNet-1.5 border rule of oracle/mfn_oracle.c has two parts that differ from the DCNv2 / torchvision zero-corner
rule -- (a) a tap with coordinate in (-1, 0) contributes ZERO (MXNet's `h_im >= 0 && w_im >= 0 && ...` test before the
bilinear read), (b) a tap in (n-1, n) collapses on the last pixel.  TVM's reference implements (a) exactly like MXNet's
im2col test (`if y < 0 or ...: continue`) and differs from MXNet in (b); on inputs that exercise only (a) it is an
independent pin of that half of the rule (torchvision pins the zero-corner mode and the interior).

Also written: affine_sampler_tvm.npz -- TVM's `affine_grid_python` (the operator TVM ported from MXNet's
GridGenerator('affine')) and `grid_sample_2d(bilinear, zeros, align_corners=True)` (= MXNet BilinearSampler semantics) on seeded
inputs with a grid that partly leaves the image: independent pins of oracle/augment_ref.grid_generator_affine and of the C
oracle's sampler outside the 'warp' use (row N4's operators).

Run from the repo root (needs only numpy):  python tests/golden/make_golden_deform_tvm.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
TVM_DEFORM = ("/opt/prime-rl/.venv/lib/python3.12/site-packages/tilelang/3rdparty/tvm/python/tvm/topi/testing/"
              "deformable_conv2d_python.py")


def load_tvm_reference():
    for name in ("tvm", "tvm.topi", "tvm.topi.nn", "tvm.topi.nn.utils"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["tvm.topi.nn.utils"].get_pad_tuple = lambda padding, kernel: (padding, padding, padding, padding)   # int padding only
    spec = importlib.util.spec_from_file_location("_tvm_deformable_conv2d_python", TVM_DEFORM)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.deformable_conv2d_nchw_python


def main():
    from oracle import cref
    tvm_deform = load_tvm_reference()
    rng = np.random.default_rng(20260924)
    N, C, H, W, F = 2, 4, 7, 9, 5
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((F, C, 3, 3)) * 0.3).astype(np.float32)
    off = rng.uniform(-2.6, 0.4, (N, 18, H, W)).astype(np.float32)         # (tap, (y, x)) pairs, MXNet / TVM layout
    # keep every tap position at or below (H - 1.01, W - 1.01): only the lower bands and the far outside are exercised
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    for k in range(9):
        kh, kw = divmod(k, 3)
        off[:, 2 * k] = np.minimum(off[:, 2 * k], (H - 1.01) - (ys - 1 + kh))
        off[:, 2 * k + 1] = np.minimum(off[:, 2 * k + 1], (W - 1.01) - (xs - 1 + kw))
    out = tvm_deform(x, off, w, 1, 1, 1, 1, 1).astype(np.float32)
    mx15 = cref.deformable_conv_forward(x, off, w, None, border_mode=0)
    zc = cref.deformable_conv_forward(x, off, w, None, border_mode=1)
    band = 0
    for k in range(9):
        kh, kw = divmod(k, 3)
        py, px = ys - 1 + kh + off[:, 2 * k], xs - 1 + kw + off[:, 2 * k + 1]
        band += int((((py > -1) & (py < 0)) | ((px > -1) & (px < 0))).sum())
    print("taps in a lower band:", band, " max |oracle(MXNet-1.5) - TVM| =", np.abs(mx15 - out).max(),
          " max |oracle(zero-corner) - TVM| =", np.abs(zc - out).max())
    assert np.abs(mx15 - out).max() < 2e-5 and np.abs(zc - out).max() > 1e-2 and band > 100
    np.savez_compressed(os.path.join(HERE, "deform_tvm_lowband.npz"), x=x, w=w, off=off, out=out)
    print("wrote", os.path.join(HERE, "deform_tvm_lowband.npz"))
    # ---- GridGenerator('affine') + BilinearSampler against TVM's affine_grid_python / grid_sample_2d ----
    from oracle import augment_ref
    spec = importlib.util.spec_from_file_location("_tvm_grid_sample_python", os.path.join(os.path.dirname(TVM_DEFORM), "grid_sample_python.py"))
    gs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gs)
    TH, TW = 11, 13
    theta = np.array([[0.9, 0.2, 0.05, -0.15, 1.1, -0.1], [1.3, -0.4, 0.3, 0.35, 0.8, 0.2], [1, 0, 0, 0, 1, 0]], np.float32)
    data = rng.standard_normal((3, 4, 9, 10)).astype(np.float32)
    grid = gs.affine_grid_python(theta.astype(np.float64).reshape(3, 2, 3), (TH, TW)).astype(np.float32)
    sampled = gs.grid_sample_2d(data.astype(np.float64), grid.astype(np.float64), "bilinear", "NCHW", "zeros", True).astype(np.float32)
    ours_grid = augment_ref.grid_generator_affine(theta, TH, TW)
    ours = cref.bilinear_sampler(data, ours_grid)
    print("affine grid max |oracle - TVM| =", np.abs(ours_grid - grid).max(), " sampler max |oracle - TVM| =", np.abs(ours - sampled).max(),
          " fraction of grid points outside [-1,1]:", float((np.abs(grid) > 1).any(axis=1).mean()))
    assert np.abs(ours_grid - grid).max() < 1e-6 and np.abs(ours - sampled).max() < 2e-5 and (np.abs(grid) > 1).any()
    np.savez_compressed(os.path.join(HERE, "affine_sampler_tvm.npz"), theta=theta, data=data, grid=grid, sampled=sampled)
    print("wrote", os.path.join(HERE, "affine_sampler_tvm.npz"))


if __name__ == "__main__":
    main()
