"""CPU tests of the oracle itself: pinned against the committed golden fixtures (independent implementations, see
tests/golden/make_golden.py), against its own second implementation (torch_ref), and against known answers."""
import os

import numpy as np
import pytest
import torch

from oracle import cref, torch_ref

G = os.path.join(os.path.dirname(__file__), "golden")


def test_correlation_matches_tvm_fixture():
    d = np.load(os.path.join(G, "corr_tvm.npz"))
    assert np.abs(cref.correlation_forward(d["a1"], d["a2"]) - d["out_a_md4"]).max() < 1e-6      # BASELINE config[0]
    assert np.abs(cref.correlation_forward(d["b1"], d["b2"], 2, 1, 2) - d["out_b_md2"]).max() < 1e-6
    assert np.abs(cref.correlation_forward(d["b1"], d["b2"], 3, 3, 2, 1, 1, 0) - d["out_b_k3_sub"]).max() < 1e-6
    t = torch_ref.correlation(torch.from_numpy(d["a1"]), torch.from_numpy(d["a2"]), 4).numpy()
    assert np.abs(t - d["out_a_md4"]).max() < 1e-6


def test_correlation_known_answers():
    one = np.ones((1, 7, 6, 9), np.float32)
    out = cref.correlation_forward(one, one)
    for q in range(81):
        dy, dx = q // 9 - 4, q % 9 - 4
        exp = np.zeros((6, 9), np.float32)
        exp[max(0, -dy):6 - max(0, dy), max(0, -dx):9 - max(0, dx)] = 1
        assert np.abs(out[0, q] - exp).max() < 1e-6
    # impulse: f1 = e(c0, y0, x0), f2 = e(c0, y1, x1)  ->  single non-zero at q(dy,dx), (y0,x0) with value 1/C
    f1 = np.zeros((1, 5, 8, 8), np.float32); f2 = np.zeros_like(f1)
    f1[0, 2, 3, 4] = 1; f2[0, 2, 5, 1] = 1
    out = cref.correlation_forward(f1, f2)
    q = (2 + 4) * 9 + (-3 + 4)
    assert abs(out[0, q, 3, 4] - 0.2) < 1e-7 and np.count_nonzero(out) == 1
    # multithreaded == single-threaded
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((2, 9, 11, 13)).astype(np.float32), rng.standard_normal((2, 9, 11, 13)).astype(np.float32)
    assert np.array_equal(cref.correlation_forward(a, b, threads=1), cref.correlation_forward(a, b, threads=4))


def test_correlation_backward_matches_autograd():
    rng = np.random.default_rng(1)
    for md in (4, 2):
        a = rng.standard_normal((2, 6, 7, 9)).astype(np.float32)
        b = rng.standard_normal((2, 6, 7, 9)).astype(np.float32)
        go = rng.standard_normal((2, (2 * md + 1) ** 2, 7, 9)).astype(np.float32)
        ta, tb = torch.from_numpy(a).requires_grad_(), torch.from_numpy(b).requires_grad_()
        torch_ref.correlation(ta, tb, md).backward(torch.from_numpy(go))
        g1, g2 = cref.correlation_backward(go, a, b, md)
        assert np.abs(g1 - ta.grad.numpy()).max() < 1e-5 and np.abs(g2 - tb.grad.numpy()).max() < 1e-5


def test_deformable_conv_fixture_and_identities():
    d = np.load(os.path.join(G, "deform_tv.npz"))
    x, w, b, off = d["x"], d["w"], d["b"], d["off"]
    zc = cref.deformable_conv_forward(x, off, w, b, border_mode=1)
    assert np.abs(zc - d["out_zero_corner"]).max() < 2e-5                       # torchvision, whole image
    mx15 = cref.deformable_conv_forward(x, off * 0.2, w, b, border_mode=0)
    zc2 = cref.deformable_conv_forward(x, off * 0.2, w, b, border_mode=1)
    assert np.abs(mx15 - zc2)[:, :, 3:-3, 3:-3].max() < 2e-5                    # both rules agree in the interior
    assert np.abs(mx15 - zc2).max() > 1e-3                                      # ... and differ in the border bands
    z = cref.deformable_conv_forward(x, np.zeros_like(off), w, b)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), padding=1).numpy()
    assert np.abs(z - ref).max() < 2e-5                                         # zero offsets == plain convolution
    for mode in (0, 1):                                                         # C oracle == torch restatement
        t = torch_ref.deformable_conv(torch.from_numpy(x), torch.from_numpy(off), torch.from_numpy(w),
                                      torch.from_numpy(b), mode).numpy()
        assert np.abs(t - cref.deformable_conv_forward(x, off, w, b, border_mode=mode)).max() < 2e-5


def test_deformable_lower_band_matches_tvm_reference():
    """MXNet-1.5 border rule, lower half -- a tap whose coordinate lies in (-1, 0) contributes zero -- against TVM's independent
    pure-python reference (tests/golden/make_golden_deform_tvm.py): 221 taps of the fixture sit in such a band, none in the
    upper band (where TVM and MXNet 1.5 differ).  The zero-corner rule must NOT match there."""
    d = np.load(os.path.join(G, "deform_tvm_lowband.npz"))
    mx15 = cref.deformable_conv_forward(d["x"], d["off"], d["w"], None, border_mode=0)
    assert np.abs(mx15 - d["out"]).max() < 2e-5
    zc = cref.deformable_conv_forward(d["x"], d["off"], d["w"], None, border_mode=1)
    assert np.abs(zc - d["out"]).max() > 1e-2
    t = torch_ref.deformable_conv(torch.from_numpy(d["x"]), torch.from_numpy(d["off"]), torch.from_numpy(d["w"]), None, 0).numpy()
    assert np.abs(t - d["out"]).max() < 2e-5                                    # the differentiable restatement too


def test_affine_grid_and_sampler_match_tvm_reference():
    """GridGenerator('affine') (oracle/augment_ref.py, MXNet-recalled) and BilinearSampler (C oracle) against TVM's
    affine_grid_python / grid_sample_2d(bilinear, zeros, align_corners) -- 19 % of the grid points lie outside the image."""
    from oracle import augment_ref
    d = np.load(os.path.join(G, "affine_sampler_tvm.npz"))
    grid = augment_ref.grid_generator_affine(d["theta"], *d["grid"].shape[2:])
    assert np.abs(grid - d["grid"]).max() < 1e-6
    assert np.abs(cref.bilinear_sampler(d["data"], grid) - d["sampled"]).max() < 2e-5


def test_deformable_border_rule_mxnet15():
    """The MXNet-1.5 rule on a 1-channel ramp with a centre-tap delta kernel: zero for coordinate < 0, last pixel (no
    blend) for H-1 < h < H, zero for h >= H."""
    H = W = 6
    x = np.arange(H * W, dtype=np.float32).reshape(1, 1, H, W) + 1
    w = np.zeros((1, 1, 3, 3), np.float32); w[0, 0, 1, 1] = 1
    def run(dy, dx):
        off = np.zeros((1, 18, H, W), np.float32); off[:, 8] = dy; off[:, 9] = dx
        return cref.deformable_conv_forward(x, off, w, None, border_mode=0)[0, 0]
    assert np.array_equal(run(0, 0), x[0, 0])
    assert (run(-0.5, 0)[0] == 0).all()                       # h = -0.5 -> zero (DCNv2 would blend with zero)
    assert np.array_equal(run(0.5, 0)[H - 1], x[0, 0, H - 1])  # h = H-0.5 -> collapses on the last row
    assert (run(1.0, 0)[H - 1] == 0).all()                    # h = H -> zero
    assert np.allclose(run(0.25, 0)[2], 0.75 * x[0, 0, 2] + 0.25 * x[0, 0, 3])


def test_upsample_and_sampler_fixtures():
    rng = np.random.default_rng(2)
    u = rng.standard_normal((2, 3, 5, 7)).astype(np.float32)
    for f in (2, 4):
        a = cref.upsample(u, f)
        assert np.abs(a - torch_ref.upsample(torch.from_numpy(u), f).numpy()).max() < 1e-6
        pad = torch.nn.functional.pad(torch.from_numpy(u), (0, 1, 0, 1), mode="replicate")
        c = torch.nn.functional.interpolate(pad, size=(f * 5 + 1, f * 7 + 1), mode="bilinear", align_corners=True)
        assert np.abs(a - c[:, :, :-1, :-1].numpy()).max() < 1e-6
    assert np.allclose(cref.upsample(u, 2)[:, :, ::2, ::2], u)
    d = np.load(os.path.join(G, "sampler_torch.npz"))
    assert np.abs(cref.reconstruction2d(d["img"], d["flow_yx"]) - d["out"]).max() < 2e-5
    t = torch_ref.reconstruction2d(torch.from_numpy(d["img"]), torch.from_numpy(d["flow_yx"])).numpy()
    assert np.abs(t - d["out"]).max() < 1e-6


def test_warp_mask_composition_matches_c_oracle():
    rng = np.random.default_rng(3)
    N, C, H, W = 1, 8, 8, 12
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((C, C, 3, 3)) * 0.2).astype(np.float32)
    b = rng.standard_normal(C).astype(np.float32)
    fl = (rng.standard_normal((N, 2, H // 2, W // 2)) * 0.3).astype(np.float32)
    mk = rng.standard_normal((N, 1, H // 2, W // 2)).astype(np.float32)
    tr = rng.standard_normal((N, C, H, W)).astype(np.float32)
    t = torch.from_numpy
    out, fu, mu = torch_ref.warp_mask(t(x), t(fl), t(mk), t(w), t(b), t(tr), 20.0, 8, 2, 0)
    fu_c, mu_c = cref.upsample(fl, 2), cref.upsample(mk, 2)
    offs = np.repeat((fu_c * 20.0 / 8)[:, None], 9, 1).reshape(N, 18, H, W)
    pre = cref.deformable_conv_forward(x, offs, w, b) * (1 / (1 + np.exp(-mu_c))) + tr
    assert np.abs(np.where(pre > 0, pre, 0.1 * pre) - out.numpy()).max() < 2e-5


def test_prepost_oracle_against_torch_interpolate_and_roundtrip(tmp_path):
    """BilinearResize2D restatement == torch F.interpolate(align_corners=True); centralize; .flo / KITTI encoders round-trip."""
    from oracle import prepost_ref
    from maskflownet_b200 import flowio
    rng = np.random.default_rng(4)
    x = rng.standard_normal((2, 3, 13, 21)).astype(np.float32)
    for oh, ow in ((64, 64), (13, 21), (7, 40), (26, 11)):
        ref = torch.nn.functional.interpolate(torch.from_numpy(x), size=(oh, ow), mode="bilinear", align_corners=True).numpy()
        assert np.abs(prepost_ref.bilinear_resize2d(x, oh, ow) - ref).max() < 2e-5
    u1 = rng.integers(0, 256, (2, 3, 20, 30), dtype=np.uint8)
    u2 = rng.integers(0, 256, (2, 3, 20, 30), dtype=np.uint8)
    a, b, m = prepost_ref.preprocess(u1, u2, prepost_ref.padded_size(20, 30))
    assert a.shape == (2, 3, 64, 64)
    assert np.abs(np.concatenate([u1, u2], 2).astype(np.float64).mean((2, 3)) / 255 - m[:, :, 0, 0]).max() < 1e-6
    a0, b0, _ = prepost_ref.preprocess(u1, u2, None)
    assert abs(float(np.concatenate([a0, b0], 2).mean())) < 1e-6          # centralised
    pred = rng.standard_normal((1, 2, 16, 16)).astype(np.float32)
    out = prepost_ref.postprocess(pred, 64, 64)                            # no resize: Upsample(4), NHWC, flip
    up = cref.upsample(pred, 4)
    assert np.array_equal(out[0, :, :, 0], up[0, 1]) and np.array_equal(out[0, :, :, 1], up[0, 0])
    out2 = prepost_ref.postprocess(pred, 50, 60)
    assert out2.shape == (1, 50, 60, 2)
    flowio.write_flo(str(tmp_path / "a.flo"), out2[0])
    assert np.array_equal(flowio.read_flo(str(tmp_path / "a.flo")), out2[0])
    k = flowio.encode_kitti(out2[0])
    f, valid = flowio.decode_kitti(k)
    assert valid.all() and np.abs(f - out2[0]).max() <= 1 / 128 + 1e-6
