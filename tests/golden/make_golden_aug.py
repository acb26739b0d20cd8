"""Generates tests/golden/aug_ref_graph.npz: the reference's OWN augmentation.py (GeometryAugmentation, ColorAugmentation;
/root/reference/augmentation.py:168-339), imported UNCHANGED and executed on a small CPU operator namespace (torch, below),
with every F.random.* draw recorded.  The fixture pins oracle/augment_ref.py (parameter derivation + image / flow / mask
arithmetic) to the reference's graph; GridGenerator('affine') and BilinearSampler inside it are served by the C oracle's
sampler (pinned separately against torch.grid_sample) and by the MXNet-recalled affine grid definition.

Needs /root/reference (this container only); run from the repo root:  python tests/golden/make_golden_aug.py
"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import augment_ref, cref  # noqa: E402


class A:
    """Minimal NDArray: just what augmentation.py touches."""

    def __init__(self, t):
        self.t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t, dtype=torch.float32)

    @staticmethod
    def r(x):
        return x.t if isinstance(x, A) else x

    def __add__(self, o): return A(self.t + A.r(o))
    def __radd__(self, o): return A(A.r(o) + self.t)
    def __sub__(self, o): return A(self.t - A.r(o))
    def __rsub__(self, o): return A(A.r(o) - self.t)
    def __mul__(self, o): return A(self.t * A.r(o))
    def __rmul__(self, o): return A(A.r(o) * self.t)
    def __truediv__(self, o): return A(self.t / A.r(o))
    def __rtruediv__(self, o): return A(A.r(o) / self.t)
    def __neg__(self): return A(-self.t)
    def cos(self): return A(self.t.cos())
    def sin(self): return A(self.t.sin())
    def clip(self, lo, hi): return A(self.t.clamp(lo, hi))
    def max(self, axis=None, keepdims=False): return A(self.t.amax(dim=axis, keepdim=keepdims))
    def min(self, axis=None, keepdims=False): return A(self.t.amin(dim=axis, keepdim=keepdims))
    def repeat(self, repeats, axis): return A(self.t.repeat_interleave(repeats, dim=axis))
    def slice_axis(self, axis, begin, end): return A(self.t.narrow(axis, begin, end - begin))
    def broadcast_like(self, o): return A(self.t.expand_as(o.t).clone())


def _mx_reshape(shape_in, codes):
    out, i = [], 0
    codes = list(codes)
    k = 0
    while k < len(codes):
        c = codes[k]
        if c == 0:
            out.append(shape_in[i]); i += 1
        elif c == -3:
            out.append(shape_in[i] * shape_in[i + 1]); i += 2
        elif c == -1:
            out.append(-1); i += 1
        else:
            out.append(c); i += 1
        k += 1
    return out


class Random:
    def __init__(self, gen):
        self.gen, self.log = gen, []

    def _shape(self, shape):
        if shape is None:
            return (1,)
        return tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)

    def uniform(self, low=0.0, high=1.0, shape=None):
        t = torch.rand(self._shape(shape), generator=self.gen) * (high - low) + low
        self.log.append(t.numpy().copy())
        return A(t)

    def normal(self, loc=0.0, scale=1.0, shape=None):
        t = torch.randn(self._shape(shape), generator=self.gen) * scale + loc
        self.log.append(t.numpy().copy())
        return A(t)


class F:
    """The operator namespace handed to hybrid_forward."""
    random = None

    @staticmethod
    def zeros_like(x): return A(torch.zeros_like(x.t))
    @staticmethod
    def ones_like(x): return A(torch.ones_like(x.t))
    @staticmethod
    def zeros(shape): return A(torch.zeros(tuple(shape)))
    @staticmethod
    def ones(shape): return A(torch.ones(tuple(shape)))
    @staticmethod
    def stack(*xs, axis=0): return A(torch.stack([x.t for x in xs], dim=axis))
    @staticmethod
    def reshape(x, shape): return A(x.t.reshape(_mx_reshape(list(x.t.shape), shape)))
    @staticmethod
    def reshape_like(x, y): return A(x.t.reshape(y.t.shape))
    @staticmethod
    def batch_dot(a, b): return A(torch.bmm(a.t, b.t))
    @staticmethod
    def abs(x): return A(x.t.abs())
    @staticmethod
    def sin(x): return A(x.t.sin())
    @staticmethod
    def cos(x): return A(x.t.cos())
    @staticmethod
    def exp(x): return A(x.t.exp())
    @staticmethod
    def minimum(a, b): return A(torch.minimum(torch.as_tensor(A.r(a), dtype=torch.float32), torch.as_tensor(A.r(b), dtype=torch.float32)))
    @staticmethod
    def maximum(a, b): return A(torch.maximum(torch.as_tensor(A.r(a), dtype=torch.float32), torch.as_tensor(A.r(b), dtype=torch.float32)))
    @staticmethod
    def concat(*xs, dim=1): return A(torch.cat([x.t for x in xs], dim=dim))
    @staticmethod
    def broadcast_mul(a, b): return A(a.t * A.r(b))
    @staticmethod
    def broadcast_div(a, b): return A(a.t / A.r(b))
    @staticmethod
    def broadcast_add(a, b): return A(a.t + A.r(b))
    @staticmethod
    def broadcast_minus(a, b): return A(a.t - A.r(b))
    @staticmethod
    def broadcast_power(a, b): return A(torch.pow(a.t, A.r(b)))
    @staticmethod
    def slice_axis(x, axis, begin, end): return x.slice_axis(axis, begin, end)
    @staticmethod
    def mean(x, keepdims=False, axis=None): return A(x.t.mean(dim=axis, keepdim=keepdims))
    @staticmethod
    def clip(x, lo, hi): return A(x.t.clamp(lo, hi))
    @staticmethod
    def arange(a, b): return A(torch.arange(a, b, dtype=torch.float32))
    @staticmethod
    def one_hot(x, depth): return A(torch.nn.functional.one_hot(x.t.long(), depth).float())
    @staticmethod
    def repeat(x, axis, repeats): return A(x.t.repeat_interleave(repeats, dim=axis))

    @staticmethod
    def GridGenerator(data, transform_type, target_shape):
        assert transform_type == "affine"
        return A(torch.from_numpy(augment_ref.grid_generator_affine(data.t.numpy(), *target_shape)))

    @staticmethod
    def BilinearSampler(data, grid):
        return A(torch.from_numpy(cref.bilinear_sampler(data.t.numpy(), grid.t.numpy())))


def load_reference_augmentation():
    from maskflownet_b200 import mx
    mx.install()                                        # `from mxnet.gluon import nn` -> the shim's HybridBlock base class
    spec = importlib.util.spec_from_file_location("_mfn_reference_augmentation", "/root/reference/augmentation.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


GEO_DRAW_NAMES = ["rotation", "aspect_ratio", "scale", "tx_unit", "tx_range", "ty_unit", "ty_range", "rel_rotation", "rel_scale",
                  "rel_translation"]                     # call order of F.random.uniform in GeometryAugmentation.hybrid_forward
COLOR_DRAW_NAMES = ["contrast", "brightness", "channel", "noise_sigma", "gamma", "alpha_u", "theta", "noise1", "noise2"]


def main():
    aug = load_reference_augmentation()
    gen = torch.Generator().manual_seed(20260924)
    N, (OH, OW), (TH, TW) = 3, (40, 56), (24, 36)
    img1 = torch.rand(N, 3, OH, OW, generator=gen)
    img2 = torch.rand(N, 3, OH, OW, generator=gen)
    flow = torch.randn(N, 2, OH, OW, generator=gen) * 3.0
    mask = (torch.rand(N, 1, OH, OW, generator=gen) > 0.15).float()
    out = {"img1": img1.numpy(), "img2": img2.numpy(), "flow": flow.numpy(), "mask": mask.numpy(),
           "orig_shape": np.array([OH, OW]), "target_shape": np.array([TH, TW])}
    # ---- GeometryAugmentation with the FlyingChairs / Things settings of main.py:412-416 (larger angles to reach the clamps)
    F.random = Random(gen)
    geo = aug.GeometryAugmentation(angle_range=(-17, 17), zoom_range=(0.5, 1 / 0.9), aspect_range=(0.9, 1 / 0.9),
                                   translation_range=0.1, target_shape=(TH, TW), orig_shape=(OH, OW), batch_size=N,
                                   relative_angle=0.25, relative_scale=(0.96, 1 / 0.96), relative_translation=0.25)
    g1, g2, gf, gm = geo.hybrid_forward(F, A(img1), A(img2), A(flow), A(mask))
    assert len(F.random.log) == len(GEO_DRAW_NAMES), len(F.random.log)
    gd = dict(zip(GEO_DRAW_NAMES, F.random.log))
    for k, v in gd.items():
        out["geo_draw_" + k] = v.reshape(N, -1).squeeze(-1) if v.size == N else v.reshape(N, 2)
    out.update(geo_img1=g1.t.numpy(), geo_img2=g2.t.numpy(), geo_flow=gf.t.numpy(), geo_mask=gm.t.numpy())
    # broadcast mask variant (train_batch's default mask: (N,1,1,1) of ones)
    F.random = Random(torch.Generator().manual_seed(7))
    b1, b2, bf, bm = geo.hybrid_forward(F, A(img1), A(img2), A(flow), A(torch.ones(N, 1, 1, 1)))
    for k, v in zip(GEO_DRAW_NAMES, F.random.log):
        out["geob_draw_" + k] = v.reshape(N, -1).squeeze(-1) if v.size == N else v.reshape(N, 2)
    out.update(geob_img1=b1.t.numpy(), geob_img2=b2.t.numpy(), geob_flow=bf.t.numpy(), geob_mask=bm.t.numpy())
    # ---- ColorAugmentation, KITTI settings (main.py:393-394: noise + gamma) on the geometry outputs
    F.random = Random(gen)
    col = aug.ColorAugmentation(contrast_range=(-0.2, 0.4), brightness_sigma=0.05, channel_range=(0.9, 1.2), batch_size=N,
                                shape=(TH, TW), noise_range=(0, 0.02), saturation=0.25, hue=0.1, gamma_range=(-0.5, 0.5),
                                eigen_aug=False)
    c1, c2 = col.hybrid_forward(F, g1, g2)
    assert len(F.random.log) == len(COLOR_DRAW_NAMES), len(F.random.log)
    for k, v in zip(COLOR_DRAW_NAMES, F.random.log):
        out["col_draw_" + k] = v.reshape(N, 3) if k == "channel" else (v if v.ndim == 4 and v.shape[1:] == (3, TH, TW) else v.reshape(-1))
    out.update(col_img1=c1.t.numpy(), col_img2=c2.t.numpy())
    # ---- ColorAugmentation with eigen_aug (spin matrix), no gamma, no noise (Sintel settings, main.py:389-390)
    F.random = Random(gen)
    col2 = aug.ColorAugmentation(contrast_range=(-0.4, 0.8), brightness_sigma=0.1, channel_range=(0.8, 1.4), batch_size=N,
                                 shape=(TH, TW), noise_range=(0, 0), saturation=0.5, hue=0.5, eigen_aug=True)
    e1, e2 = col2.hybrid_forward(F, g1, g2)
    names = ["contrast", "brightness", "channel", "noise_sigma", "alpha_u", "theta", "spin_angle", "noise1", "noise2"]
    assert len(F.random.log) == len(names)
    for k, v in zip(names, F.random.log):
        if k.startswith("noise") and k != "noise_sigma":
            continue
        out["eig_draw_" + k] = v.reshape(N, 3) if k in ("channel", "spin_angle") else v.reshape(-1)
    out.update(eig_img1=e1.t.numpy(), eig_img2=e2.t.numpy())

    # ---- the restatement reproduces the reference's graph
    gdr = {k: out["geo_draw_" + k] for k in GEO_DRAW_NAMES}
    P = augment_ref.geometry_params(gdr, (OH, OW), (TH, TW))
    o = augment_ref.geometry_augment(out["img1"], out["img2"], out["flow"], out["mask"], P, (TH, TW))
    for name, a, b in zip(("img1", "img2", "flow", "mask"), o, (g1, g2, gf, gm)):
        err = np.abs(a - b.t.numpy()).max()
        print("geometry", name, "max |oracle - reference graph| =", err)
        assert err < (2e-4 if name == "flow" else 2e-5), (name, err)
    np.savez_compressed(os.path.join(HERE, "aug_ref_graph.npz"), **out)
    print("wrote", os.path.join(HERE, "aug_ref_graph.npz"), os.path.getsize(os.path.join(HERE, "aug_ref_graph.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
