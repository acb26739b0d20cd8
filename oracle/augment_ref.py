"""CPU restatement of the reference's GPU-side augmentation.  TEST INFRASTRUCTURE ONLY.

  grid_generator_affine   MXNet F.GridGenerator(transform_type='affine') [MXNet-recalled: grid_generator-inl.h -- the target
                          grid is x = -1 + j * 2/(W-1), y = -1 + i * 2/(H-1), out = theta(2x3) . (x, y, 1), channel 0 = x]
  geometry_params         the per-sample matrices GeometryAugmentation.hybrid_forward derives from its random draws
                          (/root/reference/augmentation.py:278-303, 326, 337)
  geometry_augment        the image / flow / mask part of the same function (:305-339)
  color_params / color_augment   ColorAugmentation.hybrid_forward (:182-227)
  philox_normal           the counter-based noise stream of mfn_color_augment_forward (Philox4x32-10 + Box-Muller)

The random draws are INPUTS (a dict, in the order the reference calls F.random.*): the restatement is deterministic and is
pinned against the reference's own augmentation.py executed unchanged on a CPU operator namespace with the same draws
(tests/golden/make_golden_aug.py -> tests/golden/aug_ref_graph.npz).
"""
from __future__ import annotations

import numpy as np

from . import cref

f32 = np.float32


def grid_generator_affine(theta, TH: int, TW: int) -> np.ndarray:
    """theta (N, 6) -> grid (N, 2, TH, TW); fp32, products summed left to right (a*x + b*y + c)."""
    theta = np.asarray(theta, dtype=f32).reshape(-1, 6)
    xs = (f32(-1.0) + np.arange(TW, dtype=f32) * f32(2.0 / (TW - 1))).astype(f32)[None, None, :]
    ys = (f32(-1.0) + np.arange(TH, dtype=f32) * f32(2.0 / (TH - 1))).astype(f32)[None, :, None]
    t = theta[:, :, None, None]
    gx = (t[:, 0] * xs + t[:, 1] * ys).astype(f32) + t[:, 2]
    gy = (t[:, 3] * xs + t[:, 4] * ys).astype(f32) + t[:, 5]
    return np.stack([gx, gy], axis=1).astype(f32)


# ---------------------------------------------------------------------------------------------------------------
# GeometryAugmentation
# ---------------------------------------------------------------------------------------------------------------
def geometry_params(draws: dict, orig_shape, target_shape, aspect: bool = True, relative: bool = True,
                    relative_translation: bool = True) -> np.ndarray:
    """Draws (each (N,) float32; names = the variables of augmentation.py:279-303):
         rotation, aspect_ratio, scale, tx_unit, tx_range, ty_unit, ty_range, rel_rotation, rel_scale, rel_translation (N,2)
       -> the (N, 22) parameter block of mfn_geometry_augment_forward:
         [0:6]  affine_params (:291-293)   [6:12] affine_2 (:305)   [12:14] rel_translation (:308)
         [14:18] inverse_2 = rel_inverse . linv (:326)   [18:22] factor = (rel_inverse - I) . diag((TW-1)/2, (TH-1)/2) (:337)"""
    OH, OW = orig_shape
    TH, TW = target_shape
    g = lambda k: np.asarray(draws[k], dtype=f32)  # noqa: E731
    rot = g("rotation")
    N = rot.shape[0]
    asp = g("aspect_ratio") if aspect else np.ones(N, f32)
    scale = g("scale")
    # unit[i][j] = flip(target - 1)[i] / flip(orig - 1)[j]                                    (:246)
    ft, fo = np.array([TW - 1, TH - 1], np.float64), np.array([OW - 1, OH - 1], np.float64)
    unit = ft.reshape(2, 1) / fo.reshape(1, 2)
    os_, ts = (OH - 1, OW - 1), (TH - 1, TW - 1)
    ar = np.abs(rot)
    scale = np.minimum(scale, (f32(os_[1]) / (asp * (f32(ts[0]) * np.sin(ar) + f32(ts[1]) * np.cos(ar)))).astype(f32))   # :285
    scale = np.minimum(scale, (f32(os_[0]) / (f32(ts[0]) * np.cos(ar) + f32(ts[1]) * np.sin(ar))).astype(f32))           # :286
    pad_x, pad_y = f32(1) - scale * f32(unit[0, 0]), f32(1) - scale * f32(unit[1, 1])
    tx = g("tx_unit") * pad_x + g("tx_range")
    ty = g("ty_unit") * pad_y + g("ty_range")
    cos, sin = np.cos(rot).astype(f32), np.sin(rot).astype(f32)
    A = np.stack([scale * asp * cos * f32(unit[0, 0]), scale * asp * -sin * f32(unit[1, 0]), tx,
                  scale * sin * f32(unit[0, 1]), scale * cos * f32(unit[1, 1]), ty], axis=1).astype(f32)
    linv = np.stack([cos / (scale * asp), sin / (scale * asp), -sin / scale, cos / scale], axis=1).reshape(N, 2, 2).astype(f32)
    if relative:
        rrot, rscale = g("rel_rotation"), g("rel_scale")
    else:
        rrot, rscale = np.zeros(N, f32), np.ones(N, f32)
    ratio = f32((TH - 1) / (TW - 1))
    rc, rs = np.cos(rrot).astype(f32), np.sin(rrot).astype(f32)
    z, o = np.zeros(N, f32), np.ones(N, f32)
    rel = np.stack([rscale * rc, rscale * -rs * ratio, z, rscale * rs / ratio, rscale * rc, z, z, z, o], axis=1).reshape(N, 3, 3)
    rel_inv = np.stack([rc / rscale, rs / rscale, -rs / rscale, rc / rscale], axis=1).reshape(N, 2, 2).astype(f32)
    A2 = np.matmul(A.reshape(N, 2, 3), rel.astype(f32)).reshape(N, 6).astype(f32)
    rt = g("rel_translation").reshape(N, 2) if (relative and relative_translation) else np.zeros((N, 2), f32)
    inv2 = np.matmul(rel_inv, linv).astype(f32)
    S = np.array([[(TW - 1) / 2, 0], [0, (TH - 1) / 2]], f32)
    factor = np.matmul(rel_inv - np.eye(2, dtype=f32)[None], S[None]).astype(f32)
    return np.concatenate([A, A2, rt, inv2.reshape(N, 4), factor.reshape(N, 4)], axis=1).astype(f32)


def geometry_augment(img1, img2, flow, mask, params, target_shape):
    """augmentation.py:299-339 given the parameter block.  img1/img2 (N,3,H,W), flow (N,2,H,W), mask (N,1,H,W) or (N,1,1,1),
    all float32 (already / 255) -> img1', img2', flow', mask' on the target grid."""
    img1, img2, flow = (np.asarray(a, dtype=f32) for a in (img1, img2, flow))
    N, _, H, W = img1.shape
    TH, TW = target_shape
    P = np.asarray(params, dtype=f32).reshape(N, 22)
    mask = np.broadcast_to(np.asarray(mask, dtype=f32), (N, 1, H, W))                                   # :299
    rt = P[:, 12:14].reshape(N, 2, 1, 1)
    rel_scale = np.array([(W - 1) / 2, (H - 1) / 2], f32).reshape(1, 2, 1, 1)
    flow = flow - rt * rel_scale                                                                       # :303-307
    cat = np.concatenate([img1, mask, flow * mask], axis=1).astype(f32)                                # :309
    grid = grid_generator_affine(P[:, 0:6], TH, TW)
    ft = (np.maximum(grid.max(axis=(2, 3), keepdims=True) - f32(1), f32(0)) +
          np.minimum(grid.min(axis=(2, 3), keepdims=True) + f32(1), f32(0))).astype(f32)                # :311
    grid = np.clip(grid - ft, f32(-1), f32(1))
    s = cref.bilinear_sampler(cat, grid)
    o1, om, of = s[:, 0:3], s[:, 3:4], s[:, 4:6]
    of = of / np.maximum(om, f32(1e-8))                                                                # :318
    grid2 = grid_generator_affine(P[:, 6:12], TH, TW) - ft + rt                                        # :321-324
    o2 = cref.bilinear_sampler(img2, grid2.astype(f32))
    inv2 = P[:, 14:18].reshape(N, 2, 2)
    of = np.einsum("nij,njhw->nihw", inv2, of).astype(f32)                                             # :326-327
    ident = grid_generator_affine(np.array([[1, 0, 0, 0, 1, 0]], f32), TH, TW)                         # :336-340
    factor = P[:, 18:22].reshape(N, 2, 2)
    of = of + np.einsum("nij,jhw->nihw", factor, ident[0]).astype(f32)
    return o1.copy(), o2, of.astype(f32), om.copy()


# ---------------------------------------------------------------------------------------------------------------
# ColorAugmentation
# ---------------------------------------------------------------------------------------------------------------
def color_params(draws: dict, gamma: bool = False, eigen: bool = False) -> np.ndarray:
    """Draws (names = the variables of augmentation.py:183-204): contrast (N,), brightness (N,), channel (N,3), gamma (N,),
    alpha_u (N,) (the uniform before `1.0 +`), theta (N,), spin_angle (N,3)
       -> the (N, 26) parameter block of mfn_color_augment_forward:
         [0:9] sh_matrix (:198-200)  [9:12] contrast*channel  [12:15] channel  [15] brightness  [16] exp(gamma) or 1
         [17:26] spin_matrix (:206-208) or the identity"""
    g = lambda k: np.asarray(draws[k], dtype=f32)  # noqa: E731
    contrast = g("contrast") + f32(1)
    N = contrast.shape[0]
    alpha = f32(1.0) + g("alpha_u")
    su, sw = alpha * np.cos(g("theta")), alpha * np.sin(g("theta"))
    c = lambda v: f32(v)  # noqa: E731
    sh = np.stack([c(0.299) + c(0.701) * su + c(0.168) * sw, c(0.587) - c(0.587) * su + c(0.330) * sw, c(0.114) - c(0.114) * su - c(0.497) * sw,
                   c(0.299) - c(0.299) * su - c(0.328) * sw, c(0.587) + c(0.413) * su + c(0.035) * sw, c(0.114) - c(0.114) * su + c(0.292) * sw,
                   c(0.299) - c(0.300) * su + c(1.250) * sw, c(0.587) - c(0.588) * su - c(1.050) * sw, c(0.114) + c(0.886) * su - c(0.203) * sw],
                  axis=1).astype(f32)
    channel = g("channel").reshape(N, 3)
    if eigen:
        a = g("spin_angle").reshape(N, 3)
        c0, c1, c2 = (np.cos(a[:, k]) for k in range(3))
        s0, s1, s2 = (np.sin(a[:, k]) for k in range(3))
        spin = np.stack([c0 * c1, s1 * c2 + s0 * c1 * s2, s1 * s2 - s0 * c1 * c2,
                         -c0 * s1, c1 * c2 - s0 * s1 * s2, c1 * s2 + s0 * s1 * c2,
                         s0, -c0 * s2, c0 * c2], axis=1).astype(f32)
    else:
        spin = np.tile(np.eye(3, dtype=f32).reshape(1, 9), (N, 1))
    pw = np.exp(g("gamma")).astype(f32) if gamma else np.ones(N, f32)
    return np.concatenate([sh, contrast[:, None] * channel, channel, g("brightness")[:, None], pw[:, None], spin], axis=1).astype(f32)


def color_augment_one(img, params, noise=None, noise_sigma: float = 0.0) -> np.ndarray:
    """One image through augmentation.py:211-225."""
    img = np.asarray(img, dtype=f32)
    N = img.shape[0]
    P = np.asarray(params, dtype=f32).reshape(N, 26)
    sh = P[:, 0:9].reshape(N, 3, 3)
    aug = np.zeros_like(img)
    for i in range(3):       # sum over j in the reference's order (python sum: ((0 + t0) + t1) + t2)
        acc = np.zeros_like(img[:, 0])
        for j in range(3):
            acc = acc + img[:, j] * sh[:, i, j][:, None, None]
        aug[:, i] = acc
    if noise is not None:
        aug = aug + np.asarray(noise, dtype=f32) * f32(noise_sigma)
    mean = aug.mean(axis=(2, 3), keepdims=True, dtype=np.float64).astype(f32)
    aug = (aug - mean) * P[:, 9:12].reshape(N, 3, 1, 1)
    spin = P[:, 17:26].reshape(N, 3, 3)
    out = np.zeros_like(aug)
    for i in range(3):
        acc = np.zeros_like(aug[:, 0])
        for j in range(3):
            acc = acc + aug[:, j] * spin[:, i, j][:, None, None]
        out[:, i] = acc
    out = out + (mean * P[:, 12:15].reshape(N, 3, 1, 1) + P[:, 15].reshape(N, 1, 1, 1))
    out = np.clip(out, f32(0), f32(1))
    return np.power(out, P[:, 16].reshape(N, 1, 1, 1)).astype(f32)


# ---------------------------------------------------------------------------------------------------------------
# The in-kernel noise stream (ours, not the reference's: MXNet draws its noise from its own generator)
# ---------------------------------------------------------------------------------------------------------------
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_LO = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 on arrays of counters (uint64 holding 32-bit values); returns four uint64 arrays of 32-bit results."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & _LO for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & _LO, np.uint64(k1) & _LO
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _LO, p1 & _LO, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _LO, p0 & _LO
        k0, k1 = (k0 + _W0) & _LO, (k1 + _W1) & _LO
    return c0, c1, c2, c3


def philox_normal(N: int, H: int, W: int, seed: int, image: int) -> np.ndarray:
    """(N,3,H,W) standard normals exactly as mfn_color_augment_forward generates them: one Philox call per (pixel, image),
    counter = (pixel index lo, hi, image, 0), key = seed lo, hi; u = ((r >> 8) + 1) * 2^-24; Box-Muller on (r0, r1) gives the
    values of channels 0 and 1, on (r2, r3) its cosine branch gives channel 2."""
    idx = np.arange(N * H * W, dtype=np.uint64)
    r = philox4x32_10(idx & _LO, idx >> np.uint64(32), np.full_like(idx, image), np.zeros_like(idx),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [(((x >> np.uint64(8)) + np.uint64(1)).astype(np.float64) * 2.0 ** -24).astype(f32) for x in r]
    rad0 = np.sqrt(f32(-2.0) * np.log(u[0])).astype(f32)
    rad1 = np.sqrt(f32(-2.0) * np.log(u[2])).astype(f32)
    two_pi = f32(6.283185307179586)
    z = np.stack([rad0 * np.cos(two_pi * u[1]), rad0 * np.sin(two_pi * u[1]), rad1 * np.cos(two_pi * u[3])], axis=1)
    return z.reshape(N, H, W, 3).transpose(0, 3, 1, 2).astype(f32).copy()
