"""GPU-side training augmentation -- host side of csrc/augment.cu.

Mirrors /root/reference/augmentation.py: `GeometryAugmentation` (:229-339) and `ColorAugmentation` (:168-227) take the
reference's constructor arguments and are called like the reference's blocks (network/pipeline.py:101-102):

    geo = GeometryAugmentation(angle_range=(-17, 17), zoom_range=(0.5, 1 / 0.9), aspect_range=(0.9, 1 / 0.9),
                               translation_range=0.1, target_shape=(320, 448), orig_shape=(384, 512), batch_size=8,
                               relative_angle=0.25, relative_scale=(0.96, 1 / 0.96), relative_translation=0.25)    # main.py:412-416
    img1, img2, flow, mask = geo(img1_u8, img2_u8, flow, mask_u8)       # uint8 in: the `/ 255` of pipeline.py:100 is folded in
    img1, img2 = color(img1, img2)

What the reference computes with ~25 + ~40 MXNet operator launches per call is split here into
  * host logic (this file): the random draws (a CPU torch.Generator, a few floats per sample) and the small matrices the
    reference derives from them -- the (N,22) / (N,26) parameter blocks documented in include/maskflow_b200.h;
  * one CUDA launch (geometry) / two (colour) over the images: `geometry_augment`, `color_augment` below.
No CPU fallback: the tensor functions raise for non-CUDA tensors.  Forward only (the reference does not differentiate
through its augmentation either: inputs are data).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MaskflowError
from .ops import _call, _chk, _no_grad_path, _p

GEO_P, COL_P = 22, 26


# ----------------------------------------------------------------------------------------------------------
# tensor-level operators (C ABI)
# ----------------------------------------------------------------------------------------------------------
def geometry_augment(img1: torch.Tensor, img2: torch.Tensor, flow: torch.Tensor, mask: torch.Tensor, params: torch.Tensor,
                     target_shape: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """mfn_geometry_augment_forward: images / mask uint8 (read as x / 255) or float32, flow float32 (channel 0 = x), mask
    (N,1,H,W) or (N,1,1,1); params (N,22) float32 on the same device.  Returns img1', img2', flow', mask' on target_shape."""
    for t, nm in ((img1, "img1"), (img2, "img2"), (mask, "mask")):
        if not (t.is_cuda and t.dim() == 4 and t.dtype in (torch.uint8, torch.float32)):
            raise MaskflowError(f"geometry_augment: {nm} must be a CUDA uint8 / float32 NCHW tensor (got {t.dtype} on {t.device})")
    if not (img1.shape == img2.shape and img1.dtype == img2.dtype == mask.dtype and img1.shape[1] == 3):
        raise MaskflowError("geometry_augment: img1, img2 (N,3,H,W) and mask must agree in dtype; images in shape")
    N, _, H, W = img1.shape
    flow = _chk(flow, "geometry_augment.flow")
    params = _chk(params, "geometry_augment.params")
    if flow.shape != (N, 2, H, W):
        raise MaskflowError(f"geometry_augment: flow must be (N,2,H,W) = {(N, 2, H, W)}, got {tuple(flow.shape)}")
    if tuple(mask.shape) == (N, 1, 1, 1):
        bcast = 1
    elif tuple(mask.shape) == (N, 1, H, W):
        bcast = 0
    else:
        raise MaskflowError(f"geometry_augment: mask must be (N,1,H,W) or (N,1,1,1), got {tuple(mask.shape)}")
    if tuple(params.shape) != (N, GEO_P):
        raise MaskflowError(f"geometry_augment: params must be (N,{GEO_P}), got {tuple(params.shape)}")
    _no_grad_path("geometry_augment", img1, img2, flow, mask, params)
    img1, img2, mask = img1.contiguous(), img2.contiguous(), mask.contiguous()
    TH, TW = int(target_shape[0]), int(target_shape[1])
    dev = img1.device
    o1 = torch.empty((N, 3, TH, TW), device=dev, dtype=torch.float32)
    o2 = torch.empty_like(o1)
    of = torch.empty((N, 2, TH, TW), device=dev, dtype=torch.float32)
    om = torch.empty((N, 1, TH, TW), device=dev, dtype=torch.float32)
    _call("mfn_geometry_augment_forward", dev, _p(img1), _p(img2), 1 if img1.dtype == torch.uint8 else 0, _p(flow), _p(mask),
          bcast, _p(params), _p(o1), _p(o2), _p(of), _p(om), N, H, W, TH, TW)
    return o1, o2, of, om


def color_augment(img1: torch.Tensor, img2: torch.Tensor, params: torch.Tensor, noise_sigma: float = 0.0,
                  noise: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, seed: int = 0,
                  has_gamma: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """mfn_color_augment_forward: img1 / img2 (N,3,H,W) float32, params (N,26).  noise = (noise1, noise2) standard-normal
    tensors of the image shape, or None: the kernels then draw the noise themselves from `seed` when noise_sigma != 0."""
    a, b, params = _chk(img1, "color_augment.img1"), _chk(img2, "color_augment.img2"), _chk(params, "color_augment.params")
    if a.shape != b.shape or a.dim() != 4 or a.shape[1] != 3:
        raise MaskflowError("color_augment: img1 / img2 must both be (N,3,H,W)")
    N, _, H, W = a.shape
    if tuple(params.shape) != (N, COL_P):
        raise MaskflowError(f"color_augment: params must be (N,{COL_P}), got {tuple(params.shape)}")
    n1 = n2 = None
    if noise is not None:
        n1, n2 = _chk(noise[0], "color_augment.noise1"), _chk(noise[1], "color_augment.noise2")
        if n1.shape != a.shape or n2.shape != a.shape:
            raise MaskflowError("color_augment: the noise tensors must have the image shape")
    _no_grad_path("color_augment", a, b, params, n1, n2)
    wsb = int(_lib.lib().mfn_color_augment_workspace_bytes(N))
    ws = torch.empty(wsb // 4, device=a.device, dtype=torch.float32)
    o1, o2 = torch.empty_like(a), torch.empty_like(b)
    _call("mfn_color_augment_forward", a.device, _p(a), _p(b), _p(params), _p(n1), _p(n2), float(noise_sigma),
          int(seed) & 0x7FFFFFFFFFFFFFFF, _p(o1), _p(o2), _p(ws), wsb, N, H, W, 1 if has_gamma else 0)
    return o1, o2


# ----------------------------------------------------------------------------------------------------------
# host logic: draws -> parameter blocks (CPU float32 torch; a few floats per sample)
# ----------------------------------------------------------------------------------------------------------
def _uniform(gen, low, high, shape):
    return torch.rand(shape, generator=gen, dtype=torch.float32) * (high - low) + low


class GeometryAugmentation:
    """augmentation.GeometryAugmentation (/root/reference/augmentation.py:229-339): same constructor arguments."""

    DRAWS = ("rotation", "aspect_ratio", "scale", "tx_unit", "tx_range", "ty_unit", "ty_range", "rel_rotation", "rel_scale",
             "rel_translation")     # the order of the reference's F.random.uniform calls (:279-302)

    def __init__(self, angle_range, zoom_range, translation_range, target_shape, orig_shape, batch_size, aspect_range=None,
                 relative_angle=None, relative_scale=None, relative_translation=None, seed: Optional[int] = None):
        self._angle_range = tuple(x / 180 * math.pi for x in angle_range)
        self._scale_range = tuple(zoom_range)
        try:
            translation_range = tuple(translation_range)
            if len(translation_range) != 2:
                raise ValueError(f"expect translation range to have shape [2,], but got {translation_range}")
        except TypeError:
            translation_range = (-translation_range, translation_range)
        self._translation_range = tuple(x * 2 for x in translation_range)
        self._target_shape = (int(target_shape[0]), int(target_shape[1]))
        self._orig_shape = (int(orig_shape[0]), int(orig_shape[1]))
        self._batch_size = int(batch_size)
        if relative_angle is None:
            # the reference's hybrid_forward dereferences the relative ranges unconditionally (:300): same contract
            raise MaskflowError("GeometryAugmentation: relative_angle / relative_scale are required (as in the reference's "
                                "hybrid_forward, which has no non-relative path)")
        self._relative_scale = tuple(relative_scale)
        self._relative_angle = tuple(x / 180 * math.pi * relative_angle for x in angle_range)
        self._relative_translation = (tuple(x * relative_translation for x in self._translation_range)
                                      if relative_translation is not None else None)
        self._aspect_range = None if aspect_range is None else tuple(aspect_range)
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(int(seed))

    def sample(self) -> Dict[str, torch.Tensor]:
        """The random draws of one call, in the reference's order (CPU tensors)."""
        N, g = self._batch_size, self._gen
        d = {"rotation": _uniform(g, *self._angle_range, (N,))}
        d["aspect_ratio"] = _uniform(g, *self._aspect_range, (N,)) if self._aspect_range is not None else torch.ones(N)
        d["scale"] = _uniform(g, *self._scale_range, (N,))
        d["tx_unit"] = _uniform(g, -1, 1, (N,))
        d["tx_range"] = _uniform(g, *self._translation_range, (N,))
        d["ty_unit"] = _uniform(g, -1, 1, (N,))
        d["ty_range"] = _uniform(g, *self._translation_range, (N,))
        d["rel_rotation"] = _uniform(g, *self._relative_angle, (N,))
        d["rel_scale"] = _uniform(g, *self._relative_scale, (N,))
        d["rel_translation"] = (_uniform(g, *self._relative_translation, (N, 2)) if self._relative_translation is not None
                                else torch.zeros(N, 2))
        return d

    def params(self, draws: Dict[str, torch.Tensor]) -> torch.Tensor:
        """(N,22) parameter block of mfn_geometry_augment_forward from the draws (augmentation.py:279-303, 326, 337)."""
        f = lambda k: torch.as_tensor(draws[k], dtype=torch.float32).cpu()  # noqa: E731
        (OH, OW), (TH, TW) = self._orig_shape, self._target_shape
        rot, asp, scale = f("rotation"), f("aspect_ratio"), f("scale")
        N = rot.shape[0]
        # unit[i][j] = flip(target - 1)[i] / flip(orig - 1)[j]   (:246)
        u00, u01, u10, u11 = (TW - 1) / (OW - 1), (TW - 1) / (OH - 1), (TH - 1) / (OW - 1), (TH - 1) / (OH - 1)
        ar = rot.abs()
        scale = torch.minimum(scale, (OW - 1) / (asp * ((TH - 1) * ar.sin() + (TW - 1) * ar.cos())))       # :285
        scale = torch.minimum(scale, (OH - 1) / ((TH - 1) * ar.cos() + (TW - 1) * ar.sin()))               # :286
        pad_x, pad_y = 1 - scale * u00, 1 - scale * u11
        tx = f("tx_unit") * pad_x + f("tx_range")
        ty = f("ty_unit") * pad_y + f("ty_range")
        cos, sin = rot.cos(), rot.sin()
        A = torch.stack([scale * asp * cos * u00, scale * asp * -sin * u10, tx, scale * sin * u01, scale * cos * u11, ty], dim=1)
        linv = torch.stack([cos / (scale * asp), sin / (scale * asp), -sin / scale, cos / scale], dim=1).reshape(N, 2, 2)
        rrot, rscale = f("rel_rotation"), f("rel_scale")
        ratio = (TH - 1) / (TW - 1)
        rc, rs = rrot.cos(), rrot.sin()
        z, o = torch.zeros(N), torch.ones(N)
        rel = torch.stack([rscale * rc, rscale * -rs * ratio, z, rscale * rs / ratio, rscale * rc, z, z, z, o], dim=1).reshape(N, 3, 3)
        rel_inv = torch.stack([rc / rscale, rs / rscale, -rs / rscale, rc / rscale], dim=1).reshape(N, 2, 2)
        A2 = torch.bmm(A.reshape(N, 2, 3), rel).reshape(N, 6)
        rt = f("rel_translation").reshape(N, 2) if self._relative_translation is not None else torch.zeros(N, 2)
        inv2 = torch.bmm(rel_inv, linv).reshape(N, 4)
        S = torch.tensor([[(TW - 1) / 2, 0.0], [0.0, (TH - 1) / 2]], dtype=torch.float32)
        factor = torch.matmul(rel_inv - torch.eye(2), S).reshape(N, 4)
        return torch.cat([A, A2, rt, inv2, factor], dim=1).to(torch.float32).contiguous()

    def __call__(self, img1, img2, flow, mask, draws: Optional[Dict[str, torch.Tensor]] = None):
        if tuple(img1.shape[2:]) != self._orig_shape or img1.shape[0] != self._batch_size:
            raise MaskflowError(f"GeometryAugmentation: built for batch {self._batch_size} of {self._orig_shape}, "
                                f"got {tuple(img1.shape)}")
        P = self.params(self.sample() if draws is None else draws)
        return geometry_augment(img1, img2, flow, mask, P.to(img1.device, non_blocking=True), self._target_shape)


class ColorAugmentation:
    """augmentation.ColorAugmentation (/root/reference/augmentation.py:168-227): same constructor arguments."""

    def __init__(self, contrast_range, brightness_sigma, channel_range, batch_size, shape, noise_range, saturation, hue,
                 gamma_range=None, eigen_aug=False, seed: Optional[int] = None):
        self._contrast_range = tuple(contrast_range)
        self._brightness_sigma = float(brightness_sigma)
        self._channel_range = tuple(channel_range)
        self._batch_size = int(batch_size)
        self._shape = tuple(shape)
        self._noise_range = tuple(noise_range)
        self._gamma_range = None if gamma_range is None else tuple(gamma_range)
        self._eigen_aug = bool(eigen_aug)
        self._saturation = float(saturation)
        self._hue = float(hue)
        self._gen = torch.Generator()
        if seed is not None:
            self._gen.manual_seed(int(seed))

    def sample(self) -> Dict[str, torch.Tensor]:
        """The per-sample draws of one call in the reference's order (:183-204); the per-pixel noise is drawn in the kernel
        from `noise_seed` (the reference draws two full-size normal tensors, :214)."""
        N, g = self._batch_size, self._gen
        d = {"contrast": _uniform(g, *self._contrast_range, (N,)),
             "brightness": torch.randn(N, generator=g) * self._brightness_sigma,
             "channel": _uniform(g, *self._channel_range, (N, 3)),
             "noise_sigma": _uniform(g, *self._noise_range, (1,))}
        if self._gamma_range is not None:
            d["gamma"] = _uniform(g, *self._gamma_range, (N,))
        d["alpha_u"] = _uniform(g, -self._saturation, self._saturation, (N,))
        d["theta"] = _uniform(g, -self._hue * math.pi, self._hue * math.pi, (N,))
        if self._eigen_aug:
            d["spin_angle"] = _uniform(g, -math.pi, math.pi, (N, 3))
        d["noise_seed"] = torch.randint(0, 2 ** 62, (1,), generator=g)
        return d

    def params(self, draws: Dict[str, torch.Tensor]) -> torch.Tensor:
        """(N,26) parameter block of mfn_color_augment_forward from the draws (augmentation.py:183-208)."""
        f = lambda k: torch.as_tensor(draws[k], dtype=torch.float32).cpu()  # noqa: E731
        contrast = f("contrast") + 1
        N = contrast.shape[0]
        alpha = 1.0 + f("alpha_u")
        su, sw = alpha * f("theta").cos(), alpha * f("theta").sin()
        sh = torch.stack([0.299 + 0.701 * su + 0.168 * sw, 0.587 - 0.587 * su + 0.330 * sw, 0.114 - 0.114 * su - 0.497 * sw,
                          0.299 - 0.299 * su - 0.328 * sw, 0.587 + 0.413 * su + 0.035 * sw, 0.114 - 0.114 * su + 0.292 * sw,
                          0.299 - 0.300 * su + 1.250 * sw, 0.587 - 0.588 * su - 1.050 * sw, 0.114 + 0.886 * su - 0.203 * sw],
                         dim=1)
        channel = f("channel").reshape(N, 3)
        if self._eigen_aug:
            a = f("spin_angle").reshape(N, 3)
            c0, c1, c2 = a[:, 0].cos(), a[:, 1].cos(), a[:, 2].cos()
            s0, s1, s2 = a[:, 0].sin(), a[:, 1].sin(), a[:, 2].sin()
            spin = torch.stack([c0 * c1, s1 * c2 + s0 * c1 * s2, s1 * s2 - s0 * c1 * c2,
                                -c0 * s1, c1 * c2 - s0 * s1 * s2, c1 * s2 + s0 * s1 * c2,
                                s0, -c0 * s2, c0 * c2], dim=1)
        else:
            spin = torch.eye(3).reshape(1, 9).repeat(N, 1)
        pw = f("gamma").exp() if self._gamma_range is not None else torch.ones(N)
        return torch.cat([sh, contrast[:, None] * channel, channel, f("brightness")[:, None], pw[:, None], spin],
                         dim=1).to(torch.float32).contiguous()

    def __call__(self, img1, img2, draws: Optional[Dict[str, torch.Tensor]] = None, noise=None):
        if tuple(img1.shape[2:]) != self._shape or img1.shape[0] != self._batch_size:
            raise MaskflowError(f"ColorAugmentation: built for batch {self._batch_size} of {self._shape}, got {tuple(img1.shape)}")
        d = self.sample() if draws is None else draws
        sigma = float(torch.as_tensor(d["noise_sigma"]).reshape(-1)[0])
        seed = int(torch.as_tensor(d.get("noise_seed", 0)).reshape(-1)[0])
        return color_augment(img1, img2, self.params(d).to(img1.device, non_blocking=True), noise_sigma=sigma, noise=noise,
                             seed=seed, has_gamma=self._gamma_range is not None)
