"""CPU restatement of the steps either side of the network.  TEST INFRASTRUCTURE ONLY.

  bilinear_resize2d   MXNet contrib.BilinearResize2D [MXNet-recalled: bilinear_resize-inl.h, "align corners" mapping]
  preprocess          PipelineFlownet.predict (`/255`, network/pipeline.py:212) + centralize (:85-87) + do_batch_mx's resize (:117-130)
  postprocess         do_batch (:137-141: Upsample(4), resize back, per-channel rescale) + predict (:217-218: NHWC, flip)
"""
from __future__ import annotations

import numpy as np

from . import cref


def bilinear_resize2d(x: np.ndarray, OH: int, OW: int) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    N, C, H, W = x.shape

    def taps(n_in, n_out):
        r = np.float32((n_in - 1) / (n_out - 1)) if n_out > 1 else np.float32(0)
        s = (r * np.arange(n_out, dtype=np.float32)).astype(np.float32)
        i0 = s.astype(np.int64)
        i1 = i0 + (i0 < n_in - 1)
        return i0, i1, (s - i0.astype(np.float32)).astype(np.float32)
    y0, y1, ly = taps(H, OH)
    x0, x1, lx = taps(W, OW)
    ly, lx = ly[:, None], lx[None, :]
    a, b = x[:, :, y0][:, :, :, x0], x[:, :, y0][:, :, :, x1]
    c, d = x[:, :, y1][:, :, :, x0], x[:, :, y1][:, :, :, x1]
    return ((1 - ly) * (1 - lx) * a + (1 - ly) * lx * b + ly * (1 - lx) * c + ly * lx * d).astype(np.float32)


def padded_size(H, W, resize=None):
    if resize is not None:
        return int(resize[0]), int(resize[1])
    return H + (64 - H % 64) % 64, W + (64 - W % 64) % 64


def preprocess(img1, img2, out_hw=None):
    a = np.asarray(img1, dtype=np.float32)
    b = np.asarray(img2, dtype=np.float32)
    if np.asarray(img1).dtype == np.uint8:
        a, b = a / np.float32(255.0), b / np.float32(255.0)
    mean = np.concatenate([a, b], axis=2).mean(axis=(2, 3), dtype=np.float64).astype(np.float32)[:, :, None, None]
    a, b = a - mean, b - mean
    if out_hw is not None and tuple(out_hw) != a.shape[2:]:
        a, b = bilinear_resize2d(a, *out_hw), bilinear_resize2d(b, *out_hw)
    return a, b, mean


def postprocess(pred, H, W, flip_channels=True, is_flow=True):
    up = cref.upsample(np.asarray(pred, dtype=np.float32), 4)
    if up.shape[2:] != (H, W):
        scale = np.array([H / up.shape[2], W / up.shape[3]], dtype=np.float32).reshape(1, 2, 1, 1)
        up = bilinear_resize2d(up, H, W)
        if is_flow:
            up = up * scale
    out = np.transpose(up, (0, 2, 3, 1))
    return np.ascontiguousarray(out[..., ::-1] if flip_channels else out)
