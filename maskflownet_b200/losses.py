"""Training losses of the reference (network/MaskFlownet.py:548-611), on the CUDA Upsample operator.

EpeLossWithMask: per-sample  sum_hw( sqrt(sum_c (pred-label)^2 + eps) * mask ) / sum_hw(mask)     (q-robust variant:
(sum_c |pred-label| + eps)^q).  MultiscaleEpe('upsampling'): sum_i w_i * EpeLossWithMask(Upsample(s_i)(pred_i), flow, mask)
with s = [64, 32, 16, 8, 4] and w = [.005, .01, .02, .08, .32] (network/pipeline.py:39-45).  Out of the hot-path scope
proper (plain element-wise torch), but it is what drives the backward kernels in BASELINE configs[2] and [4].
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops

SCALES = (64, 32, 16, 8, 4)
WEIGHTS = (.005, .01, .02, .08, .32)


def epe_loss_with_mask(pred, label, mask, eps: float = 1e-8, q: Optional[float] = None):
    if q is not None:
        loss = ((pred - label).abs().sum(dim=1) + eps) ** q
    else:
        loss = torch.sqrt((pred - label).square().sum(dim=1) + eps)
    loss = loss * mask.squeeze(1)
    return loss.flatten(1).sum(dim=1) / mask.flatten(1).sum(dim=1)


def multiscale_epe(flow, mask, predictions: Sequence[torch.Tensor], scales=SCALES, weights=WEIGHTS, eps: float = 1e-8,
                   q: Optional[float] = None, upsample=ops.upsample):
    """flow (N,2,H,W) ground truth in (y,x) order, mask (N,1,H,W); returns the per-sample loss vector (N,)."""
    total = 0
    for p, w, s in zip(predictions, weights, scales):
        total = total + w * epe_loss_with_mask(upsample(p, s), flow, mask, eps, q)
    return total
