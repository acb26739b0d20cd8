"""Training losses of the reference (network/MaskFlownet.py:548-611), on the CUDA Upsample operator.

EpeLossWithMask: per-sample  sum_hw( sqrt(sum_c (pred-label)^2 + eps) * mask ) / sum_hw(mask)     (q-robust variant:
(sum_c |pred-label| + eps)^q).  MultiscaleEpe('upsampling'): sum_i w_i * EpeLossWithMask(Upsample(s_i)(pred_i), flow, mask)
with s = [64, 32, 16, 8, 4] and w = [.005, .01, .02, .08, .32] (network/pipeline.py:39-45).  It is what drives the backward
kernels in BASELINE configs[2] and [4].

`multiscale_epe` on CUDA tensors runs the FUSED kernels of csrc/loss.cu: one pass over the label and
the mask for all five scales, no up-sampled prediction is materialised, the backward is one gather launch (the composition
of operators below needs ~45 launches forward and, through Upsample(64)'s backward, a 128x128 serial gather per coarse
pixel).  `fused=False` (or a custom `upsample=`) runs the operator-by-operator composition -- the reference's own structure,
kept for A/B runs and used on CPU tensors by the tests' oracle leg.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import torch

from . import _lib, ops
from ._lib import MaskflowError

SCALES = (64, 32, 16, 8, 4)
WEIGHTS = (.005, .01, .02, .08, .32)


def epe_loss_with_mask(pred, label, mask, eps: float = 1e-8, q: Optional[float] = None):
    if q is not None:
        loss = ((pred - label).abs().sum(dim=1) + eps) ** q
    else:
        loss = torch.sqrt((pred - label).square().sum(dim=1) + eps)
    loss = loss * mask.squeeze(1)
    return loss.flatten(1).sum(dim=1) / mask.flatten(1).sum(dim=1)


def _host_arrays(tensors, scales, weights):
    n = len(tensors)
    return ((ctypes.c_void_p * n)(*[t.data_ptr() for t in tensors]), (ctypes.c_int * n)(*[int(s) for s in scales]),
            (ctypes.c_float * n)(*[float(w) for w in weights]))


class _MultiscaleEpeFn(torch.autograd.Function):
    """mfn_multiscale_epe_forward / _backward (csrc/loss.cu); gradients flow to the predictions only (label and mask are data)."""

    @staticmethod
    def forward(ctx, flow, mask, scales, weights, eps, q, *preds):
        N, _, H, W = flow.shape
        dev = flow.device
        pa, sa, wa = _host_arrays(preds, scales, weights)
        loss = torch.empty(N, device=dev, dtype=torch.float32)
        msum = torch.empty(N, device=dev, dtype=torch.float32)
        wsb = int(_lib.lib().mfn_multiscale_epe_workspace_bytes(N))
        ws = torch.empty(wsb // 4, device=dev, dtype=torch.float32)
        ops._call("mfn_multiscale_epe_forward", dev, ops._p(flow), ops._p(mask), pa, sa, wa, len(preds), float(eps), float(q),
                  ops._p(loss), ops._p(msum), ops._p(ws), wsb, N, H, W)
        ctx.cfg = (tuple(scales), tuple(weights), float(eps), float(q))
        ctx.save_for_backward(flow, mask, msum, *preds)
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        flow, mask, msum, *preds = ctx.saved_tensors
        scales, weights, eps, q = ctx.cfg
        N, _, H, W = flow.shape
        grads = [torch.empty_like(p) for p in preds]
        pa, sa, wa = _host_arrays(preds, scales, weights)
        ga = (ctypes.c_void_p * len(grads))(*[t.data_ptr() for t in grads])
        g = g.contiguous().float()
        ops._call("mfn_multiscale_epe_backward", flow.device, ops._p(flow), ops._p(mask), pa, sa, wa, len(preds), eps, q,
                  ops._p(g), ops._p(msum), ga, N, H, W)
        return (None, None, None, None, None, None, *grads)


def multiscale_epe_fused(flow, mask, predictions: Sequence[torch.Tensor], scales=SCALES, weights=WEIGHTS, eps: float = 1e-8,
                         q: Optional[float] = None) -> torch.Tensor:
    """The fused MultiscaleEpe('upsampling'): CUDA float32 tensors only (no fallback).  Returns the per-sample losses (N,)."""
    flow, mask = ops._chk(flow, "multiscale_epe.flow"), ops._chk(mask, "multiscale_epe.mask")
    preds = [ops._chk(p, "multiscale_epe.prediction") for p in predictions]
    N, C, H, W = flow.shape
    if C != 2 or tuple(mask.shape) != (N, 1, H, W):
        raise MaskflowError(f"multiscale_epe: flow must be (N,2,H,W) and mask (N,1,H,W); got {tuple(flow.shape)}, {tuple(mask.shape)}")
    if not (len(preds) == len(scales) == len(weights)):
        raise MaskflowError("multiscale_epe: one scale and one weight per prediction")
    for p, s in zip(preds, scales):
        if H % s or W % s or tuple(p.shape) != (N, 2, H // s, W // s):
            raise MaskflowError(f"multiscale_epe: prediction {tuple(p.shape)} x{s} does not up-sample to the label {tuple(flow.shape)}")
    if flow.requires_grad or mask.requires_grad:
        raise MaskflowError("multiscale_epe: the label and the mask are data (no gradient is computed for them)")
    return _MultiscaleEpeFn.apply(flow, mask, tuple(scales), tuple(weights), eps, -1.0 if q is None else float(q), *preds)


def multiscale_epe(flow, mask, predictions: Sequence[torch.Tensor], scales=SCALES, weights=WEIGHTS, eps: float = 1e-8,
                   q: Optional[float] = None, upsample=None, fused: Optional[bool] = None):
    """flow (N,2,H,W) ground truth in (y,x) order, mask (N,1,H,W); returns the per-sample loss vector (N,).
    CUDA tensors take the fused kernels unless fused=False or a custom `upsample` is given."""
    if fused is None:
        fused = upsample is None and flow.is_cuda
    if fused:
        return multiscale_epe_fused(flow, mask, predictions, scales, weights, eps, q)
    upsample = ops.upsample if upsample is None else upsample
    total = 0
    for p, w, s in zip(predictions, weights, scales):
        total = total + w * epe_loss_with_mask(upsample(p, s), flow, mask, eps, q)
    return total
