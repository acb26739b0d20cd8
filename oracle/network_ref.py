"""CPU restatement of the reference's MaskFlownet_S forward (network/MaskFlownet.py:197-315) and of the full cascade
MaskFlownet.hybrid_forward (:443-545).  TEST INFRASTRUCTURE ONLY.

Purpose: (1) checker for the product graph maskflownet_b200.network.MaskFlownetS (tests, smoke()), (2) the CPU arm of
bench.py (`cpu_baseline` and `--impl reference`): the reference's own CPU path cannot run here (MXNet is not installable,
SURVEY.md section 8c), so this port times the same computation -- torch-CPU convolutions (MXNet would use MKL-DNN) plus the
C oracle (oracle/mfn_oracle.c, OpenMP) for Correlation / DeformableConvolution / Upsample / warp.

Written operator by operator after the reference graph; independent of maskflownet_b200 (only parameter *names* are
shared: `conv1a.weight`, `deform5.weight`, `conv5f.bias`, ... = the reference's gluon prefixes).
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as tF

from . import cref

SLOPE = 0.1
STRIDES = {6: 64, 5: 32, 4: 16, 3: 8, 2: 4}


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def maskflownet_s_forward(params: Dict[str, torch.Tensor], im1: torch.Tensor, im2: torch.Tensor, scale: float = 20.0,
                          threads: int = 1, border_mode: int = 0, want_cascade_inputs: bool = False, taps=None,
                          want_srcs: bool = False):
    """params: name -> CPU float tensor; im1, im2: (N,3,H,W) CPU float tensors (already /255 and centralised).
    Returns (predictions[5], [sigmoid(mask2)], c40 or None), numpy-free torch tensors on CPU."""
    P = {k: v.detach().cpu().float() for k, v in params.items()}

    def conv(name, x, stride=1, pad=1, dil=1, act=True):
        y = tF.conv2d(x, P[name + ".weight"], P[name + ".bias"], stride, pad, dil)
        return tF.leaky_relu(y, SLOPE) if act else y

    def pyramid(x):
        feats = []
        for lvl in range(1, 7):
            x = conv(f"conv{lvl}a", x, stride=2)
            x = conv(f"conv{lvl}b", x)
            x = conv(f"conv{lvl}c", x)
            feats.append(x)
        return feats

    def corr(a, b):  # F.Correlation + LeakyReLU  (MaskFlownet.py:216-217)
        c = _t(cref.correlation_forward(a.numpy(), b.numpy(), pad_size=4, max_displacement=4, threads=threads))
        return tF.leaky_relu(c, SLOPE)

    def dense(lvl, x):
        for i in range(5):
            x = torch.cat([conv(f"conv{lvl}_{i}", x), x], dim=1)
        return x

    def up2(x):
        return _t(cref.upsample(x.numpy(), 2))

    c1, c2 = pyramid(im1), pyramid(im2)
    x = dense(6, corr(c1[5], c2[5]))
    flow = conv("pred_flow6", x, act=False)
    mask = conv("pred_mask6", x, act=False)
    flows = [flow]
    for lvl in (5, 4, 3, 2):
        feat = tF.leaky_relu(tF.conv_transpose2d(x, P[f"upfeat{lvl}.weight"], P[f"upfeat{lvl}.bias"], 2, 1), SLOPE)
        flow_up, mask_up = up2(flow), up2(mask)
        off = (flow_up * scale / STRIDES[lvl]).unsqueeze(1).repeat(1, 9, 1, 1, 1).reshape(
            flow_up.shape[0], 18, *flow_up.shape[2:])
        b = P.get(f"deform{lvl}.bias")
        warp = _t(cref.deformable_conv_forward(c2[lvl - 1].numpy(), off.numpy(), P[f"deform{lvl}.weight"].numpy(),
                                               None if b is None else b.numpy(), border_mode=border_mode,
                                               threads=threads))
        trade = conv(f"conv{lvl}f", feat, act=False)
        warp = tF.leaky_relu(warp * torch.sigmoid(mask_up) + trade, SLOPE)
        cv = corr(c1[lvl - 1], warp)
        if taps is not None:  # intermediate tensors of this level (fixtures with realistic value distributions)
            taps[lvl] = dict(c1=c1[lvl - 1], c2=c2[lvl - 1], flow_c=flow, mask_c=mask, tradeoff=trade, warp=warp, corr=cv)
        x = dense(lvl, torch.cat([cv, c1[lvl - 1], feat, flow_up], dim=1))
        flow = flow_up + conv(f"pred_flow{lvl}", x, act=False)
        if lvl > 2:
            mask = conv(f"pred_mask{lvl}", x, act=False)
        else:
            mask = mask_up
        flows.append(flow)
    y = x
    for i, d in zip(range(1, 7), (1, 2, 4, 8, 16, 1)):
        y = conv(f"dc_conv{i}", y, pad=d, dil=d)
    flow = flow + conv("dc_conv7", y, act=False)
    flows[-1] = flow
    preds = [f * scale for f in flows]
    c40 = None
    if want_cascade_inputs:
        mask0 = torch.sigmoid(_t(cref.upsample(mask.numpy(), 4))) - 0.5
        warped = _t(cref.reconstruction2d(im2.numpy(), cref.upsample(flow.numpy(), 4) * scale))
        c40 = torch.cat([warped, mask0], dim=1)
    if want_srcs:   # what the cascade consumes (MaskFlownet.py:304-314); note c2s levels 2 and 3 are IMAGE-1 features (:306)
        c2s = [c2[0], c1[1], c1[2], c2[3], c2[4], c2[5]]
        c30 = torch.cat([im1, torch.zeros_like(im1[:, :1])], dim=1)
        return preds, [torch.sigmoid(mask)], (c1, c2s, flows, c30, c40)
    return preds, [torch.sigmoid(mask)], c40


def maskflownet_forward(params: Dict[str, torch.Tensor], im1: torch.Tensor, im2: torch.Tensor, scale: float = 20.0,
                        threads: int = 1, border_mode: int = 0):
    """The cascade (reference class MaskFlownet, network/MaskFlownet.py:443-545).  params: head parameters under
    `MaskFlownet_S.<name>`, the cascade's own under `<name>`.  Returns (predictions[5], [flow2[:, 0:1]])."""
    head = {k[len("MaskFlownet_S."):]: v for k, v in params.items() if k.startswith("MaskFlownet_S.")}
    P = {k: v.detach().cpu().float() for k, v in params.items() if not k.startswith("MaskFlownet_S.")}
    _, _, srcs = maskflownet_s_forward(head, im1, im2, scale, threads, border_mode, want_cascade_inputs=True, want_srcs=True)
    c1, c2, flows_s, c30, c40 = srcs

    def conv(name, x, stride=1, pad=1, dil=1, act=True):
        y = tF.conv2d(x, P[name + ".weight"], P[name + ".bias"], stride, pad, dil)
        return tF.leaky_relu(y, SLOPE) if act else y

    def pyramid(x):   # conv{L}x (stride 2), conv{L}y, conv{L}z  (:445-457)
        feats = []
        for lvl in range(1, 7):
            x = conv(f"conv{lvl}x", x, stride=2)
            x = conv(f"conv{lvl}y", x)
            x = conv(f"conv{lvl}z", x)
            feats.append(x)
        return feats

    def corr(a, b):   # md = 2 -> 25 channels (:440-441), LeakyReLU (:467 ...)
        c = _t(cref.correlation_forward(a.numpy(), b.numpy(), pad_size=2, max_displacement=2, threads=threads))
        return tF.leaky_relu(c, SLOPE)

    def deform(lvl, x, flow):   # deformL(c2L, repeat(flow*scale/stride, 9)) + LeakyReLU, no mask (:465-466 ...)
        off = (flow * scale / STRIDES[lvl]).unsqueeze(1).repeat(1, 9, 1, 1, 1).reshape(flow.shape[0], 18, *flow.shape[2:])
        b = P.get(f"deform{lvl}.bias")
        w = _t(cref.deformable_conv_forward(x.numpy(), off.numpy(), P[f"deform{lvl}.weight"].numpy(),
                                            None if b is None else b.numpy(), border_mode=border_mode, threads=threads))
        return tF.leaky_relu(w, SLOPE)

    def dense(lvl, x):
        for i in range(5):
            x = torch.cat([conv(f"conv{lvl}_{i}", x), x], dim=1)
        return x

    c3, c4 = pyramid(c30), pyramid(c40)
    flow = flows_s[0]
    x = dense(6, torch.cat([corr(c1[5], deform(6, c2[5], flow)), corr(c3[5], c4[5]), flow], dim=1))
    flow = flow + conv("pred_flow6", x, act=False)
    flows = [flow]
    for i, lvl in enumerate((5, 4, 3, 2)):
        feat = tF.leaky_relu(tF.conv_transpose2d(x, P[f"upfeat{lvl}.weight"], P[f"upfeat{lvl}.bias"], 2, 1), SLOPE)
        flow_up = _t(cref.upsample(flow.numpy(), 2))
        cu_ = corr(c1[lvl - 1], deform(lvl, c2[lvl - 1], flow_up))
        cv_ = corr(c3[lvl - 1], c4[lvl - 1])
        x = dense(lvl, torch.cat([c1[lvl - 1], feat, cu_, cv_, flow_up, flows_s[i + 1]], dim=1))
        flow = flow_up + conv(f"pred_flow{lvl}", x, act=False)
        flows.append(flow)
    y = x
    for i, d in zip(range(1, 7), (1, 2, 4, 8, 16, 1)):
        y = conv(f"dc_conv{i}", y, pad=d, dil=d)
    flow = flow + conv("dc_conv7", y, act=False)
    flows[-1] = flow
    return [f * scale for f in flows], [flow[:, 0:1]]


def predict_flow(params, img1_u8: torch.Tensor, img2_u8: torch.Tensor, threads: int = 1) -> torch.Tensor:
    """What PipelineFlownet.do_batch computes around the network (network/pipeline.py:85-87,99,131,137):
    /255, centralize, forward, Upsample(4) of the finest prediction."""
    a, b = img1_u8.float() / 255.0, img2_u8.float() / 255.0
    mean = torch.cat([a, b], dim=2).mean(dim=(2, 3), keepdim=True)
    preds, _, _ = maskflownet_s_forward(params, a - mean, b - mean, threads=threads)
    return _t(cref.upsample(preds[-1].numpy(), 4))
