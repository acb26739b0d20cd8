"""Host-side mirror of the reference's model graphs, calling the fused B200 operators.

`MaskFlownetS` computes what `MaskFlownet_S.hybrid_forward` computes (network/MaskFlownet.py:197-315) and
`MaskFlownet` what the cascade computes (:443-545), but is organised around the hot path instead of transcribing the
unrolled reference code: one loop over pyramid levels in which

    Upsample(2) flow/mask + offset build + DeformableConvolution + sigmoid-mask multiply + trade-off + LeakyReLU
        (MaskFlownet.py:228-233)                                   -> one launch of ops.warp_mask      (K3)
    Correlation + LeakyReLU (:234-235), written straight into its slot of the decoder's concat buffer (:236)
                                                                    -> one launch of ops.correlation    (K1)
    Upsample(4) + GridGenerator + BilinearSampler + sigmoid-0.5 + concat (:308-313)
                                                                    -> one launch of ops.image_warp_concat (K5)

The dense 3x3 / transposed convolutions (about 98 % of the FLOPs; row N2) run on the tcgen05 / TMEM
kernel of csrc/conv3x3_umma.cu: f32 in / out, bf16 hi+lo split operands, fp32 accumulation.  At inference (`_fast`) with
in-place concat buffers and fused heads; with gradients enabled (`train_tc_forward`, default) every 3x3 convolution still
runs its FORWARD on that kernel and its backward through aten.convolution_backward (ops.conv3x3_train), the transposed
convolutions and the concats through torch autograd.  Sub-module names equal the reference's gluon prefixes (conv1a ... deform5, conv5f,
dc_conv7 ...) so that shipped .params checkpoints map by name (maskflownet_b200.params).

All flows are (y, x)-ordered and in units of pixels/scale, as in the reference (pipeline.py:105, MaskFlownet.py:69).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as tF

from . import ops

PYRAMID_CH = {1: 16, 2: 32, 3: 64, 4: 96, 5: 128, 6: 196}   # network/MaskFlownet.py:79-96
DECODER_CH = (128, 128, 96, 64, 32)                           # convL_0 .. convL_4 (:102-130)
STRIDES = {6: 64, 5: 32, 4: 16, 3: 8, 2: 4}                   # self.strides (:71)
SLOPE = 0.1


def _conv(cin, cout, k=3, s=1, p=1, d=1):
    return nn.Conv2d(cin, cout, k, s, p, d)


def msra_prelu_init_(module: nn.Module, slope: float = SLOPE, seed: Optional[int] = None) -> None:
    """MSRAPrelu(factor_type='avg', slope) for weights, zeros for biases -- the reference's initialiser
    (network/pipeline.py:26)."""
    gen = torch.Generator().manual_seed(seed) if seed is not None else None
    for name, p in module.named_parameters():
        if p.dim() < 2:
            with torch.no_grad():
                p.zero_()
            continue
        hw = 1
        for d in p.shape[2:]:
            hw *= d
        fan_in, fan_out = p.shape[1] * hw, p.shape[0] * hw
        std = math.sqrt(2.0 / ((1 + slope ** 2) * (fan_in + fan_out) / 2.0))
        with torch.no_grad():
            p.copy_((torch.randn(p.shape, generator=gen) * std).to(p.device))


class DeformParams(nn.Module):
    """Weights of one layer.DeformableConv2D block (network/layer.py:32-110): weight (F, C, 3, 3) and bias (F)."""

    def __init__(self, channels: int, use_bias: bool = True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(channels, channels, 3, 3))
        self.bias = nn.Parameter(torch.zeros(channels)) if use_bias else None


class _Slab:
    """Inference view of a dense block's output inside its concat buffer: x = buf[:, c0:], and -- when the block's last
    convolution also produced them -- the linear heads' partial sums over everything but that convolution's own output
    in buf[:, :c0] (see _FlowNetBase._dense_inplace)."""

    def __init__(self, buf: torch.Tensor, c0: int, last_oc: int):
        self.buf, self.c0, self.last_oc = buf, c0, last_oc

    @property
    def channels(self) -> int:
        return self.buf.shape[1] - self.c0

    def tensor(self) -> torch.Tensor:
        return self.buf[:, self.c0:]


class _FlowNetBase(nn.Module):
    fuse_heads = True   # inference: pred_flow / pred_mask over the block input ride on conv{L}_4's input pass
    use_resample_warp = True   # inference: K3 through linearity (ops.warp_mask(resample=True)) at every level
    use_tc_conv = True   # inference: decoder / context 3x3 convolutions on the fp32-accurate tensor-core kernel (row N2)
    # training (grad enabled): the 3x3 convolutions still run their FORWARD on the tcgen05 kernel (ops.conv3x3_train: bias +
    # LeakyReLU fused, the output doubles as the activation mask), the BACKWARD is aten.convolution_backward (cuDNN fp32).
    # Measured on BASELINE configs[2] (batch 8, 512x384, fwd + bwd): 66.4 -> 52.7 ms per step.  False: cuDNN both ways.
    train_tc_forward = True

    def _packed(self, name):
        """Packed split-bf16 weight image of conv `name`, rebuilt when the parameter changes."""
        cache = self.__dict__.setdefault("_pack_cache", {})
        w = getattr(self, name).weight
        key = (w.data_ptr(), w._version)
        hit = cache.get(name)
        if hit is None or hit[0] != key:
            cache[name] = (key, ops.conv3x3_pack(w))
        return cache[name][1]

    def _packed_fn(self, key, params, build):
        """Cached derived tensors (fused / re-arranged weight images), rebuilt when any source parameter changes."""
        cache = self.__dict__.setdefault("_pack_cache", {})
        ver = tuple((p.data_ptr(), p._version) for p in params if p is not None)
        hit = cache.get(key)
        if hit is None or hit[0] != ver:
            cache[key] = (ver, build())
        return cache[key][1]

    def _head_convs(self, lvl, with_mask):
        pf = getattr(self, f"pred_flow{lvl}")
        pm = getattr(self, f"pred_mask{lvl}") if with_mask and hasattr(self, f"pred_mask{lvl}") else None
        return [pf] + ([pm] if pm is not None else [])

    def _heads(self, lvl, x, with_mask):
        """pred_flow{lvl} (2 channels) and pred_mask{lvl} (1 channel) read the same block output
        (network/MaskFlownet.py:224, 226 ...): inference runs them as ONE 3-output convolution, no activation -- and, when
        the dense block left their partial sums (a _Slab with c0 > 0), only over the block's last 32 channels."""
        convs = self._head_convs(lvl, with_mask)
        pm = convs[1] if len(convs) > 1 else None
        if not isinstance(x, _Slab):
            if not self._fast(x):
                return (self._conv_act(f"pred_flow{lvl}", x, 1.0),
                        self._conv_act(f"pred_mask{lvl}", x, 1.0) if pm is not None else None)
            x = _Slab(x, 0, 0)
        nh = 2 + (1 if pm is not None else 0)
        buf, c0 = x.buf, x.c0
        N, _, H, W = buf.shape
        if c0 == nh and x.last_oc:   # partial sums present: finish with the tiny convolution over conv{L}_4's output
            oc = x.last_oc

            def build_tail():
                w = torch.cat([c.weight.detach()[:, :oc] for c in convs], dim=0).contiguous()
                b = torch.cat([c.bias.detach() for c in convs], dim=0).contiguous()
                return ops.conv3x3_pack(w), b
            packed, b = self._packed_fn(f"heads_tail{lvl}", [p for c in convs for p in (c.weight, c.bias)], build_tail)
            y = torch.empty((N, nh, H, W), device=buf.device, dtype=torch.float32)
            ops.conv3x3_slices(buf, c0, oc, packed, b, y, 0, nh, 1.0)
            y += buf[:, :nh]
        else:
            def build():
                w = torch.cat([c.weight.detach() for c in convs], dim=0).contiguous()
                b = torch.cat([c.bias.detach() for c in convs], dim=0).contiguous()
                return ops.conv3x3_pack(w), b
            packed, b = self._packed_fn(f"heads{lvl}", [p for c in convs for p in (c.weight, c.bias)], build)
            y = torch.empty((N, nh, H, W), device=buf.device, dtype=torch.float32)
            ops.conv3x3_slices(buf, c0, x.channels, packed, b, y, 0, nh, 1.0)
        if pm is None:
            return y, None
        return y[:, :2].contiguous(), y[:, 2:3].contiguous()

    def _upfeat(self, lvl, x):
        """feat = LeakyReLU(upfeat{lvl}(x)): ConvTranspose2d(4, 2, 1) as a 3x3 convolution + depth-to-space on the
        tensor-core kernel (ops.conv_transpose4x4_pack)."""
        up = getattr(self, f"upfeat{lvl}")
        if not isinstance(x, _Slab):
            if not self._fast(x):
                return tF.leaky_relu(up(x), SLOPE)
            x = _Slab(x, 0, 0)
        packed = self._packed_fn(f"upfeat{lvl}", [up.weight], lambda: ops.conv_transpose4x4_pack(up.weight))
        N, _, H, W = x.buf.shape
        F = up.out_channels
        out = torch.empty((N, F, 2 * H, 2 * W), device=x.buf.device, dtype=torch.float32)
        ops.conv3x3_slices(x.buf, x.c0, x.channels, packed, up.bias, out, 0, 4 * F, SLOPE, depth_to_space=True)
        return out

    def _plain(self, name, x):
        """3x3 convolution without activation (conv{L}f, dc_conv7)."""
        conv = getattr(self, name)
        if not self._fast(x):
            return self._conv_act(name, x, 1.0)
        return ops.conv3x3(x, self._packed(name), conv.bias, conv.out_channels, 1.0, conv.dilation[0])

    def _conv_act(self, name, x, slope):
        """Autograd path of one 3x3 convolution (+ LeakyReLU when slope != 1): torch.nn.functional (cuDNN both ways), or --
        train_tc_forward -- the tensor-core forward with the cuDNN backward."""
        conv = getattr(self, name)
        if (self.train_tc_forward and x.is_cuda and conv.kernel_size == (3, 3) and conv.padding == conv.dilation
                and conv.stride[0] == conv.stride[1] and conv.groups == 1):
            return ops.conv3x3_train(x, conv.weight, conv.bias, self._packed(name), slope, conv.dilation[0], conv.stride[0])
        y = conv(x)
        return y if slope == 1.0 else tF.leaky_relu(y, slope)

    def _fast(self, x):
        return self.use_tc_conv and x.is_cuda and not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()))

    def _pyramid(self, x, names):
        """Six levels of (3x3 stride-2, 3x3, 3x3) convolutions + LeakyReLU (network/MaskFlownet.py:200-202).  Inference:
        all eighteen run on the tcgen05 convolution kernel (bias + activation fused)."""
        feats = []
        fast = self._fast(x)
        for lvl in range(1, 7):
            for j, sfx in enumerate(names):
                name = f"conv{lvl}{sfx}"
                conv = getattr(self, name)
                if fast:
                    x = ops.conv3x3(x, self._packed(name), conv.bias, conv.out_channels, SLOPE, 1, 2 if j == 0 else 1)
                else:
                    x = self._conv_act(name, x, SLOPE)
            feats.append(x)
        return feats  # [c?1 .. c?6]

    def _pyramid_pair(self, im1, im2, names):
        """Both images through the shared pyramid; inference batches them into one pass (half the launches)."""
        if self._fast(im1):
            n = im1.shape[0]
            f = self._pyramid(torch.cat([im1, im2], dim=0), names)
            return [t[:n] for t in f], [t[n:] for t in f]
        return self._pyramid(im1, names), self._pyramid(im2, names)

    def _dense(self, lvl, x):
        """x = concat(leaky(conv_i(x)), x) five times (network/MaskFlownet.py:219-223 ...).  Inference: one pre-allocated
        buffer, every convolution reads its input channels in place and writes its output in front of them."""
        if self._fast(x):
            N, Cb, H, W = x.shape
            tot = sum(DECODER_CH) + self._heads_front(lvl)
            buf = torch.empty((N, tot + Cb, H, W), device=x.device, dtype=torch.float32)
            buf[:, tot:].copy_(x)
            return self._dense_inplace(lvl, buf, tot)
        for i in range(5):
            x = torch.cat([self._conv_act(f"conv{lvl}_{i}", x, SLOPE), x], dim=1)
        return x

    def _heads_front(self, lvl) -> int:
        """Extra leading channels of the block's concat buffer that receive the heads' partial sums (0 = not fused)."""
        if not self.fuse_heads:
            return 0
        return 2 + (1 if hasattr(self, f"pred_mask{lvl}") else 0)

    def _dense_inplace(self, lvl, buf, off):
        """buf[:, off:] holds the block's input; fills buf front to back and returns the block output as a _Slab.
        With fuse_heads the last convolution also carries the nh linear head channels over ITS input (everything the heads
        read except that convolution's own 32 output channels): weights [W_heads[:, 32:] ; W_4], written as channels
        [partial (nh, no activation) | conv{lvl}_4 (32, LeakyReLU)] -- one pass over the ~550-channel input instead of two."""
        Ctot = buf.shape[1]
        nh = self._heads_front(lvl)
        for i, oc in enumerate(DECODER_CH):
            conv = getattr(self, f"conv{lvl}_{i}")
            if i == len(DECODER_CH) - 1 and nh:
                heads = self._head_convs(lvl, True)

                def build():
                    w = torch.cat([h.weight.detach()[:, oc:] for h in heads] + [conv.weight.detach()], dim=0).contiguous()
                    b = torch.cat([torch.zeros(nh, device=w.device), conv.bias.detach()]).contiguous()
                    return ops.conv3x3_pack(w), b
                packed, b = self._packed_fn(f"conv{lvl}_4+heads", [conv.weight, conv.bias] + [h.weight for h in heads], build)
                assert off - oc - nh == 0
                ops.conv3x3_slices(buf, off, Ctot - off, packed, b, buf, 0, oc + nh, SLOPE, linear_prefix=nh)
            else:
                ops.conv3x3_slices(buf, off, Ctot - off, self._packed(f"conv{lvl}_{i}"), conv.bias, buf, off - oc, oc, SLOPE)
            off -= oc
        return _Slab(buf, nh, DECODER_CH[-1] if nh else 0)

    def _context(self, x):
        fast = isinstance(x, _Slab) or self._fast(x)
        for i in range(1, 7):
            conv = getattr(self, f"dc_conv{i}")
            if isinstance(x, _Slab):      # first layer reads the block output in place
                N, _, H, W = x.buf.shape
                y = torch.empty((N, conv.out_channels, H, W), device=x.buf.device, dtype=torch.float32)
                ops.conv3x3_slices(x.buf, x.c0, x.channels, self._packed(f"dc_conv{i}"), conv.bias, y, 0, conv.out_channels,
                                   SLOPE, conv.dilation[0])
                x = y
            elif fast:
                x = ops.conv3x3(x, self._packed(f"dc_conv{i}"), conv.bias, conv.out_channels, SLOPE, conv.dilation[0])
            else:
                x = self._conv_act(f"dc_conv{i}", x, SLOPE)
        return self._plain("dc_conv7", x)

    def _make_decoder(self, in_ch: Dict[int, int], with_mask: bool, upfeat_ch):
        for lvl in (6, 5, 4, 3, 2):
            c = in_ch[lvl]
            for i, oc in enumerate(DECODER_CH):
                setattr(self, f"conv{lvl}_{i}", _conv(c, oc))
                c += oc
            setattr(self, f"pred_flow{lvl}", _conv(c, 2))
            if with_mask and lvl > 2:
                setattr(self, f"pred_mask{lvl}", _conv(c, 1))
            if lvl > 2:
                setattr(self, f"upfeat{lvl - 1}", nn.ConvTranspose2d(c, upfeat_ch[6 - lvl], 4, 2, 1))
        c2 = in_ch[2] + sum(DECODER_CH)
        dil = (1, 2, 4, 8, 16, 1)
        chs = (128, 128, 128, 96, 64, 32)
        c = c2
        for i in range(6):
            setattr(self, f"dc_conv{i + 1}", _conv(c, chs[i], 3, 1, dil[i], dil[i]))
            c = chs[i]
        self.dc_conv7 = _conv(c, 2)


class MaskFlownetS(_FlowNetBase):
    """MaskFlownet-S (reference class MaskFlownet_S, network/MaskFlownet.py:66-315)."""

    def __init__(self, flow_multiplier: float = 1.0, deform_bias: bool = True, upfeat_ch=(16, 16, 16, 16),
                 border_mode: int = ops.BORDER_MXNET15):
        super().__init__()
        self.scale = 20.0 * flow_multiplier
        self.md = 4
        self.border_mode = border_mode
        self.upfeat_ch = tuple(upfeat_ch)
        self.event_hook = None  # optional callable(kind, level, 0|1): bench.py brackets kernels with CUDA events
        cin = 3
        for lvl in range(1, 7):
            co = PYRAMID_CH[lvl]
            setattr(self, f"conv{lvl}a", _conv(cin, co, 3, 2))
            setattr(self, f"conv{lvl}b", _conv(co, co))
            setattr(self, f"conv{lvl}c", _conv(co, co))
            cin = co
        D = (2 * self.md + 1) ** 2
        in_ch = {6: D}
        for i, lvl in enumerate((5, 4, 3, 2)):
            in_ch[lvl] = D + PYRAMID_CH[lvl] + self.upfeat_ch[i] + 2
        self._make_decoder(in_ch, with_mask=True, upfeat_ch=self.upfeat_ch)
        for i, lvl in enumerate((5, 4, 3, 2)):
            setattr(self, f"deform{lvl}", DeformParams(PYRAMID_CH[lvl], deform_bias))
            setattr(self, f"conv{lvl}f", _conv(self.upfeat_ch[i], PYRAMID_CH[lvl]))
        msra_prelu_init_(self)

    # one correlation + its consumers' concat buffer: [corr | extras...]
    def _corr_block(self, lvl, f1, f2, extras: List[torch.Tensor]):
        N, _, H, W = f1.shape
        D = (2 * self.md + 1) ** 2
        # same gate as every other layer (_fast): with grad enabled and ANY trainable parameter the autograd operators
        # run, so a frozen pyramid + trainable decoder still trains conv{lvl}_0..4
        if not self._fast(f1):
            corr = ops.correlation(f1, f2, pad_size=self.md, max_displacement=self.md, leaky_slope=SLOPE)
            return self._dense(lvl, torch.cat([corr] + extras, dim=1) if extras else corr)
        tot = D + sum(e.shape[1] for e in extras)
        # room for the dense block's outputs (written in place) and the heads' partial sums
        front = (sum(DECODER_CH) + self._heads_front(lvl)) if self.use_tc_conv else 0
        buf = torch.empty((N, front + tot, H, W), device=f1.device, dtype=torch.float32)
        hook = self.event_hook
        if hook is not None:
            hook("corr", lvl, 0)
        ops.correlation(f1, f2, pad_size=self.md, max_displacement=self.md, leaky_slope=SLOPE,
                        out=buf[:, front:front + D])
        if hook is not None:
            hook("corr", lvl, 1)
        c = front + D
        for e in extras:
            buf[:, c:c + e.shape[1]].copy_(e)
            c += e.shape[1]
        if front:
            return self._dense_inplace(lvl, buf, front)
        return self._dense(lvl, buf)

    def forward(self, im1: torch.Tensor, im2: torch.Tensor, want_cascade_inputs: bool = False):
        """Returns (predictions [flow6..flow2, each * scale], [sigmoid(mask2)], srcs or None) like the reference
        (network/MaskFlownet.py:302-315).  srcs (needed only by the cascade) is built when want_cascade_inputs."""
        c1, c2 = self._pyramid_pair(im1, im2, "abc")
        x = self._corr_block(6, c1[5], c2[5], [])   # correlation + dense block
        flow, mask = self._heads(6, x, True)
        flows = [flow]
        for lvl in (5, 4, 3, 2):
            feat = self._upfeat(lvl, x)
            dp = getattr(self, f"deform{lvl}")
            trade = self._plain(f"conv{lvl}f", feat)
            if self.event_hook is not None:
                self.event_hook("warp", lvl, 0)
            warp, flow_up, _ = ops.warp_mask(c2[lvl - 1], flow, mask, dp.weight, dp.bias, trade, self.scale,
                                             float(STRIDES[lvl]), 2, SLOPE, self.border_mode,
                                             # inference: every level is evaluated exactly through linearity (extended 3x3
                                             # convolution on tcgen05 + bilinear re-sampling + border-band tables, warp_lin.cu)
                                             packed_weight=self._packed(f"deform{lvl}") if self._fast(flow) else None,
                                             resample=self.use_resample_warp)
            if self.event_hook is not None:
                self.event_hook("warp", lvl, 1)
            x = self._corr_block(lvl, c1[lvl - 1], warp, [c1[lvl - 1], feat, flow_up])
            dflow, m = self._heads(lvl, x, lvl > 2)
            flow = flow_up + dflow
            if lvl > 2:
                mask = m
            else:
                mask_up2 = _  # Upsample(2)(mask3): the level-2 occlusion mask (network/MaskFlownet.py:283)
            flows.append(flow)
        flows[-1] = flow = flow + self._context(x)
        preds = [f * self.scale for f in flows]
        occ = [torch.sigmoid(mask_up2)]
        srcs = None
        if want_cascade_inputs:
            c30, c40 = ops.image_warp_concat(im1, im2, flow, mask_up2, self.scale)
            # quirk kept from the reference: levels 2 and 3 of c2s carry IMAGE-1 features (MaskFlownet.py:306)
            c2s = [c2[0], c1[1], c1[2], c2[3], c2[4], c2[5]]
            srcs = (c1, c2s, flows, c30, c40)
        return preds, occ, srcs


class MaskFlownet(_FlowNetBase):
    """Full cascade (reference class MaskFlownet, network/MaskFlownet.py:318-545): the S head plus a second, dual
    pyramid on [im1; 0] and [warp(im2); mask] with md=2 correlations."""

    def __init__(self, flow_multiplier: float = 1.0, deform_bias: bool = True, upfeat_ch=(16, 16, 16, 16),
                 border_mode: int = ops.BORDER_MXNET15):
        super().__init__()
        self.scale = 20.0 * flow_multiplier
        self.md = 2
        self.border_mode = border_mode
        self.MaskFlownet_S = MaskFlownetS(flow_multiplier, deform_bias, upfeat_ch, border_mode)
        cin = 4
        for lvl in range(1, 7):
            co = PYRAMID_CH[lvl]
            setattr(self, f"conv{lvl}x", _conv(cin, co, 3, 2))
            setattr(self, f"conv{lvl}y", _conv(co, co))
            setattr(self, f"conv{lvl}z", _conv(co, co))
            cin = co
        D = (2 * self.md + 1) ** 2
        in_ch = {6: 2 * D + 2}
        for i, lvl in enumerate((5, 4, 3, 2)):
            in_ch[lvl] = PYRAMID_CH[lvl] + upfeat_ch[i] + 2 * D + 4
        self._make_decoder(in_ch, with_mask=False, upfeat_ch=tuple(upfeat_ch))
        for lvl in (6, 5, 4, 3, 2):
            setattr(self, f"deform{lvl}", DeformParams(PYRAMID_CH[lvl], deform_bias))
        msra_prelu_init_(self)

    def _corr(self, a, b):
        return ops.correlation(a, b, pad_size=self.md, max_displacement=self.md, leaky_slope=SLOPE)

    def forward(self, im1, im2):
        _, _, srcs = self.MaskFlownet_S(im1, im2, want_cascade_inputs=True)
        c1, c2, flows_s, c30, c40 = srcs
        c3 = self._pyramid(c30, "xyz")
        c4 = self._pyramid(c40, "xyz")
        flow = flows_s[0]
        dp = self.deform6
        fast = self._fast(flow)
        warp, _, _ = ops.warp_mask(c2[5], flow, None, dp.weight, dp.bias, None, self.scale, float(STRIDES[6]), 1,
                                   SLOPE, self.border_mode, packed_weight=self._packed("deform6") if fast else None,
                                   resample=self.use_resample_warp)
        x = self._dense(6, torch.cat([self._corr(c1[5], warp), self._corr(c3[5], c4[5]), flow], dim=1))
        flow = flow + self._heads(6, x, False)[0]
        flows = [flow]
        for i, lvl in enumerate((5, 4, 3, 2)):
            feat = self._upfeat(lvl, x)
            dp = getattr(self, f"deform{lvl}")
            warp, flow_up, _ = ops.warp_mask(c2[lvl - 1], flow, None, dp.weight, dp.bias, None, self.scale,
                                             float(STRIDES[lvl]), 2, SLOPE, self.border_mode,
                                             packed_weight=self._packed(f"deform{lvl}") if fast else None,
                                             resample=self.use_resample_warp)
            x = self._dense(lvl, torch.cat([c1[lvl - 1], feat, self._corr(c1[lvl - 1], warp),
                                            self._corr(c3[lvl - 1], c4[lvl - 1]), flow_up, flows_s[i + 1]], dim=1))
            flow = flow_up + self._heads(lvl, x, False)[0]
            flows.append(flow)
        flows[-1] = flow = flow + self._context(x)
        return [f * self.scale for f in flows], [flow[:, 0:1]], []


# ---------------------------------------------------------------------------------------------------------------
# the step either side of the network: what PipelineFlownet.do_batch does around it (network/pipeline.py:85-87,117-147)
# ---------------------------------------------------------------------------------------------------------------
def centralize(img1: torch.Tensor, img2: torch.Tensor):
    """Subtract the per-sample RGB mean over both images (network/pipeline.py:85-87)."""
    mean = torch.cat([img1, img2], dim=2).mean(dim=(2, 3), keepdim=True)
    return img1 - mean, img2 - mean, mean


@torch.no_grad()
def predict_flow(net: nn.Module, img1_u8: torch.Tensor, img2_u8: torch.Tensor) -> torch.Tensor:
    """uint8 image pairs (N,3,H,W), H and W multiples of 64 -> full-resolution flow (N,2,H,W), (y,x)-ordered, in pixels:
    /255, centralize, network, Upsample(4) of the finest prediction (network/pipeline.py:99,117-138)."""
    if img1_u8.is_cuda and img1_u8.dtype == torch.uint8 and img1_u8.is_contiguous() and img2_u8.is_contiguous():
        a, b, _ = ops.preprocess(img1_u8, img2_u8)        # /255 + centralize in one fused op (csrc/prepost.cu)
    else:
        a, b, _ = centralize(img1_u8.float() / 255.0, img2_u8.float() / 255.0)
    preds = net(a, b)[0]
    return ops.upsample(preds[-1], 4)


@torch.no_grad()
def predict(net: nn.Module, img1: torch.Tensor, img2: torch.Tensor, resize=None):
    """PipelineFlownet.predict for one batch (network/pipeline.py:189-223), fused: uint8 (or [0,1] float) pairs (N,3,H,W) of
    ANY size -> /255 + centralize + BilinearResize2D to multiples of 64 (one op) -> network -> Upsample(4) + resize back +
    per-channel rescale + NHWC + (y,x)->(x,y) flip (one op).  Returns (flow (N,H,W,2) in (x,y) pixels -- the .flo layout,
    occlusion mask (N,H,W,1))."""
    N, _, H, W = img1.shape
    a, b, _ = ops.preprocess(img1, img2, ops.padded_size(H, W, resize))
    preds, occ, _ = net(a, b)
    flow = ops.postprocess(preds[-1], H, W, flip_channels=True, is_flow=True)
    mask = ops.postprocess(occ[0], H, W, flip_channels=False, is_flow=False) if occ and occ[0].shape[1] == 1 and \
        occ[0].shape[2] * 4 == a.shape[2] else None
    return flow, mask


class FlowPredictor:
    """predict_flow captured in a CUDA graph: one graph per input shape, static uint8 input buffers, one cudaGraphLaunch
    per call (the eager step is ~115 dependent launches; the graph removes the launch gaps between them).
    Weights are read through the packed images cached in the model: call invalidate() after changing parameters.
    The returned tensor is the graph's STATIC output buffer: the next call overwrites it -- clone() it (or copy it to the
    host) before calling again if the previous result is still needed."""

    def __init__(self, net: nn.Module, warmup: int = 2):
        self.net, self.warmup, self._graphs = net, warmup, {}

    def invalidate(self) -> None:
        self._graphs.clear()

    @torch.no_grad()
    def __call__(self, img1_u8: torch.Tensor, img2_u8: torch.Tensor) -> torch.Tensor:
        dev = next(self.net.parameters()).device      # inputs may live on the host (pinned): the static buffers do not
        key = (tuple(img1_u8.shape), img1_u8.dtype)
        entry = self._graphs.get(key)
        if entry is None:
            in1 = torch.empty(img1_u8.shape, dtype=img1_u8.dtype, device=dev)
            in2 = torch.empty(img2_u8.shape, dtype=img2_u8.dtype, device=dev)
            in1.copy_(img1_u8)
            in2.copy_(img2_u8)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up off the capture: weight packing, kernel attributes, cuDNN-free path
                for _ in range(self.warmup):
                    predict_flow(self.net, in1, in2)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = predict_flow(self.net, in1, in2)
            entry = self._graphs[key] = (graph, in1, in2, out)
        graph, in1, in2, out = entry
        in1.copy_(img1_u8, non_blocking=True)
        in2.copy_(img2_u8, non_blocking=True)
        graph.replay()
        return out



class PipelinedFlowPredictor:
    """Serving loop around FlowPredictor for HOST buffers: every call enqueues (asynchronously)
        pinned uint8 images --H2D (copy stream)--> staging --D2D--> graph inputs --graph replay--> flow --D2D--> staging
        --D2H (copy stream)--> pinned fp32 flow
    with `depth` staging slots, so the H2D copy of request i+1 and the D2H copy of result i-1 run under the forward of
    request i (PCIe is full duplex; 22 MB in / 29 MB out per batch of 8 at 1024x448 take ~0.4 / ~0.5 ms of a ~9 ms forward).
    Results are complete after synchronize() (or after waiting on the event infer() returns)."""

    def __init__(self, net: nn.Module, depth: int = 2):
        self.pred = FlowPredictor(net)
        self.depth = depth
        self._slots = None
        self._i = 0

    def _setup(self, img1, dev):
        shp = tuple(img1.shape)
        N, _, H, W = shp
        self.h2d, self.d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        self._slots = []
        for _ in range(self.depth):
            self._slots.append({
                "in1": torch.empty(shp, dtype=img1.dtype, device=dev), "in2": torch.empty(shp, dtype=img1.dtype, device=dev),
                "out": torch.empty((N, 2, H, W), dtype=torch.float32, device=dev),
                "ev_h2d": torch.cuda.Event(), "ev_in_free": torch.cuda.Event(), "ev_out": torch.cuda.Event(),
                "ev_out_free": torch.cuda.Event(), "used": False})

    @torch.no_grad()
    def infer(self, img1_host: torch.Tensor, img2_host: torch.Tensor, out_host: torch.Tensor) -> torch.cuda.Event:
        dev = next(self.pred.net.parameters()).device
        if self._slots is None:
            self._setup(img1_host, dev)
        s = self._slots[self._i % self.depth]
        self._i += 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(self.h2d):
            if s["used"]:
                self.h2d.wait_event(s["ev_in_free"])
            s["in1"].copy_(img1_host, non_blocking=True)
            s["in2"].copy_(img2_host, non_blocking=True)
            s["ev_h2d"].record(self.h2d)
        cur.wait_event(s["ev_h2d"])
        flow = self.pred(s["in1"], s["in2"])          # D2D into the graph's static inputs + replay
        s["ev_in_free"].record(cur)
        if s["used"]:
            cur.wait_event(s["ev_out_free"])
        s["out"].copy_(flow, non_blocking=True)
        s["ev_out"].record(cur)
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(s["ev_out"])
            out_host.copy_(s["out"], non_blocking=True)
            s["ev_out_free"].record(self.d2h)
        s["used"] = True
        return s["ev_out_free"]

    def synchronize(self) -> None:
        if self._slots is not None:
            self.h2d.synchronize()
            self.d2h.synchronize()
        torch.cuda.current_stream().synchronize()
