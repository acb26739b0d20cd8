"""Micro-benchmark of the training-side kernels (rows N4 and N2 of SURVEY.md 8f) on one GPU: CUDA events on the launching
stream, 3 warm-ups, working sets larger than the 126 MB L2 where the shape allows.  One JSON line per kernel:
algorithmic bytes (inputs read once + outputs written once) / time against MEASURED_PEAKS.json's HBM figure.

  python tools/train_side_bench.py [--iters 20] > gpurun_out/train_side_bench.jsonl
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import augment, losses  # noqa: E402


def peak_gbps():
    try:
        d = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        pass
    return 6572.0, "fallback (round-2 measured copy bandwidth)"


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    peak, src = peak_gbps()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    out = []
    # ---- N4 geometry: Things3D-shaped batch (BASELINE configs[4]: 540x960 frames -> 384x768 crop... main.py uses target < orig)
    N, (H, W), (TH, TW) = 8, (540, 960), (448, 832)
    i1 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, device=dev, generator=g)
    i2 = torch.randint(0, 256, (N, 3, H, W), dtype=torch.uint8, device=dev, generator=g)
    fl = torch.randn(N, 2, H, W, device=dev, generator=g) * 5
    mk = torch.full((N, 1, 1, 1), 255, dtype=torch.uint8, device=dev)
    geo = augment.GeometryAugmentation(angle_range=(-17, 17), zoom_range=(0.5, 1 / 0.9), aspect_range=(0.9, 1 / 0.9),
                                       translation_range=0.1, target_shape=(TH, TW), orig_shape=(H, W), batch_size=N,
                                       relative_angle=0.25, relative_scale=(0.96, 1 / 0.96), relative_translation=0.25, seed=1)
    P = geo.params(geo.sample()).to(dev)
    ms = timed(lambda: augment.geometry_augment(i1, i2, fl, mk, P, (TH, TW)), a.iters)
    # source pixels actually touched ~ the target footprint: count the outputs (9 fp32 planes) + as many source samples
    by = N * TH * TW * (9 * 4 + 2 * 3 * 1 + 2 * 4)
    out.append({"kernel": "geometry_augment_kernel<u8>", "shape": [N, H, W, TH, TW], "ms": round(ms, 4), "algorithmic_bytes": by,
                "gbps": round(by / ms / 1e6, 1), "frac": round(by / ms / 1e6 / peak, 3)})
    a1, a2, _, _ = augment.geometry_augment(i1, i2, fl, mk, P, (TH, TW))
    # ---- N4 colour: in-kernel noise (two passes over both images: 2 reads + 1 write per value)
    col = augment.ColorAugmentation(contrast_range=(-0.4, 0.8), brightness_sigma=0.1, channel_range=(0.8, 1.4), batch_size=N,
                                    shape=(TH, TW), noise_range=(0, 0.04), saturation=0.5, hue=0.5, seed=2)
    d = col.sample()
    Pc = col.params(d).to(dev)
    for sigma, tag in ((0.0, "no noise"), (0.03, "in-kernel Philox noise")):
        ms = timed(lambda: augment.color_augment(a1, a2, Pc, noise_sigma=sigma, seed=7), a.iters)
        by = 2 * N * 3 * TH * TW * 4 * 3
        out.append({"kernel": "color_sum_kernel + color_apply_kernel", "variant": tag, "shape": [N, 3, TH, TW], "ms": round(ms, 4),
                    "algorithmic_bytes": by, "gbps": round(by / ms / 1e6, 1), "frac": round(by / ms / 1e6 / peak, 3)})
    # ---- N2 fused MultiscaleEpe, BASELINE configs[2] shape (batch 8, 384x512), forward + backward, against the composition
    N, H, W = 8, 384, 512
    preds = [torch.randn(N, 2, H // s, W // s, device=dev, generator=g).requires_grad_() for s in losses.SCALES]
    flow = torch.randn(N, 2, H, W, device=dev, generator=g) * 3
    mask = (torch.rand(N, 1, H, W, device=dev, generator=g) > 0.2).float()

    def run(fused):
        for p in preds:
            p.grad = None
        losses.multiscale_epe(flow, mask, preds, fused=fused).sum().backward()
    for fused in (True, False):
        ms = timed(lambda: run(fused), a.iters)
        by = N * H * W * 3 * 4 * 2      # label + mask, read once forward and (per scale, from L2) backward
        out.append({"kernel": "multiscale_epe fwd+bwd", "variant": "fused (3 launches)" if fused else "operator composition",
                    "shape": [N, H, W], "ms": round(ms, 4), "algorithmic_bytes": by, "gbps": round(by / ms / 1e6, 1)})
    for o in out:
        o["peak_gbps"], o["peak_source"] = peak, src
        print(json.dumps(o))


if __name__ == "__main__":
    main()
