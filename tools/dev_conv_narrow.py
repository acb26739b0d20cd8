"""Cold-L2 timings of the narrow (pyramid level 1-2) convolutions against their HBM floor.

    python tools/dev_conv_narrow.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib  # noqa: E402

dev = "cuda"
LAYERS = [(3, 16, 448, 1024, 16, 2), (16, 16, 224, 512, 16, 1), (16, 32, 224, 512, 16, 2), (32, 32, 112, 256, 16, 1),
          (32, 64, 112, 256, 16, 2), (64, 64, 56, 128, 16, 1), (64, 32, 112, 256, 8, 1), (128, 128, 112, 256, 8, 1)]


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for ci, co, h, w, n, st in LAYERS:
        x = torch.randn(n, ci, h, w, device=dev, generator=g)
        wt = torch.randn(co, ci, 3, 3, device=dev, generator=g) * (2.0 / (9 * ci)) ** 0.5
        b = torch.randn(co, device=dev, generator=g) * 0.1
        pk = ops.conv3x3_pack(wt)
        oh, ow = (h - 1) // st + 1, (w - 1) // st + 1
        y = torch.empty(n, co, oh, ow, device=dev)
        fn = lambda: ops.conv3x3_slices(x, 0, ci, pk, b, y, 0, co, 0.1, 1, st)  # noqa: E731
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ref = torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x[:2].double(), wt.double(), b.double(), stride=st, padding=1), 0.1)
        err = (y[:2].double() - ref).abs().max().item()
        ts = []
        for _ in range(10):
            flush.zero_()
            torch.cuda._sleep(600_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        avg = sum(ts) / len(ts)
        nbytes = 4 * n * (ci * h * w + co * oh * ow)
        flops = 2 * 9 * ci * co * n * oh * ow
        print(json.dumps({"layer": f"{ci}->{co} {h}x{w} N{n} s{st}", "kernel": _lib.last_kernel(), "us": round(avg * 1e3, 1),
                          "gbs": round(nbytes / avg / 1e6, 1), "frac_hbm": round(nbytes / avg / 1e6 / 6572.2, 3),
                          "tflops": round(flops / avg / 1e9, 1), "max_err": err}), flush=True)


if __name__ == "__main__":
    main()
