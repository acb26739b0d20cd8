"""Sweep of the row-block correlation kernel (corr_rb.cu) over strip width / rows per CTA at the level-3..6 shapes, against
the tile kernel: cold-L2 timings + parity against the exact-fp32 SIMT kernel.

    python tools/dev_rb.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib  # noqa: E402

dev = "cuda"


def main():
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    md = 4
    for shape in [(8, 64, 56, 128), (8, 96, 28, 64), (8, 128, 14, 32), (8, 196, 7, 16)]:
        g = torch.Generator(device=dev).manual_seed(3)
        f1 = torch.nn.functional.leaky_relu(torch.randn(*shape, device=dev, generator=g), 0.1)
        f2 = torch.nn.functional.leaky_relu(torch.randn(*shape, device=dev, generator=g), 0.1)
        N, C, H, W = shape
        D = 81
        ref = ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=0.1, algo=ops.CORR_SIMT)
        out = torch.empty(N, D, H, W, device=dev)
        for rbk, twb, rows in [(1, 0, 0), (0, 0, 0), (2, 0, 4), (2, 0, 2), (2, 0, 1), (2, 2, 4), (2, 2, 2), (2, 2, 1)]:
            _lib.set_tuning("corr_rb", rbk)
            _lib.set_tuning("corr_rb_twb", twb)
            _lib.set_tuning("corr_rb_rows", rows)
            fn = lambda: ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=0.1,  # noqa: E731
                                         algo=ops.CORR_MMA_BF16X3, out=out)
            try:
                for _ in range(3):
                    fn()
            except Exception as e:  # noqa: BLE001
                print(json.dumps({"shape": shape, "rb": rbk, "twb": twb, "rows": rows, "error": str(e)[:100]}), flush=True)
                continue
            torch.cuda.synchronize()
            err = (out - ref).abs().max().item()
            ts = []
            for _ in range(20):
                flush.zero_()
                torch.cuda._sleep(600_000)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            avg = sum(ts) / len(ts)
            print(json.dumps({"shape": shape, "rb": rbk, "twb": twb, "rows": rows, "kernel": _lib.last_kernel(),
                              "us_avg": round(avg * 1e3, 2), "us_min": round(min(ts) * 1e3, 2), "max_err": err}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
