"""Development check of the TMA correlation kernel (corr_tma.cu): parity against the exact-fp32 SIMT kernel over shapes,
displacements, grid caps (long per-CTA runs, strip changes) and concat-slot outputs; then cold-L2 timings.

    python tools/dev_tma.py [--no-time]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib  # noqa: E402

dev = "cuda"


def feats(shape, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    f1 = torch.nn.functional.leaky_relu(torch.randn(*shape, device=dev, generator=g), 0.1)
    f2 = torch.nn.functional.leaky_relu(torch.randn(*shape, device=dev, generator=g), 0.1)
    return f1, f2


def check(shape, md, cap=0, slope=0.1, extra=0):
    f1, f2 = feats(shape, 1)
    N, C, H, W = shape
    D = (2 * md + 1) ** 2
    _lib.set_tuning("corr_grid_cap", cap)
    ref = ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=slope, algo=ops.CORR_SIMT)
    buf = torch.full((N, D + extra, H, W), -7.0, device=dev)
    ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=slope, algo=ops.CORR_MMA_BF16X3, out=buf[:, :D])
    name = _lib.last_kernel()
    torch.cuda.synchronize()
    err = (buf[:, :D] - ref).abs().max().item()
    untouched = bool((buf[:, D:] == -7.0).all().item()) if extra else True
    _lib.set_tuning("corr_grid_cap", 0)
    ok = err <= 1e-4 and untouched and "tma" in name
    print(json.dumps({"shape": shape, "md": md, "cap": cap, "slope": slope, "kernel": name, "max_err": err,
                      "untouched": untouched, "ok": ok}), flush=True)
    return ok


def main():
    allok = True
    for shape in [(1, 32, 4, 32), (1, 32, 8, 32), (2, 32, 24, 40), (1, 32, 13, 64), (2, 16, 7, 16), (1, 24, 30, 100),
                  (3, 32, 112, 256), (8, 32, 112, 256), (2, 32, 96, 128)]:
        for md in (4, 2):
            allok &= check(shape, md)
    for cap in (1, 2, 3, 7, 40):
        allok &= check((2, 32, 24, 72), 4, cap=cap)
        allok &= check((2, 32, 21, 72), 2, cap=cap, extra=3)
    allok &= check((2, 32, 16, 64), 4, slope=1.0, extra=5)
    print("ALL OK" if allok else "FAILURES", flush=True)
    if "--no-time" in sys.argv or not allok:
        return 0 if allok else 1
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for shape, md in [((8, 32, 112, 256), 4), ((8, 32, 112, 256), 2), ((8, 32, 96, 128), 4)]:
        f1, f2 = feats(shape, 0)
        N, C, H, W = shape
        D = (2 * md + 1) ** 2
        out = torch.empty(N, D, H, W, device=dev)
        for tma, rbk in ((1, 1), (2, 1), (0, 0)):     # production kernel / its development build (switches compiled in) / round-1 ring kernel
            _lib.set_tuning("corr_tma", tma)
            _lib.set_tuning("corr_rb", rbk)
            fn = lambda: ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=0.1,  # noqa: E731
                                         algo=ops.CORR_MMA_BF16X3, out=out)
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
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
            nbytes = 4 * N * H * W * (2 * C + D)
            avg = sum(ts) / len(ts)
            print(json.dumps({"shape": shape, "md": md, "kernel": _lib.last_kernel(), "us_avg": round(avg * 1e3, 2),
                              "us_min": round(min(ts) * 1e3, 2), "gbs": round(nbytes / avg / 1e6, 1),
                              "frac_6572": round(nbytes / avg / 1e6 / 6572.2, 4)}), flush=True)
        _lib.set_tuning("corr_tma", 1)
        _lib.set_tuning("corr_rb", 1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
