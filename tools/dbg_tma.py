"""Ablation of the TMA correlation kernel's pipeline stages with the corr_dbg switches (results invalid unless 0):
1 no TMA stores, 2 no staging stores, 4 no MMAs, 8 no conversion, 16 no TMA loads, 32 no data2 ldmatrix, 64 no L2 prefetch,
128 serialised store issue.   python tools/dbg_tma.py"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.nn.functional.leaky_relu(torch.randn(8, 32, 112, 256, device=dev, generator=g), 0.1)
f2 = torch.nn.functional.leaky_relu(torch.randn(8, 32, 112, 256, device=dev, generator=g), 0.1)
out = torch.empty(8, 81, 112, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def t(dbg, tma):
    _lib.set_tuning("corr_tma", tma)
    _lib.set_tuning("corr_dbg", dbg)
    fn = lambda: ops.correlation(f1, f2, leaky_slope=0.1, algo=ops.CORR_MMA_BF16X3, out=out)  # noqa: E731
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        torch.cuda._sleep(500_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    _lib.set_tuning("corr_dbg", 0)
    return round(sum(ts) / len(ts) * 1e3, 1)


names = {0: "full", 64: "no L2 prefetch", 128: "serial stores", 192: "no pf + serial (=first version)", 1: "no TMA stores",
         3: "no staging+stores", 4: "no MMA", 7: "no MMA/epilogue/stores", 8: "no conversion", 16: "no TMA loads",
         24: "no loads+conversion", 32: "no data2 ldmatrix", 36: "no MMA + no ldmatrix", 39: "loads+conversion only",
         31: "barrier protocol only (+ldmatrix)", 63: "barrier protocol only"}
names = {0: "full", 64: "no L2 prefetch", 128: "serial stores", 1: "no TMA stores", 2: "no staging stores", 3: "no staging+stores",
         4: "no MMA", 7: "no MMA/staging/stores", 8: "no conversion", 16: "no TMA loads", 24: "no loads+conversion",
         32: "no data2 ldmatrix", 39: "loads+conversion only", 63: "barrier protocol only", 512: "no proxy fence"}
for tma, label in ((2, "unrolled"),):
    for dbg, nm in names.items():
        print(json.dumps({"variant": label, "dbg": dbg, "what": nm, "us": t(dbg, tma)}), flush=True)
_lib.set_tuning("corr_tma", 1)
