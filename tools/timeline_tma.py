"""Role timeline of CTA 0 of corr_tma_kernel (development): clock64 stamps of the producer (per quantum issued), converter warp 0
(before / after the ring-slot wait, after the quantum), MMA warp 0 of each group (unit start, data ready, unit end) and the two
store issuers (per store).   python tools/timeline_tma.py [--dbg N] [--tma 1|2]"""
import argparse, ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--dbg", type=int, default=0)
ap.add_argument("--tma", type=int, default=2)
a = ap.parse_args()
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
f1 = torch.randn(8, 32, 112, 256, device=dev, generator=g)
f2 = torch.randn(8, 32, 112, 256, device=dev, generator=g)
out = torch.empty(8, 81, 112, 256, device=dev)
ts = torch.zeros(6 * 512, dtype=torch.int64, device=dev)
_lib.set_tuning("corr_tma", a.tma)
_lib.set_tuning("corr_dbg", a.dbg)
for _ in range(2):
    ops.correlation(f1, f2, leaky_slope=0.1, algo=ops.CORR_MMA_BF16X3, out=out)
torch.cuda.synchronize()
p = ts.data_ptr()
lo, hi = p & 0xffffffff, p >> 32
_lib.set_tuning("corr_ts_lo", ctypes.c_int32(lo).value)
_lib.set_tuning("corr_ts_hi", ctypes.c_int32(hi).value)
ops.correlation(f1, f2, leaky_slope=0.1, algo=ops.CORR_MMA_BF16X3, out=out)
torch.cuda.synchronize()
_lib.set_tuning("corr_ts_lo", 0)
_lib.set_tuning("corr_ts_hi", 0)
_lib.set_tuning("corr_dbg", 0)
t = ts.cpu().view(6, 512)
t0 = int(t[t > 0].min())
names = ["producer", "converter0", "mma_g0", "mma_g1", "store_g0", "store_g1"]
print("kernel", _lib.last_kernel())
for r, nm in enumerate(names):
    v = [int(x) - t0 for x in t[r] if x > 0]
    print(nm, len(v), v[:60])
