import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kbench import timeit
dev="cuda"; N,C,H,W=8,32,112,256
f1=torch.randn(N,C,H,W,device=dev); f2=torch.randn(N,C,H,W,device=dev); out=torch.empty(N,81,H,W,device=dev)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
_lib.set_tuning("corr_ring_th", 4)
for base,name in [(14,"barrier protocol only"),(14+32,"barrier protocol only, no prologue loads"),(0,"full")]:
    _lib.set_tuning("corr_dbg", base)
    avg,best=timeit(lambda: ops.correlation(f1,f2,leaky_slope=0.1,out=out,algo=3), 20, flush)
    print(f"dbg={base:3d} {name:45s} {avg*1e3:8.1f} us (best {best*1e3:.1f})", flush=True)
for skew_ns in (0, 200, 400, 600, 800, 1000, 1200, 1600, 2000, 3000):
    _lib.set_tuning("corr_dbg", (skew_ns // 32) << 16)
    avg,best=timeit(lambda: ops.correlation(f1,f2,leaky_slope=0.1,out=out,algo=3), 20, flush)
    print(f"skew {skew_ns:5d} ns: {avg*1e3:8.1f} us (best {best*1e3:.1f})", flush=True)
_lib.set_tuning("corr_dbg", 0)
