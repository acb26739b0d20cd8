import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kbench import timeit
import sys as _s
if len(_s.argv) > 1: _lib.set_tuning("corr_ring_th", int(_s.argv[1]))
dev="cuda"; N,C,H,W=8,32,112,256
f1=torch.randn(N,C,H,W,device=dev); f2=torch.randn(N,C,H,W,device=dev); out=torch.empty(N,81,H,W,device=dev)
flush=torch.empty(256<<20,dtype=torch.uint8,device=dev)
for dbg,name in [(16,"launch only (immediate return)"),(0,"full (L2 prefetch + loads before last pass)"),(512,"full (L2 prefetch + loads at tile end)"),(256,"full, old scheme (register prefetch at tile start)"),(64,"full, no global loads (convert only)"),(128,"full, loads but no convert/store"),(12+64,"convert only, no mma/epi"),(12+128,"loads only, no mma/epi"),(2,"producers idle"),(4,"no epilogue"),(8,"no mma"),(12,"no mma, no epilogue (producers+barriers only)"),(14,"barrier protocol only"),(6,"mma only (no producers, no epilogue)"),(10,"epilogue only")]:
    _lib.set_tuning("corr_dbg", dbg)
    avg,best=timeit(lambda: ops.correlation(f1,f2,leaky_slope=0.1,out=out,algo=3), 20, flush)
    print(f"dbg={dbg:2d} {name:50s} {avg*1e3:8.1f} us (best {best*1e3:.1f})", flush=True)
_lib.set_tuning("corr_dbg", 0)
a=torch.empty(133038080//8, device=dev); b=torch.empty_like(a)
avg,best=timeit(lambda: b.copy_(a), 20, flush); print(f"torch copy 66.5MB->66.5MB (133 MB traffic): {avg*1e3:.1f} us (best {best*1e3:.1f})")
small=torch.empty(1024, device=dev)
avg,best=timeit(lambda: small.zero_(), 20, flush); print(f"tiny kernel: {avg*1e3:.1f} us")
for cap in (148, 74, 37):
    _lib.set_tuning("corr_grid_cap", cap)
    avg,best=timeit(lambda: ops.correlation(f1,f2,leaky_slope=0.1,out=out,algo=3), 20, flush)
    print(f"grid cap {cap}: {avg*1e3:.1f} us")
