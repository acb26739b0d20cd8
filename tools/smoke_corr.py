"""Quick isolation check of the three correlation kernels against the oracle (each in its own try block)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
from oracle import cref
rng = np.random.default_rng(0)
res = {}
for shape in [(1, 32, 12, 32), (2, 64, 13, 20), (1, 196, 6, 8), (2, 16, 9, 15)]:
    f1 = rng.standard_normal(shape).astype(np.float32); f2 = rng.standard_normal(shape).astype(np.float32)
    for md in (4, 2):
        ref = cref.correlation_forward(f1, f2, pad_size=md, max_displacement=md)
        for name, algo in (("generic", 1), ("simt", 2), ("mma", 3)):
            try:
                out = ops.correlation(torch.from_numpy(f1).cuda(), torch.from_numpy(f2).cuda(), pad_size=md, max_displacement=md, algo=algo)
                torch.cuda.synchronize()
                err = float(np.abs(out.cpu().numpy() - ref).max())
            except Exception as e:
                err = "ERR " + str(e)[:200]
            res[f"{shape} md{md} {name}"] = err
            print(shape, md, name, err, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/smoke_corr.json", "w"), indent=1)
