"""Run the level-2 correlation (and optionally other levels) a few times -- the target of ncu captures.
    python tools/prof_corr.py [--level 2] [--md 4] [--reps 3]"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=2)
ap.add_argument("--md", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
C = {6: 196, 5: 128, 4: 96, 3: 64, 2: 32}[a.level]
H, W = 448 >> a.level, 1024 >> a.level
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.nn.functional.leaky_relu(torch.randn(8, C, H, W, device="cuda", generator=g), 0.1)
f2 = torch.nn.functional.leaky_relu(torch.randn(8, C, H, W, device="cuda", generator=g), 0.1)
out = torch.empty(8, (2 * a.md + 1) ** 2, H, W, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(a.reps):
    flush.zero_()
    ops.correlation(f1, f2, pad_size=a.md, max_displacement=a.md, leaky_slope=0.1, out=out)
torch.cuda.synchronize()
print(_lib.last_kernel())
