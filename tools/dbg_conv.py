import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kbench import timeit
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (N, Cin, Cout, H, W) in [(8, 579, 128, 112, 256), (8, 547, 32, 112, 256), (16, 16, 16, 224, 512)]:
    x = torch.randn(N, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05; b = torch.randn(Cout, device=dev)
    packed = ops.conv3x3_pack(w); out = torch.empty(N, Cout, H, W, device=dev)
    line = f"{Cin}->{Cout} {H}x{W}:"
    for dbg, name in [(0, "full"), (1, "no producer work"), (2, "no MMAs"), (3, "neither"), (4, "no stores"), (7, "barriers only")]:
        _lib.set_tuning("conv_dbg", dbg)
        avg, best = timeit(lambda: ops.conv3x3_slices(x, 0, Cin, packed, b, out, 0, Cout, 0.1), 8, flush)
        line += f"  [{name}] {avg*1e3:7.1f}"
    _lib.set_tuning("conv_dbg", 0)
    print(line, flush=True)
