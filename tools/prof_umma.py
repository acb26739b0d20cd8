import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops
dev = "cuda"
Cin, Cout = int(sys.argv[1]), int(sys.argv[2])
N, H, W = 8, 112, 256
x = torch.randn(N, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05; b = torch.randn(Cout, device=dev)
packed = ops.conv3x3_pack(w); out = torch.empty(N, Cout, H, W, device=dev)
for _ in range(3):
    ops.conv3x3_slices(x, 0, Cin, packed, b, out, 0, Cout, 0.1)
torch.cuda.synchronize()
