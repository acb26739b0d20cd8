import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import network
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = network.MaskFlownetS().cuda().eval()
a = torch.rand(8, 3, 448, 1024, device="cuda") - 0.5; b = torch.rand(8, 3, 448, 1024, device="cuda") - 0.5
with torch.no_grad():
    for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
        model(a, b)
torch.cuda.synchronize()
