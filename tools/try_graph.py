import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import network
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
model = network.MaskFlownetS().cuda().eval()
a = torch.randint(0, 255, (8, 3, 448, 1024), device="cuda", dtype=torch.uint8)
b = torch.randint(0, 255, (8, 3, 448, 1024), device="cuda", dtype=torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pred = network.FlowPredictor(model)
with torch.no_grad():
    ref = network.predict_flow(model, a, b).clone()
    out = pred(a, b).clone()
    print("graph == eager:", float((ref - out).abs().max()))
    for name, fn in (("eager", lambda: network.predict_flow(model, a, b)), ("graph", lambda: pred(a, b))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            flush.zero_(); fn()
        e1.record(); torch.cuda.synchronize()
        print(name, e0.elapsed_time(e1) / 20, "ms/step")
