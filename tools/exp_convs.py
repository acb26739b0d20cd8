"""Experiment: where does the 44 ms step go, and would channels_last help the cuDNN fp32 convolutions?"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import network, ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.benchmark = True
dev = "cuda"
torch.manual_seed(0)
model = network.MaskFlownetS().to(dev).eval()
a = torch.rand(8, 3, 448, 1024, device=dev) - 0.5; b = torch.rand(8, 3, 448, 1024, device=dev) - 0.5
def timed(fn, n=5):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
with torch.no_grad():
    print("forward NCHW fp32: %.2f ms" % timed(lambda: model(a, b)))
    # pyramid only
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(a, b); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
