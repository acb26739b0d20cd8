"""Per-call timing of every native op inside one forward of MaskFlownet-S (BASELINE configs[1]): CUDA events around each
ops.* call (serialised, warm L2) -> table sorted by time.  Shows where the step goes now that all 3x3 convolutions are ours."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import network, ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
model = network.MaskFlownetS().cuda().eval()
a = torch.randint(0, 255, (8, 3, 448, 1024), device="cuda", dtype=torch.uint8)
b = torch.randint(0, 255, (8, 3, 448, 1024), device="cuda", dtype=torch.uint8)
records = []
def wrap(name, keyfn):
    orig = getattr(ops, name)
    def f(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(*args, **kw); e1.record()
        records.append((name, keyfn(*args, **kw), e0, e1))
        return r
    setattr(ops, name, f)
wrap("conv3x3_slices", lambda bi, c0, Cin, p, bias, bo, o0, Cout, slope=0.1, dilation=1, stride=1, depth_to_space=False, linear_prefix=0:
     f"{Cin}->{Cout} {bi.shape[2]}x{bi.shape[3]} N{bi.shape[0]} s{stride} d{dilation}{' d2s' if depth_to_space else ''}{' lin' + str(linear_prefix) if linear_prefix else ''}")
wrap("correlation", lambda f1, f2, **kw: f"C{f1.shape[1]} {f1.shape[2]}x{f1.shape[3]}")
wrap("warp_mask", lambda x, *a, **kw: f"C{x.shape[1]} {x.shape[2]}x{x.shape[3]} {'resample' if kw.get('resample') else ('tc' if kw.get('packed_weight') is not None else 'simt')}")
with torch.no_grad():
    for _ in range(3):
        records.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); network.predict_flow(model, a, b); t1.record()
torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
rows = [(n, k, e0.elapsed_time(e1)) for n, k, e0, e1 in records]
native = sum(r[2] for r in rows)
print(f"step {tot:.3f} ms; native ops {native:.3f} ms in {len(rows)} calls; rest (torch glue) {tot - native:.3f} ms")
agg = collections.OrderedDict()
for n, k, t in rows:
    agg.setdefault((n, k), [0, 0.0]); agg[(n, k)][0] += 1; agg[(n, k)][1] += t
for (n, k), (cnt, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{t:8.3f} ms  x{cnt}  {n:16s} {k}")
