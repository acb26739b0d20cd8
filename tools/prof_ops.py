"""Run ONE hot-path operator a few times at its BASELINE configs[1] shape -- the target of `ncu -k regex:<kernel>` captures.
    python tools/prof_ops.py corr --level 2 | warp --level 3 | k5 | corr_bwd --level 2 | warp_bwd --level 3 | cascade_corr --level 2"""
import argparse, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("what")
ap.add_argument("--level", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--conv", default="16,16,224,512,16,1", help="conv: Cin,Cout,H,W,N,stride (input size)")
a = ap.parse_args()
L = a.level
C = {6: 196, 5: 128, 4: 96, 3: 64, 2: 32}[L]
N, H, W = 8, 448 >> L, 1024 >> L
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)  # noqa: E731
f1 = torch.nn.functional.leaky_relu(rn(N, C, H, W), 0.1)
f2 = torch.nn.functional.leaky_relu(rn(N, C, H, W), 0.1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if a.what in ("corr", "cascade_corr"):
    md = 4 if a.what == "corr" else 2
    out = torch.empty(N, (2 * md + 1) ** 2, H, W, device=dev)
    fn = lambda: ops.correlation(f1, f2, pad_size=md, max_displacement=md, leaky_slope=0.1, out=out)  # noqa: E731
elif a.what == "warp":
    w = rn(C, C, 3, 3) * (2.0 / (9 * C)) ** 0.5
    b = torch.zeros(C, device=dev)
    fc = rn(N, 2, H // 2, W // 2) * 0.4 * (2 ** L) / 20.0 / 4
    mc = rn(N, 1, H // 2, W // 2) + 0.5
    tr = rn(N, C, H, W) * 0.3
    pk = ops.conv3x3_pack(w)
    fn = lambda: ops.warp_mask(f2, fc, mc, w, b, tr, 20.0, float(2 ** L), 2, 0.1, 0, packed_weight=pk, resample=True)  # noqa: E731
elif a.what == "k5":
    im1, im2 = rn(N, 3, 448, 1024), rn(N, 3, 448, 1024)
    fq, mq = rn(N, 2, 112, 256) * 0.2, rn(N, 1, 112, 256)
    fn = lambda: ops.image_warp_concat(im1, im2, fq, mq, 20.0)  # noqa: E731
elif a.what == "corr_bwd":
    t1, t2 = f1.clone().requires_grad_(), f2.clone().requires_grad_()
    o = ops.correlation(t1, t2, leaky_slope=0.1, algo=ops.CORR_SIMT)
    go = rn(*o.shape)

    def fn():
        t1.grad = t2.grad = None
        o.backward(go, retain_graph=True)
elif a.what == "warp_bwd":
    w = (rn(C, C, 3, 3) * (2.0 / (9 * C)) ** 0.5).requires_grad_()
    b = torch.zeros(C, device=dev, requires_grad=True)
    x = f2.clone().requires_grad_()
    fc = (rn(N, 2, H // 2, W // 2) * 0.05).requires_grad_()
    mc = (rn(N, 1, H // 2, W // 2) + 0.5).requires_grad_()
    tr = (rn(N, C, H, W) * 0.3).requires_grad_()
    o, _, _ = ops.warp_mask(x, fc, mc, w, b, tr, 20.0, float(2 ** L), 2, 0.1, 0)
    go = rn(*o.shape)

    def fn():
        for t in (w, b, x, fc, mc, tr):
            t.grad = None
        o.backward(go, retain_graph=True)
elif a.what == "conv":
    ci, co, h, w_, n, st = (int(v) for v in a.conv.split(","))
    xin = rn(n, ci, h, w_)
    wt = rn(co, ci, 3, 3) * (2.0 / (9 * ci)) ** 0.5
    pk = ops.conv3x3_pack(wt)
    bo = torch.zeros(co, device=dev)
    yout = torch.empty(n, co, (h - 1) // st + 1, (w_ - 1) // st + 1, device=dev)
    fn = lambda: ops.conv3x3_slices(xin, 0, ci, pk, bo, yout, 0, co, 0.1, 1, st)  # noqa: E731
else:
    raise SystemExit("unknown op " + a.what)
with torch.set_grad_enabled(a.what.endswith("bwd")):
    for _ in range(a.reps):
        flush.zero_()
        fn()
torch.cuda.synchronize()
print(a.what, L, _lib.last_kernel())
