"""First-light check of the tcgen05 convolution: correctness vs float64 conv on a few shapes, then timing vs the mma.sync kernel."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskflownet_b200 import ops, _lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kbench import timeit
dev = "cuda"
def check(N, Cin, Cout, H, W, dil=1):
    rng = np.random.default_rng(5)
    x = rng.standard_normal((N, Cin, H, W)).astype(np.float32)
    w = (rng.standard_normal((Cout, Cin, 3, 3)) * np.sqrt(2.0 / (9 * Cin))).astype(np.float32)
    b = (rng.standard_normal(Cout) * 0.1).astype(np.float32)
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(),
                                     padding=dil, dilation=dil)
    ref = torch.nn.functional.leaky_relu(ref, 0.1).float().numpy()
    packed = ops.conv3x3_pack(torch.from_numpy(w).to(dev))
    res = {}
    for mode in (1, 0):
        _lib.set_tuning("conv_umma", mode); _lib.set_tuning("conv_umma_min_w", 1)
        got = ops.conv3x3(torch.from_numpy(x).to(dev), packed, torch.from_numpy(b).to(dev), Cout, 0.1, dilation=dil)
        torch.cuda.synchronize()
        res[mode] = (float(np.abs(got.cpu().numpy() - ref).max()), _lib.last_kernel())
    print(f"N={N} Cin={Cin} Cout={Cout} H={H} W={W} dil={dil}: umma err {res[1][0]:.2e} [{res[1][1]}]  sync err {res[0][0]:.2e} [{res[0][1]}]", flush=True)
for shp in [] if os.environ.get("SKIP_CHECK") else [(1, 16, 32, 4, 128), (1, 16, 128, 2, 128), (1, 32, 64, 5, 130), (2, 81, 128, 7, 16), (2, 131, 128, 12, 40), (1, 35, 32, 9, 33),
            (1, 579, 128, 10, 24), (1, 64, 96, 9, 256), (1, 40, 96, 21, 45, 2), (1, 40, 96, 21, 45, 4), (1, 40, 64, 21, 45, 16), (1, 128, 128, 30, 200, 8)]:
    check(*shp)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
_lib.set_tuning("conv_umma", 1); _lib.set_tuning("conv_umma_min_w", 1)
for (N, Cin, Cout, H, W, stride) in [(8, 131, 128, 112, 256, 1), (8, 565, 32, 112, 256, 1), (8, 259, 128, 56, 128, 1), (8, 341, 128, 28, 64, 1),
                                     (8, 128, 128, 112, 256, 1), (8, 597, 64, 56, 128, 1), (16, 3, 16, 448, 1024, 2), (16, 16, 16, 224, 512, 1), (16, 16, 32, 224, 512, 2), (16, 32, 32, 112, 256, 1),
                                     (16, 32, 64, 112, 256, 2), (16, 64, 64, 56, 128, 1), (16, 64, 96, 56, 128, 2), (16, 96, 96, 28, 64, 1),
                                     (16, 96, 128, 28, 64, 2), (16, 128, 128, 14, 32, 1), (16, 128, 196, 14, 32, 2), (16, 196, 196, 7, 16, 1),
                                     (8, 597, 2, 112, 256, 1), (8, 16, 32, 112, 256, 1)]:
    x = torch.randn(N, Cin, H, W, device=dev); w = torch.randn(Cout, Cin, 3, 3, device=dev) * 0.05; b = torch.randn(Cout, device=dev)
    packed = ops.conv3x3_pack(w); OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
    out = torch.empty(N, Cout, OH, OW, device=dev)
    avg, best = timeit(lambda: ops.conv3x3_slices(x, 0, Cin, packed, b, out, 0, Cout, 0.1, 1, stride), 10, flush)
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cudnn.benchmark = True
    avg2, best2 = timeit(lambda: torch.nn.functional.leaky_relu(torch.nn.functional.conv2d(x, w, b, stride=stride, padding=1), 0.1), 10, flush)
    mb = 4 * (x.numel() + out.numel()) / 1e6
    print(f"N={N} {Cin}->{Cout} {H}x{W} s{stride}: umma {avg*1e3:8.1f} us ({mb/avg/1e3:6.1f} GB/s alg)   cuDNN+leaky {avg2*1e3:8.1f} us", flush=True)
