#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config: image-pairs/sec of the MaskFlownet-S forward at
1024x448, batch 8 per GPU, fp32, synthetic images, random-init (MSRAPrelu) weights; plus the correlation kernel's
achieved HBM GB/s against the measured B200 roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (one rank per GPU, replicas)

One "step" = one pass of the hot path over one batch: uint8 pairs (already in HBM for `value`) -> /255 -> centralize ->
6-level MaskFlownet-S forward (5 correlation kernels + 4 fused warp kernels of this repo, dense convs on cuDNN fp32) ->
Upsample(4) of the finest flow.  `e2e` runs the same step through the public API with HOST buffers: pinned uint8 images
H2D, forward, full-resolution flow D2H, all inside the timed region.

--impl reference times the CPU arm: the reference's own CPU path cannot run here (MXNet is not installable, SURVEY.md
section 8c), so it is the oracle port (oracle/network_ref.py: torch-CPU convolutions + the C oracle's OpenMP correlation /
deformable convolution) on all host threads, one image pair per step (a bounded sample of the same workload).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, BATCH = 448, 1024, 8
METRIC = "image-pairs/sec (MaskFlownet-S forward, 1024x448)"
LEVEL_C = {6: 196, 5: 128, 4: 96, 3: 64, 2: 32}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dram_traffic(kernel_name: str):
    """dram read+write bytes per launch of the dominant kernel, from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    try:
        for key, rec in json.load(open(p)).items():
            if not key.startswith("_") and key in (kernel_name or ""):
                return int(rec["dram_read_bytes"]) + int(rec["dram_write_bytes"])
    except (OSError, ValueError, KeyError):
        pass
    return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def synthetic_pairs(n, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randint(0, 256, (n, 3, H, W), dtype=torch.uint8, generator=g)
    b = torch.randint(0, 256, (n, 3, H, W), dtype=torch.uint8, generator=g)
    return a, b


def cpu_arm(steps: int, warmup: int, max_threads: int):
    """Oracle port of the same forward on the host cores; one 1024x448 pair per step.  The thread count is the fastest
    of a short probe over {16, 32, 64, all} (more threads than that only add contention on the small pyramid levels)."""
    from oracle import cref, network_ref
    from maskflownet_b200.network import MaskFlownetS
    model = MaskFlownetS()
    params = {k: v.detach() for k, v in model.named_parameters()}
    a, b = synthetic_pairs(1, 0)

    def run(n, threads):
        torch.set_num_threads(threads)
        cref.lib().mfn_ref_set_num_threads(threads)
        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(n):
                network_ref.predict_flow(params, a, b, threads=threads)
        return (time.perf_counter() - t0) / n

    cands = sorted({t for t in (16, 32, 64, max_threads) if t <= max_threads})
    probe = {t: run(1, t) for t in cands}
    best = min(probe, key=probe.get)
    for _ in range(max(0, warmup - 1)):
        run(1, best)
    sec = run(steps, best)
    return 1.0 / sec, sec, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample-steps", type=int, default=2)
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    host_threads = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return
        val, sec, used = cpu_arm(max(1, K), max(1, min(Wm, 1)), host_threads)
        line = {"impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "pairs/s", "n_gpus": args.gpus,
                "steps": K, "warmup": Wm, "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "MaskFlownet-S full 6-level forward, 1024x448 synthetic, random-init weights "
                                       "(BASELINE configs[1]); CPU sample: 1 pair per step"},
                "cpu_baseline": {"value": round(val, 4), "unit": "pairs/s", "cores": used, "kind": "port",
                                 "sample": f"{max(1, K)} steps x 1 pair at 1024x448 (oracle/network_ref.py: torch-CPU "
                                           "convs + C-oracle OpenMP correlation/deformable conv)"},
                "e2e": {"value": round(val, 4), "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # ------------------------------------------------------------------------------------------ our arm (GPU)
    from maskflownet_b200 import _lib, dist as mdist, network, ops
    rank, local, world = mdist.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU implementation (use --impl reference for "
                         "the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cudnn.allow_tf32 = False           # fp32 like the reference; no reduced-precision convolutions
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True

    torch.manual_seed(0)
    model = network.MaskFlownetS().to(dev).eval()
    a_h, b_h = synthetic_pairs(BATCH, 100 + rank)
    a_h, b_h = a_h.pin_memory(), b_h.pin_memory()
    a_d, b_d = a_h.to(dev), b_h.to(dev)
    out_h = torch.empty((BATCH, 2, H, W), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak, peak_kind = hbm_peak()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def step_resident():
        return network.predict_flow(model, a_d, b_d)

    def step_e2e():
        x1 = a_h.to(dev, non_blocking=True)    # H2D from pinned host memory
        x2 = b_h.to(dev, non_blocking=True)
        flow = network.predict_flow(model, x1, x2)
        out_h.copy_(flow, non_blocking=True)   # D2H of the result
        return flow

    # kernel events inside the timed region (dominant kernel: level-2 correlation)
    ev = {}

    def hook(kind, lvl, phase):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.setdefault((kind, lvl), []).append(e)

    with torch.no_grad():
        for _ in range(max(Wm, 3)):
            step_resident()
            step_e2e()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        # ---- value: device-resident inputs ----
        model.event_hook = hook
        n0 = _lib.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()   # cudaProfilerStart: lets `ncu --profile-from-start off` list exactly the timed launches
        e0.record()
        for _ in range(K):
            flush.zero_()             # L2 flush between steps (inside the timed region, ~0.05 ms per step)
            step_resident()
        e1.record()
        barrier()
        torch.cuda.profiler.stop()
        ms_total = mdist.max_over_ranks(e0.elapsed_time(e1), dev)
        launches = _lib.launch_count() - n0
        model.event_hook = None
        kt = {}
        for key, lst in ev.items():
            durs = [lst[i].elapsed_time(lst[i + 1]) for i in range(0, len(lst) - 1, 2)]
            kt[key] = sum(durs) / len(durs)
        # ---- e2e: host buffers, copies inside the timed region ----
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(K):
            flush.zero_()
            step_e2e()
        f1.record()
        barrier()
        ms_e2e = mdist.max_over_ranks(f0.elapsed_time(f1), dev)
        sampler.stop_flag = True
        sampler.join(timeout=2)
        # ---- the dominant kernel alone, cold L2 (diagnostic, not the judged figure) ----
        f1t = torch.randn(BATCH, 32, H // 4, W // 4, device=dev)
        f2t = torch.randn(BATCH, 32, H // 4, W // 4, device=dev)
        outb = torch.empty(BATCH, 81, H // 4, W // 4, device=dev)
        iso = []
        for _ in range(10):
            flush.zero_()
            torch.cuda._sleep(400_000)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            ops.correlation(f1t, f2t, leaky_slope=0.1, out=outb)
            g1.record()
            g1.synchronize()
            iso.append(g0.elapsed_time(g1))
        corr_kernel = _lib.last_kernel()

    pairs = BATCH * K * world
    value = pairs / (ms_total * 1e-3)
    e2e_value = pairs / (ms_e2e * 1e-3)
    # roofline of the dominant hot-path kernel: level-2 correlation, algorithmic bytes 4*N*H*W*(2C+81)
    n2, h2, w2, c2 = BATCH, H // 4, W // 4, 32
    alg_bytes = 4 * n2 * h2 * w2 * (2 * c2 + 81)
    t_corr2 = kt.get(("corr", 2))
    achieved = alg_bytes / (t_corr2 * 1e-3) / 1e9 if t_corr2 else None
    hot_ms = sum(v for v in kt.values())
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
        "ms_per_step": round(ms_total / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MaskFlownet-S full 6-level forward, batch 8 per GPU, 1024x448 synthetic uint8 pairs, "
                               "random-init weights, fp32 (cuDNN TF32 off) -- BASELINE configs[1]",
                   "global_batch": BATCH * world, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "256 MiB buffer overwritten between steps (inside the timed region, ~0.05 ms/step)",
                   "value_path": "network.predict_flow, eager launches, device-resident inputs, CUDA events around the hot-path "
                                 "kernels inside the timed region",
                   "e2e_path": "network.predict_flow (eager launches), pinned host uint8 in, pinned host fp32 flow out, "
                               "copies inside the timed region"},
        "e2e": {"value": round(e2e_value, 3), "unit": "pairs/s", "h2d_bytes_per_step": int(a_h.numel() + b_h.numel()),
                "d2h_bytes_per_step": int(out_h.numel() * 4), "ms_per_step": round(ms_e2e / K, 4)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
        "roofline": {"kernel": f"{corr_kernel} (level-2 correlation, N=8 C=32 112x256, md=4)", "bound": "hbm",
                     "achieved": round(achieved, 1) if achieved else None, "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4) if achieved else None, "traffic": dram_traffic(corr_kernel),
                     "peak_source": peak_kind, "alg_bytes_per_launch": alg_bytes,
                     "launch_ms_in_step": round(t_corr2, 5) if t_corr2 else None,
                     "launch_ms_isolated_cold_l2": round(sum(iso) / len(iso), 5),
                     "hot_path_ms_per_step": round(hot_ms, 4),
                     "hot_path_share_of_step": round(hot_ms / (ms_total / K), 4),
                     "per_kernel_ms": {f"{k[0]}{k[1]}": round(v, 5) for k, v in sorted(kt.items())}},
    }
    if rank == 0 and world == 1 and args.cpu_sample_steps > 0:
        try:
            val, sec, used = cpu_arm(args.cpu_sample_steps, 1, host_threads)
            line["cpu_baseline"] = {"value": round(val, 4), "unit": "pairs/s", "cores": used, "kind": "port",
                                    "sample": f"{args.cpu_sample_steps} steps x 1 pair at 1024x448 on {used} of "
                                              f"{host_threads} host threads (oracle/network_ref.py)"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "pairs/s", "cores": host_threads, "kind": "port",
                                    "sample": f"failed: {e}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
