#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: rotations/s for log_prob over a HEALPix SO(3) grid.

Workload (BASELINE.json configs[1]): SYMSOL-I MobiusAffine conditional flow (settings/symsol.yml with the 2048-d
feature override), log-prob + arg-max + normaliser over the level-5 HEALPix grid (2 359 296 rotations) for a batch of
B images per step; synthetic features, random-init weights (seed 0).  One "step" = one pass over (grid shard x B images)
per rank: per-image conditioner hoist, the fused flow kernel, the per-image reduction and (N > 1) the one all-gather.
With N ranks every rank scores its own randomly offset copy of the level-5 grid (the way BASELINE config 5's 37 M-rotation
grid is realised: 72*8^l has no 37 M member), so the global grid has N x 2 359 296 rotations: weak scaling.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's algorithm on the host cores (oracle port), same metric
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rotations/sec log_prob (forward + log-det, fused arg-max + normaliser) over HEALPix SO(3) grid"
UNIT = "rotations/s"
TENSOR_FLOPS_PER_ROT = 21 * 57344          # conditioner GEMM flops per rotation, symsol stack (21 Mobius layers)
FP32_FLOPS_PER_ROT = 21 * (384 + 5633) + 21 * 100   # first-layer y part + mixture/Jacobian + affine layers
ALL_FLOPS_PER_ROT = 1_332_681


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d.get("hbm_gbs", 6650.0), bf16_burst=d.get("bf16_tflops", 1590.0),
                    bf16_sustained=d.get("bf16_tflops_sustained", 1400.0), sm_max_mhz=d.get("sm_max_mhz", 1965.0), src="measured")
    return dict(hbm=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def build_flow(F=2048):
    import rotationnormflow_b200 as rnf
    cfg = rnf.load_config("symsol", feature_dim=F)
    torch.manual_seed(0)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        flow = rnf.get_flow(cfg)
    return cfg, flow


def cpu_leg(cfg, state_dict, sample_rows, threads, level=5, seed=123, repeats=1):
    """The reference algorithm (oracle port, explicit-Jacobian = op-for-op restatement) on the host cores."""
    from oracle import rnf_oracle as orc
    torch.set_num_threads(threads)
    o = orc.OracleFlow(cfg, state_dict, torch.float32, explicit_jacobian=True)
    G = 72 * 8 ** level
    gen = torch.Generator().manual_seed(seed)
    start = int(torch.randint(0, G - sample_rows, (1,), generator=gen))
    grid = orc.healpix_grid(level, start, start + sample_rows)
    off = orc.random_rotations(1, gen)[0]
    feat = torch.relu(torch.randn(1, orc.feature_dim_of(cfg), generator=gen))
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        samples = grid @ off
        rows = feat.repeat(sample_rows, 1)                       # eval.py:450 materialises the repeated feature
        _, ldj = o.forward(samples, rows)
        _ = torch.argmax(ldj)
        times.append(time.perf_counter() - t0)
    return sample_rows / min(times), times


def eager_gpu_leg(cfg, state_dict, dev, rows=500000, level=5, seed=321):
    """The practical incumbent: the reference's op sequence (oracle port, explicit Jacobian) in eager PyTorch on
    the B200, one 500 000-rotation chunk as at eval.py:445, features materialised per row as at eval.py:450.  Reported
    next to the CPU number; a baseline, not part of `value`."""
    from oracle import rnf_oracle as orc
    o = orc.OracleFlow(cfg, state_dict, torch.float32, explicit_jacobian=True, device=dev)
    G = 72 * 8 ** level
    gen = torch.Generator().manual_seed(seed)
    start = int(torch.randint(0, G - rows, (1,), generator=gen))
    grid = orc.healpix_grid(level, start, start + rows).to(dev)
    off = orc.random_rotations(1, gen)[0].to(dev)
    feat = torch.relu(torch.randn(1, orc.feature_dim_of(cfg), generator=gen)).to(dev)
    times = []
    for _ in range(3):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        samples = grid @ off
        rows_f = feat.repeat(rows, 1)
        _, ldj = o.forward(samples, rows_f)
        _ = torch.argmax(ldj).item()
        torch.cuda.synchronize(dev)
        times.append(time.perf_counter() - t0)
    return rows / min(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, flow = build_flow()
    threads = os.cpu_count() or 1
    sample = args.cpu_sample
    vals = []
    for i in range(args.warmup + args.steps):
        v, _ = cpu_leg(cfg, flow.state_dict(), sample, threads, seed=100 + i)
        if i >= args.warmup:
            vals.append(v)
    ms = 1000.0 * sample / statistics.mean(vals)
    value = statistics.mean(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "symsol.yml (F=2048) log_prob over HEALPix level-5 grid; CPU arm: bounded sample per step",
                   "grid_level": 5, "sample_rows_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} consecutive level-5 grid rotations x 1 image per step, oracle port of flow/*.py (torch CPU fp32, explicit Jacobian)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=int(os.environ.get("RNF_BENCH_IMAGES", "8")))
    ap.add_argument("--level", type=int, default=5)
    ap.add_argument("--mode", default=os.environ.get("RNF_BENCH_MODE", ""))
    ap.add_argument("--cpu-sample", type=int, default=160000)      # ~10 s of host work per pass at ~16 k rot/s on 16 cores
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import rotationnormflow_b200 as rnf  # noqa: F401
    from oracle import rnf_oracle as orc  # cpu_baseline leg + synthetic input helpers only
    from rotationnormflow_b200 import dist as rdist
    from rotationnormflow_b200 import engine, grid as rgrid
    from rotationnormflow_b200.flow import _program

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode or ("tc" if engine.TC_AVAILABLE else "fp32")

    cfg, flow = build_flow()
    flow = flow.to(dev).eval()
    B, level = args.images, args.level
    G = 72 * 8 ** level
    grid = rgrid.healpix_grid(level, device=dev)                 # resident, as the reference caches it (utils/sd.py:28)
    gen = torch.Generator().manual_seed(1234 + rank)
    offset = orc.random_rotations(1, gen)[0].to(dev)             # eval.py:439: one random right-offset per batch
    feat_host = torch.relu(torch.randn(B, 2048, generator=torch.Generator().manual_seed(77))).pin_memory()
    feat_dev = feat_host.to(dev)
    prog = _program(flow, list(flow.layers), flow._perm_rows(), flow.feature_dim, dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    kern_ms = []

    def step_device(record=False):
        cond = prog.condition(feat_dev)
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        mx, am, se, _ = prog.grid_logprob(grid, rank * G, offset, cond, B, None, None, False, mode)
        if record:
            e1.record()
            kern_ms.append((e0, e1))
        return rdist.all_merge(mx, am, se)

    def step_e2e():
        f = feat_host.to(dev, non_blocking=True)
        out = rdist.sharded_grid_log_prob(flow, grid, rank * G, world * G, f, offset=offset, mlp_mode=mode)
        return out["max"].cpu(), out["argmax"].cpu(), out["log_norm"].cpu()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    evs = []
    with ClockSampler(local) as clk:
        barrier()
        t_wall0 = time.perf_counter()
        for _ in range(args.steps):
            flush.fill_(1)                                       # L2 flush between timed iterations (untimed)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step_device(record=True)
            e.record()
            evs.append((s, e))
        barrier()
        t_wall = time.perf_counter() - t_wall0
    step_ms = [s.elapsed_time(e) for s, e in evs]
    k_ms = [a.elapsed_time(b) for a, b in kern_ms]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = world * G * B / (ms_per_step * 1e-3)

    # end-to-end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step_e2e()
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * G * B / (float(e2e_t.item()) / args.steps)

    # secondary figure of the metric: sampling = Flow.inverse (BASELINE configs[3], SYMSOL-II stack, F=512), scaled to
    # 64 images x 32768 base samples per rank so the default run stays short; device-timed, not part of `value`.
    import rotationnormflow_b200 as rnf2
    torch.manual_seed(0); np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        flow_s = rnf2.get_flow(rnf2.load_config("symsol2")).to(dev).eval()
    n_img_s, n_per = 64, 32768
    base = rgrid.generate_queries(n_img_s * n_per, "random", device=dev)
    feat_s = torch.relu(torch.randn(n_img_s, 512, generator=torch.Generator().manual_seed(5))).to(dev)
    idx_s = torch.arange(n_img_s * n_per, device=dev, dtype=torch.int32) // n_per
    samp_ms = []
    with torch.no_grad():
        for i in range(2 + 3):
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            Rs, ls = flow_s.inverse(base, feat_s, feature_index=idx_s, mlp_mode=mode)
            s1.record()
            torch.cuda.synchronize()
            if i >= 2:
                samp_ms.append(s0.elapsed_time(s1))
    samp_t = torch.tensor([statistics.mean(samp_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(samp_t, op=dist.ReduceOp.MAX)
    sampling = {"value": world * n_img_s * n_per / (float(samp_t.item()) * 1e-3), "unit": "samples/s", "ms": float(samp_t.item()),
                "config": f"symsol2.yml (F=512, 42 layers) Flow.inverse, {n_img_s} images x {n_per} samples per rank (BASELINE configs[3] scaled down)"}

    pk = peaks()
    kt = statistics.mean(k_ms) * 1e-3
    rot_per_launch = G * B
    ach_tensor = TENSOR_FLOPS_PER_ROT * rot_per_launch / kt / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(mode)
    clocks = clk.summary()
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if mode == "fp32" else "f32 (conditioner GEMMs: split-fp16 tensor-core operands, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: symsol.yml (F=2048, 42 layers) log_prob + argmax + normaliser over HEALPix level-%d grid" % level,
                   "grid_rotations_per_rank": G, "images_per_step": B, "global_grid": world * G, "mlp_mode": mode,
                   "l2": "flushed between timed steps (256 MiB write)", "weights": "random init seed 0",
                   "parallelism": f"grid-sharded x{world}, one all-gather of [B,3] f64"},
        "roofline": {"bound": "tensor", "achieved": ach_tensor, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": ach_tensor / pk["bf16_sustained"], "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained",
                     # the error-compensated split issues 3 fp16 products + one K=16 block MMA per K=64 GEMM: 3.25x the useful FLOPs
                     "issued": None if mode == "fp32" else {"achieved": 3.25 * ach_tensor, "frac": 3.25 * ach_tensor / pk["bf16_sustained"],
                                                            "note": "fp16 MMAs actually issued (3-product split + bias/fc_first block)"},
                     "kernel": {"tc": "flow_t4_kernel", "tc_row": "flow_row_kernel"}.get(mode, "flow_v1_kernel"), "kernel_ms": kt * 1e3,
                     "algorithmic_flops_per_rotation": {"tensor_eligible": TENSOR_FLOPS_PER_ROT, "fp32_pipe": FP32_FLOPS_PER_ROT,
                                                        "all": ALL_FLOPS_PER_ROT},
                     "fp32_pipe": {"achieved": (ALL_FLOPS_PER_ROT if mode == "fp32" else FP32_FLOPS_PER_ROT) * rot_per_launch / kt / 1e12,
                                   "peak_nominal": fp32_peak, "unit": "TFLOP/s"}},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(feat_host.numel() * 4),
                "d2h_bytes_per_step": int(B * (4 + 8 + 4))},
        "gpu_launches": int(args.steps * (2 + 2)),
        "sampling": sampling,
        "wall_s_timed_region": t_wall,
        "check": {"argmax0": int(res[1][0]), "log_norm0": float(res[2][0])},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            del flush
            torch.cuda.empty_cache()
            v_eager = eager_gpu_leg(cfg, {k: t.detach() for k, t in flow.state_dict().items()}, dev)
            line["eager_torch_gpu_baseline"] = {"value": v_eager, "unit": UNIT, "kind": "port",
                                                "sample": "one 500 000-rotation chunk x 1 image (eval.py:445), oracle port of flow/*.py as eager PyTorch ops on the same B200, best of 3"}
        except Exception as e:  # a baseline only: never fail the bench line over it
            line["eager_torch_gpu_baseline"] = {"value": None, "error": repr(e)[:200]}
        threads = os.cpu_count() or 1
        v, times = cpu_leg(cfg, {k: t.cpu() for k, t in flow.state_dict().items()}, args.cpu_sample, threads, repeats=2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{args.cpu_sample} consecutive level-5 grid rotations x 1 image (oracle port of flow/*.py, torch CPU fp32), best of 2"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
