#!/usr/bin/env python
"""Benchmark of the hot path on BASELINE.json's metric: rotations/s for log_prob (forward + log-det, fused arg-max +
normaliser) over a HEALPix SO(3) grid, and for sampling (Flow.inverse), on 1/2/4/8 B200 with % of roofline.

    python bench.py --gpus 1 --steps 5 --warmup 3            # BASELINE configs[1] (the headline at N = 1)
    python bench.py --config {1..5} ...                     # any BASELINE.json configuration at its stated size
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # default at N > 1: configs[4], strong scaling
    python bench.py --impl reference ...                    # the reference's own CPU code path (oracle/_ref), same metric

BASELINE.json configurations (synthetic features, random-init weights seed 0):
  1  raw.yml, unconditional, Flow.forward on 100 000 uniform rotations (rows in, rows out); replicas only at N > 1
  2  symsol.yml F=2048, log-prob + arg-max + normaliser over the level-5 grid (2 359 296) x 8 images per step; at N > 1 every
     rank scores its own offset copy of the grid (weak scaling)
  3  modelnet_fisher.yml F=2080, matrix-Fisher base, grid 4 718 592 = level-5 grid under two offsets, x 256 images per step;
     the global grid index range is split over the ranks (strong scaling)
  4  symsol2.yml F=512, Flow.inverse of 1 000 000 base rotations per image x 64 images per step; images split over the ranks
  5  symsol.yml F=2048, grid 37 748 736 = level-6 grid under two offsets, split over the ranks by dist.shard_range, x 1024
     images streamed in steps of --images (default 8) images: one step = the WHOLE grid x 8 images; one all-gather of [B,3] per
     step (strong scaling: the work of a step is fixed, each rank scores 1/N of the grid)
The default run at N = 1 times config 2 as the headline and adds one short pass of configs 1, 3, 4 and of a config-5 step under
"configs", so that every BASELINE configuration is in the driver's record at its stated size.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import math
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
METRIC_ROWS = "rotations/sec log_prob (Flow.forward: rotation + log-det per row)"
METRIC_SAMPLING = "rotations/sec sampling (Flow.inverse: bisection root-solve, rotation + log-det per sample)"
UNIT = "rotations/s"
# SURVEY.md 8(d), algorithmic work per rotation and Mobius layer (FMA = 2): conditioner GEMMs 57 344 tensor-eligible FLOPs;
# FP32 pipe 384 (first-layer y part) + 5 633 (mixture, frame, Jacobian) forward, 58 040 inverse (15 bisection probes);
# ~100 per quaternion affine layer.  SFU: ~400 forward, ~2 280 inverse.
TENSOR_FLOPS_PER_MOBIUS = 57344
FP32_FWD_PER_MOBIUS = 384 + 5633
FP32_INV_PER_MOBIUS = 115384 - 57344             # the reference's algorithm: 15 bisection probes


def fp32_inv_per_mobius(evals):
    """SURVEY.md 8(d) with E evaluations of the mixture map instead of 15: first layer 384 + preparation 64*27 + E*(64*57+25) +
    Jacobian 64*12 + 65.  The kernel locates the root with Newton steps and replays the reference's halvings (csrc/flow_row.cu), so it
    EXECUTES fewer evaluations than the reference's 15; the roofline is quoted on the work executed, measured by a device counter."""
    return 384 + 64 * 27 + evals * (64 * 57 + 25) + 64 * 12 + 65


AFFINE_FLOPS = 100
SPLIT_ISSUE_FACTOR = 3.25        # error-compensated split: 3 fp16 products + one K=16 block MMA per K=64 GEMM


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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_median": statistics.median(pw)}


def build_flow(name="symsol", **ov):
    import rotationnormflow_b200 as rnf
    cfg = rnf.load_config(name, **ov)
    torch.manual_seed(0)
    np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        flow = rnf.get_flow(cfg)
    return cfg, flow


# ------------------------------------------------------------------------------------------------------------------
# CPU legs (baselines only; the one place besides tests/ and smoke() that executes oracle/)
# ------------------------------------------------------------------------------------------------------------------
def _reference_root():
    p = os.path.join(ROOT, "oracle", "_ref")
    return p if os.path.isdir(os.path.join(p, "flow")) else None


def cpu_leg(cfg, state_dict, sample_rows, threads, level=5, seed=123, repeats=1, prefer_reference=True):
    """One image x `sample_rows` consecutive grid rotations through the reference algorithm on the host cores, the way
    eval.py:444-462 runs it (feature.repeat materialised, forward, arg-max).  kind "reference": the UNMODIFIED reference modules
    from oracle/_ref (oracle/make_ref.sh) under the stub modules; kind "port": the oracle restatement (explicit Jacobian =
    op-for-op) when oracle/_ref is absent."""
    from oracle import rnf_oracle as orc
    torch.set_num_threads(threads)
    root = _reference_root() if prefer_reference else None
    kind = "port"
    fwd = None
    if root is not None:
        try:
            os.environ["RNF_REFERENCE_ROOT"] = root
            import importlib
            from oracle import ref_loader as rl
            rl = importlib.reload(rl)
            import warnings
            warnings.filterwarnings("ignore", message="Using torch.cross without specifying the dim")
            with contextlib.redirect_stdout(io.StringIO()):
                m = rl.build_reference_flow(cfg, 0)
            m.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()})
            fwd = lambda R, f: rl.run_reference(m, R, f)
            kind = "reference"
        except Exception:
            fwd = None
    if fwd is None:
        o = orc.OracleFlow(cfg, {k: v.detach().cpu() for k, v in state_dict.items()}, torch.float32, explicit_jacobian=True)
        fwd = o.forward
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
        _, ldj = fwd(samples, rows)
        _ = torch.argmax(ldj)
        times.append(time.perf_counter() - t0)
    return sample_rows / min(times), times, kind


def eager_gpu_leg(cfg, state_dict, dev, rows=500000, level=5, seed=321):
    """The practical incumbent: the reference's op sequence (oracle port, explicit Jacobian) in eager PyTorch on
    the B200, one 500 000-rotation chunk as at eval.py:445, features materialised per row as at eval.py:450."""
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
    cfg, flow = build_flow("symsol", feature_dim=2048)
    threads = os.cpu_count() or 1
    sample = args.cpu_sample
    vals, kind = [], "port"
    for i in range(args.warmup + args.steps):
        v, _, kind = cpu_leg(cfg, flow.state_dict(), sample, threads, seed=100 + i)
        if i >= args.warmup:
            vals.append(v)
    value = statistics.mean(vals)
    ms = 1000.0 * sample / value
    what = ("the unmodified reference modules flow/*.py (oracle/_ref) under the pytorch3d stub" if kind == "reference"
            else "oracle port of flow/*.py (explicit Jacobian)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: symsol.yml (F=2048, 42 layers) log_prob over HEALPix level-5 grid; CPU arm: bounded sample per step",
                   "grid_level": 5, "sample_rows_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"{sample} consecutive level-5 grid rotations x 1 image per step, {what}, torch CPU fp32, feature.repeat + forward + argmax as eval.py:444-462"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# GPU workloads
# ------------------------------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)   # > 126 MB L2
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = torch.tensor([x], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, step, steps, warmup, kernel_events=None):
        """`warmup` untimed steps, then `steps` steps timed with CUDA events on the current stream, the L2 flushed (256 MiB
        write, untimed) before every timed step; returns ms per step (max over ranks of the summed step times)."""
        for _ in range(warmup):
            step()
        self.barrier()
        evs = []
        for _ in range(steps):
            self.flush_buf.fill_(1)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            step()
            e.record()
            evs.append((s, e))
        self.barrier()
        total = sum(s.elapsed_time(e) for s, e in evs)
        return self.max_over_ranks(total) / steps

    def wall(self, step, steps, warmup):
        """End-to-end legs: wall clock around `steps` calls that end with their device->host read; -> (s per step, last result)."""
        for _ in range(warmup):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = step()
        self.barrier()
        return self.max_over_ranks(time.perf_counter() - t0) / steps, res


def _n_mobius(flow):
    return sum(1 for l in flow.layers if l.kind == "mobius")


def _roofline(pk, mode, rot_per_launch, kernel_s, n_mob, n_aff, inverse=False, kernel=None, traffic=None, evals=None):
    tensor = TENSOR_FLOPS_PER_MOBIUS * n_mob
    per_mob = (fp32_inv_per_mobius(evals) if evals is not None else FP32_INV_PER_MOBIUS) if inverse else FP32_FWD_PER_MOBIUS
    fp32 = per_mob * n_mob + AFFINE_FLOPS * n_aff
    ach = tensor * rot_per_launch / kernel_s / 1e12
    fp32_peak = 148 * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
    ach32 = fp32 * rot_per_launch / kernel_s / 1e12
    r = {"bound": "tensor", "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
         "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained", "kernel": kernel, "kernel_ms": kernel_s * 1e3,
         "issued": None if mode == "fp32" else {"achieved": SPLIT_ISSUE_FACTOR * ach, "frac": SPLIT_ISSUE_FACTOR * ach / pk["bf16_sustained"],
                                                "note": "fp16 MMAs actually issued (3-product split + bias / fc_first block)"},
         "algorithmic_flops_per_rotation": {"tensor_eligible": tensor, "fp32_pipe": fp32},
         "fp32_pipe": {"achieved": ach32, "peak_nominal": fp32_peak, "unit": "TFLOP/s", "frac_of_nominal": ach32 / fp32_peak}}
    if inverse:
        # the inverse is bound by the FP32 pipe of the CUDA cores (15 bisection probes x 64 components per layer),
        # not by the tensor pipe; no measured FP32 peak exists in MEASURED_PEAKS.json, so the nominal one is the denominator
        r["bound"] = "fp32 (CUDA cores; nominal peak 148 SMs x 128 lanes x 2 x sm_max_mhz -- MEASURED_PEAKS.json has no FP32 figure)"
        r["achieved"], r["peak"], r["frac"] = ach32, fp32_peak, ach32 / fp32_peak
        r["tensor_part"] = {"achieved": ach, "peak": pk["bf16_sustained"], "frac": ach / pk["bf16_sustained"]}
        ref32 = FP32_INV_PER_MOBIUS * n_mob + AFFINE_FLOPS * n_aff
        r["evaluations_per_sample_layer"] = {"executed": evals, "reference_algorithm": 15,
                                             "note": "Newton root + replay of the reference's 15 halvings; explicit evaluation only where the sign is within fp32 noise"}
        r["reference_algorithm_equivalent"] = {"fp32_flops_per_rotation": ref32, "achieved": ref32 * rot_per_launch / kernel_s / 1e12, "unit": "TFLOP/s",
                                               "note": "the same samples/s expressed in the FLOPs the reference's 15-probe bisection would execute"}
    return r


def _traffic(mode):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            return json.load(f).get(mode)
    return None


def _pieces(G_level, n_offsets, rank, world, level, dev):
    """This rank's share of the global grid = `n_offsets` copies of the level grid (index k*G_level + i = grid[i] @ offset_k),
    as a list of (resident grid slice, k, global index of its first rotation)."""
    from rotationnormflow_b200 import dist as rdist, grid as rgrid
    b, e = rdist.shard_range(n_offsets * G_level, rank, world)
    out = []
    for k in range(n_offsets):
        lo, hi = max(b, k * G_level), min(e, (k + 1) * G_level)
        if lo < hi:
            out.append((rgrid.healpix_grid(level, lo - k * G_level, hi - k * G_level, device=dev), k, lo))
    return out


def grid_workload(ctx, flow, level, n_offsets, B, mode, fisher_A=None, weak_copy=False, F=2048, seed=77):
    """Returns (step_device, step_e2e, info): one step = every piece of this rank's grid share x B images, reduced per image,
    merged over pieces and ranks."""
    from oracle import rnf_oracle as orc            # synthetic-input helper only (random_rotations), outside the timed region
    from rotationnormflow_b200 import dist as rdist, grid as rgrid
    from rotationnormflow_b200.fisher import fisher_constants
    from rotationnormflow_b200.flow import _program
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    G_level = 72 * 8 ** level
    gen = torch.Generator().manual_seed(1234 + (rank if weak_copy else 0))
    offsets = orc.random_rotations(n_offsets, gen).to(dev)
    if weak_copy:                                                  # config 2 at N > 1: own offset copy of the whole grid per rank
        pieces = [(rgrid.healpix_grid(level, device=dev), 0, rank * G_level)]
        G_total = world * G_level
    else:
        pieces = _pieces(G_level, n_offsets, rank, world, level, dev)
        G_total = n_offsets * G_level
    feat_host = torch.relu(torch.randn(B, F, generator=torch.Generator().manual_seed(seed))).pin_memory()
    feat_dev = feat_host.to(dev)
    A9 = c = A_host = None
    if fisher_A is not None:
        A_host = fisher_A.pin_memory()
        A9, c = fisher_constants(fisher_A.to(dev))
    prog = _program(flow, list(flow.layers), flow._perm_rows(), flow.feature_dim, dev)
    kern_ev = []

    def step_device(record=False):
        cond = prog.condition(feat_dev)
        parts = []
        for grid, k, g0 in pieces:
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            mx, am, se, _ = prog.grid_logprob(grid, g0, offsets[k], cond, B, A9, c, False, mode)
            if record:
                e1.record()
                kern_ev.append((e0, e1, grid.shape[0] * B))
            parts.append((mx, am, se))
        if len(parts) > 1:
            mx, am, se = rdist.merge_partials(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]),
                                              torch.stack([p[2] for p in parts]))
        elif parts:
            mx, am, se = parts[0]
        else:                                                      # a rank with an empty share
            mx = torch.full((B,), -math.inf, device=dev); am = torch.zeros(B, dtype=torch.int64, device=dev); se = torch.zeros(B, device=dev)
        return rdist.all_merge(mx, am, se)

    def step_e2e():
        f = feat_host.to(dev, non_blocking=True)
        A = None if A_host is None else A_host.to(dev, non_blocking=True)
        outs = [rdist.sharded_grid_log_prob(flow, grid, g0, G_total, f, offset=offsets[k], fisher_A=A, mlp_mode=mode)
                for grid, k, g0 in pieces] if len(pieces) == 1 else None
        if outs is not None:
            o = outs[0]
            return o["max"].cpu(), o["argmax"].cpu(), o["log_norm"].cpu()
        parts = [flow.grid_log_prob(grid, f, offset=offsets[k], fisher_A=A, g_index0=g0, mlp_mode=mode) for grid, k, g0 in pieces]
        mx, am, se = rdist.merge_partials(torch.stack([p["max"] for p in parts]), torch.stack([p["argmax"] for p in parts]),
                                          torch.stack([p["sumexp"] for p in parts]))
        mx, am, se = rdist.all_merge(mx, am, se)
        return mx.cpu(), am.cpu(), rdist.log_normaliser(mx, se, G_total).cpu()

    info = dict(G_total=G_total, rot_per_step=G_total * B, kern_ev=kern_ev, n_pieces=len(pieces),
                h2d=int(feat_host.numel() * 4 + (0 if A_host is None else A_host.numel() * 4)), d2h=int(B * 16),
                launches_per_step=2 + 2 * len(pieces))
    return step_device, step_e2e, info


def run_grid_config(ctx, cfg_id, mode, steps, warmup, images=None, want_e2e=True):
    """Configs 2, 3, 5."""
    pk = peaks()
    if cfg_id == 3:
        from oracle import rnf_oracle as orc
        cfg, flow = build_flow("modelnet_fisher")
        B, level, n_off, F = images or 256, 5, 2, 2080
        g = torch.Generator().manual_seed(3)
        U, V = orc.random_rotations(B, g), orc.random_rotations(B, g)
        s = torch.rand(B, 3, generator=g) * 19 + 1
        A = U @ torch.diag_embed(s) @ V.transpose(1, 2)
        weak, scaling = False, "strong"
        name = "BASELINE configs[2]: modelnet_fisher.yml (F=2080, 48 layers, matrix-Fisher base) log_prob + argmax over a 4 718 592-rotation grid (level 5 x 2 offsets) x %d images" % B
    elif cfg_id == 5:
        cfg, flow = build_flow("symsol", feature_dim=2048)
        B, level, n_off, F, A = images or 8, 6, 2, 2048, None
        weak, scaling = False, "strong"
        name = ("BASELINE configs[4]: symsol.yml (F=2048, 42 layers) log_prob + argmax + normaliser over the 37 748 736-rotation grid "
                "(level 6 x 2 offsets) split over the ranks, 1024 images streamed %d per step (one step = whole grid x %d images; 128 such steps = the full batch)" % (B, B))
    else:
        cfg, flow = build_flow("symsol", feature_dim=2048)
        B, level, n_off, F, A = images or 8, 5, 1, 2048, None
        weak, scaling = True, "weak"
        name = "BASELINE configs[1]: symsol.yml (F=2048, 42 layers) log_prob + argmax + normaliser over HEALPix level-5 grid"
    flow = flow.to(ctx.dev).eval()
    step_dev, step_e2e, info = grid_workload(ctx, flow, level, n_off, B, mode, fisher_A=A, weak_copy=weak, F=F)
    with ClockSampler(ctx.local) as clk:
        ms = ctx.timed(lambda: step_dev(record=True), steps, max(warmup, 3) if cfg_id != 3 else max(1, min(warmup, 3)))
    n_rec = len(info["kern_ev"])
    timed_ev = info["kern_ev"][-steps * info["n_pieces"]:] if n_rec else []
    k_s = sum(a.elapsed_time(b) for a, b, _ in timed_ev) * 1e-3
    k_rot = sum(n for _, _, n in timed_ev)
    value = info["rot_per_step"] / (ms * 1e-3)
    e2e_value, res = None, None
    if want_e2e:
        t, res = ctx.wall(step_e2e, max(1, min(steps, 5)), 1)
        e2e_value = info["rot_per_step"] / t
    rec = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": steps, "warmup": warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32" if mode == "fp32" else "f32 (conditioner GEMMs: split-fp16 tensor-core operands, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": name, "global_grid": info["G_total"], "images_per_step": B, "rotations_per_step": info["rot_per_step"],
                   "mlp_mode": mode, "l2": "flushed between timed steps (256 MiB write)", "weights": "random init seed 0",
                   "parallelism": (f"own offset copy of the grid per rank x{ctx.world}" if weak else f"grid index range sharded x{ctx.world}")
                                  + ", one all-gather of [B,3] f64 per step"},
        "roofline": _roofline(pk, mode, k_rot / max(1, len(timed_ev)), k_s / max(1, len(timed_ev)), _n_mobius(flow),
                              len(flow.layers) - _n_mobius(flow), kernel={"tc": "flow_t4_kernel", "tc_row": "flow_row_kernel"}.get(mode, "flow_v1_kernel"),
                              traffic=_traffic(mode) if cfg_id == 2 else None),
        "clocks": clk.summary(),
        "gpu_launches": int(steps * info["launches_per_step"]),
    }
    if want_e2e:
        rec["e2e"] = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": info["h2d"], "d2h_bytes_per_step": info["d2h"]}
        rec["check"] = {"argmax0": int(res[1][0]), "max0": float(res[0][0]), "log_norm0": float(res[2][0])}
    return rec, (cfg, flow)


def run_rows_config(ctx, mode, steps, warmup):
    """Config 1: raw.yml, Flow.forward (and inverse) on 100 000 uniform rotations; every rank runs a replica."""
    from rotationnormflow_b200 import grid as rgrid
    pk = peaks()
    cfg, flow = build_flow("raw")
    flow = flow.to(ctx.dev).eval()
    N = 100_000
    torch.manual_seed(1)
    R = rgrid.generate_queries(N, "random", ctx.dev)
    R_host = R.cpu().pin_memory()
    with torch.no_grad():
        flow(R, mlp_mode=mode)                                       # packs the weights (untimed, once)
        ev = []

        def step():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = flow(R, mlp_mode=mode)
            e1.record()
            ev.append((e0, e1))
            return out

        with ClockSampler(ctx.local) as clk:
            ms = ctx.timed(step, steps, max(warmup, 3))
        timed_ev = ev[-steps:]
        ms_inv = ctx.timed(lambda: flow.inverse(R, mlp_mode=mode), max(2, min(steps, 5)), 2)

        def e2e():
            Rz, ldj = flow(R_host.to(ctx.dev, non_blocking=True), mlp_mode=mode)
            return Rz.cpu(), ldj.cpu()

        t, (Rz, ldj) = ctx.wall(e2e, max(2, min(steps, 5)), 2)
    k_s = statistics.mean(a.elapsed_time(b) for a, b in timed_ev) * 1e-3
    n_mob = _n_mobius(flow)
    rec = {
        "metric": METRIC_ROWS, "value": ctx.world * N / (ms * 1e-3), "unit": UNIT, "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if mode == "fp32" else "f32 (conditioner GEMMs: split-fp16 tensor-core operands, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": "BASELINE configs[0]: raw.yml (unconditional, 24 Mobius + 24 affine layers) Flow.forward on 100 000 uniform rotations",
                   "rows": N, "mlp_mode": mode, "l2": "flushed between timed steps (256 MiB write)", "weights": "random init seed 0",
                   "parallelism": f"replicas x{ctx.world} (782 tiles = 1.3 waves of 148 SMs x 4 tiles: the launch is tail-bound)"},
        "roofline": _roofline(pk, mode, N, k_s, n_mob, len(flow.layers) - n_mob, kernel="flow_t4_kernel"),
        "clocks": clk.summary(),
        "e2e": {"value": ctx.world * N / t, "unit": UNIT, "h2d_bytes_per_step": N * 36, "d2h_bytes_per_step": N * 40},
        "gpu_launches": int(steps),
        "sampling": {"value": ctx.world * N / (ms_inv * 1e-3), "unit": "samples/s", "ms": ms_inv, "config": "Flow.inverse on the same 100 000 rotations"},
        "check": {"mean_exp_ldj": float(torch.exp(ldj.double()).mean()), "orthonormality": float((Rz.double() @ Rz.double().transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max())},
    }
    return rec, (cfg, flow)


def count_evaluations(flow, rows, feat, idx, mode):
    """Evaluations of the mixture map per (sample, Mobius layer) the inverse kernel executes, from its device counter (untimed run)."""
    import ctypes as C
    from rotationnormflow_b200 import _cabi
    lib = _cabi.load()
    counter = torch.zeros(1, dtype=torch.int64, device=rows.device)
    lib.rnf_debug_set_probe_counter.argtypes = [C.c_void_p]
    lib.rnf_debug_set_probe_counter.restype = None
    lib.rnf_debug_set_probe_counter(C.c_void_p(counter.data_ptr()))
    try:
        flow.inverse(rows, feat, feature_index=idx, mlp_mode=mode)
        torch.cuda.synchronize()
    finally:
        lib.rnf_debug_set_probe_counter(C.c_void_p(0))
    n = -(-rows.shape[0] // 128) * 128                              # whole tiles execute
    return float(counter.item()) / (n * _n_mobius(flow)) if mode != "fp32" else 15.0


def run_sampling_config(ctx, mode, steps, warmup, n_img=64, n_per=1_000_000):
    """Config 4: symsol2.yml Flow.inverse, n_per base rotations per image x n_img images; images split over the ranks."""
    from rotationnormflow_b200 import grid as rgrid
    pk = peaks()
    cfg, flow = build_flow("symsol2")
    flow = flow.to(ctx.dev).eval()
    lo = ctx.rank * n_img // ctx.world
    hi = (ctx.rank + 1) * n_img // ctx.world
    b = max(hi - lo, 0)
    torch.manual_seed(5)
    base = rgrid.generate_queries(n_per, "random", ctx.dev)          # one set of base rotations shared by all images (agent.py:253-256)
    feat_all = torch.relu(torch.randn(n_img, 512, generator=torch.Generator().manual_seed(5)))
    feat = feat_all[lo:hi].to(ctx.dev)
    feat_host, base_host = feat_all[lo:hi].pin_memory(), base.cpu().pin_memory()
    rows = base[None].expand(b, n_per, 3, 3).reshape(-1, 3, 3)
    idx = torch.arange(b, device=ctx.dev, dtype=torch.int32).repeat_interleave(n_per)
    ev = []
    with torch.no_grad():
        def step():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = flow.inverse(rows, feat, feature_index=idx, mlp_mode=mode)
            e1.record()
            ev.append((e0, e1))
            return out

        with ClockSampler(ctx.local) as clk:
            ms = ctx.timed(step, steps, max(1, min(warmup, 2)))
        timed_ev = ev[-steps:]

        def e2e():                                                   # agent.py:238-266: samples + arg-max of -ldj per image, host in / host out
            bs = base_host.to(ctx.dev, non_blocking=True)
            f = feat_host.to(ctx.dev, non_blocking=True)
            r = bs[None].expand(b, n_per, 3, 3).reshape(-1, 3, 3)
            S, ldj = flow.inverse(r, f, feature_index=idx, mlp_mode=mode)
            best = torch.argmax(-ldj.reshape(b, n_per), dim=-1)
            return S.reshape(b, n_per, 3, 3)[torch.arange(b, device=ctx.dev), best].cpu()

        t, _ = ctx.wall(e2e, max(1, min(steps, 3)), 1)
        S, ldj = step()
        evals = count_evaluations(flow, rows[:262144], feat, idx[:262144], mode)
        chk = min(b * n_per, 200_000)
        Rf, lf = flow(S[:chk], feat, feature_index=idx[:chk], mlp_mode=mode)
    k_s = statistics.mean(a.elapsed_time(c) for a, c in timed_ev) * 1e-3
    n_mob = _n_mobius(flow)
    total = n_img * n_per
    rec = {
        "metric": METRIC_SAMPLING, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": ctx.world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32" if mode == "fp32" else "f32 (conditioner GEMMs: split-fp16 tensor-core operands, fp32 accumulate)", "data": "synthetic",
        "config": {"workload": f"BASELINE configs[3]: symsol2.yml (F=512, 42 layers) Flow.inverse (BinFind root-solve), {n_per} base rotations per image x {n_img} images",
                   "samples_per_step": total, "mlp_mode": mode, "l2": "flushed between timed steps (256 MiB write)", "weights": "random init seed 0",
                   "parallelism": f"images split over {ctx.world} rank(s), no collective"},
        "roofline": _roofline(pk, mode, b * n_per, k_s, n_mob, len(flow.layers) - n_mob, inverse=True, kernel="flow_row_kernel<inverse>", evals=evals),
        "clocks": clk.summary(),
        "e2e": {"value": total / t, "unit": UNIT, "h2d_bytes_per_step": int(n_per * 36 + b * 512 * 4), "d2h_bytes_per_step": int(b * 36)},
        "gpu_launches": int(steps * 3),
        "check": {"round_trip_max": float((Rf - rows[:chk]).abs().max()), "ldj_antisymmetry_max": float((lf + ldj[:chk]).abs().max())},
    }
    return rec, (cfg, flow)


def run_dropin(ctx, mode, rows=500_000):
    """The literal drop-in calls: eval.py:444-453 scores a 500 000-rotation chunk with
    `flow(samples, feature.repeat(len(samples), 1))`; agent.py:240-261 calls `flow.inverse(base, feature.repeat...)`.  The
    repeated feature tensor (4 GB at F=2048) is built inside the timed region, as the reference does; the flow de-duplicates
    it on the device (csrc/dedup.cu, one streaming read) and reads the 4-byte run count back (N > the optimistic capacity)."""
    from oracle import rnf_oracle as orc
    from rotationnormflow_b200 import engine, grid as rgrid
    out = {}
    for F in (2048, 512):
        cfg, flow = build_flow("symsol", feature_dim=F)
        flow = flow.to(ctx.dev).eval()
        chunk = rgrid.healpix_grid(5, 0, rows, device=ctx.dev)
        off = orc.random_rotations(1, torch.Generator().manual_seed(9))[0].to(ctx.dev)
        f1 = torch.relu(torch.randn(1, F, generator=torch.Generator().manual_seed(10))).to(ctx.dev)
        with torch.no_grad():
            def fwd():
                samples = chunk @ off
                feats = f1.repeat(rows, 1)
                _, ldj = flow(samples, feats, mlp_mode=mode)
                return torch.argmax(ldj)

            def inv():
                feats = f1.repeat(rows, 1)
                S, ldj = flow.inverse(chunk, feats, mlp_mode=mode)
                return torch.argmax(-ldj)

            def direct():                                            # the same chunk through the N1 API (no repeated features)
                o = flow.grid_log_prob(chunk, f1, offset=off, mlp_mode=mode)
                return o["argmax"]

            ms_f = ctx.timed(fwd, 5, 3)
            ms_i = ctx.timed(inv, 3, 2)
            ms_d = ctx.timed(direct, 5, 3)
            feats = f1.repeat(rows, 1)
            ms_dd = ctx.timed(lambda: engine.dedup_rows(feats, engine.DEDUP_CAP), 5, 3)
            del feats
        out[f"F{F}"] = {"forward_rot_per_s": rows / (ms_f * 1e-3), "inverse_rot_per_s": rows / (ms_i * 1e-3),
                        "grid_log_prob_same_chunk_rot_per_s": rows / (ms_d * 1e-3), "forward_ms": ms_f, "inverse_ms": ms_i,
                        "dedup_ms": ms_dd, "dedup_read_GBps": rows * F * 4 / (ms_dd * 1e-3) / 1e9,
                        "repeated_feature_bytes": rows * F * 4}
        del flow, chunk
        torch.cuda.empty_cache()
    out["config"] = f"symsol.yml, one image, {rows}-rotation chunk of the level-5 grid; feature.repeat inside the timed region; L2 flushed between steps"
    return out


def run_train_step(ctx, batch=4096):
    """Training step of the differentiable path (agent.py:60-87: NLL of a batch, loss.backward(), optimiser step): symsol.yml
    (F = 512), one feature row per rotation as in training.  Reported next to the inference figures; not BASELINE's metric."""
    from oracle import rnf_oracle as orc
    cfg, flow = build_flow("symsol")
    flow = flow.to(ctx.dev).train()
    gen = torch.Generator().manual_seed(12)
    R = orc.random_rotations(batch, gen).to(ctx.dev)
    feat = torch.relu(torch.randn(batch, 512, generator=gen)).to(ctx.dev)
    opt = torch.optim.Adam(flow.parameters(), lr=1e-4)
    losses = []

    def step():
        _, ldj = flow(R, feat)
        loss = -ldj.mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())

    ms = ctx.timed(step, 5, 3)
    return {"rotations_per_s": batch / (ms * 1e-3), "ms_per_step": ms, "batch": batch,
            "loss_first_last": [float(losses[0]), float(losses[-1])],
            "config": "symsol.yml (F=512, 42 layers), forward + backward + Adam step through rotationnormflow_b200.train (conditioner MLP: cuBLAS FP32; Mobius mixture / calculate_16 forward + VJP: csrc/train_ops.cu)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=0, choices=[0, 1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (1-based); 0 = config 2 on one GPU, config 5 (strong scaling) under torchrun")
    ap.add_argument("--images", type=int, default=int(os.environ.get("RNF_BENCH_IMAGES", "0")), help="images per step (grid configs)")
    ap.add_argument("--mode", default=os.environ.get("RNF_BENCH_MODE", ""))
    ap.add_argument("--cpu-sample", type=int, default=160000)      # ~10 s of host work per pass at ~16 k rot/s on 16 cores
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the short passes of the other BASELINE configs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    import rotationnormflow_b200 as rnf  # noqa: F401
    from rotationnormflow_b200 import engine
    ctx = Ctx(args)
    mode = args.mode or ("tc" if engine.TC_AVAILABLE else "fp32")
    cfg_id = args.config or (2 if ctx.world == 1 else 5)
    images = args.images or None
    t_start = time.perf_counter()

    if cfg_id == 1:
        line, (cfg, flow) = run_rows_config(ctx, mode, args.steps, args.warmup)
    elif cfg_id == 4:
        line, (cfg, flow) = run_sampling_config(ctx, mode, args.steps, args.warmup)
    else:
        line, (cfg, flow) = run_grid_config(ctx, cfg_id, mode, args.steps, args.warmup, images=images)

    default_run = args.config == 0
    if default_run:
        # sampling = the second figure of BASELINE.json's metric: config 4 at its stated size (64 images x 1 000 000 samples)
        s_rec, _ = run_sampling_config(ctx, mode, 2, 1)
        line["sampling"] = {"value": s_rec["value"], "unit": "samples/s", "ms": s_rec["ms_per_step"], "config": s_rec["config"]["workload"],
                            "roofline": s_rec["roofline"], "e2e": s_rec["e2e"], "check": s_rec["check"]}
    if default_run and ctx.world == 1 and not args.no_extra_configs:
        extra = {}
        r1, _ = run_rows_config(ctx, mode, 5, 3)
        extra["1"] = {k: r1[k] for k in ("metric", "value", "ms_per_step", "config", "roofline", "e2e", "sampling", "check")}
        r3, _ = run_grid_config(ctx, 3, mode, 1, 1)
        extra["3"] = {k: r3[k] for k in ("metric", "value", "ms_per_step", "config", "roofline", "e2e", "check")}
        r5, _ = run_grid_config(ctx, 5, mode, 2, 3, want_e2e=False)
        extra["5"] = {k: r5[k] for k in ("metric", "value", "ms_per_step", "config", "roofline")}
        line["e2e_dropin"] = run_dropin(ctx, mode)
        try:
            line["train_step"] = run_train_step(ctx)
        except Exception as e:  # an extra leg: never fail the bench line over it
            line["train_step"] = {"error": repr(e)[:200]}
        extra["5"]["note"] = "one GPU scoring the whole 37.7 M grid: the N = 1 point of the strong-scaling series the default run measures under torchrun"
        line["configs"] = extra
    if ctx.rank == 0 and ctx.world == 1 and cfg_id == 2 and not args.no_cpu_baseline:
        try:
            del ctx.flush_buf
            torch.cuda.empty_cache()
            v_eager = eager_gpu_leg(cfg, {k: t.detach() for k, t in flow.state_dict().items()}, ctx.dev)
            line["eager_torch_gpu_baseline"] = {"value": v_eager, "unit": UNIT, "kind": "port",
                                                "sample": "one 500 000-rotation chunk x 1 image (eval.py:445), oracle port of flow/*.py as eager PyTorch ops on the same B200, best of 3"}
        except Exception as e:  # a baseline only: never fail the bench line over it
            line["eager_torch_gpu_baseline"] = {"value": None, "error": repr(e)[:200]}
        threads = os.cpu_count() or 1
        v, times, kind = cpu_leg(cfg, {k: t.cpu() for k, t in flow.state_dict().items()}, args.cpu_sample, threads, repeats=2)
        what = "the unmodified reference modules flow/*.py from oracle/_ref under the pytorch3d stub" if kind == "reference" else "oracle port of flow/*.py"
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"{args.cpu_sample} consecutive level-5 grid rotations x 1 image ({what}, torch CPU fp32), best of 2"}
    line["wall_s_total"] = time.perf_counter() - t_start
    if ctx.rank == 0:
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
