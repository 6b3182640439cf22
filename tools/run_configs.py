"""Runs the five BASELINE.json configurations once each on one B200 and records rate + parity on a subset
(`profiles/r01_configs.json`).  Config 5 is run at its per-GPU share (one of eight grid shards x 1024 images / 8 ... see notes)."""
import contextlib, io, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rotationnormflow_b200 as rnf
from oracle import rnf_oracle as orc
from rotationnormflow_b200 import grid as rgrid

dev = torch.device("cuda", 0)
out = {"mlp_mode": __import__("rotationnormflow_b200.engine", fromlist=["x"]).default_mlp_mode()}


def build(name, **ov):
    cfg = rnf.load_config(name, **ov)
    torch.manual_seed(0); np.random.seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        return cfg, rnf.get_flow(cfg).to(dev).eval()


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e-3, r


def rel(a, b):
    return float(((a - b).abs() / b.abs().clamp(min=1)).max())


with torch.no_grad():
    # ---- config 1: raw.yml, 100 000 uniform rotations, forward + inverse, full oracle comparison on a 2 000-row subset
    cfg, flow = build("raw")
    R = rgrid.generate_queries(100_000, "random", dev)
    t_f, (Rz, ldj) = timed(lambda: flow(R))
    t_i, (Ri, li) = timed(lambda: flow.inverse(R))
    sel = torch.arange(0, 100_000, 50)
    o = orc.OracleFlow(cfg, flow.state_dict(), torch.float64)
    Ro, lo = o.forward(R[sel].cpu())
    out["config1_raw_100k"] = dict(forward_rot_per_s=100_000 / t_f, inverse_rot_per_s=100_000 / t_i,
                                   max_abs_dR=float((Rz[sel].cpu().double() - Ro).abs().max()), rel_dldj=rel(ldj[sel].cpu().double(), lo),
                                   mean_exp_ldj=float(torch.exp(ldj.double()).mean()))
    # ---- config 2: symsol F=2048, level-5 grid, 8 images (the bench workload)
    cfg, flow = build("symsol", feature_dim=2048)
    grid = rgrid.healpix_grid(5)
    feat = torch.relu(torch.randn(8, 2048, generator=torch.Generator().manual_seed(1))).to(dev)
    off = orc.random_rotations(1, torch.Generator().manual_seed(2))[0].to(dev)
    t, res = timed(lambda: flow.grid_log_prob(grid, feat, offset=off))
    o = orc.OracleFlow(cfg, flow.state_dict(), torch.float64)
    idx = torch.arange(0, grid.shape[0], 1181)
    lp = flow.grid_log_prob(grid[idx], feat[:1], offset=off, return_logp=True)["logp"][0].cpu().double()
    _, l64 = o.forward((grid[idx] @ off).cpu(), feat[:1].cpu().expand(idx.numel(), -1))
    out["config2_symsol_grid2.4M_x8"] = dict(rot_per_s=grid.shape[0] * 8 / t, rel_dlogp_subset=rel(lp, l64),
                                             log_norm=[float(v) for v in (res["max"] + torch.log(res["sumexp"]) - math.log(grid.shape[0])).cpu()])
    # ---- config 3: modelnet_fisher (F=2080), 2 x level-5 grid (two offsets) x 256 images, Fisher base, argmax
    cfg, flow = build("modelnet_fisher")
    B = 256
    g = torch.Generator().manual_seed(3)
    feat = torch.relu(torch.randn(B, 2080, generator=g)).to(dev)
    U = orc.random_rotations(B, g); V = orc.random_rotations(B, g)
    s = torch.rand(B, 3, generator=g) * 19 + 1
    A = (U @ torch.diag_embed(s) @ V.transpose(1, 2)).to(dev)
    offs = orc.random_rotations(2, g).to(dev)
    def cfg3():
        parts = [flow.grid_log_prob(grid, feat, offset=offs[k], fisher_A=A, g_index0=k * grid.shape[0]) for k in range(2)]
        from rotationnormflow_b200 import dist as rdist
        return rdist.merge_partials(torch.stack([p["max"] for p in parts]), torch.stack([p["argmax"] for p in parts]),
                                    torch.stack([p["sumexp"] for p in parts]))
    t, (mx, am, se) = timed(cfg3, reps=1)
    # parity of the arg-max value for 2 images against the oracle evaluated at the chosen grid point
    o = orc.OracleFlow(cfg, flow.state_dict(), torch.float64)
    errs = []
    for b in (0, 100):
        k, gi = divmod(int(am[b]), grid.shape[0])
        Rq = (grid[gi:gi + 1] @ offs[k]).cpu()
        Rb, l64 = o.forward(Rq, feat[b:b + 1].cpu())
        l64 = l64 + orc.fisher_log_prob(A[b:b + 1].cpu().double(), Rb)
        errs.append(abs(float(l64) - float(mx[b])) / max(1.0, abs(float(l64))))
    out["config3_modelnet_fisher_grid4.7M_x256"] = dict(rot_per_s=2 * grid.shape[0] * B / t, seconds=t, rel_err_max_logp=max(errs))
    # ---- config 4: symsol2 (F=512) inverse sampling, 64 images x 1 000 000 base samples
    cfg, flow = build("symsol2")
    n_img, n_per = 64, 1_000_000
    feat = torch.relu(torch.randn(n_img, 512, generator=torch.Generator().manual_seed(5))).to(dev)
    base = rgrid.generate_queries(n_per, "random", dev)
    def cfg4():
        tot = 0.0
        for b0 in range(0, n_img, 8):                                  # 8 images x 1M rows per call
            rows = base[None].expand(8, n_per, 3, 3).reshape(-1, 3, 3)
            idx = torch.arange(8, device=dev, dtype=torch.int32).repeat_interleave(n_per)
            Rs, ls = flow.inverse(rows, feat[b0:b0 + 8], feature_index=idx)
            tot += float(ls[:1])
        return Rs, ls
    t, (Rs, ls) = timed(cfg4, reps=1)
    Rf, lf = flow(Rs[:200_000], feat[56:64], feature_index=torch.zeros(200_000, device=dev, dtype=torch.int32))   # image 56
    out["config4_symsol2_inverse_1M_x64"] = dict(samples_per_s=n_img * n_per / t, seconds=t,
                                                 round_trip_max=float((Rf - base[:200_000]).abs().max()),
                                                 ldj_antisymmetry_max=float((lf + ls[:200_000]).abs().max()))
    # ---- config 5 (per-GPU share of 8): one level-6 half-grid shard = 37.7M / 8 = 4 718 592 rotations x 1024 images
    cfg, flow = build("symsol", feature_dim=2048)
    B = 1024
    feat = torch.relu(torch.randn(B, 2048, generator=torch.Generator().manual_seed(6))).to(dev)
    shard = rgrid.healpix_grid(6, 0, 4_718_592)
    t, res = timed(lambda: flow.grid_log_prob(shard, feat), reps=1)
    out["config5_per_gpu_share_grid4.7M_x1024"] = dict(rot_per_s=shard.shape[0] * B / t, seconds=t,
                                                       note="one rank's shard of the 37 748 736-rotation grid (level 6, first eighth) x 1024 images")
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r01_configs.json", "w"), indent=1)
