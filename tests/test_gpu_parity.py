"""Parity of the CUDA path (through the drop-in module -> ctypes -> C ABI -> kernels) against the golden vectors of the
real reference and against the CPU oracle.  Run on the B200 box:  python -m pytest tests -m gpu

Tolerances (BASELINE.json north_star): rotations within 1e-5 max-abs, log-probs within 1e-4 relative
(|d| / max(|ref|, 1): ldj crosses zero), grid indices / arg-max bit-exact (ties: see test_grid_*).
Where the reference's own fp32 run is further than that from its fp64 run (F=2080 ModelNet: 4.5e-5 / 6.2e-5, SURVEY.md 7.2
item 1) the bound is the reference's own fp32-vs-fp64 error, read from the golden file (_tolerances).
"""
import math
import os

import numpy as np
import pytest
import torch

from conftest import ABLATION_CASES, CLU_CASES, FULL_CASES, GOLDEN, NOT_BIJECTIVE, SMALL_CASES, golden, record_error, seeded_product_flow
from oracle import rnf_oracle as orc
import rotationnormflow_b200 as rnf
from rotationnormflow_b200 import grid as rgrid
from rotationnormflow_b200.fisher import fisher_constants

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _golden_svd_convention(monkeypatch):
    """The golden vectors of the 4-D rotation layers (16Rot / 16UnRot) were minted by the reference on the CPU: LAPACK's sign
    convention defines their U^T V (engine.Program._polar_factor).  test_rotation_layers_svd_backend covers the default."""
    monkeypatch.setenv("RNF_SVD_BACKEND", "cpu")

MODES = [m for m in os.environ.get("RNF_TEST_MODES", "fp32,tc,tc_row").split(",") if m]


def _mode_available(mode):
    if mode == "fp32":
        return True
    from rotationnormflow_b200 import engine
    return getattr(engine, "TC_AVAILABLE", False)


def _product(g, dev="cuda"):
    m = seeded_product_flow(g.cfg, g.seed)
    sd = g.stored_state_dict()
    if sd is not None:
        m.load_state_dict(sd)
    return m.to(dev).eval()


def rel(a, ref):
    return ((a - ref).abs() / ref.abs().clamp(min=1.0)).max().item()


# BASELINE.json north_star: rotations within 1e-5 max-abs, log-probs within 1e-4 relative.  Where the REFERENCE's own fp32 run is
# further than that from its fp64 run (the F=2080 ModelNet stack: 4.5e-5 / 6.2e-5) no fp32 implementation can
# be held to the fixed number; the principled bound is then "no further from the fp64 reference than the reference's own fp32 run"
# (both outputs are in the golden file), see _tolerances().


# Two fp32 evaluations of the same ill-conditioned stack are two draws of the same rounding-noise distribution, and a max over the
# rows of one draw does not bound the max of the other: the product modes (tc, tc_row) get a factor 2 on the reference's own
# error, the CUDA-core cross-check kernel (mode "fp32": plain FMA chains for the GEMMs, not the product path) a factor 4.
OWN_ERROR_FACTOR = {"tc": 2.0, "tc_row": 2.0, "fp32": 4.0}


def _tolerances(g, direction="fwd", mode="tc"):
    own_R = (g.out(direction, "R", "f32").double() - g.out(direction, "R", "f64").double()).abs().max().item()
    own_l = rel(g.out(direction, "ldj", "f32").double(), g.out(direction, "ldj", "f64").double())
    # (Condition16TransLU runs in the per-layer operators whatever the mode: library GEMMs + FP32 kernels.  ConditionLU's batch
    #  diagonal makes its random-init 4x4 badly conditioned, which amplifies the different summation orders of two fp32 GEMMs: 8x)
    k = 8.0 if g.tag in CLU_CASES else OWN_ERROR_FACTOR[mode]
    return max(1e-5, k * own_R), max(1e-4, k * own_l), own_R, own_l
# 4-D rotation layers use U^T V of torch.svd(I + 1e-3 noise): nearly degenerate singular values make that matrix
# precision dependent (the reference's own fp32 and fp64 runs differ by 7e-3), so those cases are held to the fp32 run.
TRUTH = {"s_rot": "f32", "s_rotc": "f32", "s_unrot": "f32"}


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", SMALL_CASES + FULL_CASES)
def test_forward_parity(tag, mode):
    if not _mode_available(mode):
        pytest.skip("tensor-core conditioner not built in this revision")
    g = golden(tag)
    m = _product(g)
    feat = None if g.rows is None else g.rows.cuda()
    with torch.no_grad():
        R, ldj = m(g.R.cuda(), feat, mlp_mode=mode)
    R, ldj = R.cpu().double(), ldj.cpu().double()
    truth = TRUTH.get(tag, "f64")
    dR64 = (R - g.out("fwd", "R", truth).double()).abs().max().item()
    dl64 = rel(ldj, g.out("fwd", "ldj", truth).double())
    dR32 = (R - g.out("fwd", "R", "f32").double()).abs().max().item()
    print(f"\n[{tag}/{mode}] fwd  max|dR| vs ref-fp64 {dR64:.2e}  vs ref-fp32 {dR32:.2e}   rel dldj vs ref-fp64 {dl64:.2e}")
    tol_R, tol_l, own_R, own_l = _tolerances(g, "fwd", mode) if truth == "f64" else (1e-5, 1e-4, 0.0, 0.0)
    record_error(test="forward", case=tag, mode=mode, truth=truth, max_abs_dR=dR64, max_abs_dR_vs_ref_fp32=dR32, rel_dldj=dl64,
                 ref_fp32_vs_fp64_dR=own_R, ref_fp32_vs_fp64_dldj=own_l, tol_dR=tol_R, tol_dldj=tol_l)
    assert dR64 <= tol_R
    assert dl64 <= tol_l
    # outputs stay rotations
    assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 5e-6
    # the same call with features given once per image + a row index
    if feat is not None:
        with torch.no_grad():
            R2, l2 = m(g.R.cuda(), g.feat.cuda(), feature_index=g.feat_index.cuda(), mlp_mode=mode)
        # bit-identical -- except for the conditional ablation layers, whose per-image 3x3 / 6x6 matrix networks run as library
        # GEMMs whose reduction order depends on the number of rows handed over (capacity rows vs B rows)
        slack = 2e-6 if (tag in ABLATION_CASES and tag.endswith("c")) else 0.0
        assert (R2.cpu().double() - R).abs().max() <= slack and rel(l2.cpu().double(), ldj) <= 100 * slack


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", SMALL_CASES + FULL_CASES)
def test_inverse_parity(tag, mode):
    if not _mode_available(mode):
        pytest.skip("tensor-core conditioner not built in this revision")
    g = golden(tag)
    m = _product(g)
    feat = None if g.rows is None else g.rows.cuda()
    with torch.no_grad():
        R, ldj = m.inverse(g.R.cuda(), feat, mlp_mode=mode)
        Rf, lf = m(R, feat, mlp_mode=mode)
    nmob = sum(l.kind == "mobius" for l in m.layers)
    Rc, lc = R.cpu().double(), ldj.cpu().double()
    ref32, ref64 = g.out("inv", "R", "f32").double(), g.out("inv", "R", TRUTH.get(tag, "f64")).double()
    d32 = (Rc - ref32).abs().amax(dim=(1, 2))
    d64 = (Rc - ref64).abs().amax(dim=(1, 2))
    own = (ref32 - ref64).abs().amax(dim=(1, 2))           # the reference's own fp32-vs-fp64 flips
    frac32 = (d32 <= 1e-5).float().mean().item()
    frac64 = (d64 <= 1e-5).float().mean().item()
    frac_own = (own <= 1e-5).float().mean().item()
    print(f"\n[{tag}/{mode}] inv  rows within 1e-5: vs ref-fp32 {frac32:.3f}  vs ref-fp64 {frac64:.3f}  (ref fp32 vs ref fp64: {frac_own:.3f})"
          f"  worst {d64.max().item():.2e}")
    record_error(test="inverse", case=tag, mode=mode, rows_within_1e5_vs_ref_fp64=frac64, rows_within_1e5_vs_ref_fp32=frac32,
                 ref_fp32_rows_within_1e5_of_ref_fp64=frac_own, worst_row=d64.max().item(),
                 round_trip_max=(Rf.cpu().double() - g.R.double()).abs().max().item())
    # Bisection returns a dyadic angle of resolution pi/2^15; one-ulp noise at a probe flips a branch and moves the
    # row by up to 1.9e-4 per Mobius layer.  Bound: as many exact rows as the reference's own
    # fp32 run (minus slack), and no row further than the flip bound.
    assert frac64 >= min(frac_own, 0.97) - 0.05
    assert d64.max().item() <= max(2e-4 * nmob, 1e-5)
    # size-independent properties: forward(inverse(z)) ~ z to bisection resolution, ldj_inv = -ldj_fwd
    # (the random-init F=2080 ModelNet stack amplifies the pi/2^15 angle quantum to 1.4e-2 / 1.8e-2 in the reference
    #  itself, fp32 and fp64 alike -- measured with the oracle -- hence its own bound)
    rt_R, rt_l = {"modelnet": (3e-2, 4e-2)}.get(tag, (max(2e-4 * nmob, 1e-5), max(2e-3 * nmob, 1e-5)))
    if tag not in NOT_BIJECTIVE:             # (the reference's own `inverse` of those ablation layers is not the inverse map)
        assert (Rf.cpu().double() - g.R.double()).abs().max().item() <= rt_R
        assert (lf.cpu().double() + lc).abs().max().item() <= rt_l
    l64 = g.out("inv", "ldj", TRUTH.get(tag, "f64")).double()
    ok = d64 <= 1e-5
    if ok.any():
        assert rel(lc[ok], l64[ok]) <= 2e-4


@pytest.mark.parametrize("tag", ["s_uncond", "s_symsol", "s_modelnet", "s_rotc", "s_lu"])
def test_single_layer_protocol(tag):
    """layer(rotation, permute, feature) / layer.inverse(...) -- the per-layer protocol of flow/*.py -- against the oracle."""
    g = golden(tag)
    m = _product(g)
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    rows = orc.permute_rows(g.cfg, o.plan)
    feat = None if g.rows is None else g.rows
    R = g.R
    for i, layer in enumerate(m.layers):
        p = orc.PERMUTE_TABLE[rows[i]]
        perm = torch.tensor(p, dtype=torch.long, device="cuda")
        f_i = feat.cuda() if (feat is not None and layer.uses_feature) else None
        with torch.no_grad():
            Rg, lg = layer(R.cuda(), perm, f_i)
        kind, pre = o.plan[i], f"layers.{i}."
        f64 = None if feat is None else feat.double()
        if kind == "mobius":
            Ro, lo = orc.mobius_forward(o.sd, pre, R.double(), p, f64, 64)
        else:
            W, has = orc.affine_matrix(o.sd, pre, kind, f64, False)
            Ro, lo = orc.quat_affine(W, R.double(), has)
        assert (Rg.cpu().double() - Ro).abs().max() < 3e-6, (i, kind)
        assert (lg.cpu().double() - lo).abs().max() < 1e-5, (i, kind)
        R = Ro.float()


def test_edge_cases():
    g = golden("s_symsol")
    m = _product(g)
    F = g.feat.shape[1]
    with torch.no_grad():
        R0, l0 = m(torch.empty(0, 3, 3, device="cuda"), torch.empty(0, F, device="cuda"))
        assert R0.shape == (0, 3, 3) and l0.shape == (0,)
        # ragged: 1 row, 255/256/257 rows (tile boundary), each row its own feature
        for n in (1, 127, 128, 129, 255, 256, 257):
            R = g.R[:n].cuda()
            f = g.rows[:n].cuda()
            Rn, ln = m(R, f)
            Rall, lall = m(g.R.cuda(), g.rows.cuda())
            assert torch.equal(Rn, Rall[:n]) and torch.equal(ln, lall[:n])
        # every row a distinct image
        f = torch.relu(torch.randn(64, F, generator=torch.Generator().manual_seed(0))).cuda()
        Rn, ln = m(g.R[:64].cuda(), f)
        o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
        Ro, lo = o.forward(g.R[:64], f.cpu())
        assert (Rn.cpu().double() - Ro).abs().max() < 1e-5 and rel(ln.cpu().double(), lo) < 1e-4
        # inputs are not modified, non-contiguous input accepted
        Rin = g.R.cuda()
        keep = Rin.clone()
        m(Rin, g.rows.cuda())
        assert torch.equal(Rin, keep)
        Rt = g.R.cuda().transpose(1, 2).contiguous().transpose(1, 2)
        assert torch.equal(m(Rt, g.rows.cuda())[0], m(g.R.cuda(), g.rows.cuda())[0])
    with pytest.raises(AssertionError):
        m(g.R.cuda(), None)
    with pytest.raises(ValueError):
        m(g.R.cuda(), g.rows[:5].cuda())
    # checkpoint-style weight update is picked up (load_state_dict bumps parameter versions)
    with torch.no_grad():
        before = m(g.R.cuda(), g.rows.cuda())[1].clone()
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        sd["layers.1.conditioner.fc_last.bias"] += 0.25
        m.load_state_dict(sd)
        after = m(g.R.cuda(), g.rows.cuda())[1]
    assert (before - after).abs().max() > 1e-4


def test_tiles_in_flight_do_not_change_results(monkeypatch):
    """flow_t4 picks the number of tile slots (1..4, each with its own service warp) per launch; a rotation's arithmetic does not
    depend on the slot it lands in, so every choice gives bit-identical outputs -- also across several rounds per SM and with a
    ragged last tile group (launches of 1 .. 3 * 148 * 4 + 5 tiles)."""
    g = golden("s_symsol")
    m = _product(g)
    gen = torch.Generator().manual_seed(11)
    for n in (300, 128 * 148 * 2 + 77, 128 * (3 * 148 * 4 + 5) - 9):
        R = orc.random_rotations(n, gen).float().cuda()
        f = g.feat[:1].cuda().repeat(n, 1) if n < 4096 else None
        outs = []
        for act in ("", "1", "2", "3", "4"):
            if act:
                monkeypatch.setenv("RNF_T4_ACTIVE", act)
            else:
                monkeypatch.delenv("RNF_T4_ACTIVE", raising=False)
            with torch.no_grad():
                if f is not None:
                    outs.append(m(R, f, mlp_mode="tc"))
                else:                                            # one image for all rows: the broadcast-aware entry point
                    outs.append(m(R, g.feat[:1].cuda(), feature_index=torch.zeros(n, dtype=torch.int32, device="cuda"), mlp_mode="tc"))
        monkeypatch.delenv("RNF_T4_ACTIVE", raising=False)
        for Rn, ln in outs[1:]:
            assert torch.equal(Rn, outs[0][0]) and torch.equal(ln, outs[0][1])
        assert torch.isfinite(outs[0][1]).all()


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("kw", [dict(layers=1), dict(layers=2), dict(layers=3, frequent_permute=1), dict(layers=2, dist="noflow"),
                                dict(layers=1, first_affine=0), dict(layers=2, condition=1, feature_dim=16, last_affine=1)],
                         ids=["one_layer", "two_layers", "frequent_permute", "noflow", "mobius_only_1", "cond_small"])
def test_short_and_odd_stacks(kw, mode):
    """Stacks that stress the weight pipeline of the persistent kernels (1 or 2 Mobius layers: every refill targets the layer
    in flight; no Mobius layer at all) in all three directions of use: rows forward, rows inverse, grid."""
    cfg = orc.simple_config(**kw)
    m = seeded_product_flow(cfg, 5).cuda().eval()
    o = orc.OracleFlow(cfg, {k: v.cpu() for k, v in m.state_dict().items()}, torch.float64)
    gen = torch.Generator().manual_seed(2)
    N = 700                                                              # 6 tiles: groups of four with idle tiles
    R = orc.random_rotations(N, gen)
    F = orc.feature_dim_of(cfg)
    B = 3
    feat = torch.relu(torch.randn(B, F, generator=gen)) if F else None
    rows = None if feat is None else feat.repeat_interleave((N + B - 1) // B, 0)[:N]
    with torch.no_grad():
        Rg, lg = m(R.cuda(), None if rows is None else rows.cuda(), mlp_mode=mode)
        Ro, lo = o.forward(R, rows)
        assert (Rg.cpu().double() - Ro).abs().max() < 1e-5 and rel(lg.cpu().double(), lo) < 1e-4
        Ri, li = m.inverse(Rg, None if rows is None else rows.cuda(), mlp_mode=mode)
        assert (Ri.cpu() - R).abs().max() < 2e-3 and (li + lg).abs().max() < 2e-3       # bisection resolution
        grid = orc.healpix_grid(1).float()
        out = m.grid_log_prob(grid.cuda(), None if feat is None else feat.cuda(), return_logp=True, mlp_mode=mode)
        Bn = 1 if feat is None else B
        _, l64 = o.forward(grid.double().repeat(Bn, 1, 1), None if feat is None else feat.double().repeat_interleave(grid.shape[0], 0))
        assert rel(out["logp"].cpu().double().reshape(-1), l64) < 1e-4
        assert torch.equal(out["argmax"].cpu(), out["logp"].argmax(dim=-1).cpu())


def test_grid_pass_is_cuda_graph_capturable():
    """The whole grid evaluation (per-image conditioning, fused flow kernel, combine) is asynchronous on the current stream
    with no host synchronisation: it can be captured once in a CUDA graph and replayed on new features."""
    g = golden("s_symsol")
    m = _product(g)
    grid = orc.healpix_grid(1).float().cuda()
    off = orc.random_rotations(1, torch.Generator().manual_seed(1))[0].cuda()
    feat = g.feat.cuda().clone()
    with torch.no_grad():
        eager = m.grid_log_prob(grid, feat, offset=off, return_logp=True)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m.grid_log_prob(grid, feat, offset=off, return_logp=True)          # warm-up on the capture stream
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            out = m.grid_log_prob(grid, feat, offset=off, return_logp=True)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out["logp"], eager["logp"]) and torch.equal(out["argmax"], eager["argmax"])
        feat.copy_(torch.relu(torch.randn(feat.shape, generator=torch.Generator().manual_seed(9))).cuda())   # new inputs, same graph
        graph.replay()
        torch.cuda.synchronize()
        fresh = m.grid_log_prob(grid, feat, offset=off, return_logp=True)
        assert torch.equal(out["logp"], fresh["logp"]) and torch.equal(out["argmax"], fresh["argmax"])
        assert not torch.equal(fresh["logp"], eager["logp"])


def test_healpix_grid_on_device():
    z = np.load(f"{GOLDEN}/healpix_grid.npz")
    for level in (0, 1, 2):
        G = rgrid.healpix_grid(level).cpu().numpy()
        ref = z[f"full_{level}"]
        assert G.shape == ref.shape
        assert np.abs(G - ref).max() <= 1.2e-7
        big = np.abs(ref) > 1e-9
        assert (G[big] != ref[big]).mean() < 1e-3
    for level in (3, 4):
        idx = z[f"idx_{level}"]
        G = rgrid.healpix_grid(level).cpu().numpy()
        assert np.abs(G[idx] - z[f"sample_{level}"]).max() <= 1.2e-7
        b = int(idx[7])
        assert np.array_equal(rgrid.healpix_grid(level, b, b + 1000).cpu().numpy(), G[b:b + 1000])
    # full BASELINE size (level 5, 2 359 296 rotations): orthonormal, det +1, matches the CPU oracle on a slice
    G5 = rgrid.healpix_grid(5)
    assert G5.shape == (2_359_296, 3, 3)
    Gd = G5.double()
    assert (Gd @ Gd.transpose(1, 2) - torch.eye(3, device="cuda", dtype=torch.float64)).abs().max() < 1e-6
    assert (torch.linalg.det(Gd) - 1).abs().max() < 1e-6
    sl = orc.healpix_grid(5, 1_234_567, 1_234_567 + 5000)
    assert (G5[1_234_567:1_234_567 + 5000].cpu() - sl).abs().max() <= 1.2e-7
    assert rgrid.generate_queries(2_400_000, "grid").shape[0] == 2_359_296
    assert rgrid.closest_grid_level(37_000_000) == 6


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("tag", ["s_symsol", "s_modelnet", "s_uncond"])
def test_grid_log_prob(tag, mode):
    """Fused grid evaluation == forward + Fisher base + torch reductions, and == the oracle on the same inputs."""
    if not _mode_available(mode):
        pytest.skip("tensor-core conditioner not built in this revision")
    g = golden(tag)
    m = _product(g)
    grid = rgrid.healpix_grid(2)                           # 4608 rotations: not a multiple of the tile sizes? (4608 = 18*256)
    grid = grid[:4500]                                     # ragged last tile
    gen = torch.Generator().manual_seed(11)
    off = orc.random_rotations(1, gen)[0]
    feat = g.feat.cuda() if g.feat is not None else None
    B = 1 if feat is None else feat.shape[0]
    A = torch.from_numpy(g.z["fisher_A"]).cuda() if tag == "s_modelnet" else None
    with torch.no_grad():
        out = m.grid_log_prob(grid, feat, offset=off.cuda(), fisher_A=A, return_logp=True, g_index0=1000, mlp_mode=mode)
        samples = grid @ off.cuda()
        rows = None if feat is None else feat.repeat_interleave(samples.shape[0], 0)
        Rb, ldj = m(samples.repeat(B, 1, 1), rows, mlp_mode=mode)
    logp_ref = ldj.reshape(B, -1)
    if A is not None:
        A9, c = fisher_constants(A)
        logp_ref = logp_ref + (Rb.reshape(B, -1, 9) * A9[:, None, :]).sum(-1) - c[:, None]
    assert (out["logp"] - logp_ref).abs().max() < 2e-5
    # the fused reduction is exactly the reduction of the log-probs the same kernel produced
    idx, mx, lme = orc.grid_reduce(out["logp"].cpu())
    assert torch.equal(out["argmax"].cpu(), idx + 1000)
    assert torch.equal(out["max"].cpu(), mx)
    lme_k = out["max"].cpu() + torch.log(out["sumexp"].cpu()) - math.log(samples.shape[0])
    assert (lme_k - lme).abs().max() < 1e-5
    # against the oracle (fp64) on the same inputs
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    s64 = (grid.cpu().double() @ off.double())
    lp = []
    for b in range(B):
        f = None if feat is None else g.feat[b:b + 1].double().expand(s64.shape[0], -1)
        Rb64, l64 = o.forward(s64, f)
        if A is not None:
            l64 = l64 + orc.fisher_log_prob(A[b:b + 1].cpu().double(), Rb64)
        lp.append(l64)
    lp = torch.stack(lp)
    assert rel(out["logp"].cpu().double(), lp) < 1e-4
    oidx, omx, olme = orc.grid_reduce(lp)
    agree = out["argmax"].cpu() - 1000 == oidx
    # arg-max must agree unless the oracle's own top-2 are closer than fp32 noise
    for b in range(B):
        if not agree[b]:
            assert abs(float(lp[b, oidx[b]] - lp[b, out["argmax"][b].item() - 1000])) < 1e-5
    assert (lme_k.double() - olme).abs().max() < 1e-4


@pytest.mark.parametrize("tag", ["s_modelnet", "modelnet"])
def test_matrix_fisher_log_prob(tag):
    """MatrixFisherN._log_prob against the reference's own output (golden) and the oracle."""
    from rotationnormflow_b200.fisher import MatrixFisherN
    g = golden(tag)
    A = torch.from_numpy(g.z["fisher_A"])
    base = g.out("fwd", "R", "f32")
    lp = MatrixFisherN(A.cuda())._log_prob(base.cuda()).cpu()
    ref = torch.from_numpy(g.z["fisher_logp_f32"])
    assert ((lp - ref).abs() / ref.abs().clamp(min=1)).max().item() < 1e-5
    lp64 = orc.fisher_log_prob(A.double(), base.double())
    assert rel(lp.double(), lp64) < 1e-5
    with pytest.raises(Exception):
        MatrixFisherN(A.cuda())._log_prob(base[:5].cuda())      # 5 rows do not split over 4 images


def test_eval_call_sites():
    """estimate_rotation_grid / estimate_rotation_sampling (eval.py:437-462, agent.py:238-266) against the oracle."""
    from rotationnormflow_b200 import evalpath
    g = golden("s_symsol")
    m = _product(g)
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    feat = g.feat.cuda()
    B = feat.shape[0]
    # grid: level 2 (4608 rotations)
    off = orc.random_rotations(1, torch.Generator().manual_seed(3))[0]
    est, out = evalpath.estimate_rotation_grid(m, feat, 4000, offset=off.cuda())
    grid = orc.healpix_grid(2).double() @ off.double()
    for b in range(B):
        _, l64 = o.forward(grid, g.feat[b:b + 1].double().expand(grid.shape[0], -1))
        k = int(out["argmax"][b])
        assert float(l64.max() - l64[k]) < 1e-5                       # the chosen grid point is the oracle's maximum (up to fp32 ties)
        assert (est[b].cpu().double() - grid[k]).abs().max() < 1e-6
    # sampling: 300 base rotations shared by all images
    base = orc.random_rotations(300, torch.Generator().manual_seed(4))
    est, samples, logp = evalpath.estimate_rotation_sampling(m, feat, 300, base_samples=base.cuda())
    assert samples.shape == (B, 300, 3, 3) and logp.shape == (B, 300)
    for b in range(B):
        Ri, li = o.inverse(base.double(), g.feat[b:b + 1].double().expand(300, -1))
        ref = -li
        close = (logp[b].cpu().double() - ref).abs() < 5e-3                 # bisection flips move a few samples by ~1e-4 rad
        assert close.float().mean() > 0.97
        k = int(logp[b].argmax())
        assert float(ref.max() - ref[k]) < 5e-3
        assert (est[b] - samples[b, k]).abs().max() == 0
    # sampling with the matrix-Fisher base (agent.py:247-251): base samples drawn on the device, then the same path;
    # the returned log-probs are the oracle's inverse-flow log-dets of those very samples plus the base log-likelihood
    gm = golden("s_modelnet")
    mm = _product(gm)
    om = orc.OracleFlow(gm.cfg, gm.state_dict(), torch.float64)
    fm = gm.feat.cuda()
    Bm = fm.shape[0]
    A = torch.randn(Bm, 3, 3, generator=torch.Generator().manual_seed(8)) * 2.0
    est, samples, logp = evalpath.estimate_rotation_sampling(mm, fm, 200, fisher_A=A.cuda(), seed=77)
    assert samples.shape == (Bm, 200, 3, 3) and logp.shape == (Bm, 200)
    from rotationnormflow_b200.fisher import MatrixFisherN
    pre = MatrixFisherN(A.cuda())
    base = pre._sample(200, seed=77)
    for b in range(Bm):
        Ri, li = om.inverse(base[b].cpu().double(), gm.feat[b:b + 1].double().expand(200, -1))
        ref = -li + orc.fisher_log_prob(A[b:b + 1].double(), base[b].cpu().double())
        close = (logp[b].cpu().double() - ref).abs() < 5e-3 * ref.abs().clamp(min=1.0)
        assert close.float().mean() > 0.95
        k = int(logp[b].argmax())
        assert (est[b] - samples[b, k]).abs().max() == 0


@pytest.mark.parametrize("mode", MODES)
def test_spread_metric_fused_in_the_grid_pass(mode):
    """sum_g p_g d(R_g, R_gt) / sum_g p_g from the reduction epilogue == the oracle's definition on the oracle's log-probs;
    two grid slices merge to the same value; K = 1 and K = 3 ground truths per image."""
    from rotationnormflow_b200 import dist as rdist
    g = golden("s_symsol")
    m = _product(g)
    B = 3
    feat = g.feat[:B].cuda()
    grid = orc.healpix_grid(1).float()                                   # 576 rotations
    gen = torch.Generator().manual_seed(11)
    off = orc.random_rotations(1, gen)[0]
    samples = grid @ off
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    G = grid.shape[0]
    _, ldj = o.forward(samples.double().repeat(B, 1, 1), g.feat[:B].double().repeat_interleave(G, 0))
    logp = ldj.reshape(B, G)
    for K in (1, 3):
        gt = orc.random_rotations(B * K, gen).reshape(B, K, 3, 3)
        want = orc.spread(logp, samples, gt)
        with torch.no_grad():
            out = m.grid_log_prob(grid.cuda(), feat, offset=off.cuda(), gt_rotations=gt.cuda(), mlp_mode=mode)
            assert (out["spread"].cpu().double() - want).abs().max() < 2e-4 * max(1.0, float(want.max()))
            assert torch.equal(out["argmax"].cpu(), torch.argmax(logp, dim=-1))
            a = m.grid_log_prob(grid[:200].cuda(), feat, offset=off.cuda(), gt_rotations=gt.cuda(), mlp_mode=mode)
            b = m.grid_log_prob(grid[200:].cuda(), feat, offset=off.cuda(), gt_rotations=gt.cuda(), g_index0=200, mlp_mode=mode)
        mx, am, se, sn = rdist.merge_partials(torch.stack([a["max"], b["max"]]), torch.stack([a["argmax"], b["argmax"]]),
                                              torch.stack([a["sumexp"], b["sumexp"]]), torch.stack([a["spread_num"], b["spread_num"]]))
        assert torch.equal(am, out["argmax"])
        assert ((sn / se) - out["spread"]).abs().max() < 1e-5


def test_matrix_fisher_sampling_on_device():
    """MatrixFisherN._sample (csrc/fisher_sample.cu) against the oracle sampler (utils/fisher.py:117-207 restated): valid
    rotations, reproducible per seed, first and second moments equal within Monte-Carlo error, per image."""
    from rotationnormflow_b200.fisher import MatrixFisherN
    gen = torch.Generator().manual_seed(17)
    B, n = 3, 200000
    U, V = orc.random_rotations(B, gen, torch.float64), orc.random_rotations(B, gen, torch.float64)
    S = torch.tensor([[3.0, 2.0, -1.0], [12.0, 7.0, 5.0], [0.6, 0.3, 0.1]], dtype=torch.float64)
    A = U @ torch.diag_embed(S) @ V.transpose(1, 2)
    d = MatrixFisherN(A.float().cuda())
    R = d._sample(n, seed=1234)
    assert R.shape == (B, n, 3, 3)
    assert torch.equal(R, d._sample(n, seed=1234)) and not torch.equal(R, d._sample(n, seed=1235))
    Rd = R.cpu().double()
    assert (Rd.transpose(-1, -2) @ Rd - torch.eye(3, dtype=torch.float64)).abs().max() < 5e-6
    assert (torch.linalg.det(Rd) - 1).abs().max() < 5e-6
    for b in range(B):
        Ro = orc.sample_matrix_fisher(A[b], n, gen)
        tol = 5.0 / (n ** 0.5)                                          # entries of R are bounded by 1: sigma <= 1 / sqrt(n)
        assert (Rd[b].mean(0) - Ro.mean(0)).abs().max() < tol
        t_dev, t_orc = (Rd[b] * A[b]).sum((-1, -2)), (Ro * A[b]).sum((-1, -2))
        assert abs(float(t_dev.mean() - t_orc.mean())) < 5.0 * float(t_orc.std()) * (2.0 / n) ** 0.5
        assert abs(float(t_dev.std() / t_orc.std()) - 1.0) < 0.03


def test_geodesic_metrics():
    from rotationnormflow_b200 import metrics
    gen = torch.Generator().manual_seed(9)
    est = orc.random_rotations(300, gen)
    gt = orc.random_rotations(300 * 7, gen).reshape(300, 7, 3, 3)
    gt[:10, 3] = est[:10]                                   # exact hits -> angle 0 (clip protects acos)
    got = metrics.min_geodesic_distance_rotmats(est.cuda(), gt.cuda()).cpu()
    ref = orc.min_geodesic_distance_rotmats(est.double(), gt.double())
    # acos amplifies the fp32 rounding of the trace by 1/sin(angle): compare where the angle is well conditioned
    ok = (ref > 0.2) & (ref < 2.9)
    ok[:10] = False
    assert ok.sum() > 200 and (got[ok].double() - ref[ok]).abs().max() < 5e-6 and got[:10].abs().max() < 1e-3
    single = metrics.geodesic_distance_rotmats(est.cuda(), gt[:, 0].cuda()).cpu()
    ref1 = orc.min_geodesic_distance_rotmats(est.double(), gt[:, :1].double())
    ok1 = (ref1 > 0.2) & (ref1 < 2.9)
    assert (single[ok1].double() - ref1[ok1]).abs().max() < 5e-6
    assert float(metrics.acc(torch.rad2deg(got), 30.0)) == float((torch.rad2deg(got) <= 30.0).float().mean())


def test_two_rank_grid_sharding_on_one_gpu():
    """Grid sharded in two slices + merge_partials == the unsharded fused reduction (same device, no process group)."""
    from rotationnormflow_b200 import dist as rdist
    g = golden("s_symsol")
    m = _product(g)
    grid = rgrid.healpix_grid(3)                            # 36 864 rotations
    feat = g.feat.cuda()
    with torch.no_grad():
        full = m.grid_log_prob(grid, feat)
        parts = []
        for r in range(2):
            b, e = rdist.shard_range(grid.shape[0], r, 2)
            parts.append(m.grid_log_prob(grid[b:e], feat, g_index0=b))
    mx, am, se = rdist.merge_partials(torch.stack([p["max"] for p in parts]), torch.stack([p["argmax"] for p in parts]),
                                      torch.stack([p["sumexp"] for p in parts]))
    assert torch.equal(am, full["argmax"]) and torch.equal(mx, full["max"])
    assert ((se - full["sumexp"]).abs() / full["sumexp"]).max().item() < 1e-5


def test_full_size_properties():
    """BASELINE config 1 at full size (100 000 rotations, raw.yml): size-independent checks."""
    g = golden("raw")
    m = _product(g)
    R = rgrid.generate_queries(100_000, "random")
    with torch.no_grad():
        Rz, ldj = m(R)
        norm = torch.exp(ldj.double()).mean().item()       # the reference's sanity print (eval_uncondition.py:114-116)
        Ri, li = m.inverse(Rz)
    print(f"\n[raw full] mean exp(ldj) over 100k Haar rotations = {norm:.5f}")
    assert abs(norm - 1.0) < 2e-2
    assert (Rz.double() @ Rz.double().transpose(1, 2) - torch.eye(3, device="cuda", dtype=torch.float64)).abs().max() < 5e-6
    assert (torch.linalg.det(Rz.double()) - 1).abs().max() < 5e-6
    assert (Ri - R).abs().max().item() < 2e-4 * 24
    assert (li + ldj).abs().max().item() < 2e-3 * 24
    # a strided sample against the CPU oracle
    sel = torch.arange(0, 100_000, 997)
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    Ro, lo = o.forward(R[sel].cpu())
    assert (Rz[sel].cpu().double() - Ro).abs().max() < 1e-5
    assert rel(ldj[sel].cpu().double(), lo) < 1e-4


def test_dedup_rows_on_device():
    """rnf_dedup_rows: run structure of a row-aligned feature tensor, on the device (consecutive runs only: A B A is three)."""
    from rotationnormflow_b200 import engine
    f = (torch.arange(12.0).reshape(3, 4) + 0.5).cuda()
    for F in (4, 5, 2048):                                              # float4 path, scalar path, several column chunks
        base = torch.randn(3, F, generator=torch.Generator().manual_seed(F)).cuda()
        rows = base[[0, 0, 0, 1, 1, 2]].contiguous()
        idx, first, count = engine.dedup_rows(rows, 6)
        assert idx.tolist() == [0, 0, 0, 1, 1, 2] and int(count) == 3 and first[:3].tolist() == [0, 3, 5]
        idx, first, count = engine.dedup_rows(base[[0, 1, 0]].contiguous(), 3)
        assert idx.tolist() == [0, 1, 2] and int(count) == 3
        idx, first, count = engine.dedup_rows(base[:1].repeat(7, 1), 7)
        assert idx.tolist() == [0] * 7 and int(count) == 1 and int(first[0]) == 0
        one = base[[0, 0, 1]].contiguous()
        one[2, F - 1] = one[1, F - 1]                                     # rows differing in a single element elsewhere
        one[2, 0] += 1.0
        assert engine.dedup_rows(one, 3)[0].tolist() == [0, 0, 1]
    # many rows, many runs of uneven length, capacity smaller than the run count: idx is clamped, count tells
    g = torch.Generator().manual_seed(3)
    lens = torch.randint(1, 400, (300,), generator=g)
    img = torch.repeat_interleave(torch.arange(300), lens)
    feats = torch.randn(300, 64, generator=g)
    rows = feats[img].cuda()
    idx, first, count = engine.dedup_rows(rows, 300)
    assert torch.equal(idx.cpu().long(), img) and int(count) == 300
    assert torch.equal(first.cpu().long(), torch.cumsum(lens, 0) - lens)
    idx, first, count = engine.dedup_rows(rows, 100)
    assert int(count) == 300 and int(idx.max()) == 99


def test_drop_in_call_with_repeated_features():
    """flow(rotation[N,3,3], feature.repeat(...)[N,F]) -- the literal calls of eval.py:450-453 / agent.py:240-261 -- equals the
    feature_index form, without a host synchronisation for N <= the optimistic capacity, is CUDA-graph capturable, and reads the
    run count back only when N exceeds it."""
    from rotationnormflow_b200 import engine
    g = golden("s_symsol")
    m = _product(g)
    F = g.feat.shape[1]
    feat = g.feat.cuda()
    B = feat.shape[0]
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for N in (B * 50, engine.DEDUP_CAP + 1000):                      # below / above the capacity
            R = orc.random_rotations(N, gen).cuda()
            per = (N + B - 1) // B
            idx = (torch.arange(N, device="cuda") // per).to(torch.int32)
            rows = feat[idx.long()]                                        # materialised, as .repeat does
            want = m(R, feat, feature_index=idx)
            got = m(R, rows)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
            gi = m.inverse(R, rows)
            wi = m.inverse(R, feat, feature_index=idx)
            assert torch.equal(gi[0], wi[0]) and torch.equal(gi[1], wi[1])
            one = m(R, feat[:1].repeat(N, 1))                              # eval.py:450: one image repeated
            ref = m(R, feat[:1].expand(N, F))
            assert torch.equal(one[0], ref[0]) and torch.equal(one[1], ref[1])
        # capture: no host synchronisation anywhere on the path
        N = 3000
        R = orc.random_rotations(N, gen).cuda()
        rows = feat[(torch.arange(N, device="cuda") // 1000).long()].clone()
        eager = m(R, rows)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            m(R, rows)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            out = m(R, rows)
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out[0], eager[0]) and torch.equal(out[1], eager[1])
        rows.copy_(feat[(torch.arange(N, device="cuda") // 1500).long()])   # different run structure, same graph
        graph.replay()
        torch.cuda.synchronize()
        fresh = m(R, rows)
        assert torch.equal(out[0], fresh[0]) and torch.equal(out[1], fresh[1])


def test_autograd_dispatch_and_cache_hygiene():
    """no silently detached outputs in grad mode -- with autograd on and something requiring grad the call runs in the
    differentiable per-layer operators (tests/test_gpu_train.py); the packed-program cache neither blocks deepcopy / pickle nor
    survives invalidate_cache()."""
    import copy
    import pickle
    g = golden("s_symsol")
    m = _product(g)
    R, rows = g.R.cuda(), g.rows.cuda()
    Rg, lg = m(R, rows)                                                    # parameters require grad, autograd is on
    assert lg.requires_grad and lg.grad_fn is not None
    with torch.no_grad():
        a = m(R, rows)
    assert not a[1].requires_grad and (a[1] - lg.detach()).abs().max() < 1e-4
    fr = rows.clone().requires_grad_(True)
    for p in m.parameters():
        p.requires_grad_(False)
    assert m(R, fr)[1].requires_grad                                       # a backbone feature that requires grad is enough
    assert not m(R, rows)[1].requires_grad                                 # nothing requires grad: the fused kernels
    with pytest.raises(NotImplementedError):
        m.grid_log_prob(rgrid.healpix_grid(0), g.feat.cuda().requires_grad_(True))
    m2 = copy.deepcopy(m)                                                  # after the first call: the cache holds a ctypes handle
    m3 = pickle.loads(pickle.dumps(m))
    with torch.no_grad():
        assert torch.equal(m2(R, rows)[1], a[1]) and torch.equal(m3(R, rows)[1], a[1])
        # a write through .data does not bump the version counter: invalidate_cache() is the documented remedy
        m.layers[1].conditioner.fc_last.bias.data.add_(0.25)
        assert torch.equal(m(R, rows)[1], a[1])
        m.invalidate_cache()
        assert (m(R, rows)[1] - a[1]).abs().max() > 1e-4


def test_rotation_layers_svd_backend(monkeypatch):
    """16Rot / 16UnRot: U^T V is defined up to the sign convention of the SVD routine (D U^T V D).  Default backend = torch.svd on
    the CUDA tensor (what the reference calls on a GPU); RNF_SVD_BACKEND=cpu = LAPACK, the convention of the CPU-minted goldens."""
    from rotationnormflow_b200 import engine
    g = golden("s_rotc")
    m = _product(g)
    prog_of = lambda: __import__("rotationnormflow_b200.flow", fromlist=["_program"])._program(m, list(m.layers), m._perm_rows(), m.feature_dim, torch.device("cuda", 0))
    mats = {}
    for backend in ("cpu", "device"):
        monkeypatch.setenv("RNF_SVD_BACKEND", backend)
        prog = prog_of()
        cond = prog.condition(g.feat.cuda())
        base = prog.n_mob * 64
        mats[backend] = cond[:, base: base + prog.n_aff * 40].reshape(-1, prog.n_aff, 40)[:, :, :16].reshape(-1, 4, 4).cpu().double()
    for W in mats.values():                                                # 4-D rotations either way
        assert (W @ W.transpose(1, 2) - torch.eye(4, dtype=torch.float64)).abs().max() < 1e-5
    # the two conventions differ by a diagonal sign matrix on both sides at most: |entries| agree
    assert (mats["cpu"].abs() - mats["device"].abs()).abs().max() < 1e-4
    monkeypatch.setenv("RNF_SVD_BACKEND", "device")
    with torch.no_grad():
        R, ldj = m(g.R.cuda(), g.rows.cuda())
        Ri, li = m.inverse(R, g.rows.cuda())
    assert (Ri.cpu() - g.R).abs().max() < 2e-3 and (li + ldj).abs().max() < 2e-3
