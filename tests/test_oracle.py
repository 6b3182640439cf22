"""The oracle (oracle/rnf_oracle.py) against the golden vectors minted from the real reference, and -- when
/root/reference is present (build container only) -- against the reference run live."""
import numpy as np
import pytest
import torch

from conftest import FULL_CASES, GOLDEN, NOT_BIJECTIVE, SMALL_CASES, golden
from oracle import rnf_oracle as orc
from oracle import stubs

# fp64: the restatement and the reference agree to rounding; fp32: both carry the reference's own fp32 noise
TOL64 = 5e-12
TOL32_R, TOL32_LDJ = 4e-5, 8e-5


@pytest.mark.parametrize("tag", SMALL_CASES + ["raw", "symsol2"])
def test_forward_matches_golden(tag):
    g = golden(tag)
    sd = g.state_dict()
    o64 = orc.OracleFlow(g.cfg, sd, torch.float64)
    R, ldj = o64.forward(g.R, g.rows)
    assert (R - g.out("fwd", "R", "f64")).abs().max() < TOL64
    assert (ldj - g.out("fwd", "ldj", "f64")).abs().max() < TOL64
    o32 = orc.OracleFlow(g.cfg, sd, torch.float32)
    R, ldj = o32.forward(g.R, g.rows)
    assert (R - g.out("fwd", "R", "f32")).abs().max() < TOL32_R
    assert (ldj - g.out("fwd", "ldj", "f32")).abs().max() < TOL32_LDJ
    # closed-form log-det == the reference's explicit Jacobian construction
    oe = orc.OracleFlow(g.cfg, sd, torch.float64, explicit_jacobian=True)
    _, ldj_e = oe.forward(g.R, g.rows)
    assert (ldj_e - g.out("fwd", "ldj", "f64")).abs().max() < TOL64


@pytest.mark.parametrize("tag", SMALL_CASES + ["symsol2"])
def test_inverse_matches_golden(tag):
    g = golden(tag)
    o64 = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    R, ldj = o64.inverse(g.R, g.rows)
    # bisection branch decisions are identical in fp64 -> same dyadic angles
    assert (R - g.out("inv", "R", "f64")).abs().max() < 1e-9
    assert (ldj - g.out("inv", "ldj", "f64")).abs().max() < 1e-9
    if tag in NOT_BIJECTIVE:
        return
    # round trip to bisection resolution (pi / 2^15 per Mobius layer) and ldj_inv = -ldj_fwd
    Rf, lf = o64.forward(R, g.rows)
    nmob = sum(k == "mobius" for k in o64.plan)
    assert (Rf - g.R.double()).abs().max() < 2e-4 * max(nmob, 1)
    assert (lf + ldj).abs().max() < 2e-3 * max(nmob, 1)


@pytest.mark.parametrize("tag", ["symsol2048", "modelnet"])
def test_full_size_conditionals(tag):
    g = golden(tag)
    o64 = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    R, ldj = o64.forward(g.R, g.rows)
    assert (R - g.out("fwd", "R", "f64")).abs().max() < 1e-10
    assert (ldj - g.out("fwd", "ldj", "f64")).abs().max() < 1e-10


@pytest.mark.parametrize("tag", ["s_modelnet", "modelnet"])
def test_fisher_matches_golden(tag):
    g = golden(tag)
    A = torch.from_numpy(g.z["fisher_A"])
    lp = orc.fisher_log_prob(A.double(), g.out("fwd", "R", "f64"))
    assert (lp - torch.from_numpy(g.z["fisher_logp_f64"])).abs().max() < 1e-9
    lp32 = orc.fisher_log_prob(A, g.out("fwd", "R", "f32"))
    ref = torch.from_numpy(g.z["fisher_logp_f32"])
    assert ((lp32 - ref).abs() / ref.abs().clamp(min=1)).max() < 1e-5


def test_layer_plans_and_permutations():
    cfg = orc.simple_config()
    plan = orc.layer_plan(cfg)
    assert len(plan) == 48 and plan[0] == "mobius" and plan[1] == "aff_u"
    fwd, inv = orc.permute_rows(cfg, plan), orc.permute_rows(cfg, plan, inverse=True)
    assert fwd[:6] == [0, 1, 1, 2, 2, 3]                 # affine layers ignore their row
    assert [fwd[i] for i in range(0, 48, 2)] == [m % 6 for m in range(24)]
    assert all(fwd[i] == inv[i] for i in range(0, 48, 2))
    sym = orc.simple_config(condition=1, layers=21, rot="16UnTrans", frequent_permute=1, last_affine=1, first_affine=0)
    plan = orc.layer_plan(sym)
    assert len(plan) == 42 and plan[0] == "aff_c" and plan[1] == "mobius" and plan[2] == "aff_u" and plan[-1] == "mobius"
    rows = orc.permute_rows(sym, plan)
    assert rows == [i % 6 for i in range(42)]
    # forward and inverse hand every Mobius layer the same row (flow/flow.py:58-70 vs :78-88)
    inv = orc.permute_rows(sym, plan, inverse=True)
    assert all(rows[i] % 3 == inv[i] % 3 for i, k in enumerate(plan) if k == "mobius")


def test_healpix_known_answers():
    # nside = 1: z = 2/3, 0, -2/3 ; phi = pi/4 + k pi/2, k pi/2, pi/4 + k pi/2
    z, phi = stubs.pix2zphi(1, np.arange(12))
    assert np.allclose(z, [2 / 3] * 4 + [0] * 4 + [-2 / 3] * 4)
    assert np.allclose(phi[:4], np.pi / 4 + np.arange(4) * np.pi / 2)
    assert np.allclose(phi[4:8], np.arange(4) * np.pi / 2)
    # healpy docstring values (nside=16, RING)
    x, y, zz = stubs.pix2vec(16, np.array([1504, 1440, 427]))
    assert np.allclose([x[0], y[0], zz[0]], [0.9987954562051724, 0.049067674327418015, 0.0])
    assert np.allclose(x[1:], [0.99913157, 0.5000534]) and np.allclose(y[1:], [0.0, 0.5000534])
    assert np.allclose(zz[1:], [0.04166667, 0.70703125])
    for nside in (1, 2, 4, 8):
        v = np.stack(stubs.pix2vec(nside, np.arange(12 * nside * nside)), 1)
        assert np.allclose(np.linalg.norm(v, axis=1), 1) and np.abs(v.sum(0)).max() < 1e-9
        assert len(np.unique(np.round(v, 9), axis=0)) == 12 * nside * nside


def test_healpix_grid_matches_reference_golden():
    z = np.load(f"{GOLDEN}/healpix_grid.npz")
    for level in (0, 1, 2):
        G = orc.healpix_grid(level).numpy()
        ref = z[f"full_{level}"]
        assert G.shape == ref.shape
        assert np.abs(G - ref).max() <= 1.2e-7           # <= 1 ulp of fp32 after the final fp64 -> fp32 rounding
        big = np.abs(ref) > 1e-9                         # the rest are cos(pi/2)-type residues of size 1e-16
        assert (G[big] != ref[big]).mean() < 1e-3
    for level in (3, 4):
        idx = z[f"idx_{level}"]
        full = orc.healpix_grid(level).numpy()
        assert np.abs(full[idx] - z[f"sample_{level}"]).max() <= 1.2e-7
        # ranged generation == slicing
        part = orc.healpix_grid(level, int(idx[5]), int(idx[5]) + 100).numpy()
        assert np.array_equal(part, full[int(idx[5]): int(idx[5]) + 100])
    for q in (72, 500, 5000, 4096, 40000, 300000, 2_000_000, 2_400_000, 10_000_000, 37_000_000):
        assert 72 * 8 ** orc.closest_grid_level(q) == int(z[f"closest_{q}"])


def test_grid_properties():
    G = orc.healpix_grid(1).double()
    assert (G @ G.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-6
    assert (torch.linalg.det(G) - 1).abs().max() < 1e-6


def test_quaternion_roundtrip_and_scipy_agreement():
    from scipy.spatial.transform import Rotation
    R = orc.random_rotations(200, torch.Generator().manual_seed(3), torch.float64)
    q = stubs.matrix_to_quaternion(R)
    assert (stubs.quaternion_to_matrix(q) - R).abs().max() < 1e-14
    qs = torch.from_numpy(Rotation.from_matrix(R.numpy()).as_quat())[:, [3, 0, 1, 2]]
    assert torch.minimum((q - qs).abs().amax(1), (q + qs).abs().amax(1)).max() < 1e-12


def test_density_integrates_to_one():
    """The reference's own sanity check: mean exp(ldj) over Haar-uniform rotations ~ 1 (eval_uncondition.py:114-116)."""
    g = golden("s_uncond")
    o = orc.OracleFlow(g.cfg, g.state_dict(), torch.float64)
    R = orc.random_rotations(20000, torch.Generator().manual_seed(5), torch.float64)
    _, ldj = o.forward(R)
    assert abs(float(torch.exp(ldj).mean()) - 1.0) < 2e-2


def test_oracle_against_live_reference():
    from oracle import ref_loader as rl
    if not rl.available():
        pytest.skip("/root/reference not present (GPU box)")
    cfg = rl.ref_config("symsol", layers=2, feature_dim=16)
    m = rl.build_reference_flow(cfg, 21, torch.float64)
    gen = torch.Generator().manual_seed(2)
    R = orc.random_rotations(64, gen, torch.float64)
    feat = torch.randn(64, 16, generator=gen, dtype=torch.float64)
    o = orc.OracleFlow(cfg, m.state_dict(), torch.float64)
    for inv in (False, True):
        Rr, lr = rl.run_reference(m, R, feat, inverse=inv)
        Ro, lo = (o.inverse if inv else o.forward)(R, feat)
        assert (Rr - Ro).abs().max() < 1e-9 and (lr - lo).abs().max() < 1e-9


def _grad_loss(R, ldj, A, c):
    return (A * R).sum() + (c * ldj).sum()


def test_oracle_gradients_against_live_reference():
    """Autograd through the restatement == autograd through the unmodified reference, forward (training, agent.py:87) and inverse
    (BinFind.backward, flow/mobiusflow.py:248-273): parameter, feature and rotation gradients in fp64."""
    from oracle import ref_loader as rl
    if not rl.available():
        pytest.skip("/root/reference not present (GPU box)")
    for cfg_name, ov in (("symsol", dict(layers=2, feature_dim=16)), ("modelnet_uni", dict(layers=2, feature_dim=8, embedding=0)), ("raw", dict(layers=2))):
        cfg = rl.ref_config(cfg_name, **ov)
        m = rl.build_reference_flow(cfg, 33, torch.float64)
        gen = torch.Generator().manual_seed(6)
        N = 24
        F = orc.feature_dim_of(cfg)
        A = torch.randn(N, 3, 3, generator=gen, dtype=torch.float64)
        c = torch.randn(N, generator=gen, dtype=torch.float64)
        for inv in (False, True):
            R = orc.random_rotations(N, gen, torch.float64).requires_grad_(True)
            feat = torch.randn(N, F, generator=gen, dtype=torch.float64).requires_grad_(True) if F else None
            m.zero_grad()
            with rl.default_dtype(torch.float64):
                Rr, lr = (m.inverse if inv else m)(R, feat)
                _grad_loss(Rr, lr, A, c).backward()
            ref = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
            ref_R, ref_f = R.grad.clone(), (None if feat is None else feat.grad.clone())
            o = orc.OracleFlow(cfg, m.state_dict(), torch.float64)
            for k in ref:
                o.sd[k] = o.sd[k].clone().requires_grad_(True)
            R2 = R.detach().clone().requires_grad_(True)
            f2 = None if feat is None else feat.detach().clone().requires_grad_(True)
            Ro, lo = o.with_grad(R2, f2, inverse=inv)
            _grad_loss(Ro, lo, A, c).backward()
            assert ref, "the reference produced no parameter gradients"
            for k, g in ref.items():
                assert (o.sd[k].grad - g).abs().max() <= 1e-8 * max(1.0, g.abs().max().item()), (cfg_name, inv, k)
            assert (R2.grad - ref_R).abs().max() < 1e-8
            if feat is not None:
                assert (f2.grad - ref_f).abs().max() < 1e-8


def test_geodesic_metric_against_live_reference():
    import sys
    from oracle import ref_loader as rl
    if not rl.available():
        pytest.skip("/root/reference not present (GPU box)")
    if rl.REF_ROOT not in sys.path:
        sys.path.insert(0, rl.REF_ROOT)
    import utils.utils as ru                              # std-lib + torch imports only
    gen = torch.Generator().manual_seed(4)
    est = orc.random_rotations(50, gen, torch.float64)
    gt = orc.random_rotations(50 * 5, gen, torch.float64).reshape(50, 5, 3, 3)
    assert torch.equal(orc.min_geodesic_distance_rotmats(est, gt), ru.min_geodesic_distance_rotmats(est, gt))


def test_matrix_fisher_sampler_against_importance_sampling():
    """oracle.sample_matrix_fisher (utils/fisher.py:117-207 restated) draws from p(R) ~ exp(tr(A^T R)): its sample mean of R
    and of tr(A^T R) match self-normalised importance sampling from the uniform distribution."""
    gen = torch.Generator().manual_seed(3)
    U, V = orc.random_rotations(2, gen, torch.float64)
    A = U @ torch.diag(torch.tensor([3.0, 2.0, -1.0], dtype=torch.float64)) @ V.T
    R = orc.sample_matrix_fisher(A, 100000, gen)
    assert (R.transpose(1, 2) @ R - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-12 and (torch.linalg.det(R) - 1).abs().max() < 1e-12
    Ru = orc.random_rotations(400000, gen, torch.float64)
    w = torch.exp((Ru * A).sum((-1, -2)))
    mean_is = (w[:, None, None] * Ru).sum(0) / w.sum()
    tr_is = (w * (Ru * A).sum((-1, -2))).sum() / w.sum()
    assert (R.mean(0) - mean_is).abs().max() < 0.02
    assert abs(float((R * A).sum((-1, -2)).mean() - tr_is)) < 0.05
