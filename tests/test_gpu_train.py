"""The differentiable path (rotationnormflow_b200/train.py + csrc/train_ops.cu): what Flow.forward / Flow.inverse run when autograd is
on -- training (agent.py:87 loss.backward()), the nll_grad evaluation (eval.py:468-477, through BinFind.backward,
flow/mobiusflow.py:248-273) -- and for config.segments != 64.  Checked against autograd through the oracle in fp64, which
tests/test_oracle.py::test_oracle_gradients_against_live_reference pins to the unmodified reference at 1e-8."""
import numpy as np
import pytest
import torch

from conftest import golden, seeded_product_flow
from oracle import rnf_oracle as orc
import rotationnormflow_b200 as rnf

pytestmark = pytest.mark.gpu


def _vee(R, G):
    H = R.transpose(1, 2) @ G
    return torch.stack([H[:, 2, 1] - H[:, 1, 2], H[:, 0, 2] - H[:, 2, 0], H[:, 1, 0] - H[:, 0, 1]], dim=1)


def _loss(R, ldj, A, c):
    return (A * R).sum() + (c * ldj).sum()


def _compare(cfg, sd, R0, feat_rows, inverse, tol=2e-3):
    """Parameter / feature / (tangential) rotation gradients of the product against the fp64 oracle; returns the worst relative error."""
    gen = torch.Generator().manual_seed(77)
    N = R0.shape[0]
    A = torch.randn(N, 3, 3, generator=gen, dtype=torch.float64)
    c = torch.randn(N, generator=gen, dtype=torch.float64)
    m = rnf.get_flow(cfg)
    m.load_state_dict(sd)
    m = m.cuda().train()
    R = R0.clone().cuda().requires_grad_(True)
    f = None if feat_rows is None else feat_rows.clone().cuda().requires_grad_(True)
    out, ldj = (m.inverse if inverse else m)(R, f)
    assert out.requires_grad and ldj.requires_grad
    _loss(out, ldj, A.cuda().float(), c.cuda().float()).backward()
    o = orc.OracleFlow(cfg, sd, torch.float64)
    names = [k for k, _ in m.named_parameters()]
    for k in names:
        o.sd[k] = o.sd[k].clone().requires_grad_(True)
    R2 = R0.double().clone().requires_grad_(True)
    f2 = None if feat_rows is None else feat_rows.double().clone().requires_grad_(True)
    Ro, lo = o.with_grad(R2, f2, inverse=inverse)
    _loss(Ro, lo, A, c).backward()
    # values first (fp32 vs fp64; the inverse direction returns a dyadic angle: allow the flip bound on a few rows)
    dR = (out.detach().cpu().double() - Ro.detach()).abs().amax(dim=(1, 2))
    assert (dR <= 1e-5).float().mean() > (0.9 if inverse else 0.999) and dR.max() < 2e-3
    worst = 0.0
    got = dict(m.named_parameters())
    for k in names:
        ref = o.sd[k].grad
        if ref is None:
            continue
        g = got[k].grad
        assert g is not None, f"no gradient reached {k}"
        err = (g.cpu().double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
        worst = max(worst, err)
        assert err < tol, (k, err)
    if f is not None:
        err = (f.grad.cpu().double() - f2.grad).abs().max().item() / max(f2.grad.abs().max().item(), 1e-6)
        worst = max(worst, err)
        assert err < tol, ("feature", err)
    tR = _vee(R0.double(), R.grad.cpu().double())
    tref = _vee(R0.double(), R2.grad)
    err = (tR - tref).abs().max().item() / max(tref.abs().max().item(), 1e-6)
    assert err < tol, ("rotation (tangential)", err)
    return max(worst, err)


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("tag", ["s_uncond", "s_symsol", "s_modelnet", "s_pascal", "s_lu", "s_mobonly"])
def test_gradients_match_oracle_autograd(tag, inverse):
    # (the 4-D rotation layers 16Rot / 16UnRot are left out: U^T V of torch.svd(I + 1e-3 noise) has nearly degenerate singular
    #  values, so the reference's own fp32 and fp64 evaluations differ by 7e-3 -- see TRUTH in test_gpu_parity.py -- and its svd
    #  backward divides by the gaps; their differentiable path is the same torch.svd autograd the reference uses)
    g = golden(tag)
    n = 96
    worst = _compare(g.cfg, g.state_dict(), g.R[:n], None if g.rows is None else g.rows[:n], inverse)
    print(f"\n[{tag} {'inverse' if inverse else 'forward'}] worst relative gradient error {worst:.2e}")


def test_composed_path_equals_fused_path():
    """Same numbers from the per-layer operators and from the fused kernels (forward; the inverse up to bisection flips)."""
    g = golden("s_symsol")
    m = rnf.get_flow(g.cfg)
    m.load_state_dict(g.state_dict())
    m = m.cuda()
    R, rows = g.R.cuda(), g.rows.cuda()
    with torch.no_grad():
        Rf, lf = m(R, rows)
    Rc, lc = m(R, rows)                                   # autograd on, parameters require grad -> composed path
    assert lc.requires_grad and not lf.requires_grad
    assert (Rc.detach() - Rf).abs().max() < 1e-5 and (lc.detach() - lf).abs().max() < 1e-4
    assert (Rc.detach().cpu().double() - g.out("fwd", "R", "f64").double()).abs().max() < 1e-5


@pytest.mark.parametrize("K", [16, 32, 100])
def test_other_segment_counts(K):
    """config.segments != 64 (flow/mobiusflow.py:7-14 passes it through unrestricted): per-layer operators, values and gradients."""
    cfg = rnf.load_config("symsol", layers=2, feature_dim=12, segments=K)
    m = seeded_product_flow(cfg, 5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(K)
    R = orc.random_rotations(64, gen)
    rows = torch.relu(torch.randn(4, 12, generator=gen))[torch.arange(64) // 16]
    o = orc.OracleFlow(cfg, sd, torch.float64)
    mc = m.cuda().eval()
    with torch.no_grad():
        for inv in (False, True):
            Rp, lp = (mc.inverse if inv else mc)(R.cuda(), rows.cuda())
            Ro, lo = (o.inverse if inv else o.forward)(R, rows)
            d = (Rp.cpu().double() - Ro).abs().amax(dim=(1, 2))
            assert (d <= 1e-5).float().mean() > (0.9 if inv else 0.999) and d.max() < 1e-3
            if not inv:
                assert (lp.cpu().double() - lo).abs().max() < 1e-4
    _compare(cfg, sd, R, rows, False)


def test_training_step_reduces_the_loss():
    """A few optimiser steps on the NLL of a fixed batch (agent.py:60-87): the loss goes down, every parameter moves."""
    cfg = rnf.load_config("symsol", layers=3, feature_dim=16)
    m = seeded_product_flow(cfg, 9).cuda().train()
    gen = torch.Generator().manual_seed(10)
    R = orc.random_rotations(128, gen).cuda()
    feat = torch.relu(torch.randn(128, 16, generator=gen)).cuda()
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    losses = []
    for _ in range(12):
        _, ldj = m(R, feat)
        loss = -ldj.mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] - 1e-3, losses
    assert all((p.detach() - before[k]).abs().max() > 0 for k, p in m.named_parameters())
    with torch.no_grad():                                 # the fused inference path sees the trained weights (cache keyed on versions)
        _, l2 = m(R, feat)
    assert abs(float(-l2.mean()) - float(-m(R, feat)[1].mean())) < 1e-4
