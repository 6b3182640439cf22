"""CPU-side tests of the product package: state-dict contract, seeded init, packing, C-ABI surface, host logic.
No compute call is made here (there is no GPU in the build container and the package has no CPU path)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import FULL_CASES, ROOT, SMALL_CASES, golden, sd_checksum, seeded_product_flow
from oracle import rnf_oracle as orc
import rotationnormflow_b200 as rnf
from rotationnormflow_b200 import _cabi, engine
from rotationnormflow_b200 import dist as rdist


@pytest.mark.parametrize("tag", SMALL_CASES + FULL_CASES)
def test_state_dict_contract_and_seeded_init(tag):
    """Same keys and shapes as the reference's flow_state_dict (agent.py:111-153) and, because parameters are created in
    the reference's order, the same seeded initial values."""
    g = golden(tag)
    m = seeded_product_flow(g.cfg, g.seed)
    sd = m.state_dict()
    assert sorted((k, tuple(v.shape)) for k, v in sd.items()) == sorted(g.sd_keys)
    assert abs(sd_checksum(sd) - g.sd_checksum) <= 1e-9 * g.sd_checksum
    stored = g.stored_state_dict()
    if stored is not None:
        for k, v in stored.items():
            assert torch.equal(sd[k], v), k
        m.load_state_dict(stored)


def test_layer_stack_matches_oracle_plan():
    for name, ov in (("raw", {}), ("symsol", {}), ("modelnet_fisher", {}), ("pascal_uni", {}), ("raw", dict(lu=1)),
                     ("raw", dict(rot="16Rot")), ("symsol", dict(rot="16UnRot")), ("raw", dict(rot="None")),
                     ("raw", dict(dist="noflow"))):
        cfg = rnf.load_config(name, **ov)
        m = seeded_product_flow(cfg, 0)
        assert [l.kind for l in m.layers] == orc.layer_plan(cfg)
        rows = orc.permute_rows(cfg, orc.layer_plan(cfg))
        for i, l in enumerate(m.layers):
            if l.kind == "mobius":
                assert m._perm_rows()[i] == rows[i] % 3


def test_out_of_scope_layers_raise():
    # Condition9TransLU: calculate_9 with the batch-coupled ConditionLU(3) (torch.diag of a [N,3] tensor) is the one layer left out
    with pytest.raises(NotImplementedError):
        rnf.get_flow(rnf.load_config("modelnet_uni", rot="9TransLSmith", lu=1))
    assert rnf.get_flow(rnf.load_config("modelnet_uni", lu=1, layers=1)).layers[1].kind == "aff_clu"
    # config.segments != 64 is accepted (it runs in the per-layer operators of train.py, tests/test_gpu_train.py)
    assert rnf.get_flow(rnf.load_config("raw", segments=32, layers=1)).layers[0].K == 32


def test_config_defaults_follow_reference():
    c = rnf.load_config("symsol")
    assert (c.layers, c.rot, c.frequent_permute, c.last_affine, c.first_affine, c.feature_dim) == (21, "16UnTrans", 1, 1, 0, 512)
    c = rnf.load_config("modelnet_fisher")
    assert (c.condition, c.feature_dim, c.embedding, c.embedding_dim, c.pretrain_fisher) == (1, 2048, 1, 32, 1)
    c = rnf.load_config("raw")
    assert (c.condition, c.layers, c.segments, c.rot) == (0, 24, 64, "16Trans")


def _emulate_packed_mlp(blk, y, c_img):
    """NumPy walk over the packed kernel image exactly as csrc/flow_v1.cu indexes it."""
    first = blk[:256].reshape(64, 4)
    h0 = first[:, :3] @ y + first[:, 3] + c_img
    a = np.maximum(h0, 0)
    o = 256
    h = None
    for _ in range(3):
        Wt = blk[o:o + 4096].reshape(64, 64); b = blk[o + 4096:o + 4160]; o += 4160
        h = a @ Wt + b
        a = np.maximum(h, 0)
    a = np.maximum(h0 + h, 0)
    Wl = blk[o:o + 64 * 256].reshape(64, 256); bl = blk[o + 64 * 256:o + 64 * 256 + 256]
    return a @ Wl + bl


def test_mobius_packing_is_the_reference_mlp():
    g = golden("s_symsol")
    sd = {k: v.double() for k, v in g.state_dict().items()}
    pre = "layers.1.conditioner."
    sub = {k[len(pre):]: v for k, v in g.state_dict().items() if k.startswith(pre)}
    F = orc.feature_dim_of(g.cfg)
    blk, wf = engine.pack_mobius(sub, F)
    assert blk.size == engine.MOB_FLOATS and wf.shape == (64, F)
    rng = np.random.default_rng(0)
    y = rng.standard_normal(3)
    feat = rng.standard_normal(F)
    out = _emulate_packed_mlp(blk.astype(np.float64), y, wf.astype(np.float64) @ feat)
    ref = orc.conditioner(sd, pre, torch.from_numpy(np.concatenate([y, feat]))[None])[0].numpy()
    K = 64
    for c in (0, 1, 17, 63):
        assert abs(out[4 * c] - ref[c]) < 1e-12
        assert np.abs(out[4 * c + 1: 4 * c + 4] - ref[K + 3 * c: K + 3 * c + 3]).max() < 1e-12


def _unswizzle_sw128(words: np.ndarray, N: int) -> np.ndarray:
    """Inverse of engine._umma_k_major_sw128: float32 words of a [N x 64] fp16 K-major SWIZZLE_128B tile -> [N, 64] float64."""
    h = words.view(np.float16)
    n = np.arange(N)[:, None]
    k = np.arange(64)[None, :]
    off = (n // 8) * 1024 + (n % 8) * 128 + (((k // 8) ^ (n % 8)) * 16) + (k % 8) * 2
    return h[off // 2].astype(np.float64)


def _unblock(words: np.ndarray, N: int) -> np.ndarray:
    """[N x 16] no-swizzle K-major block (engine._bias_block) -> [N, 16] float64 slot values."""
    h = words.view(np.float16)
    n = np.arange(N)[:, None]
    k = np.arange(16)[None, :]
    off = (n // 8) * 256 + (k // 8) * 128 + (n % 8) * 16 + (k % 8) * 2
    return h[off // 2].astype(np.float64)


def test_tensor_core_image_is_the_reference_mlp():
    """Emulates what csrc/flow_t4.cu does with the packed image (engine.pack_mobius_tc): every GEMM = Y block x layer block
    (bias, fc_first on y, residual x0) + activations x (hi + lo planes); fc_last in the pair column layout with the logits
    scaled by log2(e).  Checked against the float64 oracle conditioner (flow/condition.py:24-30)."""
    g = golden("s_symsol")
    sd = {k: v.double() for k, v in g.state_dict().items()}
    pre = "layers.1.conditioner."
    sub = {k[len(pre):]: v for k, v in g.state_dict().items() if k.startswith(pre)}
    F = orc.feature_dim_of(g.cfg)
    img = engine.pack_mobius_tc(sub)
    assert img.size == engine.MOB_TC_FLOATS
    _, wf = engine.pack_mobius(sub, F)
    rng = np.random.default_rng(3)
    y = rng.standard_normal(3)
    y /= np.linalg.norm(y)
    feat = rng.standard_normal(F)
    c = wf.astype(np.float64) @ feat
    yh = y.astype(np.float16).astype(np.float64)
    yl = (y - yh).astype(np.float16).astype(np.float64)
    Y = np.concatenate([[1, 1], yh, yl, yh, [1, 1], [0, 0, 0]])        # the kernel's per-rotation block
    o = 0
    hid = []
    for _ in range(3):
        W = _unswizzle_sw128(img[o:o + 2048], 64) + _unswizzle_sw128(img[o + 2048:o + 4096], 64)
        blk = _unblock(img[o + 4096:o + 4608], 64)
        hid.append((W, blk))
        o += 4608
    W4 = _unswizzle_sw128(img[o:o + 8192], 256) + _unswizzle_sw128(img[o + 8192:o + 16384], 256)
    blk4 = _unblock(img[o + 16384:o + 18432], 256)
    o += 18432
    blk0 = _unblock(img[o + 256:o + 256 + 512], 64)
    assert o + 256 + 512 == img.size
    x0 = blk0 @ Y + c
    x = np.maximum(x0, 0)
    for l, (W, blk) in enumerate(hid):
        x = W @ x + blk @ Y + (c if l == 2 else 0)                        # last hidden block re-creates x0 (residual)
        x = np.maximum(x, 0)
    out = W4 @ x + blk4 @ Y
    ref = orc.conditioner(sd, pre, torch.from_numpy(np.concatenate([y, feat]))[None])[0].numpy()
    K = 64
    scale = np.abs(ref).max()
    for comp in (0, 1, 2, 17, 62, 63):
        base = 8 * (comp // 2) + (comp % 2)
        assert abs(out[base] - ref[comp] * engine.LOG2E) < 2e-6 * scale
        got_w = out[[base + 2, base + 4, base + 6]]
        assert np.abs(got_w - engine.CENTRE_SCALE * ref[K + 3 * comp: K + 3 * comp + 3]).max() < 2e-6 * scale
    perm = engine._last_layer_perm_pairs(K)
    assert sorted(perm.tolist()) == list(range(4 * K))


def test_affine_packing():
    W = torch.eye(4) + 0.1 * torch.randn(4, 4, generator=torch.Generator().manual_seed(0))
    blk = engine.pack_affine_matrix(W[None], is_rot=False)
    assert np.allclose(blk[:16].reshape(4, 4), W.numpy())
    inv = engine.AFF_INV
    assert blk.size == engine.AFF_FLOATS
    assert np.allclose(blk[inv:inv + 16].reshape(4, 4) @ W.numpy(), np.eye(4), atol=1e-6)
    assert abs(blk[16] - float(orc.det4(W.double()).abs().log())) < 1e-6
    assert abs(blk[16] + blk[inv + 16]) < 1e-6
    rot = engine.pack_affine_matrix(torch.eye(4)[None], is_rot=True)
    assert rot[16] == 0 and rot[inv + 16] == 0


def test_ablation_packing():
    """Parameter blocks of the ablation layers: forward matrix at 0, inverse-direction matrix at AFF_INV (engine.ablation_blocks)."""
    g = torch.Generator().manual_seed(1)
    M3 = torch.eye(3) + 0.2 * torch.randn(2, 3, 3, generator=g)
    M6 = torch.eye(6) + 0.1 * torch.randn(1, 6, 6, generator=g)
    inv = engine.AFF_INV
    b = engine.ablation_blocks("smith9", M3.double())
    assert torch.allclose(b[:, inv:inv + 9].reshape(2, 3, 3) @ b[:, :9].reshape(2, 3, 3), torch.eye(3).expand(2, 3, 3), atol=1e-6)
    b = engine.ablation_blocks("smith36", M6.double())
    assert torch.allclose(b[:, inv:inv + 36].reshape(1, 6, 6) @ M6, torch.eye(6)[None], atol=1e-6)
    b = engine.ablation_blocks("polar9l", M3)
    assert torch.equal(b[:, inv:inv + 9].reshape(2, 3, 3), M3.transpose(1, 2))
    b = engine.ablation_blocks("right9", M3)
    Q = b[:, :9].reshape(2, 3, 3)
    assert torch.allclose(Q @ Q.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-6) and (torch.linalg.det(Q) - 1).abs().max() < 1e-5
    assert torch.equal(b[:, inv:inv + 9].reshape(2, 3, 3), Q.transpose(1, 2))
    # the restatement of calculate_9_r_smith in the oracle builds the same Q
    R = orc.random_rotations(2, g)
    assert torch.allclose(orc.right9(M3, R, False)[0], R @ Q, atol=1e-6)


def test_cabi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rnf_abi.h")).read()
    declared = set(re.findall(r"\b(rnf_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rnf_flow_condition"} - declared          # (no-op; keeps the set literal honest)
    assert declared == set(_cabi.exported_symbols())
    lib = _cabi.load()                                       # builds with nvcc if stale; loading needs no GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rnf_abi_version() == _cabi.ABI_VERSION
    raw = ctypes.CDLL(_cabi.library_path())
    for name in declared:
        getattr(raw, name)


def test_cabi_argument_errors_without_gpu():
    lib = _cabi.load()
    assert lib.rnf_healpix_grid(9, 0, 1, None, None) == -1
    assert b"level" in lib.rnf_last_error()
    assert lib.rnf_flow_forward(None, None, 1, None, 0, None, 0, None, None, 0, None) == -1
    assert lib.rnf_grid_partial_floats(1000, 2) == ((1000 + 127) // 128) * 2 * 8    # kPartStride floats per tile
    with pytest.raises(_cabi.RnfError):
        _cabi.check(lib.rnf_flow_condition(None, None, 1, None, None))


def test_cpu_tensors_fail_loudly():
    m = seeded_product_flow(rnf.load_config("raw", layers=1), 0)
    R = torch.eye(3)[None]
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(R)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.inverse(R)
    with pytest.raises(ValueError):
        m(torch.zeros(3, 3))


def test_shard_ranges_cover_the_grid():
    for G in (0, 1, 7, 72, 2_359_296):
        for W in (1, 2, 3, 8):
            spans = [rdist.shard_range(G, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_merge_partials_with_spread_numerators():
    """Slices merged with their spread numerators give the spread of the whole grid (oracle definition)."""
    rng = torch.Generator().manual_seed(5)
    B, G = 3, 1000
    logp = torch.randn(B, G, generator=rng) * 3
    samples = orc.random_rotations(G, rng)
    gt = orc.random_rotations(B * 2, rng).reshape(B, 2, 3, 3)
    want = orc.spread(logp, samples, gt)
    prod = torch.einsum("gij,bkij->bgk", samples.double(), gt.double())
    d = torch.acos(torch.clip((prod.max(-1).values - 1.0) / 2.0, -1.0, 1.0))
    cuts = [0, 100, 100, 640, G]                                         # includes an empty slice
    mx, am, se, sn = [], [], [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        if hi == lo:
            mx.append(torch.full((B,), -float("inf"))); am.append(torch.zeros(B, dtype=torch.int64))
            se.append(torch.zeros(B)); sn.append(torch.zeros(B))
            continue
        m = logp[:, lo:hi].max(dim=-1)
        e = torch.exp(logp[:, lo:hi].double() - m.values[:, None].double())
        mx.append(m.values); am.append(m.indices + lo); se.append(e.sum(-1).float()); sn.append((e * d[:, lo:hi]).sum(-1).float())
    m, idx, s, n = rdist.merge_partials(torch.stack(mx), torch.stack(am), torch.stack(se), torch.stack(sn))
    assert torch.equal(idx, torch.argmax(logp, dim=-1))
    assert ((n.double() / s.double()) - want).abs().max() < 1e-6


def test_merge_partials_equals_global_reduction():
    gen = torch.Generator().manual_seed(0)
    B, G, W = 5, 1000, 4
    logp = torch.randn(B, G, generator=gen) * 3
    logp[1, 10] = logp[1, 900] = 50.0           # exact tie across shards -> first index wins
    logp[2] = logp[2, 0]                        # a fully flat image
    mx, am, se = [], [], []
    for r in range(W):
        b, e = rdist.shard_range(G, r, W)
        part = logp[:, b:e]
        m = part.max(1).values
        mx.append(m); am.append(part.argmax(1) + b); se.append(torch.exp(part - m[:, None]).sum(1))
    m, idx, s = rdist.merge_partials(torch.stack(mx), torch.stack(am), torch.stack(se))
    ridx, rmx, rlme = orc.grid_reduce(logp)
    assert torch.equal(idx, ridx) and torch.equal(m, rmx)
    assert (rdist.log_normaliser(m, s, G) - rlme).abs().max() < 1e-5
    # an empty shard contributes nothing
    m2, idx2, s2 = rdist.merge_partials(torch.stack(mx + [torch.full((B,), -float("inf"))]),
                                        torch.stack(am + [torch.zeros(B, dtype=torch.int64)]),
                                        torch.stack(se + [torch.zeros(B)]))
    assert torch.equal(idx2, idx) and torch.allclose(s2, s)


def test_tile_slots_per_launch_host_logic():
    """flow_t4 chooses how many of its four tile slots a launch uses (csrc/flow_t4.cu: pick_active_tiles, a host function): it minimises
    rounds x measured round duration.  Large launches use all four; launches of a few tiles per SM take the count that wastes the least
    of the last round; the choice never exceeds the work there is."""
    import ctypes as C
    from rotationnormflow_b200 import _cabi
    lib = C.CDLL(_cabi.library_path())
    f = lib.rnf_debug_pick_active_tiles
    f.argtypes, f.restype = [C.c_longlong, C.c_int], C.c_int
    sms = 148
    rel = {1: 0.589, 2: 0.772, 3: 0.834, 4: 1.0}                      # round durations, profiles/r02_active_tiles_service.txt
    def cost(n, a):
        groups = -(-n // a)
        return -(-groups // sms) * rel[a]
    for n in [1, 5, 147, 148, 149, 296, 391, 443, 444, 445, 592, 593, 782, 1563, 18432, 147456, 2_359_296]:
        a = f(n, sms)
        assert 1 <= a <= 4
        assert cost(n, a) <= min(cost(n, b) for b in (1, 2, 3, 4)) + 1e-6, (n, a)
    assert f(147456, sms) == 4 and f(18432, sms) == 4                    # bench workloads: all four slots
    assert f(782, sms) == 3                                              # config 1 (100 000 rows): two rounds of three
    assert f(100, sms) == 1                                              # fewer tiles than SMs: one slot per CTA is the shortest round
