"""TEST INFRASTRUCTURE ONLY.  Mints ``tests/golden/*.npz`` by running the UNMODIFIED reference
(/root/reference, imported through ``oracle/ref_loader.py`` with the stub modules of
``oracle/stubs.py``) on seeded inputs.  Run in the build container only:

    python -m oracle.make_golden

Each flow case stores: the config (json), the seed used before ``get_flow``, the inputs, the
reference outputs in fp32 ("what the reference computes") and fp64 ("what it means"), for the
forward and the inverse direction, and either the full state dict (small cases) or a checksum
of it (full-size cases, whose weights the tests regenerate from the same seed through the
product module, which mirrors the reference's parameter-creation order).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

from . import ref_loader as rl
from . import rnf_oracle as orc

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sd_checksum(sd) -> float:
    tot = 0.0
    for i, (k, v) in enumerate(sorted(sd.items())):
        tot += float(v.double().abs().sum()) * (1 + (i % 7))
    return tot


def flow_case(tag, cfg_name, seed, n, n_img, store_weights, with_fisher=False, **ov):
    cfg = rl.ref_config(cfg_name, **ov)
    F = orc.feature_dim_of(cfg)
    g = torch.Generator().manual_seed(1000 + seed)
    R = orc.random_rotations(n, g, torch.float32)
    feat = idx = None
    if F:
        feat = torch.relu(torch.randn(n_img, F, generator=g)).float()
        idx = torch.arange(n, dtype=torch.int64) * n_img // n          # image-major blocks
    out = dict(cfg=json.dumps(vars(cfg)), seed=seed, R=R.numpy(),
               torch_version=torch.__version__)
    if F:
        out["feat"] = feat.numpy()
        out["feat_index"] = idx.numpy().astype(np.int32)
    m32 = rl.build_reference_flow(cfg, seed, torch.float32)
    sd = {k: v.clone() for k, v in m32.state_dict().items()}
    m64 = rl.build_reference_flow(cfg, seed, torch.float32).double()     # same fp32 weights, fp64 math
    rows = None if feat is None else feat[idx]
    for name, m in (("f32", m32), ("f64", m64)):
        for inv in (False, True):
            Ro, lo = rl.run_reference(m, R, rows, inverse=inv)
            d = "inv" if inv else "fwd"
            out[f"{d}_R_{name}"] = Ro.numpy()
            out[f"{d}_ldj_{name}"] = lo.numpy()
    if with_fisher:
        fisher = rl.reference_fisher()
        A = torch.randn(n_img, 3, 3, generator=g) * 3.0
        out["fisher_A"] = A.numpy()
        base = torch.from_numpy(out["fwd_R_f32"])
        out["fisher_logp_f32"] = fisher.MatrixFisherN(A.clone())._log_prob(base).numpy()
        out["fisher_logp_f64"] = fisher.MatrixFisherN(A.double())._log_prob(
            torch.from_numpy(out["fwd_R_f64"])).numpy()
    out["sd_checksum"] = sd_checksum(sd)
    out["sd_keys"] = json.dumps([[k, list(v.shape)] for k, v in sd.items()])
    if store_weights:
        for k, v in sd.items():
            out["sd::" + k] = v.numpy()
    path = os.path.join(OUT, f"flow_{tag}.npz")
    np.savez_compressed(path, **out)
    print(f"{tag:14s} layers={len(orc.layer_plan(cfg)):2d} F={F:4d} n={n} -> {os.path.getsize(path)/1024:.0f} KiB")


def grid_case():
    sd = rl.reference_sd()
    # utils/sd.py:77-79 calls Rotation.from_euler("X", 1-D array); scipy >= 1.13 wants (N,1).  The
    # shim below only reshapes the argument -- the arithmetic is scipy's.
    from scipy.spatial.transform import Rotation
    orig = Rotation.from_euler

    def from_euler(seq, angles, degrees=False):
        a = np.asarray(angles)
        if a.ndim == 1 and len(seq) == 1:
            a = a[:, None]
        return orig(seq, a, degrees=degrees)

    sd.Rotation = type("R", (), {"from_euler": staticmethod(from_euler)})
    out = {}
    rng = np.random.default_rng(0)
    for level in range(0, 5):
        G = sd.generate_healpix_grid(recursion_level=level).numpy()
        assert G.shape[0] == 72 * 8 ** level
        if level <= 2:
            out[f"full_{level}"] = G
        else:
            idx = np.sort(rng.choice(G.shape[0], 3000, replace=False)).astype(np.int64)
            out[f"idx_{level}"] = idx
            out[f"sample_{level}"] = G[idx]
    for q in (72, 500, 5000, 4096, 40000, 300000, 2_000_000, 2_400_000, 10_000_000, 37_000_000):
        gs = 72 * 8 ** np.arange(9)
        out[f"closest_{q}"] = int(gs[np.argmin(np.abs(np.log(q) - np.log(gs)))])   # utils/sd.py:32-34
    path = os.path.join(OUT, "healpix_grid.npz")
    np.savez_compressed(path, **out)
    print(f"grid -> {os.path.getsize(path)/1024:.0f} KiB")


def small_cases():
    flow_case("s_uncond", "raw", 3, 256, 0, True, layers=3)
    flow_case("s_symsol", "symsol", 4, 256, 4, True, layers=2, feature_dim=24)
    flow_case("s_modelnet", "modelnet_fisher", 5, 256, 4, True, with_fisher=True, layers=2, feature_dim=16, embedding_dim=8)
    flow_case("s_pascal", "pascal_uni", 6, 128, 2, True, layers=1, feature_dim=16, embedding_dim=8)
    flow_case("s_lu", "raw", 7, 128, 0, True, layers=2, lu=1)
    flow_case("s_rot", "raw", 8, 128, 0, True, layers=2, rot="16Rot")
    flow_case("s_rotc", "modelnet_uni", 9, 128, 2, True, layers=1, rot="16Rot", feature_dim=8, embedding=0)
    flow_case("s_unrot", "symsol", 10, 128, 2, True, layers=2, rot="16UnRot", feature_dim=8)
    flow_case("s_mobonly", "raw", 11, 128, 0, True, layers=2, rot="None")


def clu_cases():
    """Condition16TransLU (flow/squeezetrans.py:94-144): batch-coupled in the reference; golden = what the reference computes."""
    flow_case("s_clu", "modelnet_uni", 23, 64, 2, True, layers=2, lu=1, feature_dim=8, embedding=0)


def ablation_cases():
    """Ablation replacements of the affine layer (flow/affineflow.py:27-41,55-70), unconditional and conditional."""
    flow_case("s_smith9", "raw", 12, 128, 0, True, layers=2, rot="9TransLSmith")
    flow_case("s_smith9lu", "raw", 13, 128, 0, True, layers=2, rot="9TransLSmith", lu=1)
    flow_case("s_smith36", "raw", 14, 128, 0, True, layers=2, rot="36Trans")
    flow_case("s_polar9l", "raw", 15, 128, 0, True, layers=2, rot="9TransLSVD")
    flow_case("s_polar9r", "raw", 16, 128, 0, True, layers=2, rot="9TransRSVD")
    flow_case("s_right9", "raw", 17, 128, 0, True, layers=2, rot="9TransRSmith")
    flow_case("s_smith9c", "modelnet_uni", 18, 128, 2, True, layers=2, rot="9TransLSmith", feature_dim=8, embedding=0)
    flow_case("s_smith36c", "modelnet_uni", 19, 128, 2, True, layers=2, rot="36Trans", feature_dim=8, embedding=0)
    flow_case("s_polar9lc", "modelnet_uni", 20, 128, 2, True, layers=2, rot="9TransLSVD", feature_dim=8, embedding=0)
    flow_case("s_polar9rc", "modelnet_uni", 21, 128, 2, True, layers=2, rot="9TransRSVD", feature_dim=8, embedding=0)
    flow_case("s_right9c", "modelnet_uni", 22, 128, 2, True, layers=2, rot="9TransRSmith", feature_dim=8, embedding=0)


def full_cases():
    """BASELINE.json configs 1-4: weights regenerated from the seed."""
    flow_case("raw", "raw", 0, 512, 0, False)
    flow_case("symsol2048", "symsol", 0, 512, 4, False, feature_dim=2048)
    flow_case("symsol2", "symsol2", 0, 256, 4, False)
    flow_case("modelnet", "modelnet_fisher", 0, 256, 4, False, with_fisher=True)


GROUPS = {"small": small_cases, "clu": clu_cases, "ablation": ablation_cases, "full": full_cases, "grid": grid_case}


def main():
    """`python -m oracle.make_golden [group ...]`: all groups by default (small, ablation, full, grid)."""
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name in (sys.argv[1:] or list(GROUPS)):
        GROUPS[name]()


if __name__ == "__main__":
    sys.exit(main())
