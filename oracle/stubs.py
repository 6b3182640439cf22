"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Stand-ins for the third-party modules the reference imports but which are absent
from this image (no network): ``pytorch3d.transforms``, ``healpy``, ``nflows``,
``tkinter.ttk``.  They restate the *public* definitions of the four/two
functions the hot path touches; nothing in
``/root/reference`` pins them, so parity at these boundaries is "unpinned" and
anchored only on the public formulas + the known-answer values in
``tests/test_oracle.py``.

Call sites in the reference that these serve:
  flow/squeezetrans.py:34,37   matrix_to_quaternion / quaternion_to_matrix
  flow/rottrans.py:18,20,...   same
  utils/sd.py:23,69,70         random_rotations, hp.nside2npix, hp.pix2vec
  utils/fisher.py:1            nflows.distributions.Distribution (base class only)
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch


# ----------------------------------------------------------------------------
# pytorch3d.transforms (0.7.x public source, restated)
# ----------------------------------------------------------------------------
def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k),
            two_s * (i * j - k * r),
            two_s * (i * k + j * r),
            two_s * (i * j + k * r),
            1 - two_s * (i * i + k * k),
            two_s * (j * k - i * r),
            two_s * (i * k - j * r),
            two_s * (j * k + i * r),
            1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def _sqrt_positive_part(x: torch.Tensor) -> torch.Tensor:
    ret = torch.zeros_like(x)
    pos = x > 0
    ret[pos] = torch.sqrt(x[pos])
    return ret


def matrix_to_quaternion(matrix: torch.Tensor) -> torch.Tensor:
    batch_dim = matrix.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(
        matrix.reshape(batch_dim + (9,)), dim=-1
    )
    q_abs = _sqrt_positive_part(
        torch.stack(
            [
                1.0 + m00 + m11 + m22,
                1.0 + m00 - m11 - m22,
                1.0 - m00 + m11 - m22,
                1.0 - m00 - m11 + m22,
            ],
            dim=-1,
        )
    )
    quat_by_rijk = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
        ],
        dim=-2,
    )
    flr = torch.tensor(0.1).to(dtype=q_abs.dtype, device=q_abs.device)
    quat_candidates = quat_by_rijk / (2.0 * q_abs[..., None].max(flr))
    sel = torch.nn.functional.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return quat_candidates[sel, :].reshape(batch_dim + (4,))


def random_rotations(n: int, dtype=None, device=None) -> torch.Tensor:
    o = torch.randn((n, 4), dtype=dtype, device=device)
    s = (o * o).sum(1)
    o = o / torch.copysign(torch.sqrt(s), o[:, 0])[:, None]
    return quaternion_to_matrix(o)


def random_rotation(dtype=None, device=None) -> torch.Tensor:
    return random_rotations(1, dtype, device)[0]


# ----------------------------------------------------------------------------
# healpy (RING scheme pix2vec / nside2npix, public HEALPix definition)
# ----------------------------------------------------------------------------
def nside2npix(nside: int) -> int:
    return 12 * int(nside) * int(nside)


def pix2zphi(nside: int, ipix) -> tuple[np.ndarray, np.ndarray]:
    """RING-scheme pixel centre as (z=cos(theta), phi); float64."""
    nside = int(nside)
    p = np.asarray(ipix, dtype=np.int64)
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    z = np.empty(p.shape, dtype=np.float64)
    phi = np.empty(p.shape, dtype=np.float64)

    north = p < ncap
    south = p >= npix - ncap
    belt = ~(north | south)

    # north polar cap
    pn = p[north]
    i = (1 + np.floor(np.sqrt(1 + 2 * pn.astype(np.float64))).astype(np.int64)) // 2
    # guard isqrt rounding
    i = np.where(2 * i * (i - 1) > pn, i - 1, i)
    i = np.where(2 * i * (i + 1) <= pn, i + 1, i)
    j = pn + 1 - 2 * i * (i - 1)
    z[north] = 1.0 - (i * i) * (4.0 / npix)
    phi[north] = (j - 0.5) * (np.pi / 2) / i

    # equatorial belt
    pb = p[belt] - ncap
    i = pb // (4 * nside) + nside
    j = pb % (4 * nside) + 1
    fodd = np.where(((i + nside) & 1) == 1, 1.0, 0.5)
    z[belt] = (2 * nside - i) * (2.0 / (3 * nside))
    phi[belt] = (j - fodd) * (np.pi / 2) / nside

    # south polar cap
    ps = npix - p[south]
    i = (1 + np.floor(np.sqrt((2 * ps - 1).astype(np.float64))).astype(np.int64)) // 2
    i = np.where(2 * i * (i - 1) >= ps, i - 1, i)
    i = np.where(2 * i * (i + 1) < ps, i + 1, i)
    j = 4 * i + 1 - (ps - 2 * i * (i - 1))
    z[south] = -1.0 + (i * i) * (4.0 / npix)
    phi[south] = (j - 0.5) * (np.pi / 2) / i
    return z, phi


def pix2vec(nside: int, ipix, nest: bool = False):
    assert not nest, "only RING is used by the reference (utils/sd.py:70)"
    z, phi = pix2zphi(nside, ipix)
    st = np.sqrt((1.0 - z) * (1.0 + z))
    return st * np.cos(phi), st * np.sin(phi), z


# ----------------------------------------------------------------------------
def install() -> None:
    """Register the stub modules in ``sys.modules`` (idempotent)."""
    if "pytorch3d.transforms" not in sys.modules:
        p3d = types.ModuleType("pytorch3d")
        tr = types.ModuleType("pytorch3d.transforms")
        tr.matrix_to_quaternion = matrix_to_quaternion
        tr.quaternion_to_matrix = quaternion_to_matrix
        tr.random_rotations = random_rotations
        tr.random_rotation = random_rotation
        p3d.transforms = tr
        sys.modules["pytorch3d"] = p3d
        sys.modules["pytorch3d.transforms"] = tr
    if "healpy" not in sys.modules:
        hp = types.ModuleType("healpy")
        hp.nside2npix = nside2npix
        hp.pix2vec = pix2vec
        sys.modules["healpy"] = hp
    if "nflows" not in sys.modules:
        nf = types.ModuleType("nflows")
        nfd = types.ModuleType("nflows.distributions")
        nfd.Distribution = torch.nn.Module
        nf.distributions = nfd
        sys.modules["nflows"] = nf
        sys.modules["nflows.distributions"] = nfd
    try:
        import tkinter.ttk  # noqa: F401
    except Exception:
        tk = types.ModuleType("tkinter")
        ttk = types.ModuleType("tkinter.ttk")
        ttk.Sizegrip = object
        tk.ttk = ttk
        sys.modules["tkinter"] = tk
        sys.modules["tkinter.ttk"] = ttk
