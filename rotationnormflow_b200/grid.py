"""HEALPix SO(3) query grids generated on the device (utils/sd.py:11-82)."""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _cabi

GRID_SIZES = tuple(72 * 8 ** l for l in range(9))          # utils/sd.py:32
_grids: dict = {}


def closest_grid_level(num_queries: int) -> int:
    """get_closest_available_grid (utils/sd.py:31-34): the level whose size is nearest in log space."""
    best, arg = None, 0
    for l, s in enumerate(GRID_SIZES):
        d = abs(math.log(num_queries) - math.log(s))
        if best is None or d < best:
            best, arg = d, l
    return arg


def healpix_grid(level: int, begin: int = 0, end: int | None = None, device=None) -> torch.Tensor:
    """Rotations [begin,end) of the level-`level` grid as float32 [n,3,3] on `device` (rnf_healpix_grid)."""
    lib = _cabi.load()
    total = 72 * 8 ** level
    end = total if end is None else end
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("healpix_grid generates the grid on a B200; there is no CPU path")
    out = torch.empty((max(end - begin, 0), 3, 3), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        st = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        _cabi.check(lib.rnf_healpix_grid(int(level), int(begin), int(end), C.c_void_p(out.data_ptr()), st))
    return out


def generate_healpix_grid(recursion_level=None, size=None, device=None) -> torch.Tensor:
    """Signature of utils/sd.py:48 (returns a CUDA tensor instead of a CPU one)."""
    assert not (recursion_level is None and size is None)
    if size:
        recursion_level = max(int(round(math.log(size / 72.0) / math.log(8.0))), 0)
    return healpix_grid(recursion_level, device=device)


def get_closest_available_grid(num_queries: int, device=None) -> torch.Tensor:
    level = closest_grid_level(num_queries)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    key = (level, dev.index)
    if key not in _grids:
        _grids[key] = healpix_grid(level, device=dev)
    return _grids[key]


def generate_queries(number_queries: int, mode: str = "random", device=None) -> torch.Tensor:
    """utils/sd.py:11-25.  'random': Haar-uniform rotations from normalised Gaussian quaternions (the public
    pytorch3d.random_rotations definition); 'grid': the closest HEALPix grid."""
    if mode == "grid":
        return get_closest_available_grid(number_queries, device)
    if mode != "random":
        raise ValueError(mode)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    q = torch.randn((number_queries, 4), device=dev)
    q = q / torch.copysign(q.norm(dim=1), q[:, 0])[:, None]
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return m.reshape(-1, 3, 3)
