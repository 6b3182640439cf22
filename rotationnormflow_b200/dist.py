"""Grid-sharded evaluation across the GPUs of one box: one process per GPU, ``torch.distributed`` plumbing.

The reference evaluates the grid on a single GPU in 500 000-rotation chunks (eval.py:444-462) and only ever uses
``nn.DataParallel`` (agent.py:22).  Here every rank scores a contiguous slice ``[begin,end)`` of the grid for all B
images with the fused kernel and keeps, per image, (max log p, first arg-max index, sum exp(log p - max)); ONE
all-gather of ``[B,3]`` float64 (24 B per image per rank) merges them.  There is no other data-path collective: the
per-(rotation, image) evaluations are independent.
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist


def shard_range(G: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of ``range(G)`` owned by ``rank`` (sizes differ by at most one, earlier ranks larger)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(G, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def merge_partials(mx: torch.Tensor, am: torch.Tensor, se: torch.Tensor, sn: torch.Tensor | None = None):
    """Merge per-shard partials stacked on dim 0: mx/am/se are [W,B] -> (max [B], argmax [B], sumexp [B]).

    Ties on the maximum resolve to the smallest global index (torch.argmax / first-index semantics of
    agent.py:264, eval.py:461).  Shards holding no rotations carry max = -inf and sumexp = 0.
    ``sn`` [W,B] (spread numerators, relative to each shard's max) is rescaled and summed like sumexp and returned fourth."""
    m = mx.max(dim=0).values
    cand = torch.where(mx == m[None, :], am, torch.full_like(am, torch.iinfo(torch.int64).max))
    idx = cand.min(dim=0).values
    scale = torch.where(torch.isfinite(mx), torch.exp(mx.double() - m.double()[None, :]), torch.zeros_like(mx, dtype=torch.float64))
    s = (se.double() * scale).sum(dim=0)
    if sn is not None:
        return m, idx, s.to(se.dtype), (sn.double() * scale).sum(dim=0).to(sn.dtype)
    return m, idx, s.to(se.dtype)


def all_merge(mx: torch.Tensor, am: torch.Tensor, se: torch.Tensor, group=None, sn: torch.Tensor | None = None):
    """One all-gather of the packed partials ([B,3] float64, [B,4] with spread numerators), then ``merge_partials`` on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return (mx, am, se) if sn is None else (mx, am, se, sn)
    world = dist.get_world_size(group)
    cols = [mx.double(), am.double(), se.double()] + ([] if sn is None else [sn.double()])   # indices < 2^53 are exact
    packed = torch.stack(cols, dim=1).contiguous()
    flat = torch.empty((world * packed.shape[0], len(cols)), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(flat, packed, group=group)                                  # rank-major concatenation
    gathered = flat.view(world, packed.shape[0], len(cols))
    return merge_partials(gathered[:, :, 0].to(mx.dtype), gathered[:, :, 1].to(torch.int64), gathered[:, :, 2].to(se.dtype),
                          None if sn is None else gathered[:, :, 3].to(sn.dtype))


def log_normaliser(mx: torch.Tensor, se: torch.Tensor, G_total: int) -> torch.Tensor:
    """log mean_g exp(logp)  (the ``exp(logp).mean()`` sanity value of eval.py:103-104), from (max, sumexp)."""
    return mx + torch.log(se) - math.log(G_total)


def sharded_grid_log_prob(flow, grid_shard: torch.Tensor, g_index0: int, G_total: int, feature=None, offset=None,
                          fisher_A=None, group=None, mlp_mode=None, gt_rotations=None):
    """Per-image (max, argmax, log-normaliser[, spread]) over a grid whose slices live on different ranks."""
    out = flow.grid_log_prob(grid_shard, feature, offset=offset, fisher_A=fisher_A, g_index0=g_index0, mlp_mode=mlp_mode,
                             gt_rotations=gt_rotations)
    if gt_rotations is not None:
        mx, am, se, sn = all_merge(out["max"], out["argmax"], out["sumexp"], group, sn=out["spread_num"])
        return dict(max=mx, argmax=am, sumexp=se, log_norm=log_normaliser(mx, se, G_total), spread=sn / se)
    mx, am, se = all_merge(out["max"], out["argmax"], out["sumexp"], group)
    return dict(max=mx, argmax=am, sumexp=se, log_norm=log_normaliser(mx, se, G_total))
