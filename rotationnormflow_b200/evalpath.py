"""The two evaluation call sites of the path, as single calls on [B,F] features.

    estimate_rotation_grid      eval.py:437-462  (gradient(): grid log-pdf + arg-max per image)
    estimate_rotation_sampling  agent.py:238-266 (eval_acc: inverse-flow samples + arg-max per image), uniform base

Both avoid the reference's `feature.repeat` (4 GB per 500 000-row chunk at F = 2080, eval.py:450) and its per-image
Python loop; all per-rotation arithmetic runs in the fused kernels.
"""
from __future__ import annotations

import torch

from . import grid as rgrid


def _random_rotation(device, generator=None):
    q = torch.randn(4, generator=generator, device="cpu").to(device)
    q = q / torch.copysign(q.norm(), q[0])
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum()
    return torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j))).reshape(3, 3)


@torch.no_grad()
def estimate_rotation_grid(flow, feature, number_queries, fisher_A=None, offset=None, generator=None, mlp_mode=None,
                           gt_rotations=None):
    """Returns (est_rotation [B,3,3], dict(max, argmax, sumexp[, spread])).  One random right-offset per batch
    (eval.py:439-440).  gt_rotations [B,K,3,3] adds the probability-weighted angular error (spread) of the same pass."""
    dev = feature.device
    grid = rgrid.get_closest_available_grid(number_queries, dev)
    if offset is None:
        offset = _random_rotation(dev, generator)
    out = flow.grid_log_prob(grid, feature, offset=offset, fisher_A=fisher_A, mlp_mode=mlp_mode, gt_rotations=gt_rotations)
    est = grid[out["argmax"]] @ offset.to(dev)
    return est, out


@torch.no_grad()
def estimate_rotation_sampling(flow, feature, number_queries, base_samples=None, base_ll=None, mlp_mode=None,
                               fisher_A=None, seed=None):
    """agent.py:238-266: base rotations pushed through Flow.inverse; arg-max of -ldj + base_ll per image.

    Uniform base (default): the same `number_queries` rotations for every image (sd.generate_queries + repeat,
    agent.py:253-256).  ``fisher_A`` [B,3,3] (config.pretrain_fisher, agent.py:247-251): per-image samples from
    MatrixFisherN(A) drawn on the device, with their base log-likelihood.
    Returns (est_rotation [B,3,3], samples [B,Q,3,3], log_prob [B,Q])."""
    dev = feature.device
    B = feature.shape[0]
    if fisher_A is not None:
        from .fisher import MatrixFisherN
        pre = MatrixFisherN(fisher_A.to(dev))
        base = pre._sample(number_queries, seed=seed)                  # [B,Q,3,3]
        Q = base.shape[1]
        rows = base.reshape(-1, 3, 3)
        base_ll = pre._log_prob(rows).reshape(B, Q)
        idx = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(Q)
        samples, ldj = flow.inverse(rows, feature, feature_index=idx, mlp_mode=mlp_mode)
        log_prob = -ldj.reshape(B, Q) + base_ll
        best = torch.argmax(log_prob, dim=-1)
        samples = samples.reshape(B, Q, 3, 3)
        return samples[torch.arange(B, device=dev), best], samples, log_prob
    if base_samples is None:
        base_samples = rgrid.generate_queries(number_queries, "random", dev)
    Q = base_samples.shape[0]
    rows = base_samples[None].expand(B, Q, 3, 3).reshape(-1, 3, 3)
    idx = torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(Q)
    samples, ldj = flow.inverse(rows, feature, feature_index=idx, mlp_mode=mlp_mode)
    log_prob = -ldj.reshape(B, Q)
    if base_ll is not None:
        log_prob = log_prob + base_ll
    best = torch.argmax(log_prob, dim=-1)
    samples = samples.reshape(B, Q, 3, 3)
    est = samples[torch.arange(B, device=dev), best]
    return est, samples, log_prob
