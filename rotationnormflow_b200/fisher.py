"""Matrix-Fisher base distribution: per-image constants for the fused log-prob epilogue.

Follows utils/fisher.py:67-97 (proper_svd_N, matrix_fisher_norm_N type 1) and :217-232 (_log_prob):
    log p(R) = sum_ij A_ij R_ij - (S0+S1+S2) - log c,   c = 1/sqrt(8 pi (S0+S1)(S1+S2)(S0+S2))
with S the *proper* singular values (S2 multiplied by det(U) det(V)).  The 3x3 SVD is the same library call the
reference makes, once per image; the per-rotation term sum(A*R) is evaluated inside the flow kernel.
"""
from __future__ import annotations

import math

import torch


def fisher_constants(A: torch.Tensor):
    """A [B,3,3] -> (A9 [B,9] float32 contiguous, c [B] = sum(S) + log c_norm) on A's device."""
    if A.dim() != 3 or tuple(A.shape[1:]) != (3, 3):
        raise ValueError(f"A must be [B,3,3], got {tuple(A.shape)}")
    A = A.to(torch.float32)
    U, S, V = torch.svd(A)
    S2 = S[:, 2] * torch.det(U) * torch.det(V)
    S0, S1 = S[:, 0], S[:, 1]
    log_norm = -0.5 * torch.log(8 * math.pi * (S0 + S1) * (S2 + S1) * (S0 + S2))
    c = (S0 + S1 + S2) + log_norm
    return A.reshape(-1, 9).contiguous(), c.contiguous()
