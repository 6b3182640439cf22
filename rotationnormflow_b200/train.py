"""Differentiable evaluation of the flow (training, agent.py:87 ``loss.backward()``; eval.py:468-477 ``nll_grad``) and the general
``config.segments`` path.

The fused tcgen05 kernels of the inference path keep nothing for a backward pass.  With autograd on, ``Flow.forward`` /
``Flow.inverse`` run layer by layer instead, like the reference does (flow/flow.py:53-92):
  * the conditioner MLP (flow/condition.py:24-30) as library GEMMs -- torch autograd supplies its backward, weight gradients included;
  * everything after it in hand-written CUDA operators with hand-written vector-Jacobian products (csrc/train_ops.cu through the C
    ABI ``rnf_train_*``): the Mobius mixture with its frame, log-det and, in the inverse direction, the 15-halving bisection and the
    implicit-function gradient of BinFind.backward (flow/mobiusflow.py:248-273); and the quaternion affine map calculate_16.
Rotation gradients are carried in the tangent space of SO(3) (see csrc/train_ops.cu): parameter and feature gradients equal the
reference's; ``rotation.grad`` of the input is the tangential part of the reference's (the normal part of a gradient with respect to
a rotation matrix has no meaning on the manifold)."""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd.function import once_differentiable

from . import _cabi, engine


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class MobiusMixture(torch.autograd.Function):
    """(R [N,3,3], out [N,4K]) -> (R' [N,3,3], ldj [N]) for one Mobius layer in either direction (rnf_train_mobius_*)."""

    @staticmethod
    def forward(ctx, R, out, perm: int, inverse: bool, K: int):
        lib = _cabi.load()
        R, out = R.contiguous().float(), out.contiguous().float()
        N = R.shape[0]
        R_out, ldj, theta = torch.empty_like(R), R.new_empty(N), R.new_empty(N)
        with torch.cuda.device(R.device):
            _cabi.check(lib.rnf_train_mobius_forward(_ptr(R), _ptr(out), N, K, perm, int(inverse), _ptr(R_out), _ptr(ldj), _ptr(theta), _stream(R.device)))
        ctx.save_for_backward(R, out, theta, R_out)
        ctx.meta = (perm, inverse, K)
        return R_out, ldj

    @staticmethod
    @once_differentiable
    def backward(ctx, G_Rout, g_ldj):
        R, out, theta, R_out = ctx.saved_tensors
        perm, inverse, K = ctx.meta
        lib = _cabi.load()
        N = R.shape[0]
        G_Rout = torch.zeros_like(R) if G_Rout is None else G_Rout.contiguous().float()
        g_ldj = R.new_zeros(N) if g_ldj is None else g_ldj.contiguous().float()
        G_R, G_out = torch.empty_like(R), torch.empty_like(out)
        with torch.cuda.device(R.device):
            _cabi.check(lib.rnf_train_mobius_backward(_ptr(R), _ptr(out), N, K, perm, int(inverse), _ptr(theta), _ptr(R_out), _ptr(G_Rout), _ptr(g_ldj),
                                                      _ptr(G_R), _ptr(G_out), _stream(R.device)))
        return G_R, G_out, None, None, None


class QuatAffine(torch.autograd.Function):
    """(R [N,3,3], W [N,4,4]) -> (R' = q2m(W q / |W q|), log|W q|)   (calculate_16, flow/squeezetrans.py:33-38; rnf_train_affine_*)."""

    @staticmethod
    def forward(ctx, R, W):
        lib = _cabi.load()
        R, W = R.contiguous().float(), W.contiguous().float()
        N = R.shape[0]
        R_out, loglen = torch.empty_like(R), R.new_empty(N)
        with torch.cuda.device(R.device):
            _cabi.check(lib.rnf_train_affine_forward(_ptr(R), _ptr(W), N, _ptr(R_out), _ptr(loglen), _stream(R.device)))
        ctx.save_for_backward(R, W, R_out)
        return R_out, loglen

    @staticmethod
    @once_differentiable
    def backward(ctx, G_Rout, g_loglen):
        R, W, R_out = ctx.saved_tensors
        lib = _cabi.load()
        N = R.shape[0]
        G_Rout = torch.zeros_like(R) if G_Rout is None else G_Rout.contiguous().float()
        g_loglen = R.new_zeros(N) if g_loglen is None else g_loglen.contiguous().float()
        G_R, G_W = torch.empty_like(R), torch.empty_like(W)
        with torch.cuda.device(R.device):
            _cabi.check(lib.rnf_train_affine_backward(_ptr(R), _ptr(W), N, _ptr(R_out), _ptr(G_Rout), _ptr(g_loglen), _ptr(G_R), _ptr(G_W), _stream(R.device)))
        return G_R, G_W


def _affine_matrix(layer, rows, inverse: bool):
    """The 4x4 a layer applies in this direction ([1,4,4] or [N,4,4], differentiable torch algebra) and whether it has a log-det."""
    eye = None
    kind = layer.kind
    if kind == "aff_u":
        W = layer.mat
    elif kind == "aff_lu":
        W = layer.mat()
    elif kind in ("aff_c", "rot_c"):
        eye = torch.eye(4, device=rows.device, dtype=rows.dtype)
        W = engine.conditioner_torch(layer.net, rows).reshape(-1, 4, 4) + eye
    elif kind == "aff_clu":
        W = layer.net.weight(rows)                            # batch-coupled in the reference; reproduced as written
    elif kind == "rot_u":
        W = layer.rot
    else:
        raise NotImplementedError(f"layer kind {kind!r} has no differentiable / general-segments path (ablation layers run in the fused kernels only)")
    if kind.startswith("rot"):
        U, _, V = torch.svd(W)                                # flow/rottrans.py:15-16 (note V, not V^T)
        W = U.transpose(-1, -2) @ V
        return (W.transpose(-1, -2) if inverse else W), False
    return (torch.linalg.inv(W) if inverse else W), True


def composed_run(layers, perms, rotation, feature_rows, inverse: bool):
    """flow/flow.py:53-92 layer by layer with the differentiable operators.  ``feature_rows`` [N,F] row-aligned (or None)."""
    R = rotation.to(torch.float32)
    N = R.shape[0]
    ldjs = R.new_zeros(N)
    order = range(len(layers) - 1, -1, -1) if inverse else range(len(layers))
    for i in order:
        layer, p0 = layers[i], int(perms[i]) % 3
        if layer.kind == "mobius":
            y = R[:, :, (p0 + 1) % 3]
            inp = torch.cat((y, feature_rows), dim=-1) if layer.condition else y          # flow/mobiusflow.py:52-56
            out = engine.conditioner_torch(layer.conditioner, inp)
            R, ldj = MobiusMixture.apply(R, out, p0, inverse, layer.K)
        else:
            W, has_ldj = _affine_matrix(layer, feature_rows, inverse)
            R, loglen = QuatAffine.apply(R, W.expand(N, 4, 4))
            ldj = (torch.linalg.slogdet(W)[1] - 4.0 * loglen) if has_ldj else torch.zeros_like(loglen)
        ldjs = ldjs + ldj
    return R, ldjs
