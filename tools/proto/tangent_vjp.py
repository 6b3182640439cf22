"""Prototype (float64, CPU) of the tangent-space VJPs used by the differentiable training path, checked against autograd through the
oracle restatement of the reference.  Not product code: the CUDA kernels of csrc/train_ops.cu implement exactly these formulas."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import math
import torch
from oracle import rnf_oracle as orc

torch.manual_seed(0)
K = 64


def vee(R, G):
    """tangent (body-frame) vector of a matrix gradient G at R: g_i = <G, R [e_i]x>"""
    H = R.transpose(1, 2) @ G
    return torch.stack([H[:, 2, 1] - H[:, 1, 2], H[:, 0, 2] - H[:, 2, 0], H[:, 1, 0] - H[:, 0, 1]], dim=1)


def hat(g):
    Z = torch.zeros_like(g[:, 0])
    return torch.stack([torch.stack([Z, -g[:, 2], g[:, 1]], 1), torch.stack([g[:, 2], Z, -g[:, 0]], 1), torch.stack([-g[:, 1], g[:, 0], Z], 1)], 1)


def mix_forward(R, out, p0):
    p1, p2 = (p0 + 1) % 3, (p0 + 2) % 3
    x, y, z = R[:, :, p0], R[:, :, p1], R[:, :, p2]
    a, w = out[:, :K], out[:, K:].reshape(-1, K, 3)
    al = -(w * x[:, None]).sum(-1)
    be = (w * z[:, None]).sum(-1)
    n = (al * al + be * be).sqrt()
    s = 0.7 / (1 + n)
    alp, bep = s * al, s * be
    X, Y = -1 - alp, -bep
    D2 = X * X + Y * Y
    th = math.pi + 2 * torch.atan(bep / (1 + alp))
    f = (1 - alp * alp - bep * bep) / D2
    sp = torch.nn.functional.softplus(a)
    S = sp.sum(1, keepdim=True)
    pi = sp / S
    Th = (pi * th).sum(1)
    F = (pi * f).sum(1)
    phi = Th - math.pi
    c, sn = torch.cos(phi)[:, None], torch.sin(phi)[:, None]
    tx, tz = x * c - z * sn, z * c + x * sn
    Rn = torch.zeros_like(R)
    Rn[:, :, p0], Rn[:, :, p1], Rn[:, :, p2] = tx, y, tz
    saved = dict(al=al, be=be, n=n, s=s, alp=alp, bep=bep, D2=D2, th=th, f=f, sp=sp, S=S, pi=pi, Th=Th, F=F, phi=phi, a=a, w=w)
    return Rn, F.log(), saved


def mix_backward(R, Rn, p0, sv, G, g):
    p1, p2 = (p0 + 1) % 3, (p0 + 2) % 3
    x, z = R[:, :, p0], R[:, :, p2]
    gp = vee(Rn, G)                                   # tangent of the incoming gradient, body frame of R'
    gTh = gp[:, p1]
    al, be, n, s, alp, bep, D2 = sv["al"], sv["be"], sv["n"], sv["s"], sv["alp"], sv["bep"], sv["D2"]
    th_ap, th_bp = -2 * bep / D2, 2 * (1 + alp) / D2
    om = 1 - alp * alp - bep * bep
    f_ap = (-2 * alp * D2 - om * 2 * (1 + alp)) / D2 ** 2
    f_bp = (-2 * bep * D2 - om * 2 * bep) / D2 ** 2
    t = torch.where(n > 1e-30, s / (n * (1 + n)), torch.zeros_like(n))
    J11, J12, J22 = s - t * al * al, -t * al * be, s - t * be * be
    th_a, th_b = th_ap * J11 + th_bp * J12, th_ap * J12 + th_bp * J22
    f_a, f_b = f_ap * J11 + f_bp * J12, f_ap * J12 + f_bp * J22
    pi, F = sv["pi"], sv["F"]
    A = pi * (gTh[:, None] * th_a + (g / F)[:, None] * f_a)
    B = pi * (gTh[:, None] * th_b + (g / F)[:, None] * f_b)
    sig = torch.sigmoid(sv["a"])
    d_a = sig / sv["S"] * (gTh[:, None] * (sv["th"] - sv["Th"][:, None]) + (g / F)[:, None] * (sv["f"] - F[:, None]))
    d_w = -A[:, :, None] * x[:, None] + B[:, :, None] * z[:, None]
    w = sv["w"]
    sA = R.transpose(1, 2) @ (A[:, :, None] * w).sum(1)[:, :, None]        # sum_k A_k u_k  (body frame)
    sB = R.transpose(1, 2) @ (B[:, :, None] * w).sum(1)[:, :, None]
    e = torch.eye(3, dtype=R.dtype)
    c = torch.linalg.cross(e[p0].expand_as(sA[:, :, 0]), -sA[:, :, 0]) + torch.linalg.cross(e[p2].expand_as(sB[:, :, 0]), sB[:, :, 0])
    M = R.transpose(1, 2) @ Rn                         # body rotation about e_p1 by phi
    g_om = (M @ gp[:, :, None])[:, :, 0] + c
    G_R = 0.5 * R @ hat(g_om)
    return G_R, torch.cat([d_a, d_w.reshape(-1, 3 * K)], 1)


def E_of(q):
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    return torch.stack([torch.stack([-x, -y, -z], 1), torch.stack([w, -z, y], 1), torch.stack([z, w, -x], 1), torch.stack([-y, x, w], 1)], 1)


def aff_forward(R, W):
    from oracle.stubs import matrix_to_quaternion, quaternion_to_matrix
    q = matrix_to_quaternion(R)
    p = (W @ q[:, :, None])[:, :, 0]
    l = p.norm(dim=1)
    ph = p / l[:, None]
    return quaternion_to_matrix(ph), l.log(), dict(q=q, ph=ph, l=l)


def aff_backward(R, Rn, W, sv, G, gL):
    gp = vee(Rn, G)
    q, ph, l = sv["q"], sv["ph"], sv["l"]
    d_p = (2 * (E_of(ph) @ gp[:, :, None])[:, :, 0] + gL[:, None] * ph) / l[:, None]
    G_W = d_p[:, :, None] * q[:, None, :]
    g_om = 0.5 * (E_of(q).transpose(1, 2) @ (W.transpose(1, 2) @ d_p[:, :, None]))[:, :, 0]
    return 0.5 * R @ hat(g_om), G_W


def check():
    N = 16
    g = torch.Generator().manual_seed(1)
    R = orc.random_rotations(N, g, torch.float64)
    out = torch.randn(N, 4 * K, generator=g, dtype=torch.float64) * 0.5
    A = torch.randn(N, 3, 3, generator=g, dtype=torch.float64)
    cw = torch.randn(N, generator=g, dtype=torch.float64)
    for p0 in (0, 1, 2):
        # reference formulas (oracle restatement) with autograd, loss = <A, R'> + cw . ldj
        Rr, outr = R.clone().requires_grad_(True), out.clone().requires_grad_(True)
        p = orc.PERMUTE_TABLE[p0]
        x, y = Rr[:, :, p[0]], Rr[:, :, p[1]]
        # build with the oracle's pieces
        weights = outr[:, :K]
        w = outr[:, K:].reshape(N, K, 3)
        w = w - (w * y[:, None]).sum(-1, keepdim=True) * y[:, None]
        r = -x / x.norm(dim=-1, keepdim=True)
        v = torch.linalg.cross(y, r, dim=-1)
        v = v / v.norm(dim=-1, keepdim=True)
        weights = torch.nn.functional.softplus(weights)
        weights = weights / weights.sum(-1, keepdim=True)
        w = 0.7 * w / (1 + w.norm(dim=-1, keepdim=True))
        h, _ = orc._mobius_h(x, w)
        th = orc._wrapped_angles(h, r, v)
        Th = (weights * th).sum(1, keepdim=True)
        tx = r * torch.cos(Th) + v * torch.sin(Th)
        ldj = orc._explicit_ldj(x, r, v, weights, w)
        tz = torch.linalg.cross(tx, y, dim=-1)
        tz = tz / tz.norm(dim=-1, keepdim=True)
        Rn_ref = orc._assemble(p, tx, y, tz)
        loss = (A * Rn_ref).sum() + (cw * ldj).sum()
        loss.backward()
        Rn, ldj2, sv = mix_forward(R, out, p0)
        assert (Rn - Rn_ref.detach()).abs().max() < 1e-12 and (ldj2 - ldj.detach()).abs().max() < 1e-12
        G_R, G_out = mix_backward(R, Rn, p0, sv, A, cw)
        print("perm", p0, "d_out err", (G_out - outr.grad).abs().max().item(), "tangent dR err", (vee(R, G_R) - vee(R, Rr.grad)).abs().max().item())
    # affine
    W = torch.eye(4, dtype=torch.float64) + 0.3 * torch.randn(N, 4, 4, generator=g, dtype=torch.float64)
    Rr, Wr = R.clone().requires_grad_(True), W.clone().requires_grad_(True)
    Rn_ref, ldj = orc.quat_affine(Wr, Rr)
    ((A * Rn_ref).sum() + (cw * ldj).sum()).backward()
    Rn, logl, sv = aff_forward(R, W)
    G_R, G_W = aff_backward(R, Rn, W, sv, A, -4 * cw)
    G_W = G_W + cw[:, None, None] * torch.linalg.inv(W).transpose(1, 2)
    print("affine dW err", (G_W - Wr.grad).abs().max().item(), "tangent dR err", (vee(R, G_R) - vee(R, Rr.grad)).abs().max().item())


if __name__ == "__main__":
    check()
