// train_ops.cu -- the DIFFERENTIABLE per-layer operators behind Flow.forward / Flow.inverse when autograd is on
// (training: agent.py:87 loss.backward(); eval.py:468-477 nll_grad) or when config.segments is not the 64 the fused tcgen05
// kernels are specialised for.  One thread per rotation, plain FP32 with the precise math functions, any number K of mixture
// components.  The conditioner MLP itself stays a sequence of library GEMMs on the Python side (torch autograd supplies its
// backward); these kernels are everything after it:
//   mobius_mixture  forward : (R, out[N,4K]) -> (R', ldj)            flow/mobiusflow.py:58-85 (forward) / :141-183 (inverse,
//                              with the 15 bisection halvings of BinFind.forward, :196-224)
//   mobius_mixture  backward: VJP in the TANGENT space of SO(3).  A layer maps R -> R' = R Rot(e_p1, phi) with
//                              phi = Theta(a, alpha, beta) - pi, alpha_k = -(R^T w_k)[p0], beta_k = (R^T w_k)[p2]; a body-frame
//                              perturbation omega of R gives omega' = M^T omega + (dphi) e_p1, so with g' = vee(R'^T G') the tangent
//                              gradient is g = M g' + e_p0 x (-R^T sum_k A_k w_k) + e_p2 x (R^T sum_k B_k w_k), A_k = dL/dalpha_k,
//                              B_k = dL/dbeta_k, and G_R = R [g]x / 2.  Only the tangential part of a rotation gradient reaches the
//                              parameters (every layer keeps R on the manifold), so parameter and feature gradients equal the
//                              reference's autograd; verified against autograd through the oracle (tests/test_gpu_train.py,
//                              tools/proto/tangent_vjp.py).  Inverse direction: implicit-function theorem at the returned root,
//                              which is what BinFind.backward (flow/mobiusflow.py:248-273) evaluates.
//   quat_affine     forward / backward: calculate_16 (flow/squeezetrans.py:33-38) with per-row W; returns log|Wq| (the caller adds
//                              log|det W| with torch.linalg.slogdet, which is differentiable on its own).
#include <cuda_runtime.h>
#include <math.h>

#include "mobius_math.cuh"
#include "rnf_common.cuh"

namespace rnf {
namespace {

constexpr int kMaxK = 256;          // prepared components of one rotation are kept in (L1-backed) local memory

struct Frame3 {
  float x[3], y[3], z[3];
};

__device__ __forceinline__ void load_cols(const float* __restrict__ R, int p0, Frame3& f, float Rm[9]) {
#pragma unroll
  for (int i = 0; i < 9; ++i) Rm[i] = R[i];
  const int p1 = (p0 + 1) % 3, p2 = (p0 + 2) % 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) { f.x[i] = Rm[3 * i + p0]; f.y[i] = Rm[3 * i + p1]; f.z[i] = Rm[3 * i + p2]; }
}

// prepared component: in-plane squashed centre (alpha', beta'), softplus weight
struct Prep {
  float al, be, n, s, alp, bep, sp;
};

__device__ __forceinline__ Prep prepare(const float* __restrict__ out, int K, int k, const float r[3], const float v[3]) {
  Prep c;
  const float a = out[k];
  const float w0 = out[K + 3 * k], w1 = out[K + 3 * k + 1], w2 = out[K + 3 * k + 2];
  c.al = fmaf(w2, r[2], fmaf(w1, r[1], w0 * r[0]));       // the projection (I - y y^T) w leaves these two dot products unchanged
  c.be = fmaf(w2, v[2], fmaf(w1, v[1], w0 * v[0]));
  c.n = sqrtf(fmaf(c.be, c.be, c.al * c.al));
  c.s = 0.7f / (1.0f + c.n);
  c.alp = c.s * c.al;
  c.bep = c.s * c.be;
  c.sp = softplus_torch(a);
  return c;
}

// wrapped angle of h_w'(z) and f = (1 - |w'|^2) / |z - w'|^2 at the in-plane point z = (cz, sz)
__device__ __forceinline__ void eval_point(float cz, float sz, float alp, float bep, float& theta, float& f) {
  const float X = cz - alp, Y = sz - bep;
  const float D2 = fmaf(Y, Y, X * X);
  const float om = 1.0f - fmaf(bep, bep, alp * alp);
  f = om / D2;
  const float hr = fmaf(f, X, -alp), hv = fmaf(f, Y, -bep);
  const float th = atan2f(hv, hr);
  theta = th >= 0.0f ? th : th + kTwoPi;
}

template <bool INV>
__global__ void __launch_bounds__(128) mobius_mixture_fwd_kernel(const float* __restrict__ R, const float* __restrict__ out, int64_t N, int K,
                                                                 int p0, float* __restrict__ R_out, float* __restrict__ ldj_out,
                                                                 float* __restrict__ theta_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Frame3 F;
  float Rm[9];
  load_cols(R + n * 9, p0, F, Rm);
  float r[3], v[3];
  make_frame(F.x, F.y, r, v);                              // r = -x/|x|, v = (y x r)/|y x r|   (flow/mobiusflow.py:64-67)
  const float* o = out + n * 4 * (int64_t)K;
  const float zr = dot3(F.x, r), zv = dot3(F.x, v);        // the moving column in its own frame: (-|x|, ~0)
  float pal[INV ? kMaxK : 1], pbe[INV ? kMaxK : 1], psp[INV ? kMaxK : 1];
  float S = 0.0f, Sth = 0.0f, Sf = 0.0f;
  for (int k = 0; k < K; ++k) {
    const Prep c = prepare(o, K, k, r, v);
    S += c.sp;
    if (INV) { pal[k] = c.alp; pbe[k] = c.bep; psp[k] = c.sp; }
    else {
      float th, f;
      eval_point(zr, zv, c.alp, c.bep, th, f);
      Sth = fmaf(c.sp, th, Sth);
      Sf = fmaf(c.sp, f, Sf);
    }
  }
  float theta, ldj;
  if (!INV) {
    theta = Sth / S;
    ldj = logf(Sf / S);
  } else {
    // target angle of the given column in its own frame (flow/mobiusflow.py:157-167), then BinFind.forward (:196-224)
    float ys = atan2f(zv, zr);
    ys = ys >= 0.0f ? ys : ys + kTwoPi;
    if (fabsf(ys - kTwoPi) < 1e-4f) ys = 0.0f;
    float lo = kPi / 2.0f, hi = 1.5f * kPi, x0 = 0.0f;
    for (int it = 0; it < 15; ++it) {
      x0 = (lo + hi) / 2.0f;
      float sn, cs;
      sincosf(x0, &sn, &cs);
      float Fs = 0.0f;
      for (int k = 0; k < K; ++k) {
        float th, f;
        eval_point(cs, sn, pal[k], pbe[k], th, f);
        Fs = fmaf(psp[k], th, Fs);
      }
      const float fx0 = Fs / S - ys;
      const float half_w = (hi - lo) / 2.0f;
      if (fx0 < 0.0f) lo = lo + half_w;
      else if (fx0 >= 0.0f) hi = hi - half_w;
    }
    theta = x0;
    float sn, cs;
    sincosf(x0, &sn, &cs);
    for (int k = 0; k < K; ++k) {
      float th, f;
      eval_point(cs, sn, pal[k], pbe[k], th, f);
      Sf = fmaf(psp[k], f, Sf);
    }
    ldj = -logf(Sf / S);
  }
  float nx[3], nz[3];
  circle_point(r, v, theta, nx);                           // r cos(theta) + v sin(theta)   (flow/mobiusflow.py:102,169)
  cross3(nx, F.y, nz);
  normalize3(nz);
  const int p2 = (p0 + 2) % 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) { Rm[3 * i + p0] = nx[i]; Rm[3 * i + p2] = nz[i]; }
#pragma unroll
  for (int i = 0; i < 9; ++i) R_out[n * 9 + i] = Rm[i];
  ldj_out[n] = ldj;
  theta_out[n] = theta;
}

// vee(R^T G): tangent (body frame) vector of a matrix gradient G at R, g_i = <G, R [e_i]x>
__device__ __forceinline__ void tangent_of(const float R[9], const float* __restrict__ G, float g[3]) {
  float H[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) H[3 * i + j] = R[i] * G[j] + R[3 + i] * G[3 + j] + R[6 + i] * G[6 + j];
  g[0] = H[7] - H[5];
  g[1] = H[2] - H[6];
  g[2] = H[3] - H[1];
}

// G_R = R [g]x / 2
__device__ __forceinline__ void gradient_from_tangent(const float R[9], const float g[3], float* __restrict__ G) {
  const float h[9] = {0.0f, -g[2], g[1], g[2], 0.0f, -g[0], -g[1], g[0], 0.0f};
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) G[3 * i + j] = 0.5f * (R[3 * i] * h[j] + R[3 * i + 1] * h[3 + j] + R[3 * i + 2] * h[6 + j]);
}

template <bool INV>
__global__ void __launch_bounds__(128) mobius_mixture_bwd_kernel(const float* __restrict__ R, const float* __restrict__ out, int64_t N, int K,
                                                                 int p0, const float* __restrict__ theta_in, const float* __restrict__ R_out,
                                                                 const float* __restrict__ G_Rout, const float* __restrict__ g_ldj,
                                                                 float* __restrict__ G_R, float* __restrict__ G_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Frame3 F;
  float Rm[9], Rn[9];
  load_cols(R + n * 9, p0, F, Rm);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rn[i] = R_out[n * 9 + i];
  float r[3], v[3];
  make_frame(F.x, F.y, r, v);
  const float* o = out + n * 4 * (int64_t)K;
  const int p1 = (p0 + 1) % 3, p2 = (p0 + 2) % 3;
  // evaluation point of the mixture map: the moving column (forward) or the returned root (inverse)
  float cz, sz;
  if (INV) sincosf(theta_in[n], &sz, &cz);
  else { cz = dot3(F.x, r); sz = dot3(F.x, v); }
  float gp[3];
  tangent_of(Rn, G_Rout + n * 9, gp);
  const float gL = g_ldj[n];
  // pass 1: S, Theta = sum pi theta, F = sum pi f, F_theta = sum pi df/dtheta
  float S = 0.0f, Sth = 0.0f, Sf = 0.0f, Sft = 0.0f;
  for (int k = 0; k < K; ++k) {
    const Prep c = prepare(o, K, k, r, v);
    float th, f;
    eval_point(cz, sz, c.alp, c.bep, th, f);
    S += c.sp;
    Sth = fmaf(c.sp, th, Sth);
    Sf = fmaf(c.sp, f, Sf);
    if (INV) {
      const float X = cz - c.alp, Y = sz - c.bep;
      const float D2 = fmaf(Y, Y, X * X);
      const float om = 1.0f - fmaf(c.bep, c.bep, c.alp * c.alp);
      Sft = fmaf(c.sp, -2.0f * om * (c.alp * sz - c.bep * cz) / (D2 * D2), Sft);
    }
  }
  const float Th = Sth / S, Fm = Sf / S, Ft = Sft / S;
  // coefficients of dTheta-like and dF-like parameter derivatives in dL
  //   forward: phi = Theta - pi, ldj = log F                 -> cG = g'[p1],                      cF = gL / F
  //   inverse: psi = theta* - pi with G(theta*, p) = Theta - pi = 0, ldj = -log F(theta*, p)
  //            dpsi/dp = -G_p / F, dldj/dp = -F_p / F + F_theta G_p / F^2   -> cG = -g'[p1]/F + gL F_theta / F^2, cF = -gL / F
  float cG, cF;
  if (!INV) { cG = gp[p1]; cF = gL / Fm; }
  else { cG = -gp[p1] / Fm + gL * Ft / (Fm * Fm); cF = -gL / Fm; }
  float sA[3] = {0.f, 0.f, 0.f}, sB[3] = {0.f, 0.f, 0.f};          // sum_k A_k w_k, sum_k B_k w_k (world frame)
  float* go = G_out + n * 4 * (int64_t)K;
  for (int k = 0; k < K; ++k) {
    const Prep c = prepare(o, K, k, r, v);
    float th, f;
    eval_point(cz, sz, c.alp, c.bep, th, f);
    const float X = cz - c.alp, Y = sz - c.bep;
    const float D2 = fmaf(Y, Y, X * X);
    const float om = 1.0f - fmaf(c.bep, c.bep, c.alp * c.alp);
    const float th_ap = 2.0f * Y / D2, th_bp = -2.0f * X / D2;
    const float f_ap = (-2.0f * c.alp * D2 + 2.0f * om * X) / (D2 * D2);
    const float f_bp = (-2.0f * c.bep * D2 + 2.0f * om * Y) / (D2 * D2);
    const float t = c.n > 1e-30f ? c.s / (c.n * (1.0f + c.n)) : 0.0f;     // squash Jacobian: d(alpha', beta') / d(alpha, beta)
    const float J11 = c.s - t * c.al * c.al, J12 = -t * c.al * c.be, J22 = c.s - t * c.be * c.be;
    const float th_a = th_ap * J11 + th_bp * J12, th_b = th_ap * J12 + th_bp * J22;
    const float f_a = f_ap * J11 + f_bp * J12, f_b = f_ap * J12 + f_bp * J22;
    const float pi = c.sp / S;
    const float A = pi * (cG * th_a + cF * f_a);
    const float B = pi * (cG * th_b + cF * f_b);
    const float a = o[k];
    const float sig = 1.0f / (1.0f + expf(-a));                           // d softplus / da (threshold 20: sigmoid(20) = 1 - 2e-9)
    go[k] = sig / S * (cG * (th - Th) + cF * (f - Fm));
    // alpha = w.r = -w.x, beta = w.v = w.z on the manifold
    const float w0 = o[K + 3 * k], w1 = o[K + 3 * k + 1], w2 = o[K + 3 * k + 2];
    go[K + 3 * k] = fmaf(B, F.z[0], -A * F.x[0]);
    go[K + 3 * k + 1] = fmaf(B, F.z[1], -A * F.x[1]);
    go[K + 3 * k + 2] = fmaf(B, F.z[2], -A * F.x[2]);
    sA[0] = fmaf(A, w0, sA[0]); sA[1] = fmaf(A, w1, sA[1]); sA[2] = fmaf(A, w2, sA[2]);
    sB[0] = fmaf(B, w0, sB[0]); sB[1] = fmaf(B, w1, sB[1]); sB[2] = fmaf(B, w2, sB[2]);
  }
  // body-frame versions u = R^T s, then c = e_p0 x (-uA) + e_p2 x (uB)
  float uA[3], uB[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uA[i] = Rm[i] * sA[0] + Rm[3 + i] * sA[1] + Rm[6 + i] * sA[2];
    uB[i] = Rm[i] * sB[0] + Rm[3 + i] * sB[1] + Rm[6 + i] * sB[2];
  }
  float e0[3] = {0.f, 0.f, 0.f}, e2[3] = {0.f, 0.f, 0.f}, nA[3] = {-uA[0], -uA[1], -uA[2]}, c1[3], c2[3];
  e0[p0] = 1.0f;
  e2[p2] = 1.0f;
  cross3(e0, nA, c1);
  cross3(e2, uB, c2);
  // M = R^T R' (body rotation about e_p1); g = M g' + c
  float g[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float Mij = Rm[i] * Rn[j] + Rm[3 + i] * Rn[3 + j] + Rm[6 + i] * Rn[6 + j];
      m = fmaf(Mij, gp[j], m);
    }
    g[i] = m + c1[i] + c2[i];
  }
  gradient_from_tangent(Rm, g, G_R + n * 9);
}

// E(q) omega = q (x) (0, omega): the 4x3 matrix of a body-frame perturbation of a unit quaternion (real first)
__device__ __forceinline__ void E_mul(const float q[4], const float w[3], float o[4]) {
  o[0] = -q[1] * w[0] - q[2] * w[1] - q[3] * w[2];
  o[1] = q[0] * w[0] - q[3] * w[1] + q[2] * w[2];
  o[2] = q[3] * w[0] + q[0] * w[1] - q[1] * w[2];
  o[3] = -q[2] * w[0] + q[1] * w[1] + q[0] * w[2];
}
__device__ __forceinline__ void ET_mul(const float q[4], const float d[4], float o[3]) {
  o[0] = -q[1] * d[0] + q[0] * d[1] + q[3] * d[2] - q[2] * d[3];
  o[1] = -q[2] * d[0] - q[3] * d[1] + q[0] * d[2] + q[1] * d[3];
  o[2] = -q[3] * d[0] + q[2] * d[1] - q[1] * d[2] + q[0] * d[3];
}

__global__ void __launch_bounds__(128) quat_affine_fwd_kernel(const float* __restrict__ R, const float* __restrict__ W, int64_t N,
                                                              float* __restrict__ R_out, float* __restrict__ loglen_out) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float Rm[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rm[i] = R[n * 9 + i];
  const float loglen = quat_affine(W + n * 16, Rm);
#pragma unroll
  for (int i = 0; i < 9; ++i) R_out[n * 9 + i] = Rm[i];
  loglen_out[n] = loglen;
}

// d_p = (2 E(p^) g' + gL p^) / l ;  G_W = d_p q^T ;  g = E(q)^T W^T d_p / 2        (tangent-space VJP of calculate_16)
__global__ void __launch_bounds__(128) quat_affine_bwd_kernel(const float* __restrict__ R, const float* __restrict__ W, int64_t N,
                                                              const float* __restrict__ R_out, const float* __restrict__ G_Rout,
                                                              const float* __restrict__ g_loglen, float* __restrict__ G_R,
                                                              float* __restrict__ G_W) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float Rm[9], Rn[9], q[4], p[4];
#pragma unroll
  for (int i = 0; i < 9; ++i) { Rm[i] = R[n * 9 + i]; Rn[i] = R_out[n * 9 + i]; }
  mat_to_quat(Rm, q);
  const float* Wn = W + n * 16;
#pragma unroll
  for (int a = 0; a < 4; ++a) p[a] = fmaf(Wn[4 * a + 3], q[3], fmaf(Wn[4 * a + 2], q[2], fmaf(Wn[4 * a + 1], q[1], Wn[4 * a] * q[0])));
  const float len = sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
  float ph[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) ph[a] = p[a] / len;
  float gp[3], Eg[4], dp[4];
  tangent_of(Rn, G_Rout + n * 9, gp);
  E_mul(ph, gp, Eg);
  const float gL = g_loglen[n];
#pragma unroll
  for (int a = 0; a < 4; ++a) dp[a] = (2.0f * Eg[a] + gL * ph[a]) / len;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) G_W[n * 16 + 4 * a + b] = dp[a] * q[b];
  float Wt[4];
#pragma unroll
  for (int b = 0; b < 4; ++b) Wt[b] = Wn[b] * dp[0] + Wn[4 + b] * dp[1] + Wn[8 + b] * dp[2] + Wn[12 + b] * dp[3];
  float g[3];
  ET_mul(q, Wt, g);
#pragma unroll
  for (int i = 0; i < 3; ++i) g[i] *= 0.5f;
  gradient_from_tangent(Rm, g, G_R + n * 9);
}

}  // namespace

cudaError_t launch_train_mobius_fwd(const float* R, const float* out, int64_t N, int K, int perm, bool inverse, float* R_out, float* ldj,
                                    float* theta, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((N + 127) / 128);
  if (inverse) mobius_mixture_fwd_kernel<true><<<blocks, 128, 0, st>>>(R, out, N, K, perm, R_out, ldj, theta);
  else mobius_mixture_fwd_kernel<false><<<blocks, 128, 0, st>>>(R, out, N, K, perm, R_out, ldj, theta);
  return cudaGetLastError();
}

cudaError_t launch_train_mobius_bwd(const float* R, const float* out, int64_t N, int K, int perm, bool inverse, const float* theta,
                                    const float* R_out, const float* G_Rout, const float* g_ldj, float* G_R, float* G_out, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((N + 127) / 128);
  if (inverse) mobius_mixture_bwd_kernel<true><<<blocks, 128, 0, st>>>(R, out, N, K, perm, theta, R_out, G_Rout, g_ldj, G_R, G_out);
  else mobius_mixture_bwd_kernel<false><<<blocks, 128, 0, st>>>(R, out, N, K, perm, theta, R_out, G_Rout, g_ldj, G_R, G_out);
  return cudaGetLastError();
}

cudaError_t launch_train_affine_fwd(const float* R, const float* W, int64_t N, float* R_out, float* loglen, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  quat_affine_fwd_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(R, W, N, R_out, loglen);
  return cudaGetLastError();
}

cudaError_t launch_train_affine_bwd(const float* R, const float* W, int64_t N, const float* R_out, const float* G_Rout, const float* g_loglen,
                                    float* G_R, float* G_W, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  quat_affine_bwd_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(R, W, N, R_out, G_Rout, g_loglen, G_R, G_W);
  return cudaGetLastError();
}

int train_max_components() { return kMaxK; }

}  // namespace rnf
