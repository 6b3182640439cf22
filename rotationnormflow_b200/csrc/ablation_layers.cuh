// ablation_layers.cuh -- the reference's ablation replacements of the quaternion affine layer (flow/affineflow.py:27-41,55-70):
//   RNF_LAYER_SMITH9  calculate_9        flow/squeezetrans.py:197-232   A = M R, Gram-Schmidt ("Smith") of the first two columns,
//                                                                        log-det by tangent propagation
//   RNF_LAYER_SMITH36 calculate_36       flow/squeezetrans.py:291-331   6x6 matrix on the 6-D representation (columns 0, 1 of R)
//   RNF_LAYER_POLAR9L calculate_9_l      flow/rottrans.py:69-72         orthogonal polar factor U V^T of M R; log-det 0
//   RNF_LAYER_POLAR9R calculate_9_r      flow/rottrans.py:75-78         ... of R M
//   RNF_LAYER_RIGHT9  calculate_9_r_smith flow/rottrans.py:81-91        R Q with Q = Gram-Schmidt(M) (host side); log-det 0
// One rotation per thread, everything in registers.  These layers appear in no shipped settings/*.yml; they are kept out of line
// (a by-value call) so that the hot kernels' register allocation is the one of the Mobius / quaternion path.
#pragma once
#include "rnf_common.cuh"
#include "so3_math.cuh"

namespace rnf {

struct Rot9 {
  float m[9];
  float ldj;
};

namespace ablation {

__device__ __forceinline__ void col3(const float A[9], int c, float o[3]) { o[0] = A[c]; o[1] = A[3 + c]; o[2] = A[6 + c]; }

// normalise v and propagate NT tangents (squeezetrans.py accp_normalize): t = v / |v|, dt = d/|v| - v (v.d) / |v|^3
template <int NT>
__device__ __forceinline__ void normalize_with_tangents(const float v[3], const float d[NT][3], float t[3], float dt[NT][3]) {
  const float n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  const float inv = 1.0f / sqrtf(n2);
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = inv * v[i];
#pragma unroll
  for (int k = 0; k < NT; ++k) {
    const float dn2 = 2.0f * (d[k][0] * v[0] + d[k][1] * v[1] + d[k][2] * v[2]);   // d|v|^2
    const float dnorm = dn2 * 0.5f * inv;                                          // d|v|
    const float dinv = -dnorm * inv * inv;                                         // d(1/|v|)
#pragma unroll
    for (int i = 0; i < 3; ++i) dt[k][i] = dinv * v[i] + inv * d[k][i];
  }
}

// Gram-Schmidt of (c0, c1) with the three tangents of each -> rotation T = [t0 t1 t0 x t1] and
// log |det [ (dT_k T^T)_{01}, (dT_k T^T)_{02}, (dT_k T^T)_{12} ]_k|       (squeezetrans.py:209-232 / :309-331)
__device__ __forceinline__ float smith_tail(const float c0[3], const float dc0[3][3], const float c1[3], const float dc1[3][3], float R[9]) {
  float t0[3], dt0[3][3], t1[3], dt1[3][3], raw1[3], draw1[3][3];
  normalize_with_tangents<3>(c0, dc0, t0, dt0);
  const float dot = t0[0] * c1[0] + t0[1] * c1[1] + t0[2] * c1[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) raw1[i] = c1[i] - dot * t0[i];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float ddot = (dt0[k][0] * c1[0] + dt0[k][1] * c1[1] + dt0[k][2] * c1[2]) + (t0[0] * dc1[k][0] + t0[1] * dc1[k][1] + t0[2] * dc1[k][2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) draw1[k][i] = dc1[k][i] - (ddot * t0[i] + dot * dt0[k][i]);
  }
  normalize_with_tangents<3>(raw1, draw1, t1, dt1);
  float t2[3], dt2[3][3];
  cross3(t0, t1, t2);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a[3], b[3];
    cross3(t0, dt1[k], a);
    cross3(dt0[k], t1, b);
#pragma unroll
    for (int i = 0; i < 3; ++i) dt2[k][i] = a[i] + b[i];
  }
  // delta_k = dT_k T^T with T = [t0 t1 t2] as columns: (delta_k)_{ab} = sum_c dT_k[a][c] T[b][c]
  float V[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float d01 = dt0[k][0] * t0[1] + dt1[k][0] * t1[1] + dt2[k][0] * t2[1];
    const float d02 = dt0[k][0] * t0[2] + dt1[k][0] * t1[2] + dt2[k][0] * t2[2];
    const float d12 = dt0[k][1] * t0[2] + dt1[k][1] * t1[2] + dt2[k][1] * t2[2];
    V[k][0] = d01; V[k][1] = d02; V[k][2] = d12;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) { R[3 * i] = t0[i]; R[3 * i + 1] = t1[i]; R[3 * i + 2] = t2[i]; }
  const float det = det3f(V[0][0], V[0][1], V[0][2], V[1][0], V[1][1], V[1][2], V[2][0], V[2][1], V[2][2]);
  return logf(fabsf(det));
}

// calculate_9: A = M R; tangents A G_k of a right perturbation R (I + eps G_k), G_0 = e01 - e10, G_1 = e02 - e20, G_2 = e12 - e21
__device__ __forceinline__ float smith9(const float* __restrict__ M, float R[9]) {
  float A[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) A[3 * i + j] = M[3 * i] * R[j] + M[3 * i + 1] * R[3 + j] + M[3 * i + 2] * R[6 + j];
  float a0[3], a1[3], a2[3];
  col3(A, 0, a0); col3(A, 1, a1); col3(A, 2, a2);
  float dc0[3][3], dc1[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    dc0[0][i] = -a1[i]; dc0[1][i] = -a2[i]; dc0[2][i] = 0.0f;      // column 0 of A G_k
    dc1[0][i] = a0[i];  dc1[1][i] = 0.0f;   dc1[2][i] = -a2[i];    // column 1 of A G_k
  }
  return smith_tail(a0, dc0, a1, dc1, R);
}

// calculate_36: 6-D vector (R[:,0], R[:,1]) and its tangents under the LEFT perturbation (I + eps G_k) R, mapped by M6
__device__ __forceinline__ float smith36(const float* __restrict__ M6, float R[9]) {
  float v[6], dv[3][6];
  float r0[3], r1[3];
  col3(R, 0, r0); col3(R, 1, r1);
#pragma unroll
  for (int i = 0; i < 3; ++i) { v[i] = r0[i]; v[3 + i] = r1[i]; }
  // G_0 u = (u1, -u0, 0), G_1 u = (u2, 0, -u0), G_2 u = (0, u2, -u1)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float* u = h == 0 ? r0 : r1;
    dv[0][3 * h] = u[1];  dv[0][3 * h + 1] = -u[0]; dv[0][3 * h + 2] = 0.0f;
    dv[1][3 * h] = u[2];  dv[1][3 * h + 1] = 0.0f;  dv[1][3 * h + 2] = -u[0];
    dv[2][3 * h] = 0.0f;  dv[2][3 * h + 1] = u[2];  dv[2][3 * h + 2] = -u[1];
  }
  float t[6], dt[3][6];
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    float s = 0.0f, s0 = 0.0f, s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const float w = __ldg(M6 + 6 * a + b);
      s = fmaf(w, v[b], s); s0 = fmaf(w, dv[0][b], s0); s1 = fmaf(w, dv[1][b], s1); s2 = fmaf(w, dv[2][b], s2);
    }
    t[a] = s; dt[0][a] = s0; dt[1][a] = s1; dt[2][a] = s2;
  }
  float c0[3], c1[3], dc0[3][3], dc1[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c0[i] = t[i]; c1[i] = t[3 + i];
#pragma unroll
    for (int k = 0; k < 3; ++k) { dc0[k][i] = dt[k][i]; dc1[k][i] = dt[k][3 + i]; }
  }
  return smith_tail(c0, dc0, c1, dc1, R);
}

// Orthogonal polar factor of a 3x3 matrix = U V^T of its SVD (rottrans.py:71,77), by the scaled Newton iteration
// X <- (g X + X^{-T} / g) / 2, g = sqrt(|X^{-1}|_F / |X|_F)  (Higham): quadratic convergence, ten steps cover condition numbers
// far beyond what a near-identity 3x3 times a rotation reaches; keeps the sign of det A like U V^T does.
__device__ __forceinline__ void polar3(float X[9]) {
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    // adjugate-transpose (cofactor matrix) and determinant
    float C[9];
    C[0] = X[4] * X[8] - X[5] * X[7]; C[1] = X[5] * X[6] - X[3] * X[8]; C[2] = X[3] * X[7] - X[4] * X[6];
    C[3] = X[2] * X[7] - X[1] * X[8]; C[4] = X[0] * X[8] - X[2] * X[6]; C[5] = X[1] * X[6] - X[0] * X[7];
    C[6] = X[1] * X[5] - X[2] * X[4]; C[7] = X[2] * X[3] - X[0] * X[5]; C[8] = X[0] * X[4] - X[1] * X[3];
    const float det = X[0] * C[0] + X[1] * C[1] + X[2] * C[2];
    const float idet = 1.0f / det;                    // X^{-T} = C / det
    float nx = 0.0f, nc = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) { nx = fmaf(X[i], X[i], nx); nc = fmaf(C[i], C[i], nc); }
    const float g = sqrtf(sqrtf(nc) * fabsf(idet) / sqrtf(nx));
    const float a = 0.5f * g, b = 0.5f * idet / g;
#pragma unroll
    for (int i = 0; i < 9; ++i) X[i] = fmaf(a, X[i], b * C[i]);
  }
}

__device__ __forceinline__ void matmul3(const float* __restrict__ A, const float* __restrict__ B, float out[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) out[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

}  // namespace ablation

// One ablation layer on one rotation.  W: the layer's parameter block for the requested direction (rnf_abi.h layer kinds).
static __device__ __noinline__ Rot9 ablation_layer(int kind, const float* __restrict__ W, Rot9 in) {
  Rot9 out = in;
  out.ldj = 0.0f;
  float M[9];
  if (kind != RNF_LAYER_SMITH36) {
#pragma unroll
    for (int i = 0; i < 9; ++i) M[i] = __ldg(W + i);
  }
  if (kind == RNF_LAYER_SMITH9) {
    out.ldj = ablation::smith9(M, out.m);
  } else if (kind == RNF_LAYER_SMITH36) {
    out.ldj = ablation::smith36(W, out.m);
  } else if (kind == RNF_LAYER_POLAR9L) {
    float A[9];
    ablation::matmul3(M, in.m, A);
    ablation::polar3(A);
#pragma unroll
    for (int i = 0; i < 9; ++i) out.m[i] = A[i];
  } else if (kind == RNF_LAYER_POLAR9R) {
    float A[9];
    ablation::matmul3(in.m, M, A);
    ablation::polar3(A);
#pragma unroll
    for (int i = 0; i < 9; ++i) out.m[i] = A[i];
  } else {  // RNF_LAYER_RIGHT9
    ablation::matmul3(in.m, M, out.m);
  }
  return out;
}

// Non-Mobius layer dispatch shared by the three flow kernels: FAST selects the SFU variant of calculate_16.
template <bool FAST>
__device__ __forceinline__ void affine_family_layer(const LayerDev& L, const float* __restrict__ W, float R[9], float& ldj) {
  if (L.kind == RNF_LAYER_AFFINE) {
    float Wr[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) Wr[i] = __ldg(W + i);
    const float loglen = FAST ? quat_affine_fast(Wr, R) : quat_affine(Wr, R);
    if (L.has_ldj) ldj += Wr[16] - 4.0f * loglen;
  } else {
    Rot9 io;
#pragma unroll
    for (int i = 0; i < 9; ++i) io.m[i] = R[i];
    io.ldj = 0.0f;
    io = ablation_layer(L.kind, W, io);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = io.m[i];
    ldj += io.ldj;
  }
}

}  // namespace rnf
