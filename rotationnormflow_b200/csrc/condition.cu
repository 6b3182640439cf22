// condition.cu -- per-image part of the conditioners (run once per image batch, not per rotation).
//
// The reference feeds every conditioner the row-aligned `feature [N,F]` built by `.repeat`
// (agent.py:240-244, eval.py:450) and recomputes, for each of the N rows, terms that only depend on the image:
//   * Mobius conditioner, first layer:  W0 . cat(y, feature) + b0  (flow/mobiusflow.py:54-57, flow/condition.py:25)
//       = W0[:, :3] . y + b0  +  W0[:, 3:] . feature        <- the last term is hoisted here, per image.
//   * Condition16Trans / ConditionRot: the entire network output MLP(feature).reshape(4,4) + I
//       (flow/squeezetrans.py:47-48, flow/rottrans.py:43-44), its inverse (torch.linalg.inv, squeezetrans.py:54)
//       and both log|det| (my_det_4_4, squeezetrans.py:17-22,38).
// Output layout per image (floats): [n_mobius_slots][64] | [n_affine_slots][40 = W16, ld, pad3, Winv16, ldinv, pad3]
//                                   | [n_affine_slots][64] scratch (first-layer pre-activations).
#include "rnf_common.cuh"
#include "so3_math.cuh"

namespace rnf {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

// C[b, n] = sum_f feat[b, f] * Wf[n, f]      (both operands contiguous along f)
// row_index (nullable): image m reads feature row row_index[m] (the first row of its run, csrc/dedup.cu) instead of row m;
// count (nullable): device-side number of valid images -- rows at or beyond it are skipped (B is then only the capacity).
__global__ void __launch_bounds__(256) hoist_gemm_kernel(const float* __restrict__ feat, const float* __restrict__ Wf,
                                                         int64_t B, int Ntot, int F, float* __restrict__ cond,
                                                         int64_t stride, int n_mob, int n_aff, const int32_t* __restrict__ row_index,
                                                         const int32_t* __restrict__ count) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  if (count != nullptr) {
    const int64_t c = *count;
    B = c < B ? c : B;
  }
  if (m0 >= B) return;                         // block-uniform
  const int n0 = blockIdx.x * BN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  const int lr = tid / 4, lk = (tid % 4) * 4;  // load row / k offset
  const int64_t mrow = m0 + lr;
  const int64_t src_row = (mrow < B && row_index != nullptr) ? (int64_t)__ldg(row_index + mrow) : mrow;
  for (int k0 = 0; k0 < F; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int k = k0 + lk + i;
      const int64_t m = mrow;
      const int n = n0 + lr;
      As[lk + i][lr] = (m < B && k < F) ? __ldg(feat + src_row * F + k) : 0.0f;
      Ws[lk + i][lr] = (n < Ntot && k < F) ? __ldg(Wf + (int64_t)n * F + k) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[4], wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[k][tm + i]; wv[i] = Ws[k][tn + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + tm + i;
    if (m >= B) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tn + j;
      if (n >= Ntot) continue;
      const int s = n / kH, jj = n % kH;
      float* dst = s < n_mob ? cond + m * stride + (int64_t)s * kH + jj
                             : cond + m * stride + (int64_t)n_mob * kH + (int64_t)n_aff * kAffFloats + (int64_t)(s - n_mob) * kH + jj;
      *dst = acc[i][j];
    }
  }
}

__device__ void invert4(const float* A, float* inv) {
  // Gauss-Jordan with partial pivoting (what LAPACK getrf/getri amount to on a 4x4), fp32.
  float M[4][8];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) { M[r][c] = A[4 * r + c]; M[r][4 + c] = r == c ? 1.0f : 0.0f; }
  for (int c = 0; c < 4; ++c) {
    int piv = c;
    float best = fabsf(M[c][c]);
    for (int r = c + 1; r < 4; ++r)
      if (fabsf(M[r][c]) > best) { best = fabsf(M[r][c]); piv = r; }
    if (piv != c)
      for (int k = 0; k < 8; ++k) { const float t = M[c][k]; M[c][k] = M[piv][k]; M[piv][k] = t; }
    const float d = 1.0f / M[c][c];
    for (int k = 0; k < 8; ++k) M[c][k] *= d;
    for (int r = 0; r < 4; ++r) {
      if (r == c) continue;
      const float f = M[r][c];
      for (int k = 0; k < 8; ++k) M[r][k] = fmaf(-f, M[c][k], M[r][k]);
    }
  }
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) inv[4 * r + c] = M[r][4 + c];
}

// grid (B, n_aff), 64 threads: tail of ConditionalTransform(F, 16) for one (image, conditional affine layer).
__global__ void __launch_bounds__(64) cond_affine_kernel(const float* __restrict__ caff, float* __restrict__ cond,
                                                         int64_t stride, int n_mob, int n_aff, int is_rot,
                                                         const int32_t* __restrict__ count) {
  const int64_t b = blockIdx.x;
  if (count != nullptr && b >= *count) return;
  const int s = blockIdx.y;
  const int j = threadIdx.x;
  const float* w = caff + (int64_t)s * kCaffFloats;
  float* img = cond + b * stride;
  const float* pre = img + (int64_t)n_mob * kH + (int64_t)n_aff * kAffFloats + (int64_t)s * kH;
  __shared__ float h[kH];
  __shared__ float out[16];
  const float h0 = pre[j] + w[j];  // fc_first bias
  float cur = h0;
  const float* p = w + kH;
  for (int l = 0; l < 3; ++l) {
    __syncthreads();
    h[j] = fmaxf(cur, 0.0f);
    __syncthreads();
    float acc = p[kH * kH + j];
    const float* row = p + j * kH;
    for (int k = 0; k < kH; ++k) acc = fmaf(row[k], h[k], acc);
    cur = acc;
    p += kH * kH + kH;
  }
  __syncthreads();
  h[j] = fmaxf(h0 + cur, 0.0f);
  __syncthreads();
  if (j < 16) {
    const float* row = p + j * kH;
    float acc = p[16 * kH + j];
    for (int k = 0; k < kH; ++k) acc = fmaf(row[k], h[k], acc);
    out[j] = acc + ((j / 4) == (j % 4) ? 1.0f : 0.0f);  // + I
  }
  __syncthreads();
  if (j == 0) {
    float* dst = img + (int64_t)n_mob * kH + (int64_t)s * kAffFloats;
    float W[16], Wi[16];
    for (int i = 0; i < 16; ++i) W[i] = out[i];
    for (int i = 0; i < 16; ++i) dst[i] = W[i];
    if (is_rot) {  // host replaces W by U^T V (torch.svd) and fills the transpose; nothing else to do here
      for (int i = 16; i < kAffFloats; ++i) dst[i] = 0.0f;
    } else {
      invert4(W, Wi);
      dst[16] = logf(fabsf(det4f(W)));
      dst[17] = dst[18] = dst[19] = 0.0f;
      for (int i = 0; i < 16; ++i) dst[kAffInv + i] = Wi[i];
      dst[kAffInv + 16] = logf(fabsf(det4f(Wi)));
      dst[kAffInv + 17] = dst[kAffInv + 18] = dst[kAffInv + 19] = 0.0f;
    }
  }
}

}  // namespace

cudaError_t launch_condition(const rnf_flow* f, const float* feat, int64_t B, float* cond, cudaStream_t st,
                             const int32_t* row_index, const int32_t* count) {
  const rnf_model_desc& m = f->model;
  // affine_is_rot == 2: the conditional affine slots belong to ablation layers whose blocks the caller fills (rnf_abi.h)
  const int n_aff_dev = m.affine_is_rot == 2 ? 0 : m.n_affine_slots;
  const int S = m.n_mobius_slots + n_aff_dev;
  if (S == 0 || B == 0) return cudaSuccess;
  const int Ntot = S * kH;
  dim3 grid((Ntot + BN - 1) / BN, (unsigned)((B + BM - 1) / BM));
  hoist_gemm_kernel<<<grid, 256, 0, st>>>(feat, f->weights_dev + m.wf_off, B, Ntot, m.F, cond, f->cond_floats,
                                          m.n_mobius_slots, m.n_affine_slots, row_index, count);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (n_aff_dev > 0) {
    dim3 g2((unsigned)B, m.n_affine_slots);
    cond_affine_kernel<<<g2, 64, 0, st>>>(f->weights_dev + m.caff_off, cond, f->cond_floats, m.n_mobius_slots,
                                          m.n_affine_slots, m.affine_is_rot, count);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace rnf
