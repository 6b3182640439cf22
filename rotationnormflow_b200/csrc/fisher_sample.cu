// fisher_sample.cu -- matrix-Fisher base sampling on the device.
//
// Follows utils/fisher.py:117-207 (sample_bingham, sample_matrix_fisher, Kent / Ganeiber / Mardia, arXiv:1310.8110):
//   (U, S, V) = proper_svd(A);  Bingham parameter  lambda = (0, 2(S1+S2), 2(S0+S2), 2(S0+S1));  b = 1.5
//   envelope ACG(Omega), Omega = 1 + 2 lambda / b:  y = eps / sqrt(Omega), eps ~ N(0, I4);  q = y / |y|
//   accept  w < exp(-q' Lambda q) / (M* (q' Omega q)^-2),  M* = exp(-(4 - b)/2) (4/b)^2,  w ~ U(0,1)
//   R = U quat_to_rotmat(q) V^T                                           (utils/fisher.py:14-50,196-205)
// The reference draws 8x oversampled batches with torch's host-seeded generator and keeps the first num_samples accepted
// candidates, image by image in a Python loop; accepted candidates are i.i.d., so one thread per output sample running its own
// rejection loop on a counter-based generator (Philox4x32-10, keyed by seed / image / sample / attempt) draws from the same
// distribution.  Random streams differ from torch's by construction: parity is distributional (tests compare moments against
// the oracle restatement and against importance sampling from the uniform distribution).
#include "rnf_common.cuh"

namespace rnf {
namespace {

struct Philox {
  uint32_t c[4], k[2];
  __device__ __forceinline__ void round() {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0], n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  // 10 rounds of Philox4x32 on (counter, key); the standard key schedule (Weyl constants)
  __device__ __forceinline__ void generate(uint32_t out[4]) {
    uint32_t c0[4] = {c[0], c[1], c[2], c[3]}, k0[2] = {k[0], k[1]};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      round();
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[i] = c[i]; c[i] = c0[i]; }
    k[0] = k0[0]; k[1] = k0[1];
  }
};

__device__ __forceinline__ float u01(uint32_t x) { return ((float)x + 0.5f) * 2.3283064365386963e-10f; }   // (0, 1)

// usv [B][24]: U (9, row-major), S (3, proper), V (9), pad 3
__global__ void fisher_sample_kernel(const float* __restrict__ usv, int64_t B, int64_t n, unsigned long long seed,
                                     float* __restrict__ out) {
  const int64_t total = B * n;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / n;
    const float* p = usv + b * 24;
    const float S0 = __ldg(p + 9), S1 = __ldg(p + 10), S2 = __ldg(p + 11);
    const float lam[4] = {0.0f, 2.0f * (S1 + S2), 2.0f * (S0 + S2), 2.0f * (S0 + S1)};
    const float bb = 1.5f;
    float om[4], sd[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { om[i] = 1.0f + 2.0f * lam[i] / bb; sd[i] = rsqrtf(om[i]); }
    const float log_mstar = -(4.0f - bb) * 0.5f + 2.0f * logf(4.0f / bb);
    Philox rng;
    rng.k[0] = (uint32_t)seed; rng.k[1] = (uint32_t)(seed >> 32);
    rng.c[0] = (uint32_t)idx; rng.c[1] = (uint32_t)((unsigned long long)idx >> 32); rng.c[3] = 0x5EEDu;
    float q[4] = {1.0f, 0.0f, 0.0f, 0.0f};
    for (uint32_t attempt = 0; attempt < 100000u; ++attempt) {     // acceptance >= exp(-(4-b)/2)(4/b)^2 ... ^-1 ~ 0.49 at worst
      uint32_t r0[4], r1[4];
      rng.c[2] = 2u * attempt;      rng.generate(r0);
      rng.c[2] = 2u * attempt + 1u; rng.generate(r1);
      // Box-Muller: four normals from r0, the acceptance uniform from r1
      float y[4];
      {
        const float ra = sqrtf(-2.0f * logf(u01(r0[0]))), rb = sqrtf(-2.0f * logf(u01(r0[2])));
        float sa, ca, sb, cb;
        sincospif(2.0f * u01(r0[1]), &sa, &ca);
        sincospif(2.0f * u01(r0[3]), &sb, &cb);
        y[0] = sd[0] * ra * ca; y[1] = sd[1] * ra * sa; y[2] = sd[2] * rb * cb; y[3] = sd[3] * rb * sb;
      }
      const float inv = rsqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2] + y[3] * y[3]);
      float qa = 0.0f, qo = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        q[i] = y[i] * inv;
        qa = fmaf(lam[i] * q[i], q[i], qa);
        qo = fmaf(om[i] * q[i], q[i], qo);
      }
      // w < exp(-qa) / (M* qo^-2)   <=>   log w < -qa + 2 log qo - log M*
      if (logf(u01(r1[0])) < -qa + 2.0f * logf(qo) - log_mstar) break;
    }
    // quat_to_rotmat (w, x, y, z), then U Q V^T
    const float w = q[0], x = q[1], yy = q[2], z = q[3];
    const float Q[9] = {w * w + x * x - yy * yy - z * z, 2 * x * yy - 2 * w * z, 2 * w * yy + 2 * x * z,
                        2 * w * z + 2 * x * yy, w * w - x * x + yy * yy - z * z, 2 * yy * z - 2 * w * x,
                        2 * x * z - 2 * w * yy, 2 * w * x + 2 * yy * z, w * w - x * x - yy * yy + z * z};
    float UQ[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)
        UQ[3 * i + j] = __ldg(p + 3 * i) * Q[j] + __ldg(p + 3 * i + 1) * Q[3 + j] + __ldg(p + 3 * i + 2) * Q[6 + j];
    float* o = out + idx * 9;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j)   // (UQ V^T)_ij = sum_k UQ_ik V_jk
        o[3 * i + j] = UQ[3 * i] * __ldg(p + 12 + 3 * j) + UQ[3 * i + 1] * __ldg(p + 12 + 3 * j + 1) + UQ[3 * i + 2] * __ldg(p + 12 + 3 * j + 2);
  }
}

}  // namespace

cudaError_t launch_fisher_sample(const float* usv, int64_t B, int64_t n, unsigned long long seed, float* out, cudaStream_t st) {
  const int64_t total = B * n;
  if (total <= 0) return cudaSuccess;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  fisher_sample_kernel<<<(unsigned)blocks, 256, 0, st>>>(usv, B, n, seed, out);
  return cudaGetLastError();
}

}  // namespace rnf
