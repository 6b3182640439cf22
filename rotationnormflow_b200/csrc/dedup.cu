// dedup.cu -- the reference's calling convention hands Flow.forward a row-aligned `feature [N,F]` built by `.repeat`
// (agent.py:240-244, eval.py:450): millions of identical rows per image.  These kernels recover, on the device and without a
// host synchronisation, the run structure of that tensor:
//     idx[i]   = number of row changes before row i        (row -> image index, the feat_index of rnf_flow_forward)
//     first[b] = first row of run b                         (the row the per-image conditioner reads)
//     count    = number of runs
// in one streaming pass over the N x F floats (HBM-bound: the 4 GB of a 500 000 x 2048 chunk are read once, ~0.65 ms), a
// device-wide inclusive scan of the N change flags (CUB) and a scatter of the run starts.
#include <cub/device/device_scan.cuh>

#include "rnf_common.cuh"

namespace rnf {
namespace {

// One warp per contiguous range of rows; a lane keeps the previous row's values of its columns in registers, so every element
// is read once (plus one extra row per range).  flags[i] = (row i differs from row i-1), flags[0] = 0.
template <int VEC>   // float4 loads per lane and row chunk kept in registers
__global__ void __launch_bounds__(256) row_change_kernel(const float* __restrict__ feat, int64_t N, int64_t F, int64_t rows_per_warp,
                                                         int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t r0 = warp * rows_per_warp;
  if (r0 >= N) return;
  const int64_t r1 = r0 + rows_per_warp < N ? r0 + rows_per_warp : N;
  const bool vec_ok = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
  for (int64_t r = r0 + lane; r < r1; r += 32) flags[r] = 0;
  __syncwarp();
  // column chunks of 32 lanes x VEC float4 (or scalars when the row is not 16-byte tileable)
  if (vec_ok) {
    const int64_t F4 = F / 4;
    for (int64_t c0 = 0; c0 < F4; c0 += 32 * VEC) {
      float4 prev[VEC];
      const int64_t rp = r0 > 0 ? r0 - 1 : 0;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const int64_t c = c0 + lane + 32 * v;
        prev[v] = c < F4 ? __ldg(reinterpret_cast<const float4*>(feat + rp * F) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int64_t r = r0; r < r1; ++r) {
        bool diff = false;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const int64_t c = c0 + lane + 32 * v;
          if (c < F4) {
            const float4 cur = __ldg(reinterpret_cast<const float4*>(feat + r * F) + c);
            // bit pattern comparison is not what torch's != does for NaN / signed zero; value comparison is
            diff |= (cur.x != prev[v].x) | (cur.y != prev[v].y) | (cur.z != prev[v].z) | (cur.w != prev[v].w);
            prev[v] = cur;
          }
        }
        if (__any_sync(0xffffffffu, diff) && lane == 0 && r > 0) flags[r] = 1;
      }
    }
  } else {
    for (int64_t c0 = 0; c0 < F; c0 += 32) {
      const int64_t c = c0 + lane;
      float prev = (c < F) ? __ldg(feat + (r0 > 0 ? r0 - 1 : 0) * F + c) : 0.f;
      for (int64_t r = r0; r < r1; ++r) {
        bool diff = false;
        if (c < F) {
          const float cur = __ldg(feat + r * F + c);
          diff = cur != prev;
          prev = cur;
        }
        if (__any_sync(0xffffffffu, diff) && lane == 0 && r > 0) flags[r] = 1;
      }
    }
  }
}

// idx (inclusive scan of the flags, in place) -> run starts and run count; idx clamped to cap - 1 so that an overflowing run
// count can never index outside the per-image buffers (the caller checks `count` against `cap`, see rnf_abi.h)
__global__ void run_starts_kernel(int32_t* __restrict__ idx, const int32_t* __restrict__ flags_unused, int64_t N, int32_t* __restrict__ first,
                                  int64_t cap, int32_t* __restrict__ count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int32_t me = idx[i];
  const int32_t before = i > 0 ? idx[i - 1] : -1;
  if (i == N - 1) *count = me + 1;
  if (me != before && me < cap) first[me] = (int32_t)i;
}
__global__ void clamp_idx_kernel(int32_t* __restrict__ idx, int64_t N, int32_t hi) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && idx[i] > hi) idx[i] = hi;
}

__global__ void poison_kernel(const int32_t* __restrict__ count, int64_t cap, float* __restrict__ ldj, int64_t N) {
  if (*count <= cap) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) ldj[i] = __int_as_float(0x7fc00000);
}

}  // namespace

size_t dedup_scan_bytes(int64_t N) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(N > 0 ? N : 1));
  return (bytes + 255) & ~(size_t)255;
}

cudaError_t launch_dedup(const float* feat, int64_t N, int64_t F, int32_t* idx, int32_t* first, int64_t cap, int32_t* count,
                         void* ws, size_t ws_bytes, int sm_count, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  if (N > 0x7fffffffLL) return cudaErrorInvalidValue;
  // flags live in idx (the scan runs in place); ws holds the scan's temporary storage
  const int64_t warps = (int64_t)sm_count * 32;                     // 4 CTAs x 8 warps per SM: enough loads in flight for HBM
  const int64_t rows_per_warp = (N + warps - 1) / warps;
  const int64_t n_warps = (N + rows_per_warp - 1) / rows_per_warp;
  const unsigned blocks = (unsigned)((n_warps + 7) / 8);
  row_change_kernel<4><<<blocks, 256, 0, st>>>(feat, N, F, rows_per_warp, idx);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  size_t need = ws_bytes;
  e = cub::DeviceScan::InclusiveSum(ws, need, (const int32_t*)idx, idx, (int)N, st);
  if (e != cudaSuccess) return e;
  const unsigned nb = (unsigned)((N + 255) / 256);
  if (cudaMemsetAsync(first, 0, sizeof(int32_t) * (size_t)cap, st) != cudaSuccess) return cudaGetLastError();
  run_starts_kernel<<<nb, 256, 0, st>>>(idx, nullptr, N, first, cap, count);
  clamp_idx_kernel<<<nb, 256, 0, st>>>(idx, N, (int32_t)(cap - 1));
  return cudaGetLastError();
}

cudaError_t launch_poison(const int32_t* count, int64_t cap, float* ldj, int64_t N, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  poison_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(count, cap, ldj, N);
  return cudaGetLastError();
}

}  // namespace rnf
