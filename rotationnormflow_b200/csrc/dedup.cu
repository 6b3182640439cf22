// dedup.cu -- the reference's calling convention hands Flow.forward a row-aligned `feature [N,F]` built by `.repeat`
// (agent.py:240-244, eval.py:450): millions of identical rows per image.  These kernels recover, on the device and without a
// host synchronisation, the run structure of that tensor:
//     idx[i]   = number of row changes before row i        (row -> image index, the feat_index of rnf_flow_forward)
//     first[b] = first row of run b                         (the row the per-image conditioner reads)
//     count    = number of runs
// in one streaming pass over the N x F floats (HBM-bound: the 4 GB of a 500 000 x 2048 chunk are read 1 + 1/8 times), a
// device-wide inclusive scan of the N change flags (CUB) and a scatter of the run starts.
#include <cub/device/device_scan.cuh>

#include "rnf_common.cuh"

namespace rnf {
namespace {

// flags[r] = (row r differs from row r-1), flags[0] = 0; the flags are zeroed before the launch and only ever set.
// Work item = a tile of kRB consecutive rows x 128 consecutive float4 columns (32 lanes x 4), handed to the warps in row-major
// tile order with a grid stride: at any moment the resident warps stream ONE contiguous window of the tensor (HBM pages and
// channels are visited evenly; a static split into one far-apart row range per warp made thousands of streams with identical
// alignment camp on the same channels: 0.3-0.9 TB/s).  A lane keeps the previous row of its four columns in registers, so an
// element is read once plus 1/kRB; the (kRB + 1) x 4 loads of a tile are independent and issued back to back.
constexpr int kRB = 8;

__global__ void __launch_bounds__(256) row_change_kernel(const float* __restrict__ feat, int64_t N, int64_t F4, int64_t col_tiles,
                                                         int64_t n_tiles, int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const float4* f4 = reinterpret_cast<const float4*>(feat);
  for (int64_t t = warp0; t < n_tiles; t += n_warps) {
    const int64_t rb = t / col_tiles, ct = t - rb * col_tiles;
    const int64_t r0 = rb * kRB;
    const int64_t c0 = ct * 128 + lane;
    float4 v[kRB + 1][4];
#pragma unroll
    for (int i = 0; i <= kRB; ++i) {
      int64_t r = r0 - 1 + i;
      r = r < 0 ? 0 : (r < N ? r : N - 1);               // row -1 reads row 0 (no change); rows past the end repeat the last
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t c = c0 + 32 * k;
        v[i][k] = c < F4 ? __ldg(f4 + r * F4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 1; i <= kRB; ++i) {
      bool diff = false;
#pragma unroll
      for (int k = 0; k < 4; ++k)    // value comparison (what torch's != does), not a bit-pattern comparison
        diff |= (v[i][k].x != v[i - 1][k].x) | (v[i][k].y != v[i - 1][k].y) | (v[i][k].z != v[i - 1][k].z) | (v[i][k].w != v[i - 1][k].w);
      const int64_t r = r0 - 1 + i;
      if (__any_sync(0xffffffffu, diff) && lane == 0 && r < N) flags[r] = 1;
    }
  }
}

// rows that are not 16-byte tileable (F % 4 != 0 or a misaligned base): the same tiling with scalar loads
__global__ void __launch_bounds__(256) row_change_scalar_kernel(const float* __restrict__ feat, int64_t N, int64_t F, int64_t col_tiles,
                                                                int64_t n_tiles, int32_t* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = warp0; t < n_tiles; t += n_warps) {
    const int64_t rb = t / col_tiles, ct = t - rb * col_tiles;
    const int64_t r0 = rb * kRB;
    const int64_t c0 = ct * 128 + lane;
    float v[kRB + 1][4];
#pragma unroll
    for (int i = 0; i <= kRB; ++i) {
      int64_t r = r0 - 1 + i;
      r = r < 0 ? 0 : (r < N ? r : N - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t c = c0 + 32 * k;
        v[i][k] = c < F ? __ldg(feat + r * F + c) : 0.f;
      }
    }
#pragma unroll
    for (int i = 1; i <= kRB; ++i) {
      bool diff = false;
#pragma unroll
      for (int k = 0; k < 4; ++k) diff |= v[i][k] != v[i - 1][k];
      const int64_t r = r0 - 1 + i;
      if (__any_sync(0xffffffffu, diff) && lane == 0 && r < N) flags[r] = 1;
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
  cudaError_t e = cudaMemsetAsync(idx, 0, sizeof(int32_t) * (size_t)N, st);
  if (e != cudaSuccess) return e;
  const bool vec_ok = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(feat) & 15) == 0);
  const int64_t cols = vec_ok ? F / 4 : F;                          // float4 or float columns
  const int64_t col_tiles = (cols + 127) / 128, row_blocks = (N + kRB - 1) / kRB;
  const int64_t n_tiles = col_tiles * row_blocks;
  const int64_t want = (n_tiles + 7) / 8, cap_blocks = (int64_t)sm_count * 8;     // 8 CTAs x 8 warps per SM at most
  const unsigned blocks = (unsigned)(want < cap_blocks ? want : cap_blocks);
  if (vec_ok) row_change_kernel<<<blocks, 256, 0, st>>>(feat, N, cols, col_tiles, n_tiles, idx);
  else row_change_scalar_kernel<<<blocks, 256, 0, st>>>(feat, N, cols, col_tiles, n_tiles, idx);
  e = cudaGetLastError();
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
