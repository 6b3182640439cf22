// ts_probe.cu -- standalone probe: tcgen05.mma with the A operand in tensor memory (TS form), B in shared memory.
// Verifies the A-in-TMEM layout assumed by flow_t4.cu: row r <-> lane r, K elements 2c / 2c+1 in the low / high half of column c.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../rotationnormflow_b200/csrc/tc_common.cuh"
using namespace rnf;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t desc_hi32, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A [128][64] fp16 (row-major global), B image = K-major SW128 [64 x 64] fp16 bytes, D out [128][64] fp32
__global__ void __launch_bounds__(128, 1) probe(const __half* A, const uint8_t* Bimg, float* D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8192 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(Bimg)[i];
  fence_proxy_async();
  if (tid == 0) { mbar_init(smem_u32(&s_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = s_tmem;
  const uint32_t tm = base + ((uint32_t)(warp * 32) << 16);
  // my row of A: 64 fp16 = 32 packed words -> TMEM columns 0..31 of my lane
  float packed[32];
  for (int c = 0; c < 32; ++c) {
    __half2 h = __halves2half2(A[tid * 64 + 2 * c], A[tid * 64 + 2 * c + 1]);
    packed[c] = __uint_as_float(*reinterpret_cast<uint32_t*>(&h));
  }
  tmem_st32(tm, packed);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t b = umma_desc_lo(smem_u32(smem));
      for (int k = 0; k < 4; ++k) umma_f16_ts(base + 64, base + 8 * k, b + 2 * k, kDescHi, umma_idesc(128, 64), k > 0);
      umma_commit(smem_u32(&s_bar));
    }
    __syncwarp();
  }
  mbar_wait(smem_u32(&s_bar), 0);
  tc_fence_after();
  float acc[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tm + 64 + 32 * h, acc);
    for (int j = 0; j < 32; ++j) D[tid * 64 + 32 * h + j] = acc[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory"); }
}

int main() {
  std::vector<__half> A(128 * 64);
  std::vector<float> Af(128 * 64), Bf(64 * 64), ref(128 * 64, 0.f), out(128 * 64);
  std::vector<uint8_t> img(8192, 0);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { Af[i] = (float)(rand() % 17 - 8); A[i] = __float2half(Af[i]); }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) {
      Bf[n * 64 + k] = (float)(rand() % 15 - 7);
      const int off = (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
      *reinterpret_cast<__half*>(&img[off]) = __float2half(Bf[n * 64 + k]);
    }
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 64; ++k) ref[r * 64 + n] += Af[r * 64 + k] * Bf[n * 64 + k];
  __half* dA; uint8_t* dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, 8192); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, img.data(), 8192, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 + 1024);
  probe<<<1, 128, 8192 + 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int i = 0; i < 128 * 64; ++i) { double d = fabs(out[i] - ref[i]); if (d > maxerr) maxerr = d; bad += d > 0.5; }
  printf("TS-form MMA: max |err| = %g, mismatches = %d / %d\n", maxerr, bad, 128 * 64);
  printf("row0: out %g %g %g %g | ref %g %g %g %g\n", out[0], out[1], out[2], out[3], ref[0], ref[1], ref[2], ref[3]);
  return bad != 0;
}
