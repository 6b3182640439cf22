// mma_rate.cu -- standalone probe: cycles per tcgen05.mma (M = 128, K = 16, kind::f16) as a function of N, for the A operand
// in shared memory (SS) and in tensor memory (TS).  One CTA, one issuing thread, batches of 64 MMAs per commit.
#include <cstdio>
#include <cstdlib>
#include "../rotationnormflow_b200/csrc/tc_common.cuh"
using namespace rnf;

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint32_t b_lo32, uint32_t hi32, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 db;\nsetp.ne.b32 p, %5, 0;\nmov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n}\n" ::"r"(d), "r"(a_tmem), "r"(b_lo32), "r"(hi32), "r"(idesc), "r"(acc)
      : "memory");
}

template <int N, bool TS>
__device__ long long run(uint8_t* smem, uint32_t tmem, uint32_t bar, uint32_t& parity) {
  const uint32_t a_d = umma_desc_lo(smem_u32(smem)), b_d = umma_desc_lo(smem_u32(smem + 16384));
  const uint32_t idesc = umma_idesc(128, N);
  long long t0 = 0, t1 = 0;
  __syncthreads();
  if (threadIdx.x == 0) t0 = clock64();
  for (int rep = 0; rep < 16; ++rep) {
    if (threadIdx.x < 32) {
      if (elect_one_sync()) {
#pragma unroll 1
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (TS) umma_ts(tmem + 256, tmem + 8 * k, b_d + 2 * k, kDescHi, idesc, 1);
            else umma_f16(tmem + 256, a_d + 2 * k, b_d + 2 * k, kDescHi, idesc, 1);
          }
        }
        umma_commit(bar);
      }
      __syncwarp();
    }
    mbar_wait(bar, parity);
    parity ^= 1;
  }
  if (threadIdx.x == 0) t1 = clock64();
  return t1 - t0;
}

__global__ void __launch_bounds__(128, 1) probe(long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) unsigned long long s_bar;
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&s_bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem, bar = smem_u32(&s_bar);
  uint32_t parity = 0;
  long long r[10];
  r[0] = run<32, false>(smem, tmem, bar, parity);
  r[1] = run<64, false>(smem, tmem, bar, parity);
  r[2] = run<128, false>(smem, tmem, bar, parity);
  r[3] = run<256, false>(smem, tmem, bar, parity);
  r[4] = run<32, true>(smem, tmem, bar, parity);
  r[5] = run<64, true>(smem, tmem, bar, parity);
  r[6] = run<128, true>(smem, tmem, bar, parity);
  r[7] = run<256, true>(smem, tmem, bar, parity);
  r[8] = run<16, true>(smem, tmem, bar, parity);
  r[9] = run<64, true>(smem, tmem, bar, parity);
  if (threadIdx.x == 0) for (int i = 0; i < 10; ++i) out[i] = r[i];
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory"); }
}

int main() {
  long long* d; long long h[10];
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 32768 + 1024);
  for (int it = 0; it < 2; ++it) {
    probe<<<1, 128, 16384 + 32768 + 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
  }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[10] = {"SS N=32", "SS N=64", "SS N=128", "SS N=256", "TS N=32", "TS N=64", "TS N=128", "TS N=256", "TS N=16", "TS N=64 (again)"};
  for (int i = 0; i < 10; ++i) printf("%-16s %8.1f cycles per MMA (16 batches x 64 MMAs incl. commit + wait)\n", names[i], h[i] / 1024.0);
  return 0;
}
