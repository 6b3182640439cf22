// sync_probe.cu -- what does compute-sanitizer --tool synccheck need to accept flow_t4.cu's hand-over protocol?  An mbarrier
// initialised with a count of 4 on which four producer warps arrive (one lane each) and a service warp waits.  Variants add, one by
// one, what the real kernel has on top: dynamic shared memory with 18 barriers initialised in a loop, 640 threads, setmaxnreg.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/sync_probe.cu -o tools/_build/sync_probe
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
}
// MODE 0: static smem, 160 threads; 1: dynamic smem + loop init of 18 barriers; 2: + 640 threads (4 tiles, 4 service warps); 3: + setmaxnreg;
// 4: as 2 with the roles swapped (service = warps 0..3, workers = warps 4..19): which THREADS does the tool object to?
template <int MODE>
__global__ void __launch_bounds__(MODE >= 2 ? 640 : 160, 1) k(int* out, int rounds) {
  extern __shared__ uint8_t dyn[];
  __shared__ __align__(8) unsigned long long sbar[18];
  __shared__ int data;
  uint8_t* base = MODE == 0 ? reinterpret_cast<uint8_t*>(sbar) : dyn + ((1024u - (smem_u32(dyn) & 1023u)) & 1023u) + 160256;
  const uint32_t bars = smem_u32(base);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = MODE >= 2 ? 4 : 1;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 18; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bars + 8 * i), "r"(i >= 10 ? 4 : 1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    data = 0;
  }
  __syncthreads();
  const bool service = MODE == 4 ? warp < 4 : warp >= 4 * n_tiles;
  const int tile = MODE == 4 ? (service ? warp : (warp - 4) >> 2) : (service ? warp - 4 * n_tiles : warp >> 2);
  if (MODE == 3) {
    if (service) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
  }
  uint32_t par = 0;
  for (int r = 0; r < rounds; ++r) {
    if (!service) {
      __syncwarp();
      if (lane == 0) { atomicAdd(&data, 1); arrive(bars + 8 * (10 + tile)); }
      wait(bars + 8 * (6 + tile), par);
    } else {
      wait(bars + 8 * (10 + tile), par);
      __syncwarp();
      if (lane == 0) arrive(bars + 8 * (6 + tile));
    }
    par ^= 1;
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = data;
}
template <int MODE>
void run(int* o, const char* name) {
  int h = -1;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 170000);
  k<MODE><<<1, MODE >= 2 ? 640 : 160, MODE == 0 ? 0 : 170000>>>(o, 10);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(&h, o, 4, cudaMemcpyDeviceToHost);
  printf("%-50s data %d (%s)\n", name, h, cudaGetErrorString(e));
}
int main() {
  int* o; cudaMalloc(&o, 16);
  run<0>(o, "static smem, 160 threads");
  run<1>(o, "dynamic smem, 18 barriers in a loop");
  run<2>(o, "... 640 threads, 4 tiles + 4 service warps");
  run<3>(o, "... setmaxnreg 112 / 32");
  run<4>(o, "640 threads, roles swapped (service = warps 0..3)");
  return 0;
}
