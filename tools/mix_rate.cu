// mix_rate.cu -- probe: throughput of the REAL mixture arithmetic (mobius_pair.cuh: mixture_pairs<NP, true>) and of the epilogue
// split (relu_split_pair) on one SM as a function of the number of warps per scheduler, inputs from registers (no TMEM, no MMA).
// Answers: how many cycles per mixture pair / per activation pair does an SM sub-partition need when nothing else runs?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rotationnormflow_b200/csrc tools/mix_rate.cu -o tools/_build/mix_rate
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include "mobius_pair.cuh"
using namespace rnf;

__device__ __forceinline__ void relu_split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float d0, d1;
  upk(sub2(pk(x0, x1), pk(back.x, back.y)), d0, d1);
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}

template <int MODE>   // 0: mixture NP=1, 1: NP=2, 2: NP=4, 3: epilogue split of 32 values
__global__ void k(float* out, long long* cyc, int iters) {
  Plane P;
  const float ph = threadIdx.x * 0.01f;
  P.r[0] = cosf(ph); P.r[1] = sinf(ph); P.r[2] = 0.f;
  P.v[0] = -sinf(ph); P.v[1] = cosf(ph); P.v[2] = 0.f;
  float raw[32];
  for (int i = 0; i < 32; ++i) raw[i] = 0.1f * ((i * 7 + threadIdx.x) % 13) - 0.6f;
  f32x2 S0 = 0ull, S1 = 0ull, S2 = 0ull;
  uint32_t acc_u = 0;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) mixture_pairs<1, true>(P, -1.0f, 0.f, raw + 8 * j, S0, S1, S2);
    } else if (MODE == 1) {
      mixture_pairs<2, true>(P, -1.0f, 0.f, raw, S0, S1, S2);
      mixture_pairs<2, true>(P, -1.0f, 0.f, raw + 16, S0, S1, S2);
    } else if (MODE == 2) {
      mixture_pairs<4, true>(P, -1.0f, 0.f, raw, S0, S1, S2);
    } else {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        uint32_t hi, lo;
        relu_split_pair(raw[2 * e], raw[2 * e + 1], hi, lo);
        acc_u ^= hi + lo;
      }
    }
    // perturb the inputs so nothing is hoisted (cheap: one FADD per value, counted as overhead below)
#pragma unroll
    for (int i = 0; i < 32; ++i) raw[i] += (MODE == 3) ? 1e-3f : 1e-6f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = hsum(S0) + hsum(S1) + hsum(S2) + __uint_as_float(acc_u);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float* o; long long* c; long long h;
  cudaMalloc(&o, 2048 * 4); cudaMalloc(&c, 64);
  const int iters = 4096;
  const char* names[4] = {"mixture NP=1 (4 pairs/iter)", "mixture NP=2 (4 pairs/iter)", "mixture NP=4 (4 pairs/iter)", "epilogue split (16 pairs/iter)"};
  for (int m = 0; m < 4; ++m)
    for (int wps = 1; wps <= 8; wps *= 2) {
      const int threads = 128 * wps;
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) { case 0: k<0><<<1, threads>>>(o, c, iters); break; case 1: k<1><<<1, threads>>>(o, c, iters); break;
                     case 2: k<2><<<1, threads>>>(o, c, iters); break; default: k<3><<<1, threads>>>(o, c, iters); }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      const double per_iter = (double)h / iters;
      const int units = m == 3 ? 16 : 4;
      printf("%-32s %d warps/SMSP: %8.1f cycles/iter/warp -> %6.1f cycles per pair per SMSP (incl. 32 FADD overhead/iter)\n", names[m], wps, per_iter,
             per_iter / (units * wps));
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
