// fma_rate.cu -- probe: issue / pipe throughput of FFMA, FFMA2 and their mix on one SM sub-partition (sm_100a).
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int MODE>   // 0: 8 x FFMA, 1: 8 x FFMA2, 2: 4 x FFMA2 + 4 x FFMA (same instruction count), 3: 4 x FFMA2 + 8 x FFMA
__global__ void k(float* out, long long* cyc) {
  float s[8]; u64 p[8];
  for (int i = 0; i < 8; ++i) { s[i] = threadIdx.x * 0.001f + i; p[i] = (u64)__float_as_uint(s[i]) | ((u64)__float_as_uint(s[i] + 1.f) << 32); }
  const float b = 0.999f; const u64 B = (u64)__float_as_uint(b) | ((u64)__float_as_uint(b) << 32);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 4096; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] = fma1(s[i], b, 0.5f);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], B, B);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], B, B); s[i] = fma1(s[i], b, 0.5f); }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], B, B); s[i] = fma1(s[i], b, 0.5f); s[i + 4] = fma1(s[i + 4], b, 0.5f); }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* o; long long* c; long long h;
  cudaMalloc(&o, 1024 * 4 * 4); cudaMalloc(&c, 64);
  const char* names[4] = {"8 FFMA", "8 FFMA2", "4 FFMA2 + 4 FFMA", "4 FFMA2 + 8 FFMA"};
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int m = 0; m < 4; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k<0><<<1, 32 * warps>>>(o, c); else if (m == 1) k<1><<<1, 32 * warps>>>(o, c); else if (m == 2) k<2><<<1, 32 * warps>>>(o, c); else k<3><<<1, 32 * warps>>>(o, c);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps/SM (%d per SMSP)  %-18s %7.2f cycles per loop iteration per SMSP-warp-slot => %.2f cycles/iter/warp\n", warps, warps / 4, names[m], h / 4096.0, h / 4096.0 / (warps / 4));
    }
  return 0;
}
