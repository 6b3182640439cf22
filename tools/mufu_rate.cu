// mufu_rate.cu -- probe: throughput of the SFU (XU pipe) instructions on one SM sub-partition (sm_100a), alone and mixed with FFMA2.
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ float rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float sq(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float lg2(float x) { float r; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MODE>   // 0 rcp, 1 sqrt, 2 ex2, 3 lg2, 4: 8 rcp + 8 FFMA2, 5: 8 rcp + 16 FFMA2
__global__ void k(float* out, long long* cyc) {
  float s[8]; u64 p[8];
  for (int i = 0; i < 8; ++i) { s[i] = 1.0f + threadIdx.x * 0.001f + i; p[i] = (u64)__float_as_uint(s[i]) | ((u64)__float_as_uint(s[i] + 1.f) << 32); }
  const u64 B = (u64)__float_as_uint(0.999f) | ((u64)__float_as_uint(0.999f) << 32);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 2048; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0 || MODE >= 4) s[i] = rcp(s[i]);
      if (MODE == 1) s[i] = sq(s[i]);
      if (MODE == 2) s[i] = ex2(s[i]);
      if (MODE == 3) s[i] = lg2(s[i]);
      if (MODE >= 4) p[i] = fma2(p[i], B, B);
      if (MODE == 5) p[i] = fma2(p[i], B, B);
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 8; ++i) acc += s[i] + __uint_as_float((unsigned)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* o; long long* c; long long h;
  cudaMalloc(&o, 1024 * 4 * 4); cudaMalloc(&c, 64);
  const char* names[6] = {"8 MUFU.RCP", "8 MUFU.SQRT", "8 MUFU.EX2", "8 MUFU.LG2", "8 RCP + 8 FFMA2", "8 RCP + 16 FFMA2"};
  for (int warps = 4; warps <= 16; warps *= 4)
    for (int m = 0; m < 6; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) { case 0: k<0><<<1, 32 * warps>>>(o, c); break; case 1: k<1><<<1, 32 * warps>>>(o, c); break; case 2: k<2><<<1, 32 * warps>>>(o, c); break;
                     case 3: k<3><<<1, 32 * warps>>>(o, c); break; case 4: k<4><<<1, 32 * warps>>>(o, c); break; default: k<5><<<1, 32 * warps>>>(o, c); }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps/SM (%d per SMSP)  %-18s %7.2f cycles per iteration per warp => %.2f cycles per MUFU per SMSP\n", warps, warps / 4, names[m], h / 2048.0,
             h / 2048.0 / (warps / 4) / 8);
    }
  return 0;
}
