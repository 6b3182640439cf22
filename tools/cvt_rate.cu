// cvt_rate.cu -- probe: throughput of the fp32 -> packed fp16 conversions and of the mixed-precision FHFMA used by the GEMM epilogue.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t cvt_rz(float a, float b) { uint32_t r; asm volatile("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint32_t cvt_rn(float a, float b) { uint32_t r; asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fhfma(uint32_t h, float c) {
  float d;
  asm volatile("{\n.reg .b16 h0, h1, m1;\nmov.b32 {h0, h1}, %1;\nmov.b16 m1, 0xBC00;\nfma.rn.f32.f16 %0, h0, m1, %2;\n}\n" : "=f"(d) : "r"(h), "f"(c));
  return d;
}
template <int MODE>   // 0: 8 cvt.rz, 1: 8 cvt.rn, 2: 8 fhfma, 3: full split of 8 pairs (2 cvt + 2 fhfma each)
__global__ void k(float* out, long long* cyc) {
  float s[16]; uint32_t u[8];
  for (int i = 0; i < 16; ++i) s[i] = 1.0f + threadIdx.x * 0.001f + i;
  for (int i = 0; i < 8; ++i) u[i] = 0x3C003C00u + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 2048; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { u[i] = cvt_rz(s[2 * i], s[2 * i + 1]); s[2 * i] += __uint_as_float(u[i] & 0x3f800000u); }
      if (MODE == 1) { u[i] = cvt_rn(s[2 * i], s[2 * i + 1]); s[2 * i] += __uint_as_float(u[i] & 0x3f800000u); }
      if (MODE == 2) { s[2 * i] = fhfma(u[i], s[2 * i]); }
      if (MODE == 3) { uint32_t hi = cvt_rz(s[2 * i], s[2 * i + 1]); float d0 = fhfma(hi, s[2 * i]), d1 = fhfma(hi >> 16, s[2 * i + 1]); u[i] = cvt_rn(d0, d1); s[2 * i] += __uint_as_float((hi ^ u[i]) & 0x3f800000u); }
    }
  }
  long long t1 = clock64();
  float acc = 0; for (int i = 0; i < 16; ++i) acc += s[i]; for (int i = 0; i < 8; ++i) acc += __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* o; long long* c; long long h;
  cudaMalloc(&o, 1024 * 4 * 4); cudaMalloc(&c, 64);
  const char* names[4] = {"8 cvt.rz.f16x2 (+8 LOP +8 FADD)", "8 cvt.rn.f16x2 (+8 LOP +8 FADD)", "8 FHFMA", "8 x full split (2 cvt + 2 FHFMA + ...)"};
  for (int warps = 4; warps <= 16; warps *= 4)
    for (int m = 0; m < 4; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) { case 0: k<0><<<1, 32 * warps>>>(o, c); break; case 1: k<1><<<1, 32 * warps>>>(o, c); break; case 2: k<2><<<1, 32 * warps>>>(o, c); break; default: k<3><<<1, 32 * warps>>>(o, c); }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
      printf("%2d warps/SM (%d per SMSP)  %-40s %7.2f cycles per iteration per warp => %.2f cycles per group of the 8 per SMSP\n", warps, warps / 4, names[m], h / 2048.0,
             h / 2048.0 / (warps / 4) / 8);
    }
  return 0;
}
