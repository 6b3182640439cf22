// pipe_rate.cu -- probe: reciprocal throughput (cycles per warp instruction per SM sub-partition at 8 warps per scheduler) of the
// instruction classes the flow kernels are made of, alone and in the mixes that matter (which of them share a pipe?), sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/pipe_rate.cu -o tools/_build/pipe_rate
#include <cstdio>
#include <cstdint>
typedef unsigned long long u64;
#define DI __device__ __forceinline__
DI u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
DI u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DI u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
DI float ffma(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
DI float rcp(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DI float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
DI uint32_t f2fp_rz(float a, float b) { uint32_t r; asm volatile("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
DI uint32_t f2fp_rn(float a, float b) { uint32_t r; asm volatile("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
DI float h2f_lo(uint32_t h) { float r; asm volatile("{.reg .b16 l, u; mov.b32 {l, u}, %1; cvt.f32.f16 %0, l;}" : "=f"(r) : "r"(h)); return r; }
DI uint32_t lop(uint32_t a, uint32_t b) { uint32_t r; asm volatile("and.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DI float fmnmx(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
DI float fsel(float a, float b, float c) { float r; asm volatile("{.reg .pred p; setp.lt.f32 p, %3, 0f3F000000; selp.f32 %0, %1, %2, p;}" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

enum { FFMA2_RRR, FFMA2_IMM, FMUL2, FADD2, FFMA, MUFU_RCP, MUFU_EX2, F2FP_RZ, F2FP_RN, H2F, LOP, FMNMX, SETP_SEL,
       MIX_MUFU_FFMA2_1_4, MIX_MUFU_FFMA2_1_7, MIX_F2FP_MUFU, MIX_H2F_FFMA2, MIX_F2FP_FFMA2, MIX_LOP_FFMA2, MIX_EPI_CUR, MIX_EPI_MASK, N_MODES };
const char* kNames[N_MODES] = {"FFMA2 r,r,r", "FFMA2 r,imm,imm", "FMUL2", "FADD2", "FFMA", "MUFU.RCP", "MUFU.EX2", "F2FP.rz.relu (pack 2)", "F2FP.rn.relu (pack 2)",
                               "HADD2.F32 (f16->f32)", "LOP3", "FMNMX", "FSETP+FSEL", "1 MUFU + 4 FFMA2", "1 MUFU + 7 FFMA2", "1 F2FP + 1 MUFU", "1 H2F + 1 FFMA2",
                               "1 F2FP + 1 FFMA2", "1 LOP3 + 1 FFMA2", "epilogue pair, current (F2FP, 2 H2F, FADD2, F2FP)", "epilogue pair, mask (F2FP, 2 LOP3, FADD2, F2FP)"};
const int kInstr[N_MODES] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 5, 8, 2, 2, 2, 2, 5, 5};

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float s[8]; u64 p[8]; uint32_t w[8];
  for (int i = 0; i < 8; ++i) {
    s[i] = 1.0f + threadIdx.x * 0.001f + i;
    p[i] = (u64)__float_as_uint(s[i]) | ((u64)__float_as_uint(s[i] + 1.f) << 32);
    w[i] = 0x3C003C00u + i + threadIdx.x;
  }
  const u64 B = (u64)__float_as_uint(0.999f) | ((u64)__float_as_uint(0.999f) << 32);
  u64 C = (u64)__float_as_uint(1e-3f + threadIdx.x) | ((u64)__float_as_uint(2e-3f) << 32);
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == FFMA2_RRR) p[i] = fma2(p[i], B, C);
      if (MODE == FFMA2_IMM) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(0x3F7FBE773F7FBE77ull));
      if (MODE == FMUL2) p[i] = mul2(p[i], B);
      if (MODE == FADD2) p[i] = add2(p[i], C);
      if (MODE == FFMA) s[i] = ffma(s[i], 0.999f, 1e-3f);
      if (MODE == MUFU_RCP) s[i] = rcp(s[i]);
      if (MODE == MUFU_EX2) s[i] = ex2(s[i]);
      if (MODE == F2FP_RZ) w[i] = f2fp_rz(s[i], __uint_as_float(w[i]));
      if (MODE == F2FP_RN) w[i] = f2fp_rn(s[i], __uint_as_float(w[i]));
      if (MODE == H2F) s[i] = h2f_lo(__float_as_uint(s[i]));
      if (MODE == LOP) w[i] = lop(w[i], 0xFFFFE000u + i);
      if (MODE == FMNMX) s[i] = fmnmx(s[i], 0.5f + i);
      if (MODE == SETP_SEL) s[i] = fsel(s[i], 0.25f, s[(i + 1) & 7]);
      if (MODE == MIX_MUFU_FFMA2_1_4) { s[i] = rcp(s[i]); p[i] = fma2(p[i], B, C); p[i] = fma2(p[i], B, C); p[i] = fma2(p[i], B, C); p[i] = fma2(p[i], B, C); }
      if (MODE == MIX_MUFU_FFMA2_1_7) { s[i] = rcp(s[i]); for (int r = 0; r < 7; ++r) p[i] = fma2(p[i], B, C); }
      if (MODE == MIX_F2FP_MUFU) { w[i] = f2fp_rz(s[i], __uint_as_float(w[i])); s[i] = rcp(s[i]); }
      if (MODE == MIX_H2F_FFMA2) { s[i] = h2f_lo(__float_as_uint(s[i])); p[i] = fma2(p[i], B, C); }
      if (MODE == MIX_F2FP_FFMA2) { w[i] = f2fp_rz(s[i], __uint_as_float(w[i])); p[i] = fma2(p[i], B, C); }
      if (MODE == MIX_LOP_FFMA2) { w[i] = lop(w[i], 0xFFFFE000u + i); p[i] = fma2(p[i], B, C); }
      if (MODE == MIX_EPI_CUR) {
        const float x0 = __uint_as_float((uint32_t)p[i]), x1 = __uint_as_float((uint32_t)(p[i] >> 32));
        const uint32_t hi = f2fp_rz(x1, x0);
        float b0, b1;
        asm volatile("{.reg .b16 l, u; mov.b32 {l, u}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, u;}" : "=f"(b0), "=f"(b1) : "r"(hi));
        u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(p[i]), "l"((u64)__float_as_uint(b0) | ((u64)__float_as_uint(b1) << 32)));
        const uint32_t lo = f2fp_rn(__uint_as_float((uint32_t)(d >> 32)), __uint_as_float((uint32_t)d));
        w[i] ^= hi + lo;
      }
      if (MODE == MIX_EPI_MASK) {
        const uint32_t x0 = (uint32_t)p[i], x1 = (uint32_t)(p[i] >> 32);
        const uint32_t hi = f2fp_rz(__uint_as_float(x1), __uint_as_float(x0));
        const uint32_t m0 = lop(x0, 0xFFFFE000u), m1 = lop(x1, 0xFFFFE000u);
        u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(p[i]), "l"((u64)m0 | ((u64)m1 << 32)));
        const uint32_t lo = f2fp_rn(__uint_as_float((uint32_t)(d >> 32)), __uint_as_float((uint32_t)d));
        w[i] ^= hi + lo;
      }
    }
  }
  const long long t1 = clock64();
  float acc = 0;
  for (int i = 0; i < 8; ++i) acc += s[i] + __uint_as_float((unsigned)p[i]) + __uint_as_float(w[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + __uint_as_float((unsigned)C);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(float* o, long long* c) {
  const int iters = 2048;
  for (int wps = 1; wps <= 8; wps *= 8) {
    long long h;
    for (int rep = 0; rep < 2; ++rep) { k<MODE><<<1, 128 * wps>>>(o, c, iters); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double per_group = (double)h / iters / 8.0 / wps;     // cycles per (group of kInstr instructions) per SMSP
    printf("%-52s %d warps/SMSP: %6.2f cycles per group of %d (%5.2f per instruction)\n", kNames[MODE], wps, per_group, kInstr[MODE], per_group / kInstr[MODE]);
  }
}
template <int M> struct Loop { static void go(float* o, long long* c) { run<M>(o, c); Loop<M + 1>::go(o, c); } };
template <> struct Loop<N_MODES> { static void go(float*, long long*) {} };
int main() {
  float* o; long long* c;
  cudaMalloc(&o, 2048 * 4); cudaMalloc(&c, 64);
  Loop<0>::go(o, c);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
