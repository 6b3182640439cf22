// mobius_pair.cuh -- the Mobius mixture arithmetic of mobius_fast.cuh, two components per instruction.
//
// sm_100 has packed FP32 arithmetic (PTX add / mul / fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2: one issue slot, two IEEE
// fp32 results; the second and third source may be a broadcast scalar register or a broadcast immediate).  The tcgen05 flow
// kernels are bound by instruction issue (ncu: issue slots 75 % active, FMA + ALU + XU pipes 50 / 30 / 44 %), and 2/3 of their
// instructions are this mixture, so components are evaluated in PAIRS: everything that is a multiply / add runs packed, only
// the SFU operations (ex2, lg2, sqrt, rcp), min / max and the selects stay scalar.  56 -> ~37 instructions per component.
//
// Column layout of one pair (the host permutes the fc_last rows accordingly, engine._last_layer_perm_pairs): 8 consecutive
// accumulator columns  (t_a, t_b, wx_a, wx_b, wy_a, wy_b, wz_a, wz_b)  with t = logit * log2(e) (the factor is folded into
// the weights), so that a tcgen05.ld vector delivers every packed operand as an aligned register pair.
// Prepared parameters of a pair (inverse direction, kept in TMEM for the bisection):
//   (-alpha'_a, -alpha'_b, -beta'_a, -beta'_b, 1 - |w'_a|^2, 1 - |w'_b|^2, weight_a, weight_b).
#pragma once
#include "mobius_fast.cuh"

namespace rnf {

typedef unsigned long long f32x2;   // two fp32 in a 64-bit register pair: low word = component a, high word = component b

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 bc(float x) { return pk(x, x); }
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// sign flip of both halves: a bit operation that ptxas folds into the negate modifier of the consuming FFMA2 / FADD2
__device__ __forceinline__ f32x2 neg2(f32x2 a) { return a ^ 0x8000000080000000ull; }
__device__ __forceinline__ float hsum(f32x2 v) {
  float lo, hi;
  upk(v, lo, hi);
  return lo + hi;
}
#define RNF_MAP2(out, in, fn) do { float l_, h_; upk(in, l_, h_); out = pk(fn(l_), fn(h_)); } while (0)

// Angle of a UNIT vector from its nearer axis, for a pair: asin(m) with m = min(|h.r|, |h.v|) in [0, 1/sqrt 2].  The Mobius map
// sends the unit circle to itself, so |h| = 1 (to fp32 rounding) and atan(min / max) = asin(min): no reciprocal on the SFU --
// the XU pipe (8 cycles per warp instruction, tools/mufu_rate.cu) is the tightest unit of the mixture -- and a shorter
// polynomial: asin(m) = m + m s P(s), s = m^2, P of degree 6 (minimax fit, max abs error 4.4e-8 in fp32 Horner form,
// tests/test_fastmath.py).
__device__ __forceinline__ f32x2 asin_unit2(f32x2 m) {
  const f32x2 s = mul2(m, m);
  f32x2 p = bc(0.12371734529733658f);
  p = fma2(p, s, bc(-0.11529727280139923f));
  p = fma2(p, s, bc(0.09339626878499985f));
  p = fma2(p, s, bc(0.01043224148452282f));
  p = fma2(p, s, bc(0.04762402921915054f));
  p = fma2(p, s, bc(0.07478351145982742f));
  p = fma2(p, s, bc(0.16667234897613525f));
  return fma2(mul2(p, s), m, m);
}

// NP pairs of mixture components, stage by stage.  raw[8 NP]: accumulator columns in the pair layout above.
// FWD: accumulates (sum w, sum w theta, sum w f) as packed partial sums (a-components in the low, b in the high word).
// !FWD: accumulates sum w and overwrites raw with the prepared parameters.
template <int NP, bool FWD>
__device__ __forceinline__ void mixture_pairs(const Plane& P, float zr, float zv, float* raw, f32x2& S_sp, f32x2& S_th, f32x2& S_f) {
  f32x2 sp[NP], nal[NP], nbe[NP], omw[NP];
  {
    f32x2 t[NP], e[NP], a[NP], b[NP], n2[NP], rt[NP], big[NP], small[NP], ns[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      t[j] = pk(raw[8 * j], raw[8 * j + 1]);
      RNF_MAP2(e[j], t[j], ex2_approx);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const f32x2 wx = pk(raw[8 * j + 2], raw[8 * j + 3]), wy = pk(raw[8 * j + 4], raw[8 * j + 5]), wz = pk(raw[8 * j + 6], raw[8 * j + 7]);
      a[j] = fma2(wz, bc(P.r[2]), fma2(wy, bc(P.r[1]), mul2(wx, bc(P.r[0]))));
      b[j] = fma2(wz, bc(P.v[2]), fma2(wy, bc(P.v[1]), mul2(wx, bc(P.v[0]))));
      n2[j] = fma2(b[j], b[j], mul2(a[j], a[j]));
      RNF_MAP2(rt[j], n2[j], sqrt_approx);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const f32x2 ope = add2(e[j], bc(1.0f));
      RNF_MAP2(big[j], ope, lg2_approx);
      small[j] = mul2(e[j], fma2(e[j], bc(-0.7213475204444817f), bc(1.4426950408889634f)));
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const f32x2 opr = add2(rt[j], bc(1.0f));
      f32x2 rc;
      RNF_MAP2(rc, opr, rcp_approx);
      ns[j] = mul2(rc, bc(-0.7f));                      // -0.7 / (1 + |w|)   (flow/mobiusflow.py:72, sign folded)
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      float tl, th, el, eh, bl, bh, sl, sh;
      upk(t[j], tl, th); upk(e[j], el, eh); upk(big[j], bl, bh); upk(small[j], sl, sh);
      const float vl = el < 0.0078125f ? sl : bl, vh = eh < 0.0078125f ? sh : bh;
      sp[j] = pk(tl > 28.853900817779268f ? tl : vl, th > 28.853900817779268f ? th : vh);
      nal[j] = mul2(ns[j], a[j]);
      nbe[j] = mul2(ns[j], b[j]);
      omw[j] = fma2(neg2(mul2(ns[j], ns[j])), n2[j], bc(1.0f));        // 1 - |w'|^2 = 1 - (0.7 / (1 + |w|))^2 |w|^2
    }
  }
  if (FWD) {
    f32x2 f[NP], hr[NP], hv[NP], mn[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      // forward direction: the evaluation point is the moving column itself, z = -|x| r exactly (v is orthogonal to x), so
      // zv = 0 and z - w' has the in-plane components (zr - alpha', -beta')
      const f32x2 dr = add2(nal[j], bc(zr)), dv = nbe[j];
      const f32x2 dd = fma2(dv, dv, mul2(dr, dr));
      f32x2 rc;
      RNF_MAP2(rc, dd, rcp_approx);
      f[j] = mul2(omw[j], rc);
      hr[j] = fma2(f[j], dr, nal[j]);
      hv[j] = fma2(f[j], dv, nbe[j]);
      float hrl, hrh, hvl, hvh;
      upk(hr[j], hrl, hrh);
      upk(hv[j], hvl, hvh);
      mn[j] = pk(fminf(fabsf(hvl), fabsf(hrl)), fminf(fabsf(hvh), fabsf(hrh)));
    }
    f32x2 at[NP];
#pragma unroll
    for (int j = 0; j < NP; ++j) at[j] = asin_unit2(mn[j]);
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      // angle of h from the negative r axis; hr < 0 always in the forward direction, so theta = pi - sign(hv) * that
      const f32x2 alt = fma2(at[j], bc(-1.0f), bc(1.5707963267948966f));
      float al_, ah_, bl_, bh_, hvl, hvh, hrl, hrh;
      upk(at[j], al_, ah_);
      upk(alt, bl_, bh_);
      upk(hv[j], hvl, hvh);
      upk(hr[j], hrl, hrh);
      const float ul = fabsf(hvl) > fabsf(hrl) ? bl_ : al_, uh = fabsf(hvh) > fabsf(hrh) ? bh_ : ah_;
      const f32x2 th = fma2(pk(copysignf(ul, hvl), copysignf(uh, hvh)), bc(-1.0f), bc(kPi));
      S_sp = add2(S_sp, sp[j]);
      S_th = fma2(sp[j], th, S_th);
      S_f = fma2(sp[j], f[j], S_f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      S_sp = add2(S_sp, sp[j]);
      upk(nal[j], raw[8 * j], raw[8 * j + 1]);
      upk(nbe[j], raw[8 * j + 2], raw[8 * j + 3]);
      upk(omw[j], raw[8 * j + 4], raw[8 * j + 5]);
      upk(sp[j], raw[8 * j + 6], raw[8 * j + 7]);
    }
  }
}

// Bisection probe of NP prepared pairs at the in-plane point (zr, zv) = (cos t, sin t): accumulates sum_k weight_k theta_k(z).
// Full-circle atan2: during the bisection z sweeps [pi/2, 3pi/2] and h may land anywhere (flow/mobiusflow.py:226-245).
template <int NP>
__device__ __forceinline__ void probe_pairs(float zr, float zv, const float* prm, f32x2& Fs) {
  f32x2 mn[NP];
  float hr_[2 * NP], hv_[2 * NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const f32x2 nal = pk(prm[8 * j], prm[8 * j + 1]), nbe = pk(prm[8 * j + 2], prm[8 * j + 3]), omw = pk(prm[8 * j + 4], prm[8 * j + 5]);
    const f32x2 dr = add2(nal, bc(zr)), dv = add2(nbe, bc(zv));
    const f32x2 dd = fma2(dv, dv, mul2(dr, dr));
    f32x2 rc;
    RNF_MAP2(rc, dd, rcp_approx);
    const f32x2 f = mul2(omw, rc);
    const f32x2 hr = fma2(f, dr, nal), hv = fma2(f, dv, nbe);
    upk(hr, hr_[2 * j], hr_[2 * j + 1]);
    upk(hv, hv_[2 * j], hv_[2 * j + 1]);
    mn[j] = pk(fminf(fabsf(hv_[2 * j]), fabsf(hr_[2 * j])), fminf(fabsf(hv_[2 * j + 1]), fabsf(hr_[2 * j + 1])));
  }
  f32x2 at[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) at[j] = asin_unit2(mn[j]);
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const f32x2 alt = fma2(at[j], bc(-1.0f), bc(1.5707963267948966f));
    // Wrap to [0, 2 pi) with sign bits instead of two compare / select pairs.  phi = angle of (|hr|, |hv|) from the r axis;
    // its complement c = pi/2 - phi is picked directly (asin or its complement), then
    //   pi - theta_[0,pi] = pi/2 + copysign(c, hr)            (hr >= 0: pi - phi, hr < 0: phi)        =: v in [0, pi]
    //   theta             = pi - copysign(v, hv)              (hv >= 0: pi - v,   hv < 0: pi + v)
    // which is atan2(hv, hr) wrapped exactly (tests/test_fastmath.py).
    float a_[2], b_[2], u_[2], v_[2];
    upk(at[j], a_[0], a_[1]);
    upk(alt, b_[0], b_[1]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float c = fabsf(hv_[2 * j + i]) > fabsf(hr_[2 * j + i]) ? a_[i] : b_[i];
      u_[i] = copysignf(c, hr_[2 * j + i]);
    }
    upk(add2(pk(u_[0], u_[1]), bc(1.5707963267948966f)), v_[0], v_[1]);
    const f32x2 th = fma2(pk(copysignf(v_[0], hv_[2 * j]), copysignf(v_[1], hv_[2 * j + 1])), bc(-1.0f), bc(kPi));
    Fs = fma2(pk(prm[8 * j + 6], prm[8 * j + 7]), th, Fs);
  }
}

// sum_k weight_k f_k(z) over NP prepared pairs (log-det of the inverse direction, flow/mobiusflow.py:169-181)
template <int NP>
__device__ __forceinline__ void jacobian_pairs(float zr, float zv, const float* prm, f32x2& Sf) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const f32x2 dr = add2(pk(prm[8 * j], prm[8 * j + 1]), bc(zr)), dv = add2(pk(prm[8 * j + 2], prm[8 * j + 3]), bc(zv));
    const f32x2 dd = fma2(dv, dv, mul2(dr, dr));
    f32x2 rc;
    RNF_MAP2(rc, dd, rcp_approx);
    Sf = fma2(pk(prm[8 * j + 6], prm[8 * j + 7]), mul2(pk(prm[8 * j + 4], prm[8 * j + 5]), rc), Sf);
  }
}

}  // namespace rnf
