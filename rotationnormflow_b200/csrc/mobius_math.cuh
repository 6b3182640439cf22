// mobius_math.cuh -- per-(rotation, mixture component) arithmetic of the Mobius coupling layer.
//
// Follows flow/mobiusflow.py (reference root relative):
//   :17-24   _h          h_w(z) = (1-|w|^2)/|z-w|^2 (z-w) - w
//   :62-72   parameter preparation (project w onto the plane orthogonal to y, softplus, 0.7 w/(1+|w|))
//   :94-99   angle of h in the (r, v) frame, wrapped to [0, 2 pi)
//   :104-125 the explicit Jacobian, whose norm is exactly f = (1-|w|^2)/|z-w|^2 (Householder factor and
//            |dz/dtheta| = 1 are norm preserving; SURVEY.md A.3 step 9, verified to 1e-15 in fp64)
//   :196-224 BinFind.forward bisection arithmetic
#pragma once
#include <cuda_runtime.h>

#include "so3_math.cuh"

namespace rnf {

constexpr float kTwoPi = 6.283185307179586f;  // torch.pi * 2 rounded to fp32
constexpr float kPi = 3.141592653589793f;

// torch.nn.functional.softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus_torch(float a) { return a > 20.0f ? a : log1pf(expf(a)); }

// r = -x/|x| ; v = (y x r)/|y x r|                                   (flow/mobiusflow.py:64-67)
__device__ __forceinline__ void make_frame(const float x[3], const float y[3], float r[3], float v[3]) {
  r[0] = -x[0]; r[1] = -x[1]; r[2] = -x[2];
  normalize3(r);
  cross3(y, r, v);
  normalize3(v);
}

// raw conditioner outputs (w) -> prepared component centre: w <- (I - y y^T) w ; w <- 0.7 w / (1 + |w|)
__device__ __forceinline__ void comp_prep(float w[3], const float y[3]) {
  const float yw = dot3(y, w);
  w[0] = fmaf(-yw, y[0], w[0]);
  w[1] = fmaf(-yw, y[1], w[1]);
  w[2] = fmaf(-yw, y[2], w[2]);
  const float s = 0.7f / (1.0f + sqrtf(dot3(w, w)));
  w[0] *= s; w[1] *= s; w[2] *= s;
}

// Evaluate one Mobius map at point z: wrapped angle theta of h_w(z) in the (r, v) frame and f = |dh/dtheta|.
__device__ __forceinline__ void comp_eval(const float z[3], const float w[3], const float r[3], const float v[3],
                                          float& theta, float& f) {
  const float d0 = z[0] - w[0], d1 = z[1] - w[1], d2 = z[2] - w[2];
  const float dd = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
  const float ww = dot3(w, w);
  f = (1.0f - ww) / dd;
  const float h0 = fmaf(f, d0, -w[0]), h1 = fmaf(f, d1, -w[1]), h2 = fmaf(f, d2, -w[2]);
  const float hv = fmaf(h2, v[2], fmaf(h1, v[1], h0 * v[0]));
  const float hr = fmaf(h2, r[2], fmaf(h1, r[1], h0 * r[0]));
  float th = atan2f(hv, hr);
  theta = th >= 0.0f ? th : th + kTwoPi;
}

// point on the circle: z = r cos(t) + v sin(t)                        (flow/mobiusflow.py:102,169,231)
__device__ __forceinline__ void circle_point(const float r[3], const float v[3], float t, float z[3]) {
  float s, c;
  sincosf(t, &s, &c);
  z[0] = fmaf(v[0], s, r[0] * c);
  z[1] = fmaf(v[1], s, r[1] * c);
  z[2] = fmaf(v[2], s, r[2] * c);
}

}  // namespace rnf
