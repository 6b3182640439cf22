// healpix.cu -- on-device HEALPix SO(3) grid (Yershova et al.), index -> rotation.
//
// Follows utils/sd.py:48-82 (generate_healpix_grid): nside = 2^level, pixel centres from healpy.pix2vec in RING
// order (public HEALPix definition), azimuth = atan2(y, x), polar = acos(z),
// tilts = linspace(0, 2 pi, 6*nside, endpoint=False), R[t*npix + b] = Rx(az_b) Rz(polar_b) Rx(tilt_t),
// all in float64 and rounded to float32 at the end (torch.Tensor(Rs), utils/sd.py:82).
#include "rnf_common.cuh"

namespace rnf {
namespace {

__device__ __forceinline__ long long isqrt_ll(long long v) {
  long long r = (long long)sqrt((double)v);
  while (r * r > v) --r;
  while ((r + 1) * (r + 1) <= v) ++r;
  return r;
}

__device__ void pix2zphi_ring(long long nside, long long p, double& z, double& phi) {
  const long long npix = 12 * nside * nside;
  const long long ncap = 2 * nside * (nside - 1);
  const double halfpi = 1.5707963267948966;
  if (p < ncap) {  // north polar cap
    const long long i = (1 + isqrt_ll(1 + 2 * p)) >> 1;
    const long long j = p + 1 - 2 * i * (i - 1);
    z = 1.0 - (double)(i * i) * (4.0 / (double)npix);
    phi = ((double)j - 0.5) * halfpi / (double)i;
  } else if (p < npix - ncap) {  // equatorial belt
    const long long ip = p - ncap;
    const long long i = ip / (4 * nside) + nside;
    const long long j = ip % (4 * nside) + 1;
    const double fodd = ((i + nside) & 1) ? 1.0 : 0.5;
    z = (double)(2 * nside - i) * (2.0 / (double)(3 * nside));
    phi = ((double)j - fodd) * halfpi / (double)nside;
  } else {  // south polar cap
    const long long ip = npix - p;
    const long long i = (1 + isqrt_ll(2 * ip - 1)) >> 1;
    const long long j = 4 * i + 1 - (ip - 2 * i * (i - 1));
    z = -1.0 + (double)(i * i) * (4.0 / (double)npix);
    phi = ((double)j - 0.5) * halfpi / (double)i;
  }
}

__global__ void healpix_grid_kernel(int level, int64_t begin, int64_t end, float* __restrict__ out) {
  const int64_t n = end - begin;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    const long long nside = 1LL << level;
    const long long npix = 12 * nside * nside;
    const long long ntilt = 6 * nside;
    const long long idx = begin + k;
    const long long t = idx / npix, b = idx % npix;
    double z, phi;
    pix2zphi_ring(nside, b, z, phi);
    const double sth = sqrt((1.0 - z) * (1.0 + z));
    const double az = atan2(sth * sin(phi), sth * cos(phi));
    const double polar = acos(z);
    const double tilt = (double)t * (6.283185307179586 / (double)ntilt);
    double sa, ca, sp, cp, st, ct;
    sincos(az, &sa, &ca);
    sincos(polar, &sp, &cp);
    sincos(tilt, &st, &ct);
    // Rx(az) Rz(polar), then times Rx(tilt)
    const double m00 = cp, m01 = -sp, m10 = ca * sp, m11 = ca * cp, m12 = -sa, m20 = sa * sp, m21 = sa * cp, m22 = ca;
    float* o = out + k * 9;
    o[0] = (float)m00;
    o[1] = (float)(m01 * ct);
    o[2] = (float)(-m01 * st);
    o[3] = (float)m10;
    o[4] = (float)(m11 * ct + m12 * st);
    o[5] = (float)(-m11 * st + m12 * ct);
    o[6] = (float)m20;
    o[7] = (float)(m21 * ct + m22 * st);
    o[8] = (float)(-m21 * st + m22 * ct);
  }
}

// MatrixFisherN._log_prob (utils/fisher.py:217-232): out[i] = sum(A_b * R_i) - c_b, b = image of row i (image-major rows,
// `rows_per_image` rotations each, as inputs.reshape(B, -1, 3, 3) does).  HBM bound: 36 B in + 4 B out per rotation.
__global__ void fisher_logprob_kernel(const float* __restrict__ A9, const float* __restrict__ c, const float* __restrict__ R,
                                      int64_t N, int64_t rows_per_image, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / rows_per_image;
    float tr = 0.0f;
#pragma unroll
    for (int k = 0; k < 9; ++k) tr = fmaf(__ldg(A9 + b * 9 + k), __ldg(R + i * 9 + k), tr);
    out[i] = tr - __ldg(c + b);
  }
}

// min_geodesic_distance_rotmats (utils/utils.py:231-235): angle between est[b] and the closest of its K ground-truth rotations:
// acos(clip((max_k sum_ij est_ij gt_kij - 1)/2, -1, 1)).  K = 1 is geodesic_distance_rotmats (utils/utils.py:225-228).
__global__ void min_geodesic_kernel(const float* __restrict__ est, const float* __restrict__ gt, int64_t B, int64_t K,
                                    float* __restrict__ out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float e[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) e[i] = __ldg(est + b * 9 + i);
  float best = -INFINITY;
  for (int64_t k = 0; k < K; ++k) {
    const float* g = gt + (b * K + k) * 9;
    float tr = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) tr = fmaf(e[i], __ldg(g + i), tr);
    best = fmaxf(best, tr);
  }
  out[b] = acosf(fminf(fmaxf((best - 1.0f) * 0.5f, -1.0f), 1.0f));
}

}  // namespace

cudaError_t launch_min_geodesic(const float* est, const float* gt, int64_t B, int64_t K, float* out, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  min_geodesic_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(est, gt, B, K, out);
  return cudaGetLastError();
}

cudaError_t launch_fisher_logprob(const float* A9, const float* c, const float* R, int64_t N, int64_t rows_per_image, float* out,
                                  cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t blocks = (N + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  fisher_logprob_kernel<<<(unsigned)blocks, 256, 0, st>>>(A9, c, R, N, rows_per_image, out);
  return cudaGetLastError();
}

cudaError_t launch_healpix(int level, int64_t begin, int64_t end, float* out, cudaStream_t st) {
  const int64_t n = end - begin;
  if (n <= 0) return cudaSuccess;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  healpix_grid_kernel<<<(unsigned)blocks, 256, 0, st>>>(level, begin, end, out);
  return cudaGetLastError();
}

}  // namespace rnf
