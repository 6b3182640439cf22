// flow_tc2.cu -- software-pipelined variant of the tcgen05 flow kernel (forward / grid mode), sm_100a.
//
// Same math, same packed weights, same TMEM / shared-memory operand formats as flow_tc.cu, different schedule.
// flow_tc.cu binds every thread to ONE tile, so a tile's 8 warps idle during each of the four dependent GEMM round
// trips of a Mobius layer (barrier -> issue -> MMA -> wake-up), and the other tile's warps alone cannot fill the issue
// slots (measured: 64 % issue utilisation, 25 % of warp time inside hidden-layer MMA waits, profiles/).
// Here all 512 threads work on BOTH tiles, stage by stage:
//
//     stage(A) -> arrive ready[A] (GEMM(A) is issued, runs on the tensor core) -> stage(B) -> arrive ready[B] -> stage'(A) ...
//
// so the GEMM of one tile always executes underneath the CUDA-core stage of the other tile and nobody waits on a
// named barrier inside the MLP chain: "operand ready" is an mbarrier every thread arrives on (non-blocking) and only the
// issuing thread waits for.  A thread owns row r (TMEM lane) of tile A *and* of tile B and one quarter of the columns:
// 16 hidden units and 16 mixture components; the four column groups of a row meet once per layer, through shared
// memory, to combine three partial sums per tile.
#include "tc_common.cuh"

namespace rnf {
namespace {

constexpr int kThreads = 512;
constexpr int kRows = 128;                        // rows per tile = TMEM lanes

// shared-memory image (bytes from a 1024-aligned base)
constexpr int kOffW = 0;
constexpr int kOffLastW = kHidW;                  // 49152
constexpr int kOffAux = kOffLastW + kLastW;       // 114688, double buffered (layer parity)
constexpr int kOffA = kOffAux + 2 * kAuxStride;   // 120832 : [tile][hi|lo] 128x64 fp16 (16 KB each)
constexpr int kOffXchg = kOffA + 4 * 16384;       // [tile][column group][3][128] fp32
constexpr int kOffRed = kOffXchg + 2 * 4 * 3 * 128 * 4;
constexpr int kOffBar = kOffRed + 128;
constexpr int kOffMisc = kOffBar + 8 * 16;        // tmem base, Mobius offset table
constexpr int kSmemBytes = kOffMisc + 16 + 64 * 8;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kOffA % 1024 == 0 && kOffLastW % 1024 == 0, "UMMA SW128 tiles need 1024 B alignment");
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KB shared-memory limit of an sm_100 CTA");

enum { BAR_W_FULL = 0 /* W1..W4 */, BAR_AUX_FULL = 4 /* [2] */, BAR_READY = 6 /* [tile] */, BAR_MMA = 8 /* [tile][2] */, BAR_COUNT = 12 };

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float v[16]) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]), "r"(u[8]), "r"(u[9]),
      "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 16 non-negative activations (hidden units 16cg .. 16cg+15 of row r) -> fp16 hi / lo planes of the K-major SW128 A operand
__device__ __forceinline__ void store_a16(uint8_t* a_hi, uint8_t* a_lo, int r, int cg, const float v[16]) {
  const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = v[8 * c + 2 * e], x1 = v[8 * c + 2 * e + 1];
      const __half2 hh = __floats2half2_rn(x0, x1);
      const float2 back = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(x0 - back.x, x1 - back.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const int off = rbase + (((2 * cg + c) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

template <bool GRID>
__global__ void __launch_bounds__(kThreads, 1) flow_tc2_kernel(const FlowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cg = warp >> 2;                          // column group: hidden units / mixture components 16cg .. 16cg+15
  const int rowi = (warp & 3) * 32 + lane;           // row inside a tile = TMEM lane
  const uint32_t bars = smem_u32(smem + kOffBar);
  const bool issuer0 = tid == 0, issuer1 = tid == 256;   // issue the GEMMs of tile 0 / tile 1

  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  long long* s_moff = reinterpret_cast<long long*>(smem + kOffMisc + 16);

  int n_mob = 0;
  for (int i = 0; i < a.n_layers; ++i)
    if (a.layers[i].kind == RNF_LAYER_MOBIUS) {
      if (tid == 0) s_moff[n_mob] = a.layers[i].w_off_tc;
      ++n_mob;
    }
  if (tid == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) mbar_init(bars + 8 * i, (i == BAR_READY || i == BAR_READY + 1) ? kThreads : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tm_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);     // + 256 * tile + column

  const int64_t n_pairs = (a.n_tiles + 1) / 2;
  const int64_t my_items = blockIdx.x < n_pairs ? (n_pairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int64_t total_steps = my_items * n_mob;
  const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(a.weights);
  auto load_piece = [&](int64_t mob_step, int piece, int abuf) {
    const uint8_t* src = wbytes + s_moff[mob_step % n_mob] * 4;
    uint32_t dst, bytes, bar;
    if (piece < 3) { src += piece * kW1Bytes; dst = kOffW + piece * kW1Bytes; bytes = kW1Bytes; bar = BAR_W_FULL + piece; }
    else if (piece == 3) { src += kHidW; dst = kOffLastW; bytes = kLastW; bar = BAR_W_FULL + 3; }
    else { src += kHidW + kLastW; dst = kOffAux + abuf * kAuxStride; bytes = kAuxBytes; bar = BAR_AUX_FULL + abuf; }
    mbar_expect_tx(bars + 8 * bar, bytes);
    bulk_g2s(smem_u32(smem + dst), src, bytes, bars + 8 * bar);
  };
  if (tid == 0 && total_steps > 0) {
    for (int piece = 0; piece < 4; ++piece) load_piece(0, piece, 0);
    load_piece(0, 4, 0);
    if (total_steps > 1) load_piece(1, 4, 1);
  }

  const uint32_t w_hid_d = umma_desc_lo(smem_u32(smem + kOffW)), w_last_d = umma_desc_lo(smem_u32(smem + kOffLastW));
  float* xchg = reinterpret_cast<float*>(smem + kOffXchg);           // [tile][cg][3][128]
  uint32_t par_mma = 0;      // bit t: mma[t][0], bit 2+t: mma[t][1]
  uint32_t par_ready = 0;    // issuers: phase parity of ready[my tile]
  uint32_t par_w = 0;        // issuers: bit l = weight piece l
  int64_t step = 0;

  // GEMM g of tile t: g = 0..2 hidden (piece g, N = 64), g = 3 fc_last (piece 3, two N = 128 chunks)
  auto issue_gemm = [&](int t, int g) {
    mbar_wait(bars + 8 * (BAR_READY + t), par_ready & 1u);            // every thread stored its part of the A operand
    par_ready ^= 1u;
    mbar_wait(bars + 8 * (BAR_W_FULL + g), (par_w >> g) & 1u);
    tc_fence_after();
    const uint32_t a_hi_d = umma_desc_lo(smem_u32(smem + kOffA + t * 32768)), a_lo_d = a_hi_d + (16384 >> 4);
    const uint32_t d = tmem_base + t * 256;
    const uint32_t bar0 = bars + 8 * (BAR_MMA + 2 * t);
    if (g < 3) {
      const uint32_t wb = w_hid_d + g * (kW1Bytes >> 4);
      issue_split_gemm(d, a_hi_d, a_lo_d, wb, wb + (8192 >> 4), umma_idesc(128, 64));
      umma_commit(bar0);
    } else {
      issue_split_gemm(d, a_hi_d, a_lo_d, w_last_d, w_last_d + (32768 >> 4), umma_idesc(128, 128));
      umma_commit(bar0);
      issue_split_gemm(d + 128, a_hi_d, a_lo_d, w_last_d + (16384 >> 4), w_last_d + ((32768 + 16384) >> 4), umma_idesc(128, 128));
      umma_commit(bar0 + 8);
    }
  };

  for (int64_t item = 0; item < my_items; ++item) {
    float R[2][9];
    float ldj[2] = {0.0f, 0.0f};
    bool valid[2] = {false, false};
    int64_t row[2] = {0, 0}, img[2] = {0, 0}, gidx[2] = {0, 0}, tile_idx[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      tile_idx[t] = 2 * (blockIdx.x + item * (int64_t)gridDim.x) + t;
#pragma unroll
      for (int i = 0; i < 9; ++i) R[t][i] = (i % 4 == 0) ? 1.0f : 0.0f;
      if (tile_idx[t] < a.n_tiles) {
        if (GRID) {
          img[t] = tile_idx[t] / a.tiles_per_image;
          gidx[t] = (tile_idx[t] % a.tiles_per_image) * kRows + rowi;
          valid[t] = gidx[t] < a.G;
          row[t] = img[t] * a.G + gidx[t];
          if (valid[t]) {
            float Gm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Gm[i] = __ldg(a.R_in + gidx[t] * 9 + i);
            if (a.offset != nullptr) {               // samples = grid @ random_rot (eval.py:439-440)
              float O[9];
#pragma unroll
              for (int i = 0; i < 9; ++i) O[i] = __ldg(a.offset + i);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  R[t][3 * i + j] = fmaf(Gm[3 * i + 2], O[6 + j], fmaf(Gm[3 * i + 1], O[3 + j], Gm[3 * i] * O[j]));
            } else {
#pragma unroll
              for (int i = 0; i < 9; ++i) R[t][i] = Gm[i];
            }
          }
        } else {
          row[t] = tile_idx[t] * kRows + rowi;
          valid[t] = row[t] < a.N;
          if (valid[t]) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[t][i] = __ldg(a.R_in + row[t] * 9 + i);
            if (a.cond != nullptr)
              img[t] = a.feat_index != nullptr ? (int64_t)__ldg(a.feat_index + row[t]) : row[t] / a.rows_per_image;
          }
        }
      }
    }

#pragma unroll 1
    for (int li = 0; li < a.n_layers; ++li) {
      const LayerDev L = a.layers[li];
      if (L.kind != RNF_LAYER_MOBIUS) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const float* W = L.cond_slot >= 0
                               ? a.cond + img[t] * a.cond_stride + (int64_t)a.n_mobius_slots * kH + (int64_t)L.cond_slot * kAffFloats
                               : a.weights + L.w_off;
          float Wr[17];
#pragma unroll
          for (int i = 0; i < 17; ++i) Wr[i] = __ldg(W + i);
          const float loglen = quat_affine_fast(Wr, R[t]);
          if (L.has_ldj) ldj[t] += Wr[16] - 4.0f * loglen;
        }
        continue;
      }
      // ================================ Mobius layer, both tiles ================================
      const int p0 = L.perm, p1 = (L.perm + 1) % 3, p2 = (L.perm + 2) % 3;
      const int abuf = (int)(step & 1);
      mbar_wait(bars + 8 * (BAR_AUX_FULL + abuf), (uint32_t)((step >> 1) & 1));
      const uint8_t* aux = smem + kOffAux + abuf * kAuxStride;
      const float4* sFirst = reinterpret_cast<const float4*>(aux);
      const float* sBiasHid = reinterpret_cast<const float*>(aux + 1024);
      const float* sBiasLast = reinterpret_cast<const float*>(aux + 1792);

      Plane P[2];
      float zr[2], zv[2];
      // ---- stage 0: frame + first conditioner layer (my 16 hidden units) ----
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        float x[3], y[3];
        get_col(R[t], p0, x);
        get_col(R[t], p1, y);
        make_frame_fast(x, y, P[t]);
        zr[t] = dot3(x, P[t].r);
        zv[t] = dot3(x, P[t].v);
        const float* cimg = (L.cond_slot >= 0 && a.cond != nullptr)
                                ? a.cond + img[t] * a.cond_stride + (int64_t)L.cond_slot * kH + 16 * cg : nullptr;
        float h0[16], act[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cimg != nullptr) cf = __ldg(reinterpret_cast<const float4*>(cimg) + j4);
          const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 f = sFirst[16 * cg + 4 * j4 + e];
            const float hv = fmaf(f.z, y[2], fmaf(f.y, y[1], fmaf(f.x, y[0], f.w))) + cc[e];
            h0[4 * j4 + e] = hv;
            act[4 * j4 + e] = fmaxf(hv, 0.0f);
          }
        }
        store_a16(smem + kOffA + t * 32768, smem + kOffA + t * 32768 + 16384, rowi, cg, act);
        tmem_st16(tm_lane + 256 * t + 64 + 16 * cg, h0);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bars + 8 * (BAR_READY + t));
        if (t == 0 ? issuer0 : issuer1) issue_gemm(t, 0);
      }
      // ---- stages 1..3: hidden-layer epilogues; the GEMM of one tile runs underneath the epilogue of the other ----
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_wait(bars + 8 * (BAR_MMA + 2 * t), (par_mma >> t) & 1u);
          par_mma ^= (1u << t);
          tc_fence_after();
          // piece l is dead once GEMM l of BOTH tiles has completed: this thread has now observed both
          if (t == 1 && issuer1 && step + 1 < total_steps) load_piece(step + 1, l, 0);
          float acc[16];
          tmem_ld16(tm_lane + 256 * t + 16 * cg, acc);
          const float4* bias4 = reinterpret_cast<const float4*>(sBiasHid + 64 * l + 16 * cg);
          if (l < 2) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b = bias4[j4];
              acc[4 * j4 + 0] = fmaxf(fmaf(acc[4 * j4 + 0], kWUnscale, b.x), 0.0f);
              acc[4 * j4 + 1] = fmaxf(fmaf(acc[4 * j4 + 1], kWUnscale, b.y), 0.0f);
              acc[4 * j4 + 2] = fmaxf(fmaf(acc[4 * j4 + 2], kWUnscale, b.z), 0.0f);
              acc[4 * j4 + 3] = fmaxf(fmaf(acc[4 * j4 + 3], kWUnscale, b.w), 0.0f);
            }
          } else {                                   // relu_last(x0 + x)   (flow/condition.py:29)
            float h0[16];
            tmem_ld16(tm_lane + 256 * t + 64 + 16 * cg, h0);
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b = bias4[j4];
              acc[4 * j4 + 0] = fmaxf(fmaf(acc[4 * j4 + 0], kWUnscale, b.x) + h0[4 * j4 + 0], 0.0f);
              acc[4 * j4 + 1] = fmaxf(fmaf(acc[4 * j4 + 1], kWUnscale, b.y) + h0[4 * j4 + 1], 0.0f);
              acc[4 * j4 + 2] = fmaxf(fmaf(acc[4 * j4 + 2], kWUnscale, b.z) + h0[4 * j4 + 2], 0.0f);
              acc[4 * j4 + 3] = fmaxf(fmaf(acc[4 * j4 + 3], kWUnscale, b.w) + h0[4 * j4 + 3], 0.0f);
            }
          }
          store_a16(smem + kOffA + t * 32768, smem + kOffA + t * 32768 + 16384, rowi, cg, acc);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(bars + 8 * (BAR_READY + t));
          if (t == 0 ? issuer0 : issuer1) issue_gemm(t, l + 1);
        }
      }
      if (issuer0 || issuer1) par_w ^= 0xFu;
      // ---- stage 4: mixture of my 16 components per tile, straight from TMEM ----
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int chunk = cg >> 1;                   // fc_last chunk holding my 64 columns
        mbar_wait(bars + 8 * (BAR_MMA + 2 * t + chunk), (par_mma >> (2 * chunk + t)) & 1u);
        par_mma ^= (1u << t) | (4u << t);            // both chunk barriers complete exactly once per layer
        tc_fence_after();
        // fc_last weights are dead once both chunks of BOTH tiles have completed: issuer1 (cg = 2) waits on chunk 1, which is
        // committed after chunk 0 by the same thread; tile 0's chunk 1 was observed one iteration earlier
        if (t == 1 && issuer1 && step + 1 < total_steps) load_piece(step + 1, 3, 0);
        float S_sp = 0.0f, S_th = 0.0f, S_f = 0.0f;
#pragma unroll 1
        for (int q = 0; q < 2; ++q) {
          float acc[32];
          tmem_ld32(tm_lane + 256 * t + 64 * cg + 32 * q, acc);
          const float4* b4 = reinterpret_cast<const float4*>(sBiasLast + 64 * cg + 32 * q);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 b = b4[k];
            const float sp = softplus_fast(fmaf(acc[4 * k], kWUnscale, b.x));
            float al, be, omw, th, f;
            comp_prep2(P[t], fmaf(acc[4 * k + 1], kWUnscale, b.y), fmaf(acc[4 * k + 2], kWUnscale, b.z),
                       fmaf(acc[4 * k + 3], kWUnscale, b.w), al, be, omw);
            comp_eval2(zr[t], zv[t], al, be, omw, th, f);
            S_sp += sp;
            S_th = fmaf(sp, th, S_th);
            S_f = fmaf(sp, f, S_f);
          }
        }
        float* mine = xchg + ((t * 4 + cg) * 3) * 128 + rowi;
        mine[0] = S_sp; mine[128] = S_th; mine[256] = S_f;
      }
      tc_fence_before();
      __syncthreads();
      // every thread is done with this layer's side data: refill its buffer with the layer after next
      if (tid == 0 && step + 2 < total_steps) load_piece(step + 2, 4, abuf);
      // ---- combine the four column groups of each row (fixed order) and move the rotation ----
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float* px = xchg + (t * 4 * 3) * 128 + rowi;
        const float S_sp = (px[0] + px[384]) + (px[768] + px[1152]);
        const float S_th = (px[128] + px[512]) + (px[896] + px[1280]);
        const float S_f = (px[256] + px[640]) + (px[1024] + px[1408]);
        const float inv_sp = rcp_nr(S_sp);
        float nx[3], nz[3], y[3];
        get_col(R[t], p1, y);
        circle_point(P[t].r, P[t].v, S_th * inv_sp, nx);
        ldj[t] += logf(S_f * inv_sp);
        cross3(nx, y, nz);
        normalize3_fast(nz);
        set_col(R[t], p0, nx);
        set_col(R[t], p2, nz);
      }
      ++step;
    }

    // ================================ outputs ================================
    if (!GRID) {
      if (cg == 0) {
#pragma unroll
        for (int t = 0; t < 2; ++t)
          if (valid[t]) {
#pragma unroll
            for (int i = 0; i < 9; ++i) a.R_out[row[t] * 9 + i] = R[t][i];
            a.ldj_out[row[t]] = ldj[t];
          }
      }
    } else if (cg == 0) {                            // warps 0..3 reduce both tiles (named barrier 3, 128 threads)
      float* s_v = reinterpret_cast<float*>(smem + kOffRed);
      long long* s_i = reinterpret_cast<long long*>(smem + kOffRed + 32);
#pragma unroll
      for (int t = 0; t < 2; ++t) {                  // unrolled: keeps R[t] / ldj[t] in registers (static indices)
        if (tile_idx[t] >= a.n_tiles) continue;      // uniform over the CTA
        float lp = ldj[t];
        if (a.fisher_A != nullptr) {
          float tr = 0.0f;
#pragma unroll
          for (int i = 0; i < 9; ++i) tr = fmaf(__ldg(a.fisher_A + img[t] * 9 + i), R[t][i], tr);
          lp += tr - __ldg(a.fisher_c + img[t]);
        }
        if (!valid[t]) lp = -INFINITY;
        if (a.logp_out != nullptr && valid[t]) a.logp_out[row[t]] = lp;
        float bv = lp;
        long long bi = valid[t] ? (long long)gidx[t] : 0x7fffffffffffffffLL;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
        named_bar(3, 128);
        bv = s_v[0]; bi = s_i[0];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float ov = s_v[w];
          const long long oi = s_i[w];
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        const float m = bv;
        float e = (valid[t] && m > -INFINITY) ? expf(lp - m) : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        named_bar(3, 128);
        if (lane == 0) s_v[warp] = e;
        named_bar(3, 128);
        if (rowi == 0) {
          const float s = (s_v[0] + s_v[1]) + (s_v[2] + s_v[3]);
          float* p = a.part + tile_idx[t] * 4;
          p[0] = m;
          p[1] = s;
          p[2] = __int_as_float((int)(bi & 0xffffffffLL));
          p[3] = __int_as_float((int)(bi >> 32));
        }
        named_bar(3, 128);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace

cudaError_t launch_flow_tc2(const FlowArgs& a, int sm_count, cudaStream_t st) {
  const bool grid_mode = a.G > 0;
  void (*kern)(const FlowArgs) = grid_mode ? flow_tc2_kernel<true> : flow_tc2_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
  if (e != cudaSuccess) return e;
  if (a.n_tiles <= 0) return cudaSuccess;
  const int64_t pairs = (a.n_tiles + 1) / 2;
  const int64_t grid = pairs < sm_count ? pairs : sm_count;
  kern<<<(unsigned)grid, kThreads, kSmemAlloc, st>>>(a);
  return cudaGetLastError();
}

}  // namespace rnf
