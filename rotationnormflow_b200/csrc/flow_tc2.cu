// flow_tc2.cu -- software-pipelined variant of the tcgen05 flow kernel (forward / grid mode), sm_100a.
//
// Same math, same packed weights, same TMEM / shared-memory operand formats as flow_tc.cu, different schedule.
// flow_tc.cu binds every thread to ONE tile, so a tile's 8 warps idle during each of the four dependent GEMM round
// trips of a Mobius layer (barrier -> issue -> MMA -> wake-up), and the other tile's warps alone cannot fill the issue
// slots (measured: 64 % issue utilisation, 25 % of warp time inside hidden-layer MMA waits, profiles/).
// Here all 512 threads work on BOTH tiles, and the tiles run half a Mobius layer apart.  In every "half-step" one tile
// (X) walks its MLP chain while the other (Y) evaluates its mixture, interleaved stage by stage:
//
//     chain_s(X) -> arrive ready[X] (GEMM_s(X) is issued and runs on the tensor core) -> mixture quarter s of Y -> chain_s+1(X) ...
//
// so each of the four dependent GEMM round trips of X executes underneath ~250 instructions per thread of Y's mixture
// arithmetic and nobody waits on a named barrier inside the chain: "operand ready" is an mbarrier every thread arrives on
// (non-blocking); only the issuing thread waits for it.  A thread owns row r (TMEM lane) of tile 0 *and* of tile 1 and
// one quarter of the columns: 16 hidden units and 16 mixture components; the four column groups of a row meet once per
// layer and tile, through shared memory, to combine three partial sums.
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
constexpr int kOffXchg = kOffA + 4 * 16384;       // [column group][3][128] fp32 (one tile at a time)
constexpr int kOffRed = kOffXchg + 4 * 3 * 128 * 4;
constexpr int kOffBar = kOffRed + 128;
constexpr int kOffMisc = kOffBar + 8 * 16;        // tmem base, Mobius offset table
constexpr int kSmemBytes = kOffMisc + 16 + 64 * 8;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kOffA % 1024 == 0 && kOffLastW % 1024 == 0, "UMMA SW128 tiles need 1024 B alignment");
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KB shared-memory limit of an sm_100 CTA");

enum { BAR_W_FULL = 0 /* W1..W4 */, BAR_AUX_FULL = 4 /* [2] */, BAR_READY = 6 /* [tile] */, BAR_MMA = 8 /* [tile] */, BAR_COUNT = 10 };

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
// ReLU + split of 16 pre-activations (hidden units 16cg .. 16cg+15 of row r) -> fp16 hi / lo planes of the K-major SW128 A operand
__device__ __forceinline__ void store_a16(uint8_t* a_hi, uint8_t* a_lo, int r, int cg, const float v[16]) {
  const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      relu_split2(v[8 * c + 2 * e], v[8 * c + 2 * e + 1], hi[e], lo[e]);     // ReLU is applied here
    }
    const int off = rbase + (((2 * cg + c) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Everything a thread keeps about one tile (its row of that tile), in registers.
struct Tile {
  float R[9];
  float ldj;
  Plane P;            // frame of the current Mobius layer (chain -> mixture -> finish)
  float zr, zv;       // in-plane coordinates of the moving column
  float S_sp, S_th, S_f;
  int li;             // layer cursor (uniform)
  int perm;           // p0 of the current Mobius layer
  int64_t item;       // work item (tile pair) this tile is on; < 0: finished
  int64_t img;
  bool valid;
};

template <bool GRID>
__global__ void __launch_bounds__(kThreads, 1) flow_tc2_kernel(const FlowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int cg = warp >> 2;                          // column group: hidden units / mixture components 16cg .. 16cg+15
  const int rowi = (warp & 3) * 32 + lane;           // row inside a tile = TMEM lane
  const uint32_t bars = smem_u32(smem + kOffBar);

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
  const int64_t total_steps = my_items * n_mob;      // Mobius steps per tile
  const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(a.weights);
  auto load_piece = [&](int mob_idx, int piece, int abuf) {   // mob_idx: index into the Mobius layer table
    const uint8_t* src = wbytes + s_moff[mob_idx] * 4;
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
    if (total_steps > 1) load_piece(n_mob > 1 ? 1 : 0, 4, 1);
  }

  const uint32_t w_hid_d = umma_desc_lo(smem_u32(smem + kOffW)), w_last_d = umma_desc_lo(smem_u32(smem + kOffLastW));
  float* xchg = reinterpret_cast<float*>(smem + kOffXchg);           // [cg][3][128]
  uint32_t par_mma = 0;      // bit t: phase parity of mma[t] (4 completions per Mobius step)
  uint32_t par_ready = 0;    // issuer threads: phase parity of ready[my tile]

  // ---- row I/O --------------------------------------------------------------------------------------------------
  auto load_rows = [&](Tile& T, int t) {
    const int64_t tile_idx = 2 * (blockIdx.x + T.item * (int64_t)gridDim.x) + t;
#pragma unroll
    for (int i = 0; i < 9; ++i) T.R[i] = (i % 4 == 0) ? 1.0f : 0.0f;
    T.ldj = 0.0f;
    T.valid = false;
    T.img = 0;
    if (tile_idx >= a.n_tiles) return;
    if (GRID) {
      T.img = tile_idx / a.tiles_per_image;
      const int64_t g = (tile_idx % a.tiles_per_image) * kRows + rowi;
      T.valid = g < a.G;
      if (T.valid) {
        float Gm[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) Gm[i] = __ldg(a.R_in + g * 9 + i);
        if (a.offset != nullptr) {                   // samples = grid @ random_rot (eval.py:439-440)
          float O[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) O[i] = __ldg(a.offset + i);
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
              T.R[3 * i + j] = fmaf(Gm[3 * i + 2], O[6 + j], fmaf(Gm[3 * i + 1], O[3 + j], Gm[3 * i] * O[j]));
        } else {
#pragma unroll
          for (int i = 0; i < 9; ++i) T.R[i] = Gm[i];
        }
      }
    } else {
      const int64_t row = tile_idx * kRows + rowi;
      T.valid = row < a.N;
      if (T.valid) {
#pragma unroll
        for (int i = 0; i < 9; ++i) T.R[i] = __ldg(a.R_in + row * 9 + i);
        if (a.cond != nullptr) T.img = a.feat_index != nullptr ? (int64_t)__ldg(a.feat_index + row) : row / a.rows_per_image;
      }
    }
  };

  auto store_rows = [&](Tile& T, int t) {
    const int64_t tile_idx = 2 * (blockIdx.x + T.item * (int64_t)gridDim.x) + t;
    if (tile_idx >= a.n_tiles) return;               // uniform over the CTA
    if (!GRID) {
      const int64_t row = tile_idx * kRows + rowi;
      if (cg == 0 && T.valid) {
#pragma unroll
        for (int i = 0; i < 9; ++i) a.R_out[row * 9 + i] = T.R[i];
        a.ldj_out[row] = T.ldj;
      }
      return;
    }
    if (cg != 0) return;                             // warps 0..3 reduce the tile (named barrier 3, 128 threads)
    const int64_t g = (tile_idx % a.tiles_per_image) * kRows + rowi;
    float* s_v = reinterpret_cast<float*>(smem + kOffRed);
    long long* s_i = reinterpret_cast<long long*>(smem + kOffRed + 32);
    float lp = T.ldj;
    if (a.fisher_A != nullptr) {                     // MatrixFisherN._log_prob (utils/fisher.py:217-232)
      float tr = 0.0f;
#pragma unroll
      for (int i = 0; i < 9; ++i) tr = fmaf(__ldg(a.fisher_A + T.img * 9 + i), T.R[i], tr);
      lp += tr - __ldg(a.fisher_c + T.img);
    }
    if (!T.valid) lp = -INFINITY;
    if (a.logp_out != nullptr && T.valid) a.logp_out[T.img * a.G + g] = lp;
    float bv = lp;
    long long bi = T.valid ? (long long)g : 0x7fffffffffffffffLL;
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
    float e = (T.valid && m > -INFINITY) ? expf(lp - m) : 0.0f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
    named_bar(3, 128);
    if (lane == 0) s_v[warp] = e;
    named_bar(3, 128);
    if (rowi == 0) {
      const float sum = (s_v[0] + s_v[1]) + (s_v[2] + s_v[3]);
      float* p = a.part + tile_idx * 4;
      p[0] = m;
      p[1] = sum;
      p[2] = __int_as_float((int)(bi & 0xffffffffLL));
      p[3] = __int_as_float((int)(bi >> 32));
    }
    named_bar(3, 128);
  };

  // Apply the affine layers that follow, write / reload rows at the end of the stack, stop in front of the next Mobius layer.
  auto advance = [&](Tile& T, int t) {
    while (true) {
      if (T.li == a.n_layers) {
        store_rows(T, t);
        T.item += 1;
        T.li = 0;
        if (T.item >= my_items) { T.item = -1; return; }
        load_rows(T, t);
      }
      const LayerDev L = a.layers[T.li];
      if (L.kind == RNF_LAYER_MOBIUS) { T.perm = L.perm; return; }
      const float* W = L.cond_slot >= 0
                           ? a.cond + T.img * a.cond_stride + (int64_t)a.n_mobius_slots * kH + (int64_t)L.cond_slot * kAffFloats
                           : a.weights + L.w_off;
      float Wr[17];
#pragma unroll
      for (int i = 0; i < 17; ++i) Wr[i] = __ldg(W + i);
      const float loglen = quat_affine_fast(Wr, T.R);
      if (L.has_ldj) T.ldj += Wr[16] - 4.0f * loglen;
      T.li += 1;
    }
  };

  // GEMM g of tile t: g = 0..2 hidden (piece g, N = 64), g = 3 fc_last (piece 3, N = 256 as two N = 128 instructions)
  auto issue_gemm = [&](int t, int g, int64_t m) {
    mbar_wait(bars + 8 * (BAR_READY + t), par_ready & 1u);            // every thread stored its part of the A operand
    par_ready ^= 1u;
    mbar_wait(bars + 8 * (BAR_W_FULL + g), (uint32_t)(m & 1));         // piece g of Mobius step m has landed
    tc_fence_after();
    const uint32_t a_hi_d = umma_desc_lo(smem_u32(smem + kOffA + t * 32768)), a_lo_d = a_hi_d + (16384 >> 4);
    const uint32_t d = tmem_base + t * 256;
    if (g < 3) {
      const uint32_t wb = w_hid_d + g * (kW1Bytes >> 4);
      issue_split_gemm(d, a_hi_d, a_lo_d, wb, wb + (8192 >> 4), umma_idesc(128, 64));
    } else {
      issue_split_gemm(d, a_hi_d, a_lo_d, w_last_d, w_last_d + (32768 >> 4), umma_idesc(128, 128));
      issue_split_gemm(d + 128, a_hi_d, a_lo_d, w_last_d + (16384 >> 4), w_last_d + ((32768 + 16384) >> 4), umma_idesc(128, 128));
    }
    umma_commit(bars + 8 * (BAR_MMA + t));
  };

  // ---- one stage of the MLP chain of tile t at Mobius step m (stage 0 = first layer, 1..3 = hidden epilogues) -----------
  auto chain_stage = [&](Tile& T, int t, int s, int64_t m) {
    const int abuf = (int)(m & 1);
    const uint8_t* aux = smem + kOffAux + abuf * kAuxStride;
    uint8_t* a_hi = smem + kOffA + t * 32768;
    float act[16];
    if (s == 0) {
      mbar_wait(bars + 8 * (BAR_AUX_FULL + abuf), (uint32_t)((m >> 1) & 1));
      const float4* sFirst = reinterpret_cast<const float4*>(aux);
      const int p0 = T.perm, p1 = (T.perm + 1) % 3;
      float x[3], y[3];
      get_col(T.R, p0, x);
      get_col(T.R, p1, y);
      make_frame_fast(x, y, T.P);
      T.zr = dot3(x, T.P.r);
      T.zv = dot3(x, T.P.v);
      const LayerDev L = a.layers[T.li];
      const float* cimg = (L.cond_slot >= 0 && a.cond != nullptr)
                              ? a.cond + T.img * a.cond_stride + (int64_t)L.cond_slot * kH + 16 * cg : nullptr;
      float h0[16];
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
          act[4 * j4 + e] = hv;
        }
      }
      tmem_st16(tm_lane + 256 * t + 64 + 16 * cg, h0);
    } else {
      const int l = s - 1;
      mbar_wait(bars + 8 * (BAR_MMA + t), (par_mma >> t) & 1u);
      par_mma ^= (1u << t);
      tc_fence_after();
      // piece l is dead once GEMM l of the SECOND tile to use it (tile 1) has completed
      if (t == 1 && tid == 256 && m + 1 < total_steps) load_piece((int)((uint32_t)(m + 1) % (uint32_t)n_mob), l, 0);
      tmem_ld16(tm_lane + 256 * t + 16 * cg, act);
      const float4* bias4 = reinterpret_cast<const float4*>(aux + 1024) + (64 * l + 16 * cg) / 4;
      if (l < 2) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = bias4[j4];
          act[4 * j4 + 0] = fmaf(act[4 * j4 + 0], kWUnscale, b.x);
          act[4 * j4 + 1] = fmaf(act[4 * j4 + 1], kWUnscale, b.y);
          act[4 * j4 + 2] = fmaf(act[4 * j4 + 2], kWUnscale, b.z);
          act[4 * j4 + 3] = fmaf(act[4 * j4 + 3], kWUnscale, b.w);
        }
      } else {                                       // relu_last(x0 + x)   (flow/condition.py:29)
        float h0[16];
        tmem_ld16(tm_lane + 256 * t + 64 + 16 * cg, h0);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b = bias4[j4];
          act[4 * j4 + 0] = fmaf(act[4 * j4 + 0], kWUnscale, b.x) + h0[4 * j4 + 0];
          act[4 * j4 + 1] = fmaf(act[4 * j4 + 1], kWUnscale, b.y) + h0[4 * j4 + 1];
          act[4 * j4 + 2] = fmaf(act[4 * j4 + 2], kWUnscale, b.z) + h0[4 * j4 + 2];
          act[4 * j4 + 3] = fmaf(act[4 * j4 + 3], kWUnscale, b.w) + h0[4 * j4 + 3];
        }
      }
    }
    store_a16(a_hi, a_hi + 16384, rowi, cg, act);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bars + 8 * (BAR_READY + t));
    if (tid == 256 * t) issue_gemm(t, s, m);         // thread 0 issues for tile 0, thread 256 for tile 1
  };

  // ---- quarter s of the mixture of tile t at Mobius step m: 4 of my 16 components, straight from TMEM ---------------------
  auto mix_part = [&](Tile& T, int t, int s, int64_t m) {
    const uint8_t* aux = smem + kOffAux + (int)(m & 1) * kAuxStride;
    if (s == 0) {
      mbar_wait(bars + 8 * (BAR_MMA + t), (par_mma >> t) & 1u);        // fc_last of this tile
      par_mma ^= (1u << t);
      tc_fence_after();
      // fc_last weights are dead once tile 1 (their second user) has completed its GEMM
      if (t == 1 && tid == 256 && m + 1 < total_steps) load_piece((int)((uint32_t)(m + 1) % (uint32_t)n_mob), 3, 0);
      T.S_sp = 0.0f; T.S_th = 0.0f; T.S_f = 0.0f;
    }
    float acc[16];
    tmem_ld16(tm_lane + 256 * t + 64 * cg + 16 * s, acc);
    const float4* b4 = reinterpret_cast<const float4*>(aux + 1792) + (64 * cg + 16 * s) / 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 b = b4[k];
      const float sp = softplus_fast(fmaf(acc[4 * k], kWUnscale, b.x));
      float al, be, omw, th, f;
      comp_prep2(T.P, fmaf(acc[4 * k + 1], kWUnscale, b.y), fmaf(acc[4 * k + 2], kWUnscale, b.z),
                 fmaf(acc[4 * k + 3], kWUnscale, b.w), al, be, omw);
      comp_eval2(T.zr, T.zv, al, be, omw, th, f);
      T.S_sp += sp;
      T.S_th = fmaf(sp, th, T.S_th);
      T.S_f = fmaf(sp, f, T.S_f);
    }
  };

  // ---- combine the four column groups of every row of tile t, move the rotation, walk to the next Mobius layer --------------
  auto finish = [&](Tile& T, int t, int64_t m) {
    float* mine = xchg + (cg * 3) * 128 + rowi;
    mine[0] = T.S_sp; mine[128] = T.S_th; mine[256] = T.S_f;
    tc_fence_before();
    __syncthreads();
    // the side data of step m is dead once tile 1 has finished its mixture: refill the buffer with step m + 2
    if (t == 1 && tid == 0 && m + 2 < total_steps) load_piece((int)((uint32_t)(m + 2) % (uint32_t)n_mob), 4, (int)(m & 1));
    const float* px = xchg + rowi;
    const float S_sp = (px[0] + px[384]) + (px[768] + px[1152]);
    const float S_th = (px[128] + px[512]) + (px[896] + px[1280]);
    const float S_f = (px[256] + px[640]) + (px[1024] + px[1408]);
    // (no second barrier: xchg is next written half a step later, and no thread can get there before every thread has
    //  arrived on ready[X] of the next chain stage 0, i.e. after its reads above)
    const float inv_sp = rcp_nr(S_sp);
    const int p0 = T.perm, p1 = (T.perm + 1) % 3, p2 = (T.perm + 2) % 3;
    float nx[3], nz[3], y[3];
    get_col(T.R, p1, y);
    circle_point(T.P.r, T.P.v, S_th * inv_sp, nx);
    T.ldj += logf(S_f * inv_sp);
    cross3(nx, y, nz);
    normalize3_fast(nz);
    set_col(T.R, p0, nx);
    set_col(T.R, p2, nz);
    T.li += 1;
    advance(T, t);
  };

  Tile T0, T1;
  T0.item = T1.item = 0;
  T0.li = T1.li = 0;
  if (my_items > 0) {
    load_rows(T0, 0);
    advance(T0, 0);
    load_rows(T1, 1);
    advance(T1, 1);
  }
  // Half-step h: tile (h & 1) runs the chain of its Mobius step h >> 1 while the other tile runs the mixture of the step it
  // chained half a step earlier.  Tile 0 leads tile 1 by half a step; 2 * total_steps + 1 half-steps in all.
  if (total_steps > 0) {
#pragma unroll 1
    for (int64_t h = 0; h <= 2 * total_steps; h += 2) {
      {   // even half-step: chain tile 0 (step h/2), mixture tile 1 (step h/2 - 1)
        const int64_t m = h >> 1;
        const bool chain_on = m < total_steps, mix_on = h >= 2;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (chain_on) chain_stage(T0, 0, s, m);
          if (mix_on) mix_part(T1, 1, s, m - 1);
        }
        if (mix_on) finish(T1, 1, m - 1);
      }
      if (h + 1 <= 2 * total_steps) {   // odd half-step: chain tile 1 (step h/2), mixture tile 0 (step h/2)
        const int64_t m = h >> 1;
        const bool chain_on = m < total_steps;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (chain_on) chain_stage(T1, 1, s, m);
          mix_part(T0, 0, s, m);
        }
        finish(T0, 0, m);
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

cudaError_t launch_flow_tc(const FlowArgs& a, bool inverse, int sm_count, cudaStream_t st);

cudaError_t launch_flow_tc2(const FlowArgs& a, int sm_count, cudaStream_t st, bool has_mobius) {
  if (!has_mobius) return launch_flow_tc(a, false, sm_count, st);     // affine-only stacks: nothing to pipeline
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
