// flow_t4x.cu -- forward / grid flow kernel, TWO warps per (tile, lane quarter): owner + helper (tcgen05 + TMEM), sm_100a.
//
// flow_t4.cu (one thread per rotation, four 128-rotation tiles per SM = the TMEM ceiling) leaves the issue slots 58 % busy:
// ncu shows a third of the warp samples in the mbarrier wait of the tile's own dependent GEMM chain and on average only 2.8 of
// the 4 warps of a scheduler runnable.  The tile count cannot grow (4 x 128 TMEM columns), so this kernel adds thread-level
// parallelism INSIDE a tile instead: every (tile, lane quarter) is served by two warps of the same scheduler,
//   * the OWNER (warps 0..4T-1) keeps the rotation (R, running log-det) in registers, builds the frame and the Y block, runs the
//     per-layer tail (new column, log-det, quaternion affine layers) and issues the tile's MMAs -- exactly flow_t4.cu's thread --
//   * the HELPER (warps 4T..8T-1) owns no rotation: it takes the upper half of every epilogue (accumulator columns 32..63 ->
//     ReLU -> fp16 hi / lo -> A operand) and the upper four of the eight mixture pairs of every fc_last chunk (columns 32..63),
//     reading the row's frame (r, v, zr: 7 floats) from shared memory and handing back three partial sums per layer.
// Both halves of every GEMM round trip and of every chunk's arithmetic therefore take half as long, and a scheduler has eight
// warps to pick from.  Nothing is computed twice: the per-row scalar work stays with the owner, the helper waits meanwhile.
// Synchronisation: the tile's named barrier (owner + helper warps, 256 threads) hands the accumulator / A operand to the
// tensor core as in flow_t4.cu; one more named barrier per tile returns the helper's partial sums.
// TMEM, weight pipeline (cp.async.bulk pieces refilled by the last tile done with them), MMA issue, grid reduction: flow_t4.cu.
#include "mobius_pair.cuh"
#include "tc_common.cuh"

namespace rnf {
namespace {

#ifndef RNF_X2_TILES
#define RNF_X2_TILES 4
#endif
#ifndef RNF_X2_WAIT_BAR
#define RNF_X2_WAIT_BAR 1
#endif
#ifndef RNF_X2_REGS_OWNER
#define RNF_X2_REGS_OWNER 0      // 0: no setmaxnreg (every warp runs with the launch-bound register count)
#endif
#ifndef RNF_X2_REGS_HELPER
#define RNF_X2_REGS_HELPER 0
#endif
constexpr int kTiles = RNF_X2_TILES;
constexpr int kOwnerThreads = kTiles * 128;
constexpr int kThreads = 2 * kOwnerThreads;
constexpr int kRows = 128;
constexpr int kTileThreads = 256;                         // owner + helper threads of a tile

// shared memory (bytes from a 1024-aligned base)
constexpr int kOffW = 0;                                  // W1 | W2 | W3 pieces
constexpr int kOffLastW = kHidW;                          // W4 piece
constexpr int kOffAux = kOffLastW + kLastW;               // aux piece, double buffered on the layer parity
constexpr int kOffY = kOffAux + 2 * kAuxStride;           // [tile] Y block [128 x 16] fp16, no swizzle (4 KB)
constexpr int kOffC = kOffY + kTiles * 4096;              // [tile] per-image block [64 x 16] fp16, no swizzle (2 KB)
constexpr int kOffP = kOffC + kTiles * 2048;              // [7][tiles * 128] fp32: frame (r, v) and zr of every row (owner -> helper)
constexpr int kOffS = kOffP + 7 * kOwnerThreads * 4;      // [3][tiles * 128] fp32: the helper's partial sums (helper -> owner)
constexpr int kOffRed = kOffS + 3 * kOwnerThreads * 4;    // [tile] reduction scratch
constexpr int kOffBar = kOffRed + kTiles * 128;
constexpr int kOffMisc = kOffBar + 8 * 16;
constexpr int kSmemBytes = kOffMisc + 64 + 64 * 8;
constexpr int kSmemAlloc = kSmemBytes + 1024;
static_assert(kOffLastW % 1024 == 0 && kW1Bytes % 1024 == 0, "UMMA SW128 tiles need 1024 B alignment");
static_assert(kOffY % 16 == 0 && kOffC % 16 == 0 && kOffAux % 16 == 0, "no-swizzle blocks need 16 B alignment");
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KB shared-memory limit of an sm_100 CTA");

enum { BAR_W_FULL = 0 /* W1,W2,W3,W4 */, BAR_AUX_FULL = 4 /* [2] */, BAR_MMA = 6 /* [tile] */, BAR_COUNT = 6 + RNF_X2_TILES };
// named barriers: 1 + tile = hand-over (256 threads), 5 + tile = partial sums (256; the owners' grid reduction between
// items, 128), 9 + tile = sleeping through a GEMM round trip (256)

constexpr uint32_t kColAhi = 0, kColAlo = 32, kColD = 64, kColsPerTile = 128;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t desc_hi32, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo32), "r"(desc_hi32), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void issue_split_ts(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t idesc) {
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16_ts(d, a_lo + 8 * k, b_hi + 2 * k, kDescHi, idesc, 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16_ts(d, a_hi + 8 * k, b_lo + 2 * k, kDescHi, idesc, 1);
#pragma unroll
  for (int k = 0; k < 4; ++k) umma_f16_ts(d, a_hi + 8 * k, b_hi + 2 * k, kDescHi, idesc, 1);
}

__device__ __forceinline__ uint32_t pack_h2(__half lo, __half hi) {
  const __half2 h = __halves2half2(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__device__ __forceinline__ void relu_split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float2 back = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float d0, d1;
  upk(sub2(pk(x0, x1), pk(back.x, back.y)), d0, d1);
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(d1), "f"(d0));
}

__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t u[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(u[0]), "r"(u[1]),
               "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
               : "memory");
}

// Half of one GEMM's epilogue: accumulator columns 32 half .. 32 half + 31 of my row -> (+ c) -> ReLU -> fp16 hi / lo ->
// A operand columns 16 half .. 16 half + 15 of both planes.  Two 16-column pieces, the second load in flight under the first
// piece's arithmetic; the stores are waited for once.
__device__ __forceinline__ void epilogue_half(uint32_t tm, int half, const float* cadd) {
  float acc0[16], acc1[16];
  tmem_ld16_async(tm + kColD + 32 * half, acc0);
  tmem_ld16_async(tm + kColD + 32 * half + 16, acc1);
  tmem_ld_wait16(acc0);
  tmem_ld_wait16(acc1);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    float* acc = p == 0 ? acc0 : acc1;
    if (cadd != nullptr) {
#pragma unroll
      for (int j4 = 0; j4 < 4; ++j4) {
        const float4 c = __ldg(reinterpret_cast<const float4*>(cadd + 32 * half + 16 * p) + j4);
        acc[4 * j4] += c.x; acc[4 * j4 + 1] += c.y; acc[4 * j4 + 2] += c.z; acc[4 * j4 + 3] += c.w;
      }
    }
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) relu_split_pair(acc[2 * e], acc[2 * e + 1], hi[e], lo[e]);
    tmem_st8_nowait(tm + kColAhi + 16 * half + 8 * p, hi);
    tmem_st8_nowait(tm + kColAlo + 16 * half + 8 * p, lo);
  }
  tmem_st_wait();
}

template <bool GRID>
__global__ void __launch_bounds__(kThreads, 1) flow_t4x_kernel(const FlowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int half = warp / (4 * kTiles);              // 0 = owner, 1 = helper (warp-uniform)
  const int tile = (warp >> 2) % kTiles;
  const int rowi = (warp & 3) * 32 + lane;           // row inside the tile = TMEM lane
  const int slot = tile * kRows + rowi;              // row slot in the shared-memory exchange arrays
  const bool issuer_warp = half == 0 && (warp & 3) == (tile & 3);   // one owner warp per tile, on a different scheduler per tile
  const bool elected = issuer_warp && lane == 0;
  const uint32_t bars = smem_u32(smem + kOffBar);

  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  int* s_cnt = reinterpret_cast<int*>(smem + kOffMisc + 4);                // tiles done with: [0..2] W1..W3, [3] W4, [4 + buf] aux
  long long* s_moff = reinterpret_cast<long long*>(smem + kOffMisc + 64);
  float* s_P = reinterpret_cast<float*>(smem + kOffP);
  float* s_S = reinterpret_cast<float*>(smem + kOffS);

  int n_mob = 0;
  for (int i = 0; i < a.n_layers; ++i)
    if (a.layers[i].kind == RNF_LAYER_MOBIUS) {
      if (tid == 0) s_moff[n_mob] = a.layers[i].w_off_tc;
      ++n_mob;
    }
  if (tid == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) mbar_init(bars + 8 * i, 1);
    for (int i = 0; i < 6; ++i) s_cnt[i] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < (kTiles * (4096 + 2048)) / 16; i += kThreads) reinterpret_cast<uint4*>(smem + kOffY)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
#if RNF_X2_REGS_OWNER > 0
  if (half == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(RNF_X2_REGS_HELPER));
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(RNF_X2_REGS_OWNER));
#endif
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tm_tile = tmem_base + (uint32_t)tile * kColsPerTile;                   // lane 0 of the tile (MMA addresses)
  const uint32_t tm = tm_tile + ((uint32_t)((warp & 3) * 32) << 16);                    // my warp's lane quarter

  const int64_t n_groups = (a.n_tiles + kTiles - 1) / kTiles;
  const int64_t my_items = blockIdx.x < n_groups ? (n_groups - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int64_t total_steps = my_items * n_mob;
  const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(a.weights);
  auto load_piece = [&](int mob_idx, int piece, int abuf) {
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

  // bar_red (owners only, 128, between items) shares the id of bar_sum: the helpers' next arrival there lies behind the next
  // item's first GEMM, i.e. behind the reduction; the hand-over id would NOT do (helpers arrive early for the next Y block)
  const int bar_tile = 1 + tile, bar_sum = 5 + tile, bar_red = 5 + tile, bar_wait = 9 + tile;
  const uint32_t bar_mma = bars + 8 * (BAR_MMA + tile);
  uint32_t par_mma = 0;
  int64_t step = 0;
  constexpr uint32_t kIdesc = umma_idesc(128, 64);

  auto wait_mma = [&]() {
    mbar_wait(bar_mma, par_mma);
    par_mma ^= 1;
    tc_fence_after();
  };
  // long waits (the dependent GEMM round trips): only the issuing warp polls the mbarrier, the tile's other seven warps sleep in
  // a hardware named barrier (no issue slots; the polling loop was 45 % of this kernel's executed instructions otherwise)
  auto wait_mma_long = [&]() {
#if RNF_X2_WAIT_BAR
    if (issuer_warp) {
      mbar_wait(bar_mma, par_mma);
      tc_fence_before();
      named_arrive(bar_wait, kTileThreads);
    } else {
      named_bar(bar_wait, kTileThreads);
    }
    par_mma ^= 1;
    tc_fence_after();
#else
    wait_mma();
#endif
  };

  if (half == 1) {
    // =========================================== helper warps ===========================================
    for (int64_t item = 0; item < my_items; ++item) {
      const int64_t tile_idx = kTiles * (blockIdx.x + item * (int64_t)gridDim.x) + tile;
      const float* cond_img = nullptr;
      if (!GRID && a.cond != nullptr && tile_idx < a.n_tiles) {       // row mode adds the per-image term in the epilogue
        const int64_t row = tile_idx * kRows + rowi;
        int64_t img = 0;
        if (row < a.N) img = a.feat_index != nullptr ? (int64_t)__ldg(a.feat_index + row) : row / a.rows_per_image;
        cond_img = a.cond + img * a.cond_stride;
      }
#pragma unroll 1
      for (int li = 0; li < a.n_layers; ++li) {
        const LayerDev L = a.layers[li];
        if (L.kind != RNF_LAYER_MOBIUS) continue;
        const float* cadd = (!GRID && L.cond_slot >= 0 && cond_img != nullptr) ? cond_img + (int64_t)L.cond_slot * kH : nullptr;
        tc_fence_before();                            // my TMEM reads of the previous layer are done
        named_arrive(bar_tile, kTileThreads);         // Y block hand-over (the owners' part)
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
          wait_mma_long();
          epilogue_half(tm, 1, (l == 0 || l == 3) ? cadd : nullptr);
          tc_fence_before();
          named_arrive(bar_tile, kTileThreads);
        }
        // the owner wrote the frame before its Y hand-over, which every GEMM of this layer is ordered after
        Plane P;
        P.r[0] = s_P[0 * kOwnerThreads + slot]; P.r[1] = s_P[1 * kOwnerThreads + slot]; P.r[2] = s_P[2 * kOwnerThreads + slot];
        P.v[0] = s_P[3 * kOwnerThreads + slot]; P.v[1] = s_P[4 * kOwnerThreads + slot]; P.v[2] = s_P[5 * kOwnerThreads + slot];
        const float zr = s_P[6 * kOwnerThreads + slot];
        f32x2 S_sp2 = 0ull, S_at2 = 0ull, S_f2 = 0ull;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float buf0[16], buf1[16];
          if (c == 0) wait_mma_long(); else wait_mma();
          tmem_ld16_async(tm + kColD + 32, buf0);
          tmem_ld16_async(tm + kColD + 48, buf1);
          tmem_ld_wait16(buf0);
          tmem_ld_wait16(buf1);
          if (c < 3) {
            tc_fence_before();
            named_arrive(bar_tile, kTileThreads);
          }
          mixture_pairs<2, true>(P, zr, 0.0f, buf0, S_sp2, S_at2, S_f2);
          mixture_pairs<2, true>(P, zr, 0.0f, buf1, S_sp2, S_at2, S_f2);
        }
        s_S[0 * kOwnerThreads + slot] = hsum(S_sp2);
        s_S[1 * kOwnerThreads + slot] = hsum(S_at2);
        s_S[2 * kOwnerThreads + slot] = hsum(S_f2);
        named_arrive(bar_sum, kTileThreads);
        ++step;
      }
    }
  } else {
    // =========================================== owner warps ===========================================
    uint8_t* y_blk = smem + kOffY + tile * 4096;
    uint8_t* c_blk = smem + kOffC + tile * 2048;
    const uint32_t y_d = umma_desc_lo_ns(smem_u32(y_blk)), c_d = umma_desc_lo_ns(smem_u32(c_blk));
    const uint32_t w_hid_d = umma_desc_lo(smem_u32(smem + kOffW)), w_last_d = umma_desc_lo(smem_u32(smem + kOffLastW));
    const uint32_t bias_hid_d = umma_desc_lo_ns(smem_u32(smem + kOffW + 16384)), bias_last_d = umma_desc_lo_ns(smem_u32(smem + kOffLastW + 65536));
    const uint32_t aux_blk_d = umma_desc_lo_ns(smem_u32(smem + kOffAux + kAuxFirst));
    uint32_t par_w = 0;
    int mob_cur = 0;
    auto hand_over = [&]() {
      tc_fence_before();
      if (issuer_warp) named_bar(bar_tile, kTileThreads); else named_arrive(bar_tile, kTileThreads);
    };

    for (int64_t item = 0; item < my_items; ++item) {
      const int64_t tile_idx = kTiles * (blockIdx.x + item * (int64_t)gridDim.x) + tile;
      int64_t row = 0, img = 0, g = 0;
      bool valid = false;
      float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
      if (tile_idx < a.n_tiles) {
        if (GRID) {
          img = tile_idx / a.tiles_per_image;
          g = (tile_idx % a.tiles_per_image) * kRows + rowi;
          valid = g < a.G;
          row = img * a.G + g;
          if (valid) {
            float Gm[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) Gm[i] = __ldg(a.R_in + g * 9 + i);
            if (a.offset != nullptr) {                 // samples = grid @ random_rot (eval.py:439-440)
              float O[9];
#pragma unroll
              for (int i = 0; i < 9; ++i) O[i] = __ldg(a.offset + i);
#pragma unroll
              for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j)
                  R[3 * i + j] = fmaf(Gm[3 * i + 2], O[6 + j], fmaf(Gm[3 * i + 1], O[3 + j], Gm[3 * i] * O[j]));
            } else {
#pragma unroll
              for (int i = 0; i < 9; ++i) R[i] = Gm[i];
            }
          }
        } else {
          row = tile_idx * kRows + rowi;
          valid = row < a.N;
          if (valid) {
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = __ldg(a.R_in + row * 9 + i);
            if (a.cond != nullptr) img = a.feat_index != nullptr ? (int64_t)__ldg(a.feat_index + row) : row / a.rows_per_image;
          }
        }
      }
      const float* cond_img = a.cond != nullptr ? a.cond + img * a.cond_stride : nullptr;
      float ldj = 0.0f;
      float dgt = 0.0f;
      if (GRID && a.gt != nullptr && valid) dgt = gt_distance(a.gt + img * a.gt_k * 9, a.gt_k, R);

#pragma unroll 1
      for (int li = 0; li < a.n_layers; ++li) {
        const LayerDev L = a.layers[li];
        if (L.kind != RNF_LAYER_MOBIUS) {
          const float* W = L.cond_slot >= 0 ? cond_img + (int64_t)a.n_mobius_slots * kH + (int64_t)L.cond_slot * kAffFloats
                                            : a.weights + L.w_off;
          float Wr[17];
#pragma unroll
          for (int i = 0; i < 17; ++i) Wr[i] = __ldg(W + i);
          const float loglen = quat_affine_fast(Wr, R);
          if (L.has_ldj) ldj += Wr[16] - 4.0f * loglen;
          continue;
        }
        // ================================ Mobius layer ================================
        const int p0 = L.perm, p1 = (L.perm + 1) % 3, p2 = (L.perm + 2) % 3;
        float x[3], y[3];
        Plane P;
        get_col(R, p0, x);
        get_col(R, p1, y);
        make_frame_fast(x, y, P);
        const float zr = dot3(x, P.r);                      // in-plane coordinate of the moving column (its v coordinate is 0)
        const float* cimg = (L.cond_slot >= 0 && cond_img != nullptr) ? cond_img + (int64_t)L.cond_slot * kH : nullptr;
        const bool c_by_mma = GRID && cimg != nullptr;      // warp- and tile-uniform
        const float* cadd = GRID ? nullptr : cimg;
        const int abuf = (int)(step & 1);
        const int mob_n1 = mob_cur + 1 >= n_mob ? mob_cur + 1 - n_mob : mob_cur + 1;
        const int mob_n2 = mob_n1 + 1 >= n_mob ? mob_n1 + 1 - n_mob : mob_n1 + 1;

        // ---- frame for the helper; my row of the Y block: (1, 1, y_hi, y_lo | y_hi, 1, 1, 0, 0, 0); per-image block ----
        s_P[0 * kOwnerThreads + slot] = P.r[0]; s_P[1 * kOwnerThreads + slot] = P.r[1]; s_P[2 * kOwnerThreads + slot] = P.r[2];
        s_P[3 * kOwnerThreads + slot] = P.v[0]; s_P[4 * kOwnerThreads + slot] = P.v[1]; s_P[5 * kOwnerThreads + slot] = P.v[2];
        s_P[6 * kOwnerThreads + slot] = zr;
        {
          __half yh[3], yl[3];
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            yh[i] = __float2half_rn(y[i]);
            yl[i] = __float2half_rn(y[i] - __half2float(yh[i]));
          }
          const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
          uint8_t* yrow = y_blk + (rowi >> 3) * 256 + (rowi & 7) * 16;
          *reinterpret_cast<uint4*>(yrow) = make_uint4(pack_h2(one, one), pack_h2(yh[0], yh[1]), pack_h2(yh[2], yl[0]), pack_h2(yl[1], yl[2]));
          *reinterpret_cast<uint4*>(yrow + 128) = make_uint4(pack_h2(yh[0], yh[1]), pack_h2(yh[2], one), pack_h2(one, zero), 0u);
          if (c_by_mma && rowi < 64) {
            const float cv = __ldg(cimg + rowi);
            const __half ch = __float2half_rn(cv);
            const __half cl = __float2half_rn(cv - __half2float(ch));
            *reinterpret_cast<uint32_t*>(c_blk + (rowi >> 3) * 256 + (rowi & 7) * 16) = pack_h2(ch, cl);
          }
        }
        fence_proxy_async();
        hand_over();
        // ---- fc_first and three hidden layers: four dependent GEMM round trips ----
#pragma unroll 1
        for (int l = 0; l < 4; ++l) {
          if (issuer_warp) {
            if (l == 0) mbar_wait(bars + 8 * (BAR_AUX_FULL + abuf), (uint32_t)((step >> 1) & 1));
            else mbar_wait(bars + 8 * (BAR_W_FULL + l - 1), (par_w >> (l - 1)) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
              const uint32_t d = tm_tile + kColD;
              if (l == 0) {
                umma_f16(d, y_d, aux_blk_d + abuf * (kAuxStride >> 4), kDescHiNS, kIdesc, 0);
              } else {
                const uint32_t wb = w_hid_d + (l - 1) * (kW1Bytes >> 4);
                umma_f16(d, y_d, bias_hid_d + (l - 1) * (kW1Bytes >> 4), kDescHiNS, kIdesc, 0);
                issue_split_ts(d, tm_tile + kColAhi, tm_tile + kColAlo, wb, wb + (8192 >> 4), kIdesc);
              }
              if (c_by_mma && (l == 0 || l == 3)) umma_f16(d, y_d, c_d, kDescHiNS, kIdesc, 1);
              umma_commit(bar_mma);
            }
            __syncwarp();
          }
          wait_mma_long();
          if (elected) {
            if (l == 0) { if ((atomicAdd(&s_cnt[4 + abuf], 1) % kTiles) == kTiles - 1 && step + 2 < total_steps) load_piece(mob_n2, 4, abuf); }
            else if ((atomicAdd(&s_cnt[l - 1], 1) % kTiles) == kTiles - 1 && step + 1 < total_steps) load_piece(mob_n1, l - 1, 0);
          }
          epilogue_half(tm, 0, (l == 0 || l == 3) ? cadd : nullptr);
          hand_over();
        }
        // ---- fc_last in four N = 64 chunks through the single accumulator; my half: pairs 0..3 (columns 0..31) ----
        auto issue_chunk = [&](int c) {               // issuing warp only, right after the hand-over barrier
          if (c == 0) mbar_wait(bars + 8 * (BAR_W_FULL + 3), (par_w >> 3) & 1u);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t d = tm_tile + kColD;
            const uint32_t wb = w_last_d + c * (8192 >> 4);                  // rows 64c .. 64c+63 of the hi plane
            umma_f16(d, y_d, bias_last_d + c * (2048 >> 4), kDescHiNS, kIdesc, 0);
            issue_split_ts(d, tm_tile + kColAhi, tm_tile + kColAlo, wb, wb + (32768 >> 4), kIdesc);
            umma_commit(bar_mma);
          }
          __syncwarp();
        };
        if (issuer_warp) issue_chunk(0);
        f32x2 S_sp2 = 0ull, S_at2 = 0ull, S_f2 = 0ull;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float buf0[16], buf1[16];
          if (c == 0) wait_mma_long(); else wait_mma();
          if (c == 3 && elected && (atomicAdd(&s_cnt[3], 1) % kTiles) == kTiles - 1 && step + 1 < total_steps) load_piece(mob_n1, 3, 0);
          tmem_ld16_async(tm + kColD, buf0);
          tmem_ld16_async(tm + kColD + 16, buf1);
          tmem_ld_wait16(buf0);
          tmem_ld_wait16(buf1);
          if (c < 3) {                                  // accumulator drained: the next chunk runs under the math below
            hand_over();
            if (issuer_warp) issue_chunk(c + 1);
          }
          mixture_pairs<2, true>(P, zr, 0.0f, buf0, S_sp2, S_at2, S_f2);
          mixture_pairs<2, true>(P, zr, 0.0f, buf1, S_sp2, S_at2, S_f2);
        }
        if (issuer_warp) par_w ^= 0xFu;
        named_bar(bar_sum, kTileThreads);               // the helper's partial sums are in shared memory
        const float S_sp = hsum(S_sp2) + s_S[0 * kOwnerThreads + slot];
        const float S_at = hsum(S_at2) + s_S[1 * kOwnerThreads + slot];
        const float S_f = hsum(S_f2) + s_S[2 * kOwnerThreads + slot];
        float nx[3], nz[3];
        const float inv_sp = rcp_nr(S_sp);
        circle_point_fast(P.r, P.v, mixture_angle(S_at, inv_sp), nx);
        ldj += log_fast(S_f * inv_sp);
        cross3(nx, y, nz);
        normalize3_fast(nz);
        set_col(R, p0, nx);
        set_col(R, p2, nz);
        ++step;
        mob_cur = mob_n1;
      }

      // ================================ outputs ================================
      if (!GRID) {
        if (valid) {
#pragma unroll
          for (int i = 0; i < 9; ++i) a.R_out[row * 9 + i] = R[i];
          a.ldj_out[row] = ldj;
        }
      } else if (tile_idx < a.n_tiles) {
        float lp = ldj;
        if (a.fisher_A != nullptr) {
          float tr = 0.0f;
#pragma unroll
          for (int i = 0; i < 9; ++i) tr = fmaf(__ldg(a.fisher_A + img * 9 + i), R[i], tr);
          lp += tr - __ldg(a.fisher_c + img);
        }
        if (!valid) lp = -INFINITY;
        if (a.logp_out != nullptr && valid) a.logp_out[row] = lp;
        float* s_v = reinterpret_cast<float*>(smem + kOffRed + tile * 128);
        long long* s_i = reinterpret_cast<long long*>(smem + kOffRed + tile * 128 + 32);
        const int w4 = warp & 3;
        float bv = lp;
        long long bi = valid ? (long long)g : 0x7fffffffffffffffLL;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) { s_v[w4] = bv; s_i[w4] = bi; }
        named_bar(bar_red, 128);
        bv = s_v[0]; bi = s_i[0];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float ov = s_v[w];
          const long long oi = s_i[w];
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        const float m = bv;
        float e = (valid && m > -INFINITY) ? expf(lp - m) : 0.0f;
        float ed = e * dgt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          e += __shfl_xor_sync(0xffffffffu, e, o);
          ed += __shfl_xor_sync(0xffffffffu, ed, o);
        }
        named_bar(bar_red, 128);
        if (lane == 0) { s_v[w4] = e; s_v[4 + w4] = ed; }
        named_bar(bar_red, 128);
        if (rowi == 0) {
          const float s = (s_v[0] + s_v[1]) + (s_v[2] + s_v[3]);
          float* p = a.part + tile_idx * kPartStride;
          p[0] = m;
          p[1] = s;
          p[4] = (s_v[4] + s_v[5]) + (s_v[6] + s_v[7]);
          p[2] = __int_as_float((int)(bi & 0xffffffffLL));
          p[3] = __int_as_float((int)(bi >> 32));
        }
        named_bar(bar_red, 128);
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

cudaError_t launch_flow_t4x(const FlowArgs& a, int sm_count, cudaStream_t st) {
  const bool grid_mode = a.G > 0;
  void (*kern)(const FlowArgs) = grid_mode ? flow_t4x_kernel<true> : flow_t4x_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
  if (e != cudaSuccess) return e;
  if (a.n_tiles <= 0) return cudaSuccess;
  const int64_t groups = (a.n_tiles + kTiles - 1) / kTiles;
  const int64_t grid = groups < sm_count ? groups : sm_count;
  kern<<<(unsigned)grid, kThreads, kSmemAlloc, st>>>(a);
  return cudaGetLastError();
}

}  // namespace rnf
