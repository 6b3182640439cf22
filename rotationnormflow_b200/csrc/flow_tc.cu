// flow_tc.cu -- fused flow kernel with the conditioner GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// One CTA (512 threads, one per SM, persistent) works on a PAIR of 128-rotation tiles at a time.  A tile's 128 rows are
// the 128 TMEM lanes of its accumulator; every row is owned by two threads (column halves h = 0/1) that both keep the
// rotation's 3x3 matrix and running log|det J| in registers across the whole layer stack (flow/flow.py:53-72).
//
// Per Mobius layer (flow/mobiusflow.py:46-125) and tile:
//   CUDA cores : frame (r, v), first conditioner layer  h0 = W0[:, :3].y + b0 + c_img  (flow/condition.py:25)
//                -> ReLU -> split into fp16 (hi, lo) -> K-major 128B-swizzled A operand in shared memory
//   tensor core: D[128 x 64] = A . W^T as three tcgen05.mma products  Alo.Whi + Ahi.Wlo + Ahi.Whi  (fp32 accumulate in
//                TMEM; error-compensated split => fp32-level accuracy), three times (layers.1/3/5)
//   CUDA cores : tcgen05.ld -> + bias, ReLU (residual on the last) -> split -> A operand
//   tensor core: fc_last  D[128 x 256] in two N=128 chunks, each committed to its own mbarrier
//   CUDA cores : each half walks its 32 mixture components straight out of TMEM (softplus weights, Mobius maps, atan2,
//                analytic Jacobian), the halves exchange three partial sums through shared memory.
// Weights: the host packs, per Mobius layer, the exact shared-memory image (UMMA canonical K-major SWIZZLE_128B tiles of
// fp16 hi / lo planes, scaled by 2^8 so that the lo plane stays in the fp16 normal range, plus the fp32 biases); one
// cp.async.bulk (TMA bulk copy) per region brings it in, signalled on an mbarrier, and is re-issued for the next layer
// as soon as both tiles' MMAs that read the region have completed -- so weight traffic overlaps the mixture math.
// While one tile waits for its MMA, the other tile's 8 warps own the issue slots (ping-pong).
//
// Quaternion affine layers, grid mode (offset, Fisher base term, per-tile max / arg-max / sum-exp) and the row <-> image
// mapping are identical to flow_v1.cu.
#include "tc_common.cuh"

namespace rnf {
namespace {

#ifndef RNF_TC_TURN_RELEASE
#define RNF_TC_TURN_RELEASE -1     // >= 0: explicit ping-pong, GEMM index (0..2 hidden, 3 fc_last) whose issue hands the turn to the other tile (measured: no gain)
#endif

#ifndef RNF_TC_TRACE
#define RNF_TC_TRACE 0
#endif
#if RNF_TC_TRACE
#define TRACE(i) do { if (tr_on) tr[(i)] = clock64(); } while (0)
#else
#define TRACE(i) do { } while (0)
#endif

#ifndef RNF_TC_INTERLEAVE_TILES
#define RNF_TC_INTERLEAVE_TILES 0      // measured: 2 % slower than contiguous warp groups per tile
#endif

#ifndef RNF_TC_NAP_NS
#define RNF_TC_NAP_NS 100
#endif
#ifndef RNF_TC_MIX_YIELD
#define RNF_TC_MIX_YIELD 0      // nanosleep(0) per 4 mixture components = a scheduler yield: +3 % (measured 0..150 ns: same)
#endif
#ifndef RNF_TC_YIELD
#define RNF_TC_YIELD 0      // mixture warps back off while the other tile runs a chain burst (see BUSY below)
#endif

constexpr int kThreads = 512;
constexpr int kRows = 128;                        // rows per tile = TMEM lanes

// ---- shared-memory image (bytes from a 1024-aligned base).  The packed global image of one Mobius conditioner is
//      [W1 | W2 | W3 : each (hi 64x64, lo 64x64) fp16 SW128][W4 : hi 256x64, lo 256x64][aux : first[64][4], b1..b3, b4' fp32]
//      and every piece is brought in by its own bulk copy as soon as its previous contents are dead. ----
constexpr int kOffW = 0;
constexpr int kOffLastW = kHidW;                  // 55296
constexpr int kOffAux = kOffLastW + kLastW;       // 129024, double buffered (layer parity)
constexpr int kOffA = kOffAux + 2 * kAuxStride;   // 131072 = 128 * 1024 : [tile][hi|lo] 128x64 fp16 (16 KB each)
constexpr int kOffOnes = kOffA + 4 * 16384;       // constant [128 x 16] fp16 tile, ones in K columns 0 and 1 (no swizzle)
constexpr int kOffXchg = kOffOnes + 4096;         // [tile][half][3][128] fp32
constexpr int kOffRed = kOffXchg + 2 * 2 * 3 * 128 * 4;   // [tile] reduction scratch
constexpr int kOffBar = kOffRed + 2 * 128;
constexpr int kOffMisc = kOffBar + 8 * 16;        // tmem base, counters, Mobius offset table
constexpr int kSmemBytes = kOffMisc + 32 + 64 * 8;
constexpr int kSmemAlloc = kSmemBytes + 1024;     // slack for manual 1024 B alignment
static_assert(kOffA % 1024 == 0 && kOffLastW % 1024 == 0 && kW1Bytes % 1024 == 0, "UMMA SW128 tiles need 1024 B alignment");
static_assert(kSmemAlloc <= 232448, "exceeds the 227 KB shared-memory limit of an sm_100 CTA");

// mbarrier slots
enum { BAR_W_FULL = 0 /* W1,W2,W3,W4 */, BAR_AUX_FULL = 4 /* [2] */, BAR_MMA = 6 /* [tile][2] */, BAR_COUNT = 10 };

// ReLU + split 32 fp32 pre-activations (columns 32h .. 32h+31 of row r) into fp16 hi / lo and store them into the
// K-major SW128 A operand: element (r, k) lives at (r/8)*1024 + (r%8)*128 + ((k/8) ^ (r%8))*16 + (k%8)*2.
__device__ __forceinline__ void store_a_operand(uint8_t* a_hi, uint8_t* a_lo, int r, int h, const float v[32]) {
  const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      relu_split2(v[8 * c + 2 * e], v[8 * c + 2 * e + 1], hi[e], lo[e]);     // ReLU is applied here
    }
    const int off = rbase + (((4 * h + c) ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct TileCtx {
  int tile;        // 0 / 1 inside the CTA
  int half;        // column half owned by this thread
  int row;         // 0..127 = TMEM lane
  bool elected;    // tile-local thread 0: issues MMAs and weight copies
  uint32_t tmem_d; // TMEM address of this tile's accumulator, lane field = this warp's quarter
  uint32_t bars;   // shared address of the mbarrier array
  uint32_t par_mma0, par_mma1, par_w;               // phase parities (par_w: bit l = weight region l, elected thread only)
};

template <bool INV, bool GRID>
__global__ void __launch_bounds__(kThreads, 1) flow_tc_kernel(const FlowArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // manual 1024 B alignment (SWIZZLE_128B atoms) done on the shared-window address so the pointer stays a shared pointer
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  TileCtx c;
#if RNF_TC_INTERLEAVE_TILES
  // warps {0-3, 8-11} = tile 0, {4-7, 12-15} = tile 1: the scheduler prefers high warp ids, so interleaving the tiles'
  // warp groups keeps either tile from always winning the issue slot
  c.tile = (warp >> 2) & 1;
  c.half = warp >> 3;
  c.row = (warp & 3) * 32 + lane;
  c.elected = tid == 128 * c.tile;
#else
  c.tile = warp >> 3;
  c.half = (warp >> 2) & 1;
  c.row = (warp & 3) * 32 + lane;
  c.elected = (tid & 255) == 0;
#endif
  c.bars = smem_u32(smem + kOffBar);
  // warp-uniform (provably: broadcast from lane 0) flag of the warp that issues this tile's MMAs
  const bool issuer_warp = __shfl_sync(0xffffffffu, (int)c.elected, 0) != 0;
  c.par_mma0 = c.par_mma1 = c.par_w = 0;

  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffMisc);
  int* s_cnt = reinterpret_cast<int*>(smem + kOffMisc + 4);                // consumers done: [0..2] W1..W3, [3] W4, [4] aux
  long long* s_moff = reinterpret_cast<long long*>(smem + kOffMisc + 32);
  // Issue-slot arbitration between the tiles.  The hardware scheduler keeps issuing from the (always ready) mixture warps and
  // starves the other tile's short, latency-critical chain bursts (first layer / epilogues / MMA issue): measured with
  // tools/tc_timeline.py, a tile made no chain progress at all while the other tile was in its mixture.  The chain tile
  // therefore raises s_busy[tile] around its bursts and the mixture loop of the other tile naps while it is up.
  volatile int* s_busy = reinterpret_cast<volatile int*>(smem + kOffMisc + 24);  // w_off_tc of Mobius layers, execution order

  // ---- one-time setup -------------------------------------------------------------------------------------------------
  int n_mob = 0;
  for (int i = 0; i < a.n_layers; ++i) {
    const int li = INV ? a.n_layers - 1 - i : i;
    if (a.layers[li].kind == RNF_LAYER_MOBIUS) {
      if (tid == 0) s_moff[n_mob] = a.layers[li].w_off_tc;
      ++n_mob;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < BAR_COUNT; ++i) mbar_init(c.bars + 8 * i, 1);
    for (int i = 0; i < 5; ++i) s_cnt[i] = 0;
    s_busy[0] = s_busy[1] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // constant ones tile (bias MMA): row r holds 1.0 in K columns 0 and 1, zero elsewhere
  // (no-swizzle layout: per 8-row group 128 B for K 0..7 then 128 B for K 8..15; thread tid writes bytes 8 tid .. 8 tid + 7)
  reinterpret_cast<uint2*>(smem + kOffOnes)[tid] = make_uint2(((tid & 1) == 0 && (tid & 31) < 16) ? 0x3C003C00u : 0u, 0u);
  fence_proxy_async();
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(s_tmem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  c.tmem_d = tmem_base + (uint32_t)(c.tile * 256) + ((uint32_t)((warp & 3) * 32) << 16);

  const int64_t n_pairs = (a.n_tiles + 1) / 2;
  const int64_t my_items = blockIdx.x < n_pairs ? (n_pairs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
  const int64_t total_steps = my_items * n_mob;     // Mobius layer executions of this CTA
  const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(a.weights);
  // One bulk copy per piece of a layer image; `piece` 0..2 = W1..W3, 3 = W4, 4 = aux (into buffer `abuf`).
  auto load_piece = [&](int mob_idx, int piece, int abuf) {   // mob_idx: index into the Mobius layer table
    const uint8_t* src = wbytes + s_moff[mob_idx] * 4;
    uint32_t dst, bytes, bar;
    if (piece < 3) { src += piece * kW1Bytes; dst = kOffW + piece * kW1Bytes; bytes = kW1Bytes; bar = BAR_W_FULL + piece; }
    else if (piece == 3) { src += kHidW; dst = kOffLastW; bytes = kLastW; bar = BAR_W_FULL + 3; }
    else { src += kHidW + kLastW; dst = kOffAux + abuf * kAuxStride; bytes = kAuxBytes; bar = BAR_AUX_FULL + abuf; }
    mbar_expect_tx(c.bars + 8 * bar, bytes);
    bulk_g2s(smem_u32(smem + dst), src, bytes, c.bars + 8 * bar);
  };
  if (tid == 0 && total_steps > 0) {                // prime every region with the first Mobius layer(s)
    for (int piece = 0; piece < 4; ++piece) load_piece(0, piece, 0);
    load_piece(0, 4, 0);
    if (total_steps > 1) load_piece(n_mob > 1 ? 1 : 0, 4, 1);
  }

  uint8_t* a_hi = smem + kOffA + c.tile * 32768;
  uint8_t* a_lo = a_hi + 16384;
  const uint32_t a_hi_d = umma_desc_lo(smem_u32(a_hi)), a_lo_d = umma_desc_lo(smem_u32(a_lo));
  const uint32_t w_hid_d = umma_desc_lo(smem_u32(smem + kOffW)), w_last_d = umma_desc_lo(smem_u32(smem + kOffLastW));
  const uint32_t ones_d = umma_desc_lo_ns(smem_u32(smem + kOffOnes));
  const uint32_t bias_hid_d = umma_desc_lo_ns(smem_u32(smem + kOffW + 16384)), bias_last_d = umma_desc_lo_ns(smem_u32(smem + kOffLastW + 65536));
  float* xchg = reinterpret_cast<float*>(smem + kOffXchg) + c.tile * (2 * 3 * 128);   // [half][slot 0..2][row]
  float* x_mine = xchg + c.half * 384 + c.row;
  const float* x_lo = xchg + c.row;
  const float* x_hi = xchg + 384 + c.row;
  const int bar_tile = 1 + c.tile;                   // named barrier of the tile's 256 threads
  const uint32_t bar_mma0 = c.bars + 8 * (BAR_MMA + 2 * c.tile), bar_mma1 = bar_mma0 + 8;   // hidden GEMMs + chunk A | chunk B
  const uint32_t tm_stash = c.tmem_d + 64 + 32 * c.half;     // h0 of my 32 hidden columns (free TMEM columns)
  const uint32_t tm_mine = c.tmem_d + 128 * c.half;          // my 128 fc_last columns = 32 mixture components
  int64_t step = 0;                                  // Mobius executions finished by this tile (same on both tiles)
  int mob_cur = 0;                                   // step % n_mob, maintained without a 64-bit modulo
  // Ping-pong of the two tiles: the MLP chain of one tile (latency bound: four dependent GEMM round trips) is made to run
  // against the mixture arithmetic of the other (issue bound).  Named barrier 5 + t = "tile t may start its chain".
  const int turn_mine = 5 + c.tile, turn_other = 6 - c.tile;
  if (RNF_TC_TURN_RELEASE >= 0 && c.tile == 1) named_arrive(5, 512);

  for (int64_t item = 0; item < my_items; ++item) {
    const int64_t tile_idx = 2 * (blockIdx.x + item * (int64_t)gridDim.x) + c.tile;
    int64_t row = 0, img = 0, g = 0;
    bool valid = false;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (tile_idx < a.n_tiles) {
      if (GRID) {
        img = tile_idx / a.tiles_per_image;
        g = (tile_idx % a.tiles_per_image) * kRows + c.row;
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
        row = tile_idx * kRows + c.row;
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

#pragma unroll 1
    for (int lstep = 0; lstep < a.n_layers; ++lstep) {
      const int li = INV ? a.n_layers - 1 - lstep : lstep;
      const LayerDev L = a.layers[li];
      if (L.kind != RNF_LAYER_MOBIUS) {
        const float* W = L.cond_slot >= 0 ? cond_img + (int64_t)a.n_mobius_slots * kH + (int64_t)L.cond_slot * kAffFloats
                                          : a.weights + L.w_off;
        if (INV) W += kAffInv;
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
      const float* cimg = (L.cond_slot >= 0 && cond_img != nullptr) ? cond_img + (int64_t)L.cond_slot * kH : nullptr;

#if RNF_TC_TRACE
      const bool tr_on = a.trace != nullptr && blockIdx.x == 0 && c.elected && step >= 40 && step < 48;
      long long* tr = a.trace + ((c.tile * 8 + (step - 40)) * 32);
#endif
      TRACE(0);
      if (RNF_TC_TURN_RELEASE >= 0) named_bar(turn_mine, 512);
      TRACE(1);
      // fp32 side data of this layer (first-layer columns, biases): double buffered on the layer parity, every thread
      // observes the bulk copy itself
      const int abuf = (int)(step & 1);
      const int mob_n1 = mob_cur + 1 >= n_mob ? mob_cur + 1 - n_mob : mob_cur + 1;     // (step + 1) % n_mob
      const int mob_n2 = mob_n1 + 1 >= n_mob ? mob_n1 + 1 - n_mob : mob_n1 + 1;         // (step + 2) % n_mob
      mbar_wait(c.bars + 8 * (BAR_AUX_FULL + abuf), (uint32_t)((step >> 1) & 1));
      if (RNF_TC_YIELD && c.elected) s_busy[c.tile] = 1;
      const uint8_t* aux = smem + kOffAux + abuf * kAuxStride;
      const float4* sFirst = reinterpret_cast<const float4*>(aux);

      // ---- first conditioner layer for my 32 columns; pre-activation stashed in TMEM for the residual ----
      {
        float h0[32], act[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4 cf = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cimg != nullptr) cf = __ldg(reinterpret_cast<const float4*>(cimg + 32 * c.half) + j4);
          const float cc[4] = {cf.x, cf.y, cf.z, cf.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 f = sFirst[32 * c.half + 4 * j4 + e];
            const float hv = fmaf(f.z, y[2], fmaf(f.y, y[1], fmaf(f.x, y[0], f.w))) + cc[e];
            h0[4 * j4 + e] = hv;
            act[4 * j4 + e] = hv;                    // ReLU happens inside the fp16 split
          }
        }
        store_a_operand(a_hi, a_lo, c.row, c.half, act);
        tmem_st32(tm_stash, h0);
      }
      TRACE(2);
      // ---- three hidden layers on the tensor core ----
#pragma unroll 1
      for (int l = 0; l < 3; ++l) {
        fence_proxy_async();
        tc_fence_before();
        named_bar(bar_tile, 256);
        TRACE(3 + 4 * l);
        if (RNF_TC_TURN_RELEASE == l) named_arrive(turn_other, 512);
        if (issuer_warp) {
          mbar_wait(c.bars + 8 * (BAR_W_FULL + l), (c.par_w >> l) & 1u);
          tc_fence_after();
          if (elect_one_sync()) {
            const uint32_t wb = w_hid_d + l * (kW1Bytes >> 4);
            issue_split_gemm(tmem_base + c.tile * 256, a_hi_d, a_lo_d, wb, wb + (8192 >> 4), ones_d, bias_hid_d + l * (kW1Bytes >> 4),
                             umma_idesc(128, 64));
            umma_commit(bar_mma0);
            if (RNF_TC_YIELD) s_busy[c.tile] = 0;
          }
          __syncwarp();
        }
        TRACE(4 + 4 * l);
        mbar_wait(bar_mma0, c.par_mma0);
        c.par_mma0 ^= 1;
        tc_fence_after();
        TRACE(5 + 4 * l);
        if (RNF_TC_YIELD && c.elected) s_busy[c.tile] = 1;
        // W_l is dead once BOTH tiles' GEMM l has completed: the second tile to get here refills it for the next layer
        if (c.elected && (atomicAdd(&s_cnt[l], 1) & 1) && step + 1 < total_steps) load_piece(mob_n1, l, 0);
        float acc[32];
        tmem_ld32(c.tmem_d + 32 * c.half, acc);
        if (l == 2) {                                // relu_last(x0 + x)   (flow/condition.py:29); biases come out of the GEMM
          float h0[32];
          tmem_ld32(tm_stash, h0);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += h0[j];
        }
        store_a_operand(a_hi, a_lo, c.row, c.half, acc);
        TRACE(6 + 4 * l);
      }
      // ---- fc_last: two N = 128 chunks, own barrier each ----
      fence_proxy_async();
      tc_fence_before();
      named_bar(bar_tile, 256);
      TRACE(15);
      if (RNF_TC_TURN_RELEASE == 3) named_arrive(turn_other, 512);
      if (issuer_warp) {
        mbar_wait(c.bars + 8 * (BAR_W_FULL + 3), (c.par_w >> 3) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t d = tmem_base + c.tile * 256;
          issue_split_gemm(d, a_hi_d, a_lo_d, w_last_d, w_last_d + (32768 >> 4), ones_d, bias_last_d, umma_idesc(128, 128));
          umma_commit(bar_mma0);
          issue_split_gemm(d + 128, a_hi_d, a_lo_d, w_last_d + (16384 >> 4), w_last_d + ((32768 + 16384) >> 4), ones_d,
                           bias_last_d + (4096 >> 4), umma_idesc(128, 128));
          umma_commit(bar_mma1);
          if (RNF_TC_YIELD) s_busy[c.tile] = 0;
        }
        __syncwarp();
        c.par_w ^= 0xFu;                             // every lane of the issuing warp keeps the weight-phase parities
      }
      TRACE(16);
      if (c.half == 0) { mbar_wait(bar_mma0, c.par_mma0); } else { mbar_wait(bar_mma1, c.par_mma1); }
      c.par_mma0 ^= 1;                               // both barriers complete exactly once here
      c.par_mma1 ^= 1;
      tc_fence_after();

      TRACE(17);
      // ---- mixture of my 32 components, 8 at a time straight from TMEM ----
      float S_sp = 0.0f, S_th = 0.0f, S_f = 0.0f;
      const float zr = dot3(x, P.r), zv = dot3(x, P.v);   // in-plane coordinates of the moving column
      {
        // 8 chunks of 16 columns (4 components each); the TMEM load of the next chunk is in flight while the current one is
        // evaluated (two register buffers).  Rolled into 4 iterations of 2 chunks: the fully unrolled body (32 KB of SASS)
        // overflowed the instruction cache (ncu: stall_no_instruction 0.53 per issue).
        float buf0[16], buf1[16];
        auto eval4 = [&](float* acc, int col) {      // col: first of the 16 fc_last columns (of my half) in `acc`
          if (RNF_TC_YIELD) {
            for (int nap = 0; nap < 64 && s_busy[1 - c.tile]; ++nap) __nanosleep(RNF_TC_NAP_NS);
          }
#if RNF_TC_MIX_YIELD >= 0
          __nanosleep(RNF_TC_MIX_YIELD);               // scheduler hint: let the other tile's chain warps in
#endif
          mixture4<!INV>(P, zr, zv, acc, S_sp, S_th, S_f);
          if (INV) tmem_st16(tm_mine + col, acc);      // prepared parameters stay in my TMEM lane for the bisection
        };
        tmem_ld16_async(tm_mine, buf0);
        tmem_ld_wait16(buf0);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          tmem_ld16_async(tm_mine + 32 * j + 16, buf1);
          eval4(buf0, 32 * j);
          tmem_ld_wait16(buf1);
          if (j < 3) tmem_ld16_async(tm_mine + 32 * j + 32, buf0);
          eval4(buf1, 32 * j + 16);
          if (j < 3) tmem_ld_wait16(buf0);
        }
      }
      TRACE(18);
      // W4 is dead once both chunks of BOTH tiles have completed (the elected thread sits in half 0: check chunk B too)
      if (c.elected) {
        mbar_wait(bar_mma1, c.par_mma1 ^ 1u);
        if ((atomicAdd(&s_cnt[3], 1) & 1) && step + 1 < total_steps) load_piece(mob_n1, 3, 0);
      }
      // ---- exchange partial sums between the two halves of the row (fixed summation order) ----
      x_mine[0] = S_sp; x_mine[128] = S_th; x_mine[256] = S_f;
      tc_fence_before();
      named_bar(bar_tile, 256);
      // Past this barrier every thread of the tile is done with this layer's aux buffer: when both tiles are, it is
      // refilled with the side data of the layer after next.
      TRACE(19);
      if (c.elected && (atomicAdd(&s_cnt[4], 1) & 1) && step + 2 < total_steps) load_piece(mob_n2, 4, abuf);
      S_sp = x_lo[0] + x_hi[0];
      float nx[3], nz[3];
      if (!INV) {
        S_th = x_lo[128] + x_hi[128];
        S_f = x_lo[256] + x_hi[256];
        const float inv_sp = rcp_nr(S_sp);
        circle_point(P.r, P.v, S_th * inv_sp, nx);
        ldj += logf(S_f * inv_sp);
      } else {
        // target angle of the given column in its own frame (flow/mobiusflow.py:157-167); ~pi by construction
        float ys = atan2f(zv, zr);
        ys = ys >= 0.0f ? ys : ys + kTwoPi;
        if (fabsf(ys - kTwoPi) < 1e-4f) ys = 0.0f;
        // BinFind.forward (flow/mobiusflow.py:196-224): bracket [pi/2, 3pi/2], 15 halvings, return the last probe.
        // Each half sums its 32 components; partial sums ping-pong through two exchange slots (one barrier per probe).
        float lo = kPi / 2.0f, hi = 1.5f * kPi, x0 = 0.0f;
#pragma unroll 1
        for (int it = 0; it < 15; ++it) {
          x0 = (lo + hi) / 2.0f;
          float sn, cs;
          sincosf(x0, &sn, &cs);
          float Fs = 0.0f;
          {
            float buf0[16], buf1[16];
            tmem_ld16_async(tm_mine, buf0);
            tmem_ld_wait16(buf0);
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
              tmem_ld16_async(tm_mine + 32 * j + 16, buf1);
              probe4(cs, sn, buf0, Fs);
              tmem_ld_wait16(buf1);
              if (j < 3) tmem_ld16_async(tm_mine + 32 * j + 32, buf0);
              probe4(cs, sn, buf1, Fs);
              if (j < 3) tmem_ld_wait16(buf0);
            }
          }
          const int slot = (1 + (it & 1)) * 128;     // slots 1 / 2 (slot 0 still holds S_sp of slow readers)
          x_mine[slot] = Fs;
          named_bar(bar_tile, 256);
          const float fx0 = (x_lo[slot] + x_hi[slot]) / S_sp - ys;
          const float half_w = (hi - lo) / 2.0f;
          if (fx0 < 0.0f) lo = lo + half_w;
          else if (fx0 >= 0.0f) hi = hi - half_w;
        }
        float sn, cs;
        sincosf(x0, &sn, &cs);
        nx[0] = fmaf(P.v[0], sn, P.r[0] * cs);
        nx[1] = fmaf(P.v[1], sn, P.r[1] * cs);
        nx[2] = fmaf(P.v[2], sn, P.r[2] * cs);
        float Sf = 0.0f;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          float prm[32];
          tmem_ld32(tm_mine + 32 * q, prm);
#pragma unroll
          for (int k = 0; k < 8; ++k) Sf = fmaf(prm[4 * k + 3], comp_f2(cs, sn, prm[4 * k], prm[4 * k + 1], prm[4 * k + 2]), Sf);
        }
        x_mine[0] = Sf;                              // slot 0: everybody read S_sp before the first probe barrier
        tc_fence_before();
        named_bar(bar_tile, 256);
        ldj -= logf((x_lo[0] + x_hi[0]) / S_sp);
      }
      cross3(nx, y, nz);
      normalize3_fast(nz);
      set_col(R, p0, nx);
      set_col(R, p2, nz);
      TRACE(20);
      ++step;
      mob_cur = mob_n1;
    }

    // ================================ outputs ================================
    if (!GRID) {
      if (valid && c.half == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) a.R_out[row * 9 + i] = R[i];
        a.ldj_out[row] = ldj;
      }
    } else if (tile_idx < a.n_tiles) {
      // the per-tile reduction runs on half 0 (4 warps = 128 rows); named barrier id 3 + tile
      if (c.half == 0) {
        float lp = ldj;
        if (a.fisher_A != nullptr) {
          float tr = 0.0f;
#pragma unroll
          for (int i = 0; i < 9; ++i) tr = fmaf(__ldg(a.fisher_A + img * 9 + i), R[i], tr);
          lp += tr - __ldg(a.fisher_c + img);
        }
        if (!valid) lp = -INFINITY;
        if (a.logp_out != nullptr && valid) a.logp_out[row] = lp;
        float* s_v = reinterpret_cast<float*>(smem + kOffRed + c.tile * 128);
        long long* s_i = reinterpret_cast<long long*>(smem + kOffRed + c.tile * 128 + 32);
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
        named_bar(3 + c.tile, 128);
        bv = s_v[0]; bi = s_i[0];
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float ov = s_v[w];
          const long long oi = s_i[w];
          if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        const float m = bv;
        float e = (valid && m > -INFINITY) ? expf(lp - m) : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        named_bar(3 + c.tile, 128);                  // everybody has read s_v / s_i
        if (lane == 0) s_v[w4] = e;
        named_bar(3 + c.tile, 128);
        if (c.row == 0) {
          const float s = (s_v[0] + s_v[1]) + (s_v[2] + s_v[3]);
          float* p = a.part + tile_idx * 4;
          p[0] = m;
          p[1] = s;
          p[2] = __int_as_float((int)(bi & 0xffffffffLL));
          p[3] = __int_as_float((int)(bi >> 32));
        }
        named_bar(3 + c.tile, 128);                  // scratch reusable by the next item
      }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace

bool flow_tc_supported(const rnf_flow* f) {
  for (int i = 0; i < f->model.n_layers; ++i)
    if (f->layers_host[i].kind == RNF_LAYER_MOBIUS && f->layers_host[i].w_off_tc < 0) return false;
  return true;
}

cudaError_t launch_flow_tc(const FlowArgs& a, bool inverse, int sm_count, cudaStream_t st) {
  const bool grid_mode = a.G > 0;
  void (*kern)(const FlowArgs) = grid_mode ? flow_tc_kernel<false, true>
                                           : (inverse ? flow_tc_kernel<true, false> : flow_tc_kernel<false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAlloc);
  if (e != cudaSuccess) return e;
  if (a.n_tiles <= 0) return cudaSuccess;
  const int64_t pairs = (a.n_tiles + 1) / 2;
  const int64_t grid = pairs < sm_count ? pairs : sm_count;
  kern<<<(unsigned)grid, kThreads, kSmemAlloc, st>>>(a);
  return cudaGetLastError();
}

}  // namespace rnf
