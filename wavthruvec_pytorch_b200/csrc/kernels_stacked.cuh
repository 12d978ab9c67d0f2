// Fused multi-receptive-field stage for C <= 64 (ResBlock2), N-stacked:  same algebra as
// kernels_fused.cuh,
//
//     sum_j out_j = [ sum_j conv2_j(A1_j) + conv_sum(A0) ] + sum_j (b1_j + b2_j) + nk * X0
//     A1_j = half(lrelu(conv1_j(A0) + b1_j + X0))                  (vec2wav/models.py:65-70,135-141)
//
// but the tensor-core work is laid out for what actually bounds it at small C.  One
// tcgen05.mma (M = 128, K = 16) streams its 4 KB A operand from shared memory at ~64 B/clk, i.e.
// ~64 clk per MMA however small N is: at N = C = 16..64 the tensor pipe idles behind operand
// delivery (measured: the first fused kernel ran C = 32 at 96 % of that 64-clk bound).  Every conv1
// of the stage, and conv_sum, read the SAME input tile at the same row shifts, so the taps that share
// a time offset are issued as ONE MMA whose B operand stacks their weight matrices along N
// (N up to 4C): 11 entries instead of 21 + 11 for kernel sizes 3/7/11.  That needs
//   * one row alignment for every accumulator: window row i <-> time q0c - hc + i for T1 AND T2
//     (conv2 then reads A1 rows i + off, off in [-hc, hc]; rows it cannot see are the discarded
//     halo rows i < hc, i >= hc + rc of the window);
//   * TMEM columns [T2 | T1_{nb-1} | ... | T1_0] per 128-row sub-tile with branch slots sorted by
//     kernel size, so the blocks sharing an offset are a contiguous column range starting at 0.
//
// The MMA warp interprets a host-built program (model.cpp plan_stacked, include/v2w.h): segments of
// entries with wait / commit bit masks, phase-A segments (read the input tile) first, then the
// phase-B segments (conv2, read A1).  C <= 32 has two accumulator sets (tile k+1's conv1 never
// waits for tile k's final epilogue), C = 64 fills TMEM with one.  (Issuing conv1 of tile k+1
// AHEAD of conv2 of tile k was measured 20-25 % slower: with one A1 buffer per branch the
// epilogue of tile k+1 then waits for conv2_j(k) and the two sides ping-pong.)
//
// The epilogue is latency-, not throughput-bound: what matters is that nothing it waits for is a
// cold global load.  For C <= 32 a thread keeps its fp32 X0 row in registers for all passes of a
// tile and fetches the NEXT tile's row one tile ahead.
//
// Roles: warp 0 weight producer (resident block once for C <= 32, ring for C = 64), warp 2 input
// tile producer, warp 1 MMA issuer, warps 4-11 epilogue (thread = window row).
#pragma once

#include "kernels_fused.cuh"
#include "kernels_simt.cuh"

// Developer timeline (build with V2W_NVCC_EXTRA=-DV2W_TRACE): CTA 0 prints SM-clock stamps of its MMA
// warp and of epilogue warp 4 for the first tiles.  Compiled out otherwise.
#ifdef V2W_TRACE
#define V2W_TR_DECL long long tr_t[96]; int tr_tag[96]; int tr_n = 0; const long long tr_base = clock64();
#ifndef V2W_TR_K0
#define V2W_TR_K0 30
#endif
#define V2W_TR(tag) do { if (blockIdx.x == 0 && tr_n < 96 && (tag) >= V2W_TR_K0 * 100 && (tag) < (V2W_TR_K0 + 2) * 100) { tr_tag[tr_n] = (tag); tr_t[tr_n++] = clock64() - tr_base; } } while (0)
#define V2W_TR_DUMP(name) do { if (blockIdx.x == 0 && lane == 0) for (int q = 0; q < tr_n; ++q) \
    printf("TR %s %d %d %lld\n", name, q, tr_tag[q], tr_t[q]); } while (0)
#else
#define V2W_TR_DECL
#define V2W_TR(tag) do { } while (0)
#define V2W_TR_DUMP(name) do { } while (0)
#endif

namespace v2w {
namespace stacked {

using namespace v2w::tc;
using fused::kBoxRows;
using fused::kWin;
using fused::st_shared_v4;

constexpr int kMaxSeg = V2W_FUSED_MAX_SEG;
constexpr int kMaxEnt = V2W_FUSED_MAX_TAPS;
constexpr int kBitT2 = V2W_FUSED_BIT_T2, kBitA0 = V2W_FUSED_BIT_A0;

// 32-bit fields only (profiles/r01_ptxas_param_miscompile.md)
struct Segment {
  int phase;        // 0: reads the input tile (conv1 / sum), 1: reads an A1 slot (conv2)
  int src;          // phase 1: A1 slot
  int dcol, n;      // accumulator columns inside a sub-tile, MMA N
  int count, first; // entries
  int acc;          // 0: the first MMA overwrites the accumulator
  int wait, commit; // V2W_FUSED_BIT_*
  uint32_t idesc;   // tcgen05 instruction descriptor (M = 128, N = n, operand format)
};

struct Params {
  const int* row_limit;         // variable-length batches: see tc::Params
  int rows_per_frame;
  int batch, rows, tiles_per_b;
  int rc, hc, h1;
  int a0_boxes;
  int nseg, nbranch, nent;
  int t_bufs, a0_bufs;
  int mma_warps;                // 1, or 2 (alternate tiles; needs t_bufs == a0_bufs == 2, resident weights)
  int sub_cols, tile_cols, tmem_cols;
  int stages_w, stage_w_bytes, w_region_bytes;
  int w_blocks;                 // resident: C-row blocks of the weight matrix
  int fp16;
  Segment seg[kMaxSeg];
  int a_off16[kMaxEnt];         // operand row shift of the entry in 16-byte units (negative for conv2 taps < 0)
  int w_off16[kMaxEnt];         // resident weights: offset of the entry's block in 16-byte units
  int w_row[kMaxEnt];           // first row of the entry's weight block (TMA coordinate)
  const float* x0;              // fp32 [B, L, C], or row-blocked (x0_blocked): [B][ceil(L / 32)][C / 4][32 rows][4]
  // Row-blocked X0: a warp's row-per-thread float4 read of one column quad is 512 contiguous bytes (4 L1
  // wavefronts) instead of 32 rows 4 C bytes apart (32 wavefronts) -- the epilogue passes of the fused stages
  // were bound by exactly those wavefronts (profiles/r02_notes.md, timeline of the C = 128 stage).
  int x0_blocked;
  // The same layout for the fp32 OUTPUT and for the running MRF sum (ResBlock1 conv pairs: the fp32 stream between
  // the pairs of a block, and the sum between blocks, never leave it): the final pass's row-per-thread float4
  // stores / loads become 512-byte runs instead of one 16-byte piece of 32 different lines each.
  int out_blocked, acc_blocked;
  const float* bias1[V2W_MAX_KERNELS];   // by slot
  const float* bsum;
  const float* accum_in;        // fp32 [B, L, C] added in the final pass (running MRF sum), or nullptr
  // conv_post folded into the final pass of the LAST stage (C <= 32): leaky_relu(0.01) -> Conv1d(C -> 1,
  // k = 7) -> tanh (models.py:143-145).  The stage's halo is 3 rows wider so that every central row
  // finds its 7 neighbours in the same window; the waveform is the kernel's only output.
  int post;
  float post_w[7 * 32];         // [tap][channel], stride C
  float post_bias;
  simt::PostOut post_out;
  simt::RaggedOffsets post_ro;
  int e1_add_x0;                // 1: the middle activation sees the residual (ResBlock2); 0: ResBlock1 pair
  // C = 64: the epilogue PRE-LOADS T1_j with X0 + b1_j (tcgen05.st) while conv2 of the previous tile
  // runs, conv1 accumulates on top, and the per-slot pass is load-activate-store only.  The
  // "T1 of slot 0 free" barrier then means "T1 of this tile pre-loaded" (waited with the full parity).
  int preload;
  float mid_slope, res_scale, out_scale, out_slope;
  float* out_f32;
  __nv_bfloat16* out_bf16;
};

// C = 16 is latency-bound (four serial epilogue passes per tile, 64 small MMAs) and needs only 73 KB
// of shared memory and 256 TMEM columns: two CTAs share an SM and hide each other's waits.
// kPair (C = 64): CTA pairs in lockstep, each with its own window and buffers, sharing the weight
// operand through cta_group::2 MMAs -- the same construction as mrf_stage_kernel<128, true>.
// kBlk: the instantiation whose final pass can write its fp32 output / read the running sum row-blocked
// (Params::out_blocked, acc_blocked: ResBlock1 conv pairs).  A template parameter because the extra address
// arithmetic, compiled into the ResBlock2 stage kernels as run-time selects, cost them 1-7 % (stage 3: 108 -> 115 us).
// kX0B: stage program whose residual stream is row-blocked (compile-time there, run-time in the kBlk instantiations)
template <int C, bool kPair = false, bool kBlk = false, bool kX0B = false>
__global__ void __launch_bounds__(kThreads, (C == 16 && !kPair) ? 2 : 1)
mrf_stacked_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                   const __grid_constant__ Params p) {
  // kBlk = false is THE ResBlock2 stage program of this width: pre-loaded accumulators iff C = 64, a folded conv_post
  // only at C <= 32, no running sum, residual in the middle activation -- compile-time facts there.  (As run-time
  // tests in these loops `preload` alone cost the C = 32 / 16 stages 7 / 15 %: it guards global loads.)  kBlk = true
  // keeps every test at run time: conv-pair programs, and a C = 64 stage program without pre-load ($V2W_FUSED_NO_PRELOAD).
  const bool preload = kBlk ? p.preload != 0 : (C == 64);
  const bool post = (!kBlk && C <= 32) ? p.post != 0 : false;

  constexpr int kRowBytes = C * 2;              // one K-chunk: the whole channel dimension
  constexpr int kKSteps = C / 16;
  constexpr bool kResident = C <= 32;
  constexpr bool kCacheX0 = C <= 32;
  constexpr uint32_t kA1Bytes = kWin * kRowBytes;
  constexpr uint32_t kBoxBytes = kBoxRows * kRowBytes;
  constexpr uint32_t kWBlockBytes = C * kRowBytes;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a0_bytes = p.a0_boxes * kBoxBytes;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_a1 = smem_a0 + p.a0_bufs * a0_bytes;
  const uint32_t smem_w = smem_a1 + p.nbranch * kA1Bytes;
  const uint32_t bars = smem_w + p.w_region_bytes;
  const uint32_t a0_full = bars, a0_empty = bars + 16, w_full = bars + 32, w_empty = bars + 96;
  const uint32_t t1_full = bars + 160, t1_empty = bars + 224;     // [t_buf][slot]
  const uint32_t t2_full = bars + 288, t2_empty = bars + 304;     // [t_buf]
  const uint32_t a1_full = bars + 320;                            // [t_buf][slot]: one waiter each
  const uint32_t w_all = bars + 416, tmem_slot = bars + 424;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.batch * p.tiles_per_b;
  static_assert(!kPair || C == 64, "CTA pairs are wired for the streamed-weight configuration");
  // Tile slots: a slot is one tile, or with CTA pairs the two tiles 2 s (leader) and 2 s + 1 (peer;
  // possibly past the end: that CTA then computes on zero-filled loads and stores nothing)
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int cta = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nctas = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_slots = kPair ? (total_tiles + 1) / 2 : total_tiles;
  const int n_my = (num_slots - cta + nctas - 1) / nctas;
  auto tile_of = [&](int kt) {                  // this CTA's tile of its kt-th slot
    const int slot = cta + kt * nctas;
    return kPair ? 2 * slot + (int)rank : slot;
  };
  auto tile_dead = [&](int tile) {
    if (tile >= total_tiles) return true;
    const int b = tile / p.tiles_per_b;
    return tile_skipped(p.row_limit, p.rows_per_frame, b, (tile - b * p.tiles_per_b) * p.rc);
  };
  // Variable-length batches: first local tile index >= kt whose rows some utterance still needs.
  // Every role walks the same sequence  kt = live(0), live(kt + 1), ...  and numbers the tiles it
  // meets k = 0, 1, ... (buffer indices and barrier phases follow k, addresses follow kt).
  auto live = [&](int kt) {
    for (; kt < n_my; ++kt) {
      const int slot = cta + kt * nctas;
      const bool dead = kPair ? (tile_dead(2 * slot) && tile_dead(2 * slot + 1)) : tile_dead(slot);
      if (!dead) break;   // a pair skips a slot only when neither CTA needs it: lockstep
    }
    return kt;
  };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a0_full + 8 * i, 1);
      mbar_init(a0_empty + 8 * i, 1);
      mbar_init(t2_full + 8 * i, 1);
      mbar_init(t2_empty + 8 * i, kPair ? 16 : 8);   // epilogue warps (of both CTAs)
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(w_full + 8 * i, 1);
      mbar_init(w_empty + 8 * i, 1);
      mbar_init(t1_full + 8 * i, 1);
      mbar_init(t1_empty + 8 * i, kPair ? 16 : 8);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(a1_full + 8 * i, kPair ? 16 : 8);
      mbar_init(a1_full + 8 * (4 + i), kPair ? 16 : 8);
    }
    mbar_init(w_all, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, p.tmem_cols);
    else tmem_alloc(tmem_slot, p.tmem_cols);
  }
  // biases of every pass in shared memory ([slot][C], then bsum[C]): an LDS broadcast in the
  // epilogue instead of an L1 round trip on its critical path
  const uint32_t smem_bias = bars + 512;
  const uint32_t smem_part = smem_bias + (((p.nbranch + 1) * C * 4 + 15) & ~15);   // conv_post partials, 2 x [7][256] (post only)
  for (int t = threadIdx.x; t < (p.nbranch + 1) * C; t += kThreads) {
    const int j = t / C, ch = t - j * C;
    const float v = j < p.nbranch ? p.bias1[j][ch] : p.bsum[ch];
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(smem_bias + 4 * t), "f"(v) : "memory");
  }
  tcgen05_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers exist before anything remote touches them
  __syncthreads();
  // everything above touched only kernel parameters and the (constant) packed weights: from here
  // on the previous kernel's outputs are read and its inputs overwritten
  pdl_launch_dependents();
  pdl_wait();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // =================================== weight producer ====================================
    if (kResident) {
      if (elect_one()) {
        mbar_expect_tx(w_all, p.w_blocks * kWBlockBytes);
        for (int t = 0; t < p.w_blocks; ++t)
          tma_load_2d(smem_w + t * kWBlockBytes, &tmW, 0, t * C, w_all);
      }
      __syncwarp();
    } else {
      Ring rw;
      for (int kt = live(0); kt < n_my; kt = live(kt + 1)) {
        for (int s = 0; s < p.nseg; ++s) {
          const Segment sg = p.seg[s];
          for (int t = 0; t < sg.count; ++t) {
            mbar_wait(w_empty + 8 * rw.s, rw.ph ^ 1u);
            if (elect_one()) {
              const uint32_t bar = w_full + 8 * rw.s;
              const uint32_t dst = smem_w + rw.s * p.stage_w_bytes;
              const int row = p.w_row[sg.first + t];
              if constexpr (kPair) {
                // this CTA's half of the stacked block: rows [rank * n/2, (rank + 1) * n/2), C/2-row boxes
                if (rank == 0) mbar_expect_tx(bar, sg.n * kRowBytes);
                const int half = sg.n / 2;
                for (int q = 0; q < half / (C / 2); ++q)
                  tma_load_2d_pair(dst + q * (kWBlockBytes / 2), &tmW, 0, row + (int)rank * half + q * (C / 2), bar);
              } else {
                mbar_expect_tx(bar, sg.n * kRowBytes);
                for (int q = 0; q < sg.n / C; ++q)
                  tma_load_2d(dst + q * kWBlockBytes, &tmW, 0, row + q * C, bar);
              }
            }
            __syncwarp();
            rw.advance(p.stages_w);
          }
        }
      }
    }
  } else if (warp == 2) {
    // ================================= input tile producer ==================================
    for (int kt = live(0), k = 0; kt < n_my; kt = live(kt + 1), ++k) {
      const int tile = tile_of(kt);
      const int b = tile / p.tiles_per_b;            // == batch for a peer tile past the end: zero fill
      const int ti = tile - b * p.tiles_per_b;
      const int row0 = ti * p.rc - p.hc - p.h1;      // first row of the input tile
      const uint32_t buf = k % p.a0_bufs, ph = (k / p.a0_bufs) & 1u;
      mbar_wait(a0_empty + 8 * buf, ph ^ 1u);
      if (elect_one()) {
        const uint32_t bar = a0_full + 8 * buf;
        if constexpr (kPair) {
          if (rank == 0) mbar_expect_tx(bar, 2 * p.a0_boxes * kBoxBytes);
          for (int bx = 0; bx < p.a0_boxes; ++bx)
            tma_load_3d_pair(smem_a0 + buf * a0_bytes + bx * kBoxBytes, &tmA, 0, row0 + bx * kBoxRows, b, bar);
        } else {
          mbar_expect_tx(bar, p.a0_boxes * kBoxBytes);
          for (int bx = 0; bx < p.a0_boxes; ++bx)
            tma_load_3d(smem_a0 + buf * a0_bytes + bx * kBoxBytes, &tmA, 0, row0 + bx * kBoxRows, b, bar);
        }
      }
      __syncwarp();
    }
  } else if ((warp == 1 || (warp == 3 && p.mma_warps == 2)) && rank == 0) {
    // ====================================== MMA issuer ======================================
    // With two accumulator sets (C <= 32) warps 1 and 3 take alternate tiles: each thread's
    // tcgen05.commit covers exactly its own tile's MMAs, the sets and input buffers are indexed by
    // tile parity, and a1_full is per (set, slot) so that each barrier has ONE waiter that sees its
    // phases in order (a parity wait can only tell a phase from its neighbour: a warp waiting for
    // tile k+1's A1 on a barrier shared with tile k would pass on tile k-1's phase).
    Ring rw;
    V2W_TR_DECL
#ifdef V2W_TRACE
    unsigned long long tr_g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tr_g0));
#endif
    const uint32_t dhi = smem_desc_hi<C>();
    if (kResident) mbar_wait(w_all, 0);
    // This warp's instruction latency is what bounds the small-C stages (a tile is 64 MMAs at
    // C = 16), so everything that can be is precomputed: per-entry descriptor offsets and the
    // instruction descriptor come from the host, buffer indices are bit operations, and only
    // the barriers a segment names are touched.
    const uint32_t tmask = p.t_bufs - 1, amask = p.a0_bufs - 1;     // depths are 1 or 2
    const uint32_t w_lo = (smem_w & 0x3FFFFu) >> 4;
    constexpr uint32_t kSubLo = (kMTile * kRowBytes) >> 4;           // second 128-row sub-tile
    for (int kt = live(0), kk = 0; kt < n_my; kt = live(kt + 1), ++kk) {
      if (p.mma_warps == 2 && (kk & 1) != (warp == 3 ? 1 : 0)) continue;
      const uint32_t k = kk;          // ordinal of this tile among the CTA's live tiles
      const uint32_t tb = k & tmask, tph = (k >> tmask) & 1u;
      const uint32_t ab = k & amask, aph = (k >> amask) & 1u;
      const uint32_t a0_lo = ((smem_a0 + ab * a0_bytes) & 0x3FFFFu) >> 4;
      for (int s = 0; s < p.nseg; ++s) {
        const Segment sg = p.seg[s];
        for (uint32_t w = sg.wait; w; w &= w - 1) {
          const int bit = __ffs(w) - 1;
          if (bit < 4) mbar_wait(t1_empty + 8 * (tb * 4 + bit), preload ? tph : (tph ^ 1u));
          else if (bit == 4) mbar_wait(t2_empty + 8 * tb, tph ^ 1u);
          else if (bit == 5) mbar_wait(a0_full + 8 * ab, aph);
          else mbar_wait(a1_full + 8 * (tb * 4 + bit - 8), tph);
        }
        tcgen05_fence_after();
        V2W_TR(k * 100 + s * 2);
        // One elected lane issues the whole segment.
        if (elect_one()) {
          const uint32_t a_lo = sg.phase ? ((smem_a1 + sg.src * kA1Bytes) & 0x3FFFFu) >> 4 : a0_lo;
          const uint32_t d_tmem = tmem_base + tb * p.tile_cols + sg.dcol;
          const uint32_t d_tmem1 = d_tmem + p.sub_cols;
          const uint32_t idesc = sg.idesc;
          Ring r = rw;   // private copy; all lanes advance `rw` below
          for (int t = 0; t < sg.count; ++t) {
            const int e = sg.first + t;
            // 14-bit start-address fields; conv2 entries reach up to hc rows before their tile
            const uint32_t ad_lo = ((a_lo + p.a_off16[e]) & 0x3FFFu) | (1u << 16);
            const uint32_t ad1_lo = ((a_lo + p.a_off16[e] + kSubLo) & 0x3FFFu) | (1u << 16);
            uint32_t bd_lo;
            if (kResident) {
              bd_lo = (w_lo + p.w_off16[e]) | (1u << 16);
            } else {
              mbar_wait(w_full + 8 * r.s, r.ph);
              tcgen05_fence_after();
              bd_lo = (w_lo + ((r.s * p.stage_w_bytes) >> 4)) | (1u << 16);
            }
            const uint64_t hi = static_cast<uint64_t>(dhi) << 32;
            const uint32_t acc0 = sg.acc | (t != 0);
#pragma unroll
            for (int kk = 0; kk < kKSteps; ++kk)
              if constexpr (kPair) umma_bf16_pair(d_tmem, (hi | ad_lo) + 2 * kk, (hi | bd_lo) + 2 * kk, idesc, acc0 | (kk != 0));
              else umma_bf16(d_tmem, (hi | ad_lo) + 2 * kk, (hi | bd_lo) + 2 * kk, idesc, acc0 | (kk != 0));
#pragma unroll
            for (int kk = 0; kk < kKSteps; ++kk)
              if constexpr (kPair) umma_bf16_pair(d_tmem1, (hi | ad1_lo) + 2 * kk, (hi | bd_lo) + 2 * kk, idesc, acc0 | (kk != 0));
              else umma_bf16(d_tmem1, (hi | ad1_lo) + 2 * kk, (hi | bd_lo) + 2 * kk, idesc, acc0 | (kk != 0));
            if (!kResident) {
              if constexpr (kPair) umma_commit_pair(w_empty + 8 * r.s);
              else umma_commit(w_empty + 8 * r.s);
              r.advance(p.stages_w);
            }
          }
          for (uint32_t w = sg.commit; w; w &= w - 1) {
            const int bit = __ffs(w) - 1;
            const uint32_t bar = bit < 4 ? t1_full + 8 * (tb * 4 + bit) : (bit == 4 ? t2_full + 8 * tb : a0_empty + 8 * ab);
            if constexpr (kPair) umma_commit_pair(bar);
            else umma_commit(bar);
          }
        }
        __syncwarp();
        V2W_TR(k * 100 + s * 2 + 1);
        if (!kResident) {
          for (int t = 0; t < sg.count; ++t) rw.advance(p.stages_w);
        }
      }
    }
    if (warp == 1) V2W_TR_DUMP("mma");
#ifdef V2W_TRACE
    if (warp == 1 && lane == 0 && (blockIdx.x % 21) == 0) {
      unsigned long long g1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
      printf("TRC cta %d tiles %d clocks %lld ns %llu\n", (int)blockIdx.x, n_my, clock64() - tr_base, g1 - tr_g0);
    }
#endif
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int sub = (warp - 4) >> 2;              // which 128-row accumulator
    const int quarter = warp & 3;                 // TMEM lane quarter
    const int i = sub * kMTile + quarter * 32 + lane;   // window row of this thread
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + sub * p.sub_cols;
    const uint32_t swz = C == 64 ? (i & 7) : (C == 32 ? ((i >> 1) & 3) : ((i >> 2) & 1));
    constexpr int kNC = C / 16;                   // 16-column chunks per row
    constexpr int kXc = kCacheX0 ? C / 4 : 8;     // cached row, or a 2-chunk look-ahead window
    constexpr int kPair = kNC >= 2 ? 2 : 1;
    const bool central = i >= p.hc && i < p.hc + p.rc;
    // fp32 X0 row of tile `k` for this thread (clamped to a valid row; `valid` says whether it counts)
    // X0 row of this thread for tile `kt`: pointer to its first column quad; quad q is at ptr + q * xs.
    // `lin` = b * rows + row (linear row index: addresses of every row-major tensor of the stage).
    const bool x0_blocked = kBlk ? p.x0_blocked != 0 : kX0B;
    const int xs = x0_blocked ? 32 : 1;
    const int x_nblk = ((p.rows + 31) >> 5) + tc::kBlockedPad;
    size_t boff = 0;     // the row's first column quad in a row-blocked [B, L, C] tensor, in float4 units (quad q: + 32 q)
    auto row_of = [&](int kt, bool& valid, size_t& lin) -> const float4* {
      const int tile = tile_of(kt);
      const bool live_tile = tile < total_tiles;       // a peer tile past the end stores nothing
      const int b = live_tile ? tile / p.tiles_per_b : 0;
      const int ti = live_tile ? tile - b * p.tiles_per_b : 0;
      const int gr = ti * p.rc - p.hc + i;
      valid = live_tile && gr >= 0 && gr < p.rows;
      const int r = valid ? gr : 0;
      lin = static_cast<size_t>(b) * p.rows + r;
      if constexpr (kBlk) boff = (static_cast<size_t>(b) * x_nblk + (r >> 5)) * (C / 4) * 32 + (r & 31);
      if (x0_blocked)
        return reinterpret_cast<const float4*>(p.x0) +
               (static_cast<size_t>(b) * x_nblk + (r >> 5)) * (C / 4) * 32 + (r & 31);
      return reinterpret_cast<const float4*>(p.x0 + lin * C);
    };
    float4 xn[kCacheX0 ? kXc : 1];     // C <= 32: the next tile's row, fetched one tile ahead
    if constexpr (kCacheX0) {
      bool v;
      size_t l0;
      const float4* x = row_of(min(live(0), n_my - 1), v, l0);
#pragma unroll
      for (int e = 0; e < kXc; ++e) xn[e] = __ldg(x + e * xs);
    }
    V2W_TR_DECL
    float4 xc[kXc];
    if constexpr (!kCacheX0) {         // C = 64: chunks 0 and 1 of the first pass
      bool v;
      size_t l0;
      const float4* x = row_of(min(live(0), n_my - 1), v, l0);
#pragma unroll
      for (int e = 0; e < 8; ++e) xc[e] = __ldg(x + e * xs);
    }
    // P(kt): T1_j := X0 + b1_j for every slot of the tile at local index kt (its accumulator set)
    auto preload_t1 = [&](int kt, uint32_t tb_p) {
      bool v;
      size_t l0;
      const float4* x = row_of(kt, v, l0);
      const uint32_t t1s = t_row + tb_p * p.tile_cols;
#pragma unroll
      for (int ci = 0; ci < kNC; ++ci) {
        float4 xr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) xr[e] = __ldg(x + (ci * 4 + e) * xs);
        for (int j = 0; j < p.nbranch; ++j) {
          uint32_t r[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 bb = ld_shared_f4(smem_bias + (j * C + ci * 16 + 4 * e) * 4);
            r[4 * e] = __float_as_uint(xr[e].x + bb.x);
            r[4 * e + 1] = __float_as_uint(xr[e].y + bb.y);
            r[4 * e + 2] = __float_as_uint(xr[e].z + bb.z);
            r[4 * e + 3] = __float_as_uint(xr[e].w + bb.w);
          }
          tmem_st16(t1s + (p.nbranch - j) * C + ci * 16, r);
        }
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kPair) mbar_arrive_leader(t1_empty + 8 * (tb_p * 4));
        else mbar_arrive(t1_empty + 8 * (tb_p * 4));
      }
    };
    if (preload && live(0) < n_my) preload_t1(live(0), 0);
    for (int kt = live(0), k = 0; kt < n_my; ++k) {
      bool valid, next_valid = false;
      size_t lin, lin_next;
      const float4* xrow = row_of(kt, valid, lin);
      const size_t cur_boff = boff;       // (the look-ahead call below overwrites it)
      const int kt_next = live(kt + 1);
      const float4* xnext = kt_next < n_my ? row_of(kt_next, next_valid, lin_next) : xrow;
      const uint32_t tb = k % p.t_bufs, tph = (k / p.t_bufs) & 1u;
      if constexpr (kCacheX0) {
#pragma unroll
        for (int e = 0; e < kXc; ++e) xc[e] = xn[e];
        if (kt_next < n_my) {
#pragma unroll
          for (int e = 0; e < kXc; ++e) xn[e] = __ldg(xnext + e * xs);
        }
      }
      // ---- per slot: T1 -> A1 = half(lrelu(conv1 + b1 + X0)) in shared memory ----
      for (int j = 0; j < p.nbranch; ++j) {
        V2W_TR(k * 100 + j * 10);
        mbar_wait(t1_full + 8 * (tb * 4 + j), tph);
        V2W_TR(k * 100 + j * 10 + 1);
        // A1_j is free: this thread has already seen t2_full of the previous tile (final pass
        // below), whose commit covers every conv2 MMA that read the A1 buffers.
        tcgen05_fence_after();
        V2W_TR(k * 100 + j * 10 + 2);
        const uint32_t a1_row = smem_a1 + j * kA1Bytes + i * kRowBytes;
        const uint32_t t1 = t_row + tb * p.tile_cols + (p.nbranch - j) * C;
        const uint32_t b1 = smem_bias + j * (C * 4);
#pragma unroll
        for (int cp = 0; cp < kNC; cp += kPair) {
          // two 16-column chunks per TMEM round trip
          uint32_t rr[kPair][16];
#pragma unroll
          for (int u = 0; u < kPair; ++u) tmem_ld16(t1 + (cp + u) * 16, rr[u]);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < kPair; ++u) {
            const int ci = cp + u;
            const int cc = ci * 16;
            const uint32_t (&r)[16] = rr[u];
            float4 xr[4];
            if constexpr (kCacheX0) {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[ci * 4 + e];
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[(ci & 1) * 4 + e];
              // C = 64: a two-chunk window that runs ahead of the consumer -- past the end of
              // this pass it wraps to chunks 0, 1 of the next pass (same row, L1/L2 hits)
              const int nc = ci + 2 < kNC ? ci + 2 : ci + 2 - kNC;
              if (!preload) {   // (pre-loaded accumulators: this pass does not read X0 at all)
#pragma unroll
                for (int e = 0; e < 4; ++e) xc[(ci & 1) * 4 + e] = __ldg(xrow + (nc * 4 + e) * xs);
              }
            }
            uint32_t h[8];
            if (valid) {
              float v[16];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 bb = ld_shared_f4(b1 + (cc + 4 * e) * 4);
                // selects, not multiplies by 0: in pre-load mode this pass does not load xr
                const bool ax = (kBlk ? p.e1_add_x0 != 0 : true) && !preload, ab = !preload;
                v[4 * e] = __uint_as_float(r[4 * e]) + (ab ? bb.x : 0.f) + (ax ? xr[e].x : 0.f);
                v[4 * e + 1] = __uint_as_float(r[4 * e + 1]) + (ab ? bb.y : 0.f) + (ax ? xr[e].y : 0.f);
                v[4 * e + 2] = __uint_as_float(r[4 * e + 2]) + (ab ? bb.z : 0.f) + (ax ? xr[e].z : 0.f);
                v[4 * e + 3] = __uint_as_float(r[4 * e + 3]) + (ab ? bb.w : 0.f) + (ax ? xr[e].w : 0.f);
              }
              tc::lrelu_pack16(v, p.mid_slope, p.fp16 != 0, h);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) h[e] = 0u;   // conv2's zero padding outside [0, L)
            }
            const uint32_t q = cc >> 3;            // 16-byte unit inside the row
            st_shared_v4(a1_row + ((q ^ swz) << 4), h[0], h[1], h[2], h[3]);
            st_shared_v4(a1_row + (((q + 1) ^ swz) << 4), h[4], h[5], h[6], h[7]);
          }
        }
        fence_proxy_async();          // generic-proxy writes -> visible to the tensor core
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {              // one arrival per warp: 256 arrivals serialise on the barrier
          if constexpr (kPair) {
            mbar_arrive_leader(a1_full + 8 * (tb * 4 + j));
            if (!preload) mbar_arrive_leader(t1_empty + 8 * (tb * 4 + j));
          } else {
            mbar_arrive(a1_full + 8 * (tb * 4 + j));
            if (!preload) mbar_arrive(t1_empty + 8 * (tb * 4 + j));
          }
        }
        V2W_TR(k * 100 + j * 10 + 3);
      }
      // every T1 of this tile has been consumed: pre-load the next tile's (single accumulator set),
      // in the time the tensor pipe spends on this tile's conv2 segments
      if (preload && kt_next < n_my) preload_t1(kt_next, 0);
      // ---- final pass with conv_post folded in (last stage, C <= 32) ----
      if constexpr (C <= 32) {
        if (post) {
          // rows conv_post needs from this window: the central rows and 3 on either side (all of them
          // have complete conv2 sums: the true halo is hc - 3)
          const bool need = valid && i >= p.hc - 3 && i < p.hc + p.rc + 3;
          mbar_wait(t2_full + 8 * tb, tph);
          tcgen05_fence_after();
          const uint32_t t2 = t_row + tb * p.tile_cols;
          float part[7];
#pragma unroll
          for (int j = 0; j < 7; ++j) part[j] = 0.f;
#pragma unroll
          for (int cp = 0; cp < kNC; cp += kPair) {
            uint32_t rr[kPair][16];
#pragma unroll
            for (int u = 0; u < kPair; ++u) tmem_ld16(t2 + (cp + u) * 16, rr[u]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < kPair; ++u) {
              const int ci = cp + u;
              const int cc = ci * 16;
              if (need) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 bb = ld_shared_f4(smem_bias + (p.nbranch * C + cc + 4 * e) * 4);
                  const float4 xr = xc[ci * 4 + e];
                  float a[4];
                  a[0] = lrelu((__uint_as_float(rr[u][4 * e]) + bb.x + p.res_scale * xr.x) * p.out_scale, p.out_slope);
                  a[1] = lrelu((__uint_as_float(rr[u][4 * e + 1]) + bb.y + p.res_scale * xr.y) * p.out_scale, p.out_slope);
                  a[2] = lrelu((__uint_as_float(rr[u][4 * e + 2]) + bb.z + p.res_scale * xr.z) * p.out_scale, p.out_slope);
                  a[3] = lrelu((__uint_as_float(rr[u][4 * e + 3]) + bb.w + p.res_scale * xr.w) * p.out_scale, p.out_slope);
#pragma unroll
                  for (int j = 0; j < 7; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) part[j] = fmaf(a[q], p.post_w[j * C + cc + 4 * e + q], part[j]);
                }
              }
            }
          }
          // T2 is drained: hand it back before the exchange through shared memory
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t2_empty + 8 * tb);
          const uint32_t pb = smem_part + (k & 1) * (7 * kWin * 4);
#pragma unroll
          for (int j = 0; j < 7; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(pb + (j * kWin + i) * 4), "f"(part[j]) : "memory");
          asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight epilogue warps
          if (valid && central) {
            // y[t] = bias + sum_j w[j] . x[t + j - 3]; row t + j - 3 is window row i + j - 3
            float sum = p.post_bias;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
              float v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pb + (j * kWin + i + j - 3) * 4));
              sum += v;
            }
            const float y = tanhf(sum);
            const size_t row_idx = lin;   // b * rows + t
            const int b = static_cast<int>(row_idx / p.rows);
            const int t = static_cast<int>(row_idx - static_cast<size_t>(b) * p.rows);
            if (!p.post_out.ragged) {
              p.post_out.y[row_idx] = y;
            } else {
              const int first = p.post_ro.off[b];
              if (t < (p.post_ro.off[b + 1] - first) * p.post_out.up) {
                const size_t o = static_cast<size_t>(first) * p.post_out.up + t;
                if (p.post_out.pcm16)
                  reinterpret_cast<short*>(p.post_out.y)[o] =
                      static_cast<short>(__float2int_rn(fminf(fmaxf(y * 32768.f, -32768.f), 32767.f)));
                else
                  p.post_out.y[o] = y;
              }
            }
          }
          kt = kt_next;
          continue;
        }
      }
      // ---- final: out = (T2 + bsum + nk * X0) / nk on the central rows of the window ----
      {
        const bool out_valid = valid && central;
        const size_t row_base = lin * C;
        const size_t out_boff = cur_boff;
        if constexpr (!kCacheX0) {
          if (preload) {   // the X0 window was not kept running by the per-slot passes: prime it here
#pragma unroll
            for (int e = 0; e < 8; ++e) xc[e] = __ldg(xrow + e * xs);
          }
        }
        V2W_TR(k * 100 + 50);
        mbar_wait(t2_full + 8 * tb, tph);
        tcgen05_fence_after();
        V2W_TR(k * 100 + 51);
        const uint32_t t2 = t_row + tb * p.tile_cols;
#pragma unroll
        for (int cp = 0; cp < kNC; cp += kPair) {
          uint32_t rr[kPair][16];
#pragma unroll
          for (int u = 0; u < kPair; ++u) tmem_ld16(t2 + (cp + u) * 16, rr[u]);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < kPair; ++u) {
            const int ci = cp + u;
            const int cc = ci * 16;
            const uint32_t (&r)[16] = rr[u];
            float4 xr[4];
            if constexpr (kCacheX0) {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[ci * 4 + e];
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[(ci & 1) * 4 + e];
              // past the end of the tile's last pass the window moves on to the next tile's row
              const float4* src = ci + 2 < kNC ? xrow + (ci + 2) * 4 * xs : xnext + (ci + 2 - kNC) * 4 * xs;
              if (!preload || ci + 2 < kNC) {
#pragma unroll
                for (int e = 0; e < 4; ++e) xc[(ci & 1) * 4 + e] = __ldg(src + e * xs);
              }
            }
            if (out_valid) {
              float v[16];
              float4 acc4[4];
              if (kBlk && p.accum_in) {      // (only the conv-pair programs carry a running sum)
                const float4* ap = reinterpret_cast<const float4*>(p.accum_in + row_base + cc);
                int as = 1;
                if constexpr (kBlk) {
                  if (p.acc_blocked) {
                    ap = reinterpret_cast<const float4*>(p.accum_in) + out_boff + ci * 4 * 32;
                    as = 32;
                  }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) acc4[e] = __ldg(ap + e * as);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) acc4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 bb = ld_shared_f4(smem_bias + (p.nbranch * C + cc + 4 * e) * 4);
                v[4 * e] = (__uint_as_float(r[4 * e]) + bb.x + p.res_scale * xr[e].x + acc4[e].x) * p.out_scale;
                v[4 * e + 1] = (__uint_as_float(r[4 * e + 1]) + bb.y + p.res_scale * xr[e].y + acc4[e].y) * p.out_scale;
                v[4 * e + 2] = (__uint_as_float(r[4 * e + 2]) + bb.z + p.res_scale * xr[e].z + acc4[e].z) * p.out_scale;
                v[4 * e + 3] = (__uint_as_float(r[4 * e + 3]) + bb.w + p.res_scale * xr[e].w + acc4[e].w) * p.out_scale;
              }
              if (kBlk && p.out_f32) {      // (a stage program with an fp32 output runs the kBlk instantiation)
                float4* op = reinterpret_cast<float4*>(p.out_f32 + row_base + cc);
                int os = 1;
                if constexpr (kBlk) {
                  if (p.out_blocked) {
                    op = reinterpret_cast<float4*>(p.out_f32) + out_boff + ci * 4 * 32;
                    os = 32;
                  }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  op[e * os] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
              }
              if (p.out_bf16) {
                uint32_t h[8];
                tc::lrelu_pack16(v, p.out_slope, p.fp16 != 0, h);
                uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + row_base + cc);
                op[0] = make_uint4(h[0], h[1], h[2], h[3]);
                op[1] = make_uint4(h[4], h[5], h[6], h[7]);
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_leader(t2_empty + 8 * tb);
          else mbar_arrive(t2_empty + 8 * tb);
        }
        V2W_TR(k * 100 + 52);
      }
      kt = kt_next;
    }
    if (warp == 4) V2W_TR_DUMP("epi");
  }

  tcgen05_fence_before();
  if constexpr (kPair) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, p.tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace stacked
}  // namespace v2w
