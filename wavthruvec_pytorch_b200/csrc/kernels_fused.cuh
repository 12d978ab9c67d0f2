// Fused multi-receptive-field stage (ResBlock2), plain program:  one kernel computes
//
//     x_out = (1/nk) * sum_j  ResBlock2_j(x)            (vec2wav/models.py:65-70,135-141)
//
// for a 226-row tile without the intermediates ever leaving the SM.  Unfused, one stage moves
// ~74 B per output element through HBM (six convs x [16-bit operand + fp32 residual in, fp32 +
// 16-bit out]); fused it is ~10 B (A0 + fp32 X0 in, 16-bit or fp32 out).
//
// In the shipped plans this kernel runs the C = 128 stage (two K-chunks of 64 channels; TMEM and
// shared memory are full there, everything is single buffered) and the C = 128 conv pairs of
// ResBlock1 (Params::pair).  C <= 64 runs kernels_stacked.cuh; the C <= 64 instantiations below are
// kept as the A/B baseline of that kernel (V2W_FUSED_NO_STACK=1).
//
// Algebra (per branch j: conv1_j dilation d0, conv2_j dilation d1_j; A0 = half(lrelu(X0))):
//     Y1_j  = conv1_j(A0) + b1_j + X0                       (fp32)
//     A1_j  = half(lrelu(Y1_j)), zero outside [0, L)        (stays in shared memory)
//     out_j = conv2_j(A1_j) + b2_j + Y1_j
//     sum_j out_j = [ sum_j conv2_j(A1_j) + conv_sum(A0) ] + sum_j (b1_j + b2_j) + nk * X0
// where conv_sum is ONE convolution whose taps are the per-offset sums of the conv1_j weights
// (built at pack time).  Everything in brackets accumulates in a single TMEM accumulator T2; the
// only per-branch epilogue is T1_j -> A1_j.
//
// Geometry: window = 256 rows (two 128-lane accumulators) = central rows [q0c, q0c + rc) plus hc rows
// of halo per side (hc = max_j d1_j (k_j - 1)/2 = 15, rc = 226).  conv1_j is evaluated on the whole
// window (its output feeds conv2's halo), conv_sum and conv2_j on the 256 rows starting at q0c (the
// last 2*hc of them are discarded).  A0 tile = window + h1 rows per side (h1 = d0 (kmax - 1)/2).
//
// Roles: warp 0 TMA producer (A0 tile, weight taps), warp 1 MMA issuer (one elected lane issues a
// whole segment), warps 4-11 epilogue (thread = window row).  The segment program comes from the
// host (model.cpp plan_fused_stage / plan_pair): for C = 128  c1_0, SUM, then c2_j, c1_{j+1}, so that
// conv_sum covers the first epilogue; for a ResBlock1 pair  c1, c2  (no SUM: conv2 is T2's first
// writer and waits for the previous tile's final pass).
//
// Weights stream per (tap, K-chunk) through a ring for C >= 64; the C <= 32 instantiations keep all
// taps resident.  The conv1 biases are staged in shared memory; barrier arrivals are one per warp.
#pragma once

#include "kernels_tc.cuh"

namespace v2w {
namespace fused {

using namespace v2w::tc;

constexpr int kWin = 256;          // window rows per tile
constexpr int kBoxRows = 64;
constexpr int kMaxSeg = 12;
constexpr int kMaxTaps = 80;
constexpr int kSegConv1 = 0, kSegSum = 1, kSegConv2 = 2;
constexpr int kSegFlagReleaseA0 = 1;   // last reader of the A0 tile
constexpr int kSegFlagLast = 2;        // last segment of the tile
constexpr int kSegFlagAcc = 4;         // accumulate from the first MMA (a later piece of the `sum` convolution)

// 32-bit fields on purpose: with four packed int16 ptxas (12.9) merged the constant-bank loads
// of a dynamically indexed p.seg[s] and then read `first` from an unrelated register whenever
// the array sat at a 4-byte (not 8-byte) aligned parameter offset -- wrong taps, no diagnostics
// (profiles/r01_ptxas_param_miscompile.md).
struct Segment {
  int type, branch, ntaps, first, flags;
};

struct Params {
  const int* row_limit;         // variable-length batches: see tc::Params
  int rows_per_frame;
  int batch, rows, tiles_per_b;
  int rc, hc, h1;               // central rows per tile, conv2 halo, conv1 halo
  int a0_boxes;                 // TMA boxes (64 rows) per A0 tile and K-chunk
  int nseg, nbranch, ntaps_total;
  int stages_w, stage_w_bytes;
  int w_region_bytes;           // bytes of the weight region (ring or resident block)
  int fp16;                     // 16-bit operand format: 0 = bf16, 1 = fp16
  Segment seg[kMaxSeg];
  int shift[kMaxTaps];          // operand row (sub-tile 0) of each tap inside its A tile
  const float* x0;              // fp32 [B, L, C], or row-blocked (x0_blocked): [B][ceil(L / 32)][C / 4][32 rows][4]
  int x0_blocked;               // see stacked::Params
  int out_blocked, acc_blocked; // fp32 output / running sum in the same row-blocked layout (stacked::Params)
  const float* bias1[V2W_MAX_KERNELS];
  const float* bsum;            // sum_j (b1_j + b2_j)   (ResBlock1 pair: b2)
  const float* accum_in;        // fp32 [B, L, C] added in the final pass (running MRF sum), or nullptr
  int pair;                     // 1: ResBlock1 conv pair -- no `sum` segment (the first conv2 MMA overwrites T2)
                                //    and the middle activation does not see the residual
  float mid_slope;              // leaky-relu between the convs of a block (0.1)
  float res_scale;              // nk: multiplier of X0 in the final sum
  float out_scale;              // 1/nk
  float out_slope;
  float* out_f32;
  __nv_bfloat16* out_bf16;
};

// kSub: 128-row accumulators per window.  2 = the 256-row window described above.  1 (C = 128 only, "narrow"):
// a 128-row window, so that T1 + T2 of TWO windows fit the 512 TMEM columns and every buffer is double
// buffered like at C <= 64 -- the MMAs of one window overlap the epilogue passes of the other, at the price of
// 98 instead of 226 central rows per window (15 % more MMA work per output row).  The eight epilogue warps
// then split a pass by COLUMN halves instead of row halves.
template <int C, int kSub = 2>
struct Cfg {
  static constexpr int kKC = C < 64 ? C : 64;            // channels per K-chunk
  static constexpr int kChunks = C / kKC;
  static constexpr int kRowBytes = kKC * 2;
  static constexpr int kKSteps = kKC / 16;
  static constexpr int kAccCols = C < 32 ? 32 : C;       // TMEM columns per 128-row accumulator
  static constexpr bool kResident = C <= 32;             // weights stay in shared memory
  static constexpr int kSlots = (C <= 64 || kSub == 1) ? 2 : 1;   // buffering depth of A0 / T1 / A1 / T2
  static constexpr int kWinRows = kSub * 128;            // window rows per tile
  static constexpr bool kCacheX0 = C <= 32;              // keep the thread's X0 row in registers
  static constexpr uint32_t kT1Cols = kSub * kAccCols;   // kSub sub-tiles
  static constexpr uint32_t kTmemCols = 2 * kSlots * kT1Cols;   // 512 / 512 / 256 / 256
  static constexpr uint32_t kA1ChunkBytes = kWinRows * kRowBytes;
  static constexpr uint32_t kA1Bytes = kChunks * kA1ChunkBytes;
  static constexpr uint32_t kBoxBytes = kBoxRows * kRowBytes;
  static constexpr uint32_t kWBytes = C * kRowBytes;     // one (tap, K-chunk) block
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                             uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}

// kPair (C = 128): two CTAs of a cluster walk the tile list in lockstep, each with its OWN window,
// input tile, A1 buffers and accumulators; what they share is the weight operand -- every MMA is one
// cta_group::2 instruction over both windows with half of the tap matrix in each CTA's shared memory
// (6 KB instead of 8 KB of operand reads per CTA and MMA at N = 128).  Barrier plumbing as in
// tapgemm_tc_kernel: loads land on the leader's full barriers, commits are multicast, the epilogue
// warps of both CTAs arrive on the leader's barriers.
// kBlk: see stacked::mrf_stacked_kernel (row-blocked fp32 output / running sum of the ResBlock1 conv pairs)
// kX0B: stage program whose residual stream is row-blocked (compile-time there, run-time in the kBlk instantiations)
template <int C, bool kPair = false, int kSub = 2, bool kBlk = false, bool kX0B = false>
__global__ void __launch_bounds__(kThreads, 1)
mrf_stage_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ Params p) {
  using K = Cfg<C, kSub>;
  static_assert(kSub == 2 || (kSub == 1 && C == 128), "narrow windows are wired for C = 128");
  constexpr int kRowBytes = K::kRowBytes;
  constexpr int kSlots = K::kSlots;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a0_chunk_bytes = p.a0_boxes * K::kBoxBytes;
  const uint32_t a0_bytes = K::kChunks * a0_chunk_bytes;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_a1 = smem_a0 + kSlots * a0_bytes;
  const uint32_t smem_w = smem_a1 + kSlots * K::kA1Bytes;
  const uint32_t bars = smem_w + p.w_region_bytes;
  // barriers: a0_full[2] a0_empty[2] w_full[8] w_empty[8] t1_full[2] t1_empty[2] a1_full[2]
  //           a1_empty[2] t2_full[2] t2_empty[2] w_all
  const uint32_t a0_full = bars, a0_empty = bars + 16, w_full = bars + 32, w_empty = bars + 96;
  const uint32_t t1_full = bars + 160, t1_empty = bars + 176, a1_full = bars + 192;
  const uint32_t a1_empty = bars + 208, t2_full = bars + 224, t2_empty = bars + 240;
  const uint32_t w_all = bars + 256, tmem_slot = bars + 264;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.batch * p.tiles_per_b;
  static_assert(!kPair || !K::kResident, "CTA pairs are wired for the streamed-weight configuration");
  // Tile slots: without pairs a slot is a tile; with pairs slot ts holds tiles 2 ts (leader) and
  // 2 ts + 1 (peer; possibly past the end -- that CTA then runs on zero-filled loads and stores nothing)
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int slot_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int slot_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int num_slots = kPair ? (total_tiles + 1) / 2 : total_tiles;
  auto tile_dead = [&](int tile) {   // past the end, or past its utterance's row limit
    if (tile >= total_tiles) return true;
    const int b = tile / p.tiles_per_b;
    return tile_skipped(p.row_limit, p.rows_per_frame, b, (tile - b * p.tiles_per_b) * p.rc);
  };
  auto slot_skipped = [&](int ts) {  // every role evaluates the same predicate: the pair stays in lockstep
    return kPair ? (tile_dead(2 * ts) && tile_dead(2 * ts + 1)) : tile_dead(ts);
  };

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    for (int i = 0; i < 2; ++i) {
      mbar_init(a0_full + 8 * i, 1);
      mbar_init(a0_empty + 8 * i, 1);
      mbar_init(t1_full + 8 * i, 1);
      mbar_init(t1_empty + 8 * i, kPair ? 16 : 8);   // epilogue warps (of both CTAs)
      mbar_init(a1_full + 8 * i, kPair ? 16 : 8);
      mbar_init(a1_empty + 8 * i, 1);
      mbar_init(t2_full + 8 * i, 1);
      mbar_init(t2_empty + 8 * i, kPair ? 16 : 8);
    }
    for (int i = 0; i < kMaxStagesW; ++i) {
      mbar_init(w_full + 8 * i, 1);
      mbar_init(w_empty + 8 * i, 1);
    }
    mbar_init(w_all, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, K::kTmemCols);
    else tmem_alloc(tmem_slot, K::kTmemCols);
  }
  // conv1 biases in shared memory ([branch][C]): an LDS broadcast in the per-branch epilogue
  // instead of an L1 round trip on its critical path (bsum stays in global memory: C = 128 needs
  // every remaining byte for the fifth weight stage)
  const uint32_t smem_bias = bars + 512;
  for (int t = threadIdx.x; t < p.nbranch * C; t += kThreads) {
    const int j = t / C, ch = t - j * C;
    const float v = p.bias1[j][ch];
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
  const uint32_t tmem_t2 = tmem_base + kSlots * K::kT1Cols;   // T2[slot] = tmem_t2 + slot * kT1Cols

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    Ring rw;
    uint32_t tile_iter = 0;
    if (K::kResident) {
      if (elect_one()) {
        mbar_expect_tx(w_all, p.ntaps_total * K::kWBytes);
        for (int t = 0; t < p.ntaps_total; ++t)
          tma_load_2d(smem_w + t * K::kWBytes, &tmW, 0, t * C, w_all);
      }
      __syncwarp();
    }
    for (int ts = slot_first; ts < num_slots; ts += slot_step) {
      if (slot_skipped(ts)) continue;
      const int tile = kPair ? 2 * ts + (int)rank : ts;
      const int b = tile / p.tiles_per_b;            // == batch for a peer tile past the end: zero fill
      const int ti = tile - b * p.tiles_per_b;
      const int row0 = ti * p.rc - p.hc - p.h1;      // first row of the A0 tile
      const uint32_t buf = tile_iter % kSlots, ph = (tile_iter / kSlots) & 1u;
      ++tile_iter;   // processed slots only (every role counts the same way)
      mbar_wait(a0_empty + 8 * buf, ph ^ 1u);
      if (elect_one()) {
        const uint32_t bar = a0_full + 8 * buf;
        const uint32_t bytes = K::kChunks * p.a0_boxes * K::kBoxBytes;
        if constexpr (kPair) {
          if (rank == 0) mbar_expect_tx(bar, 2 * bytes);   // both CTAs' tiles land on the leader's barrier
        } else {
          mbar_expect_tx(bar, bytes);
        }
        for (int c = 0; c < K::kChunks; ++c)
          for (int bx = 0; bx < p.a0_boxes; ++bx) {
            const uint32_t dst = smem_a0 + buf * a0_bytes + c * a0_chunk_bytes + bx * K::kBoxBytes;
            if constexpr (kPair) tma_load_3d_pair(dst, &tmA, c * K::kKC, row0 + bx * kBoxRows, b, bar);
            else tma_load_3d(dst, &tmA, c * K::kKC, row0 + bx * kBoxRows, b, bar);
          }
      }
      __syncwarp();
      if (!K::kResident) {
        for (int t = 0; t < p.ntaps_total; ++t) {
          for (int c = 0; c < K::kChunks; ++c) {
            mbar_wait(w_empty + 8 * rw.s, rw.ph ^ 1u);
            if (elect_one()) {
              const uint32_t bar = w_full + 8 * rw.s;
              if constexpr (kPair) {
                // this CTA's half of the tap matrix (C / 2 rows of n)
                if (rank == 0) mbar_expect_tx(bar, K::kWBytes);
                tma_load_2d_pair(smem_w + rw.s * p.stage_w_bytes, &tmW, c * K::kKC,
                                 t * C + (int)rank * (C / 2), bar);
              } else {
                mbar_expect_tx(bar, K::kWBytes);
                tma_load_2d(smem_w + rw.s * p.stage_w_bytes, &tmW, c * K::kKC, t * C, bar);
              }
            }
            __syncwarp();
            rw.advance(p.stages_w);
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ====================================== MMA issuer ======================================
    Ring rw;
    V2W_TRTC_DECL
    uint32_t tile_iter = 0, n1 = 0, n2 = 0;   // conv1 / conv2 segments issued so far
    const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kMTile : kMTile, C, p.fp16);
    const uint32_t dhi = smem_desc_hi<K::kKC>();
    if (K::kResident) mbar_wait(w_all, 0);
    for (int ts = slot_first; ts < num_slots; ts += slot_step) {
      if (slot_skipped(ts)) continue;
      const uint32_t buf = tile_iter % kSlots, tph = (tile_iter / kSlots) & 1u;
      ++tile_iter;
      V2W_TRTC(tile_iter - 1, 1);
      mbar_wait(a0_full + 8 * buf, tph);
      V2W_TRTC(tile_iter - 1, 2);
      const uint32_t a0 = smem_a0 + buf * a0_bytes;
      for (int s = 0; s < p.nseg; ++s) {
        const Segment sg = p.seg[s];
        uint32_t d_tmem, a_tile, a_chunk_bytes, slot = 0;
        if (sg.type == kSegConv1) {
          slot = n1 % kSlots;
          mbar_wait(t1_empty + 8 * slot, ((n1 / kSlots) & 1u) ^ 1u);
          d_tmem = tmem_base + slot * K::kT1Cols;
          a_tile = a0;
          a_chunk_bytes = a0_chunk_bytes;
        } else if (sg.type == kSegSum) {
          mbar_wait(t2_empty + 8 * buf, tph ^ 1u);
          d_tmem = tmem_t2 + buf * K::kT1Cols;
          a_tile = a0;
          a_chunk_bytes = a0_chunk_bytes;
        } else {
          slot = n2 % kSlots;
          if (kBlk && p.pair) mbar_wait(t2_empty + 8 * buf, tph ^ 1u);   // no `sum` segment: conv2 is T2's first writer
          mbar_wait(a1_full + 8 * slot, (n2 / kSlots) & 1u);
          d_tmem = tmem_t2 + buf * K::kT1Cols;
          a_tile = smem_a1 + slot * K::kA1Bytes;
          a_chunk_bytes = K::kA1ChunkBytes;
        }
        tcgen05_fence_after();
        V2W_TRTC(tile_iter - 1, 10 + s * 10);
        // One elected lane issues the whole segment: per tap only two descriptor words change
        // (A: row shift inside the tile, W: next tap matrix), everything else is loop-invariant.
        if (elect_one()) {
          const uint32_t a_lo = ((a_tile & 0x3FFFFu) >> 4) | (1u << 16);
          Ring r = rw;   // the elected lane walks a private copy; all lanes advance `rw` below
          for (int t = 0; t < sg.ntaps; ++t) {
            const uint32_t a_t = a_lo + p.shift[sg.first + t] * (kRowBytes >> 4);
#pragma unroll
            for (int c = 0; c < K::kChunks; ++c) {
              uint32_t w_addr;
              if (K::kResident) {
                w_addr = smem_w + (sg.first + t) * K::kWBytes;
              } else {
                mbar_wait(w_full + 8 * r.s, r.ph);
                tcgen05_fence_after();
                w_addr = smem_w + r.s * p.stage_w_bytes;
              }
              const uint64_t bd = smem_desc(dhi, w_addr);
              const uint32_t acc0 = (sg.type == kSegConv2 && !(kBlk && p.pair)) || (sg.flags & kSegFlagAcc) || (t != 0) || (c != 0);
#pragma unroll
              for (int m = 0; m < kSub; ++m) {
                const uint64_t ad = (static_cast<uint64_t>(dhi) << 32) |
                                    (a_t + ((c * a_chunk_bytes + m * (kMTile * kRowBytes)) >> 4));
#pragma unroll
                for (int kk = 0; kk < K::kKSteps; ++kk)
                  if constexpr (kPair)
                    umma_bf16_pair(d_tmem + m * K::kAccCols, ad + 2 * kk, bd + 2 * kk, idesc,
                                   acc0 | (kk != 0));
                  else
                    umma_bf16(d_tmem + m * K::kAccCols, ad + 2 * kk, bd + 2 * kk, idesc,
                              acc0 | (kk != 0));
              }
              if (!K::kResident) {
                if constexpr (kPair) umma_commit_pair(w_empty + 8 * r.s);
                else umma_commit(w_empty + 8 * r.s);
                r.advance(p.stages_w);
              }
            }
          }
          if constexpr (kPair) {
            if (sg.type == kSegConv1) umma_commit_pair(t1_full + 8 * slot);
            if (sg.type == kSegConv2) umma_commit_pair(a1_empty + 8 * slot);
            if (sg.flags & kSegFlagReleaseA0) umma_commit_pair(a0_empty + 8 * buf);
            if (sg.flags & kSegFlagLast) umma_commit_pair(t2_full + 8 * buf);
          } else {
            if (sg.type == kSegConv1) umma_commit(t1_full + 8 * slot);
            if (sg.type == kSegConv2) umma_commit(a1_empty + 8 * slot);
            if (sg.flags & kSegFlagReleaseA0) umma_commit(a0_empty + 8 * buf);
            if (sg.flags & kSegFlagLast) umma_commit(t2_full + 8 * buf);
          }
        }
        __syncwarp();
        V2W_TRTC(tile_iter - 1, 11 + s * 10);
        if (!K::kResident) {
          for (int t = 0; t < sg.ntaps * K::kChunks; ++t) rw.advance(p.stages_w);
        }
        if (sg.type == kSegConv1) ++n1;
        if (sg.type == kSegConv2) ++n2;
      }
    }
    V2W_TRTC_DUMP("fmma");
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    V2W_TRTC_DECL
    const int half = (warp - 4) >> 2;             // kSub = 2: which 128-row accumulator; kSub = 1: which column half
    const int sub = kSub == 2 ? half : 0;
    const int quarter = warp & 3;                 // TMEM lane quarter
    const int i = sub * kMTile + quarter * 32 + lane;   // window row of this thread
    const uint32_t t_lane = (static_cast<uint32_t>(quarter * 32) << 16) + sub * K::kAccCols;
    const uint32_t swz = K::kKC == 64 ? (i & 7) : (K::kKC == 32 ? ((i >> 1) & 3) : ((i >> 2) & 1));
    constexpr int kNCrow = C / 16;                // 16-column chunks per row
    constexpr int kNC = kSub == 2 ? kNCrow : kNCrow / 2;   // ... per thread and pass
    const int ci0 = kSub == 2 ? 0 : half * kNC;   // first chunk of this thread (even: the look-ahead window's parity holds)
    constexpr int kXc = K::kCacheX0 ? C / 4 : 8;  // cached row, or a 2-chunk look-ahead window
    uint32_t tile_iter = 0, e1 = 0;
    // (compile-time only in the C = 128, 256-row-window stage program -- the one that runs by default)
    const bool x0_blocked = (kBlk || C != 128 || kSub != 2) ? p.x0_blocked != 0 : kX0B;
    const int xs = x0_blocked ? 32 : 1;            // float4 units between column quads of an X0 row
    const int x_nblk = ((p.rows + 31) >> 5) + tc::kBlockedPad;
    for (int ts = slot_first; ts < num_slots; ts += slot_step) {
      if (slot_skipped(ts)) continue;
      const int tile = kPair ? 2 * ts + (int)rank : ts;
      const bool live = tile < total_tiles;          // a peer tile past the end computes on zeros, stores nothing
      const int b = live ? tile / p.tiles_per_b : 0;
      const int ti = live ? tile - b * p.tiles_per_b : 0;
      const int q0c = ti * p.rc;
      const uint32_t tile_it = tile_iter++;
      // ---- per branch: T1 -> A1 = half(lrelu(conv1 + b1 + X0)) in shared memory ----
      {
        const int gr = q0c - p.hc + i;
        const bool valid = live && gr >= 0 && gr < p.rows;
        const int xr_ = valid ? gr : 0;
        const float4* xrow = (x0_blocked
            ? reinterpret_cast<const float4*>(p.x0) + (static_cast<size_t>(b) * x_nblk + (xr_ >> 5)) * (C / 4) * 32 + (xr_ & 31)
            : reinterpret_cast<const float4*>(p.x0 + (static_cast<size_t>(b) * p.rows + xr_) * C)) + ci0 * 4 * xs;   // first chunk of this thread
        float4 xc[kXc];
        if constexpr (K::kCacheX0) if (valid) {
          // this row of X0 is the same for every branch: read it once per tile, before the first
          // accumulator is even ready, so that no global latency sits between the MMAs
#pragma unroll
          for (int e = 0; e < kXc; ++e) xc[e] = __ldg(xrow + e * xs);
        }
        for (int j = 0; j < p.nbranch; ++j, ++e1) {
          const uint32_t slot = e1 % kSlots, ph = (e1 / kSlots) & 1u;
          if constexpr (!K::kCacheX0) if (valid) {   // look-ahead: chunks 0 and 1, before the waits
#pragma unroll
            for (int e = 0; e < 8; ++e) xc[e] = __ldg(xrow + e * xs);
          }
          V2W_TRTC(tile_it, 100 + j * 10);
          mbar_wait(t1_full + 8 * slot, ph);
          mbar_wait(a1_empty + 8 * slot, ph ^ 1u);
          tcgen05_fence_after();
          V2W_TRTC(tile_it, 101 + j * 10);
          const uint32_t a1_row = smem_a1 + slot * K::kA1Bytes + i * kRowBytes;
          const uint32_t b1 = smem_bias + j * (C * 4);
#pragma unroll
          for (int ci = 0; ci < kNC; ++ci) {
            const int cc = (ci0 + ci) * 16;
            float4 xr[4];
            if constexpr (K::kCacheX0) {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[ci * 4 + e];
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) xr[e] = xc[(ci & 1) * 4 + e];
              if (valid && ci + 2 < kNC) {
#pragma unroll
                for (int e = 0; e < 4; ++e) xc[(ci & 1) * 4 + e] = __ldg(xrow + ((ci + 2) * 4 + e) * xs);
              }
            }
            uint32_t r[16];
            tmem_ld16(tmem_base + slot * K::kT1Cols + t_lane + cc, r);
            tmem_ld_wait();
            uint32_t h[8];
            if (valid) {
              float v[16];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 bb = ld_shared_f4(b1 + (cc + 4 * e) * 4);
                const float ax = (kBlk && p.pair) ? 0.f : 1.f;
                v[4 * e] = fmaf(ax, xr[e].x, __uint_as_float(r[4 * e]) + bb.x);
                v[4 * e + 1] = fmaf(ax, xr[e].y, __uint_as_float(r[4 * e + 1]) + bb.y);
                v[4 * e + 2] = fmaf(ax, xr[e].z, __uint_as_float(r[4 * e + 2]) + bb.z);
                v[4 * e + 3] = fmaf(ax, xr[e].w, __uint_as_float(r[4 * e + 3]) + bb.w);
              }
              tc::lrelu_pack16(v, p.mid_slope, p.fp16 != 0, h);
            } else {
#pragma unroll
              for (int e = 0; e < 8; ++e) h[e] = 0u;   // conv2's zero padding outside [0, L)
            }
            // column cc lives in K-chunk cc / kKC; q = 16-byte unit inside the chunk's row
            const uint32_t dst = a1_row + (cc / K::kKC) * K::kA1ChunkBytes;
            const uint32_t q = (cc % K::kKC) >> 3;
            st_shared_v4(dst + ((q ^ swz) << 4), h[0], h[1], h[2], h[3]);
            st_shared_v4(dst + (((q + 1) ^ swz) << 4), h[4], h[5], h[6], h[7]);
          }
          fence_proxy_async();          // generic-proxy writes -> visible to the tensor core
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {              // one arrival per warp
            if constexpr (kPair) {
              mbar_arrive_leader(a1_full + 8 * slot);
              mbar_arrive_leader(t1_empty + 8 * slot);
            } else {
              mbar_arrive(a1_full + 8 * slot);
              mbar_arrive(t1_empty + 8 * slot);
            }
          }
          V2W_TRTC(tile_it, 102 + j * 10);
        }
      }
      // ---- final: out = (T2 + bsum + nk * X0) / nk ----
      {
        const uint32_t t2s = tile_it % kSlots, t2ph = (tile_it / kSlots) & 1u;
        const int gr = q0c + i;
        const bool valid = live && i < p.rc && gr < p.rows;
        const size_t row_base = (static_cast<size_t>(b) * p.rows + (valid ? gr : 0)) * C;
        const int xr_ = valid ? gr : 0;
        const float4* xrow = (x0_blocked
            ? reinterpret_cast<const float4*>(p.x0) + (static_cast<size_t>(b) * x_nblk + (xr_ >> 5)) * (C / 4) * 32 + (xr_ & 31)
            : reinterpret_cast<const float4*>(p.x0 + row_base)) + ci0 * 4 * xs;
        // the row's first column quad in a row-blocked [B, L, C] tensor, float4 units (quad q: + 32 q)
        size_t boff = 0;
        if constexpr (kBlk) boff = (static_cast<size_t>(b) * x_nblk + (xr_ >> 5)) * (C / 4) * 32 + (xr_ & 31);
        float4 xc[kXc];
        if (valid) {   // issued before the wait: overlaps the last conv2 MMAs
#pragma unroll
          for (int e = 0; e < (K::kCacheX0 ? kXc : 8); ++e) xc[e] = __ldg(xrow + e * xs);
        }
        V2W_TRTC(tile_it, 200);
        mbar_wait(t2_full + 8 * t2s, t2ph);
        tcgen05_fence_after();
        V2W_TRTC(tile_it, 201);
#pragma unroll
        for (int ci = 0; ci < kNC; ++ci) {
          const int cc = (ci0 + ci) * 16;
          float4 xr[4];
          if constexpr (K::kCacheX0) {
#pragma unroll
            for (int e = 0; e < 4; ++e) xr[e] = xc[ci * 4 + e];
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) xr[e] = xc[(ci & 1) * 4 + e];
            if (valid && ci + 2 < kNC) {
#pragma unroll
              for (int e = 0; e < 4; ++e) xc[(ci & 1) * 4 + e] = __ldg(xrow + ((ci + 2) * 4 + e) * xs);
            }
          }
          uint32_t r[16];
          tmem_ld16(tmem_t2 + t2s * K::kT1Cols + t_lane + cc, r);
          tmem_ld_wait();
          if (valid) {
            float v[16];
            float4 acc4[4];
            if (kBlk && p.accum_in) {      // (only the conv-pair programs carry a running sum)
              const float4* ap = reinterpret_cast<const float4*>(p.accum_in + row_base + cc);
              int as = 1;
              if constexpr (kBlk) {
                if (p.acc_blocked) {
                  ap = reinterpret_cast<const float4*>(p.accum_in) + boff + (cc >> 2) * 32;
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
              const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bsum + cc + 4 * e));
              v[4 * e] = (__uint_as_float(r[4 * e]) + bb.x + p.res_scale * xr[e].x + acc4[e].x) * p.out_scale;
              v[4 * e + 1] = (__uint_as_float(r[4 * e + 1]) + bb.y + p.res_scale * xr[e].y + acc4[e].y) * p.out_scale;
              v[4 * e + 2] = (__uint_as_float(r[4 * e + 2]) + bb.z + p.res_scale * xr[e].z + acc4[e].z) * p.out_scale;
              v[4 * e + 3] = (__uint_as_float(r[4 * e + 3]) + bb.w + p.res_scale * xr[e].w + acc4[e].w) * p.out_scale;
            }
            if ((kBlk || C != 128 || kSub != 2) && p.out_f32) {   // (the default stage program has no fp32 output)
              float4* op = reinterpret_cast<float4*>(p.out_f32 + row_base + cc);
              int os = 1;
              if constexpr (kBlk) {
                if (p.out_blocked) {
                  op = reinterpret_cast<float4*>(p.out_f32) + boff + (cc >> 2) * 32;
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
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (kPair) mbar_arrive_leader(t2_empty + 8 * t2s);
          else mbar_arrive(t2_empty + 8 * t2s);
        }
        V2W_TRTC(tile_it, 202);
      }
    }
    if (warp == 4) V2W_TRTC_DUMP("fepi");
  }

  tcgen05_fence_before();
  if constexpr (kPair) {
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem_base, K::kTmemCols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, K::kTmemCols);
  }
}

}  // namespace fused
}  // namespace v2w
