// tcgen05 / TMEM / TMA tap-GEMM for sm_100a.
//
//   out[b, q, n] = epilogue( sum_taps sum_c  A[b, q + off_j, c] * W[j][n][c] )
//
// A  : bf16 channels-last activations [B, rows, cin]  (already leaky-relu'd by the producer op)
// W  : bf16 [taps][n_total][cin]  (K-major, like A) -- see v2w_layer_desc in include/v2w.h
// acc: fp32 in TMEM, 128 time steps per accumulator (TMEM lane = time step), N = n_tile columns.
//
// Persistent, warp-specialised CTA (one per SM).  A work item is a SUPER-TILE: `msub` (1, 2 or 4)
// consecutive 128-row sub-tiles of one utterance x one column tile.  Every W block that TMA
// brings in is used by `msub` MMAs, which divides the L2->SMEM weight traffic (64 B/clk/SM at
// msub = 1, above what L2 sustains for all SMs) and amortises barrier round trips.
//
//   warp 0      TMA producer   : A halo tile once per K-chunk, W block once per (K-chunk, tap)
//   warp 1      MMA issuer     : owns the TMEM allocation; one elected lane issues tcgen05.mma
//   warps 4-11  epilogue       : two groups of 4 warps; group g drains accumulator stage g, so the
//                                epilogue of super-tile i overlaps the MMAs of super-tile i+1
//
// Role loops are warp-uniform (all 32 lanes walk the loop, `elect.sync` picks the issuing lane):
// issuing uniform-datapath instructions (UTMALDG / UTCHMMA / UTCBAR) from inside a divergent
// `if (lane == 0)` region makes ptxas wrap each of them in a BRA.U.ANY waterfall loop, and the
// round-1 profile showed the single issuing thread instruction-bound at ~900 clk per step.
// Stage indices and phase bits are running counters (no div/mod in the loops); K-chunk width is
// a template parameter.
//
// Round 2 (DESIGN.md sections 3.1, 3.13, 3.16-3.18), all in Params:
//   * several taps per W stage (tap_group): one barrier round trip serves up to 16 KB of taps;
//   * K-grouped / N-stacked operands for the two launches of a wide stage (kgroups, w_row, res_cols, h_pitch);
//   * split operands (split, template kSplit): fp32-grade products as A_hi W_hi + A_lo W_hi + A_hi W_lo over pairs of
//     16-bit values (fp32tc plans: fp16 pairs; ResBlock1 bf16 plans: bf16 pairs), or the weights only (split == 2);
//   * the epilogue's chunk loop compiled per op flavour (epi_mode), 32 columns per step for the residual-free ones
//     (epi_wide), and the phase-interleaved walk of a transposed conv whose fp32 output is row-blocked (epi_phased).
//
// Halo reuse: the A tile for one K-chunk holds rows [q0 + off_min, q0 + 128*msub + off_max) and is
// loaded ONCE; every tap (and every dilation) is a row-shifted view of it -- the UMMA shared-memory
// descriptor's start address is advanced by shift * row_bytes.  Measured on B200 (tools/tc_probe.py):
// with the 128B/64B/32B swizzles the XOR pattern is a function of the absolute shared-memory
// address, so a shifted view is consistent with what TMA wrote, and the descriptor's base_offset
// field must stay 0 (filling it from the address gives wrong results).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "v2w.h"

// Developer timeline of the generic kernel (build with V2W_NVCC_EXTRA=-DV2W_TRACE_TC): CTA 0 prints SM-clock
// stamps of its MMA warp and of epilogue warp 4 for tile V2W_TRTC_TILE (default 1: steady state).
#ifdef V2W_TRACE_TC
#ifndef V2W_TRTC_TILE
#define V2W_TRTC_TILE 1
#endif
#define V2W_TRTC_DECL long long trc_t[120]; int trc_tag[120]; int trc_n = 0; const long long trc_base = clock64();
#define V2W_TRTC(tile_it, tag) do { if (blockIdx.x == 0 && trc_n < 120 && (int)(tile_it) == V2W_TRTC_TILE) { trc_tag[trc_n] = (tag); trc_t[trc_n++] = clock64() - trc_base; } } while (0)
#define V2W_TRTC_DUMP(name) do { if (blockIdx.x == 0 && lane == 0) for (int q = 0; q < trc_n; ++q) \
    printf("TRTC %s %d %d %lld\n", name, q, trc_tag[q], trc_t[q]); } while (0)
#else
#define V2W_TRTC_DECL
#define V2W_TRTC(tile_it, tag) do { } while (0)
#define V2W_TRTC_DUMP(name) do { } while (0)
#endif

namespace v2w {
namespace tc {

constexpr int kBlockedPad = 4;     // spare 32-row blocks per utterance of a row-blocked tensor (>= the largest upsampling rate)
constexpr int kThreads = 384;
constexpr int kMTile = 128;
constexpr int kMaxStagesA = 4;
constexpr int kMaxStagesW = 8;
constexpr int kMaxSub = 4;
// Output staging per epilogue warp: two buffers per output tensor, each one TMA store box of 32 rows x
// y_cols fp32 / h_cols 16-bit columns (Params::y_cols, h_cols: 16, 32 or 64; sized per op on the host).
// Boxes are as wide as the tile and shared memory allow (plan.cu prepare_tc): one fence, one store issue and
// one wait per box instead of per 16-column chunk.
// per epilogue warp: this tile's bias | scale | shift column vectors (n_tile <= 256 floats each),
// staged once per tile so that the chunk loop reads them with LDS broadcasts instead of L1 round trips
// (3 * n_tile * 4 bytes per warp: Params::cvec_warp_bytes)

struct Params {
  // Variable-length batches (v2w_forward_ragged): utterance b only needs rows below
  // row_limit[b] * rows_per_frame; tiles that start at or past it are skipped by every role.
  // nullptr = uniform batch.
  const int* row_limit;
  int rows_per_frame;
  int cvec_warp_bytes;          // per epilogue warp: bias | scale | shift of the tile's columns
  int batch, rows, cin, n_total, cmod, n_tile, n_tiles;
  int msub;             // 128-row sub-tiles per super-tile
  int s_tiles;          // super-tiles per utterance
  int num_chunks;
  int box_rows;         // rows per TMA box of A
  int stages_a, stages_w;
  int stage_a_bytes, stage_w_bytes;
  // Taps per W stage: one barrier round trip (a ~90 clk try_wait on each side, a commit, the producer's wait for
  // the slot) serves `tap_group` consecutive taps of the list -- each still its own TMA load, into consecutive
  // blocks of the stage.  The narrow layers (C <= 64), whose MMAs are short, were bound by these round trips:
  // ~500 clk per (K-chunk, tap) step whatever the MMA does (tools/trace_small.py).
  int tap_group;
  int acc_cols;         // TMEM columns per accumulator (per sub-tile)
  int fp16;             // 16-bit operand format: 0 = bf16, 1 = fp16
  int tma_out;          // outputs leave through shared-memory staging + TMA stores
  int y_cols, h_cols;   // TMA store box widths (columns) of the fp32 / 16-bit output; multiples of 16 dividing n_tile
  int stage_warp_bytes; // staging per epilogue warp: 2 * 32 * y_cols * 4 (if fp32 out) + 2 * 32 * h_cols * 2 (if 16-bit out)
  // 32-bit entries on purpose: see the note at fused::Segment (ptxas merges 16-bit loads of
  // dynamically indexed parameter arrays incorrectly at some alignments).
  int num_taps[V2W_MAX_TILES];
  int off_min[V2W_MAX_TILES];
  int halo[V2W_MAX_TILES];
  int tap_off[V2W_MAX_TILES][V2W_MAX_TAPS];
  // First row of the (tap list, tap) weight block in the W tensor map (rows of `w_cols` 16-bit
  // elements), column tile included: tap_row * n_total + nt * n_tile for a layer's own [taps][n][cin]
  // array; arbitrary for stacked ops whose blocks live in several arrays (plan.cu TcExtra).
  int w_row[V2W_MAX_TILES][V2W_MAX_TAPS];
  // K-grouped operand (kgroups > 0): the input's channels are `kgroups` groups of `cpg`
  // K-chunks, each with its OWN tap list (list g instead of list nt) and its own weight array whose
  // rows hold only the group's channels -- the branches of a multi-receptive-field stage side by
  // side along K (conv2_j reads A1_j), accumulated in ONE accumulator.  off_min / halo of list 0 span
  // every group.
  int kgroups, cpg;
  // Residual geometry: row pitch and column period of the residual tensor (0 = n_total: dense, no
  // wrap) -- the N-stacked conv1 of a stage adds the SAME [rows, C] residual to every branch's columns.
  int res_pitch, res_cols;
  float res_scale;
  int h_pitch;          // row pitch of the 16-bit output in elements (0 = n_total: dense)
  // Pitched 16-bit output of a polyphase (transposed-conv) op: the GEMM row [up * cout] is `up` consecutive
  // rows of the channels-last signal, each of which starts h_pitch elements after the previous one -- a 4-D
  // tensor (cout, up, rows, batch).  h_cout = cout in that case (4-D tensor map), else 0.
  int h_cout;
  // Row-blocked fp32 output (y_blocked): [batch][nblk][cout / 4][32 rows][4], nblk = ceil(L / 32) + kBlockedPad, with
  // L = rows * up output rows of cout = 1 << y_cout_log2 channels, up = 1 << y_up_log2 -- the layout in which the
  // fused stage kernels read their residual stream with coalesced row-per-thread float4 loads
  // (stacked::Params::x0_blocked).  Written through the same staging + TMA store as a row-major output, with a 5-D
  // tensor map (4 floats, phase, row / up inside a block, column quad, block): every (block, quad) run of the box
  // is 512 contiguous bytes.  (Direct 16-byte stores were measured 1.2-2.2x slower for the whole launch:
  // half-filled lines.)  Output rows of a box past the utterance's end land in the kBlockedPad spare blocks.
  int y_blocked, y_cout_log2, y_up_log2;
  // Split-fp16 operands (fp32 values on the tensor cores, V2W_PREC_FP32_TC): every operand is the pair
  // hi = fp16(v), lo = fp16(v - hi) -- 22 significant bits -- and the product is the three-term sum
  //     A_hi W_hi + A_lo W_hi + A_hi W_lo        (A_lo W_lo ~ 2^-22 of the result, dropped)
  // accumulated in the one fp32 accumulator, i.e. a GEMM over K' = 3 cin walked as three sections of
  // `sec_chunks` K-chunks.  The operand tensor is [rows, hi(cin) | lo(cin)] (lo at column a_lo_col), the weight
  // array [taps][n][cin] of hi values followed w_lo_row rows later by the same array of lo values.  The 16-bit
  // output is written in the same form: hi at its column, lo h_lo_col columns further (a second staging box).
  // split == 2: the weights only -- A (W_hi + W_lo), two sections, plain 16-bit activations in and out.
  int split, sec_chunks, a_lo_col, w_lo_row, h_lo_col;
  int epi_mode;         // compiled copy of the epilogue's chunk loop (0 = all tests at run time), see the kernel
  int epi_wide;         // epi_mode != 0 and y_cols == h_cols == 32: the 32-columns-per-step copy of that loop
  int epi_phased;       // transposed conv, row-blocked fp32 output, one column tile: all phases of 16 channels per step
  const float* bias;
  const float* scale;
  const float* shift;
  const float* residual;   // may alias out_f32 (element-wise in-place update)
  const float* accum_in;   // may alias out_f32
  float out_scale;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  float out_slope;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|p, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must surface as a trapped kernel (sticky CUDA error reported
// through the C-ABI), never as a hung GPU.  The bound is a build-time option: 2^V2W_WAIT_TRAP_LOG2 polls
// (default 2^28, several seconds -- a debugger stop or an MPS time slice must not kill a healthy kernel;
// round 1 trapped after 2^24, ~0.3 s); -DV2W_WAIT_TRAP_LOG2=0 compiles the trap out.
#ifndef V2W_WAIT_TRAP_LOG2
#define V2W_WAIT_TRAP_LOG2 28
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if V2W_WAIT_TRAP_LOG2 > 0
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << V2W_WAIT_TRAP_LOG2)) __trap();
  }
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}
// Programmatic dependent launch (no-ops when the kernel was launched without the attribute)
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const void* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tm, int c0, int c1,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tm, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tm)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tm)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const void* tm, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(
          reinterpret_cast<uint64_t>(tm)),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c,
                                              uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
               "r"(d)
               : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster execute one M = 256 MMA; each holds its own
// 128 rows of A and HALF of the B tile, so shared-memory operand traffic per CTA drops from
// 4 KB + N * 32 B to 4 KB + N * 16 B per MMA.  Shared-memory addresses from cvta are already
// cluster-window addresses of the executing CTA; clearing bit 24 names the same offset in the
// even (leader) CTA of the pair.
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const void* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)),
      "r"(bar & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const void* tm, int c0, int c1, int c2,
                                                 uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)),
      "r"(bar & kPeerMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {   // arrive on the leader CTA's barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & kPeerMask) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // same barrier offset in both CTAs
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(bar), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32"
      "[%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Upper 32 bits of the UMMA shared-memory descriptor of a K-major operand in a swizzled
// row-major tile (cute/arch/mma_sm100_desc.hpp SmemDescriptor: SBO[32,46) version[46,48)=1
// base_offset[49,52)=0 layout_type[61,64)); the lower word is (address >> 4) | LBO(=1) << 16.
template <int KC>
__device__ __forceinline__ uint32_t smem_desc_hi() {
  constexpr uint32_t sbo = 8 * KC * 2;                          // 8 rows of one swizzle span
  constexpr uint32_t layout = KC == 64 ? 2u : (KC == 32 ? 4u : 6u);  // 128B / 64B / 32B swizzle
  return (sbo >> 4) | (1u << 14) | (layout << 29);
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t hi, uint32_t saddr) {
  return (static_cast<uint64_t>(hi) << 32) | ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16 (format 1) or F16 (format 0), both
// K-major, dense (cute/arch/mma_sm100_desc.hpp InstrDescriptor).
__device__ __forceinline__ uint32_t make_idesc_bf16(int m, int n, int fp16 = 0) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// leaky-relu for slopes in [0, 1] (0.1, 0.01 and 1 = identity here): max(v, slope * v) -- two instructions
// instead of compare + multiply + select, same value for every input (NaN and the infinities included)
__device__ __forceinline__ float lrelu(float v, float slope) { return fmaxf(v, v * slope); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// two fp32 -> one packed pair of 16-bit operands (bf16, or fp16 clamped to the finite range)
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi, int fp16) {
  if (fp16) {
    lo = fminf(fmaxf(lo, -65504.f), 65504.f);
    hi = fminf(fmaxf(hi, -65504.f), 65504.f);
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16x2(lo, hi);
}

// sixteen values -> leaky-relu -> eight packed 16-bit pairs with ONE format branch (inlined per pair, the
// `fp16` test becomes a uniform branch per pair: eight per 16-column chunk on the epilogue's critical path)
__device__ __forceinline__ void lrelu_pack16(const float (&v)[16], float slope, bool fp16, uint32_t (&h)[8]) {
  if (fp16) {
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = pack_half2(lrelu(v[2 * e], slope), lrelu(v[2 * e + 1], slope), 1);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) h[e] = pack_bf16x2(lrelu(v[2 * e], slope), lrelu(v[2 * e + 1], slope));
  }
}

// sixteen values -> leaky-relu -> the split-fp16 pair: hi = fp16(v) (clamped to the finite range), lo = fp16(v - hi)
__device__ __forceinline__ void lrelu_split16(const float (&v)[16], float slope, bool fp16, uint32_t (&h)[8],
                                              uint32_t (&l)[8]) {
  if (!fp16) {      // bf16 pairs: 16 significant bits, no range clamp
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float a = lrelu(v[2 * e], slope), b = lrelu(v[2 * e + 1], slope);
      const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
      const float2 back = __bfloat1622float2(hi);
      const __nv_bfloat162 lo = __floats2bfloat162_rn(a - back.x, b - back.y);
      h[e] = *reinterpret_cast<const uint32_t*>(&hi);
      l[e] = *reinterpret_cast<const uint32_t*>(&lo);
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float a = fminf(fmaxf(lrelu(v[2 * e], slope), -65504.f), 65504.f);
    const float b = fminf(fmaxf(lrelu(v[2 * e + 1], slope), -65504.f), 65504.f);
    const __half2 hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    const __half2 lo = __floats2half2_rn(a - back.x, b - back.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hi);
    l[e] = *reinterpret_cast<const uint32_t*>(&lo);
  }
}

struct Ring {   // stage index + phase bit of a circular mbarrier pipeline
  uint32_t s = 0, ph = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++s == n) {
      s = 0;
      ph ^= 1u;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool tile_skipped(const int* row_limit, int rows_per_frame, int b, int row0) {
  return row_limit != nullptr && row0 >= __ldg(row_limit + b) * rows_per_frame;
}

template <int KC, bool kPair = false, bool kSplit = false>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmH,
                  const __grid_constant__ Params p) {
  constexpr int kRowBytes = KC * 2;
  constexpr int kKSteps = KC / 16;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_w = smem_a + p.stages_a * p.stage_a_bytes;
  const uint32_t bars = smem_w + p.stages_w * p.stage_w_bytes;
  // barrier block (8 B each): fullA[4] emptyA[4] fullW[8] emptyW[8] accFull[2] accEmpty[2]
  const uint32_t full_a = bars, empty_a = bars + 32, full_w = bars + 64, empty_w = bars + 128;
  const uint32_t acc_full = bars + 192, acc_empty = bars + 208, tmem_slot = bars + 224;
  const uint32_t stage_base = bars + 1024;   // 1024-aligned: bars is, by construction

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // kPair: a tile is 2 x msub x 128 rows, the CTA of rank r owns the r-th half; both CTAs walk the
  // same tile sequence (tile_first, tile_step in units of pairs)
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const int tile_first = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_step = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tile_rows = (kPair ? 2 : 1) * p.msub * kMTile;
  const int tiles_per_b = p.s_tiles * p.n_tiles;
  const int total_tiles = p.batch * tiles_per_b;
  const uint32_t stage_cols = p.msub * p.acc_cols;   // TMEM columns per accumulator stage
  const uint32_t tmem_cols = 2u * stage_cols;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmA);
    prefetch_tensormap(&tmW);
    for (int i = 0; i < kMaxStagesA; ++i) {
      mbar_init(full_a + 8 * i, 1);
      mbar_init(empty_a + 8 * i, 1);
    }
    for (int i = 0; i < kMaxStagesW; ++i) {
      mbar_init(full_w + 8 * i, 1);
      mbar_init(empty_w + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, kPair ? 8 : 4);   // epilogue warps of both CTAs arrive on the leader's
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    if constexpr (kPair) tmem_alloc_pair(tmem_slot, tmem_cols);
    else tmem_alloc(tmem_slot, tmem_cols);
  }
  tcgen05_fence_before();
  if constexpr (kPair) cluster_sync_all();   // the peer's barriers exist before anything remote touches them
  __syncthreads();   // (also in pair mode: the CTA-level barrier is what orders the TMEM-address read below for racecheck)
  // everything above touched only kernel parameters and the (constant) packed weights: from here
  // on the previous kernel's outputs are read and its inputs overwritten
  pdl_launch_dependents();
  pdl_wait();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    Ring ra, rw;
    const uint32_t box_bytes = p.box_rows * kRowBytes;
    const int w_rows = kPair ? p.n_tile / 2 : p.n_tile;     // rows of the B tile this CTA holds
    const uint32_t w_bytes = w_rows * kRowBytes;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int b = tile / tiles_per_b;
      const int rem = tile - b * tiles_per_b;
      const int st = rem / p.n_tiles;
      const int nt = rem - st * p.n_tiles;
      const int q0t = st * tile_rows;                       // first row of the (pair's) tile
      if (tile_skipped(p.row_limit, p.rows_per_frame, b, q0t)) continue;
      const int q0 = q0t + (int)rank * p.msub * kMTile;     // first row of this CTA's half
      // sub-tiles the MMA warp issues: set by the leader's half (the peer's rows may all lie past
      // the end -- its loads are then zero-filled, its stores clipped)
      const int nsub = min(p.msub, (p.rows - q0t + kMTile - 1) / kMTile);
      const int nbox = (nsub * kMTile + p.halo[nt] + p.box_rows - 1) / p.box_rows;
      const int row0 = q0 + p.off_min[nt];
      // (K-grouped ops list their blocks per group: the column tile's rows are added here)
      const int wrank = (int)rank * w_rows + (p.kgroups ? nt * p.n_tile : 0);
      for (int c = 0; c < p.num_chunks; ++c) {
        const int li = p.kgroups ? c / p.cpg : nt;          // tap list of this K-chunk
        int wc = p.kgroups ? (c - li * p.cpg) * KC : c * KC;   // column in the weight array
        int ac = c * KC, wlo = 0;                           // column in the operand tensor, extra weight rows
        if (p.split) {
          // split == 1, sections 0: A_hi W_hi, 1: A_lo W_hi, 2: A_hi W_lo; split == 2 (weights only), 0: A W_hi, 1: A W_lo
          const int sec = c / p.sec_chunks;
          wc = ac = (c - sec * p.sec_chunks) * KC;
          if (p.split == 1) {
            if (sec == 1) ac += p.a_lo_col;
            if (sec == 2) wlo = p.w_lo_row;
          } else if (sec == 1) {
            wlo = p.w_lo_row;
          }
        }
        const int ntaps = p.num_taps[li];
        mbar_wait(empty_a + 8 * ra.s, ra.ph ^ 1u);
        if (elect_one()) {
          const uint32_t bar = full_a + 8 * ra.s;
          const uint32_t dst = smem_a + ra.s * p.stage_a_bytes;
          if constexpr (kPair) {
            // both CTAs' bytes land on the LEADER's barrier; only the leader arms it
            if (rank == 0) mbar_expect_tx(bar, 2 * nbox * box_bytes);
            for (int bx = 0; bx < nbox; ++bx)
              tma_load_3d_pair(dst + bx * box_bytes, &tmA, ac, row0 + bx * p.box_rows, b, bar);
          } else {
            mbar_expect_tx(bar, nbox * box_bytes);
            for (int bx = 0; bx < nbox; ++bx)
              tma_load_3d(dst + bx * box_bytes, &tmA, ac, row0 + bx * p.box_rows, b, bar);
          }
        }
        __syncwarp();
        ra.advance(p.stages_a);
        for (int j0 = 0; j0 < ntaps; j0 += p.tap_group) {
          const int gn = min(p.tap_group, ntaps - j0);
          mbar_wait(empty_w + 8 * rw.s, rw.ph ^ 1u);
          if (elect_one()) {
            const uint32_t bar = full_w + 8 * rw.s;
            const uint32_t dst = smem_w + rw.s * p.stage_w_bytes;
            if constexpr (kPair) {
              if (rank == 0) mbar_expect_tx(bar, 2 * gn * w_bytes);
              for (int t = 0; t < gn; ++t)
                tma_load_2d_pair(dst + t * w_bytes, &tmW, wc, p.w_row[li][j0 + t] + wrank + wlo, bar);
            } else {
              mbar_expect_tx(bar, gn * w_bytes);
              for (int t = 0; t < gn; ++t)
                tma_load_2d(dst + t * w_bytes, &tmW, wc, p.w_row[li][j0 + t] + wrank + wlo, bar);
            }
          }
          __syncwarp();
          rw.advance(p.stages_w);
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ====================================== MMA issuer ======================================
    // (pair mode: the leader CTA issues M = 256 MMAs for both; commits are multicast)
    Ring ra, rw;
    uint32_t tile_iter = 0;
    V2W_TRTC_DECL
    const uint32_t idesc = make_idesc_bf16(kPair ? 2 * kMTile : kMTile, p.n_tile, p.fp16);
    const uint32_t dhi = smem_desc_hi<KC>();
    const uint32_t w_bytes = (kPair ? p.n_tile / 2 : p.n_tile) * kRowBytes;   // one tap's block in a W stage
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int b = tile / tiles_per_b;
      const int rem = tile - b * tiles_per_b;
      const int st = rem / p.n_tiles;
      const int nt = rem - st * p.n_tiles;
      const int off_min = p.off_min[nt];
      const int q0 = st * tile_rows;
      if (tile_skipped(p.row_limit, p.rows_per_frame, b, q0)) continue;
      const int nsub = min(p.msub, (p.rows - q0 + kMTile - 1) / kMTile);
      const uint32_t as = tile_iter & 1u, aph = (tile_iter >> 1) & 1u;
      ++tile_iter;   // counts processed tiles only
      V2W_TRTC(tile_iter - 1, 1);
      mbar_wait(acc_empty + 8 * as, aph ^ 1u);
      tcgen05_fence_after();
      V2W_TRTC(tile_iter - 1, 2);
      const uint32_t d_tmem = tmem_base + as * stage_cols;
      for (int c = 0; c < p.num_chunks; ++c) {
        const int li = p.kgroups ? c / p.cpg : nt;
        const int ntaps = p.num_taps[li];
        mbar_wait(full_a + 8 * ra.s, ra.ph);
        V2W_TRTC(tile_iter - 1, 10 + c);
        const uint32_t a_stage = smem_a + ra.s * p.stage_a_bytes;
        for (int j0 = 0; j0 < ntaps; j0 += p.tap_group) {
          const int gn = min(p.tap_group, ntaps - j0);
          mbar_wait(full_w + 8 * rw.s, rw.ph);
          tcgen05_fence_after();
          V2W_TRTC(tile_iter - 1, 100 + c * 16 + j0);
          if (elect_one()) {
            for (int t = 0; t < gn; ++t) {
              const int j = j0 + t;
              const uint32_t a_addr = a_stage + (p.tap_off[li][j] - off_min) * kRowBytes;
              const uint32_t w_addr = smem_w + rw.s * p.stage_w_bytes + t * w_bytes;
              const uint32_t acc0 = (c | j) != 0;
              for (int m = 0; m < nsub; ++m) {
                const uint64_t ad = smem_desc(dhi, a_addr + m * (kMTile * kRowBytes));
                const uint64_t bd = smem_desc(dhi, w_addr);
#pragma unroll
                for (int kk = 0; kk < kKSteps; ++kk) {
                  if constexpr (kPair)
                    umma_bf16_pair(d_tmem + m * p.acc_cols, ad + 2 * kk, bd + 2 * kk, idesc,
                                   acc0 | (kk != 0));
                  else
                    umma_bf16(d_tmem + m * p.acc_cols, ad + 2 * kk, bd + 2 * kk, idesc,
                              acc0 | (kk != 0));
                }
              }
            }
            const bool last = j0 + gn == ntaps;
            if constexpr (kPair) {
              umma_commit_pair(empty_w + 8 * rw.s);
              if (last) umma_commit_pair(empty_a + 8 * ra.s);
              if (last && c == p.num_chunks - 1) umma_commit_pair(acc_full + 8 * as);
            } else {
              umma_commit(empty_w + 8 * rw.s);
              if (last) umma_commit(empty_a + 8 * ra.s);
              if (last && c == p.num_chunks - 1) umma_commit(acc_full + 8 * as);
            }
          }
          __syncwarp();
          rw.advance(p.stages_w);
        }
        ra.advance(p.stages_a);
      }
      V2W_TRTC(tile_iter - 1, 3);
    }
    V2W_TRTC_DUMP("mma");
  } else if (warp >= 4) {
    // ======================================= epilogue =======================================
    const int group = (warp - 4) >> 2;  // accumulator stage this warp drains
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const bool pow2 = (p.cmod & (p.cmod - 1)) == 0;
    const uint32_t ybuf = stage_base + (warp - 4) * p.stage_warp_bytes;
    const uint32_t cvec = stage_base + (p.tma_out ? 8 * p.stage_warp_bytes : 0) + (warp - 4) * p.cvec_warp_bytes;
    const uint32_t y_box = p.out_f32 ? 32u * p.y_cols * 4u : 0u;      // bytes of one staging box
    const uint32_t h_box = 32u * p.h_cols * 2u;
    const uint32_t hbuf = ybuf + 2 * y_box;
    const uint32_t lbuf = hbuf + 2 * h_box;                            // split operands: the lo halves' boxes
    const uint32_t y_row = p.y_cols * 4u, h_row = p.h_cols * 2u;      // 64 / 128 and 32 / 64 / 128 bytes
    // XOR of the 16-byte unit index inside a row, as the TMA swizzle of that row size applies it
    const uint32_t y_sw = y_row == 128 ? (lane & 7) : ((lane >> 1) & 3);
    const uint32_t h_sw = h_row == 128 ? (lane & 7) : (h_row == 64 ? ((lane >> 1) & 3) : ((lane >> 2) & 1));
    const int y_sub = p.y_cols >> 4, h_sub = p.h_cols >> 4;           // 16-column chunks per box
    uint32_t sy = 0, sh = 0;                                           // staging buffer in use
    V2W_TRTC_DECL
    uint32_t tile_iter = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step) {
      const int b = tile / tiles_per_b;
      const int rem = tile - b * tiles_per_b;
      const int st = rem / p.n_tiles;
      const int nt = rem - st * p.n_tiles;
      const int q0t = st * tile_rows;
      if (tile_skipped(p.row_limit, p.rows_per_frame, b, q0t)) continue;
      const int q0 = q0t + (int)rank * p.msub * kMTile;     // this CTA's rows
      const uint32_t it = tile_iter++;   // processed tiles only, like the MMA warp's count
      if ((it & 1u) != static_cast<uint32_t>(group)) continue;
      const int nsub = max(0, min(p.msub, (p.rows - q0 + kMTile - 1) / kMTile));
      const uint32_t aph = (it >> 1) & 1u;
      // column vectors of this tile (overlaps the wait for the accumulator)
      for (int c = lane; c < p.n_tile; c += 32) {
        const int col = nt * p.n_tile + c;
        const int cb = pow2 ? (col & (p.cmod - 1)) : (col % p.cmod);
        const float bv = p.bias ? __ldg(p.bias + cb) : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(cvec + 4 * c), "f"(bv) : "memory");
        if (p.scale) {
          const float sv = __ldg(p.scale + static_cast<size_t>(b) * p.cmod + cb);
          const float hv = __ldg(p.shift + static_cast<size_t>(b) * p.cmod + cb);
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(cvec + 4 * (p.n_tile + c)), "f"(sv) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(cvec + 4 * (2 * p.n_tile + c)), "f"(hv) : "memory");
        }
      }
      __syncwarp();
      V2W_TRTC(it, 1);
      mbar_wait(acc_full + 8 * group, aph);
      tcgen05_fence_after();
      V2W_TRTC(it, 2);
      const uint32_t t_lane = tmem_base + group * stage_cols +
                              (static_cast<uint32_t>(quarter * 32) << 16);
      for (int m = 0; m < nsub; ++m) {
        const int row0w = q0 + m * kMTile + quarter * 32;   // first row of this warp's slab
        const int row = row0w + lane;
        const bool valid = row < p.rows;
        const size_t row_lin = static_cast<size_t>(b) * p.rows + row;
        const size_t row_base = row_lin * p.n_total;
        const size_t res_base = p.res_pitch ? row_lin * p.res_pitch : row_base;
        // ---- chunk loop: 16 accumulator columns per step ----
        // Instruction-latency-bound: two epilogue warps per scheduler, one dependent chain each.  The timeline
        // of tools/trace_tc.py showed ~2000 clk per chunk in round 1's loop -- ~1000 clk of ALU work with a
        // uniform branch per packed pair (the 16-bit format test), ~390 clk to issue ONE TMA store from under
        // `if (lane == 0)` (ptxas wraps it in a BRA.U.ANY waterfall loop) and ~300 clk of loop overhead -- which
        // made every k = 3 convolution and every transposed conv epilogue-bound.  So: loop invariants hoisted,
        // the format branch taken once per chunk, no runtime division, stores issued by an ELECTED lane.
        const int col0 = nt * p.n_tile;
        const size_t res_off = res_base + (p.res_cols ? col0 % p.res_cols : col0);   // once per sub-tile
        const size_t acc_off = row_base + col0;
        const uint32_t y_mask = y_sub - 1, h_mask = h_sub - 1;    // boxes are 1, 2 or 4 chunks wide
        const float out_scale = p.out_scale, out_slope = p.out_slope, res_scale = p.res_scale;
        const bool fp16 = p.fp16 != 0;
        constexpr bool split = kSplit;       // (a template parameter: the plain kernels' chunk loop carries none of it)
        // The loop body tests eight op properties (bias, affine, residual, running sum, scale, which outputs, how
        // they leave).  As run-time flags each test is a constant-bank load -> uniform compare -> uniform branch
        // chain per CHUNK (the SASS of round 2's loop: 26 LDCU, 19 UISETP, 20 BRA for ~100 ALU instructions), and
        // these chains, not the arithmetic, were most of the ~870-1700 clk per chunk.  The combinations the plans
        // actually emit are compiled as their own copies of the loop (Params::epi_mode, chosen on the host);
        // mode 0 keeps every test at run time.
        auto chunk_loop = [&](auto mode_c) {
        constexpr int M = decltype(mode_c)::value;
        const bool has_res = valid && (M == 0 ? p.residual != nullptr : M >= 3);
        const bool has_acc = valid && (M == 0 ? p.accum_in != nullptr : false);
        const bool do_scale = M == 0 ? out_scale != 1.0f : M == 4;
        const bool has_bias = M == 0 ? p.bias != nullptr : true;
        const bool has_affine = M == 0 ? p.scale != nullptr : M == 2;
        const bool st_y = M == 0 ? p.out_f32 != nullptr : (M == 2 || M == 5);
        const bool st_h = M == 0 ? p.out_bf16 != nullptr : true;
        const bool tma_out = M == 0 ? p.tma_out != 0 : true;
        for (int cc = 0; cc < p.n_tile; cc += 16) {
          const int col = col0 + cc;
          // issue the global reads first so that they overlap the TMEM load
          float4 res[4], acc[4];
          if (has_res) {
            const float4* rp = reinterpret_cast<const float4*>(p.residual + res_off + cc);
#pragma unroll
            for (int e = 0; e < 4; ++e) res[e] = rp[e];
          }
          if (has_acc) {
            const float4* ap = reinterpret_cast<const float4*>(p.accum_in + acc_off + cc);
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[e] = ap[e];
          }
          uint32_t r[16];
          V2W_TRTC(it, 1000 + cc);
          tmem_ld16(t_lane + m * p.acc_cols + cc, r);
          tmem_ld_wait();
          V2W_TRTC(it, 2000 + cc);
          if (valid) {
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(r[e]);
            if (has_bias) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 t = ld_shared_f4(cvec + (cc + e) * 4);
                v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
              }
            }
            V2W_TRTC(it, 2100 + cc);
            if (has_affine) {
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 s4 = ld_shared_f4(cvec + (p.n_tile + cc + e) * 4);
                const float4 h4 = ld_shared_f4(cvec + (2 * p.n_tile + cc + e) * 4);
                v[e] = fmaf(v[e], s4.x, h4.x);
                v[e + 1] = fmaf(v[e + 1], s4.y, h4.y);
                v[e + 2] = fmaf(v[e + 2], s4.z, h4.z);
                v[e + 3] = fmaf(v[e + 3], s4.w, h4.w);
              }
            }
            V2W_TRTC(it, 2200 + cc);
            if (has_res) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[4 * e] = fmaf(res_scale, res[e].x, v[4 * e]); v[4 * e + 1] = fmaf(res_scale, res[e].y, v[4 * e + 1]);
                v[4 * e + 2] = fmaf(res_scale, res[e].z, v[4 * e + 2]); v[4 * e + 3] = fmaf(res_scale, res[e].w, v[4 * e + 3]);
              }
            }
            if (has_acc) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[4 * e] += acc[e].x; v[4 * e + 1] += acc[e].y;
                v[4 * e + 2] += acc[e].z; v[4 * e + 3] += acc[e].w;
              }
            }
            if (do_scale) {
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] *= out_scale;
            }
            V2W_TRTC(it, 2300 + cc);
            uint32_t h[8], hl[8];
            if (st_h) {
              if (split) lrelu_split16(v, out_slope, fp16, h, hl);
              else lrelu_pack16(v, out_slope, fp16, h);       // one format branch per chunk, not one per packed pair
            }
            V2W_TRTC(it, 2400 + cc);
            if (tma_out) {
              // own row -> staging box in the TMA swizzle layout (conflict-free 16-byte stores)
              if (st_y && p.y_blocked) {
                // staging box in the blocked order [block][quad][row / up inside the block][4 floats]
                const uint32_t rl = 5u - p.y_up_log2;                                  // log2(32 / up)
                const uint32_t q0 = (static_cast<uint32_t>(cc >> 4) & y_mask) * 4u;     // first quad of this chunk in the box
                const uint32_t nq = static_cast<uint32_t>(p.y_cols) >> 2;
                const uint32_t base = ybuf + sy * y_box + (((lane >> rl) * nq) << (rl + 4)) + ((lane & ((1u << rl) - 1u)) << 4);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  st_shared_v4u(base + ((q0 + e) << (rl + 4)), __float_as_uint(v[4 * e]),
                                __float_as_uint(v[4 * e + 1]), __float_as_uint(v[4 * e + 2]),
                                __float_as_uint(v[4 * e + 3]));
              }
              if (st_y && !p.y_blocked) {
                const uint32_t yb = ybuf + sy * y_box + lane * y_row;
                const uint32_t u0 = (static_cast<uint32_t>(cc >> 4) & y_mask) * 4u;   // first 16-byte unit of this chunk
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  st_shared_v4u(yb + (((u0 + e) ^ y_sw) << 4), __float_as_uint(v[4 * e]),
                                __float_as_uint(v[4 * e + 1]), __float_as_uint(v[4 * e + 2]),
                                __float_as_uint(v[4 * e + 3]));
              }
              if (st_h) {
                const uint32_t hb = hbuf + sh * h_box + lane * h_row;
                const uint32_t u0 = (static_cast<uint32_t>(cc >> 4) & h_mask) * 2u;
                st_shared_v4u(hb + ((u0 ^ h_sw) << 4), h[0], h[1], h[2], h[3]);
                st_shared_v4u(hb + (((u0 + 1) ^ h_sw) << 4), h[4], h[5], h[6], h[7]);
                if (split) {
                  const uint32_t lb = lbuf + sh * h_box + lane * h_row;
                  st_shared_v4u(lb + ((u0 ^ h_sw) << 4), hl[0], hl[1], hl[2], hl[3]);
                  st_shared_v4u(lb + (((u0 + 1) ^ h_sw) << 4), hl[4], hl[5], hl[6], hl[7]);
                }
              }
            } else {
              if (st_y && !p.y_blocked) {
                float4* op = reinterpret_cast<float4*>(p.out_f32 + row_base + col);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  op[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
              }
              if (st_h) {
                size_t ho = row_base + col;
                if (p.h_cout) {
                  const int ph = col / p.h_cout;
                  ho = (row_lin * (p.n_total / p.h_cout) + ph) * p.h_pitch + (col - ph * p.h_cout);
                } else if (p.h_pitch) {
                  ho = row_lin * p.h_pitch + col;
                }
                uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + ho);
                op[0] = make_uint4(h[0], h[1], h[2], h[3]);
                op[1] = make_uint4(h[4], h[5], h[6], h[7]);
                if (split) {
                  uint4* lp = reinterpret_cast<uint4*>(p.out_bf16 + ho + p.h_lo_col);
                  lp[0] = make_uint4(hl[0], hl[1], hl[2], hl[3]);
                  lp[1] = make_uint4(hl[4], hl[5], hl[6], hl[7]);
                }
              }
            }
          }
          if (tma_out) {
            // A box is complete when its last 16-column chunk has been staged: the elected lane hands the
            // 32-row box to TMA.  Rows past the utterance end are clipped by the tensor map.  Double buffered
            // per output tensor: after every commit the lane waits until at most the NEWEST group still reads
            // shared memory, so the box written two stores ago (the buffer about to be reused) has been read.
            const uint32_t cidx = static_cast<uint32_t>(cc >> 4);
            const bool y_done = st_y && (cidx & y_mask) == y_mask;
            const bool h_done = st_h && (cidx & h_mask) == h_mask;
            V2W_TRTC(it, 3000 + cc);
            if (y_done || h_done) {
              fence_proxy_async();
              __syncwarp();
              V2W_TRTC(it, 4000 + cc);
              if (row0w < p.rows) {          // warp-uniform
                if (elect_one()) {
                  if (y_done && p.y_blocked) {
                    const int yc0 = col + 16 - p.y_cols;                      // first column of the box
                    const int ph = yc0 >> p.y_cout_log2, c = yc0 - (ph << p.y_cout_log2);
                    const int nblk = (((p.rows << p.y_up_log2) + 31) >> 5) + kBlockedPad;
                    tma_store_5d(&tmY, ybuf + sy * y_box, 0, ph, 0, c >> 2, b * nblk + (row0w >> (5 - p.y_up_log2)));
                  } else if (y_done) {
                    tma_store_3d(&tmY, ybuf + sy * y_box, col + 16 - p.y_cols, row0w, b);
                  }
                  if (h_done) {
                    const int hc0 = col + 16 - p.h_cols;
                    if (p.h_cout) {
                      const int ph = hc0 / p.h_cout;
                      tma_store_4d(&tmH, hbuf + sh * h_box, hc0 - ph * p.h_cout, ph, row0w, b);
                      if (split) tma_store_4d(&tmH, lbuf + sh * h_box, hc0 - ph * p.h_cout + p.h_lo_col, ph, row0w, b);
                    } else {
                      tma_store_3d(&tmH, hbuf + sh * h_box, hc0, row0w, b);
                      if (split) tma_store_3d(&tmH, lbuf + sh * h_box, hc0 + p.h_lo_col, row0w, b);
                    }
                  }
                  tma_store_commit();
                  V2W_TRTC(it, 5000 + cc);
                  tma_store_wait_read<1>();
                }
              }
              __syncwarp();
              V2W_TRTC(it, 6000 + cc);
              if (y_done) sy ^= 1u;
              if (h_done) sh ^= 1u;
            }
          }
        }
        };
        // The same loop, 32 columns per step, for the compiled flavours when both staging boxes are 32 columns wide
        // (Params::epi_wide): two independent 16-column computations per step for the scheduler to interleave (a lone
        // warp issues ~0.25 instructions per clock in the 16-column loop: dependent fixed-latency chains), and the box
        // bookkeeping, fence, store issue and wait once per 32 columns without any "is the box complete" test.
        auto chunk_loop_wide = [&](auto mode_c) {
        constexpr int M = decltype(mode_c)::value;
        static_assert(M >= 1 && M <= 5, "compiled flavours only");
        constexpr bool kRes = M >= 3;
        const bool has_res = valid && kRes;
        constexpr bool do_scale = M == 4, has_affine = M == 2, st_y = (M == 2 || M == 5);
        for (int cc = 0; cc < p.n_tile; cc += 32) {
          const int col = col0 + cc;
          float4 res[8];
          if (has_res) {
            const float4* rp = reinterpret_cast<const float4*>(p.residual + res_off + cc);
#pragma unroll
            for (int e = 0; e < 8; ++e) res[e] = rp[e];
          }
          uint32_t r0[16], r1[16];
          tmem_ld16(t_lane + m * p.acc_cols + cc, r0);
          tmem_ld16(t_lane + m * p.acc_cols + cc + 16, r1);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              float v[16];
#pragma unroll
              for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(u == 0 ? r0[e] : r1[e]);
              const int cu = cc + 16 * u;
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 t = ld_shared_f4(cvec + (cu + e) * 4);
                v[e] += t.x; v[e + 1] += t.y; v[e + 2] += t.z; v[e + 3] += t.w;
              }
              if constexpr (has_affine) {
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                  const float4 s4 = ld_shared_f4(cvec + (p.n_tile + cu + e) * 4);
                  const float4 h4 = ld_shared_f4(cvec + (2 * p.n_tile + cu + e) * 4);
                  v[e] = fmaf(v[e], s4.x, h4.x);
                  v[e + 1] = fmaf(v[e + 1], s4.y, h4.y);
                  v[e + 2] = fmaf(v[e + 2], s4.z, h4.z);
                  v[e + 3] = fmaf(v[e + 3], s4.w, h4.w);
                }
              }
              if constexpr (kRes) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float4 q = res[4 * u + e];
                  v[4 * e] = fmaf(res_scale, q.x, v[4 * e]); v[4 * e + 1] = fmaf(res_scale, q.y, v[4 * e + 1]);
                  v[4 * e + 2] = fmaf(res_scale, q.z, v[4 * e + 2]); v[4 * e + 3] = fmaf(res_scale, q.w, v[4 * e + 3]);
                }
              }
              if constexpr (do_scale) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] *= out_scale;
              }
              uint32_t h[8];
              lrelu_pack16(v, out_slope, fp16, h);
              if constexpr (st_y) {
                if (p.y_blocked) {
                  const uint32_t rl = 5u - p.y_up_log2;
                  const uint32_t base = ybuf + sy * y_box + (((lane >> rl) * 8u) << (rl + 4)) + ((lane & ((1u << rl) - 1u)) << 4);
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    st_shared_v4u(base + ((4u * u + e) << (rl + 4)), __float_as_uint(v[4 * e]),
                                  __float_as_uint(v[4 * e + 1]), __float_as_uint(v[4 * e + 2]),
                                  __float_as_uint(v[4 * e + 3]));
                } else {
                  const uint32_t yb = ybuf + sy * y_box + lane * 128u;
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    st_shared_v4u(yb + (((4u * u + e) ^ (lane & 7u)) << 4), __float_as_uint(v[4 * e]),
                                  __float_as_uint(v[4 * e + 1]), __float_as_uint(v[4 * e + 2]),
                                  __float_as_uint(v[4 * e + 3]));
                }
              }
              const uint32_t hb = hbuf + sh * h_box + lane * 64u;
              const uint32_t hs = (lane >> 1) & 3u;
              st_shared_v4u(hb + (((2u * u) ^ hs) << 4), h[0], h[1], h[2], h[3]);
              st_shared_v4u(hb + (((2u * u + 1u) ^ hs) << 4), h[4], h[5], h[6], h[7]);
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (row0w < p.rows) {          // warp-uniform
            if (elect_one()) {
              if constexpr (st_y) {
                if (p.y_blocked) {
                  const int ph = col >> p.y_cout_log2, c = col - (ph << p.y_cout_log2);
                  const int nblk = (((p.rows << p.y_up_log2) + 31) >> 5) + kBlockedPad;
                  tma_store_5d(&tmY, ybuf + sy * y_box, 0, ph, 0, c >> 2, b * nblk + (row0w >> (5 - p.y_up_log2)));
                } else {
                  tma_store_3d(&tmY, ybuf + sy * y_box, col, row0w, b);
                }
              }
              if (p.h_cout) {
                const int ph = col / p.h_cout;
                tma_store_4d(&tmH, hbuf + sh * h_box, col - ph * p.h_cout, ph, row0w, b);
              } else {
                tma_store_3d(&tmH, hbuf + sh * h_box, col, row0w, b);
              }
              tma_store_commit();
              tma_store_wait_read<1>();
            }
          }
          __syncwarp();
          if constexpr (st_y) sy ^= 1u;
          sh ^= 1u;
        }
        };
        // Transposed conv with a row-blocked fp32 output (Params::epi_phased): one step = ALL phases of 16 output
        // channels, so that every (32-row block, column quad) run of the blocked tensor leaves complete -- 512
        // contiguous bytes per run, `up` blocks x 4 quads per TMA store -- instead of the 16-byte fragments a
        // phase-by-phase walk produces (measured 1.2-2.2x slower for the whole launch).  The 16-bit operand leaves
        // as one (16 channels, up phases, 32 rows) box.  Needs the column tile to hold every phase (n_tile == n_total).
        auto chunk_loop_phased = [&]() {
          const uint32_t upl = p.y_up_log2, up = 1u << upl, rl = 5u - upl;
          const int cout = 1 << p.y_cout_log2;
          const uint32_t yb_bytes = up * 2048u, hb_bytes = up * 1024u, nbuf = up == 2u ? 2u : 1u;
          const uint32_t ybase = ybuf, hbase = ybuf + nbuf * yb_bytes;
          const uint32_t blk = lane >> rl, pos0 = (lane & ((1u << rl) - 1u)) << upl;
          const int nblk = (((p.rows << upl) + 31) >> 5) + kBlockedPad;
          for (int c0 = 0; c0 < cout; c0 += 16) {
            for (uint32_t ph = 0; ph < up; ++ph) {
              const int cc = static_cast<int>(ph) * cout + c0;     // accumulator column of this (phase, channel chunk)
              uint32_t r[16];
              tmem_ld16(t_lane + m * p.acc_cols + cc, r);
              tmem_ld_wait();
              if (valid) {
                float v[16];
#pragma unroll
                for (int e = 0; e < 16; e += 4) {
                  const float4 t = ld_shared_f4(cvec + (cc + e) * 4);
                  const float4 s4 = ld_shared_f4(cvec + (p.n_tile + cc + e) * 4);
                  const float4 h4 = ld_shared_f4(cvec + (2 * p.n_tile + cc + e) * 4);
                  v[e] = fmaf(__uint_as_float(r[e]) + t.x, s4.x, h4.x);
                  v[e + 1] = fmaf(__uint_as_float(r[e + 1]) + t.y, s4.y, h4.y);
                  v[e + 2] = fmaf(__uint_as_float(r[e + 2]) + t.z, s4.z, h4.z);
                  v[e + 3] = fmaf(__uint_as_float(r[e + 3]) + t.w, s4.w, h4.w);
                }
                uint32_t h[8];
                lrelu_pack16(v, out_slope, fp16, h);
                const uint32_t ya = ybase + sy * yb_bytes + ((blk * 4u) << 9) + ((pos0 + ph) << 4);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  st_shared_v4u(ya + (static_cast<uint32_t>(e) << 9), __float_as_uint(v[4 * e]), __float_as_uint(v[4 * e + 1]),
                                __float_as_uint(v[4 * e + 2]), __float_as_uint(v[4 * e + 3]));
                const uint32_t ha = hbase + sh * hb_bytes + lane * (up * 32u) + ph * 32u;
                st_shared_v4u(ha, h[0], h[1], h[2], h[3]);
                st_shared_v4u(ha + 16u, h[4], h[5], h[6], h[7]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (row0w < p.rows) {          // warp-uniform
              if (elect_one()) {
                tma_store_3d(&tmY, ybase + sy * yb_bytes, 0, c0 >> 2, b * nblk + (row0w >> rl));
                tma_store_4d(&tmH, hbase + sh * hb_bytes, c0, 0, row0w, b);
                tma_store_commit();
                if (nbuf == 2u) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
              }
            }
            __syncwarp();
            if (nbuf == 2u) { sy ^= 1u; sh ^= 1u; }
          }
        };
        if constexpr (!kSplit) {
          if (p.epi_phased) {
            chunk_loop_phased();
            continue;
          }
        }
        if constexpr (!kSplit) {
          if (p.epi_wide) {
            // (flavours with a residual stay on the 16-column loop: 32 residual registers on top of two chunks
            // in flight spill at this kernel's 168-register cap)
            if (p.epi_mode == 1) chunk_loop_wide(std::integral_constant<int, 1>{});
            else chunk_loop_wide(std::integral_constant<int, 2>{});
            continue;
          }
        }
        switch (p.epi_mode) {
          case 1: chunk_loop(std::integral_constant<int, 1>{}); break;    // bias; 16-bit out
          case 2: chunk_loop(std::integral_constant<int, 2>{}); break;    // bias, affine; fp32 + 16-bit out (transposed convs)
          case 3: chunk_loop(std::integral_constant<int, 3>{}); break;    // bias, residual; 16-bit out
          case 4: chunk_loop(std::integral_constant<int, 4>{}); break;    // bias, residual, scale; 16-bit out
          case 5: chunk_loop(std::integral_constant<int, 5>{}); break;    // bias, residual; fp32 + 16-bit out
          default: chunk_loop(std::integral_constant<int, 0>{}); break;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {                                      // one arrival per warp
        if constexpr (kPair) mbar_arrive_leader(acc_empty + 8 * group);
        else mbar_arrive(acc_empty + 8 * group);
      }
    }
    if (p.tma_out) {
      if (elect_one()) tma_store_wait_all();      // the same lane that issued (and committed) every store
    }
    if (warp == 4 || warp == 8) V2W_TRTC_DUMP(warp == 4 ? "epi4" : "epi8");
  }

  tcgen05_fence_before();
  if constexpr (kPair) {
    cluster_sync_all();   // nobody frees TMEM or exits while the peer can still signal or read it
    if (warp == 1) tmem_dealloc_pair(tmem_base, tmem_cols);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace tc
}  // namespace v2w
