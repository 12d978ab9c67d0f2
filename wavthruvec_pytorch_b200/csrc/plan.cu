// Plan construction (host), device binding and the forward executor of libv2w.
//
// A plan turns Generator.forward (vec2wav/models.py:116-147) for one (batch, frames, precision)
// into an ordered list of ops over a caller-provided workspace:
//
//   TRANSPOSE_IN   x[B,Cin,T] -> [B,T,Cin]                                  (models.py:123 input)
//   COND_AFFINE    cat(spk,noise) -> fcs[i] -> SN-Linear -> BN fold        (models.py:120,131-134)
//   TAPGEMM        conv_pre / ups[i] (+CBN affine) / resblock convs (+residual, +MRF sum, /3)
//   CONV_POST      leaky_relu(0.01) -> conv_post -> tanh                    (models.py:143-145)
//
// Every elementwise step of the reference forward is folded into the producing GEMM's epilogue or
// the consuming GEMM's load; see DESIGN.md for the table.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>

#include "common.h"
#include "kernels_simt.cuh"
#include "kernels_tc.cuh"
#include "kernels_fused.cuh"
#include "kernels_stacked.cuh"

namespace v2w {

#define V2W_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return V2W_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

struct Act {        // one logical activation tensor [B, rows, C]
  int64_t f32 = -1;   // fp32 copy (pre-activation values)
  int64_t bf16 = -1;  // bf16 copy of leaky_relu(value, slope)
  float slope = 1.f;  // the activation the consumer must see
  int pitch16 = 0;    // row pitch of the 16-bit copy in elements (0 = dense): a column range of a wider tensor
  bool f32_blocked = false;   // fp32 copy in the row-blocked layout (v2w_op_desc::f32_blocked)
};

struct TcLaunch {
  CUtensorMap tm_a, tm_w, tm_y, tm_h;
  tc::Params prm;
  int grid = 0;
  int smem = 0;
  int kc = 64;
  bool pair = false;   // CTA pairs (cta_group::2): cluster of 2, each CTA holds half of the B tile
};

struct FusedLaunch {
  CUtensorMap tm_a, tm_w;
  fused::Params prm;       // plain program (kernels_fused.cuh)
  stacked::Params sprm;    // stacked program (kernels_stacked.cuh)
  bool stacked = false;
  bool pair = false;       // plain program as CTA pairs (cta_group::2), C = 128
  bool narrow = false;     // C = 128 with 128-row windows, two in flight (fused::Cfg<128, 1>)
  int grid = 0;
  int smem = 0;
  int channels = 0;
};

struct Op {
  v2w_op_desc d;
  v2w_tapgemm_args g;   // resolved at bind time (TAPGEMM ops)
  TcLaunch tcl;         // TC engine only
  FusedLaunch fl;       // MRF_FUSED ops
  bool tc_ready = false;
  // train() mode (v2w_forward_train): the same transposed conv with bias only, fp32 result into the
  // statistics scratch -- no affine, no 16-bit copy
  v2w_tapgemm_args g_raw;
  TcLaunch tcl_raw;
  bool has_raw = false;
};

// workspace of a wide stage: the [rows, (nk + 1) * C] 16-bit operand tensor [A1_0 | ... | A1_{nk-1} | A0]
struct WideBuf {
  int64_t a1cat = -1;
  int width = 0;
};

}  // namespace v2w

struct v2w_plan {
  const v2w_model* model = nullptr;
  int batch = 0, frames = 0, precision = 0;
  int packed_format = V2W_HALF_BF16;   // what v2w_plan_bind accepts (v2w_plan_packed_format)
  bool split_operands = false;         // every 16-bit operand tensor is [rows, hi | lo] (v2w_op_desc::split == 1)
  std::vector<v2w::Op> ops;
  int64_t workspace_bytes = 0;
  // workspace offsets
  int64_t stage_x = -1, stage_spk = -1, stage_noise = -1, stage_out = -1;
  int64_t cond_scale[V2W_MAX_STAGES], cond_shift[V2W_MAX_STAGES];
  int64_t xt = -1;        // transposed input
  int64_t post_in = -1;   // fp32 input of conv_post
  int64_t row_limit = -1; // int32 [batch]: frames each utterance needs (variable-length batches)
  // train() mode scratch: raw ups output (fp32, aliases the first-conv buffer of the resblocks, dead at
  // that point), per-stage double sums [2][C], batch mean / rstd [C], 1/sigma
  int64_t tr_y = -1;
  int64_t tr_sums[V2W_MAX_STAGES], tr_isg[V2W_MAX_STAGES];
  int64_t tr_ones = -1, tr_zeros = -1;   // fp32 [max C]: rstd = 1, mean = 0 for the gamma / beta pass
  bool bound = false;
  const uint8_t* packed = nullptr;
  uint8_t* ws = nullptr;
  std::vector<float> post_w;   // host copy of conv_post's folded weights [7][C] (kernel parameter)
  float post_bias = 0.f;
  bool profiling = false;
  std::vector<cudaEvent_t> events;   // ops.size() + 1 when profiling
  // The conditioning kernel only feeds the epilogue of ups[0]: it runs on a side stream, forked at
  // the start of a forward and joined right before the first op that reads its scale/shift.
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // CUDA-graph replay (v2w_forward_graph): one instantiated graph per pointer tuple, LRU
  struct GraphEntry {
    const void *x, *spk, *noise, *out;
    cudaGraphExec_t exec;
    int launches;
    uint64_t stamp;
  };
  std::vector<GraphEntry> graphs;
  uint64_t graph_clock = 0;
  int graph_miss_streak = 0;
  bool ran_once = false;          // lazy per-device initialisation happened outside any capture
  cudaStream_t cap = nullptr;     // capture stream
  void drop_graphs() {
    for (GraphEntry& g : graphs) cudaGraphExecDestroy(g.exec);
    graphs.clear();
    graph_miss_streak = 0;
  }
  ~v2w_plan() {
    drop_graphs();
    if (cap) cudaStreamDestroy(cap);
    for (cudaEvent_t e : events) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    if (side) cudaStreamDestroy(side);
  }
};

namespace v2w {

// ---------------------------------------------------------------------------------------------
// device / driver helpers
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    V2W_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (q != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled not available from the driver");
      return V2W_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return V2W_OK;
}

static int sm_count(int* out) {
  static int n = 0;
  if (!n) {
    int dev = 0;
    V2W_CUDA(cudaGetDevice(&dev));
    int major = 0;
    V2W_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) {
      set_error("libv2w is built for sm_100a only; device %d has compute capability %d.x", dev,
                major);
      return V2W_ERR_CUDA;
    }
    V2W_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  *out = n;
  return V2W_OK;
}

// cudaFuncSetAttribute is per DEVICE: a process that drives more than one GPU must raise the dynamic
// shared-memory limit of every kernel once per device, not once per process.
static uint64_t device_bit() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;   // 0: never cached
  return 1ull << dev;
}

static int pow2_ceil(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// ---------------------------------------------------------------------------------------------
// tap-GEMM launchers
// ---------------------------------------------------------------------------------------------
static int validate_gemm(const v2w_tapgemm_args& a) {
  if (a.batch < 1 || a.rows < 1 || a.cin < 16 || a.cin % 16 || a.n_total < 16 || a.n_total % 16 ||
      a.n_tile < 16 || a.n_tile % 16 || a.n_tiles < 1 || a.n_tiles > V2W_MAX_TILES ||
      a.n_tile * a.n_tiles != a.n_total || a.cmod < 16 || a.cmod % 16) {
    set_error("tapgemm: bad shape batch=%d rows=%d cin=%d n_total=%d n_tile=%d n_tiles=%d cmod=%d",
              a.batch, a.rows, a.cin, a.n_total, a.n_tile, a.n_tiles, a.cmod);
    return V2W_ERR_INVALID;
  }
  for (int t = 0; t < a.n_tiles; ++t) {
    if (a.num_taps[t] < 1 || a.num_taps[t] > V2W_MAX_TAPS) {
      set_error("tapgemm: tile %d has %d taps", t, a.num_taps[t]);
      return V2W_ERR_INVALID;
    }
    for (int j = 1; j < a.num_taps[t]; ++j)
      if (a.tap_off[t][j] < a.tap_off[t][j - 1]) {
        set_error("tapgemm: tap offsets must be ascending");
        return V2W_ERR_INVALID;
      }
  }
  if (!a.in_act || (!a.out_f32 && !a.out_bf16)) {
    set_error("tapgemm: null input or no output");
    return V2W_ERR_INVALID;
  }
  if ((a.scale == nullptr) != (a.shift == nullptr)) {
    set_error("tapgemm: scale and shift must be given together");
    return V2W_ERR_INVALID;
  }
  return V2W_OK;
}

// Geometry beyond a single layer's own arrays (all fields optional, zero = the plain layout): operand /
// output tensors that are column ranges of wider tensors, weights gathered from several packed arrays,
// K-grouped tap lists, residual column wrap -- what the two launches of a wide stage need (v2w_wide_desc).
struct TcExtra {
  int a_pitch = 0;       // elements per row of the input tensor (0 = cin)
  int h_pitch = 0;       // elements per row of the 16-bit output tensor (0 = n_total)
  int h_cout = 0;        // polyphase op with a pitched output: channels of one output row (4-D output tensor)
  int y_blocked_cout = 0; // > 0: fp32 output in the row-blocked layout, rows of this many channels (tc::Params::y_blocked)
  int res_pitch = 0, res_cols = 0;
  float res_scale = 1.f;
  int kgroups = 0, cpg = 0;
  int w_cols = 0;        // 16-bit elements per row of the weight arrays (0 = cin)
  bool abs_rows = false; // w_row[list][tap]: first row of the block relative to `w16`, in rows of w_cols elements
  int w_row[V2W_MAX_TILES][V2W_MAX_TAPS];
  int64_t w_total_rows = 0;
  // split-fp16 operands (tc::Params::split): input [rows, hi | lo] (a_pitch = 2 cin), 16-bit output likewise
  // (h_pitch = 2 cout), the lo weights w_lo_row rows after the hi ones
  int split = 0;         // tc::Params::split: 1 = operands and weights, 2 = weights only
  int w_lo_row = 0;
};

static int prepare_tc(const v2w_tapgemm_args& a, TcLaunch* L, const TcExtra* ex = nullptr) {
  if (!a.in_is_bf16 || a.in_slope != 1.0f) {
    set_error("tapgemm(TC): input must be pre-activated bf16");
    return V2W_ERR_INVALID;
  }
  if (!a.w16) {
    set_error("tapgemm(TC): w16 is null");
    return V2W_ERR_INVALID;
  }
  const int kc = a.cin >= 64 ? 64 : a.cin;
  if ((kc != 16 && kc != 32 && kc != 64) || a.cin % kc) {
    set_error("tapgemm(TC): cin=%d unsupported (16, 32 or a multiple of 64)", a.cin);
    return V2W_ERR_INVALID;
  }
  if (a.n_tile > 256) {
    set_error("tapgemm(TC): n_tile=%d > 256", a.n_tile);
    return V2W_ERR_INVALID;
  }
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  tc::Params& p = L->prm;
  std::memset(&p, 0, sizeof(p));
  p.batch = a.batch; p.rows = a.rows; p.cin = a.cin; p.n_total = a.n_total; p.cmod = a.cmod;
  p.n_tile = a.n_tile; p.n_tiles = a.n_tiles;
  p.num_chunks = a.cin / kc;
  p.res_scale = 1.f;
  const bool split = ex && ex->split == 1;      // operands and weights (weights only: ex->split == 2)
  if (split) {
    if (ex->kgroups || ex->abs_rows || ex->a_pitch != 2 * a.cin ||
        (a.out_bf16 && ex->h_pitch != 2 * (ex->h_cout ? ex->h_cout : a.n_total))) {
      set_error("tapgemm(TC): split operands need [hi | lo] pitches and a plain weight array");
      return V2W_ERR_INVALID;
    }
    p.split = 1;
    p.sec_chunks = p.num_chunks;
    p.num_chunks *= 3;
    p.a_lo_col = a.cin;
    p.w_lo_row = ex->w_lo_row;
    p.h_lo_col = ex->h_pitch / 2;
  } else if (ex && ex->split == 2) {
    if (ex->kgroups || ex->abs_rows) {
      set_error("tapgemm(TC): split weights need a plain weight array");
      return V2W_ERR_INVALID;
    }
    p.split = 2;
    p.sec_chunks = p.num_chunks;
    p.num_chunks *= 2;
    p.w_lo_row = ex->w_lo_row;
  }
  const int nlists = (ex && ex->kgroups) ? ex->kgroups : a.n_tiles;
  if (ex) {
    p.kgroups = ex->kgroups; p.cpg = ex->cpg;
    p.res_pitch = ex->res_pitch; p.res_cols = ex->res_cols; p.res_scale = ex->res_scale;
    if (ex->y_blocked_cout) {
      int lg = 0;
      while ((1 << lg) < ex->y_blocked_cout) ++lg;
      if ((1 << lg) != ex->y_blocked_cout || ex->y_blocked_cout < 16 || a.n_total % ex->y_blocked_cout) {
        set_error("tapgemm(TC): row-blocked fp32 output needs a power-of-two channel count >= 16 (got %d)", ex->y_blocked_cout);
        return V2W_ERR_INVALID;
      }
      const int up = a.n_total / ex->y_blocked_cout;
      int ul = 0;
      while ((1 << ul) < up) ++ul;
      if ((1 << ul) != up || up > 32 || up > tc::kBlockedPad) {
        set_error("tapgemm(TC): row-blocked fp32 output needs a power-of-two upsampling rate <= %d (got %d)", tc::kBlockedPad, up);
        return V2W_ERR_INVALID;
      }
      p.y_blocked = 1;
      p.y_cout_log2 = lg;
      p.y_up_log2 = ul;
    }
    p.h_pitch = ex->h_pitch;
    p.h_cout = (ex->h_pitch && ex->h_cout && ex->h_cout != a.n_total) ? ex->h_cout : 0;
    if (p.h_cout && (a.n_total % p.h_cout || p.h_cout % 16)) {
      set_error("tapgemm(TC): bad pitched polyphase output (n_total %d, cout %d)", a.n_total, p.h_cout);
      return V2W_ERR_INVALID;
    }
    if (ex->kgroups && (ex->kgroups > V2W_MAX_TILES || ex->kgroups * ex->cpg != p.num_chunks)) {
      set_error("tapgemm(TC): bad K-group geometry (%d groups x %d chunks, %d chunks, %d column tiles)",
                ex->kgroups, ex->cpg, p.num_chunks, a.n_tiles);
      return V2W_ERR_INVALID;
    }
  }
  int halo_max = 0;
  for (int t = 0; t < nlists; ++t) {
    p.num_taps[t] = a.num_taps[t];
    p.off_min[t] = a.tap_off[t][0];
    const int halo = a.tap_off[t][a.num_taps[t] - 1] - a.tap_off[t][0];
    p.halo[t] = halo;
    halo_max = std::max(halo_max, halo);
    for (int j = 0; j < a.num_taps[t]; ++j) {
      p.tap_off[t][j] = a.tap_off[t][j];
      p.w_row[t][j] = (ex && ex->abs_rows) ? ex->w_row[t][j] : a.tap_row[t][j] * a.n_total + t * a.n_tile;
    }
  }
  if (ex && ex->kgroups) {
    // one input tile serves every group: its first row / extent span all the tap lists
    int lo = p.off_min[0], hi = p.off_min[0] + p.halo[0];
    for (int t = 1; t < nlists; ++t) {
      lo = std::min(lo, p.off_min[t]);
      hi = std::max(hi, p.off_min[t] + p.halo[t]);
    }
    for (int t = 0; t < V2W_MAX_TILES; ++t) { p.off_min[t] = lo; p.halo[t] = hi - lo; }
    halo_max = hi - lo;
  }
  p.acc_cols = std::max(32, pow2_ceil(a.n_tile));
  p.fp16 = (a.tc_flags & 0x100) ? 1 : 0;   // V2W_TC_FLAG_FP16
  // CTA pairs for wide tiles: the layers with n_tile >= 128 are bound by shared-memory operand
  // delivery (64 + N/2 clk per MMA); a pair shares B, which takes N/4 clk off every MMA.
  static const bool no_pair = [] { const char* e = std::getenv("V2W_TC_NO_PAIR"); return e && std::atoi(e); }();
  L->pair = !no_pair && a.n_tile >= 128 && sms >= 2;
  const int w_rows = L->pair ? a.n_tile / 2 : a.n_tile;
  // Taps per W stage (tc::Params::tap_group): as many as 16 KB hold, at most the longest tap list.
  // $V2W_TC_TAP_GROUP overrides (1 = one tap per barrier round trip, the round-1 behaviour).
  int max_taps = 1;
  for (int t = 0; t < nlists; ++t) max_taps = std::max(max_taps, (int)a.num_taps[t]);
  static const int group_env = [] { const char* e = std::getenv("V2W_TC_TAP_GROUP"); return e ? std::atoi(e) : 0; }();
  // (a row-blocked output stages 12 KB per epilogue warp -- see epi_phased below: 8 KB of taps per stage there, so
  // that two-sub-tile operand stages still fit)
  const int w_stage_cap = (ex && ex->y_blocked_cout) ? 8192 : 16384;
  p.tap_group = std::max(1, std::min(max_taps, w_stage_cap / (w_rows * kc * 2)));
  if (group_env >= 1) p.tap_group = std::min(group_env, max_taps);
  p.stage_w_bytes = (int)align_up((int64_t)p.tap_group * w_rows * kc * 2, 1024);
  // Super-tile height: as many 128-row sub-tiles as TMEM (2 stages x msub x acc_cols <= 512
  // columns) and shared memory allow, while keeping >= 2 work items per SM.
  p.tma_out = std::getenv("V2W_TC_NO_TMA_OUT") ? 0 : 1;
  p.cvec_warp_bytes = 3 * a.n_tile * 4;
  // compiled copies of the epilogue's chunk loop (tc::Params::epi_mode); $V2W_TC_EPI_GENERIC=1: run-time tests only
  static const bool epi_generic = [] { const char* e = std::getenv("V2W_TC_EPI_GENERIC"); return e && std::atoi(e); }();
  p.epi_mode = 0;
  if (!epi_generic && p.tma_out && a.bias && !a.accum_in && a.out_bf16) {
    const bool scaled = a.out_scale != 1.0f;
    if (!a.scale && !a.residual && !scaled && !a.out_f32) p.epi_mode = 1;
    else if (a.scale && !a.residual && !scaled && a.out_f32) p.epi_mode = 2;
    else if (!a.scale && a.residual && !scaled && !a.out_f32) p.epi_mode = 3;
    else if (!a.scale && a.residual && scaled && !a.out_f32) p.epi_mode = 4;
    else if (!a.scale && a.residual && !scaled && a.out_f32) p.epi_mode = 5;
  }
  // TMA store boxes: as wide as the column tile allows, up to 128-byte rows (fp32: 32 columns; 16-bit: 64
  // columns when it is the only output, else 32 -- two outputs with 128-byte rows would take 128 KB of
  // staging): one fence / store issue / wait per box instead of per 16-column chunk.  Measured with the
  // round-2 epilogue (profiles/r02_notes.md): marginal k = 3 tile 7.7 -> 6.8 us, ups[3] 37.0 -> 34.7 us.
  // $V2W_TC_BOX16=1: 16-column boxes.
  static const bool box16 = [] { const char* e = std::getenv("V2W_TC_BOX16"); return e && std::atoi(e); }();
  const int h_unit = (ex && ex->h_cout && ex->h_pitch) ? ex->h_cout : a.n_tile;   // a pitched polyphase box stays inside one phase
  p.y_cols = (!box16 && a.n_tile % 32 == 0) ? 32 : 16;
  p.h_cols = 16;
  if (!box16) {
    if (!a.out_f32 && a.n_tile % 64 == 0 && h_unit % 64 == 0) p.h_cols = 64;
    else if (a.n_tile % 32 == 0 && h_unit % 32 == 0) p.h_cols = 32;
  }
  if (ex && ex->y_blocked_cout && p.y_cols > ex->y_blocked_cout) p.y_cols = 16;   // a blocked box stays inside one phase
  auto staging = [&] {
    return (a.out_f32 ? 2 * 32 * p.y_cols * 4 : 0) + (a.out_bf16 ? (split ? 4 : 2) * 32 * p.h_cols * 2 : 0);
  };
  // (split operands stage two 16-bit boxes per buffer: keep the epilogue's share of shared memory at the 12 KB per
  // warp the widest plain op takes)
  while (staging() > 12 * 1024 && (p.h_cols > 16 || p.y_cols > 16)) {
    if (a.out_bf16 && p.h_cols > 16) p.h_cols >>= 1;
    else p.y_cols >>= 1;
  }
  // 32 columns per epilogue step (tc::Params::epi_wide): both boxes exactly 32 columns.  $V2W_TC_EPI_NARROW=1: off.
  static const bool epi_narrow = [] { const char* e = std::getenv("V2W_TC_EPI_NARROW"); return e && std::atoi(e); }();
  p.epi_wide = 0;
  if (!epi_narrow && (p.epi_mode == 1 || p.epi_mode == 2) && !split && a.n_tile % 32 == 0 && h_unit % 32 == 0 && !box16 &&
      (!a.out_f32 || p.y_cols == 32)) {
    p.h_cols = 32;
    if (staging() <= 12 * 1024) p.epi_wide = 1;
  }
  p.stage_warp_bytes = staging();
  // Row-blocked fp32 output of a transposed conv whose one column tile holds every phase: the phase-interleaved
  // walk (tc::Params::epi_phased).  Staging per warp: nbuf x up x (2 KB fp32 + 1 KB 16-bit).  $V2W_TC_NO_PHASED=1: off.
  static const bool no_phased = [] { const char* e = std::getenv("V2W_TC_NO_PHASED"); return e && std::atoi(e); }();
  p.epi_phased = 0;
  if (!no_phased && p.y_blocked && p.epi_mode == 2 && a.n_tiles == 1 && !split && (ex->h_pitch == 0) &&
      (p.y_up_log2 == 1 || p.y_up_log2 == 2) && ex->y_blocked_cout % 16 == 0) {
    p.epi_phased = 1;
    p.epi_wide = 0;
    p.stage_warp_bytes = (p.y_up_log2 == 1 ? 2 : 1) * (1 << p.y_up_log2) * 3072;
  }
  // barriers (+ output staging) + per-warp column vectors
  const int tail_bytes = 1024 + (p.tma_out ? 8 * p.stage_warp_bytes : 0) + (int)align_up(8 * p.cvec_warp_bytes, 1024);
  const int budget = 227 * 1024 - 1024 - tail_bytes;
  const int forced = std::getenv("V2W_TC_MSUB") ? std::atoi(std::getenv("V2W_TC_MSUB")) : 0;
  int msub = std::min(tc::kMaxSub, 256 / p.acc_cols);
  if (forced >= 1 && forced <= msub) msub = forced;
  // Wave quantisation: among the heights that pass the test above, take the one whose last round wastes
  // the least -- the launch lasts rounds(msub) x msub sub-tile times (cfg2, ups[3]: 500 super-tiles of 4 on
  // 148 SMs = 4 rounds x 4, against 1000 of 2 = 7 rounds x 2).  Ties go to the taller tile (fewer W fetches);
  // $V2W_TC_NO_QUANT=1 restores "tallest that leaves two items per SM".
  static const bool no_step_cost = [] { const char* e = std::getenv("V2W_TC_NO_STEP_COST"); return e && std::atoi(e); }();
  static const bool no_quant = [] { const char* e = std::getenv("V2W_TC_NO_QUANT"); return e && std::atoi(e); }();
  auto geometry = [&](int ms) {
    p.msub = ms;
    p.box_rows = ms == 1 ? 32 : 64;
    const int boxes = (ms * tc::kMTile + halo_max + p.box_rows - 1) / p.box_rows;
    p.stage_a_bytes = boxes * p.box_rows * kc * 2;
    const int tile_rows = (L->pair ? 2 : 1) * ms * tc::kMTile;
    p.s_tiles = (a.rows + tile_rows - 1) / tile_rows;
    return (int64_t)a.batch * p.s_tiles * a.n_tiles;
  };
  const int slots = L->pair ? sms / 2 : sms;
  int best = 0;
  int64_t best_cost = 0;
  for (int ms = msub; ms >= 1; ms >>= 1) {
    const int64_t tiles = geometry(ms);
    const bool fits = 2 * p.stage_a_bytes + 3 * p.stage_w_bytes <= budget;
    const bool ok = ms == 1 || (fits && (tiles >= (L->pair ? sms : 2 * sms) || forced));
    if (!ok) continue;
    if (ms == 1 && !fits) {
      set_error("tapgemm(TC): tile does not fit shared memory (A %d B, W %d B)", p.stage_a_bytes, p.stage_w_bytes);
      return V2W_ERR_INVALID;
    }
    // Duration of a tile, in clocks, fitted to timelines and to launches of one shape at every height
    // (tools/trace_tc.py, trace_small.py; profiles/r02_notes.md section 8).  MMA side: ~1000 clk per tile (accumulator
    // and operand waits, commits) + ~450 clk of barrier round trip per W stage + per tap and sub-tile kc / 16
    // instructions of ~(64 + N / 2) clk (N / 4 for a pair).  Epilogue side: ~2000 clk per tile (column vectors from
    // global memory, accumulator wait, hand-back) + ~1400 clk per 16-column chunk and sub-tile; two warp groups
    // alternate tiles.  Without the fixed parts the narrow layers chose msub = 1 and ran at the barrier rate.
    const int64_t mma_clk = (int64_t)(kc / 16) * (64 + (L->pair ? a.n_tile / 4 : a.n_tile / 2));
    const int64_t w_stages = (int64_t)p.num_chunks * ((max_taps + p.tap_group - 1) / p.tap_group);
    const int64_t mma_tile = 1000 + w_stages * 450 + (int64_t)p.num_chunks * max_taps * ms * mma_clk;
    const int64_t epi_tile = (2000 + (int64_t)ms * ((a.n_tile / 16) * 1400 + 100)) / 2;
    const int64_t cost = ((tiles + slots - 1) / slots) * (no_step_cost ? ms : std::max(mma_tile, epi_tile));
    if (!best || cost < best_cost) {
      best = ms;
      best_cost = cost;
    }
    if (forced || no_quant) break;     // first admissible height, as before
  }
  geometry(best);
  p.stages_a = 2;
  p.stages_w = std::min((budget - 2 * p.stage_a_bytes) / p.stage_w_bytes, tc::kMaxStagesW);
  p.stages_a = std::max(2, std::min((budget - p.stages_w * p.stage_w_bytes) / p.stage_a_bytes,
                                    tc::kMaxStagesA));
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift;
  p.residual = a.residual; p.accum_in = a.accum_in;
  p.out_scale = a.out_scale; p.out_f32 = a.out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  p.out_slope = a.out_slope;
  L->kc = kc;
  L->smem = p.stages_a * p.stage_a_bytes + p.stages_w * p.stage_w_bytes + tail_bytes + 1024;
  const int64_t tiles = (int64_t)a.batch * p.s_tiles * a.n_tiles;
  L->grid = L->pair ? 2 * (int)std::min<int64_t>(tiles, sms / 2) : (int)std::min<int64_t>(tiles, sms);

  EncodeTiledFn enc = nullptr;
  rc = get_encode_fn(&enc);
  if (rc) return rc;
  const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    const cuuint64_t a_pitch = (ex && ex->a_pitch) ? ex->a_pitch : a.cin;
    cuuint64_t dims[3] = {(cuuint64_t)(split ? 2 * a.cin : a.cin), (cuuint64_t)a.rows, (cuuint64_t)a.batch};
    cuuint64_t strides[2] = {a_pitch * 2, (cuuint64_t)a.rows * a_pitch * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)p.box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&L->tm_a, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void*>(a.in_act),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(A) failed with CUresult %d", (int)r);
      return V2W_ERR_CUDA;
    }
  }
  {
    int max_row = 0;
    for (int t = 0; t < a.n_tiles; ++t)
      for (int j = 0; j < a.num_taps[t]; ++j) max_row = std::max(max_row, a.tap_row[t][j]);
    const cuuint64_t w_cols = (ex && ex->w_cols) ? ex->w_cols : a.cin;
    const cuuint64_t w_rows_total = (ex && ex->abs_rows) ? (cuuint64_t)ex->w_total_rows
                                                        : (cuuint64_t)(max_row + 1) * a.n_total + ((ex && ex->split) ? ex->w_lo_row : 0);
    cuuint64_t dims[2] = {w_cols, w_rows_total};
    cuuint64_t strides[1] = {w_cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)w_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L->tm_w, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(a.w16), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(W) failed with CUresult %d", (int)r);
      return V2W_ERR_CUDA;
    }
  }
  std::memset(&L->tm_y, 0, sizeof(L->tm_y));
  std::memset(&L->tm_h, 0, sizeof(L->tm_h));
  if (p.tma_out) {
    // output slabs: 32 rows x 16 columns per epilogue warp and chunk (kernels_tc.cuh)
    cuuint64_t dims[3] = {(cuuint64_t)a.n_total, (cuuint64_t)a.rows, (cuuint64_t)a.batch};
    cuuint64_t dims_h[3] = {(cuuint64_t)(split ? 2 * a.n_total : a.n_total), (cuuint64_t)a.rows, (cuuint64_t)a.batch};
    cuuint32_t box[3] = {(cuuint32_t)p.y_cols, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    auto swz_of = [](int row_bytes) {
      return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                              : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    };
    if (a.out_f32 && p.epi_phased) {
      // (one 512-byte run = 32 rows x 4 floats, column quad, block): boxes of `up` blocks x 4 quads
      const cuuint64_t cout = 1ull << p.y_cout_log2, up = 1ull << p.y_up_log2;
      const cuuint64_t nblk = (((cuuint64_t)a.rows * up + 31) >> 5) + tc::kBlockedPad;
      cuuint64_t dims3[3] = {128, cout / 4, nblk * (cuuint64_t)a.batch};
      cuuint64_t strides3[2] = {512, (cout / 4) * 512};
      cuuint32_t box3[3] = {128, 4, (cuuint32_t)up};
      cuuint32_t estr3[3] = {1, 1, 1};
      CUresult r = enc(&L->tm_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.out_f32, dims3, strides3, box3, estr3,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(out f32, row-blocked runs) failed with CUresult %d", (int)r);
        return V2W_ERR_CUDA;
      }
    } else if (a.out_f32 && p.y_blocked) {
      // (4 floats, phase, GEMM row inside a block, column quad, block): see tc::Params::y_blocked
      const cuuint64_t cout = 1ull << p.y_cout_log2, up = 1ull << p.y_up_log2;
      const cuuint64_t nblk = (((cuuint64_t)a.rows * up + 31) >> 5) + tc::kBlockedPad;
      cuuint64_t dims5[5] = {4, up, 32 / up, cout / 4, nblk * (cuuint64_t)a.batch};
      cuuint64_t strides5[4] = {16, up * 16, 512, (cout / 4) * 512};
      cuuint32_t box5[5] = {4, 1, (cuuint32_t)(32 / up), (cuuint32_t)(p.y_cols / 4), (cuuint32_t)up};
      cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&L->tm_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, a.out_f32, dims5, strides5, box5, estr5,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(out f32, row-blocked) failed with CUresult %d", (int)r);
        return V2W_ERR_CUDA;
      }
    } else if (a.out_f32) {
      cuuint64_t strides[2] = {(cuuint64_t)a.n_total * 4, (cuuint64_t)a.rows * a.n_total * 4};
      CUresult r = enc(&L->tm_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, a.out_f32, dims, strides, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz_of(p.y_cols * 4),
                       CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(out f32) failed with CUresult %d", (int)r);
        return V2W_ERR_CUDA;
      }
    }
    if (a.out_bf16) {
      const cuuint64_t h_pitch = (ex && ex->h_pitch) ? ex->h_pitch : a.n_total;
      cuuint64_t strides[2] = {h_pitch * 2, (cuuint64_t)a.rows * h_pitch * 2};
      CUresult r;
      if (p.epi_phased) {
        // (cout, up, rows, batch) over the dense [rows][up * cout] operand: boxes of 16 channels x every phase x 32 rows
        const cuuint64_t cout = 1ull << p.y_cout_log2, up = 1ull << p.y_up_log2;
        cuuint64_t dims4[4] = {cout, up, (cuuint64_t)a.rows, (cuuint64_t)a.batch};
        cuuint64_t strides4[3] = {cout * 2, up * cout * 2, (cuuint64_t)a.rows * up * cout * 2};
        cuuint32_t box4[4] = {16, (cuuint32_t)up, 32, 1};
        cuuint32_t estr4[4] = {1, 1, 1, 1};
        r = enc(&L->tm_h, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                4, a.out_bf16, dims4, strides4, box4, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else if (p.h_cout) {
        // (cout, up, rows, batch): one GEMM row is `up` output rows, h_pitch elements apart
        const cuuint64_t up = (cuuint64_t)(a.n_total / p.h_cout);
        cuuint64_t dims4[4] = {(cuuint64_t)(split ? 2 * p.h_cout : p.h_cout), up, (cuuint64_t)a.rows, (cuuint64_t)a.batch};
        cuuint64_t strides4[3] = {h_pitch * 2, up * h_pitch * 2, (cuuint64_t)a.rows * up * h_pitch * 2};
        cuuint32_t box4[4] = {(cuuint32_t)p.h_cols, 1, 32, 1};
        cuuint32_t estr4[4] = {1, 1, 1, 1};
        r = enc(&L->tm_h, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                4, a.out_bf16, dims4, strides4, box4, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE,
                swz_of(p.h_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      } else {
        cuuint32_t boxh[3] = {(cuuint32_t)p.h_cols, 32, 1};
        r = enc(&L->tm_h, p.fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                3, a.out_bf16, dims_h, strides, boxh, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                swz_of(p.h_cols * 2), CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      }
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(out 16-bit) failed with CUresult %d", (int)r);
        return V2W_ERR_CUDA;
      }
    }
  }
  return V2W_OK;
}

// Programmatic dependent launch: the tcgen05 kernels run their prologue (barrier init, TMEM
// allocation, descriptor prefetch, bias staging) while the previous kernel's last CTAs are still
// draining, and block in griddepcontrol.wait before they touch anything that kernel wrote.
static bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("V2W_NO_PDL");
    v = (e && std::atoi(e)) ? 0 : 1;
  }
  return v == 1;
}

template <int kCluster, typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st,
                             Args&&... args) {
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (kCluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = kCluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st,
                              Args&&... args) {
  return launch_ex<1>(kernel, grid, block, smem, st, std::forward<Args>(args)...);
}

template <int KC, bool kPair, bool kSplit>
static int launch_tc_as(const TcLaunch& L, cudaStream_t st) {
  auto kernel = tc::tapgemm_tc_kernel<KC, kPair, kSplit>;
  static std::atomic<uint64_t> attr_devs{0};     // (one per instantiation)
  const uint64_t dev_bit = device_bit();
  if (!(attr_devs.load() & dev_bit) || !dev_bit) {
    V2W_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_devs.fetch_or(dev_bit);
  }
  V2W_CUDA(launch_ex<kPair ? 2 : 1>(kernel, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.tm_y, L.tm_h, L.prm));
  V2W_CUDA(cudaGetLastError());
  return V2W_OK;
}

template <bool kPair, bool kSplit>
static int launch_tc_kc(const TcLaunch& L, cudaStream_t st) {
  if (L.kc == 64) return launch_tc_as<64, kPair, kSplit>(L, st);
  if (L.kc == 32) return launch_tc_as<32, kPair, kSplit>(L, st);
  return launch_tc_as<16, kPair, kSplit>(L, st);
}

static int launch_tc(const TcLaunch& L0, cudaStream_t st, const int* row_limit = nullptr,
                     int rows_per_frame = 0) {
  TcLaunch L = L0;
  L.prm.row_limit = row_limit;
  L.prm.rows_per_frame = rows_per_frame;
  if (L.prm.split == 1) return L.pair ? launch_tc_kc<true, true>(L, st) : launch_tc_kc<false, true>(L, st);
  return L.pair ? launch_tc_kc<true, false>(L, st) : launch_tc_kc<false, false>(L, st);
}

template <typename TIn>
static void launch_simt_bn(const simt::Params& p, int bn, dim3 grid, cudaStream_t st) {
  if (bn == 64) simt::tapgemm_simt_kernel<TIn, 64><<<grid, simt::kThreads, 0, st>>>(p);
  else if (bn == 32) simt::tapgemm_simt_kernel<TIn, 32><<<grid, simt::kThreads, 0, st>>>(p);
  else simt::tapgemm_simt_kernel<TIn, 16><<<grid, simt::kThreads, 0, st>>>(p);
}

static int launch_simt(const v2w_tapgemm_args& a, cudaStream_t st, const int* row_limit = nullptr,
                       int rows_per_frame = 0) {
  if (!a.w32) {
    set_error("tapgemm(SIMT): w32 is null");
    return V2W_ERR_INVALID;
  }
  simt::Params p;
  std::memset(&p, 0, sizeof(p));
  p.row_limit = row_limit; p.rows_per_frame = rows_per_frame;
  p.batch = a.batch; p.rows = a.rows; p.cin = a.cin; p.n_total = a.n_total; p.cmod = a.cmod;
  p.n_tile = a.n_tile;
  for (int t = 0; t < a.n_tiles; ++t) {
    p.num_taps[t] = a.num_taps[t];
    p.off_min[t] = a.tap_off[t][0];
    const int halo = a.tap_off[t][a.num_taps[t] - 1] - a.tap_off[t][0];
    if (halo > simt::kMaxHalo) {
      set_error("tapgemm(SIMT): halo %d > %d", halo, simt::kMaxHalo);
      return V2W_ERR_INVALID;
    }
    p.halo[t] = halo;
    for (int j = 0; j < a.num_taps[t]; ++j) {
      p.tap_off[t][j] = a.tap_off[t][j];
      p.tap_row[t][j] = a.tap_row[t][j];
    }
  }
  p.in_act = a.in_act; p.in_slope = a.in_slope; p.w32 = a.w32;
  p.bias = a.bias; p.scale = a.scale; p.shift = a.shift;
  p.residual = a.residual; p.accum_in = a.accum_in;
  p.out_scale = a.out_scale; p.out_f32 = a.out_f32;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a.out_bf16);
  p.out_slope = a.out_slope;
  const int bn = a.n_tile % 64 == 0 ? 64 : (a.n_tile % 32 == 0 ? 32 : 16);
  dim3 grid((a.rows + simt::kBM - 1) / simt::kBM, a.n_total / bn, a.batch);
  if (grid.z > 65535 || grid.y > 65535) {
    set_error("tapgemm(SIMT): grid too large");
    return V2W_ERR_INVALID;
  }
  if (a.in_is_bf16) launch_simt_bn<__nv_bfloat16>(p, bn, grid, st);
  else launch_simt_bn<float>(p, bn, grid, st);
  V2W_CUDA(cudaGetLastError());
  return V2W_OK;
}

// ---------------------------------------------------------------------------------------------
// fused MRF stage launcher
// ---------------------------------------------------------------------------------------------
static int encode_fused_maps(const v2w_fused_desc& F, const v2w_op_desc& d, const uint8_t* packed,
                             uint8_t* ws, bool fp16, int kc, FusedLaunch* L);

// Stacked program (C <= 64): buffering depths and the shared-memory / TMEM layout of
// kernels_stacked.cuh.
static int prepare_stacked(const v2w_model* m, const v2w_fused_desc& F, const v2w_op_desc& d,
                           const uint8_t* packed, uint8_t* ws, bool fp16, int sms, FusedLaunch* L) {
  const int C = F.channels, nb = F.num_branches;
  stacked::Params& p = L->sprm;
  std::memset(&p, 0, sizeof(p));
  L->stacked = true;
  L->channels = C;
  p.batch = d.batch; p.rows = d.rows;
  p.tiles_per_b = (d.rows + F.rc - 1) / F.rc;
  p.rc = F.rc; p.hc = F.hc; p.h1 = F.h1; p.a0_boxes = F.a0_boxes;
  p.nseg = F.num_segments; p.nbranch = nb; p.nent = F.num_taps;
  p.t_bufs = F.t_bufs; p.a0_bufs = F.a0_bufs;
  p.mma_warps = (F.t_bufs == 2 && F.a0_bufs == 2 && C <= 32) ? 2 : 1;
  if (const char* e = std::getenv("V2W_FUSED_MMA_WARPS")) p.mma_warps = std::min(p.mma_warps, std::max(1, std::atoi(e)));
  p.sub_cols = (nb + 1) * C;
  p.tile_cols = 2 * p.sub_cols;
  int cols = 32;
  while (cols < p.t_bufs * p.tile_cols) cols *= 2;
  if (cols > 512) {
    set_error("fused stage: accumulators (%d columns) exceed tensor memory", p.t_bufs * p.tile_cols);
    return V2W_ERR_INVALID;
  }
  p.tmem_cols = cols;
  // C = 64 (streamed weights, one accumulator set): CTA pairs share the weight operand
  static const bool no_cpair = [] { const char* e = std::getenv("V2W_FUSED_NO_PAIR"); return e && std::atoi(e); }();
  L->pair = C == 64 && !no_cpair && sms >= 2;
  int max_n = 0;
  for (int s = 0; s < F.num_segments; ++s) {
    stacked::Segment& g = p.seg[s];
    g.phase = F.seg_type[s] == V2W_FUSED_SEG_CONV2;
    g.src = g.phase ? F.seg_branch[s] : 0;
    g.dcol = F.seg_dcol[s]; g.n = F.seg_n[s];
    g.count = F.seg_ntaps[s]; g.first = F.seg_first[s];
    g.acc = F.seg_acc[s]; g.wait = F.seg_wait[s]; g.commit = F.seg_commit[s];
    // tcgen05 instruction descriptor, kind::f16: D fp32, A/B bf16 (1) or fp16 (0), K-major, M = 128
    const uint32_t fmt = fp16 ? 0u : 1u;
    g.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(g.n >> 3) << 17) | (((L->pair ? 256u : 128u) >> 4) << 24);
    max_n = std::max(max_n, g.n);
  }
  const int row_bytes = C * 2;
  for (int t = 0; t < F.num_taps; ++t) {
    p.a_off16[t] = F.shift[t] * row_bytes / 16;
    p.w_off16[t] = F.w_row[t] * row_bytes / 16;
    p.w_row[t] = F.w_row[t];
  }
  const int a0_bytes = F.a0_boxes * fused::kBoxRows * row_bytes;
  const int a1_bytes = fused::kWin * row_bytes;
  const int budget = 227 * 1024 - 1024 - 512 - (nb + 1) * C * 4 - (F.post ? 2 * 7 * fused::kWin * 4 + 16 : 0);
  const int left = budget - p.a0_bufs * a0_bytes - nb * a1_bytes;
  if (C <= 32) {
    p.stages_w = 1;
    p.w_blocks = F.w_rows / C;
    p.w_region_bytes = (int)align_up((int64_t)F.w_rows * row_bytes, 1024);
    if (p.w_region_bytes > left) {
      set_error("fused stage: resident weights (%d B) do not fit shared memory", p.w_region_bytes);
      return V2W_ERR_INVALID;
    }
  } else {
    // (CTA pairs: each CTA holds HALF of every stacked weight block, so a ring stage is half as big and the
    // ring twice as deep -- 5 instead of 2 stages at C = 64; the timeline showed ~10k of 31k clk per tile in
    // gaps between segments waiting for streamed weights)
    p.stage_w_bytes = (int)align_up((int64_t)(L->pair ? max_n / 2 : max_n) * row_bytes, 1024);
    p.stages_w = std::min(left / p.stage_w_bytes, tc::kMaxStagesW);
    if (p.stages_w < 2) {
      set_error("fused stage: does not fit shared memory");
      return V2W_ERR_INVALID;
    }
    p.w_region_bytes = p.stages_w * p.stage_w_bytes;
  }
  p.x0 = reinterpret_cast<const float*>(ws + d.residual);
  p.x0_blocked = d.f32_blocked;
  p.out_blocked = d.out_blocked;
  p.acc_blocked = d.acc_blocked;
  for (int j = 0; j < nb; ++j)
    p.bias1[j] = reinterpret_cast<const float*>(packed + m->layers[F.first_layer + 2 * F.order[j]].bias_off);
  p.bsum = reinterpret_cast<const float*>(packed + F.bsum_off);
  p.mid_slope = kLreluSlope;
  p.res_scale = F.pair ? 1.0f : (float)nb;
  p.e1_add_x0 = F.pair ? 0 : 1;
  p.preload = F.preload;
  p.accum_in = d.accum_in >= 0 ? reinterpret_cast<const float*>(ws + d.accum_in) : nullptr;
  p.out_scale = d.out_scale;
  p.out_slope = d.out_slope;
  p.out_f32 = d.out_f32 >= 0 ? reinterpret_cast<float*>(ws + d.out_f32) : nullptr;
  p.out_bf16 = d.out_bf16 >= 0 ? reinterpret_cast<__nv_bfloat16*>(ws + d.out_bf16) : nullptr;
  p.fp16 = fp16 ? 1 : 0;
  // + barriers (512) + bias table + 1 KB alignment slack
  const int post_bytes = F.post ? 2 * 7 * fused::kWin * 4 + 16 : 0;   // conv_post partials (kernels_stacked.cuh)
  L->smem = p.a0_bufs * a0_bytes + nb * a1_bytes + p.w_region_bytes + 512 + (nb + 1) * C * 4 + post_bytes + 1024;
  // C = 16: two CTAs per SM (kernels_stacked.cuh); they need 2 x smem and 2 x tmem_cols <= 512
  int ctas_per_sm = 1;
  if (C == 16 && 2 * L->smem <= 227 * 1024 && 2 * p.tmem_cols <= 512) ctas_per_sm = 2;
  if (const char* e = std::getenv("V2W_FUSED_CTAS_PER_SM")) ctas_per_sm = std::min(ctas_per_sm, std::max(1, std::atoi(e)));
  {
    const int64_t tiles = (int64_t)d.batch * p.tiles_per_b;
    L->grid = L->pair ? 2 * (int)std::min<int64_t>((tiles + 1) / 2, sms / 2)
                      : (int)std::min<int64_t>(tiles, (int64_t)sms * ctas_per_sm);
  }
  return encode_fused_maps(F, d, packed, ws, fp16, C, L);
}

static int prepare_fused(const v2w_model* m, const v2w_fused_desc& F, const v2w_op_desc& d,
                         const uint8_t* packed, uint8_t* ws, bool fp16, FusedLaunch* L) {
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  const int C = F.channels;
  if (F.stacked) return prepare_stacked(m, F, d, packed, ws, fp16, sms, L);
  fused::Params& p = L->prm;
  std::memset(&p, 0, sizeof(p));
  p.batch = d.batch; p.rows = d.rows;
  p.tiles_per_b = (d.rows + F.rc - 1) / F.rc;
  p.rc = F.rc; p.hc = F.hc; p.h1 = F.h1; p.a0_boxes = F.a0_boxes;
  p.nseg = F.num_segments; p.nbranch = F.num_branches; p.ntaps_total = F.num_taps;
  for (int s = 0; s < F.num_segments; ++s) {
    p.seg[s].type = F.seg_type[s]; p.seg[s].branch = F.seg_branch[s];
    p.seg[s].ntaps = F.seg_ntaps[s]; p.seg[s].first = F.seg_first[s];
    p.seg[s].flags = F.seg_flags[s] | (F.seg_acc[s] ? fused::kSegFlagAcc : 0);
  }
  for (int t = 0; t < F.num_taps; ++t) p.shift[t] = F.shift[t];
  const int kc = C < 64 ? C : 64;        // K-chunk (fused::Cfg<C>)
  const int win = F.rc + 2 * F.hc;       // window rows: 256, or 128 for the narrow C = 128 variant
  L->narrow = C == 128 && win == 128 && !F.pair;
  const int slots = (C <= 64 || L->narrow) ? 2 : 1;
  static const bool no_pair_f = [] { const char* e = std::getenv("V2W_FUSED_NO_PAIR"); return e && std::atoi(e); }();
  const bool will_pair = C == 128 && !no_pair_f && sms >= 2;
  // one (tap, K-chunk) block; a CTA of a pair holds half of it (twice the ring depth in the same space)
  p.stage_w_bytes = (int)align_up((int64_t)(will_pair ? C / 2 : C) * kc * 2, 1024);
  const int a0_bytes = F.a0_boxes * fused::kBoxRows * C * 2;
  const int a1_bytes = win * C * 2;
  const int bias_bytes = F.num_branches * C * 4;
  const int budget = 227 * 1024 - 1024 - 512 - bias_bytes;
  const int left = budget - slots * (a0_bytes + a1_bytes);
  int w_bytes_total;
  if (C <= 32) {
    // resident weights: every tap of the stage lives in shared memory (kernels_fused.cuh)
    p.stages_w = 1;
    w_bytes_total = (int)align_up((int64_t)F.num_taps * C * C * 2, 1024);
    if (w_bytes_total > left) {
      set_error("fused stage: resident weights (%d B) do not fit shared memory", w_bytes_total);
      return V2W_ERR_INVALID;
    }
  } else {
    p.stages_w = std::min(left / p.stage_w_bytes, tc::kMaxStagesW);
    if (p.stages_w < 2) {
      set_error("fused stage: does not fit shared memory");
      return V2W_ERR_INVALID;
    }
    w_bytes_total = p.stages_w * p.stage_w_bytes;
  }
  p.x0 = reinterpret_cast<const float*>(ws + d.residual);
  p.x0_blocked = d.f32_blocked;
  p.out_blocked = d.out_blocked;
  p.acc_blocked = d.acc_blocked;
  for (int j = 0; j < F.num_branches; ++j)
    p.bias1[j] = reinterpret_cast<const float*>(packed + m->layers[F.first_layer + 2 * j].bias_off);
  p.bsum = reinterpret_cast<const float*>(packed + F.bsum_off);
  p.mid_slope = kLreluSlope;
  p.res_scale = F.pair ? 1.0f : (float)F.num_branches;
  p.pair = F.pair;
  p.accum_in = d.accum_in >= 0 ? reinterpret_cast<const float*>(ws + d.accum_in) : nullptr;
  p.out_scale = d.out_scale;
  p.out_slope = d.out_slope;
  p.out_f32 = d.out_f32 >= 0 ? reinterpret_cast<float*>(ws + d.out_f32) : nullptr;
  p.out_bf16 = d.out_bf16 >= 0 ? reinterpret_cast<__nv_bfloat16*>(ws + d.out_bf16) : nullptr;
  L->channels = C;
  p.w_region_bytes = w_bytes_total;
  p.fp16 = fp16 ? 1 : 0;
  L->smem = slots * (a0_bytes + a1_bytes) + w_bytes_total + 512 + bias_bytes + 1024;
  // C = 128: CTA pairs share the weight operand (kernels_fused.cuh); each CTA holds C / 2 rows of a tap matrix
  static const bool no_pair = [] { const char* e = std::getenv("V2W_FUSED_NO_PAIR"); return e && std::atoi(e); }();
  L->pair = C == 128 && !no_pair && sms >= 2;
  const int64_t tiles = (int64_t)d.batch * p.tiles_per_b;
  L->grid = L->pair ? 2 * (int)std::min<int64_t>((tiles + 1) / 2, sms / 2) : (int)std::min<int64_t>(tiles, sms);

  return encode_fused_maps(F, d, packed, ws, fp16, kc, L);
}

static int encode_fused_maps(const v2w_fused_desc& F, const v2w_op_desc& d, const uint8_t* packed,
                             uint8_t* ws, bool fp16, int kc, FusedLaunch* L) {
  const int C = F.channels;
  EncodeTiledFn enc = nullptr;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const CUtensorMapSwizzle swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                          : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)d.rows, (cuuint64_t)d.batch};
    cuuint64_t strides[2] = {(cuuint64_t)C * 2, (cuuint64_t)d.rows * C * 2};
    cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)fused::kBoxRows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&L->tm_a, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     ws + d.in_act, dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(fused A) failed with CUresult %d", (int)r);
      return V2W_ERR_CUDA;
    }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)F.w_rows};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)(L->pair ? C / 2 : C)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&L->tm_w, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<uint8_t*>(packed + F.w_off), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled(fused W) failed with CUresult %d", (int)r);
      return V2W_ERR_CUDA;
    }
  }
  return V2W_OK;
}

struct PostFold {            // conv_post folded into a fused stage (v2w_op_desc::post)
  const float* w = nullptr;  // host [7][C]
  float bias = 0.f;
  simt::PostOut out;
  const simt::RaggedOffsets* ro = nullptr;
};

static int launch_fused(const FusedLaunch& L0, cudaStream_t st, const int* row_limit = nullptr,
                        int rows_per_frame = 0, const PostFold* post = nullptr) {
  FusedLaunch L = L0;
  L.prm.row_limit = row_limit; L.prm.rows_per_frame = rows_per_frame;
  L.sprm.row_limit = row_limit; L.sprm.rows_per_frame = rows_per_frame;
  if (post) {
    if (!L.stacked || L.channels > 32) {
      set_error("conv_post folding needs the stacked kernel with C <= 32");
      return V2W_ERR_STATE;
    }
    L.sprm.post = 1;
    std::memcpy(L.sprm.post_w, post->w, sizeof(float) * 7 * L.channels);
    L.sprm.post_bias = post->bias;
    L.sprm.post_out = post->out;
    if (post->ro) L.sprm.post_ro = *post->ro;
  }
  static std::atomic<uint64_t> attr_devs{0};
  const uint64_t dev_bit = device_bit();
  if (!(attr_devs.load() & dev_bit) || !dev_bit) {
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, false, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, true, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<64>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<32>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<32>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<16>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    // the instantiations with a row-blocked fp32 output / running sum (ResBlock1 conv pairs)
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, false, 2, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, true, 2, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64, true, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<32, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<16, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, true, 2, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(fused::mrf_stage_kernel<128, false, 2, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64, true, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<64, false, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<32, false, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    V2W_CUDA(cudaFuncSetAttribute(stacked::mrf_stacked_kernel<16, false, false, true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_devs.fetch_or(dev_bit);
  }
  // (every conv-pair program runs the kBlk instantiation; the other one is the ResBlock2 stage program with the pair /
  // running-sum tests compiled out)
  const bool blk = L.stacked ? (L.sprm.out_blocked || L.sprm.acc_blocked || !L.sprm.e1_add_x0 ||
                                (L.channels == 64) != (L.sprm.preload != 0) || L.sprm.out_f32 != nullptr)
                             : (L.prm.out_blocked || L.prm.acc_blocked || L.prm.pair ||
                                (L.channels == 128 && !L.narrow && L.prm.out_f32 != nullptr));
  if (blk) {
    if (L.stacked && L.channels == 64 && L.pair)
      V2W_CUDA(launch_ex<2>(stacked::mrf_stacked_kernel<64, true, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.stacked && L.channels == 64)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<64, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.stacked && L.channels == 32)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<32, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.stacked && L.channels == 16)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<16, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (!L.stacked && L.channels == 128 && !L.narrow && L.pair)
      V2W_CUDA(launch_ex<2>(fused::mrf_stage_kernel<128, true, 2, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
    else if (!L.stacked && L.channels == 128 && !L.narrow)
      V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<128, false, 2, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
    else {
      set_error("fused stage: no kernel with a row-blocked output for C = %d%s", L.channels, L.narrow ? " (narrow windows)" : "");
      return V2W_ERR_STATE;
    }
    V2W_CUDA(cudaGetLastError());
    return V2W_OK;
  }
  if (L.stacked && L.sprm.x0_blocked) {      // stage programs with a row-blocked residual stream
    if (L.channels == 64 && L.pair)
      V2W_CUDA(launch_ex<2>(stacked::mrf_stacked_kernel<64, true, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.channels == 64)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<64, false, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.channels == 32)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<32, false, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<16, false, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    V2W_CUDA(cudaGetLastError());
    return V2W_OK;
  }
  if (L.stacked) {
    if (L.channels == 64 && L.pair)
      V2W_CUDA(launch_ex<2>(stacked::mrf_stacked_kernel<64, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.channels == 64)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<64>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else if (L.channels == 32)
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<32>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    else
      V2W_CUDA(launch_pdl(stacked::mrf_stacked_kernel<16>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.sprm));
    V2W_CUDA(cudaGetLastError());
    return V2W_OK;
  }
  if (L.channels == 128 && !L.narrow && L.prm.x0_blocked) {      // stage program with a row-blocked residual stream
    if (L.pair)
      V2W_CUDA(launch_ex<2>(fused::mrf_stage_kernel<128, true, 2, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
    else
      V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<128, false, 2, false, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
    V2W_CUDA(cudaGetLastError());
    return V2W_OK;
  }
  if (L.channels == 128 && L.narrow && L.pair)
    V2W_CUDA(launch_ex<2>(fused::mrf_stage_kernel<128, true, 1>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else if (L.channels == 128 && L.narrow)
    V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<128, false, 1>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else if (L.channels == 128 && L.pair)
    V2W_CUDA(launch_ex<2>(fused::mrf_stage_kernel<128, true>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else if (L.channels == 128)
    V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<128>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else if (L.channels == 64)
    V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<64>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else if (L.channels == 32)
    V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<32>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  else
    V2W_CUDA(launch_pdl(fused::mrf_stage_kernel<16>, L.grid, tc::kThreads, L.smem, st, L.tm_a, L.tm_w, L.prm));
  V2W_CUDA(cudaGetLastError());
  return V2W_OK;
}

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
struct Bump {
  int64_t off = 0;
  int64_t take(int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  }
};

// Pick the column tile of a layer.  Candidates: whole phase groups (or divisors of cout) of 16..256
// columns.  Cost model from the round-1 measurements: one M=128, K=16 MMA costs ~(64 + N/2) clk
// (shared-memory operand delivery ~64 B/clk: profiles/r01_notes.md), so a tile costs
// taps * (128 + n_tile) units and the launch takes ceil(tiles / SMs) rounds of it.  Zero polyphase
// blocks count as issued taps.  Ties go to the larger tile (fewer A re-reads).
static void choose_tiles(const v2w_layer_desc& L, int batch, int rows, v2w_op_desc* d) {
  int best = 0;
  double best_cost = 0;
  const int nt_max = std::getenv("V2W_NTILE_MAX") ? std::atoi(std::getenv("V2W_NTILE_MAX")) : 256;
  const int sms = 148;
  const int m_tiles = (rows + 127) / 128;
  for (int nt = 16; nt <= nt_max; nt <<= 1) {
    if (L.n_total % nt) continue;
    if (nt % L.cout && L.cout % nt) continue;
    const int n_tiles = L.n_total / nt;
    if (n_tiles > V2W_MAX_TILES) continue;
    int64_t taps_sum = 0;
    for (int t = 0; t < n_tiles; ++t) {
      const int p_lo = t * nt / L.cout, p_hi = ((t + 1) * nt - 1) / L.cout;
      uint32_t pm = 0;
      for (int p = p_lo; p <= p_hi; ++p) pm |= 1u << p;
      for (int j = 0; j < L.num_taps; ++j) taps_sum += (L.tap_mask[j] & pm) != 0;
    }
    const int64_t tiles = (int64_t)batch * m_tiles * n_tiles;
    const double rounds = (double)((tiles + sms - 1) / sms);
    const double cost = rounds * ((double)taps_sum / n_tiles) * (128 + nt);
    if (!best || cost < best_cost || (cost == best_cost && nt > best)) {
      best = nt;
      best_cost = cost;
    }
  }
  d->n_tile = best;
  d->n_tiles = L.n_total / best;
  for (int t = 0; t < d->n_tiles; ++t) {
    const int p_lo = t * best / L.cout, p_hi = ((t + 1) * best - 1) / L.cout;
    uint32_t pm = 0;
    for (int p = p_lo; p <= p_hi; ++p) pm |= 1u << p;
    int n = 0;
    for (int j = 0; j < L.num_taps; ++j)
      if (L.tap_mask[j] & pm) d->tap_index[t][n++] = j;
    d->num_taps[t] = n;
  }
}

}  // namespace v2w

using namespace v2w;

extern "C" {

int32_t v2w_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess &&
        major == 10)
      ++ok;
  }
  return ok;
}

int32_t v2w_plan_create(const v2w_model* m, int32_t batch, int32_t frames, int32_t precision,
                        v2w_plan** out) {
  if (!m || !out) {
    set_error("v2w_plan_create: null argument");
    return V2W_ERR_INVALID;
  }
  *out = nullptr;
  if (batch < 1 || frames < 1 || batch > 65535) {
    set_error("v2w_plan_create: batch=%d frames=%d out of range", batch, frames);
    return V2W_ERR_INVALID;
  }
  if (precision != V2W_PREC_FP32 && precision != V2W_PREC_BF16 && precision != V2W_PREC_FP16 &&
      precision != V2W_PREC_FP32_TC) {
    set_error("v2w_plan_create: unknown precision %d", precision);
    return V2W_ERR_INVALID;
  }
  if ((int64_t)frames * m->total_up > (1 << 30)) {
    set_error("v2w_plan_create: %d frames is too long", frames);
    return V2W_ERR_INVALID;
  }
  v2w_plan* P = new (std::nothrow) v2w_plan();
  if (!P) {
    set_error("out of memory");
    return V2W_ERR_NOMEM;
  }
  P->model = m;
  P->batch = batch;
  P->frames = frames;
  P->precision = precision;
  const bool bf = precision != V2W_PREC_FP32;   // 16-bit tensor-core operands (bf16 or fp16)
  // fp32 values on the tensor cores: every operand tensor is [rows, hi | lo] fp16 (tc::Params::split), one
  // tap-GEMM per convolution (the fused stage kernels keep their operands in shared memory as single 16-bit values)
  const v2w_config& c = m->cfg;
  // ResBlock1 with bf16 operands: the same mechanism with bf16 pairs (16 significant bits), see below
  const char* rb1_env = std::getenv("V2W_RB1_BF16");
  const bool rb1_bf16 = precision == V2W_PREC_BF16 && c.resblock_type == 1 &&
                        !(std::getenv("V2W_RB1_BF16_FUSED") && std::atoi(std::getenv("V2W_RB1_BF16_FUSED")));
  const bool rb1_weights_only = rb1_bf16 && rb1_env && std::strcmp(rb1_env, "weights") == 0;
  const bool split = precision == V2W_PREC_FP32_TC || (rb1_bf16 && !rb1_weights_only);
  const int hsz = split ? 4 : 2;                // bytes per element of a 16-bit operand tensor
  // ResBlock1 with bf16 operands: single bf16 values take its 97-conv chain below the 45 dB bar (40.8 dB on the
  // stress fixture; rounding the WEIGHTS alone accounts for it).  With the weights as hi + lo (two MMAs per product,
  // v2w_op_desc::split == 2, $V2W_RB1_BF16=weights) it is 46.2-46.6 dB there but 40 dB on a harsher state -- what is
  // left is the 8-bit rounding of the activations.  Default: operands AND weights as bf16 pairs (split == 1, three
  // MMAs per product, 16 significant bits), one tap-GEMM per convolution.  $V2W_RB1_BF16_FUSED=1: the single-rounded
  // weights in the fused conv pairs, as in round 1 (fastest, 40.8 dB).
  const bool wsplit = rb1_weights_only;
  P->split_operands = split;
  P->packed_format = precision == V2W_PREC_FP32_TC ? V2W_HALF_FP16_SPLIT
                   : (rb1_bf16 ? V2W_HALF_BF16_SPLIT : (precision == V2W_PREC_FP16 ? V2W_HALF_FP16 : V2W_HALF_BF16));
  const int U = c.num_upsamples;
  const int B = batch, T = frames;
  const bool force_simt = std::getenv("V2W_FORCE_SIMT") != nullptr;
  const bool no_fuse = split || wsplit || std::getenv("V2W_NO_FUSE") != nullptr;
  bool post_folded = false;   // conv_post runs inside the last fused stage
  const int fuse_mask = std::getenv("V2W_FUSE_MASK") ? std::atoi(std::getenv("V2W_FUSE_MASK")) : -1;

  Bump ws;
  P->stage_x = ws.take((int64_t)B * c.num_wv_feat * T * 4);
  P->stage_spk = ws.take((int64_t)B * c.spk_dim * 4);
  P->stage_noise = ws.take((int64_t)B * std::max(1, c.noise_dim) * 4);
  P->stage_out = ws.take((int64_t)B * T * m->total_up * 4);
  for (int i = 0; i < U; ++i) {
    P->cond_scale[i] = ws.take((int64_t)B * m->stage_channels[i] * 4);
    P->cond_shift[i] = ws.take((int64_t)B * m->stage_channels[i] * 4);
  }
  int64_t max_e = 0;
  for (int i = 0; i < U; ++i)
    max_e = std::max(max_e, (int64_t)B * T * m->stage_up[i] * m->stage_channels[i]);
  const int64_t esz = bf ? hsz : 4;
  P->row_limit = ws.take((int64_t)B * 4);
  P->xt = ws.take((int64_t)B * T * c.num_wv_feat * esz);
  Act pre;   // conv_pre output, leaky-relu'd for ups[0] (models.py:128)
  pre.slope = kLreluSlope;
  if (bf) pre.bf16 = ws.take((int64_t)B * T * c.upsample_initial_channel * hsz);
  else pre.f32 = ws.take((int64_t)B * T * c.upsample_initial_channel * 4);
  Act x0, y1, t1;
  x0.slope = y1.slope = t1.slope = kLreluSlope;
  // (+ 32 rows per utterance: the row-blocked layout of the fused stages' residual stream pads L to 32)
  int64_t cmax_stage = 0;
  for (int i = 0; i < U; ++i) cmax_stage = std::max<int64_t>(cmax_stage, m->stage_channels[i]);
  const int64_t blocked_pad = (int64_t)B * 32 * (1 + tc::kBlockedPad) * cmax_stage * 4;
  x0.f32 = ws.take(max_e * 4 + blocked_pad);
  y1.f32 = ws.take(max_e * 4 + blocked_pad);
  const int64_t sum_f32 = ws.take(max_e * 4 + blocked_pad);
  if (bf) {
    x0.bf16 = ws.take(max_e * hsz);
    y1.bf16 = ws.take(max_e * hsz);
  }
  Act y2;    // ResBlock1, fused conv pairs: a pair reads its input WITH halo rows, so it cannot run in place
  y2.slope = kLreluSlope;
  if (c.resblock_type == 1) {
    if (bf) t1.bf16 = ws.take(max_e * hsz);
    else t1.f32 = ws.take(max_e * 4);
    if (bf) {
      y2.f32 = ws.take(max_e * 4 + blocked_pad);
      y2.bf16 = ws.take(max_e * hsz);
    }
  }
  const int64_t next_bf16 = bf ? ws.take(max_e * hsz) : -1;
  // Wide stages (two stacked tap-GEMM launches, v2w_wide_desc): by default the stages no fused kernel
  // covers (C = 256); $V2W_WIDE_MASK = bitmask of stages overrides (0 = the six separate tap-GEMMs).
  const int wide_mask = std::getenv("V2W_WIDE_MASK") ? std::atoi(std::getenv("V2W_WIDE_MASK")) : -1;
  std::vector<WideBuf> widebuf((size_t)U);
  for (int i = 0; i < U; ++i) {
    const bool fusable = m->fused[i].ok && ((fuse_mask >> i) & 1);
    const bool want = wide_mask >= 0 ? ((wide_mask >> i) & 1) : !fusable;
    if (bf && !force_simt && !no_fuse && m->wide[i].ok && want) {
      widebuf[i].width = (m->wide[i].num_branches + 1) * m->stage_channels[i];
      widebuf[i].a1cat = ws.take((int64_t)B * T * m->stage_up[i] * widebuf[i].width * 2);
    }
  }
  P->tr_y = y1.f32;
  for (int i = 0; i < U; ++i) {
    P->tr_sums[i] = ws.take((int64_t)2 * m->stage_channels[i] * 8);
    P->tr_isg[i] = ws.take(4);
  }
  {
    int cmax = 0;
    for (int i = 0; i < U; ++i) cmax = std::max(cmax, m->stage_channels[i]);
    P->tr_ones = ws.take((int64_t)cmax * 4);
    P->tr_zeros = ws.take((int64_t)cmax * 4);
  }
  P->workspace_bytes = ws.off;
  P->post_in = sum_f32;

  auto blank = [](int kind) {
    Op o;
    std::memset(&o.d, 0, sizeof(o.d));
    std::memset(&o.g, 0, sizeof(o.g));
    o.d.kind = kind;
    o.d.layer = -1;
    o.d.in_act = o.d.residual = o.d.accum_in = o.d.out_f32 = o.d.out_bf16 = -1;
    o.d.cond_stage = -1;
    o.d.stage = -1;
    o.d.pair = -1;
    o.d.post = 0;
    o.d.in_slope = 1.f;
    o.d.out_scale = 1.f;
    o.d.out_slope = 1.f;
    return o;
  };
  {
    Op o = blank(V2W_OP_TRANSPOSE_IN);
    o.d.batch = B; o.d.rows = T;
    o.d.out_f32 = bf ? -1 : P->xt;
    o.d.out_bf16 = bf ? P->xt : -1;
    P->ops.push_back(o);
  }
  {
    Op o = blank(V2W_OP_COND_AFFINE);
    o.d.batch = B;
    P->ops.push_back(o);
  }
  auto emit = [&](int layer, int rows, const Act& in, int cond_stage, int64_t residual,
                  int64_t accum, float out_scale, const Act& out) {
    const v2w_layer_desc& L = m->layers[layer];
    Op o = blank(V2W_OP_TAPGEMM);
    o.d.engine = (bf && (!force_simt || precision == V2W_PREC_FP16 || split)) ? V2W_ENGINE_TC : V2W_ENGINE_SIMT;
    o.d.layer = layer;
    o.d.batch = B;
    o.d.rows = rows;
    choose_tiles(L, B, rows, &o.d);
    if (bf && out.f32 >= 0 && out.f32_blocked && o.d.n_tiles > 1 && L.n_total <= 256) {
      // a row-blocked output leaves complete 512-byte runs only when one column tile holds every phase
      // (tc::Params::epi_phased): one tile with the union of the taps (dead (tap, phase) blocks are zero weights)
      o.d.n_tile = L.n_total;
      o.d.n_tiles = 1;
      o.d.num_taps[0] = L.num_taps;
      for (int j = 0; j < L.num_taps; ++j) o.d.tap_index[0][j] = j;
    }
    if (bf) {
      o.d.in_act = in.bf16; o.d.in_is_bf16 = 1; o.d.in_slope = 1.f;
      o.d.out_f32 = out.f32; o.d.out_bf16 = out.bf16; o.d.out_slope = out.slope;
      o.d.out_bf16_pitch = out.pitch16;
      o.d.f32_blocked = (out.f32 >= 0 && out.f32_blocked) ? 1 : 0;
      if (split) {
        o.d.split = 1;
        o.d.in_pitch = 2 * L.cin;
        o.d.out_bf16_pitch = out.bf16 >= 0 ? 2 * L.cout : 0;
      } else if (wsplit && o.d.engine == V2W_ENGINE_TC) {
        o.d.split = 2;
      }
    } else {
      o.d.in_act = in.f32; o.d.in_is_bf16 = 0; o.d.in_slope = in.slope;
      o.d.out_f32 = out.f32; o.d.out_bf16 = -1; o.d.out_slope = 1.f;
    }
    o.d.cond_stage = cond_stage;
    o.d.residual = residual;
    o.d.accum_in = accum;
    o.d.out_scale = out_scale;
    o.d.flops = 2 * L.macs_per_row * (int64_t)rows * B;
    P->ops.push_back(o);
  };

  // layer indices follow v2w_model_create's push order
  int li = 0;
  Act xin;  // transposed input, no activation (models.py:123)
  xin.slope = 1.f;
  if (bf) xin.bf16 = P->xt; else xin.f32 = P->xt;
  emit(li++, T, xin, -1, -1, -1, 1.f, pre);
  Act cur_in = pre;
  int rows = T;
  const int nk = c.num_kernels;
  for (int i = 0; i < U; ++i) {
    const bool last_stage = i == U - 1;
    // x = cbns[i](ups[i](leaky_relu(x)), fcs[i](spk_noise))   (models.py:128-134)
    const bool wide = widebuf[i].a1cat >= 0;
    // A fused stage reads its residual stream row-per-thread; in the row-blocked layout (v2w_op_desc::f32_blocked)
    // those reads are coalesced, which takes 12-15 % off the C = 128 and 9 % off the C = 32 stage kernel.  The
    // transposed conv that PRODUCES the tensor must then emit complete 512-byte runs: with one column tile that
    // holds every phase it walks its accumulator phase-interleaved (tc::Params::epi_phased; rate-2 stages: same
    // speed as the row-major output), otherwise it stores 16-byte fragments and loses more than the stage gains
    // except at C = 128.  Measured per stage (profiles/r02_notes.md section 9): stage 1 -10 us at cfg2, stage 3
    // -9 us (-100 us at cfg5 per GPU), stage 2 (rate 4: single-buffered staging) and stage 4 a wash.
    // $V2W_BLOCKED_X0_MASK = bitmask of stages; default: stages 1 and 3.
    const char* blocked_env = std::getenv("V2W_BLOCKED_X0_MASK");
    // (ResBlock1: the conv-pair kernels of the C <= 128 stages read it the same way, three times per stage -- the
    // first pair of every block -- so that every fused stage wins: default there stages 1-4)
    const bool rb1_pairs = c.resblock_type == 1 && bf && !force_simt && !no_fuse && ((fuse_mask >> i) & 1) &&
                           m->stage_channels[i] <= 128 && m->stage_channels[i] >= 16;
    const int blocked_mask = blocked_env ? std::atoi(blocked_env) : (c.resblock_type == 1 ? 30 : 10);
    const bool stage_fused = bf && !force_simt && !no_fuse && !wide && m->fused[i].ok && ((fuse_mask >> i) & 1);
    const bool x0_blocked = (stage_fused || rb1_pairs) && ((blocked_mask >> i) & 1) && m->stage_channels[i] >= 16 &&
                            c.upsample_rates[i] <= tc::kBlockedPad && (c.upsample_rates[i] & (c.upsample_rates[i] - 1)) == 0;
    if (x0_blocked) {
      Act x0b = x0;
      x0b.f32_blocked = true;
      emit(li++, rows, cur_in, i, -1, -1, 1.f, x0b);
    } else if (wide) {
      // the 16-bit copy of the stage input goes straight into the last C columns of the stage's operand tensor
      Act x0w = x0;
      x0w.bf16 = widebuf[i].a1cat + (int64_t)m->wide[i].num_branches * m->stage_channels[i] * 2;
      x0w.pitch16 = widebuf[i].width;
      emit(li++, rows, cur_in, i, -1, -1, 1.f, x0w);
    } else {
      emit(li++, rows, cur_in, i, -1, -1, 1.f, x0);
    }
    rows *= c.upsample_rates[i];
    Act stage_out;   // x = xs / num_kernels (models.py:141), as the next consumer wants it
    stage_out.slope = last_stage ? kLreluSlopePost : kLreluSlope;
    stage_out.f32 = (!bf || last_stage) ? sum_f32 : -1;
    stage_out.bf16 = (bf && !last_stage) ? next_bf16 : -1;
    Act partial;     // running MRF sum xs (models.py:136-140)
    partial.f32 = sum_f32;
    if (wide) {
      const v2w_wide_desc& W = m->wide[i];
      const int C = W.channels, nb = W.num_branches;
      for (int half = 0; half < 2; ++half) {
        Op o = blank(half == 0 ? V2W_OP_MRF_WIDE1 : V2W_OP_MRF_WIDE2);
        o.d.engine = V2W_ENGINE_TC;
        o.d.layer = li;
        o.d.stage = i;
        o.d.batch = B;
        o.d.rows = rows;
        o.d.n_tile = C;
        o.d.in_is_bf16 = 1;
        o.d.in_pitch = widebuf[i].width;
        o.d.residual = x0.f32;
        if (half == 0) {
          o.d.n_tiles = nb;
          o.d.in_act = widebuf[i].a1cat + (int64_t)nb * C * 2;     // A0
          o.d.out_bf16 = widebuf[i].a1cat;                        // A1_0 | ... | A1_{nb-1}
          o.d.out_bf16_pitch = widebuf[i].width;
          o.d.out_slope = kLreluSlope;
        } else {
          // one column tile, unless that leaves most SM pairs idle (cfg1: 2 pair tiles, cfg4: 30 of 74): then two
          // column tiles of C / 2 -- every tile walks the whole K = (nk + 1) C, so halving N halves its time
          const int64_t pair_tiles = (int64_t)B * ((rows + 255) / 256);
          o.d.n_tiles = (C >= 256 && pair_tiles <= 148 / 4) ? 2 : 1;
          o.d.n_tile = C / o.d.n_tiles;
          o.d.in_act = widebuf[i].a1cat;
          o.d.out_scale = 1.0f / nk;
          o.d.out_f32 = stage_out.f32; o.d.out_bf16 = stage_out.bf16; o.d.out_slope = stage_out.slope;
        }
        for (int q = 0; q < nb; ++q)
          o.d.flops += 2 * m->layers[li + 2 * q + half].macs_per_row * (int64_t)rows * B;
        P->ops.push_back(o);
      }
      li += 2 * nk;
      cur_in = stage_out;
      continue;
    }
    if (bf && !force_simt && !no_fuse && m->fused[i].ok && ((fuse_mask >> i) & 1)) {
      // whole stage in one kernel (kernels_fused.cuh)
      Op o = blank(V2W_OP_MRF_FUSED);
      o.d.engine = V2W_ENGINE_TC;
      o.d.layer = li;
      o.d.stage = i;
      o.d.batch = B;
      o.d.rows = rows;
      o.d.n_tile = m->stage_channels[i];
      o.d.n_tiles = 1;
      o.d.in_act = x0.bf16; o.d.in_is_bf16 = 1;
      o.d.residual = x0.f32;
      o.d.f32_blocked = x0_blocked ? 1 : 0;
      o.d.out_scale = 1.0f / nk;
      o.d.out_f32 = stage_out.f32; o.d.out_bf16 = stage_out.bf16; o.d.out_slope = stage_out.slope;
      for (int q = 0; q < 2 * nk; ++q)
        o.d.flops += 2 * m->layers[li + q].macs_per_row * (int64_t)rows * B;
      if (last_stage && m->fused[i].post) {
        // conv_post + tanh inside the stage's final pass: the waveform is this op's output
        o.d.post = 1;
        o.d.out_f32 = o.d.out_bf16 = -1;
        o.d.flops += 2 * m->layers[li + 2 * nk].macs_per_row * (int64_t)rows * B;
        post_folded = true;
      }
      P->ops.push_back(o);
      li += 2 * nk;
      cur_in = stage_out;
      continue;
    }
    for (int j = 0; j < nk; ++j) {
      const bool last_branch = j == nk - 1;
      const int64_t accum = j > 0 ? sum_f32 : -1;
      const float scale = last_branch ? 1.0f / nk : 1.0f;
      const Act& tail_out = last_branch ? stage_out : partial;
      if (c.resblock_type == 2) {
        // ResBlock2 (models.py:65-70): x = c(lrelu(x)) + x, twice
        emit(li++, rows, x0, -1, x0.f32, -1, 1.f, y1);
        emit(li++, rows, y1, -1, y1.f32, accum, scale, tail_out);
      } else {
        // ResBlock1 (models.py:37-44): x = c2(lrelu(c1(lrelu(x)))) + x, three times
        Act cur = x0;
        // The fp32 stream between the conv pairs of a block and the running MRF sum between blocks stay in the
        // row-blocked layout (conv-pair kernels, C <= 128): their final passes store / load row-per-thread, which
        // in the row-major layout is one 16-byte piece of 32 different lines per instruction.  Not the last stage's
        // sum, which is conv_post's row-major input.  $V2W_RB1_NO_BLOCKED=1: row-major everywhere.
        // cfg2, fp16 operands: 3.85 -> 3.28 ms per forward (profiles/r02_notes.md section 10).
        const char* nb_env = std::getenv("V2W_RB1_NO_BLOCKED");      // (read per plan, like $V2W_BLOCKED_X0_MASK)
        const bool rb1_no_blocked = nb_env && std::atoi(nb_env);
        const bool rb1_blk = !rb1_no_blocked && m->stage_channels[i] <= 128;
        bool cur_blocked = x0_blocked;
        for (int s = 0; s < 3; ++s) {
          const int pi = (bf && !force_simt && !no_fuse && ((fuse_mask >> i) & 1)) ? m->pair_of_layer[li] : -1;
          if (pi >= 0) {
            // conv pair in one kernel (kernels_stacked.cuh, pair program): the intermediate t1 never
            // leaves the SM
            const bool tail = s == 2;
            const Act& dst = tail ? tail_out : (s == 0 ? y1 : y2);
            Op o = blank(V2W_OP_PAIR_FUSED);
            o.d.engine = V2W_ENGINE_TC;
            o.d.layer = li;
            o.d.stage = i;
            o.d.pair = pi;
            o.d.batch = B;
            o.d.rows = rows;
            o.d.n_tile = m->stage_channels[i];
            o.d.n_tiles = 1;
            o.d.in_act = cur.bf16; o.d.in_is_bf16 = 1;
            o.d.residual = cur.f32;
            o.d.accum_in = tail ? accum : -1;
            o.d.out_scale = tail ? scale : 1.f;
            o.d.out_f32 = dst.f32; o.d.out_bf16 = dst.bf16; o.d.out_slope = dst.slope;
            o.d.f32_blocked = cur_blocked ? 1 : 0;
            o.d.out_blocked = (rb1_blk && dst.f32 >= 0 && !(tail && last_stage)) ? 1 : 0;
            o.d.acc_blocked = (rb1_blk && o.d.accum_in >= 0 && !last_stage) ? 1 : 0;
            cur_blocked = o.d.out_blocked != 0;
            o.d.flops = 2 * (m->layers[li].macs_per_row + m->layers[li + 1].macs_per_row) * (int64_t)rows * B;
            P->ops.push_back(o);
            li += 2;
            if (!tail) cur = dst;
            continue;
          }
          emit(li++, rows, cur, -1, -1, -1, 1.f, t1);
          if (s < 2) {
            emit(li++, rows, t1, -1, cur.f32, -1, 1.f, y1);
            cur = y1;
          } else {
            emit(li++, rows, t1, -1, cur.f32, accum, scale, tail_out);
          }
        }
      }
    }
    cur_in = stage_out;
  }
  if (!post_folded) {
    Op o = blank(V2W_OP_CONV_POST);
    o.d.layer = li;
    o.d.batch = B; o.d.rows = rows;
    o.d.in_act = sum_f32;
    o.d.in_slope = kLreluSlopePost;
    o.d.flops = 2 * m->layers[li].macs_per_row * (int64_t)rows * B;
    P->ops.push_back(o);
  }
  *out = P;
  return V2W_OK;
}

void v2w_plan_destroy(v2w_plan* p) { delete p; }

int64_t v2w_plan_workspace_bytes(const v2w_plan* p) { return p ? p->workspace_bytes : 0; }
int32_t v2w_plan_packed_format(const v2w_plan* p) { return p ? p->packed_format : V2W_HALF_BF16; }
int32_t v2w_plan_num_ops(const v2w_plan* p) { return p ? (int32_t)p->ops.size() : 0; }

int32_t v2w_plan_get_op(const v2w_plan* p, int32_t i, v2w_op_desc* out) {
  if (!p || !out || i < 0 || i >= (int32_t)p->ops.size()) {
    set_error("v2w_plan_get_op: bad index %d", i);
    return V2W_ERR_INVALID;
  }
  *out = p->ops[i].d;
  return V2W_OK;
}

int32_t v2w_plan_get_launch(const v2w_plan* p, int32_t i, v2w_launch_desc* out) {
  if (!p || !out || i < 0 || i >= (int32_t)p->ops.size()) {
    set_error("v2w_plan_get_launch: bad index %d", i);
    return V2W_ERR_INVALID;
  }
  if (!p->bound) {
    set_error("v2w_plan_get_launch: plan is not bound (launch geometry is decided in v2w_plan_bind)");
    return V2W_ERR_STATE;
  }
  std::memset(out, 0, sizeof(*out));
  const Op& o = p->ops[(size_t)i];
  if (o.d.kind == V2W_OP_MRF_FUSED || o.d.kind == V2W_OP_PAIR_FUSED) {
    out->grid = o.fl.grid; out->smem_bytes = o.fl.smem; out->cta_pair = o.fl.pair ? 1 : 0;
    return V2W_OK;
  }
  if (!o.tc_ready) return V2W_OK;     // not a launch of the generic tensor-core kernel
  const TcLaunch& L = o.tcl;
  out->grid = L.grid; out->smem_bytes = L.smem; out->cta_pair = L.pair ? 1 : 0; out->k_chunk = L.kc;
  out->msub = L.prm.msub; out->stages_a = L.prm.stages_a; out->stages_w = L.prm.stages_w;
  out->tap_group = L.prm.tap_group; out->y_cols = L.prm.y_cols; out->h_cols = L.prm.h_cols;
  out->epi_mode = L.prm.epi_mode; out->epi_wide = L.prm.epi_wide; out->epi_phased = L.prm.epi_phased;
  out->split = L.prm.split;
  return V2W_OK;
}

int32_t v2w_plan_set_engine(v2w_plan* p, int32_t i, int32_t engine) {
  if (!p || i < 0 || i >= (int32_t)p->ops.size() ||
      (engine != V2W_ENGINE_SIMT && engine != V2W_ENGINE_TC)) {
    set_error("v2w_plan_set_engine: bad argument");
    return V2W_ERR_INVALID;
  }
  if (p->bound) {
    set_error("v2w_plan_set_engine: plan already bound");
    return V2W_ERR_STATE;
  }
  Op& o = p->ops[i];
  if (o.d.kind != V2W_OP_TAPGEMM) {
    set_error("v2w_plan_set_engine: op %d is not a tap-GEMM", i);
    return V2W_ERR_INVALID;
  }
  if (engine == V2W_ENGINE_TC && p->precision == V2W_PREC_FP32) {
    set_error("v2w_plan_set_engine: the tensor-core engine needs a BF16/FP16 plan");
    return V2W_ERR_INVALID;
  }
  if (engine == V2W_ENGINE_SIMT && (o.d.out_bf16_pitch || o.d.in_pitch || o.d.f32_blocked)) {
    set_error("v2w_plan_set_engine: op %d writes into a wide stage's operand tensor (tensor-core engine only)", i);
    return V2W_ERR_INVALID;
  }
  if (engine == V2W_ENGINE_SIMT && (p->precision == V2W_PREC_FP16 || p->split_operands)) {
    set_error("v2w_plan_set_engine: the FFMA engine reads fp32 or bf16 activations, not fp16");
    return V2W_ERR_INVALID;
  }
  o.d.engine = engine;
  // (the FFMA engine reads the fp32 weights: nothing to split)
  if (p->packed_format == V2W_HALF_BF16_SPLIT && !p->split_operands) o.d.split = engine == V2W_ENGINE_TC ? 2 : 0;
  return V2W_OK;
}

// The two launches of a wide stage (v2w_wide_desc) as tap-GEMMs over the packed arrays that already
// exist: launch 1 reads the conv1 weights of every branch in place (column tile t = branch t), launch 2
// the conv2 weights (K-group j = branch j) and the `sum` convolution (last K-group).  All of them are
// [taps][C][C] arrays, so one weight tensor map with rows of C elements covers them; a block's row is its
// byte distance from the lowest array divided by the row size (every section is 1024-byte aligned).
static int prepare_wide(v2w_plan* p, Op& o, int tc_flags) {
  const v2w_model* m = p->model;
  const v2w_wide_desc& W = m->wide[o.d.stage];
  const int C = W.channels, nb = W.num_branches;
  const bool first = o.d.kind == V2W_OP_MRF_WIDE1;
  v2w_tapgemm_args& g = o.g;
  std::memset(&g, 0, sizeof(g));
  TcExtra ex;
  std::memset(ex.w_row, 0, sizeof(ex.w_row));
  g.engine = V2W_ENGINE_TC;
  g.batch = o.d.batch; g.rows = o.d.rows;
  g.in_is_bf16 = 1; g.in_slope = 1.f;
  g.tc_flags = tc_flags;
  auto at = [&](int64_t off) -> uint8_t* { return off < 0 ? nullptr : p->ws + off; };
  g.in_act = at(o.d.in_act);
  g.residual = reinterpret_cast<const float*>(at(o.d.residual));
  g.out_scale = o.d.out_scale;
  g.out_f32 = reinterpret_cast<float*>(at(o.d.out_f32));
  g.out_bf16 = at(o.d.out_bf16);
  g.out_slope = o.d.out_slope;
  ex.a_pitch = o.d.in_pitch;
  ex.h_pitch = o.d.out_bf16_pitch;
  ex.res_pitch = C; ex.res_cols = C;
  ex.w_cols = C;
  ex.abs_rows = true;
  const int64_t row_bytes = (int64_t)C * 2;
  // lowest / highest byte of the weight arrays this launch reads
  int64_t lo = -1, hi = -1;
  auto span = [&](int64_t off, int64_t bytes) {
    lo = lo < 0 ? off : std::min(lo, off);
    hi = std::max(hi, off + bytes);
  };
  for (int j = 0; j < nb; ++j) {
    const v2w_layer_desc& L = m->layers[W.first_layer + 2 * j + (first ? 0 : 1)];
    span(L.w16_off, (int64_t)L.num_taps * C * row_bytes);
  }
  if (!first) span(W.wsum_off, (int64_t)W.sum_taps * C * row_bytes);
  if (lo % row_bytes) {
    set_error("wide stage %d: weight arrays are not aligned to their row size", o.d.stage);
    return V2W_ERR_STATE;
  }
  g.w16 = p->packed + lo;
  ex.w_total_rows = (hi - lo) / row_bytes;
  auto rows_of = [&](int64_t off) { return (int)((off - lo) / row_bytes); };
  if (first) {
    g.cin = C; g.n_total = nb * C; g.cmod = nb * C; g.n_tile = C; g.n_tiles = nb;
    g.bias = reinterpret_cast<const float*>(p->packed + W.b1cat_off);
    ex.res_scale = 1.f;
    for (int t = 0; t < nb; ++t) {
      const v2w_layer_desc& L = m->layers[W.first_layer + 2 * t];
      g.num_taps[t] = L.num_taps;
      for (int j = 0; j < L.num_taps; ++j) {
        g.tap_off[t][j] = L.tap_off[j];
        g.tap_row[t][j] = j;
        ex.w_row[t][j] = rows_of(L.w16_off) + j * C;
      }
    }
  } else {
    g.cin = (nb + 1) * C; g.n_total = C; g.cmod = C; g.n_tile = o.d.n_tile; g.n_tiles = o.d.n_tiles;
    g.bias = reinterpret_cast<const float*>(p->packed + W.bsum_off);
    ex.res_scale = (float)nb;
    ex.kgroups = nb + 1;
    ex.cpg = C / 64;
    for (int t = 0; t < nb; ++t) {
      const v2w_layer_desc& L = m->layers[W.first_layer + 2 * t + 1];
      g.num_taps[t] = L.num_taps;
      for (int j = 0; j < L.num_taps; ++j) {
        g.tap_off[t][j] = L.tap_off[j];
        g.tap_row[t][j] = j;
        ex.w_row[t][j] = rows_of(L.w16_off) + j * C;
      }
    }
    g.num_taps[nb] = W.sum_taps;
    for (int j = 0; j < W.sum_taps; ++j) {
      g.tap_off[nb][j] = W.sum_off[j];
      g.tap_row[nb][j] = j;
      ex.w_row[nb][j] = rows_of(W.wsum_off) + j * C;
    }
  }
  int rc = prepare_tc(g, &o.tcl, &ex);
  if (rc) return rc;
  o.tc_ready = true;
  o.has_raw = false;
  return V2W_OK;
}

int32_t v2w_plan_bind(v2w_plan* p, const void* packed_dev, void* workspace_dev) {
  if (!p || !packed_dev || !workspace_dev) {
    set_error("v2w_plan_bind: null argument");
    return V2W_ERR_INVALID;
  }
  p->drop_graphs();   // captured launches hold the old buffers' addresses
  p->ran_once = false;
  if ((reinterpret_cast<uintptr_t>(packed_dev) & 255) ||
      (reinterpret_cast<uintptr_t>(workspace_dev) & 255)) {
    set_error("v2w_plan_bind: buffers must be 256-byte aligned");
    return V2W_ERR_INVALID;
  }
  int sms = 0;
  int rc = sm_count(&sms);
  if (rc) return rc;
  p->packed = static_cast<const uint8_t*>(packed_dev);
  p->ws = static_cast<uint8_t*>(workspace_dev);
  const v2w_model* m = p->model;
  {
    uint32_t hdr[3] = {0, 0, 0};
    V2W_CUDA(cudaMemcpy(hdr, packed_dev, sizeof(hdr), cudaMemcpyDeviceToHost));
    if (hdr[0] != kPackedMagic || hdr[1] != V2W_ABI_VERSION) {
      set_error("v2w_plan_bind: packed_dev is not a v2w_model_pack buffer of this ABI");
      return V2W_ERR_INVALID;
    }
    static const char* const kFormatName[4] = {"bf16", "fp16", "split-fp16", "split-bf16"};
    const int want = p->packed_format;
    if (p->precision != V2W_PREC_FP32 && (int)hdr[2] != want) {
      set_error("v2w_plan_bind: packed weights hold %s operands but the plan computes in %s",
                hdr[2] < 4 ? kFormatName[hdr[2]] : "unknown", kFormatName[want]);
      return V2W_ERR_INVALID;
    }
  }
  const bool split = p->split_operands;
  const int tc_flags = (p->precision == V2W_PREC_FP16 || p->precision == V2W_PREC_FP32_TC) ? 0x100 : 0;
  for (Op& o : p->ops) {
    if (o.d.kind == V2W_OP_MRF_FUSED || o.d.kind == V2W_OP_PAIR_FUSED) {
      const v2w_fused_desc& F = o.d.kind == V2W_OP_MRF_FUSED ? m->fused[o.d.stage] : m->pairs[o.d.pair];
      rc = prepare_fused(m, F, o.d, p->packed, p->ws, p->precision == V2W_PREC_FP16, &o.fl);
      if (rc) return rc;
      continue;
    }
    if (o.d.kind == V2W_OP_MRF_WIDE1 || o.d.kind == V2W_OP_MRF_WIDE2) {
      rc = prepare_wide(p, o, tc_flags);
      if (rc) return rc;
      continue;
    }
    if (o.d.kind != V2W_OP_TAPGEMM) continue;
    const v2w_layer_desc& L = m->layers[o.d.layer];
    v2w_tapgemm_args& g = o.g;
    std::memset(&g, 0, sizeof(g));
    g.engine = o.d.engine;
    g.batch = o.d.batch; g.rows = o.d.rows; g.cin = L.cin; g.n_total = L.n_total; g.cmod = L.cout;
    g.n_tile = o.d.n_tile; g.n_tiles = o.d.n_tiles;
    for (int t = 0; t < o.d.n_tiles; ++t) {
      g.num_taps[t] = o.d.num_taps[t];
      for (int j = 0; j < o.d.num_taps[t]; ++j) {
        g.tap_row[t][j] = o.d.tap_index[t][j];
        g.tap_off[t][j] = L.tap_off[o.d.tap_index[t][j]];
      }
    }
    auto at = [&](int64_t off) -> uint8_t* { return off < 0 ? nullptr : p->ws + off; };
    g.in_act = at(o.d.in_act);
    g.in_is_bf16 = o.d.in_is_bf16;
    g.in_slope = o.d.in_slope;
    g.w16 = L.w16_off >= 0 ? p->packed + L.w16_off : nullptr;
    g.w32 = reinterpret_cast<const float*>(p->packed + L.w32_off);
    g.bias = reinterpret_cast<const float*>(p->packed + L.bias_off);
    if (o.d.cond_stage >= 0) {
      g.scale = reinterpret_cast<const float*>(at(p->cond_scale[o.d.cond_stage]));
      g.shift = reinterpret_cast<const float*>(at(p->cond_shift[o.d.cond_stage]));
    }
    g.residual = reinterpret_cast<const float*>(at(o.d.residual));
    g.accum_in = reinterpret_cast<const float*>(at(o.d.accum_in));
    g.out_scale = o.d.out_scale;
    g.out_f32 = reinterpret_cast<float*>(at(o.d.out_f32));
    g.out_bf16 = at(o.d.out_bf16);
    g.out_slope = o.d.out_slope;
    g.tc_flags = tc_flags;
    rc = validate_gemm(g);
    if (rc) return rc;
    o.tc_ready = false;
    if (g.engine == V2W_ENGINE_TC) {
      TcExtra ex;
      ex.h_pitch = o.d.out_bf16_pitch;
      ex.h_cout = L.cout;       // a transposed conv's GEMM row is `up` output rows of cout channels
      ex.y_blocked_cout = o.d.f32_blocked ? L.cout : 0;
      if (split) {
        ex.split = 1;
        ex.a_pitch = 2 * L.cin;
        ex.h_pitch = 2 * L.cout;
        ex.w_lo_row = L.num_taps * L.n_total;
      } else if (o.d.split == 2) {
        ex.split = 2;
        ex.w_lo_row = L.num_taps * L.n_total;
      }
      rc = prepare_tc(g, &o.tcl, (o.d.out_bf16_pitch || o.d.f32_blocked || o.d.split) ? &ex : nullptr);
      if (rc) return rc;
      o.tc_ready = true;
    } else if (o.d.out_bf16_pitch || o.d.f32_blocked) {
      set_error("v2w_plan_bind: op writes a pitched / row-blocked output, which needs the tensor-core engine");
      return V2W_ERR_STATE;
    }
    o.has_raw = false;
    if (o.d.cond_stage >= 0) {
      o.g_raw = g;
      o.g_raw.scale = o.g_raw.shift = nullptr;
      o.g_raw.out_f32 = reinterpret_cast<float*>(at(p->tr_y));
      o.g_raw.out_bf16 = nullptr;
      o.g_raw.out_slope = 1.f;
      rc = validate_gemm(o.g_raw);
      if (rc) return rc;
      if (g.engine == V2W_ENGINE_TC) {
        TcExtra exr;
        exr.split = o.d.split;
        exr.a_pitch = split ? 2 * L.cin : 0;
        exr.h_cout = L.cout;
        exr.w_lo_row = L.num_taps * L.n_total;
        rc = prepare_tc(o.g_raw, &o.tcl_raw, o.d.split ? &exr : nullptr);
        if (rc) return rc;
      }
      o.has_raw = true;
    }
  }
  {
    // conv_post's 7 x C weights travel as a by-value kernel parameter (constant bank operands)
    const v2w_layer_desc& L = m->layers.back();
    p->post_w.assign((size_t)L.num_taps * L.cin, 0.f);
    V2W_CUDA(cudaMemcpy(p->post_w.data(), p->packed + L.w32_off, p->post_w.size() * 4,
                        cudaMemcpyDeviceToHost));
    V2W_CUDA(cudaMemcpy(&p->post_bias, p->packed + L.bias_off, 4, cudaMemcpyDeviceToHost));
  }
  {
    int cmax = 0;
    for (int c : m->stage_channels) cmax = std::max(cmax, c);
    std::vector<float> ones((size_t)cmax, 1.f);
    V2W_CUDA(cudaMemcpy(p->ws + p->tr_ones, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
    V2W_CUDA(cudaMemset(p->ws + p->tr_zeros, 0, (size_t)cmax * 4));
  }
  p->bound = true;
  return V2W_OK;
}

int32_t v2w_plan_refresh_weights(v2w_plan* p, void* stream) {
  if (!p || !p->bound) {
    set_error("v2w_plan_refresh_weights: plan is not bound");
    return V2W_ERR_STATE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const v2w_layer_desc& L = p->model->layers.back();
  p->post_w.assign((size_t)L.num_taps * L.cin, 0.f);
  V2W_CUDA(cudaMemcpyAsync(p->post_w.data(), p->packed + L.w32_off, p->post_w.size() * 4,
                           cudaMemcpyDeviceToHost, st));
  V2W_CUDA(cudaMemcpyAsync(&p->post_bias, p->packed + L.bias_off, 4, cudaMemcpyDeviceToHost, st));
  V2W_CUDA(cudaStreamSynchronize(st));
  p->drop_graphs();
  return V2W_OK;
}

// train() mode, conditioning that does not depend on the activations (vec2wav/modules.py:16,24 with
// module.training set): the power iteration of every stage and gamma / beta = SN-Linear(fcs(spk|noise)),
// on the side stream while conv_pre runs.  cond_affine_kernel runs with rstd = 1, mean = 0, so its
// scale / shift outputs hold gamma / beta until bn_finalize_kernel folds the batch statistics in.
static int32_t train_conditioning(v2w_plan* p, const v2w_train_state& tr, const float* spk_dev,
                                  const float* noise_dev, cudaStream_t cs, int* launches) {
  const v2w_model* m = p->model;
  const v2w_config& c = m->cfg;
  simt::SnTrainParams sp;
  std::memset(&sp, 0, sizeof(sp));
  simt::CondParams cp;
  std::memset(&cp, 0, sizeof(cp));
  int cmax = 0;
  for (int i = 0; i < c.num_upsamples; ++i) {
    const v2w_train_stage& T = tr.stage[i];
    const int C = m->stage_channels[i];
    if (C < 4 || C > 512 || (C & (C - 1))) {
      set_error("v2w_forward_train: %d channels unsupported by the statistics kernels", C);
      return V2W_ERR_INVALID;
    }
    cmax = std::max(cmax, C);
    sp.st[i].W = T.sn_weight_orig; sp.st[i].u = T.sn_u; sp.st[i].v = T.sn_v;
    sp.st[i].inv_sigma = reinterpret_cast<float*>(p->ws + p->tr_isg[i]);
    sp.st[i].C = C;
    simt::CondStage& S = cp.st[i];
    S.fcw = T.fc_weight; S.fcb = T.fc_bias; S.snw = T.sn_weight_orig; S.snb = T.sn_bias;
    S.rstd = reinterpret_cast<const float*>(p->ws + p->tr_ones);
    S.mean = reinterpret_cast<const float*>(p->ws + p->tr_zeros);
    S.inv_sigma = sp.st[i].inv_sigma;
    S.scale = reinterpret_cast<float*>(p->ws + p->cond_scale[i]);
    S.shift = reinterpret_cast<float*>(p->ws + p->cond_shift[i]);
    S.C = C;
  }
  simt::sn_train_kernel<<<c.num_upsamples, simt::kSnThreads, 0, cs>>>(sp);
  V2W_CUDA(cudaGetLastError());
  cp.spk = spk_dev; cp.noise = noise_dev;
  cp.spk_dim = c.spk_dim; cp.noise_dim = c.noise_dim;
  const int smem = (c.spk_dim + c.noise_dim + 128 + 2 * cmax) * 4;
  simt::cond_affine_kernel<<<dim3((unsigned)c.num_upsamples, p->batch), simt::kCondThreads, smem, cs>>>(cp);
  V2W_CUDA(cudaGetLastError());
  *launches += 2;
  return V2W_OK;
}

// train() mode, one upsample stage (vec2wav/modules.py:20-30 with module.training set): raw transposed
// conv -> per-channel batch statistics -> scale/shift of this batch (+ running-stat update).  The
// plan's own op (same GEMM, affine epilogue) follows.
static int32_t train_stage_affine(v2w_plan* p, Op& o, const v2w_train_state& tr, cudaStream_t st, int* launches) {
  const v2w_model* m = p->model;
  const int i = o.d.cond_stage;
  const v2w_train_stage& T = tr.stage[i];
  const int C = m->stage_channels[i];
  if (!o.has_raw) {
    set_error("v2w_forward_train: stage %d has no raw launch", i);
    return V2W_ERR_STATE;
  }
  int rc = o.g_raw.engine == V2W_ENGINE_TC ? launch_tc(o.tcl_raw, st) : launch_simt(o.g_raw, st);
  if (rc) return rc;
  double* sums = reinterpret_cast<double*>(p->ws + p->tr_sums[i]);
  V2W_CUDA(cudaMemsetAsync(sums, 0, (size_t)2 * C * 8, st));
  const long long R = (long long)p->batch * o.d.rows * m->layers[o.d.layer].up;   // rows of [R][C]
  const int groups = simt::kStatThreads / (C / 4);
  const long long per_cta = (long long)groups * simt::kStatRun;
  int sms = 148;
  sm_count(&sms);
  const int grid = (int)std::max<long long>(1, std::min<long long>((R + per_cta - 1) / per_cta, (long long)sms * 8));
  simt::bn_stats_kernel<<<grid, simt::kStatThreads, 0, st>>>(reinterpret_cast<const float*>(p->ws + p->tr_y), R, C, sums);
  V2W_CUDA(cudaGetLastError());
  simt::BnFinalizeParams fp;
  fp.sums = sums;
  fp.scale = reinterpret_cast<float*>(p->ws + p->cond_scale[i]);
  fp.shift = reinterpret_cast<float*>(p->ws + p->cond_shift[i]);
  fp.run_mean = T.bn_running_mean; fp.run_var = T.bn_running_var;
  fp.C = C; fp.batch = p->batch;
  fp.count = R;
  fp.momentum = tr.momentum;
  fp.eps = kBnEps;
  simt::bn_finalize_kernel<<<(p->batch * C + 255) / 256, 256, 0, st>>>(fp);
  V2W_CUDA(cudaGetLastError());
  *launches += 3;
  return V2W_OK;
}

// One pass over the plan's op list.  `ragged`: x_dev is packed time-major rows and the waveform is
// written un-padded (v2w_forward_ragged); the row offsets are already in the workspace.
static int32_t run_ops(v2w_plan* p, const float* x_dev, const float* spk_dev, const float* noise_dev,
                       void* out_dev, const simt::RaggedOffsets* ro, int32_t out_format, void* stream,
                       int32_t* gpu_launches, const v2w_train_state* tr = nullptr) {
  const bool ragged = ro != nullptr;
  static const simt::RaggedOffsets no_offsets = {};
  // variable-length batches: every row-tiled kernel skips the tiles no utterance needs
  const int* lim = ragged ? reinterpret_cast<const int*>(p->ws + p->row_limit) : nullptr;
  if (!p || !x_dev || !spk_dev || !out_dev) {
    set_error("v2w_forward: null argument");
    return V2W_ERR_INVALID;
  }
  if (!p->bound) {
    set_error("v2w_forward: plan is not bound to device buffers");
    return V2W_ERR_STATE;
  }
  const v2w_model* m = p->model;
  const v2w_config& c = m->cfg;
  if (c.noise_dim > 0 && !noise_dev) {
    set_error("v2w_forward: noise is required (vec2wav/models.py:120)");
    return V2W_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf = p->precision != V2W_PREC_FP32;
  int launches = 0;
  size_t op_index = 0;
  static const bool no_fork = [] { const char* e = std::getenv("V2W_NO_FORK"); return e && std::atoi(e); }();
  const bool fork = !p->profiling && !no_fork && !tr;
  bool joined = true;
  if (fork && !p->side) {
    V2W_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
    V2W_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    V2W_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
  }
  for (Op& o : p->ops) {
    if (p->profiling) V2W_CUDA(cudaEventRecord(p->events[op_index], st));
    ++op_index;
    if (!joined && o.d.kind == V2W_OP_TAPGEMM && o.d.cond_stage >= 0) {
      V2W_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));
      joined = true;
    }
    switch (o.d.kind) {
      case V2W_OP_TRANSPOSE_IN: {
        if (ragged) {
          dim3 grid((p->frames + 7) / 8, p->batch);
          if (p->split_operands && p->precision == V2W_PREC_BF16)
            simt::gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                x_dev, *ro, reinterpret_cast<__nv_bfloat16*>(p->ws + p->xt), c.num_wv_feat, p->frames,
                reinterpret_cast<int*>(p->ws + p->row_limit), v2w_model_halo_frames(m), c.num_wv_feat);
          else if (p->split_operands)
            simt::gather_rows_kernel<__half><<<grid, 256, 0, st>>>(
                x_dev, *ro, reinterpret_cast<__half*>(p->ws + p->xt), c.num_wv_feat, p->frames,
                reinterpret_cast<int*>(p->ws + p->row_limit), v2w_model_halo_frames(m), c.num_wv_feat);
          else if (p->precision == V2W_PREC_FP16)
            simt::gather_rows_kernel<__half><<<grid, 256, 0, st>>>(
                x_dev, *ro, reinterpret_cast<__half*>(p->ws + p->xt), c.num_wv_feat, p->frames,
                reinterpret_cast<int*>(p->ws + p->row_limit), v2w_model_halo_frames(m));
          else if (bf)
            simt::gather_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
                x_dev, *ro, reinterpret_cast<__nv_bfloat16*>(p->ws + p->xt), c.num_wv_feat, p->frames,
                reinterpret_cast<int*>(p->ws + p->row_limit), v2w_model_halo_frames(m));
          else
            simt::gather_rows_kernel<float><<<grid, 256, 0, st>>>(
                x_dev, *ro, reinterpret_cast<float*>(p->ws + p->xt), c.num_wv_feat, p->frames,
                reinterpret_cast<int*>(p->ws + p->row_limit), v2w_model_halo_frames(m));
          V2W_CUDA(cudaGetLastError());
          ++launches;
          break;
        }
        dim3 grid((p->frames + 31) / 32, (c.num_wv_feat + 31) / 32, p->batch);
        if (p->split_operands && p->precision == V2W_PREC_BF16)
          simt::transpose_in_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
              x_dev, reinterpret_cast<__nv_bfloat16*>(p->ws + p->xt), c.num_wv_feat, p->frames, c.num_wv_feat);
        else if (p->split_operands)
          simt::transpose_in_kernel<__half><<<grid, 256, 0, st>>>(
              x_dev, reinterpret_cast<__half*>(p->ws + p->xt), c.num_wv_feat, p->frames, c.num_wv_feat);
        else if (p->precision == V2W_PREC_FP16)
          simt::transpose_in_kernel<__half><<<grid, 256, 0, st>>>(
              x_dev, reinterpret_cast<__half*>(p->ws + p->xt), c.num_wv_feat, p->frames);
        else if (bf)
          simt::transpose_in_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
              x_dev, reinterpret_cast<__nv_bfloat16*>(p->ws + p->xt), c.num_wv_feat, p->frames);
        else
          simt::transpose_in_kernel<float><<<grid, 256, 0, st>>>(
              x_dev, reinterpret_cast<float*>(p->ws + p->xt), c.num_wv_feat, p->frames);
        V2W_CUDA(cudaGetLastError());
        ++launches;
        break;
      }
      case V2W_OP_COND_AFFINE: {
        if (tr) {
          // train(): power iterations and gamma / beta of every stage on the side stream; the batch
          // statistics are folded in stage by stage (bn_finalize_kernel)
          if (!p->side) {
            V2W_CUDA(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
            V2W_CUDA(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
            V2W_CUDA(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
          }
          V2W_CUDA(cudaEventRecord(p->ev_fork, st));
          V2W_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
          int rc = train_conditioning(p, *tr, spk_dev, noise_dev, p->side, &launches);
          if (rc) return rc;
          V2W_CUDA(cudaEventRecord(p->ev_join, p->side));
          joined = false;
          break;
        }
        simt::CondParams cp;
        std::memset(&cp, 0, sizeof(cp));
        int cmax = 0;
        for (size_t i = 0; i < m->cond.size(); ++i) {
          const CondLayout& cl = m->cond[i];
          simt::CondStage& S = cp.st[i];
          S.fcw = reinterpret_cast<const float*>(p->packed + cl.fcw);
          S.fcb = reinterpret_cast<const float*>(p->packed + cl.fcb);
          S.snw = reinterpret_cast<const float*>(p->packed + cl.snw);
          S.snb = reinterpret_cast<const float*>(p->packed + cl.snb);
          S.rstd = reinterpret_cast<const float*>(p->packed + cl.rstd);
          S.mean = reinterpret_cast<const float*>(p->packed + cl.mean);
          S.scale = reinterpret_cast<float*>(p->ws + p->cond_scale[i]);
          S.shift = reinterpret_cast<float*>(p->ws + p->cond_shift[i]);
          S.C = cl.C;
          cmax = std::max(cmax, cl.C);
        }
        cp.spk = spk_dev; cp.noise = noise_dev;
        cp.spk_dim = c.spk_dim; cp.noise_dim = c.noise_dim;
        const int smem = (c.spk_dim + c.noise_dim + 128 + 2 * cmax) * 4;
        dim3 grid((unsigned)m->cond.size(), p->batch);
        cudaStream_t cs = st;
        if (fork) {
          // everything this stream did before (the previous forward read scale/shift) is ordered first
          V2W_CUDA(cudaEventRecord(p->ev_fork, st));
          V2W_CUDA(cudaStreamWaitEvent(p->side, p->ev_fork, 0));
          cs = p->side;
        }
        simt::cond_affine_kernel<<<grid, simt::kCondThreads, smem, cs>>>(cp);
        V2W_CUDA(cudaGetLastError());
        if (fork) {
          V2W_CUDA(cudaEventRecord(p->ev_join, p->side));
          joined = false;
        }
        ++launches;
        break;
      }
      case V2W_OP_TAPGEMM: {
        const int rpf = o.d.rows / p->frames;   // rows of this op per input frame
        if (tr && o.d.cond_stage >= 0) {
          int rc = train_stage_affine(p, o, *tr, st, &launches);
          if (rc) return rc;
        }
        int rc = o.g.engine == V2W_ENGINE_TC ? launch_tc(o.tcl, st, lim, rpf) : launch_simt(o.g, st, lim, rpf);
        if (rc) return rc;
        ++launches;
        break;
      }
      case V2W_OP_MRF_WIDE1:
      case V2W_OP_MRF_WIDE2: {
        int rc = launch_tc(o.tcl, st, lim, o.d.rows / p->frames);
        if (rc) return rc;
        ++launches;
        break;
      }
      case V2W_OP_PAIR_FUSED:
      case V2W_OP_MRF_FUSED: {
        PostFold pf;
        if (o.d.post) {
          pf.w = p->post_w.data();
          pf.bias = p->post_bias;
          pf.out.y = static_cast<float*>(out_dev);
          pf.out.ragged = ragged ? 1 : 0;
          pf.out.up = m->total_up;
          pf.out.pcm16 = out_format == V2W_OUT_PCM16;
          pf.ro = ro;
        }
        int rc = launch_fused(o.fl, st, lim, o.d.rows / p->frames, o.d.post ? &pf : nullptr);
        if (rc) return rc;
        ++launches;
        break;
      }
      case V2W_OP_CONV_POST: {
        const v2w_layer_desc& L = m->layers[o.d.layer];
        const float* in = reinterpret_cast<const float*>(p->ws + o.d.in_act);
        dim3 grid((o.d.rows + simt::kPostOut - 1) / simt::kPostOut, p->batch);
        simt::PostOut po;
        po.y = static_cast<float*>(out_dev);
        po.ragged = ragged ? 1 : 0;
        const simt::RaggedOffsets& offs = ragged ? *ro : no_offsets;
        po.up = m->total_up;
        po.pcm16 = out_format == V2W_OUT_PCM16;
        if (L.cin == 16) {
          simt::PostWeights<16> pw;
          std::memcpy(pw.w, p->post_w.data(), sizeof(pw.w));
          pw.bias = p->post_bias;
          simt::conv_post_kernel<16><<<grid, simt::kPostThreads, 0, st>>>(in, pw, po, offs, o.d.rows,
                                                                         o.d.in_slope);
        } else if (L.cin == 32) {
          simt::PostWeights<32> pw;
          std::memcpy(pw.w, p->post_w.data(), sizeof(pw.w));
          pw.bias = p->post_bias;
          simt::conv_post_kernel<32><<<grid, simt::kPostThreads, 0, st>>>(in, pw, po, offs, o.d.rows,
                                                                         o.d.in_slope);
        } else {
          set_error("conv_post with %d input channels unsupported", L.cin);
          return V2W_ERR_INVALID;
        }
        V2W_CUDA(cudaGetLastError());
        ++launches;
        break;
      }
      default:
        set_error("v2w_forward: unknown op kind %d", o.d.kind);
        return V2W_ERR_STATE;
    }
  }
  if (!joined) V2W_CUDA(cudaStreamWaitEvent(st, p->ev_join, 0));   // no op consumed it (not reached in practice)
  if (p->profiling) V2W_CUDA(cudaEventRecord(p->events[op_index], st));
  if (gpu_launches) *gpu_launches = launches;
  return V2W_OK;
}

int32_t v2w_forward(v2w_plan* p, const float* x_dev, const float* spk_dev, const float* noise_dev,
                    float* out_dev, void* stream, int32_t* gpu_launches) {
  int32_t rc = run_ops(p, x_dev, spk_dev, noise_dev, out_dev, nullptr, V2W_OUT_F32, stream, gpu_launches);
  if (rc == V2W_OK) p->ran_once = true;
  return rc;
}

int32_t v2w_forward_graph(v2w_plan* p, const float* x_dev, const float* spk_dev, const float* noise_dev,
                          float* out_dev, void* stream, int32_t* gpu_launches, int32_t* replayed) {
  constexpr size_t kMaxGraphs = 8;
  constexpr int kMissLimit = 16;
  if (replayed) *replayed = 0;
  if (!p) {
    set_error("v2w_forward_graph: null plan");
    return V2W_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->profiling)   // per-op events are recorded by the plain path only
    return v2w_forward(p, x_dev, spk_dev, noise_dev, out_dev, stream, gpu_launches);
  for (v2w_plan::GraphEntry& g : p->graphs) {
    if (g.x == x_dev && g.spk == spk_dev && g.noise == noise_dev && g.out == out_dev) {
      V2W_CUDA(cudaGraphLaunch(g.exec, st));
      g.stamp = ++p->graph_clock;
      p->graph_miss_streak = 0;
      if (gpu_launches) *gpu_launches = g.launches;
      if (replayed) *replayed = 1;
      return V2W_OK;
    }
  }
  // First run of a plan, profiling, or pointer tuples that never repeat: plain stream launches.
  if (!p->ran_once || p->profiling || p->graph_miss_streak >= kMissLimit)
    return v2w_forward(p, x_dev, spk_dev, noise_dev, out_dev, stream, gpu_launches);
  ++p->graph_miss_streak;
  // Captured on a stream of the plan's own (the caller's may be the legacy default stream, which cannot
  // be captured); the instantiated graph is then launched on the caller's stream.
  if (!p->cap) V2W_CUDA(cudaStreamCreateWithFlags(&p->cap, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  V2W_CUDA(cudaStreamBeginCapture(p->cap, cudaStreamCaptureModeThreadLocal));
  int32_t launches = 0;
  const int32_t rc = run_ops(p, x_dev, spk_dev, noise_dev, out_dev, nullptr, V2W_OUT_F32, p->cap, &launches);
  const cudaError_t ce = cudaStreamEndCapture(p->cap, &graph);   // always leave capture mode
  if (rc != V2W_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess || !graph) {
    cudaGetLastError();
    set_error("v2w_forward_graph: stream capture failed: %s", cudaGetErrorString(ce));
    return V2W_ERR_CUDA;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) {
    cudaGetLastError();
    set_error("v2w_forward_graph: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
    return V2W_ERR_CUDA;
  }
  if (p->graphs.size() >= kMaxGraphs) {
    size_t victim = 0;
    for (size_t i = 1; i < p->graphs.size(); ++i)
      if (p->graphs[i].stamp < p->graphs[victim].stamp) victim = i;
    cudaGraphExecDestroy(p->graphs[victim].exec);
    p->graphs.erase(p->graphs.begin() + victim);
  }
  p->graphs.push_back({x_dev, spk_dev, noise_dev, out_dev, exec, launches, ++p->graph_clock});
  V2W_CUDA(cudaGraphLaunch(exec, st));
  if (gpu_launches) *gpu_launches = launches;
  if (replayed) *replayed = 1;
  return V2W_OK;
}

int32_t v2w_forward_train(v2w_plan* p, const float* x_dev, const float* spk_dev, const float* noise_dev,
                          float* out_dev, const v2w_train_state* state, void* stream,
                          int32_t* gpu_launches) {
  if (!p || !state) {
    set_error("v2w_forward_train: null argument");
    return V2W_ERR_INVALID;
  }
  if (state->num_stages != p->model->cfg.num_upsamples) {
    set_error("v2w_forward_train: state has %d stages, the model %d", state->num_stages,
              p->model->cfg.num_upsamples);
    return V2W_ERR_INVALID;
  }
  if (!(state->momentum >= 0.f && state->momentum <= 1.f)) {
    set_error("v2w_forward_train: momentum %g out of [0, 1]", (double)state->momentum);
    return V2W_ERR_INVALID;
  }
  for (int i = 0; i < state->num_stages; ++i) {
    const v2w_train_stage& T = state->stage[i];
    if (!T.fc_weight || !T.fc_bias || !T.sn_weight_orig || !T.sn_bias || !T.sn_u || !T.sn_v ||
        !T.bn_running_mean || !T.bn_running_var) {
      set_error("v2w_forward_train: null pointer in stage %d", i);
      return V2W_ERR_INVALID;
    }
  }
  return run_ops(p, x_dev, spk_dev, noise_dev, out_dev, nullptr, V2W_OUT_F32, stream, gpu_launches, state);
}

int32_t v2w_forward_ragged(v2w_plan* p, const float* x_rows_dev, const int32_t* row_offsets_host,
                           const float* spk_dev, const float* noise_dev, void* out_packed_dev,
                           int32_t out_format, void* stream, int32_t* gpu_launches) {
  if (!p || !row_offsets_host) {
    set_error("v2w_forward_ragged: null argument");
    return V2W_ERR_INVALID;
  }
  if (!p->bound) {
    set_error("v2w_forward_ragged: plan is not bound to device buffers");
    return V2W_ERR_STATE;
  }
  if (out_format != V2W_OUT_F32 && out_format != V2W_OUT_PCM16) {
    set_error("v2w_forward_ragged: unknown output format %d", out_format);
    return V2W_ERR_INVALID;
  }
  if (p->batch > simt::kMaxRaggedBatch) {
    set_error("v2w_forward_ragged: at most %d utterances per call (plan has %d)", simt::kMaxRaggedBatch, p->batch);
    return V2W_ERR_INVALID;
  }
  if (row_offsets_host[0] != 0) {
    set_error("v2w_forward_ragged: row_offsets[0] must be 0");
    return V2W_ERR_INVALID;
  }
  for (int b = 0; b < p->batch; ++b) {
    const int64_t len = (int64_t)row_offsets_host[b + 1] - row_offsets_host[b];
    if (len < 1 || len > p->frames) {
      set_error("v2w_forward_ragged: utterance %d has %lld frames, the plan holds 1..%d", b,
                (long long)len, p->frames);
      return V2W_ERR_INVALID;
    }
  }
  // the offsets travel as a by-value kernel parameter: nothing to upload, nothing to wait for
  simt::RaggedOffsets ro;
  std::memset(&ro, 0, sizeof(ro));
  for (int b = 0; b <= p->batch; ++b) ro.off[b] = row_offsets_host[b];
  return run_ops(p, x_rows_dev, spk_dev, noise_dev, out_packed_dev, &ro, out_format, stream,
                 gpu_launches);
}

int32_t v2w_plan_set_profiling(v2w_plan* p, int32_t enabled) {
  if (!p) {
    set_error("v2w_plan_set_profiling: null plan");
    return V2W_ERR_INVALID;
  }
  if (enabled && p->events.empty()) {
    p->events.resize(p->ops.size() + 1);
    for (auto& e : p->events) V2W_CUDA(cudaEventCreate(&e));
  }
  p->profiling = enabled != 0;
  return V2W_OK;
}

float v2w_plan_op_time_ms(v2w_plan* p, int32_t i) {
  if (!p || p->events.empty() || i < 0 || i >= (int32_t)p->ops.size()) {
    set_error("v2w_plan_op_time_ms: profiling not enabled or bad index");
    return -1.f;
  }
  float ms = -1.f;
  if (cudaEventElapsedTime(&ms, p->events[i], p->events[i + 1]) != cudaSuccess) {
    set_error("v2w_plan_op_time_ms: events not complete");
    cudaGetLastError();
    return -1.f;
  }
  return ms;
}

int32_t v2w_forward_host(v2w_plan* p, const float* x_host, const float* spk_host,
                         const float* noise_host, float* out_host, void* stream,
                         int32_t* gpu_launches) {
  if (!p || !x_host || !spk_host || !out_host) {
    set_error("v2w_forward_host: null argument");
    return V2W_ERR_INVALID;
  }
  if (!p->bound) {
    set_error("v2w_forward_host: plan is not bound to device buffers");
    return V2W_ERR_STATE;
  }
  const v2w_config& c = p->model->cfg;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* dx = reinterpret_cast<float*>(p->ws + p->stage_x);
  float* ds = reinterpret_cast<float*>(p->ws + p->stage_spk);
  float* dn = reinterpret_cast<float*>(p->ws + p->stage_noise);
  float* dout = reinterpret_cast<float*>(p->ws + p->stage_out);
  const size_t B = p->batch;
  V2W_CUDA(cudaMemcpyAsync(dx, x_host, B * c.num_wv_feat * p->frames * 4, cudaMemcpyHostToDevice, st));
  V2W_CUDA(cudaMemcpyAsync(ds, spk_host, B * c.spk_dim * 4, cudaMemcpyHostToDevice, st));
  if (c.noise_dim > 0) {
    if (!noise_host) {
      set_error("v2w_forward_host: noise is required (vec2wav/models.py:120)");
      return V2W_ERR_INVALID;
    }
    V2W_CUDA(cudaMemcpyAsync(dn, noise_host, B * c.noise_dim * 4, cudaMemcpyHostToDevice, st));
  }
  int rc = v2w_forward(p, dx, ds, dn, dout, stream, gpu_launches);
  if (rc) return rc;
  V2W_CUDA(cudaMemcpyAsync(out_host, dout, B * p->frames * p->model->total_up * 4,
                           cudaMemcpyDeviceToHost, st));
  V2W_CUDA(cudaStreamSynchronize(st));
  return V2W_OK;
}

// ---------------------------------------------------------------------------------------------
// pipelined host API
// ---------------------------------------------------------------------------------------------
struct v2w_pipeline {
  v2w_plan* plan = nullptr;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  float *x[2] = {nullptr, nullptr}, *spk[2] = {nullptr, nullptr}, *noise[2] = {nullptr, nullptr};
  float* out[2] = {nullptr, nullptr};
  cudaEvent_t in_ready[2], run_done[2], out_done[2];
  bool events = false;
  int64_t step = 0, launches = 0;
  size_t x_bytes = 0, spk_bytes = 0, noise_bytes = 0, out_bytes = 0;
};

void v2w_pipeline_destroy(v2w_pipeline* q) {
  if (!q) return;
  if (q->s_run) cudaStreamSynchronize(q->s_run);
  if (q->s_out) cudaStreamSynchronize(q->s_out);
  for (int i = 0; i < 2; ++i) {
    cudaFree(q->x[i]); cudaFree(q->spk[i]); cudaFree(q->noise[i]); cudaFree(q->out[i]);
    if (q->events) {
      cudaEventDestroy(q->in_ready[i]); cudaEventDestroy(q->run_done[i]); cudaEventDestroy(q->out_done[i]);
    }
  }
  if (q->s_in) cudaStreamDestroy(q->s_in);
  if (q->s_run) cudaStreamDestroy(q->s_run);
  if (q->s_out) cudaStreamDestroy(q->s_out);
  delete q;
}

int32_t v2w_pipeline_create(v2w_plan* p, v2w_pipeline** out) {
  if (!p || !out) {
    set_error("v2w_pipeline_create: null argument");
    return V2W_ERR_INVALID;
  }
  *out = nullptr;
  if (!p->bound) {
    set_error("v2w_pipeline_create: plan is not bound to device buffers");
    return V2W_ERR_STATE;
  }
  v2w_pipeline* q = new (std::nothrow) v2w_pipeline();
  if (!q) {
    set_error("out of memory");
    return V2W_ERR_NOMEM;
  }
  q->plan = p;
  const v2w_config& c = p->model->cfg;
  const size_t B = p->batch;
  q->x_bytes = B * c.num_wv_feat * p->frames * 4;
  q->spk_bytes = B * c.spk_dim * 4;
  q->noise_bytes = B * (size_t)std::max(1, c.noise_dim) * 4;
  q->out_bytes = B * p->frames * p->model->total_up * 4;
  auto fail = [&](cudaError_t e, const char* what) {
    set_error("v2w_pipeline_create: %s failed: %s", what, cudaGetErrorString(e));
    v2w_pipeline_destroy(q);
    return V2W_ERR_CUDA;
  };
  cudaError_t e;
  if ((e = cudaStreamCreateWithFlags(&q->s_in, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
  if ((e = cudaStreamCreateWithFlags(&q->s_run, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
  if ((e = cudaStreamCreateWithFlags(&q->s_out, cudaStreamNonBlocking)) != cudaSuccess) return fail(e, "stream");
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaMalloc(&q->x[i], q->x_bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&q->spk[i], q->spk_bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&q->noise[i], q->noise_bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
    if ((e = cudaMalloc(&q->out[i], q->out_bytes)) != cudaSuccess) return fail(e, "cudaMalloc");
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventCreateWithFlags(&q->in_ready[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&q->run_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&q->out_done[i], cudaEventDisableTiming);
  }
  q->events = true;
  *out = q;
  return V2W_OK;
}

int32_t v2w_pipeline_submit(v2w_pipeline* q, const float* x_host, const float* spk_host,
                            const float* noise_host, float* out_host) {
  if (!q || !x_host || !spk_host || !out_host) {
    set_error("v2w_pipeline_submit: null argument");
    return V2W_ERR_INVALID;
  }
  const v2w_config& c = q->plan->model->cfg;
  if (c.noise_dim > 0 && !noise_host) {
    set_error("v2w_pipeline_submit: noise is required (vec2wav/models.py:120)");
    return V2W_ERR_INVALID;
  }
  const int s = (int)(q->step & 1);
  // input slot s was last read by the kernels of step-2, output slot s by the D2H of step-2
  if (q->step >= 2) {
    V2W_CUDA(cudaStreamWaitEvent(q->s_in, q->run_done[s], 0));
    V2W_CUDA(cudaStreamWaitEvent(q->s_run, q->out_done[s], 0));
  }
  V2W_CUDA(cudaMemcpyAsync(q->x[s], x_host, q->x_bytes, cudaMemcpyHostToDevice, q->s_in));
  V2W_CUDA(cudaMemcpyAsync(q->spk[s], spk_host, q->spk_bytes, cudaMemcpyHostToDevice, q->s_in));
  if (c.noise_dim > 0)
    V2W_CUDA(cudaMemcpyAsync(q->noise[s], noise_host, (size_t)q->plan->batch * c.noise_dim * 4,
                             cudaMemcpyHostToDevice, q->s_in));
  V2W_CUDA(cudaEventRecord(q->in_ready[s], q->s_in));
  V2W_CUDA(cudaStreamWaitEvent(q->s_run, q->in_ready[s], 0));
  int32_t n = 0;
  int rc = v2w_forward(q->plan, q->x[s], q->spk[s], q->noise[s], q->out[s], q->s_run, &n);
  if (rc) return rc;
  q->launches += n;
  V2W_CUDA(cudaEventRecord(q->run_done[s], q->s_run));
  V2W_CUDA(cudaStreamWaitEvent(q->s_out, q->run_done[s], 0));
  V2W_CUDA(cudaMemcpyAsync(out_host, q->out[s], q->out_bytes, cudaMemcpyDeviceToHost, q->s_out));
  V2W_CUDA(cudaEventRecord(q->out_done[s], q->s_out));
  ++q->step;
  return V2W_OK;
}

int32_t v2w_pipeline_wait(v2w_pipeline* q, int64_t* gpu_launches) {
  if (!q) {
    set_error("v2w_pipeline_wait: null argument");
    return V2W_ERR_INVALID;
  }
  V2W_CUDA(cudaStreamSynchronize(q->s_in));
  V2W_CUDA(cudaStreamSynchronize(q->s_run));
  V2W_CUDA(cudaStreamSynchronize(q->s_out));
  if (gpu_launches) *gpu_launches = q->launches;
  return V2W_OK;
}

int32_t v2w_tapgemm(const v2w_tapgemm_args* a, void* stream) {
  if (!a) {
    set_error("v2w_tapgemm: null argument");
    return V2W_ERR_INVALID;
  }
  int rc = validate_gemm(*a);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a->engine == V2W_ENGINE_TC) {
    TcLaunch L;
    rc = prepare_tc(*a, &L);
    if (rc) return rc;
    return launch_tc(L, st);
  }
  if (a->engine == V2W_ENGINE_SIMT) return launch_simt(*a, st);
  set_error("v2w_tapgemm: unknown engine %d", a->engine);
  return V2W_ERR_INVALID;
}

}  // extern "C"
