// Host half of libv2w: hyper-parameters -> tap-GEMM layer table, and state_dict -> packed
// weights (weight-norm / spectral-norm / batch-norm folding).  No CUDA in this file, so it is
// exercised by the CPU test-suite on a box without a GPU.
//
// Reference semantics restated here (never copied):
//   Generator.__init__        vec2wav/models.py:78-114
//   get_padding               vec2wav/utils.py:35-36
//   weight_norm(dim=0)        torch.nn.utils.weight_norm as used at models.py:18-33,58-61,83,90-92,100
//   ConditionalBatchNorm1d    vec2wav/modules.py:5-30  (eval: running stats, sigma = u.(W v))
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.h"

namespace v2w {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

static void add_conv1d(v2w_model* m, int role, int stage, int conv_index, int cin, int cout,
                       int k, int dil) {
  v2w_layer_desc L;
  std::memset(&L, 0, sizeof(L));
  L.role = role;
  L.stage = stage;
  L.conv_index = conv_index;
  L.cin = cin;
  L.cout = cout;
  L.up = 1;
  L.n_total = cout;
  L.kernel_size = k;
  L.dilation = dil;
  L.padding = (k * dil - dil) / 2;  // utils.py:35-36
  L.num_taps = k;
  for (int j = 0; j < k; ++j) {
    L.tap_off[j] = j * dil - L.padding;
    L.tap_mask[j] = 1u;
  }
  L.macs_per_row = (int64_t)k * cin * cout;
  m->layers.push_back(L);
}

// ConvTranspose1d(cin, cout, k, stride=u, padding=(k-u)//2) as a polyphase tap-GEMM.
//   y[t*u + kk - pad] += x[t] * w[kk]   ==>   for output row o = q*u + p:
//   kk = u*m + p + pad with input row t = q - m, valid while 0 <= kk < k.
static void add_convT(v2w_model* m, int stage, int conv_index, int cin, int cout, int k, int u) {
  v2w_layer_desc L;
  std::memset(&L, 0, sizeof(L));
  const int pad = (k - u) / 2;  // models.py:92
  L.role = V2W_ROLE_UPSAMPLE;
  L.stage = stage;
  L.conv_index = conv_index;
  L.cin = cin;
  L.cout = cout;
  L.up = u;
  L.n_total = u * cout;
  L.kernel_size = k;
  L.dilation = 1;
  L.padding = pad;
  int m_lo = 1 << 30, m_hi = -(1 << 30);
  for (int p = 0; p < u; ++p)
    for (int mm = -k; mm <= k; ++mm) {
      int kk = u * mm + p + pad;
      if (kk >= 0 && kk < k) {
        m_lo = std::min(m_lo, mm);
        m_hi = std::max(m_hi, mm);
      }
    }
  // ascending input-row offset: offset = -m, so walk m from high to low
  int nt = 0;
  int64_t pairs = 0;
  for (int mm = m_hi; mm >= m_lo; --mm) {
    uint32_t mask = 0;
    for (int p = 0; p < u; ++p) {
      int kk = u * mm + p + pad;
      if (kk >= 0 && kk < k) {
        mask |= 1u << p;
        ++pairs;
      }
    }
    L.tap_off[nt] = -mm;
    L.tap_mask[nt] = mask;
    ++nt;
  }
  L.num_taps = nt;
  L.macs_per_row = pairs * cin * cout;  // == k*cin*cout
  m->layers.push_back(L);
}

// Stacked program (C <= 64): see include/v2w.h.  Branch slots are sorted by kernel size so that the
// set of branches whose conv1 reaches time offset o is always a suffix {jmin..nb-1}; with the TMEM
// column order [T2 | T1_{nb-1} | ... | T1_0] that suffix plus the `sum` block is columns
// [0, (nb - jmin + 1) C).  `split_sum`: the first offset group feeds T1 only (columns [C, (nb+1) C))
// and its `sum` taps run as a separate N = C segment -- T2 is then not touched by the first MMAs of
// a tile, which takes the previous tile's final epilogue off the tensor pipe's critical path.
static bool plan_stacked(const v2w_model* m, int i, v2w_fused_desc& F) {
  const v2w_config& c = m->cfg;
  const int C = F.channels, nb = F.num_branches;
  const int d0 = c.resblock_dilation_sizes[0][0];
  if (C > 64 || (nb + 1) * C > 256 || 2 * (nb + 1) * C > 512) return false;
  int split_sum = C == 64 ? 1 : 0;
  if (const char* e = std::getenv("V2W_FUSED_SPLIT_SUM")) split_sum = std::atoi(e) != 0;
  // slots by ascending kernel size (stable)
  int order[V2W_MAX_KERNELS];
  for (int j = 0; j < nb; ++j) order[j] = j;
  std::stable_sort(order, order + nb, [&](int a, int b) {
    return c.resblock_kernel_sizes[a] < c.resblock_kernel_sizes[b];
  });
  int r[V2W_MAX_KERNELS];
  int hc = 0, entries = 0;
  for (int j = 0; j < nb; ++j) {
    const int k = c.resblock_kernel_sizes[order[j]];
    r[j] = (k - 1) / 2;
    hc = std::max(hc, c.resblock_dilation_sizes[order[j]][1] * r[j]);
    entries += k;
  }
  // Last stage with few channels: conv_post (C -> 1, k = 7) runs inside the final pass; 3 more halo
  // rows so that every central row finds its 7 neighbours in the window (kernels_stacked.cuh)
  static const bool no_post = [] { const char* e = std::getenv("V2W_NO_POST_FUSE"); return e && std::atoi(e); }();
  if (i == c.num_upsamples - 1 && C <= 32 && !no_post) {
    F.post = 1;
    hc += 3;
  }
  const int h1 = d0 * r[nb - 1];
  entries += 2 * r[nb - 1] + 1 + (split_sum ? 2 * r[0] + 1 : 0);
  const int rc = 256 - 2 * hc;
  const int a0_rows = 256 + 2 * h1;
  if (rc < 64 || a0_rows > 64 * 7 || entries > V2W_FUSED_MAX_TAPS || 2 * nb + 2 > V2W_FUSED_MAX_SEG)
    return false;
  F.stacked = 1;
  {
    // single accumulator set with `sum` split off (C = 64): T1 can be pre-loaded by the epilogue
    static const bool no_preload = [] { const char* e = std::getenv("V2W_FUSED_NO_PRELOAD"); return e && std::atoi(e); }();
    F.preload = (C == 64 && split_sum && !no_preload) ? 1 : 0;
  }
  for (int j = 0; j < nb; ++j) F.order[j] = order[j];
  F.hc = hc; F.h1 = h1; F.rc = rc;
  F.a0_boxes = (a0_rows + 63) / 64;
  F.t_bufs = 4 * (nb + 1) * C <= 512 ? 2 : 1;
  F.a0_bufs = F.t_bufs;
  int nseg = 0, nent = 0, wrow = 0;
  auto begin = [&](int type, int branch, int dcol, int n, int acc, int wait, int commit) {
    F.seg_type[nseg] = type; F.seg_branch[nseg] = branch;
    F.seg_dcol[nseg] = dcol; F.seg_n[nseg] = n; F.seg_acc[nseg] = acc;
    F.seg_wait[nseg] = wait; F.seg_commit[nseg] = commit;
    F.seg_first[nseg] = nent; F.seg_ntaps[nseg] = 0;
  };
  auto entry = [&](int off, int shift) {
    F.offset[nent] = off; F.shift[nent] = shift; F.w_row[nent] = wrow;
    wrow += F.seg_n[nseg];
    ++nent; ++F.seg_ntaps[nseg];
  };
  // phase A: offset groups g_j = { r[j-1] < |t| <= r[j] }, group 0 first (T1_0 completes first)
  int last_a = -1;
  for (int j = 0; j < nb; ++j) {
    const int lo = j == 0 ? -1 : r[j - 1];
    if (r[j] == lo) {                      // same kernel size as the previous slot: no new offsets
      F.seg_commit[last_a] |= V2W_FUSED_BIT_T1(j);
      continue;
    }
    const bool first = j == 0;
    const int nblk = nb - j + ((first && split_sum) ? 0 : 1);
    // The epilogue drains T1_0 .. T1_{nb-1}, then T2, in that order: "the last one is free" implies
    // the earlier ones are, so the first segment of a tile names ONE accumulator barrier.
    // preload: the accumulators hold x + b1 when conv1 starts, so its first MMA accumulates too, and
    // the barrier it waits for is "T1 pre-loaded" (bit T1(0), full-parity wait in the kernel)
    begin(V2W_FUSED_SEG_CONV1, -1, (first && split_sum) ? C : 0, nblk * C, (first && !F.preload) ? 0 : 1,
          first ? (V2W_FUSED_BIT_A0 | (F.preload ? V2W_FUSED_BIT_T1(0)
                                                 : (split_sum ? V2W_FUSED_BIT_T1(nb - 1) : V2W_FUSED_BIT_T2))) : 0,
          V2W_FUSED_BIT_T1(j));
    for (int t = -r[j]; t <= r[j]; ++t)
      if (std::abs(t) > lo) entry(t * d0, h1 + t * d0);
    last_a = nseg++;
    if (first && split_sum) {
      begin(V2W_FUSED_SEG_SUM, -1, 0, C, 0, V2W_FUSED_BIT_T2, 0);
      for (int t = -r[0]; t <= r[0]; ++t) entry(t * d0, h1 + t * d0);
      last_a = nseg++;
    }
  }
  F.seg_commit[last_a] |= V2W_FUSED_BIT_A0;
  // phase B: conv2 of every slot out of its A1 buffer
  for (int j = 0; j < nb; ++j) {
    const int d1 = c.resblock_dilation_sizes[order[j]][1];
    // no per-slot "A1 no longer read" commit: the epilogue overwrites A1 only after it has seen the
    // tile's T2 complete, and that commit covers every conv2 MMA
    begin(V2W_FUSED_SEG_CONV2, j, 0, C, 1, V2W_FUSED_BIT_A1(j), j == nb - 1 ? V2W_FUSED_BIT_T2 : 0);
    for (int t = -r[j]; t <= r[j]; ++t) entry(t * d1, t * d1);
    ++nseg;
  }
  F.num_segments = nseg;
  F.num_taps = nent;
  F.w_rows = wrow;
  F.ok = 1;
  return true;
}

// ResBlock1 conv pair  x' = c2(lrelu(c1(lrelu(x)) + b1)) + b2 + x  (models.py:37-44) as a two-segment
// program of the stacked kernel: conv1 (dilation d) of the whole 256-row window into T1, the
// epilogue turns it into A1 in shared memory, conv2 (dilation 1) into T2.  Halo = conv2's reach only
// ((k - 1) / 2 rows: 246 of 256 window rows are central for k = 11).
static v2w_fused_desc plan_pair(const v2w_model* m, int stage, int first_layer, int k, int d) {
  v2w_fused_desc F;
  std::memset(&F, 0, sizeof(F));
  const int C = m->stage_channels[stage];
  F.channels = C;
  F.num_branches = 1;
  F.first_layer = first_layer;
  F.pair = 1;
  F.stacked = 1;
  if (m->cfg.resblock_type != 1 || (C != 16 && C != 32 && C != 64 && C != 128)) return F;
  const int r = (k - 1) / 2;
  const int hc = r, h1 = d * r;
  if (C == 128) {
    // plain program of kernels_fused.cuh (two K-chunks; T2 rows start hc rows into the window)
    F.stacked = 0;
    const int rc = 256 - 2 * hc, a0_rows = 256 + hc + 2 * h1;
    if (rc < 64 || a0_rows > 64 * 5 || 2 * k > V2W_FUSED_MAX_TAPS) return F;   // A0 80 KB + A1 64 KB + ring
    F.hc = hc; F.h1 = h1; F.rc = rc;
    F.a0_boxes = (a0_rows + 63) / 64;
    for (int sgi = 0; sgi < 2; ++sgi) {
      F.seg_type[sgi] = sgi == 0 ? V2W_FUSED_SEG_CONV1 : V2W_FUSED_SEG_CONV2;
      F.seg_branch[sgi] = 0;
      F.seg_ntaps[sgi] = k;
      F.seg_first[sgi] = sgi * k;
      F.seg_n[sgi] = C;
    }
    F.seg_flags[0] = 1;   // conv1 is the last reader of the input tile
    F.seg_flags[1] = 2;   // conv2 is the last segment
    for (int t = 0; t < k; ++t) {
      F.offset[t] = (t - r) * d;  F.shift[t] = h1 + (t - r) * d;      F.w_row[t] = t * C;
      F.offset[k + t] = t - r;    F.shift[k + t] = hc + (t - r);      F.w_row[k + t] = (k + t) * C;
    }
    F.num_segments = 2;
    F.num_taps = 2 * k;
    F.w_rows = 2 * k * C;
    F.ok = 1;
    return F;
  }
  const int rc = 256 - 2 * hc, a0_rows = 256 + 2 * h1;
  if (rc < 64 || a0_rows > 64 * 7 || 2 * k > V2W_FUSED_MAX_TAPS) return F;
  F.hc = hc; F.h1 = h1; F.rc = rc;
  F.a0_boxes = (a0_rows + 63) / 64;
  F.t_bufs = 2; F.a0_bufs = 2;
  // TMEM columns of a sub-tile: [T2 | T1_0]
  F.seg_type[0] = V2W_FUSED_SEG_CONV1; F.seg_branch[0] = -1;
  F.seg_dcol[0] = C; F.seg_n[0] = C; F.seg_acc[0] = 0;
  F.seg_wait[0] = V2W_FUSED_BIT_A0 | V2W_FUSED_BIT_T1(0);
  F.seg_commit[0] = V2W_FUSED_BIT_T1(0) | V2W_FUSED_BIT_A0;
  F.seg_first[0] = 0; F.seg_ntaps[0] = k;
  F.seg_type[1] = V2W_FUSED_SEG_CONV2; F.seg_branch[1] = 0;
  F.seg_dcol[1] = 0; F.seg_n[1] = C; F.seg_acc[1] = 0;          // the first conv2 MMA overwrites T2
  F.seg_wait[1] = V2W_FUSED_BIT_A1(0) | V2W_FUSED_BIT_T2;
  F.seg_commit[1] = V2W_FUSED_BIT_T2;
  F.seg_first[1] = k; F.seg_ntaps[1] = k;
  for (int t = 0; t < k; ++t) {
    F.offset[t] = (t - r) * d;     F.shift[t] = h1 + (t - r) * d;  F.w_row[t] = t * C;
    F.offset[k + t] = t - r;       F.shift[k + t] = t - r;         F.w_row[k + t] = (k + t) * C;
  }
  F.num_segments = 2;
  F.num_taps = 2 * k;
  F.w_rows = 2 * k * C;
  F.ok = 1;
  return F;
}

// Decide whether stage `i` can run as one fused kernel and lay out its program.
// `first_layer` = index of branch 0's first conv in m->layers.
static v2w_fused_desc plan_fused_stage(const v2w_model* m, int i, int first_layer) {
  v2w_fused_desc F;
  std::memset(&F, 0, sizeof(F));
  const v2w_config& c = m->cfg;
  const int C = m->stage_channels[i];
  const int nb = c.num_kernels;
  F.channels = C;
  F.num_branches = nb;
  F.first_layer = first_layer;
  for (int j = 0; j < nb; ++j) F.order[j] = j;
  if (c.resblock_type != 2 || (C != 16 && C != 32 && C != 64 && C != 128)) return F;
  const int d0 = c.resblock_dilation_sizes[0][0];
  int kmax = 0, hc = 0, total = 0;
  for (int j = 0; j < nb; ++j) {
    if (c.resblock_dilation_sizes[j][0] != d0) return F;   // conv_sum needs aligned conv1 taps
    const int k = c.resblock_kernel_sizes[j];
    kmax = std::max(kmax, k);
    hc = std::max(hc, c.resblock_dilation_sizes[j][1] * (k - 1) / 2);
    total += 2 * k;
  }
  const char* no_stack = std::getenv("V2W_FUSED_NO_STACK");
  if (!(no_stack && std::atoi(no_stack)) && plan_stacked(m, i, F)) return F;
  const int h1 = d0 * (kmax - 1) / 2;
  total += kmax;
  // C = 128, "narrow" windows ($V2W_FUSED_NARROW=1): 128 rows per window instead of 256, two windows in flight
  // (kernels_fused.cuh Cfg<128, 1>)
  const char* narrow_env = std::getenv("V2W_FUSED_NARROW");
  const bool narrow = C == 128 && narrow_env && std::atoi(narrow_env);
  const int win = narrow ? 128 : 256;
  const int rc = win - 2 * hc;
  const int a0_rows = win + hc + 2 * h1;
  if (rc < 64 || a0_rows > 64 * 7 || total > V2W_FUSED_MAX_TAPS || 2 * nb + 1 > V2W_FUSED_MAX_SEG)
    return F;
  F.hc = hc; F.h1 = h1; F.rc = rc;
  F.a0_boxes = (a0_rows + 63) / 64;
  // issue order with double-buffered T1/A1 (C <= 64): c1_0, c1_1, SUM, then c2_j, c1_{j+2};
  // single-buffered (C = 128):                            c1_0, SUM, then c2_j, c1_{j+1}
  const int slots = (C <= 64 || narrow) ? 2 : 1;
  int nseg = 0, ntap = 0;
  // A segment = taps [t_lo, t_hi) of one convolution.  The `sum` convolution may be issued in PIECES
  // (seg_acc = 1 on every piece but the first: it accumulates from its first MMA).
  auto push = [&](int type, int j, int t_lo = 0, int t_hi = -1) {
    const int k = type == V2W_FUSED_SEG_SUM ? kmax : c.resblock_kernel_sizes[j];
    const int dil = type == V2W_FUSED_SEG_CONV2 ? c.resblock_dilation_sizes[j][1] : d0;
    if (t_hi < 0) t_hi = k;
    F.seg_type[nseg] = type;
    F.seg_branch[nseg] = j;
    F.seg_ntaps[nseg] = t_hi - t_lo;
    F.seg_first[nseg] = ntap;
    F.seg_n[nseg] = C;
    F.seg_acc[nseg] = (type == V2W_FUSED_SEG_SUM && t_lo > 0) ? 1 : 0;
    for (int t = t_lo; t < t_hi; ++t) {
      const int off = (t - (k - 1) / 2) * dil;
      F.offset[ntap] = off;
      F.shift[ntap] = type == V2W_FUSED_SEG_CONV1 ? h1 + off
                      : type == V2W_FUSED_SEG_SUM ? hc + h1 + off
                                                  : hc + off;
      F.w_row[ntap] = ntap * C;
      ++ntap;
    }
    ++nseg;
  };
  // Single-buffered (C = 128, 256-row windows), optional ($V2W_FUSED_SUM_PIECES=n): the `sum` convolution cut into
  // n pieces, one behind every conv1 --  c1_0 SUM_a c2_0 c1_1 SUM_b c2_1 c1_2 SUM_c c2_2  -- so that independent
  // MMAs are queued while the per-branch pass T1_j -> A1_j runs.  MEASURED NEUTRAL (cfg2 119.5 vs 119.5 us, cfg5/GPU
  // 1059 vs 1047 us): the stage is not waiting for those passes (profiles/r02_notes.md), so the default stays
  // round 1's  c1_0 SUM c2_0 c1_1 c2_1 c1_2 c2_2.
  int pieces = 1;
  if (const char* e = std::getenv("V2W_FUSED_SUM_PIECES")) pieces = std::max(1, std::min(nb, std::atoi(e)));
  if (slots != 1) pieces = 1;
  if (2 * nb + pieces > V2W_FUSED_MAX_SEG) pieces = 1;
  auto sum_piece = [&](int q) {     // taps [q * kmax / pieces, (q + 1) * kmax / pieces)
    push(V2W_FUSED_SEG_SUM, -1, q * kmax / pieces, (q + 1) * kmax / pieces);
  };
  push(V2W_FUSED_SEG_CONV1, 0);
  if (slots == 2 && nb > 1) push(V2W_FUSED_SEG_CONV1, 1);
  sum_piece(0);
  for (int j = 0; j < nb; ++j) {
    push(V2W_FUSED_SEG_CONV2, j);
    if (j + slots < nb) {
      push(V2W_FUSED_SEG_CONV1, j + slots);
      if (j + 1 < pieces) sum_piece(j + 1);
    }
  }
  int last_a0 = 0;
  for (int s = 0; s < nseg; ++s)
    if (F.seg_type[s] != V2W_FUSED_SEG_CONV2) last_a0 = s;
  F.seg_flags[last_a0] |= 1;
  F.seg_flags[nseg - 1] |= 2;
  F.num_segments = nseg;
  F.num_taps = ntap;
  F.w_rows = ntap * C;
  F.ok = 1;
  return F;
}

}  // namespace v2w

using namespace v2w;

// Wide stage (v2w_wide_desc, include/v2w.h): ResBlock2 stages with C >= 64 can run as two stacked
// tap-GEMM launches.  Called right after the stage's layers were pushed.
static v2w_wide_desc plan_wide_stage(const v2w_model* m, int stage, int first_layer) {
  v2w_wide_desc W;
  std::memset(&W, 0, sizeof(W));
  const v2w_config& c = m->cfg;
  const int C = m->stage_channels[stage];
  const int nb = c.num_kernels;
  W.channels = C;
  W.num_branches = nb;
  W.first_layer = first_layer;
  if (c.resblock_type != 2 || C < 64 || C % 64 || C > 256 || nb + 1 > V2W_MAX_TILES) return W;
  // union of the conv1 offsets, ascending
  std::vector<int> offs;
  for (int j = 0; j < nb; ++j) {
    const v2w_layer_desc& L = m->layers[first_layer + 2 * j];
    for (int t = 0; t < L.num_taps; ++t)
      if (std::find(offs.begin(), offs.end(), L.tap_off[t]) == offs.end()) offs.push_back(L.tap_off[t]);
  }
  std::sort(offs.begin(), offs.end());
  if ((int)offs.size() > V2W_MAX_TAPS) return W;
  W.sum_taps = (int)offs.size();
  for (int t = 0; t < W.sum_taps; ++t) W.sum_off[t] = offs[t];
  W.ok = 1;
  return W;
}

extern "C" {

int32_t v2w_abi_version(void) { return V2W_ABI_VERSION; }

const char* v2w_last_error(void) { return g_last_error.c_str(); }

int32_t v2w_model_create(const v2w_config* c, v2w_model** out) {
  if (!c || !out) {
    set_error("v2w_model_create: null argument");
    return V2W_ERR_INVALID;
  }
  *out = nullptr;
  const int U = c->num_upsamples;
  if (U < 1 || U > V2W_MAX_STAGES) {
    set_error("num_upsamples=%d out of range [1,%d]", U, V2W_MAX_STAGES);
    return V2W_ERR_INVALID;
  }
  if (c->num_kernels < 1 || c->num_kernels > V2W_MAX_KERNELS) {
    set_error("num_kernels=%d out of range [1,%d]", c->num_kernels, V2W_MAX_KERNELS);
    return V2W_ERR_INVALID;
  }
  if (c->resblock_type != 1 && c->resblock_type != 2) {
    set_error("resblock_type must be 1 or 2 (got %d)", c->resblock_type);
    return V2W_ERR_INVALID;
  }
  // ResBlock2 reads dilation[0..1], ResBlock1 all three (models.py:14,54); the constructors index
  // them positionally, so fewer entries is an IndexError in the reference.
  const int need_dil = c->resblock_type == 1 ? 3 : 2;
  if (c->num_dilations < need_dil || c->num_dilations > V2W_MAX_DILATIONS) {
    set_error("resblock_dilation_sizes rows need >= %d entries (got %d, max %d)", need_dil,
              c->num_dilations, V2W_MAX_DILATIONS);
    return V2W_ERR_INVALID;
  }
  // models.py:113 hard-codes the CBN widths 256>>i, which equals the ups[i] output width only
  // when upsample_initial_channel == 512; anything else raises a shape error in the reference.
  if (c->upsample_initial_channel != 512) {
    set_error("upsample_initial_channel must be 512 (CBN widths are hard-coded 256>>i at "
              "vec2wav/models.py:113); got %d", c->upsample_initial_channel);
    return V2W_ERR_INVALID;
  }
  if (c->num_wv_feat < 16 || c->num_wv_feat % 16 != 0) {
    set_error("num_wv_feat must be a positive multiple of 16 (got %d)", c->num_wv_feat);
    return V2W_ERR_INVALID;
  }
  if (c->num_wv_feat > 64 && c->num_wv_feat % 64 != 0) {
    set_error("num_wv_feat above 64 must be a multiple of 64 (got %d)", c->num_wv_feat);
    return V2W_ERR_INVALID;
  }
  if (c->spk_dim < 1 || c->noise_dim < 0 || c->spk_dim + c->noise_dim > 4096) {
    set_error("spk_dim/noise_dim out of range (%d, %d)", c->spk_dim, c->noise_dim);
    return V2W_ERR_INVALID;
  }
  if ((c->upsample_initial_channel >> U) < 16) {
    set_error("too many upsample stages: last stage would have %d channels (< 16)",
              c->upsample_initial_channel >> U);
    return V2W_ERR_INVALID;
  }
  if ((c->upsample_initial_channel >> U) > 32) {
    set_error("conv_post kernel supports 16 or 32 input channels; %d stages leave %d", U,
              c->upsample_initial_channel >> U);
    return V2W_ERR_INVALID;
  }
  for (int i = 0; i < U; ++i) {
    int k = c->upsample_kernel_sizes[i], u = c->upsample_rates[i];
    if (u < 1 || k < u || ((k - u) & 1)) {
      set_error("upsample stage %d: kernel %d / rate %d unsupported (need k >= u and k-u even so "
                "that L_out == L_in*u)", i, k, u);
      return V2W_ERR_INVALID;
    }
    if (u > 8) {
      set_error("upsample rate %d > 8 unsupported", u);
      return V2W_ERR_INVALID;
    }
  }
  for (int j = 0; j < c->num_kernels; ++j) {
    int k = c->resblock_kernel_sizes[j];
    if (k < 1 || !(k & 1) || k > V2W_MAX_TAPS) {
      set_error("resblock kernel size %d unsupported (odd, <= %d)", k, V2W_MAX_TAPS);
      return V2W_ERR_INVALID;
    }
    for (int d = 0; d < need_dil; ++d)
      if (c->resblock_dilation_sizes[j][d] < 1 || c->resblock_dilation_sizes[j][d] > 32) {
        set_error("resblock dilation %d unsupported", c->resblock_dilation_sizes[j][d]);
        return V2W_ERR_INVALID;
      }
  }

  v2w_model* m = new (std::nothrow) v2w_model();
  if (!m) {
    set_error("out of memory");
    return V2W_ERR_NOMEM;
  }
  m->cfg = *c;

  // ---- conv index bookkeeping: conv_pre, ups[0..U), resblock convs (execution order), conv_post
  const int convs_per_block = c->resblock_type == 1 ? 6 : 2;
  int rb_conv_base = 1 + U;
  add_conv1d(m, V2W_ROLE_CONV_PRE, -1, 0, c->num_wv_feat, c->upsample_initial_channel, 7, 1);
  int ch = c->upsample_initial_channel;
  int up_total = 1;
  for (int i = 0; i < U; ++i) {
    const int cin = ch;
    ch = c->upsample_initial_channel >> (i + 1);
    add_convT(m, i, 1 + i, cin, ch, c->upsample_kernel_sizes[i], c->upsample_rates[i]);
    if (m->layers.back().num_taps > V2W_MAX_TAPS) {
      set_error("upsample stage %d needs too many taps", i);
      delete m;
      return V2W_ERR_INVALID;
    }
    up_total *= c->upsample_rates[i];
    m->stage_channels.push_back(ch);
    m->stage_up.push_back(up_total);
    m->fused.push_back(plan_fused_stage(m, i, (int)m->layers.size()));
    const int stage_first_layer = (int)m->layers.size();
    for (int j = 0; j < c->num_kernels; ++j) {
      const int k = c->resblock_kernel_sizes[j];
      const int block = i * c->num_kernels + j;
      const int base = rb_conv_base + block * convs_per_block;
      if (c->resblock_type == 2) {
        // ResBlock2.forward (models.py:65-70): dilations d[0], d[1]
        for (int s = 0; s < 2; ++s)
          add_conv1d(m, V2W_ROLE_RESBLOCK, i, base + s, ch, ch, k,
                     c->resblock_dilation_sizes[j][s]);
      } else {
        // ResBlock1.forward (models.py:37-44): (convs1[s] dil d[s], convs2[s] dil 1) x 3
        for (int s = 0; s < 3; ++s) {
          const int first = (int)m->layers.size();
          add_conv1d(m, V2W_ROLE_RESBLOCK, i, base + 2 * s, ch, ch, k,
                     c->resblock_dilation_sizes[j][s]);
          add_conv1d(m, V2W_ROLE_RESBLOCK, i, base + 2 * s + 1, ch, ch, k, 1);
          const v2w_fused_desc P = plan_pair(m, i, first, k, c->resblock_dilation_sizes[j][s]);
          m->pair_of_layer.resize(m->layers.size(), -1);
          if (P.ok) {
            m->pair_of_layer[first] = (int)m->pairs.size();
            m->pairs.push_back(P);
          }
        }
      }
    }
    m->wide.push_back(plan_wide_stage(m, i, stage_first_layer));
  }
  m->total_up = up_total;
  m->num_convs = rb_conv_base + U * c->num_kernels * convs_per_block + 1;
  add_conv1d(m, V2W_ROLE_CONV_POST, -1, m->num_convs - 1, ch, 1, 7, 1);

  // ---- packed layout (after the format header)
  int64_t off = V2W_PACKED_HEADER_BYTES;
  int64_t flops = 0;
  for (auto& L : m->layers) {
    if (L.role != V2W_ROLE_CONV_POST) {
      L.w16_off = off;
      off = align_up(off + (int64_t)2 * L.num_taps * L.n_total * L.cin * 2, 1024);   // operands | lo halves
    } else {
      L.w16_off = -1;
    }
    L.w32_off = off;
    off = align_up(off + (int64_t)L.num_taps * L.cin * L.n_total * 4, 1024);
    L.bias_off = off;
    off = align_up(off + (int64_t)L.cout * 4, 1024);
    int rows_per_frame = 1;
    if (L.role == V2W_ROLE_UPSAMPLE) rows_per_frame = L.stage == 0 ? 1 : m->stage_up[L.stage - 1];
    else if (L.role == V2W_ROLE_RESBLOCK) rows_per_frame = m->stage_up[L.stage];
    else if (L.role == V2W_ROLE_CONV_POST) rows_per_frame = up_total;
    flops += 2 * L.macs_per_row * rows_per_frame;
  }
  const int D = c->spk_dim + c->noise_dim;
  for (int i = 0; i < U; ++i) {
    CondLayout cl;
    cl.C = m->stage_channels[i];
    cl.D = D;
    cl.fcw = off; off = align_up(off + (int64_t)kZ * D * 4, 256);
    cl.fcb = off; off = align_up(off + (int64_t)kZ * 4, 256);
    cl.snw = off; off = align_up(off + (int64_t)2 * cl.C * kZ * 4, 256);
    cl.snb = off; off = align_up(off + (int64_t)2 * cl.C * 4, 256);
    cl.rstd = off; off = align_up(off + (int64_t)cl.C * 4, 256);
    cl.mean = off; off = align_up(off + (int64_t)cl.C * 4, 256);
    m->cond.push_back(cl);
  }
  for (auto* vec : {&m->fused, &m->pairs}) {
    for (auto& F : *vec) {
      if (!F.ok) continue;
      F.w_off = off;
      off = align_up(off + (int64_t)F.w_rows * F.channels * 2, 1024);
      F.bsum_off = off;
      off = align_up(off + (int64_t)F.channels * 4, 1024);
    }
  }
  for (auto& W : m->wide) {
    if (!W.ok) continue;
    W.wsum_off = off;
    off = align_up(off + (int64_t)W.sum_taps * W.channels * W.channels * 2, 1024);
    W.b1cat_off = off;
    off = align_up(off + (int64_t)W.num_branches * W.channels * 4, 1024);
    W.bsum_off = off;
    off = align_up(off + (int64_t)W.channels * 4, 1024);
  }
  m->pair_of_layer.resize(m->layers.size(), -1);
  m->packed_bytes = align_up(off, 1024);
  m->flops_per_frame = flops;
  *out = m;
  return V2W_OK;
}

void v2w_model_destroy(v2w_model* m) { delete m; }

int32_t v2w_model_num_layers(const v2w_model* m) { return m ? (int32_t)m->layers.size() : 0; }

int32_t v2w_model_get_layer(const v2w_model* m, int32_t i, v2w_layer_desc* out) {
  if (!m || !out || i < 0 || i >= (int32_t)m->layers.size()) {
    set_error("v2w_model_get_layer: bad index %d", i);
    return V2W_ERR_INVALID;
  }
  *out = m->layers[i];
  return V2W_OK;
}

int32_t v2w_model_get_cond(const v2w_model* m, int32_t i, v2w_cond_desc* out) {
  if (!m || !out || i < 0 || i >= (int32_t)m->cond.size()) {
    set_error("v2w_model_get_cond: bad stage %d", i);
    return V2W_ERR_INVALID;
  }
  const CondLayout& c = m->cond[i];
  out->channels = c.C; out->in_dim = c.D;
  out->fcw_off = c.fcw; out->fcb_off = c.fcb; out->snw_off = c.snw; out->snb_off = c.snb;
  out->rstd_off = c.rstd; out->mean_off = c.mean;
  return V2W_OK;
}

int32_t v2w_model_halo_frames(const v2w_model* m) {
  if (!m) return 0;
  const v2w_config& c = m->cfg;
  // walk back from the waveform: R = rows past the utterance's end (at the current resolution)
  // that the last valid row depends on
  int64_t R = 3;                                    // conv_post k = 7 (models.py:100)
  for (int i = c.num_upsamples - 1; i >= 0; --i) {
    int reach = 0;                                  // one MRF stage: the widest branch
    for (int j = 0; j < c.num_kernels; ++j) {
      const int half = (c.resblock_kernel_sizes[j] - 1) / 2;
      int r = 0;
      if (c.resblock_type == 2) {
        r = (c.resblock_dilation_sizes[j][0] + c.resblock_dilation_sizes[j][1]) * half;   // models.py:54-70
      } else {
        for (int d = 0; d < 3; ++d) r += (c.resblock_dilation_sizes[j][d] + 1) * half;    // models.py:14-44
      }
      reach = std::max(reach, r);
    }
    R += reach;
    // ConvTranspose1d(k, stride u, padding p): output row t reads input rows <= floor((t + p) / u)
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i], pad = (k - u) / 2;
    R = (R + pad + u - 1) / u;
  }
  R += 3;                                           // conv_pre k = 7 (models.py:83)
  return (int32_t)R;
}

int32_t v2w_model_num_pairs(const v2w_model* m) { return m ? (int32_t)m->pairs.size() : 0; }

int32_t v2w_model_get_pair(const v2w_model* m, int32_t i, v2w_fused_desc* out) {
  if (!m || !out || i < 0 || i >= (int32_t)m->pairs.size()) {
    set_error("v2w_model_get_pair: bad index %d", i);
    return V2W_ERR_INVALID;
  }
  *out = m->pairs[i];
  return V2W_OK;
}

int32_t v2w_model_get_fused(const v2w_model* m, int32_t i, v2w_fused_desc* out) {
  if (!m || !out || i < 0 || i >= (int32_t)m->fused.size()) {
    set_error("v2w_model_get_fused: bad stage %d", i);
    return V2W_ERR_INVALID;
  }
  *out = m->fused[i];
  return V2W_OK;
}

int32_t v2w_model_get_wide(const v2w_model* m, int32_t i, v2w_wide_desc* out) {
  if (!m || !out || i < 0 || i >= (int32_t)m->wide.size()) {
    set_error("v2w_model_get_wide: bad stage %d", i);
    return V2W_ERR_INVALID;
  }
  *out = m->wide[i];
  return V2W_OK;
}

int32_t v2w_model_num_convs(const v2w_model* m) { return m ? m->num_convs : 0; }
}  // extern "C"

// Conditioning section of one stage: fcs copied, spectral norm folded with the STORED u, v (eval()
// semantics of modules.py:16: sigma = u . (W v), no power iteration), BN running statistics as
// mean / rstd.  `base` + the layout's offsets is where the section goes.
static int pack_cond_stage(uint8_t* base, const v2w::CondLayout& cl, const v2w_cond_params& q, int stage) {
  using namespace v2w;
  if (!q.fc_weight || !q.fc_bias || !q.sn_weight_orig || !q.sn_u || !q.sn_v || !q.sn_bias ||
      !q.bn_mean || !q.bn_var) {
    set_error("v2w_model_pack: cond stage %d has a null pointer", stage);
    return V2W_ERR_INVALID;
  }
  std::memcpy(base + cl.fcw, q.fc_weight, (size_t)kZ * cl.D * 4);
  std::memcpy(base + cl.fcb, q.fc_bias, (size_t)kZ * 4);
  double sigma = 0.0;
  for (int r = 0; r < 2 * cl.C; ++r) {
    double wv = 0.0;
    for (int c2 = 0; c2 < kZ; ++c2) wv += (double)q.sn_weight_orig[(size_t)r * kZ + c2] * q.sn_v[c2];
    sigma += (double)q.sn_u[r] * wv;
  }
  const float fsigma = (float)sigma;
  float* snw = reinterpret_cast<float*>(base + cl.snw);
  for (int e = 0; e < 2 * cl.C * kZ; ++e) snw[e] = q.sn_weight_orig[e] / fsigma;
  std::memcpy(base + cl.snb, q.sn_bias, (size_t)2 * cl.C * 4);
  float* rstd = reinterpret_cast<float*>(base + cl.rstd);
  float* mean = reinterpret_cast<float*>(base + cl.mean);
  for (int ch = 0; ch < cl.C; ++ch) {
    rstd[ch] = 1.0f / std::sqrt(q.bn_var[ch] + kBnEps);
    mean[ch] = q.bn_mean[ch];
  }
  return V2W_OK;
}

extern "C" {

int32_t v2w_model_cond_region(const v2w_model* m, int64_t* offset, int64_t* bytes) {
  if (!m || !offset || !bytes || m->cond.empty()) {
    v2w::set_error("v2w_model_cond_region: null argument");
    return V2W_ERR_INVALID;
  }
  const v2w::CondLayout& last = m->cond.back();
  *offset = m->cond.front().fcw;
  *bytes = v2w::align_up(last.mean + (int64_t)last.C * 4, 256) - *offset;
  return V2W_OK;
}

int32_t v2w_model_pack_cond(const v2w_model* m, const v2w_cond_params* cond, int32_t num_stages,
                            void* region_host) {
  if (!m || !cond || !region_host || num_stages != (int32_t)m->cond.size()) {
    v2w::set_error("v2w_model_pack_cond: bad argument (stages %d, model %d)", num_stages,
                   m ? (int)m->cond.size() : -1);
    return V2W_ERR_INVALID;
  }
  int64_t off = 0, bytes = 0;
  int rc = v2w_model_cond_region(m, &off, &bytes);
  if (rc) return rc;
  std::memset(region_host, 0, (size_t)bytes);
  uint8_t* base = static_cast<uint8_t*>(region_host) - off;   // the layouts hold whole-buffer offsets
  for (size_t i = 0; i < m->cond.size(); ++i) {
    rc = pack_cond_stage(base, m->cond[i], cond[i], (int)i);
    if (rc) return rc;
  }
  return V2W_OK;
}

}  // extern "C"

extern "C" {

int64_t v2w_model_packed_bytes(const v2w_model* m) { return m ? m->packed_bytes : 0; }
int64_t v2w_model_flops_per_frame(const v2w_model* m) { return m ? m->flops_per_frame : 0; }

int32_t v2w_model_pack(const v2w_model* m, const v2w_params* P, void* packed_host) {
  return v2w_model_pack_half(m, P, packed_host, V2W_HALF_BF16);
}

int32_t v2w_model_pack_half(const v2w_model* m, const v2w_params* P, void* packed_host,
                            int32_t half_format) {
  if (!m || !P || !packed_host) {
    set_error("v2w_model_pack: null argument");
    return V2W_ERR_INVALID;
  }
  if (half_format < V2W_HALF_BF16 || half_format > V2W_HALF_BF16_SPLIT) {
    set_error("v2w_model_pack: unknown 16-bit format %d", half_format);
    return V2W_ERR_INVALID;
  }
  const bool f16 = half_format == V2W_HALF_FP16 || half_format == V2W_HALF_FP16_SPLIT;
  const bool split = half_format == V2W_HALF_FP16_SPLIT || half_format == V2W_HALF_BF16_SPLIT;
  auto half_lo_bits = [f16](float w) {     // the part of w its 16-bit value does not hold, in the same format
    return f16 ? fp16_bits(w - fp16_value(fp16_bits(w))) : bf16_bits(w - bf16_value(bf16_bits(w)));
  };
  auto half_bits = [f16](float w) { return f16 ? fp16_bits(w) : bf16_bits(w); };
  if (P->num_convs != m->num_convs || P->num_stages != (int)m->cond.size()) {
    set_error("v2w_model_pack: expected %d convs / %d stages, got %d / %d", m->num_convs,
              (int)m->cond.size(), P->num_convs, P->num_stages);
    return V2W_ERR_INVALID;
  }
  uint8_t* base = static_cast<uint8_t*>(packed_host);
  std::memset(base, 0, (size_t)m->packed_bytes);
  {
    uint32_t hdr[3] = {kPackedMagic, V2W_ABI_VERSION, (uint32_t)half_format};
    std::memcpy(base, hdr, sizeof(hdr));
  }
  for (const auto& L : m->layers) {
    const v2w_conv_params& cp = P->convs[L.conv_index];
    if (!cp.v || !cp.bias) {
      set_error("v2w_model_pack: conv %d has null weight/bias", L.conv_index);
      return V2W_ERR_INVALID;
    }
  }
  // Two parallel phases over independent tasks (8.6 - 16.8 M weights in all; a single thread takes
  // 110 - 240 ms, which every parameter update pays before its next forward):
  //   1. per layer: the weight-norm scales (one sequential pass over v) and the bias copy;
  //   2. per (layer, tap): the folded weights in both layouts -- conv_pre alone is a third of the
  //      weights, so whole layers would leave the other threads idle behind it.
  const int nthreads = pack_threads();
  const int nl = (int)m->layers.size();
  std::vector<std::vector<float>> scales((size_t)nl);
  parallel_for(nl, nthreads, [&](int li) {
    const v2w_layer_desc& L = m->layers[li];
    const v2w_conv_params& cp = P->convs[L.conv_index];
    const bool transposed = L.role == V2W_ROLE_UPSAMPLE;
    // weight_norm(dim=0): one norm per index of dim 0 -- Cout for Conv1d [Cout,Cin,K],
    // Cin for ConvTranspose1d [Cin,Cout,K].
    const int d0 = transposed ? L.cin : L.cout;
    const int inner = (transposed ? L.cout : L.cin) * L.kernel_size;
    std::vector<float>& scale = scales[(size_t)li];
    scale.assign(d0, 1.0f);
    if (cp.g) {
      for (int r = 0; r < d0; ++r) {
        double ss = 0.0;
        const float* row = cp.v + (size_t)r * inner;
        for (int e = 0; e < inner; ++e) ss += (double)row[e] * row[e];
        scale[r] = (float)((double)cp.g[r] / std::sqrt(ss));
      }
    }
    float* bias = reinterpret_cast<float*>(base + L.bias_off);
    for (int co = 0; co < L.cout; ++co) bias[co] = cp.bias[co];
  });
  std::vector<std::pair<int, int>> tasks;   // (layer, tap)
  for (int li = 0; li < nl; ++li)
    for (int j = 0; j < m->layers[li].num_taps; ++j) tasks.emplace_back(li, j);
  parallel_for((int)tasks.size(), nthreads, [&](int ti) {
    const int li = tasks[(size_t)ti].first, j = tasks[(size_t)ti].second;
    const v2w_layer_desc& L = m->layers[li];
    const v2w_conv_params& cp = P->convs[L.conv_index];
    const std::vector<float>& scale = scales[(size_t)li];
    const int K = L.kernel_size;
    const bool transposed = L.role == V2W_ROLE_UPSAMPLE;
    uint16_t* w16 = L.w16_off >= 0 ? reinterpret_cast<uint16_t*>(base + L.w16_off) : nullptr;
    uint16_t* w16lo = (w16 && split) ? w16 + (size_t)L.num_taps * L.n_total * L.cin : nullptr;
    float* w32 = reinterpret_cast<float*>(base + L.w32_off);
    for (int p = 0; p < L.up; ++p) {
      if (!(L.tap_mask[j] >> p & 1u)) continue;
      // kernel element addressed by (tap j, phase p)
      const int kk = transposed ? (L.up * (-L.tap_off[j]) + p + L.padding) : j;
      for (int co = 0; co < L.cout; ++co) {
        const int n = p * L.cout + co;
        for (int ci = 0; ci < L.cin; ++ci) {
          float w;
          if (transposed) w = cp.v[((size_t)ci * L.cout + co) * K + kk] * scale[ci];
          else w = cp.v[((size_t)co * L.cin + ci) * K + kk] * scale[co];
          if (w16) w16[((size_t)j * L.n_total + n) * L.cin + ci] = half_bits(w);
          if (w16lo) w16lo[((size_t)j * L.n_total + n) * L.cin + ci] = half_lo_bits(w);
          w32[((size_t)j * L.cin + ci) * L.n_total + n] = w;
        }
      }
    }
  });
  for (size_t i = 0; i < m->cond.size(); ++i) {
    int rc = pack_cond_stage(base, m->cond[i], P->cond[i], (int)i);
    if (rc) return rc;
  }
  // fused stages: tap matrices in issue order, from the folded fp32 weights packed above
  std::vector<const v2w_fused_desc*> all_fused;
  for (const auto& F : m->fused) all_fused.push_back(&F);
  for (const auto& F : m->pairs) all_fused.push_back(&F);
  parallel_for((int)all_fused.size(), nthreads, [&](int fi) {
    const v2w_fused_desc& F = *all_fused[fi];
    if (!F.ok) return;
    const int C = F.channels;
    uint16_t* wf = reinterpret_cast<uint16_t*>(base + F.w_off);
    float* bsum = reinterpret_cast<float*>(base + F.bsum_off);
    for (int ch = 0; ch < C; ++ch) bsum[ch] = 0.f;
    // a ResBlock1 pair adds only conv2's bias at the end (b1 went into the middle activation)
    for (int j = F.pair ? 1 : 0; j < 2 * F.num_branches; ++j) {
      const float* bl = reinterpret_cast<const float*>(base + m->layers[F.first_layer + j].bias_off);
      for (int ch = 0; ch < C; ++ch) bsum[ch] += bl[ch];
    }
    // weight matrix [w_rows][C]: per entry one C x C block (n, c) per accumulator block it feeds
    std::vector<float> acc((size_t)C * C);
    const int nb = F.num_branches;
    for (int s = 0; s < F.num_segments; ++s) {
      const int type = F.seg_type[s];
      const int nblk = F.seg_n[s] / C;
      for (int t = 0; t < F.seg_ntaps[s]; ++t) {
        const int ent = F.seg_first[s] + t;
        for (int q = 0; q < nblk; ++q) {
          // which conv feeds block q: -1 = sum of all conv1, else model branch `br`, conv `which`
          int br = -1, which = 0;
          if (type == V2W_FUSED_SEG_CONV2) {
            br = F.stacked ? F.order[F.seg_branch[s]] : F.seg_branch[s];
            which = 1;
          } else if (type == V2W_FUSED_SEG_CONV1) {
            if (F.stacked) {
              const int col = F.seg_dcol[s] / C + q;       // 0 = T2 (sum), c = T1 of slot nb - c
              br = col == 0 ? -1 : F.order[nb - col];
            } else {
              br = F.seg_branch[s];
            }
          }
          std::fill(acc.begin(), acc.end(), 0.f);
          for (int j = 0; j < nb; ++j) {
            if (br >= 0 && br != j) continue;
            const v2w_layer_desc& L = m->layers[F.first_layer + 2 * j + which];
            int lt = -1;
            for (int u = 0; u < L.num_taps; ++u)
              if (L.tap_off[u] == F.offset[ent]) lt = u;
            if (lt < 0) continue;   // this branch's kernel does not reach that offset
            const float* w32 = reinterpret_cast<const float*>(base + L.w32_off) + (size_t)lt * C * C;
            for (int ci = 0; ci < C; ++ci)
              for (int n = 0; n < C; ++n) acc[(size_t)n * C + ci] += w32[(size_t)ci * C + n];
          }
          uint16_t* dst = wf + ((size_t)F.w_row[ent] + (size_t)q * C) * C;
          for (size_t e = 0; e < (size_t)C * C; ++e) dst[e] = half_bits(acc[e]);
        }
      }
    }
  });
  // wide stages: the `sum` convolution (per-offset sums of the conv1 tap matrices, branch order), the
  // concatenated first-conv biases and the total bias
  parallel_for((int)m->wide.size(), nthreads, [&](int wi) {
    const v2w_wide_desc& W = m->wide[(size_t)wi];
    if (!W.ok) return;
    const int C = W.channels, nb = W.num_branches;
    uint16_t* ws = reinterpret_cast<uint16_t*>(base + W.wsum_off);
    std::vector<float> acc((size_t)C * C);
    for (int t = 0; t < W.sum_taps; ++t) {
      std::fill(acc.begin(), acc.end(), 0.f);
      for (int j = 0; j < nb; ++j) {
        const v2w_layer_desc& L = m->layers[W.first_layer + 2 * j];
        int lt = -1;
        for (int u = 0; u < L.num_taps; ++u)
          if (L.tap_off[u] == W.sum_off[t]) lt = u;
        if (lt < 0) continue;
        const float* w32 = reinterpret_cast<const float*>(base + L.w32_off) + (size_t)lt * C * C;
        for (int ci = 0; ci < C; ++ci)
          for (int n = 0; n < C; ++n) acc[(size_t)n * C + ci] += w32[(size_t)ci * C + n];
      }
      uint16_t* dst = ws + (size_t)t * C * C;
      for (size_t e = 0; e < (size_t)C * C; ++e) dst[e] = half_bits(acc[e]);
    }
    float* b1cat = reinterpret_cast<float*>(base + W.b1cat_off);
    float* bsum = reinterpret_cast<float*>(base + W.bsum_off);
    for (int ch = 0; ch < C; ++ch) bsum[ch] = 0.f;
    for (int j = 0; j < 2 * nb; ++j) {
      const float* bl = reinterpret_cast<const float*>(base + m->layers[W.first_layer + j].bias_off);
      for (int ch = 0; ch < C; ++ch) bsum[ch] += bl[ch];
      if (!(j & 1)) {
        // (a bias of 0.f + b is b: the device packer's `a = 0.f; a += b` gives the same bits)
        for (int ch = 0; ch < C; ++ch) b1cat[(size_t)(j / 2) * C + ch] = 0.f + bl[ch];
      }
    }
  });
  return V2W_OK;
}

}  // extern "C"
