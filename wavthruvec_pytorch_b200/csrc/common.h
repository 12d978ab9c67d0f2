// Internal declarations shared by the host (model/pack/plan) and device (kernels) halves of
// libv2w.  Nothing here is part of the ABI; the ABI is include/v2w.h.
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#include "v2w.h"

namespace v2w {

void set_error(const char* fmt, ...);

static inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// round-to-nearest-even fp32 -> bf16 bit pattern (NaN-safe enough for weights)
static inline uint16_t bf16_bits(float f) {
  uint32_t u;
  __builtin_memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // quiet NaN
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

// round-to-nearest-even fp32 -> IEEE fp16 bit pattern (overflow saturates to +-65504)
static inline uint16_t fp16_bits(float f) {
  uint32_t u;
  __builtin_memcpy(&u, &f, 4);
  const uint32_t sign = (u >> 16) & 0x8000u;
  u &= 0x7fffffffu;
  if (u > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
  if (u >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);         // >= 65520 -> max finite
  if (u < 0x38800000u) {                                           // subnormal or zero in fp16
    if (u < 0x33000000u) return (uint16_t)sign;                    // < 2^-25 -> 0
    const uint32_t shift = 126u - (u >> 23);                       // 14 .. 24
    const uint32_t mant = (u & 0x7fffffu) | 0x800000u;
    uint32_t h = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((u - 0x38000000u) >> 13);
  const uint32_t rem = u & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}

static inline float bf16_value(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
}

// IEEE fp16 bit pattern -> fp32 (exact)
static inline float fp16_value(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu;
  uint32_t u;
  if (exp == 0) {
    if (man == 0) {
      u = sign;
    } else {                                                       // subnormal: man * 2^-24
      float f = (float)man * 5.9604644775390625e-8f;
      __builtin_memcpy(&u, &f, 4);
      u |= sign;
    }
  } else if (exp == 31) {
    u = sign | 0x7f800000u | (man << 13);
  } else {
    u = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  __builtin_memcpy(&f, &u, 4);
  return f;
}

// fn(i) for i in [0, n) on up to `threads` host threads (dynamic scheduling; the caller's thread works
// too).  For independent, infallible tasks -- the per-layer weight folds of v2w_model_pack.
template <typename Fn>
static inline void parallel_for(int n, int threads, Fn fn) {
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i);
  };
  const int nt = threads < 1 ? 1 : (threads > n ? n : threads);
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
}

// Host threads of the weight fold: $V2W_PACK_THREADS, else min(8, hardware threads).
static inline int pack_threads() {
  if (const char* e = std::getenv("V2W_PACK_THREADS")) {
    const int v = std::atoi(e);
    if (v >= 1) return v > 64 ? 64 : v;
  }
  const unsigned hw = std::thread::hardware_concurrency();
  return hw == 0 ? 1 : (hw > 8 ? 8 : (int)hw);
}

constexpr uint32_t kPackedMagic = 0x50573256u;  // "V2WP"

constexpr int kZ = 128;          // z_channels of ConditionalBatchNorm1d (modules.py:9)
constexpr float kBnEps = 1e-5f;  // nn.BatchNorm1d default eps (modules.py:14)
constexpr float kLreluSlope = 0.1f;      // LRELU_SLOPE (models.py:10)
constexpr float kLreluSlopePost = 0.01f; // F.leaky_relu default before conv_post (models.py:143)

struct CondLayout {
  int C = 0;           // channels of this stage's CBN
  int D = 0;           // spk_dim + noise_dim
  int64_t fcw = 0;     // fp32 [128][D]
  int64_t fcb = 0;     // fp32 [128]
  int64_t snw = 0;     // fp32 [2C][128], already divided by sigma
  int64_t snb = 0;     // fp32 [2C]
  int64_t rstd = 0;    // fp32 [C]  1/sqrt(running_var + eps)
  int64_t mean = 0;    // fp32 [C]
};

}  // namespace v2w

struct v2w_model {
  v2w_config cfg;
  std::vector<v2w_layer_desc> layers;
  std::vector<v2w::CondLayout> cond;
  std::vector<v2w_fused_desc> fused;   // per stage; ok == 0 when not fusable
  std::vector<v2w_wide_desc> wide;     // per stage; ok == 0 when the stage cannot run as two stacked tap-GEMMs
  std::vector<v2w_fused_desc> pairs;   // ResBlock1: one per fusable (stage, branch, step) conv pair
  std::vector<int> pair_of_layer;      // layer index of a pair's first conv -> index into `pairs`, else -1
  std::vector<int> stage_channels;  // C_i
  std::vector<int> stage_up;        // cumulative upsampling AFTER stage i
  int total_up = 1;
  int num_convs = 0;
  int64_t packed_bytes = 0;
  int64_t flops_per_frame = 0;
};
