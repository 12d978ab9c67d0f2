// Internal declarations shared by the host (model/pack/plan) and device (kernels) halves of
// libv2w.  Nothing here is part of the ABI; the ABI is include/v2w.h.
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
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
  std::vector<int> stage_channels;  // C_i
  std::vector<int> stage_up;        // cumulative upsampling AFTER stage i
  int total_up = 1;
  int num_convs = 0;
  int64_t packed_bytes = 0;
  int64_t flops_per_frame = 0;
};
