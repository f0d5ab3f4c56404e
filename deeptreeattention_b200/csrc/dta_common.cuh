// Shared declarations for the Hang2020 hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dta_b200.h"

namespace dta {

constexpr int kSide = DTA_IMAGE_SIZE;  // 11
constexpr int kHW = kSide * kSide;     // 121
constexpr float kBnEps = 1e-5f;        // nn.BatchNorm2d default (Hang2020.py:19)
constexpr float kBnMomentum = 0.1f;
constexpr int kTcGuard = 16;           // zero rows before the first / after the last crop of a packed position stream (dta_conv_tc.cuh)

// Attention flavour of a branch.
enum BranchType : int { BR_NONE = 0, BR_SPECTRAL = 1, BR_SPATIAL = 2 };

// How a convolution kernel produces its input element on the fly.
enum SrcMode : int {
  SRC_RAW = 0,  // plain tensor (the crops, conv1)
  SRC_ACT = 1,  // relu(bn(z)) [2x2 max-pooled] * attention gate   (input of conv2 / conv3)
  SRC_DZ = 2    // BatchNorm backward on the fly: k0*da + k1*z + k2 (input of dgrad / wgrad)
};

struct ConvSrc {
  int mode;
  int cin;        // channels per group
  int ctot;       // channels per crop in the source tensor
  int src_hw;     // positions per channel plane in the source tensor
  int pool;       // SRC_ACT: source plane has side 2*S+1 and is max-pooled 2x2 (floor)
  const float* a; // RAW: x    ACT: z     DZ: da
  const float* b; //                      DZ: z
  const float* k0;
  const float* k1;
  const float* k2;
  const unsigned char* arg;   // SRC_DZ with pool = 1: `a` holds one value per POOLED cell ([B][ctot][SP*SP], SP = S/2) and `arg` the
                              // slot (0..3) of its 2x2 window that the value belongs to -- da is zero everywhere else
  const float* gate;  // SRC_ACT: attention rows [B][G][gate_ld], gate value at gate_off + (c | p)
  int gate_ld;
  int gate_off;
  int gate_mode[2];   // per group: BranchType
};

struct Ptr2 {
  const float* p[2];
};
struct MutPtr2 {
  float* p[2];
};

// Programmatic dependent launch: first statement of every kernel.  launch_dependents lets the NEXT kernel of the stream be
// scheduled as soon as every CTA of this grid has started (its CTAs then park in their own wait); wait blocks until the
// PREVIOUS grid has completed and its writes are visible, so stream order is preserved exactly -- only the launch latency
// between the ~60 short kernels of a step is hidden.  Both are no-ops for a launch without the attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ float sigmoidf_acc(float v) { return 1.0f / (1.0f + expf(-v)); }

// Fetch conv input element (crop b, group g, channel ci within group, position p on an
// S x S plane).  G = number of groups (for the gate row index).
template <int S>
__device__ __forceinline__ float conv_src_load(const ConvSrc& s, int b, int g, int G, int ci, int p) {
  const int ch = g * s.cin + ci;
  if (s.mode == SRC_RAW) {
    return __ldg(s.a + ((size_t)b * s.ctot + ch) * s.src_hw + p);
  } else if (s.mode == SRC_DZ) {
    const size_t idx = ((size_t)b * s.ctot + ch) * s.src_hw + p;
    return __ldg(s.k0 + ch) * __ldg(s.a + idx) + __ldg(s.k1 + ch) * __ldg(s.b + idx) + __ldg(s.k2 + ch);
  } else {
    const float sc = __ldg(s.k0 + ch), sh = __ldg(s.k1 + ch);
    const float* zp = s.a + ((size_t)b * s.ctot + ch) * s.src_hw;
    float v;
    if (s.pool) {
      constexpr int SP = 2 * S + 1;
      const int y = p / S, x = p - y * S;
      const float* q = zp + (2 * y) * SP + 2 * x;
      const float v0 = fmaxf(fmaf(__ldg(q), sc, sh), 0.f);
      const float v1 = fmaxf(fmaf(__ldg(q + 1), sc, sh), 0.f);
      const float v2 = fmaxf(fmaf(__ldg(q + SP), sc, sh), 0.f);
      const float v3 = fmaxf(fmaf(__ldg(q + SP + 1), sc, sh), 0.f);
      v = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
    } else {
      v = fmaxf(fmaf(__ldg(zp + p), sc, sh), 0.f);
    }
    const int gm = s.gate_mode[g];
    if (gm == BR_SPECTRAL) {
      v *= __ldg(s.gate + ((size_t)b * G + g) * s.gate_ld + s.gate_off + ci);
    } else if (gm == BR_SPATIAL) {
      v *= __ldg(s.gate + ((size_t)b * G + g) * s.gate_ld + s.gate_off + p);
    }
    return v;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dta
