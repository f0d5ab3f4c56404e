// Crop preprocessing on the device: raw int16 crown crops -> the float32 tensors the networks take.
// Reference: utils.preprocess_image (src/utils.py:36-57): drop the first and last 10 bands, cast to float32,
// per-pixel min-max scaling over the bands with sklearn.preprocessing.minmax_scale(data, axis=1), i.e. in float32
//     scale = 1 / range (range < 10*eps -> 1),  offset = 0 - min*scale,  out = x*scale + offset   (two roundings).
// The same IEEE operations are issued here (no FMA contraction), so the result is bit-identical.
#pragma once
#include "dta_common.cuh"

namespace dta {

constexpr int kPrepSlices = 4;   // band slices per pixel in the min/max pass

// grid = crops, block = 128 pixels x kPrepSlices.  raw (B, C, HW) int16, out (B, C - 2*clip, HW) float32.
__global__ void __launch_bounds__(128 * kPrepSlices)
preprocess_crops_kernel(const short* __restrict__ raw, int C, int HW, int clip, float* __restrict__ out) {
  pdl_prologue();
  __shared__ float s_min[kPrepSlices][128], s_max[kPrepSlices][128];
  const int p = threadIdx.x & 127, sl = threadIdx.x >> 7;
  const int b = blockIdx.x;
  const int c_lo = clip, c_hi = C - clip;
  const short* src = raw + (size_t)b * C * HW;
  float mn = INFINITY, mx = -INFINITY;
  if (p < HW) {
    for (int c = c_lo + sl; c < c_hi; c += kPrepSlices) {
      const float v = (float)__ldg(src + (size_t)c * HW + p);
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  s_min[sl][p] = mn;
  s_max[sl][p] = mx;
  __syncthreads();
  if (p >= HW) return;
#pragma unroll
  for (int k = 0; k < kPrepSlices; ++k) { mn = fminf(mn, s_min[k][p]); mx = fmaxf(mx, s_max[k][p]); }
  float range = __fsub_rn(mx, mn);
  if (range < 10.f * 1.1920929e-07f) range = 1.f;          // sklearn _handle_zeros_in_scale for float32
  const float scale = __fdiv_rn(1.f, range);
  const float offset = __fsub_rn(0.f, __fmul_rn(mn, scale));
  float* dst = out + (size_t)b * (c_hi - c_lo) * HW;
  for (int c = c_lo + sl; c < c_hi; c += kPrepSlices) {
    const float v = (float)__ldg(src + (size_t)c * HW + p);
    dst[(size_t)(c - c_lo) * HW + p] = __fadd_rn(__fmul_rn(v, scale), offset);
  }
}

}  // namespace dta
