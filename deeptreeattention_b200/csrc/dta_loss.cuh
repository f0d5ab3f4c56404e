// Weighted cross-entropy over several classifier heads in one pass: the loss either side of the
// hot path.  Reference: TreeModel.training_step  F.cross_entropy(y_hat, y, weight=loss_weight)
// (src/main.py:78, weights :66-69), i.e. per head  sum_b w[y_b] * (-log softmax(s_b)[y_b]) / sum_b w[y_b];
// the north-star regime sums that over the heads.  Produces the loss AND d(loss)/d(scores), so the
// backward pass needs no further kernels.
#pragma once
#include "dta_common.cuh"

namespace dta {

struct HeadPtrs {
  const float* s[8];
  float* ds[8];
};

// den = sum_b w[y_b] over the valid labels, computed by every CTA for itself (fp64, fixed order: every CTA gets the same
// bits) instead of by a one-block kernel in front: one launch less on the step's critical path.
__device__ __forceinline__ double ce_den_block(const long long* __restrict__ y, const float* __restrict__ w, int B, int classes, double* s_red) {
  double a = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long c = y[b];
    if (c >= 0 && c < classes) a += w ? (double)__ldg(w + c) : 1.0;
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = a;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_red[i];
  return t;
}

// One warp: stable log-softmax of one score row, its weighted NLL term (returned in every lane) and, through emit(c, v), the
// gradient of  sum_b w[y_b] nll_b / dsum  with respect to the row.  Shared by ce_rows_kernel and by the attention-backward
// kernel of a fused training step (dta_train_step), which therefore produce the same bits.
template <class Emit>
__device__ __forceinline__ float ce_row_warp(const float* __restrict__ s, int classes, long long yc, const float* __restrict__ w,
                                             double dsum, bool want_grad, Emit emit) {
  const int lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int c = lane; c < classes; c += 32) m = fmaxf(m, __ldg(s + c));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float z = 0.f;
  for (int c = lane; c < classes; c += 32) z = __fadd_rn(z, expf(__fsub_rn(__ldg(s + c), m)));
  z = warp_sum(z);
  const bool ok = yc >= 0 && yc < classes;
  const float wy = ok ? (w ? __ldg(w + yc) : 1.f) : 0.f;
  const float lse = __fadd_rn(m, logf(z));
  if (want_grad) {
    const float k = __fmul_rn(wy, (float)(1.0 / dsum));
    for (int c = lane; c < classes; c += 32) {
      const float p = expf(__fsub_rn(__ldg(s + c), lse));
      emit(c, __fmul_rn(k, __fsub_rn(p, c == yc ? 1.f : 0.f)));
    }
  }
  return ok ? __fmul_rn(wy, __fsub_rn(lse, __ldg(s + yc))) : 0.f;
}

// den = sum_b w[y_b] as its own (one-CTA) launch: a fused training step computes it beside the forward pass so that the
// attention-backward kernel of block 3 can form its heads' score gradients itself.  Same bits as ce_den_block anywhere else.
__global__ void __launch_bounds__(256) ce_den_kernel(const long long* __restrict__ y, const float* __restrict__ w, int B, int classes, double* __restrict__ den) {
  pdl_prologue();
  __shared__ double s_red[8];
  const double d = ce_den_block(y, w, B, classes, s_red);
  if (threadIdx.x == 0) den[0] = d;
}

// One warp per (head, crop).  The CTA that finishes LAST (ticket counter in context-owned memory, reset by that CTA) also sums
// the rows into the losses -- fixed order, fp64 -- so the loss costs one launch on the step's critical path instead of two.
__global__ void __launch_bounds__(256)
ce_rows_kernel(HeadPtrs h, int n_heads, const long long* __restrict__ y, const float* __restrict__ w, int B, int classes,
               float* __restrict__ row_loss /*[n_heads][B]*/, float* __restrict__ loss /*[n_heads + 1]*/, unsigned int* __restrict__ ticket) {
  pdl_prologue();
  __shared__ double s_red[8];
  __shared__ double s_head[8];
  __shared__ unsigned int s_last;
  const double dsum = ce_den_block(y, w, B, classes, s_red);
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw < n_heads * B) {
    const int head = gw / B, b = gw - head * B;
    float* ds = h.ds[head];
    const float rl = ce_row_warp(h.s[head] + (size_t)b * classes, classes, y[b], w, dsum, ds != nullptr,
                                 [&](int c, float v) { ds[(size_t)b * classes + c] = v; });
    if (lane == 0) row_loss[(size_t)head * B + b] = rl;
  }
  // ---- last CTA: loss[head] = sum_b row_loss / den (fp64, fixed order), loss[n_heads] = sum over heads ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int warp = threadIdx.x >> 5;
  if (warp < n_heads) {
    double a = 0.0;
#pragma unroll 4
    for (int b = lane; b < B; b += 32) a += (double)__ldcg(row_loss + (size_t)warp * B + b);
    a = warp_sum(a);
    if (lane == 0) {
      const double v = a / dsum;
      loss[warp] = (float)v;
      s_head[warp] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int hd = 0; hd < n_heads; ++hd) total += s_head[hd];
    loss[n_heads] = (float)total;
    *ticket = 0u;
  }
}

// What the attention-backward kernel needs to form its heads' score gradients from the scores themselves (fused training
// step, block 3): scores == nullptr for a branch = read the gradient from memory as usual.
struct CeInline {
  const float* scores[2];     // per branch: (B, classes) scores of this block's head
  const long long* y;         // (B) labels
  const float* w;             // (classes) class weights or nullptr
  const double* den;          // sum_b w[y_b] (ce_den_kernel)
};

}  // namespace dta
