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

// den = sum_b w[y_b]  (fp64, fixed order; one block)
__global__ void ce_den_kernel(const long long* __restrict__ y, const float* __restrict__ w, int B, int classes, double* __restrict__ den,
                              int* __restrict__ bad_label) {
  pdl_prologue();
  __shared__ double s[256];
  double a = 0.0;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const long long c = y[b];
    if (c < 0 || c >= classes) { *bad_label = 1; continue; }
    a += w ? (double)w[c] : 1.0;
  }
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) den[0] = s[0];
}

// One warp per (head, crop): stable log-softmax, weighted NLL term and the score gradient.
__global__ void ce_rows_kernel(HeadPtrs h, int n_heads, const long long* __restrict__ y, const float* __restrict__ w, int B, int classes,
                               const double* __restrict__ den, float* __restrict__ row_loss /*[n_heads][B]*/) {
  pdl_prologue();
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_heads * B) return;
  const int head = gw / B, b = gw - head * B;
  const float* s = h.s[head] + (size_t)b * classes;
  float m = -INFINITY;
  for (int c = lane; c < classes; c += 32) m = fmaxf(m, __ldg(s + c));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float z = 0.f;
  for (int c = lane; c < classes; c += 32) z += expf(__ldg(s + c) - m);
  z = warp_sum(z);
  const long long yc = y[b];
  const bool ok = yc >= 0 && yc < classes;
  const float wy = ok ? (w ? __ldg(w + yc) : 1.f) : 0.f;
  const float lse = m + logf(z);
  if (lane == 0) row_loss[(size_t)head * B + b] = ok ? wy * (lse - __ldg(s + yc)) : 0.f;
  float* ds = h.ds[head];
  if (ds != nullptr) {
    const float inv = (float)(1.0 / den[0]);
    for (int c = lane; c < classes; c += 32) {
      const float p = expf(__ldg(s + c) - lse);
      ds[(size_t)b * classes + c] = wy * inv * (p - (c == yc ? 1.f : 0.f));
    }
  }
}

// loss[head] = sum_b row_loss / den (fp64, fixed order), loss[n_heads] = sum over heads.  One block.
__global__ void ce_finish_kernel(const float* __restrict__ row_loss, int n_heads, int B, const double* __restrict__ den, float* __restrict__ loss) {
  pdl_prologue();
  __shared__ double s[256];
  __shared__ double total;
  if (threadIdx.x == 0) total = 0.0;
  for (int head = 0; head < n_heads; ++head) {
    double a = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) a += (double)row_loss[(size_t)head * B + b];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const double v = s[0] / den[0];
      loss[head] = (float)v;
      total += v;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[n_heads] = (float)total;
}

}  // namespace dta
