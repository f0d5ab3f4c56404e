// Stand-alone building blocks of the reference's src/models/Hang2020.py for ARBITRARY plane sizes and channel
// counts: global_spectral_pool (:7-12), conv_module (:14-31), Classifier (:55-66), spatial_attention (:68-124),
// spectral_attention (:126-168).  The reference's own tests call these modules on their own
// (tests/test_Hang2020.py:8-33); the networks never do -- their forward/backward is the fused tensor-core pipeline
// of dta_conv_tc.cuh / dta_attention.cuh.  These kernels are therefore plain fp32 CUDA-core code, written for
// exactness and generality (any H x W, any channel count), one pass per tensor, fixed-order reductions (no atomics).
#pragma once
#include "dta_common.cuh"

namespace dta {

constexpr int kBlkThreads = 256;

// Block-wide sum in fp64, fixed order; every thread gets the result.  s_red: >= 32 doubles of shared memory.
__device__ __forceinline__ double blk_sum(double v, double* s_red) {
  v = warp_sum(v);
  __syncthreads();   // s_red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += s_red[i];
  return t;
}

// ---- global_spectral_pool (Hang2020.py:7-12): out[row] = mean_p in[row][p], one warp per (crop, channel) row ----
__global__ void blk_plane_mean_kernel(const float* __restrict__ in, size_t rows, int HW, float* __restrict__ out) {
  const size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* r = in + row * HW;
  float s = 0.f;
  for (int p = lane; p < HW; p += 32) s += r[p];
  s = warp_sum(s);
  if (lane == 0) out[row] = s / (float)HW;
}
__global__ void blk_plane_mean_bwd_kernel(const float* __restrict__ dout, size_t n, int HW, float* __restrict__ din) {
  const float inv = 1.f / (float)HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) din[i] = dout[i / HW] * inv;
}

// ---- conv_module (Hang2020.py:14-31) --------------------------------------------------------------------------
// 3x3 "same" convolution, one thread per output element.
//   transposed = 0: out[b][o][y][x] = bias[o] + sum_{i,dy,dx} in[b][i][y+dy-1][x+dx-1] * w[o][i][dy][dx]     (w: (nout, nin, 3, 3))
//   transposed = 1: out[b][o][y][x] =           sum_{i,dy,dx} in[b][i][y+dy-1][x+dx-1] * w[i][o][2-dy][2-dx] (w: (nin, nout, 3, 3))
// the second form is the input gradient of the first with in = dz.
__global__ void blk_conv3x3_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                                   float* __restrict__ out, int B, int nin, int nout, int H, int W, int transposed) {
  const int HW = H * W;
  const size_t total = (size_t)B * nout * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(idx % HW);
    const int o = (int)((idx / HW) % nout);
    const int b = (int)(idx / ((size_t)HW * nout));
    const int y = p / W, x = p - y * W;
    float acc = (bias != nullptr && !transposed) ? __ldg(bias + o) : 0.f;
    const float* ib = in + (size_t)b * nin * HW;
    for (int i = 0; i < nin; ++i) {
      const float* ip = ib + (size_t)i * HW;
      const float* wp = transposed ? w + ((size_t)i * nout + o) * 9 : w + ((size_t)o * nin + i) * 9;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = x + dx - 1;
          if (xx < 0 || xx >= W) continue;
          const float wv = transposed ? __ldg(wp + (2 - dy) * 3 + (2 - dx)) : __ldg(wp + dy * 3 + dx);
          acc = fmaf(__ldg(ip + yy * W + xx), wv, acc);
        }
      }
    }
    out[idx] = acc;
  }
}

// BatchNorm2d statistics, one CTA per channel (fp64 sums, fixed order).  stat[c] = mean, stat[C + c] = 1/sqrt(var + eps).
// training: batch statistics (biased variance) + running-stat update (momentum 0.1, unbiased variance) + num_batches_tracked;
// eval: the running statistics.
__global__ void __launch_bounds__(kBlkThreads)
blk_bn_stats_kernel(const float* __restrict__ z, int B, int C, int HW, int training, float* __restrict__ rm, float* __restrict__ rv,
                    long long* __restrict__ nbt, float* __restrict__ stat) {
  __shared__ double s_red[32];
  const int c = blockIdx.x;
  if (!training) {
    if (threadIdx.x == 0) {
      stat[c] = rm[c];
      stat[C + c] = (float)(1.0 / sqrt((double)rv[c] + (double)kBnEps));
    }
    return;
  }
  double s = 0.0, q = 0.0;
  const size_t n = (size_t)B * HW;
  for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
    const size_t b = e / HW, p = e - b * HW;
    const double v = (double)z[(b * C + c) * HW + p];
    s += v;
    q += v * v;
  }
  s = blk_sum(s, s_red);
  q = blk_sum(q, s_red);
  if (threadIdx.x != 0) return;
  const double count = (double)n;
  const double mu = s / count;
  double var = q / count - mu * mu;
  if (var < 0.0) var = 0.0;
  stat[c] = (float)mu;
  stat[C + c] = (float)(1.0 / sqrt(var + (double)kBnEps));
  const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
  rm[c] = (float)((1.0 - kBnMomentum) * (double)rm[c] + kBnMomentum * mu);
  rv[c] = (float)((1.0 - kBnMomentum) * (double)rv[c] + kBnMomentum * unbiased);
  if (c == 0 && nbt != nullptr) nbt[0] += 1;
}

__device__ __forceinline__ float blk_bn_relu(float z, float mean, float istd, float gamma, float beta) {
  return fmaxf((z - mean) * istd * gamma + beta, 0.f);
}

// out = [maxpool (ph, pw), stride = kernel, floor]( relu( bn(z) ) ); ph = pw = 1: no pooling.
__global__ void blk_bn_relu_pool_kernel(const float* __restrict__ z, const float* __restrict__ stat, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, int B, int C, int H, int W, int ph, int pw,
                                        float* __restrict__ out) {
  const int Ho = H / ph, Wo = W / pw;
  const size_t total = (size_t)B * C * Ho * Wo;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int px = (int)(idx % Wo);
    const int py = (int)((idx / Wo) % Ho);
    const size_t bc = idx / ((size_t)Wo * Ho);
    const int c = (int)(bc % C);
    const float m = stat[c], is = stat[C + c], g = gamma[c], be = beta[c];
    const float* zp = z + bc * (size_t)H * W;
    float best = -INFINITY;
    for (int dy = 0; dy < ph; ++dy)
      for (int dx = 0; dx < pw; ++dx) best = fmaxf(best, blk_bn_relu(zp[(py * ph + dy) * W + px * pw + dx], m, is, g, be));
    out[idx] = best;
  }
}

// da[b][c][y][x]: gradient w.r.t. the BatchNorm output = dout routed through the max-pool (first maximum of the window in
// row-major order, ATen max_pool2d_with_indices) and masked by the ReLU.  Positions outside every window get 0.
__global__ void blk_relu_pool_bwd_kernel(const float* __restrict__ z, const float* __restrict__ stat, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, const float* __restrict__ dout, int B, int C, int H, int W,
                                         int ph, int pw, float* __restrict__ da) {
  const int Ho = H / ph, Wo = W / pw;
  const size_t total = (size_t)B * C * H * W;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const int y = (int)((idx / W) % H);
    const size_t bc = idx / ((size_t)W * H);
    const int c = (int)(bc % C);
    const int py = y / ph, px = x / pw;
    float g = 0.f;
    if (py < Ho && px < Wo) {
      const float m = stat[c], is = stat[C + c], ga = gamma[c], be = beta[c];
      const float* zp = z + bc * (size_t)H * W;
      float best = -INFINITY;
      int arg = -1;
      for (int dy = 0; dy < ph; ++dy)
        for (int dx = 0; dx < pw; ++dx) {
          const int q = (py * ph + dy) * W + px * pw + dx;
          const float v = blk_bn_relu(zp[q], m, is, ga, be);
          if (v > best) { best = v; arg = q; }
        }
      if (arg == y * W + x && best > 0.f) g = dout[(bc * Ho + py) * Wo + px];
    }
    da[idx] = g;
  }
}

// BatchNorm backward reductions, one CTA per channel: dgamma = sum da*zhat, dbeta = sum da, and the coefficients of
//   dz = k0*da + k1*z + k2   (train: k0 = gamma*istd, k1 = -k0*istd*dgamma/N, k2 = -k0*dbeta/N - k1*mean; eval: k1 = k2 = 0)
// plus the conv bias gradient sum_{b,p} dz (identically 0 under batch statistics).  coef: [3][C].
__global__ void __launch_bounds__(kBlkThreads)
blk_bn_bwd_stats_kernel(const float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ stat,
                        const float* __restrict__ gamma, int B, int C, int HW, int training, float* __restrict__ dgamma,
                        float* __restrict__ dbeta, float* __restrict__ dconv_b, float* __restrict__ coef) {
  __shared__ double s_red[32];
  const int c = blockIdx.x;
  const double mu = stat[c], is = stat[C + c];
  double s1 = 0.0, s2 = 0.0;
  const size_t n = (size_t)B * HW;
  for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
    const size_t b = e / HW, p = e - b * HW;
    const size_t i = (b * C + c) * HW + p;
    const double d = da[i];
    s1 += d;
    s2 += d * ((double)z[i] - mu) * is;
  }
  s1 = blk_sum(s1, s_red);
  s2 = blk_sum(s2, s_red);
  if (threadIdx.x != 0) return;
  const double count = (double)n;
  const double a = (double)gamma[c] * is;
  double b1 = 0.0, b2 = 0.0, dbias;
  if (training) {
    b1 = -a * is * s2 / count;
    b2 = -a * s1 / count - b1 * mu;
    dbias = 0.0;
  } else {
    dbias = a * s1;
  }
  coef[c] = (float)a;
  coef[C + c] = (float)b1;
  coef[2 * C + c] = (float)b2;
  if (dgamma) dgamma[c] = (float)s2;
  if (dbeta) dbeta[c] = (float)s1;
  if (dconv_b) dconv_b[c] = (float)dbias;
}

// dz = k0*da + k1*z + k2, in place over da.
__global__ void blk_bn_dz_kernel(float* __restrict__ da, const float* __restrict__ z, const float* __restrict__ coef, int C, int HW,
                                 size_t total) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((idx / HW) % C);
    da[idx] = coef[c] * da[idx] + coef[C + c] * z[idx] + coef[2 * C + c];
  }
}

// dW[co][ci][dy][dx] = sum_{b,y,x} dz[b][co][y][x] * in[b][ci][y+dy-1][x+dx-1]; one CTA per (ci, co).
__global__ void __launch_bounds__(kBlkThreads)
blk_conv3x3_wgrad_kernel(const float* __restrict__ in, const float* __restrict__ dz, int B, int Cin, int Cout, int H, int W,
                         float* __restrict__ dw) {
  __shared__ double s_red[32];
  const int ci = blockIdx.x, co = blockIdx.y;
  const int HW = H * W;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  const size_t n = (size_t)B * HW;
  for (size_t e = threadIdx.x; e < n; e += blockDim.x) {
    const size_t b = e / HW;
    const int p = (int)(e - b * HW);
    const int y = p / W, x = p - y * W;
    const float d = dz[(b * Cout + co) * HW + p];
    const float* ip = in + (b * Cin + ci) * HW;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int yy = y + dy - 1;
      if (yy < 0 || yy >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int xx = x + dx - 1;
        if (xx < 0 || xx >= W) continue;
        acc[dy * 3 + dx] = fmaf(d, ip[yy * W + xx], acc[dy * 3 + dx]);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const double s = blk_sum((double)acc[t], s_red);
    if (threadIdx.x == 0) dw[((size_t)co * Cin + ci) * 9 + t] = (float)s;
  }
}

// ---- Classifier (Hang2020.py:55-66): scores = feat W^T + b ------------------------------------------------------
__global__ void blk_linear_kernel(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias, int B,
                                  int F, int K, float* __restrict__ out) {
  const size_t total = (size_t)B * K;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / K;
    const int k = (int)(idx - b * K);
    float acc = bias ? bias[k] : 0.f;
    const float* f = feat + b * F;
    const float* wr = w + (size_t)k * F;
    for (int j = 0; j < F; ++j) acc = fmaf(f[j], wr[j], acc);
    out[idx] = acc;
  }
}
// dfeat[b][j] = sum_k dout[b][k] * W[k][j]
__global__ void blk_linear_dinput_kernel(const float* __restrict__ dout, const float* __restrict__ w, int B, int F, int K,
                                         float* __restrict__ dfeat) {
  const size_t total = (size_t)B * F;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t b = idx / F;
    const int j = (int)(idx - b * F);
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(dout[b * K + k], w[(size_t)k * F + j], acc);
    dfeat[idx] = acc;
  }
}

// ---- spectral_attention (Hang2020.py:126-168) -------------------------------------------------------------------
// One CTA per crop.  Conv1d over a length-1 sequence with "same" padding only ever sees its centre tap.
// saved[b] = [g (C) | h1 (C) | s (C)]: squeeze, hidden activation, gate.  Dynamic shared memory: 3*C floats.
__global__ void __launch_bounds__(kBlkThreads)
blk_spectral_fwd_kernel(const float* __restrict__ x, int C, int HW, int ks, const float* __restrict__ w0, const float* __restrict__ b0,
                        const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ out, float* __restrict__ feat,
                        float* __restrict__ saved) {
  extern __shared__ float s_f[];
  float* s_g = s_f;
  float* s_h = s_f + C;
  float* s_s = s_f + 2 * C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const float* xb = x + (size_t)b * C * HW;
  const int mid = ks / 2;
  for (int c = warp; c < C; c += nwarp) {
    float s = 0.f;
    for (int p = lane; p < HW; p += 32) s += xb[(size_t)c * HW + p];
    s = warp_sum(s);
    if (lane == 0) s_g[c] = s / (float)HW;
  }
  __syncthreads();
  for (int i = tid; i < C; i += blockDim.x) {
    float u = b0[i];
    for (int j = 0; j < C; ++j) u = fmaf(w0[((size_t)i * C + j) * ks + mid], s_g[j], u);
    s_h[i] = fmaxf(u, 0.f);
  }
  __syncthreads();
  for (int i = tid; i < C; i += blockDim.x) {
    float u = b1[i];
    for (int j = 0; j < C; ++j) u = fmaf(w1[((size_t)i * C + j) * ks + mid], s_h[j], u);
    s_s[i] = sigmoidf_acc(u);
  }
  __syncthreads();
  float* ob = out + (size_t)b * C * HW;
  for (int c = warp; c < C; c += nwarp) {
    const float g = s_s[c];
    float s = 0.f;
    for (int p = lane; p < HW; p += 32) {
      const float o = xb[(size_t)c * HW + p] * g;
      ob[(size_t)c * HW + p] = o;
      s += o;
    }
    s = warp_sum(s);
    if (lane == 0) feat[(size_t)b * C + c] = s / (float)HW;
  }
  for (int i = tid; i < 3 * C; i += blockDim.x) saved[(size_t)b * 3 * C + i] = s_f[i];
}

// dout: gradient of the gated map (may be null), dfeat: gradient of the pooled features (may be null).
// prow[b] = [du2 (C) | du1 (C)]: pre-activation gradients of the second / first Conv1d, reduced over the batch afterwards.
// Dynamic shared memory: 6*C floats.
__global__ void __launch_bounds__(kBlkThreads)
blk_spectral_bwd_kernel(const float* __restrict__ x, int C, int HW, int ks, const float* __restrict__ w0, const float* __restrict__ w1,
                        const float* __restrict__ saved, const float* __restrict__ dout, const float* __restrict__ dfeat,
                        float* __restrict__ dx, float* __restrict__ prow) {
  extern __shared__ float s_f[];
  float* s_h = s_f;           // h1
  float* s_s = s_f + C;       // gate
  float* s_ds = s_f + 2 * C;  // d gate
  float* s_u2 = s_f + 3 * C;
  float* s_u1 = s_f + 4 * C;
  float* s_dg = s_f + 5 * C;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const float* xb = x + (size_t)b * C * HW;
  const float* sv = saved + (size_t)b * 3 * C;
  const int mid = ks / 2;
  const float inv = 1.f / (float)HW;
  for (int i = tid; i < C; i += blockDim.x) { s_h[i] = sv[C + i]; s_s[i] = sv[2 * C + i]; }
  for (int c = warp; c < C; c += nwarp) {
    const float df = dfeat ? dfeat[(size_t)b * C + c] * inv : 0.f;
    float s = 0.f;
    for (int p = lane; p < HW; p += 32) {
      const size_t i = (size_t)c * HW + p;
      const float g = (dout ? dout[(size_t)b * C * HW + i] : 0.f) + df;
      s = fmaf(g, xb[i], s);
    }
    s = warp_sum(s);
    if (lane == 0) s_ds[c] = s;
  }
  __syncthreads();
  for (int i = tid; i < C; i += blockDim.x) s_u2[i] = s_ds[i] * s_s[i] * (1.f - s_s[i]);
  __syncthreads();
  for (int j = tid; j < C; j += blockDim.x) {
    float d = 0.f;
    for (int i = 0; i < C; ++i) d = fmaf(w1[((size_t)i * C + j) * ks + mid], s_u2[i], d);
    s_u1[j] = s_h[j] > 0.f ? d : 0.f;
  }
  __syncthreads();
  for (int j = tid; j < C; j += blockDim.x) {
    float d = 0.f;
    for (int i = 0; i < C; ++i) d = fmaf(w0[((size_t)i * C + j) * ks + mid], s_u1[i], d);
    s_dg[j] = d * inv;
  }
  __syncthreads();
  if (dx != nullptr) {
    float* db = dx + (size_t)b * C * HW;
    for (int c = warp; c < C; c += nwarp) {
      const float df = dfeat ? dfeat[(size_t)b * C + c] * inv : 0.f;
      const float g = s_s[c], add = s_dg[c];
      for (int p = lane; p < HW; p += 32) {
        const size_t i = (size_t)c * HW + p;
        db[i] = ((dout ? dout[(size_t)b * C * HW + i] : 0.f) + df) * g + add;
      }
    }
  }
  for (int i = tid; i < C; i += blockDim.x) {
    prow[(size_t)b * 2 * C + i] = s_u2[i];
    prow[(size_t)b * 2 * C + C + i] = s_u1[i];
  }
}

// ---- spatial_attention (Hang2020.py:68-124) ----------------------------------------------------------------------
// k x k "same" stencil over one plane held in shared memory (zero padding).
__device__ __forceinline__ float blk_stencil(const float* s_plane, const float* __restrict__ w, int ks, int H, int W, int y, int x) {
  const int r = ks / 2;
  float acc = 0.f;
  for (int dy = 0; dy < ks; ++dy) {
    const int yy = y + dy - r;
    if (yy < 0 || yy >= H) continue;
    for (int dx = 0; dx < ks; ++dx) {
      const int xx = x + dx - r;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(w[dy * ks + dx], s_plane[yy * W + xx], acc);
    }
  }
  return acc;
}
// transpose of the stencil: sum_tap w[tap] * plane[p - (tap - r)]
__device__ __forceinline__ float blk_stencil_t(const float* s_plane, const float* __restrict__ w, int ks, int H, int W, int y, int x) {
  const int r = ks / 2;
  float acc = 0.f;
  for (int dy = 0; dy < ks; ++dy) {
    const int yy = y - dy + r;
    if (yy < 0 || yy >= H) continue;
    for (int dx = 0; dx < ks; ++dx) {
      const int xx = x - dx + r;
      if (xx < 0 || xx >= W) continue;
      acc = fmaf(w[dy * ks + dx], s_plane[yy * W + xx], acc);
    }
  }
  return acc;
}

// One CTA per crop.  saved[b] = [q (HW) | t (HW) | s (HW)].  feat[b] = flatten_{c,i,j} maxpool_P(out), floor.
// Dynamic shared memory: 3*HW floats.
__global__ void __launch_bounds__(kBlkThreads)
blk_spatial_fwd_kernel(const float* __restrict__ x, int C, int H, int W, int ks, int P, const float* __restrict__ pool_w,
                       const float* __restrict__ pool_b, const float* __restrict__ w0, const float* __restrict__ b0,
                       const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ out, float* __restrict__ feat,
                       float* __restrict__ saved) {
  extern __shared__ float s_f[];
  const int HW = H * W;
  float* s_q = s_f;
  float* s_t = s_f + HW;
  float* s_s = s_f + 2 * HW;
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xb = x + (size_t)b * C * HW;
  for (int p = tid; p < HW; p += blockDim.x) {
    float u = pool_b[0];
    for (int c = 0; c < C; ++c) u = fmaf(pool_w[c], xb[(size_t)c * HW + p], u);
    s_q[p] = fmaxf(u, 0.f);
  }
  __syncthreads();
  for (int p = tid; p < HW; p += blockDim.x) s_t[p] = fmaxf(b0[0] + blk_stencil(s_q, w0, ks, H, W, p / W, p % W), 0.f);
  __syncthreads();
  for (int p = tid; p < HW; p += blockDim.x) s_s[p] = sigmoidf_acc(b1[0] + blk_stencil(s_t, w1, ks, H, W, p / W, p % W));
  __syncthreads();
  float* ob = out + (size_t)b * C * HW;
  for (int i = tid; i < C * HW; i += blockDim.x) ob[i] = xb[i] * s_s[i % HW];
  const int Ho = H / P, Wo = W / P;
  float* fb = feat + (size_t)b * C * Ho * Wo;
  for (int i = tid; i < C * Ho * Wo; i += blockDim.x) {
    const int px = i % Wo, py = (i / Wo) % Ho, c = i / (Wo * Ho);
    float best = -INFINITY;
    for (int dy = 0; dy < P; ++dy)
      for (int dx = 0; dx < P; ++dx) {
        const int q = (py * P + dy) * W + px * P + dx;
        best = fmaxf(best, xb[(size_t)c * HW + q] * s_s[q]);
      }
    fb[i] = best;
  }
  for (int i = tid; i < 3 * HW; i += blockDim.x) saved[(size_t)b * 3 * HW + i] = s_f[i];
}

// Upstream gradient of the gated map at (c, p): dout plus dfeat routed to the first maximum of its class-pool window.
__device__ __forceinline__ float blk_spatial_gout(const float* __restrict__ xb, const float* s_s, const float* __restrict__ doutb,
                                                  const float* __restrict__ dfeatb, int c, int p, int H, int W, int P) {
  const int HW = H * W;
  float g = doutb ? doutb[(size_t)c * HW + p] : 0.f;
  if (dfeatb != nullptr) {
    const int Ho = H / P, Wo = W / P;
    const int y = p / W, x = p - y * W;
    const int py = y / P, px = x / P;
    if (py < Ho && px < Wo) {
      float best = -INFINITY;
      int arg = -1;
      for (int dy = 0; dy < P; ++dy)
        for (int dx = 0; dx < P; ++dx) {
          const int q = (py * P + dy) * W + px * P + dx;
          const float v = xb[(size_t)c * HW + q] * s_s[q];
          if (v > best) { best = v; arg = q; }
        }
      if (arg == p) g += dfeatb[((size_t)c * Ho + py) * Wo + px];
    }
  }
  return g;
}

// prow[b] = [dA1 (k*k) | db0 | dA2 (k*k) | db1 | dpool_w (C) | dpool_b]: per-crop parameter gradients, summed over the batch
// afterwards.  Dynamic shared memory: 6*HW floats.
__global__ void __launch_bounds__(kBlkThreads)
blk_spatial_bwd_kernel(const float* __restrict__ x, int C, int H, int W, int ks, int P, const float* __restrict__ pool_w,
                       const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ saved,
                       const float* __restrict__ dout, const float* __restrict__ dfeat, float* __restrict__ dx,
                       float* __restrict__ prow) {
  extern __shared__ float s_f[];
  const int HW = H * W, kk = ks * ks, r = ks / 2;
  float* s_q = s_f;
  float* s_t = s_f + HW;
  float* s_s = s_f + 2 * HW;
  float* s_v2 = s_f + 3 * HW;   // gradient at the pre-activation of the sigmoid stencil
  float* s_v1 = s_f + 4 * HW;   // ... of the first stencil
  float* s_dq = s_f + 5 * HW;   // ... of the channel pool
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* xb = x + (size_t)b * C * HW;
  const float* doutb = dout ? dout + (size_t)b * C * HW : nullptr;
  const int Ho = H / P, Wo = W / P;
  const float* dfeatb = dfeat ? dfeat + (size_t)b * C * Ho * Wo : nullptr;
  for (int i = tid; i < 3 * HW; i += blockDim.x) s_f[i] = saved[(size_t)b * 3 * HW + i];
  __syncthreads();
  for (int p = tid; p < HW; p += blockDim.x) {
    float ds = 0.f;
    for (int c = 0; c < C; ++c) ds = fmaf(blk_spatial_gout(xb, s_s, doutb, dfeatb, c, p, H, W, P), xb[(size_t)c * HW + p], ds);
    s_v2[p] = ds * s_s[p] * (1.f - s_s[p]);
  }
  __syncthreads();
  for (int p = tid; p < HW; p += blockDim.x) {
    const float dt = blk_stencil_t(s_v2, w1, ks, H, W, p / W, p % W);
    s_v1[p] = s_t[p] > 0.f ? dt : 0.f;
  }
  __syncthreads();
  for (int p = tid; p < HW; p += blockDim.x) {
    const float dq = blk_stencil_t(s_v1, w0, ks, H, W, p / W, p % W);
    s_dq[p] = s_q[p] > 0.f ? dq : 0.f;
  }
  __syncthreads();
  if (dx != nullptr) {
    float* db = dx + (size_t)b * C * HW;
    for (int i = tid; i < C * HW; i += blockDim.x) {
      const int c = i / HW, p = i - c * HW;
      db[i] = blk_spatial_gout(xb, s_s, doutb, dfeatb, c, p, H, W, P) * s_s[p] + pool_w[c] * s_dq[p];
    }
  }
  float* pr = prow + (size_t)b * (2 * kk + 2 + C + 1);
  // stencil weight gradients: dA[tap] = sum_p v[p] * src[p + tap - r]
  for (int i = tid; i < 2 * kk; i += blockDim.x) {
    const bool second = i >= kk;
    const int tap = second ? i - kk : i;
    const int dy = tap / ks, dxx = tap - dy * ks;
    const float* v = second ? s_v2 : s_v1;
    const float* src = second ? s_t : s_q;
    float acc = 0.f;
    for (int y = 0; y < H; ++y) {
      const int yy = y + dy - r;
      if (yy < 0 || yy >= H) continue;
      for (int xq = 0; xq < W; ++xq) {
        const int xx = xq + dxx - r;
        if (xx < 0 || xx >= W) continue;
        acc = fmaf(v[y * W + xq], src[yy * W + xx], acc);
      }
    }
    pr[second ? kk + 1 + tap : tap] = acc;
  }
  for (int c = tid; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc = fmaf(s_dq[p], xb[(size_t)c * HW + p], acc);
    pr[2 * kk + 2 + c] = acc;
  }
  if (tid < 3) {
    const float* v = tid == 0 ? s_v1 : (tid == 1 ? s_v2 : s_dq);
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += v[p];
    pr[tid == 0 ? kk : (tid == 1 ? 2 * kk + 1 : 2 * kk + 2 + C)] = acc;
  }
}

// ---- batch reductions of per-crop rows into parameter gradients (fixed order over the batch) -----------------------
// out[i*si + j*sj] = sum_b U[b*ldu + i] * V[b*ldv + j]   (V == nullptr: sum_b U[b*ldu + i], nj = 1)
__global__ void blk_batch_sum_kernel(const float* __restrict__ U, size_t ldu, const float* __restrict__ V, size_t ldv, int B, int ni,
                                     int nj, float* __restrict__ out, size_t si, size_t sj) {
  const size_t total = (size_t)ni * nj;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / nj, j = idx - i * nj;
    float acc = 0.f;
    if (V != nullptr) {
      for (int b = 0; b < B; ++b) acc = fmaf(U[(size_t)b * ldu + i], V[(size_t)b * ldv + j], acc);
    } else {
      for (int b = 0; b < B; ++b) acc += U[(size_t)b * ldu + i];
    }
    out[i * si + j * sj] = acc;
  }
}

// ---- year ensemble on the device (src/models/year.py:24-33, src/models/multi_stage.py:302,314) ----------------------------
constexpr int kMaxYears = 16;
constexpr int kYearSumBlocks = 1024;   // partial sums per year (fixed order)
struct YearPtrs {
  const float* p[kMaxYears];
};

// partial[y][blk] = sum of one grid-stride slice of crops[y]
__global__ void __launch_bounds__(kBlkThreads) crops_sum_partial_kernel(YearPtrs crops, size_t elems, float* __restrict__ partial) {
  __shared__ float s_w[kBlkThreads / 32];
  const float* src = crops.p[blockIdx.y];
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (elems & 3u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems / 4; i += stride) {
      const float4 v = __ldg(s4 + i);
      s += (v.x + v.y) + (v.z + v.w);
    }
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < elems; i += stride) s += __ldg(src + i);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < kBlkThreads / 32; ++i) t += s_w[i];
    partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
  }
}
// one warp per year: flags[y] = (sum != 0)
__global__ void crops_nonzero_finish_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ flags) {
  const int y = blockIdx.x, lane = threadIdx.x;
  float s = 0.f;
  for (int i = lane; i < nblk; i += 32) s += partial[(size_t)y * nblk + i];
  s = warp_sum(s);
  if (lane == 0) flags[y] = (s != 0.f) ? 1.f : 0.f;
}

// One warp per crop row: mean of the active years' scores, optional row softmax.
__global__ void __launch_bounds__(kBlkThreads)
ensemble_mean_kernel(YearPtrs scores, const float* __restrict__ flags, int n, int B, int K, int softmax, float* __restrict__ out) {
  const int row = blockIdx.x * (kBlkThreads / 32) + (threadIdx.x >> 5);
  if (row >= B) return;
  const int lane = threadIdx.x & 31;
  float count = 0.f;
  for (int y = 0; y < n; ++y) count += (flags == nullptr || flags[y] != 0.f) ? 1.f : 0.f;
  float* o = out + (size_t)row * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) {
    float acc = 0.f;
    for (int y = 0; y < n; ++y)
      if (flags == nullptr || flags[y] != 0.f) acc += scores.p[y][(size_t)row * K + k];
    const float v = acc / count;
    o[k] = v;
    mx = fmaxf(mx, v);
  }
  if (!softmax) return;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float e = expf(o[k] - mx);
    o[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  for (int k = lane; k < K; k += 32) o[k] = o[k] / sum;
}

// ---- fused Adam over a table of parameter tensors (torch.optim.Adam as configured in src/main.py:135-136 and
// src/models/multi_stage.py:258-262: lr from the config, betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) ------------
constexpr int kAdamMaxTensors = 96;
constexpr int kAdamChunk = 2048;   // elements per CTA
struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  int chunk_begin[kAdamMaxTensors + 1];   // first CTA of tensor i
  int elem_begin[kAdamMaxTensors];        // offset of tensor i in the flat moment buffers
  int numel[kAdamMaxTensors];
  int n;
};
struct AdamHyper {
  float lr, beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay;   // float32 roundings of the host doubles, as torch passes them
  double lr64, beta1_64, beta2_64, eps64;
  long long step;   // 1-based step number when step_dev is null
};

__global__ void adam_tick_kernel(long long* step_dev) { step_dev[0] += 1; }

// torch.optim.Adam single-tensor arithmetic (torch/optim/adam.py _single_tensor_adam):
//   m <- lerp(m, g, 1-b1);  v <- v*b2 + (1-b2)*g*g;  p <- p - (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps),  bc_i = 1 - b_i^step
__global__ void __launch_bounds__(kBlkThreads)
adam_step_kernel(const __grid_constant__ AdamTable tab, AdamHyper h, const long long* __restrict__ step_dev,
                 const float* __restrict__ lr_dev, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, double* p64,
                 const double* g64, double* m64 /*[2]: exp_avg, exp_avg_sq of the float64 scalar*/) {
  // which tensor does this CTA work on: binary search over the chunk prefix
  int lo = 0, hi = tab.n - 1;
  const int blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.chunk_begin[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  const double step = (double)(step_dev ? step_dev[0] : h.step);
  const double lr = lr_dev ? (double)lr_dev[0] : h.lr64;
  const double bc1 = 1.0 - pow(h.beta1_64, step);
  const double bc2 = 1.0 - pow(h.beta2_64, step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  if (tab.n > 0 && blk < tab.chunk_begin[tab.n]) {
    const int numel = tab.numel[lo];
    const int e0 = (blk - tab.chunk_begin[lo]) * kAdamChunk;
    float* p = tab.p[lo];
    const float* g = tab.g[lo];
    float* m = exp_avg + tab.elem_begin[lo];
    float* v = exp_avg_sq + tab.elem_begin[lo];
    for (int e = e0 + threadIdx.x; e < e0 + kAdamChunk && e < numel; e += blockDim.x) {
      float gr = g[e];
      const float pv = p[e];
      if (h.weight_decay != 0.f) gr = fmaf(h.weight_decay, pv, gr);
      float mv = m[e], vv = v[e];
      mv = mv + h.one_minus_beta1 * (gr - mv);
      vv = vv * h.beta2 + h.one_minus_beta2 * gr * gr;
      m[e] = mv;
      v[e] = vv;
      const float denom = sqrtf(vv) / bc2_sqrt + h.eps;
      p[e] = pv - step_size * (mv / denom);
    }
  }
  if (blk == 0 && threadIdx.x == 0 && p64 != nullptr && g64 != nullptr) {   // Hang2020.alpha: the one float64 parameter
    double gr = g64[0];
    const double pv = p64[0];
    if (h.weight_decay != 0.f) gr += (double)h.weight_decay * pv;
    const double mv = m64[0] + (1.0 - h.beta1_64) * (gr - m64[0]);
    const double vv = m64[1] * h.beta2_64 + (1.0 - h.beta2_64) * gr * gr;
    m64[0] = mv;
    m64[1] = vv;
    p64[0] = pv - (lr / bc1) * (mv / (sqrt(vv) / sqrt(bc2) + h.eps64));
  }
}

}  // namespace dta
