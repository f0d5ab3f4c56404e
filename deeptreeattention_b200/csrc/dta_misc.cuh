// Small kernels around the hot stages: parameter packing, BatchNorm statistics
// finalisation (forward and backward), deterministic batch reductions for the small
// parameter gradients, and the alpha blend of Hang2020.forward (Hang2020.py:256-261).
#pragma once
#include "dta_common.cuh"

namespace dta {

// Wp[g][ci][tap][co]: forward weight table.  merged=1: one group whose output channels are
// the concatenation of both branches (conv1 reads the same crops for both branches).
__global__ void pack_conv_w_kernel(Ptr2 w, int nb, int cout_b, int cin, int merged, float* __restrict__ wp) {
  pdl_prologue();
  const size_t per = (size_t)cout_b * cin * 9;
  const size_t total = per * nb;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int br = (int)(i / per);
    const size_t r = i - (size_t)br * per;
    const int co = (int)(r / ((size_t)cin * 9));
    const int rr = (int)(r - (size_t)co * cin * 9);
    const int ci = rr / 9, tap = rr - ci * 9;
    const float v = __ldg(w.p[br] + r);
    size_t o;
    if (merged) o = ((size_t)ci * 9 + tap) * (nb * cout_b) + br * cout_b + co;
    else o = (((size_t)br * cin + ci) * 9 + tap) * cout_b + co;
    wp[o] = v;
  }
}

// Wd[g][co][8-tap][ci]: the transposed + flipped table that turns the forward kernel into
// the input-gradient kernel.  merged=1: single group with nb*cout_b "input" channels.
__global__ void pack_conv_wd_kernel(Ptr2 w, int nb, int cout_b, int cin, int merged, float* __restrict__ wd) {
  pdl_prologue();
  const size_t per = (size_t)cout_b * cin * 9;
  const size_t total = per * nb;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int br = (int)(i / per);
    const size_t r = i - (size_t)br * per;
    const int co = (int)(r / ((size_t)cin * 9));
    const int rr = (int)(r - (size_t)co * cin * 9);
    const int ci = rr / 9, tap = rr - ci * 9;
    const float v = __ldg(w.p[br] + r);
    (void)merged;  // both cases index the "input" channel as br*cout_b+co
    const size_t o = (((size_t)br * cout_b + co) * 9 + (8 - tap)) * cin + ci;
    wd[o] = v;
  }
}

// Dense centre taps of the Conv1d attention weights (C,C,ks): only tap ks/2 sees the length-1
// sequence (Hang2020.py:146-147,155-158).  d[i*C+j] and its transpose t[j*C+i].
__global__ void pack_spectral_kernel(const float* __restrict__ w, int C, int ks, float* __restrict__ d,
                                     float* __restrict__ t) {
  pdl_prologue();
  const int total = C * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / C, c = i - r * C;
    const float v = __ldg(w + (size_t)i * ks + ks / 2);
    d[i] = v;
    t[c * C + r] = v;
  }
}

struct BnParams {
  const float* gamma[2];
  const float* beta[2];
  float* rm[2];
  float* rv[2];
  long long* nbt[2];
  int c_per_branch;
};

constexpr int kBnCh = 8;         // channels per block in the BatchNorm finalize kernels
constexpr int kBnSlices = 128;   // row slices per block (block = 8 channels x 128 slices = 1024 threads)

// Sum over the kBnSlices row slices of a block for every channel (fixed order): xor-shuffles over the 4 slices a warp holds,
// then the 32 per-warp partials through shared memory.  Valid in the threads with ry == 0.
__device__ __forceinline__ void bn_block_sum2(double& a, double& b, double (*sa)[kBnCh], double (*sb)[kBnCh]) {
  a += __shfl_xor_sync(0xffffffffu, a, 8);  b += __shfl_xor_sync(0xffffffffu, b, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 16); b += __shfl_xor_sync(0xffffffffu, b, 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < kBnCh) { sa[warp][lane] = a; sb[warp][lane] = b; }
  __syncthreads();
  if (threadIdx.x < kBnCh) {
    a = 0.0; b = 0.0;
#pragma unroll 8
    for (int w = 0; w < kBnSlices * kBnCh / 32; ++w) { a += sa[w][threadIdx.x]; b += sb[w][threadIdx.x]; }
  }
}

// Block = 8 channels x 128 row slices, fp64 accumulation in fixed order (few rows per thread: the kernel is pure latency).
// training: reduce the conv kernels' partial sums, produce mean / invstd / scale / shift and update the running
// statistics (momentum 0.1, unbiased variance) exactly like nn.BatchNorm2d in train(); eval: use running statistics.
__global__ void __launch_bounds__(kBnCh * kBnSlices)
bn_fwd_finalize_kernel(const float* __restrict__ part /*[nblk][ctot][2]*/, int nblk, int ctot, double count, BnParams bn,
                       int training, float* __restrict__ mean, float* __restrict__ istd, float* __restrict__ scale,
                       float* __restrict__ shift, const float* __restrict__ update_gate /*null, or: running statistics only move if *gate != 0*/) {
  pdl_prologue();
  __shared__ double ss[kBnSlices * kBnCh / 32][kBnCh], sq[kBnSlices * kBnCh / 32][kBnCh];
  const int cx = threadIdx.x & (kBnCh - 1), ry = threadIdx.x / kBnCh;
  const int ch = blockIdx.x * kBnCh + cx;
  // the finishing threads fetch their channel's parameters up front: the loads travel under the partial-sum loop
  const bool fin = ry == 0 && ch < ctot;
  const int br = fin ? ch / bn.c_per_branch : 0, c = fin ? ch - br * bn.c_per_branch : 0;
  float p_gamma = 0.f, p_beta = 0.f, p_rm = 0.f, p_rv = 1.f;
  if (fin) { p_gamma = bn.gamma[br][c]; p_beta = bn.beta[br][c]; p_rm = bn.rm[br][c]; p_rv = bn.rv[br][c]; }
  double s = 0.0, q = 0.0;
  if (training && ch < ctot) {
#pragma unroll 4
    for (int k = ry; k < nblk; k += kBnSlices) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(part + ((size_t)k * ctot + ch) * 2));
      s += (double)v.x;
      q += (double)v.y;
    }
  }
  bn_block_sum2(s, q, ss, sq);
  if (!fin) return;
  float m, is;
  if (training) {
    const double mu = s / count;
    double var = q / count - mu * mu;
    if (var < 0.0) var = 0.0;
    m = (float)mu;
    is = (float)(1.0 / sqrt(var + (double)kBnEps));
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    if (update_gate == nullptr || __ldg(update_gate) != 0.f) {
      bn.rm[br][c] = (float)((1.0 - kBnMomentum) * (double)p_rm + kBnMomentum * mu);
      bn.rv[br][c] = (float)((1.0 - kBnMomentum) * (double)p_rv + kBnMomentum * unbiased);
      if (c == 0 && bn.nbt[br] != nullptr) bn.nbt[br][0] += 1;
    }
  } else {
    m = p_rm;
    is = (float)(1.0 / sqrt((double)p_rv + (double)kBnEps));
  }
  const float sc = p_gamma * is;
  mean[ch] = m;
  istd[ch] = is;
  scale[ch] = sc;
  shift[ch] = p_beta - m * sc;
}

struct BnGrads {
  float* dgamma[2];
  float* dbeta[2];
  float* dconv_b[2];
};

// Block = 8 channels x 128 crop slices.  rows[b][g][2C] hold per-crop (sum da | sum da*zhat); reduce over the batch in
// fp64 (fixed order), emit dgamma / dbeta / dbias and the coefficients of
//   dz = k0*da + k1*z + k2     (train: k0 = gamma*istd, k1 = -k0*istd*dgamma/N,
//                               k2 = -k0*dbeta/N - k1*mean;  eval: k1 = k2 = 0)
__global__ void __launch_bounds__(kBnCh * kBnSlices)
bn_bwd_finalize_kernel(const float* __restrict__ rows, int B, int G, int C, double count, BnParams bn, const float* __restrict__ mean,
                       const float* __restrict__ istd, int training, BnGrads gr, float* __restrict__ k0, float* __restrict__ k1,
                       float* __restrict__ k2) {
  pdl_prologue();
  __shared__ double sa[kBnSlices * kBnCh / 32][kBnCh], sb[kBnSlices * kBnCh / 32][kBnCh];
  const int cx = threadIdx.x & (kBnCh - 1), ry = threadIdx.x / kBnCh;
  const int ctot = G * C;
  const int ch = blockIdx.x * kBnCh + cx;
  const int g = ch / C, c = ch - g * C;
  const bool fin = ry == 0 && ch < ctot;
  const int br = fin ? ch / bn.c_per_branch : 0, cb = fin ? ch - br * bn.c_per_branch : 0;
  float p_gamma = 0.f, p_istd = 0.f, p_mean = 0.f;
  if (fin) { p_gamma = bn.gamma[br][cb]; p_istd = istd[ch]; p_mean = mean[ch]; }
  double s1 = 0.0, s2 = 0.0;
  if (ch < ctot) {
#pragma unroll 4
    for (int b = ry; b < B; b += kBnSlices) {
      const float* r = rows + ((size_t)b * G + g) * 2 * C;
      s1 += (double)__ldg(r + c);
      s2 += (double)__ldg(r + C + c);
    }
  }
  bn_block_sum2(s1, s2, sa, sb);
  if (!fin) return;
  const double gam = p_gamma;
  const double is = p_istd, mu = p_mean;
  const double a = gam * is;
  double b1 = 0.0, b2 = 0.0, dbias;
  if (training) {
    b1 = -a * is * s2 / count;
    b2 = -a * s1 / count - b1 * mu;
    dbias = 0.0;  // sum of dz over the batch vanishes identically under batch statistics
  } else {
    dbias = a * s1;
  }
  k0[ch] = (float)a;
  k1[ch] = (float)b1;
  k2[ch] = (float)b2;
  if (gr.dgamma[br]) gr.dgamma[br][cb] = (float)s2;
  if (gr.dbeta[br]) gr.dbeta[br][cb] = (float)s1;
  if (gr.dconv_b[br]) gr.dconv_b[br][cb] = (float)dbias;
}

// out[j*ostride] = sum_b rows[b*ld + j], j < n.  Block = 32 columns x 8 batch slices.
__global__ void colsum_kernel(const float* __restrict__ rows, size_t ld, int B, int n, float* __restrict__ out,
                              int ostride) {
  pdl_prologue();
  __shared__ float s[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (j < n)
    for (int b = threadIdx.y; b < B; b += 8) a += rows[(size_t)b * ld + j];
  s[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && j < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s[k][threadIdx.x];
    out[(size_t)j * ostride] = t;
  }
}

// out[i*si + j*sj] = sum_b U[b*ldu + i] * V[b*ldv + j]  (i < ni, j < nj); 16x16 outputs per
// block, batch walked in shared-memory tiles of 32.
__global__ void outer_sum_kernel(const float* __restrict__ U, size_t ldu, const float* __restrict__ V, size_t ldv,
                                 int B, int ni, int nj, float* __restrict__ out, size_t si, size_t sj) {
  pdl_prologue();
  __shared__ float su[32][17];
  __shared__ float sv[32][17];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 16 x 16
  const int i0 = blockIdx.y * 16, j0 = blockIdx.x * 16;
  const int t = ty * 16 + tx;
  float acc = 0.f;
  for (int b0 = 0; b0 < B; b0 += 32) {
    for (int e = t; e < 32 * 16; e += 256) {
      const int bb = e >> 4, k = e & 15;
      const int b = b0 + bb;
      su[bb][k] = (b < B && i0 + k < ni) ? U[(size_t)b * ldu + i0 + k] : 0.f;
      sv[bb][k] = (b < B && j0 + k < nj) ? V[(size_t)b * ldv + j0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) acc = fmaf(su[bb][ty], sv[bb][tx], acc);
    __syncthreads();
  }
  if (i0 + ty < ni && j0 + tx < nj) out[(size_t)(i0 + ty) * si + (size_t)(j0 + tx) * sj] = acc;
}

// joint = s_spec * float(w) + s_spat * float(1-w),  w = sigmoid(alpha) in fp64 (Hang2020.py:259-260)
__global__ void joint_fwd_kernel(const float* __restrict__ spec, const float* __restrict__ spat,
                                 const double* __restrict__ alpha, float* __restrict__ joint, float* __restrict__ keep_spec,
                                 float* __restrict__ keep_spat, size_t n) {
  pdl_prologue();
  const double w = 1.0 / (1.0 + exp(-alpha[0]));
  const float wf = (float)w, vf = (float)(1.0 - w);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float a = spec[i], b = spat[i];
    joint[i] = a * wf + b * vf;
    keep_spec[i] = a;   // copies of the two last-head scores for the alpha gradient (the caller owns `scores` and may reuse them)
    keep_spat[i] = b;
  }
}

// dS_spec = dscores_spec + djoint*w ; dS_spat = dscores_spat + djoint*(1-w)
__global__ void joint_bwd_kernel(const float* __restrict__ dspec, const float* __restrict__ dspat,
                                 const float* __restrict__ djoint, const double* __restrict__ alpha,
                                 float* __restrict__ out_spec, float* __restrict__ out_spat, size_t n) {
  pdl_prologue();
  const double w = 1.0 / (1.0 + exp(-alpha[0]));
  const float wf = (float)w, vf = (float)(1.0 - w);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float dj = djoint[i];
    out_spec[i] = (dspec ? dspec[i] : 0.f) + dj * wf;
    out_spat[i] = (dspat ? dspat[i] : 0.f) + dj * vf;
  }
}

// dalpha = w(1-w) * sum djoint * (s_spec - s_spat); single block, fixed-order fp64 tree.
__global__ void alpha_grad_kernel(const float* __restrict__ djoint, const float* __restrict__ spec,
                                  const float* __restrict__ spat, const double* __restrict__ alpha, size_t n,
                                  double* __restrict__ dalpha) {
  pdl_prologue();
  __shared__ double s[1024];
  double a = 0.0;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) a += (double)(djoint[i] * spec[i]) - (double)(djoint[i] * spat[i]);
  s[threadIdx.x] = a;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double w = 1.0 / (1.0 + exp(-alpha[0]));
    dalpha[0] = s[0] * w * (1.0 - w);
  }
}

// ---------------------------------------------------------------------------------------
// All small parameter gradients of a backward pass in two launches.  Each task is a batch reduction
//   out[i*si + j*sj] = sum_b U[b*ldu + i] * V[b*ldv + j]   (V == nullptr: column sums of U, nj = 1):
// attention / classifier weight and bias gradients.  Stage 1: a CTA owns one 32x32 output tile (outer products) or
// 32 columns (column sums) of one task for one of kReduceSplits slices of the batch and writes a partial tile;
// stage 2 adds the slices in fixed order (deterministic, no atomics).
// ---------------------------------------------------------------------------------------
struct ReduceTask {
  const float* U;
  const float* V;
  float* out;
  long long ldu, ldv, si, sj;
  int ni, nj;
  int tile_begin;   // first tile of this task
  int tiles_j;      // tiles along j
};
constexpr int kMaxReduceTasks = 96;
constexpr int kReduceSplits = 8;
struct ReduceTaskTable {
  ReduceTask t[kMaxReduceTasks];
  int n;
};

// grid = (tiles, kReduceSplits); partial[tile][split][32*32]
__global__ void __launch_bounds__(256) batched_reduce_kernel(const __grid_constant__ ReduceTaskTable tab, int B, float* __restrict__ partial) {
  pdl_prologue();
  __shared__ float su[32][33];
  __shared__ float sv[32][33];
  int ti = 0;
  while (ti + 1 < tab.n && tab.t[ti + 1].tile_begin <= (int)blockIdx.x) ++ti;
  const ReduceTask& T = tab.t[ti];
  const int tile = blockIdx.x - T.tile_begin;
  const int tid = threadIdx.x;
  const int per = (B + kReduceSplits - 1) / kReduceSplits;
  const int b_begin = blockIdx.y * per, b_end = min(B, b_begin + per);
  float* pout = partial + ((size_t)blockIdx.x * kReduceSplits + blockIdx.y) * 1024;
  if (T.V == nullptr) {
    // column sums: 32 columns x 8 batch slices
    const int tx = tid & 31, ty = tid >> 5;
    const int j = tile * 32 + tx;
    float a = 0.f;
    if (j < T.ni)
      for (int b = b_begin + ty; b < b_end; b += 8) a += __ldg(T.U + (size_t)b * T.ldu + j);
    su[ty][tx] = a;
    __syncthreads();
    if (ty == 0) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += su[k][tx];
      pout[tx] = t;
    }
    return;
  }
  const int i0 = (tile / T.tiles_j) * 32, j0 = (tile % T.tiles_j) * 32;
  const int tx = tid & 15, ty = tid >> 4;   // each thread: outputs (2*ty + {0,1}, 2*tx + {0,1})
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int b0 = b_begin; b0 < b_end; b0 += 32) {
    for (int e = tid; e < 32 * 32; e += 256) {
      const int bb = e >> 5, k = e & 31;
      const int b = b0 + bb;
      su[bb][k] = (b < b_end && i0 + k < T.ni) ? __ldg(T.U + (size_t)b * T.ldu + i0 + k) : 0.f;
      sv[bb][k] = (b < b_end && j0 + k < T.nj) ? __ldg(T.V + (size_t)b * T.ldv + j0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float u0 = su[bb][2 * ty], u1 = su[bb][2 * ty + 1];
      const float v0 = sv[bb][2 * tx], v1 = sv[bb][2 * tx + 1];
      acc[0][0] = fmaf(u0, v0, acc[0][0]); acc[0][1] = fmaf(u0, v1, acc[0][1]);
      acc[1][0] = fmaf(u1, v0, acc[1][0]); acc[1][1] = fmaf(u1, v1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) pout[(2 * ty + a) * 32 + 2 * tx + c] = acc[a][c];
}

// grid = tiles; sums the kReduceSplits partial tiles in fixed order and scatters to the gradient tensors.
__global__ void __launch_bounds__(256) batched_reduce_finish_kernel(const __grid_constant__ ReduceTaskTable tab, const float* __restrict__ partial) {
  pdl_prologue();
  int ti = 0;
  while (ti + 1 < tab.n && tab.t[ti + 1].tile_begin <= (int)blockIdx.x) ++ti;
  const ReduceTask& T = tab.t[ti];
  const int tile = blockIdx.x - T.tile_begin;
  const float* p = partial + (size_t)blockIdx.x * kReduceSplits * 1024;
  if (T.V == nullptr) {
    const int tx = threadIdx.x;
    const int j = tile * 32 + tx;
    if (tx < 32 && j < T.ni) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < kReduceSplits; ++k) t += p[k * 1024 + tx];
      T.out[(size_t)j * T.si] = t;
    }
    return;
  }
  const int i0 = (tile / T.tiles_j) * 32, j0 = (tile % T.tiles_j) * 32;
  for (int e = threadIdx.x; e < 1024; e += 256) {
    const int i = i0 + (e >> 5), j = j0 + (e & 31);
    if (i < T.ni && j < T.nj) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < kReduceSplits; ++k) t += p[k * 1024 + e];
      T.out[(size_t)i * T.si + (size_t)j * T.sj] = t;
    }
  }
}

__global__ void fill_zero_kernel(float* __restrict__ p, size_t n) {
  pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

}  // namespace dta
