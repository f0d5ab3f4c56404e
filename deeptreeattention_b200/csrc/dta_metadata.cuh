// Site-metadata branch and late fusion of BASELINE config 5 (reference src/models/metadata.py:9-44):
//   metadata:                Embedding(sites,16) -> BatchNorm1d(16) -> Dropout(p=0.7) -> Linear(16,classes) -> ReLU
//   metadata_sensor_fusion:  cat([metadata(site), Hang2020(images)], 1) -> Linear(2*classes, classes) -> ReLU
// The sensor scores come from dta_forward (they are an input here and dsensor is handed to dta_backward as `djoint`).
// Everything is tiny next to the crops (a few MFLOP per batch): exact fp32 CUDA-core arithmetic, fixed-order reductions.
#pragma once
#include "dta_common.cuh"

namespace dta {

constexpr int kMetaDim = 16;          // embedding width, metadata.py:12
constexpr float kMetaDropP = 0.7f;    // nn.Dropout(p=0.7), metadata.py:15
constexpr int kMetaCrops = 4;         // crops per CTA in the row kernels
constexpr int kMetaThreads = 256;
constexpr int kMetaStatThreads = 512; // 16 features x 32 batch slices

struct MetaTensors {                  // mirror of dta_metadata_tensors with typed pointers
  float* emb; float* bn_w; float* bn_b; float* bn_rm; float* bn_rv; long long* bn_nbt;
  float* mlp_w; float* mlp_b; float* fc_w; float* fc_b;
};

// Layout of `saved` (floats): [mean 16 | istd 16 | pad to 64][d: B x 16][scale: B x 16][m: B x classes]
struct MetaSaved {
  float* mean; float* istd; float* d; float* scale; float* m;
  size_t floats;
};
inline MetaSaved meta_saved_layout(void* base, int B, int classes) {
  MetaSaved s{};
  float* p = static_cast<float*>(base);
  s.mean = p; s.istd = p ? p + kMetaDim : nullptr;
  size_t off = 64;
  s.d = p ? p + off : nullptr; off += (size_t)B * kMetaDim;
  s.scale = p ? p + off : nullptr; off += (size_t)B * kMetaDim;
  s.m = p ? p + off : nullptr; off += (size_t)B * classes;
  s.floats = off;
  return s;
}
// Layout of the backward workspace (floats): [dgamma 16 | dbeta 16 | pad to 64][U1: B x C][U2: B x C][DY: B x 16][DYX: B x 16]
struct MetaWork {
  float* dgamma; float* dbeta; float* u1; float* u2; float* dy; float* dyx;
  size_t floats;
};
inline MetaWork meta_work_layout(void* base, int B, int classes) {
  MetaWork w{};
  float* p = static_cast<float*>(base);
  w.dgamma = p; w.dbeta = p ? p + kMetaDim : nullptr;
  size_t off = 64;
  w.u1 = p ? p + off : nullptr; off += (size_t)B * classes;
  w.u2 = p ? p + off : nullptr; off += (size_t)B * classes;
  w.dy = p ? p + off : nullptr; off += (size_t)B * kMetaDim;
  w.dyx = p ? p + off : nullptr; off += (size_t)B * kMetaDim;
  w.floats = off;
  return w;
}

// Counter-based dropout stream: splitmix64 of (seed, element index) -> 24-bit uniform in [0,1).
__device__ __forceinline__ float meta_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// BatchNorm1d statistics of the embedded sites (one CTA; fp64, fixed order).  training: batch mean / biased variance,
// running statistics updated with momentum 0.1 and the unbiased variance, num_batches_tracked += 1 (nn.BatchNorm1d);
// eval: running statistics.
__global__ void __launch_bounds__(kMetaStatThreads)
meta_bn_stats_kernel(const long long* __restrict__ site, int B, int sites, MetaTensors p, int training, float* __restrict__ mean,
                     float* __restrict__ istd) {
  __shared__ double ss[kMetaStatThreads / kMetaDim][kMetaDim], sq[kMetaStatThreads / kMetaDim][kMetaDim];
  const int f = threadIdx.x & (kMetaDim - 1), slice = threadIdx.x / kMetaDim;
  double s = 0.0, q = 0.0;
  if (training) {
    for (int b = slice; b < B; b += kMetaStatThreads / kMetaDim) {
      long long sidx = site[b];
      sidx = sidx < 0 ? 0 : (sidx >= sites ? sites - 1 : sidx);
      const double v = (double)p.emb[sidx * kMetaDim + f];
      s += v;
      q += v * v;
    }
  }
  ss[slice][f] = s;
  sq[slice][f] = q;
  __syncthreads();
  if (threadIdx.x >= kMetaDim) return;
  float m, is;
  if (training) {
    s = 0.0; q = 0.0;
    for (int k = 0; k < kMetaStatThreads / kMetaDim; ++k) { s += ss[k][f]; q += sq[k][f]; }
    const double mu = s / B;
    double var = q / B - mu * mu;
    if (var < 0.0) var = 0.0;
    m = (float)mu;
    is = (float)(1.0 / sqrt(var + (double)kBnEps));
    const double unbiased = B > 1 ? var * B / (B - 1.0) : var;
    if (p.bn_rm) p.bn_rm[f] = (float)((1.0 - kBnMomentum) * (double)p.bn_rm[f] + kBnMomentum * mu);
    if (p.bn_rv) p.bn_rv[f] = (float)((1.0 - kBnMomentum) * (double)p.bn_rv[f] + kBnMomentum * unbiased);
    if (f == 0 && p.bn_nbt) p.bn_nbt[0] += 1;
  } else {
    m = p.bn_rm[f];
    is = (float)(1.0 / sqrt((double)p.bn_rv[f] + (double)kBnEps));
  }
  mean[f] = m;
  istd[f] = is;
}

// kMetaCrops crops per CTA.  Dynamic shared memory: kMetaCrops * (16 + 2*classes) floats.
__global__ void __launch_bounds__(kMetaThreads)
meta_fwd_kernel(const long long* __restrict__ site, const float* __restrict__ sensor /*[B][C] or null*/, int B, int sites, int C,
                MetaTensors p, int training, const unsigned char* __restrict__ keep_mask /*[B][16] or null*/, unsigned long long seed,
                MetaSaved sv, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* s_d = smem;                                // [kMetaCrops][16]
  float* s_cat = smem + kMetaCrops * kMetaDim;      // [kMetaCrops][2C]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = blockIdx.x * kMetaCrops;
  if (tid < kMetaCrops * kMetaDim) {
    const int crop = tid / kMetaDim, f = tid - crop * kMetaDim, b = b0 + crop;
    float d = 0.f;
    if (b < B) {
      long long sidx = site[b];
      sidx = sidx < 0 ? 0 : (sidx >= sites ? sites - 1 : sidx);
      const float xhat = (p.emb[sidx * kMetaDim + f] - sv.mean[f]) * sv.istd[f];
      const float yv = fmaf(xhat, p.bn_w[f], p.bn_b[f]);
      float scale = 1.f;
      if (training) {
        const bool keep = keep_mask ? keep_mask[(size_t)b * kMetaDim + f] != 0
                                    : meta_uniform(seed, (unsigned long long)b * kMetaDim + f) >= kMetaDropP;
        scale = keep ? 1.f / (1.f - kMetaDropP) : 0.f;
      }
      d = yv * scale;
      sv.d[(size_t)b * kMetaDim + f] = d;
      sv.scale[(size_t)b * kMetaDim + f] = scale;
    }
    s_d[tid] = d;
  }
  __syncthreads();
  for (int e = tid; e < kMetaCrops * C; e += kMetaThreads) {
    const int crop = e / C, i = e - crop * C, b = b0 + crop;
    float a = p.mlp_b[i];
    const float* w = p.mlp_w + (size_t)i * kMetaDim;
#pragma unroll
    for (int f = 0; f < kMetaDim; ++f) a = fmaf(w[f], s_d[crop * kMetaDim + f], a);
    a = fmaxf(a, 0.f);
    s_cat[crop * 2 * C + i] = a;
    if (b < B) {
      sv.m[(size_t)b * C + i] = a;
      if (sensor == nullptr) out[(size_t)b * C + i] = a;     // stand-alone metadata module
    }
    if (sensor != nullptr) s_cat[crop * 2 * C + C + i] = b < B ? sensor[(size_t)b * C + i] : 0.f;
  }
  if (sensor == nullptr) return;
  __syncthreads();
  // fusion layer: a warp per output class, the weight row read once for the CTA's crops
  for (int i = warp; i < C; i += kMetaThreads / 32) {
    float acc[kMetaCrops];
#pragma unroll
    for (int c = 0; c < kMetaCrops; ++c) acc[c] = 0.f;
    const float* w = p.fc_w + (size_t)i * 2 * C;
    for (int j = lane; j < 2 * C; j += 32) {
      const float wv = w[j];
#pragma unroll
      for (int c = 0; c < kMetaCrops; ++c) acc[c] = fmaf(wv, s_cat[c * 2 * C + j], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < kMetaCrops; ++c) acc[c] = warp_sum(acc[c]);
    if (lane == 0) {
      const float bias = p.fc_b[i];
#pragma unroll
      for (int c = 0; c < kMetaCrops; ++c)
        if (b0 + c < B) out[(size_t)(b0 + c) * C + i] = fmaxf(acc[c] + bias, 0.f);
    }
  }
}

// Per-crop part of the backward.  Dynamic shared memory: kMetaCrops * (2*classes) floats.
//   U1 = dout * (out > 0)                       (gradient at the fusion layer's pre-activation; fused only)
//   dcat = U1 * fc_w ; dsensor = dcat[:, C:] ;   U2 = dcat[:, :C] * (m > 0)      (stand-alone: U2 = dout * (out > 0))
//   dd = U2 * mlp_w ; DY = dd * scale ; DYX = DY * xhat
__global__ void __launch_bounds__(kMetaThreads)
meta_bwd_rows_kernel(const long long* __restrict__ site, int B, int sites, int C, int fused, MetaTensors p, MetaSaved sv,
                     const float* __restrict__ out, const float* __restrict__ dout, MetaWork wk, float* __restrict__ dsensor) {
  extern __shared__ __align__(16) float smem[];
  float* s_u1 = smem;                       // [kMetaCrops][C]
  float* s_u2 = smem + kMetaCrops * C;      // [kMetaCrops][C]
  const int tid = threadIdx.x;
  const int b0 = blockIdx.x * kMetaCrops;
  for (int e = tid; e < kMetaCrops * C; e += kMetaThreads) {
    const int crop = e / C, i = e - crop * C, b = b0 + crop;
    float g = 0.f;
    if (b < B) {
      g = out[(size_t)b * C + i] > 0.f ? dout[(size_t)b * C + i] : 0.f;
      if (fused) wk.u1[(size_t)b * C + i] = g; else wk.u2[(size_t)b * C + i] = g;
    }
    (fused ? s_u1 : s_u2)[e] = g;
  }
  __syncthreads();
  if (fused) {
    for (int j = tid; j < 2 * C; j += kMetaThreads) {
      float acc[kMetaCrops];
#pragma unroll
      for (int c = 0; c < kMetaCrops; ++c) acc[c] = 0.f;
      for (int i = 0; i < C; ++i) {
        const float wv = p.fc_w[(size_t)i * 2 * C + j];
#pragma unroll
        for (int c = 0; c < kMetaCrops; ++c) acc[c] = fmaf(s_u1[c * C + i], wv, acc[c]);
      }
#pragma unroll
      for (int c = 0; c < kMetaCrops; ++c) {
        const int b = b0 + c;
        if (j < C) {
          const float g = (b < B && sv.m[(size_t)b * C + j] > 0.f) ? acc[c] : 0.f;
          s_u2[c * C + j] = g;
          if (b < B) wk.u2[(size_t)b * C + j] = g;
        } else if (b < B && dsensor != nullptr) {
          dsensor[(size_t)b * C + (j - C)] = acc[c];
        }
      }
    }
    __syncthreads();
  }
  if (tid < kMetaCrops * kMetaDim) {
    const int crop = tid / kMetaDim, f = tid - crop * kMetaDim, b = b0 + crop;
    if (b < B) {
      float dd = 0.f;
      for (int i = 0; i < C; ++i) dd = fmaf(s_u2[crop * C + i], p.mlp_w[(size_t)i * kMetaDim + f], dd);
      long long sidx = site[b];
      sidx = sidx < 0 ? 0 : (sidx >= sites ? sites - 1 : sidx);
      const float xhat = (p.emb[sidx * kMetaDim + f] - sv.mean[f]) * sv.istd[f];
      const float dy = dd * sv.scale[(size_t)b * kMetaDim + f];
      wk.dy[(size_t)b * kMetaDim + f] = dy;
      wk.dyx[(size_t)b * kMetaDim + f] = dy * xhat;
    }
  }
}

// Embedding gradient through the BatchNorm1d backward (one CTA per site: 16 features x 16 batch slices, slices added in order):
//   train: dx = gamma*istd * (dy - dbeta/B - xhat*dgamma/B) ; eval: dx = gamma*istd*dy ;  dE[s][f] = sum_{b: site[b]=s} dx[b][f]
__global__ void __launch_bounds__(256)
meta_bwd_embed_kernel(const long long* __restrict__ site, int B, int sites, MetaTensors p, MetaSaved sv, MetaWork wk,
                      int training, float* __restrict__ demb) {
  __shared__ float part[16][kMetaDim];
  const int s = blockIdx.x, f = threadIdx.x & (kMetaDim - 1), slice = threadIdx.x / kMetaDim;
  const float k0 = p.bn_w[f] * sv.istd[f];
  const float xhat = (p.emb[(size_t)s * kMetaDim + f] - sv.mean[f]) * sv.istd[f];
  const float c1 = training ? wk.dbeta[f] / (float)B : 0.f, c2 = training ? xhat * wk.dgamma[f] / (float)B : 0.f;
  float acc = 0.f;
  for (int b = slice; b < B; b += 16) {
    long long sidx = site[b];
    sidx = sidx < 0 ? 0 : (sidx >= sites ? sites - 1 : sidx);
    if (sidx == s) acc += k0 * (wk.dy[(size_t)b * kMetaDim + f] - c1 - c2);
  }
  part[slice][f] = acc;
  __syncthreads();
  if (slice == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) t += part[k][f];
    demb[(size_t)s * kMetaDim + f] = t;
  }
}

// All batch reductions of the backward in ONE launch: task t computes out[i*si + j*sj] = sum_b U[b*ldu + i] * V[b*ldv + j]
// (V == nullptr: column sums of U).  A CTA owns one 32x32 output tile of one task and walks the whole batch in order
// (fixed summation order, no atomics); 256 threads, 2x2 outputs each.
struct MetaReduceTask {
  const float* U; const float* V; float* out;
  int ldu, ldv, si, sj, ni, nj, tile_begin, tiles_j;
};
constexpr int kMetaMaxTasks = 8;
struct MetaReduceTable {
  MetaReduceTask t[kMetaMaxTasks];
  int n;
};
__global__ void __launch_bounds__(256) meta_reduce_kernel(const __grid_constant__ MetaReduceTable tab, int B) {
  __shared__ float su[32][33];
  __shared__ float sv[32][33];
  int ti = 0;
  while (ti + 1 < tab.n && tab.t[ti + 1].tile_begin <= (int)blockIdx.x) ++ti;
  const MetaReduceTask& T = tab.t[ti];
  const int tile = blockIdx.x - T.tile_begin;
  const int tid = threadIdx.x;
  if (T.V == nullptr) {
    const int tx = tid & 31, ty = tid >> 5;
    const int i = tile * 32 + tx;
    float a = 0.f;
    if (i < T.ni)
      for (int b = ty; b < B; b += 8) a += T.U[(size_t)b * T.ldu + i];
    su[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && i < T.ni) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += su[k][tx];
      T.out[(size_t)i * T.si] = t;
    }
    return;
  }
  const int i0 = (tile / T.tiles_j) * 32, j0 = (tile % T.tiles_j) * 32;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int b0 = 0; b0 < B; b0 += 32) {
    for (int e = tid; e < 32 * 32; e += 256) {
      const int bb = e >> 5, k = e & 31;
      const int b = b0 + bb;
      su[bb][k] = (b < B && i0 + k < T.ni) ? T.U[(size_t)b * T.ldu + i0 + k] : 0.f;
      sv[bb][k] = (b < B && j0 + k < T.nj) ? T.V[(size_t)b * T.ldv + j0 + k] : 0.f;
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
    for (int c = 0; c < 2; ++c) {
      const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + c;
      if (i < T.ni && j < T.nj) T.out[(size_t)i * T.si + (size_t)j * T.sj] = acc[a][c];
    }
}

__global__ void meta_copy16_kernel(const float* __restrict__ a, float* __restrict__ da, const float* __restrict__ b, float* __restrict__ db) {
  const int f = threadIdx.x;
  if (f < kMetaDim) {
    if (da) da[f] = a[f];
    if (db) db[f] = b[f];
  }
}

}  // namespace dta
