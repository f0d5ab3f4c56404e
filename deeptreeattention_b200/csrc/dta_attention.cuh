// Per-crop fused stages between the convolutions: BatchNorm apply + ReLU + max-pool +
// spectral / spatial attention gate + classifier head (forward), and their backward.
// One CTA per (crop, branch).  Reference: conv_module.forward Hang2020.py:24-31,
// spectral_attention.forward :149-168, spatial_attention.forward :105-124, Classifier :63-66.
#pragma once
#include "dta_common.cuh"
#include "dta_loss.cuh"
#include "dta_tc.cuh"

namespace dta {

constexpr int kAttnThreads = 256;

// Per-branch parameter views used by the attention kernels (device pointers).
struct AttnParams {
  int btype[2];
  // spectral: dense centre-tap matrices packed by pack_spectral_kernel:
  //   w0d[i*C+j] = attention_conv1.weight[i][j][ks/2], w0t = transpose; same for w1
  const float* w0d[2]; const float* w0t[2]; const float* w1d[2]; const float* w1t[2];
  // spatial: raw reference tensors (1,1,ks,ks)
  const float* st0[2]; const float* st1[2];
  const float* b0[2]; const float* b1[2];
  const float* pool_w[2]; const float* pool_b[2];
  const float* fc_w[2]; const float* fc_b[2];   // head; nullptr = no head at this block
  int classes_g[2];                             // forward only: per-branch class count (0 = the launch's common `classes`); the
                                                // inference fan-out pairs networks of different levels (dta_forward_pair)
};

template <int C, int SPRE, bool POOL>
struct AttnCfg {
  static constexpr int S = POOL ? SPRE / 2 : SPRE;
  static constexpr int HW = S * S;
  static constexpr int HWPRE = SPRE * SPRE;
  static constexpr int KS_SPATIAL = (C == 32) ? 7 : (C == 64 ? 5 : 3);   // Hang2020.py:77-82
  static constexpr int WIN = (C == 32) ? 4 : (C == 64 ? 2 : 1);         // Hang2020.py:91-99
  static constexpr int FEAT_SPATIAL = 4 * C;                             // 128 / 256 / 512
  static constexpr int ROW = (C > HW) ? C : HW;   // stride of one vector in the saved attention row
  static constexpr int ATT_LD = 3 * ROW;
  static constexpr int FEAT_LD = 4 * C;
};

// Build r = [maxpool2x2](relu(z*scale+shift)) in shared memory (and optionally the argmax
// slot of every pooled cell, first-max rule of ATen max_pool2d, with the conv output at that slot).
template <int C, int SPRE, bool POOL>
__device__ __forceinline__ void build_r(const float* __restrict__ zsrc, const float* __restrict__ scale,
                                        const float* __restrict__ shift, float* s_r, unsigned char* s_arg, float* s_zarg) {
  // POOL: the 2x2 windows are read straight from global memory (each conv output element is needed once, so staging
  // the pre-pool plane would only cost shared memory and occupancy).  All loads of a thread are issued before the first
  // use: the loop is bound by global-load latency, not by arithmetic.
  // !POOL: s_r already holds the crop's conv output, bulk-copied by the caller (attn_stage_in), and is rectified in place.
  using Cfg = AttnCfg<C, SPRE, POOL>;
  const int tid = threadIdx.x;
  if (POOL) {
    constexpr int N = C * Cfg::HW;
    constexpr int NIT = (N + kAttnThreads - 1) / kAttnThreads;
    float zv[NIT][4];
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = tid + it * kAttnThreads;
      if (i < N) {
        const int c = i / Cfg::HW, p = i - c * Cfg::HW;
        const int y = p / Cfg::S, x = p - y * Cfg::S;
        const float* q = zsrc + c * Cfg::HWPRE + (2 * y) * SPRE + 2 * x;
        zv[it][0] = __ldg(q); zv[it][1] = __ldg(q + 1); zv[it][2] = __ldg(q + SPRE); zv[it][3] = __ldg(q + SPRE + 1);
      }
    }
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = tid + it * kAttnThreads;
      if (i < N) {
        const int c = i / Cfg::HW;
        const float sc = __ldg(scale + c), sh = __ldg(shift + c);
        float best = -INFINITY, zbest = 0.f;
        int arg = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = fmaxf(fmaf(zv[it][k], sc, sh), 0.f);
          if (a > best) { best = a; arg = k; zbest = zv[it][k]; }
        }
        s_r[i] = best;
        if (s_arg != nullptr) { s_arg[i] = (unsigned char)arg; s_zarg[i] = zbest; }
      }
    }
  } else {
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kAttnThreads / 32, NC = (C + NW - 1) / NW;   // channels per warp
    float scv[NC], shv[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = warp + k * NW;
      scv[k] = c < C ? __ldg(scale + c) : 0.f;
      shv[k] = c < C ? __ldg(shift + c) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = warp + k * NW;
      if (c < C)
        for (int p = lane; p < Cfg::HW; p += 32) s_r[c * Cfg::HW + p] = fmaxf(fmaf(s_r[c * Cfg::HW + p], scv[k], shv[k]), 0.f);
    }
  }
  __syncthreads();
}

// out[i] = sum_{j < nin} w[j * nout + i] * s_vec[j] for i < nout, with ALL threads of the CTA: when nout < kAttnThreads
// (nout a power of two) the j range is split over kAttnThreads / nout thread groups whose partials are added in fixed
// order through s_part (kAttnThreads floats); loads are unrolled so several are in flight per thread (these loops are
// bound by L2 latency).  s_out may alias nothing read here.  Ends with a barrier: s_out is ready for every thread.
__device__ __forceinline__ void fc_cols(const float* __restrict__ w, int nout, int nin, const float* s_vec, float* s_out, float* s_part) {
  const int tid = threadIdx.x;
  if (nout >= kAttnThreads) {
    for (int i = tid; i < nout; i += kAttnThreads) {
      float a = 0.f;
#pragma unroll 16
      for (int j = 0; j < nin; ++j) a = fmaf(__ldg(w + (size_t)j * nout + i), s_vec[j], a);
      s_out[i] = a;
    }
    __syncthreads();
    return;
  }
  const int T = kAttnThreads / nout;
  const int per = (nin + T - 1) / T;
  const int i = tid & (nout - 1), q = tid / nout;
  const int j0 = q * per, j1 = min(nin, j0 + per);
  float a = 0.f;
#pragma unroll 16
  for (int j = j0; j < j1; ++j) a = fmaf(__ldg(w + (size_t)j * nout + i), s_vec[j], a);
  s_part[tid] = a;
  __syncthreads();
  if (tid < nout) {
    float r = 0.f;
    for (int k = 0; k < T; ++k) r += s_part[k * nout + tid];
    s_out[tid] = r;
  }
  __syncthreads();
}

// Stage contiguous per-crop blocks global -> shared with the bulk-copy engine (one instruction per block instead of a
// load/store loop per thread; the kernels were stalled on exactly those loops).  Blocks are multiples of 16 bytes and
// 16-byte aligned on both sides by construction of the saved / workspace layouts.  Thread 0 issues, everybody waits.
__device__ __forceinline__ void attn_stage_in(uint64_t* bar, float* dst0, const float* src0, int n0, float* dst1, const float* src1,
                                              int n1) {
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_fence_init();
    tc::mbar_arrive_expect_tx(bar, (uint32_t)(n0 + (src1 ? n1 : 0)) * 4u);
    tc::bulk_g2s(dst0, src0, (uint32_t)n0 * 4u, bar);
    if (src1) tc::bulk_g2s(dst1, src1, (uint32_t)n1 * 4u, bar);
  }
  __syncthreads();            // barrier initialised before anyone polls it
  tc::mbar_wait(bar, 0);
}

// k x k "same" stencil on an S x S plane held in shared memory (zero padding).
template <int S, int KS>
__device__ __forceinline__ float stencil_at(const float* s_plane, const float* __restrict__ w, int y, int x) {
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < KS; ++u) {
    const int yy = y + u - KS / 2;
    if (yy < 0 || yy >= S) continue;
#pragma unroll
    for (int v = 0; v < KS; ++v) {
      const int xx = x + v - KS / 2;
      if (xx < 0 || xx >= S) continue;
      acc = fmaf(__ldg(w + u * KS + v), s_plane[yy * S + xx], acc);
    }
  }
  return acc;
}

// ------------------------------------------------------------------------------ forward
template <int C, int SPRE, bool POOL>
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const float* __restrict__ z /*[B][G*C][HWPRE]*/, const float* __restrict__ scale,
                const float* __restrict__ shift, AttnParams prm, int classes,
                float* __restrict__ att /*[B][G][ATT_LD]*/, float* __restrict__ feat /*[B][G][FEAT_LD]*/,
                MutPtr2 scores /*per branch [B][classes]*/,
                __nv_bfloat16* __restrict__ packed /*split-bf16 position stream of the gated output (next conv's operand) or null*/,
                size_t packed_rows, int packed_nchunk) {
  pdl_prologue();
  using Cfg = AttnCfg<C, SPRE, POOL>;
  constexpr int S = Cfg::S, HW = Cfg::HW;
  extern __shared__ __align__(16) float smem[];
  float* s_r = smem;                       // C*HW rectified (pooled) activations; !POOL: receives z first
  float* s_v = s_r + C * HW;               // 3*ROW vectors
  float* s_feat = s_v + 3 * Cfg::ROW;      // FEAT_LD
  float* s_part = s_feat + Cfg::FEAT_LD;   // kAttnThreads partials of the split matrix-vector products

  const int b = blockIdx.x, g = blockIdx.y, G = gridDim.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int btype = prm.btype[g];
  __shared__ uint64_t stage_bar;

  const float* zg = z + ((size_t)b * G + g) * C * Cfg::HWPRE;
  if (!POOL) attn_stage_in(&stage_bar, s_r, zg, C * Cfg::HWPRE, nullptr, nullptr, 0);
  build_r<C, SPRE, POOL>(zg, scale + g * C, shift + g * C, s_r, nullptr, nullptr);

  float* att_row = att + ((size_t)b * G + g) * Cfg::ATT_LD;
  float* feat_row = feat + ((size_t)b * G + g) * Cfg::FEAT_LD;
  int F = 0;
  if (btype == BR_SPECTRAL) {
    float* s_g = s_v; float* s_h = s_v + Cfg::ROW; float* s_s = s_v + 2 * Cfg::ROW;
    if (HW <= 32) {   // few positions per plane: one thread per channel walks them (a warp per channel would idle most lanes)
      for (int c = tid; c < C; c += kAttnThreads) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < HW; ++p) a += s_r[c * HW + p];
        s_g[c] = a / (float)HW;
      }
    } else {
      for (int c = warp; c < C; c += kAttnThreads / 32) {
        float a = 0.f;
        for (int p = lane; p < HW; p += 32) a += s_r[c * HW + p];
        a = warp_sum(a);
        if (lane == 0) s_g[c] = a / (float)HW;
      }
    }
    __syncthreads();
    fc_cols(prm.w0t[g], C, C, s_g, s_h, s_part);
    if (tid < C) s_h[tid] = fmaxf(s_h[tid] + __ldg(prm.b0[g] + tid), 0.f);
    __syncthreads();
    fc_cols(prm.w1t[g], C, C, s_h, s_s, s_part);
    if (tid < C) s_s[tid] = sigmoidf_acc(s_s[tid] + __ldg(prm.b1[g] + tid));
    __syncthreads();
    if (HW <= 32) {
      for (int c = tid; c < C; c += kAttnThreads) {
        const float sv = s_s[c];
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < HW; ++p) a += s_r[c * HW + p] * sv;
        s_feat[c] = a / (float)HW;
      }
    } else {
      for (int c = warp; c < C; c += kAttnThreads / 32) {
        const float sv = s_s[c];
        float a = 0.f;
        for (int p = lane; p < HW; p += 32) a += s_r[c * HW + p] * sv;
        a = warp_sum(a);
        if (lane == 0) s_feat[c] = a / (float)HW;
      }
    }
    F = C;
    for (int i = tid; i < C; i += kAttnThreads) {
      att_row[i] = s_g[i]; att_row[Cfg::ROW + i] = s_h[i]; att_row[2 * Cfg::ROW + i] = s_s[i];
    }
  } else if (btype == BR_SPATIAL) {
    constexpr int KS = Cfg::KS_SPATIAL;
    float* s_q = s_v; float* s_t = s_v + Cfg::ROW; float* s_s = s_v + 2 * Cfg::ROW;
    for (int p = tid; p < HW; p += kAttnThreads) {
      float a = __ldg(prm.pool_b[g]);
      const float* w = prm.pool_w[g];
#pragma unroll 8
      for (int c = 0; c < C; ++c) a = fmaf(__ldg(w + c), s_r[c * HW + p], a);
      s_q[p] = fmaxf(a, 0.f);
    }
    __syncthreads();
    for (int p = tid; p < HW; p += kAttnThreads)
      s_t[p] = fmaxf(stencil_at<S, KS>(s_q, prm.st0[g], p / S, p % S) + __ldg(prm.b0[g]), 0.f);
    __syncthreads();
    for (int p = tid; p < HW; p += kAttnThreads)
      s_s[p] = sigmoidf_acc(stencil_at<S, KS>(s_t, prm.st1[g], p / S, p % S) + __ldg(prm.b1[g]));
    __syncthreads();
    constexpr int WIN = Cfg::WIN;
    for (int f = tid; f < 4 * C; f += kAttnThreads) {
      const int c = f >> 2, i = (f >> 1) & 1, j = f & 1;
      float best = -INFINITY;
#pragma unroll
      for (int u = 0; u < WIN; ++u)
#pragma unroll
        for (int v = 0; v < WIN; ++v) {
          const int p = (i * WIN + u) * S + j * WIN + v;
          best = fmaxf(best, s_r[c * HW + p] * s_s[p]);
        }
      s_feat[f] = best;
    }
    F = 4 * C;
    for (int p = tid; p < HW; p += kAttnThreads) {
      att_row[p] = s_q[p]; att_row[Cfg::ROW + p] = s_t[p]; att_row[2 * Cfg::ROW + p] = s_s[p];
    }
  } else {
    // vanilla_CNN: flatten(r) feeds fc1 (Hang2020.py:50-51); only launched for block 3
    for (int f = tid; f < C * HW; f += kAttnThreads) s_feat[f] = s_r[f];
    F = C * HW;
  }
  __syncthreads();
  if (packed != nullptr) {
    // The gated feature map r * s is the next convolution's input: write it straight into that kernel's operand
    // format (dta_conv_tc.cuh position stream: pad row above / pad column right of the plane, 8-channel chunks,
    // hi and lo bf16 planes) -- same arithmetic as tc_pack_stream_kernel<SRC_ACT>, without re-reading z.
    constexpr int PT = S + 1, PC = (S + 1) * PT;
    const float* s_s = s_v + 2 * Cfg::ROW;
    uint4* dst = reinterpret_cast<uint4*>(packed);
    const size_t lo_off = packed_rows * packed_nchunk;
    for (int u = tid; u < (C / 8) * PC; u += kAttnThreads) {
      const int c8 = u / PC, r = u - c8 * PC;
      const int yy = r / PT, xx = r - yy * PT;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (yy >= 1 && xx < S) {
        const int p = (yy - 1) * S + xx;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = s_r[(c8 * 8 + j) * HW + p];
          v[j] = btype == BR_SPECTRAL ? a * s_s[c8 * 8 + j] : (btype == BR_SPATIAL ? a * s_s[p] : a);
        }
      }
      uint4 hi, lo;
      tc::split2(v[0], v[1], hi.x, lo.x);
      tc::split2(v[2], v[3], hi.y, lo.y);
      tc::split2(v[4], v[5], hi.z, lo.z);
      tc::split2(v[6], v[7], hi.w, lo.w);
      const size_t idx = (size_t)(g * (C / 8) + c8) * packed_rows + kTcGuard + (size_t)b * PC + r;
      dst[idx] = hi;
      dst[lo_off + idx] = lo;
    }
  }
  for (int f = tid; f < F; f += kAttnThreads) feat_row[f] = s_feat[f];
  const float* fw = prm.fc_w[g];
  if (fw != nullptr) {
    // classifier head (Hang2020.py:64): a warp per class, FOUR classes in flight per warp so that the loads and the
    // shuffle reductions of independent classes overlap
    const int ncls = prm.classes_g[g] > 0 ? prm.classes_g[g] : classes;
    float* sc_out = scores.p[g] + (size_t)b * ncls;
    constexpr int NW = kAttnThreads / 32;
    for (int cls0 = warp; cls0 < ncls; cls0 += 4 * NW) {
      float a[4] = {0.f, 0.f, 0.f, 0.f};
      const float* w0 = fw + (size_t)cls0 * F;
      const bool v1 = cls0 + NW < ncls, v2 = cls0 + 2 * NW < ncls, v3 = cls0 + 3 * NW < ncls;
      const float* w1 = v1 ? w0 + (size_t)NW * F : w0;
      const float* w2 = v2 ? w0 + (size_t)2 * NW * F : w0;
      const float* w3 = v3 ? w0 + (size_t)3 * NW * F : w0;
#pragma unroll 4
      for (int f = lane; f < F; f += 32) {
        const float x = s_feat[f];
        a[0] = fmaf(x, __ldg(w0 + f), a[0]);
        a[1] = fmaf(x, __ldg(w1 + f), a[1]);
        a[2] = fmaf(x, __ldg(w2 + f), a[2]);
        a[3] = fmaf(x, __ldg(w3 + f), a[3]);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
      }
      if (lane == 0) {
        sc_out[cls0] = a[0] + __ldg(prm.fc_b[g] + cls0);
        if (v1) sc_out[cls0 + NW] = a[1] + __ldg(prm.fc_b[g] + cls0 + NW);
        if (v2) sc_out[cls0 + 2 * NW] = a[2] + __ldg(prm.fc_b[g] + cls0 + 2 * NW);
        if (v3) sc_out[cls0 + 3 * NW] = a[3] + __ldg(prm.fc_b[g] + cls0 + 3 * NW);
      }
    }
  }
}

template <int C, int SPRE, bool POOL>
constexpr size_t attn_fwd_smem() {
  using Cfg = AttnCfg<C, SPRE, POOL>;
  return sizeof(float) * (C * Cfg::HW + 3 * Cfg::ROW + Cfg::FEAT_LD + kAttnThreads);
}

// ------------------------------------------------------------------------------ backward
// Row layout of the per-crop parameter-gradient partials written by attn_bwd_kernel:
//   spectral: [du2 (C) | du1 (C)]
//   spatial : [dA1 (KS*KS) | db_a1 | dA2 (KS*KS) | db_a2 | dpool_w (C) | dpool_b]
template <int C, int SPRE, bool POOL>
struct AttnBwdRow {
  using Cfg = AttnCfg<C, SPRE, POOL>;
  static constexpr int KK = Cfg::KS_SPATIAL * Cfg::KS_SPATIAL;
  static constexpr int SPATIAL_LEN = 2 * KK + 2 + C + 1;
  static constexpr int SPECTRAL_LEN = 2 * C;
  static constexpr int LEN = SPATIAL_LEN > SPECTRAL_LEN ? SPATIAL_LEN : SPECTRAL_LEN;
  static constexpr int LD = (LEN + 3) / 4 * 4;
};

template <int C, int SPRE, bool POOL>
__global__ void __launch_bounds__(kAttnThreads)
attn_bwd_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                const float* __restrict__ mean, const float* __restrict__ istd, AttnParams prm, int classes,
                const float* __restrict__ att, const float* __restrict__ feat_unused,
                Ptr2 dscores /*per branch [B][classes] or null*/, const float* __restrict__ dout /*[B][G][C][HW] or null*/,
                float* __restrict__ da /*[B][G*C][HWPRE]; compact: [B][G*C][HW] values, then [B][G*C][HW] arg-max bytes*/,
                float* __restrict__ bnrow /*[B][G][2C]*/, float* __restrict__ prow /*[B][G][ROW_LD]*/, int compact,
                CeInline ce /*scores[g] != null: this head's score gradient is formed here from the scores (fused training step)*/) {
  pdl_prologue();
  using Cfg = AttnCfg<C, SPRE, POOL>;
  using Row = AttnBwdRow<C, SPRE, POOL>;
  constexpr int S = Cfg::S, HW = Cfg::HW, HWPRE = Cfg::HWPRE;
  extern __shared__ __align__(16) float smem[];
  float* s_r = smem;                       // C*HW rectified (pooled) activations; !POOL: receives z first
  float* s_D = s_r + C * HW;               // C*HW  gradient wrt gated output, then wrt r
  float* s_v = s_D + C * HW;               // 3*ROW saved attention vectors
  float* s_w = s_v + 3 * Cfg::ROW;         // 4*ROW work vectors
  float* s_dfeat = s_w + 4 * Cfg::ROW;     // FEAT_LD
  float* s_part = s_dfeat + Cfg::FEAT_LD;  // kAttnThreads partials of the split matrix-vector products
  float* s_ds = s_part + kAttnThreads;     // classes (rounded up by the launcher)
  float* s_zarg = s_ds + ((classes + 3) / 4) * 4;                                            // POOL: C*HW conv outputs at the arg-max
  unsigned char* s_arg = reinterpret_cast<unsigned char*>(s_zarg + (POOL ? C * HW : 0));     // POOL: C*HW arg-max slots

  const int b = blockIdx.x, g = blockIdx.y, G = gridDim.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int btype = prm.btype[g];
  const float* sc = scale + g * C;
  const float* sh = shift + g * C;
  (void)feat_unused;
  __shared__ uint64_t stage_bar;

  // conv output of this crop and the upstream gradient of its gated feature map, one bulk copy each
  const float* dsrc = dout ? dout + ((size_t)b * G + g) * C * HW : nullptr;
  const float* zg = z + ((size_t)b * G + g) * C * HWPRE;
  // small per-crop vectors first: their loads travel while the bulk copies are in flight
  const float* att_row = att + ((size_t)b * G + g) * Cfg::ATT_LD;
  for (int i = tid; i < 3 * Cfg::ROW; i += kAttnThreads) s_v[i] = __ldg(att_row + i);
  const float* dsc = dscores.p[g];
  const bool ce_here = ce.scores[g] != nullptr;
  const bool has_head = (dsc != nullptr || ce_here) && (prm.fc_w[g] != nullptr);
  if (has_head && ce_here) {
    // fused training step: the weighted cross-entropy gradient of this crop's head, by warp 0, with the arithmetic of
    // ce_rows_kernel (which writes the same values to memory off the critical path for the head's weight gradient)
    if (warp == 0)
      ce_row_warp(ce.scores[g] + (size_t)b * classes, classes, ce.y[b], ce.w, ce.den[0], true, [&](int c, float v) { s_ds[c] = v; });
  } else if (has_head) {
    for (int i = tid; i < classes; i += kAttnThreads) s_ds[i] = __ldg(dsc + (size_t)b * classes + i);
  }
  if (!POOL) attn_stage_in(&stage_bar, s_r, zg, C * HWPRE, s_D, dsrc, C * HW);
  else if (dsrc != nullptr) attn_stage_in(&stage_bar, s_D, dsrc, C * HW, nullptr, nullptr, 0);
  build_r<C, SPRE, POOL>(zg, sc, sh, s_r, POOL ? s_arg : nullptr, POOL ? s_zarg : nullptr);   // ends with a barrier

  const int F = (btype == BR_SPECTRAL) ? C : (btype == BR_SPATIAL ? 4 * C : (has_head ? C * HW : 0));
  // gradient of the head features: dfeat = Wc^T dscores   (Classifier, Hang2020.py:64)
  if (has_head) {
    fc_cols(prm.fc_w[g], F, classes, s_ds, s_dfeat, s_part);
  } else {
    for (int f = tid; f < F; f += kAttnThreads) s_dfeat[f] = 0.f;
  }
  // upstream gradient of the gated feature map from the next conv's dgrad: already staged; zero when absent
  if (dsrc == nullptr)
    for (int i = tid; i < C * HW; i += kAttnThreads) s_D[i] = 0.f;
  __syncthreads();

  float* prow_row = prow + ((size_t)b * G + g) * Row::LD;

  if (btype == BR_SPECTRAL) {
    const float* s_g = s_v; const float* s_h = s_v + Cfg::ROW; const float* s_s = s_v + 2 * Cfg::ROW;
    float* s_dsv = s_w; float* s_du2 = s_w + Cfg::ROW; float* s_du1 = s_w + 2 * Cfg::ROW; float* s_dg = s_w + 3 * Cfg::ROW;
    (void)s_g;
    // D += dfeat/HW (mean over HW, global_spectral_pool :7-12);  ds[c] = sum_p D*r;  du2 = ds * s(1-s)
    if (HW <= 32) {
      for (int c = tid; c < C; c += kAttnThreads) {
        const float df = s_dfeat[c] / (float)HW;
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < HW; ++p) {
          const float d = s_D[c * HW + p] + df;
          s_D[c * HW + p] = d;
          a = fmaf(d, s_r[c * HW + p], a);
        }
        const float sv = s_s[c];
        s_du2[c] = a * sv * (1.f - sv);
      }
    } else {
      for (int c = warp; c < C; c += kAttnThreads / 32) {
        const float df = s_dfeat[c] / (float)HW;
        float a = 0.f;
        for (int p = lane; p < HW; p += 32) {
          const float d = s_D[c * HW + p] + df;
          s_D[c * HW + p] = d;
          a = fmaf(d, s_r[c * HW + p], a);
        }
        a = warp_sum(a);
        const float sv = s_s[c];
        if (lane == 0) s_du2[c] = a * sv * (1.f - sv);
      }
    }
    (void)s_dsv;
    __syncthreads();
    fc_cols(prm.w1d[g], C, C, s_du2, s_du1, s_part);
    if (tid < C) s_du1[tid] = (s_h[tid] > 0.f) ? s_du1[tid] : 0.f;
    __syncthreads();
    fc_cols(prm.w0d[g], C, C, s_du1, s_dg, s_part);
    if (tid < C) s_dg[tid] = s_dg[tid] / (float)HW;
    __syncthreads();
    for (int i = tid; i < C * HW; i += kAttnThreads) {
      const int c = i / HW;
      s_D[i] = fmaf(s_D[i], s_s[c], s_dg[c]);
    }
    for (int i = tid; i < C; i += kAttnThreads) { prow_row[i] = s_du2[i]; prow_row[C + i] = s_du1[i]; }
  } else if (btype == BR_SPATIAL) {
    constexpr int KS = Cfg::KS_SPATIAL, KK = KS * KS, WIN = Cfg::WIN;
    const float* s_q = s_v; const float* s_t = s_v + Cfg::ROW; const float* s_s = s_v + 2 * Cfg::ROW;
    float* s_dv2 = s_w; float* s_dv1 = s_w + Cfg::ROW; float* s_dq = s_w + 2 * Cfg::ROW;
    // route head-feature gradient through the class max-pool (first-max rule)
    for (int f = tid; f < 4 * C; f += kAttnThreads) {
      const int c = f >> 2, i = (f >> 1) & 1, j = f & 1;
      float best = -INFINITY; int barg = 0;
#pragma unroll
      for (int u = 0; u < WIN; ++u)
#pragma unroll
        for (int v = 0; v < WIN; ++v) {
          const int p = (i * WIN + u) * S + j * WIN + v;
          const float o = s_r[c * HW + p] * s_s[p];
          if (o > best) { best = o; barg = p; }
        }
      s_D[c * HW + barg] += s_dfeat[f];   // windows are disjoint: no race
    }
    __syncthreads();
    // ds[p] = sum_c D*r ; dv2 = ds * s(1-s)
    for (int p = tid; p < HW; p += kAttnThreads) {
      float a = 0.f;
#pragma unroll 8
      for (int c = 0; c < C; ++c) a = fmaf(s_D[c * HW + p], s_r[c * HW + p], a);
      const float sv = s_s[p];
      s_dv2[p] = a * sv * (1.f - sv);
    }
    __syncthreads();
    // dt = A2^T * dv2 (transposed stencil); dv1 = dt * (t>0)
    for (int p = tid; p < HW; p += kAttnThreads) {
      const int y = p / S, x = p % S;
      float a = 0.f;
      for (int u = 0; u < KS; ++u) {
        const int yy = y - u + KS / 2;
        if (yy < 0 || yy >= S) continue;
        for (int v = 0; v < KS; ++v) {
          const int xx = x - v + KS / 2;
          if (xx < 0 || xx >= S) continue;
          a = fmaf(__ldg(prm.st1[g] + u * KS + v), s_dv2[yy * S + xx], a);
        }
      }
      s_dv1[p] = (s_t[p] > 0.f) ? a : 0.f;
    }
    __syncthreads();
    for (int p = tid; p < HW; p += kAttnThreads) {
      const int y = p / S, x = p % S;
      float a = 0.f;
      for (int u = 0; u < KS; ++u) {
        const int yy = y - u + KS / 2;
        if (yy < 0 || yy >= S) continue;
        for (int v = 0; v < KS; ++v) {
          const int xx = x - v + KS / 2;
          if (xx < 0 || xx >= S) continue;
          a = fmaf(__ldg(prm.st0[g] + u * KS + v), s_dv1[yy * S + xx], a);
        }
      }
      s_dq[p] = (s_q[p] > 0.f) ? a : 0.f;
    }
    __syncthreads();
    // stencil / bias / channel-pool parameter partials for this crop
    for (int e = tid; e < 2 * KK + 2; e += kAttnThreads) {
      const bool second = e >= KK + 1;
      const int k = second ? e - (KK + 1) : e;
      const float* dv = second ? s_dv2 : s_dv1;
      const float* src = second ? s_t : s_q;
      float a = 0.f;
      if (k == KK) {
        for (int p = 0; p < HW; ++p) a += dv[p];
      } else {
        const int u = k / KS, v = k % KS;
        for (int y = 0; y < S; ++y) {
          const int yy = y + u - KS / 2;
          if (yy < 0 || yy >= S) continue;
          for (int x = 0; x < S; ++x) {
            const int xx = x + v - KS / 2;
            if (xx < 0 || xx >= S) continue;
            a = fmaf(dv[y * S + x], src[yy * S + xx], a);
          }
        }
      }
      prow_row[e] = a;
    }
    if (HW <= 32) {
      for (int c = tid; c < C; c += kAttnThreads) {
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < HW; ++p) a = fmaf(s_dq[p], s_r[c * HW + p], a);
        prow_row[2 * KK + 2 + c] = a;
      }
    } else {
      for (int c = warp; c < C; c += kAttnThreads / 32) {
        float a = 0.f;
        for (int p = lane; p < HW; p += 32) a = fmaf(s_dq[p], s_r[c * HW + p], a);
        a = warp_sum(a);
        if (lane == 0) prow_row[2 * KK + 2 + c] = a;
      }
    }
    if (warp == kAttnThreads / 32 - 1) {
      float a = 0.f;
      for (int p = lane; p < HW; p += 32) a += s_dq[p];
      a = warp_sum(a);
      if (lane == 0) prow_row[2 * KK + 2 + C] = a;
    }
    __syncthreads();
    // dr = D*s + pool_w[c]*dq
    for (int i = tid; i < C * HW; i += kAttnThreads) {
      const int c = i / HW, p = i - c * HW;
      s_D[i] = fmaf(s_D[i], s_s[p], __ldg(prm.pool_w[g] + c) * s_dq[p]);
    }
  } else {
    // vanilla: flatten feeds fc1 at block 3; blocks 1/2 pass the upstream gradient through
    if (has_head)
      for (int i = tid; i < C * HW; i += kAttnThreads) s_D[i] += s_dfeat[i];
  }
  if (POOL)   // BatchNorm mean / inverse std of this branch's channels for the per-cell loop below (s_part is idle by now)
    for (int i = tid; i < C; i += kAttnThreads) { s_part[i] = __ldg(mean + g * C + i); s_part[C + i] = __ldg(istd + g * C + i); }
  __syncthreads();

  // route through max-pool (argmax) and ReLU; emit da and the BatchNorm-backward partials
  float* da_out = da + ((size_t)b * G + g) * C * HWPRE;
  float* bn_row = bnrow + ((size_t)b * G + g) * 2 * C;
  if (POOL && compact) {
    // Pooled block, compact form: da is zero except at the arg-max of every 2x2 window, so only (value, slot) per pooled cell
    // travel to the pack kernel (tc_pack_stream_kernel<S, SRC_DZ, true>): a quarter of the bytes, contiguous stores.
    float* dvc = da + ((size_t)b * G + g) * C * HW;
    unsigned char* darg = reinterpret_cast<unsigned char*>(da + (size_t)gridDim.x * G * C * HW) + ((size_t)b * G + g) * C * HW;
    for (int i = tid; i < C * HW; i += kAttnThreads) {
      const int c = i / HW;
      const float dv = s_r[i] > 0.f ? s_D[i] : 0.f;
      dvc[i] = dv;
      darg[i] = s_arg[i];
      s_D[i] = dv;
      s_r[i] = dv * ((s_zarg[i] - s_part[c]) * s_part[C + c]);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kAttnThreads) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 5
      for (int cell = 0; cell < HW; ++cell) { s1 += s_D[c * HW + cell]; s2 += s_r[c * HW + cell]; }
      bn_row[c] = s1;
      bn_row[C + c] = s2;
    }
  } else if (POOL) {
    // One thread per pooled cell: only the arg-max position of its 2x2 window receives gradient.  The ReLU mask is r > 0
    // (r is the rectified maximum itself) and the conv output at the arg-max was kept by build_r: no second read of z.
    // Per-cell terms of the BatchNorm sums go back to shared memory (s_D: dv, s_r: dv * zhat) and are added per channel
    // by one thread each, in cell order.
    for (int e = tid; e < C * (2 * SPRE - 1); e += kAttnThreads) {   // row / column dropped by the floor pooling
      const int c = e / (2 * SPRE - 1), k = e - c * (2 * SPRE - 1);
      const int idx = k < SPRE ? (SPRE - 1) * SPRE + k : (k - SPRE) * SPRE + (SPRE - 1);
      da_out[c * HWPRE + idx] = 0.f;
    }
    for (int i = tid; i < C * HW; i += kAttnThreads) {
      const int c = i / HW, cell = i - c * HW;
      const int cy = cell / S, cx = cell - cy * S;
      const int base = c * HWPRE + (2 * cy) * SPRE + 2 * cx;
      const int arg = s_arg[i];
      const float dv = s_r[i] > 0.f ? s_D[i] : 0.f;
      da_out[base] = arg == 0 ? dv : 0.f;
      da_out[base + 1] = arg == 1 ? dv : 0.f;
      da_out[base + SPRE] = arg == 2 ? dv : 0.f;
      da_out[base + SPRE + 1] = arg == 3 ? dv : 0.f;
      s_D[i] = dv;
      s_r[i] = dv * ((s_zarg[i] - s_part[c]) * s_part[C + c]);
    }
    __syncthreads();
    for (int c = tid; c < C; c += kAttnThreads) {
      float s1 = 0.f, s2 = 0.f;
#pragma unroll 5
      for (int cell = 0; cell < HW; ++cell) { s1 += s_D[c * HW + cell]; s2 += s_r[c * HW + cell]; }
      bn_row[c] = s1;
      bn_row[C + c] = s2;
    }
  } else {
    for (int c = warp; c < C; c += kAttnThreads / 32) {
      const float mu = __ldg(mean + g * C + c), is = __ldg(istd + g * C + c);
      constexpr int NIT = (HWPRE + 31) / 32;
      float zv[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {      // second read of z (L2-resident) instead of keeping it in smem; all loads first
        const int pp = lane + 32 * it;
        zv[it] = pp < HWPRE ? __ldg(zg + c * HWPRE + pp) : 0.f;
      }
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int pp = lane + 32 * it;
        if (pp < HWPRE) {
          const float dv = s_r[c * HW + pp] > 0.f ? s_D[c * HW + pp] : 0.f;   // r = relu(bn(z)) > 0  <=>  bn(z) > 0
          da_out[c * HWPRE + pp] = dv;
          s1 += dv;
          s2 = fmaf(dv, (zv[it] - mu) * is, s2);
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) { bn_row[c] = s1; bn_row[C + c] = s2; }
    }
  }
}

template <int C, int SPRE, bool POOL>
size_t attn_bwd_smem(int classes) {
  using Cfg = AttnCfg<C, SPRE, POOL>;
  const size_t floats = (size_t)2 * C * Cfg::HW + 7 * Cfg::ROW + Cfg::FEAT_LD + kAttnThreads + ((classes + 3) / 4) * 4 +
                        (POOL ? (size_t)C * Cfg::HW : 0);
  return floats * sizeof(float) + (POOL ? (size_t)C * Cfg::HW : 0);
}

}  // namespace dta
