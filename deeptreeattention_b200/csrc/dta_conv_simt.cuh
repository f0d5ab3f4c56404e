// fp32 CUDA-core direct 3x3 "same" convolution (forward / input-gradient / weight-gradient).
// This is conv_impl = 0: the exact-fp32 path, also the in-library cross-check for the
// tcgen05 implicit-GEMM path.  Reference semantics: nn.Conv2d(k=3, padding="same"),
// Hang2020.py:18.
#pragma once
#include "dta_common.cuh"

namespace dta {

// ---------------------------------------------------------------------------------------
// Forward-type convolution: out[b][g*COUT+co][p] = bias + sum_{ci,tap} in[b][g][ci][p+tap] *
// Wp[g][ci][tap][co].  The weight-transposed/flipped table makes the same kernel the dgrad.
//   S      plane side (11 or 5), NCROP crops per CTA, thread tile TP positions x TC channels
//   CK     input channels staged per shared-memory chunk
// Epilogue (stats != nullptr): per-CTA per-channel sum / sum-of-squares for BatchNorm.
// ---------------------------------------------------------------------------------------
template <int S, int NCROP, int TP, int TC, int COUT, int CK>
struct FpropCfg {
  static constexpr int PS = S + 2;
  static constexpr int HW = S * S;
  static constexpr int P = NCROP * HW;
  static constexpr int PG = (P + TP - 1) / TP;
  static constexpr int CG = COUT / TC;
  static constexpr int NT = ((PG * CG + 31) / 32) * 32;
  static constexpr int IN_PLANE = NCROP * PS * PS;
  static constexpr int IN_FLOATS = CK * IN_PLANE;
  static constexpr int W_FLOATS = CK * 9 * COUT;
  static constexpr int RED_FLOATS = PG * COUT * 2;
  static constexpr int STAGE_FLOATS = IN_FLOATS + W_FLOATS;
  static constexpr int SMEM_FLOATS = STAGE_FLOATS > RED_FLOATS ? STAGE_FLOATS : RED_FLOATS;
  static constexpr size_t SMEM_BYTES = (size_t)SMEM_FLOATS * sizeof(float);
  static_assert(COUT % TC == 0, "COUT must be a multiple of the channel tile");
  static_assert(TC % 4 == 0, "channel tile must allow float4 weight loads");
};

template <int S, int NCROP, int TP, int TC, int COUT, int CK>
__global__ void __launch_bounds__(FpropCfg<S, NCROP, TP, TC, COUT, CK>::NT)
conv3x3_fprop_simt(ConvSrc src, const float* __restrict__ wp /*[G][Cin][9][COUT]*/, Ptr2 bias,
                   int bias_split /*channels per bias pointer*/, float* __restrict__ out, int out_ctot,
                   float* __restrict__ stats /*[gridDim.x][G*COUT][2] or null*/, int B) {
  pdl_prologue();
  using Cfg = FpropCfg<S, NCROP, TP, TC, COUT, CK>;
  constexpr int PS = Cfg::PS, HW = Cfg::HW;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_w = smem + Cfg::IN_FLOATS;

  const int tid = threadIdx.x;
  const int g = blockIdx.y, G = gridDim.y;
  const int b0 = blockIdx.x * NCROP;
  const int cg = tid % Cfg::CG;
  const int pg = tid / Cfg::CG;
  const bool active = pg < Cfg::PG;
  const int Cin = src.cin;

  int off[TP];
  bool valid[TP];
#pragma unroll
  for (int j = 0; j < TP; ++j) {
    const int m = pg * TP + j;
    const int crop = m / HW;
    const int p = m - crop * HW;
    const int y = p / S, x = p - y * S;
    valid[j] = active && (m < Cfg::P) && (b0 + crop < B);
    off[j] = (m < Cfg::P) ? (crop * PS * PS + y * PS + x) : 0;
  }

  float acc[TP][TC];
#pragma unroll
  for (int j = 0; j < TP; ++j)
#pragma unroll
    for (int c = 0; c < TC; ++c) acc[j][c] = 0.f;

  // zero the padded planes once; only interiors are rewritten per chunk
  for (int i = tid; i < Cfg::IN_FLOATS; i += Cfg::NT) s_in[i] = 0.f;

  const float* wg = wp + (size_t)g * Cin * 9 * COUT;
  for (int c0 = 0; c0 < Cin; c0 += CK) {
    __syncthreads();
    for (int i = tid; i < CK * NCROP * HW; i += Cfg::NT) {
      const int ci = i / (NCROP * HW);
      const int r = i - ci * (NCROP * HW);
      const int crop = r / HW;
      const int p = r - crop * HW;
      const int y = p / S, x = p - y * S;
      float v = 0.f;
      if (c0 + ci < Cin && b0 + crop < B) v = conv_src_load<S>(src, b0 + crop, g, G, c0 + ci, p);
      s_in[ci * Cfg::IN_PLANE + crop * PS * PS + (y + 1) * PS + (x + 1)] = v;
    }
    {
      const int nck = min(CK, Cin - c0);
      const float* wsrc = wg + (size_t)c0 * 9 * COUT;
      for (int i = tid; i < CK * 9 * COUT; i += Cfg::NT) s_w[i] = (i < nck * 9 * COUT) ? __ldg(wsrc + i) : 0.f;
    }
    __syncthreads();
    if (active) {
#pragma unroll 2
      for (int ci = 0; ci < CK; ++ci) {
        const float* ip = s_in + ci * Cfg::IN_PLANE;
        const float* wq = s_w + ci * 9 * COUT + cg * TC;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int toff = (t / 3) * PS + (t % 3);
          float wv[TC];
#pragma unroll
          for (int c = 0; c < TC; c += 4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wq + t * COUT + c);
            wv[c] = w4.x; wv[c + 1] = w4.y; wv[c + 2] = w4.z; wv[c + 3] = w4.w;
          }
#pragma unroll
          for (int j = 0; j < TP; ++j) {
            const float xv = ip[off[j] + toff];
#pragma unroll
            for (int c = 0; c < TC; ++c) acc[j][c] = fmaf(xv, wv[c], acc[j][c]);
          }
        }
      }
    }
  }

  // epilogue: bias, store, optional BatchNorm partial statistics
  float bsum[TC], bsq[TC];
#pragma unroll
  for (int c = 0; c < TC; ++c) {
    const int ch = g * COUT + cg * TC + c;
    float bv = 0.f;
    if (bias.p[0] != nullptr) bv = __ldg(bias.p[ch / bias_split] + (ch % bias_split));
    bsum[c] = 0.f; bsq[c] = 0.f;
#pragma unroll
    for (int j = 0; j < TP; ++j) {
      if (valid[j]) {
        const int m = pg * TP + j;
        const int crop = m / HW;
        const int p = m - crop * HW;
        const float v = acc[j][c] + bv;
        out[((size_t)(b0 + crop) * out_ctot + ch) * HW + p] = v;
        bsum[c] += v;
        bsq[c] = fmaf(v, v, bsq[c]);
      }
    }
  }
  if (stats != nullptr) {
    __syncthreads();  // staging buffers are dead, reuse as reduction scratch
    float* s_red = smem;
    if (active) {
#pragma unroll
      for (int c = 0; c < TC; ++c) {
        s_red[(pg * COUT + cg * TC + c) * 2 + 0] = bsum[c];
        s_red[(pg * COUT + cg * TC + c) * 2 + 1] = bsq[c];
      }
    }
    __syncthreads();
    for (int i = tid; i < COUT * 2; i += Cfg::NT) {
      float s = 0.f;
      for (int q = 0; q < Cfg::PG; ++q) s += s_red[q * COUT * 2 + i];
      stats[((size_t)blockIdx.x * G * COUT + g * COUT) * 2 + i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradient: part[split][g][co][ci][tap] = sum_{b in split, p} dz[b][g][co][p] *
// in[b][g][ci][p+tap]; a second kernel adds the splits in fixed order (deterministic).
//   block = CIK input channels x (COUT/TCO) channel groups; 9*TCO accumulators per thread.
// ---------------------------------------------------------------------------------------
template <int S, int CIK, int COUT, int TCO>
struct WgradCfg {
  static constexpr int PS = S + 2;
  static constexpr int HW = S * S;
  static constexpr int CG = COUT / TCO;
  static constexpr int NT = CIK * CG;
  static constexpr int IN_FLOATS = CIK * PS * PS;
  static constexpr int DZ_FLOATS = HW * COUT;
  static constexpr size_t SMEM_BYTES = (size_t)(IN_FLOATS + DZ_FLOATS) * sizeof(float);
  static_assert(TCO % 4 == 0 && COUT % TCO == 0, "bad channel tile");
  static_assert(NT % 32 == 0 && NT <= 1024, "bad block size");
};

template <int S, int CIK, int COUT, int TCO>
__global__ void __launch_bounds__(WgradCfg<S, CIK, COUT, TCO>::NT)
conv3x3_wgrad_simt(ConvSrc in, ConvSrc dz, float* __restrict__ part /*[nsplit][G][COUT][Cin][9]*/, int B,
                   int crops_per_split) {
  pdl_prologue();
  using Cfg = WgradCfg<S, CIK, COUT, TCO>;
  constexpr int PS = Cfg::PS, HW = Cfg::HW;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;
  float* s_dz = smem + Cfg::IN_FLOATS;  // [HW][COUT]

  const int tid = threadIdx.x;
  const int g = blockIdx.z, G = gridDim.z;
  const int split = blockIdx.y;
  const int c0 = blockIdx.x * CIK;
  const int ci_l = tid % CIK;
  const int cg = tid / CIK;
  const int Cin = in.cin;

  float acc[9][TCO];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < TCO; ++c) acc[t][c] = 0.f;

  for (int i = tid; i < Cfg::IN_FLOATS; i += Cfg::NT) s_in[i] = 0.f;

  const int b_begin = split * crops_per_split;
  const int b_end = min(B, b_begin + crops_per_split);
  for (int b = b_begin; b < b_end; ++b) {
    __syncthreads();
    for (int i = tid; i < CIK * HW; i += Cfg::NT) {
      const int ci = i / HW;
      const int p = i - ci * HW;
      const int y = p / S, x = p - y * S;
      float v = 0.f;
      if (c0 + ci < Cin) v = conv_src_load<S>(in, b, g, G, c0 + ci, p);
      s_in[ci * PS * PS + (y + 1) * PS + (x + 1)] = v;
    }
    for (int i = tid; i < COUT * HW; i += Cfg::NT) {
      const int co = i / HW;
      const int p = i - co * HW;
      s_dz[p * COUT + co] = conv_src_load<S>(dz, b, g, G, co, p);
    }
    __syncthreads();
    const float* ip = s_in + ci_l * PS * PS;
    for (int y = 0; y < S; ++y) {
      for (int x = 0; x < S; ++x) {
        float dv[TCO];
        const float* dq = s_dz + (y * S + x) * COUT + cg * TCO;
#pragma unroll
        for (int c = 0; c < TCO; c += 4) {
          const float4 d4 = *reinterpret_cast<const float4*>(dq + c);
          dv[c] = d4.x; dv[c + 1] = d4.y; dv[c + 2] = d4.z; dv[c + 3] = d4.w;
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const float iv = ip[(y + t / 3) * PS + x + (t % 3)];
#pragma unroll
          for (int c = 0; c < TCO; ++c) acc[t][c] = fmaf(iv, dv[c], acc[t][c]);
        }
      }
    }
  }

  if (c0 + ci_l < Cin) {
#pragma unroll
    for (int c = 0; c < TCO; ++c) {
      const int co = cg * TCO + c;
      float* dst = part + ((((size_t)split * G + g) * COUT + co) * Cin + (c0 + ci_l)) * 9;
#pragma unroll
      for (int t = 0; t < 9; ++t) dst[t] = acc[t][c];
    }
  }
}

// dW[g-th pointer][i] = sum_s part[s][g][i]   (fixed order)
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int G, size_t per_group,
                                    MutPtr2 dw, size_t ptr_split /*elements per destination tensor*/) {
  pdl_prologue();
  const size_t total = (size_t)G * per_group;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(size_t)k * total + i];
    // destination: group-major tensors, or one group split across two tensors (conv1)
    const size_t q = i / ptr_split;
    float* d = dw.p[q];
    if (d != nullptr) d[i - q * ptr_split] = s;
  }
}

}  // namespace dta
