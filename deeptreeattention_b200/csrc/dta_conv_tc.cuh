// tcgen05 (5th-gen tensor core) implicit-GEMM 3x3 "same" convolution for sm_100a: forward and
// weight gradient, fp32-faithful through split-bf16 operands (v = hi + lo, products
// hi*hi + hi*lo + lo*hi [+ lo*lo], fp32 accumulation in tensor memory).  conv_impl = 1.
// Reference semantics: nn.Conv2d(k=3, padding="same"), Hang2020.py:18.
//
// Data layout ("position stream").  A crop plane of side S is laid out with one zero pad row
// above it and one zero pad column to its right: pitch PT = S+1, PC = (S+1)*PT positions per
// crop (S=11: 12 x 12 = 144).  The next crop's pad row is this crop's bottom padding, the pad
// column is the right padding of row y and the left padding of row y+1.  Crops follow each other,
// so the whole batch is ONE stream of positions q = b*PC + (y+1)*PT + x and every 3x3 tap is a
// constant shift s = dy*PT + dx of that stream.  Operands are pre-packed (dta_pack_* kernels) as
//     buf[half][chunk][row = GUARD + q][8 x bf16]     half: 0 = hi, 1 = lo;  chunk = 8 channels
// so (a) a tile's rows are contiguous in global memory -> plain 1-D bulk copies (TMA engine, no
// tensor map), (b) in shared memory the same bytes are a canonical no-swizzle UMMA operand, K-major
// (forward: row = M) or MN-major (wgrad: row = K), and (c) a tap is `start address += s * 16 B`
// (dta_tc.cuh; verified on hardware by tools/tc_probe.cu).
#pragma once
#include <type_traits>
#include "dta_common.cuh"
#include "dta_tc.cuh"

namespace dta {

constexpr int kTcThreads = 192;    // warp 0: bulk-copy producer, warp 1: MMA issuer, warps 2-5: epilogue

__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t desc_from(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" : : : "memory"); }

// Geometry of the position stream for plane side S.
template <int S>
struct Stream {
  static constexpr int PT = S + 1;          // pitch
  static constexpr int PC = (S + 1) * PT;   // positions per crop
};

// Rows of a packed operand buffer for `B` crops read in tiles of `tile` positions.
inline size_t tc_rows(int B, int pc, int tile) {
  const size_t n = (size_t)B * pc;
  return (n + tile - 1) / tile * tile + 2 * kTcGuard;
}

// =======================================================================================
// Packing kernels (HBM-bound elementwise; one thread = 8 channels x 1 stream position)
// =======================================================================================
// fp32 source -> split-bf16 position stream, dst[half][chunk][rows][8] (chunk = channel / 8 over the concatenated
// groups).  One thread = 8 channels x 1 stream position; all loads of a thread are issued before any use (the kernels
// are latency-bound otherwise).  Channels >= ctot, pad positions and guard rows are written as zeros.
//   MODE SRC_RAW: the crops.   SRC_ACT: relu(z*scale+shift) [2x2 max-pooled] * attention gate.
//   SRC_DZ: BatchNorm backward on the fly, k0*da + k1*z + k2.
template <int S, int MODE, bool POOL>
__global__ void __launch_bounds__(256)
tc_pack_stream_kernel(ConvSrc src, int G, int B, int nchunk, size_t rows, __nv_bfloat16* __restrict__ dst) {
  pdl_prologue();
  using St = Stream<S>;
  const size_t total = rows * nchunk;
  const int ctot = G * src.cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i / rows);
    const size_t row = i - (size_t)chunk * rows;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    const long long q = (long long)row - kTcGuard;
    const int ch0 = chunk * 8;
    if (q >= 0 && q < (long long)B * St::PC && ch0 < ctot) {
      const int b = (int)(q / St::PC);
      const int r = (int)(q - (long long)b * St::PC);
      const int yy = r / St::PT, xx = r - yy * St::PT;
      if (yy >= 1 && xx < S) {
        const int y = yy - 1, p = y * S + xx;
        const int nv = min(8, ctot - ch0);                  // live channels of this chunk
        if (MODE == SRC_RAW) {
          const float* a = src.a + ((size_t)b * src.ctot + ch0) * src.src_hw + p;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (j < nv) v[j] = __ldg(a + (size_t)j * src.src_hw);
        } else if (MODE == SRC_DZ) {
          const size_t base = ((size_t)b * src.ctot + ch0) * src.src_hw + p;
          float da[8], zz[8];
          if (POOL) {
            // compact upstream gradient of a pooled block: one value + arg-max slot per 2x2 window (3/4 of da is zero)
            constexpr int SP = S / 2;
            const bool in = y < 2 * SP && xx < 2 * SP;
            const int cell = (y >> 1) * SP + (xx >> 1), slot = ((y & 1) << 1) | (xx & 1);
            const size_t cb = ((size_t)b * src.ctot + ch0) * (SP * SP) + cell;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v = in ? __ldg(src.a + cb + (size_t)j * (SP * SP)) : 0.f;
              const int k = in ? (int)__ldg(src.arg + cb + (size_t)j * (SP * SP)) : -1;
              da[j] = k == slot ? v : 0.f;
              zz[j] = __ldg(src.b + base + (size_t)j * src.src_hw);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) { da[j] = __ldg(src.a + base + (size_t)j * src.src_hw); zz[j] = __ldg(src.b + base + (size_t)j * src.src_hw); }
          }
          const float4 k0a = __ldg(reinterpret_cast<const float4*>(src.k0 + ch0)), k0b = __ldg(reinterpret_cast<const float4*>(src.k0 + ch0) + 1);
          const float4 k1a = __ldg(reinterpret_cast<const float4*>(src.k1 + ch0)), k1b = __ldg(reinterpret_cast<const float4*>(src.k1 + ch0) + 1);
          const float4 k2a = __ldg(reinterpret_cast<const float4*>(src.k2 + ch0)), k2b = __ldg(reinterpret_cast<const float4*>(src.k2 + ch0) + 1);
          const float k0[8] = {k0a.x, k0a.y, k0a.z, k0a.w, k0b.x, k0b.y, k0b.z, k0b.w};
          const float k1[8] = {k1a.x, k1a.y, k1a.z, k1a.w, k1b.x, k1b.y, k1b.z, k1b.w};
          const float k2[8] = {k2a.x, k2a.y, k2a.z, k2a.w, k2b.x, k2b.y, k2b.z, k2b.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = k0[j] * da[j] + k1[j] * zz[j] + k2[j];   // same association as conv_src_load
        } else {
          constexpr int SP = POOL ? 2 * S + 1 : S;          // side of the source plane
          const int g = ch0 / src.cin;                       // cin is a multiple of 8 for activations
          const float* zp = src.a + ((size_t)b * src.ctot + ch0) * src.src_hw + (POOL ? (2 * y) * SP + 2 * xx : p);
          float z0[8], z1[8], z2[8], z3[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* zj = zp + (size_t)j * src.src_hw;
            z0[j] = __ldg(zj);
            if (POOL) { z1[j] = __ldg(zj + 1); z2[j] = __ldg(zj + SP); z3[j] = __ldg(zj + SP + 1); }
          }
          const float4 sca = __ldg(reinterpret_cast<const float4*>(src.k0 + ch0)), scb = __ldg(reinterpret_cast<const float4*>(src.k0 + ch0) + 1);
          const float4 sha = __ldg(reinterpret_cast<const float4*>(src.k1 + ch0)), shb = __ldg(reinterpret_cast<const float4*>(src.k1 + ch0) + 1);
          const float sc[8] = {sca.x, sca.y, sca.z, sca.w, scb.x, scb.y, scb.z, scb.w};
          const float sh[8] = {sha.x, sha.y, sha.z, sha.w, shb.x, shb.y, shb.z, shb.w};
          const int gm = src.gate_mode[g];
          const float* grow = src.gate + ((size_t)b * G + g) * src.gate_ld + src.gate_off;
          float gate_p = 1.f;
          if (gm == BR_SPATIAL) gate_p = __ldg(grow + p);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a = fmaxf(fmaf(z0[j], sc[j], sh[j]), 0.f);
            if (POOL) {
              const float a1 = fmaxf(fmaf(z1[j], sc[j], sh[j]), 0.f), a2 = fmaxf(fmaf(z2[j], sc[j], sh[j]), 0.f),
                          a3 = fmaxf(fmaf(z3[j], sc[j], sh[j]), 0.f);
              a = fmaxf(fmaxf(a, a1), fmaxf(a2, a3));
            }
            float gv = gate_p;
            if (gm == BR_SPECTRAL) gv = __ldg(grow + (ch0 - g * src.cin) + j);
            v[j] = (gm == BR_NONE) ? a : a * gv;
          }
        }
      }
    }
    uint4 hi, lo;
    tc::split2(v[0], v[1], hi.x, lo.x);
    tc::split2(v[2], v[3], hi.y, lo.y);
    tc::split2(v[4], v[5], hi.z, lo.z);
    tc::split2(v[6], v[7], hi.w, lo.w);
    reinterpret_cast<uint4*>(dst)[i] = hi;
    reinterpret_cast<uint4*>(dst)[total + i] = lo;
  }
}

// Zero the guard rows [0, GUARD) and the tail rows [GUARD + valid, rows) of every (half, chunk) plane of a packed stream
// whose crop rows are written by another kernel (the attention kernels pack their own output).
__global__ void tc_zero_guards_kernel(__nv_bfloat16* __restrict__ dst, size_t rows, int nchunk, size_t valid) {
  pdl_prologue();
  const size_t tail0 = kTcGuard + valid;
  const size_t per_plane = kTcGuard + (rows - tail0);
  const size_t total = per_plane * nchunk * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t plane = i / per_plane, r = i - plane * per_plane;
    const size_t row = r < (size_t)kTcGuard ? r : tail0 + (r - kTcGuard);
    reinterpret_cast<uint4*>(dst)[plane * rows + row] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// Forward-type weights -> per group, per 16-input-channel stage:
//     dst[group][stage][tap 9][kchunk 2][row 2*NCO: hi rows | lo rows][8 k]
// mode 0 (merged forward, conv1): one group, row = br*cout_b + co over both branches, k = ci
// mode 1 (grouped forward)      : group = branch, row = co, k = ci
// mode 2 (grouped input-gradient): group = branch, row = ci, k = co, tap flipped: the same GEMM
//         kernel then computes dIn[ci][p] = sum_{co,tap} dz[co][p - s_tap] W[co][ci][tap]
// w.p[br] = conv_layer.weight (cout_b, cin, 3, 3); rows / k beyond the tensor are zero.
template <int NCO>
__global__ void tc_pack_w_fprop_kernel(Ptr2 w, int nb, int cout_b, int cin, int nstage, int mode, __nv_bfloat16* __restrict__ dst) {
  pdl_prologue();
  const int G = mode == 0 ? 1 : nb;
  // one thread = one weight element: a single (gathering) load yields both its hi and its lo row
  const size_t total = (size_t)G * nstage * 9 * 2 * NCO * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int j = (int)(r % 8); r /= 8;
    const int n = (int)(r % NCO); r /= NCO;
    const int kc = (int)(r % 2); r /= 2;
    const int tap = (int)(r % 9); r /= 9;
    const int stage = (int)(r % nstage); r /= nstage;
    const int g = (int)r;
    const int k = stage * 16 + kc * 8 + j;
    float v = 0.f;
    if (mode == 0) {
      const int br = n / cout_b;
      if (br < nb && k < cin) v = __ldg(w.p[br] + ((size_t)(n - br * cout_b) * cin + k) * 9 + tap);
    } else if (mode == 1) {
      if (n < cout_b && k < cin) v = __ldg(w.p[g] + ((size_t)n * cin + k) * 9 + tap);
    } else {
      if (n < cin && k < cout_b) v = __ldg(w.p[g] + ((size_t)k * cin + n) * 9 + (8 - tap));
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const size_t o = ((((((size_t)g * nstage + stage) * 9 + tap) * 2 + kc) * (2 * NCO)) + n) * 8 + j;
    dst[o] = h;
    dst[o + (size_t)NCO * 8] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// =======================================================================================
// Forward-type convolution (also the input gradient, with mode-2 weights):
//   out[b][g*cout_g + n][p] = bias + sum_{k,tap} in[b][g][k][p + tap] * Wp[g][n][k][tap]
//   GEMM view: M = stream positions (tiles of SUB x 128), N = NCO, K = 16 input channels per stage x 9 taps.
//   Per (subtile, tap, stage): MMA1  A_hi x [W_hi ; W_lo]  (N = 2*NCO)  -> cols [0,NCO) += hi*hi, [NCO,2NCO) += hi*lo
//                              MMA2  A_lo x  W_hi          (N = NCO)    -> cols [0,NCO) += lo*hi
//   Epilogue: out = cols[0,NCO) + cols[NCO,2NCO) + bias.  Work items = (group, tile), persistent CTAs.
// =======================================================================================
constexpr int kTcFpropThreads = 320;   // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue (two per TMEM lane quadrant)

// ACC2: two accumulator stages of half a tile each, so the epilogue of one tile overlaps the MMAs of the
// next (used where K is short: conv2, conv3 and the input gradients); conv1 (K = 24 stages) keeps one
// full-size accumulator so the packed weights are re-streamed half as often.
template <int S, int NCO, bool ACC2>
struct TcFprop {
  using St = Stream<S>;
  static constexpr int NACC = ACC2 ? 2 : 1;
  static constexpr int ACC_COLS = 512 / NACC;
  static constexpr int SUB = ACC_COLS / (2 * NCO);       // 128-position subtiles per tile
  static constexpr int TILE = SUB * 128;
  static constexpr int AROWS = TILE + 2 * kTcGuard;      // rows staged per chunk
  static constexpr int A_BYTES = 2 * 2 * AROWS * 16;     // [half][kchunk][AROWS][16 B]
  static constexpr int W_BYTES = 9 * 2 * (2 * NCO) * 16; // [tap][kchunk][2*NCO][16 B]
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;
  static constexpr int NSTAGE = (222 * 1024) / STAGE_BYTES >= 4 ? 4 : (222 * 1024) / STAGE_BYTES;
  static constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_BYTES + 1024;
  static_assert(SUB >= 1, "accumulator stage too small");
  static_assert(NSTAGE >= 2, "stage too large");
  static_assert(kTcGuard >= St::PT + 1, "guard must cover the largest tap shift");
};

// FUSEX (conv1): the A operand is not read pre-packed; eight extra "converter" warps read the raw fp32 crops, split them
// into hi/lo bf16 straight into the shared-memory stage (generic-proxy stores + fence.proxy.async) and, when asked,
// also emit the packed position stream to global memory for the weight-gradient kernel -- the separate pack pass over
// the crops (183 MB read + 227 MB written) disappears into this kernel's shadow.
constexpr int kTcConvWarps = 8;
#ifndef DTA_X_PREFETCH
#define DTA_X_PREFETCH 2
#endif
constexpr int kXPrefetch = DTA_X_PREFETCH;   // stages the producer warp's L2 prefetch of the raw crops runs ahead (0 = off)
struct FuseX {
  const float* x;             // crops (B, bands, S, S) float32
  int bands;
  __nv_bfloat16* xp_out;      // packed stream to write ([2][nchunk][rows][8]) or null
};

template <int S, int NCO, bool ACC2, bool FUSEX = false>
__global__ void __launch_bounds__(kTcFpropThreads + (FUSEX ? 32 * kTcConvWarps : 0), 1)
tc_conv_fprop_kernel(const __nv_bfloat16* __restrict__ xp /*[2][nchunk][rows][8]*/, size_t rows, int nchunk, int chunks_per_group,
                     const __nv_bfloat16* __restrict__ wp /*[G][nstage][W_BYTES]*/, int nstage, Ptr2 bias, int bias_split,
                     float* __restrict__ out /*[B][out_ctot][S*S]*/, int out_ctot, int cout_g, int B, int ntiles, int G,
                     float* __restrict__ stats /*[gridDim.x*4][out_ctot][2] or null: BatchNorm partial sums of out*/, FuseX fx) {
  pdl_prologue();
  using Cfg = TcFprop<S, NCO, ACC2>;
  using St = Stream<S>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  // Accumulator hand-off.  ACC2: one full / empty barrier pair per accumulator stage.  !ACC2 (conv1, one accumulator of SUB
  // subtiles): one pair per SUBTILE -- the issuer signals subtile s as soon as its last MMA is queued and re-enters it for the
  // next tile as soon as it is drained, so the bubble between two tiles is one subtile's drain instead of the whole epilogue.
  constexpr int NHB = ACC2 ? 2 : Cfg::SUB;
  __shared__ uint64_t full_bar[4], empty_bar[4], tmem_full[NHB], tmem_empty[NHB];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[2][NCO];   // per group (G <= 2), filled once: no per-tile barrier in the epilogue

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int nwork = ntiles * G;

  if (tid == 0) {
    for (int i = 0; i < Cfg::NSTAGE; ++i) {
      tc::mbar_init(&full_bar[i], FUSEX ? 1 + 32 * kTcConvWarps : 1);   // W bulk copy (+ every converter thread)
      tc::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < NHB; ++i) { tc::mbar_init(&tmem_full[i], 1); tc::mbar_init(&tmem_empty[i], 8); }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  const bool has_bias = bias.p[0] != nullptr;
  for (int i = tid; i < 2 * NCO; i += blockDim.x) {
    const int g = i / NCO, n = i - g * NCO;
    const int chg = g * cout_g + n;
    s_bias[g][n] = (has_bias && g < G && n < cout_g) ? __ldg(bias.p[chg / bias_split] + (chg % bias_split)) : 0.f;
  }
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const size_t half_stride = (size_t)nchunk * rows * 8;   // elements between the hi and lo planes

  if (warp == 0) {
    // ---------------- producer: bulk copies global -> shared ----------------
    // FUSEX: the whole warp also prefetches the raw crops of the stage kXPrefetch steps ahead into L2 (one elected lane still
    // owns the barrier and the bulk copy of the weights).  The converter warps hold one stage of loads in registers at a
    // time, so their period is load latency + conversion; with the lines already in L2 that latency drops from a DRAM to an
    // L2 round trip and the MMAs stop waiting for operands.
    const uint32_t leader = elect_one_sync();
    auto prefetch_x = [&](int work, int ks) {
      if constexpr (FUSEX && kXPrefetch > 0) {
      while (ks >= nstage) { ks -= nstage; work += gridDim.x; }
      if (work >= nwork) return;
      const long long q0 = (long long)work * Cfg::TILE - kTcGuard, q1 = q0 + Cfg::AROWS;   // stream positions staged (FUSEX: one group)
      const int b_lo = (int)(q0 > 0 ? q0 / St::PC : 0);
      const int b_hi = (int)min((long long)B - 1, (q1 - 1) / St::PC);
      const int ch0 = ks * 16;
      const int nch = min(16, fx.bands - ch0);
      if (nch <= 0) return;
      const uintptr_t xbase = reinterpret_cast<uintptr_t>(fx.x);
      for (int b = b_lo; b <= b_hi; ++b) {
        const uintptr_t beg = xbase + ((size_t)b * fx.bands + ch0) * (S * S) * sizeof(float);
        const uintptr_t end = beg + (size_t)nch * (S * S) * sizeof(float);
        uintptr_t a = (beg & ~uintptr_t(127));
        if (a < xbase) a = xbase;
        for (a += (uintptr_t)lane * 128; a < end; a += 32 * 128) asm volatile("prefetch.global.L2 [%0];" : : "l"(a));
      }
      }
    };
    if (FUSEX)
      for (int ks = 1; ks < kXPrefetch; ++ks) prefetch_x(blockIdx.x, ks);
    uint32_t it = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      const int g = work / ntiles, tile = work - g * ntiles;
      const size_t row0 = (size_t)tile * Cfg::TILE;       // first staged row (= GUARD + q0 - GUARD)
      for (int ks = 0; ks < nstage; ++ks, ++it) {
        if (FUSEX) prefetch_x(work, ks + kXPrefetch);
        if (leader) {
          const int st = it % Cfg::NSTAGE;
          const uint32_t ph = (it / Cfg::NSTAGE) & 1;
          tc::mbar_wait(&empty_bar[st], ph ^ 1);
          unsigned char* sa = smem + (size_t)st * Cfg::STAGE_BYTES;
          tc::mbar_arrive_expect_tx(&full_bar[st], FUSEX ? Cfg::W_BYTES : Cfg::STAGE_BYTES);
          if (!FUSEX) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int kc = 0; kc < 2; ++kc)
                tc::bulk_g2s(sa + (size_t)(h * 2 + kc) * Cfg::AROWS * 16,
                             xp + h * half_stride + ((size_t)(g * chunks_per_group + ks * 2 + kc) * rows + row0) * 8, Cfg::AROWS * 16,
                             &full_bar[st]);
          }
          tc::bulk_g2s(sa + Cfg::A_BYTES, wp + ((size_t)g * nstage + ks) * (Cfg::W_BYTES / 2), Cfg::W_BYTES, &full_bar[st]);
        }
        if (FUSEX) __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (warp-uniform loop, one elected lane issues) ----------------
    const uint32_t leader = elect_one_sync();
    constexpr uint32_t idesc1 = tc::make_idesc_bf16(128, 2 * NCO, 0, 0);
    constexpr uint32_t idesc2 = tc::make_idesc_bf16(128, NCO, 0, 0);
    uint32_t it = 0, tile_it = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++tile_it) {
      const uint32_t acc = tile_it % Cfg::NACC, acc_use = tile_it / Cfg::NACC;
      if (ACC2) {
        tc::mbar_wait(&tmem_empty[acc], (acc_use & 1) ^ 1);
        tc::fence_after_sync();
      }
      const uint32_t dbase = tmem + acc * Cfg::ACC_COLS;
      for (int ks = 0; ks < nstage; ++ks, ++it) {
        const int st = it % Cfg::NSTAGE;
        const uint32_t ph = (it / Cfg::NSTAGE) & 1;
        tc::mbar_wait(&full_bar[st], ph);
        tc::fence_after_sync();
        // descriptors of row 0; the address field is the low 14 bits, so "+ rows" below moves the start
        const uint32_t sa = tc::smem_u32(smem + (size_t)st * Cfg::STAGE_BYTES);
        const uint64_t a_d = tc::sdesc_kmajor(sa, Cfg::AROWS);                              // A_hi plane
        const uint64_t b_d = tc::sdesc_kmajor(sa + Cfg::A_BYTES, 2 * NCO);
        const uint32_t a_lo32 = (uint32_t)a_d, a_hi32 = (uint32_t)(a_d >> 32);
        const uint32_t al_lo32 = a_lo32 + 2 * Cfg::AROWS;                                   // A_lo plane
        const uint32_t b_lo32 = (uint32_t)b_d, b_hi32 = (uint32_t)(b_d >> 32);
#pragma unroll
        for (int s = 0; s < Cfg::SUB; ++s) {
          if (!ACC2 && ks == 0) {             // subtile s of the previous tile must have been drained
            tc::mbar_wait(&tmem_empty[s], (tile_it & 1) ^ 1);
            tc::fence_after_sync();
          }
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int shift = kTcGuard + s * 128 + (t / 3 - 1) * St::PT + (t % 3 - 1);   // rows == 16-byte units
            const uint32_t boff = t * (2 * (2 * NCO));                                    // 16-byte units per tap
            if (leader) {
              tc::mma_bf16(dbase + s * (2 * NCO), desc_from(a_lo32 + shift, a_hi32), desc_from(b_lo32 + boff, b_hi32), idesc1,
                           (ks | t) ? 1u : 0u);
              tc::mma_bf16(dbase + s * (2 * NCO), desc_from(al_lo32 + shift, a_hi32), desc_from(b_lo32 + boff, b_hi32), idesc2, 1u);
            }
          }
          if (!ACC2 && ks == nstage - 1) {    // subtile s is complete once everything issued so far has run
            __syncwarp();
            if (leader) tc::mma_commit(&tmem_full[s]);
          }
        }
        __syncwarp();
        if (leader) tc::mma_commit(&empty_bar[st]);
      }
      if (ACC2 && leader) tc::mma_commit(&tmem_full[acc]);
      __syncwarp();
    }
  } else if (FUSEX && warp >= 10) {
    // ---------------- converters: fp32 crops -> split-bf16 A operand in shared memory (+ packed copy in global) ----------------
    constexpr int NT = 32 * kTcConvWarps;
    constexpr int UNITS = 2 * Cfg::AROWS;                       // (kchunk, row) pairs of 8 channels per stage
    constexpr int PER = (UNITS + NT - 1) / NT;
    const int ct = tid - kTcFpropThreads;
    const size_t lo_off = (size_t)nchunk * rows;                // uint4 units between the hi and lo planes in global memory
    uint4* xo = reinterpret_cast<uint4*>(fx.xp_out);
    uint32_t it = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x) {
      const int tile = work;                                    // FUSEX runs with one group
      const size_t row0 = (size_t)tile * Cfg::TILE;
      for (int ks = 0; ks < nstage; ++ks, ++it) {
        const int st = it % Cfg::NSTAGE;
        const uint32_t ph = (it / Cfg::NSTAGE) & 1;
        float v[PER][8];
        // 1. all global loads of this thread's units first (the stage buffer is not needed yet)
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int u = ct + k * NT;
          const int kc = u / Cfg::AROWS, r = u - kc * Cfg::AROWS;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
          const long long q = (long long)row0 + r - kTcGuard;
          if (u < UNITS && q >= 0 && q < (long long)B * St::PC) {
            const int b = (int)(q / St::PC);
            const int rr = (int)(q - (long long)b * St::PC);
            const int yy = rr / St::PT, xx = rr - yy * St::PT;
            const int ch0 = ks * 16 + kc * 8;
            if (yy >= 1 && xx < S && ch0 < fx.bands) {
              const float* src = fx.x + ((size_t)b * fx.bands + ch0) * (S * S) + (yy - 1) * S + xx;
              const int nv = min(8, fx.bands - ch0);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nv) v[k][j] = __ldg(src + (size_t)j * (S * S));
            }
          }
        }
        // 2. the stage buffer must be free (MMAs that read its previous contents have completed)
        tc::mbar_wait(&empty_bar[st], ph ^ 1);
        uint4* sa = reinterpret_cast<uint4*>(smem + (size_t)st * Cfg::STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < PER; ++k) {
          const int u = ct + k * NT;
          if (u < UNITS) {
            const int kc = u / Cfg::AROWS, r = u - kc * Cfg::AROWS;
            uint4 hi, lo;
            tc::split2(v[k][0], v[k][1], hi.x, lo.x);
            tc::split2(v[k][2], v[k][3], hi.y, lo.y);
            tc::split2(v[k][4], v[k][5], hi.z, lo.z);
            tc::split2(v[k][6], v[k][7], hi.w, lo.w);
            sa[kc * Cfg::AROWS + r] = hi;                              // [half 0][kc][AROWS]
            sa[(2 + kc) * Cfg::AROWS + r] = lo;                        // [half 1][kc][AROWS]
            // the tile's own rows go to the packed global copy; the first / last tile also own the stream's guard rows
            const bool own = (r >= kTcGuard && r < kTcGuard + Cfg::TILE) || (tile == 0 && r < kTcGuard) ||
                             (tile == ntiles - 1 && r >= kTcGuard + Cfg::TILE);
            if (xo != nullptr && own) {
              const size_t gi = (size_t)(ks * 2 + kc) * rows + row0 + r;
              xo[gi] = hi;
              xo[lo_off + gi] = lo;
            }
          }
        }
        tc::fence_async_smem();          // generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc::mbar_arrive(&full_bar[st]);
      }
    }
  } else if (warp < 10) {
    // ---------------- epilogue: TMEM -> registers -> out (NCHW fp32) ----------------
    const int ew = warp - 2;                   // 0..7
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int chalf = ew >> 2;                 // which half of the output channels
    constexpr int CH_PER = NCO / 2;
    constexpr int NGRP = (CH_PER + 31) / 32;   // 32-channel groups per thread (lane l ends up owning channel 32*grp + l)
    float run_sum[2][NGRP], run_sq[2][NGRP];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < NGRP; ++c) { run_sum[a][c] = 0.f; run_sq[a][c] = 0.f; }
    // BatchNorm partial sums of this thread's positions, kept in registers ACROSS the CTA's tiles: the warp-level butterfly
    // (62 shuffles per 32 channels, a quarter of the epilogue's instructions and most of its MIO pressure) runs once per
    // group instead of once per tile.  A CTA walks its work items in order, so the group changes at most G - 1 times.
    constexpr int GW = CH_PER < 32 ? CH_PER : 32;        // channels in a group
    float ps[NGRP][GW], pq[NGRP][GW];
#pragma unroll
    for (int grp = 0; grp < NGRP; ++grp)
#pragma unroll
      for (int j = 0; j < GW; ++j) { ps[grp][j] = 0.f; pq[grp][j] = 0.f; }
    auto flush_stats = [&](int gflush) {
#pragma unroll
      for (int grp = 0; grp < NGRP; ++grp) {
        // halving butterfly: after log2(32) rounds lane l holds the warp total of channel cg0 + (l % GW)
        float rs = 0.f, rq = 0.f;
        if (GW == 32) {
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) {
            const bool upper = (lane & w) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float keep_s = upper ? ps[grp][j + w] : ps[grp][j], send_s = upper ? ps[grp][j] : ps[grp][j + w];
              const float keep_q = upper ? pq[grp][j + w] : pq[grp][j], send_q = upper ? pq[grp][j] : pq[grp][j + w];
              ps[grp][j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, w);
              pq[grp][j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, w);
            }
          }
          rs = ps[grp][0]; rq = pq[grp][0];
        } else {   // GW == 16: plain xor-reduce of 16 values, lane l reports channel l % 16
#pragma unroll
          for (int j = 0; j < GW; ++j) {
            float a = ps[grp][j], c = pq[grp][j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); c += __shfl_xor_sync(0xffffffffu, c, o); }
            if ((lane % GW) == j) { rs = a; rq = c; }
          }
        }
        run_sum[gflush & 1][grp] += rs;
        run_sq[gflush & 1][grp] += rq;
#pragma unroll
        for (int j = 0; j < GW; ++j) { ps[grp][j] = 0.f; pq[grp][j] = 0.f; }
      }
    };
    int cur_g = -1;
    uint32_t tile_it = 0;
    for (int work = blockIdx.x; work < nwork; work += gridDim.x, ++tile_it) {
      const int g = work / ntiles, tile = work - g * ntiles;
      if (stats != nullptr && g != cur_g && cur_g >= 0) flush_stats(cur_g);
      cur_g = g;
      const uint32_t acc = tile_it % Cfg::NACC, acc_use = tile_it / Cfg::NACC;
      const float* sb = s_bias[g & 1];
      if (ACC2) {
        tc::mbar_wait(&tmem_full[acc], acc_use & 1);
        tc::fence_after_sync();
      }
      constexpr int NCC = GW / 16;                 // 16-channel blocks per group
      constexpr int NIT = Cfg::SUB * NCC;          // (subtile, 16-channel block) steps per group, walked two at a time
      constexpr int CCB = NCC == 2 ? 16 : 0;       // channel offset of the odd steps
      static_assert(NIT % 2 == 0 && (NCC == 1 || NCC == 2), "the epilogue walks its steps in pairs");
      static_assert(ACC2 || (NGRP == 1 && NCC == 2), "per-subtile hand-off: one step pair per subtile");
      const uint32_t tbase = tmem + ((uint32_t)(quad * 32) << 16) + acc * Cfg::ACC_COLS;
#pragma unroll
      for (int grp = 0; grp < NGRP; ++grp) {
        const int cg0 = chalf * CH_PER + grp * 32;
        const bool live = cg0 < cout_g;            // warp-uniform: this warp's channels exist
        auto issue = [&](int k, int cc, float* v0, float* v1) {
          const int s = k / NCC;
          const uint32_t taddr = tbase + s * (2 * NCO) + cg0 + cc;
          tc::tmem_ld16(taddr, v0);
          tc::tmem_ld16(taddr + NCO, v1);
        };
        auto process = [&](int k, auto ccv, const float* v0, const float* v1) {
          constexpr int cc = decltype(ccv)::value;
          const int s = k / NCC;
          const int c0 = cg0 + cc;
          const long long q = (long long)tile * Cfg::TILE + s * 128 + quad * 32 + lane;
          const int b = (int)(q / St::PC);
          const int r = (int)(q - (long long)b * St::PC);
          const int yy = r / St::PT, xx = r - yy * St::PT;
          const bool valid = b < B && yy >= 1 && xx < S;
          if (valid && c0 < cout_g) {
            float* orow = out + ((size_t)b * out_ctot + (size_t)g * cout_g + c0) * (S * S) + (yy - 1) * S + xx;
            float bs[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + j);
              bs[j] = b4.x; bs[j + 1] = b4.y; bs[j + 2] = b4.z; bs[j + 3] = b4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float o = v0[j] + v1[j] + bs[j];
              orow[(size_t)j * (S * S)] = o;
              ps[grp][cc + j] += o;
              pq[grp][cc + j] = fmaf(o, o, pq[grp][cc + j]);
            }
          }
        };
        float va0[16], va1[16];
#pragma unroll 1
        for (int k = 0; k < NIT; k += 2) {
          if (!ACC2) {                             // one pair of steps = one subtile
            tc::mbar_wait(&tmem_full[k / 2], tile_it & 1);
            tc::fence_after_sync();
          }
          if (live) {
            issue(k, 0, va0, va1);
            tc::tmem_ld_wait();
            process(k, std::integral_constant<int, 0>{}, va0, va1);
            issue(k + 1, CCB, va0, va1);
            tc::tmem_ld_wait();
            process(k + 1, std::integral_constant<int, CCB>{}, va0, va1);
          }
          if (!ACC2) {
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&tmem_empty[k / 2]);
          }
        }
      }
      if (ACC2) {
        tc::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&tmem_empty[acc]);
      }
    }
    if (stats != nullptr && cur_g >= 0) flush_stats(cur_g);
    if (stats != nullptr) {
      // partial statistics of this warp's positions: row (cta*4 + quad), channel (g*cout_g + chalf*CH_PER + 32*grp + lane)
      float* srow = stats + ((size_t)blockIdx.x * 4 + quad) * out_ctot * 2;
      for (int g = 0; g < G && g < 2; ++g)
#pragma unroll
        for (int grp = 0; grp < NGRP; ++grp) {
          const int ch = chalf * CH_PER + grp * 32 + lane;
          if (lane < GW && ch < cout_g) {
            srow[(size_t)(g * cout_g + ch) * 2 + 0] = run_sum[g][grp];
            srow[(size_t)(g * cout_g + ch) * 2 + 1] = run_sq[g][grp];
          }
        }
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// =======================================================================================
// Weight gradient: dW[g][co][ci][tap] = sum_q dz[q][g][co] * in[q + s_tap][g][ci]   (q over the whole stream)
//   GEMM view per tap: M = 128 rows of dz (A, MN-major, start shifted by -s_tap), N = NCI input channels
//   of this CTA's slice (B, MN-major), K = stream positions (KROWS per stage).
//   STACK (64 output channels per group): A rows = [dz_hi co ; dz_lo co]; MMAs with B = in_hi and in_lo
//       -> rows co hold dz_hi*in, rows 64+co hold dz_lo*in (all four partial products), summed in the epilogue.
//   !STACK (128 output channels per group): MMAs (dz_hi,in_hi), (dz_hi,in_lo), (dz_lo,in_hi) into the same rows.
//   9 taps x NCI columns of tensor memory.  Split-K over the stream: part[split][g][co][ci][tap], summed in
//   fixed order by wgrad_reduce_kernel.
// =======================================================================================
//   NTG tap groups: the 9 taps are split over NTG CTAs (blockIdx.x = slice*NTG + group) that stream the same operands
//   (the second reader hits L2); fewer taps per CTA leave tensor-memory room for wider N, and the per-MMA cost is
//   32 + N/4 cycles of shared-memory operand reads against an N/2 issue floor (tools/tc_rate.cu), so wider N is faster.
//   WIDE: the staged input slice is [in_hi chunks | in_lo chunks] with ONE chunk stride, so an MMA of N = 2 * NCI reads both
//   halves as a single operand: columns [0, NCI) collect dz * in_hi, [NCI, 2 NCI) dz * in_lo, added in the epilogue.  For the
//   32-channel slices of conv2 / conv3 that is one 48-cycle MMA instead of two 40-cycle ones per tap and K step.
template <int NCI_, int KROWS_, bool STACK_, int NTG_ = 1, int NSTAGE_ = 3, bool WIDE_ = false>
struct TcWgrad {
  static constexpr int NCI = NCI_;                         // input channels per CTA slice
  static constexpr bool WIDE = WIDE_;
  static constexpr int NW = WIDE ? 2 * NCI : NCI;          // fp32 columns of tensor memory per tap
  static constexpr int KROWS = KROWS_;                     // stream positions per stage
  static constexpr bool STACK = STACK_;
  static constexpr int NTG = NTG_;
  static constexpr int MAXT = (9 + NTG - 1) / NTG;         // taps per CTA (first groups take the larger share)
  static constexpr int COUT = STACK ? 64 : 128;            // output channels per group
  static constexpr int AROWS = KROWS + 2 * kTcGuard;       // dz rows staged per chunk
  static constexpr int ACH = STACK ? 16 : 32;              // A chunks: [hi | lo]
  static constexpr int A_BYTES = ACH * AROWS * 16;
  static constexpr int BCH = NCI / 8;
  static constexpr int B_BYTES = 2 * BCH * KROWS * 16;     // [half][BCH][KROWS][16 B]
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int NSTAGE = NSTAGE_;
  static constexpr int RED_BYTES = STACK ? 64 * NCI * 4 : 0;   // lo-half partials for the epilogue; aliases stage 0 (idle by then)
  static constexpr size_t SMEM_BYTES = (size_t)NSTAGE * STAGE_BYTES + 1024;
  static_assert(RED_BYTES <= STAGE_BYTES, "epilogue scratch must fit one stage");
  static_assert(MAXT * NW <= 512 && NCI % 16 == 0 && NW <= 256, "taps x columns per tap must fit tensor memory");
  static_assert(SMEM_BYTES <= 227 * 1024, "stage too large");
};

template <int S, class Cfg>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_conv_wgrad_kernel(const __nv_bfloat16* __restrict__ dzp /*[2][dz_chunks][rows][8]*/, int dz_chunks,
                     const __nv_bfloat16* __restrict__ xp /*[2][x_chunks][rows][8]*/, int x_chunks, size_t rows,
                     int cin_g, int cout_g, int nkstage_total, int stages_per_split, float* __restrict__ part /*[nsplit][G][cout_g][cin_g][9]*/) {
  pdl_prologue();
  using St = Stream<S>;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  float* s_red = reinterpret_cast<float*>(smem);   // stage 0, reused once every MMA has completed (after tmem_full)
  __shared__ uint64_t full_bar[Cfg::NSTAGE], empty_bar[Cfg::NSTAGE], tmem_full;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const int slice = blockIdx.x / Cfg::NTG, tgroup = blockIdx.x % Cfg::NTG;
  const int split = blockIdx.y, g = blockIdx.z, G = gridDim.z;
  const int tap0 = tgroup * Cfg::MAXT;                          // this CTA's taps [tap0, tap0 + ntap)
  const int ntap = min(Cfg::MAXT, 9 - tap0);
  const int ks_begin = split * stages_per_split;
  const int ks_end = min(nkstage_total, ks_begin + stages_per_split);
  const int nks = max(0, ks_end - ks_begin);
  constexpr int DZ_CPG = Cfg::COUT / 8;            // dz chunks per group

  if (tid == 0) {
    for (int i = 0; i < Cfg::NSTAGE; ++i) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], 1); }
    tc::mbar_init(&tmem_full, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_s, 512);
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (elect_one_sync()) {
      const size_t dz_half = (size_t)dz_chunks * rows * 8, x_half = (size_t)x_chunks * rows * 8;
      const int x_chunk0 = g * (cin_g / 8) + slice * Cfg::BCH;
      for (int i = 0; i < nks; ++i) {
        const int st = i % Cfg::NSTAGE;
        const uint32_t ph = (i / Cfg::NSTAGE) & 1;
        tc::mbar_wait(&empty_bar[st], ph ^ 1);
        unsigned char* sa = smem + (size_t)st * Cfg::STAGE_BYTES;
        unsigned char* sb = sa + Cfg::A_BYTES;
        const size_t k0 = (size_t)(ks_begin + i) * Cfg::KROWS;     // first stream position of the stage
        tc::mbar_arrive_expect_tx(&full_bar[st], Cfg::STAGE_BYTES);
        // dz rows [GUARD + k0 - GUARD, +AROWS): hi chunks then lo chunks
        for (int c = 0; c < Cfg::ACH; ++c) {
          const int h = c / DZ_CPG, cc = c - h * DZ_CPG;
          tc::bulk_g2s(sa + (size_t)c * Cfg::AROWS * 16, dzp + h * dz_half + ((size_t)(g * DZ_CPG + cc) * rows + k0) * 8, Cfg::AROWS * 16,
                       &full_bar[st]);
        }
        // input rows [GUARD + k0, +KROWS) of this slice's chunks
        for (int h = 0; h < 2; ++h)
          for (int c = 0; c < Cfg::BCH; ++c)
            tc::bulk_g2s(sb + (size_t)(h * Cfg::BCH + c) * Cfg::KROWS * 16,
                         xp + h * x_half + ((size_t)(x_chunk0 + c) * rows + kTcGuard + k0) * 8, Cfg::KROWS * 16, &full_bar[st]);
      }
    }
  } else if (warp == 1) {
    const uint32_t leader = elect_one_sync();
    constexpr uint32_t idesc = tc::make_idesc_bf16(128, Cfg::NCI, 1, 1);
    constexpr uint32_t idesc_wide = tc::make_idesc_bf16(128, Cfg::NW, 1, 1);
    for (int i = 0; i < nks; ++i) {
      const int st = i % Cfg::NSTAGE;
      const uint32_t ph = (i / Cfg::NSTAGE) & 1;
      tc::mbar_wait(&full_bar[st], ph);
      tc::fence_after_sync();
      const uint32_t sa = tc::smem_u32(smem + (size_t)st * Cfg::STAGE_BYTES);
      const uint64_t a_d = tc::sdesc_mnmajor(sa, Cfg::AROWS);
      const uint64_t b_d = tc::sdesc_mnmajor(sa + Cfg::A_BYTES, Cfg::KROWS);
      const uint32_t a_lo32 = (uint32_t)a_d, a_hi32 = (uint32_t)(a_d >> 32);
      const uint32_t b_lo32 = (uint32_t)b_d, b_hi32 = (uint32_t)(b_d >> 32);
      constexpr uint32_t A_LO_PLANE = 16 * Cfg::AROWS;          // !STACK: rows (16 B units) from dz_hi to dz_lo
      constexpr uint32_t B_LO_PLANE = Cfg::BCH * Cfg::KROWS;    // in_hi -> in_lo
#pragma unroll
      for (int kk = 0; kk < Cfg::KROWS / 16; ++kk) {
#pragma unroll
        for (int tl = 0; tl < Cfg::MAXT; ++tl) {
          const int t = tap0 + tl;
          const int arow = kTcGuard + kk * 16 - ((t / 3 - 1) * St::PT + (t % 3 - 1));   // dz row for input row kk*16
          if (leader && tl < ntap) {
            const uint32_t dcol = tmem + tl * Cfg::NW;
            if (Cfg::WIDE) {
              tc::mma_bf16(dcol, desc_from(a_lo32 + arow, a_hi32), desc_from(b_lo32 + kk * 16, b_hi32), idesc_wide, (i | kk) ? 1u : 0u);
            } else {
              tc::mma_bf16(dcol, desc_from(a_lo32 + arow, a_hi32), desc_from(b_lo32 + kk * 16, b_hi32), idesc, (i | kk) ? 1u : 0u);
              tc::mma_bf16(dcol, desc_from(a_lo32 + arow, a_hi32), desc_from(b_lo32 + B_LO_PLANE + kk * 16, b_hi32), idesc, 1u);
            }
            if (!Cfg::STACK)
              tc::mma_bf16(dcol, desc_from(a_lo32 + A_LO_PLANE + arow, a_hi32), desc_from(b_lo32 + kk * 16, b_hi32), idesc, 1u);
          }
        }
      }
      __syncwarp();
      if (leader) tc::mma_commit(&empty_bar[st]);
    }
    if (leader) tc::mma_commit(&tmem_full);
    __syncwarp();
  } else {
    // epilogue -> part[split][g][co][ci][tap]
    const int quad = warp & 3;
    const int row = quad * 32 + lane;          // TMEM lane = A row
    if (nks > 0) {
      tc::mbar_wait(&tmem_full, 0);
      tc::fence_after_sync();
    }
    for (int tl = 0; tl < ntap; ++tl) {
      const int t = tap0 + tl;
      float v[Cfg::NCI];
      if (nks > 0) {
#pragma unroll
        for (int c0 = 0; c0 < Cfg::NCI; c0 += 16) tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + tl * Cfg::NW + c0, v + c0);
        if (Cfg::WIDE) {      // add the dz * in_lo columns
          float w[Cfg::NCI];
#pragma unroll
          for (int c0 = 0; c0 < Cfg::NCI; c0 += 16) tc::tmem_ld16(tmem + ((uint32_t)(quad * 32) << 16) + tl * Cfg::NW + Cfg::NCI + c0, w + c0);
          tc::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < Cfg::NCI; ++j) v[j] += w[j];
        }
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < Cfg::NCI; ++j) v[j] = 0.f;
      }
      if (Cfg::STACK) {
        if (row >= 64) {
#pragma unroll
          for (int j = 0; j < Cfg::NCI; ++j) s_red[(row - 64) * Cfg::NCI + j] = v[j];
        }
        epi_bar_sync();
        if (row < 64) {
#pragma unroll
          for (int j = 0; j < Cfg::NCI; ++j) v[j] += s_red[row * Cfg::NCI + j];
        }
      }
      if (row < Cfg::COUT && row < cout_g) {
        // partial layout [split][g][tap][ci][co]: a warp's 32 rows (co) are contiguous -> coalesced stores
        float* dst = part + ((((size_t)split * G + g) * 9 + t) * cin_g + (size_t)slice * Cfg::NCI) * cout_g + row;
#pragma unroll
        for (int j = 0; j < Cfg::NCI; ++j)
          if (slice * Cfg::NCI + j < cin_g) dst[(size_t)j * cout_g] = v[j];
      }
      if (Cfg::STACK) epi_bar_sync();
    }
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, 512);
}

// dW[g][co][ci][tap] = sum_s part[s][g][tap][ci][co] (fixed order).  Reads are coalesced along co; the destination is the
// reference's (cout, cin, 3, 3) tensor, or two of them when one group covers both branches (conv1).
__global__ void tc_wgrad_reduce_kernel(const float* __restrict__ part, int nsplit, int G, int cout_g, int cin_g, MutPtr2 dw,
                                       size_t ptr_split /*elements per destination tensor*/) {
  pdl_prologue();
  const size_t per_split = (size_t)G * 9 * cin_g * cout_g;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_split; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += __ldg(part + (size_t)k * per_split + i);
    size_t r = i;
    const int co = (int)(r % cout_g); r /= cout_g;
    const int ci = (int)(r % cin_g); r /= cin_g;
    const int t = (int)(r % 9); r /= 9;
    const int g = (int)r;
    const size_t flat = (((size_t)g * cout_g + co) * cin_g + ci) * 9 + t;   // index in [G][cout][cin][9]
    const size_t q = flat / ptr_split;
    float* d = dw.p[q];
    if (d != nullptr) d[flat - q * ptr_split] = s;
  }
}

}  // namespace dta
