// C ABI of libdta_b200.so: context, buffer-size arithmetic and the forward / backward
// launch sequences (see include/dta_b200.h for the contract and DESIGN.md for the plan).
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "dta_attention.cuh"
#include "dta_collective.cuh"
#include "dta_common.cuh"
#include "dta_ctx.cuh"
#include "dta_conv_simt.cuh"
#include "dta_conv_tc.cuh"
#include "dta_loss.cuh"
#include "dta_misc.cuh"
#include "dta_preprocess.cuh"

using namespace dta;

static std::string g_create_error;

namespace {

constexpr int kC[3] = {32, 64, 128};
constexpr int kHWpre[3] = {121, 121, 25};   // conv output plane per block
constexpr int kAttRow[3] = {121, 64, 128};  // AttnCfg::ROW
constexpr int kFeatLd[3] = {128, 256, 512};
constexpr int kProwLd[3] = {AttnBwdRow<32, 11, false>::LD, AttnBwdRow<64, 11, true>::LD, AttnBwdRow<128, 5, true>::LD};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct NetDesc {
  int nb;
  int btype[2];
  int n_heads;
};

bool describe(int kind, NetDesc* d) {
  switch (kind) {
    case DTA_NET_HANG2020: *d = {2, {BR_SPECTRAL, BR_SPATIAL}, 6}; return true;
    case DTA_NET_SPECTRAL: *d = {1, {BR_SPECTRAL, BR_NONE}, 3}; return true;
    case DTA_NET_SPATIAL: *d = {1, {BR_SPATIAL, BR_NONE}, 3}; return true;
    case DTA_NET_VANILLA: *d = {1, {BR_NONE, BR_NONE}, 1}; return true;
    case DTA_NET_SPECTRAL_PAIR: *d = {2, {BR_SPECTRAL, BR_SPECTRAL}, 6}; return true;
    case DTA_NET_SPATIAL_PAIR: *d = {2, {BR_SPATIAL, BR_SPATIAL}, 6}; return true;
    default: return false;
  }
}

// Carves float regions (256-byte aligned) out of one caller-owned buffer.
struct Carver {
  size_t off = 0;
  char* base;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t nfloats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += align_up(nfloats * sizeof(float), 256);
    return p;
  }
};

// Tensor-core path geometry.  Two position streams: side 11 (conv1, conv2 and their gradients) and
// side 5 (conv3).  Buffers are padded so that every kernel's tile size divides the stream length.
using WgradCfg1 = TcWgrad<128, 128, true, 3, 2>;  // conv1: merged 64 output channels (hi/lo stacked), 128-channel slices, 3 tap groups
using WgradCfg2 = TcWgrad<32, 128, true, 2, 2, true>;    // conv2: 64 output channels per branch, 32 input channels; hi|lo halves as one N = 64 operand, 2 tap groups
using WgradCfg3 = TcWgrad<32, 64, false, 2, 2, true>;    // conv3: 128 output channels per branch, 2 slices of 32 input channels; N = 64, 2 tap groups
constexpr int kTcSmCount = 148;              // split-K factors are sized for the B200's 148 SMs (dta_query_sizes has no device)

struct TcSplit { int nkstage, nslices, nsplit, per; };
TcSplit tc_split(size_t rows, int krows, int cin_g, int nci, int G, int ntg = 1) {
  TcSplit t{};
  t.nkstage = (int)((rows - 2 * kTcGuard) / krows);
  t.nslices = (cin_g + nci - 1) / nci;
  int want = kTcSmCount / (t.nslices * G * ntg);
  if (want < 1) want = 1;
  if (want > t.nkstage) want = t.nkstage;
  t.per = (t.nkstage + want - 1) / want;
  t.nsplit = (t.nkstage + t.per - 1) / t.per;
  return t;
}
struct TcGeom {
  size_t rows11, rows5;   // rows of the packed streams
  int nstage1;            // conv1 forward K stages (16 input channels each)
  int nchunk1;            // 8-channel chunks allocated for the packed crops (multiple of 6 and 2)
  TcSplit w1, w2, w3;     // weight-gradient decompositions
};
TcGeom tc_geom(int B, int bands, int nb) {
  TcGeom g{};
  g.rows11 = tc_rows(B, Stream<11>::PC, 1024);
  g.rows5 = tc_rows(B, Stream<5>::PC, 512);
  g.nstage1 = (bands + 15) / 16;
  g.nchunk1 = (2 * g.nstage1 + WgradCfg1::BCH - 1) / WgradCfg1::BCH * WgradCfg1::BCH;   // whole weight-gradient slices
  g.w1 = tc_split(g.rows11, WgradCfg1::KROWS, g.nchunk1 * 8, WgradCfg1::NCI, 1, WgradCfg1::NTG);
  g.w2 = tc_split(g.rows11, WgradCfg2::KROWS, 32, WgradCfg2::NCI, nb, WgradCfg2::NTG);
  g.w3 = tc_split(g.rows5, WgradCfg3::KROWS, 64, WgradCfg3::NCI, nb, WgradCfg3::NTG);
  return g;
}
inline __nv_bfloat16* take_bf16(Carver& c, size_t n16) { return reinterpret_cast<__nv_bfloat16*>(c.take(n16 * 4)); }

struct SavedLayout {
  __nv_bfloat16* xp;    // split-bf16 position stream of the crops (conv1 forward and weight gradient)
  __nv_bfloat16* a1p;   // ... of the gated block-1 activations (conv2 forward and weight gradient)
  __nv_bfloat16* a2p;   // ... of the gated, pooled block-2 activations (conv3)
  float* z[3];
  float* bn_mean[3]; float* bn_istd[3]; float* bn_scale[3]; float* bn_shift[3];
  float* att[3];
  float* feat[3];
  float* s3[2];
  float* wp[3];
  float* spec_pack[2][3][4];  // [branch][block][w0d, w0t, w1d, w1t]
  size_t bytes;
};

SavedLayout layout_saved(const dta_shape& s, const NetDesc& d, void* base) {
  SavedLayout L{};
  Carver c(base);
  const size_t B = s.batch;
  for (int k = 0; k < 3; ++k) L.z[k] = c.take(B * d.nb * kC[k] * kHWpre[k]);
  for (int k = 0; k < 3; ++k) {
    L.bn_mean[k] = c.take(d.nb * kC[k]); L.bn_istd[k] = c.take(d.nb * kC[k]);
    L.bn_scale[k] = c.take(d.nb * kC[k]); L.bn_shift[k] = c.take(d.nb * kC[k]);
  }
  for (int k = 0; k < 3; ++k) L.att[k] = c.take(B * d.nb * 3 * kAttRow[k]);
  for (int k = 0; k < 3; ++k) L.feat[k] = c.take(B * d.nb * kFeatLd[k]);
  for (int g = 0; g < 2; ++g) L.s3[g] = c.take(B * s.classes);
  L.wp[0] = c.take((size_t)s.bands * 9 * d.nb * 32);
  L.wp[1] = c.take((size_t)d.nb * 32 * 9 * 64);
  L.wp[2] = c.take((size_t)d.nb * 64 * 9 * 128);
  for (int g = 0; g < 2; ++g)
    for (int k = 0; k < 3; ++k)
      for (int q = 0; q < 4; ++q) L.spec_pack[g][k][q] = c.take((size_t)kC[k] * kC[k]);
  {
    const TcGeom g = tc_geom(s.batch, s.bands, d.nb);
    L.xp = take_bf16(c, (size_t)2 * g.nchunk1 * g.rows11);   // 16 B per (half, chunk, row)
    L.a1p = take_bf16(c, (size_t)2 * d.nb * 4 * g.rows11);
    L.a2p = take_bf16(c, (size_t)2 * d.nb * 8 * g.rows5);
  }
  L.bytes = c.off;
  return L;
}

struct FwdWork {
  float* stats;
  __nv_bfloat16* wpf[3];   // packed forward weights of conv1..3 for the tensor-core path
  size_t bytes;
};
FwdWork layout_fwd(const dta_shape& s, const NetDesc& d, void* base) {
  FwdWork W{};
  Carver c(base);
  {
    size_t n = (size_t)s.batch * d.nb * 128 * 2;                       // SIMT path: one row per CTA (<= batch)
    const size_t n_tc = (size_t)kTcSmCount * 4 * d.nb * 128 * 2;      // tensor-core path: four rows per persistent CTA
    W.stats = c.take(n > n_tc ? n : n_tc);
  }
  W.wpf[0] = take_bf16(c, (size_t)tc_geom(s.batch, s.bands, d.nb).nstage1 * (TcFprop<11, 64, false>::W_BYTES / 16));
  W.wpf[1] = take_bf16(c, (size_t)d.nb * 2 * (TcFprop<11, 64, false>::W_BYTES / 16));
  W.wpf[2] = take_bf16(c, (size_t)d.nb * 4 * (TcFprop<5, 128, true>::W_BYTES / 16));
  W.bytes = c.off;
  return W;
}

// wgrad split-K factors (batch slices per weight-gradient CTA column)
struct Splits { int n[3]; int per[3]; };
Splits wgrad_splits(int B) {
  Splits sp;
  const int want[3] = {24, 128, 64};
  for (int k = 0; k < 3; ++k) {
    int n = want[k] < B ? want[k] : B;
    int per = (B + n - 1) / n;
    n = (B + per - 1) / per;
    sp.n[k] = n; sp.per[k] = per;
  }
  return sp;
}

// Upper bound of the 32x32 tiles the batched small-gradient reduction needs (sizes the partial buffer).
int reduce_tiles_bound(int classes, int nb) {
  const int ct = (classes + 31) / 32;
  // per branch: heads (F <= 512 -> 16 tiles along j, three heads: 4 + 8 + 16), their biases, spectral CxC (1 + 4 + 16) x 2, ~30 column sums
  int per_branch = ct * (4 + 8 + 16) + 3 * ct + 2 * (1 + 4 + 16) + 40;
  const int vanilla = ct * 16 + ct;
  if (vanilla > per_branch) per_branch = vanilla;
  return nb * per_branch;
}

struct BwdWork {
  float* dS[6];
  float* da[3];
  float* dout[2];   // gradient wrt gated output of block 1, 2
  float* bnrows;
  float* k0[3]; float* k1[3]; float* k2[3];
  float* prow[3];   // per-crop partial parameter gradients of each attention block
  float* wpart;
  float* wd[3];
  __nv_bfloat16* dzp[3];   // split-bf16 position streams of each block's conv-output gradient (one per block: block k's weight
                           // gradient may still read its stream on the side stream while block k-1's is being packed)
  __nv_bfloat16* wdp[2];   // packed input-gradient weights of conv2, conv3 (tensor-core path)
  float* rpart;            // partial tiles of the batched small-gradient reduction
  size_t bytes;
};
BwdWork layout_bwd(const dta_shape& s, const NetDesc& d, void* base) {
  BwdWork W{};
  Carver c(base);
  const size_t B = s.batch;
  for (int h = 0; h < 6; ++h) W.dS[h] = c.take(B * s.classes);
  for (int k = 0; k < 3; ++k) W.da[k] = c.take(B * d.nb * kC[k] * kHWpre[k]);
  W.dout[0] = c.take(B * d.nb * 32 * 121);
  W.dout[1] = c.take(B * d.nb * 64 * 25);
  W.bnrows = c.take(B * d.nb * 256);
  for (int k = 0; k < 3; ++k) { W.k0[k] = c.take(d.nb * kC[k]); W.k1[k] = c.take(d.nb * kC[k]); W.k2[k] = c.take(d.nb * kC[k]); }
  for (int k = 0; k < 3; ++k) W.prow[k] = c.take(B * d.nb * 256);
  const Splits sp = wgrad_splits(s.batch);
  size_t wp = (size_t)sp.n[0] * d.nb * 32 * s.bands * 9;
  const size_t wp2 = (size_t)sp.n[1] * d.nb * 64 * 32 * 9, wp3 = (size_t)sp.n[2] * d.nb * 128 * 64 * 9;
  if (wp2 > wp) wp = wp2;
  if (wp3 > wp) wp = wp3;
  const TcGeom tg = tc_geom(s.batch, s.bands, d.nb);
  const size_t wp_tc[3] = {(size_t)tg.w1.nsplit * 64 * s.bands * 9, (size_t)tg.w2.nsplit * d.nb * 64 * 32 * 9,
                           (size_t)tg.w3.nsplit * d.nb * 128 * 64 * 9};
  for (int k = 0; k < 3; ++k)
    if (wp_tc[k] > wp) wp = wp_tc[k];
  W.wpart = c.take(wp);
  W.dzp[0] = take_bf16(c, (size_t)2 * 8 * tg.rows11);           // conv1: 64 merged channels
  W.dzp[1] = take_bf16(c, (size_t)2 * d.nb * 8 * tg.rows11);    // conv2: 64 per branch
  W.dzp[2] = take_bf16(c, (size_t)2 * d.nb * 16 * tg.rows5);    // conv3: 128 per branch
  W.wdp[0] = take_bf16(c, (size_t)d.nb * 4 * (TcFprop<11, 32, true>::W_BYTES / 16));
  W.wdp[1] = take_bf16(c, (size_t)d.nb * 8 * (TcFprop<5, 64, true>::W_BYTES / 16));
  W.rpart = c.take((size_t)reduce_tiles_bound(s.classes, d.nb) * kReduceSplits * 1024);
  W.wd[0] = c.take((size_t)d.nb * 32 * 9 * s.bands);
  W.wd[1] = c.take((size_t)d.nb * 64 * 9 * 32);
  W.wd[2] = c.take((size_t)d.nb * 128 * 9 * 64);
  W.bytes = c.off;
  return W;
}

inline bool tc_enabled_for_exchange(const dta_ctx* ctx) { return ctx->ex_mc != nullptr && ctx->ex_sync != nullptr; }

int check_shape(dta_ctx* ctx, const dta_shape* s, NetDesc* d) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!s) return fail(ctx, DTA_ERR_INVALID_ARG, "shape is NULL");
  if (!describe(s->net_kind, d)) return fail(ctx, DTA_ERR_INVALID_ARG, "unknown net_kind");
  if (s->batch <= 0 || s->bands <= 0 || s->classes <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "batch, bands and classes must be positive");
  if (s->classes > 4096) return fail(ctx, DTA_ERR_UNSUPPORTED, "classes > 4096 not supported");
  return DTA_OK;
}

// Opt a kernel in to `bytes` of dynamic shared memory -- once per (device, kernel, size), not on every launch.
template <typename K>
cudaError_t allow_smem(K kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  cudaGetDevice(&dev);
  const std::pair<int, const void*> key(dev, reinterpret_cast<const void*>(kernel));
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done[key] = bytes;
  return e;
}

// ---- convolution launchers (conv_impl 0) ----------------------------------------------
template <int S, int NCROP, int TP, int TC, int COUT, int CK>
cudaError_t launch_fprop(const ConvSrc& src, const float* wp, Ptr2 bias, int bias_split, float* out, int out_ctot,
                         float* stats, int B, int G, cudaStream_t st, int* nblk) {
  using Cfg = FpropCfg<S, NCROP, TP, TC, COUT, CK>;
  auto kern = conv3x3_fprop_simt<S, NCROP, TP, TC, COUT, CK>;
  cudaError_t e = allow_smem(kern, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  dim3 grid((B + NCROP - 1) / NCROP, G);
  if (nblk) *nblk = grid.x;
  launch_k(kern, grid, Cfg::NT, Cfg::SMEM_BYTES, st, src, wp, bias, bias_split, out, out_ctot, stats, B);
  return cudaGetLastError();
}

template <int S, int CIK, int COUT, int TCO>
cudaError_t launch_wgrad(const ConvSrc& in, const ConvSrc& dz, float* part, int B, int nsplit, int per, int G,
                         cudaStream_t st) {
  using Cfg = WgradCfg<S, CIK, COUT, TCO>;
  auto kern = conv3x3_wgrad_simt<S, CIK, COUT, TCO>;
  cudaError_t e = allow_smem(kern, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  dim3 grid((in.cin + CIK - 1) / CIK, nsplit, G);
  launch_k(kern, grid, Cfg::NT, Cfg::SMEM_BYTES, st, in, dz, part, B, per);
  return cudaGetLastError();
}

// ---- convolution launchers (conv_impl 1, tcgen05) ---------------------------------------
template <int S, int NCO, bool ACC2, bool FUSEX = false>
cudaError_t run_tc_fprop(dta_ctx* ctx, cudaStream_t st, const __nv_bfloat16* xp, size_t rows, int nchunk, int chunks_per_group,
                         const __nv_bfloat16* wp, int nstage, Ptr2 bias, int bias_split, float* out, int out_ctot, int cout_g, int B,
                         int G, float* stats = nullptr, int* nblk = nullptr, FuseX fx = FuseX{nullptr, 0, nullptr}) {
  using Cfg = TcFprop<S, NCO, ACC2>;
  auto kern = tc_conv_fprop_kernel<S, NCO, ACC2, FUSEX>;
  cudaError_t e = allow_smem(kern, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  const int ntiles = (int)((rows - 2 * kTcGuard) / Cfg::TILE);
  const int nwork = ntiles * G;
  const int sms = ctx->sm_count < kTcSmCount ? ctx->sm_count : kTcSmCount;   // the statistics workspace holds kTcSmCount * 4 rows
  const int grid = nwork < sms ? nwork : sms;
  if (nblk) *nblk = grid * 4;
  launch_k(kern, grid, kTcFpropThreads + (FUSEX ? 32 * kTcConvWarps : 0), Cfg::SMEM_BYTES, st, xp, rows, nchunk, chunks_per_group, wp, nstage, bias,
                                                                                         bias_split, out, out_ctot, cout_g, B, ntiles, G, stats, fx);
  return cudaGetLastError();
}
template <int S, class Cfg>
cudaError_t run_tc_wgrad(cudaStream_t st, const __nv_bfloat16* dzp, int dz_chunks, const __nv_bfloat16* xp, int x_chunks, size_t rows,
                         int cin_g, int cout_g, int G, const TcSplit& sp, float* part) {
  auto kern = tc_conv_wgrad_kernel<S, Cfg>;
  cudaError_t e = allow_smem(kern, Cfg::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  launch_k(kern, dim3(sp.nslices * Cfg::NTG, sp.nsplit, G), kTcThreads, Cfg::SMEM_BYTES, st, dzp, dz_chunks, xp, x_chunks, rows, cin_g, cout_g, sp.nkstage,
                                                                          sp.per, part);
  return cudaGetLastError();
}
template <int S>
cudaError_t run_tc_pack(dta_ctx* ctx, cudaStream_t st, const ConvSrc& src, int G, int B, int nchunk, size_t rows, __nv_bfloat16* dst) {
  const size_t total = rows * nchunk;
  size_t blocks = (total + 255) / 256;
  if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
  const int grid = (int)blocks;
  if (src.mode == SRC_RAW) launch_k(tc_pack_stream_kernel<S, SRC_RAW, false>, grid, 256, 0, st, src, G, B, nchunk, rows, dst);
  else if (src.mode == SRC_DZ && src.pool) launch_k(tc_pack_stream_kernel<S, SRC_DZ, true>, grid, 256, 0, st, src, G, B, nchunk, rows, dst);
  else if (src.mode == SRC_DZ) launch_k(tc_pack_stream_kernel<S, SRC_DZ, false>, grid, 256, 0, st, src, G, B, nchunk, rows, dst);
  else if (src.pool) launch_k(tc_pack_stream_kernel<S, SRC_ACT, true>, grid, 256, 0, st, src, G, B, nchunk, rows, dst);
  else launch_k(tc_pack_stream_kernel<S, SRC_ACT, false>, grid, 256, 0, st, src, G, B, nchunk, rows, dst);
  return cudaGetLastError();
}
ConvSrc src_raw(const float* x, int cin, int hw) {
  ConvSrc s{};
  s.mode = SRC_RAW; s.cin = cin; s.ctot = cin; s.src_hw = hw; s.a = x;
  return s;
}
ConvSrc src_act(const float* z, int cin, int nb, int src_hw, int pool, const float* scale, const float* shift,
                const float* gate, int gate_ld, int gate_off, const int* btype) {
  ConvSrc s{};
  s.mode = SRC_ACT; s.cin = cin; s.ctot = nb * cin; s.src_hw = src_hw; s.pool = pool; s.a = z;
  s.k0 = scale; s.k1 = shift; s.gate = gate; s.gate_ld = gate_ld; s.gate_off = gate_off;
  s.gate_mode[0] = btype[0]; s.gate_mode[1] = btype[1];
  return s;
}
// compact_cells > 0: `da` is the compact form the pooled attention backward writes -- B * ctot * compact_cells values followed
// by as many arg-max bytes (compact_cells = pooled cells per channel plane)
ConvSrc src_dz(const float* da, const float* z, int cin, int ctot, int hw, const float* k0, const float* k1,
               const float* k2, int B = 0, int compact_cells = 0) {
  ConvSrc s{};
  s.mode = SRC_DZ; s.cin = cin; s.ctot = ctot; s.src_hw = hw; s.a = da; s.b = z; s.k0 = k0; s.k1 = k1; s.k2 = k2;
  if (compact_cells > 0) {
    s.pool = 1;
    s.arg = reinterpret_cast<const unsigned char*>(da + (size_t)B * ctot * compact_cells);
  }
  return s;
}

AttnParams attn_params(const dta_tensors* p, const SavedLayout& L, const NetDesc& d, int k, bool vanilla_head, int classes_second = 0) {
  AttnParams a{};
  a.classes_g[1] = classes_second;
  for (int g = 0; g < 2; ++g) {
    a.btype[g] = d.btype[g];
    if (g >= d.nb) continue;
    const dta_branch& br = p->branch[g];
    if (d.btype[g] == BR_SPECTRAL) {
      a.w0d[g] = L.spec_pack[g][k][0]; a.w0t[g] = L.spec_pack[g][k][1];
      a.w1d[g] = L.spec_pack[g][k][2]; a.w1t[g] = L.spec_pack[g][k][3];
      a.b0[g] = br.attn[k].b0; a.b1[g] = br.attn[k].b1;
    } else if (d.btype[g] == BR_SPATIAL) {
      a.st0[g] = br.attn[k].w0; a.st1[g] = br.attn[k].w1;
      a.b0[g] = br.attn[k].b0; a.b1[g] = br.attn[k].b1;
      a.pool_w[g] = br.attn[k].pool_w; a.pool_b[g] = br.attn[k].pool_b;
    }
    if (d.btype[g] != BR_NONE || vanilla_head) { a.fc_w[g] = br.fc_w[k]; a.fc_b[g] = br.fc_b[k]; }
  }
  return a;
}

BnParams bn_params(const dta_tensors* p, const NetDesc& d, int k) {
  BnParams b{};
  for (int g = 0; g < d.nb; ++g) {
    const dta_conv_block& cb = p->branch[g].conv[k];
    b.gamma[g] = cb.bn_w; b.beta[g] = cb.bn_b; b.rm[g] = cb.bn_rm; b.rv[g] = cb.bn_rv;
    b.nbt[g] = reinterpret_cast<long long*>(cb.bn_nbt);
  }
  b.c_per_branch = kC[k];
  return b;
}

int validate_params(dta_ctx* ctx, const dta_tensors* p, const NetDesc& d, int kind, bool need_running) {
  if (!p) return fail(ctx, DTA_ERR_INVALID_ARG, "params is NULL");
  if (kind == DTA_NET_HANG2020 && !p->alpha) return fail(ctx, DTA_ERR_INVALID_ARG, "alpha is NULL");
  for (int g = 0; g < d.nb; ++g) {
    const dta_branch& br = p->branch[g];
    for (int k = 0; k < 3; ++k) {
      const dta_conv_block& cb = br.conv[k];
      if (!cb.conv_w || !cb.conv_b || !cb.bn_w || !cb.bn_b) return fail(ctx, DTA_ERR_INVALID_ARG, "conv block parameter is NULL");
      if (need_running && (!cb.bn_rm || !cb.bn_rv)) return fail(ctx, DTA_ERR_INVALID_ARG, "BatchNorm running statistics are NULL");
      const dta_attention& at = br.attn[k];
      if (d.btype[g] != BR_NONE && (!at.w0 || !at.b0 || !at.w1 || !at.b1)) return fail(ctx, DTA_ERR_INVALID_ARG, "attention parameter is NULL");
      if (d.btype[g] == BR_SPATIAL && (!at.pool_w || !at.pool_b)) return fail(ctx, DTA_ERR_INVALID_ARG, "channel_pool parameter is NULL");
      if ((d.btype[g] != BR_NONE || k == 2) && (!br.fc_w[k] || !br.fc_b[k])) return fail(ctx, DTA_ERR_INVALID_ARG, "classifier parameter is NULL");
    }
  }
  return DTA_OK;
}

}  // namespace

// =======================================================================================
extern "C" {

int dta_abi_version(void) { return DTA_ABI_VERSION; }

int dta_create(dta_ctx** out, int device) {
  if (!out) return DTA_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
    cudaGetLastError();
    return DTA_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { g_create_error = "device index out of range"; return DTA_ERR_INVALID_ARG; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return DTA_ERR_CUDA; }
  if (prop.major != 10) {
    g_create_error = "device is sm_" + std::to_string(prop.major * 10 + prop.minor) + "; this library is built for sm_100a only";
    return DTA_ERR_NO_DEVICE;
  }
  dta_ctx* c = new dta_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  // side stream (highest priority: its one-CTA-per-SM tensor kernels must get their shared memory before a co-scheduled
  // attention grid fills the SMs) and the events that fork / join it
  {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi) != cudaSuccess) c->side = nullptr;
    if (cudaStreamCreateWithPriority(&c->aux, cudaStreamNonBlocking, hi) != cudaSuccess) c->aux = nullptr;
    for (int i = 0; i < 32 && c->side; ++i) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess) c->sync_events.push_back(e);
    }
    // "last CTA done" ticket counters (self-resetting, so they are zeroed exactly once): the only device memory the
    // library owns; every per-call buffer stays the caller's
    if (cudaMalloc(&c->tickets, 64 * sizeof(unsigned int)) == cudaSuccess) cudaMemset(c->tickets, 0, 64 * sizeof(unsigned int));
    else c->tickets = nullptr;
    cudaGetLastError();
    cudaSetDevice(prev);
  }
  *out = c;
  return DTA_OK;
}

void dta_destroy(dta_ctx* ctx) {
  if (!ctx) return;
  fold_spans(ctx);
  for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->sync_events) cudaEventDestroy(e);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->tickets) cudaFree(ctx->tickets);
  delete ctx;
}

int dta_profile_read(dta_ctx* ctx, dta_stage_time* out, int capacity, int* count, int reset) {
  if (!ctx || !count) return DTA_ERR_INVALID_ARG;
  fold_spans(ctx);
  const int n = (int)ctx->stage_names.size();
  *count = n;
  for (int i = 0; i < n && i < capacity && out; ++i) {
    memset(out[i].name, 0, sizeof(out[i].name));
    strncpy(out[i].name, ctx->stage_names[i].c_str(), sizeof(out[i].name) - 1);
    out[i].total_ms = ctx->stage_ms[i];
    out[i].calls = ctx->stage_calls[i];
  }
  if (reset) {
    for (int i = 0; i < n; ++i) { ctx->stage_ms[i] = 0.0; ctx->stage_calls[i] = 0; }
  }
  return DTA_OK;
}

const char* dta_last_error(const dta_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dta_set_option(dta_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return DTA_ERR_INVALID_ARG;
  if (!strcmp(key, "conv_impl")) {
    if (value != 0 && value != 1) return fail(ctx, DTA_ERR_INVALID_ARG, "conv_impl must be 0 (fp32 direct) or 1 (tcgen05 split-bf16)");
    ctx->conv_impl = (int)value;
    return DTA_OK;
  }
  if (!strcmp(key, "profile")) {
    ctx->profile = value != 0;
    return DTA_OK;
  }
  if (!strcmp(key, "small_tiles")) {
    if (value < 0 || value > 2) return fail(ctx, DTA_ERR_INVALID_ARG, "small_tiles: 0 = never, 1 = eval mode (default), 2 = training too");
    ctx->small_tiles = (int)value;
    return DTA_OK;
  }
  if (!strcmp(key, "fuse_x")) {
    ctx->fuse_x = value != 0;
    return DTA_OK;
  }
  if (!strcmp(key, "overlap")) {
    if (value < 0 || value > 2) return fail(ctx, DTA_ERR_INVALID_ARG, "overlap must be 0, 1 or 2");
    ctx->overlap = (int)value;
    return DTA_OK;
  }
  if (!strcmp(key, "pdl")) {
    ctx->pdl = value != 0;
    return DTA_OK;
  }
  return fail(ctx, DTA_ERR_INVALID_ARG, std::string("unknown option ") + key);
}

int dta_set_update_gate(dta_ctx* ctx, const float* gate) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  ctx->update_gate = gate;
  return DTA_OK;
}

int dta_get_option(const dta_ctx* ctx, const char* key, int64_t* value) {
  if (!ctx || !key || !value) return DTA_ERR_INVALID_ARG;
  if (!strcmp(key, "conv_impl")) { *value = ctx->conv_impl; return DTA_OK; }
  if (!strcmp(key, "launches")) { *value = ctx->launches; return DTA_OK; }
  if (!strcmp(key, "launches_total")) { *value = ctx->launches_total + ctx->launches; return DTA_OK; }
  if (!strcmp(key, "profile")) { *value = ctx->profile; return DTA_OK; }
  if (!strcmp(key, "sm_count")) { *value = ctx->sm_count; return DTA_OK; }
  if (!strcmp(key, "overlap")) { *value = ctx->overlap; return DTA_OK; }
  if (!strcmp(key, "pdl")) { *value = ctx->pdl; return DTA_OK; }
  if (!strcmp(key, "fuse_x")) { *value = ctx->fuse_x; return DTA_OK; }
  if (!strcmp(key, "small_tiles")) { *value = ctx->small_tiles; return DTA_OK; }
  if (!strcmp(key, "exchanged")) { *value = ctx->exchanged; return DTA_OK; }
  return DTA_ERR_INVALID_ARG;
}

int dta_query_sizes(const dta_shape* shape, dta_sizes* out) {
  NetDesc d;
  if (!shape || !out || !describe(shape->net_kind, &d)) return DTA_ERR_INVALID_ARG;
  if (shape->batch <= 0 || shape->bands <= 0 || shape->classes <= 0) return DTA_ERR_INVALID_ARG;
  out->saved_bytes = layout_saved(*shape, d, nullptr).bytes;
  out->workspace_fwd = layout_fwd(*shape, d, nullptr).bytes;
  out->workspace_bwd = layout_bwd(*shape, d, nullptr).bytes;
  out->n_heads = d.n_heads;
  return DTA_OK;
}

int dta_saved_region(const dta_shape* shape, int block, int region, size_t* offset_bytes, size_t* n_floats) {
  NetDesc d;
  if (!shape || !offset_bytes || !n_floats || !describe(shape->net_kind, &d)) return DTA_ERR_INVALID_ARG;
  if (shape->batch <= 0 || shape->bands <= 0 || shape->classes <= 0 || block < 0 || block > 2 || region < 0 || region > 2) return DTA_ERR_INVALID_ARG;
  char* const base = reinterpret_cast<char*>(uintptr_t(256));   // any non-null base: only differences are used
  const SavedLayout L = layout_saved(*shape, d, base);
  const float* p = region == 0 ? L.z[block] : (region == 1 ? L.bn_scale[block] : L.bn_shift[block]);
  *offset_bytes = (size_t)(reinterpret_cast<const char*>(p) - base);
  *n_floats = region == 0 ? (size_t)shape->batch * d.nb * kC[block] * kHWpre[block] : (size_t)d.nb * kC[block];
  return DTA_OK;
}

// Shared body of dta_forward and dta_forward_pair (classes_second > 0: branch 1's heads have that many classes).
static int forward_impl(dta_ctx* ctx, const dta_shape* shape, int classes_second, const float* x, const dta_tensors* params,
                        float* const scores[6], float* joint, void* saved, void* workspace, void* cuda_stream, bool joint_optional = false) {
  NetDesc d;
  int rc = check_shape(ctx, shape, &d);
  if (rc != DTA_OK) return rc;
  if (!x || !scores || !saved || !workspace) return fail(ctx, DTA_ERR_INVALID_ARG, "x, scores, saved and workspace are required");
  if ((reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(workspace)) & 255u)
    return fail(ctx, DTA_ERR_INVALID_ARG, "saved and workspace must be 256-byte aligned (bulk copies and vector stores rely on it)");
  rc = validate_params(ctx, params, d, shape->net_kind, true);
  if (rc != DTA_OK) return rc;
  for (int h = 0; h < d.n_heads; ++h)
    if (!scores[h]) return fail(ctx, DTA_ERR_INVALID_ARG, "scores[h] is NULL for an existing head");
  if (shape->net_kind == DTA_NET_HANG2020 && !joint && !joint_optional) return fail(ctx, DTA_ERR_INVALID_ARG, "joint is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  pdl_enabled() = ctx->pdl;

  const int B = shape->batch, nb = d.nb, bands = shape->bands, classes = shape->classes;
  const bool vanilla = shape->net_kind == DTA_NET_VANILLA;
  SavedLayout L = layout_saved(*shape, d, saved);
  FwdWork W = layout_fwd(*shape, d, workspace);

  // 1. pack parameters into kernel-friendly tables (a few MB, once per step).  Everything that depends on the parameters
  //    only (attention tables, conv2/conv3 operand tables, guard rows of the activation streams) goes to the side stream and
  //    runs under conv1; it is joined before the first consumer (block 1's attention kernel).
  SideStream side(ctx, st);
  side.wait_main();
  cudaStream_t ss = side.stream();
  const bool tcp = ctx->conv_impl == 1;
  const TcGeom tg = tc_geom(shape->batch, shape->bands, d.nb);
  StageScope pack_scope(ctx, "fwd.pack_params", st);
  for (int k = 0; k < 3 && ctx->conv_impl == 0; ++k) {
    Ptr2 w{{params->branch[0].conv[k].conv_w, nb > 1 ? params->branch[1].conv[k].conv_w : nullptr}};
    const int cin = k == 0 ? bands : kC[k - 1];
    launch_k(pack_conv_w_kernel, ctx->sm_count * 2, 256, 0, st, w, nb, kC[k], cin, k == 0 ? 1 : 0, L.wp[k]);
    DTA_CHECK_LAUNCH(ctx, "pack_conv_w");
  }
  for (int g = 0; g < nb; ++g) {
    if (d.btype[g] != BR_SPECTRAL) continue;
    for (int k = 0; k < 3; ++k) {
      const int ks = k == 0 ? 3 : (k == 1 ? 5 : 7);  // Hang2020.py:136-141
      launch_k(pack_spectral_kernel, (kC[k] * kC[k] + 255) / 256, 256, 0, ss, params->branch[g].attn[k].w0, kC[k], ks, L.spec_pack[g][k][0], L.spec_pack[g][k][1]);
      DTA_CHECK_LAUNCH(ctx, "pack_spectral");
      launch_k(pack_spectral_kernel, (kC[k] * kC[k] + 255) / 256, 256, 0, ss, params->branch[g].attn[k].w1, kC[k], ks, L.spec_pack[g][k][2], L.spec_pack[g][k][3]);
      DTA_CHECK_LAUNCH(ctx, "pack_spectral");
    }
  }
  auto conv_w = [&](int k) { return Ptr2{{params->branch[0].conv[k].conv_w, nb > 1 ? params->branch[1].conv[k].conv_w : nullptr}}; };
  auto conv_b = [&](int k) { return Ptr2{{params->branch[0].conv[k].conv_b, nb > 1 ? params->branch[1].conv[k].conv_b : nullptr}}; };
  if (tcp) {
    launch_k(tc_pack_w_fprop_kernel<64>, ctx->sm_count, 256, 0, ss, conv_w(1), nb, 64, 32, 2, 1, W.wpf[1]);
    DTA_CHECK_LAUNCH(ctx, "tc_pack_w_fprop");
    launch_k(tc_pack_w_fprop_kernel<128>, ctx->sm_count, 256, 0, ss, conv_w(2), nb, 128, 64, 4, 1, W.wpf[2]);
    DTA_CHECK_LAUNCH(ctx, "tc_pack_w_fprop");
    if (!vanilla) {   // the attention kernels write the crop rows of a1p / a2p; the guard and tail rows are zeroed here
      launch_k(tc_zero_guards_kernel, 32, 256, 0, ss, L.a1p, tg.rows11, nb * 4, (size_t)B * Stream<11>::PC);
      DTA_CHECK_LAUNCH(ctx, "tc_zero_guards");
      launch_k(tc_zero_guards_kernel, 32, 256, 0, ss, L.a2p, tg.rows5, nb * 8, (size_t)B * Stream<5>::PC);
      DTA_CHECK_LAUNCH(ctx, "tc_zero_guards");
    }
  }

  pack_scope.end();
  cudaError_t e;
  int nblk = 0;
  // 2. block 1: conv1 over the crops (both branches share the read of x)
  if (tcp) {
    // tensor-core path: pack crops + weights, implicit GEMM, batch statistics of z
    {
      StageScope sc(ctx, "fwd.conv1_pack", st);
      if (!ctx->fuse_x) DTA_TC_CHECK(run_tc_pack<11>(ctx, st, src_raw(x, bands, kHW), 1, B, tg.nchunk1, tg.rows11, L.xp), "tc_pack_stream(x)");
      // on the critical path in front of conv1: one thread per weight element, the whole table in flight at once
      launch_k(tc_pack_w_fprop_kernel<64>, (tg.nstage1 * 9 * 2 * 64 * 8 + 255) / 256, 256, 0, st, conv_w(0), nb, 32, bands, tg.nstage1, 0, W.wpf[0]);
      DTA_CHECK_LAUNCH(ctx, "tc_pack_w_fprop");
    }
    {
      StageScope sc(ctx, "fwd.conv1", st);
      if (ctx->fuse_x) {
        // the crops are converted inside the convolution; the packed copy it leaves in L.xp feeds the weight gradient
        if (2 * tg.nstage1 < tg.nchunk1) {   // chunks beyond the forward's K range (weight-gradient slice padding) stay zero
          const size_t used = (size_t)2 * tg.nstage1 * tg.rows11 * 16, all = (size_t)tg.nchunk1 * tg.rows11 * 16;
          cudaMemsetAsync(reinterpret_cast<char*>(L.xp) + used, 0, all - used, st);
          cudaMemsetAsync(reinterpret_cast<char*>(L.xp) + all + used, 0, all - used, st);
        }
        // 512-position tiles (one accumulator of all 512 tensor-memory columns: the weights stream half as often) while they
        // fill the SMs; at small batches 256-position tiles (two accumulator stages) keep twice as many SMs busy
        const bool big_tiles = (long long)(tg.rows11 - 2 * kTcGuard) / TcFprop<11, 64, false>::TILE >= (ctx->sm_count < kTcSmCount ? ctx->sm_count : kTcSmCount);
        // (training keeps the 512-position tiles unless small_tiles = 2: the BatchNorm partial sums are grouped per CTA, so the
        // tile size moves the batch statistics in their last bits; in eval mode the two forms are bit-identical)
        const bool small = !big_tiles && (ctx->small_tiles == 2 || (ctx->small_tiles == 1 && !shape->training));
        if (!small)
          DTA_TC_CHECK((run_tc_fprop<11, 64, false, true>(ctx, st, L.xp, tg.rows11, tg.nchunk1, 0, W.wpf[0], tg.nstage1, conv_b(0), 32, L.z[0], nb * 32,
                                                         nb * 32, B, 1, shape->training ? W.stats : nullptr, &nblk, FuseX{x, bands, L.xp})),
                       "tc_conv_fprop(conv1, fused crops)");
        else
          DTA_TC_CHECK((run_tc_fprop<11, 64, true, true>(ctx, st, L.xp, tg.rows11, tg.nchunk1, 0, W.wpf[0], tg.nstage1, conv_b(0), 32, L.z[0], nb * 32,
                                                        nb * 32, B, 1, shape->training ? W.stats : nullptr, &nblk, FuseX{x, bands, L.xp})),
                       "tc_conv_fprop(conv1, fused crops, 256-position tiles)");
      } else {
        DTA_TC_CHECK((run_tc_fprop<11, 64, false>(ctx, st, L.xp, tg.rows11, tg.nchunk1, 0, W.wpf[0], tg.nstage1, conv_b(0), 32, L.z[0], nb * 32, nb * 32, B, 1,
                                                 shape->training ? W.stats : nullptr, &nblk)),
                     "tc_conv_fprop(conv1)");
      }
    }
  } else {
    StageScope sc(ctx, "fwd.conv1", st);
    ConvSrc src = src_raw(x, bands, kHW);
    Ptr2 bias{{params->branch[0].conv[0].conv_b, nb > 1 ? params->branch[1].conv[0].conv_b : nullptr}};
    float* stats = shape->training ? W.stats : nullptr;
    if (nb == 2) e = launch_fprop<11, 1, 8, 8, 64, 16>(src, L.wp[0], bias, 32, L.z[0], 64, stats, B, 1, st, &nblk);
    else e = launch_fprop<11, 1, 8, 8, 32, 16>(src, L.wp[0], bias, 32, L.z[0], 32, stats, B, 1, st, &nblk);
    if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv1 fprop: ") + cudaGetErrorString(e));
    ctx->launches++;
  }
  auto bn_finalize = [&](int k, int nblk_k) -> int {
    StageScope sc(ctx, "fwd.bn_finalize", st);
    const int ctot = nb * kC[k];
    launch_k(bn_fwd_finalize_kernel, (ctot + kBnCh - 1) / kBnCh, kBnCh * kBnSlices, 0, st, W.stats, nblk_k, ctot, (double)B * kHWpre[k], bn_params(params, d, k),
                                                                    shape->training, L.bn_mean[k], L.bn_istd[k], L.bn_scale[k], L.bn_shift[k], ctx->update_gate);
    DTA_CHECK_LAUNCH(ctx, "bn_fwd_finalize");
    return DTA_OK;
  };
  auto score_ptrs = [&](int k) {
    MutPtr2 m{{nullptr, nullptr}};
    if (vanilla) { if (k == 2) m.p[0] = scores[0]; return m; }
    for (int g = 0; g < nb; ++g) m.p[g] = scores[g * 3 + k];
    return m;
  };
  if ((rc = bn_finalize(0, nblk)) != DTA_OK) return rc;
  side.join();   // parameter tables are ready from here on
  if (!vanilla) {
    StageScope sc(ctx, "fwd.attn1", st);
    auto kern = attn_fwd_kernel<32, 11, false>;
    const size_t sm = attn_fwd_smem<32, 11, false>();
    allow_smem(kern, sm);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[0], L.bn_scale[0], L.bn_shift[0], attn_params(params, L, d, 0, false, classes_second), classes, L.att[0], L.feat[0], score_ptrs(0),
                                              tcp ? L.a1p : nullptr, tg.rows11, nb * 4);
    DTA_CHECK_LAUNCH(ctx, "attn_fwd<1>");
  }
  // 3. block 2
  if (tcp) {
    ConvSrc src = src_act(L.z[0], 32, nb, 121, 0, L.bn_scale[0], L.bn_shift[0], L.att[0], 3 * kAttRow[0], 2 * kAttRow[0], d.btype);
    {
      StageScope sc(ctx, "fwd.conv2_pack", st);
      // block 1's attention kernel already wrote the crop rows of a1p (vanilla_CNN has no attention kernel: pack here)
      if (vanilla) DTA_TC_CHECK(run_tc_pack<11>(ctx, st, src, nb, B, nb * 4, tg.rows11, L.a1p), "tc_pack_stream(act1)");
    }
    {
      StageScope sc(ctx, "fwd.conv2", st);
      DTA_TC_CHECK((run_tc_fprop<11, 64, true>(ctx, st, L.a1p, tg.rows11, nb * 4, 4, W.wpf[1], 2, conv_b(1), 64, L.z[1], nb * 64, 64, B, nb,
                                              shape->training ? W.stats : nullptr, &nblk)), "tc_conv_fprop(conv2)");
    }
  } else {
    StageScope sc(ctx, "fwd.conv2", st);
    ConvSrc src = src_act(L.z[0], 32, nb, 121, 0, L.bn_scale[0], L.bn_shift[0], L.att[0], 3 * kAttRow[0], 2 * kAttRow[0], d.btype);
    Ptr2 bias{{params->branch[0].conv[1].conv_b, nb > 1 ? params->branch[1].conv[1].conv_b : nullptr}};
    e = launch_fprop<11, 1, 8, 8, 64, 16>(src, L.wp[1], bias, 64, L.z[1], nb * 64, shape->training ? W.stats : nullptr, B, nb, st, &nblk);
    if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv2 fprop: ") + cudaGetErrorString(e));
    ctx->launches++;
  }
  if ((rc = bn_finalize(1, nblk)) != DTA_OK) return rc;
  if (!vanilla) {
    StageScope sc(ctx, "fwd.attn2", st);
    auto kern = attn_fwd_kernel<64, 11, true>;
    const size_t sm = attn_fwd_smem<64, 11, true>();
    allow_smem(kern, sm);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[1], L.bn_scale[1], L.bn_shift[1], attn_params(params, L, d, 1, false, classes_second), classes, L.att[1], L.feat[1], score_ptrs(1),
                                              tcp ? L.a2p : nullptr, tg.rows5, nb * 8);
    DTA_CHECK_LAUNCH(ctx, "attn_fwd<2>");
  }
  // 4. block 3
  if (tcp) {
    ConvSrc src = src_act(L.z[1], 64, nb, 121, 1, L.bn_scale[1], L.bn_shift[1], L.att[1], 3 * kAttRow[1], 2 * kAttRow[1], d.btype);
    {
      StageScope sc(ctx, "fwd.conv3_pack", st);
      if (vanilla) DTA_TC_CHECK(run_tc_pack<5>(ctx, st, src, nb, B, nb * 8, tg.rows5, L.a2p), "tc_pack_stream(act2)");
    }
    {
      StageScope sc(ctx, "fwd.conv3", st);
      DTA_TC_CHECK((run_tc_fprop<5, 128, true>(ctx, st, L.a2p, tg.rows5, nb * 8, 8, W.wpf[2], 4, conv_b(2), 128, L.z[2], nb * 128, 128, B, nb,
                                              shape->training ? W.stats : nullptr, &nblk)), "tc_conv_fprop(conv3)");
    }
  } else {
    StageScope sc(ctx, "fwd.conv3", st);
    ConvSrc src = src_act(L.z[1], 64, nb, 121, 1, L.bn_scale[1], L.bn_shift[1], L.att[1], 3 * kAttRow[1], 2 * kAttRow[1], d.btype);
    Ptr2 bias{{params->branch[0].conv[2].conv_b, nb > 1 ? params->branch[1].conv[2].conv_b : nullptr}};
    e = launch_fprop<5, 4, 4, 8, 128, 8>(src, L.wp[2], bias, 128, L.z[2], nb * 128, shape->training ? W.stats : nullptr, B, nb, st, &nblk);
    if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv3 fprop: ") + cudaGetErrorString(e));
    ctx->launches++;
  }
  if ((rc = bn_finalize(2, nblk)) != DTA_OK) return rc;
  {
    StageScope sc(ctx, "fwd.attn3", st);
    auto kern = attn_fwd_kernel<128, 5, true>;
    const size_t sm = attn_fwd_smem<128, 5, true>();
    allow_smem(kern, sm);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[2], L.bn_scale[2], L.bn_shift[2], attn_params(params, L, d, 2, vanilla, classes_second), classes, L.att[2], L.feat[2], score_ptrs(2), nullptr, 0, 0);
    DTA_CHECK_LAUNCH(ctx, "attn_fwd<3>");
  }
  // 5. alpha blend + copies of the last-head scores for dalpha
  if (shape->net_kind == DTA_NET_HANG2020 && joint != nullptr) {
    StageScope sc(ctx, "fwd.joint", st);
    const size_t n = (size_t)B * classes;
    launch_k(joint_fwd_kernel, (int)((n + 255) / 256), 256, 0, st, scores[2], scores[5], params->alpha, joint, L.s3[0], L.s3[1], n);
    DTA_CHECK_LAUNCH(ctx, "joint_fwd");
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("forward: ") + cudaGetErrorString(e));
  return DTA_OK;
}

int dta_forward(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                float* const scores[6], float* joint, void* saved, void* workspace, void* cuda_stream) {
  if (ctx && shape && (shape->net_kind == DTA_NET_SPECTRAL_PAIR || shape->net_kind == DTA_NET_SPATIAL_PAIR))
    return fail(ctx, DTA_ERR_INVALID_ARG, "pair kinds are inference fan-out only: use dta_forward_pair");
  return forward_impl(ctx, shape, 0, x, params, scores, joint, saved, workspace, cuda_stream);
}

int dta_forward_pair(dta_ctx* ctx, const dta_shape* shape, int classes_second, const float* x, const dta_tensors* params,
                     float* const scores[6], void* saved, void* workspace, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!shape || (shape->net_kind != DTA_NET_SPECTRAL_PAIR && shape->net_kind != DTA_NET_SPATIAL_PAIR))
    return fail(ctx, DTA_ERR_INVALID_ARG, "dta_forward_pair needs net_kind DTA_NET_SPECTRAL_PAIR or DTA_NET_SPATIAL_PAIR");
  if (shape->training) return fail(ctx, DTA_ERR_UNSUPPORTED, "dta_forward_pair is eval-mode only (prediction fan-out)");
  if (classes_second <= 0 || classes_second > 4096) return fail(ctx, DTA_ERR_INVALID_ARG, "classes_second must be in [1, 4096]");
  return forward_impl(ctx, shape, classes_second, x, params, scores, nullptr, saved, workspace, cuda_stream);
}

int dta_preprocess_crops(dta_ctx* ctx, const int16_t* raw, int batch, int bands_in, int clip, float* out, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!raw || !out) return fail(ctx, DTA_ERR_INVALID_ARG, "raw and out are required");
  if (batch <= 0 || bands_in <= 0 || clip < 0 || bands_in - 2 * clip <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "batch and bands_in - 2*clip must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  pdl_enabled() = ctx->pdl;
  StageScope sc(ctx, "data.preprocess_crops", st);
  launch_k(preprocess_crops_kernel, batch, 128 * kPrepSlices, 0, st, reinterpret_cast<const short*>(raw), bands_in, kHW, clip, out);
  DTA_CHECK_LAUNCH(ctx, "preprocess_crops");
  return DTA_OK;
}

int dta_grad_allreduce_sizes(size_t n_float4, size_t n_double, int world, size_t* buffer_bytes, size_t* flags_offset, size_t* scratch_bytes) {
  if (world < 1 || world > kMaxPeers || !buffer_bytes || !scratch_bytes) return DTA_ERR_INVALID_ARG;
  *buffer_bytes = ar_buffer_bytes(n_float4, n_double, world);
  if (flags_offset) *flags_offset = ar_flags_offset(n_float4, n_double);
  *scratch_bytes = n_float4 * 16 + n_double * 8 + 16;
  return DTA_OK;
}

static int launch_allreduce(dta_ctx* ctx, cudaStream_t st, int rank, int world, void* const peer_buffers[], const void* mc, size_t n4, size_t nd,
                            void* scratch, void* sync_words, const ArRanges& rg, int flag_set) {
  PeerPtrs pp{};
  for (int r = 0; r < world; ++r) pp.buf[r] = static_cast<float*>(peer_buffers[r]);
  // every CTA spins on flags, so the grid must be co-resident: far below one CTA per SM
  size_t work = n4;
  if (mc != nullptr) {     // NVLS path: a rank only touches its own slice
    if (rg.n > 0) { work = 0; for (int k = 0; k < rg.n; ++k) work += rg.hi[k] - rg.lo[k]; }
    work = (work + world - 1) / world;
  }
  int grid = (int)((work + kArThreads - 1) / kArThreads);
  if (grid > 64) grid = 64;
  if (grid < 1) grid = 1;
  grad_allreduce_kernel<<<grid, kArThreads, 0, st>>>(pp, static_cast<const float*>(mc), rank, world, n4, nd, static_cast<float*>(scratch),
                                                     static_cast<uint32_t*>(sync_words), rg, flag_set);
  DTA_CHECK_LAUNCH(ctx, "grad_allreduce");
  return DTA_OK;
}

int dta_grad_allreduce(dta_ctx* ctx, int rank, int world, void* const peer_buffers[], const void* multicast_buffer, size_t n_float4,
                       size_t n_double, void* scratch, void* sync_words, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return fail(ctx, DTA_ERR_INVALID_ARG, "need 0 <= rank < world <= 16");
  if (!peer_buffers || !scratch || !sync_words) return fail(ctx, DTA_ERR_INVALID_ARG, "peer_buffers, scratch and sync_words are required");
  for (int r = 0; r < world; ++r)
    if (!peer_buffers[r]) return fail(ctx, DTA_ERR_INVALID_ARG, "peer_buffers[r] is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  pdl_enabled() = ctx->pdl;
  StageScope sc(ctx, "dist.grad_allreduce", st);
  return launch_allreduce(ctx, st, rank, world, peer_buffers, multicast_buffer, n_float4, n_double, scratch, sync_words, ArRanges{}, 0);
}

int dta_set_grad_exchange(dta_ctx* ctx, int rank, int world, void* const peer_buffers[], const void* multicast_buffer, size_t n_float4,
                          size_t n_double, void* sync_words) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  ctx->ex_world = 0;
  ctx->exchanged = 0;
  if (world <= 1 || !peer_buffers) return DTA_OK;
  if (world > kMaxPeers || rank < 0 || rank >= world) return fail(ctx, DTA_ERR_INVALID_ARG, "need 0 <= rank < world <= 16");
  if (!multicast_buffer || !sync_words) return fail(ctx, DTA_ERR_INVALID_ARG, "the in-backward exchange needs a multicast mapping and sync_words");
  for (int r = 0; r < world; ++r) {
    if (!peer_buffers[r]) return fail(ctx, DTA_ERR_INVALID_ARG, "peer_buffers[r] is NULL");
    ctx->ex_peers[r] = peer_buffers[r];
  }
  ctx->ex_rank = rank; ctx->ex_world = world; ctx->ex_mc = multicast_buffer; ctx->ex_n4 = n_float4; ctx->ex_nd = n_double; ctx->ex_sync = sync_words;
  return DTA_OK;
}

int dta_loss_workspace_bytes(int batch, int n_heads, size_t* out) {
  if (!out || batch <= 0 || n_heads <= 0 || n_heads > 7) return DTA_ERR_INVALID_ARG;
  *out = 256 + (size_t)n_heads * batch * sizeof(float);
  return DTA_OK;
}

int dta_cross_entropy_heads(dta_ctx* ctx, int batch, int classes, int n_heads, const float* const scores[], const int64_t* labels,
                            const float* class_weight, float* loss, float* const dscores[], void* workspace, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (batch <= 0 || classes <= 0 || n_heads <= 0 || n_heads > 7) return fail(ctx, DTA_ERR_INVALID_ARG, "batch, classes must be positive and 1 <= n_heads <= 7");
  if (!scores || !labels || !loss || !workspace) return fail(ctx, DTA_ERR_INVALID_ARG, "scores, labels, loss and workspace are required");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  pdl_enabled() = ctx->pdl;
  HeadPtrs h{};
  for (int i = 0; i < n_heads; ++i) {
    if (!scores[i]) return fail(ctx, DTA_ERR_INVALID_ARG, "scores[i] is NULL");
    h.s[i] = scores[i];
    h.ds[i] = dscores ? dscores[i] : nullptr;
  }
  if (!ctx->tickets) return fail(ctx, DTA_ERR_CUDA, "context has no ticket counters (allocation failed at dta_create)");
  float* rows = reinterpret_cast<float*>(static_cast<char*>(workspace) + 256);
  StageScope sc(ctx, "loss.cross_entropy", st);
  const long long warps = (long long)n_heads * batch;
  launch_k(ce_rows_kernel, (unsigned)((warps * 32 + 255) / 256), 256, 0, st, h, n_heads, reinterpret_cast<const long long*>(labels), class_weight, batch,
           classes, rows, loss, ctx->tickets + 0);
  DTA_CHECK_LAUNCH(ctx, "ce_rows");
  return DTA_OK;
}

// ce3: fused training step -- the score gradients of block 3's heads are formed inside its attention-backward kernel from
// the scores (the loss kernel that writes all of dscores runs on the side stream and is joined before block 2 reads them).
static int backward_impl(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                         const void* saved, const float* const dscores[6], const float* djoint,
                         const dta_tensors* grads, float* dx, void* workspace, void* cuda_stream, const CeInline* ce3) {
  NetDesc d;
  int rc = check_shape(ctx, shape, &d);
  if (rc != DTA_OK) return rc;
  if (shape->net_kind == DTA_NET_SPECTRAL_PAIR || shape->net_kind == DTA_NET_SPATIAL_PAIR)
    return fail(ctx, DTA_ERR_UNSUPPORTED, "pair kinds are inference fan-out only: no backward");
  if (!x || !saved || !workspace || !grads || !dscores) return fail(ctx, DTA_ERR_INVALID_ARG, "x, saved, workspace, dscores and grads are required");
  if ((reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(workspace)) & 255u)
    return fail(ctx, DTA_ERR_INVALID_ARG, "saved and workspace must be 256-byte aligned (bulk copies and vector stores rely on it)");
  rc = validate_params(ctx, params, d, shape->net_kind, false);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  pdl_enabled() = ctx->pdl;

  const int B = shape->batch, nb = d.nb, bands = shape->bands, classes = shape->classes;
  const bool vanilla = shape->net_kind == DTA_NET_VANILLA;
  const bool hang = shape->net_kind == DTA_NET_HANG2020;
  SavedLayout L = layout_saved(*shape, d, const_cast<void*>(saved));
  BwdWork W = layout_bwd(*shape, d, workspace);
  const Splits sp = wgrad_splits(B);
  const size_t nsc = (size_t)B * classes;
  cudaError_t e;
  // In-backward gradient exchange (dta_set_grad_exchange): only when the gradient table IS the registered symmetric buffer
  // (first float32 gradient at its base, conv1 weight gradients inside it) -- otherwise the caller reduces afterwards.
  ctx->exchanged = 0;
  bool exchange = false;
  unsigned long long ex_lo[2] = {0, 0}, ex_hi[2] = {0, 0};
  int ex_nw = 0;
  if (ctx->ex_world > 1 && tc_enabled_for_exchange(ctx) && grads->branch[0].conv[0].conv_w == ctx->ex_peers[ctx->ex_rank]) {
    exchange = true;
    const char* base = static_cast<const char*>(ctx->ex_peers[ctx->ex_rank]);
    for (int g = 0; g < nb && exchange; ++g) {
      const char* w = reinterpret_cast<const char*>(grads->branch[g].conv[0].conv_w);
      const size_t bytes = (size_t)32 * bands * 9 * sizeof(float);
      if (!w || w < base || w + bytes > base + ctx->ex_n4 * 16) { exchange = false; break; }
      ex_lo[ex_nw] = (unsigned long long)(w - base) / 16;                 // rounded outward to whole float4
      ex_hi[ex_nw] = (unsigned long long)((w - base) + bytes + 15) / 16;
      if (ex_nw > 0 && ex_lo[ex_nw] < ex_hi[ex_nw - 1]) ex_lo[ex_nw] = ex_hi[ex_nw - 1];
      ++ex_nw;
    }
  }

  // Side stream: everything off the critical path dgrad -> attention backward -> BatchNorm backward -> pack runs there:
  // parameter packing and gradient zero-fill first, then each block's weight gradient (+ split-K reduce) under the NEXT
  // block's attention backward, the batched small-parameter reduction under conv1's weight gradient.
  SideStream side(ctx, st);
  SideStream aux(ctx, st, 1);
  side.wait_main();
  cudaStream_t ss = side.stream();

  // upstream gradients per head (the alpha blend folds djoint into the two last heads)
  const float* dS[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int h = 0; h < d.n_heads; ++h) dS[h] = dscores[h];
  StageScope pre_scope(ctx, "bwd.prologue", st);
  if (hang && djoint) {
    launch_k(joint_bwd_kernel, (int)((nsc + 255) / 256), 256, 0, st, dscores[2], dscores[5], djoint, params->alpha, W.dS[2], W.dS[5], nsc);
    DTA_CHECK_LAUNCH(ctx, "joint_bwd");
    dS[2] = W.dS[2]; dS[5] = W.dS[5];
    if (grads->alpha) {
      launch_k(alpha_grad_kernel, 1, 1024, 0, ss, djoint, L.s3[0], L.s3[1], params->alpha, nsc, grads->alpha);   // off the critical path
      DTA_CHECK_LAUNCH(ctx, "alpha_grad");
    }
  } else if (hang && grads->alpha) {
    cudaMemsetAsync(grads->alpha, 0, sizeof(double), ss);   // off the critical path like every other zero-fill
  }
  auto head_ds = [&](int g, int k) -> const float* {
    if (vanilla) return k == 2 ? dS[0] : nullptr;
    return dS[g * 3 + k];
  };

  // dgrad weight tables
  const bool tcp = ctx->conv_impl == 1;
  const TcGeom tg = tc_geom(B, bands, nb);
  for (int k = 1; k < 3; ++k) {
    Ptr2 w{{params->branch[0].conv[k].conv_w, nb > 1 ? params->branch[1].conv[k].conv_w : nullptr}};
    if (tcp) {
      if (k == 1) launch_k(tc_pack_w_fprop_kernel<32>, ctx->sm_count, 256, 0, ss, w, nb, 64, 32, 4, 2, W.wdp[0]);
      else launch_k(tc_pack_w_fprop_kernel<64>, ctx->sm_count, 256, 0, ss, w, nb, 128, 64, 8, 2, W.wdp[1]);
    } else {
      launch_k(pack_conv_wd_kernel, ctx->sm_count * 2, 256, 0, ss, w, nb, kC[k], kC[k - 1], 0, W.wd[k]);
    }
    DTA_CHECK_LAUNCH(ctx, "pack_conv_wd");
  }
  if (dx) {
    Ptr2 w{{params->branch[0].conv[0].conv_w, nb > 1 ? params->branch[1].conv[0].conv_w : nullptr}};
    launch_k(pack_conv_wd_kernel, ctx->sm_count * 2, 256, 0, st, w, nb, 32, bands, 1, W.wd[0]);
    DTA_CHECK_LAUNCH(ctx, "pack_conv_wd");
  }

  // zero-fill gradients that are only partially written (dead Conv1d taps) or may stay unreached
  for (int g = 0; g < nb; ++g) {
    const dta_branch& gb = grads->branch[g];
    for (int k = 0; k < 3; ++k) {
      if (d.btype[g] == BR_SPECTRAL) {
        const int ks = k == 0 ? 3 : (k == 1 ? 5 : 7);
        if (gb.attn[k].w0) cudaMemsetAsync(gb.attn[k].w0, 0, sizeof(float) * kC[k] * kC[k] * ks, ss);
        if (gb.attn[k].w1) cudaMemsetAsync(gb.attn[k].w1, 0, sizeof(float) * kC[k] * kC[k] * ks, ss);
      }
      const int F = d.btype[g] == BR_SPECTRAL ? kC[k] : (d.btype[g] == BR_SPATIAL ? 4 * kC[k] : 512);
      if (head_ds(g, k) == nullptr) {
        if (gb.fc_w[k]) cudaMemsetAsync(gb.fc_w[k], 0, sizeof(float) * classes * F, ss);
        if (gb.fc_b[k]) cudaMemsetAsync(gb.fc_b[k], 0, sizeof(float) * classes, ss);
      }
    }
  }

  pre_scope.end();
  auto bn_grads = [&](int k) {
    BnGrads g{};
    for (int b = 0; b < nb; ++b) {
      g.dgamma[b] = grads->branch[b].conv[k].bn_w; g.dbeta[b] = grads->branch[b].conv[k].bn_b;
      g.dconv_b[b] = grads->branch[b].conv[k].conv_b;
    }
    return g;
  };
  auto bn_bwd = [&](int k) -> int {
    StageScope sc(ctx, "bwd.bn_finalize", st);
    const int ctot = nb * kC[k];
    launch_k(bn_bwd_finalize_kernel, (ctot + kBnCh - 1) / kBnCh, kBnCh * kBnSlices, 0, st, W.bnrows, B, nb, kC[k], (double)B * kHWpre[k], bn_params(params, d, k),
                                                                    L.bn_mean[k], L.bn_istd[k], shape->training, bn_grads(k), W.k0[k], W.k1[k], W.k2[k]);
    DTA_CHECK_LAUNCH(ctx, "bn_bwd_finalize");
    return DTA_OK;
  };
  // small parameter gradients (attention blocks + heads): batch reductions queued here and
  // run by ONE batched_reduce_kernel launch at the end of the backward pass
  ReduceTaskTable rtab{};
  int rtiles = 0;
  auto add_task = [&](const float* U, size_t ldu, const float* V, size_t ldv, int ni, int nj, float* out, size_t si, size_t sj) -> int {
    if (out == nullptr) return DTA_OK;
    if (rtab.n >= kMaxReduceTasks) return fail(ctx, DTA_ERR_UNSUPPORTED, "too many reduction tasks");
    ReduceTask& t = rtab.t[rtab.n++];
    t.U = U; t.V = V; t.out = out; t.ldu = (long long)ldu; t.ldv = (long long)ldv; t.si = (long long)si; t.sj = (long long)sj;
    t.ni = ni; t.nj = nj; t.tile_begin = rtiles;
    t.tiles_j = V ? (nj + 31) / 32 : 1;
    rtiles += V ? ((ni + 31) / 32) * t.tiles_j : (ni + 31) / 32;
    return DTA_OK;
  };
  auto attn_param_grads = [&](int k) -> int {
    const int C = kC[k], ld = kProwLd[k];
    int r = DTA_OK;
    for (int g = 0; g < nb && r == DTA_OK; ++g) {
      const dta_branch& gb = grads->branch[g];
      const float* prow = W.prow[k] + (size_t)g * ld;
      const size_t prow_ld = (size_t)nb * ld;
      const float* att = L.att[k] + (size_t)g * 3 * kAttRow[k];
      const size_t att_ld = (size_t)nb * 3 * kAttRow[k];
      if (d.btype[g] == BR_SPECTRAL) {
        const int ks = k == 0 ? 3 : (k == 1 ? 5 : 7);
        // dW2[i][j][mid] = sum_b du2[b][i] * h[b][j];  dW1[i][j][mid] = sum_b du1[b][i] * g[b][j]
        if (!r) r = add_task(prow, prow_ld, att + kAttRow[k], att_ld, C, C, gb.attn[k].w1 ? gb.attn[k].w1 + ks / 2 : nullptr, (size_t)C * ks, ks);
        if (!r) r = add_task(prow + C, prow_ld, att, att_ld, C, C, gb.attn[k].w0 ? gb.attn[k].w0 + ks / 2 : nullptr, (size_t)C * ks, ks);
        if (!r) r = add_task(prow, prow_ld, nullptr, 0, C, 1, gb.attn[k].b1, 1, 0);
        if (!r) r = add_task(prow + C, prow_ld, nullptr, 0, C, 1, gb.attn[k].b0, 1, 0);
      } else if (d.btype[g] == BR_SPATIAL) {
        const int ks = k == 0 ? 7 : (k == 1 ? 5 : 3), kk = ks * ks;
        if (!r) r = add_task(prow, prow_ld, nullptr, 0, kk, 1, gb.attn[k].w0, 1, 0);
        if (!r) r = add_task(prow + kk, prow_ld, nullptr, 0, 1, 1, gb.attn[k].b0, 1, 0);
        if (!r) r = add_task(prow + kk + 1, prow_ld, nullptr, 0, kk, 1, gb.attn[k].w1, 1, 0);
        if (!r) r = add_task(prow + 2 * kk + 1, prow_ld, nullptr, 0, 1, 1, gb.attn[k].b1, 1, 0);
        if (!r) r = add_task(prow + 2 * kk + 2, prow_ld, nullptr, 0, C, 1, gb.attn[k].pool_w, 1, 0);
        if (!r) r = add_task(prow + 2 * kk + 2 + C, prow_ld, nullptr, 0, 1, 1, gb.attn[k].pool_b, 1, 0);
      }
      const float* ds = head_ds(g, k);
      if (ds != nullptr && (d.btype[g] != BR_NONE || k == 2)) {
        const int F = d.btype[g] == BR_SPECTRAL ? C : (d.btype[g] == BR_SPATIAL ? 4 * C : 512);
        const float* feat = L.feat[k] + (size_t)g * kFeatLd[k];
        if (!r) r = add_task(ds, classes, feat, (size_t)nb * kFeatLd[k], classes, F, gb.fc_w[k], F, 1);
        if (!r) r = add_task(ds, classes, nullptr, 0, classes, 1, gb.fc_b[k], 1, 0);
      }
    }
    return r;
  };
  auto attn_prm = [&](int k) {
    AttnParams a = attn_params(params, L, d, k, vanilla && k == 2);
    return a;
  };
  auto ds_ptrs = [&](int k) {
    Ptr2 p{{nullptr, nullptr}};
    for (int g = 0; g < nb; ++g) p.p[g] = head_ds(g, k);
    return p;
  };
  auto reduce_w = [&](int k, int cin, int nsplit) -> int {   // on the side stream, right behind the weight gradient it reduces
    StageScope sc(ctx, "bwd.wgrad_reduce", ss);
    MutPtr2 dw{{grads->branch[0].conv[k].conv_w, nb > 1 ? grads->branch[1].conv[k].conv_w : nullptr}};
    const size_t per_branch = (size_t)kC[k] * cin * 9;
    if (tcp) {
      // tensor-core partials are [split][g][tap][ci][co]; conv1 is one group over both branches' output channels
      const int G = k == 0 ? 1 : nb, cout_g = k == 0 ? nb * kC[0] : kC[k];
      launch_k(tc_wgrad_reduce_kernel, ctx->sm_count * 4, 256, 0, ss, W.wpart, nsplit, G, cout_g, cin, dw, per_branch);
    } else
    launch_k(wgrad_reduce_kernel, ctx->sm_count * 4, 256, 0, ss, W.wpart, nsplit, 1, per_branch * nb, dw, per_branch);
    DTA_CHECK_LAUNCH(ctx, "wgrad_reduce");
    return DTA_OK;
  };

  // ---- block 3 ----
  {
    auto kern = attn_bwd_kernel<128, 5, true>;
    const size_t sm = attn_bwd_smem<128, 5, true>(classes);
    allow_smem(kern, sm);
    StageScope asc(ctx, "bwd.attn3", st);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[2], L.bn_scale[2], L.bn_shift[2], L.bn_mean[2], L.bn_istd[2], attn_prm(2), classes, L.att[2], L.feat[2],
                                              ds_ptrs(2), nullptr, W.da[2], W.bnrows, W.prow[2], tcp ? 1 : 0, ce3 ? *ce3 : CeInline{});
    asc.end();
    DTA_CHECK_LAUNCH(ctx, "attn_bwd<3>");
    if ((rc = attn_param_grads(2)) != DTA_OK) return rc;
    if ((rc = bn_bwd(2)) != DTA_OK) return rc;
    ConvSrc dz = src_dz(W.da[2], L.z[2], 128, nb * 128, 25, W.k0[2], W.k1[2], W.k2[2], B, tcp ? 4 : 0);
    ConvSrc in = src_act(L.z[1], 64, nb, 121, 1, L.bn_scale[1], L.bn_shift[1], L.att[1], 3 * kAttRow[1], 2 * kAttRow[1], d.btype);
    side.join();   // packed input-gradient weights (and the gradient zero-fill) are done
    if (tcp) {
      { StageScope sc(ctx, "bwd.conv3_pack", st); DTA_TC_CHECK(run_tc_pack<5>(ctx, st, dz, nb, B, nb * 16, tg.rows5, W.dzp[2]), "tc_pack_stream(dz3)"); }
      {
        StageScope sc(ctx, "bwd.conv3_dgrad", st);
        DTA_TC_CHECK((run_tc_fprop<5, 64, true>(ctx, st, W.dzp[2], tg.rows5, nb * 16, 16, W.wdp[1], 8, Ptr2{{nullptr, nullptr}}, 64, W.dout[1], nb * 64, 64, B, nb)),
                     "tc_conv_fprop(conv3 dgrad)");
      }
      side.wait_main();   // the weight gradient starts when the input gradient has finished, next to block 2's attention backward
      {
        StageScope sc(ctx, "bwd.conv3_wgrad", ss);
        DTA_TC_CHECK((run_tc_wgrad<5, WgradCfg3>(ss, W.dzp[2], nb * 16, L.a2p, nb * 8, tg.rows5, 64, 128, nb, tg.w3, W.wpart)), "tc_conv_wgrad(conv3)");
      }
      if ((rc = reduce_w(2, 64, tg.w3.nsplit)) != DTA_OK) return rc;
    } else {
      { StageScope sc(ctx, "bwd.conv3_dgrad", st); e = launch_fprop<5, 4, 4, 8, 64, 8>(dz, W.wd[2], Ptr2{{nullptr, nullptr}}, 64, W.dout[1], nb * 64, nullptr, B, nb, st, nullptr); }
      if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv3 dgrad: ") + cudaGetErrorString(e));
      ctx->launches++;
      side.wait_main();
      { StageScope sc(ctx, "bwd.conv3_wgrad", ss); e = launch_wgrad<5, 32, 128, 8>(in, dz, W.wpart, B, sp.n[2], sp.per[2], nb, ss); }
      if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv3 wgrad: ") + cudaGetErrorString(e));
      ctx->launches++;
      if ((rc = reduce_w(2, 64, sp.n[2])) != DTA_OK) return rc;
    }
  }
  // ---- block 2 ----
  if (ce3 != nullptr && aux.on) {   // fused training step: the loss kernel on the auxiliary stream has written dscores by now
    aux.pending = true;
    aux.join();
  }
  {
    auto kern = attn_bwd_kernel<64, 11, true>;
    const size_t sm = attn_bwd_smem<64, 11, true>(classes);
    allow_smem(kern, sm);
    StageScope asc(ctx, "bwd.attn2", st);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[1], L.bn_scale[1], L.bn_shift[1], L.bn_mean[1], L.bn_istd[1], attn_prm(1), classes, L.att[1], L.feat[1],
                                              ds_ptrs(1), W.dout[1], W.da[1], W.bnrows, W.prow[1], tcp ? 1 : 0, CeInline{});
    asc.end();
    DTA_CHECK_LAUNCH(ctx, "attn_bwd<2>");
    if ((rc = attn_param_grads(1)) != DTA_OK) return rc;
    if ((rc = bn_bwd(1)) != DTA_OK) return rc;
    ConvSrc dz = src_dz(W.da[1], L.z[1], 64, nb * 64, 121, W.k0[1], W.k1[1], W.k2[1], B, tcp ? 25 : 0);
    ConvSrc in = src_act(L.z[0], 32, nb, 121, 0, L.bn_scale[0], L.bn_shift[0], L.att[0], 3 * kAttRow[0], 2 * kAttRow[0], d.btype);
    if (tcp) {
      { StageScope sc(ctx, "bwd.conv2_pack", st); DTA_TC_CHECK(run_tc_pack<11>(ctx, st, dz, nb, B, nb * 8, tg.rows11, W.dzp[1]), "tc_pack_stream(dz2)"); }
      {
        StageScope sc(ctx, "bwd.conv2_dgrad", st);
        DTA_TC_CHECK((run_tc_fprop<11, 32, true>(ctx, st, W.dzp[1], tg.rows11, nb * 8, 8, W.wdp[0], 4, Ptr2{{nullptr, nullptr}}, 32, W.dout[0], nb * 32, 32, B, nb)),
                     "tc_conv_fprop(conv2 dgrad)");
      }
      side.wait_main();   // next to block 1's attention backward
      {
        StageScope sc(ctx, "bwd.conv2_wgrad", ss);
        DTA_TC_CHECK((run_tc_wgrad<11, WgradCfg2>(ss, W.dzp[1], nb * 8, L.a1p, nb * 4, tg.rows11, 32, 64, nb, tg.w2, W.wpart)), "tc_conv_wgrad(conv2)");
      }
      if ((rc = reduce_w(1, 32, tg.w2.nsplit)) != DTA_OK) return rc;
    } else {
      { StageScope sc(ctx, "bwd.conv2_dgrad", st); e = launch_fprop<11, 1, 8, 4, 32, 16>(dz, W.wd[1], Ptr2{{nullptr, nullptr}}, 32, W.dout[0], nb * 32, nullptr, B, nb, st, nullptr); }
      if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv2 dgrad: ") + cudaGetErrorString(e));
      ctx->launches++;
      side.wait_main();
      { StageScope sc(ctx, "bwd.conv2_wgrad", ss); e = launch_wgrad<11, 32, 64, 8>(in, dz, W.wpart, B, sp.n[1], sp.per[1], nb, ss); }
      if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv2 wgrad: ") + cudaGetErrorString(e));
      ctx->launches++;
      if ((rc = reduce_w(1, 32, sp.n[1])) != DTA_OK) return rc;
    }
  }
  // ---- block 1 ----
  {
    auto kern = attn_bwd_kernel<32, 11, false>;
    const size_t sm = attn_bwd_smem<32, 11, false>(classes);
    allow_smem(kern, sm);
    StageScope asc(ctx, "bwd.attn1", st);
    launch_k(kern, dim3(B, nb), kAttnThreads, sm, st, L.z[0], L.bn_scale[0], L.bn_shift[0], L.bn_mean[0], L.bn_istd[0], attn_prm(0), classes, L.att[0], L.feat[0],
                                              ds_ptrs(0), W.dout[0], W.da[0], W.bnrows, W.prow[0], 0, CeInline{});
    asc.end();
    DTA_CHECK_LAUNCH(ctx, "attn_bwd<1>");
    if ((rc = attn_param_grads(0)) != DTA_OK) return rc;
    // every attention block has left its per-crop partials: ONE batched reduction for all small parameter gradients, on the
    // side stream (under BatchNorm backward / pack / conv1's weight gradient)
    if (rtiles > 0) {
      // on the auxiliary stream when there is one: queued behind conv2's weight gradient it would hold back conv1's, the
      // last kernel of the step (the gradient zero-fill it depends on was joined into the caller's stream in block 3)
      cudaStream_t rs = aux.on ? aux.stream() : ss;
      if (aux.on) aux.wait_main(); else side.wait_main();
      StageScope sc(ctx, "bwd.small_param_grads", rs);
      if (rtiles > reduce_tiles_bound(classes, nb)) return fail(ctx, DTA_ERR_UNSUPPORTED, "reduction tile bound exceeded");
      launch_k(batched_reduce_kernel, dim3(rtiles, kReduceSplits), 256, 0, rs, rtab, B, W.rpart);
      DTA_CHECK_LAUNCH(ctx, "batched_reduce");
      launch_k(batched_reduce_finish_kernel, rtiles, 256, 0, rs, rtab, W.rpart);
      DTA_CHECK_LAUNCH(ctx, "batched_reduce_finish");
    }
    if ((rc = bn_bwd(0)) != DTA_OK) return rc;
    // Gradient exchange, part A (dta_set_grad_exchange): every gradient except conv1's weights is final once the kernels
    // enqueued so far have run -- reduce them over the ranks now, on the auxiliary stream, under conv1's weight gradient.
    if (exchange) {
      ArRanges ra{};
      unsigned long long pos = 0;
      for (int k = 0; k < ex_nw; ++k) {                       // complement of conv1's weight ranges
        if (ex_lo[k] > pos) { ra.lo[ra.n] = pos; ra.hi[ra.n] = ex_lo[k]; ++ra.n; }
        pos = ex_hi[k];
      }
      if (pos < ctx->ex_n4) { ra.lo[ra.n] = pos; ra.hi[ra.n] = ctx->ex_n4; ++ra.n; }
      ra.with_doubles = 1;
      cudaStream_t xs = aux.on ? aux.stream() : (side.on ? ss : st);
      if (aux.on) {
        aux.wait_main();
        if (side.on) { cudaEvent_t e2 = side.next_event(); cudaEventRecord(e2, ss); cudaStreamWaitEvent(xs, e2, 0); }
      } else if (side.on) {
        side.wait_main();
      }
      StageScope sc(ctx, "dist.grad_allreduce", xs);
      if (ra.n > 0 && (rc = launch_allreduce(ctx, xs, ctx->ex_rank, ctx->ex_world, ctx->ex_peers, ctx->ex_mc, ctx->ex_n4, ctx->ex_nd, nullptr, ctx->ex_sync, ra, 0)) != DTA_OK)
        return rc;
    }
    ConvSrc dz = src_dz(W.da[0], L.z[0], nb * 32, nb * 32, 121, W.k0[0], W.k1[0], W.k2[0]);
    ConvSrc in = src_raw(x, bands, kHW);
    int conv1_nsplit = sp.n[0];
    if (tcp) {
      { StageScope sc(ctx, "bwd.conv1_pack", st); DTA_TC_CHECK(run_tc_pack<11>(ctx, st, dz, 1, B, 8, tg.rows11, W.dzp[0]), "tc_pack_stream(dz1)"); }
      side.wait_main();   // the split-K partial buffer is shared with conv2's weight gradient: stay behind it on the side stream
      StageScope sc(ctx, "bwd.conv1_wgrad", ss);
      e = run_tc_wgrad<11, WgradCfg1>(ss, W.dzp[0], 8, L.xp, tg.nchunk1, tg.rows11, bands, nb * 32, 1, tg.w1, W.wpart);
      conv1_nsplit = tg.w1.nsplit;
    } else {
      side.wait_main();
      StageScope sc(ctx, "bwd.conv1_wgrad", ss);
      if (nb == 2) e = launch_wgrad<11, 32, 64, 8>(in, dz, W.wpart, B, sp.n[0], sp.per[0], 1, ss);
      else e = launch_wgrad<11, 32, 32, 8>(in, dz, W.wpart, B, sp.n[0], sp.per[0], 1, ss);
    }
    if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("conv1 wgrad: ") + cudaGetErrorString(e));
    ctx->launches++;
    if ((rc = reduce_w(0, bands, conv1_nsplit)) != DTA_OK) return rc;
    if (exchange) {   // part B: conv1's weight gradient, right behind the kernel that produced it
      ArRanges rb{};
      for (int k = 0; k < ex_nw; ++k) { rb.lo[k] = ex_lo[k]; rb.hi[k] = ex_hi[k]; }
      rb.n = ex_nw;
      StageScope sc(ctx, "dist.grad_allreduce", ss);
      if ((rc = launch_allreduce(ctx, ss, ctx->ex_rank, ctx->ex_world, ctx->ex_peers, ctx->ex_mc, ctx->ex_n4, ctx->ex_nd, nullptr, ctx->ex_sync, rb, 1)) != DTA_OK) return rc;
      ctx->exchanged = 1;
    }
    if (dx) return fail(ctx, DTA_ERR_UNSUPPORTED, "gradient of the crops (dx) is not built yet; the reference feeds requires_grad=False inputs");
  }
  side.join();
  aux.join();
  e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("backward: ") + cudaGetErrorString(e));
  return DTA_OK;
}

int dta_backward(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                 const void* saved, const float* const dscores[6], const float* djoint,
                 const dta_tensors* grads, float* dx, void* workspace, void* cuda_stream) {
  return backward_impl(ctx, shape, x, params, saved, dscores, djoint, grads, dx, workspace, cuda_stream, nullptr);
}

// Forward + loss (sum over the heads of the weighted cross-entropy) + backward as ONE call.  Same kernels and the same bits
// as dta_forward, dta_cross_entropy_heads(all heads), dta_backward(dscores, djoint = NULL) in sequence; what changes is the
// schedule: the loss kernel (losses + every head's score gradient in memory) runs on the side stream, and block 3's
// attention-backward kernel -- the next kernel of the critical chain -- forms its two heads' score gradients itself from the
// scores, so neither the loss kernel nor the alpha blend sits between the forward and the backward pass.
int dta_train_step(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params, const int64_t* labels,
                   const float* class_weight, float* const scores[6], float* joint, float* loss, float* const dscores[6],
                   const dta_tensors* grads, void* saved, void* workspace_fwd, void* workspace_bwd, void* workspace_loss,
                   void* cuda_stream) {
  NetDesc d;
  int rc = check_shape(ctx, shape, &d);
  if (rc != DTA_OK) return rc;
  if (shape->net_kind == DTA_NET_SPECTRAL_PAIR || shape->net_kind == DTA_NET_SPATIAL_PAIR)
    return fail(ctx, DTA_ERR_UNSUPPORTED, "pair kinds are inference fan-out only: no training step");
  if (!labels || !loss || !dscores || !workspace_loss || !scores) return fail(ctx, DTA_ERR_INVALID_ARG, "labels, scores, loss, dscores and workspace_loss are required");
  if (reinterpret_cast<uintptr_t>(workspace_loss) & 255u) return fail(ctx, DTA_ERR_INVALID_ARG, "workspace_loss must be 256-byte aligned");
  for (int h = 0; h < d.n_heads; ++h)
    if (!dscores[h]) return fail(ctx, DTA_ERR_INVALID_ARG, "dscores[h] is NULL for an existing head");
  if (!ctx->tickets) return fail(ctx, DTA_ERR_CUDA, "context has no ticket counters (allocation failed at dta_create)");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  pdl_enabled() = ctx->pdl;
  const int B = shape->batch, classes = shape->classes;
  double* den = reinterpret_cast<double*>(workspace_loss);                      // first 256 bytes of the loss workspace
  float* rows = reinterpret_cast<float*>(static_cast<char*>(workspace_loss) + 256);
  const long long* y = reinterpret_cast<const long long*>(labels);
  long long extra_launches = 0;
  {
    // sum of the label weights: beside the forward pass (its side stream is joined in front of block 1's attention kernel)
    SideStream pre(ctx, st);
    pre.wait_main();
    launch_k(ce_den_kernel, 1, 256, 0, pre.stream(), y, class_weight, B, classes, den);
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "ce_den launch failed");
    ++extra_launches;
    pre.pending = false;   // joined by forward_impl's own join of the same stream
  }
  rc = forward_impl(ctx, shape, 0, x, params, scores, joint, saved, workspace_fwd, cuda_stream, true);
  if (rc != DTA_OK) return rc;
  {
    // losses and all score gradients, off the critical path: on the auxiliary stream (idle until the end of the backward
    // pass; the side stream is busy with the backward pass's parameter tables), joined by backward_impl in front of block 2's
    // attention kernel, the first reader of the gradients in memory
    SideStream post(ctx, st, 1);
    post.wait_main();
    HeadPtrs h{};
    for (int i = 0; i < d.n_heads; ++i) { h.s[i] = scores[i]; h.ds[i] = dscores[i]; }
    StageScope sc(ctx, "loss.cross_entropy", post.stream());
    launch_k(ce_rows_kernel, (unsigned)(((long long)d.n_heads * B * 32 + 255) / 256), 256, 0, post.stream(), h, d.n_heads, y, class_weight, B, classes, rows, loss,
             ctx->tickets + 0);
    if (cudaGetLastError() != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "ce_rows launch failed");
    ++extra_launches;
    if (post.on) post.pending = false;   // joined by backward_impl (same stream), not here
  }
  CeInline ce{};
  const float* ds_const[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int h = 0; h < d.n_heads; ++h) ds_const[h] = dscores[h];
  const bool vanilla = shape->net_kind == DTA_NET_VANILLA;
  for (int g = 0; g < d.nb; ++g) ce.scores[g] = vanilla ? scores[0] : scores[g * 3 + 2];
  ce.y = y; ce.w = class_weight; ce.den = den;
  rc = backward_impl(ctx, shape, x, params, saved, ds_const, nullptr, grads, nullptr, workspace_bwd, cuda_stream, &ce);
  ctx->launches += extra_launches;
  return rc;
}

}  // extern "C"
