// C ABI of the stand-alone building blocks (conv_module, attention modules, Classifier, global_spectral_pool) and of the
// fused Adam step: argument checks and launch sequences for the kernels of dta_blocks.cuh.  Contract: include/dta_b200.h.
#include "dta_blocks.cuh"
#include "dta_ctx.cuh"
#include "dta_metadata.cuh"

using namespace dta;

namespace {

inline int grid_for(size_t total, int sm_count) {
  size_t blocks = (total + kBlkThreads - 1) / kBlkThreads;
  const size_t cap = (size_t)sm_count * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int begin_call(dta_ctx* ctx) {
  if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, "cudaSetDevice failed");
  cudaGetLastError();
  ctx->launches_total += ctx->launches; ctx->launches = 0;
  return DTA_OK;
}

int check_plane(dta_ctx* ctx, const dta_plane* in) {
  if (!in) return fail(ctx, DTA_ERR_INVALID_ARG, "plane shape is NULL");
  if (in->batch <= 0 || in->channels <= 0 || in->height <= 0 || in->width <= 0)
    return fail(ctx, DTA_ERR_INVALID_ARG, "batch, channels, height and width must be positive");
  if ((size_t)in->height * in->width > 4096) return fail(ctx, DTA_ERR_UNSUPPORTED, "planes above 4096 positions are not supported");
  return DTA_OK;
}

// kernel size and class pool of the attention modules (Hang2020.py:77-99, 136-141)
bool attention_geometry(int kind, int filters, int* ks, int* pool) {
  int idx;
  if (filters == 32) idx = 0; else if (filters == 64) idx = 1; else if (filters == 128) idx = 2; else return false;
  if (kind == DTA_ATTN_SPECTRAL) { const int k[3] = {3, 5, 7}; *ks = k[idx]; *pool = 1; return true; }
  if (kind == DTA_ATTN_SPATIAL) { const int k[3] = {7, 5, 3}; const int p[3] = {4, 2, 1}; *ks = k[idx]; *pool = p[idx]; return true; }
  return false;
}

void batch_sum(dta_ctx* ctx, cudaStream_t st, const float* U, size_t ldu, const float* V, size_t ldv, int B, int ni, int nj, float* out,
               size_t si, size_t sj) {
  if (out == nullptr) return;
  blk_batch_sum_kernel<<<grid_for((size_t)ni * nj, ctx->sm_count), kBlkThreads, 0, st>>>(U, ldu, V, ldv, B, ni, nj, out, si, sj);
  ctx->launches++;
}

}  // namespace

extern "C" {

int dta_plane_mean(dta_ctx* ctx, const float* in, size_t rows, int hw, float* out, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!in || !out || rows == 0 || hw <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "in, out, rows and hw are required");
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int wpb = kBlkThreads / 32;
  blk_plane_mean_kernel<<<(unsigned)((rows + wpb - 1) / wpb), kBlkThreads, 0, st>>>(in, rows, hw, out);
  DTA_CHECK_LAUNCH(ctx, "plane_mean");
  return DTA_OK;
}

int dta_plane_mean_backward(dta_ctx* ctx, const float* dout, size_t rows, int hw, float* din, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!dout || !din || rows == 0 || hw <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "dout, din, rows and hw are required");
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  blk_plane_mean_bwd_kernel<<<grid_for(rows * hw, ctx->sm_count), kBlkThreads, 0, st>>>(dout, rows * (size_t)hw, hw, din);
  DTA_CHECK_LAUNCH(ctx, "plane_mean_bwd");
  return DTA_OK;
}

int dta_conv_module_workspace_bytes(const dta_plane* in, int filters, size_t* out) {
  if (!in || !out || filters <= 0 || in->batch <= 0 || in->height <= 0 || in->width <= 0) return DTA_ERR_INVALID_ARG;
  *out = ((size_t)in->batch * filters * in->height * in->width + 3 * (size_t)filters) * sizeof(float);
  return DTA_OK;
}

int dta_conv_module_forward(dta_ctx* ctx, const dta_plane* in, int filters, int pool_h, int pool_w, int training, const float* x,
                            const dta_conv_block* params, float* z, float* stat, float* out, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = check_plane(ctx, in);
  if (rc != DTA_OK) return rc;
  if (filters <= 0 || pool_h <= 0 || pool_w <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "filters and the pooling kernel must be positive");
  if (pool_h > in->height || pool_w > in->width) return fail(ctx, DTA_ERR_INVALID_ARG, "pooling kernel larger than the plane");
  if (!x || !params || !z || !stat || !out) return fail(ctx, DTA_ERR_INVALID_ARG, "x, params, z, stat and out are required");
  if (!params->conv_w || !params->conv_b || !params->bn_w || !params->bn_b || !params->bn_rm || !params->bn_rv)
    return fail(ctx, DTA_ERR_INVALID_ARG, "conv block parameter is NULL");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int B = in->batch, Cin = in->channels, H = in->height, W = in->width, C = filters;
  StageScope sc(ctx, "block.conv_module_fwd", st);
  blk_conv3x3_kernel<<<grid_for((size_t)B * C * H * W, ctx->sm_count), kBlkThreads, 0, st>>>(x, params->conv_w, params->conv_b, z, B, Cin, C, H, W, 0);
  DTA_CHECK_LAUNCH(ctx, "blk_conv3x3");
  blk_bn_stats_kernel<<<C, kBlkThreads, 0, st>>>(z, B, C, H * W, training, params->bn_rm, params->bn_rv,
                                                 reinterpret_cast<long long*>(params->bn_nbt), stat);
  DTA_CHECK_LAUNCH(ctx, "blk_bn_stats");
  blk_bn_relu_pool_kernel<<<grid_for((size_t)B * C * (H / pool_h) * (W / pool_w), ctx->sm_count), kBlkThreads, 0, st>>>(
      z, stat, params->bn_w, params->bn_b, B, C, H, W, pool_h, pool_w, out);
  DTA_CHECK_LAUNCH(ctx, "blk_bn_relu_pool");
  return DTA_OK;
}

int dta_conv_module_backward(dta_ctx* ctx, const dta_plane* in, int filters, int pool_h, int pool_w, int training, const float* x,
                             const dta_conv_block* params, const float* z, const float* stat, const float* dout,
                             const dta_conv_block* grads, float* dx, void* workspace, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = check_plane(ctx, in);
  if (rc != DTA_OK) return rc;
  if (filters <= 0 || pool_h <= 0 || pool_w <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "filters and the pooling kernel must be positive");
  if (!x || !params || !z || !stat || !dout || !grads || !workspace)
    return fail(ctx, DTA_ERR_INVALID_ARG, "x, params, z, stat, dout, grads and workspace are required");
  if (!params->conv_w || !params->bn_w || !params->bn_b) return fail(ctx, DTA_ERR_INVALID_ARG, "conv block parameter is NULL");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int B = in->batch, Cin = in->channels, H = in->height, W = in->width, C = filters;
  const size_t nz = (size_t)B * C * H * W;
  float* da = static_cast<float*>(workspace);
  float* coef = da + nz;
  StageScope sc(ctx, "block.conv_module_bwd", st);
  blk_relu_pool_bwd_kernel<<<grid_for(nz, ctx->sm_count), kBlkThreads, 0, st>>>(z, stat, params->bn_w, params->bn_b, dout, B, C, H, W, pool_h,
                                                                                pool_w, da);
  DTA_CHECK_LAUNCH(ctx, "blk_relu_pool_bwd");
  blk_bn_bwd_stats_kernel<<<C, kBlkThreads, 0, st>>>(da, z, stat, params->bn_w, B, C, H * W, training, grads->bn_w, grads->bn_b,
                                                     grads->conv_b, coef);
  DTA_CHECK_LAUNCH(ctx, "blk_bn_bwd_stats");
  blk_bn_dz_kernel<<<grid_for(nz, ctx->sm_count), kBlkThreads, 0, st>>>(da, z, coef, C, H * W, nz);
  DTA_CHECK_LAUNCH(ctx, "blk_bn_dz");
  if (grads->conv_w) {
    blk_conv3x3_wgrad_kernel<<<dim3(Cin, C), kBlkThreads, 0, st>>>(x, da, B, Cin, C, H, W, grads->conv_w);
    DTA_CHECK_LAUNCH(ctx, "blk_conv3x3_wgrad");
  }
  if (dx) {
    blk_conv3x3_kernel<<<grid_for((size_t)B * Cin * H * W, ctx->sm_count), kBlkThreads, 0, st>>>(da, params->conv_w, nullptr, dx, B, C, Cin, H, W, 1);
    DTA_CHECK_LAUNCH(ctx, "blk_conv3x3(dgrad)");
  }
  return DTA_OK;
}

int dta_attention_sizes(int kind, const dta_plane* in, size_t* feat_per_crop, size_t* saved_floats_per_crop, size_t* workspace_bytes) {
  if (!in || in->batch <= 0 || in->channels <= 0 || in->height <= 0 || in->width <= 0) return DTA_ERR_INVALID_ARG;
  int ks = 0, P = 1;
  if (!attention_geometry(kind, in->channels, &ks, &P)) return DTA_ERR_UNSUPPORTED;
  const size_t C = in->channels, HW = (size_t)in->height * in->width;
  if (kind == DTA_ATTN_SPECTRAL) {
    if (feat_per_crop) *feat_per_crop = C;
    if (saved_floats_per_crop) *saved_floats_per_crop = 3 * C;
    if (workspace_bytes) *workspace_bytes = (size_t)in->batch * 2 * C * sizeof(float);
  } else {
    if (feat_per_crop) *feat_per_crop = C * (in->height / P) * (in->width / P);
    if (saved_floats_per_crop) *saved_floats_per_crop = 3 * HW;
    if (workspace_bytes) *workspace_bytes = (size_t)in->batch * (2 * ks * ks + 2 + C + 1) * sizeof(float);
  }
  return DTA_OK;
}

int dta_attention_forward(dta_ctx* ctx, int kind, const dta_plane* in, const float* x, const dta_attention* params, float* out,
                          float* feat, float* saved, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = check_plane(ctx, in);
  if (rc != DTA_OK) return rc;
  int ks = 0, P = 1;
  if (!attention_geometry(kind, in->channels, &ks, &P))
    return fail(ctx, DTA_ERR_UNSUPPORTED, "Unknown incoming kernel size for attention layers: filters must be 32, 64 or 128");
  if (!x || !params || !out || !feat || !saved) return fail(ctx, DTA_ERR_INVALID_ARG, "x, params, out, feat and saved are required");
  if (!params->w0 || !params->b0 || !params->w1 || !params->b1) return fail(ctx, DTA_ERR_INVALID_ARG, "attention parameter is NULL");
  if (kind == DTA_ATTN_SPATIAL && (!params->pool_w || !params->pool_b)) return fail(ctx, DTA_ERR_INVALID_ARG, "channel_pool parameter is NULL");
  if (kind == DTA_ATTN_SPATIAL && (in->height < P || in->width < P)) return fail(ctx, DTA_ERR_INVALID_ARG, "plane smaller than the class pool");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int B = in->batch, C = in->channels, H = in->height, W = in->width;
  StageScope sc(ctx, "block.attention_fwd", st);
  if (kind == DTA_ATTN_SPECTRAL) {
    blk_spectral_fwd_kernel<<<B, kBlkThreads, 3 * C * sizeof(float), st>>>(x, C, H * W, ks, params->w0, params->b0, params->w1, params->b1, out,
                                                                        feat, saved);
  } else {
    const size_t sm = (size_t)3 * H * W * sizeof(float);
    cudaFuncSetAttribute(blk_spatial_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    blk_spatial_fwd_kernel<<<B, kBlkThreads, sm, st>>>(x, C, H, W, ks, P, params->pool_w, params->pool_b, params->w0, params->b0, params->w1,
                                                       params->b1, out, feat, saved);
  }
  DTA_CHECK_LAUNCH(ctx, "blk_attention_fwd");
  return DTA_OK;
}

int dta_attention_backward(dta_ctx* ctx, int kind, const dta_plane* in, const float* x, const dta_attention* params, const float* saved,
                           const float* dout, const float* dfeat, float* dx, const dta_attention* grads, void* workspace,
                           void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = check_plane(ctx, in);
  if (rc != DTA_OK) return rc;
  int ks = 0, P = 1;
  if (!attention_geometry(kind, in->channels, &ks, &P))
    return fail(ctx, DTA_ERR_UNSUPPORTED, "Unknown incoming kernel size for attention layers: filters must be 32, 64 or 128");
  if (!x || !params || !saved || !grads || !workspace) return fail(ctx, DTA_ERR_INVALID_ARG, "x, params, saved, grads and workspace are required");
  if (!params->w0 || !params->w1) return fail(ctx, DTA_ERR_INVALID_ARG, "attention parameter is NULL");
  if (kind == DTA_ATTN_SPATIAL && !params->pool_w) return fail(ctx, DTA_ERR_INVALID_ARG, "channel_pool parameter is NULL");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const int B = in->batch, C = in->channels, H = in->height, W = in->width;
  float* prow = static_cast<float*>(workspace);
  StageScope sc(ctx, "block.attention_bwd", st);
  if (kind == DTA_ATTN_SPECTRAL) {
    blk_spectral_bwd_kernel<<<B, kBlkThreads, 6 * C * sizeof(float), st>>>(x, C, H * W, ks, params->w0, params->w1, saved, dout, dfeat, dx, prow);
    DTA_CHECK_LAUNCH(ctx, "blk_spectral_bwd");
    // dW2[i][j][mid] = sum_b du2[b][i] * h1[b][j];  dW1[i][j][mid] = sum_b du1[b][i] * g[b][j]; off-centre taps see only padding
    const size_t wn = (size_t)C * C * ks;
    if (grads->w1) cudaMemsetAsync(grads->w1, 0, wn * sizeof(float), st);
    if (grads->w0) cudaMemsetAsync(grads->w0, 0, wn * sizeof(float), st);
    batch_sum(ctx, st, prow, 2 * C, saved + C, 3 * C, B, C, C, grads->w1 ? grads->w1 + ks / 2 : nullptr, (size_t)C * ks, ks);
    batch_sum(ctx, st, prow + C, 2 * C, saved, 3 * C, B, C, C, grads->w0 ? grads->w0 + ks / 2 : nullptr, (size_t)C * ks, ks);
    batch_sum(ctx, st, prow, 2 * C, nullptr, 0, B, C, 1, grads->b1, 1, 0);
    batch_sum(ctx, st, prow + C, 2 * C, nullptr, 0, B, C, 1, grads->b0, 1, 0);
  } else {
    const size_t sm = (size_t)6 * H * W * sizeof(float);
    cudaFuncSetAttribute(blk_spatial_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    blk_spatial_bwd_kernel<<<B, kBlkThreads, sm, st>>>(x, C, H, W, ks, P, params->pool_w, params->w0, params->w1, saved, dout, dfeat, dx, prow);
    DTA_CHECK_LAUNCH(ctx, "blk_spatial_bwd");
    const int kk = ks * ks;
    const size_t ld = (size_t)2 * kk + 2 + C + 1;
    batch_sum(ctx, st, prow, ld, nullptr, 0, B, kk, 1, grads->w0, 1, 0);
    batch_sum(ctx, st, prow + kk, ld, nullptr, 0, B, 1, 1, grads->b0, 1, 0);
    batch_sum(ctx, st, prow + kk + 1, ld, nullptr, 0, B, kk, 1, grads->w1, 1, 0);
    batch_sum(ctx, st, prow + 2 * kk + 1, ld, nullptr, 0, B, 1, 1, grads->b1, 1, 0);
    batch_sum(ctx, st, prow + 2 * kk + 2, ld, nullptr, 0, B, C, 1, grads->pool_w, 1, 0);
    batch_sum(ctx, st, prow + 2 * kk + 2 + C, ld, nullptr, 0, B, 1, 1, grads->pool_b, 1, 0);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("attention backward: ") + cudaGetErrorString(e));
  return DTA_OK;
}

int dta_classifier_forward(dta_ctx* ctx, int batch, int in_features, int classes, const float* feat, const float* w, const float* b,
                           float* scores, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (batch <= 0 || in_features <= 0 || classes <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "batch, in_features and classes must be positive");
  if (!feat || !w || !scores) return fail(ctx, DTA_ERR_INVALID_ARG, "feat, w and scores are required");
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  blk_linear_kernel<<<grid_for((size_t)batch * classes, ctx->sm_count), kBlkThreads, 0, st>>>(feat, w, b, batch, in_features, classes, scores);
  DTA_CHECK_LAUNCH(ctx, "blk_linear");
  return DTA_OK;
}

int dta_classifier_backward(dta_ctx* ctx, int batch, int in_features, int classes, const float* feat, const float* w,
                            const float* dscores, float* dfeat, float* dw, float* db, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (batch <= 0 || in_features <= 0 || classes <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "batch, in_features and classes must be positive");
  if (!feat || !w || !dscores) return fail(ctx, DTA_ERR_INVALID_ARG, "feat, w and dscores are required");
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (dfeat) {
    blk_linear_dinput_kernel<<<grid_for((size_t)batch * in_features, ctx->sm_count), kBlkThreads, 0, st>>>(dscores, w, batch, in_features, classes, dfeat);
    DTA_CHECK_LAUNCH(ctx, "blk_linear_dinput");
  }
  batch_sum(ctx, st, dscores, classes, feat, in_features, batch, classes, in_features, dw, in_features, 1);
  batch_sum(ctx, st, dscores, classes, nullptr, 0, batch, classes, 1, db, 1, 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("classifier backward: ") + cudaGetErrorString(e));
  return DTA_OK;
}

int dta_crops_nonzero(dta_ctx* ctx, int n_years, const float* const crops[], size_t elems, float* flags, void* workspace,
                      void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (n_years <= 0 || n_years > kMaxYears) return fail(ctx, DTA_ERR_INVALID_ARG, "need 1 <= n_years <= 16");
  if (!crops || !flags || !workspace || elems == 0) return fail(ctx, DTA_ERR_INVALID_ARG, "crops, flags, workspace and elems are required");
  YearPtrs yp{};
  for (int y = 0; y < n_years; ++y) {
    if (!crops[y]) return fail(ctx, DTA_ERR_INVALID_ARG, "crops[y] is NULL");
    yp.p[y] = crops[y];
  }
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  StageScope sc(ctx, "year.crops_nonzero", st);
  size_t want = (elems / 4 + kBlkThreads - 1) / kBlkThreads;
  int nblk = (int)(want < 1 ? 1 : (want > (size_t)kYearSumBlocks ? (size_t)kYearSumBlocks : want));
  float* partial = static_cast<float*>(workspace);
  crops_sum_partial_kernel<<<dim3(nblk, n_years), kBlkThreads, 0, st>>>(yp, elems, partial);
  DTA_CHECK_LAUNCH(ctx, "crops_sum_partial");
  crops_nonzero_finish_kernel<<<n_years, 32, 0, st>>>(partial, nblk, flags);
  DTA_CHECK_LAUNCH(ctx, "crops_nonzero_finish");
  return DTA_OK;
}

int dta_ensemble_mean(dta_ctx* ctx, int n_years, const float* const scores[], const float* flags, int batch, int classes,
                      int softmax, float* out, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (n_years <= 0 || n_years > kMaxYears) return fail(ctx, DTA_ERR_INVALID_ARG, "need 1 <= n_years <= 16");
  if (!scores || !out || batch <= 0 || classes <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "scores, out, batch and classes are required");
  YearPtrs yp{};
  for (int y = 0; y < n_years; ++y) {
    if (!scores[y]) return fail(ctx, DTA_ERR_INVALID_ARG, "scores[y] is NULL");
    yp.p[y] = scores[y];
  }
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  StageScope sc(ctx, "year.ensemble_mean", st);
  const int rows_per_block = kBlkThreads / 32;
  ensemble_mean_kernel<<<(batch + rows_per_block - 1) / rows_per_block, kBlkThreads, 0, st>>>(yp, flags, n_years, batch, classes, softmax, out);
  DTA_CHECK_LAUNCH(ctx, "ensemble_mean");
  return DTA_OK;
}

int dta_adam_step(dta_ctx* ctx, int n_tensors, float* const params[], const float* const grads[], const int64_t numel[],
                  const int64_t offset[], float* exp_avg, float* exp_avg_sq, double* param64, const double* grad64,
                  double* moments64, const dta_adam_hyper* h, int64_t* step_device, const float* lr_device, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  if (!h) return fail(ctx, DTA_ERR_INVALID_ARG, "hyper-parameters are NULL");
  if (n_tensors < 0 || n_tensors > kAdamMaxTensors) return fail(ctx, DTA_ERR_UNSUPPORTED, "at most 96 tensors per dta_adam_step call");
  if (n_tensors > 0 && (!params || !grads || !numel || !offset || !exp_avg || !exp_avg_sq))
    return fail(ctx, DTA_ERR_INVALID_ARG, "params, grads, numel, offset and the moment buffers are required");
  if ((param64 != nullptr) != (grad64 != nullptr) || (param64 && !moments64))
    return fail(ctx, DTA_ERR_INVALID_ARG, "param64, grad64 and moments64 go together");
  if (n_tensors == 0 && !param64) return fail(ctx, DTA_ERR_INVALID_ARG, "nothing to update");
  if (!step_device && h->step < 1) return fail(ctx, DTA_ERR_INVALID_ARG, "step is 1-based");
  if (!(h->beta1 >= 0.0 && h->beta1 < 1.0 && h->beta2 >= 0.0 && h->beta2 < 1.0 && h->eps >= 0.0))
    return fail(ctx, DTA_ERR_INVALID_ARG, "need 0 <= beta < 1 and eps >= 0");
  AdamTable tab{};
  int chunks = 0;
  for (int i = 0; i < n_tensors; ++i) {
    if (!params[i] || !grads[i] || numel[i] <= 0 || offset[i] < 0 || numel[i] > 0x7fffffff || offset[i] + numel[i] > 0x7fffffff)
      return fail(ctx, DTA_ERR_INVALID_ARG, "bad tensor entry (NULL pointer, empty tensor or offset beyond 2^31)");
    tab.p[i] = params[i];
    tab.g[i] = grads[i];
    tab.chunk_begin[i] = chunks;
    tab.elem_begin[i] = (int)offset[i];
    tab.numel[i] = (int)numel[i];
    chunks += (int)((numel[i] + kAdamChunk - 1) / kAdamChunk);
  }
  tab.n = n_tensors;
  tab.chunk_begin[n_tensors] = chunks;
  int rc = begin_call(ctx);
  if (rc != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  StageScope sc(ctx, "optim.adam", st);
  if (step_device) {
    adam_tick_kernel<<<1, 1, 0, st>>>(reinterpret_cast<long long*>(step_device));
    DTA_CHECK_LAUNCH(ctx, "adam_tick");
  }
  AdamHyper hy{};
  hy.lr = (float)h->lr;
  hy.beta1 = (float)h->beta1;
  hy.beta2 = (float)h->beta2;
  hy.one_minus_beta1 = (float)(1.0 - h->beta1);
  hy.one_minus_beta2 = (float)(1.0 - h->beta2);
  hy.eps = (float)h->eps;
  hy.weight_decay = (float)h->weight_decay;
  hy.lr64 = h->lr;
  hy.beta1_64 = h->beta1;
  hy.beta2_64 = h->beta2;
  hy.eps64 = h->eps;
  hy.step = h->step;
  adam_step_kernel<<<chunks > 0 ? chunks : 1, kBlkThreads, 0, st>>>(tab, hy, reinterpret_cast<const long long*>(step_device), lr_device, exp_avg,
                                                                   exp_avg_sq, param64, grad64, moments64);
  DTA_CHECK_LAUNCH(ctx, "adam_step");
  return DTA_OK;
}

// ---- metadata / metadata_sensor_fusion (src/models/metadata.py:9-44) -----------------------------------------------
int dta_metadata_sizes(int batch, int classes, size_t* saved_bytes, size_t* workspace_bytes) {
  if (batch <= 0 || classes <= 0 || !saved_bytes || !workspace_bytes) return DTA_ERR_INVALID_ARG;
  *saved_bytes = meta_saved_layout(nullptr, batch, classes).floats * sizeof(float);
  *workspace_bytes = meta_work_layout(nullptr, batch, classes).floats * sizeof(float);
  return DTA_OK;
}

static MetaTensors meta_tensors(const dta_metadata_tensors* t) {
  MetaTensors m{};
  m.emb = t->embedding; m.bn_w = t->bn_w; m.bn_b = t->bn_b; m.bn_rm = t->bn_rm; m.bn_rv = t->bn_rv;
  m.bn_nbt = reinterpret_cast<long long*>(t->bn_nbt);
  m.mlp_w = t->mlp_w; m.mlp_b = t->mlp_b; m.fc_w = t->fc_w; m.fc_b = t->fc_b;
  return m;
}

static int meta_check(dta_ctx* ctx, int batch, int sites, int classes, const int64_t* site, const float* sensor,
                      const dta_metadata_tensors* p, int need_running) {
  if (batch <= 0 || sites <= 0 || classes <= 0) return fail(ctx, DTA_ERR_INVALID_ARG, "batch, sites and classes must be positive");
  if (classes > 4096) return fail(ctx, DTA_ERR_UNSUPPORTED, "classes > 4096 not supported");
  if (!site || !p) return fail(ctx, DTA_ERR_INVALID_ARG, "site and params are required");
  if (!p->embedding || !p->bn_w || !p->bn_b || !p->mlp_w || !p->mlp_b) return fail(ctx, DTA_ERR_INVALID_ARG, "metadata parameter is NULL");
  if (need_running && (!p->bn_rm || !p->bn_rv)) return fail(ctx, DTA_ERR_INVALID_ARG, "BatchNorm1d running statistics are NULL");
  if (sensor && (!p->fc_w || !p->fc_b)) return fail(ctx, DTA_ERR_INVALID_ARG, "fusion layer parameter is NULL");
  return DTA_OK;
}

int dta_metadata_forward(dta_ctx* ctx, int batch, int sites, int classes, int training, const int64_t* site, const float* sensor,
                         const dta_metadata_tensors* params, const uint8_t* keep_mask, uint64_t seed, float* out, void* saved,
                         void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = meta_check(ctx, batch, sites, classes, site, sensor, params, !training);
  if (rc != DTA_OK) return rc;
  if (!out || !saved) return fail(ctx, DTA_ERR_INVALID_ARG, "out and saved are required");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  StageScope sc(ctx, "meta.forward", st);
  const MetaTensors p = meta_tensors(params);
  const MetaSaved sv = meta_saved_layout(saved, batch, classes);
  const long long* sidx = reinterpret_cast<const long long*>(site);
  meta_bn_stats_kernel<<<1, kMetaStatThreads, 0, st>>>(sidx, batch, sites, p, training, sv.mean, sv.istd);
  DTA_CHECK_LAUNCH(ctx, "meta_bn_stats");
  const size_t smem = sizeof(float) * kMetaCrops * (kMetaDim + 2 * (size_t)classes);
  cudaFuncSetAttribute(meta_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  meta_fwd_kernel<<<(batch + kMetaCrops - 1) / kMetaCrops, kMetaThreads, smem, st>>>(sidx, sensor, batch, sites, classes, p, training, keep_mask,
                                                                                  (unsigned long long)seed, sv, out);
  DTA_CHECK_LAUNCH(ctx, "meta_fwd");
  return DTA_OK;
}

int dta_metadata_backward(dta_ctx* ctx, int batch, int sites, int classes, int training, const int64_t* site, const float* sensor,
                          const dta_metadata_tensors* params, const void* saved, const float* out, const float* dout,
                          const dta_metadata_tensors* grads, float* dsensor, void* workspace, void* cuda_stream) {
  if (!ctx) return DTA_ERR_INVALID_ARG;
  int rc = meta_check(ctx, batch, sites, classes, site, sensor, params, 0);
  if (rc != DTA_OK) return rc;
  if (!saved || !out || !dout || !grads || !workspace) return fail(ctx, DTA_ERR_INVALID_ARG, "saved, out, dout, grads and workspace are required");
  if ((rc = begin_call(ctx)) != DTA_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  StageScope sc(ctx, "meta.backward", st);
  const MetaTensors p = meta_tensors(params);
  const MetaSaved sv = meta_saved_layout(const_cast<void*>(saved), batch, classes);
  const MetaWork wk = meta_work_layout(workspace, batch, classes);
  const long long* sidx = reinterpret_cast<const long long*>(site);
  const int fused = sensor != nullptr;
  const int C = classes;
  const size_t smem = sizeof(float) * kMetaCrops * 2 * (size_t)C;
  cudaFuncSetAttribute(meta_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  meta_bwd_rows_kernel<<<(batch + kMetaCrops - 1) / kMetaCrops, kMetaThreads, smem, st>>>(sidx, batch, sites, C, fused, p, sv, out, dout, wk, dsensor);
  DTA_CHECK_LAUNCH(ctx, "meta_bwd_rows");
  // every batch reduction in one launch (fixed order): fusion weights / bias, site-MLP weights / bias, BatchNorm1d sums
  MetaReduceTable tab{};
  int tiles = 0;
  auto add = [&](const float* U, int ldu, const float* V, int ldv, int ni, int nj, float* out, int si, int sj) {
    if (!out) return;
    MetaReduceTask& t = tab.t[tab.n++];
    t.U = U; t.V = V; t.out = out; t.ldu = ldu; t.ldv = ldv; t.si = si; t.sj = sj; t.ni = ni; t.nj = nj;
    t.tile_begin = tiles; t.tiles_j = V ? (nj + 31) / 32 : 1;
    tiles += V ? ((ni + 31) / 32) * t.tiles_j : (ni + 31) / 32;
  };
  if (fused) {
    // dfc_w[i][j] = sum_b U1[b][i] * cat[b][j], cat = [m | sensor]
    add(wk.u1, C, sv.m, C, C, C, grads->fc_w, 2 * C, 1);
    add(wk.u1, C, sensor, C, C, C, grads->fc_w ? grads->fc_w + C : nullptr, 2 * C, 1);
    add(wk.u1, C, nullptr, 0, C, 1, grads->fc_b, 1, 0);
  }
  add(wk.u2, C, sv.d, kMetaDim, C, kMetaDim, grads->mlp_w, kMetaDim, 1);
  add(wk.u2, C, nullptr, 0, C, 1, grads->mlp_b, 1, 0);
  add(wk.dyx, kMetaDim, nullptr, 0, kMetaDim, 1, wk.dgamma, 1, 0);
  add(wk.dy, kMetaDim, nullptr, 0, kMetaDim, 1, wk.dbeta, 1, 0);
  meta_reduce_kernel<<<tiles, 256, 0, st>>>(tab, batch);
  DTA_CHECK_LAUNCH(ctx, "meta_reduce");
  if (grads->bn_w || grads->bn_b) {
    meta_copy16_kernel<<<1, 32, 0, st>>>(wk.dgamma, grads->bn_w, wk.dbeta, grads->bn_b);
    DTA_CHECK_LAUNCH(ctx, "meta_copy16");
  }
  if (grads->embedding) {
    meta_bwd_embed_kernel<<<sites, 256, 0, st>>>(sidx, batch, sites, p, sv, wk, training, grads->embedding);
    DTA_CHECK_LAUNCH(ctx, "meta_bwd_embed");
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string("metadata backward: ") + cudaGetErrorString(e));
  return DTA_OK;
}

}  // extern "C"
