// Context object behind the C ABI (include/dta_b200.h) and the helpers every translation unit of the library shares:
// error reporting, launch checks and the optional per-stage CUDA-event timing.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dta_common.cuh"

// Stage timing (option "profile"): CUDA events recorded on the caller's stream around each
// named stage; dta_profile_read() folds them into per-stage totals.
struct ProfSpan {
  int stage;
  cudaEvent_t t0, t1;
};
struct dta_ctx {
  int device = 0;
  int sm_count = 0;
  int conv_impl = 1;   // 1: tcgen05 split-bf16 implicit GEMM where built (conv1 forward + weight gradient), 0: fp32 SIMT
  long long launches = 0;
  int profile = 0;
  int fuse_x = 1;      // conv1 forward converts the raw crops itself (no separate pack pass); 0 = pack kernel + pre-packed operand
  std::vector<std::string> stage_names;
  std::vector<double> stage_ms;
  std::vector<long long> stage_calls;
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> free_events;
  std::string err;
  // Side stream for work off the critical path (parameter packing, weight gradients): forked from / joined to the caller's
  // stream with the events below, so the caller still sees one stream-ordered call (and a CUDA-graph capture sees a DAG).
  int overlap = 1;
  cudaStream_t side = nullptr;
  std::vector<cudaEvent_t> sync_events;
  size_t sync_next = 0;
};


namespace dta {

inline int stage_index(dta_ctx* ctx, const char* name) {
  for (size_t i = 0; i < ctx->stage_names.size(); ++i)
    if (ctx->stage_names[i] == name) return (int)i;
  ctx->stage_names.push_back(name);
  ctx->stage_ms.push_back(0.0);
  ctx->stage_calls.push_back(0);
  return (int)ctx->stage_names.size() - 1;
}
inline cudaEvent_t take_event(dta_ctx* ctx) {
  if (!ctx->free_events.empty()) {
    cudaEvent_t e = ctx->free_events.back();
    ctx->free_events.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// RAII bracket around the launches of one stage.
struct StageScope {
  dta_ctx* ctx;
  cudaStream_t st;
  ProfSpan span{};
  bool on;
  StageScope(dta_ctx* c, const char* name, cudaStream_t s) : ctx(c), st(s), on(c->profile != 0) {
    if (!on) return;
    span.stage = stage_index(c, name);
    span.t0 = take_event(c);
    span.t1 = take_event(c);
    cudaEventRecord(span.t0, st);
  }
  // Closes the bracket (idempotent): the destructor calls it, long stages call it early.
  void end() {
    if (!on) return;
    on = false;
    cudaEventRecord(span.t1, st);
    ctx->spans.push_back(span);
  }
  ~StageScope() { end(); }
};
inline void fold_spans(dta_ctx* ctx) {
  for (ProfSpan& sp : ctx->spans) {
    float ms = 0.f;
    if (cudaEventSynchronize(sp.t1) == cudaSuccess && cudaEventElapsedTime(&ms, sp.t0, sp.t1) == cudaSuccess) {
      ctx->stage_ms[sp.stage] += ms;
      ctx->stage_calls[sp.stage] += 1;
    }
    ctx->free_events.push_back(sp.t0);
    ctx->free_events.push_back(sp.t1);
  }
  ctx->spans.clear();
  cudaGetLastError();
}

inline int fail(dta_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define DTA_CHECK_LAUNCH(ctx, what)                                                        \
  do {                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess)                                                                \
      return fail(ctx, DTA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e__));  \
    (ctx)->launches++;                                                                     \
  } while (0)

// Same for the launch helpers that return the launch status.
#define DTA_TC_CHECK(expr, what)                                                                          \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess) return fail(ctx, DTA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e__)); \
    ctx->launches++;                                                                                      \
  } while (0)


}  // namespace dta
