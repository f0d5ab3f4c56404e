// Context object behind the C ABI (include/dta_b200.h) and the helpers every translation unit of the library shares:
// error reporting, launch checks and the optional per-stage CUDA-event timing.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
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
  int conv_impl = 1;   // 1: tcgen05 split-bf16 implicit GEMM (every convolution, forward / input gradient / weight gradient), 0: fp32 SIMT
  long long launches = 0;         // kernels launched by the current / last C-ABI call
  long long launches_total = 0;   // ... by all earlier calls on this context (option "launches_total" = both)
  int profile = 0;
  int fuse_x = 1;      // conv1 forward converts the raw crops itself (no separate pack pass); 0 = pack kernel + pre-packed operand
#ifndef DTA_SMALL_TILES_DEFAULT
#define DTA_SMALL_TILES_DEFAULT 1
#endif
  int small_tiles = DTA_SMALL_TILES_DEFAULT;   // conv1 forward (fused crops): 256-position tiles when 512-position ones would leave SMs idle (small batches): 0 never, 1 eval mode, 2 training too
  std::vector<std::string> stage_names;
  std::vector<double> stage_ms;
  std::vector<long long> stage_calls;
  std::vector<ProfSpan> spans;
  std::vector<cudaEvent_t> free_events;
  std::string err;
  // Side stream for work off the critical path (parameter packing, weight gradients): forked from / joined to the caller's
  // stream with the events below, so the caller still sees one stream-ordered call (and a CUDA-graph capture sees a DAG).
  int overlap = 2;     // 0: caller's stream only; 1: one side stream; 2: + auxiliary stream for the small-parameter reduction
  int pdl = 1;         // programmatic dependent launch between consecutive kernels (launch_k below)
  cudaStream_t side = nullptr;
  cudaStream_t aux = nullptr;   // second side stream (option "overlap" = 2): work that must not queue behind the weight gradients
  std::vector<cudaEvent_t> sync_events;
  size_t sync_next = 0;
  // gradient exchange performed by dta_backward itself (dta_set_grad_exchange); world <= 1: none registered
  int ex_rank = 0, ex_world = 0;
  void* ex_peers[16] = {};
  const void* ex_mc = nullptr;
  size_t ex_n4 = 0, ex_nd = 0;
  void* ex_sync = nullptr;
  int exchanged = 0;   // the last dta_backward exchanged its gradients
  unsigned int* tickets = nullptr;      // 64 self-resetting "last CTA done" counters (allocated once at dta_create)
  const float* update_gate = nullptr;   // dta_set_update_gate: device flag that gates the BatchNorm running-statistics update
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

// Kernel launch with the programmatic-dependent-launch attribute (option "pdl"): consecutive kernels of a stream overlap
// their launch latency; every kernel launched through here starts with pdl_prologue() (dta_common.cuh), which restores
// exact stream order before it touches memory.  Captured into CUDA graphs as programmatic edges.
inline int& pdl_enabled() {
  static thread_local int v = 1;
  return v;
}
template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

// Fork/join of the library-owned side stream around one C-ABI call (option "overlap").  Inactive (everything on the caller's
// stream) when the option is 0, when no side stream exists, or while stage profiling is on (stage times stay attributable).
struct SideStream {
  dta_ctx* ctx;
  cudaStream_t main;
  cudaStream_t side;
  bool on;
  bool pending = false;   // side work enqueued since the last join
  // which = 0: the weight-gradient side stream; which = 1: the auxiliary side stream (active from "overlap" = 2)
  SideStream(dta_ctx* c, cudaStream_t m, int which = 0)
      : ctx(c), main(m), side(which == 0 ? c->side : c->aux),
        on(c->overlap > which && c->profile == 0 && (which == 0 ? c->side : c->aux) != nullptr && !c->sync_events.empty()) {}
  cudaStream_t stream() const { return on ? side : main; }
  cudaEvent_t next_event() {
    cudaEvent_t e = ctx->sync_events[ctx->sync_next];
    ctx->sync_next = (ctx->sync_next + 1) % ctx->sync_events.size();
    return e;
  }
  // the side stream waits for everything enqueued on the caller's stream so far
  void wait_main() {
    if (!on) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, main);
    cudaStreamWaitEvent(side, e, 0);
    pending = true;
  }
  // the caller's stream waits for everything enqueued on the side stream so far
  void join() {
    if (!on || !pending) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, side);
    cudaStreamWaitEvent(main, e, 0);
    pending = false;
  }
  ~SideStream() { join(); }   // error paths: never leave the caller's stream (or a graph capture) with unjoined work
};

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
