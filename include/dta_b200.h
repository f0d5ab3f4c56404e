/*
 * dta_b200.h -- C ABI of the B200-native Hang2020 hot path (libdta_b200.so).
 *
 * The reference (weecology/DeepTreeAttention) has no FFI for this path: the model is a
 * stack of torch.nn layers (src/models/Hang2020.py:14-263) and the "interface" is the
 * Python call  self.model.forward(images)  in TreeModel.training_step (src/main.py:77).
 * This header is the boundary a maintainer would bind instead of that layer stack: one
 * forward and one backward entry point over raw device pointers, no torch types, no C++
 * types, no exceptions.  deeptreeattention_b200/_capi.py is the ctypes binding;
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless named host_*; all tensors are dense,
 *    row-major, in the reference's own layouts (NCHW crops, state_dict() parameter shapes);
 *  - the caller owns every buffer (inputs, outputs, saved activations, workspace); the
 *    library allocates nothing per call and only ENQUEUES work on the given stream;
 *    `saved` and `workspace` must be 256-byte aligned (any cudaMalloc / torch allocation is);
 *  - functions return DTA_OK (0) or a negative dta_status; dta_last_error() gives the text;
 *  - a dta_ctx belongs to one device and one host thread at a time (one per process rank).
 */
#ifndef DTA_B200_H_
#define DTA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DTA_ABI_VERSION 1
#define DTA_IMAGE_SIZE 11   /* config.yml:50 image_size; crops are (bands, 11, 11) */

typedef enum dta_status {
  DTA_OK = 0,
  DTA_ERR_INVALID_ARG = -1,  /* NULL where a pointer is required, bad enum, batch <= 0 ... */
  DTA_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels cover                      */
  DTA_ERR_CUDA = -3,         /* a CUDA runtime call or launch failed                      */
  DTA_ERR_NO_DEVICE = -4     /* no sm_100 device / wrong architecture                     */
} dta_status;

/* Which reference module the call stands in for. */
typedef enum dta_net_kind {
  DTA_NET_HANG2020 = 0, /* Hang2020.Hang2020          (Hang2020.py:242-263): 2 branches + alpha */
  DTA_NET_SPECTRAL = 1, /* Hang2020.spectral_network  (:206-240)                               */
  DTA_NET_SPATIAL = 2,  /* Hang2020.spatial_network   (:170-204)                               */
  DTA_NET_VANILLA = 3   /* Hang2020.vanilla_CNN       (:33-53)                                 */
} dta_net_kind;

typedef struct dta_shape {
  int32_t net_kind; /* dta_net_kind */
  int32_t batch;    /* crops in this call (per GPU)                                   */
  int32_t bands;    /* constructor arg `bands` (369 / 349 / 3 ...)                     */
  int32_t classes;  /* constructor arg `classes`                                       */
  int32_t training; /* 1: BatchNorm batch statistics + running-stat update; 0: eval()  */
} dta_shape;

/* conv_module (Hang2020.py:14-31): Conv2d 3x3 "same" + BatchNorm2d. */
typedef struct dta_conv_block {
  float* conv_w;   /* conv_layer.weight (C, Cin, 3, 3)                      */
  float* conv_b;   /* conv_layer.bias   (C)                                 */
  float* bn_w;     /* bn1.weight (C)                                        */
  float* bn_b;     /* bn1.bias   (C)                                        */
  float* bn_rm;    /* bn1.running_mean (C)  -- read (eval) / updated (train); NULL in a grads table */
  float* bn_rv;    /* bn1.running_var  (C)                                  */
  int64_t* bn_nbt; /* bn1.num_batches_tracked ()  -- incremented in training; may be NULL        */
} dta_conv_block;

/* spectral_attention (Hang2020.py:126-168) or spatial_attention (:68-124).
 *   spectral: w0 = attention_conv1.weight (C,C,ks) b0 = .bias (C)
 *             w1 = attention_conv2.weight (C,C,ks) b1 = .bias (C)          ks = 3/5/7
 *             pool_w = pool_b = NULL
 *   spatial:  pool_w = channel_pool.weight (1,C,1,1)  pool_b = .bias (1)
 *             w0 = attention_conv1.weight (1,1,ks,ks) b0 = .bias (1)
 *             w1 = attention_conv2.weight (1,1,ks,ks) b1 = .bias (1)       ks = 7/5/3   */
typedef struct dta_attention {
  float* pool_w;
  float* pool_b;
  float* w0;
  float* b0;
  float* w1;
  float* b1;
} dta_attention;

/* One branch = spectral_network / spatial_network / vanilla_CNN parameter tree.
 * vanilla_CNN: attn[] unused, fc_w[2]/fc_b[2] = fc1 (classes, 512), fc_w[0..1] NULL. */
typedef struct dta_branch {
  dta_conv_block conv[3];
  dta_attention attn[3];
  float* fc_w[3]; /* classifier{k}.fc1.weight (classes, F_k) */
  float* fc_b[3]; /* classifier{k}.fc1.bias   (classes)      */
} dta_branch;

/* Parameter table (and, with the same shape, the gradient table written by
 * dta_backward).  HANG2020: branch[0] = spectral_network, branch[1] = spatial_network.
 * Other kinds: branch[0] only. */
typedef struct dta_tensors {
  double* alpha; /* Hang2020.alpha, float64 scalar (Hang2020.py:249); NULL for other kinds */
  dta_branch branch[2];
} dta_tensors;

typedef struct dta_sizes {
  size_t saved_bytes;     /* activations kept from forward for backward          */
  size_t workspace_fwd;   /* scratch for dta_forward                              */
  size_t workspace_bwd;   /* scratch for dta_backward                             */
  int32_t n_heads;        /* 6 (HANG2020), 3 (SPECTRAL/SPATIAL), 1 (VANILLA)      */
} dta_sizes;

typedef struct dta_ctx dta_ctx;

int dta_abi_version(void);

/* Creates a context on CUDA device `device`.  Fails with DTA_ERR_NO_DEVICE when there
 * is no GPU: there is no CPU fallback in this library. */
int dta_create(dta_ctx** out, int device);
void dta_destroy(dta_ctx* ctx);
const char* dta_last_error(const dta_ctx* ctx); /* ctx may be NULL: last create error */

/* Tuning / debugging knobs.  key "conv_impl": 0 = fp32 CUDA-core direct convolution,
 * 1 = tcgen05 split-bf16 implicit GEMM (default where implemented).
 * key "launches": read-only counter of kernels launched by the last forward/backward.
 * key "profile": 1 = record per-stage CUDA-event timings (see dta_profile_read). */
int dta_set_option(dta_ctx* ctx, const char* key, int64_t value);
int dta_get_option(const dta_ctx* ctx, const char* key, int64_t* value);

/* Stage timing.  With option "profile" = 1 every forward/backward brackets its stages
 * (conv1 fprop, attention block k, conv1 wgrad, ...) with CUDA events on the caller's
 * stream.  dta_profile_read synchronises the recorded events, folds them into per-stage
 * totals and copies up to `capacity` rows to `out` (`*count` = rows available); `reset`
 * != 0 clears the totals afterwards.  Used by bench.py for the roofline of the dominant
 * kernel; costs two event records per stage, nothing when the option is 0. */
typedef struct dta_stage_time {
  char name[40];
  double total_ms;
  int64_t calls;
} dta_stage_time;
int dta_profile_read(dta_ctx* ctx, dta_stage_time* out, int capacity, int* count, int reset);

/* Pure host arithmetic: buffer sizes for a shape (no GPU needed). */
int dta_query_sizes(const dta_shape* shape, dta_sizes* out);

/*
 * Forward.  Replaces Hang2020.forward / spectral_network.forward / spatial_network.forward /
 * vanilla_CNN.forward (Hang2020.py:251-263, 226-240, 190-204, 45-53).
 *   x        : (batch, bands, 11, 11) float32
 *   scores[i]: (batch, classes) float32 per head, branch-major (spectral 1..3, spatial 1..3);
 *              entries beyond n_heads are ignored, no entry below n_heads may be NULL
 *   joint    : (batch, classes) float32 alpha-blend (HANG2020 only, else may be NULL)
 *   saved    : saved_bytes; must be handed unchanged to dta_backward
 * In training mode the running_mean/var/num_batches_tracked tensors in `params` are updated
 * (momentum 0.1, unbiased variance), like nn.BatchNorm2d.
 */
int dta_forward(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                float* const scores[6], float* joint, void* saved, void* workspace,
                void* cuda_stream);

/*
 * Backward of the forward above (what autograd derives in the reference).
 *   dscores[i]: upstream gradient per head, (batch, classes) float32, NULL = head unused
 *   djoint    : upstream gradient of `joint`, NULL = unused
 *   grads     : same table shape as params; every non-NULL entry is OVERWRITTEN with the
 *               gradient of the matching parameter (zeros where nothing reaches it, e.g. the
 *               dead Conv1d taps); bn_rm/bn_rv/bn_nbt entries are ignored
 *   dx        : (batch, bands, 11, 11) gradient of the crops, or NULL (the reference feeds
 *               requires_grad=False inputs)
 */
int dta_backward(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                 const void* saved, const float* const dscores[6], const float* djoint,
                 const dta_tensors* grads, float* dx, void* workspace, void* cuda_stream);

/*
 * Weighted cross-entropy over n_heads (<= 7) score tensors and its gradient, in one pass.
 * Replaces  F.cross_entropy(y_hat, y, weight=self.loss_weight)  of TreeModel.training_step /
 * validation_step (src/main.py:78,89; weights :66-69), summed over the heads handed in (the
 * reference's regime is n_heads = 1 with the joint scores).
 *   scores[i] : (batch, classes) float32        labels : (batch) int64, 0 <= y < classes
 *   class_weight : (classes) float32 or NULL (= ones, what the reference uses on a CPU host)
 *   loss      : n_heads + 1 floats: per-head  sum_b w[y_b] nll_b / sum_b w[y_b], then their sum
 *   dscores[i]: (batch, classes) d(sum of head losses)/d(scores[i]); dscores or entries may be NULL
 *   workspace : dta_loss_workspace_bytes(batch, n_heads) bytes
 * Labels outside [0, classes) contribute nothing (PyTorch raises; check on the host if needed).
 */
int dta_loss_workspace_bytes(int batch, int n_heads, size_t* out);
int dta_cross_entropy_heads(dta_ctx* ctx, int batch, int classes, int n_heads, const float* const scores[],
                            const int64_t* labels, const float* class_weight, float* loss,
                            float* const dscores[], void* workspace, void* cuda_stream);

/*
 * Raw crown crops -> network input, on the device.  Replaces utils.preprocess_image (src/utils.py:36-57) for
 * crops that are already 11 x 11: drop `clip` bands at each end (the reference drops 10 when there are more than
 * 3 bands), cast int16 -> float32, scale every pixel's spectrum to [0, 1] with the float32 arithmetic of
 * sklearn.preprocessing.minmax_scale(axis=1) (bit-identical; constant spectra map to 0).
 *   raw : (batch, bands_in, 11, 11) int16        out : (batch, bands_in - 2*clip, 11, 11) float32
 * Lets the crops cross PCIe as int16 (half the bytes of the float32 tensors the reference's loader ships).
 */
int dta_preprocess_crops(dta_ctx* ctx, const int16_t* raw, int batch, int bands_in, int clip, float* out,
                         void* cuda_stream);

/*
 * Gradient exchange of data-parallel training as ONE kernel over NVLink / NVSwitch peer memory.  Replaces the
 * gradient all-reduce (mean) that Lightning's DDP wrapper would issue through NCCL when the reference trains with
 * gpus > 1 (train.py:89-98).
 *   peer_buffers[r]  : rank r's gradient buffer as mapped in THIS process (symmetric memory; [rank] is local).
 *                      Layout: n_float4 * 16 bytes of float32 gradients, n_double float64 values, then the flag words
 *                      at flags_offset (dta_grad_allreduce_sizes); the flag words must be zero before the first call.
 *   multicast_buffer : NVSwitch multicast mapping of the same buffers (multimem.ld_reduce sums in the switch), or NULL
 *                      (plain loads from every peer, summed in rank order)
 *   scratch          : scratch_bytes of local device memory; sync_words: 4 zero-initialised uint32 of local memory
 * On return (stream order) the local buffer holds the MEAN over ranks; every rank must make the same call.
 */
int dta_grad_allreduce_sizes(size_t n_float4, size_t n_double, int world, size_t* buffer_bytes, size_t* flags_offset,
                             size_t* scratch_bytes);
int dta_grad_allreduce(dta_ctx* ctx, int rank, int world, void* const peer_buffers[], const void* multicast_buffer,
                       size_t n_float4, size_t n_double, void* scratch, void* sync_words, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DTA_B200_H_ */
