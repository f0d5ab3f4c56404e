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
  DTA_NET_VANILLA = 3,  /* Hang2020.vanilla_CNN       (:33-53)                                 */
  /* inference fan-out (dta_forward_pair, dta_query_sizes): TWO networks of one kind on the same crops */
  DTA_NET_SPECTRAL_PAIR = 4,
  DTA_NET_SPATIAL_PAIR = 5
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
 * key "launches": read-only counter of kernels launched by the last C-ABI call; "launches_total": by every call on this
 * context so far (difference of two reads = kernels of the calls in between).
 * key "profile": 1 = record per-stage CUDA-event timings (see dta_profile_read).
 * key "overlap": 2 (default) / 1 = work off the critical path (parameter packing, weight gradients, small reductions) runs on a
 * library-owned side stream, forked from and joined back to the caller's stream with events inside each call (still one
 * stream-ordered call for the caller; a CUDA-graph capture records a DAG); 2 adds a second side stream so that the batched
 * small-parameter reduction does not queue between two weight gradients; 0 = everything on the caller's stream.
 * key "pdl": 1 (default) = consecutive kernels of a call are launched with programmatic dependent launch (the next kernel's
 * CTAs are scheduled while the previous grid drains and wait for its completion before touching memory): same stream order,
 * less launch latency between the ~60 short kernels of a step; 0 = plain launches.
 * key "fuse_x": 1 (default) = conv1's forward kernel converts the raw float32 crops into its split-bf16 operand itself (and
 * leaves the packed copy the weight gradient reads); 0 = a separate pack pass in front.  Same bits either way.
 * key "small_tiles": conv1 forward at batches whose 512-position tiles would leave SMs idle (< ~526 crops): 256-position tiles
 * and two accumulator stages.  1 (default) = in eval mode (bit-identical there), 2 = in training too (the per-CTA grouping
 * of the BatchNorm partial sums then moves the batch statistics in their last bits), 0 = never. */
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

/* Pure host arithmetic: where inside `saved` the forward leaves, for convolution block `block` (0..2),
 *   region 0: the convolution output z -- the pre-BatchNorm value of conv_module.forward (Hang2020.py:25), bias included --
 *             as NCHW float32 (batch, branches * C_block, S, S), branch-major on the channel axis (spectral then spatial for
 *             HANG2020), S = 11, 11, 5;
 *   region 1 / 2: the per-channel BatchNorm scale = gamma * invstd and shift = beta - mean * scale (branches * C_block floats
 *             each) every kernel applies as a = fmaf(z, scale, shift).
 * Diagnostic view for parity tests (they hand these values to the float64 oracle so that both sides take every ReLU /
 * max-pool decision on the same numbers); nothing on the product path calls it. */
int dta_saved_region(const dta_shape* shape, int block, int region, size_t* offset_bytes, size_t* n_floats);

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
 * Prediction fan-out.  MultiStage.predict_step (src/models/multi_stage.py:306-318) runs every level's model on the SAME
 * crops; per year that is one spectral_network per level on one crop tensor.  This call evaluates TWO such networks in one
 * pass: the crops are read once and block 1's convolution runs with both networks' filters side by side (N = 64, exactly
 * the Hang2020 two-branch shape of dta_forward).  Eval mode only (running statistics; nothing is updated).
 *   shape      : net_kind DTA_NET_SPECTRAL_PAIR / DTA_NET_SPATIAL_PAIR, training = 0, classes = class count of network 0
 *   classes_second : class count of network 1 (levels have different label sets)
 *   params     : branch[0] = network 0, branch[1] = network 1 (alpha ignored)
 *   scores     : [0..2] = heads of network 0 (batch, classes), [3..5] = heads of network 1 (batch, classes_second)
 *   saved / workspace : dta_query_sizes(shape) with the pair kind and classes = max of the two
 */
int dta_forward_pair(dta_ctx* ctx, const dta_shape* shape, int classes_second, const float* x, const dta_tensors* params,
                     float* const scores[6], void* saved, void* workspace, void* cuda_stream);

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
 * One training step of the regime the north star names, in ONE call: forward, loss = sum over the network's heads of the
 * class-weighted cross-entropy, backward.  Replaces, for a caller that sums the head losses,
 *     y_hat = self.model.forward(images); loss = F.cross_entropy(y_hat, y, weight=self.loss_weight); loss.backward()
 * of TreeModel.training_step (src/main.py:71-80) plus the autograd pass Lightning runs behind it.  Results are bit-identical to
 * dta_forward + dta_cross_entropy_heads(every head) + dta_backward(dscores, djoint = NULL); only the schedule differs: the
 * loss kernel leaves the critical path (side stream) and block 3's attention-backward kernel forms its heads' score gradients
 * from the scores itself, so nothing sits between the last forward kernel and the first backward kernel.
 *   labels (batch) int64, class_weight (classes) float32 or NULL; scores / joint / saved / workspace_fwd as dta_forward
 *   (joint may be NULL: the alpha blend is skipped; alpha's gradient is zero in this regime either way)
 *   loss      : n_heads + 1 floats, as dta_cross_entropy_heads
 *   dscores[h]: (batch, classes) float32 per existing head, caller-owned, receives d loss / d scores[h]
 *   grads / workspace_bwd as dta_backward;  workspace_loss : dta_loss_workspace_bytes(batch, n_heads) bytes, 256-byte aligned
 */
int dta_train_step(dta_ctx* ctx, const dta_shape* shape, const float* x, const dta_tensors* params,
                   const int64_t* labels, const float* class_weight, float* const scores[6], float* joint,
                   float* loss, float* const dscores[6], const dta_tensors* grads, void* saved,
                   void* workspace_fwd, void* workspace_bwd, void* workspace_loss, void* cuda_stream);

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
 *   scratch          : scratch_bytes of local device memory; sync_words: 8 zero-initialised uint32 of local memory
 * On return (stream order) the local buffer holds the MEAN over ranks; every rank must make the same call.
 *
 * dta_set_grad_exchange registers the same buffers with the context (multicast mapping required).  From then on a
 * dta_backward whose gradient table starts at peer_buffers[rank] performs the exchange ITSELF, in two launches: everything
 * but conv1's weight gradient as soon as it is final -- on a side stream, under conv1's weight-gradient kernel, the longest
 * of the step -- and conv1's slice (0.85 of 2.9 MB) right behind that kernel.  dta_get_option("exchanged") tells whether the
 * last backward did so (the caller then must not reduce again).  world <= 1 or peer_buffers == NULL unregisters.
 * (Measured slower than the single dta_grad_allreduce launch after the backward pass on 2 and 8 B200s -- the spinning
 * exchange CTAs compete with the tensor kernel they hide under -- so the Python side only registers it on request.)
 */
int dta_grad_allreduce_sizes(size_t n_float4, size_t n_double, int world, size_t* buffer_bytes, size_t* flags_offset,
                             size_t* scratch_bytes);
int dta_grad_allreduce(dta_ctx* ctx, int rank, int world, void* const peer_buffers[], const void* multicast_buffer,
                       size_t n_float4, size_t n_double, void* scratch, void* sync_words, void* cuda_stream);
int dta_set_grad_exchange(dta_ctx* ctx, int rank, int world, void* const peer_buffers[], const void* multicast_buffer,
                          size_t n_float4, size_t n_double, void* sync_words);

/*
 * Year ensemble on the device.  learned_ensemble.forward (src/models/year.py:24-33) skips a year whose crop tensor sums to
 * zero (`if x.sum() == 0: continue`, a device->host sync per year) and averages the last-head scores of the others.
 *   dta_crops_nonzero : flags[y] = (sum of crops[y] != 0) ? 1 : 0 for n_years <= 16 tensors of `elems` floats each
 *                       (workspace: n_years * 4096 bytes) -- the same test, without leaving the device
 *   dta_ensemble_mean : out[b][k] = mean over the years with flags[y] != 0 (flags NULL: all) of scores[y][b][k], followed by
 *                       F.softmax(dim=1) when softmax != 0 (MultiStage.validation_step / predict_step,
 *                       multi_stage.py:302,314).  No active year gives NaN rows (the reference raises on the empty stack).
 */
/* Device-side skip for the year ensemble in TRAINING: the reference does not run a year network whose crops are all zero
 * (year.py:27), so its BatchNorm buffers stay untouched.  With a gate registered (one float in device memory, e.g. one
 * element of dta_crops_nonzero's flags), training-mode dta_forward calls still compute the network but update
 * running_mean / running_var / num_batches_tracked only if *gate != 0 when the kernel runs -- the caller masks the year out
 * of the ensemble mean (and with it every gradient) with the same flag, and no value ever travels to the host.
 * gate == NULL (default) restores the unconditional update. */
int dta_set_update_gate(dta_ctx* ctx, const float* gate);
int dta_crops_nonzero(dta_ctx* ctx, int n_years, const float* const crops[], size_t elems, float* flags, void* workspace,
                      void* cuda_stream);
int dta_ensemble_mean(dta_ctx* ctx, int n_years, const float* const scores[], const float* flags, int batch, int classes,
                      int softmax, float* out, void* cuda_stream);

/*
 * Site metadata branch and late fusion (BASELINE config 5).  Replaces metadata.forward and the layers after the sensor
 * model in metadata_sensor_fusion.forward (src/models/metadata.py:17-24, 37-44):
 *   meta  = relu(Linear(16, classes)(Dropout(0.7)(BatchNorm1d(16)(Embedding(sites, 16)(site)))))
 *   out   = relu(Linear(2*classes, classes)(cat([meta, sensor], 1)))        (sensor = Hang2020 joint scores, dta_forward)
 * With sensor == NULL the call is the stand-alone `metadata` module (out = meta; fc_w / fc_b unused).
 *   site      : (batch) int64 site index, device memory; values outside [0, sites) are clamped (nn.Embedding raises)
 *   params    : device pointers of the state_dict tensors; training != 0 updates bn_rm / bn_rv / bn_nbt
 *   keep_mask : (batch, 16) uint8, non-zero = element kept by the dropout, or NULL: the mask is drawn from a counter-based
 *               generator keyed by `seed`.  Ignored in eval mode (dropout is the identity)
 *   saved     : saved_bytes; backward needs it together with the same site / sensor and the forward's `out`
 * Backward: `grads` entries may be NULL (skipped); dsensor (batch, classes) is the gradient to hand to dta_backward as
 * `djoint` (NULL when sensor was NULL).  All reductions run in fixed order.
 */
typedef struct dta_metadata_tensors {
  float* embedding;  /* (sites, 16)                      metadata_model.embedding.weight */
  float* bn_w;       /* (16)                             metadata_model.batch_norm.weight */
  float* bn_b;
  float* bn_rm;      /* running_mean / running_var / num_batches_tracked (ignored in a gradient table) */
  float* bn_rv;
  int64_t* bn_nbt;
  float* mlp_w;      /* (classes, 16)                    metadata_model.mlp.weight */
  float* mlp_b;
  float* fc_w;       /* (classes, 2*classes)             fc1.weight (fusion only) */
  float* fc_b;
} dta_metadata_tensors;
int dta_metadata_sizes(int batch, int classes, size_t* saved_bytes, size_t* workspace_bytes);
int dta_metadata_forward(dta_ctx* ctx, int batch, int sites, int classes, int training, const int64_t* site, const float* sensor,
                         const dta_metadata_tensors* params, const uint8_t* keep_mask, uint64_t seed, float* out, void* saved,
                         void* cuda_stream);
int dta_metadata_backward(dta_ctx* ctx, int batch, int sites, int classes, int training, const int64_t* site, const float* sensor,
                          const dta_metadata_tensors* params, const void* saved, const float* out, const float* dout,
                          const dta_metadata_tensors* grads, float* dsensor, void* workspace, void* cuda_stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Stand-alone building blocks.  The reference's tests and notebooks call conv_module, spatial_attention,
 * spectral_attention and Classifier on their own (tests/test_Hang2020.py:8-33); inside the networks they run fused
 * (dta_forward).  These entry points cover ANY plane size / channel count with exact-fp32 CUDA-core kernels.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dta_plane {
  int32_t batch, channels, height, width; /* a (batch, channels, height, width) float32 NCHW tensor */
} dta_plane;

typedef enum dta_attention_kind {
  DTA_ATTN_SPECTRAL = 1, /* Hang2020.spectral_attention (Hang2020.py:126-168) */
  DTA_ATTN_SPATIAL = 2   /* Hang2020.spatial_attention  (Hang2020.py:68-124)  */
} dta_attention_kind;

/* global_spectral_pool (Hang2020.py:7-12): out[row] = mean over the hw positions of row `row` of `in`; rows = batch * channels.
 * Backward: din[row][p] = dout[row] / hw. */
int dta_plane_mean(dta_ctx* ctx, const float* in, size_t rows, int hw, float* out, void* cuda_stream);
int dta_plane_mean_backward(dta_ctx* ctx, const float* dout, size_t rows, int hw, float* din, void* cuda_stream);

/*
 * conv_module.forward(x, pool) (Hang2020.py:24-31): Conv2d 3x3 "same" -> BatchNorm2d -> ReLU -> optional MaxPool2d.
 *   in      : shape of x; `filters` output channels; pool_h = pool_w = 1 means no pooling (else kernel = stride, floor)
 *   z       : (batch, filters, H, W)   convolution output, kept for backward
 *   stat    : 2 * filters floats       BatchNorm mean | 1/sqrt(var + eps) used by this call, kept for backward
 *   out     : (batch, filters, H / pool_h, W / pool_w)
 * training != 0 uses batch statistics and updates params->bn_rm / bn_rv / bn_nbt like nn.BatchNorm2d.
 * Backward: grads->conv_w, conv_b, bn_w, bn_b are overwritten (NULL entries skipped); dx may be NULL;
 * workspace: dta_conv_module_workspace_bytes().
 */
int dta_conv_module_workspace_bytes(const dta_plane* in, int filters, size_t* out);
int dta_conv_module_forward(dta_ctx* ctx, const dta_plane* in, int filters, int pool_h, int pool_w, int training, const float* x,
                            const dta_conv_block* params, float* z, float* stat, float* out, void* cuda_stream);
int dta_conv_module_backward(dta_ctx* ctx, const dta_plane* in, int filters, int pool_h, int pool_w, int training, const float* x,
                             const dta_conv_block* params, const float* z, const float* stat, const float* dout,
                             const dta_conv_block* grads, float* dx, void* workspace, void* cuda_stream);

/*
 * spectral_attention.forward / spatial_attention.forward (Hang2020.py:146-168 / 103-124): returns the gated map
 * `out` (same shape as x) and the pooled head features `feat` (batch, feat_per_crop).  channels must be 32, 64 or 128
 * (kernel sizes 3/5/7 resp. 7/5/3, class pool 4/2/1; DTA_ERR_UNSUPPORTED otherwise, the reference raises too).
 *   saved : saved_floats_per_crop * batch floats kept for backward
 * Backward: dout / dfeat are the upstream gradients of the two outputs (either may be NULL); dx may be NULL;
 * grads uses the dta_attention layout of `params` (NULL entries skipped; Conv1d off-centre taps are written as exact zeros).
 */
int dta_attention_sizes(int kind, const dta_plane* in, size_t* feat_per_crop, size_t* saved_floats_per_crop, size_t* workspace_bytes);
int dta_attention_forward(dta_ctx* ctx, int kind, const dta_plane* in, const float* x, const dta_attention* params, float* out,
                          float* feat, float* saved, void* cuda_stream);
int dta_attention_backward(dta_ctx* ctx, int kind, const dta_plane* in, const float* x, const dta_attention* params, const float* saved,
                           const float* dout, const float* dfeat, float* dx, const dta_attention* grads, void* workspace,
                           void* cuda_stream);

/* Classifier.forward (Hang2020.py:63-66): scores = feat W^T + b.  Backward: dfeat (may be NULL), dw, db (may be NULL). */
int dta_classifier_forward(dta_ctx* ctx, int batch, int in_features, int classes, const float* feat, const float* w, const float* b,
                           float* scores, void* cuda_stream);
int dta_classifier_backward(dta_ctx* ctx, int batch, int in_features, int classes, const float* feat, const float* w,
                            const float* dscores, float* dfeat, float* dw, float* db, void* cuda_stream);

/*
 * Fused Adam step over a table of parameter tensors: ONE kernel launch for the whole model.  Replaces the optimizer the reference
 * configures around the path, torch.optim.Adam(self.model.parameters(), lr=config["lr"]) (src/main.py:135-136,
 * src/models/multi_stage.py:258-262): betas / eps / weight_decay as in torch (L2 form), no amsgrad.
 *   params[i], grads[i] : n_tensors float32 tensors of numel[i] elements (n_tensors <= 96); tensors whose gradient is None are
 *                         simply left out of the table, as torch skips them
 *   exp_avg, exp_avg_sq : flat float32 moment buffers; tensor i occupies [offset[i], offset[i] + numel[i])
 *   param64 / grad64 / moments64 : the one float64 parameter of the path (Hang2020.alpha) with its 2 moments, or NULL
 *   step         : 1-based step number (host value) -- or step_device != NULL: an int64 device counter that the call
 *                  increments first and then reads, and lr_device (may be NULL) a float32 device scalar overriding h->lr, so
 *                  that a captured CUDA graph replays correct bias corrections and a scheduler can change lr without re-capture
 */
typedef struct dta_adam_hyper {
  double lr, beta1, beta2, eps, weight_decay;
  int64_t step;
} dta_adam_hyper;
int dta_adam_step(dta_ctx* ctx, int n_tensors, float* const params[], const float* const grads[], const int64_t numel[],
                  const int64_t offset[], float* exp_avg, float* exp_avg_sq, double* param64, const double* grad64,
                  double* moments64, const dta_adam_hyper* h, int64_t* step_device, const float* lr_device, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* DTA_B200_H_ */
