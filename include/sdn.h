/* sdn.h - C ABI of libsdn_b200.so: the B200-native (sm_100a) stereo U-Net step.
 *
 * The reference (sdfgeoff/stereo_depth_estimation) is pure Python and has no
 * FFI; the boundary it exposes for this hot path is the Python surface of
 *   src/foundation_stereo_depth/model.py:48-104   (StereoUNet.forward)
 *   src/foundation_stereo_depth/train.py:320-357  (loss + metric sums in run_epoch)
 *   src/foundation_stereo_depth/dataset.py:23-30,184-270,302-311 (sample pipeline)
 *   src/live_camera/depth_live_dl.py:516-529      (single-pair inference)
 * Every entry point below names the reference interface it replaces.  The
 * Python mirror (stereo_depth_estimation_b200/, ctypes) is the host side.
 *
 * Conventions: plain pointers and sizes only (no torch / C++ types); all
 * pointers are DEVICE pointers unless the name says host; every function
 * returns 0 on success and a negative code on failure, with a thread-local
 * message available from sdn_last_error().  Work is enqueued on the caller's
 * CUDA stream (passed as void*, a cudaStream_t) and is asynchronous.  The
 * library owns only its context/workspace; parameters, gradients, inputs and
 * outputs stay owned by the caller (PyTorch storages in practice).
 */
#ifndef SDN_H_
#define SDN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdn_ctx sdn_ctx;

#define SDN_NUM_PARAMS 66 /* len(list(StereoUNet().parameters())), model.py:59-77 */
#define SDN_NUM_BN 18     /* BatchNorm2d layers, model.py:37,40 (x9 blocks) */
#define SDN_NUM_STAGES 5  /* backward stages == gradient all-reduce buckets */

/* Per-view photometric augmentation parameters, the explicit form of the
 * reference samplers dataset.py:214-246 consumed by dataset.py:248-270. */
typedef struct sdn_aug_params {
    float brightness;    /* _sample_jitter_factor(brightness_jitter) */
    float contrast;      /* _sample_jitter_factor(contrast_jitter) */
    float saturation;    /* _sample_jitter_factor(saturation_jitter) */
    float hue;           /* _sample_hue_shift() */
    float gamma;         /* _sample_gamma_factor() */
    float blur_sigma;    /* > 0: gaussian_blur(k=5, sigma); 0: no blur (coin came up tails) */
    float noise_std;     /* _sample_noise_std() */
    uint32_t noise_seed; /* stream id of the device Philox generator */
} sdn_aug_params;

const char* sdn_last_error(void);
int sdn_version(void);

/* Context: workspace for batches up to max_batch at resolution H x W
 * (both multiples of 16, model.py:79-95 pooling/upsampling constraint). */
#define SDN_CTX_PREPROCESS_ONLY 1u /* flags: allocate only what sdn_preprocess needs */
int sdn_create(sdn_ctx** out, int device, int max_batch, int H, int W, unsigned flags);
int sdn_destroy(sdn_ctx* ctx);
/* Bytes of device workspace the context holds. */
int64_t sdn_workspace_bytes(const sdn_ctx* ctx);

/* Borrow the 66 fp32 parameter tensors in StereoUNet.parameters() order
 * (model.py:59-77), optional gradient destinations (same order, same layouts,
 * entries may be NULL), and the 18 BatchNorm running buffers in module order
 * (state_dict keys *.running_mean / *.running_var / *.num_batches_tracked). */
int sdn_set_params(sdn_ctx* ctx, const float* const* params, float* const* grads, float* const* bn_running_mean,
                   float* const* bn_running_var, int64_t* const* bn_num_batches_tracked);

/* StereoUNet.forward(x, return_uncertainty) (model.py:79-104).
 * x: fp32 NCHW [B,6,H,W]; disp / logvar: fp32 [B,1,H,W] (logvar may be NULL).
 * training != 0: BatchNorm uses batch statistics and updates the running
 * buffers, activations are kept for sdn_backward*.  params_dirty != 0: the fp32
 * parameters changed since the last call (repack the bf16 operand cache). */
int sdn_forward(sdn_ctx* ctx, const float* x, float* disp, float* logvar, int B, int training, int params_dirty,
                void* stream);

/* autograd of forward (what loss.backward() at train.py:342 reaches): seeds with
 * dL/d(disp), dL/d(logvar) (fp32 [B,1,H,W], g_logvar may be NULL) and runs the
 * head backward.  Then sdn_backward_stage(s) for s = 0..SDN_NUM_STAGES-1 in
 * order; when stage s has run, the gradients of the parameters in
 * sdn_stage_param_range(s) are final in the caller's grad tensors.
 * Stream semantics: all work is ordered on `stream`.  Internally the weight gradients of a stage run
 * on a context-owned low-priority side stream (forked after each layer's BatchNorm backward, joined
 * back into `stream` before sdn_backward_stage returns), so whatever the caller enqueues on `stream`
 * after the call - an all-reduce of the stage's gradient bucket, the optimizer - sees final values. */
int sdn_backward_begin(sdn_ctx* ctx, const float* g_disp, const float* g_logvar, int accumulate, void* stream);
int sdn_backward_stage(sdn_ctx* ctx, int stage, void* stream);
int sdn_stage_param_range(int stage, int* first_param, int* num_params);

/* Fused heteroscedastic Laplace loss (train.py:329-357) on the forward just run:
 * mask = valid_mask & isfinite(target); n = sum(mask) (device side);
 * sums[0..3] (device fp64, like the reference's Python-float running sums) += sum nll, sum |diff|,
 * sum diff^2, sum exp(0.5*logvar) over mask, *count += n; seeds the backward with dL/dz for L = sum(nll)/n_norm where
 * n_norm = *n_norm_dev (a device u64, e.g. the all-reduced global count) or n if
 * NULL.  with_backward == 0 gives the validation path (metrics only).
 * disp / logvar outputs are optional (NULL skips the store). */
int sdn_loss_begin(sdn_ctx* ctx, const float* target, const uint8_t* valid_mask, float* disp, float* logvar,
                   double* sums4, unsigned long long* count, const unsigned long long* n_norm_dev, int with_backward,
                   int accumulate, void* stream);
/* n = sum(valid_mask & isfinite(target)) into *count_out (device u64, pre-zeroed by the callee). */
int sdn_count_valid(sdn_ctx* ctx, const float* target, const uint8_t* valid_mask, int B, unsigned long long* count_out,
                    void* stream);

/* ---- one-call steps --------------------------------------------------------------------------------
 * sdn_train_step = the loop body of run_epoch for one assembled batch, train.py:325-342 (everything but
 * optimizer.step(), which is sdn_adamw_step): train-mode forward -> n = sum(valid_mask & isfinite(target))
 * (skipped with SDN_STEP_HAVE_COUNT: *n_norm already holds this rank's count, e.g. from sdn_preprocess) ->
 * [all-reduce of n: the loss normaliser is the GLOBAL valid count] -> fused loss, metric sums and head
 * backward (sdn_loss_begin semantics: sums4 / count accumulate) -> the SDN_NUM_STAGES backward stages; with a
 * communicator (sdn_comm_init) each stage's gradient bucket is all-reduced (sum) on the context's
 * communicator stream while the next stage runs, and `stream` waits for the last bucket before the call's
 * work ends.  Gradients go to the destinations given to sdn_set_params; for data parallelism the
 * destinations of each stage (sdn_stage_param_range) must be consecutive views of one buffer.
 * On return *n_norm (device) holds the global count: 0 means "skip the optimizer step" (train.py:331-332);
 * sdn_adamw_step takes it as its gate, so no host synchronisation is needed.
 * sdn_eval_step = run_epoch with optimizer=None (train.py:618) / log_epoch_previews (train.py:268-272):
 * eval-mode forward (running statistics) + the same metric sums; disp / logvar are optional outputs. */
#define SDN_STEP_HAVE_COUNT 1u
#define SDN_STEP_NO_OVERLAP 2u /* all-reduce on `stream` itself (bit-identical result; the dp_check leg of bench.py) */
int sdn_train_step(sdn_ctx* ctx, const float* x, const float* target, const uint8_t* valid_mask, int B,
                   int params_dirty, double* sums4, unsigned long long* count, unsigned long long* n_norm,
                   unsigned flags, void* stream);
int sdn_eval_step(sdn_ctx* ctx, const float* x, const float* target, const uint8_t* valid_mask, int B,
                  int params_dirty, float* disp, float* logvar, double* sums4, unsigned long long* count,
                  void* stream);

/* ---- data parallelism (new: the reference is single-process) -------------------------------
 * One process per GPU, one NCCL communicator per context.  Rank 0 calls sdn_comm_unique_id and hands the
 * 128 bytes to the other ranks by any means (torch.distributed store, MPI, a file); every rank then calls
 * sdn_comm_init (collective).  NCCL is resolved with dlopen("libnccl.so.2") at that moment - inside a
 * PyTorch process that is the copy torch already mapped - so single-GPU users never need it.
 * sdn_comm_allreduce: in-place sum over ranks on `stream` (metric sums, checks); no-op for world 1. */
#define SDN_F32 0
#define SDN_F64 1
#define SDN_U64 2
int sdn_comm_unique_id(void* out128_host);
int sdn_comm_init(sdn_ctx* ctx, const void* unique_id_128_host, int rank, int world);
int sdn_comm_destroy(sdn_ctx* ctx);
int sdn_comm_world(const sdn_ctx* ctx);
int sdn_comm_allreduce(sdn_ctx* ctx, void* buf, int64_t count, int dtype, void* stream);

/* FoundationStereoDataset.__getitem__ + default collate (dataset.py:184-212,
 * 248-270, 302-311) for a batch of raw uint8 HWC images already on the device:
 * left / right RGB and the RGB-encoded disparity, each [B,Hs,Ws,3].
 * Outputs: input fp32 [B,6,H,W], target fp32 [B,1,H,W], mask u8 [B,1,H,W],
 * *valid_count (device u64, optional) = sum(mask & isfinite(target)).
 * aug: 2*B parameter structs on the DEVICE (left, right per sample; pinned host
 * memory with SDN_PREPROCESS_AUG_HOST) or NULL for augment=False.
 * flags: SDN_RESIZE_FOURTERM selects the other of the two fused-multiply-add
 * orderings that torch's CPU bilinear kernel is compiled with (it is the one a
 * single-threaded DataLoader worker runs on small / 3-channel images; results
 * differ by <= 1 ulp, indices and weights are identical); 0 = canonical form. */
#define SDN_RESIZE_FOURTERM 1u
#define SDN_PREPROCESS_DIRECT 2u /* flags: force the un-staged kernel (tests compare both) */
/* flags: `aug` is PINNED HOST memory (cudaHostAlloc / torch pin_memory, device-readable under UVA).  A small
 * kernel stages it into the context, so the step issues no copy-engine transfer that would queue behind
 * the bulk host->device prefetch of the next batch.  The caller keeps the buffer unchanged until the
 * stream has passed this call. */
#define SDN_PREPROCESS_AUG_HOST 4u
int sdn_preprocess(sdn_ctx* ctx, const uint8_t* left, const uint8_t* right, const uint8_t* disparity, int B, int Hs,
                   int Ws, const sdn_aug_params* aug_dev, float* input, float* target, uint8_t* mask,
                   unsigned long long* valid_count, unsigned flags, void* stream);

/* The same sample on a CACHE HIT (dataset.py:86-106 load_cached_sample, then 302-311): the npz read-through
 * cache holds uint8 HWC views [B,H,W,3] and a FLOAT16 disparity [B,H,W] already at the context's H x W, so
 * there is no resize and no disparity rescale: views = u8 / 255, target = (float)f16, then the same
 * augmentation / cat / valid_mask / count as sdn_preprocess (aug, flags: as there). */
int sdn_preprocess_cached(sdn_ctx* ctx, const uint8_t* left, const uint8_t* right, const uint16_t* disparity_f16, int B,
                          const sdn_aug_params* aug_dev, float* input, float* target, uint8_t* mask,
                          unsigned long long* valid_count, unsigned flags, void* stream);

/* Live viewer, per frame (src/live_camera/depth_live_dl.py).
 * sdn_live_preprocess = preprocess_rgb(view_l), preprocess_rgb(view_r), cat (:225-229, 516-520): two BGR uint8
 * frames [Hs,Ws,3] on the device -> the model input [1,6,H,W] float32 (H x W = the context's); the uint8
 * INTER_LINEAR resize reproduces cv2.resize's fixed-point arithmetic bit-exactly.
 * sdn_live_postprocess (:531-538, 371-381): optional EMA of the disparity (ema_state: device float[n], kept by
 * the caller between frames; ema_has_state == 0 on the first frame; ema_alpha <= 0 or NULL state = off),
 * depth = focal_px * baseline_m / d where d is finite and > 1e-6 (NaN elsewhere), confidence =
 * exp(-0.5 * logvar).  disp_out / depth_out / conf_out are optional (NULL skips). */
int sdn_live_preprocess(sdn_ctx* ctx, const uint8_t* frame_left_bgr, const uint8_t* frame_right_bgr, int Hs, int Ws,
                        float* input, void* stream);
int sdn_live_postprocess(sdn_ctx* ctx, const float* disp, const float* logvar, int64_t n_pixels, float* ema_state,
                         int ema_has_state, double ema_alpha, double focal_px, double baseline_m, float* disp_out,
                         float* depth_out, float* conf_out, void* stream);

/* Test / debug access to the NHWC bf16 activations of the last forward:
 * which = conv layer index 0..17 (100..103: the ConvTranspose2d outputs); kind 0 = pre-BN conv output, 1 = post
 * BN+ReLU, 2 = gradient w.r.t. the pre-BN output (after backward), 3 = gradient w.r.t. the post-ReLU output,
 * 4 / 5 (pooled layers 1, 3, 5, 7) = gradient w.r.t. / value of the 2x2 max-pooled output.
 * Copies to a host fp32 buffer of B*H_l*W_l*C_l elements; returns the dims. */
int sdn_debug_read(sdn_ctx* ctx, int which, int kind, float* host_out, int64_t capacity, int* dims4);

/* torch.optim.AdamW.step() (train.py:343,578; lr 1e-3, betas (0.9, 0.999), eps 1e-8, decoupled weight
 * decay) over up to 66 fp32 tensors in one launch.  step_dev: device int64 counter of applied steps
 * (incremented here); gate_dev (nullable): device u64, 0 = skip the whole update (the reference skips the
 * batch when no pixel is valid, train.py:331-332) - no host synchronisation is needed for that rule. */
int sdn_adamw_step(sdn_ctx* ctx, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, int n, double lr, double beta1, double beta2,
                   double eps, double weight_decay, long long* step_dev, const unsigned long long* gate_dev,
                   void* stream);

/* Timing forensics (SDN_DEBUG_TRACE_LAYER=<conv layer>): clock64() stamps of the three
 * warp roles of CTA 0 for its first 16 tiles, [role][tile][event] as 384 int64. */
int sdn_debug_trace(sdn_ctx* ctx, long long* host_out);

/* Per-op device timing: when enabled every op is bracketed by CUDA events on the
 * launching stream.  sdn_profile_dump synchronises the device and writes a CSV
 * "name,layer,calls,total_ms,flops,bytes" (algorithmic flops / bytes of SURVEY 8d
 * per op) into buf, then clears the records. */
int sdn_profile_enable(sdn_ctx* ctx, int enable);
int sdn_profile_dump(sdn_ctx* ctx, char* buf, int64_t capacity);

/* Kernels launched by this context since creation (bench.py's gpu_launches). */
int64_t sdn_launch_count(const sdn_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SDN_H_ */
