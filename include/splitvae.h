/* splitvae.h — C-ABI of libsplitvae.so: the SPLIT-VAE / SPLIT-GMVAE train step on B200 (sm_100a).
 *
 * The reference (51616/split-vae) has no FFI: its boundary is the Python surface of
 * vae/model.py + vae/trainer.py.  Each entry point below states which reference lines it
 * replaces.  The Python host (the modules under splitvae_b200/) binds these with ctypes and re-exposes the
 * reference's classes/functions (LGVae, LGGMVae, kl_divergence, ...) on top.
 *
 * Conventions
 *   - every call returns sv_status (0 = OK); text via sv_last_error(); no exceptions cross.
 *   - all pointers named *_dev are CUDA device pointers owned by the CALLER (torch tensors in
 *     the Python host) and must stay alive until the stream work completes.
 *   - `stream` is a cudaStream_t passed as void*; all compute calls are asynchronous on it and
 *     are CUDA-graph capturable (no allocation, no synchronisation inside).
 *   - one handle per GPU; a handle is not thread-safe; handles are independent.
 *   - there is NO CPU fallback: sv_create fails on a device that is not compute capability 10.x.
 *   - tensors are NHWC, float32 at the boundary, exactly as the reference's TF tensors.
 */
#ifndef SPLITVAE_H_
#define SPLITVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t sv_status;
enum {
  SV_OK = 0,
  SV_ERR_INVALID = 1,      /* bad argument / unsupported configuration */
  SV_ERR_DEVICE = 2,       /* not an sm_100 device, or CUDA runtime error */
  SV_ERR_STATE = 3,        /* call made before sv_bind, or buffers too small */
  SV_ERR_NOT_IMPLEMENTED = 4
};

enum { SV_MODEL_LGVAE = 0, SV_MODEL_LGGMVAE = 1,        /* vae/main.py:63-69  --model lgvae | lggmvae */
       SV_MODEL_GMVAE = 2 };                            /* vae/main.py:70-73  --model gmvae: gm encoder + ONE decoder fed by z_x (vae/model.py:277-299);
                                                           outputs / scalars of the x_hat path are not available (sv_output_ptr fails for them) */
enum { SV_PRECISION_BF16_TC = 0,   /* single-bf16 operands on tcgen05 tensor cores, fp32 accumulate: fastest; gradients 5-10 % from
                                      fp32 (ReLU masks of near-zero units flip under 2^-9 forward roundings) */
       SV_PRECISION_FP32_REF = 1,  /* fp32 SIMT reference kernels (debug / tight parity) */
       SV_PRECISION_BF16X3 = 2     /* the parity mode: forward products on bf16 PAIRS (hi + lo, three tcgen05 MMAs hi*hi + lo*hi +
                                      hi*lo, fp32 accumulate), backward on single bf16: gradients < 1e-2, KL / ELBO < 1e-3 from fp64 */ };

/* Mirrors the reference's CLI/config (vae/main.py:16-31) + model ctor args (vae/model.py:175,222). */
typedef struct sv_config {
  int32_t model;                /* SV_MODEL_*                                  --model        */
  int32_t height, width;        /* image_shape[1], image_shape[2]              --dataset      */
  int32_t batch;                /* per-GPU batch                               --batch_size   */
  int32_t global_latent_dims;   /* 128 (only value supported)                  --global_latent_dims */
  int32_t local_latent_dims;    /* 128 (only value supported)                  --local_latent_dims  */
  int32_t y_size;               /* <= 32                                       --y_size       */
  float tau;                    /*                                             --tau          */
  float beta;                   /*                                             --beta         */
  float alpha;                  /*                                             --alpha        */
  float learning_rate;          /*                                             --learning_rate*/
  int32_t world_size;           /* data-parallel ranks; gradients are scaled 1/(batch*world_size) */
  int32_t precision;            /* SV_PRECISION_*                                              */
  int32_t flags;                /* reserved, 0                                                 */
} sv_config;

typedef struct sv_handle sv_handle;

/* Per-variable description, Keras layout (conv HWIO, dense [in,out], bias [out]); order =
 * model.trainable_variables of the reference (vae/model.py:36-42,49-76,152-156,182-186,230-234). */
typedef struct sv_param_desc {
  char name[64];                /* e.g. "encoder_x.e1.kernel" */
  int32_t ndim;
  int32_t shape[4];
  int64_t offset;               /* in floats, into the parameter / gradient / Adam arenas */
  int64_t count;                /* number of floats */
} sv_param_desc;

/* Named device-side results of a forward / loss pass (the reference's output tuple,
 * vae/model.py:200,248 and the scalars of vae/trainer.py:125-135,151-164). */
enum {
  SV_OUT_DEC_X = 0,        /* [B,H,W,6] fp32: channels 0-2 x_mean, 3-5 x_log_scale (model.py:169) */
  SV_OUT_DEC_X_HAT = 1,    /* [B,H,W,6] fp32 */
  SV_OUT_Z_X = 2,          /* [B,128] */
  SV_OUT_Z_MEAN_X = 3,
  SV_OUT_Z_SIG_X = 4,
  SV_OUT_Z_X_HAT = 5,
  SV_OUT_Z_MEAN_X_HAT = 6,
  SV_OUT_Z_SIG_X_HAT = 7,
  SV_OUT_Y = 8,            /* [B,32] (first y_size columns valid), lggmvae only */
  SV_OUT_Y_LOGITS = 9,     /* [B,32] */
  SV_OUT_Z_PRIOR_MEAN = 10,
  SV_OUT_Z_PRIOR_SIG = 11,
  SV_OUT_SCALARS = 12,     /* float[8]: see SV_SCALAR_* */
  SV_OUT_SCALAR_SUMS = 13, /* float[16]: running sums of SV_SCALAR_* [0..5] over every sv_loss_fwd_bwd since the caller last zeroed them,
                              [8] = number of passes (the reference's Keras Mean metrics, vae/trainer.py:140-144); caller may memset */
  SV_OUT_COUNT = 14
};
enum { SV_SCALAR_RECON_X = 0, SV_SCALAR_RECON_X_HAT = 1, SV_SCALAR_KL_X = 2, SV_SCALAR_KL_X_HAT = 3,
       SV_SCALAR_TOTAL_KL = 4, /* lgvae: beta*(kl_x+kl_x_hat) (trainer.py:130); lggmvae: y_kl (trainer.py:161) */
       SV_SCALAR_TOTAL = 5, SV_SCALAR_COUNT = 8 };

/* ---- lifetime ------------------------------------------------------------------------- */
/* Replaces model construction, vae/main.py:63-74 (+ Adam / ExponentialDecay selection 65-68). */
sv_status sv_create(const sv_config* cfg, sv_handle** out);
sv_status sv_destroy(sv_handle* h);
const char* sv_last_error(const sv_handle* h);   /* h may be NULL: last sv_create error */
const char* sv_version(void);

/* ---- parameter inventory (model.trainable_variables) ----------------------------------- */
int32_t sv_param_count(const sv_handle* h);
sv_status sv_param_describe(const sv_handle* h, int32_t index, sv_param_desc* out);
int64_t sv_arena_floats(const sv_handle* h);      /* floats in each of params/grads/adam_m/adam_v */
int64_t sv_workspace_bytes(const sv_handle* h);   /* activations, packed weights, partials */

/* Caller-allocated device buffers.  params/grads/adam_m/adam_v: sv_arena_floats() floats each
 * (256-byte aligned); workspace: sv_workspace_bytes() bytes (1024-byte aligned), zero-filled. */
sv_status sv_bind(sv_handle* h, float* params_dev, float* grads_dev, float* adam_m_dev,
                  float* adam_v_dev, void* workspace_dev, int64_t workspace_bytes);

/* Must be called (on `stream`) after the caller writes new values into params_dev from outside
 * (initialisation, checkpoint load): refreshes the tensor-core operand copies of the weights. */
sv_status sv_params_updated(sv_handle* h, void* stream);

/* ---- the hot path ---------------------------------------------------------------------- */
/* LGVae.call / LGGMVae.call (vae/model.py:189-200, 237-248).  inputs_dev: [B,H,W,6] fp32 in
 * [-1,1] (x in channels 0-2, x_hat in 3-5).  eps_g/eps_l: [B,128] N(0,1) noise of the two
 * Sampling layers (model.py:12); u_dev: [B,y_size] U(0,1) gumbel noise (model.py:122), NULL for
 * lgvae.  Passing NULL for a noise pointer draws it in-kernel (Philox, device-side counter). */
sv_status sv_forward(sv_handle* h, const float* inputs_dev, const float* eps_g_dev,
                     const float* eps_l_dev, const float* u_dev, void* stream);

/* Loss block of train_step_lg_vae / train_step_lg_gm_vae (vae/trainer.py:125-135, 151-164) on
 * the outputs of the last sv_forward, fused with its own backward: writes SV_OUT_SCALARS and
 * the gradients w.r.t. the decoder outputs and the latent heads. */
sv_status sv_loss_fwd_bwd(sv_handle* h, const float* inputs_dev, void* stream);

/* tape.gradient (vae/trainer.py:137,166), split into segments so the host can overlap a
 * gradient all-reduce with the remaining backward work.  Segment s finishes the gradients of
 * the arena ranges returned by sv_segment_range. Segments must run in order 0..n-1. */
int32_t sv_num_segments(const sv_handle* h);
/* Segment s owns sv_segment_num_ranges(h, s) contiguous arena ranges (its layers' variables; one or two: the two encoders are not
 * adjacent in the Keras variable order).  Segments in backward order: 0 = the decoders, 1 = the encoders' last layers (dense heads and
 * the last conv), 2 = the encoders' first convs. */
int32_t sv_segment_num_ranges(const sv_handle* h, int32_t segment);
sv_status sv_segment_range(const sv_handle* h, int32_t segment, int32_t range_index, int64_t* offset_floats, int64_t* count_floats);
sv_status sv_backward_segment(sv_handle* h, int32_t segment, void* stream);
/* The same work, but the segment's gradients are final on `done_stream` (the stream the host reduces / updates them on) instead of
 * `stream`: the library's internal weight- and bias-gradient streams are joined into done_stream, so the backward chain of segment
 * s+1 issued on `stream` starts while segment s's weight gradients are still running.  `stream` only carries the chain itself
 * (activation gradients).  done_stream must already be ordered behind `stream` (and, under stream capture, part of the capture);
 * done_stream == stream or NULL is sv_backward_segment.  The LAST segment must be issued with sv_backward_segment. */
sv_status sv_backward_segment_deferred(sv_handle* h, int32_t segment, void* stream, void* done_stream);

/* optimizer.apply_gradients (vae/trainer.py:138,167): multi-tensor Keras-Adam over the arena,
 * step counter and (lggmvae) staircase LR schedule kept on the device (vae/main.py:65-68). */
sv_status sv_adam_step(sv_handle* h, void* stream);
/* The same update restricted to one backward segment's arena ranges (segment 0 also advances the step counter / bias-corrected
 * step size, so segments must be applied in order 0..n-1, each exactly once per step).  Lets the host update segment k
 * (its gradients are final - and all-reduced - first) while the backward pass of segment k+1 is still running.
 * sv_adam_step == sv_adam_segment(0) ; ... ; sv_adam_segment(n-1). */
sv_status sv_adam_segment(sv_handle* h, int32_t segment, void* stream);

/* Data parallel over NVLS (NVLink 5 / NVSwitch multicast): the gradient all-reduce FUSED with the optimizer.  mc_grads_dev / mc_params_dev
 * are the MULTICAST addresses of symmetric allocations that back every rank's gradient / parameter arena (the arenas given to sv_bind
 * must be those allocations).  One kernel per arena range of the segment: multimem.ld_reduce sums the ranks' gradients in the switch
 * (reduce-scatter), this rank applies Keras Adam to its 1/world shard (its m / v shard only), multimem.st writes the new weights into
 * every rank's arena (all-gather).  The caller provides the cross-rank ordering: a barrier after the segment's backward pass and
 * before this call, another after it and before sv_repack_segment (the operand re-pack of the segment's layers) or any other read of
 * the parameters.  Segment 0 advances the step counter; segments in order 0..n-1, once per step, on every rank. */
sv_status sv_nvls_adam_segment(sv_handle* h, int32_t segment, const float* mc_grads_dev, float* mc_params_dev, int32_t rank,
                               int32_t world, int32_t write_reduced_grads, void* stream);
sv_status sv_repack_segment(sv_handle* h, int32_t segment, void* stream);

/* Whole train_step_* (vae/trainer.py:120-144 / 146-173) = forward + loss + all segments + Adam. */
sv_status sv_train_step(sv_handle* h, const float* inputs_dev, const float* eps_g_dev,
                        const float* eps_l_dev, const float* u_dev, void* stream);

/* The whole train step as a CUDA graph owned by the handle, for hosts without a graph API of their own (the reference's step is
 * one @tf.function graph execution, vae/trainer.py:120,146).  sv_capture_graph records ONE sv_train_step issued on `stream` (a
 * non-default stream; nothing executes) with the given device pointers baked in: keep those buffers alive and refill them in
 * place between replays; NULL noise pointers keep the in-kernel Philox draws, which advance on the device at every replay.
 * sv_replay launches the captured step n_steps times on `stream`.  Re-capturing replaces the previous graph. */
sv_status sv_capture_graph(sv_handle* h, const float* inputs_dev, const float* eps_g_dev, const float* eps_l_dev,
                           const float* u_dev, void* stream);
sv_status sv_replay(sv_handle* h, int32_t n_steps, void* stream);

/* Device pointer of a named result (valid after sv_bind; contents after the producing call). */
sv_status sv_output_ptr(const sv_handle* h, int32_t which, void** dev_ptr, int64_t* count_floats);

/* LGVae.decode / LGGMVae.decode (vae/model.py:211-218, 259-266) without rescale: runs both
 * decoders on caller latents ([B,128] each); results land in SV_OUT_DEC_X / SV_OUT_DEC_X_HAT. */
sv_status sv_decode(sv_handle* h, const float* z_x_dev, const float* z_x_hat_dev, void* stream);

/* Encoder.encode_y (vae/model.py:137-140): prior heads on y [B,y_size] -> SV_OUT_Z_PRIOR_*. */
sv_status sv_encode_y(sv_handle* h, const float* y_dev, void* stream);

/* Optimizer iteration counter (optimizer.iterations), device-resident.  get / set are ordered on `stream` and return once it has
 * drained (they synchronise that stream, nothing else). */
sv_status sv_get_iterations(sv_handle* h, int64_t* iterations, void* stream);
sv_status sv_set_iterations(sv_handle* h, int64_t iterations, void* stream);

/* Number of kernels this library launched since creation (for bench.py's gpu_launches). */
int64_t sv_launch_count(const sv_handle* h);

/* ---- stand-alone operators (used by the parity tests and by splitvae_b200.trainer) ------ */
/* discretised_logistic_loss (vae/trainer.py:21-38): elementwise NLL over n elements. */
sv_status sv_discretised_logistic_loss(const float* x_dev, const float* mean_dev, const float* log_scale_dev,
                                       float* out_dev, int64_t n, void* stream);
/* Keras Adam on a flat range (TF ResourceApplyAdam arithmetic); alpha = lr*sqrt(1-b2^t)/(1-b1^t). */
sv_status sv_adam_flat(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int64_t n,
                       float alpha, void* stream);
/* scramble (augmentation.py:43-57) + uint8 -> [-1,1] scaling (vae/data.py:52) on device:
 * u8_dev [B,H,W,3], perm_dev [B,(H/p)*(W/p)] int32 -> inputs_dev [B,H,W,6] fp32. */
sv_status sv_stage_scramble(const uint8_t* u8_dev, const int32_t* perm_dev, float* inputs_dev,
                            int32_t batch, int32_t height, int32_t width, int32_t patch, void* stream);

/* CelebA preprocessing (vae/data.py:82-87: decode_jpeg -> resize_with_crop_or_pad(178,178) -> tf.image.resize([64,64]) -> /255*2-1)
 * fused with the scramble: u8_dev [B,Hs,Ws,3] decoded images, centre crop [crop_y, crop_y+crop_h) x [crop_x, crop_x+crop_w), bilinear
 * resize (half-pixel centres, no antialiasing) to height x width, perm_dev as above -> inputs_dev [B,height,width,6] fp32. */
sv_status sv_stage_resize_scramble(const uint8_t* u8_dev, const int32_t* perm_dev, float* inputs_dev, int32_t batch, int32_t src_height,
                                   int32_t src_width, int32_t crop_y, int32_t crop_x, int32_t crop_h, int32_t crop_w, int32_t height,
                                   int32_t width, int32_t patch, void* stream);
/* tf.random.shuffle of the patches (augmentation.py:49) on the device: perm_dev [B, n_patch] int32, one uniform permutation per image
 * (n_patch <= 4096), Philox stream (seed, step). */
sv_status sv_draw_permutations(int32_t* perm_dev, int32_t batch, int32_t n_patch, uint64_t seed, uint64_t step, void* stream);

/* ---- test hooks: run ONE layer of the plan with the reference (SIMT) or the tensor-core kernel on the
 * handle's own buffers, so the parity tests can check each tcgen05 kernel against its reference. --------- */
typedef struct sv_layer_info {
  char name[64];
  int32_t kh, kw, stride, Hi, Wi, Ci, Ho, Wo, Co;
  int32_t in_ld, in_coff, out_ld, dout_ld, din_ld;
  int32_t in_dt, out_dt, act_dt;          /* 0 = fp32, 1 = bf16 */
  int32_t has_dgrad, tc_fwd, tc_dgrad, tc_wgrad;
  void *in, *out, *dout, *din;            /* device pointers; in == NULL: the layer reads the caller's inputs */
  int64_t in_elems, out_elems, dout_elems, din_elems;
  /* which tensor-core kernel serves each pass (SV_KERN_*; 0 = reference SIMT kernel) */
  int32_t kern_fwd, kern_dgrad, kern_wgrad;
  int32_t split_fwd;                      /* 1: the forward product multiplies bf16 pairs (SV_PRECISION_BF16X3) */
  void *in_lo, *out_lo;                   /* lo planes of in / out (value = hi + lo) in the bf16x3 mode, else NULL */
  int32_t wgrad_ctas;                     /* CTAs of the weight-gradient launch (the halo kernel runs on a fixed few, one per SM, beside the dgrad chain) */
  int32_t reserved;
} sv_layer_info;
enum { SV_KERN_NONE = 0, SV_KERN_IGEMM = 1, SV_KERN_HALO_CONV = 2, SV_KERN_NSCONV = 3, SV_KERN_WGRAD = 4, SV_KERN_HALO_WGRAD = 5, SV_KERN_PCONV = 6 };
enum { SV_PASS_FWD = 0, SV_PASS_DGRAD = 1, SV_PASS_WGRAD = 2 };
enum { SV_IMPL_REF = 0, SV_IMPL_TC = 1 };
int32_t sv_debug_layer_count(const sv_handle* h);
sv_status sv_debug_layer_info(const sv_handle* h, int32_t index, sv_layer_info* out);
sv_status sv_debug_run_layer(sv_handle* h, int32_t index, int32_t pass, int32_t impl, const float* inputs_dev, void* stream);
/* The pixel likelihood kernel of sv_loss_fwd_bwd ALONE (no scalar reduction behind it): what bench.py times for the kernel's HBM roofline. */
sv_status sv_debug_pixel_loss(sv_handle* h, const float* inputs_dev, void* stream);
/* With SV_HALO_TRACE=1 in the environment at sv_create, every CTA of the halo convolution kernel records the SM clock at its
 * phase boundaries; this copies 8 words per CTA of the LAST such launch to host memory (synchronises). Returns #CTAs or -1. */
int32_t sv_debug_halo_trace(uint64_t* out_host, int32_t max_ctas);

#ifdef __cplusplus
}
#endif
#endif /* SPLITVAE_H_ */
