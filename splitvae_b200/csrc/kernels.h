// Host-callable launchers of every kernel in libsplitvae (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sv {

// ---- simt_kernels.cu --------------------------------------------------------------------------
void ref_conv_fwd(const ConvGeom& g, const void* in, int in_dt, const float* params, void* out, int out_dt,
                  bool round_w, cudaStream_t s);
void ref_conv_dgrad(const ConvGeom& g, const void* dout, int dt, const float* params, void* din,
                    const void* mask_src, int mask_dt, int mask_act, bool round_w, cudaStream_t s);
void ref_conv_wgrad(const ConvGeom& g, const void* in, int in_dt, const void* dout, int dout_dt, float* grads, bool round_in,
                    cudaStream_t s);
int colsum_chunks(long long rows);
void bias_grad(const ConvGeom& g, const void* dout, int dt, float* partial_ws, float* grads, cudaStream_t s);
// multi-tensor bias gradients (bf16 dY): all layers of one backward branch in two launches
// ext_chunks > 0: the per-chunk column partials [ext_chunks][ncols] of this layer are written by the kernel that PRODUCES dY
// (upsample2x_bwd for the decoders' d2-d4, pixel_loss for d5: the values are still in registers there), the table only finishes them.
struct ColsumSpec { ConvGeom g; const void* dout; int ext_chunks; };
struct ColsumTable;
bool colsum_multi_supported(const ColsumSpec* specs, int n);
long long colsum_table_partial_floats(const ColsumSpec* specs, int n);
ColsumTable* colsum_table_create(const ColsumSpec* specs, int n, float* partial_ws, const char** err);
void colsum_table_destroy(ColsumTable* t);
int colsum_table_run(ColsumTable* t, float* grads, cudaStream_t s);   // returns #launches
float* colsum_table_ext_partial(ColsumTable* t, int spec_index);      // where the producer of spec i writes partial[chunk][colsum_ext_cols]
int colsum_ext_cols(int dout_ld);                                      // columns of one external partial row (the padded pitch)
int upsample2x_bwd_blocks(int B, int H, int W, int C);                 // grid of the bf16 upsample2x_bwd launch = its ext_chunks
void upsample2x_fwd(const void* in, void* out, int dt, int B, int H, int W, int C, cudaStream_t s);
// bf16x3: input and output are bf16 pairs (value = hi + lo); out_lo may be NULL (the consumer reads single bf16: d5)
void upsample2x_fwd_pair(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C, cudaStream_t s);
// colsum_partial (bf16 only, may be NULL): per-block column sums of din, [upsample2x_bwd_blocks][C] (the bias gradient of the producer)
void upsample2x_bwd(const void* dout, void* din, const void* mask_src, int mask_act, int dt, int B, int H, int W,
                    int C, cudaStream_t s, float* colsum_partial = nullptr);

// ---- fused_kernels.cu -------------------------------------------------------------------------
struct LatentBufs {
  // forward
  const float* heads_g;   // [B,256] (z_mean | z_sig) of encoder_x
  const float* heads_l;   // [B,256] of encoder_x_hat
  float* eps_g;           // [B,128] saved noise
  float* eps_l;
  float* z_g;             // [B,128] fp32 outputs
  float* z_l;
  float* zm_g; float* zs_g; float* zm_l; float* zs_l;   // contiguous copies for the output tuple
  void* zcat;             // [B,256] activation dtype: decoder_x input (z_g | z_l)
  bf16* zcat_lo;          // bf16x3: lo plane of zcat (NULL otherwise)
  // backward
  const void* dzcat;      // [B,256] activation dtype: dgrad of decoder_x.d1
  const void* dzl2;       // [B,128] activation dtype: dgrad of decoder_x_hat.d1
  void* dheads_g;         // [B,256] activation dtype, gradient w.r.t. pre-activation heads
  void* dheads_l;
  // gm prior (NULL for lgvae)
  const float* yheads;    // [B,768] (h_top | z_prior_mean | z_prior_sig), fp32
};

// also writes the per-block partial sums of the two Gaussian KLs to kl_partials[2 * reparam_blocks(B)]
int reparam_blocks(int B);
void reparam(const LatentBufs& L, int B, int act_dt, const float* user_eps_g, const float* user_eps_l,
             unsigned long long seed, const unsigned long long* counter_dev, float* kl_partials, cudaStream_t s);
void latent_bwd(const LatentBufs& L, int B, int act_dt, int gm, float beta, float inv_batch, cudaStream_t s);

// (y_act_lo / *_lo: lo planes of the bf16x3 mode, NULL otherwise)
void gumbel_fwd(const float* logits, const float* user_u, float* u_saved, float* y, void* y_act, void* y_act_lo, int act_dt, int B,
                int K, float tau, unsigned long long seed, const unsigned long long* counter_dev, cudaStream_t s);
// h = e1out + h_top  (vae/model.py:130)
void gm_add(const void* yb0e1_out, const void* yb0e1_lo, const float* yheads, void* hsum, void* hsum_lo, int act_dt, int B, cudaStream_t s);
// glue A: gradients entering the (y_block.0|e1) and (h_top|z_prior_mean|z_prior_sig) fused layers
void gm_glue_a(const void* dhsum, const void* yb0e1_out, const float* yheads, const float* zm_g, const float* zs_g,
               void* d_yb0e1, void* d_yheads, int act_dt, int B, float beta, float inv_batch, cudaStream_t s);
// glue B: gumbel-softmax backward + categorical-KL gradient -> d y_logits
void gm_glue_b(const void* dy, const float* y, const float* logits, void* dlogits, int act_dt, int B, int K, float tau,
               float alpha, float inv_batch, cudaStream_t s);

// fused reconstruction likelihood fwd+bwd for both decoders; partial sums -> loss_partials[2*nblocks]
int pixel_loss_blocks(long long npix);
void pixel_loss(const float* inputs, const float* dec_x, const float* dec_xh, void* dout_x, void* dout_xh, int dout_dt,
                int dout_ld, long long npix, float grad_scale, float* loss_partials, bool fast_math, cudaStream_t s,
                float* colsum_x = nullptr, float* colsum_xh = nullptr);   // [pixel_loss_blocks][16] bias-gradient partials of the two d5 layers
// final reduction of the KL partials (reparam) and pixel partials (pixel_loss) + categorical KL -> scalars[8]
void loss_scalars(const float* kl_partials, int kl_blocks, const float* y_logits, int B, int K, int gm, float beta, float alpha,
                  const float* loss_partials, int nblocks, float* scalars, cudaStream_t s);

void dll_elementwise(const float* x, const float* m, const float* ls, float* out, long long n, cudaStream_t s);

// Keras Adam.  state_dev: {iterations (u64), alpha (f32)}; adam_prepare advances it on the device.
struct AdamState { unsigned long long iterations; float alpha; float pad; };
void adam_prepare(AdamState* st, float lr, int staircase, cudaStream_t s);
void adam_apply(float* p, const float* g, float* m, float* v, long long n, const AdamState* st, float alpha_host,
                cudaStream_t s);

// fused reduce-scatter + Adam(shard) + all-gather over NVLS multicast addresses of the gradient / parameter arenas (see fused_kernels.cu)
void nvls_adam(const float* mc_grads, float* mc_params, const float* params, float* m, float* v, float* mc_grads_out, long long off,
               long long cnt, int rank, int world, const AdamState* st, cudaStream_t s);

void stage_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int B, int H, int W, int p, cudaStream_t s);
// CelebA: centre crop [cy, cy+ch) x [cx, cx+cw) of a [B,Hs,Ws,3] uint8 batch, bilinear resize to H x W, scaling and scramble in one pass
void stage_resize_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int B, int Hs, int Ws, int cy, int cx, int ch, int cw,
                           int H, int W, int p, cudaStream_t s);
// uniform patch permutations [B, n] (n <= 4096), Philox stream (seed, step)
void draw_permutations(int32_t* perm, int B, int n, unsigned long long seed, unsigned long long step, cudaStream_t s);

}  // namespace sv
