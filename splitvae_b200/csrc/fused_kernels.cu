// HBM-bound fused kernels of the train step: reparameterisation, the fused reconstruction
// log-likelihood forward+backward, KL scalars, latent/gumbel gradient glue, multi-tensor Keras
// Adam, and the on-device scramble staging.
//
// Reference semantics: Sampling (vae/model.py:9-13), gumbel-softmax (vae/model.py:122-123),
// kl_divergence / kl_divergence_two_gauss / discretised_logistic_loss (vae/trainer.py:11-38),
// loss assembly (vae/trainer.py:125-135, 151-164), tf.keras.optimizers.Adam (vae/main.py:65-68),
// Augmentator.scramble (augmentation.py:43-57).  Analytic backward: SURVEY.md section 9.2.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace sv {

// ------------------------------------------------------------------ Philox4x32-10 counter RNG
struct U4 { uint32_t x, y, z, w; };
__device__ __forceinline__ U4 philox4x32(U4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = U4{hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f); }
__device__ __forceinline__ float normal_from(uint32_t a, uint32_t b) {
  return sqrtf(-2.f * logf(u01(a))) * cospif(2.f * u01(b));
}

template <typename T>
__device__ __forceinline__ T ld_as(const void* p, long long i) { return ((const T*)p)[i]; }

// ------------------------------------------------------------------ reparameterisation
// Also emits the per-block partial sums of the two Gaussian KL terms (vae/trainer.py:11-18) into kl_partials[2*block]
// (fixed-order block reduction; loss_scalars adds the blocks up), so the scalar kernel does not re-read the latents.
template <typename T>
__global__ void __launch_bounds__(256) reparam_kernel(LatentBufs L, int B, const float* __restrict__ ueg, const float* __restrict__ uel,
                                                      unsigned long long seed, const unsigned long long* __restrict__ counter,
                                                      float* __restrict__ kl_partials) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  float kl_g = 0.f, kl_l = 0.f;
  if (idx < B * 128) {
  const int b = idx >> 7, d = idx & 127;
  float eg, el;
  if (ueg && uel) {
    eg = ueg[idx];
    el = uel[idx];
  } else {
    const unsigned long long step = *counter;
    const U4 r = philox4x32(U4{(uint32_t)idx, 0x5A17u, (uint32_t)step, (uint32_t)(step >> 32)}, (uint32_t)seed,
                            (uint32_t)(seed >> 32));
    eg = ueg ? ueg[idx] : normal_from(r.x, r.y);
    el = uel ? uel[idx] : normal_from(r.z, r.w);
  }
  const bool has_l = L.heads_l != nullptr;        // (plain GMVAE: no x_hat encoder; z_l = 0, sigma_l = 1 -> KL_l = 0)
  const float mg = L.heads_g[b * 256 + d], sg = L.heads_g[b * 256 + 128 + d];
  const float ml = has_l ? L.heads_l[b * 256 + d] : 0.f, sl = has_l ? L.heads_l[b * 256 + 128 + d] : 1.f;
  if (!has_l) el = 0.f;
  const float zg = mg + sg * eg, zl = ml + sl * el;  // vae/model.py:13
  L.eps_g[idx] = eg; L.eps_l[idx] = el;
  L.z_g[idx] = zg; L.z_l[idx] = zl;
  L.zm_g[idx] = mg; L.zs_g[idx] = sg; L.zm_l[idx] = ml; L.zs_l[idx] = sl;
  ((T*)L.zcat)[b * 256 + d] = from_f32<T>(zg);
  ((T*)L.zcat)[b * 256 + 128 + d] = from_f32<T>(zl);
  if (L.zcat_lo) {   // bf16x3: z as a bf16 pair
    L.zcat_lo[b * 256 + d] = __float2bfloat16_rn(zg - round_bf16(zg));
    L.zcat_lo[b * 256 + 128 + d] = __float2bfloat16_rn(zl - round_bf16(zl));
  }
  if (L.yheads) {   // q(z_g) || p(z_g | y), q(z_l) || N(0, I)   (vae/trainer.py:157-158)
    const float pm = L.yheads[b * 768 + 512 + d], ps = L.yheads[b * 768 + 640 + d];
    kl_g = logf(ps) - logf(sg) + (sg * sg + (mg - pm) * (mg - pm)) / (2.f * ps * ps) - 0.5f;
    kl_l = -logf(sl) + (sl * sl + ml * ml) * 0.5f - 0.5f;
  } else {          // vae/trainer.py:11-15
    kl_g = -0.5f * (1.f + logf(sg * sg) - mg * mg - sg * sg);
    kl_l = -0.5f * (1.f + logf(sl * sl) - ml * ml - sl * sl);
  }
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    kl_g += __shfl_xor_sync(0xffffffffu, kl_g, o);
    kl_l += __shfl_xor_sync(0xffffffffu, kl_l, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = kl_g; red[1][threadIdx.x >> 5] = kl_l; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += red[0][i]; c += red[1][i]; }
    kl_partials[2 * blockIdx.x] = a;
    kl_partials[2 * blockIdx.x + 1] = c;
  }
}

int reparam_blocks(int B) { return (B * 128 + 255) / 256; }

void reparam(const LatentBufs& L, int B, int act_dt, const float* ueg, const float* uel, unsigned long long seed,
             const unsigned long long* counter, float* kl_partials, cudaStream_t s) {
  const int nb = reparam_blocks(B);
  if (act_dt == DT_F32) launch_pdl(reparam_kernel<float>, dim3(nb), dim3(256), 0, s, L, B, ueg, uel, seed, counter, kl_partials);
  else launch_pdl(reparam_kernel<bf16>, dim3(nb), dim3(256), 0, s, L, B, ueg, uel, seed, counter, kl_partials);
}

// gradient w.r.t. the pre-activation encoder heads: reparam adjoint + KL gradient + softplus'
template <typename T>
__global__ void latent_bwd_kernel(LatentBufs L, int B, int gm, float beta, float inv_batch) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 128) return;
  const int b = idx >> 7, d = idx & 127;
  const float dzg = to_f32(ld_as<T>(L.dzcat, b * 256 + d));
  const bool has_l = L.dheads_l != nullptr;
  const float dzl = has_l ? to_f32(ld_as<T>(L.dzcat, b * 256 + 128 + d)) + to_f32(ld_as<T>(L.dzl2, idx)) : 0.f;
  const float mg = L.zm_g[idx], sg = L.zs_g[idx], ml = L.zm_l[idx], sl = L.zs_l[idx];
  const float k = beta * inv_batch;
  float dmg, dsg;
  if (gm) {
    const float pm = L.yheads[b * 768 + 512 + d], ps = L.yheads[b * 768 + 640 + d];
    const float ip2 = 1.f / (ps * ps);
    dmg = dzg + k * (mg - pm) * ip2;
    dsg = dzg * L.eps_g[idx] + k * (sg * ip2 - 1.f / sg);
  } else {
    dmg = dzg + k * mg;
    dsg = dzg * L.eps_g[idx] + k * (sg - 1.f / sg);
  }
  const float dml = dzl + k * ml;
  const float dsl = dzl * L.eps_l[idx] + k * (sl - 1.f / sl);
  ((T*)L.dheads_g)[b * 256 + d] = from_f32<T>(dmg);
  ((T*)L.dheads_g)[b * 256 + 128 + d] = from_f32<T>(dsg * (1.f - expf(-sg)));
  if (has_l) {
    ((T*)L.dheads_l)[b * 256 + d] = from_f32<T>(dml);
    ((T*)L.dheads_l)[b * 256 + 128 + d] = from_f32<T>(dsl * (1.f - expf(-sl)));
  }
}

void latent_bwd(const LatentBufs& L, int B, int act_dt, int gm, float beta, float inv_batch, cudaStream_t s) {
  const int n = B * 128;
  if (act_dt == DT_F32) launch_pdl(latent_bwd_kernel<float>, dim3((n + 255) / 256), dim3(256), 0, s, L, B, gm, beta, inv_batch);
  else launch_pdl(latent_bwd_kernel<bf16>, dim3((n + 255) / 256), dim3(256), 0, s, L, B, gm, beta, inv_batch);
}

// ------------------------------------------------------------------ gumbel softmax (one warp per row)
template <typename T>
__global__ void gumbel_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ user_u,
                                  float* __restrict__ u_saved, float* __restrict__ y, T* __restrict__ y_act, bf16* __restrict__ y_act_lo, int B,
                                  int K, float tau, unsigned long long seed,
                                  const unsigned long long* __restrict__ counter) {
  pdl_enter();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int k = threadIdx.x & 31;
  if (row >= B) return;
  float u = 0.5f, l = -INFINITY;
  if (k < K) {
    if (user_u) u = user_u[row * K + k];
    else {
      const unsigned long long step = *counter;
      const U4 r = philox4x32(U4{(uint32_t)(row * 32 + k), 0x6B31u, (uint32_t)step, (uint32_t)(step >> 32)},
                              (uint32_t)seed, (uint32_t)(seed >> 32));
      u = u01(r.x);
    }
    l = (logits[row * 32 + k] - logf(-logf(u))) / tau;  // vae/model.py:123
  }
  float mx = l;
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float e = k < K ? expf(l - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float yv = e / sum;
  u_saved[row * 32 + k] = u;
  y[row * 32 + k] = yv;
  y_act[row * 32 + k] = from_f32<T>(yv);
  if (y_act_lo) y_act_lo[row * 32 + k] = __float2bfloat16_rn(yv - round_bf16(yv));
}

void gumbel_fwd(const float* logits, const float* user_u, float* u_saved, float* y, void* y_act, void* y_act_lo, int act_dt, int B,
                int K, float tau, unsigned long long seed, const unsigned long long* counter, cudaStream_t s) {
  const int rows_per_block = 8;
  dim3 grid((B + rows_per_block - 1) / rows_per_block), block(32 * rows_per_block);
  if (act_dt == DT_F32) launch_pdl(gumbel_fwd_kernel<float>, dim3(grid), dim3(block), 0, s, logits, user_u, u_saved, y, (float*)y_act, nullptr, B, K, tau, seed, counter);
  else launch_pdl(gumbel_fwd_kernel<bf16>, dim3(grid), dim3(block), 0, s, logits, user_u, u_saved, y, (bf16*)y_act, (bf16*)y_act_lo, B, K, tau, seed, counter);
}

template <typename T>
__global__ void gm_add_kernel(const T* __restrict__ yb0e1, const bf16* __restrict__ yb0e1_lo, const float* __restrict__ yheads,
                              T* __restrict__ hsum, bf16* __restrict__ hsum_lo, int B) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 512) return;
  const int b = idx >> 9, j = idx & 511;
  float e1o = to_f32(yb0e1[(long long)b * 1536 + 1024 + j]);
  if (yb0e1_lo) e1o += __bfloat162float(yb0e1_lo[(long long)b * 1536 + 1024 + j]);       // bf16x3: e1's output is a bf16 pair
  const float h = e1o + yheads[b * 768 + j];                                               // model.py:130
  hsum[idx] = from_f32<T>(h);
  if (hsum_lo) hsum_lo[idx] = __float2bfloat16_rn(h - round_bf16(h));
}
void gm_add(const void* yb0e1_out, const void* yb0e1_lo, const float* yheads, void* hsum, void* hsum_lo, int act_dt, int B, cudaStream_t s) {
  const int n = B * 512;
  if (act_dt == DT_F32) launch_pdl(gm_add_kernel<float>, dim3((n + 255) / 256), dim3(256), 0, s, (const float*)yb0e1_out, nullptr, yheads, (float*)hsum, nullptr, B);
  else launch_pdl(gm_add_kernel<bf16>, dim3((n + 255) / 256), dim3(256), 0, s, (const bf16*)yb0e1_out, (const bf16*)yb0e1_lo, yheads, (bf16*)hsum, (bf16*)hsum_lo, B);
}

template <typename T>
__global__ void gm_glue_a_kernel(const T* __restrict__ dhsum, const T* __restrict__ yb0e1, const float* __restrict__ yheads,
                                 const float* __restrict__ zm_g, const float* __restrict__ zs_g, T* __restrict__ d_yb0e1,
                                 T* __restrict__ d_yheads, int B, float beta, float inv_batch) {
  pdl_enter();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 640) return;
  const int b = idx / 640, j = idx % 640;
  if (j < 512) {
    const float dh = to_f32(dhsum[b * 512 + j]);
    const float e1o = to_f32(yb0e1[(long long)b * 1536 + 1024 + j]);
    const float ht = yheads[b * 768 + j];
    d_yb0e1[(long long)b * 1536 + 1024 + j] = from_f32<T>(dh * act_grad_from_out(e1o, ACT_ELU));
    d_yheads[b * 768 + j] = from_f32<T>(dh * act_grad_from_out(ht, ACT_ELU));
  } else {
    const int d = j - 512;
    const float pm = yheads[b * 768 + 512 + d], ps = yheads[b * 768 + 640 + d];
    const float mg = zm_g[b * 128 + d], sg = zs_g[b * 128 + d];
    const float k = beta * inv_batch, dl = mg - pm, ip2 = 1.f / (ps * ps);
    const float dpm = -k * dl * ip2;
    const float dps = k * (1.f / ps - (sg * sg + dl * dl) * ip2 / ps);
    d_yheads[b * 768 + 512 + d] = from_f32<T>(dpm);
    d_yheads[b * 768 + 640 + d] = from_f32<T>(dps * (1.f - expf(-ps)));
  }
}
void gm_glue_a(const void* dhsum, const void* yb0e1_out, const float* yheads, const float* zm_g, const float* zs_g,
               void* d_yb0e1, void* d_yheads, int act_dt, int B, float beta, float inv_batch, cudaStream_t s) {
  const int n = B * 640;
  if (act_dt == DT_F32)
    launch_pdl(gm_glue_a_kernel<float>, dim3((n + 255) / 256), dim3(256), 0, s, (const float*)dhsum, (const float*)yb0e1_out, yheads, zm_g, zs_g, (float*)d_yb0e1, (float*)d_yheads, B, beta, inv_batch);
  else
    launch_pdl(gm_glue_a_kernel<bf16>, dim3((n + 255) / 256), dim3(256), 0, s, (const bf16*)dhsum, (const bf16*)yb0e1_out, yheads, zm_g, zs_g, (bf16*)d_yb0e1, (bf16*)d_yheads, B, beta, inv_batch);
}

template <typename T>
__global__ void gm_glue_b_kernel(const T* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ logits,
                                 T* __restrict__ dlogits, int B, int K, float tau, float alpha, float inv_batch) {
  pdl_enter();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int k = threadIdx.x & 31;
  if (row >= B) return;
  const bool ok = k < K;
  const float yv = ok ? y[row * 32 + k] : 0.f;
  const float dyv = ok ? to_f32(dy[row * 32 + k]) : 0.f;
  float t = yv * dyv;
#pragma unroll
  for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  const float d_gs = yv * (dyv - t) / tau;
  // categorical KL (vae/trainer.py:160-161)
  const float l = ok ? logits[row * 32 + k] : -INFINITY;
  float mx = l;
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float e = ok ? expf(l - mx) : 0.f;
  float sum = e;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float py = e / sum;
  const float dfdp = ok ? logf(py + 1e-8f) + logf((float)K) + py / (py + 1e-8f) : 0.f;
  float w = py * dfdp;
#pragma unroll
  for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  const float d_kl = alpha * inv_batch * py * (dfdp - w);
  dlogits[row * 32 + k] = from_f32<T>(ok ? d_gs + d_kl : 0.f);
}
void gm_glue_b(const void* dy, const float* y, const float* logits, void* dlogits, int act_dt, int B, int K, float tau,
               float alpha, float inv_batch, cudaStream_t s) {
  dim3 grid((B + 7) / 8), block(256);
  if (act_dt == DT_F32) launch_pdl(gm_glue_b_kernel<float>, dim3(grid), dim3(block), 0, s, (const float*)dy, y, logits, (float*)dlogits, B, K, tau, alpha, inv_batch);
  else launch_pdl(gm_glue_b_kernel<bf16>, dim3(grid), dim3(block), 0, s, (const bf16*)dy, y, logits, (bf16*)dlogits, B, K, tau, alpha, inv_batch);
}

// ------------------------------------------------------------------ discretised logistic likelihood
// One element: returns the NLL (= -log_prob) and its derivatives w.r.t. mean and log_scale.
// Branch structure follows vae/trainer.py:37 exactly; the arithmetic is arranged so that only
// exp(-|.|) is ever exponentiated (no overflow) and a single log serves the three main branches.
template <bool FAST>
__device__ __forceinline__ float dll_elem(float x, float m, float ls, float& g_m, float& g_ls) {
  const float s = FAST ? __expf(-ls) : expf(-ls);
  const float c = x - m;
  const float plus = s * (c + (1.f / 255.f));
  const float mn = s * (c - (1.f / 255.f));
  const float tp = FAST ? __expf(-fabsf(plus)) : expf(-fabsf(plus));
  const float tm = FAST ? __expf(-fabsf(mn)) : expf(-fabsf(mn));
  const float rp = FAST ? __fdividef(1.f, 1.f + tp) : 1.f / (1.f + tp);
  const float rm = FAST ? __fdividef(1.f, 1.f + tm) : 1.f / (1.f + tm);
  const float sp = plus >= 0.f ? rp : tp * rp, csp = plus >= 0.f ? tp * rp : rp;  // sigmoid(plus), 1-sigmoid(plus)
  const float sm = mn >= 0.f ? rm : tm * rm, csm = mn >= 0.f ? tm * rm : rm;
  const float delta = mn >= 0.f ? csm - csp : sp - sm;
  float lp, dm, dls;
  if (x < -0.999f) {            // log_cdf_plus = plus - softplus(plus) = log sigmoid(plus)
    const float lg = FAST ? __logf(1.f + tp) : log1pf(tp);
    lp = fminf(plus, 0.f) - lg;
    dm = -s * csp;
    dls = -plus * csp;
  } else if (x > 0.999f) {      // log_one_minus_cdf_min = -softplus(min)
    const float lg = FAST ? __logf(1.f + tm) : log1pf(tm);
    lp = -fmaxf(mn, 0.f) - lg;
    dm = s * sm;
    dls = mn * sm;
  } else if (delta > 1e-5f) {   // log(max(cdf_delta, 1e-12))
    lp = FAST ? __logf(delta) : logf(delta);
    const float dsp = sp * csp, dsm = sm * csm;
    const float inv = FAST ? __fdividef(1.f, delta) : 1.f / delta;
    dm = -s * (dsp - dsm) * inv;
    dls = -(plus * dsp - mn * dsm) * inv;
  } else {                      // log_pdf_mid - log(127.5)
    const float mid = s * c;
    const float tmid = expf(-fabsf(mid));
    const float spm = fmaxf(mid, 0.f) + log1pf(tmid);
    lp = mid - ls - 2.f * spm - 4.8481163645f;
    const float sg = mid >= 0.f ? 1.f / (1.f + tmid) : tmid / (1.f + tmid);
    const float q = 1.f - 2.f * sg;
    dm = -s * q;
    dls = -mid * q - 1.f;
  }
  g_m = -dm;
  g_ls = -dls;
  return -lp;
}

__global__ void dll_elementwise_kernel(const float* __restrict__ x, const float* __restrict__ m,
                                       const float* __restrict__ ls, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float a, b;
    out[i] = dll_elem<false>(x[i], m[i], ls[i], a, b);
  }
}
void dll_elementwise(const float* x, const float* m, const float* ls, float* out, long long n, cudaStream_t s) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  dll_elementwise_kernel<<<(int)blocks, 256, 0, s>>>(x, m, ls, out, n);
}

template <typename T, int LD>
__device__ __forceinline__ void store_dout(T* p, const float* g) {  // g[6] -> LD channels (zero padded)
  if constexpr (sizeof(T) == 4) {
#pragma unroll
    for (int i = 0; i < LD; ++i) p[i] = i < 6 ? g[i] : 0.f;
  } else {
    static_assert(LD % 8 == 0 || sizeof(T) == 4, "bf16 dout pitch must be a multiple of 8");
    __nv_bfloat162 v[LD / 2];
#pragma unroll
    for (int i = 0; i < LD / 2; ++i)
      v[i] = __floats2bfloat162_rn(2 * i < 6 ? g[2 * i] : 0.f, 2 * i + 1 < 6 ? g[2 * i + 1] : 0.f);
    uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
    for (int i = 0; i < LD / 8; ++i) q[i] = reinterpret_cast<uint4*>(v)[i];
  }
}

// One thread-iteration = ONE pixel of both streams (x | x_hat): three 8-byte loads per tensor (a pixel's 6 floats are 24 bytes; the
// lanes of a warp cover 768 contiguous bytes per tensor, every sector is used).  The grid is ONE balanced wave: 148 x 4 blocks of 256
// threads (<= 64 registers), so at CelebA64 x 256 every thread takes 6.92 -> 7 pixels; the two-pixel version ran 1184 blocks of 80
// registers = 2.67 waves of blocks with 1 or 2 iterations each and left the HBM pipe 39 % busy.
// Writes d(loss)/d(decoder output) (scaled by grad_scale = 1/(B*world)) and per-block partial
// sums of the two reconstruction losses.
constexpr int kLossThreads = 256;
template <typename T, int LD, bool FAST>
__global__ void __launch_bounds__(kLossThreads, 4) pixel_loss_kernel(const float* __restrict__ inputs,
                                                                     const float* __restrict__ dec_x,
                                                                     const float* __restrict__ dec_xh,
                                                                     T* __restrict__ dout_x, T* __restrict__ dout_xh,
                                                                     long long npix, float grad_scale,
                                                                     float* __restrict__ partials, float* __restrict__ cs_x,
                                                                     float* __restrict__ cs_xh) {
  pdl_enter();
  float sum_x = 0.f, sum_xh = 0.f;
  // cs_x / cs_xh (may be NULL): [gridDim.x][16] per-block column sums of the written gradients = d5's bias-gradient partials
  float bsx[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, bsh[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bool two = dec_xh != nullptr;            // (plain GMVAE: one decoder, one likelihood term)
#pragma unroll 1
  for (long long px = blockIdx.x * (long long)kLossThreads + threadIdx.x; px < npix; px += (long long)gridDim.x * kLossThreads) {
    float in[6], ox[6], oh[6];
    const float2* ip = reinterpret_cast<const float2*>(inputs + px * 6);
    const float2* xp = reinterpret_cast<const float2*>(dec_x + px * 6);
    const float2* hp = reinterpret_cast<const float2*>(dec_xh + px * 6);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      reinterpret_cast<float2*>(in)[i] = __ldg(ip + i);
      reinterpret_cast<float2*>(ox)[i] = __ldg(xp + i);
      reinterpret_cast<float2*>(oh)[i] = two ? __ldg(hp + i) : make_float2(0.f, 0.f);
    }
    float gx[6], gh[6];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float gm_, gl_;
      sum_x += dll_elem<FAST>(in[c], ox[c], ox[3 + c], gm_, gl_);
      gx[c] = gm_ * grad_scale; gx[3 + c] = gl_ * grad_scale;
      if (two) {
        sum_xh += dll_elem<FAST>(in[3 + c], oh[c], oh[3 + c], gm_, gl_);
        gh[c] = gm_ * grad_scale; gh[3 + c] = gl_ * grad_scale;
      }
    }
    store_dout<T, LD>(dout_x + px * LD, gx);
    if (two) store_dout<T, LD>(dout_xh + px * LD, gh);
#pragma unroll
    for (int c = 0; c < 6; ++c) { bsx[c] += gx[c]; if (two) bsh[c] += gh[c]; }
  }
  if (cs_x) {
    __shared__ float cred[2][kLossThreads / 32][6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        bsx[c] += __shfl_xor_sync(0xffffffffu, bsx[c], o);
        bsh[c] += __shfl_xor_sync(0xffffffffu, bsh[c], o);
      }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
      for (int c = 0; c < 6; ++c) { cred[0][threadIdx.x >> 5][c] = bsx[c]; cred[1][threadIdx.x >> 5][c] = bsh[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int which = threadIdx.x >> 4, c = threadIdx.x & 15;
      float t = 0.f;
      if (c < 6) {
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) t += cred[which][w][c];
      }
      float* dst = which ? cs_xh : cs_x;
      if (dst) dst[(size_t)blockIdx.x * 16 + c] = t;
    }
    __syncthreads();
  }
  // deterministic block reduction
  __shared__ float red[2][kLossThreads / 32];
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sum_x += __shfl_xor_sync(0xffffffffu, sum_x, o);
    sum_xh += __shfl_xor_sync(0xffffffffu, sum_xh, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sum_x; red[1][threadIdx.x >> 5] = sum_xh; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < kLossThreads / 32; ++i) { a += red[0][i]; b += red[1][i]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
  }
}

int pixel_loss_blocks(long long npix) {
  long long b = (npix + kLossThreads - 1) / kLossThreads;
  if (b > 148 * 4) b = 148 * 4;                    // one wave: 4 resident blocks per SM (launch bounds)
  if (b < 1) b = 1;
  return (int)b;
}

void pixel_loss(const float* inputs, const float* dec_x, const float* dec_xh, void* dout_x, void* dout_xh, int dout_dt,
                int dout_ld, long long npix, float grad_scale, float* partials, bool fast_math, cudaStream_t s, float* cs_x, float* cs_xh) {
  if (dout_ld != 16 || dout_dt != DT_BF16) { cs_x = nullptr; cs_xh = nullptr; }     // (the external partial rows are 16 columns wide)
  const int blocks = pixel_loss_blocks(npix);
#define LAUNCH(T, LD, F) launch_pdl(pixel_loss_kernel<T, LD, F>, dim3(blocks), dim3(kLossThreads), 0, s, inputs, dec_x, dec_xh, (T*)dout_x, (T*)dout_xh, npix, grad_scale, partials, cs_x, cs_xh)
  if (dout_dt == DT_F32) {
    if (fast_math) LAUNCH(float, 6, true); else LAUNCH(float, 6, false);
  } else if (dout_ld == 8) {
    if (fast_math) LAUNCH(bf16, 8, true); else LAUNCH(bf16, 8, false);
  } else {
    if (fast_math) LAUNCH(bf16, 16, true); else LAUNCH(bf16, 16, false);
  }
#undef LAUNCH
}

// Final scalars: fixed-order sum of the per-block partials of the pixel likelihood (pixel_loss_kernel) and of the Gaussian
// KLs (reparam_kernel), plus the categorical KL of q(y|x) (vae/trainer.py:160-161, one thread per row), in double.
__global__ void __launch_bounds__(1024) loss_scalars_kernel(const float* __restrict__ kl_partials, int kl_blocks,
                                                            const float* __restrict__ y_logits, int B, int K,
                                                            int gm, float beta, float alpha,
                                                            const float* __restrict__ partials, int nblocks,
                                                            float* __restrict__ scalars) {
  pdl_enter();
  __shared__ double red[5][32];
  double acc[5] = {0, 0, 0, 0, 0};  // kl_x, kl_x_hat, y_kl, recon_x, recon_x_hat
  for (int i = threadIdx.x; i < kl_blocks; i += blockDim.x) { acc[0] += kl_partials[2 * i]; acc[1] += kl_partials[2 * i + 1]; }
  if (gm) {
    for (int row = threadIdx.x; row < B; row += blockDim.x) {
      float mx = -INFINITY;
      for (int k = 0; k < K; ++k) mx = fmaxf(mx, y_logits[row * 32 + k]);
      float sum = 0.f;
      for (int k = 0; k < K; ++k) sum += expf(y_logits[row * 32 + k] - mx);
      float t = 0.f;
      for (int k = 0; k < K; ++k) {
        const float py = expf(y_logits[row * 32 + k] - mx) / sum;
        t += py * (logf(py + 1e-8f) - logf(1.0f / (float)K));
      }
      acc[2] += t;
    }
  }
  for (int i = threadIdx.x; i < nblocks; i += blockDim.x) { acc[3] += partials[2 * i]; acc[4] += partials[2 * i + 1]; }
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    double v = acc[q];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t[5];
    for (int q = 0; q < 5; ++q) {
      double v = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[q][i];
      t[q] = v / B;
    }
    const double klsum = beta * (t[0] + t[1]);
    scalars[0] = (float)t[3];
    scalars[1] = (float)t[4];
    scalars[2] = (float)t[0];
    scalars[3] = (float)t[1];
    scalars[4] = gm ? (float)t[2] : (float)klsum;
    scalars[5] = (float)(t[3] + t[4] + klsum + (gm ? alpha * t[2] : 0.0));
    scalars[6] = 0.f;
    scalars[7] = 0.f;
    // running sums of the six terms + the number of passes (Keras Mean metrics, vae/trainer.py:140-144, 169-173: updated EVERY step
    // inside the step; the host reads and clears them at report time - no per-step synchronisation)
#pragma unroll
    for (int q = 0; q < 6; ++q) scalars[8 + q] += scalars[q];
    scalars[16] += 1.f;
  }
}

void loss_scalars(const float* kl_partials, int kl_blocks, const float* y_logits, int B, int K, int gm, float beta, float alpha,
                  const float* partials, int nblocks, float* scalars, cudaStream_t s) {
  launch_pdl(loss_scalars_kernel, dim3(1), dim3(1024), 0, s, kl_partials, kl_blocks, y_logits, B, K, gm, beta, alpha, partials, nblocks, scalars);
}

// ------------------------------------------------------------------ Keras Adam (ResourceApplyAdam)
__global__ void adam_prepare_kernel(AdamState* st, float lr, int staircase) {
  pdl_enter();
  const unsigned long long it = st->iterations;
  const double t = (double)(it + 1);
  double lr_t = (double)lr;
  if (staircase) lr_t *= pow(0.4, floor((double)it / 1000000.0));  // ExponentialDecay(lr,1e6,0.4,staircase) main.py:67
  st->alpha = (float)(lr_t * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
  st->iterations = it + 1;
}
void adam_prepare(AdamState* st, float lr, int staircase, cudaStream_t s) { launch_pdl(adam_prepare_kernel, dim3(1), dim3(1), 0, s, st, lr, staircase); }

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float alpha) {
  // exact op order of TF's ApplyAdam functor; explicit roundings forbid FMA contraction so the
  // result is bit-identical to the fp32 numpy oracle.
  const float omb1 = __fsub_rn(1.0f, 0.9f), omb2 = __fsub_rn(1.0f, 0.999f);
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), omb1));
  v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), omb2));
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(m, alpha), __fadd_rn(__fsqrt_rn(v), 1e-7f)));
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, const AdamState* __restrict__ st,
                                                   float alpha_host) {
  pdl_enter();
  const float alpha = st ? st->alpha : alpha_host;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, alpha);
    adam_one(pp.y, gg.y, mm.y, vv.y, alpha);
    adam_one(pp.z, gg.z, mm.z, vv.z, alpha);
    adam_one(pp.w, gg.w, mm.w, vv.w, alpha);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) adam_one(p[i], g[i], m[i], v[i], alpha);
}

void adam_apply(float* p, const float* g, float* m, float* v, long long n, const AdamState* st, float alpha_host,
                cudaStream_t s) {
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(adam_kernel, dim3((int)blocks), dim3(256), 0, s, p, g, m, v, n, st, alpha_host);
}

// ------------------------------------------------------------------ data parallel: reduce-scatter + Adam(shard) + all-gather in ONE kernel
// over NVLS multicast memory (NVLink 5 / NVSwitch).  Every rank's gradient arena and parameter arena are symmetric allocations mapped
// behind one multicast address each: multimem.ld_reduce on the gradient multicast address returns the SUM over all ranks (the switch
// reduces in flight = the reduce-scatter), the rank applies Keras Adam to ITS 1/world shard of the range (m, v live only for that
// shard: 1/world of the optimizer traffic), and multimem.st on the parameter multicast address writes the new weights into every
// rank's arena (the all-gather).  No NCCL kernel, no second pass over the gradients; a few small CTAs that co-reside with the
// persistent convolution kernels instead of taking whole SMs.  The caller brackets the kernel with cross-rank barriers.
__device__ __forceinline__ float4 multimem_ld_reduce_f32x4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int kNvlsUnroll = 4;      // independent multimem.ld_reduce round trips (SM -> switch -> every GPU -> SM, a few us) in flight per thread
__global__ void __launch_bounds__(256) nvls_adam_kernel(const float* __restrict__ mc_grads, float* __restrict__ mc_params,
                                                        const float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                                        float* __restrict__ mc_grads_out, long long off, long long cnt, int rank, int world,
                                                        const AdamState* __restrict__ st) {
  const float alpha = st->alpha;
  const long long n4 = cnt >> 2;                                   // (arena slots are multiples of 64 floats)
  const long long per = (n4 + world - 1) / world;
  const long long begin = (long long)rank * per, end = begin + per < n4 ? begin + per : n4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = begin + blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < end; i0 += stride * kNvlsUnroll) {
    float4 g[kNvlsUnroll];
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i < end) g[u] = multimem_ld_reduce_f32x4(mc_grads + off + 4 * i);      // sum over ranks, reduced by the switch
    }
#pragma unroll
    for (int u = 0; u < kNvlsUnroll; ++u) {
      const long long i = i0 + u * stride;
      if (i >= end) break;
      const long long e = off + 4 * i;
      float4 pp = *reinterpret_cast<const float4*>(params + e);
      float4 mm = *reinterpret_cast<const float4*>(m + e), vv = *reinterpret_cast<const float4*>(v + e);
      adam_one(pp.x, g[u].x, mm.x, vv.x, alpha);
      adam_one(pp.y, g[u].y, mm.y, vv.y, alpha);
      adam_one(pp.z, g[u].z, mm.z, vv.z, alpha);
      adam_one(pp.w, g[u].w, mm.w, vv.w, alpha);
      *reinterpret_cast<float4*>(m + e) = mm;
      *reinterpret_cast<float4*>(v + e) = vv;
      multimem_st_f32x4(mc_params + e, pp);                           // new weights into every rank's arena
      if (mc_grads_out) multimem_st_f32x4(mc_grads_out + e, g[u]);    // (tests: leave the reduced gradient in every arena)
    }
  }
}

void nvls_adam(const float* mc_grads, float* mc_params, const float* params, float* m, float* v, float* mc_grads_out, long long off,
               long long cnt, int rank, int world, const AdamState* st, cudaStream_t s) {
  const long long per = ((cnt >> 2) + world - 1) / world;
  // small CTAs (256 threads, ~32 registers, no shared memory) that co-reside with the persistent convolution kernels; enough of
  // them to keep ~100 k vector loads in flight across the switch
  long long blocks = (per + 256 * kNvlsUnroll - 1) / (256 * kNvlsUnroll);
  if (blocks > 128) blocks = 128;
  if (blocks < 1) blocks = 1;
  nvls_adam_kernel<<<(int)blocks, 256, 0, s>>>(mc_grads, mc_params, params, m, v, mc_grads_out, off, cnt, rank, world, st);
}

// ------------------------------------------------------------------ scramble staging
__global__ void stage_scramble_kernel(const uint8_t* __restrict__ u8, const int32_t* __restrict__ perm,
                                      float* __restrict__ inputs, int B, int H, int W, int p) {
  const long long total = (long long)B * H * W;
  const int G = W / p, npatch = (H / p) * G;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W), y = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
    const int q = perm[(long long)b * npatch + (y / p) * G + (x / p)];
    const int sy = (q / G) * p + y % p, sx = (q % G) * p + x % p;  // augmentation.py:47-54
    const uint8_t* a = u8 + idx * 3;
    const uint8_t* h = u8 + (((long long)b * H + sy) * W + sx) * 3;
    float* o = inputs + idx * 6;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      o[c] = (float)((double)a[c] / 255.0 * 2.0 - 1.0);      // vae/data.py:52: float64 math, then astype(float32)
      o[3 + c] = (float)((double)h[c] / 255.0 * 2.0 - 1.0);
    }
  }
}
// CelebA staging (vae/data.py:82-87 + augmentation.py:43-57 in one pass): centre crop of the decoded uint8 image
// (tf.image.resize_with_crop_or_pad to 178x178), bilinear resize to H x W (tf.image.resize: half-pixel centres, no antialiasing, edge
// clamp), /255*2-1, and the patch scramble - x_hat is the resized image sampled at the permuted patch's pixel.
__device__ __forceinline__ void bilinear_rgb(const uint8_t* img, int Ws, int cy, int cx, int ch, int cw, float sy, float sx, int oy, int ox, float* out) {
  const float fy = ((float)oy + 0.5f) * sy - 0.5f, fx = ((float)ox + 0.5f) * sx - 0.5f;
  const float y0f = floorf(fy), x0f = floorf(fx);
  const float wy = fy - y0f, wx = fx - x0f;
  const int y0 = min(max((int)y0f, 0), ch - 1), y1 = min(max((int)y0f + 1, 0), ch - 1);
  const int x0 = min(max((int)x0f, 0), cw - 1), x1 = min(max((int)x0f + 1, 0), cw - 1);
  const uint8_t* r0 = img + ((size_t)(cy + y0) * Ws + cx) * 3;
  const uint8_t* r1 = img + ((size_t)(cy + y1) * Ws + cx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float top = (float)r0[x0 * 3 + c] + ((float)r0[x1 * 3 + c] - (float)r0[x0 * 3 + c]) * wx;
    const float bot = (float)r1[x0 * 3 + c] + ((float)r1[x1 * 3 + c] - (float)r1[x0 * 3 + c]) * wx;
    out[c] = (top + (bot - top) * wy) / 255.f * 2.f - 1.f;         // vae/data.py:86
  }
}
__global__ void stage_resize_scramble_kernel(const uint8_t* __restrict__ u8, const int32_t* __restrict__ perm, float* __restrict__ inputs,
                                             int B, int Hs, int Ws, int cy, int cx, int ch, int cw, int H, int W, int p) {
  const long long total = (long long)B * H * W;
  const int G = W / p, npatch = (H / p) * G;
  const float sy = (float)ch / (float)H, sx = (float)cw / (float)W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W), y = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
    const int q = perm[(long long)b * npatch + (y / p) * G + (x / p)];
    const int py = (q / G) * p + y % p, px = (q % G) * p + x % p;
    const uint8_t* img = u8 + (size_t)b * Hs * Ws * 3;
    float* o = inputs + idx * 6;
    bilinear_rgb(img, Ws, cy, cx, ch, cw, sy, sx, y, x, o);
    bilinear_rgb(img, Ws, cy, cx, ch, cw, sy, sx, py, px, o + 3);
  }
}
void stage_resize_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int B, int Hs, int Ws, int cy, int cx, int ch, int cw,
                           int H, int W, int p, cudaStream_t s) {
  const long long total = (long long)B * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stage_resize_scramble_kernel<<<(int)blocks, 256, 0, s>>>(u8, perm, inputs, B, Hs, Ws, cy, cx, ch, cw, H, W, p);
}

// One uniform random permutation of n patches per image (tf.random.shuffle, augmentation.py:49), drawn on the device: Philox keys,
// bitonic sort of (key, index) in shared memory, one block per image.  n <= 4096 (64x64 image, patch_size 1).
__global__ void __launch_bounds__(1024) draw_permutations_kernel(int32_t* __restrict__ perm, int n, int npow2, unsigned long long seed,
                                                                 unsigned long long step) {
  extern __shared__ unsigned long long keys[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long k = ~0ull;                       // padding sorts last
    if (i < n) {
      const U4 r = philox4x32(U4{(uint32_t)i, (uint32_t)b, (uint32_t)step, (uint32_t)(step >> 32)}, (uint32_t)seed, (uint32_t)(seed >> 32) ^ 0x7065726Du);
      k = ((unsigned long long)r.x << 32 | (unsigned long long)(r.y & 0xFFFFF000u)) | (unsigned long long)i;     // 52 random bits, index in the low 12
    }
    keys[i] = k;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < npow2 / 2; i += blockDim.x) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[(long long)b * n + i] = (int32_t)(keys[i] & 0xFFFull);
}
void draw_permutations(int32_t* perm, int B, int n, unsigned long long seed, unsigned long long step, cudaStream_t s) {
  int npow2 = 1;
  while (npow2 < n) npow2 <<= 1;
  const int threads = npow2 / 2 < 32 ? 32 : npow2 / 2 > 1024 ? 1024 : npow2 / 2;
  draw_permutations_kernel<<<B, threads, (size_t)npow2 * sizeof(unsigned long long), s>>>(perm, n, npow2, seed, step);
}

void stage_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int B, int H, int W, int p, cudaStream_t s) {
  const long long total = (long long)B * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stage_scramble_kernel<<<(int)blocks, 256, 0, s>>>(u8, perm, inputs, B, H, W, p);
}

}  // namespace sv
