// CUDA-core (SIMT) kernels, two roles:
//  (1) production kernels of every precision mode: the bilinear 2x resize of the decoders and its adjoint (vectorised bf16 / bf16-pair
//      variants; the adjoint also emits the producer layer's bias-gradient partials) and the multi-tensor bias-gradient column sums;
//  (2) the fp32-accumulate reference of every layer - Conv2D/Dense forward, dgrad, wgrad, bias gradient - which is the whole device path
//      of SV_PRECISION_FP32_REF and the checker the tensor-core kernels in tc_kernels.cu are tested against layer by layer.
//
// Reference semantics: Keras Conv2D(padding='same') / Dense (vae/model.py:36-42,49-76,152-156),
// tf.image.resize bilinear half-pixel (vae/model.py:163-167), tape.gradient (vae/trainer.py:137).
#include "common.cuh"
#include <stdlib.h>
#include <vector>

#include "kernels.h"

namespace sv {

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) ref_conv_fwd_kernel(ConvGeom g, const TI* __restrict__ in,
                                                           const float* __restrict__ params,
                                                           TO* __restrict__ out, int round_w) {
  const long long total = (long long)g.B * g.Ho * g.Wo * g.Co;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % g.Co);
    long long pix = idx / g.Co;
    const int wo = (int)(pix % g.Wo);
    const int ho = (int)((pix / g.Wo) % g.Ho);
    const int n = (int)(pix / ((long long)g.Wo * g.Ho));
    int lc;
    const int j = part_of(g, co, lc);
    const float* w = params + g.part_w[j] + lc;
    const int ldw = g.part_n[j];
    float acc = 0.f;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int y = ho * g.stride + kh - g.pt;
      if (y < 0 || y >= g.Hi) continue;
      for (int kw = 0; kw < g.kw; ++kw) {
        const int x = wo * g.stride + kw - g.pl;
        if (x < 0 || x >= g.Wi) continue;
        const TI* ip = in + (((long long)n * g.Hi + y) * g.Wi + x) * g.in_ld + g.in_coff;
        const float* wp = w + (long long)((kh * g.kw + kw) * g.Ci) * ldw;
        for (int ci = 0; ci < g.Ci; ++ci) {
          float wv = wp[(long long)ci * ldw];
          if (round_w) wv = round_bf16(wv);
          float iv = to_f32(ip[ci]);
          if (round_w && sizeof(TI) == 4) iv = round_bf16(iv);  // bf16 path: the image is a bf16 operand too
          acc = fmaf(iv, wv, acc);
        }
      }
    }
    acc += params[g.part_b[j] + lc];
    out[pix * g.out_ld + co] = from_f32<TO>(apply_act(acc, g.part_act[j]));
  }
}

// dX[n,h,w,ci] = sum_{kh,kw,co} dY[n,(h+pt-kh)/s,(w+pl-kw)/s,co] * W[kh,kw,ci,co], then multiplied by
// the derivative of the activation that produced X (taken from X itself).
template <typename T, typename TM>
__global__ void __launch_bounds__(256) ref_conv_dgrad_kernel(ConvGeom g, const T* __restrict__ dout,
                                                             const float* __restrict__ params,
                                                             T* __restrict__ din, const TM* __restrict__ mask_src,
                                                             int mask_act, int round_w) {
  const long long total = (long long)g.B * g.Hi * g.Wi * g.Ci;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(idx % g.Ci);
    long long pix = idx / g.Ci;
    const int x = (int)(pix % g.Wi);
    const int y = (int)((pix / g.Wi) % g.Hi);
    const int n = (int)(pix / ((long long)g.Wi * g.Hi));
    float acc = 0.f;
    for (int kh = 0; kh < g.kh; ++kh) {
      const int ty = y + g.pt - kh;
      if (ty < 0 || ty % g.stride) continue;
      const int ho = ty / g.stride;
      if (ho >= g.Ho) continue;
      for (int kw = 0; kw < g.kw; ++kw) {
        const int tx = x + g.pl - kw;
        if (tx < 0 || tx % g.stride) continue;
        const int wo = tx / g.stride;
        if (wo >= g.Wo) continue;
        const T* dp = dout + (((long long)n * g.Ho + ho) * g.Wo + wo) * g.dout_ld;
        int co = 0;
        for (int j = 0; j < g.nparts; ++j) {
          const int ldw = g.part_n[j];
          const float* wp = params + g.part_w[j] + ((long long)(kh * g.kw + kw) * g.Ci + ci) * ldw;
          for (int lc = 0; lc < ldw; ++lc, ++co) {
            float wv = wp[lc];
            if (round_w) wv = round_bf16(wv);
            acc = fmaf(to_f32(dp[co]), wv, acc);
          }
        }
      }
    }
    if (mask_act != ACT_NONE)
      acc *= act_grad_from_out(to_f32(mask_src[pix * g.in_ld + g.in_coff + ci]), mask_act);
    din[pix * g.din_ld + ci] = from_f32<T>(acc);
  }
}

// dW[kh,kw,ci,co] = sum_{n,ho,wo} X[n,s*ho+kh-pt,s*wo+kw-pl,ci] * dY[n,ho,wo,co]; one thread per
// weight, fixed summation order (deterministic).
template <typename TI, typename TD>
__global__ void __launch_bounds__(256) ref_conv_wgrad_kernel(ConvGeom g, const TI* __restrict__ in,
                                                             const TD* __restrict__ dout,
                                                             float* __restrict__ grads, int round_in) {
  const long long total = (long long)g.kh * g.kw * g.Ci * g.Co;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % g.Co);
    const int ci = (int)((idx / g.Co) % g.Ci);
    const int tap = (int)(idx / ((long long)g.Co * g.Ci));
    const int kh = tap / g.kw, kw = tap % g.kw;
    float acc = 0.f;
    for (int n = 0; n < g.B; ++n)
      for (int ho = 0; ho < g.Ho; ++ho) {
        const int y = ho * g.stride + kh - g.pt;
        if (y < 0 || y >= g.Hi) continue;
        for (int wo = 0; wo < g.Wo; ++wo) {
          const int x = wo * g.stride + kw - g.pl;
          if (x < 0 || x >= g.Wi) continue;
          float a = to_f32(in[(((long long)n * g.Hi + y) * g.Wi + x) * g.in_ld + g.in_coff + ci]);
          if (round_in && sizeof(TI) == 4) a = round_bf16(a);
          const float d = to_f32(dout[(((long long)n * g.Ho + ho) * g.Wo + wo) * g.dout_ld + co]);
          acc = fmaf(a, d, acc);
        }
      }
    int lc;
    const int j = part_of(g, co, lc);
    grads[g.part_w[j] + ((long long)tap * g.Ci + ci) * g.part_n[j] + lc] = acc;
  }
}

// Column sums of dY [rows, C] (pitch ld) -> bias gradients.  Two deterministic stages:
// partial[chunk][c] over row chunks, then a fixed-order sum over chunks.
template <typename T>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const T* __restrict__ d, long long rows, int C, int ld,
                                                             int rows_per_chunk, float* __restrict__ partial) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f32(d[r * ld + c]);
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    partial[(long long)blockIdx.y * C + c] = s;
  }
}

// block = 32 columns x 8 chunk lanes; fixed summation order (deterministic)
__global__ void __launch_bounds__(256) colsum_final_kernel(ConvGeom g, const float* __restrict__ partial, int nchunks,
                                                           float* __restrict__ grads) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < g.Co)
    for (int k = threadIdx.y; k < nchunks; k += 8) s += partial[(long long)k * g.Co + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < g.Co) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    int lc;
    const int j = part_of(g, c, lc);
    grads[g.part_b[j] + lc] = t;
  }
}

// tf.image.resize(x, [2H, 2W]) (bilinear, half-pixel centres):
//   out[2i] = .25*in[max(i-1,0)] + .75*in[i];  out[2i+1] = .75*in[i] + .25*in[min(i+1,n-1)]
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                             int B, int H, int W, int C) {
  const long long total = (long long)B * 2 * H * 2 * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int ox = (int)(p % (2 * W));
    const int oy = (int)((p / (2 * W)) % (2 * H));
    const int n = (int)(p / ((long long)4 * W * H));
    const int iy = oy >> 1, ix = ox >> 1;
    const int y0 = (oy & 1) ? iy : max(iy - 1, 0), y1 = (oy & 1) ? min(iy + 1, H - 1) : iy;
    const int x0 = (ox & 1) ? ix : max(ix - 1, 0), x1 = (ox & 1) ? min(ix + 1, W - 1) : ix;
    const float ly = (oy & 1) ? 0.25f : 0.75f;  // weight of y1
    const float lx = (ox & 1) ? 0.25f : 0.75f;  // weight of x1
    const T* b = in + (long long)n * H * W * C + c;
    const float tl = to_f32(b[((long long)y0 * W + x0) * C]), tr = to_f32(b[((long long)y0 * W + x1) * C]);
    const float bl = to_f32(b[((long long)y1 * W + x0) * C]), br = to_f32(b[((long long)y1 * W + x1) * C]);
    const float top = tl + (tr - tl) * lx;
    const float bot = bl + (br - bl) * lx;
    out[idx] = from_f32<T>(top + (bot - top) * ly);
  }
}

// Adjoint of the above as a gather (deterministic), fused with the activation derivative of the
// layer that produced the low-resolution tensor.
template <typename T>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const T* __restrict__ dout, T* __restrict__ din,
                                                             const T* __restrict__ mask_src, int mask_act,
                                                             int B, int H, int W, int C) {
  const long long total = (long long)B * H * W * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    long long p = idx / C;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int n = (int)(p / ((long long)W * H));
    // 1-D adjoint taps: rows {2y-1 (or 0 at the edge), 2y, 2y+1, 2y+2 (or 2H-1 at the edge)}, weights .25,.75,.75,.25
    int ry[4] = {y > 0 ? 2 * y - 1 : 0, 2 * y, 2 * y + 1, y < H - 1 ? 2 * y + 2 : 2 * H - 1};
    int rx[4] = {x > 0 ? 2 * x - 1 : 0, 2 * x, 2 * x + 1, x < W - 1 ? 2 * x + 2 : 2 * W - 1};
    const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    const T* b = dout + (long long)n * 4 * H * W * C + c;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float row = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) row += wt[q] * to_f32(b[((long long)ry[a] * 2 * W + rx[q]) * C]);
      acc += wt[a] * row;
    }
    if (mask_act != ACT_NONE) acc *= act_grad_from_out(to_f32(mask_src[idx]), mask_act);
    din[idx] = from_f32<T>(acc);
  }
}


// ---- vectorised bf16 variants: one thread = 8 channels (one 128-bit access) of one pixel; same arithmetic order ----
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ld_bf16x8(const bf16* p) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    r.v[2 * i] = __uint_as_float(w[i] << 16);
    r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
  return r;
}
__device__ __forceinline__ void st_bf16x8(bf16* p, const F8& f) {
  uint4 q;
  uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f.v[2 * i], f.v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = q;
}

__global__ void __launch_bounds__(256) upsample2x_fwd_vec_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                                 int B, int H, int W, int C8) {
  const int total = B * 2 * H * 2 * W * C8;   // < 2^31 for every supported shape (checked by the launcher)
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c8 = idx % C8;
    int p = idx / C8;
    const int ox = p % (2 * W);
    p /= 2 * W;
    const int oy = p % (2 * H);
    const int n = p / (2 * H);
    const int iy = oy >> 1, ix = ox >> 1;
    const int y0 = (oy & 1) ? iy : max(iy - 1, 0), y1 = (oy & 1) ? min(iy + 1, H - 1) : iy;
    const int x0 = (ox & 1) ? ix : max(ix - 1, 0), x1 = (ox & 1) ? min(ix + 1, W - 1) : ix;
    const float ly = (oy & 1) ? 0.25f : 0.75f, lx = (ox & 1) ? 0.25f : 0.75f;
    const bf16* b = in + ((size_t)n * H * W) * C8 * 8 + c8 * 8;
    const F8 tl = ld_bf16x8(b + ((size_t)y0 * W + x0) * C8 * 8), tr = ld_bf16x8(b + ((size_t)y0 * W + x1) * C8 * 8);
    const F8 bl = ld_bf16x8(b + ((size_t)y1 * W + x0) * C8 * 8), br = ld_bf16x8(b + ((size_t)y1 * W + x1) * C8 * 8);
    F8 o;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float top = tl.v[i] + (tr.v[i] - tl.v[i]) * lx;
      const float bot = bl.v[i] + (br.v[i] - bl.v[i]) * lx;
      o.v[i] = top + (bot - top) * ly;
    }
    st_bf16x8(out + (size_t)idx * 8, o);
  }
}

// Quad variant: one thread = the 2x2 output pixels that lie BETWEEN four input pixels (iy, iy+1) x (ix, ix+1), i in [-1, n-1]
// with clamped reads: out[2i+1] = in[i] + (in[i+1]-in[i])*.25, out[2i+2] = in[i] + (in[i+1]-in[i])*.75 - the same (x0, x1, weight)
// triples as the per-pixel kernel above, so results are bit-identical, with 4 loads per 4 outputs instead of 16.
__global__ void __launch_bounds__(256) upsample2x_fwd_quad_kernel(const bf16* __restrict__ in, bf16* __restrict__ out,
                                                                  int B, int H, int W, int C8) {
  pdl_enter();
  const int total = B * (H + 1) * (W + 1) * C8;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c8 = idx % C8;
    int p = idx / C8;
    const int qx = p % (W + 1) - 1;
    p /= W + 1;
    const int qy = p % (H + 1) - 1;
    const int n = p / (H + 1);
    const int y0 = max(qy, 0), y1 = min(qy + 1, H - 1), x0 = max(qx, 0), x1 = min(qx + 1, W - 1);
    const bf16* b = in + ((size_t)n * H * W) * C8 * 8 + c8 * 8;
    const F8 tl = ld_bf16x8(b + ((size_t)y0 * W + x0) * C8 * 8), tr = ld_bf16x8(b + ((size_t)y0 * W + x1) * C8 * 8);
    const F8 bl = ld_bf16x8(b + ((size_t)y1 * W + x0) * C8 * 8), br = ld_bf16x8(b + ((size_t)y1 * W + x1) * C8 * 8);
    bf16* ob = out + ((size_t)n * 4 * H * W) * C8 * 8 + c8 * 8;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int oy = 2 * qy + 1 + dy;
      if (oy < 0 || oy >= 2 * H) continue;
      const float ly = dy ? 0.75f : 0.25f;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ox = 2 * qx + 1 + dx;
        if (ox < 0 || ox >= 2 * W) continue;
        const float lx = dx ? 0.75f : 0.25f;
        F8 o;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float top = tl.v[i] + (tr.v[i] - tl.v[i]) * lx;
          const float bot = bl.v[i] + (br.v[i] - bl.v[i]) * lx;
          o.v[i] = top + (bot - top) * ly;
        }
        st_bf16x8(ob + ((size_t)oy * 2 * W + ox) * C8 * 8, o);
      }
    }
  }
}

// bf16x3 forward: the low-resolution tensor is a bf16 pair (hi + lo), interpolated in fp32 and written back as a pair
// (out_lo == NULL: the consumer multiplies single bf16 - the input of d5).  Same quad decomposition as upsample2x_fwd_quad_kernel.
__device__ __forceinline__ F8 ld_pair8(const bf16* hi, const bf16* lo, size_t off) {
  F8 a = ld_bf16x8(hi + off);
  const F8 b = ld_bf16x8(lo + off);
#pragma unroll
  for (int i = 0; i < 8; ++i) a.v[i] += b.v[i];
  return a;
}
// Launch shape: one block per row of quads (blockIdx.x = n * (H + 1) + quad row), one thread per (quad column, channel group) with
// C8 = 1 << c8_shift - no per-thread division (the flat-index form spent a fifth of its instructions on three runtime div/mods).
__global__ void __launch_bounds__(256) upsample2x_fwd_pair_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo,
                                                                  bf16* __restrict__ out_hi, bf16* __restrict__ out_lo, int B, int H, int W, int C8,
                                                                  int c8_shift) {
  pdl_enter();
  const int c8 = threadIdx.x & (C8 - 1);
  const int n = blockIdx.x / (H + 1);
  const int qy = (int)blockIdx.x - n * (H + 1) - 1;
  for (int qx = (int)(threadIdx.x >> c8_shift) - 1; qx < W; qx += (int)(blockDim.x >> c8_shift)) {   // (one pass unless (W + 1) * C8 > 256)
    const int y0 = max(qy, 0), y1 = min(qy + 1, H - 1), x0 = max(qx, 0), x1 = min(qx + 1, W - 1);
    const size_t base = ((size_t)n * H * W) * C8 * 8 + c8 * 8;
    const F8 tl = ld_pair8(in_hi, in_lo, base + ((size_t)y0 * W + x0) * C8 * 8), tr = ld_pair8(in_hi, in_lo, base + ((size_t)y0 * W + x1) * C8 * 8);
    const F8 bl = ld_pair8(in_hi, in_lo, base + ((size_t)y1 * W + x0) * C8 * 8), br = ld_pair8(in_hi, in_lo, base + ((size_t)y1 * W + x1) * C8 * 8);
    const size_t obase = ((size_t)n * 4 * H * W) * C8 * 8 + c8 * 8;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int oy = 2 * qy + 1 + dy;
      if (oy < 0 || oy >= 2 * H) continue;
      const float ly = dy ? 0.75f : 0.25f;
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ox = 2 * qx + 1 + dx;
        if (ox < 0 || ox >= 2 * W) continue;
        const float lx = dx ? 0.75f : 0.25f;
        F8 o, l;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float top = tl.v[i] + (tr.v[i] - tl.v[i]) * lx;
          const float bot = bl.v[i] + (br.v[i] - bl.v[i]) * lx;
          o.v[i] = top + (bot - top) * ly;
          l.v[i] = o.v[i] - round_bf16(o.v[i]);
        }
        const size_t off = obase + ((size_t)oy * 2 * W + ox) * C8 * 8;
        st_bf16x8(out_hi + off, o);
        if (out_lo) st_bf16x8(out_lo + off, l);
      }
    }
  }
}

// colsum (may be NULL): [gridDim.x][8 * C8] per-block column sums of din (before its bf16 rounding) = the producer's bias-gradient
// partials.  gridDim.x * 256 is a multiple of C8, so a thread keeps ONE channel group over its grid-stride items.
__global__ void __launch_bounds__(256) upsample2x_bwd_vec_kernel(const bf16* __restrict__ dout, bf16* __restrict__ din,
                                                                 const bf16* __restrict__ mask_src, int mask_act,
                                                                 int B, int H, int W, int C8, float* __restrict__ colsum) {
  pdl_enter();
  const int total = B * H * W * C8;
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c8 = idx % C8;
    int p = idx / C8;
    const int x = p % W;
    p /= W;
    const int y = p % H;
    const int n = p / H;
    const int ry[4] = {y > 0 ? 2 * y - 1 : 0, 2 * y, 2 * y + 1, y < H - 1 ? 2 * y + 2 : 2 * H - 1};
    const int rx[4] = {x > 0 ? 2 * x - 1 : 0, 2 * x, 2 * x + 1, x < W - 1 ? 2 * x + 2 : 2 * W - 1};
    const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
    const bf16* b = dout + ((size_t)n * 4 * H * W) * C8 * 8 + c8 * 8;
    F8 acc;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc.v[i] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      F8 row;
#pragma unroll
      for (int i = 0; i < 8; ++i) row.v[i] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const F8 t = ld_bf16x8(b + ((size_t)ry[a] * 2 * W + rx[q]) * C8 * 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) row.v[i] += wt[q] * t.v[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] += wt[a] * row.v[i];
    }
    if (mask_act != ACT_NONE) {
      const F8 m = ld_bf16x8(mask_src + (size_t)idx * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc.v[i] *= act_grad_from_out(m.v[i], mask_act);
    }
    st_bf16x8(din + (size_t)idx * 8, acc);
#pragma unroll
    for (int i = 0; i < 8; ++i) csum[i] += acc.v[i];
  }
  if (colsum) {
    // fixed-order block reduction: lanes of equal channel group (lane % C8) by xor-shuffles, then the 8 warps through shared memory
    __shared__ float red[8][16][8];
    for (int o = C8; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) csum[i] += __shfl_xor_sync(0xffffffffu, csum[i], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < C8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane][i] = csum[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < C8 * 8) {
      const int g = threadIdx.x >> 3, i = threadIdx.x & 7;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][g][i];
      colsum[(size_t)blockIdx.x * (C8 * 8) + threadIdx.x] = t;
    }
  }
}

// Same adjoint, one thread = 8 channels of a 2x2 BLOCK of low-resolution pixels: the 6x6 high-resolution window of the block is
// read once (36 loads for 4 outputs instead of 64) and streamed row by row - per input row the two horizontal 4-tap sums, then the
// vertical accumulation into the block's two output rows - in exactly the operation order of upsample2x_bwd_vec_kernel, so the
// results are bit-identical.  H and W even.  colsum as above.
__global__ void __launch_bounds__(256) upsample2x_bwd_blk_kernel(const bf16* __restrict__ dout, bf16* __restrict__ din,
                                                                 const bf16* __restrict__ mask_src, int mask_act,
                                                                 int B, int H, int W, int C8, float* __restrict__ colsum) {
  pdl_enter();
  const int Hb = H >> 1, Wb = W >> 1;
  const int total = B * Hb * Wb * C8;
  float csum[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const float wt[4] = {0.25f, 0.75f, 0.75f, 0.25f};
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int c8 = idx % C8;
    int p = idx / C8;
    const int bx = p % Wb;
    p /= Wb;
    const int by = p % Hb;
    const int n = p / Hb;
    const int y0 = 2 * by, x0 = 2 * bx;
    // window rows / columns: [0] = tap 0 of the first output, [1..4] = 2*y0 .. 2*y0+3, [5] = tap 3 of the second output (edges clamp)
    const int wr[6] = {y0 > 0 ? 2 * y0 - 1 : 0, 2 * y0, 2 * y0 + 1, 2 * y0 + 2, 2 * y0 + 3, y0 + 1 < H - 1 ? 2 * y0 + 4 : 2 * H - 1};
    const int wc[6] = {x0 > 0 ? 2 * x0 - 1 : 0, 2 * x0, 2 * x0 + 1, 2 * x0 + 2, 2 * x0 + 3, x0 + 1 < W - 1 ? 2 * x0 + 4 : 2 * W - 1};
    const bf16* b = dout + ((size_t)n * 4 * H * W) * C8 * 8 + c8 * 8;
    F8 acc[2][2];                         // [output row][output column]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][j].v[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const bf16* rowp = b + (size_t)wr[r] * 2 * W * C8 * 8;
      F8 t[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) t[q] = ld_bf16x8(rowp + (size_t)wc[q] * C8 * 8);
      F8 h0, h1;                          // horizontal sums of output columns x0 (window columns 0..3) and x0 + 1 (2..5)
#pragma unroll
      for (int k = 0; k < 8; ++k) { h0.v[k] = 0.f; h1.v[k] = 0.f; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { h0.v[k] += wt[q] * t[q].v[k]; h1.v[k] += wt[q] * t[q + 2].v[k]; }
      }
      if (r < 4) {                        // window rows 0..3 are taps 0..3 of output row y0
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[0][0].v[k] += wt[r] * h0.v[k]; acc[0][1].v[k] += wt[r] * h1.v[k]; }
      }
      if (r >= 2) {                       // window rows 2..5 are taps 0..3 of output row y0 + 1
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[1][0].v[k] += wt[r - 2] * h0.v[k]; acc[1][1].v[k] += wt[r - 2] * h1.v[k]; }
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const size_t o = ((((size_t)n * H + (y0 + i)) * W + (x0 + j)) * C8 + c8) * 8;
        if (mask_act != ACT_NONE) {
          const F8 m = ld_bf16x8(mask_src + o);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[i][j].v[k] *= act_grad_from_out(m.v[k], mask_act);
        }
        st_bf16x8(din + o, acc[i][j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) csum[k] += acc[i][j].v[k];
      }
  }
  if (colsum) {
    __shared__ float red[8][16][8];
    for (int o = C8; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) csum[i] += __shfl_xor_sync(0xffffffffu, csum[i], o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < C8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[warp][lane][i] = csum[i];
    }
    __syncthreads();
    if ((int)threadIdx.x < C8 * 8) {
      const int g = threadIdx.x >> 3, i = threadIdx.x & 7;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w][g][i];
      colsum[(size_t)blockIdx.x * (C8 * 8) + threadIdx.x] = t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
static inline int grid_for(long long total, int block = 256, int cap = 148 * 16) {
  long long b = (total + block - 1) / block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define SV_DISPATCH2(dtA, dtB, CALL)                                   \
  do {                                                                 \
    if ((dtA) == DT_F32 && (dtB) == DT_F32) { CALL(float, float); }    \
    else if ((dtA) == DT_F32 && (dtB) == DT_BF16) { CALL(float, bf16); } \
    else if ((dtA) == DT_BF16 && (dtB) == DT_F32) { CALL(bf16, float); } \
    else { CALL(bf16, bf16); }                                         \
  } while (0)

void ref_conv_fwd(const ConvGeom& g, const void* in, int in_dt, const float* params, void* out, int out_dt,
                  bool round_w, cudaStream_t s) {
  const long long total = (long long)g.B * g.Ho * g.Wo * g.Co;
#define CALL(TI, TO) ref_conv_fwd_kernel<TI, TO><<<grid_for(total), 256, 0, s>>>(g, (const TI*)in, params, (TO*)out, round_w)
  SV_DISPATCH2(in_dt, out_dt, CALL);
#undef CALL
}

void ref_conv_dgrad(const ConvGeom& g, const void* dout, int dt, const float* params, void* din,
                    const void* mask_src, int mask_dt, int mask_act, bool round_w, cudaStream_t s) {
  const long long total = (long long)g.B * g.Hi * g.Wi * g.Ci;
  if (mask_act == ACT_NONE) { mask_src = din; mask_dt = dt; }
#define CALL(T, TM) ref_conv_dgrad_kernel<T, TM><<<grid_for(total), 256, 0, s>>>(g, (const T*)dout, params, (T*)din, (const TM*)mask_src, mask_act, round_w)
  SV_DISPATCH2(dt, mask_dt, CALL);
#undef CALL
}

void ref_conv_wgrad(const ConvGeom& g, const void* in, int in_dt, const void* dout, int dout_dt, float* grads, bool round_in,
                    cudaStream_t s) {
  const long long total = (long long)g.kh * g.kw * g.Ci * g.Co;
#define CALL(TI, TD) ref_conv_wgrad_kernel<TI, TD><<<grid_for(total, 256, 1 << 20), 256, 0, s>>>(g, (const TI*)in, (const TD*)dout, grads, round_in)
  SV_DISPATCH2(in_dt, dout_dt, CALL);
#undef CALL
}

// bf16, ld % 8 == 0: each thread sums 8 adjacent columns (one 128-bit load per row); block = (ld/8 column groups) x row lanes
__global__ void __launch_bounds__(256) colsum_partial_vec_kernel(const bf16* __restrict__ d, long long rows, int C, int ld,
                                                                 int rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float red_v[];   // [blockDim.y][ld]
  const int groups = ld >> 3;
  const int gi = threadIdx.x;        // column group
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  long long r1 = r0 + rows_per_chunk;
  if (r1 > rows) r1 = rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (gi < groups)
    for (long long r = r0 + threadIdx.y; r < r1; r += blockDim.y) {
      const F8 v = ld_bf16x8(d + r * ld + gi * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v.v[i];
    }
  if (gi < groups) {
#pragma unroll
    for (int i = 0; i < 8; ++i) red_v[threadIdx.y * ld + gi * 8 + i] = acc[i];
  }
  __syncthreads();
  for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y) {
    float s = 0.f;
    for (int y = 0; y < (int)blockDim.y; ++y) s += red_v[y * ld + c];
    partial[(long long)blockIdx.x * C + c] = s;
  }
}

int colsum_chunks(long long rows) {
  long long c = (rows + 511) / 512;
  if (c > 256) c = 256;
  if (c < 1) c = 1;
  return (int)c;
}

void bias_grad(const ConvGeom& g, const void* dout, int dt, float* partial_ws, float* grads, cudaStream_t s) {
  const long long rows = (long long)g.B * g.Ho * g.Wo;
  const int nch = colsum_chunks(rows);
  const int rpc = (int)((rows + nch - 1) / nch);
  if (dt == DT_BF16 && (g.dout_ld % 8) == 0 && g.dout_ld <= 2048) {
    const int groups = g.dout_ld / 8;
    int bx = 1;
    while (bx < groups && bx < 256) bx <<= 1;
    const int by = 256 / bx > 0 ? 256 / bx : 1;
    colsum_partial_vec_kernel<<<nch, dim3(bx, by), (size_t)by * g.dout_ld * sizeof(float), s>>>((const bf16*)dout, rows, g.Co, g.dout_ld, rpc,
                                                                                               partial_ws);
    colsum_final_kernel<<<(g.Co + 31) / 32, dim3(32, 8), 0, s>>>(g, partial_ws, nch, grads);
    return;
  }
  dim3 grid((g.Co + 31) / 32, nch), block(32, 8);
  if (dt == DT_F32)
    colsum_partial_kernel<float><<<grid, block, 0, s>>>((const float*)dout, rows, g.Co, g.dout_ld, rpc, partial_ws);
  else
    colsum_partial_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)dout, rows, g.Co, g.dout_ld, rpc, partial_ws);
  colsum_final_kernel<<<(g.Co + 31) / 32, dim3(32, 8), 0, s>>>(g, partial_ws, nch, grads);
}

// ---------------------------------------------------------------------------------------------
// Multi-tensor bias gradients: the column sums of every dY of one backward branch (a decoder / an encoder) in two launches
// instead of two per layer.  Job = (layer, slab of <= 256 columns); block = (job, row chunk): 256 threads = column groups of
// 8 bf16 (one 128-bit load per row) x row lanes, fixed-order shared-memory reduction over the row lanes, then a second
// kernel adds the chunks in order -> deterministic.
// ---------------------------------------------------------------------------------------------
struct ColsumJob {
  const bf16* d;
  long long rows;
  int ld, col0, ncols, c_valid;     // slab [col0, col0 + ncols) of a [rows, ld] matrix; columns >= c_valid are padding
  int rows_per_chunk, nchunks;
  long long partial_off;            // floats; partial[chunk][ncols]
  int block_start, fblock_start;    // first block of this job in the partial / final kernels
  int nparts, part_n[3];
  long long part_b[3];
};

__global__ void __launch_bounds__(256) colsum_multi_partial_kernel(const ColsumJob* __restrict__ jobs, int njobs, float* __restrict__ partial) {
  pdl_enter();
  __shared__ float red[2048];
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].block_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const ColsumJob J = jobs[lo];
  const int chunk = blockIdx.x - J.block_start;
  const int ngroups = J.ncols >> 3;              // 2 .. 32, power of two
  const int lanes = 256 / ngroups;
  const int gi = threadIdx.x % ngroups, ry = threadIdx.x / ngroups;
  const long long r0 = (long long)chunk * J.rows_per_chunk;
  long long r1 = r0 + J.rows_per_chunk;
  if (r1 > J.rows) r1 = J.rows;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const bf16* base = J.d + J.col0 + gi * 8;
  long long r = r0 + ry;
  for (; r + 3 * lanes < r1; r += 4 * lanes) {    // four independent 128-bit loads in flight per thread
    const F8 a = ld_bf16x8(base + r * J.ld), b = ld_bf16x8(base + (r + lanes) * J.ld);
    const F8 c = ld_bf16x8(base + (r + 2 * lanes) * J.ld), d = ld_bf16x8(base + (r + 3 * lanes) * J.ld);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += (a.v[i] + b.v[i]) + (c.v[i] + d.v[i]);
  }
  for (; r < r1; r += lanes) {
    const F8 a = ld_bf16x8(base + r * J.ld);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += a.v[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry * J.ncols + gi * 8 + i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < J.ncols; c += 256) {
    float t = 0.f;
    for (int y = 0; y < lanes; ++y) t += red[y * J.ncols + c];
    partial[J.partial_off + (long long)chunk * J.ncols + c] = t;
  }
}

constexpr int kColsumFinalLanes = 32;   // chunk lanes per column: the jobs fed by upsample2x_bwd / pixel_loss bring 592-1184 chunks for 16-128 columns
constexpr int kColsumFinalCols = 8;     // columns per block (one 32-byte sector per partial row); 256-thread blocks fit beside the persistent
                                        // chain kernels' registers, 1024-thread ones waited 70 us for a drained SM
__global__ void __launch_bounds__(kColsumFinalCols * kColsumFinalLanes) colsum_multi_final_kernel(const ColsumJob* __restrict__ jobs, int njobs,
                                                                                                  const float* __restrict__ partial,
                                                                                                  float* __restrict__ grads) {
  pdl_enter();
  int lo = 0, hi = njobs - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].fblock_start <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  // block = 8 columns x 32 chunk lanes (lane y sums chunks y, y+32, ... with four loads in flight), fixed-order combine through shared memory
  __shared__ float red[kColsumFinalLanes][kColsumFinalCols + 1];
  const ColsumJob& J = jobs[lo];
  const int tx = threadIdx.x % kColsumFinalCols, ty = threadIdx.x / kColsumFinalCols;
  const int c = (blockIdx.x - J.fblock_start) * kColsumFinalCols + tx;
  float t = 0.f;
  if (c < J.ncols) {
    const float* p = partial + J.partial_off + c;
    const long long st = (long long)kColsumFinalLanes * J.ncols;
    int k = ty;
    for (; k + 3 * kColsumFinalLanes < J.nchunks; k += 4 * kColsumFinalLanes) {
      const float* q = p + (long long)k * J.ncols;
      const float a = q[0], b = q[st], d = q[2 * st], e = q[3 * st];
      t += (a + b) + (d + e);
    }
    for (; k < J.nchunks; k += kColsumFinalLanes) t += p[(long long)k * J.ncols];
  }
  red[ty][tx] = t;
  __syncthreads();
  if (ty == 0 && c < J.ncols && J.col0 + c < J.c_valid) {
    float u = 0.f;
#pragma unroll
    for (int i = 0; i < kColsumFinalLanes; ++i) u += red[i][tx];
    int lc = J.col0 + c, j = 0;
    while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
    grads[J.part_b[j] + lc] = u;
  }
}

struct ColsumTable {
  ColsumJob* dev = nullptr;
  int njobs = 0, nblocks = 0, nfblocks = 0;
  float* partial = nullptr;
  std::vector<long long> ext_off;     // per spec: float offset of its external partial block, or -1
};

int colsum_ext_cols(int dout_ld) { return dout_ld; }
static void colsum_build(const ColsumSpec* specs, int n, std::vector<ColsumJob>& jobs, long long& partial_floats, int& nblocks, int& nfblocks,
                         std::vector<long long>* ext_off = nullptr) {
  jobs.clear();
  partial_floats = 0; nblocks = 0; nfblocks = 0;
  for (int i = 0; i < n; ++i) {
    const ConvGeom& g = specs[i].g;
    const long long rows = (long long)g.B * g.Ho * g.Wo;
    const int ld = g.dout_ld;
    if (specs[i].ext_chunks > 0) {                  // partials come from the kernel that produces dY: only the final pass runs here
      ColsumJob J{};
      J.d = nullptr;
      J.rows = rows; J.ld = ld; J.col0 = 0; J.c_valid = g.Co;
      J.ncols = colsum_ext_cols(ld);
      J.rows_per_chunk = 0; J.nchunks = specs[i].ext_chunks;
      J.partial_off = partial_floats;
      partial_floats += (long long)J.nchunks * J.ncols;
      J.block_start = nblocks;                       // (no blocks in the partial kernel)
      J.fblock_start = nfblocks; nfblocks += (J.ncols + kColsumFinalCols - 1) / kColsumFinalCols;
      J.nparts = g.nparts;
      for (int k = 0; k < 3; ++k) { J.part_n[k] = g.part_n[k]; J.part_b[k] = g.part_b[k]; }
      jobs.push_back(J);
      if (ext_off) ext_off->push_back(J.partial_off);
      continue;
    }
    if (ext_off) ext_off->push_back(-1);
    for (int col0 = 0; col0 < ld && col0 < ((g.Co + 7) & ~7); col0 += 256) {
      ColsumJob J{};
      J.d = (const bf16*)specs[i].dout;
      J.rows = rows; J.ld = ld; J.col0 = col0; J.c_valid = g.Co;
      int ncols = ld - col0 < 256 ? ld - col0 : 256;
      int p2 = 16;                                   // slab width: power of two in [16, 256] (ld is one, or a multiple of 256)
      while (p2 < ncols) p2 <<= 1;
      J.ncols = p2 > 256 ? 256 : p2;
      // chunks of ~32 KB of dY (enough blocks to fill the GPU: the kernel is latency-bound per block), at most 512 per job
      const long long rows_per = (32768 / (J.ncols * 2)) < 64 ? 64 : 32768 / (J.ncols * 2);
      long long nch = (rows + rows_per - 1) / rows_per;
      if (nch > 512) nch = 512;
      if (nch < 1) nch = 1;
      J.rows_per_chunk = (int)((rows + nch - 1) / nch);
      J.nchunks = (int)((rows + J.rows_per_chunk - 1) / J.rows_per_chunk);
      J.partial_off = partial_floats;
      partial_floats += (long long)J.nchunks * J.ncols;
      J.block_start = nblocks; nblocks += J.nchunks;
      J.fblock_start = nfblocks; nfblocks += (J.ncols + kColsumFinalCols - 1) / kColsumFinalCols;
      J.nparts = g.nparts;
      for (int k = 0; k < 3; ++k) { J.part_n[k] = g.part_n[k]; J.part_b[k] = g.part_b[k]; }
      jobs.push_back(J);
    }
  }
}

bool colsum_multi_supported(const ColsumSpec* specs, int n) {
  for (int i = 0; i < n; ++i) {
    const int ld = specs[i].g.dout_ld;
    const bool pow2 = ld >= 16 && (ld & (ld - 1)) == 0;
    if (!(pow2 || (ld % 256) == 0)) return false;
  }
  return n > 0;
}

long long colsum_table_partial_floats(const ColsumSpec* specs, int n) {
  std::vector<ColsumJob> jobs;
  long long pf; int nb, nf;
  colsum_build(specs, n, jobs, pf, nb, nf);
  return pf;
}

ColsumTable* colsum_table_create(const ColsumSpec* specs, int n, float* partial_ws, const char** err) {
  std::vector<ColsumJob> jobs;
  long long pf; int nb, nf;
  std::vector<long long> ext;
  colsum_build(specs, n, jobs, pf, nb, nf, &ext);
  ColsumTable* T = new ColsumTable();
  T->ext_off = ext;
  T->njobs = (int)jobs.size(); T->nblocks = nb; T->nfblocks = nf; T->partial = partial_ws;
  if (T->njobs && (cudaMalloc(&T->dev, jobs.size() * sizeof(ColsumJob)) != cudaSuccess ||
                   cudaMemcpy(T->dev, jobs.data(), jobs.size() * sizeof(ColsumJob), cudaMemcpyHostToDevice) != cudaSuccess)) {
    *err = "colsum table upload failed";
    delete T;
    return nullptr;
  }
  return T;
}

void colsum_table_destroy(ColsumTable* t) {
  if (!t) return;
  if (t->dev) cudaFree(t->dev);
  delete t;
}

int colsum_table_run(ColsumTable* t, float* grads, cudaStream_t s) {
  if (!t || !t->njobs) return 0;
  if (t->nblocks) launch_pdl(colsum_multi_partial_kernel, dim3(t->nblocks), dim3(256), 0, s, t->dev, t->njobs, t->partial);
  launch_pdl(colsum_multi_final_kernel, dim3(t->nfblocks), dim3(kColsumFinalCols * kColsumFinalLanes), 0, s, t->dev, t->njobs, t->partial, grads);
  return t->nblocks ? 2 : 1;
}
float* colsum_table_ext_partial(ColsumTable* t, int i) {
  return (t && i >= 0 && i < (int)t->ext_off.size() && t->ext_off[i] >= 0) ? t->partial + t->ext_off[i] : nullptr;
}

void upsample2x_fwd(const void* in, void* out, int dt, int B, int H, int W, int C, cudaStream_t s) {
  const long long total = (long long)B * 4 * H * W * C;
  if (dt == DT_BF16 && (C % 8) == 0 && total / 8 < (1ll << 31)) {
    const long long quads = (long long)B * (H + 1) * (W + 1) * (C / 8);
    launch_pdl(upsample2x_fwd_quad_kernel, dim3(grid_for(quads, 256, 148 * 32)), dim3(256), 0, s, (const bf16*)in, (bf16*)out, B, H, W, C / 8);
    return;
  }
  if (dt == DT_F32)
    upsample2x_fwd_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)in, (float*)out, B, H, W, C);
  else
    upsample2x_fwd_kernel<bf16><<<grid_for(total), 256, 0, s>>>((const bf16*)in, (bf16*)out, B, H, W, C);
}

void upsample2x_fwd_pair(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int B, int H, int W, int C, cudaStream_t s) {
  const int C8 = C / 8;                                 // (a power of two for every decoder tensor: 128 / 64 / 32 channels)
  int shift = 0;
  while ((1 << shift) < C8) ++shift;
  int threads = ((W + 1) * C8 + 31) / 32 * 32;          // 160 for the CelebA64 decoder (9 x 16, 17 x 8, 33 x 4 quads x channel groups)
  if (threads > 256) threads = 256;
  launch_pdl(upsample2x_fwd_pair_kernel, dim3(B * (H + 1)), dim3(threads), 0, s, (const bf16*)in_hi, (const bf16*)in_lo, (bf16*)out_hi,
             (bf16*)out_lo, B, H, W, C8, shift);
}

// grid of the bf16 launch: at most 8 blocks per SM (a multiple of every C / 8, so the grid-stride keeps a thread's channel group)
static bool upsample2x_bwd_use_blk(int H, int W) {
  const char* v = getenv("SV_UPS_BWD_BLK");          // (read per engine / per launch: the A/B test toggles it inside one process)
  return !(v && v[0] == '0') && (H % 2) == 0 && (W % 2) == 0;
}
int upsample2x_bwd_blocks(int B, int H, int W, int C) {
  if (upsample2x_bwd_use_blk(H, W)) return grid_for((long long)B * (H / 2) * (W / 2) * (C / 8), 256, 148 * 8);
  return grid_for((long long)B * H * W * (C / 8), 256, 148 * 8);
}
void upsample2x_bwd(const void* dout, void* din, const void* mask_src, int mask_act, int dt, int B, int H, int W,
                    int C, cudaStream_t s, float* colsum_partial) {
  const long long total = (long long)B * H * W * C;
  if (dt == DT_BF16 && (C % 8) == 0 && total / 8 < (1ll << 29)) {
    const bool fold = colsum_partial && C / 8 <= 16 && (256 % (C / 8)) == 0;
    if (upsample2x_bwd_use_blk(H, W) && C / 8 <= 16 && (256 % (C / 8)) == 0) {
      // (the grid is upsample2x_bwd_blocks() whether or not the partials are wanted: one launch shape per layer)
      launch_pdl(upsample2x_bwd_blk_kernel, dim3(upsample2x_bwd_blocks(B, H, W, C)), dim3(256), 0, s, (const bf16*)dout, (bf16*)din,
                 (const bf16*)mask_src, mask_act, B, H, W, C / 8, fold ? colsum_partial : (float*)nullptr);
      return;
    }
    launch_pdl(upsample2x_bwd_vec_kernel, dim3(fold ? upsample2x_bwd_blocks(B, H, W, C) : grid_for(total / 8, 256, 148 * 32)), dim3(256), 0, s,
               (const bf16*)dout, (bf16*)din, (const bf16*)mask_src, mask_act, B, H, W, C / 8, fold ? colsum_partial : (float*)nullptr);
    return;
  }
  if (dt == DT_F32)
    upsample2x_bwd_kernel<float><<<grid_for(total), 256, 0, s>>>((const float*)dout, (float*)din, (const float*)mask_src, mask_act, B, H, W, C);
  else
    upsample2x_bwd_kernel<bf16><<<grid_for(total), 256, 0, s>>>((const bf16*)dout, (bf16*)din, (const bf16*)mask_src, mask_act, B, H, W, C);
}

}  // namespace sv
