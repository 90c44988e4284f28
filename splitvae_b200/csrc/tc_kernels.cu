// Tensor-core implicit-GEMM convolution / dense kernels for sm_100a.
//
//   * operands are staged by TMA (cp.async.bulk.tensor) into 128B/64B/32B-swizzled shared memory:
//     the A tile of a conv tap is ONE 4-D box of the NHWC activation tensor shifted by the tap offset
//     (out-of-bounds rows/cols are zero-filled by the TMA unit = TF 'same' padding; stride-2 layers use the
//     tensor map's element strides), the B tile is a 2-D box of the packed bf16 weights;
//   * one elected thread issues tcgen05.mma (M=128, N=tile_cols, K=16 per instruction, bf16 x bf16 -> fp32)
//     with the accumulator in TMEM; tcgen05.commit releases smem stages / signals the epilogue;
//   * four epilogue warps read the accumulator with tcgen05.ld (thread = output pixel), add the bias, apply the
//     activation (forward) or the activation derivative of the producer layer (dgrad), convert and store NHWC.
//
// Reference semantics: Keras Conv2D / Dense forward (vae/model.py:36-42,49-76,152-156) and their
// input gradients under tape.gradient (vae/trainer.py:137,166).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_device.cuh"
#include "tc_kernels.h"

namespace sv {

namespace {

constexpr int kMaxStages = 8;
constexpr int kThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
char g_tc_error[256] = "";

struct SmemCtl {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// bf16x3 forward (TcLaunch::split): a tap's K loop in LOGICAL chunks -> physical A chunk (hi chunks 0..kc-1, lo chunks kc..2kc-1)
// and physical weight k-block (per tap [W_hi(ch) W_lo(ch)] interleaved by chunk: block 2*ch + sec).
//   split 1: logical (hi,Whi) x kc, (lo,Whi) x kc, (hi,Wlo) x kc;  split 2 (first layer, pairs inside the pixel): (x, [Whi Whi]), (x, [Wlo 0])
__device__ __forceinline__ void logical_chunk(int split, int kc, int c, int& ap, int& bp) {
  if (split == 1) { const int cls = c / kc, ch = c - cls * kc; ap = cls == 1 ? kc + ch : ch; bp = 2 * ch + (cls == 2 ? 1 : 0); }
  else if (split == 2) { ap = 0; bp = c; }
  else { ap = c; bp = c; }
}
// the same products driven by the PHYSICAL weight block jj of a tap (streamed-weight kernels read every block once): the A chunks
// it multiplies; returns how many (1 or 2)
__device__ __forceinline__ int phys_block_chunks(int split, int kc, int jj, int& a0, int& a1) {
  if (split == 1) { const int ch = jj >> 1; a0 = ch; a1 = kc + ch; return (jj & 1) ? 1 : 2; }
  a1 = 0;
  if (split == 2) { a0 = 0; return 1; }
  a0 = jj;
  return 1;
}


// One epilogue thread = one accumulator row (TMEM lane) = one output pixel: reads tile_cols fp32 columns, adds the bias,
// applies the activation (forward) or multiplies by the activation derivative of the producer layer (dgrad), stores NHWC.
// The activation / mask kind / output type are compile-time so each instantiation is a short straight-line loop
// (a single generic body with every activation inlined was ~6000 instructions and I-cache bound).
template <int ACT>
__device__ __forceinline__ float act_t(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == ACT_ELU) return x > 0.f ? x : expm1f(x);
  if (ACT == ACT_SOFTPLUS) return softplus_f(x);
  return x;
}
template <int MASK>
__device__ __forceinline__ float mask_t(float y) {
  if (MASK == ACT_RELU) return y > 0.f ? 1.f : 0.f;
  if (MASK == ACT_ELU) return y > 0.f ? 1.f : y + 1.f;
  if (MASK == ACT_SOFTPLUS) return 1.f - expf(-y);
  return 1.f;
}

// Loop-invariant epilogue parameters, read ONCE into registers.  (ncu, round 1: with the launch descriptor read in place, every
// asm("memory") clobber - each tcgen05.ld / wait - forced the fields to be re-read through a generic pointer, LD.E + long
// scoreboard stall per field per block, and the epilogue took ~70 % of a halo CTA's lifetime.)
struct EpiRegs {
  const float* bias;
  const bf16* mask_src;
  void* out;
  bf16* out_lo;          // bf16x3 forward: lo plane of the output (value = hi + lo), same indexing as `out`
  int tile_cols, n_valid, out_ld, mask_ld, mask_coff;
  int stack_off;         // N-stacked pair accumulators (TcLaunch::nstack2): the correction columns sit stack_off TMEM columns further
};
__device__ __forceinline__ EpiRegs load_epi_regs(const TcLaunch& P) {
  EpiRegs E;
  E.bias = P.bias; E.mask_src = (const bf16*)P.mask_src; E.out = P.out; E.out_lo = (bf16*)P.out_lo;
  E.tile_cols = P.tile_cols; E.n_valid = P.n_valid; E.out_ld = P.out_ld; E.mask_ld = P.mask_ld; E.mask_coff = P.mask_coff;
  E.stack_off = P.nstack2 ? P.tile_cols : 0;
  return E;
}

// Finishes one block of 16 accumulator columns whose tcgen05.ld into v[] is in flight: the block's mask loads are issued first
// (they overlap the TMEM load), then the wait, then the NEXT block's tcgen05.ld is started into vn[] so that its latency hides
// behind this block's arithmetic and stores.  (Blocks of 16: two register sets of 32 columns spilled at 3 CTAs per SM.)
constexpr int kEpiBW = 16;
template <int ACT, int MASK, bool OUT_F32>
__device__ __forceinline__ void epi_block_t(const EpiRegs& E, uint32_t (&v)[kEpiBW], int cfirst, long long opix, bool valid,
                                            uint32_t next_taddr, uint32_t (&vn)[kEpiBW], uint64_t* release_bar = nullptr, uint32_t cur_taddr = 0u) {
  const int esz = OUT_F32 ? 4 : 2;
  const bool vec_ok = ((E.out_ld * esz) % 16) == 0;
  const bool mask_vec = MASK != ACT_NONE && (E.mask_ld % 8) == 0 && (E.mask_coff % 8) == 0 && cfirst + kEpiBW <= E.n_valid;
  const bf16* mrow = MASK != ACT_NONE ? E.mask_src + opix * E.mask_ld + E.mask_coff + cfirst : nullptr;
  uint4 mq[kEpiBW / 8];
  if (MASK != ACT_NONE && mask_vec && valid) {
#pragma unroll
    for (int i = 0; i < kEpiBW / 8; ++i) mq[i] = __ldg(reinterpret_cast<const uint4*>(mrow + i * 8));
  }
  tc::tmem_ld_wait();
  if (E.stack_off) {            // D = D_main + D_corr (hi*W_lo accumulated in its own columns by the N-stacked MMA)
    uint32_t w[kEpiBW];
    tc::tmem_ld16(cur_taddr + (uint32_t)E.stack_off, w);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < kEpiBW; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
  }
  if (next_taddr != 0u) tc::tmem_ld16(next_taddr, vn);
  else if (release_bar) {       // persistent kernel: the last block has left TMEM -> hand the accumulator set back to the MMA warp
    tc::tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tc::mbar_arrive(release_bar);
  }
  if (!valid) return;
  float f[kEpiBW];
#pragma unroll
  for (int i = 0; i < kEpiBW; ++i) f[i] = __uint_as_float(v[i]);
  if (E.bias) {   // packed bias is padded to n_pad (zeros beyond n_valid), 16-byte aligned; L1-resident after the first block
    const float4* bp = reinterpret_cast<const float4*>(E.bias + cfirst);
#pragma unroll
    for (int i = 0; i < kEpiBW / 4; ++i) {
      const float4 b = __ldg(bp + i);
      f[4 * i] += b.x; f[4 * i + 1] += b.y; f[4 * i + 2] += b.z; f[4 * i + 3] += b.w;
    }
  }
  if (ACT != ACT_NONE) {
#pragma unroll
    for (int i = 0; i < kEpiBW; ++i) f[i] = act_t<ACT>(f[i]);
  }
  if (MASK != ACT_NONE) {
    if (mask_vec) {
#pragma unroll
      for (int i = 0; i < kEpiBW / 8; ++i) {
        const uint4 q = mq[i];
        const int o = 8 * i;
        f[o + 0] *= mask_t<MASK>(__uint_as_float(q.x << 16)); f[o + 1] *= mask_t<MASK>(__uint_as_float(q.x & 0xffff0000u));
        f[o + 2] *= mask_t<MASK>(__uint_as_float(q.y << 16)); f[o + 3] *= mask_t<MASK>(__uint_as_float(q.y & 0xffff0000u));
        f[o + 4] *= mask_t<MASK>(__uint_as_float(q.z << 16)); f[o + 5] *= mask_t<MASK>(__uint_as_float(q.z & 0xffff0000u));
        f[o + 6] *= mask_t<MASK>(__uint_as_float(q.w << 16)); f[o + 7] *= mask_t<MASK>(__uint_as_float(q.w & 0xffff0000u));
      }
    } else {
#pragma unroll
      for (int i = 0; i < kEpiBW; ++i)
        if (cfirst + i < E.n_valid) f[i] *= mask_t<MASK>(__bfloat162float(mrow[i]));
    }
  }
  if (OUT_F32) {
    float* o = (float*)E.out + opix * E.out_ld + cfirst;
#pragma unroll
    for (int i = 0; i < kEpiBW; i += 4) {
      if (vec_ok && cfirst + i + 4 <= E.n_valid) {
        *reinterpret_cast<float4*>(o + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (cfirst + i + q < E.n_valid) o[i + q] = f[i + q];
      }
    }
  } else {
    bf16* o = (bf16*)E.out + opix * E.out_ld + cfirst;
#pragma unroll
    for (int i = 0; i < kEpiBW; i += 8) {
      if (vec_ok && cfirst + i + 8 <= E.n_valid) {
        uint4 pk;
        pk.x = pack_bf16x2(f[i], f[i + 1]); pk.y = pack_bf16x2(f[i + 2], f[i + 3]);
        pk.z = pack_bf16x2(f[i + 4], f[i + 5]); pk.w = pack_bf16x2(f[i + 6], f[i + 7]);
        *reinterpret_cast<uint4*>(o + i) = pk;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (cfirst + i + q < E.n_valid) o[i + q] = __float2bfloat16_rn(f[i + q]);
      }
    }
    if (E.out_lo) {     // bf16 pair: lo = bf16(v - hi)
      bf16* ol = E.out_lo + opix * E.out_ld + cfirst;
#pragma unroll
      for (int i = 0; i < kEpiBW; ++i) f[i] -= round_bf16(f[i]);
#pragma unroll
      for (int i = 0; i < kEpiBW; i += 8) {
        if (vec_ok && cfirst + i + 8 <= E.n_valid) {
          uint4 pk;
          pk.x = pack_bf16x2(f[i], f[i + 1]); pk.y = pack_bf16x2(f[i + 2], f[i + 3]);
          pk.z = pack_bf16x2(f[i + 4], f[i + 5]); pk.w = pack_bf16x2(f[i + 6], f[i + 7]);
          *reinterpret_cast<uint4*>(ol + i) = pk;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (cfirst + i + q < E.n_valid) ol[i + q] = __float2bfloat16_rn(f[i + q]);
        }
      }
    }
  }
}

// Drains `n_acc` accumulators (each tile_cols columns - a multiple of 16 -, `acc_stride` TMEM columns apart) of this thread's TMEM
// lane.  Accumulator a belongs to output pixel opix0 + (a / mtx) * step_ty + (a % mtx) * step_tx (igemm: one accumulator; halo
// kernel: mtx x mty).  Blocks of 16 columns are software-pipelined through two register sets (see epi_block_t).
template <int ACT, int MASK, bool OUT_F32>
__device__ __forceinline__ void epilogue_acc_t(const EpiRegs& E, uint32_t tmem_lane, int n_acc, int acc_stride, int mtx, long long opix0,
                                               long long step_tx, long long step_ty, bool valid, int col_base, uint64_t* release_bar = nullptr) {
  const int cpb = E.tile_cols / kEpiBW;                // column blocks per accumulator
  const int nblk = n_acc * cpb;
  uint32_t va[kEpiBW], vb[kEpiBW];
  int cb = 0, tx = 0;                                  // current block: column block cb of the accumulator at (.., tx)
  uint32_t taddr = tmem_lane;                          // TMEM address of the current block
  long long opix = opix0;
  const uint32_t acc_skip = (uint32_t)(acc_stride - (cpb - 1) * kEpiBW);   // last block of an accumulator -> first of the next
  tc::tmem_ld16(taddr, va);
  for (int j = 0; j < nblk; j += 2) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (j + half >= nblk) break;
      const bool last_cb = cb + 1 == cpb;
      const uint32_t next = taddr + (last_cb ? acc_skip : (uint32_t)kEpiBW);
      const uint32_t next_taddr = j + half + 1 < nblk ? next : 0u;
      if (half == 0) epi_block_t<ACT, MASK, OUT_F32>(E, va, col_base + cb * kEpiBW, opix, valid, next_taddr, vb, release_bar, taddr);
      else epi_block_t<ACT, MASK, OUT_F32>(E, vb, col_base + cb * kEpiBW, opix, valid, next_taddr, va, release_bar, taddr);
      taddr = next;
      if (last_cb) {
        cb = 0;
        if (++tx == mtx) { tx = 0; opix += step_ty - (long long)(mtx - 1) * step_tx; } else opix += step_tx;
      } else ++cb;
    }
  }
}

// warp-uniform dispatch on (activation of this column tile, mask kind, output type)
struct EpiSel { int act, mask, f32; };
__device__ __forceinline__ EpiSel epilogue_select(const TcLaunch& P, int n_tile) {
  int c = n_tile * P.tile_cols, j = 0;
  while (j + 1 < P.nparts && c >= P.part_n[j]) { c -= P.part_n[j]; ++j; }
  return EpiSel{P.part_act[j], P.mask_act, P.out_f32};
}
#define SV_EPI_CALL(A, M, F) epilogue_acc_t<A, M, F>(E, tmem_lane, n_acc, acc_stride, mtx, opix0, step_tx, step_ty, valid, col_base, release_bar)
__device__ __forceinline__ void epilogue_dispatch(const EpiRegs& E, EpiSel e, uint32_t tmem_lane, int n_acc, int acc_stride, int mtx, long long opix0,
                                                  long long step_tx, long long step_ty, bool valid, int col_base, uint64_t* release_bar = nullptr) {
  if (e.mask != ACT_NONE) {            // dgrad: linear, bf16 out
    if (e.mask == ACT_RELU) SV_EPI_CALL(ACT_NONE, ACT_RELU, false);
    else if (e.mask == ACT_ELU) SV_EPI_CALL(ACT_NONE, ACT_ELU, false);
    else SV_EPI_CALL(ACT_NONE, ACT_SOFTPLUS, false);
  } else if (e.f32) {
    if (e.act == ACT_NONE) SV_EPI_CALL(ACT_NONE, ACT_NONE, true);
    else if (e.act == ACT_SOFTPLUS) SV_EPI_CALL(ACT_SOFTPLUS, ACT_NONE, true);
    else if (e.act == ACT_ELU) SV_EPI_CALL(ACT_ELU, ACT_NONE, true);
    else SV_EPI_CALL(ACT_RELU, ACT_NONE, true);
  } else {
    if (e.act == ACT_RELU) SV_EPI_CALL(ACT_RELU, ACT_NONE, false);
    else if (e.act == ACT_ELU) SV_EPI_CALL(ACT_ELU, ACT_NONE, false);
    else if (e.act == ACT_SOFTPLUS) SV_EPI_CALL(ACT_SOFTPLUS, ACT_NONE, false);
    else SV_EPI_CALL(ACT_NONE, ACT_NONE, false);
  }
}
#undef SV_EPI_CALL

// (FAT is a template parameter: a run-time flag inside the producer / issue loops cost the plain path 30 % - d2 dgrad 20.5 -> 26.6 us)
template <bool FAT>
__device__ __forceinline__ void igemm_body_t(const TcLaunch& P, const int zsplit) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_bytes = 128 * P.bk * 2;
  const int b_bytes = (P.tile_cols * P.bk * 2 + 1023) & ~1023;
  constexpr bool fat = FAT;
  const int stage_bytes = fat ? 2 * (a_bytes + b_bytes) : a_bytes + b_bytes;
  SmemCtl* ctl = reinterpret_cast<SmemCtl*>(smem + (size_t)P.stages * stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y;
  const int tiles_per_img = P.grid_h / P.tile_h;   // tile_w == grid_w
  const int n0 = (m_tile / tiles_per_img) * P.tile_n_img;
  const int y0 = (m_tile % tiles_per_img) * P.tile_h;
  const int units_per_tap = fat ? P.kc : P.kcl;
  const int kb_total = P.taps_h * P.taps_w * units_per_tap;                       // logical k-blocks (see TcLaunch::split); fat: (tap, chunk) units
  const int kb_first = P.k_splits > 1 ? zsplit * P.kb_per_split : 0;              // split-K: this CTA's k-block range
  const int num_kb = P.k_splits > 1 ? min(P.kb_per_split, kb_total - kb_first) : kb_total;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map_a);
    tc::prefetch_tmap(&P.map_b);
    if (P.kca > P.kc) tc::prefetch_tmap(&P.map_a_lo);
    for (int i = 0; i < P.stages; ++i) { tc::mbar_init(&ctl->full[i], 1); tc::mbar_init(&ctl->empty[i], 1); }
    tc::mbar_init(&ctl->tmem_full, 1);
    tc::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)P.tile_cols) tmem_cols <<= 1;
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (tc::elect_one()) {
      for (int it = 0; it < num_kb; ++it) {
        const int stage = it % P.stages, phase = (it / P.stages) & 1;
        tc::mbar_wait(&ctl->empty[stage], phase ^ 1);
        const int kb = kb_first + it;
        if constexpr (FAT) {
          const int tap = kb / P.kc, ch = kb - tap * P.kc;
          const int ta = tap / P.taps_w, tb = tap - ta * P.taps_w;
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          const int wb = P.tile_cols * P.bk * 2;
          tc::mbar_expect_tx(&ctl->full[stage], 2 * a_bytes + 2 * wb);
          tc::tma_load_4d(sa, &P.map_a, &ctl->full[stage], ch * P.bk, tb - P.pad_l, y0 * P.a_stride + ta - P.pad_t, n0);
          tc::tma_load_4d(sa + a_bytes, &P.map_a_lo, &ctl->full[stage], ch * P.bk, tb - P.pad_l, y0 * P.a_stride + ta - P.pad_t, n0);
          tc::tma_load_2d(sa + 2 * a_bytes, &P.map_b, &ctl->full[stage], (tap * P.kcb + 2 * ch) * P.bk, n_tile * P.tile_cols);
          tc::tma_load_2d(sa + 2 * a_bytes + b_bytes, &P.map_b, &ctl->full[stage], (tap * P.kcb + 2 * ch + 1) * P.bk, n_tile * P.tile_cols);
        } else {
        const int tap = kb / P.kcl, chunk = kb - tap * P.kcl;                 // logical chunk -> physical A chunk / weight k-block
        int ap, bp;
        logical_chunk(P.split, P.kc, chunk, ap, bp);
        const bool a_is_lo = ap >= P.kc;
        const int ta = tap / P.taps_w, tb = tap - ta * P.taps_w;
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        tc::mbar_expect_tx(&ctl->full[stage], a_bytes + P.tile_cols * P.bk * 2);
        tc::tma_load_4d(sa, a_is_lo ? &P.map_a_lo : &P.map_a, &ctl->full[stage], (a_is_lo ? ap - P.kc : ap) * P.bk, tb - P.pad_l,
                        y0 * P.a_stride + ta - P.pad_t, n0);
        tc::tma_load_2d(sa + a_bytes, &P.map_b, &ctl->full[stage], (tap * P.kcb + bp) * P.bk, n_tile * P.tile_cols);
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, P.tile_cols, 0, 0);
      const uint32_t lt = tc::layout_type_for(P.swizzle);
      const uint32_t sbo = 16u * P.bk;  // 8 rows x (bk*2) bytes
      const uint64_t tmpl = tc::make_smem_desc(0, 16, sbo, lt);   // + start address >> 4 in bits 0-13
      const int ksteps = P.bk / 16;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int stage = kb % P.stages, phase = (kb / P.stages) & 1;
        tc::mbar_wait(&ctl->full[stage], phase);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + (size_t)stage * stage_bytes);
        if constexpr (FAT) {      // hi*hi + lo*hi + hi*lo from one stage
          const uint64_t dah = tmpl + (sa >> 4), dal = tmpl + ((sa + a_bytes) >> 4);
          const uint64_t dbh = tmpl + ((sa + 2 * a_bytes) >> 4), dbl = tmpl + ((sa + 2 * a_bytes + b_bytes) >> 4);
          for (int k = 0; k < ksteps; ++k) {
            tc::umma_bf16(tmem_base, dah + 2u * k, dbh + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
            tc::umma_bf16(tmem_base, dal + 2u * k, dbh + 2u * k, idesc, 1u);
            tc::umma_bf16(tmem_base, dah + 2u * k, dbl + 2u * k, idesc, 1u);
          }
        } else {
        const uint64_t da = tmpl + (sa >> 4), db = tmpl + ((sa + a_bytes) >> 4);
        tc::umma_bf16(tmem_base, da, db, idesc, kb != 0);
        if (ksteps > 1) tc::umma_bf16(tmem_base, da + 2, db + 2, idesc, 1u);
        if (ksteps > 2) { tc::umma_bf16(tmem_base, da + 4, db + 4, idesc, 1u); tc::umma_bf16(tmem_base, da + 6, db + 6, idesc, 1u); }
        }
        tc::umma_commit(&ctl->empty[stage]);   // frees this smem stage once the MMAs above have read it
      }
      tc::umma_commit(&ctl->tmem_full);        // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> global -----------------------------------------
    const EpiRegs E = load_epi_regs(P);         // (before the wait: these loads overlap the main loop)
    const EpiSel esel = epilogue_select(P, n_tile);
    tc::mbar_wait(&ctl->tmem_full, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3;               // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    if (P.k_splits > 1) {                       // split-K: raw fp32 partial, finished by splitk_finish_kernel
      float* dst = P.partial + ((size_t)zsplit * P.m_pad + (size_t)m_tile * 128 + row) * P.n_pad + (size_t)n_tile * P.tile_cols;
      for (int c0 = 0; c0 < P.tile_cols; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        const int ncol = min(32, P.tile_cols - c0);
        if (ncol >= 32) tc::tmem_ld32(taddr, v); else tc::tmem_ld16(taddr, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (i < ncol) *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    } else {
    const int per_img = P.tile_h * P.tile_w;
    const int nn = row / per_img, rem = row - nn * per_img;
    const int hh = rem / P.tile_w, ww = rem - hh * P.tile_w;
    const int n = n0 + nn, y = y0 + hh, x = ww;
    const bool valid = n < P.n_img;
    const long long opix = ((long long)n * P.OH + (y * P.osy + P.ooy)) * P.OW + (x * P.osx + P.oox);
    epilogue_dispatch(E, esel, tmem_base + ((uint32_t)(quarter * 32) << 16), 1, 0, 1, opix, 0, 0, valid, n_tile * P.tile_cols);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

__device__ __forceinline__ void igemm_body(const TcLaunch& P, const int zsplit) { igemm_body_t<false>(P, zsplit); }
__global__ void __launch_bounds__(kThreads, 3) igemm_kernel(const __grid_constant__ TcLaunch P) { pdl_enter(); igemm_body_t<false>(P, blockIdx.z); }
// bf16x3 forward with fat ring stages (TcLaunch::fat)
__global__ void __launch_bounds__(kThreads, 3) igemm_fat_kernel(const __grid_constant__ TcLaunch P) { pdl_enter(); igemm_body_t<true>(P, blockIdx.z); }

// The s*s parity classes of a stride-s dgrad (each a small stride-1 convolution scattering into its own output parity)
// as ONE launch: blockIdx.z selects the class.  4 x more CTAs in flight for layers whose single class does not fill the GPU.
struct TcLaunch4 { TcLaunch l[4]; };
__global__ void __launch_bounds__(kThreads, 3) igemm4_kernel(const __grid_constant__ TcLaunch4 P4) { pdl_enter(); igemm_body(P4.l[blockIdx.z], 0); }

// Split-K finish for dense layers (one output row per image): out[row][col] = epilogue(sum_z partial[z][row][col]).
// Fixed summation order -> deterministic.  One thread per output element; consecutive threads = consecutive columns.
__global__ void __launch_bounds__(256) splitk_finish_kernel(const __grid_constant__ TcLaunch P) {
  pdl_enter();
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int col = (int)(idx % P.n_valid);
  const long long row = idx / P.n_valid;
  if (row >= P.n_img) return;
  const size_t zstride = (size_t)P.m_pad * P.n_pad;
  const float* p = P.partial + (size_t)row * P.n_pad + col;
  float acc = 0.f;
  for (int z = 0; z < P.k_splits; ++z) acc += p[z * zstride];
  if (P.bias) acc += P.bias[col];
  int c = col, j = 0;
  while (j + 1 < P.nparts && c >= P.part_n[j]) { c -= P.part_n[j]; ++j; }
  acc = apply_act(acc, P.part_act[j]);
  if (P.mask_act != ACT_NONE)
    acc *= act_grad_from_out(__bfloat162float(((const bf16*)P.mask_src)[row * P.mask_ld + P.mask_coff + col]), P.mask_act);
  const long long o = row * P.out_ld + col;
  if (P.out_f32) ((float*)P.out)[o] = acc;
  else {
    const bf16 hi = __float2bfloat16_rn(acc);
    ((bf16*)P.out)[o] = hi;
    if (P.out_lo) ((bf16*)P.out_lo)[o] = __float2bfloat16_rn(acc - __bfloat162float(hi));
  }
}


// ------------------------------------------------------------------------------------------------
// Halo-resident implicit GEMM (stride-1 convolutions: decoder forward layers, every dgrad).
//   * warp 0 loads the (TH+taps_h-1) x (TW+taps_w-1) input halo of a TW x TH output tile ONCE (one 4-D TMA box per
//     64/32/16-channel chunk; OOB zero fill = TF 'same' padding) and then streams the packed weights through a ring
//     (kb_per_stage 2-D TMA boxes of [tile_cols][chunk] K-major per stage, one mbarrier);
//   * the A operand of filter tap (a,b) for accumulator (tx,ty) is the SAME shared-memory halo addressed through a
//     shifted UMMA descriptor: start = halo + ((ty*16+a)*TWp + tx*8 + b) * pixel_bytes, SBO = TWp * pixel_bytes
//     (an M row group = 8 horizontally adjacent pixels, 16 groups = 16 image rows); the 128B/64B/32B swizzle is a
//     function of absolute shared-memory address bits for TMA and UMMA alike, so shifted views stay consistent;
//   * (TW/8)*(TH/16) accumulators of 128 x tile_cols live in TMEM; the epilogue drains them as in igemm_kernel.
// L2->SMEM traffic per output tile drops from taps*(A+B) to halo + weights (10-16x for the 6x6 layers).
// ------------------------------------------------------------------------------------------------
// MMAs of one k-block (tap, channel chunk) for MTX (compile-time) x mty accumulators: k-step outermost so that consecutive
// MMAs target different accumulators; the tile-column loop is unrolled and everything stays in (uniform) registers.
template <int MTX>
__device__ __forceinline__ void halo_issue_kb(uint64_t da0, uint64_t db0, uint32_t tmem_base, uint32_t tile_cols, uint32_t idesc, int mty, int ksteps,
                                              uint32_t ty_step, uint32_t tx_step, uint32_t accum0) {
  for (int k = 0; k < ksteps; ++k) {
    const uint64_t dbk = db0 + 2u * k;
    const uint32_t accum = k ? 1u : accum0;
    uint64_t da_row = da0 + 2u * k;
    uint32_t acc = tmem_base;
    for (int ty = 0; ty < mty; ++ty, da_row += ty_step, acc += MTX * tile_cols) {
#pragma unroll
      for (int tx = 0; tx < MTX; ++tx) tc::umma_bf16(acc + tx * tile_cols, da_row + tx * tx_step, dbk, idesc, accum);
    }
  }
}

// SV_HALO_TRACE=1: every halo CTA records the SM clock at its phase boundaries (sv_debug_halo_trace reads the buffer back):
//   [0] smid  [1] CTA start  [2] setup done  [3] halo landed (MMA warp)  [4] last MMA issued  [5] accumulators complete
//   (epilogue warp)  [6] epilogue done  [7] globaltimer at start
constexpr int kTraceSlots = 8, kTraceCtas = 8192;
__device__ unsigned long long g_halo_trace[kTraceSlots * kTraceCtas];
__device__ __forceinline__ void trace_mark(int on, int slot) {
  if (on) {
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta < kTraceCtas) g_halo_trace[cta * kTraceSlots + slot] = (unsigned long long)clock64();
  }
}

// bf16x3, N-stacked weight pairs: all MMAs of ONE filter tap (= one weight-ring stage: [W_hi(ch) W_lo(ch)] per channel chunk), fully
// unrolled - per chunk A_hi x [W_hi ; W_lo] (2N columns) then A_lo x W_hi (N columns), KS k-steps, MTX accumulators.  The generic
// nest spends ~150 cycles of single-thread issue per tcgen05.mma (phase trace, d4: 85 k cycles for 576 MMAs); this one a few.
template <int KS, int MTX, int KC>
__device__ __forceinline__ void halo_issue_tap_pairs(uint64_t da_hi, uint64_t db0, uint32_t acc0, uint32_t spacing, uint32_t idesc1, uint32_t idesc2,
                                                     uint32_t tx_step, uint32_t chunk_step, uint32_t lo_off, uint32_t blk_step, uint32_t first) {
#pragma unroll
  for (int ch = 0; ch < KC; ++ch) {
    const uint64_t da = da_hi + (uint64_t)(ch * chunk_step), db = db0 + (uint64_t)(2 * ch * blk_step);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
#pragma unroll
      for (int tx = 0; tx < MTX; ++tx)
        tc::umma_bf16(acc0 + tx * spacing, da + (uint64_t)(tx * tx_step + 2u * k), db + 2u * k, idesc2, (ch | k) != 0 ? 1u : first);
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
#pragma unroll
      for (int tx = 0; tx < MTX; ++tx)
        tc::umma_bf16(acc0 + tx * spacing, da + (uint64_t)(lo_off + tx * tx_step + 2u * k), db + 2u * k, idesc1, 1u);
    }
  }
}

struct HaloCtl {
  uint64_t halo_full;
  uint64_t w_full[kMaxStages];
  uint64_t w_empty[kMaxStages];
  uint64_t tmem_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ void halo_body(const TcLaunch& P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunks = P.kcl;                    // LOGICAL chunks per tap; the halo holds P.kca physical chunks (TcLaunch::split)
  const int nphys = P.kca;
  const int sx = P.halo_sx, sy = P.halo_sy;     // input stride: sx parity planes per chunk, rows sy apart (1 = plain stride-1 conv)
  uint8_t* halo = smem;
  uint8_t* wring = smem + (size_t)nphys * sx * P.chunk_bytes;
  HaloCtl* ctl = reinterpret_cast<HaloCtl*>(wring + (size_t)P.w_stages * P.w_stage_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  int t = blockIdx.x;
  const int tile_x = t % P.tiles_x; t /= P.tiles_x;
  const int tile_y = t % P.tiles_y;
  const int n = t / P.tiles_y;
  const int x0 = tile_x * P.TW, y0 = tile_y * P.TH;
  const int num_kb = P.taps_h * P.taps_w * P.kcb;                      // PHYSICAL weight k-blocks: each is streamed once per tile
  const int num_stages_total = (num_kb + P.kb_per_stage - 1) / P.kb_per_stage;
  const int MT = P.mtx * P.mty;
  // Every CTA streams the same weights: started together they all hit the same L2 lines at the same time (d4 bf16x3: 1.0 TB/s
  // L2->SMEM, 562 us).  Each CTA therefore starts the (commutative) K loop at its own ring stage.
  const int rot = P.split ? (int)((blockIdx.x * 11u + blockIdx.y * 5u) % (unsigned)num_stages_total) : 0;

  const int tr = P.trace;
  if (threadIdx.x == 0) {
    if (tr) {
      const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
      if (cta < kTraceCtas) {
        unsigned smid; unsigned long long gt;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g_halo_trace[cta * kTraceSlots + 0] = smid; g_halo_trace[cta * kTraceSlots + 7] = gt;
      }
      trace_mark(tr, 1);
    }
    tc::prefetch_tmap(&P.map_a);
    tc::prefetch_tmap(&P.map_b);
    if (P.kca > P.kc) tc::prefetch_tmap(&P.map_a_lo);
    tc::mbar_init(&ctl->halo_full, 1);
    for (int i = 0; i < P.w_stages; ++i) { tc::mbar_init(&ctl->w_full[i], 1); tc::mbar_init(&ctl->w_empty[i], 1); }
    tc::mbar_init(&ctl->tmem_full, 1);
    tc::fence_barrier_init();
  }
  const int acc_cols1 = P.nstack2 ? 2 * P.tile_cols : P.tile_cols;     // TMEM columns per accumulator ([main | correction] when N-stacked)
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(MT * acc_cols1)) tmem_cols <<= 1;
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int kb_bytes = P.tile_cols * P.bk * 2;   // one k-block of weights

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&ctl->halo_full, (uint32_t)(nphys * sx * P.THp * P.TWp * P.bk * 2));
      for (int p = 0; p < sx; ++p)         // plane p holds input columns sx*x0 - pad_l + p + sx*j (TMA element stride sx)
        for (int c = 0; c < nphys; ++c)    // physical chunks: the hi plane's chunks, then the lo plane's
          tc::tma_load_4d(halo + (size_t)(p * nphys + c) * P.chunk_bytes, c < P.kc ? &P.map_a : &P.map_a_lo, &ctl->halo_full,
                          (c < P.kc ? c : c - P.kc) * P.bk, sx * x0 - P.pad_l + p, sy * y0 - P.pad_t, n);
      const int w_stages = P.w_stages, kb_per_stage = P.kb_per_stage, w_stage_bytes = P.w_stage_bytes, bk = P.bk, col0 = n_tile * P.tile_cols;
      const CUtensorMap* map_b = &P.map_b;     // (registers: the TMA asm statements clobber memory)
      const int box3 = P.w_box3;
      for (int st = 0; st < num_stages_total; ++st) {
        const int slot = st % w_stages, phase = (st / w_stages) & 1;
        tc::mbar_wait(&ctl->w_empty[slot], phase ^ 1);
        const int se = st + rot < num_stages_total ? st + rot : st + rot - num_stages_total;
        const int kb0 = se * kb_per_stage;
        const int nkb = min(kb_per_stage, num_kb - kb0);
        tc::mbar_expect_tx(&ctl->w_full[slot], (uint32_t)(nkb * kb_bytes));
        if (box3) tc::tma_load_3d(wring + (size_t)slot * w_stage_bytes, map_b, &ctl->w_full[slot], 0, col0, kb0);   // the whole stage in one box
        else
          for (int j = 0; j < nkb; ++j)
            tc::tma_load_2d(wring + (size_t)slot * w_stage_bytes + (size_t)j * kb_bytes, map_b, &ctl->w_full[slot], (kb0 + j) * bk, col0);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      // every launch parameter the loop needs lives in a register: each tcgen05.mma carries a "memory" clobber, so a field read
      // through `P` inside the loop is re-loaded (long-scoreboard stall) after every MMA - 148 cycles per MMA on d4 bf16x3
      const int tile_n = P.tile_cols, bk = P.bk, TWp = P.TWp, taps_w = P.taps_w, mtx = P.mtx, mty = P.mty, kcb = P.kcb, kc = P.kc;
      const int split = P.split, nstack2 = P.nstack2, kb_per_stage = P.kb_per_stage, w_stages = P.w_stages;
      const uint32_t chunk_bytes = (uint32_t)P.chunk_bytes, w_stage_bytes = (uint32_t)P.w_stage_bytes;
      const uint32_t idesc1 = tc::make_idesc_bf16(128, tile_n, 0, 0), idesc2 = tc::make_idesc_bf16(128, 2 * tile_n, 0, 0);
      const uint32_t lt = tc::layout_type_for(P.swizzle);
      const uint32_t pix = (uint32_t)bk * 2u;                // bytes per pixel inside a chunk
      const uint32_t a_sbo = (uint32_t)(sy * TWp) * pix;     // next row group = next output row = sy input rows
      const uint32_t b_sbo = 8u * pix;                       // weights: dense [tile_cols][bk]
      const uint32_t halo_addr = tc::smem_u32(halo), wring_addr = tc::smem_u32(wring);
      trace_mark(tr, 2);
      tc::mbar_wait(&ctl->halo_full, 0);
      tc::tc_fence_after();
      trace_mark(tr, 3);
      // Descriptors differ only in their 14-bit start-address field (bits 0-13, units of 16 B): build one template per
      // operand and add offsets, so the single issuing thread spends a handful of instructions per MMA.
      const uint64_t a_tmpl = tc::make_smem_desc(0, 16, a_sbo, lt), b_tmpl = tc::make_smem_desc(0, 16, b_sbo, lt);
      const uint32_t tx_step = (8u * pix) >> 4, ty_step = (16u * (uint32_t)(sy * TWp) * pix) >> 4;
      const uint32_t tile_cols = (uint32_t)acc_cols1;      // accumulator spacing in TMEM
      const int ksteps = bk / 16;
      const int kb_inc = nstack2 ? 2 : 1;
      const uint32_t kb_step = (uint32_t)(kb_inc * kb_bytes) >> 4;
      const bool tap_unrolled = nstack2 && sx == 1 && mty == 1 && ksteps == 4 && (kb_per_stage % kcb) == 0 && (kc == 1 || kc == 2) &&
                                (mtx == 1 || mtx == 2 || mtx == 4) && !(P.trace & 2);
      const int taps_per_stage = kb_per_stage / kcb;
      uint32_t first = 0u;                  // 0 for the very first MMA group of the tile (overwrites the accumulators)
      for (int st = 0; st < num_stages_total; ++st) {
        const int slot = st % w_stages, phase = (st / w_stages) & 1;
        tc::mbar_wait(&ctl->w_full[slot], phase);
        tc::tc_fence_after();
        const int se = st + rot < num_stages_total ? st + rot : st + rot - num_stages_total;
        int kb = se * kb_per_stage;
        const int kb_end = min(num_kb, kb + kb_per_stage);
        uint32_t b_addr = (wring_addr + (uint32_t)slot * w_stage_bytes) >> 4;
        int tap = kb / kcb, jj = kb - tap * kcb;                        // physical weight block jj of filter tap (ta, tb)
        int ta = tap / taps_w, tb = tap - ta * taps_w;
        if (tap_unrolled)              // one stage = a whole number of taps (the planner made kb_per_stage a multiple of kcb)
         for (int tt = 0; tt < taps_per_stage; ++tt) {
          const uint64_t da_hi = a_tmpl + ((halo_addr + (uint32_t)(ta * TWp + tb) * pix) >> 4);
          const uint32_t cstep = chunk_bytes >> 4, lo_off = (uint32_t)kc * cstep, bstep = (uint32_t)kb_bytes >> 4;
          const uint64_t db0 = b_tmpl + b_addr + (uint32_t)(tt * kcb) * bstep;
          if (kc == 1) {
            if (mtx == 2) halo_issue_tap_pairs<4, 2, 1>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
            else if (mtx == 1) halo_issue_tap_pairs<4, 1, 1>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
            else halo_issue_tap_pairs<4, 4, 1>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
          } else {
            if (mtx == 2) halo_issue_tap_pairs<4, 2, 2>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
            else if (mtx == 1) halo_issue_tap_pairs<4, 1, 2>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
            else halo_issue_tap_pairs<4, 4, 2>(da_hi, db0, tmem_base, tile_cols, idesc1, idesc2, tx_step, cstep, lo_off, bstep, first);
          }
          first = 1u;
          kb = kb_end;
          if (++tb == taps_w) { tb = 0; ++ta; }
         }
        for (; kb < kb_end; kb += kb_inc, b_addr += kb_step) {
         int a0, a1;
         const int na = phys_block_chunks(split, kc, jj, a0, a1);
         // N-stacked pair: blocks (W_hi(ch), W_lo(ch)) are adjacent in the ring = ONE B operand of 2N rows:
         //   q = 0: A_hi x [W_hi ; W_lo] -> columns [main | correction],  q = 1: A_lo x W_hi -> main
         for (int q = 0; q < na; ++q, first = 1u) {
          const uint32_t idesc = (nstack2 && q == 0) ? idesc2 : idesc1;
          // filter column tb = sx * b' + p: parity plane p, shifted by b' plane columns
          const int ap = q ? a1 : a0;          // physical halo chunk multiplied with this weight block
          const uint32_t a_tap = sx == 1 ? (halo_addr + (uint32_t)ap * chunk_bytes + (uint32_t)(ta * TWp + tb) * pix) >> 4
                                         : (halo_addr + (uint32_t)((tb % sx) * nphys + ap) * chunk_bytes + (uint32_t)(ta * TWp + tb / sx) * pix) >> 4;
          const uint64_t db = b_tmpl + b_addr;
          switch (mtx) {
            case 1: halo_issue_kb<1>(a_tmpl + a_tap, db, tmem_base, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            case 2: halo_issue_kb<2>(a_tmpl + a_tap, db, tmem_base, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            case 4: halo_issue_kb<4>(a_tmpl + a_tap, db, tmem_base, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            case 8: halo_issue_kb<8>(a_tmpl + a_tap, db, tmem_base, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            default:
              for (int k = 0; k < ksteps; ++k) {
                uint32_t a_row = a_tap + 2u * k, acc = tmem_base;
                const uint64_t dbk = db + 2u * k;
                const uint32_t accum = k ? 1u : first;
                for (int ty = 0; ty < mty; ++ty, a_row += ty_step) {
                  uint64_t da = a_tmpl + a_row;
                  for (int tx = 0; tx < mtx; ++tx, da += tx_step, acc += tile_cols) tc::umma_bf16(acc, da, dbk, idesc, accum);
                }
              }
          }
         }
         jj += kb_inc;
         if (jj >= kcb) { jj = 0; if (++tb == taps_w) { tb = 0; ++ta; } }
        }
        tc::umma_commit(&ctl->w_empty[slot]);
      }
      tc::umma_commit(&ctl->tmem_full);
      trace_mark(tr, 4);
    }
  } else {
    const EpiRegs E = load_epi_regs(P);         // (before the wait: these loads overlap the main loop)
    const EpiSel esel = epilogue_select(P, n_tile);
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const int y = y0 + (row >> 3), x = x0 + (row & 7);             // accumulator (0,0): row group = image row, 8 pixels wide
    const long long opix0 = ((long long)n * P.OH + (y * P.osy + P.ooy)) * P.OW + (x * P.osx + P.oox);
    const long long step_tx = 8LL * P.osx, step_ty = 16LL * P.osy * P.OW;
    const int mtx = P.mtx, n_acc = MT, acc_stride = acc_cols1, col_base = n_tile * P.tile_cols;
    tc::mbar_wait(&ctl->tmem_full, 0);
    tc::tc_fence_after();
    if (threadIdx.x == 64) trace_mark(tr, 5);
    epilogue_dispatch(E, esel, tmem_base + ((uint32_t)(quarter * 32) << 16), n_acc, acc_stride, mtx, opix0, step_tx, step_ty, true, col_base);
    if (threadIdx.x == 64) trace_mark(tr, 6);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void __launch_bounds__(kThreads, 3) halo_conv_kernel(const __grid_constant__ TcLaunch P) { pdl_enter(); halo_body(P); }
// the 4 parity classes of a stride-2 dgrad as one launch (blockIdx.z = class), as igemm4_kernel
__global__ void __launch_bounds__(kThreads, 3) halo4_kernel(const __grid_constant__ TcLaunch4 P4) { pdl_enter(); halo_body(P4.l[blockIdx.z]); }

// ------------------------------------------------------------------------------------------------
// Persistent, pipelined variant of the halo convolution (same operand addressing as halo_body).
// The one-tile-per-CTA kernel above runs its phases back to back and co-resident CTAs fall into lockstep (per-CTA phase trace,
// d5 dgrad: 900 setup + 2400 halo wait + 23500 MMA phase shared by 3 CTAs + 4000 epilogue of a 31400-cycle lifetime): the tensor
// pipe idles a third of the time and every CTA re-fetches the weights.  Here one CTA per SM keeps the packed weights resident,
// streams halos through an `p_stages`-deep ring and double-buffers the accumulators in TMEM:
//   warp 0      TMA producer: weights once, then one halo per tile
//   warp 1      single-thread MMA issuer: tile i -> accumulator set i % 2
//   warps 2-5   epilogue of the even tiles (set 0),  warps 6-9: odd tiles (set 1); a set hands its accumulators back as soon as
//               its last TMEM read has completed, so load(i+1), MMA(i) and the epilogues of i-1 / i-2 overlap.
// blockIdx.y selects one of up to 4 launches (the parity classes of a stride-2 dgrad) that share the geometry.
// ------------------------------------------------------------------------------------------------
constexpr int kPcThreads = 64 + 8 * 32;
constexpr int kPcMaxStages = 6;
struct PcCtl {
  uint64_t w_full, halo_full[kPcMaxStages], halo_empty[kPcMaxStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// MMAs of one tile, fully unrolled for the shapes of this model family (one channel chunk, unit column stride, one accumulator
// row): KH x KW taps x KS k-steps x MTX accumulators.  The single issuing thread must spend only a few instructions per
// tcgen05.mma - the generic nest costs ~60 per k-block and capped d5 dgrad at 85 cycles per MMA (44 is the pipe's own rate).
// NB: weight k-blocks per tap that multiply the SAME A operand (2 for the bf16x3 first layer: [Whi Whi 0 0] and [Wlo 0 0 0] against
// the staged pixel [x_hi x_lo 0 0])
template <int KH, int KW, int KS, int MTX, int SX = 1, int NB = 1>
__device__ __forceinline__ void pc_issue_tile(uint64_t da0, uint64_t db0, uint32_t acc, uint32_t tile_cols, uint32_t idesc, uint32_t row_step,
                                              uint32_t pix_step, uint32_t kb_step, uint32_t tx_step, uint32_t plane_step = 0) {
#pragma unroll
  for (int a = 0; a < KH; ++a) {
#pragma unroll
    for (int b = 0; b < KW; ++b) {
      // (SX parity planes: filter column b lives in plane b % SX, shifted by b / SX plane columns)
      const uint64_t da = da0 + (uint64_t)(a * row_step + (b / SX) * pix_step + (b % SX) * plane_step);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const uint64_t db = db0 + (uint64_t)(((a * KW + b) * NB + j) * kb_step);
#pragma unroll
        for (int k = 0; k < KS; ++k) {
#pragma unroll
          for (int tx = 0; tx < MTX; ++tx)
            tc::umma_bf16(acc + tx * tile_cols, da + (uint64_t)(tx * tx_step + 2u * k), db + 2u * k, idesc, (a | b | j | k) != 0 ? 1u : 0u);
        }
      }
    }
  }
}

template <int ACT, int MASK, bool OUT_F32>
__device__ __noinline__ void pconv_epilogue_t(const TcLaunch& P, PcCtl* ctl, uint32_t tmem_base, int set, int quarter, int lane) {
  const int row = quarter * 32 + lane;
  const EpiRegs E = load_epi_regs(P);
  const long long step_tx = 8LL * P.osx, step_ty = 16LL * P.osy * P.OW;
  const int mtx = P.mtx, MT = P.mtx * P.mty, TW = P.TW, TH = P.TH, tiles_x = P.tiles_x, OH = P.OH, OW = P.OW;
  const int osy = P.osy, ooy = P.ooy, osx = P.osx, oox = P.oox;
  const int acc_stride = P.tile_cols, acc_cols = MT * P.tile_cols, col_base = P.p_ntile * P.tile_cols;
  const int tiles_per_img = P.tiles_x * P.tiles_y, tiles = tiles_per_img * P.n_img;
  const uint32_t tmem_lane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(set * acc_cols);
  uint64_t* full = &ctl->acc_full[set];
  uint64_t* empty = &ctl->acc_empty[set];
  int i = set;
  for (int t = blockIdx.x + set * gridDim.x; t < tiles; t += 2 * gridDim.x, i += 2) {
    const int aph = (i >> 1) & 1;
    const int n = t / tiles_per_img, r = t - n * tiles_per_img;
    const int tile_y = r / tiles_x, tile_x = r - tile_y * tiles_x;
    const int y = tile_y * TH + (row >> 3), x = tile_x * TW + (row & 7);
    const long long opix0 = ((long long)n * OH + (y * osy + ooy)) * OW + (x * osx + oox);
    tc::mbar_wait(full, aph);
    tc::tc_fence_after();
    epilogue_acc_t<ACT, MASK, OUT_F32>(E, tmem_lane, MT, acc_stride, mtx, opix0, step_tx, step_ty, true, col_base, empty);
  }
}

__device__ __forceinline__ void pconv_body(const TcLaunch& P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nchunks = P.kcl, nphys = P.kca, sx = P.halo_sx, sy = P.halo_sy, nst = P.p_stages;   // logical / physical chunks (TcLaunch::split)
  const int stage_bytes = nphys * sx * P.chunk_bytes;
  uint8_t* wsm = smem;
  uint8_t* halo = smem + P.p_wbytes;
  PcCtl* ctl = reinterpret_cast<PcCtl*>(halo + (size_t)nst * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, ncta = gridDim.x;
  const int num_kb = P.taps_h * P.taps_w * nchunks;
  const int MT = P.mtx * P.mty;
  const int acc_cols = MT * P.tile_cols;
  const int tiles_per_img = P.tiles_x * P.tiles_y;
  const int tiles = tiles_per_img * P.n_img;
  const int kb_bytes = P.tile_cols * P.bk * 2;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map_a);
    tc::prefetch_tmap(&P.map_b);
    if (P.kca > P.kc) tc::prefetch_tmap(&P.map_a_lo);
    tc::mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < nst; ++i) { tc::mbar_init(&ctl->halo_full[i], 1); tc::mbar_init(&ctl->halo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&ctl->acc_full[i], 1); tc::mbar_init(&ctl->acc_empty[i], 4); }
    tc::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(2 * acc_cols)) tmem_cols <<= 1;
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (tc::elect_one()) {
      const int num_kb_phys = P.taps_h * P.taps_w * P.kcb;          // resident weights: the PHYSICAL k-blocks
      tc::mbar_expect_tx(&ctl->w_full, (uint32_t)(num_kb_phys * kb_bytes));
      for (int kb = 0; kb < num_kb_phys; ++kb)
        tc::tma_load_2d(wsm + (size_t)kb * kb_bytes, &P.map_b, &ctl->w_full, kb * P.bk, P.p_ntile * P.tile_cols);
      const uint32_t halo_tx = (uint32_t)(nphys * sx * P.THp * P.TWp * P.bk * 2);
      int i = 0;
      for (int t = cta; t < tiles; t += ncta, ++i) {
        const int st = i % nst, ph = (i / nst) & 1;
        const int n = t / tiles_per_img, r = t - n * tiles_per_img;
        const int tile_y = r / P.tiles_x, tile_x = r - tile_y * P.tiles_x;
        const int x0 = tile_x * P.TW, y0 = tile_y * P.TH;
        tc::mbar_wait(&ctl->halo_empty[st], ph ^ 1);
        tc::mbar_expect_tx(&ctl->halo_full[st], halo_tx);
        for (int p = 0; p < sx; ++p)
          for (int c = 0; c < nphys; ++c)
            tc::tma_load_4d(halo + (size_t)st * stage_bytes + (size_t)(p * nphys + c) * P.chunk_bytes, c < P.kc ? &P.map_a : &P.map_a_lo,
                            &ctl->halo_full[st], (c < P.kc ? c : c - P.kc) * P.bk, sx * x0 - P.pad_l + p, sy * y0 - P.pad_t, n);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, P.tile_cols, 0, 0);
      const uint32_t lt = tc::layout_type_for(P.swizzle);
      const uint32_t pix = (uint32_t)P.bk * 2u;
      const uint32_t a_sbo = (uint32_t)(sy * P.TWp) * pix, b_sbo = 8u * pix;
      const uint64_t a_tmpl = tc::make_smem_desc(0, 16, a_sbo, lt), b_tmpl = tc::make_smem_desc(0, 16, b_sbo, lt);
      const uint32_t tx_step = (8u * pix) >> 4, ty_step = (16u * (uint32_t)(sy * P.TWp) * pix) >> 4;
      const uint32_t tile_cols = (uint32_t)P.tile_cols, chunk_bytes = (uint32_t)P.chunk_bytes;
      const int ksteps = P.bk / 16, mtx = P.mtx, mty = P.mty, taps_w = P.taps_w, TWp = P.TWp;
      const int split_kind = P.split, kc_phys = P.kc, kcb = P.kcb;
      const uint32_t halo_addr = tc::smem_u32(halo), w_addr = tc::smem_u32(wsm) >> 4, kb_step = (uint32_t)kb_bytes >> 4;
      const bool env_shape_off = (P.trace & 2) != 0;      // SV_HALO_TRACE=2: generic issue loop (A/B)
      // unrolled issue sequences: 1 = 6x6 taps, K 16, 4 accumulators (d5 dgrad); 2 = 6x3 pair taps (first layer, 64x64 images);
      // 3 = 3x3 taps, K 64, 2 accumulators (e2 dgrad classes); 4 = first layer, 32x32 images
      // 5 / 6 = 6x6 stride-2 forward over two parity planes, K 32, one / two accumulators (e2 forward)
      // 8 / 9 = shapes 2 / 4 of the bf16x3 first layer: two weight k-blocks per pair tap against the same staged pixel pair
      const int shape = (P.split == 2 && nphys == 1 && mty == 1 && sx == 1 && !env_shape_off && P.taps_h == 6 && taps_w == 3 && ksteps == 1 &&
                         (mtx == 4 || mtx == 2)) ? (mtx == 4 ? 8 : 9)
                        : (nchunks != 1 || mty != 1 || env_shape_off) ? 0
                        : sx == 2 ? ((P.taps_h == 6 && taps_w == 6 && ksteps == 2 && mtx == 1) ? 5
                                     : (P.taps_h == 6 && taps_w == 6 && ksteps == 2 && mtx == 2) ? 6 : 0)
                        : sx != 1 ? 0
                        : (P.taps_h == 6 && taps_w == 6 && ksteps == 1 && mtx == 4) ? 1
                        : (P.taps_h == 6 && taps_w == 3 && ksteps == 1 && mtx == 4) ? 2
                        : (P.taps_h == 3 && taps_w == 3 && ksteps == 4 && mtx == 2) ? 3
                        : (P.taps_h == 6 && taps_w == 3 && ksteps == 1 && mtx == 2) ? 4
                        : (P.taps_h == 6 && taps_w == 3 && ksteps == 4 && mtx == 1) ? 7 : 0;   // 7 = e2 forward in the pixel-pair view
      tc::mbar_wait(&ctl->w_full, 0);
      int i = 0;
      for (int t = cta; t < tiles; t += ncta, ++i) {
        const int st = i % nst, ph = (i / nst) & 1;
        const int ab = i & 1, aph = (i >> 1) & 1;
        tc::mbar_wait(&ctl->halo_full[st], ph);
        tc::mbar_wait(&ctl->acc_empty[ab], aph ^ 1);
        tc::tc_fence_after();
        const uint32_t h_addr = halo_addr + (uint32_t)st * (uint32_t)stage_bytes;
        const uint32_t acc = tmem_base + (uint32_t)(ab * acc_cols);
        uint32_t b_addr = w_addr;
        int ta = 0, tb = 0, chunk = 0;
        if (shape) {
          const uint64_t da0 = a_tmpl + (h_addr >> 4), db0 = b_tmpl + w_addr;
          const uint32_t row_step = ((uint32_t)TWp * pix) >> 4, pix_step = pix >> 4;
          if (shape == 1) pc_issue_tile<6, 6, 1, 4>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 2) pc_issue_tile<6, 3, 1, 4>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 3) pc_issue_tile<3, 3, 4, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 4) pc_issue_tile<6, 3, 1, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 7) pc_issue_tile<6, 3, 4, 1>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 8) pc_issue_tile<6, 3, 1, 4, 1, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 9) pc_issue_tile<6, 3, 1, 2, 1, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step);
          else if (shape == 5) pc_issue_tile<6, 6, 2, 1, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step, chunk_bytes >> 4);
          else pc_issue_tile<6, 6, 2, 2, 2>(da0, db0, acc, tile_cols, idesc, row_step, pix_step, kb_step, tx_step, chunk_bytes >> 4);
        } else
        for (int kb = 0; kb < num_kb; ++kb) {
          int ap, bp;                                          // logical chunk -> physical halo chunk / weight k-block
          logical_chunk(split_kind, kc_phys, chunk, ap, bp);
          b_addr = w_addr + (uint32_t)((ta * taps_w + tb) * kcb + bp) * kb_step;
          const uint32_t a_tap = sx == 1 ? (h_addr + (uint32_t)ap * chunk_bytes + (uint32_t)(ta * TWp + tb) * pix) >> 4
                                         : (h_addr + (uint32_t)((tb % sx) * nphys + ap) * chunk_bytes + (uint32_t)(ta * TWp + tb / sx) * pix) >> 4;
          const uint64_t db = b_tmpl + b_addr;
          const uint32_t first = kb != 0;
          switch (mtx) {
            case 1: halo_issue_kb<1>(a_tmpl + a_tap, db, acc, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            case 2: halo_issue_kb<2>(a_tmpl + a_tap, db, acc, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            case 4: halo_issue_kb<4>(a_tmpl + a_tap, db, acc, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
            default: halo_issue_kb<8>(a_tmpl + a_tap, db, acc, tile_cols, idesc, mty, ksteps, ty_step, tx_step, first); break;
          }
          if (++chunk == nchunks) { chunk = 0; if (++tb == taps_w) { tb = 0; ++ta; } }
        }
        tc::umma_commit(&ctl->halo_empty[st]);
        tc::umma_commit(&ctl->acc_full[ab]);
      }
    }
  } else {
    const int set = (warp - 2) >> 2;            // accumulator set this warp drains (tiles i with i % 2 == set)
    const EpiSel e = epilogue_select(P, P.p_ntile);
    // one dispatch per thread with the tile loop INSIDE each instantiation (with the dispatch inside the loop the compiler
    // hoisted the loop invariants of all eleven variants at once: 168 registers and 20 KB of spill code)
#define SV_PC_CALL(A, M, F) pconv_epilogue_t<A, M, F>(P, ctl, tmem_base, set, warp & 3, lane)
    if (e.mask != ACT_NONE) {
      if (e.mask == ACT_RELU) SV_PC_CALL(ACT_NONE, ACT_RELU, false);
      else if (e.mask == ACT_ELU) SV_PC_CALL(ACT_NONE, ACT_ELU, false);
      else SV_PC_CALL(ACT_NONE, ACT_SOFTPLUS, false);
    } else if (e.f32) {
      if (e.act == ACT_NONE) SV_PC_CALL(ACT_NONE, ACT_NONE, true);
      else if (e.act == ACT_ELU) SV_PC_CALL(ACT_ELU, ACT_NONE, true);
      else SV_PC_CALL(ACT_RELU, ACT_NONE, true);
    } else {
      if (e.act == ACT_RELU) SV_PC_CALL(ACT_RELU, ACT_NONE, false);
      else if (e.act == ACT_ELU) SV_PC_CALL(ACT_ELU, ACT_NONE, false);
      else SV_PC_CALL(ACT_NONE, ACT_NONE, false);
    }
#undef SV_PC_CALL
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}
__global__ void __launch_bounds__(kPcThreads, 1) pconv_kernel(const __grid_constant__ TcLaunch4 P4) { pdl_enter(); pconv_body(P4.l[blockIdx.y]); }

// ------------------------------------------------------------------------------------------------
// N-stacked persistent convolution (see TcNsConv in tc_kernels.h).
//   warp 0   : TMA producer - resident weights once, then one (R + kh - 1)-row input halo per tile (2 stages)
//   warp 1   : single-thread tcgen05.mma issuer - kh * (C / 16) MMAs of N = kw * nb per tile (x ng N groups)
//   warps 2-5: epilogue - for every block of 8 output channels: kw TMEM loads, lane shuffles by (b - pad_l) along x
//              (W = 64: the 2-3 lanes that cross the two warps of an image row go through shared memory), bias,
//              activation / mask, NHWC store.  Accumulators are double-buffered so this overlaps the next tile's MMAs.
// ------------------------------------------------------------------------------------------------
constexpr int kNsMaxStages = 8;
constexpr int kNsMaxGroups = 4;                  // epilogue groups = accumulator buffers in TMEM
constexpr int kNsEpiWarps = 16;                  // epilogue warps per CTA = 4 TMEM quarters x groups x chunk lanes
constexpr int kNsThreads = 64 + 32 * kNsEpiWarps;
struct NsCtl {
  uint64_t w_full, halo_full[kNsMaxStages], halo_empty[kNsMaxStages], acc_full[kNsMaxGroups], acc_empty[kNsMaxGroups];
  uint32_t tmem_base;
};
constexpr int kNsXchSlots = 8;   // lanes 0..3 -> slots 0..3, lanes 28..31 -> slots 4..7

// Epilogue of the N-stacked kernel for one warp.  A single warp per scheduler runs this dependent shuffle / FMA chain at ~8
// cycles per instruction, so the work is spread over 16 warps: `groups` warp sets take whole tiles round-robin (group g owns
// accumulator buffer g), and inside a group `lanes` warps per TMEM quarter q split a tile's 8-channel blocks.  The accumulator
// is handed back to the MMA warp as soon as this warp's last block has been read from TMEM (before the shuffles and stores).
// Everything that shapes the instruction stream is a template parameter (the first, fully runtime version was ~1700 SASS
// instructions per 8-channel block and instruction-cache bound; this one is ~230).
template <int KW, int ACT, bool WIDE, bool OUT_F32>
__device__ __forceinline__ void ns_epilogue_t(const TcNsConv& P, NsCtl* ctl, float* xch_all, uint32_t tmem_base, int q, int grp, int cl, int lane) {
  // loop-invariant parameters in registers (the asm memory clobbers would otherwise force reloads through the generic pointer)
  const int W = P.W, H = P.H, R = P.R, nb = P.nb, pad_l = P.pad_l, tiles = P.tiles, tiles_per_img = P.tiles_per_img;
  const int groups = P.groups, lanes = P.lanes, n_valid = P.n_valid, out_ld = P.out_ld, mask_act = P.mask_act;
  const int mask_ld = P.mask_ld, mask_coff = P.mask_coff;
  const float* bias = P.bias;
  void* out = P.out;
  const bf16* mask_src = (const bf16*)P.mask_src;
  const int p = q * 32 + lane;
  const int r = p / W, x = p - r * W;
  const int nchunk8 = nb >> 3;
  const int mb = P.mb;                          // 128-pixel blocks per tile (each R rows), one accumulator per block
  const int acc1 = P.acc1, acc_cols = mb * acc1;
  const int pair = P.pair, corr_off = P.n_total;      // pair mode: correction columns (hi * W_lo) sit n_total columns after the main ones
  bf16* out_lo = (bf16*)P.out_lo;
  const int split = blockIdx.x % P.co_splits, cta = blockIdx.x / P.co_splits, ncta = gridDim.x / P.co_splits;
  const int cbase = split * nb;               // first output channel of this CTA's split
  const int my_last = cl + ((nchunk8 - 1 - cl) / lanes) * lanes;   // last block this warp reads (< 0: none)
  float* xch = xch_all + (size_t)(grp * lanes + cl) * (2 * 4 * KW * kNsXchSlots * 8);
  const int bar_id = 1 + (grp * lanes + cl) * 2 + (q >> 1);     // one named barrier per (warp set, image row of the tile)
  // per filter column b: source lane and whether the source pixel x + b - pad_l lies inside the image row (else: zero padding)
  float keep[KW];
  int src[KW];
#pragma unroll
  for (int b = 0; b < KW; ++b) {
    const int d = b - pad_l;
    keep[b] = (x + d >= 0 && x + d < W) ? 1.f : 0.f;
    src[b] = (lane + d) & 31;
  }
  int xbuf = 0, i = 0;
#ifdef SV_NS_TRACE
  const bool etr = (P.debug & 4) != 0 && q == 0 && grp == 0 && cl == 0;
#else
  constexpr bool etr = false;
#endif
  long long t_ewait = 0, t_ework = 0, tq = 0;
  for (int t = cta; t < tiles; t += ncta, ++i) {
    if (i % groups != grp) continue;
    const int aph = (i / groups) & 1;
    const int n = t / tiles_per_img, y0 = (t - n * tiles_per_img) * R * mb;
    if (etr) { const long long c = clock64(); if (i >= groups) t_ework += c - tq; tq = c; }
    tc::mbar_wait(&ctl->acc_full[grp], aph);
    if (etr) { const long long c = clock64(); t_ewait += c - tq; tq = c; }
    tc::tc_fence_after();
    if (cl >= nchunk8) {                          // more chunk lanes than blocks: nothing to read, but the barrier counts every warp
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&ctl->acc_empty[grp]);
      continue;
    }
    for (int mbi = 0; mbi < mb; ++mbi) {
    const long long opix = ((long long)n * H + (y0 + mbi * R + r)) * W + x;
    const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * acc_cols + mbi * acc1);
    for (int cc = cl; cc < nchunk8; cc += lanes) {
      uint32_t v[KW][8];                          // [filter column b][channel]
#pragma unroll
      for (int b = 0; b < KW; ++b) tc::tmem_ld8(tacc + (uint32_t)(b * nb + cc * 8), v[b]);
      tc::tmem_ld_wait();
      if (pair) {                                   // D = main + correction
        uint32_t w[KW][8];
#pragma unroll
        for (int b = 0; b < KW; ++b) tc::tmem_ld8(tacc + (uint32_t)(corr_off + b * nb + cc * 8), w[b]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < KW; ++b) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[b][j] = __float_as_uint(__uint_as_float(v[b][j]) + __uint_as_float(w[b][j]));
        }
      }
      if (cc == my_last && mbi == mb - 1) {       // this warp is done with the accumulators: hand them back to the MMA warp
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&ctl->acc_empty[grp]);
      }
      if (P.debug & 1) continue;
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
      float* xb = xch + (size_t)xbuf * (4 * KW * kNsXchSlots * 8);
      if (WIDE) {
        // an image row spans two warps (q even: left half, q odd: right half): lanes near the boundary publish their values
        // first, the in-warp shuffles below hide the latency of the pair barrier
        const int slot = lane < 4 ? lane : lane >= 28 ? lane - 24 : -1;
        if (slot >= 0) {
          float* xw = xb + (size_t)q * (KW * kNsXchSlots * 8);
#pragma unroll
          for (int b = 0; b < KW; ++b) {
            uint4* dst = reinterpret_cast<uint4*>(xw + (b * kNsXchSlots + slot) * 8);
            dst[0] = make_uint4(v[b][0], v[b][1], v[b][2], v[b][3]);
            dst[1] = make_uint4(v[b][4], v[b][5], v[b][6], v[b][7]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < KW; ++b) {
        const int ls = lane + b - pad_l;
        const float k = WIDE ? ((ls >= 0 && ls < 32) ? keep[b] : 0.f) : keep[b];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(k, __shfl_sync(0xffffffffu, __uint_as_float(v[b][j]), src[b]), o[j]);
      }
      if (WIDE) {
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");      // the two warps of this image row: partner's values visible
#pragma unroll
        for (int b = 0; b < KW; ++b) {
          const int ls = lane + b - pad_l;
          if (keep[b] != 0.f && (ls < 0 || ls > 31)) {   // source pixel lives in the neighbouring warp of this image row
            const int qn = ls < 0 ? q - 1 : q + 1;
            const int sl = ls < 0 ? ls + 8 : ls - 32;
            const float4* xr = reinterpret_cast<const float4*>(xb + (size_t)qn * (KW * kNsXchSlots * 8) + (b * kNsXchSlots + sl) * 8);
            const float4 lo = xr[0], hi = xr[1];
            o[0] += lo.x; o[1] += lo.y; o[2] += lo.z; o[3] += lo.w; o[4] += hi.x; o[5] += hi.y; o[6] += hi.z; o[7] += hi.w;
          }
        }
        xbuf ^= 1;
      }
      const int c0 = cbase + cc * 8;
      if (bias) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
        o[0] += b0.x; o[1] += b0.y; o[2] += b0.z; o[3] += b0.w; o[4] += b1.x; o[5] += b1.y; o[6] += b1.z; o[7] += b1.w;
      }
      if (ACT != ACT_NONE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = act_t<ACT>(o[j]);
      }
      if (mask_act != ACT_NONE) {
        const bf16* mrow = mask_src + opix * mask_ld + mask_coff + c0;
        for (int j = 0; j < 8; ++j)
          if (c0 + j < n_valid) o[j] *= act_grad_from_out(__bfloat162float(mrow[j]), mask_act);
      }
      if (OUT_F32) {
        float* dst = (float*)out + opix * out_ld + c0;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < n_valid) dst[j] = o[j];
      } else {
        bf16* dst = (bf16*)out + opix * out_ld + c0;
        if ((out_ld & 7) == 0 && c0 + 8 <= n_valid) {
          uint4 pk;
          pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]); pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(dst) = pk;
        } else {
          for (int j = 0; j < 8; ++j)
            if (c0 + j < n_valid) dst[j] = __float2bfloat16_rn(o[j]);
        }
        if (out_lo) {                                 // output pair: lo = bf16(v - hi)
          bf16* dl = out_lo + opix * out_ld + c0;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] -= round_bf16(o[j]);
          if ((out_ld & 7) == 0 && c0 + 8 <= n_valid) {
            uint4 pk;
            pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]); pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
            *reinterpret_cast<uint4*>(dl) = pk;
          } else {
            for (int j = 0; j < 8; ++j)
              if (c0 + j < n_valid) dl[j] = __float2bfloat16_rn(o[j]);
          }
        }
      }
    }
    }
  }
  if (etr && lane == 0 && blockIdx.x < kTraceCtas) {
    t_ework += clock64() - tq;
    g_halo_trace[blockIdx.x * kTraceSlots + 6] = (unsigned long long)t_ewait;
    g_halo_trace[blockIdx.x * kTraceSlots + 7] = (unsigned long long)t_ework;
  }
}

#define SV_NS_EPI_ARGS P, ctl, xch, tmem_base, q, grp, cl, lane
template <int KW, int ACT>
__device__ __forceinline__ void ns_epilogue_kw_act(const TcNsConv& P, NsCtl* ctl, float* xch, uint32_t tmem_base, int q, int grp, int cl, int lane) {
  if (P.W > 32) {
    if (P.out_f32) ns_epilogue_t<KW, ACT, true, true>(SV_NS_EPI_ARGS);
    else ns_epilogue_t<KW, ACT, true, false>(SV_NS_EPI_ARGS);
  } else {
    if (P.out_f32) ns_epilogue_t<KW, ACT, false, true>(SV_NS_EPI_ARGS);
    else ns_epilogue_t<KW, ACT, false, false>(SV_NS_EPI_ARGS);
  }
}
template <int KW>
__device__ __forceinline__ void ns_epilogue_kw(const TcNsConv& P, NsCtl* ctl, float* xch, uint32_t tmem_base, int q, int grp, int cl, int lane) {
  if (P.act == ACT_RELU) ns_epilogue_kw_act<KW, ACT_RELU>(SV_NS_EPI_ARGS);
  else if (P.act == ACT_ELU) ns_epilogue_kw_act<KW, ACT_ELU>(SV_NS_EPI_ARGS);
  else ns_epilogue_kw_act<KW, ACT_NONE>(SV_NS_EPI_ARGS);   // (the planner admits relu / elu / linear only)
}
__device__ __forceinline__ void ns_epilogue(const TcNsConv& P, NsCtl* ctl, float* xch, uint32_t tmem_base, int q, int grp, int cl, int lane) {
  if (P.kw == 6) ns_epilogue_kw<6>(SV_NS_EPI_ARGS);
  else ns_epilogue_kw<4>(SV_NS_EPI_ARGS);                  // (the planner admits kw 4 and 6 only)
}
#undef SV_NS_EPI_ARGS

// MMAs of one tile, fully unrolled: KH filter rows x KS k-steps of 16 channels x NG column groups.
// (NCH channel chunks per filter row: A chunk c sits chunk_step further, its weight k-block is (a * NCH + c); accum0 = 1: the first MMA
//  accumulates too - the lo-plane pass of the pair mode)
template <int KH, int KS, int NG, int NCH = 1>
__device__ __forceinline__ void ns_issue_tile(uint64_t da0, uint64_t db0, uint32_t acc, uint32_t idesc, uint32_t row_step, uint32_t wk_step,
                                              uint32_t grp_step, uint32_t ncols, uint32_t accum0 = 0u, uint32_t chunk_step = 0u) {
#pragma unroll
  for (int a = 0; a < KH; ++a) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const uint64_t da_a = da0 + (uint64_t)(a * row_step + c * chunk_step), db_a = db0 + (uint64_t)((a * NCH + c) * wk_step);
#pragma unroll
      for (int k = 0; k < KS; ++k) {
#pragma unroll
        for (int g = 0; g < NG; ++g)
          tc::umma_bf16(acc + g * ncols, da_a + 2u * k, db_a + (uint64_t)(g * grp_step) + 2u * k, idesc, (a | c | k) != 0 ? 1u : accum0);
      }
    }
  }
}

__global__ void __launch_bounds__(kNsThreads, 1) nsconv_kernel(const __grid_constant__ TcNsConv P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;
  uint8_t* halo = smem + P.w_bytes;
  NsCtl* ctl = reinterpret_cast<NsCtl*>(halo + (size_t)P.nstages * P.stage_bytes);
  float* xch = reinterpret_cast<float*>(halo + (size_t)P.nstages * P.stage_bytes + P.xch_off);
  const int split = blockIdx.x % P.co_splits, cta = blockIdx.x / P.co_splits, ncta = gridDim.x / P.co_splits;
  // SV_NS_DEBUG bit 2: per-CTA cycle sums of where each role waits (read back through sv_debug_halo_trace, scripts/ns_trace.py):
  //   [0] CTA lifetime  [1] MMA thread: weights landed  [2] MMA: sum of halo_full waits  [3] MMA: sum of acc_empty waits
  //   [4] MMA: sum of issue + commit  [5] producer: sum of halo_empty waits  [6] epilogue warp 0: sum of acc_full waits  [7] its work
  //   bit 3 (instead of bit 2): [0] CTA lifetime  [1] globaltimer at CTA start  [2] globaltimer at CTA end  [3] smid
  // (compiled in only with -DSV_NS_TRACE - SV_BUILD_DEFINES=-DSV_NS_TRACE python splitvae_b200/build.py --force: the counters cost
  //  registers in a kernel that already spills at its 96-register cap)
#ifdef SV_NS_TRACE
  const bool trc = (P.debug & 4) != 0, trg = (P.debug & 8) != 0;
#else
  constexpr bool trc = false, trg = false;
#endif
  const long long t_cta0 = (trc || trg) ? clock64() : 0;
  if (trg && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
    unsigned smid; unsigned long long gt;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_halo_trace[blockIdx.x * kTraceSlots + 1] = gt; g_halo_trace[blockIdx.x * kTraceSlots + 3] = smid;
  }

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map_x);
    tc::prefetch_tmap(&P.map_w);
    if (P.pair) tc::prefetch_tmap(&P.map_x_lo);
    tc::mbar_init(&ctl->w_full, 1);
    for (int i = 0; i < P.nstages; ++i) { tc::mbar_init(&ctl->halo_full[i], 1); tc::mbar_init(&ctl->halo_empty[i], 1); }
    for (int i = 0; i < P.groups; ++i) { tc::mbar_init(&ctl->acc_full[i], 1); tc::mbar_init(&ctl->acc_empty[i], 4 * P.lanes); }
    tc::fence_barrier_init();
  }
  const int acc1 = P.acc1, acc_cols = P.mb * acc1;
  const int pair = P.pair, wrows = pair ? 2 * P.n_total : P.n_total;      // rows of one weight k-block
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(P.groups * acc_cols)) tmem_cols <<= 1;
  tc::pdl_trigger();
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (tc::elect_one()) {
      // (the packed weights are written by the optimizer's re-pack, which every kernel of the step depends on in full: loading them
      //  before pdl_wait() overlaps the previous kernel's tail)
      tc::mbar_expect_tx(&ctl->w_full, (uint32_t)(P.num_kb * P.wk_bytes));
      for (int j = 0; j < P.num_kb; ++j)
        for (int r0 = 0; r0 < wrows; r0 += P.n_box)
          tc::tma_load_2d(wsm + (size_t)j * P.wk_bytes + (size_t)r0 * P.pixB, &P.map_w, &ctl->w_full, j * P.ck, split * wrows + r0);
      tc::pdl_wait();
      const uint32_t halo_tx = (uint32_t)(P.nchunks * P.halo_rows * P.W * P.pixB);
      const int passes = pair ? 2 : 1;            // pair: the hi plane and the lo plane of a tile are two consecutive ring stages
      long long t_prod_wait = 0;
      int i = 0;
      for (int t = cta; t < P.tiles; t += ncta) {
        const int n = t / P.tiles_per_img, y0 = (t - n * P.tiles_per_img) * P.R * P.mb;
        for (int pl = 0; pl < passes; ++pl, ++i) {
          const int st = i % P.nstages, ph = (i / P.nstages) & 1;
          const long long tw = trc ? clock64() : 0;
          tc::mbar_wait(&ctl->halo_empty[st], ph ^ 1);
          if (trc) t_prod_wait += clock64() - tw;
          tc::mbar_expect_tx(&ctl->halo_full[st], halo_tx);
          for (int c = 0; c < P.nchunks; ++c)
            tc::tma_load_4d(halo + (size_t)st * P.stage_bytes + (size_t)c * P.chunk_bytes, pl ? &P.map_x_lo : &P.map_x, &ctl->halo_full[st],
                            c * P.ck, 0, y0 - P.pad_t, n);
        }
      }
      if (trc && blockIdx.x < kTraceCtas) g_halo_trace[blockIdx.x * kTraceSlots + 5] = (unsigned long long)t_prod_wait;
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, P.ncols, 0, 0);
      const uint32_t lt = tc::layout_type_for(P.pixB);
      const uint64_t tmpl = tc::make_smem_desc(0, 16, 8u * (uint32_t)P.pixB, lt);   // K-major, dense 8-row groups (A and B alike)
      const uint32_t w_addr = tc::smem_u32(wsm) >> 4;
      const uint32_t row_step = ((uint32_t)P.W * (uint32_t)P.pixB) >> 4;             // one image row of the halo
      const uint32_t grp_step = ((uint32_t)P.ncols * (uint32_t)P.pixB) >> 4;         // next N group inside a weight k-block
      const uint32_t wk_step = (uint32_t)P.wk_bytes >> 4, chunk_step = (uint32_t)P.chunk_bytes >> 4;
      const int ksteps = P.ck / 16, kh = P.kh, nchunks = P.nchunks, ng = P.ng, tiles = P.tiles, nstages = P.nstages, groups = P.groups;
      const uint32_t ncols = (uint32_t)P.ncols;
      const uint32_t halo_addr = tc::smem_u32(halo) >> 4, stage_step = (uint32_t)P.stage_bytes >> 4;
      const int mb = P.mb;
      const uint32_t blk_step = ((uint32_t)(P.R * P.W) * (uint32_t)P.pixB) >> 4;     // R image rows: next 128-pixel block of the tile
      const int shape = nchunks != 1 || kh != 6 ? 0
                        : ng == 1 ? (ksteps == 4 ? 1 : ksteps == 2 ? 2 : ksteps == 1 ? 3 : 0)
                                  : ng == 2 ? (ksteps == 2 ? 4 : ksteps == 4 ? 5 : 0) : 0;
      // pair mode: pass 0 = hi plane x the whole k-block (N = 2 * n_total: [main | correction]), pass 1 = lo plane x the W_hi rows
      const uint32_t idesc_hi = tc::make_idesc_bf16(128, 2 * P.n_total, 0, 0);
      const int passes = pair ? 2 : 1;
      const int pshape = !pair ? 0 : (kh == 6 && nchunks == 1 && ksteps == 4) ? 1 : (kh == 4 && nchunks == 2 && ksteps == 4) ? 2 : 0;
      tc::mbar_wait(&ctl->w_full, 0);
      long long t_hw = 0, t_aw = 0, t_is = 0, tq = 0;
      if (trc && blockIdx.x < kTraceCtas) g_halo_trace[blockIdx.x * kTraceSlots + 1] = (unsigned long long)(clock64() - t_cta0);
      int i = 0, it = 0;                        // i: ring stage counter, it: tile counter
      for (int t = cta; t < tiles; t += ncta, ++it) {
        const int ab = it % groups, aph = (it / groups) & 1;
        for (int pl = 0; pl < passes; ++pl, ++i) {
        const int st = i % nstages, ph = (i / nstages) & 1;
        if (trc) tq = clock64();
        tc::mbar_wait(&ctl->halo_full[st], ph);
        if (trc) { const long long c = clock64(); t_hw += c - tq; tq = c; }
        if (pl == 0) tc::mbar_wait(&ctl->acc_empty[ab], aph ^ 1);
        if (trc) { const long long c = clock64(); t_aw += c - tq; tq = c; }
        tc::tc_fence_after();
        for (int mbi = 0; mbi < mb; ++mbi) {    // block mbi of the tile: rows mbi*R .. of the shared halo, its own accumulator
        const uint32_t h_addr = halo_addr + (uint32_t)st * stage_step + (uint32_t)mbi * blk_step;
        const uint32_t acc = tmem_base + (uint32_t)(ab * acc_cols + mbi * acc1);
        const uint64_t da0 = tmpl + h_addr, db0 = tmpl + w_addr;
        // fully unrolled issue sequences for the shapes of this model family (the single issuing thread must spend only a
        // few instructions per tcgen05.mma: the generic nest below costs ~75 and caps the tensor pipe at ~35 %)
        if (P.debug & 2) { }
        else if (pair) {
          const uint32_t id = pl ? idesc : idesc_hi, acc0 = pl ? 1u : 0u;
          if (pshape == 1) ns_issue_tile<6, 4, 1>(da0, db0, acc, id, row_step, wk_step, grp_step, ncols, acc0);
          else if (pshape == 2) ns_issue_tile<4, 4, 1, 2>(da0, db0, acc, id, row_step, wk_step, grp_step, ncols, acc0, chunk_step);
          else {
            uint32_t b_addr = w_addr, accum = acc0;
            for (int a = 0; a < kh; ++a) {
              uint32_t a_addr = h_addr + (uint32_t)a * row_step;
              for (int c = 0; c < nchunks; ++c, a_addr += chunk_step, b_addr += wk_step)
                for (int k = 0; k < ksteps; ++k) {
                  tc::umma_bf16(acc, tmpl + a_addr + 2u * k, tmpl + b_addr + 2u * k, id, accum);
                  accum = 1u;
                }
            }
          }
        }
        else if (shape == 1) ns_issue_tile<6, 4, 1>(da0, db0, acc, idesc, row_step, wk_step, grp_step, ncols);
        else if (shape == 2) ns_issue_tile<6, 2, 1>(da0, db0, acc, idesc, row_step, wk_step, grp_step, ncols);
        else if (shape == 3) ns_issue_tile<6, 1, 1>(da0, db0, acc, idesc, row_step, wk_step, grp_step, ncols);
        else if (shape == 4) ns_issue_tile<6, 2, 2>(da0, db0, acc, idesc, row_step, wk_step, grp_step, ncols);
        else if (shape == 5) ns_issue_tile<6, 4, 2>(da0, db0, acc, idesc, row_step, wk_step, grp_step, ncols);
        else {
          uint32_t b_addr = w_addr, accum = 0;
          for (int a = 0; a < kh; ++a) {
            uint32_t a_addr = h_addr + (uint32_t)a * row_step;
            for (int c = 0; c < nchunks; ++c, a_addr += chunk_step, b_addr += wk_step) {
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t da = tmpl + a_addr + 2u * k;
                uint64_t db = tmpl + b_addr + 2u * k;
                uint32_t d = acc;
                for (int g = 0; g < ng; ++g, db += grp_step, d += ncols) tc::umma_bf16(d, da, db, idesc, accum);
                accum = 1u;
              }
            }
          }
        }
        }
        tc::umma_commit(&ctl->halo_empty[st]);
        if (trc) t_is += clock64() - tq;
        }
        tc::umma_commit(&ctl->acc_full[ab]);
      }
      if (trc && blockIdx.x < kTraceCtas) {
        unsigned long long* tr = g_halo_trace + blockIdx.x * kTraceSlots;
        tr[2] = (unsigned long long)t_hw; tr[3] = (unsigned long long)t_aw; tr[4] = (unsigned long long)t_is;
      }
    }
  } else {
    tc::pdl_wait();                              // (stores, mask reads: only after the previous kernel of the stream has completed)
    const int e = (warp - 2) >> 2;               // epilogue warp set: (group, chunk lane); the TMEM quarter is warp % 4
    if (e < P.groups * P.lanes) ns_epilogue(P, ctl, xch, tmem_base, warp & 3, e / P.lanes, e % P.lanes, lane);
  }
  tc::tc_fence_before();
  __syncthreads();
  if ((trc || trg) && threadIdx.x == 0 && blockIdx.x < kTraceCtas) {
    g_halo_trace[blockIdx.x * kTraceSlots + 0] = (unsigned long long)(clock64() - t_cta0);
    if (trg) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); g_halo_trace[blockIdx.x * kTraceSlots + 2] = gt; }
  }
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad: D[(tap,ci), co] = sum over pixels; K axis = pixels (64 per stage), both operands MN-major.
// ------------------------------------------------------------------------------------------------
constexpr int kWgMaxA = 6, kWgMaxB = 2;
struct WgCtl {
  uint64_t a_full[kWgMaxA], a_empty[kWgMaxA], b_full[kWgMaxB], b_empty[kWgMaxB], tmem_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads) wgrad_kernel(const __grid_constant__ TcWgradLaunch P) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kPix = 64;
  const int a_box = kPix * P.cb * 2;             // one (tap, channel-block) box
  const int a_stage = 128 * kPix * 2;            // nsub boxes = 128 rows of the M axis
  const int b_box = kPix * P.cbn * 2;
  const int b_stage = P.tile_cols * kPix * 2;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)P.a_stages * a_stage;
  WgCtl* ctl = reinterpret_cast<WgCtl*>(b_ring + (size_t)P.b_stages * b_stage);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = blockIdx.x * P.groups_per_cta;
  const int ng = min(P.groups_per_cta, P.groups - g0);
  const int n_tile = blockIdx.y;
  const int c_begin = blockIdx.z * P.chunks_per_split;
  const int c_end = min(P.nchunks, c_begin + P.chunks_per_split);
  const int chunks_per_img = P.grid_h / P.tile_h;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map_a);
    tc::prefetch_tmap(&P.map_b);
    for (int i = 0; i < P.a_stages; ++i) { tc::mbar_init(&ctl->a_full[i], 1); tc::mbar_init(&ctl->a_empty[i], 1); }
    for (int i = 0; i < P.b_stages; ++i) { tc::mbar_init(&ctl->b_full[i], 1); tc::mbar_init(&ctl->b_empty[i], 1); }
    tc::mbar_init(&ctl->tmem_full, 1);
    tc::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(P.groups_per_cta * P.tile_cols)) tmem_cols <<= 1;
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    if (tc::elect_one()) {
      int ia = 0;  // running A-stage counter
      for (int c = c_begin; c < c_end; ++c) {
        const int ib = c - c_begin;
        const int bs = ib % P.b_stages, bph = (ib / P.b_stages) & 1;
        const int n0 = P.tile_n_img > 1 ? c * P.tile_n_img : c / chunks_per_img;
        const int y0 = P.tile_n_img > 1 ? 0 : (c % chunks_per_img) * P.tile_h;
        tc::mbar_wait(&ctl->b_empty[bs], bph ^ 1);
        tc::mbar_expect_tx(&ctl->b_full[bs], b_stage);
        for (int j = 0; j < P.tile_cols / P.cbn; ++j)
          tc::tma_load_4d(b_ring + (size_t)bs * b_stage + (size_t)j * b_box, &P.map_b, &ctl->b_full[bs],
                          n_tile * P.tile_cols + j * P.cbn, 0, y0, n0);
        for (int g = 0; g < ng; ++g, ++ia) {
          const int as = ia % P.a_stages, aph = (ia / P.a_stages) & 1;
          tc::mbar_wait(&ctl->a_empty[as], aph ^ 1);
          const int sb0 = (g0 + g) * P.nsub;
          const int nv = min(P.nsub, P.total_sb - sb0);
          tc::mbar_expect_tx(&ctl->a_full[as], nv * a_box);
          for (int j = 0; j < nv; ++j) {
            const int sb = sb0 + j;
            const int tap = sb / P.ncb, cblk = sb - tap * P.ncb;
            const int ta = tap / P.taps_w, tb = tap - ta * P.taps_w;
            tc::tma_load_4d(a_ring + (size_t)as * a_stage + (size_t)j * a_box, &P.map_a, &ctl->a_full[as], cblk * P.cb,
                            tb - P.pad_l, y0 * P.a_stride + ta - P.pad_t, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, P.tile_cols, 1, 1);
      const uint32_t lta = tc::layout_type_for(P.a_swizzle), ltb = tc::layout_type_for(P.b_swizzle);
      const uint32_t a_row = P.cb * 2, b_row = P.cbn * 2;  // bytes per pixel row inside a box
      const uint64_t a_tmpl = tc::make_smem_desc(0, a_box, 8 * a_row, lta), b_tmpl = tc::make_smem_desc(0, b_box, 8 * b_row, ltb);
      int ia = 0;
      for (int c = c_begin; c < c_end; ++c) {
        const int ib = c - c_begin;
        const int bs = ib % P.b_stages, bph = (ib / P.b_stages) & 1;
        tc::mbar_wait(&ctl->b_full[bs], bph);
        const uint32_t sb_addr = tc::smem_u32(b_ring + (size_t)bs * b_stage);
        for (int g = 0; g < ng; ++g, ++ia) {
          const int as = ia % P.a_stages, aph = (ia / P.a_stages) & 1;
          tc::mbar_wait(&ctl->a_full[as], aph);
          tc::tc_fence_after();
          const uint32_t sa_addr = tc::smem_u32(a_ring + (size_t)as * a_stage);
          // MN-major: LBO = distance between channel blocks (one box), SBO = 8 pixel rows; a K step of 16 pixels advances
          // the start address by 16 pixel rows (a_row / b_row in units of 16 B)
          const uint64_t da = a_tmpl + (sa_addr >> 4), db = b_tmpl + (sb_addr >> 4);
          const uint32_t acc = tmem_base + (uint32_t)(g * P.tile_cols);
          tc::umma_bf16(acc, da, db, idesc, c > c_begin);
          tc::umma_bf16(acc, da + a_row, db + b_row, idesc, 1u);
          tc::umma_bf16(acc, da + 2 * a_row, db + 2 * b_row, idesc, 1u);
          tc::umma_bf16(acc, da + 3 * a_row, db + 3 * b_row, idesc, 1u);
          tc::umma_commit(&ctl->a_empty[as]);
        }
        tc::umma_commit(&ctl->b_empty[bs]);
      }
      tc::umma_commit(&ctl->tmem_full);
    }
  } else {
    tc::mbar_wait(&ctl->tmem_full, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int g = 0; g < ng; ++g) {
      float* dst = P.partial + ((size_t)blockIdx.z * P.m_pad + (size_t)(g0 + g) * 128 + row) * P.n_pad + (size_t)n_tile * P.tile_cols;
      for (int c0 = 0; c0 < P.tile_cols; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * P.tile_cols + c0);
        const int ncol = min(32, P.tile_cols - c0);
        if (ncol >= 32) tc::tmem_ld32(taddr, v); else tc::tmem_ld16(taddr, v);
        tc::tmem_ld_wait();
        if (c_end > c_begin) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (i < ncol) *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (i < ncol) *reinterpret_cast<uint4*>(dst + c0 + i) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}


// ------------------------------------------------------------------------------------------------
// Halo-resident wgrad (see TcHaloWgrad in tc_kernels.h).
// ------------------------------------------------------------------------------------------------
struct HwCtl {
  uint64_t full[4], empty[4], tmem_full;
  uint32_t tmem_base;
  uint32_t goff[32];   // per row group: byte offset >> 4 of its first sub-block inside a stage's X halo
};

// Issues the MMAs of one pixel tile for NG (compile-time) row groups: descriptors live in registers and the group loop is
// fully unrolled, so the single issuing thread spends only a few (uniform-datapath) instructions per tcgen05.mma.
template <int NG>
__device__ __forceinline__ void hw_issue_tile(uint64_t da0, uint64_t db0, const uint32_t* goff, uint32_t tmem_base, uint32_t n_pad, uint32_t idesc,
                                              int TH, int ksx, uint32_t a_row, uint32_t b_row, uint32_t a_xs, uint32_t b_xs, uint32_t accum0) {
  uint64_t dag[NG];
  uint32_t acc[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) { dag[g] = da0 + goff[g]; acc[g] = tmem_base + (uint32_t)g * n_pad; }
  uint32_t accum = accum0;
  uint32_t ao = 0, bo = 0;
  for (int yy = 0; yy < TH; ++yy, ao += a_row, bo += b_row) {
    uint32_t ax = ao, bx = bo;
    for (int xs = 0; xs < ksx; ++xs, ax += a_xs, bx += b_xs) {
      const uint64_t db = db0 + bx;
#pragma unroll
      for (int g = 0; g < NG; ++g) tc::umma_bf16(acc[g], dag[g] + ax, db, idesc, accum);
      accum = 1u;
    }
  }
}

__global__ void __launch_bounds__(kThreads) halo_wgrad_kernel(const __grid_constant__ TcHaloWgrad P) {
  pdl_enter();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  HwCtl* ctl = reinterpret_cast<HwCtl*>(smem + (size_t)P.stages * P.stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g0 = blockIdx.x * P.groups_per_cta;
  const int ng = min(P.groups_per_cta, P.groups - g0);
  const int t_begin = blockIdx.z * P.tiles_per_split;
  const int t_end = min(P.tiles, t_begin + P.tiles_per_split);

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map_x);
    tc::prefetch_tmap(&P.map_dy);
    for (int i = 0; i < P.stages; ++i) { tc::mbar_init(&ctl->full[i], 1); tc::mbar_init(&ctl->empty[i], 1); }
    tc::mbar_init(&ctl->tmem_full, 1);
    tc::fence_barrier_init();
  }
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)(P.groups_per_cta * P.n_pad)) tmem_cols <<= 1;
  if (warp == 1) tc::tmem_alloc(&ctl->tmem_base, tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  const int x_bytes = P.nchunks * P.x_chunk_bytes;

  if (warp == 0) {
    if (tc::elect_one()) {
      const uint32_t tx_bytes = (uint32_t)(P.nchunks * P.x_th * P.x_tw * P.cb * 2 + P.nbchunks * P.dy_th * P.dy_tw * P.cbn * 2);
      for (int t = t_begin; t < t_end; ++t) {
        const int i = t - t_begin, st = i % P.stages, ph = (i / P.stages) & 1;
        int r = t;
        const int tile_x = r % P.tiles_x; r /= P.tiles_x;
        const int tile_y = r % P.tiles_y;
        const int n = r / P.tiles_y;
        const int x0 = tile_x * P.TW, y0 = tile_y * P.TH;
        uint8_t* sx = smem + (size_t)st * P.stage_bytes;
        tc::mbar_wait(&ctl->empty[st], ph ^ 1);
        tc::mbar_expect_tx(&ctl->full[st], tx_bytes);
        for (int c = 0; c < P.nchunks; ++c)
          tc::tma_load_4d(sx + (size_t)c * P.x_chunk_bytes, &P.map_x, &ctl->full[st], c * P.cb, x0 + P.x_dx, y0 * P.sy + P.x_dy, n);
        for (int c = 0; c < P.nbchunks; ++c)
          tc::tma_load_4d(sx + x_bytes + (size_t)c * P.dy_chunk_bytes, &P.map_dy, &ctl->full[st], c * P.cbn, x0 + P.dy_dx, y0, n);
      }
    }
  } else if (warp == 1) {
    const uint32_t pix = (uint32_t)P.cb * 2u, pixb = (uint32_t)P.cbn * 2u;
    if (lane < ng) ctl->goff[lane] = P.goff[g0 + lane];   // host-computed group offsets
    __syncwarp();
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_bf16(128, P.n_pad, 1, 1);
      // A K step is 16 pixels = two groups of 8 (the descriptors' stride-dimension offset): 16 horizontally adjacent pixels of one
      // output row, or - 8-pixel-wide images (krows = 2) - the same 8 columns of two consecutive output rows, one tile-row pitch apart.
      const int kr = P.krows;
      const uint32_t a_sbo = kr == 2 ? (uint32_t)(P.sy * P.x_tw) * pix : 8u * pix, b_sbo = kr == 2 ? (uint32_t)P.dy_tw * pixb : 8u * pixb;
      const uint64_t a_tmpl = tc::make_smem_desc(0, (uint32_t)P.a_lbo, a_sbo, tc::layout_type_for(P.x_swizzle));
      const uint64_t b_tmpl = tc::make_smem_desc(0, (uint32_t)P.b_lbo, b_sbo, tc::layout_type_for(P.dy_swizzle));
      const uint32_t n_pad = (uint32_t)P.n_pad;
      const int ksx = kr == 2 ? 1 : P.TW / 16, ny = P.TH / kr;
      const uint32_t a_xs = (16u * pix) >> 4, b_xs = (16u * pixb) >> 4;
      const uint32_t a_row = ((uint32_t)(kr * P.sy * P.x_tw) * pix) >> 4, b_row = ((uint32_t)(kr * P.dy_tw) * pixb) >> 4;   // (sy = 2: every other X row)
      const volatile uint32_t* goff = ctl->goff;
      for (int t = t_begin; t < t_end; ++t) {
        const int i = t - t_begin, st = i % P.stages, ph = (i / P.stages) & 1;
        tc::mbar_wait(&ctl->full[st], ph);
        tc::tc_fence_after();
        const uint32_t sx = tc::smem_u32(smem + (size_t)st * P.stage_bytes);
        const uint64_t da0 = a_tmpl + (sx >> 4), db0 = b_tmpl + ((sx + (uint32_t)x_bytes) >> 4);
        const uint32_t accum0 = t > t_begin ? 1u : 0u;
        if (ng <= 4) {
          uint32_t go[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) go[g] = goff[g < ng ? g : 0];
          switch (ng) {
            case 1: hw_issue_tile<1>(da0, db0, go, tmem_base, n_pad, idesc, ny, ksx, a_row, b_row, a_xs, b_xs, accum0); break;
            case 2: hw_issue_tile<2>(da0, db0, go, tmem_base, n_pad, idesc, ny, ksx, a_row, b_row, a_xs, b_xs, accum0); break;
            case 3: hw_issue_tile<3>(da0, db0, go, tmem_base, n_pad, idesc, ny, ksx, a_row, b_row, a_xs, b_xs, accum0); break;
            default: hw_issue_tile<4>(da0, db0, go, tmem_base, n_pad, idesc, ny, ksx, a_row, b_row, a_xs, b_xs, accum0); break;
          }
        } else {
          uint64_t da_y = da0, db_y = db0;
          uint32_t accum = accum0;
          for (int yy = 0; yy < ny; ++yy, da_y += a_row, db_y += b_row) {
            uint64_t da = da_y, db = db_y;
            for (int xs = 0; xs < ksx; ++xs, da += a_xs, db += b_xs) {
              uint32_t acc = tmem_base;
              int g = 0;
              for (; g + 4 <= ng; g += 4) {
                const uint32_t o0 = goff[g], o1 = goff[g + 1], o2 = goff[g + 2], o3 = goff[g + 3];
                tc::umma_bf16(acc, da + o0, db, idesc, accum);
                tc::umma_bf16(acc + n_pad, da + o1, db, idesc, accum);
                tc::umma_bf16(acc + 2 * n_pad, da + o2, db, idesc, accum);
                tc::umma_bf16(acc + 3 * n_pad, da + o3, db, idesc, accum);
                acc += 4 * n_pad;
              }
              for (; g < ng; ++g, acc += n_pad) tc::umma_bf16(acc, da + goff[g], db, idesc, accum);
              accum = 1u;
            }
          }
        }
        tc::umma_commit(&ctl->empty[st]);
      }
      tc::umma_commit(&ctl->tmem_full);
    }
  } else {
    tc::mbar_wait(&ctl->tmem_full, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    for (int g = 0; g < ng; ++g) {
      float* dst = P.partial + ((size_t)blockIdx.z * P.m_pad + (size_t)(g0 + g) * 128 + row) * P.n_pad;
      for (int c0 = 0; c0 < P.n_pad; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(g * P.n_pad + c0);
        const int ncol = min(32, P.n_pad - c0);
        if (ncol >= 32) tc::tmem_ld32(taddr, v); else tc::tmem_ld16(taddr, v);
        tc::tmem_ld_wait();
        const bool have = t_end > t_begin;
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          if (i < ncol)
            *reinterpret_cast<uint4*>(dst + c0 + i) = have ? make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, tmem_cols);
  }
}

// row of (tap, ci) inside the split-K partial buffer for each wgrad flavour
struct WgRowMap { int mode, kw, ci_pad, cb, nsub, gw, gpt, nb; int pair_c = 0, pair_sh = 0, kwp = 0; };
// mode 0: tap*ci_pad+ci; 1: first layer; 2: halo, taps stacked along the filter row; 3: halo, chunks of one tap;
// 4/5: N-stacked halo (rows enumerate (filter row a, ci); columns (kw-1-b)*nb + co): 4 = vertical taps stacked, 5 = chunks stacked
// pair_c > 0 (stride-2 layer on the halo kernel through the pixel-PAIR view [W/2][2 * pair_c]): filter column b of the layer is
// pair tap u >> 1 and parity u & 1 with u = b + pair_sh; the parity selects the half of the pair's 2 * pair_c channels
__device__ __forceinline__ void wg_tap(const WgRowMap& R, int tap, int& a, int& b, int& ci, int& kw) {
  a = tap / R.kw; b = tap - a * R.kw; kw = R.kw;
  if (R.pair_c) { const int u = b + R.pair_sh; b = u >> 1; ci += (u & 1) * R.pair_c; kw = R.kwp; }
}
__device__ __forceinline__ size_t wg_row(const WgRowMap& R, int tap0, int ci) {
  int a, b, kw;
  wg_tap(R, tap0, a, b, ci, kw);
  const int tap = a * kw + b;
  switch (R.mode) {
    case 1: return (size_t)a * 64 + b * 8 + ci;
    case 2: return (size_t)(a * R.gw + b / R.nsub) * 128 + (b % R.nsub) * R.cb + ci;
    case 3: { const int c = ci / R.cb; return (size_t)(tap * R.gpt + c / R.nsub) * 128 + (c % R.nsub) * R.cb + ci % R.cb; }
    case 4: return (size_t)(a / R.nsub) * 128 + (a % R.nsub) * R.cb + ci;
    case 5: { const int c = ci / R.cb; return (size_t)(a * R.gpt + c / R.nsub) * 128 + (c % R.nsub) * R.cb + ci % R.cb; }
    default: return (size_t)tap * R.ci_pad + ci;
  }
}
__device__ __forceinline__ int wg_col(const WgRowMap& R, int tap, int co) {
  if (R.mode < 4) return co;
  int a, b, kw, ci = 0;
  wg_tap(R, tap, a, b, ci, kw);
  return (kw - 1 - b) * R.nb + co;
}

// sums the split-K partials in a fixed order and scatters into the Keras-layout gradient arena.
// block = 32 consecutive output columns x 8 split lanes: lane y sums splits y, y+8, ... (coalesced 128-byte rows), then the
// 8 partial sums are combined in a fixed order through shared memory -> deterministic, no atomics.
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(ConvGeom g, const float* __restrict__ partial, int k_splits, int m_pad, int n_pad,
                                                           WgRowMap R, float* __restrict__ grads) {
  pdl_enter();
  __shared__ float red[8][33];
  const int rows = g.kh * g.kw * g.Ci;
  const int cblocks = (g.Co + 31) / 32;
  const int r = blockIdx.x / cblocks, cb = blockIdx.x - r * cblocks;
  const int co = cb * 32 + threadIdx.x;
  const int ci = r % g.Ci, tap = r / g.Ci;
  const size_t row = wg_row(R, tap, ci);
  float s = 0.f;
  if (co < g.Co && r < rows)
    for (int k = threadIdx.y; k < k_splits; k += 8) s += partial[((size_t)k * m_pad + row) * n_pad + wg_col(R, tap, co)];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && co < g.Co && r < rows) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    int lc;
    const int j = part_of(g, co, lc);
    grads[g.part_w[j] + ((long long)tap * g.Ci + ci) * g.part_n[j] + lc] = t;
  }
}

// Many splits, VEC consecutive output columns per thread (128/64/32-bit loads): block = 32 column vectors x `blockDim.y` split
// lanes; lane y sums splits y, y+L, ... with four independent loads in flight, then the L partial sums are combined in a fixed
// order through shared memory (deterministic).  4x fewer threads / instructions per byte than wgrad_reduce_kernel, whose
// scalar 128-byte rows ran at 1.1-1.8 TB/s out of L2.  Requires Co, every part width % VEC == 0.
template <int VEC> struct VecF;
template <> struct VecF<4> { typedef float4 T; };
template <> struct VecF<2> { typedef float2 T; };
template <> struct VecF<1> { typedef float T; };
template <int VEC>
__global__ void __launch_bounds__(1024) wgrad_reduce_vec_kernel(ConvGeom g, const float* __restrict__ partial, int k_splits, int m_pad, int n_pad,
                                                                WgRowMap R, float* __restrict__ grads) {
  pdl_enter();
  typedef typename VecF<VEC>::T V;
  extern __shared__ float red_dyn[];   // [blockDim.y][32 * VEC]
  const int cov = g.Co / VEC;
  const long long items = (long long)g.kh * g.kw * g.Ci * cov;
  const long long item = (long long)blockIdx.x * 32 + threadIdx.x;
  const bool live = item < items;
  float s[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) s[i] = 0.f;
  int tap = 0, ci = 0, co = 0;
  if (live) {
    co = (int)(item % cov) * VEC;
    const int r = (int)(item / cov);
    ci = r % g.Ci; tap = r / g.Ci;
    const size_t zstride = (size_t)m_pad * n_pad;
    const float* p = partial + wg_row(R, tap, ci) * n_pad + wg_col(R, tap, co);
    const int L = blockDim.y;
    int k = threadIdx.y;
    for (; k + 3 * L < k_splits; k += 4 * L) {
      const V a = *reinterpret_cast<const V*>(p + (size_t)k * zstride);
      const V b = *reinterpret_cast<const V*>(p + (size_t)(k + L) * zstride);
      const V c = *reinterpret_cast<const V*>(p + (size_t)(k + 2 * L) * zstride);
      const V d = *reinterpret_cast<const V*>(p + (size_t)(k + 3 * L) * zstride);
      const float* fa = reinterpret_cast<const float*>(&a); const float* fb = reinterpret_cast<const float*>(&b);
      const float* fc = reinterpret_cast<const float*>(&c); const float* fd = reinterpret_cast<const float*>(&d);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[i] += (fa[i] + fb[i]) + (fc[i] + fd[i]);
    }
    for (; k < k_splits; k += L) {
      const V a = *reinterpret_cast<const V*>(p + (size_t)k * zstride);
      const float* fa = reinterpret_cast<const float*>(&a);
#pragma unroll
      for (int i = 0; i < VEC; ++i) s[i] += fa[i];
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) red_dyn[(threadIdx.y * 32 + threadIdx.x) * VEC + i] = s[i];
  __syncthreads();
  if (threadIdx.y == 0 && live) {
    float t[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) t[i] = 0.f;
    for (int y = 0; y < (int)blockDim.y; ++y) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) t[i] += red_dyn[(y * 32 + threadIdx.x) * VEC + i];
    }
    int lc;
    const int j = part_of(g, co, lc);
    float* o = grads + g.part_w[j] + ((long long)tap * g.Ci + ci) * g.part_n[j] + lc;
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(o) = make_float4(t[0], t[1], t[2], t[3]);
    else if constexpr (VEC == 2) *reinterpret_cast<float2*>(o) = make_float2(t[0], t[1]);
    else o[0] = t[0];
  }
}

// few splits, 4 consecutive output columns per thread (128-bit loads / stores); requires Co, every part width and n_pad to be
// multiples of 4 and a row map with contiguous columns (mode < 4)
__global__ void __launch_bounds__(256) wgrad_reduce_few4_kernel(ConvGeom g, const float* __restrict__ partial, int k_splits, int m_pad, int n_pad,
                                                                WgRowMap R, float* __restrict__ grads) {
  pdl_enter();
  const int co4 = g.Co >> 2;
  const long long total = (long long)g.kh * g.kw * g.Ci * co4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % co4) * 4;
    const int r = (int)(idx / co4);
    const int ci = r % g.Ci, tap = r / g.Ci;
    const size_t row = wg_row(R, tap, ci);
    float4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = k < k_splits ? *reinterpret_cast<const float4*>(partial + ((size_t)k * m_pad + row) * n_pad + co) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { t.x += v[k].x; t.y += v[k].y; t.z += v[k].z; t.w += v[k].w; }
    int lc;
    const int j = part_of(g, co, lc);
    *reinterpret_cast<float4*>(grads + g.part_w[j] + ((long long)tap * g.Ci + ci) * g.part_n[j] + lc) = t;
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_few_kernel(ConvGeom g, const float* __restrict__ partial, int k_splits, int m_pad, int n_pad,
                                                               WgRowMap R, float* __restrict__ grads) {
  pdl_enter();
  const long long total = (long long)g.kh * g.kw * g.Ci * g.Co;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % g.Co);
    const int r = (int)(idx / g.Co);
    const int ci = r % g.Ci, tap = r / g.Ci;
    const size_t row = wg_row(R, tap, ci);
    float v[8];
#pragma unroll
    const int pc = wg_col(R, tap, co);
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = k < k_splits ? partial[((size_t)k * m_pad + row) * n_pad + pc] : 0.f;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += v[k];
    int lc;
    const int j = part_of(g, co, lc);
    grads[g.part_w[j] + ((long long)tap * g.Ci + ci) * g.part_n[j] + lc] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: one multi-tensor kernel refreshes every bf16 operand copy from the fp32 masters
// ------------------------------------------------------------------------------------------------
constexpr int kPackKT = 4;
constexpr int kPackDgIter = 4;      // kind 1 (dgrad copies): 8-element groups per thread
struct PackJob {
  int kind;                 // 0: fwd weights, 1: dgrad weights (one parity class), 2: bias, 4 / 5: N-stacked fwd / dgrad weights
  int KH, KW, Ci, Co;       // layer geometry (logical)
  int nparts, part_n[3];
  long long part_w[3], part_b[3];
  int rows_pad, taps_h, taps_w, k_pad;   // dst = [rows_pad][taps_h*taps_w][k_pad]
  // kind 0: the K axis of a tap (a, b') is a sequence of k-blocks of ppx pixels x cpp channels, [channel chunk][section]:
  // k_pad = chunks * nsec * ppx * cpp; tap (a, b') covers the filter columns b = b' * ppx + px.  ppx = 1: one pixel per tap;
  // 2: the pixel-pair views; 8: the first layer's window view.  nsec = 2 (bf16x3): section 0 = bf16(W) = W_hi, section 1 =
  // bf16(W - W_hi) = W_lo.  first_cat (first layer, bf16x3): the staged pixel holds [x_hi(3) x_lo(3) 0 0], so section 0 =
  // [W_hi W_hi 0 0] (x_hi*W_hi + x_lo*W_hi) and section 1 = [W_lo 0 0 0].
  int ppx, cpp, nsec, first_cat;
  int fast;                 // kind 0 with ppx == 1, cpp % 32 == 0, no first_cat: one block converts kPackKT 32-channel tiles and writes all sections
  int stride, rh, rw;       // dgrad: kh = stride*(taps_h-1-a) + rh  (stride 1: rh = 0)
  void* dst;
  long long count;
  int block_start;
};

__device__ __forceinline__ float master_w(const PackJob& J, const float* params, int kh, int kw, int ci, int co) {
  int j = 0, lc = co;
  while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
  return params[J.part_w[j] + ((long long)(kh * J.KW + kw) * J.Ci + ci) * J.part_n[j] + lc];
}

__global__ void __launch_bounds__(256) pack_kernel(const PackJob* __restrict__ jobs, const uint16_t* __restrict__ block_job,
                                                   const float* __restrict__ params) {
  pdl_enter();
  const int lo = block_job[blockIdx.x];        // (one load: the binary search over block_start was six dependent L2 round trips per block)
  __shared__ PackJob Js;                       // the job descriptor is read hundreds of times: keep it on chip
  if (threadIdx.x < sizeof(PackJob) / 4) reinterpret_cast<uint32_t*>(&Js)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&jobs[lo])[threadIdx.x];
  __syncthreads();
  const PackJob& J = Js;
  if (J.kind == 1) {
    // dgrad weights: dst[ci][tap'][co] = W[kh][kw][ci][co] (flipped sub-kernel) keeps co contiguous on both sides: one thread
    // converts 8 consecutive co (two 128-bit loads when the run lies inside one Keras variable, one 128-bit store)
    const int k8 = J.k_pad >> 3, ntap = J.taps_h * J.taps_w;
    const int base8 = (blockIdx.x - J.block_start) * (256 * kPackDgIter);      // 2048 * kPackDgIter elements per block
    float v[kPackDgIter][8];
#pragma unroll
    for (int it = 0; it < kPackDgIter; ++it) {                                 // all loads of the block's groups in flight before the first store
      const int i8 = base8 + it * 256 + threadIdx.x;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[it][i] = 0.f;
      if ((long long)i8 * 8 >= J.count) continue;
      const int kk = (i8 % k8) * 8;
      const int rt = i8 / k8;
      const int tap = rt % ntap, r = rt / ntap;
      const int a = tap / J.taps_w, b = tap - a * J.taps_w;
      const int kh = J.stride * (J.taps_h - 1 - a) + J.rh, kw = J.stride * (J.taps_w - 1 - b) + J.rw;
      if (r < J.Ci) {
        int j = 0, lc = kk;
        while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
        const float* src = params + J.part_w[j] + ((long long)(kh * J.KW + kw) * J.Ci + r) * J.part_n[j] + lc;
        if (lc + 8 <= J.part_n[j] && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(src)), p1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
          v[it][0] = p0.x; v[it][1] = p0.y; v[it][2] = p0.z; v[it][3] = p0.w; v[it][4] = p1.x; v[it][5] = p1.y; v[it][6] = p1.z; v[it][7] = p1.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (kk + i < J.Co) v[it][i] = master_w(J, params, kh, kw, r, kk + i);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kPackDgIter; ++it) {
      const int i8 = base8 + it * 256 + threadIdx.x;
      if ((long long)i8 * 8 >= J.count) continue;
      uint4 pk;
      pk.x = pack_bf16x2(v[it][0], v[it][1]); pk.y = pack_bf16x2(v[it][2], v[it][3]);
      pk.z = pack_bf16x2(v[it][4], v[it][5]); pk.w = pack_bf16x2(v[it][6], v[it][7]);
      reinterpret_cast<uint4*>(J.dst)[i8] = pk;
    }
    return;
  }
  if (J.kind == 0 && J.fast) {
    // forward weights, common case (one pixel per k-block, 32 | channels per k-block): a 32-channel tile lies inside one channel chunk,
    // so the k-block arithmetic is per tile instead of per element, W_hi and W_lo come from ONE fp32 read, and a block converts
    // kPackKT tiles (the generic path below spent ~70 us per optimizer segment on index divisions at 0.7 TB/s)
    __shared__ float tile[32][33];
    const int lk = J.k_pad / J.nsec;                  // logical channels per tap (padded)
    const int tkl_n = lk >> 5, tg_n = (tkl_n + kPackKT - 1) / kPackKT, tr_n = (J.rows_pad + 31) >> 5;
    int t = blockIdx.x - J.block_start;
    const int tg = t % tg_n; t /= tg_n;
    const int tr = t % tr_n;
    const int tap = t / tr_n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int co = tr * 32 + tx;
    int j = 0, lc = co;
    while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
    const long long pn = J.part_n[j];
    const float* src = params + J.part_w[j] + lc + (long long)tap * J.Ci * pn;     // (ppx == 1: tap = a * KW + b)
    const int ntap = J.taps_h * J.taps_w;
    for (int q = 0; q < kPackKT; ++q) {
      const int tk = tg * kPackKT + q;
      if (tk >= tkl_n) break;
      const int c0 = tk << 5;
      const int chunk = c0 / J.cpp, w0 = c0 - chunk * J.cpp;
      if (q) __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ci = c0 + ty + 8 * i;
        tile[ty + 8 * i][tx] = (co < J.Co && ci < J.Ci) ? src[(long long)ci * pn] : 0.f;
      }
      __syncthreads();
      {
        // thread = (section, output row, group of 8 consecutive channels): one 128-bit store each (2-byte stores were 8x the
        // store instructions); tile[ci][co] column reads: bank = (8 g + i + r) mod 32, distinct over the warp's (g, r)
        const int sec = threadIdx.x >> 7, rr = (threadIdx.x & 127) >> 2, g8 = threadIdx.x & 3;
        const int r = tr * 32 + rr;
        if (sec < J.nsec && r < J.rows_pad) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float w = tile[g8 * 8 + i][rr];
            v[i] = sec ? w - round_bf16(w) : w;
          }
          uint4 pk;
          pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
          bf16* d = (bf16*)J.dst + ((long long)r * ntap + tap) * J.k_pad + (long long)chunk * J.nsec * J.cpp + (long long)sec * J.cpp + w0 + g8 * 8;
          *reinterpret_cast<uint4*>(d) = pk;
        }
      }
    }
    return;
  }
  if (J.kind == 0) {
    // forward weights: dst[co][tap][k] = W[tap][ci(k)][co] is a transpose per tap -> 32x32 tiles through shared memory so
    // that both the fp32 reads (along co) and the bf16 writes (along k) are coalesced
    __shared__ float tile[32][33];
    const int tk_n = (J.k_pad + 31) >> 5, tr_n = (J.rows_pad + 31) >> 5;
    int t = blockIdx.x - J.block_start;
    const int tk = t % tk_n; t /= tk_n;
    const int tr = t % tr_n;
    const int tap = t / tr_n;
    const int a = tap / J.taps_w, bq = tap - a * J.taps_w;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int co = tr * 32 + tx;
    int j = 0, lc = co;
    while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
    const float* src = params + J.part_w[j] + lc;
    const int blk_len = J.ppx * J.cpp;                // one k-block: ppx pixels x cpp channels
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = tk * 32 + ty + 8 * i;
      const int blk = kk / blk_len, rem = kk - blk * blk_len;      // k-block of the tap: channel chunk blk / nsec, section blk % nsec
      const int sec = blk % J.nsec, chunk = blk / J.nsec;
      const int px = rem / J.cpp, c = chunk * J.cpp + (rem - px * J.cpp);
      const int b = bq * J.ppx + px;
      int ci = c, plane = sec;                       // plane: 0 = hi, 1 = lo, -1 = zero
      if (J.first_cat) {
        if (c < 3) { ci = c; plane = sec; }
        else if (c < 6) { ci = c - 3; plane = sec == 0 ? 0 : -1; }
        else plane = -1;
      }
      float v = 0.f;
      if (kk < J.k_pad && plane >= 0 && co < J.Co && ci < J.Ci && b < J.KW) {
        v = src[((long long)(a * J.KW + b) * J.Ci + ci) * J.part_n[j]];
        if (plane == 1) v -= round_bf16(v);
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
    const int kk = tk * 32 + tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = tr * 32 + ty + 8 * i;
      if (r < J.rows_pad && kk < J.k_pad)
        ((bf16*)J.dst)[((long long)r * (J.taps_h * J.taps_w) + tap) * J.k_pad + kk] = __float2bfloat16_rn(tile[tx][ty + 8 * i]);
    }
    return;
  }
  const int base = (blockIdx.x - J.block_start) * 2048;     // every packed operand has < 2^31 elements: 32-bit index math
  for (int t = threadIdx.x; t < 2048; t += 256) {
    const int idx = base + t;
    if (idx >= (int)J.count) return;
    if (J.kind == 2) {
      int j = 0, lc = (int)idx;
      float v = 0.f;
      if (idx < J.Co) {
        while (j + 1 < J.nparts && lc >= J.part_n[j]) { lc -= J.part_n[j]; ++j; }
        v = params[J.part_b[j] + lc];
      }
      ((float*)J.dst)[idx] = v;
      continue;
    }
    if (J.kind >= 4) {           // N-stacked conv: dst[(b, n)][(a, c)] with rows_pad = nb channels per filter column, k_pad = C_pad
      const int k_total = J.taps_h * J.k_pad;
      const int k = (int)(idx % k_total);
      int nrow = (int)(idx / k_total);
      const int nsec = J.kind == 4 && J.nsec == 2 ? 2 : 1;
      const int rows_per_sec = J.taps_w * J.rows_pad;            // J.rows_pad = channels per filter column per split
      const int rows_per_split = nsec * rows_per_sec;
      const int split = nrow / rows_per_split;
      nrow -= split * rows_per_split;
      const int sec = nrow / rows_per_sec;                       // 0: W_hi rows, 1: W_lo rows (bf16x3 pair mode)
      nrow -= sec * rows_per_sec;
      const int b = nrow / J.rows_pad, nn = split * J.rows_pad + (nrow - b * J.rows_pad);
      const int a = k / J.k_pad, c = k - a * J.k_pad;
      float v = 0.f;
      if (J.kind == 4) { if (nn < J.Co && c < J.Ci) v = master_w(J, params, a, b, c, nn); }            // fwd: n = co, k = ci
      else if (nn < J.Ci && c < J.Co) v = master_w(J, params, J.KH - 1 - a, J.KW - 1 - b, nn, c);      // dgrad: n = ci, k = co, flipped
      if (sec) v -= round_bf16(v);
      ((bf16*)J.dst)[idx] = __float2bfloat16_rn(v);
      continue;
    }
    const int kk = (int)(idx % J.k_pad);
    const int tap = (int)((idx / J.k_pad) % (J.taps_h * J.taps_w));
    const int r = idx / (J.k_pad * J.taps_h * J.taps_w);
    const int a = tap / J.taps_w, b = tap % J.taps_w;
    float v = 0.f;
    {                            // dgrad (generic path): rows = ci, k = co, flipped (sub-)kernel
      const int kh = J.stride * (J.taps_h - 1 - a) + J.rh, kw = J.stride * (J.taps_w - 1 - b) + J.rw;
      if (r < J.Ci && kk < J.Co) v = master_w(J, params, kh, kw, r, kk);
    }
    ((bf16*)J.dst)[idx] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: TMA descriptors
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

CUtensorMapSwizzle swz(int bytes) {
  return bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                                                                            : CU_TENSOR_MAP_SWIZZLE_NONE;
}

// activation tensor [N][H][W][ld] (bf16), channels [coff, coff+C): box = {bk, tw*s, th*s, tn}, element strides {1,s,s,1}
const char* make_act_map(CUtensorMap* m, const void* base, int N, int H, int W, int ld, int coff, int C, int bk, int tw, int th, int tn,
                         int s, int swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(tw * s), (cuuint32_t)(th * s), (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
  if (s > 1) { box[1] -= (s - 1); box[2] -= (s - 1); }  // ceil(box/stride) elements are loaded
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)((const bf16*)base + coff), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(act) failed: %d (dims %d,%d,%d,%d ld %d box %u,%u,%u,%u s %d)", (int)r,
             C, W, H, N, ld, box[0], box[1], box[2], box[3], s);
    return g_tc_error;
  }
  return nullptr;
}

// halo tile of an activation tensor [N][H][W][ld]: twp plane columns at element stride sx, thp consecutive rows
const char* make_halo_map(CUtensorMap* m, const void* base, int N, int H, int W, int ld, int coff, int C, int bk, int twp, int thp, int sx,
                          int swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
  cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(twp * sx - (sx - 1)), (cuuint32_t)thp, 1};   // ceil(box/stride) elements are loaded
  cuuint32_t estr[4] = {1, (cuuint32_t)sx, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)((const bf16*)base + coff), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, swz(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(halo) failed: %d (dims %d,%d,%d,%d ld %d box %u,%u,%u sx %d)", (int)r,
             C, W, H, N, ld, box[0], box[1], box[2], sx);
    return g_tc_error;
  }
  return nullptr;
}

// packed weights [rows][k_total] bf16: box = {bk, tile_rows}
const char* make_w_map(CUtensorMap* m, const void* base, int rows, long long k_total, int bk, int tile_rows, int swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  cuuint64_t dims[2] = {(cuuint64_t)k_total, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)k_total * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)tile_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(weights) failed: %d (rows %d k %lld bk %d tile %d)", (int)r, rows,
             k_total, bk, tile_rows);
    return g_tc_error;
  }
  return nullptr;
}

// packed weights seen as [rows][blocks][bk]: box = {bk, tile_rows, nblk} = `nblk` consecutive k-blocks of `tile_rows` rows in ONE TMA
// operation, landing block-major in shared memory (exactly the ring layout of halo_conv_kernel).  The TMA unit works through small
// boxes one after the other (~700 cycles per 4 KB box of 32 rows: 6 B/clk per SM, phase trace of d4 bf16x3); large boxes stream.
const char* make_w_map3(CUtensorMap* m, const void* base, int rows, long long k_total, int bk, int tile_rows, int nblk, int swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  cuuint64_t dims[3] = {(cuuint64_t)bk, (cuuint64_t)rows, (cuuint64_t)(k_total / bk)};
  cuuint64_t strides[2] = {(cuuint64_t)k_total * 2, (cuuint64_t)bk * 2};
  cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)tile_rows, (cuuint32_t)nblk};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swz(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(weights3) failed: %d (rows %d k %lld bk %d tile %d x %d)", (int)r, rows,
             k_total, bk, tile_rows, nblk);
    return g_tc_error;
  }
  return nullptr;
}

// first-layer input as overlapping 8-pixel windows: element (k, wo, y, n) = xp[n][y][2*wo + k/8][k%8]
const char* make_window_map(CUtensorMap* m, const void* xp, int N, int H, int W, int Wo, int tw, int th, int tn) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  const cuuint64_t row_pitch = (cuuint64_t)(W + 8) * 16;
  cuuint64_t dims[4] = {64, (cuuint64_t)Wo, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {32, row_pitch, row_pitch * H};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)(th * 2 - 1), (cuuint32_t)tn};
  cuuint32_t estr[4] = {1, 1, 2, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)xp, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(window) failed: %d (H %d W %d Wo %d box %u,%u,%u)", (int)r, H, W, Wo, box[1],
             box[2], box[3]);
    return g_tc_error;
  }
  return nullptr;
}

// natural pixel pairs of the staged first-layer image [N][H][W + 8][8] (pixel x in column x + 2): [N][H][W / 2][16], 32-byte swizzle
const char* make_first_pair_map(CUtensorMap* m, const void* xp, int N, int H, int W, int tw, int th) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return "cuTensorMapEncodeTiled unavailable";
  const cuuint64_t row_pitch = (cuuint64_t)(W + 8) * 16;
  cuuint64_t dims[4] = {16, (cuuint64_t)(W / 2), (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {32, row_pitch, row_pitch * H};
  cuuint32_t box[4] = {16, (cuuint32_t)tw, (cuuint32_t)th, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)((const bf16*)xp + 16), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_tc_error, sizeof(g_tc_error), "cuTensorMapEncodeTiled(first pair) failed: %d (H %d W %d box %u,%u)", (int)r, H, W, box[1], box[2]);
    return g_tc_error;
  }
  return nullptr;
}

__global__ void __launch_bounds__(256) stage_first_kernel(const float* __restrict__ inputs, bf16* __restrict__ xp, int coff, int B, int H, int W,
                                                          int split) {
  pdl_enter();
  const long long total = (long long)B * H * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % W);
    const long long row = idx / W;  // n*H + y
    const float* ip = inputs + idx * 6 + coff;
    const float v0 = ip[0], v1 = ip[1], v2 = ip[2];
    uint4 pk;
    pk.x = pack_bf16x2(v0, v1);
    if (split) {      // bf16 pairs inside the pixel: channels 0-2 hi, 3-5 lo = bf16(v - hi)
      const float l0 = v0 - round_bf16(v0), l1 = v1 - round_bf16(v1), l2 = v2 - round_bf16(v2);
      pk.y = pack_bf16x2(v2, l0);
      pk.z = pack_bf16x2(l1, l2);
    } else {
      pk.y = pack_bf16x2(v2, 0.f);
      pk.z = 0u;
    }
    pk.w = 0u;
    *reinterpret_cast<uint4*>(xp + (row * (W + 8) + x + 2) * 8) = pk;
  }
}

int round_up(int v, int m) { return (v + m - 1) / m * m; }
bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// K chunking for a channel count: returns bk (64/32/16) and the padded channel count
void choose_bk(int C, int& bk, int& c_pad) {
  if (C >= 64) { bk = 64; c_pad = round_up(C, 64); }
  else if (C > 16) { bk = 32; c_pad = 32; }
  else { bk = 16; c_pad = 16; }
}

int pad_cols(int n) { return n <= 16 ? 16 : n <= 32 ? 32 : n <= 64 ? 64 : round_up(n, 128); }

// fills the geometry part of a launch for an output grid GH x GW (per image)
bool tile_grid(TcLaunch& L, int GH, int GW, int n_img) {
  if (!is_pow2(GW) || !is_pow2(GH) || GW > 128) return false;
  L.tile_w = GW;
  L.tile_h = GH < 128 / GW ? GH : 128 / GW;
  L.tile_n_img = 128 / (L.tile_w * L.tile_h);
  L.grid_h = GH; L.grid_w = GW; L.n_img = n_img;
  return true;
}

int env_int(const char* name, int dflt);
// logical / physical chunk bookkeeping of a launch (TcLaunch::split: 0 plain, 1 bf16 pairs, 2 first layer with in-pixel pairs)
void set_chunks(TcLaunch& L) {
  if (L.split == 1) { L.kcl = 3 * L.kc; L.kca = 2 * L.kc; L.kcb = 2 * L.kc; }
  else if (L.split == 2) { L.kcl = 2; L.kca = 1; L.kcb = 2; }      // (kc == 1)
  else { L.kcl = L.kca = L.kcb = L.kc; }
}
void finish_launch(TcLaunch& L, int n_cols_pad, int concurrent = 1) {   // concurrent: launches of this shape sharing the GPU
  set_chunks(L);
  L.tile_cols = n_cols_pad < 128 ? n_cols_pad : 128;
  L.n_tiles = n_cols_pad / L.tile_cols;
  L.fat = (L.split == 1 && env_int("SV_IGEMM_FAT", 1)) ? 1 : 0;
  const int a_bytes = (L.fat ? 2 : 1) * 128 * L.bk * 2, b_bytes = (L.fat ? 2 : 1) * round_up(L.tile_cols * L.bk * 2, 1024);
  // ring budget: 100 KB keeps two CTAs per SM; a launch with at most one CTA per SM anyway (the 8x8-pixel layers: 128 CTAs)
  // takes the whole shared memory for a deeper ring instead
  const int tiles_per_img = L.tile_h > 0 ? L.grid_h / L.tile_h : 1;
  const long long m_tiles = L.tile_n_img > 1 ? (L.n_img + L.tile_n_img - 1) / L.tile_n_img : (long long)L.n_img * tiles_per_img;
  const bool one_wave = m_tiles * L.n_tiles * concurrent <= 148 && env_int("SV_IGEMM_DEEP", 1);
  int stages = ((one_wave ? 192 : 100) * 1024) / (a_bytes + b_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  const int num_kb = L.taps_h * L.taps_w * (L.fat ? L.kc : L.kcl);
  if (stages > num_kb) stages = num_kb < 1 ? 1 : num_kb;
  L.stages = stages;
  L.smem_bytes = (size_t)stages * (a_bytes + b_bytes) + sizeof(SmemCtl) + 1024;
}


int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

// Split-K plan for a dense layer (1x1 "image", one tap) with few output tiles and a long K axis: heads (K = 8192 -> 256),
// d1 dgrad, the GM y_block; without it 2-4 CTAs stream the whole weight matrix at TMA latency (78 us for 4 MB).
void plan_split_k(TcLaunch& L, int n_img, size_t& off) {
  L.k_splits = 1; L.kb_per_split = 0; L.partial = nullptr;
  if (env_int("SV_NO_SPLITK", 0)) return;
  if (L.halo || L.grid_h != 1 || L.grid_w != 1 || L.taps_h * L.taps_w != 1 || L.osy != 1 || L.osx != 1) return;
  const int num_kb = L.fat ? L.kc : L.kcl;
  const int m_tiles = (n_img + 127) / 128, ctas = m_tiles * L.n_tiles;
  if (num_kb < 8 || ctas >= 74) return;
  int splits = (148 + ctas - 1) / ctas;
  if (splits > num_kb / 2) splits = num_kb / 2;
  const int kbps = (num_kb + splits - 1) / splits;
  splits = (num_kb + kbps - 1) / kbps;
  if (splits < 2) return;
  L.k_splits = splits; L.kb_per_split = kbps;
  L.m_pad = m_tiles * 128; L.n_pad = L.n_tiles * L.tile_cols;
  if (L.stages > kbps) L.stages = kbps;
  off = (off + 1023) / 1024 * 1024;
  // (the caller records `off` as the partial-buffer offset)
}
size_t split_k_bytes(const TcLaunch& L) {
  return L.k_splits > 1 ? ((size_t)L.k_splits * L.m_pad * L.n_pad * 4 + 1023) / 1024 * 1024 : 0;
}

// Converts an igemm launch (stride-1 A addressing) into a halo-resident launch when the output grid allows 8x16 row
// groups.  Tile choice: minimise estimated L2->SMEM bytes per output pixel (halo + streamed weights), with a 25 % penalty
// for configurations that leave a single CTA per SM (no cross-CTA overlap of the load / MMA / epilogue phases).

void try_halo(TcLaunch& L, int GH, int GW, int n_img, int sx = 1, int sy = 1, bool force = false, size_t max_halo_bytes = 0,
              int tmem_limit = 512) {
  L.halo = 0;
  L.nstack2 = 0;
  L.halo_sx = 1; L.halo_sy = 1;
  if (env_int("SV_NO_HALO", 0)) return;
  if ((sx == 1 && sy == 1 && L.a_stride != 1) || (GH % 16) || (GW % 8) || L.taps_h * L.taps_w < 2) return;
  // measured on B200 (scripts/bench_layers.py): the halo kernel wins where the per-tap kernel is L2-bound with little
  // tensor work per byte (<= 32 channels per pixel and N <= 32: d5 forward 150 vs 221 us, d5 dgrad 103 vs 234 us); with 64+ channel
  // chunks the per-tap kernel (2 CTAs/SM, deep ring) is as fast or faster (d4 forward 97 vs 130 us) -> keep it there.
  // `force` (strided forward layers, stride-2 dgrad classes): the per-tap kernel re-reads every activation tile once per tap
  // from L2 (e2 forward: 226 MB in 31 us), the halo kernel reads it once.
  if ((L.bk > 32 || L.tile_cols > 32) && !force && !env_int("SV_HALO_ALL", 0)) return;
  set_chunks(L);
  const int pix = L.bk * 2, nch = L.kca;                 // physical chunks held by the halo
  const int kb_bytes = L.tile_cols * L.bk * 2;
  const int num_kb = L.taps_h * L.taps_w * L.kcb;        // physical weight k-blocks, each streamed once per tile
  const double w_total = (double)num_kb * kb_bytes;
  int KB = 8192 / kb_bytes;
  if (KB < 1) KB = 1;
  if (KB > num_kb) KB = num_kb;
  // bf16x3: one CTA per SM anyway (two halo planes), and the weight stream is the bottleneck (latency-bound with a 24 KB ring:
  // d4 562 us) -> the ring takes all the shared memory the halo leaves, in up to kMaxStages stages
  const bool deep = L.split != 0;
  const size_t smem_cap = deep ? 226 * 1024 : 200 * 1024;
  int best_tw = 0, best_th = 0, best_stages = 0;
  double best_cost = 1e30;
  const int force_tw = env_int("SV_HALO_TW", 0), force_th = env_int("SV_HALO_TH", 0);
  const int wtaps = (L.taps_w + sx - 1) / sx;
  for (int TH = 16; TH <= 32 && TH <= GH; TH += 16) {
    if (GH % TH) continue;
    for (int TW = 8; TW <= 64 && TW <= GW; TW += 8) {
      if (GW % TW) continue;
      if (force_tw && TW != force_tw) continue;
      if (force_th && TH != force_th) continue;
      const int MT = (TW / 8) * (TH / 16);
      if (MT * L.tile_cols > tmem_limit) continue;
      const int TWp = TW + wtaps - 1, THp = (TH - 1) * sy + L.taps_h;
      if (TWp * sx > 256 || THp > 256) continue;
      const size_t chunk = ((size_t)THp * TWp * pix + 1023) / 1024 * 1024;
      if (max_halo_bytes && chunk * nch * sx > max_halo_bytes) continue;
      for (int stages = 3; stages >= 2; --stages) {
        const size_t smem = chunk * nch * sx + (size_t)stages * KB * kb_bytes + sizeof(HaloCtl) + 1024;
        if (smem > smem_cap) continue;
        if (deep && chunk * nch * sx + (size_t)env_int("SV_HALO_MIN_RING_KB", 64) * 1024 > smem_cap) continue;   // leave room for the weight ring
        int tmem_cols = 32;
        while (tmem_cols < MT * L.tile_cols) tmem_cols <<= 1;
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
        if (per_sm < 1) continue;
        const double traffic = ((double)chunk * nch * sx + w_total) / (TW * TH);
        double cost = traffic * (per_sm == 1 ? 1.6 : 1.0);
        if (TW == 32 && TH == 16) cost *= 0.5;   // measured best shape for the 64x64 layers
        if (cost < best_cost) { best_cost = cost; best_tw = TW; best_th = TH; best_stages = stages; }
        break;
      }
    }
  }
  if (!best_tw) return;
  L.halo = 1;
  L.trace = env_int("SV_HALO_TRACE", 0);
  L.halo_sx = sx; L.halo_sy = sy;
  L.TW = best_tw; L.TH = best_th; L.TWp = best_tw + wtaps - 1; L.THp = (best_th - 1) * sy + L.taps_h;
  L.mtx = best_tw / 8; L.mty = best_th / 16;
  L.chunk_bytes = (int)(((size_t)L.THp * L.TWp * pix + 1023) / 1024 * 1024);
  L.kb_per_stage = KB; L.w_stages = best_stages; L.w_stage_bytes = KB * kb_bytes;
  if (deep) {     // widen / deepen the ring into the free shared memory: <= kMaxStages stages of an even number of k-blocks
    const size_t free_bytes = smem_cap - ((size_t)L.chunk_bytes * nch * sx + sizeof(HaloCtl) + 1024);
    int total_kb = (int)(free_bytes / kb_bytes);
    if (total_kb > num_kb) total_kb = num_kb;
    int st = kMaxStages, per = total_kb / st;
    while (st > 2 && per < 2) { --st; per = total_kb / st; }
    if (per > 1) per -= per & 1;
    if (per >= 1) { L.kb_per_stage = per; L.w_stages = st; L.w_stage_bytes = per * kb_bytes; }
  }
  // N-stacked weight pairs (one MMA of 2N columns for A_hi x [W_hi ; W_lo]): two MMAs per k-step instead of three where the MMA rate
  // is bound by the A-operand read (N <= 64)
  L.nstack2 = (L.split == 1 && (L.kb_per_stage % 2) == 0 && 2 * L.tile_cols <= 256 && 2 * (best_tw / 8) * (best_th / 16) * L.tile_cols <= 512 &&
               env_int("SV_NSTACK2", 1)) ? 1 : 0;
  L.w_box3 = 0;
  if (L.nstack2) {
    // one ring stage = a whole number of filter taps ([W_hi(ch) W_lo(ch)] x kc each; the kernel's unrolled per-tap issue sequence),
    // loaded by ONE 3-D TMA box: ~24 KB stages, >= 3 of them
    const size_t free_bytes = smem_cap - ((size_t)L.chunk_bytes * nch * sx + sizeof(HaloCtl) + 1024);
    const size_t tap_bytes = (size_t)L.kcb * kb_bytes;
    const int taps = L.taps_h * L.taps_w;
    int tps = 1;
    for (int c = 1; c <= taps; ++c)
      if (taps % c == 0 && c * tap_bytes <= 24 * 1024 && free_bytes / (c * tap_bytes) >= 3 && c * L.kcb <= 256) tps = c;
    int st = (int)(free_bytes / (tps * tap_bytes));
    if (st > kMaxStages) st = kMaxStages;
    if (st >= 2) {
      L.kb_per_stage = tps * L.kcb; L.w_stages = st; L.w_stage_bytes = (int)(tps * tap_bytes);
      L.w_box3 = env_int("SV_W_BOX3", 1);
    }
  }
  L.tiles_x = GW / best_tw; L.tiles_y = GH / best_th;
  L.n_img = n_img;
  L.smem_bytes = (size_t)L.chunk_bytes * nch * sx + (size_t)L.w_stages * L.w_stage_bytes + sizeof(HaloCtl) + 1024;
}


// Upgrades a halo launch to the persistent pipelined kernel when the packed weights fit beside two halo stages and two
// accumulator sets fit in TMEM.  `n_classes` launches share the GPU (stride-2 dgrad: 4 parity classes, 37 CTAs each).
void plan_persist(TcLaunch& L, int n_classes) {
  L.persist = 0;
  if (!L.halo || !env_int("SV_PCONV", 1)) return;
  const int MT = L.mtx * L.mty;
  if (L.n_tiles > 4 || (L.n_tiles > 1 && n_classes != L.n_tiles) || 2 * MT * L.tile_cols > 512 || (L.tile_cols % kEpiBW) || L.nparts != 1) return;
  set_chunks(L);
  const int num_kb = L.taps_h * L.taps_w * L.kcb, kb_bytes = L.tile_cols * L.bk * 2;     // resident: the physical k-blocks
  const size_t w_bytes = ((size_t)num_kb * kb_bytes + 1023) / 1024 * 1024;
  const size_t stage_bytes = (size_t)L.kca * L.halo_sx * L.chunk_bytes;
  const size_t budget = 225 * 1024 - 1024 - sizeof(PcCtl);
  if (w_bytes + 2 * stage_bytes > budget) return;
  int nst = (int)((budget - w_bytes) / stage_bytes);
  const int cap = env_int("SV_PCONV_STAGES", kPcMaxStages);
  if (nst > cap) nst = cap;
  if (nst > kPcMaxStages) nst = kPcMaxStages;
  const int tiles = L.tiles_x * L.tiles_y * L.n_img;
  int grid = env_int("SV_PCONV_GRID", 148) / n_classes;
  if (grid > tiles) grid = tiles;
  if (grid < 1) grid = 1;
  L.persist = 1; L.p_stages = nst; L.p_wbytes = (int)w_bytes; L.p_grid = grid; L.p_ntile = 0;
  L.nstack2 = 0;                         // (the persistent kernel issues the three products separately)
  L.w_box3 = 0;
  L.p_smem = w_bytes + (size_t)nst * stage_bytes + sizeof(PcCtl) + 1024;
}

// Plans the halo-resident wgrad for a stride-1 convolution; returns false when the layer is not eligible.
bool plan_halo_wgrad(TcHaloWgrad& H, const ConvGeom& g0, int cb, int cipad, int cbn, int copad) {
  if (env_int("SV_NO_HALO_WGRAD", 0)) return false;
  ConvGeom g = g0;
  H.sy = 1; H.pair_c = 0; H.pair_sh = 0;
  if (g0.stride == 2) {
    // Stride 2 through the pixel-PAIR view of X: [H][W/2][2 * ld] - consecutive output pixels are consecutive pairs, the filter column
    // b = 2 q + r (relative to the padding) becomes pair tap q and selects the parity-r half of the pair's channels, so horizontally
    // this is a stride-1 layer with 2 * Ci channels and ceil-ish(kw / 2) + 1 taps; vertically the MMA loop steps two X rows per
    // output row.  (The per-tap kernel re-read the input once per tap: e2's 8 MB input became ~290 MB of L2 -> SMEM traffic.)
    if (env_int("SV_NO_PAIR_WGRAD", 0)) return false;
    if (g0.in_coff || g0.in_ld != g0.Ci || (g0.Ci % 8) || (g0.Wi & 1) || g0.Wo * 2 != g0.Wi || g0.Ho * 2 != g0.Hi) return false;
    const int plp = (g0.pl + 1) / 2;
    H.pair_c = g0.Ci;
    H.pair_sh = 2 * plp - g0.pl;                                  // u = b - pl + 2 plp >= 0
    g.kw = (g0.kw - 1 + H.pair_sh) / 2 + 1;
    g.pl = plp;
    g.Ci = 2 * g0.Ci; g.in_ld = 2 * g0.in_ld;
    g.Wi = g0.Wi / 2; g.Hi = g.Ho;                                // (checked below as a same-size layer)
    g.stride = 1;
    choose_bk(g.Ci, cb, cipad);
    H.sy = 2;
  }
  const int sy = H.sy;
  const bool narrow = g.Wo == 8 && (g.Ho % 2) == 0 && !env_int("SV_NO_NARROW_WGRAD", 0);     // 8-pixel-wide images: a K step = 8 columns x 2 rows
  if (g.stride != 1 || ((g.Wo % 16) && !narrow) || g.Wo != g.Wi || g.Ho != g.Hi || copad > 256 || g.kh * g.kw < 2) return false;
  H.krows = narrow ? 2 : 1;
  H.taps_h = g.kh; H.taps_w = g.kw; H.pad_t = g.pt; H.pad_l = g.pl;
  H.cb = cb; H.nchunks = cipad / cb; H.x_swizzle = cb * 2;
  H.cbn = cbn; H.nbchunks = copad / cbn; H.dy_swizzle = cbn * 2;
  H.nsub = 128 / cb;
  // N-stacked variant: dY must be a single swizzle atom per pixel and kw blocks must fit one MMA
  H.nstack = (H.nbchunks == 1 && g.kw * copad <= 256 && !env_int("SV_NO_NSTACK", 0)) ? 1 : 0;
  H.nb = copad;
  H.n_pad = H.nstack ? g.kw * copad : copad;
  if (H.nchunks == 1) {
    H.mode = 0;
    H.gw = H.nstack ? 0 : (g.kw + H.nsub - 1) / H.nsub;
    H.gpt = 0;
    H.groups = H.nstack ? (g.kh + H.nsub - 1) / H.nsub : g.kh * H.gw;
  } else {
    if (H.nchunks % H.nsub) return false;
    H.mode = 1;
    H.gpt = H.nchunks / H.nsub;
    H.gw = 0;
    H.groups = (H.nstack ? g.kh : g.kh * g.kw) * H.gpt;
  }
  if (H.groups > 32) return false;
  H.m_pad = H.groups * 128;
  int gpc = 512 / H.n_pad;
  if (gpc > H.groups) gpc = H.groups;
  H.m_splits = (H.groups + gpc - 1) / gpc;
  H.groups_per_cta = (H.groups + H.m_splits - 1) / H.m_splits;   // balanced
  // tile: TW = 32 when it divides the width (16 otherwise); TH = largest divisor of the height whose two stages fit
  const int force_th = env_int("SV_HWG_TH", 0), force_tw = env_int("SV_HWG_TW", 0);
  H.TW = narrow ? 8 : force_tw ? force_tw : ((g.Wo % 32) == 0 ? 32 : 16);
  if (g.Wo % H.TW) return false;
  H.stages = env_int("SV_HWG_STAGES", 2);
  if (H.stages < 2) H.stages = 2;
  if (H.stages > 4) H.stages = 4;
  const int x_tw = H.nstack ? H.TW : H.TW + g.kw - 1, dy_tw = H.nstack ? H.TW + g.kw - 1 : H.TW;
  // garbage sub-blocks of a partially filled group read up to (nsub-1) rows (N-stack) / pixels (plain) past the X tile
  const size_t slack = H.nstack ? (size_t)(H.nsub - 1) * x_tw * cb * 2 : (size_t)H.nsub * cb * 2;
  int best_th = 0;
  for (int th = 1; th <= g.Ho && th <= 32; ++th) {
    if (g.Ho % th || th % H.krows) continue;
    if (force_th && th != force_th) continue;
    const int thp = (th - 1) * sy + g.kh;
    const size_t xc = ((size_t)thp * x_tw * cb * 2 + 1023) / 1024 * 1024, dc = ((size_t)th * dy_tw * cbn * 2 + 1023) / 1024 * 1024;
    const size_t smem = (size_t)H.stages * (xc * H.nchunks + dc * H.nbchunks) + slack + sizeof(HwCtl) + 1024;
    if (smem <= 190 * 1024) best_th = th;
  }
  if (!best_th) return false;
  H.TH = best_th;
  H.x_tw = x_tw; H.x_th = (H.TH - 1) * sy + g.kh; H.dy_tw = dy_tw; H.dy_th = H.TH;
  H.x_dx = H.nstack ? 0 : -g.pl; H.x_dy = -g.pt; H.dy_dx = H.nstack ? -(g.kw - 1) + g.pl : 0;
  H.x_chunk_bytes = (int)(((size_t)H.x_th * H.x_tw * cb * 2 + 1023) / 1024 * 1024);
  H.dy_chunk_bytes = (int)(((size_t)H.dy_th * H.dy_tw * cbn * 2 + 1023) / 1024 * 1024);
  H.stage_bytes = H.x_chunk_bytes * H.nchunks + H.dy_chunk_bytes * H.nbchunks;
  H.smem_bytes = (size_t)H.stages * H.stage_bytes + slack + sizeof(HwCtl) + 1024;
  const int pix = cb * 2;
  H.a_lbo = H.mode == 1 ? H.x_chunk_bytes : (H.nstack ? H.x_tw * pix : pix);
  H.b_lbo = H.nstack ? cbn * 2 : H.dy_chunk_bytes;
  for (int gi = 0; gi < H.groups; ++gi) {
    long long off;
    if (H.nstack) {
      if (H.mode == 0) off = (long long)(gi * H.nsub) * H.x_tw * pix;
      else { const int a = gi / H.gpt, c0 = (gi % H.gpt) * H.nsub; off = (long long)c0 * H.x_chunk_bytes + (long long)a * H.x_tw * pix; }
    } else {
      if (H.mode == 0) { const int a = gi / H.gw, b0 = (gi % H.gw) * H.nsub; off = (long long)(a * H.x_tw + b0) * pix; }
      else { const int tap = gi / H.gpt, c0 = (gi % H.gpt) * H.nsub; off = (long long)c0 * H.x_chunk_bytes + (long long)((tap / g.kw) * H.x_tw + tap % g.kw) * pix; }
    }
    H.goff[gi] = (uint32_t)(off >> 4);
  }
  H.tiles_x = g.Wo / H.TW; H.tiles_y = g.Ho / H.TH; H.n_img = g.B;
  H.tiles = H.tiles_x * H.tiles_y * g.B;
  // CTAs per launch (the weight gradients share the GPU with the dgrad chain and with each other).  Measured on the C2 step once every
  // convolution's weight gradient ran on this kernel: stride-1 layers 74 -> 1.731 ms, 56 -> 1.711, 37 -> 1.695, 28 -> 1.681, 20 -> 1.679
  // (with 28 on the pair-view layers); pair-view layers 111 -> 1.718, 74 -> 1.695, 37 -> 1.681, 28 -> 1.667 = 37 within noise.
  int ks = (H.pair_c ? env_int("SV_HWG_SPLITS_S2", 37) : env_int("SV_HWG_SPLITS", 28)) / H.m_splits;
  if (ks < 1) ks = 1;
  if (ks > H.tiles) ks = H.tiles;
  H.tiles_per_split = (H.tiles + ks - 1) / ks;
  H.k_splits = (H.tiles + H.tiles_per_split - 1) / H.tiles_per_split;
  return true;
}

// Plans the N-stacked persistent kernel for a stride-1 convolution C -> n_out over an H x W image (W in {16, 32, 64}).
// Returns false when the layer is not eligible (then the per-tap / halo kernels are used).
bool plan_nsconv(TcNsConv& P, int kh, int kw, int pad_t, int pad_l, int H, int W, int n_img, int C, int n_out, bool pair = false) {
  if (env_int("SV_NO_NSCONV", 0)) return false;
  if (pair && !env_int("SV_NS_PAIR", 1)) return false;
  P.pair = pair ? 1 : 0;
  if (!(kw == 4 || kw == 6) || pad_l > 4 || kw - 1 - pad_l > 4) return false;
  if (!(W == 16 || W == 32 || W == 64)) return false;
  const int R = 128 / W;
  if (H % R) return false;
  int ck, cpad;
  choose_bk(C, ck, cpad);
  const int pixB = ck * 2, nchunks = cpad / ck, num_kb = kh * nchunks;
  int nb_all = round_up(n_out, 8);
  if ((kw * nb_all) % 16) nb_all = round_up(n_out, 16);
  // Blocks per tile: a tile of mb * R rows shares one (mb*R + kh - 1)-row halo.  With R = 2 (64-wide images) a single block
  // re-reads its input 3.5x through the 7-row halo of a 6x6 filter and d5 forward is L2->SMEM bound (235 MB per launch); two
  // blocks per tile cut that to 2.25x.  Needs two accumulators per buffer in TMEM (two buffers: 4 * N <= 512) and one channel chunk.
  int mb = 1;
  {
    const bool can2 = nchunks == 1 && (H % (2 * R)) == 0 && 4 * kw * nb_all <= 512;
    if (R + kh - 1 > 3 * R && can2) mb = 2;
    const int f = env_int("SV_NS_MB", 0);
    if (f == 1 || (f == 2 && can2)) mb = f;
    // four blocks per tile (a 21-row halo for 16 rows of a 6x6 filter): 1.31x re-read; two accumulator buffers of 4 blocks in TMEM
    if (f == 4 && nchunks == 1 && (H % (4 * R)) == 0 && 8 * kw * nb_all <= 512) mb = 4;
  }
  // split the output channels over CTAs when the resident weights would leave room for fewer than 4 halo stages
  const int chunk_bytes = round_up((mb * R + kh - 1) * W * pixB, 1024), stage_bytes = chunk_bytes * nchunks;
  const size_t xch_bytes = W > 32 ? (size_t)(kNsEpiWarps / 4) * 2 * 4 * kw * kNsXchSlots * 8 * 4 : 0;   // one area per (group, chunk lane)
  const size_t budget = 227 * 1024 - 1024 - 256 - xch_bytes;
  // (measured: the split costs more than it buys whenever the whole weight set fits beside two halo stages - d4 forward
  // 35.6 us unsplit vs 41.8 us split - so it is only used where the layer would otherwise not fit at all: d3)
  int co_splits = 1;
  if (pair) {
    // bf16x3: resident [W_hi-stack ; W_lo-stack] k-blocks (twice the bytes) beside >= 2 single-plane halo stages, and the hi pass is ONE
    // MMA of 2 * kw * nb <= 256 columns: split the output channels over 2 / 4 CTA classes until both hold
    mb = 1;
    const int want = env_int("SV_NS_PAIR_SPLITS", 0);
    for (co_splits = 1; co_splits <= 8; co_splits *= 2) {
      if (nb_all % (8 * co_splits)) return false;
      const int nbs = nb_all / co_splits;
      if ((kw * nbs) % 16 || 2 * kw * nbs > 256) continue;
      if (want && co_splits != want) continue;
      if ((size_t)2 * kw * nbs * num_kb * pixB + 2 * (size_t)round_up((R + kh - 1) * W * pixB, 1024) * nchunks <= budget) break;
    }
    if (co_splits > 8) return false;
  } else
  if ((size_t)kw * nb_all * num_kb * pixB + 2 * (size_t)stage_bytes > budget && (nb_all % 16) == 0 && ((kw * nb_all / 2) % 16) == 0 &&
      !env_int("SV_NS_NOSPLIT", 0))
    co_splits = 2;
  // N wider than one MMA (d4 dgrad: 6 x 64 = 384 columns) leaves room for a single accumulator set in TMEM, so the MMAs and the
  // epilogue of a tile serialise; splitting the output channels over CTA pairs halves N and restores the double buffering
  if (!pair && co_splits == 1 && kw * nb_all > 256 && (nb_all % 16) == 0 && ((kw * nb_all / 2) % 16) == 0 && env_int("SV_NS_SPLIT_WIDE", 1))   // d4 dgrad 55 -> 45 us, 1.538 -> 1.514 ms/step
    co_splits = 2;
  const int nb = nb_all / co_splits;
  const int n_total = kw * nb;
  const int chunk_bytes_f = round_up((mb * R + kh - 1) * W * pixB, 1024), stage_bytes_f = chunk_bytes_f * nchunks;   // (pair mode forces mb = 1)
  int ng = 1;
  if (n_total > 256) {
    if (pair || (kw % 2) || n_total / 2 > 256 || (n_total / 2) % 16) return false;
    ng = 2;
  }
  P.kh = kh; P.kw = kw; P.pad_t = pad_t; P.pad_l = pad_l;
  P.W = W; P.R = R; P.H = H; P.n_img = n_img;
  P.mb = mb;
  P.tiles_per_img = H / (R * mb); P.tiles = n_img * P.tiles_per_img;
  P.ck = ck; P.nchunks = nchunks; P.pixB = pixB;
  P.nb = nb; P.ng = ng; P.ncols = n_total / ng; P.n_total = n_total; P.co_splits = co_splits;
  // epilogue warp sets: `groups` accumulator buffers / tile round-robin, `lanes` warps per TMEM quarter splitting the 8-channel blocks
  P.acc1 = pair ? 2 * n_total : n_total;          // (ng * ncols == n_total)
  int groups = 512 / (P.acc1 * mb);
  if (groups > kNsMaxGroups) groups = kNsMaxGroups;
  if (groups < 1) return false;
  const int sets = kNsEpiWarps / 4;
  int lanes = sets / groups;
  if (lanes > nb / 8) lanes = nb / 8;
  if (lanes < 1) lanes = 1;
  { const int fg = env_int("SV_NS_GROUPS", 0), fl = env_int("SV_NS_LANES", 0);
    if (fg >= 1 && fg <= groups) groups = fg;
    if (fl >= 1 && fl * groups <= sets) lanes = fl; }
  P.groups = groups; P.lanes = lanes;
  P.nacc = groups;
  P.halo_rows = mb * R + kh - 1;
  P.chunk_bytes = chunk_bytes_f;
  P.stage_bytes = stage_bytes_f;
  P.num_kb = num_kb;
  P.wk_bytes = (pair ? 2 : 1) * n_total * pixB;
  if (P.wk_bytes % 1024 || (n_total * pixB) % 1024) return false;            // every weight k-block (and its W_lo half) starts on a swizzle-atom boundary
  P.w_bytes = round_up(P.num_kb * P.wk_bytes, 1024);
  P.n_box = pair ? 2 * n_total : n_total <= 256 ? n_total : P.ncols;
  P.xch_off = 256;                                // after NsCtl
  if ((size_t)P.w_bytes + 2 * (size_t)stage_bytes_f > budget) return false;
  int nst = (int)((budget - P.w_bytes) / stage_bytes_f);
  const int cap = env_int("SV_NS_STAGES", 6);
  if (nst > cap) nst = cap;
  if (nst > kNsMaxStages) nst = kNsMaxStages;
  P.nstages = nst;
  P.smem_bytes = (size_t)P.w_bytes + (size_t)nst * stage_bytes_f + 256 + xch_bytes + 1024;
  P.grid = 148;
  const int work = P.tiles * co_splits;
  if (work < P.grid) P.grid = work;
  P.grid -= P.grid % co_splits;
  P.debug = env_int("SV_NS_DEBUG", 0);
  return true;
}

}  // namespace

const char* tc_last_error() { return g_tc_error; }

size_t tc_first_stage_bytes(int B, int H, int W) { return (size_t)B * H * (W + 8) * 16; }

void tc_stage_first(const float* inputs, void* xp, int coff, int B, int H, int W, bool split, cudaStream_t s) {
  const long long total = (long long)B * H * W;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(stage_first_kernel, dim3((int)blocks), dim3(256), 0, s, inputs, (bf16*)xp, coff, B, H, W, split ? 1 : 0);
}

// First conv of an encoder (6x6, stride 2, 3 input channels): the staged image is read as overlapping
// 8-pixel x 8-channel windows, one K block of 64 per kernel row (48 of the 64 K entries carry weights).
static void plan_first_layer(TcLayer& t, const ConvGeom& g, int out_dt, size_t& off, bool split_fwd) {
  if (!(g.Ci == 3 && g.kh == 6 && g.kw == 6 && g.stride == 2 && g.pl == 2 && is_pow2(g.Ho) && is_pow2(g.Wo) && g.Wo <= 64 &&
        g.Wi == 2 * g.Wo && g.Hi == 2 * g.Ho && (g.dout_ld % 8) == 0))
    return;
  t.first = true;
  {
    TcLaunch& L = t.fwd;
    if (!tile_grid(L, g.Ho, g.Wo, g.B)) return;
    L.taps_h = 6; L.taps_w = 1; L.pad_t = g.pt; L.pad_l = 0; L.a_stride = 2;
    L.bk = 64; L.swizzle = 128; L.kc = 1;
    L.split = split_fwd ? 2 : 0;          // the staged pixel carries [hi lo] itself: two weight k-blocks per tap, one A chunk
    t.ci_pad = 64;
    t.n_pad_fwd = pad_cols(g.Co);
    L.n_valid = g.Co;
    L.OH = g.Ho; L.OW = g.Wo; L.osy = 1; L.ooy = 0; L.osx = 1; L.oox = 0;
    L.out_ld = g.out_ld; L.out_f32 = out_dt == DT_F32;
    L.nparts = g.nparts;
    for (int j = 0; j < 3; ++j) { L.part_n[j] = g.part_n[j]; L.part_act[j] = g.part_act[j]; }
    L.mask_act = ACT_NONE;
    finish_launch(L, t.n_pad_fwd);
    t.fwd_ok = true;
    t.fwd_launches = 1;
    t.w_fwd_off = off;
    off += round_up(t.n_pad_fwd * 6 * 64 * 2 * (split_fwd ? 2 : 1), 1024);
    t.bias_off = off;
    off += round_up(t.n_pad_fwd * 4, 1024);
    // Pixel-pair view on the halo kernel: the staged row [W+8][8 ch] is also [(W+8)/2][16 ch], so horizontally the 6x6 stride-2
    // convolution is a 3-tap stride-1 convolution over pairs (K = 16 per pair-tap: one MMA) and vertically the UMMA row-group
    // stride skips every other input row.  The tile's input halo is read once (the window kernel re-reads it per filter row).
    if ((g.Wi % 2) == 0 && env_int("SV_FIRST_PAIR", 1)) {
      TcLaunch P = L;
      P.taps_h = 6; P.taps_w = 3; P.pad_l = 0; P.bk = 16; P.swizzle = 32; P.kc = 1;
      finish_launch(P, t.n_pad_fwd);
      try_halo(P, g.Ho, g.Wo, g.B, 1, 2, true);
      if (P.halo) plan_persist(P, 1);
      if (P.halo && !P.persist) {       // wide first layers (GM h_block.0: 128 columns): a tile whose two accumulator sets fit in TMEM
        TcLaunch Q = L;
        Q.taps_h = 6; Q.taps_w = 3; Q.pad_l = 0; Q.bk = 16; Q.swizzle = 32; Q.kc = 1;
        finish_launch(Q, t.n_pad_fwd);
        try_halo(Q, g.Ho, g.Wo, g.B, 1, 2, true, 0, 256);
        if (Q.halo) { plan_persist(Q, 1); if (Q.persist) P = Q; }
      }
      if (P.halo) { L = P; t.first_pair = true; t.ci_pad = 16; }
    }
  }
  {
    TcWgradLaunch& L = t.wg;
    int cbn, copad;
    choose_bk(g.Co, cbn, copad);
    if (copad > g.dout_ld) return;
    L.first = 1;
    L.taps_h = 6; L.taps_w = 1; L.pad_t = g.pt; L.pad_l = 0; L.a_stride = 2;
    L.cb = 64; L.ncb = 1; L.a_swizzle = 128;
    L.cbn = cbn; L.b_swizzle = cbn * 2;
    L.nsub = 2; L.total_sb = 6; L.groups = 3;
    L.n_pad = copad;
    L.tile_cols = copad < 256 ? copad : 256;
    L.n_tiles = copad / L.tile_cols;
    L.groups_per_cta = 512 / L.tile_cols < 3 ? 512 / L.tile_cols : 3;
    L.m_pad = L.groups * 128;
    L.tile_w = g.Wo; L.tile_h = g.Ho < 64 / g.Wo ? g.Ho : 64 / g.Wo; L.tile_n_img = 64 / (L.tile_w * L.tile_h);
    L.grid_h = g.Ho; L.n_img = g.B;
    L.nchunks = L.tile_n_img > 1 ? (g.B + L.tile_n_img - 1) / L.tile_n_img : g.B * (g.Ho / L.tile_h);
    const int m_splits = (L.groups + L.groups_per_cta - 1) / L.groups_per_cta;
    int ks = (296 + m_splits * L.n_tiles - 1) / (m_splits * L.n_tiles);
    if (ks > L.nchunks) ks = L.nchunks;
    if (ks < 1) ks = 1;
    L.chunks_per_split = (L.nchunks + ks - 1) / ks;
    L.k_splits = (L.nchunks + L.chunks_per_split - 1) / L.chunks_per_split;
    L.a_stages = 4; L.b_stages = 2;
    L.smem_bytes = (size_t)L.a_stages * 128 * 64 * 2 + (size_t)L.b_stages * L.tile_cols * 64 * 2 + sizeof(WgCtl) + 1024;
    t.wgrad_ok = true;
    t.wgrad_launches = 2;
    size_t partial_bytes = (size_t)L.k_splits * L.m_pad * L.n_pad * 4;
    // Halo kernel through the pixel-pair view of the staged image (natural pairs: pixel x sits in column x + 2 of the [W + 8][8] row,
    // so the map starts two pixels in and the zero borders become TMA out-of-bounds fill): M rows = (filter row, parity, 8-channel
    // pixel slot), one 96-column MMA per 16 pixels instead of 6 x 2 per-tap box loads of the whole input.
    {
      ConvGeom g1 = g;
      g1.Ci = 8; g1.in_ld = 8; g1.in_coff = 0;
      int cb1 = 0, cipad1 = 0;
      if (env_int("SV_FIRST_PAIR_WGRAD", 1) && plan_halo_wgrad(t.hw, g1, cb1, cipad1, cbn, copad)) {
        t.wg_halo = true;
        const size_t hb = (size_t)t.hw.k_splits * t.hw.m_pad * t.hw.n_pad * 4;
        if (hb > partial_bytes) partial_bytes = hb;
      }
    }
    t.wg_partial_off = off;
    off += (partial_bytes + 1023) / 1024 * 1024;
  }
}

void tc_plan_layer(TcLayer& t, const ConvGeom& g, int in_dt, int out_dt, bool has_internal_input, bool has_dgrad, bool first_layer,
                   bool split_fwd) {
  t.in_dt = in_dt;
  t.out_dt = out_dt;
  t.split_fwd = split_fwd;
  size_t off = 0;
  if (first_layer) {
    plan_first_layer(t, g, out_dt, off, split_fwd);
    t.bytes = off;
    return;
  }
  // ---- forward: A = layer input (must be an internal bf16 tensor) ----
  if (has_internal_input && in_dt == DT_BF16 && (g.in_ld % 8) == 0 && (g.in_coff % 8) == 0) {
    TcLaunch& L = t.fwd;
    int bk, cpad;
    choose_bk(g.Ci, bk, cpad);
    if (cpad <= g.in_ld - g.in_coff || cpad == round_up(g.Ci, 8)) {
      if (tile_grid(L, g.Ho, g.Wo, g.B)) {
        L.taps_h = g.kh; L.taps_w = g.kw; L.pad_t = g.pt; L.pad_l = g.pl; L.a_stride = g.stride;
        L.bk = bk; L.swizzle = bk * 2; L.kc = cpad / bk;
        L.split = split_fwd ? 1 : 0;
        t.ci_pad = cpad;
        t.n_pad_fwd = pad_cols(g.Co);
        L.n_valid = g.Co;
        L.OH = g.Ho; L.OW = g.Wo; L.osy = 1; L.ooy = 0; L.osx = 1; L.oox = 0;
        L.out_ld = g.out_ld; L.out_f32 = out_dt == DT_F32;
        L.nparts = g.nparts;
        for (int j = 0; j < 3; ++j) { L.part_n[j] = g.part_n[j]; L.part_act[j] = g.part_act[j]; }
        L.mask_act = ACT_NONE;
        finish_launch(L, t.n_pad_fwd);
        if (cpad <= g.in_ld - g.in_coff) {
          // (stride-2 forward on the halo kernel: parity-tested, but the weights of e2 - 144 KB - cannot stay resident beside the
          //  halo, and streaming them through the 3-stage ring is latency-bound: 39 us vs 32 us per-tap -> off by default)
          // (bf16x3: twice the operand bytes per tap make the per-tap kernel L2-bound everywhere -> halo-resident wherever the grid allows)
          if (g.stride == 1) try_halo(L, g.Ho, g.Wo, g.B, 1, 1, split_fwd);
          else if (g.stride == 2 && g.Hi == 2 * g.Ho && g.Wi == 2 * g.Wo && env_int("SV_S2_FWD_HALO", 0)) try_halo(L, g.Ho, g.Wo, g.B, 2, 2, true);
          plan_persist(L, 1);
          // Stride-2 forward on the persistent kernel: split the output channels over up to 4 CTA classes until the class's
          // weights stay resident beside two halo stages (e2: 2 x 32 columns, 72 KB of weights, 8x16 tiles); the halo is then
          // read once per class instead of once per filter tap (per-tap kernel: 226 MB L2->SMEM, 31 us).
          if (g.stride == 2 && !L.persist && g.Hi == 2 * g.Ho && g.Wi == 2 * g.Wo && g.nparts == 1 && env_int("SV_S2_FWD_PCONV", 1)) {
            // Pixel-pair view (as for the first layer): with an even filter width and an even left pad the NHWC row [W][C] read
            // as [W/2][2C] turns the stride-2 convolution into kw/2 unit-stride pair taps of K = 2C, so the halo is ONE
            // contiguous TMA box (the parity-plane form needs element-strided boxes, which the TMA unit gathers at ~11 cycles per
            // 64-byte pixel: e2 forward stayed at 29 us).  Needs a dense input tensor and 2C <= 64 channels per swizzle row.
            const bool pair_ok = (g.kw % 2) == 0 && (g.pl % 2) == 0 && g.in_coff == 0 && g.in_ld == cpad && L.kc == 1 && 2 * cpad <= 64 &&
                                 (g.Wi % 2) == 0 && env_int("SV_S2_FWD_PAIR", 1);
            for (int split = 1; split <= 4; split *= 2) {
              const int cols = t.n_pad_fwd / split;
              if (cols < 16 || (cols % 16) || cols > 128) continue;
              TcLaunch Q = L;
              Q.tile_cols = cols; Q.n_tiles = split;
              if (pair_ok) { Q.taps_w = g.kw / 2; Q.pad_l = g.pl / 2; Q.bk = 2 * L.bk; Q.swizzle = 2 * L.swizzle; }
              const size_t w_bytes = ((size_t)g.kh * g.kw * L.kc * cols * L.bk * 2 * (split_fwd ? 2 : 1) + 1023) / 1024 * 1024;
              const size_t budget = 225 * 1024 - 1024 - sizeof(PcCtl);
              if (w_bytes + 4096 >= budget) continue;
              try_halo(Q, g.Ho, g.Wo, g.B, pair_ok ? 1 : 2, 2, true, (budget - w_bytes) / 2);
              if (!Q.halo) continue;
              plan_persist(Q, split);
              if (Q.persist) { L = Q; t.fwd_pair = pair_ok; if (pair_ok) t.ci_pad = 2 * cpad; break; }
            }
            // bf16x3: the pair's weights (2 sections) never stay resident beside the halo pairs; the one-tile-per-CTA halo kernel streams
            // them instead, still reading the input once per tile (the per-tap kernel re-reads both planes per tap: e2 forward 86 us)
            if (split_fwd && !L.persist && pair_ok && t.n_pad_fwd <= 128) {
              TcLaunch Q = L;
              Q.taps_w = g.kw / 2; Q.pad_l = g.pl / 2; Q.bk = 2 * L.bk; Q.swizzle = 2 * L.swizzle;
              finish_launch(Q, t.n_pad_fwd);
              try_halo(Q, g.Ho, g.Wo, g.B, 1, 2, true);
              if (Q.halo) { L = Q; t.fwd_pair = true; t.ci_pad = 2 * cpad; }
            }
          }
        }
        // (bf16x3: the pair mode of the N-stacked kernel, W <= 32 only - d3 / d4; the 64-wide cross-warp epilogue is not paired)
        if ((!split_fwd || g.Wo <= 32) && g.stride == 1 && g.nparts == 1 && g.part_act[0] != ACT_SOFTPLUS && cpad <= g.in_ld - g.in_coff &&
            g.Ho == g.Hi && g.Wo == g.Wi && plan_nsconv(t.ns_fwd, g.kh, g.kw, g.pt, g.pl, g.Ho, g.Wo, g.B, g.Ci, g.Co, split_fwd)) {
          TcNsConv& P = t.ns_fwd;
          P.n_valid = g.Co; P.out_ld = g.out_ld; P.out_f32 = out_dt == DT_F32; P.act = g.part_act[0]; P.mask_act = ACT_NONE;
          t.fwd_ns = true;
          t.w_nsf_off = off;
          off += (size_t)P.w_bytes * P.co_splits;
        }
        plan_split_k(L, g.B, off);
        t.sk_fwd_off = off;
        off += split_k_bytes(L);
        t.fwd_ok = true;
        t.fwd_launches = L.k_splits > 1 ? 2 : 1;
        t.w_fwd_off = off;
        off += round_up((int)((size_t)t.n_pad_fwd * g.kh * g.kw * cpad * 2 * (split_fwd ? 2 : 1)), 1024);
        t.bias_off = off;
        off += round_up(t.n_pad_fwd * 4, 1024);
      }
    }
  }
  // ---- dgrad: A = dY (bf16, pitch dout_ld), output = dX ----
  if (has_dgrad && (g.dout_ld % 8) == 0 && (g.stride == 1 || g.stride == 2)) {
    int bk, copad;
    choose_bk(g.Co, bk, copad);
    const int s = g.stride;
    const int GH = g.Hi / s, GW = g.Wi / s;
    bool ok = copad <= g.dout_ld && (g.kh % s) == 0 && (g.kw % s) == 0 && (g.Hi % s) == 0 && (g.Wi % s) == 0;
    t.n_dgrad = s * s;
    t.dg_taps_h = g.kh / s; t.dg_taps_w = g.kw / s;
    t.co_pad = copad;
    t.n_pad_dg = pad_cols(g.Ci);
    for (int cls = 0; ok && cls < s * s; ++cls) {
      TcLaunch& L = t.dgrad[cls];
      const int ph = cls / s, pw = cls % s;
      if (!tile_grid(L, GH, GW, g.B)) { ok = false; break; }
      const int rh = (ph + g.pt) % s, rw = (pw + g.pl) % s;
      const int qh = (ph + g.pt - rh) / s, qw = (pw + g.pl - rw) / s;
      L.taps_h = t.dg_taps_h; L.taps_w = t.dg_taps_w;
      L.pad_t = t.dg_taps_h - 1 - qh; L.pad_l = t.dg_taps_w - 1 - qw;
      L.a_stride = 1;
      L.bk = bk; L.swizzle = bk * 2; L.kc = copad / bk;
      L.n_valid = g.Ci;
      L.OH = g.Hi; L.OW = g.Wi; L.osy = s; L.ooy = ph; L.osx = s; L.oox = pw;
      L.out_ld = g.din_ld; L.out_f32 = 0;
      L.nparts = 1; L.part_n[0] = t.n_pad_dg; L.part_act[0] = ACT_NONE;
      finish_launch(L, t.n_pad_dg, s * s);
      try_halo(L, GH, GW, g.B, 1, 1, s == 2 && L.tile_cols <= 64 && env_int("SV_S2_DGRAD_HALO", 1));
      plan_persist(L, s * s);
      if (s == 1) {
        plan_split_k(L, g.B, off);
        t.sk_dgrad_off = off;
        off += split_k_bytes(L);
      }
    }
    // (W = 64 dgrads - d5 - stay on the halo kernel: 4 shuffle blocks per tile with the cross-warp exchange make the
    // N-stacked epilogue the bottleneck there, 124 us vs 102 us)
    if (ok && s == 1 && g.Ho == g.Hi && g.Wo == g.Wi && (g.Wi <= 32 || env_int("SV_NS_WIDE_DGRAD", 0)) &&
        plan_nsconv(t.ns_dgrad, g.kh, g.kw, g.kh - 1 - g.pt, g.kw - 1 - g.pl, g.Hi, g.Wi, g.B, g.Co, g.Ci)) {
      TcNsConv& P = t.ns_dgrad;
      P.n_valid = g.Ci; P.out_ld = g.din_ld; P.out_f32 = 0; P.act = ACT_NONE;
      if (P.nchunks * P.ck <= g.dout_ld) {
        t.dgrad_ns = true;
        off = (off + 1023) / 1024 * 1024;
        t.w_nsd_off = off;
        off += (size_t)P.w_bytes * P.co_splits;
      }
    }
    if (ok && s == 2 && !env_int("SV_NO_DGRAD_MERGE", 0)) {
      bool same = true;
      for (int cls = 0; cls < 4; ++cls) {
        const TcLaunch &A = t.dgrad[0], &C = t.dgrad[cls];
        same = same && C.halo == A.halo && C.k_splits <= 1 && C.smem_bytes == A.smem_bytes && C.n_tiles == A.n_tiles && C.tile_h == A.tile_h &&
               C.tile_n_img == A.tile_n_img && C.grid_h == A.grid_h && C.n_img == A.n_img &&
               (!C.halo || (C.TW == A.TW && C.TH == A.TH && C.tiles_x == A.tiles_x && C.tiles_y == A.tiles_y)) &&
               C.persist == A.persist && (!C.persist || (C.p_smem == A.p_smem && C.p_grid == A.p_grid));
      }
      t.dgrad_merged = same;
    }
    if (ok) {
      t.dgrad_ok = true;
      t.dgrad_launches = t.dgrad_ns || t.dgrad_merged ? 1 : s * s + (s == 1 && t.dgrad[0].k_splits > 1 ? 1 : 0);
      t.w_dgrad_off = off;
      off += (size_t)s * s * round_up((int)((size_t)t.n_pad_dg * t.dg_taps_h * t.dg_taps_w * copad * 2), 1024);
    }
  }
  // ---- wgrad: A = layer input, B = dY, K axis = pixels ----
  if (has_internal_input && in_dt == DT_BF16 && (g.in_ld % 8) == 0 && (g.in_coff % 8) == 0 && (g.dout_ld % 8) == 0 &&
      is_pow2(g.Ho) && is_pow2(g.Wo) && g.Wo <= 64) {
    TcWgradLaunch& L = t.wg;
    int cb, cipad, cbn, copad;
    choose_bk(g.Ci, cb, cipad);
    choose_bk(g.Co, cbn, copad);
    const bool a_fits = cipad <= g.in_ld - g.in_coff, b_fits = copad <= g.dout_ld;
    if (a_fits && b_fits) {
      L.taps_h = g.kh; L.taps_w = g.kw; L.pad_t = g.pt; L.pad_l = g.pl; L.a_stride = g.stride;
      L.cb = cb; L.ncb = cipad / cb; L.a_swizzle = cb * 2;
      L.cbn = cbn; L.b_swizzle = cbn * 2;
      L.nsub = 128 / cb;
      L.total_sb = g.kh * g.kw * L.ncb;
      L.groups = (L.total_sb + L.nsub - 1) / L.nsub;
      L.n_pad = copad;
      L.tile_cols = copad < 256 ? copad : 256;
      if (copad % L.tile_cols) L.tile_cols = 128;
      L.n_tiles = copad / L.tile_cols;
      // Parallelism comes from the M axis first (few row groups per CTA) and only then from split-K: every K split costs a
      // full fp32 copy of the weight gradient in the partial buffer, written once and re-read by the reduce kernel (64-way
      // splits made that ~1.2 GB of traffic per step).
      L.groups_per_cta = 512 / L.tile_cols;
      if (L.groups_per_cta > L.groups) L.groups_per_cta = L.groups;
      if (L.groups_per_cta > 8) L.groups_per_cta = 8;
      {
        const int cap = env_int("SV_WG_GPC", 1);
        if (cap >= 1 && L.groups_per_cta > cap) L.groups_per_cta = cap;
      }
      L.m_pad = L.groups * 128;
      L.tile_w = g.Wo; L.tile_h = g.Ho < 64 / g.Wo ? g.Ho : 64 / g.Wo; L.tile_n_img = 64 / (L.tile_w * L.tile_h);
      L.grid_h = g.Ho; L.n_img = g.B;
      L.nchunks = L.tile_n_img > 1 ? (g.B + L.tile_n_img - 1) / L.tile_n_img : g.B * (g.Ho / L.tile_h);
      const int m_splits = (L.groups + L.groups_per_cta - 1) / L.groups_per_cta;
      // measured (B200, C2): convolutions like ~2 CTAs per SM (e2 48 vs 73 us), the dense layers ~1 (38 vs 48 us: their K axis
      // is only the batch, so extra splits just add partial-buffer traffic)
      const int target_ctas = env_int("SV_WG_CTAS", g.kh * g.kw > 1 ? 296 : 148);
      int ks = (target_ctas + m_splits * L.n_tiles - 1) / (m_splits * L.n_tiles);
      if (ks > L.nchunks) ks = L.nchunks;
      if (ks < 1) ks = 1;
      L.chunks_per_split = (L.nchunks + ks - 1) / ks;
      L.k_splits = (L.nchunks + L.chunks_per_split - 1) / L.chunks_per_split;
      L.a_stages = 4; L.b_stages = 2;
      L.smem_bytes = (size_t)L.a_stages * 128 * 64 * 2 + (size_t)L.b_stages * L.tile_cols * 64 * 2 + sizeof(WgCtl) + 1024;
      t.wgrad_ok = true;
      t.wgrad_launches = 2;
      size_t partial_bytes = (size_t)L.k_splits * L.m_pad * L.n_pad * 4;
      if (plan_halo_wgrad(t.hw, g, cb, cipad, cbn, copad)) {
        t.wg_halo = true;
        const size_t hb = (size_t)t.hw.k_splits * t.hw.m_pad * t.hw.n_pad * 4;
        if (hb > partial_bytes) partial_bytes = hb;
      }
      t.wg_partial_off = off;
      off += (partial_bytes + 1023) / 1024 * 1024;
    }
  }
  t.bytes = off;
}

size_t tc_workspace_bytes(const TcLayer& t, const ConvGeom&) { return t.bytes; }

const char* tc_bind_layer(TcLayer& t, const ConvGeom& g, const void* in, void* out, void* dout, void* din, const void* mask_src,
                          int mask_act, char* ws, const void* in_lo, void* out_lo) {
  t.ws = ws;
  if (t.fwd_ok) {
    TcLaunch& L = t.fwd;
    auto make_a = [&](CUtensorMap* m, const void* base) -> const char* {
      // (first_pair: the staged image [B][H][W+8][8] seen as [B][H][(W+8)/2][16] - one K = 16 row per pixel pair)
      return t.first_pair ? make_halo_map(m, base, g.B, g.Hi, (g.Wi + 8) / 2, 16, 0, 16, 16, L.TWp, L.THp, 1, 32)
             : t.fwd_pair ? make_halo_map(m, base, g.B, g.Hi, g.Wi / 2, 2 * g.in_ld, 0, t.ci_pad, L.bk, L.TWp, L.THp, 1, L.swizzle)
             : t.first    ? make_window_map(m, base, g.B, g.Hi, g.Wi, g.Wo, L.tile_w, L.tile_h, L.tile_n_img)
             : L.halo     ? make_halo_map(m, base, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, t.ci_pad, L.bk, L.TWp, L.THp, L.halo_sx, L.swizzle)
                          : make_act_map(m, base, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, t.ci_pad <= g.in_ld - g.in_coff ? t.ci_pad : g.Ci, L.bk,
                                         L.tile_w, L.tile_h, L.tile_n_img, g.stride, L.swizzle);
    };
    const char* e = make_a(&L.map_a, in);
    if (e) return e;
    if (L.kca > L.kc) {          // bf16x3: the lo plane of the input, same geometry
      if (!in_lo) return "bf16x3 forward needs the lo plane of the layer input";
      e = make_a(&L.map_a_lo, in_lo);
      if (e) return e;
    }
    e = L.w_box3 ? make_w_map3(&L.map_b, ws + t.w_fwd_off, t.n_pad_fwd, (long long)L.taps_h * L.taps_w * L.kcb * L.bk, L.bk, L.tile_cols,
                               L.kb_per_stage, L.swizzle)
                 : make_w_map(&L.map_b, ws + t.w_fwd_off, t.n_pad_fwd, (long long)L.taps_h * L.taps_w * L.kcb * L.bk, L.bk, L.tile_cols, L.swizzle);
    if (e) return e;
    L.bias = (const float*)(ws + t.bias_off);
    L.out = out;
    L.out_lo = (t.split_fwd && !L.out_f32) ? out_lo : nullptr;
    L.mask_src = nullptr;
    L.partial = L.k_splits > 1 ? (float*)(ws + t.sk_fwd_off) : nullptr;
    if (cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
    if (cudaFuncSetAttribute(igemm_fat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
  }
  if (t.fwd_ns || t.dgrad_ns) {
    if (cudaFuncSetAttribute(nsconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
    for (int which = 0; which < 2; ++which) {
      if (!(which ? t.dgrad_ns : t.fwd_ns)) continue;
      TcNsConv& P = which ? t.ns_dgrad : t.ns_fwd;
      const char* e = which ? make_act_map(&P.map_x, dout, g.B, g.Ho, g.Wo, g.dout_ld, 0, P.nchunks * P.ck, P.ck, P.W, P.halo_rows, 1, 1, P.pixB)
                            : make_act_map(&P.map_x, in, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, P.nchunks * P.ck, P.ck, P.W, P.halo_rows, 1, 1, P.pixB);
      if (e) return e;
      if (P.pair) {
        if (which || !in_lo) return "bf16x3 N-stacked forward needs the lo plane of the layer input";
        e = make_act_map(&P.map_x_lo, in_lo, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, P.nchunks * P.ck, P.ck, P.W, P.halo_rows, 1, 1, P.pixB);
        if (e) return e;
      }
      e = make_w_map(&P.map_w, ws + (which ? t.w_nsd_off : t.w_nsf_off), P.co_splits * (P.pair ? 2 : 1) * P.n_total, (long long)P.kh * P.nchunks * P.ck,
                     P.ck, P.n_box, P.pixB);
      if (e) return e;
      P.out_lo = (!which && P.pair) ? out_lo : nullptr;
      P.bias = which ? nullptr : (const float*)(ws + t.bias_off);
      P.out = which ? din : out;
      P.mask_src = which ? mask_src : nullptr;
      P.mask_act = which ? mask_act : ACT_NONE;
      P.mask_ld = g.in_ld; P.mask_coff = g.in_coff;
      if (env_int("SV_TC_VERBOSE", 0))
        fprintf(stderr, "[tc] %dx%d s%d Ci%d Co%d %s: NSCONV block %dx%d halo rows %d chunks %d x %d B, N %d = %d x %d (ng %d, groups %d x lanes %d, co_splits %d), weights %d B resident, %d halo stages, smem %zu, grid %d\n",
                g.kh, g.kw, g.stride, g.Ci, g.Co, which ? "dgrad" : "fwd", P.R, P.W, P.halo_rows, P.nchunks, P.chunk_bytes, P.n_total, P.kw, P.nb, P.ng,
                P.groups, P.lanes, P.co_splits, P.w_bytes, P.nstages, P.smem_bytes, P.grid);
    }
  }
  if (t.dgrad_ok) {
    const size_t per_class = round_up((int)((size_t)t.n_pad_dg * t.dg_taps_h * t.dg_taps_w * t.co_pad * 2), 1024);
    for (int cls = 0; cls < t.n_dgrad; ++cls) {
      TcLaunch& L = t.dgrad[cls];
      const char* e = L.halo ? make_halo_map(&L.map_a, dout, g.B, g.Ho, g.Wo, g.dout_ld, 0, t.co_pad, L.bk, L.TWp, L.THp, 1, L.swizzle)
                             : make_act_map(&L.map_a, dout, g.B, g.Ho, g.Wo, g.dout_ld, 0, t.co_pad, L.bk, L.tile_w, L.tile_h, L.tile_n_img, 1,
                                            L.swizzle);
      if (e) return e;
      e = make_w_map(&L.map_b, ws + t.w_dgrad_off + cls * per_class, t.n_pad_dg, (long long)t.dg_taps_h * t.dg_taps_w * t.co_pad, L.bk,
                     L.tile_cols, L.swizzle);
      if (e) return e;
      L.bias = nullptr;
      L.out = din;
      L.mask_src = mask_src;
      L.mask_act = mask_act;
      L.mask_ld = g.in_ld;
      L.mask_coff = g.in_coff;
      L.partial = L.k_splits > 1 ? (float*)(ws + t.sk_dgrad_off) : nullptr;
    }
    if (cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
    if (cudaFuncSetAttribute(igemm_fat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
    if (cudaFuncSetAttribute(igemm4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
  }
  if (cudaFuncSetAttribute(halo_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
  if (cudaFuncSetAttribute(halo4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
  if (cudaFuncSetAttribute(pconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
  if (env_int("SV_TC_VERBOSE", 0)) {
    auto show = [&](const char* what, const TcLaunch& L) {
      if (L.halo && L.persist)
        fprintf(stderr, "[tc] %dx%d s%d Ci%d Co%d %s: PCONV tile %dx%d halo %dx%d x%d planes, chunks %d x %d B, N %d, weights %d B resident, %d halo stages, smem %zu, grid %d\n",
                g.kh, g.kw, g.stride, g.Ci, g.Co, what, L.TW, L.TH, L.TWp, L.THp, L.halo_sx, L.kc, L.chunk_bytes, L.tile_cols, L.p_wbytes, L.p_stages,
                L.p_smem, L.p_grid);
      else if (L.halo)
        fprintf(stderr, "[tc] %dx%d s%d Ci%d Co%d %s: HALO tile %dx%d halo %dx%d chunks %d x %d B, N %d x%d, w ring %d x %d B (%d kb/stage), smem %zu\n",
                g.kh, g.kw, g.stride, g.Ci, g.Co, what, L.TW, L.TH, L.TWp, L.THp, L.kc, L.chunk_bytes, L.tile_cols, L.n_tiles, L.w_stages,
                L.w_stage_bytes, L.kb_per_stage, L.smem_bytes);
      else
        fprintf(stderr, "[tc] %dx%d s%d Ci%d Co%d %s: per-tap tile %dx%dx%d bk %d N %d x%d stages %d smem %zu split-K %d x %d kb\n", g.kh, g.kw, g.stride, g.Ci, g.Co,
                what, L.tile_n_img, L.tile_h, L.tile_w, L.bk, L.tile_cols, L.n_tiles, L.stages, L.smem_bytes, L.k_splits, L.kb_per_split);
    };
    if (t.fwd_ok) show("fwd", t.fwd);
    if (t.dgrad_ok) for (int c = 0; c < t.n_dgrad; ++c) show("dgrad", t.dgrad[c]);
  }
  if (t.wgrad_ok) {
    TcWgradLaunch& L = t.wg;
    const char* e = t.first ? make_window_map(&L.map_a, in, g.B, g.Hi, g.Wi, g.Wo, L.tile_w, L.tile_h, L.tile_n_img)
                            : make_act_map(&L.map_a, in, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, L.ncb * L.cb, L.cb, L.tile_w, L.tile_h,
                                           L.tile_n_img, g.stride, L.a_swizzle);
    if (e) return e;
    e = make_act_map(&L.map_b, dout, g.B, g.Ho, g.Wo, g.dout_ld, 0, L.n_pad, L.cbn, L.tile_w, L.tile_h, L.tile_n_img, 1, L.b_swizzle);
    if (e) return e;
    L.partial = (float*)(ws + t.wg_partial_off);
    if (cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
    if (t.wg_halo) {
      TcHaloWgrad& H = t.hw;
      e = t.first  ? make_first_pair_map(&H.map_x, in, g.B, g.Hi, g.Wi, H.x_tw, H.x_th)
          : H.pair_c ? make_act_map(&H.map_x, in, g.B, g.Hi, g.Wi / 2, 2 * g.in_ld, 0, H.nchunks * H.cb, H.cb, H.x_tw, H.x_th, 1, 1, H.x_swizzle)
                   : make_act_map(&H.map_x, in, g.B, g.Hi, g.Wi, g.in_ld, g.in_coff, H.nchunks * H.cb, H.cb, H.x_tw, H.x_th, 1, 1, H.x_swizzle);
      if (e) return e;
      e = make_act_map(&H.map_dy, dout, g.B, g.Ho, g.Wo, g.dout_ld, 0, H.nbchunks * H.cbn, H.cbn, H.dy_tw, H.dy_th, 1, 1, H.dy_swizzle);
      if (e) return e;
      H.partial = (float*)(ws + t.wg_partial_off);
      if (cudaFuncSetAttribute(halo_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return "cudaFuncSetAttribute failed";
      if (env_int("SV_TC_VERBOSE", 0))
        fprintf(stderr, "[tc] %dx%d s%d Ci%d Co%d wgrad: HALO mode %d nstack %d pair %d tile %dx%d groups %d (%d/CTA, m_splits %d) N %d tiles %d k_splits %d smem %zu\n",
                g.kh, g.kw, g.stride, g.Ci, g.Co, H.mode, H.nstack, H.pair_c, H.TW, H.TH, H.groups, H.groups_per_cta, H.m_splits, H.n_pad, H.tiles, H.k_splits,
                H.smem_bytes);
    }
  }
  return nullptr;
}

struct TcPackTable {
  PackJob* dev = nullptr;
  uint16_t* block_job = nullptr;      // block index -> job
  int njobs = 0, nblocks = 0;
};
void tc_pack_table_destroy(TcPackTable* t);

TcPackTable* tc_pack_table_create(TcLayer* const* layers, const ConvGeom* const* geoms, int n, const char** err) {
  std::vector<PackJob> jobs;
  int blocks = 0;
  auto push = [&](PackJob J) {
    J.block_start = blocks;
    if (J.kind == 0 && J.fast) blocks += J.taps_h * J.taps_w * ((J.rows_pad + 31) / 32) * ((J.k_pad / J.nsec / 32 + kPackKT - 1) / kPackKT);
    else if (J.kind == 0) blocks += J.taps_h * J.taps_w * ((J.rows_pad + 31) / 32) * ((J.k_pad + 31) / 32);   // 32x32 transpose tiles
    else if (J.kind == 1) blocks += (int)((J.count + 2048 * kPackDgIter - 1) / (2048 * kPackDgIter));
    else blocks += (int)((J.count + 2047) / 2048);
    if (env_int("SV_PACK_DEBUG", 0))
      fprintf(stderr, "[pack] job %zu kind %d fast %d %dx%d Ci %d Co %d rows_pad %d taps %dx%d k_pad %d nsec %d ppx %d count %lld blocks %d\n", jobs.size(), J.kind,
              J.fast, J.KH, J.KW, J.Ci, J.Co, J.rows_pad, J.taps_h, J.taps_w, J.k_pad, J.nsec, J.ppx, J.count, blocks - J.block_start);
    jobs.push_back(J);
  };
  for (int i = 0; i < n; ++i) {
    const TcLayer& t = *layers[i];
    const ConvGeom& g = *geoms[i];
    PackJob B{};
    B.KH = g.kh; B.KW = g.kw; B.Ci = g.Ci; B.Co = g.Co; B.nparts = g.nparts;
    for (int j = 0; j < 3; ++j) { B.part_n[j] = g.part_n[j]; B.part_w[j] = g.part_w[j]; B.part_b[j] = g.part_b[j]; }
    if (t.fwd_ns) {
      PackJob J = B;
      J.kind = 4; J.rows_pad = t.ns_fwd.nb; J.taps_h = g.kh; J.taps_w = g.kw; J.k_pad = t.ns_fwd.nchunks * t.ns_fwd.ck;
      J.nsec = t.ns_fwd.pair ? 2 : 1;               // pair: rows [W_hi-stack ; W_lo-stack] per output-channel split
      J.dst = t.ws + t.w_nsf_off;
      J.count = (long long)t.ns_fwd.co_splits * J.nsec * t.ns_fwd.n_total * g.kh * J.k_pad;
      push(J);
    }
    if (t.dgrad_ns) {
      PackJob J = B;
      J.kind = 5; J.rows_pad = t.ns_dgrad.nb; J.taps_h = g.kh; J.taps_w = g.kw; J.k_pad = t.ns_dgrad.nchunks * t.ns_dgrad.ck;
      J.dst = t.ws + t.w_nsd_off;
      J.count = (long long)t.ns_dgrad.co_splits * t.ns_dgrad.n_total * g.kh * J.k_pad;
      push(J);
    }
    if (t.fwd_ok) {
      PackJob J = B;
      const TcLaunch& L = t.fwd;
      J.kind = t.fwd_ns ? -1 : 0;                   // (the N-stacked kernel serves this layer's forward: the plain copy would never be read)
      J.rows_pad = t.n_pad_fwd; J.taps_h = L.taps_h; J.taps_w = L.taps_w;
      // pixels per k-block: 8 for the first layer's window view (one k-block per filter row), 2 for the pixel-pair views
      J.ppx = (t.first && !t.first_pair) ? 8 : (t.first_pair || t.fwd_pair) ? 2 : 1;
      J.cpp = L.bk / J.ppx;                           // channels per pixel slot of one k-block
      J.nsec = L.kcb / L.kc;                          // [W_hi | W_lo] sections per channel chunk (bf16x3) or one
      J.first_cat = L.split == 2;
      J.k_pad = L.kcb * L.bk;
      J.fast = (J.ppx == 1 && !J.first_cat && (J.cpp % 32) == 0 && (J.nsec == 1 || J.nsec == 2) && !env_int("SV_PACK_SLOW", 0)) ? 1 : 0;
      J.dst = t.ws + t.w_fwd_off;
      J.count = (long long)t.n_pad_fwd * J.taps_h * J.taps_w * J.k_pad;
      if (J.kind == 0) push(J);
      PackJob Jb = B;
      Jb.kind = 2; Jb.dst = t.ws + t.bias_off; Jb.count = t.n_pad_fwd;
      push(Jb);
    }
    if (t.dgrad_ok && !t.dgrad_ns) {              // (dgrad_ns: the N-stacked copy above is the one tc_conv_dgrad uses)
      const int s = g.stride;
      const size_t per_class = round_up((int)((size_t)t.n_pad_dg * t.dg_taps_h * t.dg_taps_w * t.co_pad * 2), 1024);
      for (int cls = 0; cls < t.n_dgrad; ++cls) {
        PackJob J = B;
        J.kind = 1; J.rows_pad = t.n_pad_dg; J.taps_h = t.dg_taps_h; J.taps_w = t.dg_taps_w; J.k_pad = t.co_pad;
        J.stride = s; J.rh = (cls / s + g.pt) % s; J.rw = (cls % s + g.pl) % s;
        J.dst = t.ws + t.w_dgrad_off + cls * per_class;
        J.count = (long long)t.n_pad_dg * t.dg_taps_h * t.dg_taps_w * t.co_pad;
        push(J);
      }
    }
  }
  TcPackTable* T = new TcPackTable();
  T->njobs = (int)jobs.size();
  T->nblocks = blocks;
  if (T->njobs && env_int("SV_PACK_DEBUG", 0) != 2) {       // (2: dry run on a plan-only handle, job list on stderr only)
    std::vector<uint16_t> bj((size_t)blocks);
    for (size_t j = 0; j < jobs.size(); ++j) {
      const int end = j + 1 < jobs.size() ? jobs[j + 1].block_start : blocks;
      for (int b = jobs[j].block_start; b < end; ++b) bj[b] = (uint16_t)j;
    }
    if (cudaMalloc(&T->dev, jobs.size() * sizeof(PackJob)) != cudaSuccess ||
        cudaMemcpy(T->dev, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMalloc(&T->block_job, bj.size() * sizeof(uint16_t)) != cudaSuccess ||
        cudaMemcpy(T->block_job, bj.data(), bj.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess) {
      *err = "pack table upload failed";
      tc_pack_table_destroy(T);
      return nullptr;
    }
  }
  return T;
}

void tc_pack_table_destroy(TcPackTable* t) {
  if (!t) return;
  if (t->dev) cudaFree(t->dev);
  if (t->block_job) cudaFree(t->block_job);
  delete t;
}

int tc_repack_all(TcPackTable* t, const float* params, cudaStream_t s) {
  if (!t || !t->njobs) return 0;
  launch_pdl(pack_kernel, dim3(t->nblocks), dim3(256), 0, s, (const PackJob*)t->dev, (const uint16_t*)t->block_job, params);
  return 1;
}

static void launch(const TcLaunch& L, cudaStream_t s) {
  if (L.halo && L.persist) {
    TcLaunch4 P4;
    for (int j = 0; j < L.n_tiles; ++j) { P4.l[j] = L; P4.l[j].p_ntile = j; }   // (n_tiles > 1: one CTA class per column tile)
    launch_pdl(pconv_kernel, dim3(L.p_grid, L.n_tiles), dim3(kPcThreads), L.p_smem, s, P4);
    return;
  }
  if (L.halo) {
    dim3 grid(L.tiles_x * L.tiles_y * L.n_img, L.n_tiles);
    launch_pdl(halo_conv_kernel, dim3(grid), dim3(kThreads), L.smem_bytes, s, L);
    return;
  }
  const int tiles_per_img = L.grid_h / L.tile_h;
  const int m_tiles = L.tile_n_img > 1 ? (L.n_img + L.tile_n_img - 1) / L.tile_n_img : L.n_img * tiles_per_img;
  dim3 grid(m_tiles, L.n_tiles, L.k_splits > 1 ? L.k_splits : 1);
  if (L.fat) launch_pdl(igemm_fat_kernel, dim3(grid), dim3(kThreads), L.smem_bytes, s, L);
  else launch_pdl(igemm_kernel, dim3(grid), dim3(kThreads), L.smem_bytes, s, L);
  if (L.k_splits > 1) {
    const long long total = (long long)L.n_img * L.n_valid;
    launch_pdl(splitk_finish_kernel, dim3((int)((total + 255) / 256)), dim3(256), 0, s, L);
  }
}

static void launch_ns(const TcNsConv& P, cudaStream_t s) { launch_pdl(nsconv_kernel, dim3(P.grid), dim3(kNsThreads), P.smem_bytes, s, P); }

int tc_halo_trace_read(unsigned long long* out, int max_ctas) {
  const int n = max_ctas < kTraceCtas ? max_ctas : kTraceCtas;
  if (cudaMemcpyFromSymbol(out, g_halo_trace, (size_t)n * kTraceSlots * sizeof(unsigned long long)) != cudaSuccess) return -1;
  return n;
}

void tc_conv_fwd(TcLayer& t, cudaStream_t s) {
  if (t.fwd_ns) launch_ns(t.ns_fwd, s); else launch(t.fwd, s);
}
void tc_conv_dgrad(TcLayer& t, cudaStream_t s) {
  if (t.dgrad_ns) { launch_ns(t.ns_dgrad, s); return; }
  if (t.dgrad_merged) {
    TcLaunch4 P4;
    for (int c = 0; c < 4; ++c) P4.l[c] = t.dgrad[c];
    const TcLaunch& L = t.dgrad[0];
    const int tiles_per_img = L.grid_h / L.tile_h;
    const int m_tiles = L.tile_n_img > 1 ? (L.n_img + L.tile_n_img - 1) / L.tile_n_img : L.n_img * tiles_per_img;
    if (L.halo && L.persist) launch_pdl(pconv_kernel, dim3(L.p_grid, 4), dim3(kPcThreads), L.p_smem, s, P4);
    else if (L.halo) launch_pdl(halo4_kernel, dim3(L.tiles_x * L.tiles_y * L.n_img, L.n_tiles, 4), dim3(kThreads), L.smem_bytes, s, P4);
    else launch_pdl(igemm4_kernel, dim3(m_tiles, L.n_tiles, 4), dim3(kThreads), L.smem_bytes, s, P4);
    return;
  }
  for (int c = 0; c < t.n_dgrad; ++c) launch(t.dgrad[c], s);
}
static void launch_wgrad_reduce(const ConvGeom& g, const float* partial, int k_splits, int m_pad, int n_pad, const WgRowMap& R, float* grads,
                                cudaStream_t s) {
  if (k_splits <= 8) {   // few splits, many outputs (dense layers): one thread per output, all splits in flight at once
    const long long total = (long long)g.kh * g.kw * g.Ci * g.Co;
    bool vec = R.mode < 4 && (g.Co % 4) == 0 && (n_pad % 4) == 0;
    for (int j = 0; j < g.nparts; ++j) vec = vec && (g.part_n[j] % 4) == 0 && (g.part_w[j] % 4) == 0;
    long long blocks = ((vec ? total / 4 : total) + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (vec) launch_pdl(wgrad_reduce_few4_kernel, dim3((int)blocks), dim3(256), 0, s, g, partial, k_splits, m_pad, n_pad, R, grads);
    else launch_pdl(wgrad_reduce_few_kernel, dim3((int)blocks), dim3(256), 0, s, g, partial, k_splits, m_pad, n_pad, R, grads);
  } else if (!env_int("SV_OLD_REDUCE", 0)) {
    int vec = 4;
    while (vec > 1) {
      bool ok = (g.Co % vec) == 0 && (n_pad % vec) == 0 && (R.mode < 4 || (R.nb % vec) == 0);
      for (int j = 0; j < g.nparts; ++j) ok = ok && (g.part_n[j] % vec) == 0 && (g.part_w[j] % vec) == 0;
      if (ok) break;
      vec >>= 1;
    }
    const long long items = (long long)g.kh * g.kw * g.Ci * (g.Co / vec);
    const int blocks = (int)((items + 31) / 32);
    // split lanes: enough that every lane has <= ~8 loads, more when the grid alone would not fill the GPU
    // 256-thread blocks: the reduce runs on a low-priority stream beside the persistent chain kernels (nsconv: 576 threads x 96 registers per
    // SM), where a 1024-thread block finds no SM with enough free registers until that kernel has drained - in the step's timeline the
    // 9 us reduce of d3 took 50-80 us and held up every weight gradient queued behind it on its stream
    const int max_lanes = env_int("SV_REDUCE_LANES", 8);
    int lanes = 8;
    while (lanes < max_lanes && (k_splits > 8 * lanes || (long long)blocks * lanes < 148 * 16)) lanes <<= 1;
    const size_t smem = (size_t)lanes * 32 * vec * sizeof(float);
    if (vec == 4) launch_pdl(wgrad_reduce_vec_kernel<4>, dim3(blocks), dim3(32, lanes), smem, s, g, partial, k_splits, m_pad, n_pad, R, grads);
    else if (vec == 2) launch_pdl(wgrad_reduce_vec_kernel<2>, dim3(blocks), dim3(32, lanes), smem, s, g, partial, k_splits, m_pad, n_pad, R, grads);
    else launch_pdl(wgrad_reduce_vec_kernel<1>, dim3(blocks), dim3(32, lanes), smem, s, g, partial, k_splits, m_pad, n_pad, R, grads);
  } else {
    const int rblocks = g.kh * g.kw * g.Ci * ((g.Co + 31) / 32);
    launch_pdl(wgrad_reduce_kernel, dim3(rblocks), dim3(32, 8), 0, s, g, partial, k_splits, m_pad, n_pad, R, grads);
  }
}

void tc_conv_wgrad(TcLayer& t, const ConvGeom& g, float* grads, cudaStream_t s) {
  if (t.wg_halo) {
    const TcHaloWgrad& H = t.hw;
    dim3 grid(H.m_splits, 1, H.k_splits);
    launch_pdl(halo_wgrad_kernel, dim3(grid), dim3(kThreads), H.smem_bytes, s, H);
    const WgRowMap R{(H.nstack ? 4 : 2) + (H.mode ? 1 : 0), g.kw, 0, H.cb, H.nsub, H.gw, H.gpt, H.nb, H.pair_c, H.pair_sh, H.taps_w};
    launch_wgrad_reduce(g, H.partial, H.k_splits, H.m_pad, H.n_pad, R, grads, s);
    return;
  }
  const TcWgradLaunch& L = t.wg;
  const int m_splits = (L.groups + L.groups_per_cta - 1) / L.groups_per_cta;
  dim3 grid(m_splits, L.n_tiles, L.k_splits);
  launch_pdl(wgrad_kernel, dim3(grid), dim3(kThreads), L.smem_bytes, s, L);
  const WgRowMap R{L.first ? 1 : 0, g.kw, L.ncb * L.cb, 0, 0, 0, 0, 0};
  launch_wgrad_reduce(g, L.partial, L.k_splits, L.m_pad, L.n_pad, R, grads, s);
}

}  // namespace sv
