// Tensor-core (tcgen05 + TMA) implicit-GEMM path: per-layer plan and launchers (internal header).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sv {

// One launch of the implicit-GEMM kernel  D[pixels, N] = A[pixels, taps*C] * B[N, taps*C]^T.
//   A = an NHWC bf16 activation tensor read through a 4-D TMA map (C, W, H, N) with a per-tap shift,
//   B = a packed bf16 weight matrix [n_pad][taps*c_pad] read through a 2-D TMA map.
struct TcLaunch {
  CUtensorMap map_a, map_b;
  CUtensorMap map_a_lo;                         // bf16x3 forward: the lo plane of the A tensor (same geometry as map_a)
  int taps_h, taps_w, pad_t, pad_l, a_stride;   // A coordinate of (out y, tap a) = y*a_stride + a - pad_t
  int kc;                                       // channel chunks per tap (c_pad / bk)
  // bf16x3 forward (operands are bf16 pairs hi + lo, product = hi*hi + lo*hi + hi*lo): per tap the K loop runs over `kcl` LOGICAL
  // chunks that address `kca` physical A chunks (the hi plane's chunks, then the lo plane's) and `kcb` physical weight k-blocks
  // ([W_hi(ch) W_lo(ch)] per channel chunk); the mapping is logical_chunk() / phys_block_chunks() in tc_kernels.cu.
  //   split 0: plain, kcl = kca = kcb = kc;   split 1: kcl 3kc, kca 2kc, kcb 2kc;
  //   split 2 (first layer): the 8-channel staged pixel holds [hi(3) lo(3) 0 0]; k-block 0 = [Whi Whi 0 0], 1 = [Wlo 0 0 0]: kcl 2, kca 1, kcb 2
  int split, kcl, kca, kcb;
  // per-tap kernel, split 1 ("fat" ring stages): one stage = [A_hi(ch) | A_lo(ch) | W_hi(ch) | W_lo(ch)] of one (tap, channel chunk) and
  // feeds the three products, so every operand box is loaded ONCE per tap (the logical-chunk ring loaded A_hi and W_hi twice: the
  // 8x8-pixel layers d2 / e3 and the dense layers are L2->SMEM bound).  k-block units (split-K ranges, stages) are then (tap, chunk).
  int fat;
  int w_box3;                                   // halo kernel: map_b is the 3-D [bk][rows][blocks] view, one TMA box per ring stage
  int nstack2;                                  // halo kernel, split 1: A_hi x [W_hi ; W_lo] as ONE MMA of 2N columns (accumulator = [main | correction])
  void* out_lo;                                 // lo plane of a bf16 output (NULL: single bf16 / fp32 output)
  int bk, swizzle;                              // K elements per stage, swizzle bytes (= 2*bk)
  int tile_n_img, tile_h, tile_w;               // M tile = tile_n_img x tile_h x tile_w = 128 output pixels
  int grid_h, grid_w, n_img;                    // output grid per image (before scatter) and image count
  int tile_cols;                                // UMMA N of one CTA
  int n_tiles;                                  // CTAs along N
  int n_valid;                                  // valid output columns in total
  int stages;
  // epilogue: out pixel (n,y,x) -> ((n*OH + y*osy + ooy)*OW + x*osx + oox)*out_ld + col
  int OH, OW, osy, ooy, osx, oox, out_ld, out_f32;
  int nparts, part_n[3], part_act[3];
  int mask_act, mask_ld, mask_coff;
  const float* bias;                            // packed fp32 bias [n_pad] or NULL
  void* out;
  const void* mask_src;
  size_t smem_bytes;
  // halo-resident variant (halo_conv_kernel): the (TH+taps_h-1) x (TW+taps_w-1) input halo of a TW x TH output tile is
  // loaded ONCE into shared memory; every tap's A operand is a shifted UMMA view of it; only weights stream.
  int halo;                                     // 1: use halo_conv_kernel
  int TW, TH, TWp, THp;                         // output tile and halo extents (pixels)
  int mtx, mty;                                 // 128-row accumulators per CTA: (TW/8) x (TH/16)
  int chunk_bytes;                              // bytes of one channel-chunk halo (1024-aligned)
  int kb_per_stage, w_stages, w_stage_bytes;    // weight ring: k-blocks (tap, chunk) per stage
  int tiles_x, tiles_y;
  // strided input (stride-2 forward convolutions): halo_sx parity planes per channel chunk (plane p = input columns
  // sx*x0 - pad_l + p + sx*j, loaded with TMA element stride sx; filter column b = sx*b' + p reads plane p shifted by b'),
  // halo_sy = input rows per output row (the UMMA row-group stride).  TWp = plane width, THp = (TH-1)*sy + taps_h.
  int halo_sx, halo_sy;
  int trace;                                    // SV_HALO_TRACE: record per-CTA phase clocks (debugging)
  // persistent pipelined variant (pconv_kernel): weights resident in shared memory, halo ring, two accumulator sets in TMEM
  int persist, p_stages, p_wbytes, p_grid;
  int p_ntile;                                  // column tile (of n_tiles) this launch descriptor serves: weight rows / output columns p_ntile * tile_cols
  size_t p_smem;
  // split-K (skinny dense GEMMs: few output tiles, long K): blockIdx.z owns kb_per_split k-blocks and stores its raw fp32
  // accumulator to partial[z][m_pad][n_pad]; splitk_finish_kernel sums the splits in a fixed order and applies the epilogue
  int k_splits, kb_per_split, m_pad, n_pad;
  float* partial;
};

// Weight gradient  dW[(tap,ci), co] = sum_pixels X[pixel+tap, ci] * dY[pixel, co]  as a GEMM whose K axis is the
// pixel axis: both operands are "MN-major" (channels contiguous) 64-pixel TMA boxes of NHWC tensors.
struct TcWgradLaunch {
  CUtensorMap map_a, map_b;       // A = layer input (box: cb channels x 64 pixels), B = dY (box: cbn channels x 64 pixels)
  int taps_h, taps_w, pad_t, pad_l, a_stride;
  int cb, ncb, a_swizzle;         // A channel block (64/32/16), blocks per tap, swizzle bytes
  int cbn, b_swizzle;             // B channel block
  int nsub;                       // A sub-boxes per 128-row group (128 / cb)
  int total_sb;                   // taps * ncb
  int groups, groups_per_cta;     // 128-row groups in total / per CTA
  int tile_cols, n_tiles;         // UMMA N per CTA, CTAs along N
  int tile_n_img, tile_h, tile_w; // 64-pixel chunk decomposition
  int grid_h, n_img;
  int nchunks, chunks_per_split, k_splits;
  int a_stages, b_stages;
  int m_pad, n_pad;               // partial buffer dims
  float* partial;
  int first;                      // first-layer row mapping: row = kh*64 + kw*8 + ci
  size_t smem_bytes;
};

// Halo-resident weight gradient (stride-1 convolutions with W % 16 == 0): per pixel tile the X halo and the dY tile are loaded
// once; every (filter tap, channel block) A operand is a shifted MN-major view of the same shared-memory halo.
// An M = 128 row group stacks `nsub` sub-blocks of `cb` channels: adjacent taps (mode 0) or channel chunks of one tap (mode 1).
// K steps are 16 horizontally adjacent pixels.  Two variants:
//   plain   : D[(a,b,ci), co]          - groups enumerate (a,b); taps stacked horizontally (LBO = one pixel); N = Cout_pad
//   N-stack : D[(a,ci), (kw-1-b, co)]  - the filter COLUMN moves to the N axis: B = kw shifted views of the dY tile
//             (LBO = one pixel), taps stacked vertically in M (LBO = one tile row).  kw x fewer, kw x wider MMAs: the MMA
//             rate is bound by the 128x16 A-operand read, so wide N is what makes the small-Cout layers efficient.
struct TcHaloWgrad {
  CUtensorMap map_x, map_dy;
  int taps_h, taps_w, pad_t, pad_l;
  int cb, nchunks, x_swizzle;       // X channel chunk (elements), chunks per pixel
  int cbn, nbchunks, dy_swizzle;    // dY channel chunk, chunks
  int n_pad;                        // UMMA N = columns of the partial buffer
  int nstack, nb;                   // N-stacked variant: N = taps_w blocks of nb channels (block j = filter column taps_w-1-j)
  int TW, TH;                       // pixel tile (K steps = 16 horizontally adjacent pixels)
  int x_tw, x_th, dy_tw, dy_th;     // shared-memory tile extents (pixels) of X and dY
  int x_dx, x_dy, dy_dx;            // TMA start = tile origin + these offsets
  int sy;                           // X rows per output row (2: stride-2 layer through the pixel-pair view)
  int krows;                        // output rows per K step: 1 (16 adjacent pixels of a row) or 2 (8-pixel-wide images: 8 columns x 2 rows)
  int pair_c, pair_sh;              // pair view: channels per pixel of the layer, u = b + pair_sh -> (pair tap u >> 1, parity u & 1)
  int a_lbo, b_lbo;                 // byte distance between the stacked sub-blocks of A (M) / B (N)
  uint32_t goff[32];                // per 128-row group: (byte offset of its first sub-block inside the X tile) >> 4
  int x_chunk_bytes, dy_chunk_bytes, stage_bytes, stages;
  int mode, nsub, gw, gpt;          // grouping, used by the reduce kernel's row map
  int groups, groups_per_cta, m_splits;
  int tiles_x, tiles_y, n_img, tiles, tiles_per_split, k_splits;
  int m_pad;
  float* partial;
  size_t smem_bytes;
};

// N-stacked persistent convolution (stride-1 layers with few output channels: d4 / d5 forward and their dgrads).
//   D[pixel, (b, co)] = sum_{a, ci} X[y + a - pad_t, x, ci] * W[a, b, ci, co]      (one accumulator, N = kw * nb columns)
//   out[y, x, co]     = act(bias[co] + sum_b D[(y, x + b - pad_l), (b, co)])          (epilogue: lane shuffles along x)
// The filter COLUMN b lives on the UMMA N axis, so one 128x16 A-operand read from shared memory feeds kw x more MACs: the
// small-Cout layers stop being bound by the A read (44 cycles per MMA for N <= 32) and run at the tensor pipe's own rate.
// The filter ROW a is a shifted view (whole image rows: offset a * W pixels) of a (R + kh - 1)-row halo that TMA loads once
// per tile; the packed weights [kw*nb][kh*C] stay resident in shared memory for the whole persistent CTA; accumulators are
// double-buffered in TMEM so tile i's epilogue overlaps tile i+1's MMAs.
struct TcNsConv {
  CUtensorMap map_x, map_w;
  // bf16x3 forward ("pair" mode): the input is a bf16 pair (hi plane map_x, lo plane map_x_lo); every weight k-block holds the rows
  // [W_hi-stack ; W_lo-stack] (2 * n_total rows).  A tile runs as two half-tile passes through the halo ring: the hi plane against the
  // whole k-block (ONE MMA of 2 * n_total columns -> accumulator columns [main | correction]), then the lo plane against the W_hi rows
  // (n_total columns -> main).  The epilogue adds main + correction and stores the output pair.
  CUtensorMap map_x_lo;
  int pair, acc1;                   // acc1: TMEM columns of one accumulator (pair: 2 * n_total, else ng * ncols)
  void* out_lo;
  int kh, kw, pad_t, pad_l;
  int W, R, H, n_img;               // block = R rows x W pixels (R * W = 128)
  int mb;                           // blocks per tile (tile = mb * R rows sharing one halo; one accumulator per block)
  int tiles_per_img, tiles;
  int ck, nchunks, pixB;            // channel chunk (elements), chunks per pixel, bytes per pixel per chunk (= swizzle span)
  int nb;                           // N block per filter column (padded output channels)
  int ng, ncols, n_total;           // UMMA N groups (each ncols <= 256 columns), kw * nb
  int nacc;                         // accumulator buffers in TMEM (= groups)
  int groups, lanes;                // epilogue: `groups` warp sets take tiles round-robin (one accumulator each); inside a set
                                    // `lanes` warps per TMEM quarter split the tile's 8-channel blocks
  int co_splits;                    // output channels split over CTAs (CTA c serves split c % co_splits for its whole life): halves the
                                    // resident weights so that more halo stages fit; nb / n_total / ncols are PER SPLIT
  int nstages;                      // halo ring depth (the kernel is latency-bound on the halo loads below ~3 stages)
  int halo_rows, chunk_bytes, stage_bytes;
  int num_kb, wk_bytes, w_bytes;    // weight k-blocks (kh * nchunks), bytes per k-block, total (1024-aligned)
  int n_box;                        // rows per weight TMA box (<= 256)
  int xch_off;                      // W = 64: shared-memory exchange area for the shuffles that cross the two warps of a row
  int n_valid, out_ld, out_f32, act, mask_act, mask_ld, mask_coff;
  const float* bias;
  void* out;
  const void* mask_src;
  size_t smem_bytes;
  int grid;
  int debug;                        // SV_NS_DEBUG bit 0: epilogue only drains TMEM (no shuffles / stores); bit 1: no MMAs issued
};

struct TcLayer {
  bool fwd_ok = false, dgrad_ok = false, wgrad_ok = false;
  bool first = false;                 // first conv of an encoder: reads the staged, padded bf16 image (tc_stage_first)
  int fwd_launches = 0, dgrad_launches = 0, wgrad_launches = 0;
  int n_dgrad = 0;
  TcLaunch fwd{}, dgrad[4]{};
  TcWgradLaunch wg{};
  bool wg_halo = false;
  TcHaloWgrad hw{};
  bool split_fwd = false;                  // bf16x3: the forward pass multiplies bf16 pairs (see TcLaunch::split)
  bool fwd_ns = false, dgrad_ns = false;   // N-stacked persistent kernel replaces the per-tap / halo kernel
  bool dgrad_merged = false;               // stride-2 dgrad: the 4 parity classes run as one launch (igemm4_kernel / halo4_kernel)
  bool fwd_pair = false;                   // stride-2 forward on the persistent kernel through the pixel-pair view [W/2][2C]
  bool first_pair = false;                 // first layer forward on the halo kernel through the pixel-PAIR view of the staged image
  TcNsConv ns_fwd{}, ns_dgrad{};
  size_t w_nsf_off = 0, w_nsd_off = 0;
  size_t wg_partial_off = 0;
  // geometry of the packed operands
  int n_pad_fwd = 0, ci_pad = 0;      // fwd : B = [n_pad_fwd][taps][ci_pad]
  int n_pad_dg = 0, co_pad = 0;       // dgrad: B = [classes][n_pad_dg][taps'][co_pad]
  int dg_taps_h = 0, dg_taps_w = 0;
  size_t w_fwd_off = 0, w_dgrad_off = 0, bias_off = 0, bytes = 0;   // offsets inside this layer's workspace slice
  size_t sk_fwd_off = 0, sk_dgrad_off = 0;                            // split-K partial buffers
  char* ws = nullptr;
  int in_dt = 0, out_dt = 0;
};

// plan (no CUDA calls), workspace size, bind (creates TMA descriptors; returns NULL or an error text)
void tc_plan_layer(TcLayer& t, const ConvGeom& g, int in_dt, int out_dt, bool has_internal_input, bool has_dgrad, bool first_layer,
                   bool split_fwd = false);
// staged first-layer input: [B][H][W+8][8] bf16, real pixel x at column x+2, channels 3..7 and the borders zero
// (split: channels 0-2 = hi, 3-5 = lo of the image value, 6-7 zero)
size_t tc_first_stage_bytes(int B, int H, int W);
void tc_stage_first(const float* inputs, void* xp, int coff, int B, int H, int W, bool split, cudaStream_t s);
size_t tc_workspace_bytes(const TcLayer& t, const ConvGeom& g);
// in_lo / out_lo: the lo planes of the layer input / output (bf16x3 forward; NULL otherwise)
const char* tc_bind_layer(TcLayer& t, const ConvGeom& g, const void* in, void* out, void* dout, void* din,
                          const void* mask_src, int mask_act, char* ws, const void* in_lo = nullptr, void* out_lo = nullptr);
// One multi-tensor launch that refreshes every packed operand from the fp32 master weights.
struct TcPackTable;
TcPackTable* tc_pack_table_create(TcLayer* const* layers, const ConvGeom* const* geoms, int n, const char** err);
void tc_pack_table_destroy(TcPackTable* t);
int tc_repack_all(TcPackTable* t, const float* params, cudaStream_t s);  // returns #launches

void tc_conv_fwd(TcLayer& t, cudaStream_t s);
void tc_conv_dgrad(TcLayer& t, cudaStream_t s);
void tc_conv_wgrad(TcLayer& t, const ConvGeom& g, float* grads, cudaStream_t s);
const char* tc_last_error();
int tc_halo_trace_read(unsigned long long* out, int max_ctas);   // 8 words per CTA, see g_halo_trace

}  // namespace sv
