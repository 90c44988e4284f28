// libsplitvae host side: model plan (layers, buffers, parameter arena), the C-ABI of
// include/splitvae.h, and the orchestration of one train step.  All device work is issued on the
// caller's stream with no allocation or synchronisation so a whole step can be graph-captured.
//
// Mirrors: vae/main.py:63-74 (model/optimizer construction), vae/model.py:100-135,158-169,189-200,
// 237-248 (forward), vae/trainer.py:120-173 (train steps).
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/splitvae.h"
#include "common.cuh"
#include "kernels.h"
#include "tc_kernels.h"

using namespace sv;

namespace {

constexpr int SV_FLAG_PLAN_ONLY = 1;  // build the plan without touching CUDA (inventory / sizes only)
constexpr int SV_FLAG_NO_TC = 2;      // bf16 precision but reference kernels everywhere (A/B debugging)

char g_create_error[512] = "";

struct Buf { size_t off = 0, bytes = 0; };

struct Layer {
  std::string name;
  ConvGeom g{};
  int in = -1, out = -1, dout = -1, din = -1;  // buffer ids (-1: external input / no dgrad)
  int in_dt = DT_F32, out_dt = DT_F32;
  int mask_act = ACT_NONE;                    // activation of the producer of `in`
  int xp = -1;                                // first conv of an encoder: staged padded bf16 image (tensor-core path)
  bool split_fwd = false;                     // bf16x3: the forward product multiplies bf16 pairs (every layer but the decoders' d5)
  int in_lo = -1, out_lo = -1;                // lo planes of the input / output activation (bf16x3)
  TcLayer tc;                                 // tensor-core plan (tc_kernels.h)
};

struct Var { std::string name; int ndim; int shape[4]; long long off, count; };

struct Decoder { int d1, d2, d3, d4, d5; int D1, D2, U1, D3, U2, D4, U3, OUT; int dD1, dD2, dU1, dD3, dU2, dD4, dU3, dOUT; int dz; };
struct ConvEnc { int e1, e2, e3, heads; int A1, A2, A3, HEADS; int dA1, dA2, dA3, dHEADS; };
struct GmEnc {
  int h1, h2, h3, yb0e1, yb2, ydense, yheads, zheads;
  int A1, A2, A3, YB0E1, YH2, LOGITS, Y, YT, U, YHEADS, HSUM, HEADS;
  int dA1, dA2, dA3, dYB0E1, dYH2, dLOGITS, dY, dYHEADS, dHSUM, dHEADS;
};

}  // namespace

struct sv_handle {
  sv_config cfg{};
  int act_dt = DT_BF16;
  bool round_w = true, plan_only = false, use_tc = true;
  bool split = false;                         // SV_PRECISION_BF16X3: forward operands are bf16 pairs (hi + lo), backward single bf16
  std::vector<int> lo_of;                     // buffer id -> id of its lo plane (-1: none)
  int B = 0, H = 0, W = 0, F = 0, K = 0;
  std::vector<Var> vars;
  long long arena_floats = 0;
  std::vector<Buf> bufs;
  size_t ws_bytes = 0;
  std::vector<Layer> layers;
  ConvEnc enc_x{}, enc_xh{};
  GmEnc gm_enc{};
  Decoder dec_x{}, dec_xh{};
  int ZCAT = -1, EPS_G = -1, EPS_L = -1, Z_G = -1, Z_L = -1, ZM_G = -1, ZS_G = -1, ZM_L = -1, ZS_L = -1;
  int ZPM_OUT = -1, ZPS_OUT = -1, SCALARS = -1, PARTIALS = -1, KLPART = -1, COLSUM = -1, ADAM = -1, TCWS = -1;
  int XP[2] = {-1, -1};     // staged first-layer images (x, x_hat) for the tensor-core path
  long long seg_split = 0;  // arena offset where the decoders start
  // bound buffers
  float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr;
  char* ws = nullptr;
  bool bound = false;
  long long launches = 0;
  char err[512] = "";
  const float* last_inputs = nullptr;  // inputs of the step in flight (first-layer wgrad reads them)
  unsigned long long seed = 0x5EEDull;
  TcPackTable* pack = nullptr;
  // Backward / optimizer SEGMENTS (gradient buckets of the data-parallel all-reduce, in the order their gradients become final):
  //   0 decoders   1 encoder tops (heads, e3 | the GM dense stack, h_block.2)   2 encoder bottoms (e2, e1 | h_block.1, h_block.0)
  // Each owns its layers' arena ranges (<= one contiguous range per encoder) and a re-pack table of its layers' operand copies.
  static constexpr int kSegs = 3;
  std::vector<std::pair<long long, long long>> seg_ranges[kSegs];
  TcPackTable* pack_seg[kSegs] = {nullptr, nullptr, nullptr};
  cudaStream_t opt = nullptr;                      // optimizer stream: Adam + re-pack of segment 0 overlap the encoders' backward
  cudaEvent_t ev_opt_fork = nullptr, ev_opt_join = nullptr;
  // the x / x_hat encoders and the two decoders are independent: they run on two streams (fork/join with events, which
  // CUDA-graph capture turns into parallel branches)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int COLSUM2 = -1;
  bool two_streams = true;
  bool gm = false;            // gmvae-type encoder_x (lggmvae, gmvae)
  bool has_local = true;      // the x_hat encoder / decoder pair exists (false: plain GMVAE, vae/model.py:277-299)
  // weight gradients run on one auxiliary stream per branch stream (dgrad chain = critical path, wgrad + reduce fill the gaps)
  // (SV_WGRAD_STREAMS = n streams per branch, used round-robin by consecutive layers; 0 = none.  Two per branch let wgrad(L-1) start
  // while wgrad(L) and its split-K reduce are still queued: on one in-order stream the encoders' last wgrads formed a serial tail)
  static constexpr int kMaxAux = 8;
  cudaStream_t aux[2][kMaxAux] = {};
  cudaEvent_t ev_aux[2][kMaxAux] = {}, ev_aux_join[2][kMaxAux] = {};
  int aux_n = 0, aux_rr[2] = {0, 0};
  bool aux_dirty[2][kMaxAux] = {};                            // forked since the last join
  bool defer_join = false;                                    // inside sv_backward_segment_deferred
  bool wgrad_streams = false;
  // multi-tensor bias gradients, one table per backward branch: 0 decoder_x, 1 decoder_x_hat, 2 encoder_x, 3 encoder_x_hat
  bool cs_on = false;
  // (encoders: part 0 = the layers of segment 1, part 1 = segment 2; decoders: part 0 only)
  int CSP[4][2] = {{-1, -1}, {-1, -1}, {-1, -1}, {-1, -1}};
  ColsumTable* cs[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  float* cs_fold[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};   // [decoder][d2..d5]: external partials
  cudaEvent_t ev_seg_done = nullptr;
  cudaGraph_t graph = nullptr;            // sv_capture_graph: one captured sv_train_step
  cudaGraphExec_t graph_exec = nullptr;
  long long graph_kernels = 0;
};

namespace {

}  // namespace
// Off by default: measured on the C2 step (two branches interleaving on two streams) PDL is a LOSS - 1.906 -> 1.930 ms - because the
// next kernel's CTAs take their SM as soon as the previous kernel's CTA leaves it and then idle until that whole grid has drained,
// where the other branch's kernel would have run; back-to-back launches of ONE chain gain ~3 us each (profiles/r02_pdl_*.txt).
bool sv::pdl_enabled() { static const bool on = getenv("SV_PDL") && getenv("SV_PDL")[0] == '1'; return on; }
namespace {
bool getenv_off(const char* name) { const char* v = getenv(name); return v && *v == '0'; }

sv_status fail(sv_handle* h, sv_status code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(h ? h->err : g_create_error, 512, fmt, ap);
  va_end(ap);
  return code;
}

void same_pad(int in, int k, int s, int& before) {
  const int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  before = total / 2;
}

int new_buf(sv_handle* h, size_t bytes) {
  Buf b;
  b.off = h->ws_bytes;
  b.bytes = bytes;
  h->ws_bytes += (bytes + 1023) / 1024 * 1024;
  h->bufs.push_back(b);
  return (int)h->bufs.size() - 1;
}
size_t esz(int dt) { return dt == DT_F32 ? 4 : 2; }
int act_buf(sv_handle* h, long long elems) { return new_buf(h, (size_t)elems * esz(h->act_dt)); }
// forward activation that feeds another forward layer: in the bf16x3 mode it is a bf16 PAIR, the lo plane is a twin buffer
int fwd_buf(sv_handle* h, long long elems) {
  const int hi = act_buf(h, elems);
  if (h->split) {
    const int lo = act_buf(h, elems);
    h->lo_of.resize(h->bufs.size(), -1);
    h->lo_of[hi] = lo;
  }
  return hi;
}
int lo_buf(const sv_handle* h, int id) { return (id >= 0 && id < (int)h->lo_of.size()) ? h->lo_of[id] : -1; }
int f32_buf(sv_handle* h, long long elems) { return new_buf(h, (size_t)elems * 4); }

long long add_var(sv_handle* h, const std::string& name, int ndim, const int* shape) {
  Var v;
  v.name = name;
  v.ndim = ndim;
  v.count = 1;
  for (int i = 0; i < 4; ++i) { v.shape[i] = i < ndim ? shape[i] : 0; if (i < ndim) v.count *= shape[i]; }
  v.off = h->arena_floats;
  h->arena_floats += (v.count + 63) / 64 * 64;  // 256-byte aligned variables
  h->vars.push_back(v);
  return v.off;
}

struct PartSpec { const char* name; int n; int act; };

// Creates the Keras variables (kernel, bias per part, in Keras order) and the layer record.
int add_layer(sv_handle* h, const std::string& prefix, int kh, int kw, int stride, int Hi, int Wi, int Ci,
              std::vector<PartSpec> parts) {
  Layer L;
  ConvGeom& g = L.g;
  g.B = h->B; g.Hi = Hi; g.Wi = Wi; g.Ci = Ci;
  g.kh = kh; g.kw = kw; g.stride = stride;
  g.Ho = (Hi + stride - 1) / stride; g.Wo = (Wi + stride - 1) / stride;
  same_pad(Hi, kh, stride, g.pt);
  same_pad(Wi, kw, stride, g.pl);
  g.nparts = (int)parts.size();
  g.Co = 0;
  L.name = prefix + "." + parts[0].name;
  for (int j = 0; j < g.nparts; ++j) {
    const std::string vn = prefix + "." + parts[j].name;
    g.part_n[j] = parts[j].n;
    g.part_act[j] = parts[j].act;
    if (kh == 1 && kw == 1 && Hi == 1 && Wi == 1) {
      int shp[2] = {Ci, parts[j].n};
      g.part_w[j] = add_var(h, vn + ".kernel", 2, shp);
    } else {
      int shp[4] = {kh, kw, Ci, parts[j].n};
      g.part_w[j] = add_var(h, vn + ".kernel", 4, shp);
    }
    int bs[1] = {parts[j].n};
    g.part_b[j] = add_var(h, vn + ".bias", 1, bs);
    g.Co += parts[j].n;
  }
  g.in_ld = Ci; g.in_coff = 0; g.out_ld = g.Co; g.dout_ld = g.Co; g.din_ld = Ci;
  h->layers.push_back(L);
  return (int)h->layers.size() - 1;
}

void wire(sv_handle* h, int li, int in, int in_dt, int in_ld, int in_coff, int out, int out_dt, int out_ld, int dout,
          int dout_ld, int din, int din_ld, int mask_act) {
  Layer& L = h->layers[li];
  L.in = in; L.in_dt = in_dt; L.g.in_ld = in_ld; L.g.in_coff = in_coff;
  L.out = out; L.out_dt = out_dt; L.g.out_ld = out_ld;
  L.dout = dout; L.g.dout_ld = dout_ld;
  L.din = din; L.g.din_ld = din_ld;
  L.mask_act = mask_act;
}

ConvEnc build_conv_encoder(sv_handle* h, const char* prefix, int coff) {
  ConvEnc e{};
  const int B = h->B, H = h->H, W = h->W, T = h->act_dt;
  e.e1 = add_layer(h, prefix, 6, 6, 2, H, W, 3, {{"e1", 32, ACT_RELU}});
  e.e2 = add_layer(h, prefix, 6, 6, 2, H / 2, W / 2, 32, {{"e2", 64, ACT_RELU}});
  e.e3 = add_layer(h, prefix, 4, 4, 2, H / 4, W / 4, 64, {{"e3", 128, ACT_RELU}});
  e.heads = add_layer(h, prefix, 1, 1, 1, 1, 1, h->F, {{"e4_mean", 128, ACT_NONE}, {"e4_sd", 128, ACT_SOFTPLUS}});
  e.A1 = fwd_buf(h, (long long)B * (H / 2) * (W / 2) * 32); e.dA1 = act_buf(h, (long long)B * (H / 2) * (W / 2) * 32);
  e.A2 = fwd_buf(h, (long long)B * (H / 4) * (W / 4) * 64); e.dA2 = act_buf(h, (long long)B * (H / 4) * (W / 4) * 64);
  e.A3 = fwd_buf(h, (long long)B * h->F); e.dA3 = act_buf(h, (long long)B * h->F);
  e.HEADS = f32_buf(h, (long long)B * 256); e.dHEADS = act_buf(h, (long long)B * 256);
  wire(h, e.e1, -1, DT_F32, 6, coff, e.A1, T, 32, e.dA1, 32, -1, 0, ACT_NONE);
  wire(h, e.e2, e.A1, T, 32, 0, e.A2, T, 64, e.dA2, 64, e.dA1, 32, ACT_RELU);
  wire(h, e.e3, e.A2, T, 64, 0, e.A3, T, 128, e.dA3, 128, e.dA2, 64, ACT_RELU);
  wire(h, e.heads, e.A3, T, h->F, 0, e.HEADS, DT_F32, 256, e.dHEADS, 256, e.dA3, h->F, ACT_RELU);
  return e;
}

GmEnc build_gm_encoder(sv_handle* h, const char* prefix) {
  GmEnc e{};
  const int B = h->B, H = h->H, W = h->W, T = h->act_dt, F = h->F, K = h->K;
  // Keras variable order = attribute order of Encoder.__init__ (vae/model.py:49-76)
  e.h1 = add_layer(h, prefix, 6, 6, 2, H, W, 3, {{"h_block.0", 128, ACT_ELU}});
  e.h2 = add_layer(h, prefix, 6, 6, 2, H / 2, W / 2, 128, {{"h_block.1", 128, ACT_ELU}});
  e.h3 = add_layer(h, prefix, 4, 4, 2, H / 4, W / 4, 128, {{"h_block.2", 128, ACT_ELU}});
  // y_block.0 and e1 share their input and are fused along the output axis, but Keras creates
  // y_block.0, y_block.2, y_dense, h_top_dense, z_prior_mean, z_prior_sig, e1, z_mean, z_sig in
  // that order; the variables of the fused layer are therefore declared part by part below.
  e.yb0e1 = add_layer(h, prefix, 1, 1, 1, 1, 1, F, {{"y_block.0", 1024, ACT_ELU}});
  e.yb2 = add_layer(h, prefix, 1, 1, 1, 1, 1, 1024, {{"y_block.2", 128, ACT_ELU}});
  e.ydense = add_layer(h, prefix, 1, 1, 1, 1, 1, 128, {{"y_dense", K, ACT_NONE}});
  e.yheads = add_layer(h, prefix, 1, 1, 1, 1, 1, K,
                       {{"h_top_dense", 512, ACT_ELU}, {"z_prior_mean", 128, ACT_NONE}, {"z_prior_sig", 128, ACT_SOFTPLUS}});
  {  // append e1 as the second part of the fused (y_block.0 | e1) layer
    Layer& L = h->layers[e.yb0e1];
    int shp[2] = {F, 512};
    L.g.part_w[1] = add_var(h, std::string(prefix) + ".e1.kernel", 2, shp);
    int bs[1] = {512};
    L.g.part_b[1] = add_var(h, std::string(prefix) + ".e1.bias", 1, bs);
    L.g.part_n[1] = 512; L.g.part_act[1] = ACT_ELU; L.g.nparts = 2; L.g.Co = 1536;
  }
  e.zheads = add_layer(h, prefix, 1, 1, 1, 1, 1, 512, {{"z_mean", 128, ACT_NONE}, {"z_sig", 128, ACT_SOFTPLUS}});

  e.A1 = fwd_buf(h, (long long)B * (H / 2) * (W / 2) * 128); e.dA1 = act_buf(h, (long long)B * (H / 2) * (W / 2) * 128);
  e.A2 = fwd_buf(h, (long long)B * (H / 4) * (W / 4) * 128); e.dA2 = act_buf(h, (long long)B * (H / 4) * (W / 4) * 128);
  e.A3 = fwd_buf(h, (long long)B * F); e.dA3 = act_buf(h, (long long)B * F);
  e.YB0E1 = fwd_buf(h, (long long)B * 1536); e.dYB0E1 = act_buf(h, (long long)B * 1536);
  e.YH2 = fwd_buf(h, (long long)B * 128); e.dYH2 = act_buf(h, (long long)B * 128);
  e.LOGITS = f32_buf(h, (long long)B * 32); e.dLOGITS = act_buf(h, (long long)B * 32);
  e.Y = f32_buf(h, (long long)B * 32); e.YT = fwd_buf(h, (long long)B * 32); e.U = f32_buf(h, (long long)B * 32);
  e.dY = act_buf(h, (long long)B * 32);
  e.YHEADS = f32_buf(h, (long long)B * 768); e.dYHEADS = act_buf(h, (long long)B * 768);
  e.HSUM = fwd_buf(h, (long long)B * 512); e.dHSUM = act_buf(h, (long long)B * 512);
  e.HEADS = f32_buf(h, (long long)B * 256); e.dHEADS = act_buf(h, (long long)B * 256);

  wire(h, e.h1, -1, DT_F32, 6, 0, e.A1, T, 128, e.dA1, 128, -1, 0, ACT_NONE);
  wire(h, e.h2, e.A1, T, 128, 0, e.A2, T, 128, e.dA2, 128, e.dA1, 128, ACT_ELU);
  wire(h, e.h3, e.A2, T, 128, 0, e.A3, T, 128, e.dA3, 128, e.dA2, 128, ACT_ELU);
  wire(h, e.yb0e1, e.A3, T, F, 0, e.YB0E1, T, 1536, e.dYB0E1, 1536, e.dA3, F, ACT_ELU);
  wire(h, e.yb2, e.YB0E1, T, 1536, 0, e.YH2, T, 128, e.dYH2, 128, e.dYB0E1, 1536, ACT_ELU);
  wire(h, e.ydense, e.YH2, T, 128, 0, e.LOGITS, DT_F32, 32, e.dLOGITS, 32, e.dYH2, 128, ACT_ELU);
  wire(h, e.yheads, e.YT, T, 32, 0, e.YHEADS, DT_F32, 768, e.dYHEADS, 768, e.dY, 32, ACT_NONE);
  wire(h, e.zheads, e.HSUM, T, 512, 0, e.HEADS, DT_F32, 256, e.dHEADS, 256, e.dHSUM, 512, ACT_NONE);
  return e;
}

Decoder build_decoder(sv_handle* h, const char* prefix, int L, int zcoff, int dz_buf, int dz_ld, int dout_ld) {
  Decoder d{};
  const int B = h->B, H = h->H, W = h->W, T = h->act_dt, F = h->F;
  d.d1 = add_layer(h, prefix, 1, 1, 1, 1, 1, L, {{"d1", F, ACT_RELU}});
  d.d2 = add_layer(h, prefix, 4, 4, 1, H / 8, W / 8, 128, {{"d2", 128, ACT_RELU}});
  d.d3 = add_layer(h, prefix, 4, 4, 1, H / 4, W / 4, 128, {{"d3", 64, ACT_RELU}});
  d.d4 = add_layer(h, prefix, 6, 6, 1, H / 2, W / 2, 64, {{"d4", 32, ACT_RELU}});
  d.d5 = add_layer(h, prefix, 6, 6, 1, H, W, 32, {{"d5", 6, ACT_NONE}});
  const long long p8 = (long long)B * (H / 8) * (W / 8), p4 = p8 * 4, p2 = p8 * 16, p1 = p8 * 64;
  d.D1 = fwd_buf(h, p8 * 128); d.dD1 = act_buf(h, p8 * 128);
  d.D2 = fwd_buf(h, p8 * 128); d.dD2 = act_buf(h, p8 * 128);
  d.U1 = fwd_buf(h, p4 * 128); d.dU1 = act_buf(h, p4 * 128);
  d.D3 = fwd_buf(h, p4 * 64); d.dD3 = act_buf(h, p4 * 64);
  d.U2 = fwd_buf(h, p2 * 64); d.dU2 = act_buf(h, p2 * 64);
  d.D4 = fwd_buf(h, p2 * 32); d.dD4 = act_buf(h, p2 * 32);
  d.U3 = act_buf(h, p1 * 32); d.dU3 = act_buf(h, p1 * 32);      // (input of d5: single bf16 in every mode)
  d.OUT = f32_buf(h, p1 * 6); d.dOUT = act_buf(h, p1 * dout_ld);
  d.dz = dz_buf;
  wire(h, d.d1, h->ZCAT, T, 256, zcoff, d.D1, T, F, d.dD1, F, dz_buf, dz_ld, ACT_NONE);
  wire(h, d.d2, d.D1, T, 128, 0, d.D2, T, 128, d.dD2, 128, d.dD1, 128, ACT_RELU);
  wire(h, d.d3, d.U1, T, 128, 0, d.D3, T, 64, d.dD3, 64, d.dU1, 128, ACT_NONE);
  wire(h, d.d4, d.U2, T, 64, 0, d.D4, T, 32, d.dD4, 32, d.dU2, 64, ACT_NONE);
  wire(h, d.d5, d.U3, T, 32, 0, d.OUT, DT_F32, 6, d.dOUT, dout_ld, d.dU3, 32, ACT_NONE);
  return d;
}

void* bp(sv_handle* h, int id) { return id < 0 ? nullptr : (void*)(h->ws + h->bufs[id].off); }

// layers of backward branch `which` (0 decoder_x, 1 decoder_x_hat, 2 encoder_x, 3 encoder_x_hat) in backward order;
// part 0 / 1 = the encoder layers of segment 1 / 2 (decoders: everything is part 0), part -1 = all
std::vector<int> branch_layers(const sv_handle* h, int which, int part) {
  auto dec = [&](const Decoder& d) { return part == 1 ? std::vector<int>{} : std::vector<int>{d.d5, d.d4, d.d3, d.d2, d.d1}; };
  auto enc = [&](const ConvEnc& e) {
    return part == 0 ? std::vector<int>{e.heads, e.e3} : part == 1 ? std::vector<int>{e.e2, e.e1} : std::vector<int>{e.heads, e.e3, e.e2, e.e1};
  };
  if (which == 0) return dec(h->dec_x);
  if (which == 1) return h->has_local ? dec(h->dec_xh) : std::vector<int>{};
  if (which == 3) return h->has_local ? enc(h->enc_xh) : std::vector<int>{};
  if (h->gm) {
    const GmEnc& e = h->gm_enc;
    if (part == 0) return {e.zheads, e.yheads, e.ydense, e.yb2, e.yb0e1, e.h3};
    if (part == 1) return {e.h2, e.h1};
    return {e.zheads, e.yheads, e.ydense, e.yb2, e.yb0e1, e.h3, e.h2, e.h1};
  }
  return enc(h->enc_x);
}
std::vector<int> segment_layers(const sv_handle* h, int seg) {
  std::vector<int> v;
  for (int b = seg == 0 ? 0 : 2; b < (seg == 0 ? 2 : 4); ++b)
    for (int li : branch_layers(h, b, seg == 0 ? 0 : seg - 1)) v.push_back(li);
  return v;
}
// Decoder layers whose dY is produced by a CUDA-core kernel of the backward pass get their bias-gradient partials from that kernel
// (d2-d4: upsample2x_bwd, d5: pixel_loss) instead of a second pass over dY; index into `fold` = 0..3 for d2..d5.
std::vector<ColsumSpec> branch_specs(sv_handle* h, int which, int part, bool with_ptrs, int* fold_index = nullptr) {
  std::vector<ColsumSpec> v;
  const bool fold = which < 2 && !getenv_off("SV_FOLD_COLSUM");
  const Decoder* d = which == 0 ? &h->dec_x : which == 1 ? &h->dec_xh : nullptr;
  if (fold_index) for (int k = 0; k < 4; ++k) fold_index[k] = -1;
  for (int li : branch_layers(h, which, part)) {
    ColsumSpec sp{};
    sp.g = h->layers[li].g;
    sp.dout = with_ptrs ? bp(h, h->layers[li].dout) : nullptr;
    if (fold && d) {
      const int B = h->B, H = h->H, W = h->W;
      int k = -1;
      if (li == d->d2) { k = 0; sp.ext_chunks = upsample2x_bwd_blocks(B, H / 8, W / 8, 128); }
      else if (li == d->d3) { k = 1; sp.ext_chunks = upsample2x_bwd_blocks(B, H / 4, W / 4, 64); }
      else if (li == d->d4) { k = 2; sp.ext_chunks = upsample2x_bwd_blocks(B, H / 2, W / 2, 32); }
      else if (li == d->d5 && sp.g.dout_ld == 16) { k = 3; sp.ext_chunks = pixel_loss_blocks((long long)B * H * W); }
      if (k >= 0 && fold_index) fold_index[k] = (int)v.size();
    }
    v.push_back(sp);
  }
  return v;
}

unsigned long long* noise_counter(sv_handle* h) { return (unsigned long long*)((char*)bp(h, h->ADAM) + 64); }

LatentBufs latent_bufs(sv_handle* h) {
  LatentBufs L{};
  const bool gm = h->gm;
  L.heads_g = (const float*)bp(h, gm ? h->gm_enc.HEADS : h->enc_x.HEADS);
  L.heads_l = h->has_local ? (const float*)bp(h, h->enc_xh.HEADS) : nullptr;
  L.eps_g = (float*)bp(h, h->EPS_G); L.eps_l = (float*)bp(h, h->EPS_L);
  L.z_g = (float*)bp(h, h->Z_G); L.z_l = (float*)bp(h, h->Z_L);
  L.zm_g = (float*)bp(h, h->ZM_G); L.zs_g = (float*)bp(h, h->ZS_G);
  L.zm_l = (float*)bp(h, h->ZM_L); L.zs_l = (float*)bp(h, h->ZS_L);
  L.zcat = bp(h, h->ZCAT);
  L.zcat_lo = (bf16*)bp(h, lo_buf(h, h->ZCAT));
  L.dzcat = bp(h, h->dec_x.dz);
  L.dzl2 = h->has_local ? bp(h, h->dec_xh.dz) : nullptr;
  L.dheads_g = bp(h, gm ? h->gm_enc.dHEADS : h->enc_x.dHEADS);
  L.dheads_l = h->has_local ? bp(h, h->enc_xh.dHEADS) : nullptr;
  L.yheads = gm ? (const float*)bp(h, h->gm_enc.YHEADS) : nullptr;
  return L;
}

// ---- per-layer execution ------------------------------------------------------------------
void layer_fwd(sv_handle* h, int li, const float* ext_in, cudaStream_t s) {
  Layer& L = h->layers[li];
  const void* in = L.in < 0 ? (const void*)ext_in : bp(h, L.in);
  if (h->use_tc && L.tc.fwd_ok) {
    tc_conv_fwd(L.tc, s);
    h->launches += L.tc.fwd_launches;
  } else {
    ref_conv_fwd(L.g, in, L.in_dt, h->params, bp(h, L.out), L.out_dt, h->round_w, s);
    h->launches += 1;
  }
}

// auxiliary stream paired with branch stream s (or s itself when the feature is off); makes it wait for everything issued on s
cudaStream_t wgrad_stream(sv_handle* h, cudaStream_t s) {
  if (!h->wgrad_streams || !h->cs_on) return s;
  const int k = (s == h->side) ? 1 : 0;
  const int j = h->aux_rr[k]++ % h->aux_n;
  cudaEventRecord(h->ev_aux[k][j], s);
  cudaStreamWaitEvent(h->aux[k][j], h->ev_aux[k][j], 0);
  h->aux_dirty[k][j] = true;
  return h->aux[k][j];
}
void join_wgrad_stream(sv_handle* h, cudaStream_t s) {
  if (!h->wgrad_streams || !h->cs_on) return;
  const int k = (s == h->side) ? 1 : 0;
  const int used = h->aux_rr[k] < h->aux_n ? h->aux_rr[k] : h->aux_n;   // (an untouched stream is not part of a graph capture)
  for (int j = 0; j < used; ++j) {
    cudaEventRecord(h->ev_aux_join[k][j], h->aux[k][j]);
    cudaStreamWaitEvent(s, h->ev_aux_join[k][j], 0);
    h->aux_dirty[k][j] = false;
  }
  h->aux_rr[k] = 0;
}
// Deferred mode (sv_backward_segment_deferred): the weight-gradient / bias-gradient streams of BOTH branches are joined into
// `target` (the optimizer / reduction stream) instead of the chain's stream, so the dgrad chain of the next segment starts while
// this segment's weight gradients are still running (they occupy 37 SMs each: the chain used to idle ~150 us behind them at
// the decoder -> encoder boundary and ~60 us between the encoder halves).
void join_all_wgrad_streams(sv_handle* h, cudaStream_t target) {
  if (!h->wgrad_streams || !h->cs_on) return;
  for (int k = 0; k < 2; ++k) {
    for (int j = 0; j < h->aux_n; ++j)
      if (h->aux_dirty[k][j]) {
        cudaEventRecord(h->ev_aux_join[k][j], h->aux[k][j]);
        cudaStreamWaitEvent(target, h->ev_aux_join[k][j], 0);
        h->aux_dirty[k][j] = false;
      }
    h->aux_rr[k] = 0;
  }
}

void layer_bwd(sv_handle* h, int li, const float* ext_in, cudaStream_t s) {
  Layer& L = h->layers[li];
  const void* in = L.in < 0 ? (const void*)ext_in : bp(h, L.in);
  const int T = h->act_dt;
  if (h->use_tc && L.tc.wgrad_ok) {
    tc_conv_wgrad(L.tc, L.g, h->grads, wgrad_stream(h, s));   // dY(L) and X(L) are complete on s at this point
    h->launches += L.tc.wgrad_launches;
  } else {
    ref_conv_wgrad(L.g, in, L.in_dt, bp(h, L.dout), T, h->grads, h->round_w, s);
    h->launches += 1;
  }
  if (!h->cs_on) {   // (bf16 mode: one multi-tensor launch pair per branch instead, see branch_bias_grads)
    bias_grad(L.g, bp(h, L.dout), T, (float*)bp(h, s == h->side ? h->COLSUM2 : h->COLSUM), h->grads, s);
    h->launches += 2;
  }
  if (L.din >= 0) {
    if (h->use_tc && L.tc.dgrad_ok) {
      tc_conv_dgrad(L.tc, s);
      h->launches += L.tc.dgrad_launches;
    } else {
      ref_conv_dgrad(L.g, bp(h, L.dout), T, h->params, bp(h, L.din), in, L.in_dt, L.mask_act, h->round_w, s);
      h->launches += 1;
    }
  }
}

void branch_bias_grads(sv_handle* h, int which, int part, cudaStream_t s) {
  if (h->defer_join) {     // every dY of the branch is complete on s here: the column sums leave the chain too
    if (h->cs_on && h->cs[which][part]) h->launches += colsum_table_run(h->cs[which][part], h->grads, wgrad_stream(h, s));
    return;
  }
  if (h->cs_on && h->cs[which][part]) h->launches += colsum_table_run(h->cs[which][part], h->grads, s);
  join_wgrad_stream(h, s);
}

void conv_encoder_fwd(sv_handle* h, const ConvEnc& e, const float* inputs, cudaStream_t s) {
  layer_fwd(h, e.e1, inputs, s);
  layer_fwd(h, e.e2, nullptr, s);
  layer_fwd(h, e.e3, nullptr, s);
  layer_fwd(h, e.heads, nullptr, s);
}
void conv_encoder_bwd(sv_handle* h, const ConvEnc& e, const float* inputs, int part, cudaStream_t s) {
  if (part == 0) {
    layer_bwd(h, e.heads, nullptr, s);
    layer_bwd(h, e.e3, nullptr, s);
  } else {
    layer_bwd(h, e.e2, nullptr, s);
    layer_bwd(h, e.e1, inputs, s);
  }
}

void gm_encoder_fwd(sv_handle* h, const float* inputs, const float* u, cudaStream_t s) {
  const GmEnc& e = h->gm_enc;
  layer_fwd(h, e.h1, inputs, s);
  layer_fwd(h, e.h2, nullptr, s);
  layer_fwd(h, e.h3, nullptr, s);
  layer_fwd(h, e.yb0e1, nullptr, s);
  layer_fwd(h, e.yb2, nullptr, s);
  layer_fwd(h, e.ydense, nullptr, s);
  gumbel_fwd((const float*)bp(h, e.LOGITS), u, (float*)bp(h, e.U), (float*)bp(h, e.Y), bp(h, e.YT), bp(h, lo_buf(h, e.YT)), h->act_dt, h->B,
             h->K, h->cfg.tau, h->seed, noise_counter(h), s);
  layer_fwd(h, e.yheads, nullptr, s);
  gm_add(bp(h, e.YB0E1), bp(h, lo_buf(h, e.YB0E1)), (const float*)bp(h, e.YHEADS), bp(h, e.HSUM), bp(h, lo_buf(h, e.HSUM)), h->act_dt, h->B, s);
  layer_fwd(h, e.zheads, nullptr, s);
  h->launches += 2;
}

void gm_encoder_bwd(sv_handle* h, const float* inputs, int part, cudaStream_t s) {
  const GmEnc& e = h->gm_enc;
  const float inv_batch = 1.f / ((float)h->B * (float)h->cfg.world_size);
  if (part == 1) {
    layer_bwd(h, e.h2, nullptr, s);
    layer_bwd(h, e.h1, inputs, s);
    return;
  }
  layer_bwd(h, e.zheads, nullptr, s);
  gm_glue_a(bp(h, e.dHSUM), bp(h, e.YB0E1), (const float*)bp(h, e.YHEADS), (const float*)bp(h, h->ZM_G),
            (const float*)bp(h, h->ZS_G), bp(h, e.dYB0E1), bp(h, e.dYHEADS), h->act_dt, h->B, h->cfg.beta, inv_batch, s);
  layer_bwd(h, e.yheads, nullptr, s);
  gm_glue_b(bp(h, e.dY), (const float*)bp(h, e.Y), (const float*)bp(h, e.LOGITS), bp(h, e.dLOGITS), h->act_dt, h->B,
            h->K, h->cfg.tau, h->cfg.alpha, inv_batch, s);
  layer_bwd(h, e.ydense, nullptr, s);
  layer_bwd(h, e.yb2, nullptr, s);   // writes columns 0..1023 of dYB0E1; glue A wrote 1024..1535
  layer_bwd(h, e.yb0e1, nullptr, s);
  layer_bwd(h, e.h3, nullptr, s);
  h->launches += 2;
}

void decoder_fwd(sv_handle* h, const Decoder& d, cudaStream_t s) {
  const int B = h->B, H = h->H, W = h->W, T = h->act_dt;
  auto up = [&](int src, int dst, int hh, int ww, int c) {
    if (h->split) upsample2x_fwd_pair(bp(h, src), bp(h, lo_buf(h, src)), bp(h, dst), bp(h, lo_buf(h, dst)), B, hh, ww, c, s);
    else upsample2x_fwd(bp(h, src), bp(h, dst), T, B, hh, ww, c, s);
  };
  layer_fwd(h, d.d1, nullptr, s);
  layer_fwd(h, d.d2, nullptr, s);
  up(d.D2, d.U1, H / 8, W / 8, 128);
  layer_fwd(h, d.d3, nullptr, s);
  up(d.D3, d.U2, H / 4, W / 4, 64);
  layer_fwd(h, d.d4, nullptr, s);
  up(d.D4, d.U3, H / 2, W / 2, 32);
  layer_fwd(h, d.d5, nullptr, s);
  h->launches += 3;
}

void decoder_bwd(sv_handle* h, const Decoder& d, cudaStream_t s) {
  const int B = h->B, H = h->H, W = h->W, T = h->act_dt;
  float* const* fold = h->cs_on ? h->cs_fold[&d == &h->dec_x ? 0 : 1] : nullptr;   // bias-gradient partials of d2..d4 ride along
  layer_bwd(h, d.d5, nullptr, s);
  upsample2x_bwd(bp(h, d.dU3), bp(h, d.dD4), bp(h, d.D4), ACT_RELU, T, B, H / 2, W / 2, 32, s, fold ? fold[2] : nullptr);
  layer_bwd(h, d.d4, nullptr, s);
  upsample2x_bwd(bp(h, d.dU2), bp(h, d.dD3), bp(h, d.D3), ACT_RELU, T, B, H / 4, W / 4, 64, s, fold ? fold[1] : nullptr);
  layer_bwd(h, d.d3, nullptr, s);
  upsample2x_bwd(bp(h, d.dU1), bp(h, d.dD2), bp(h, d.D2), ACT_RELU, T, B, H / 8, W / 8, 128, s, fold ? fold[0] : nullptr);
  layer_bwd(h, d.d2, nullptr, s);
  layer_bwd(h, d.d1, nullptr, s);
  h->launches += 3;
}

__global__ void pack_z_kernel(const float* __restrict__ zg, const float* __restrict__ zl, void* zcat, bf16* zcat_lo, int dt, int B) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 128) return;
  const int b = idx >> 7, d = idx & 127;
  const float g = zg[idx], l = zl ? zl[idx] : 0.f;           // (zl == NULL: plain GMVAE, one latent)
  if (dt == DT_F32) { ((float*)zcat)[b * 256 + d] = g; ((float*)zcat)[b * 256 + 128 + d] = l; }
  else { ((bf16*)zcat)[b * 256 + d] = __float2bfloat16_rn(g); ((bf16*)zcat)[b * 256 + 128 + d] = __float2bfloat16_rn(l); }
  if (zcat_lo) {
    zcat_lo[b * 256 + d] = __float2bfloat16_rn(g - round_bf16(g));
    zcat_lo[b * 256 + 128 + d] = __float2bfloat16_rn(l - round_bf16(l));
  }
}
__global__ void pack_y_kernel(const float* __restrict__ y, void* yt, bf16* yt_lo, int dt, int B, int K) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 32) return;
  const int b = idx >> 5, k = idx & 31;
  const float v = k < K ? y[b * K + k] : 0.f;
  if (dt == DT_F32) ((float*)yt)[idx] = v; else ((bf16*)yt)[idx] = __float2bfloat16_rn(v);
  if (yt_lo) yt_lo[idx] = __float2bfloat16_rn(v - round_bf16(v));
}
// The Philox counter of the in-kernel noise (Sampling eps, gumbel u): its own device-side word, bumped once per forward pass, so
// forward-only engines (evaluation, model(x), encode, get_y) draw fresh noise on every call like tf.random does - the optimizer's
// iteration count, which only training advances, used to serve as the counter.
__global__ void bump_counter_kernel(unsigned long long* ctr) { pdl_enter(); *ctr += 1ull; }
__global__ void copy_cols_kernel(const float* __restrict__ src, int ld, int coff, float* __restrict__ dst, int B, int n) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * n) return;
  dst[idx] = src[(idx / n) * ld + coff + idx % n];
}

void stage_first_inputs(sv_handle* h, const float* inputs, cudaStream_t s) {
  if (!h->use_tc) return;
  for (int i = 0; i < 2; ++i)
    if (h->XP[i] >= 0) {
      tc_stage_first(inputs, bp(h, h->XP[i]), i ? 3 : 0, h->B, h->H, h->W, h->split, s);
      h->launches += 1;
    }
}

// side-stream branch: returns the stream the second branch should use (the main stream if two-stream mode is off)
cudaStream_t fork_side(sv_handle* h, cudaStream_t s) {
  if (!h->two_streams || !h->side) return s;
  cudaEventRecord(h->ev_fork, s);
  cudaStreamWaitEvent(h->side, h->ev_fork, 0);
  return h->side;
}
void join_side(sv_handle* h, cudaStream_t s) {
  if (!h->two_streams || !h->side) return;
  cudaEventRecord(h->ev_join, h->side);
  cudaStreamWaitEvent(s, h->ev_join, 0);
}

sv_status check_launch(sv_handle* h, const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, SV_ERR_DEVICE, "%s: %s", what, cudaGetErrorString(e));
  }
  return SV_OK;
}

#define REQUIRE_BOUND(h)                                                         \
  do {                                                                           \
    if (!(h)) return SV_ERR_INVALID;                                             \
    if ((h)->plan_only) return fail((h), SV_ERR_STATE, "handle is plan-only");   \
    if (!(h)->bound) return fail((h), SV_ERR_STATE, "sv_bind has not been called"); \
  } while (0)

}  // namespace

// =============================================================================================
extern "C" {

const char* sv_version(void) { return "splitvae-b200 0.1 (sm_100a)"; }

const char* sv_last_error(const sv_handle* h) { return h ? h->err : g_create_error; }

sv_status sv_create(const sv_config* cfg, sv_handle** out) {
  if (!cfg || !out) return fail(nullptr, SV_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->model != SV_MODEL_LGVAE && cfg->model != SV_MODEL_LGGMVAE && cfg->model != SV_MODEL_GMVAE)
    return fail(nullptr, SV_ERR_INVALID, "unknown model %d", cfg->model);
  if (cfg->height < 16 || cfg->width < 16 || cfg->height % 16 || cfg->width % 16 || cfg->height > 256 || cfg->width > 256)
    return fail(nullptr, SV_ERR_INVALID, "image size %dx%d unsupported (multiples of 16 in [16,256])", cfg->height, cfg->width);
  if (cfg->batch < 1 || cfg->batch > 65536) return fail(nullptr, SV_ERR_INVALID, "batch %d unsupported", cfg->batch);
  if (cfg->global_latent_dims != 128 || cfg->local_latent_dims != 128)
    return fail(nullptr, SV_ERR_INVALID, "latent dims must be 128/128 (got %d/%d)", cfg->global_latent_dims, cfg->local_latent_dims);
  if (cfg->model != SV_MODEL_LGVAE && (cfg->y_size < 2 || cfg->y_size > 32))
    return fail(nullptr, SV_ERR_INVALID, "y_size %d unsupported (2..32)", cfg->y_size);
  if (cfg->world_size < 1) return fail(nullptr, SV_ERR_INVALID, "world_size must be >= 1");
  if (cfg->precision != SV_PRECISION_BF16_TC && cfg->precision != SV_PRECISION_FP32_REF && cfg->precision != SV_PRECISION_BF16X3)
    return fail(nullptr, SV_ERR_INVALID, "unknown precision %d", cfg->precision);
  const bool plan_only = (cfg->flags & SV_FLAG_PLAN_ONLY) != 0;
  if (!plan_only) {
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
      cudaGetLastError();
      return fail(nullptr, SV_ERR_DEVICE, "no CUDA device available: libsplitvae has no CPU fallback");
    }
    if (prop.major != 10)
      return fail(nullptr, SV_ERR_DEVICE, "device '%s' is sm_%d%d; libsplitvae is built for sm_100a only", prop.name, prop.major, prop.minor);
  }

  sv_handle* h = new sv_handle();
  h->cfg = *cfg;
  h->plan_only = plan_only;
  h->act_dt = cfg->precision == SV_PRECISION_FP32_REF ? DT_F32 : DT_BF16;
  h->round_w = h->act_dt == DT_BF16;
  h->split = cfg->precision == SV_PRECISION_BF16X3;
  h->use_tc = (cfg->precision == SV_PRECISION_BF16_TC && !(cfg->flags & SV_FLAG_NO_TC)) || h->split;
  h->B = cfg->batch; h->H = cfg->height; h->W = cfg->width;
  h->F = ((cfg->height / 8) * cfg->width) / 8 * 128;  // vae/model.py:152 precedence
  h->gm = cfg->model != SV_MODEL_LGVAE;
  h->has_local = cfg->model != SV_MODEL_GMVAE;
  h->K = h->gm ? cfg->y_size : 0;
  h->seed = 0x5EEDull + 0x9E3779B97F4A7C15ull * (unsigned long long)(unsigned)(cfg->flags >> 8);
  const int B = h->B;
  const int dout_ld = h->act_dt == DT_F32 ? 6 : 16;

  // shared latent buffers first (decoders reference ZCAT)
  h->ZCAT = fwd_buf(h, (long long)B * 256);
  h->EPS_G = f32_buf(h, B * 128); h->EPS_L = f32_buf(h, B * 128);
  h->Z_G = f32_buf(h, B * 128); h->Z_L = f32_buf(h, B * 128);
  h->ZM_G = f32_buf(h, B * 128); h->ZS_G = f32_buf(h, B * 128);
  h->ZM_L = f32_buf(h, B * 128); h->ZS_L = f32_buf(h, B * 128);
  h->ZPM_OUT = f32_buf(h, B * 128); h->ZPS_OUT = f32_buf(h, B * 128);
  h->SCALARS = f32_buf(h, 64);
  h->PARTIALS = f32_buf(h, 2 * 148 * 8 + 64);
  h->KLPART = f32_buf(h, 2 * reparam_blocks(B) + 64);
  h->COLSUM = f32_buf(h, 256 * 8192);
  h->COLSUM2 = f32_buf(h, 256 * 8192);
  h->ADAM = new_buf(h, 1024);
  const int dzcat = act_buf(h, (long long)B * 256), dzl2 = act_buf(h, (long long)B * 128);

  // Keras variable order: encoder_x, encoder_x_hat, decoder_x, decoder_x_hat (model.py:182-186, 230-234)
  if (cfg->model == SV_MODEL_LGVAE) h->enc_x = build_conv_encoder(h, "encoder_x", 0);
  else h->gm_enc = build_gm_encoder(h, "encoder_x");
  if (h->has_local) h->enc_xh = build_conv_encoder(h, "encoder_x_hat", 3);
  h->seg_split = h->arena_floats;
  // (plain GMVAE: ONE decoder whose d1 reads z_x = columns 0..127 of the [B,256] latent buffer, vae/model.py:286,295)
  h->dec_x = build_decoder(h, "decoder_x", h->has_local ? 256 : 128, 0, dzcat, 256, dout_ld);
  if (h->has_local) h->dec_xh = build_decoder(h, "decoder_x_hat", 128, 128, dzl2, 128, dout_ld);

  if (h->act_dt == DT_BF16 && !getenv("SV_NO_MULTI_COLSUM")) {
    bool ok = true;
    for (int b = 0; b < 4 && ok; ++b)
      for (int part = 0; part < 2 && ok; ++part) {
        const std::vector<ColsumSpec> sp = branch_specs(h, b, part, false);
        ok = sp.empty() || colsum_multi_supported(sp.data(), (int)sp.size());
      }
    if (ok) {
      h->cs_on = true;
      for (int b = 0; b < 4; ++b)
        for (int part = 0; part < 2; ++part) {
          const std::vector<ColsumSpec> sp = branch_specs(h, b, part, false);
          if (!sp.empty()) h->CSP[b][part] = f32_buf(h, colsum_table_partial_floats(sp.data(), (int)sp.size()) + 64);
        }
    }
  }
  for (int seg = 0; seg < sv_handle::kSegs; ++seg) {     // arena ranges of each segment: its layers' variables, adjacent slots merged
    std::vector<std::pair<long long, long long>> r;
    auto slot = [](long long n) { return (n + 63) / 64 * 64; };
    for (int li : segment_layers(h, seg)) {
      const ConvGeom& g = h->layers[li].g;
      for (int j = 0; j < g.nparts; ++j) {
        r.push_back({g.part_w[j], slot((long long)g.kh * g.kw * g.Ci * g.part_n[j])});
        r.push_back({g.part_b[j], slot(g.part_n[j])});
      }
    }
    std::sort(r.begin(), r.end());
    for (const auto& x : r) {
      auto& out = h->seg_ranges[seg];
      if (!out.empty() && out.back().first + out.back().second == x.first) out.back().second += x.second;
      else out.push_back(x);
    }
  }
  if (h->split) {     // every forward product but the decoders' last layer multiplies bf16 pairs
    h->lo_of.resize(h->bufs.size(), -1);
    for (size_t li = 0; li < h->layers.size(); ++li) {
      Layer& L = h->layers[li];
      L.split_fwd = (int)li != h->dec_x.d5 && !(h->has_local && (int)li == h->dec_xh.d5);
      if (L.split_fwd) { L.in_lo = lo_buf(h, L.in); L.out_lo = lo_buf(h, L.out); }
    }
  }
  if (h->use_tc) {
    for (auto& L : h->layers) {
      const bool first = L.in < 0;
      tc_plan_layer(L.tc, L.g, L.in_dt, L.out_dt, L.in >= 0, L.din >= 0, first, L.split_fwd);
      if (h->split && !L.tc.fwd_ok) {
        const sv_status st = fail(nullptr, SV_ERR_NOT_IMPLEMENTED, "bf16x3: no tensor-core forward plan for layer %s at this shape", L.name.c_str());
        delete h;
        return st;
      }
      if (first && (L.tc.fwd_ok || L.tc.wgrad_ok)) {
        if (h->XP[L.g.in_coff ? 1 : 0] < 0) h->XP[L.g.in_coff ? 1 : 0] = new_buf(h, tc_first_stage_bytes(h->B, h->H, h->W));
        L.xp = h->XP[L.g.in_coff ? 1 : 0];
      }
    }
    size_t tcws = 0;
    for (auto& L : h->layers) tcws += tc_workspace_bytes(L.tc, L.g);
    h->TCWS = new_buf(h, tcws + 1024);
    if (plan_only && getenv("SV_PACK_DEBUG") && atoi(getenv("SV_PACK_DEBUG")) == 2) {      // job list of the per-segment operand re-pack, host only
      for (int seg = 0; seg < sv_handle::kSegs; ++seg) {
        std::vector<TcLayer*> sl;
        std::vector<const ConvGeom*> sg;
        for (int li : segment_layers(h, seg)) { sl.push_back(&h->layers[li].tc); sg.push_back(&h->layers[li].g); }
        const char* perr = nullptr;
        fprintf(stderr, "[pack] segment %d\n", seg);
        tc_pack_table_destroy(tc_pack_table_create(sl.data(), sg.data(), (int)sl.size(), &perr));
      }
    }
  }
  *out = h;
  return SV_OK;
}

sv_status sv_destroy(sv_handle* h) {
  if (h) {
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    tc_pack_table_destroy(h->pack);
    for (int seg = 0; seg < sv_handle::kSegs; ++seg) tc_pack_table_destroy(h->pack_seg[seg]);
    if (h->ev_seg_done) cudaEventDestroy(h->ev_seg_done);
    if (h->ev_opt_fork) cudaEventDestroy(h->ev_opt_fork);
    if (h->ev_opt_join) cudaEventDestroy(h->ev_opt_join);
    if (h->opt) cudaStreamDestroy(h->opt);
    for (int b = 0; b < 4; ++b)
      for (int part = 0; part < 2; ++part) colsum_table_destroy(h->cs[b][part]);
    for (int k = 0; k < 2; ++k)
      for (int j = 0; j < sv_handle::kMaxAux; ++j) {
        if (h->ev_aux[k][j]) cudaEventDestroy(h->ev_aux[k][j]);
        if (h->ev_aux_join[k][j]) cudaEventDestroy(h->ev_aux_join[k][j]);
        if (h->aux[k][j]) cudaStreamDestroy(h->aux[k][j]);
      }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->side) cudaStreamDestroy(h->side);
  }
  delete h;
  return SV_OK;
}

int32_t sv_param_count(const sv_handle* h) { return h ? (int32_t)h->vars.size() : 0; }

sv_status sv_param_describe(const sv_handle* h, int32_t i, sv_param_desc* out) {
  if (!h || !out || i < 0 || i >= (int32_t)h->vars.size()) return SV_ERR_INVALID;
  const Var& v = h->vars[i];
  memset(out, 0, sizeof(*out));
  snprintf(out->name, sizeof(out->name), "%s", v.name.c_str());
  out->ndim = v.ndim;
  for (int k = 0; k < 4; ++k) out->shape[k] = v.shape[k];
  out->offset = v.off;
  out->count = v.count;
  return SV_OK;
}

int64_t sv_arena_floats(const sv_handle* h) { return h ? h->arena_floats : 0; }
int64_t sv_workspace_bytes(const sv_handle* h) { return h ? (int64_t)h->ws_bytes : 0; }

sv_status sv_bind(sv_handle* h, float* params, float* grads, float* adam_m, float* adam_v, void* ws, int64_t ws_bytes) {
  if (!h) return SV_ERR_INVALID;
  if (h->plan_only) return fail(h, SV_ERR_STATE, "handle is plan-only");
  if (!params || !grads || !adam_m || !adam_v || !ws) return fail(h, SV_ERR_INVALID, "null buffer");
  if (ws_bytes < (int64_t)h->ws_bytes) return fail(h, SV_ERR_STATE, "workspace too small: %lld < %lld", (long long)ws_bytes, (long long)h->ws_bytes);
  if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)adam_m | (uintptr_t)adam_v) & 255) return fail(h, SV_ERR_INVALID, "arenas must be 256-byte aligned");
  if ((uintptr_t)ws & 1023) return fail(h, SV_ERR_INVALID, "workspace must be 1024-byte aligned");
  h->params = params; h->grads = grads; h->adam_m = adam_m; h->adam_v = adam_v; h->ws = (char*)ws;
  if (h->use_tc) {
    char* tcws = (char*)bp(h, h->TCWS);
    for (auto& L : h->layers) {
      const char* err = tc_bind_layer(L.tc, L.g, L.in >= 0 ? bp(h, L.in) : bp(h, L.xp), bp(h, L.out), bp(h, L.dout),
                                      L.din >= 0 ? bp(h, L.din) : nullptr, L.in >= 0 ? bp(h, L.in) : nullptr, L.mask_act, tcws,
                                      bp(h, L.in_lo), bp(h, L.out_lo));
      if (err) return fail(h, SV_ERR_DEVICE, "tensor-core plan for %s: %s", L.name.c_str(), err);
      tcws += tc_workspace_bytes(L.tc, L.g);
    }
    std::vector<TcLayer*> tl;
    std::vector<const ConvGeom*> tg;
    for (auto& L : h->layers) { tl.push_back(&L.tc); tg.push_back(&L.g); }
    const char* perr = nullptr;
    tc_pack_table_destroy(h->pack);
    h->pack = tc_pack_table_create(tl.data(), tg.data(), (int)tl.size(), &perr);
    if (!h->pack) return fail(h, SV_ERR_DEVICE, "tensor-core pack table: %s", perr ? perr : "?");
    for (int seg = 0; seg < sv_handle::kSegs; ++seg) {
      std::vector<TcLayer*> sl;
      std::vector<const ConvGeom*> sg;
      for (int li : segment_layers(h, seg)) { sl.push_back(&h->layers[li].tc); sg.push_back(&h->layers[li].g); }
      tc_pack_table_destroy(h->pack_seg[seg]);
      h->pack_seg[seg] = tc_pack_table_create(sl.data(), sg.data(), (int)sl.size(), &perr);
      if (!h->pack_seg[seg]) return fail(h, SV_ERR_DEVICE, "tensor-core pack table: %s", perr ? perr : "?");
    }
  }
  if (h->cs_on) {
    for (int b = 0; b < 4; ++b)
      for (int part = 0; part < 2; ++part) {
        int fold_index[4];
        const std::vector<ColsumSpec> sp = branch_specs(h, b, part, true, fold_index);
        if (sp.empty()) continue;
        const char* cerr = nullptr;
        colsum_table_destroy(h->cs[b][part]);
        h->cs[b][part] = colsum_table_create(sp.data(), (int)sp.size(), (float*)bp(h, h->CSP[b][part]), &cerr);
        if (!h->cs[b][part]) return fail(h, SV_ERR_DEVICE, "bias-gradient table: %s", cerr ? cerr : "?");
        if (b < 2 && part == 0)
          for (int k = 0; k < 4; ++k) h->cs_fold[b][k] = colsum_table_ext_partial(h->cs[b][part], fold_index[k]);
      }
  }
  // Stream priorities (only the order in which PENDING blocks are dispatched; nothing is pre-empted): the forward / dgrad chain
  // (caller's stream + side stream) goes first, the weight-gradient streams fill the SMs the chain leaves, and the optimizer stream
  // (Adam + operand re-pack: thousands of short blocks that otherwise sit in front of the chain's next kernel) comes last.  The
  // caller's stream keeps its own priority: trainer.StepRunner and bench.py capture on a highest-priority stream.
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  const char* pv = getenv("SV_STREAM_PRIO");
  const bool use_prio = !(pv && *pv == '0');
  // (-3 = the highest level torch.cuda.Stream(priority=...) hands out, so the caller's capture stream and the side stream are equals)
  const int prio_side = use_prio ? (prio_greatest > -3 ? prio_greatest : -3) : 0;
  int prio_aux = use_prio ? prio_side + 1 : 0;
  if (const char* av = getenv("SV_AUX_PRIO")) if (*av) prio_aux = atoi(av);
  if (prio_aux > prio_least) prio_aux = prio_least;
  if (prio_aux < prio_greatest) prio_aux = prio_greatest;
  // optimizer stream: level with the weight-gradient streams.  (At the lowest level the decoders' Adam starved behind every other
  // kernel until the end of the step and the whole optimizer tail - 90 us - was exposed after the last weight gradient.)
  int prio_opt = use_prio ? prio_aux : 0;
  if (const char* ov = getenv("SV_OPT_PRIO")) if (*ov) prio_opt = atoi(ov);
  if (prio_opt > prio_least) prio_opt = prio_least;
  if (prio_opt < prio_greatest) prio_opt = prio_greatest;
  if (!h->side) {
    const char* one = getenv("SV_ONE_STREAM");
    h->two_streams = !(one && *one == '1');
    if (h->two_streams &&
        (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_side) != cudaSuccess ||
         cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
         cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess))
      return fail(h, SV_ERR_DEVICE, "side stream / event creation failed");
  }
  if (!h->opt) {
    const char* off = getenv("SV_OPT_STREAM");
    if (!(off && *off == '0') &&
        (cudaStreamCreateWithPriority(&h->opt, cudaStreamNonBlocking, prio_opt) != cudaSuccess ||
         cudaEventCreateWithFlags(&h->ev_opt_fork, cudaEventDisableTiming) != cudaSuccess ||
         cudaEventCreateWithFlags(&h->ev_seg_done, cudaEventDisableTiming) != cudaSuccess ||
         cudaEventCreateWithFlags(&h->ev_opt_join, cudaEventDisableTiming) != cudaSuccess))
      return fail(h, SV_ERR_DEVICE, "optimizer stream / event creation failed");
  }
  if (!h->aux[0][0]) {
    const char* off = getenv("SV_WGRAD_STREAMS");
    h->aux_n = off && *off ? atoi(off) : 3;      // (measured on the C2 step: 2 -> 1.698, 3 -> 1.695, 4 = 2, 6 slower)
    if (h->aux_n > sv_handle::kMaxAux) h->aux_n = sv_handle::kMaxAux;
    h->wgrad_streams = h->aux_n > 0;
    for (int k = 0; k < 2; ++k)
      for (int j = 0; j < h->aux_n; ++j)
        if (cudaStreamCreateWithPriority(&h->aux[k][j], cudaStreamNonBlocking, prio_aux) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_aux[k][j], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_aux_join[k][j], cudaEventDisableTiming) != cudaSuccess)
          return fail(h, SV_ERR_DEVICE, "auxiliary stream / event creation failed");
  }
  h->bound = true;
  return SV_OK;
}

sv_status sv_params_updated(sv_handle* h, void* stream) {
  REQUIRE_BOUND(h);
  if (h->use_tc) {
    h->launches += tc_repack_all(h->pack, h->params, (cudaStream_t)stream);
  }
  return check_launch(h, "sv_params_updated");
}

static sv_status forward_impl(sv_handle* h, const float* inputs, const float* eps_g, const float* eps_l, const float* u,
                              void* stream, bool prior_copies) {
  REQUIRE_BOUND(h);
  if (!inputs) return fail(h, SV_ERR_INVALID, "inputs is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const bool gm = h->gm, loc = h->has_local;
  stage_first_inputs(h, inputs, s);
  cudaStream_t s2 = loc ? fork_side(h, s) : s;
  if (gm) gm_encoder_fwd(h, inputs, u, s); else conv_encoder_fwd(h, h->enc_x, inputs, s);
  if (loc) { conv_encoder_fwd(h, h->enc_xh, inputs, s2); join_side(h, s); }
  reparam(latent_bufs(h), h->B, h->act_dt, eps_g, eps_l, h->seed, noise_counter(h), (float*)bp(h, h->KLPART), s);
  h->launches += 1;
  s2 = loc ? fork_side(h, s) : s;
  decoder_fwd(h, h->dec_x, s);
  if (loc) { decoder_fwd(h, h->dec_xh, s2); join_side(h, s); }
  launch_pdl(bump_counter_kernel, dim3(1), dim3(1), 0, s, noise_counter(h));
  h->launches += 1;
  if (gm && prior_copies) {  // contiguous z_prior_mean / z_prior_sig for the reference's output tuple
    copy_cols_kernel<<<(h->B * 128 + 255) / 256, 256, 0, s>>>((const float*)bp(h, h->gm_enc.YHEADS), 768, 512, (float*)bp(h, h->ZPM_OUT), h->B, 128);
    copy_cols_kernel<<<(h->B * 128 + 255) / 256, 256, 0, s>>>((const float*)bp(h, h->gm_enc.YHEADS), 768, 640, (float*)bp(h, h->ZPS_OUT), h->B, 128);
    h->launches += 2;
  }
  return check_launch(h, "sv_forward");
}

sv_status sv_forward(sv_handle* h, const float* inputs, const float* eps_g, const float* eps_l, const float* u, void* stream) {
  return forward_impl(h, inputs, eps_g, eps_l, u, stream, true);
}

static void run_pixel_loss(sv_handle* h, const float* inputs, cudaStream_t s) {
  const bool loc = h->has_local;
  const long long npix = (long long)h->B * h->H * h->W;
  const float inv_batch = 1.f / ((float)h->B * (float)h->cfg.world_size);
  pixel_loss(inputs, (const float*)bp(h, h->dec_x.OUT), loc ? (const float*)bp(h, h->dec_xh.OUT) : nullptr, bp(h, h->dec_x.dOUT),
             loc ? bp(h, h->dec_xh.dOUT) : nullptr, h->act_dt, h->layers[h->dec_x.d5].g.dout_ld, npix, inv_batch,
             (float*)bp(h, h->PARTIALS), h->act_dt == DT_BF16, s, h->cs_on ? h->cs_fold[0][3] : nullptr,
             h->cs_on && loc ? h->cs_fold[1][3] : nullptr);
}

sv_status sv_debug_pixel_loss(sv_handle* h, const float* inputs, void* stream) {
  REQUIRE_BOUND(h);
  if (!inputs) return fail(h, SV_ERR_INVALID, "inputs is NULL");
  run_pixel_loss(h, inputs, (cudaStream_t)stream);
  return check_launch(h, "sv_debug_pixel_loss");
}

sv_status sv_loss_fwd_bwd(sv_handle* h, const float* inputs, void* stream) {
  REQUIRE_BOUND(h);
  if (!inputs) return fail(h, SV_ERR_INVALID, "inputs is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const bool gm = h->gm;
  const long long npix = (long long)h->B * h->H * h->W;
  run_pixel_loss(h, inputs, s);
  loss_scalars((const float*)bp(h, h->KLPART), reparam_blocks(h->B), gm ? (const float*)bp(h, h->gm_enc.LOGITS) : nullptr, h->B, h->K, gm, h->cfg.beta, h->cfg.alpha,
               (const float*)bp(h, h->PARTIALS), pixel_loss_blocks(npix), (float*)bp(h, h->SCALARS), s);
  h->launches += 2;
  h->last_inputs = inputs;
  return check_launch(h, "sv_loss_fwd_bwd");
}

int32_t sv_num_segments(const sv_handle* h) { return h ? sv_handle::kSegs : 0; }

int32_t sv_segment_num_ranges(const sv_handle* h, int32_t seg) {
  return (h && seg >= 0 && seg < sv_handle::kSegs) ? (int32_t)h->seg_ranges[seg].size() : 0;
}

sv_status sv_segment_range(const sv_handle* h, int32_t seg, int32_t index, int64_t* off, int64_t* cnt) {
  if (!h || !off || !cnt || seg < 0 || seg >= sv_handle::kSegs || index < 0 || index >= (int32_t)h->seg_ranges[seg].size()) return SV_ERR_INVALID;
  *off = h->seg_ranges[seg][index].first;
  *cnt = h->seg_ranges[seg][index].second;
  return SV_OK;
}

static sv_status backward_segment_impl(sv_handle* h, int32_t seg, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const float* inputs = h->last_inputs;
  const bool loc = h->has_local;
  if (seg == 0) {
    cudaStream_t s2 = loc ? fork_side(h, s) : s;
    decoder_bwd(h, h->dec_x, s);
    branch_bias_grads(h, 0, 0, s);
    if (loc) {
      decoder_bwd(h, h->dec_xh, s2);
      branch_bias_grads(h, 1, 0, s2);
      join_side(h, s);
    }
  } else if (seg == 1 || seg == 2) {
    if (!inputs) return fail(h, SV_ERR_STATE, "sv_loss_fwd_bwd must precede sv_backward_segment");
    const bool gm = h->gm;
    const int part = seg - 1;          // 0: heads .. e3 (gradients final first), 1: e2, e1
    if (part == 0) {
      const float inv_batch = 1.f / ((float)h->B * (float)h->cfg.world_size);
      latent_bwd(latent_bufs(h), h->B, h->act_dt, gm, h->cfg.beta, inv_batch, s);
      h->launches += 1;
    }
    cudaStream_t s2 = loc ? fork_side(h, s) : s;
    if (loc) {
      conv_encoder_bwd(h, h->enc_xh, inputs, part, s2);
      branch_bias_grads(h, 3, part, s2);
    }
    if (gm) gm_encoder_bwd(h, inputs, part, s); else conv_encoder_bwd(h, h->enc_x, inputs, part, s);
    branch_bias_grads(h, 2, part, s);
    if (loc) join_side(h, s);
  } else {
    return fail(h, SV_ERR_INVALID, "segment %d out of range", seg);
  }
  return check_launch(h, "sv_backward_segment");
}

sv_status sv_backward_segment(sv_handle* h, int32_t seg, void* stream) {
  REQUIRE_BOUND(h);
  h->defer_join = false;
  const sv_status st = backward_segment_impl(h, seg, stream);
  if (st == SV_OK) join_all_wgrad_streams(h, (cudaStream_t)stream);   // (streams a deferred call left unjoined)
  return st;
}

sv_status sv_backward_segment_deferred(sv_handle* h, int32_t seg, void* stream, void* done_stream) {
  REQUIRE_BOUND(h);
  if (!done_stream || done_stream == stream || getenv_off("SV_DEFER_JOIN")) return sv_backward_segment(h, seg, stream);
  h->defer_join = true;
  const sv_status st = backward_segment_impl(h, seg, stream);
  h->defer_join = false;
  if (st != SV_OK) return st;
  cudaStream_t s = (cudaStream_t)stream, d = (cudaStream_t)done_stream;
  cudaEventRecord(h->ev_seg_done, s);               // the chain's own part: dgrads (readers of the packed weights), glue kernels
  cudaStreamWaitEvent(d, h->ev_seg_done, 0);
  join_all_wgrad_streams(h, d);
  return check_launch(h, "sv_backward_segment_deferred");
}

sv_status sv_adam_segment(sv_handle* h, int32_t seg, void* stream) {
  REQUIRE_BOUND(h);
  if (seg < 0 || seg >= sv_handle::kSegs) return fail(h, SV_ERR_INVALID, "segment %d out of range", seg);
  cudaStream_t s = (cudaStream_t)stream;
  AdamState* st = (AdamState*)bp(h, h->ADAM);
  if (seg == 0) {
    adam_prepare(st, h->cfg.learning_rate, h->gm, s);     // staircase schedule for both GM models (vae/main.py:66-72)
    h->launches += 1;
  }
  for (const auto& r : h->seg_ranges[seg]) {
    adam_apply(h->params + r.first, h->grads + r.first, h->adam_m + r.first, h->adam_v + r.first, r.second, st, 0.f, s);
    h->launches += 1;
  }
  if (h->use_tc) h->launches += tc_repack_all(h->pack_seg[seg], h->params, s);
  return check_launch(h, "sv_adam_segment");
}

sv_status sv_nvls_adam_segment(sv_handle* h, int32_t seg, const float* mc_grads, float* mc_params, int32_t rank, int32_t world,
                               int32_t write_reduced_grads, void* stream) {
  REQUIRE_BOUND(h);
  if (seg < 0 || seg >= sv_handle::kSegs) return fail(h, SV_ERR_INVALID, "segment %d out of range", seg);
  if (!mc_grads || !mc_params || world < 1 || rank < 0 || rank >= world) return fail(h, SV_ERR_INVALID, "bad multicast pointers / rank");
  if (world != h->cfg.world_size) return fail(h, SV_ERR_INVALID, "world %d differs from the handle's world_size %d (gradient scale)", world, h->cfg.world_size);
  cudaStream_t s = (cudaStream_t)stream;
  AdamState* st = (AdamState*)bp(h, h->ADAM);
  if (seg == 0) {
    adam_prepare(st, h->cfg.learning_rate, h->gm, s);
    h->launches += 1;
  }
  for (const auto& r : h->seg_ranges[seg]) {
    nvls_adam(mc_grads, mc_params, h->params, h->adam_m, h->adam_v, write_reduced_grads ? const_cast<float*>(mc_grads) : nullptr, r.first, r.second,
              rank, world, st, s);
    h->launches += 1;
  }
  return check_launch(h, "sv_nvls_adam_segment");
}

sv_status sv_repack_segment(sv_handle* h, int32_t seg, void* stream) {
  REQUIRE_BOUND(h);
  if (seg < 0 || seg >= sv_handle::kSegs) return fail(h, SV_ERR_INVALID, "segment %d out of range", seg);
  if (h->use_tc) h->launches += tc_repack_all(h->pack_seg[seg], h->params, (cudaStream_t)stream);
  return check_launch(h, "sv_repack_segment");
}

sv_status sv_adam_step(sv_handle* h, void* stream) {
  for (int seg = 0; seg < sv_handle::kSegs; ++seg) {
    const sv_status st = sv_adam_segment(h, seg, stream);
    if (st) return st;
  }
  return SV_OK;
}

sv_status sv_train_step(sv_handle* h, const float* inputs, const float* eps_g, const float* eps_l, const float* u, void* stream) {
  sv_status st = forward_impl(h, inputs, eps_g, eps_l, u, stream, false);
  if (st) return st;
  if ((st = sv_loss_fwd_bwd(h, inputs, stream))) return st;
  cudaStream_t s = (cudaStream_t)stream;
  // segment k's gradients are final after its backward: Adam + re-pack of segment k run on the optimizer stream while the
  // backward of segment k+1 runs; only the last (smallest) segment's update is exposed
  for (int seg = 0; seg < sv_handle::kSegs; ++seg) {
    if (h->opt && seg + 1 < sv_handle::kSegs) {
      if (seg == 0) {                                  // (brings the optimizer stream into a capture before it is a join target)
        cudaEventRecord(h->ev_opt_fork, s);
        cudaStreamWaitEvent(h->opt, h->ev_opt_fork, 0);
      }
      if ((st = sv_backward_segment_deferred(h, seg, stream, h->opt))) return st;
      if ((st = sv_adam_segment(h, seg, h->opt))) return st;
      if (seg + 2 == sv_handle::kSegs) cudaEventRecord(h->ev_opt_join, h->opt);
    } else {
      if ((st = sv_backward_segment(h, seg, stream))) return st;
      if (h->opt) cudaStreamWaitEvent(s, h->ev_opt_join, 0);
      if ((st = sv_adam_segment(h, seg, stream))) return st;
    }
  }
  return SV_OK;
}

sv_status sv_capture_graph(sv_handle* h, const float* inputs, const float* eps_g, const float* eps_l, const float* u, void* stream) {
  REQUIRE_BOUND(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (!s) return fail(h, SV_ERR_INVALID, "sv_capture_graph needs a non-default stream");
  if (h->graph_exec) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
  if (h->graph) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
  if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return fail(h, SV_ERR_DEVICE, "cudaStreamBeginCapture failed");
  }
  const long long before = h->launches;
  const sv_status st = sv_train_step(h, inputs, eps_g, eps_l, u, stream);
  h->graph_kernels = h->launches - before;                 // kernels of one captured step (counted by sv_replay)
  h->launches = before;                                    // (capture records, nothing ran)
  cudaGraph_t g = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(s, &g);
  if (st != SV_OK) { if (g) cudaGraphDestroy(g); return st; }
  if (ce != cudaSuccess || !g) { cudaGetLastError(); return fail(h, SV_ERR_DEVICE, "cudaStreamEndCapture failed: %s", cudaGetErrorString(ce)); }
  h->graph = g;
  if (cudaGraphInstantiate(&h->graph_exec, g, 0) != cudaSuccess) {
    cudaGetLastError();
    return fail(h, SV_ERR_DEVICE, "cudaGraphInstantiate failed");
  }
  return SV_OK;
}

sv_status sv_replay(sv_handle* h, int32_t n_steps, void* stream) {
  REQUIRE_BOUND(h);
  if (!h->graph_exec) return fail(h, SV_ERR_STATE, "sv_capture_graph has not been called");
  if (n_steps < 0) return fail(h, SV_ERR_INVALID, "n_steps < 0");
  for (int i = 0; i < n_steps; ++i)
    if (cudaGraphLaunch(h->graph_exec, (cudaStream_t)stream) != cudaSuccess) {
      cudaGetLastError();
      return fail(h, SV_ERR_DEVICE, "cudaGraphLaunch failed");
    }
  h->launches += (long long)n_steps * h->graph_kernels;
  return SV_OK;
}

sv_status sv_output_ptr(const sv_handle* hc, int32_t which, void** ptr, int64_t* count) {
  sv_handle* h = const_cast<sv_handle*>(hc);
  if (!h || !ptr || !count) return SV_ERR_INVALID;
  if (!h->bound) return fail(h, SV_ERR_STATE, "sv_bind has not been called");
  const bool gm = h->gm, loc = h->has_local;
  const long long B = h->B;
  int id = -1;
  long long n = 0;
  switch (which) {
    case SV_OUT_DEC_X: id = h->dec_x.OUT; n = B * h->H * h->W * 6; break;
    case SV_OUT_DEC_X_HAT: if (loc) { id = h->dec_xh.OUT; n = B * h->H * h->W * 6; } break;
    case SV_OUT_Z_X: id = h->Z_G; n = B * 128; break;
    case SV_OUT_Z_MEAN_X: id = h->ZM_G; n = B * 128; break;
    case SV_OUT_Z_SIG_X: id = h->ZS_G; n = B * 128; break;
    case SV_OUT_Z_X_HAT: if (loc) { id = h->Z_L; n = B * 128; } break;
    case SV_OUT_Z_MEAN_X_HAT: if (loc) { id = h->ZM_L; n = B * 128; } break;
    case SV_OUT_Z_SIG_X_HAT: if (loc) { id = h->ZS_L; n = B * 128; } break;
    case SV_OUT_Y: if (gm) { id = h->gm_enc.Y; n = B * 32; } break;
    case SV_OUT_Y_LOGITS: if (gm) { id = h->gm_enc.LOGITS; n = B * 32; } break;
    case SV_OUT_Z_PRIOR_MEAN: if (gm) { id = h->ZPM_OUT; n = B * 128; } break;
    case SV_OUT_Z_PRIOR_SIG: if (gm) { id = h->ZPS_OUT; n = B * 128; } break;
    case SV_OUT_SCALARS: id = h->SCALARS; n = SV_SCALAR_COUNT; break;
    case SV_OUT_SCALAR_SUMS: id = h->SCALARS; n = 16; break;          // (offset applied below)
    default: break;
  }
  if (id < 0) return fail(h, SV_ERR_INVALID, "output %d not available for this model", which);
  *ptr = bp(h, id);
  if (which == SV_OUT_SCALAR_SUMS) *ptr = (float*)*ptr + 8;
  *count = n;
  return SV_OK;
}

sv_status sv_decode(sv_handle* h, const float* z_x, const float* z_x_hat, void* stream) {
  REQUIRE_BOUND(h);
  if (!z_x || (!z_x_hat && h->has_local)) return fail(h, SV_ERR_INVALID, "null latent");
  cudaStream_t s = (cudaStream_t)stream;
  pack_z_kernel<<<(h->B * 128 + 255) / 256, 256, 0, s>>>(z_x, z_x_hat, bp(h, h->ZCAT), (bf16*)bp(h, lo_buf(h, h->ZCAT)), h->act_dt, h->B);
  h->launches += 1;
  if (h->has_local) {
    cudaStream_t s2 = fork_side(h, s);
    decoder_fwd(h, h->dec_x, s);
    decoder_fwd(h, h->dec_xh, s2);
    join_side(h, s);
  } else decoder_fwd(h, h->dec_x, s);
  return check_launch(h, "sv_decode");
}

sv_status sv_encode_y(sv_handle* h, const float* y, void* stream) {
  REQUIRE_BOUND(h);
  if (!h->gm) return fail(h, SV_ERR_INVALID, "encode_y needs a gmvae-type encoder (lggmvae / gmvae)");
  if (!y) return fail(h, SV_ERR_INVALID, "null y");
  cudaStream_t s = (cudaStream_t)stream;
  pack_y_kernel<<<(h->B * 32 + 255) / 256, 256, 0, s>>>(y, bp(h, h->gm_enc.YT), (bf16*)bp(h, lo_buf(h, h->gm_enc.YT)), h->act_dt, h->B, h->K);
  layer_fwd(h, h->gm_enc.yheads, nullptr, s);
  copy_cols_kernel<<<(h->B * 128 + 255) / 256, 256, 0, s>>>((const float*)bp(h, h->gm_enc.YHEADS), 768, 512, (float*)bp(h, h->ZPM_OUT), h->B, 128);
  copy_cols_kernel<<<(h->B * 128 + 255) / 256, 256, 0, s>>>((const float*)bp(h, h->gm_enc.YHEADS), 768, 640, (float*)bp(h, h->ZPS_OUT), h->B, 128);
  h->launches += 3;
  return check_launch(h, "sv_encode_y");
}

// Both are ordered on the caller's stream (the legacy default stream is not ordered against non-blocking streams) and return after
// that stream has drained: a read sees every step queued before it, a write cannot race an in-flight adam_prepare.
sv_status sv_get_iterations(sv_handle* h, int64_t* it, void* stream) {
  REQUIRE_BOUND(h);
  if (!it) return SV_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long v = 0;
  if (cudaMemcpyAsync(&v, bp(h, h->ADAM), 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return fail(h, SV_ERR_DEVICE, "memcpy failed");
  *it = (int64_t)v;
  return SV_OK;
}

sv_status sv_set_iterations(sv_handle* h, int64_t it, void* stream) {
  REQUIRE_BOUND(h);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long v = (unsigned long long)it;
  if (cudaMemcpyAsync(bp(h, h->ADAM), &v, 8, cudaMemcpyHostToDevice, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
    return fail(h, SV_ERR_DEVICE, "memcpy failed");
  return SV_OK;
}

int64_t sv_launch_count(const sv_handle* h) { return h ? h->launches : 0; }

sv_status sv_discretised_logistic_loss(const float* x, const float* m, const float* ls, float* out, int64_t n, void* stream) {
  if (!x || !m || !ls || !out || n < 0) return SV_ERR_INVALID;
  dll_elementwise(x, m, ls, out, n, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? SV_OK : SV_ERR_DEVICE;
}

sv_status sv_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, float alpha, void* stream) {
  if (!p || !g || !m || !v || n < 0) return SV_ERR_INVALID;
  adam_apply(p, g, m, v, n, nullptr, alpha, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? SV_OK : SV_ERR_DEVICE;
}

int32_t sv_debug_layer_count(const sv_handle* h) { return h ? (int32_t)h->layers.size() : 0; }

sv_status sv_debug_layer_info(const sv_handle* hc, int32_t i, sv_layer_info* o) {
  sv_handle* h = const_cast<sv_handle*>(hc);
  if (!h || !o || i < 0 || i >= (int32_t)h->layers.size()) return SV_ERR_INVALID;
  const Layer& L = h->layers[i];
  const ConvGeom& g = L.g;
  memset(o, 0, sizeof(*o));
  snprintf(o->name, sizeof(o->name), "%s", L.name.c_str());
  o->kh = g.kh; o->kw = g.kw; o->stride = g.stride; o->Hi = g.Hi; o->Wi = g.Wi; o->Ci = g.Ci; o->Ho = g.Ho; o->Wo = g.Wo; o->Co = g.Co;
  o->in_ld = g.in_ld; o->in_coff = g.in_coff; o->out_ld = g.out_ld; o->dout_ld = g.dout_ld; o->din_ld = g.din_ld;
  o->in_dt = L.in_dt; o->out_dt = L.out_dt; o->act_dt = h->act_dt;
  o->has_dgrad = L.din >= 0; o->tc_fwd = L.tc.fwd_ok; o->tc_dgrad = L.tc.dgrad_ok; o->tc_wgrad = L.tc.wgrad_ok;
  if (h->bound) { o->in = bp(h, L.in); o->out = bp(h, L.out); o->dout = bp(h, L.dout); o->din = bp(h, L.din); o->in_lo = bp(h, L.in_lo); o->out_lo = bp(h, L.out_lo); }
  o->split_fwd = L.split_fwd;
  if (h->use_tc) {
    o->kern_fwd = !L.tc.fwd_ok ? SV_KERN_NONE : L.tc.fwd_ns ? SV_KERN_NSCONV : L.tc.fwd.halo ? (L.tc.fwd.persist ? SV_KERN_PCONV : SV_KERN_HALO_CONV) : SV_KERN_IGEMM;
    o->kern_dgrad = !L.tc.dgrad_ok ? SV_KERN_NONE : L.tc.dgrad_ns ? SV_KERN_NSCONV : L.tc.dgrad[0].halo ? (L.tc.dgrad[0].persist ? SV_KERN_PCONV : SV_KERN_HALO_CONV) : SV_KERN_IGEMM;
    o->kern_wgrad = !L.tc.wgrad_ok ? SV_KERN_NONE : L.tc.wg_halo ? SV_KERN_HALO_WGRAD : SV_KERN_WGRAD;
    if (L.tc.wgrad_ok) {
      const TcWgradLaunch& W = L.tc.wg;
      o->wgrad_ctas = L.tc.wg_halo ? L.tc.hw.m_splits * L.tc.hw.k_splits
                                   : ((W.groups + W.groups_per_cta - 1) / W.groups_per_cta) * W.n_tiles * W.k_splits;
    }
  }
  o->in_elems = (int64_t)g.B * g.Hi * g.Wi * g.in_ld; o->out_elems = (int64_t)g.B * g.Ho * g.Wo * g.out_ld;
  o->dout_elems = (int64_t)g.B * g.Ho * g.Wo * g.dout_ld; o->din_elems = (int64_t)g.B * g.Hi * g.Wi * g.din_ld;
  return SV_OK;
}

sv_status sv_debug_run_layer(sv_handle* h, int32_t i, int32_t pass, int32_t impl, const float* inputs, void* stream) {
  REQUIRE_BOUND(h);
  if (i < 0 || i >= (int32_t)h->layers.size()) return fail(h, SV_ERR_INVALID, "layer %d out of range", i);
  Layer& L = h->layers[i];
  cudaStream_t s = (cudaStream_t)stream;
  const void* in = L.in < 0 ? (const void*)inputs : bp(h, L.in);
  if (L.in < 0 && !inputs && pass != SV_PASS_DGRAD) return fail(h, SV_ERR_INVALID, "layer %d reads the external inputs", i);
  const int T = h->act_dt;
  if (impl == SV_IMPL_TC) {
    if (L.in < 0 && inputs) stage_first_inputs(h, inputs, s);
    if (pass == SV_PASS_FWD) { if (!L.tc.fwd_ok) return fail(h, SV_ERR_NOT_IMPLEMENTED, "no tensor-core fwd for %s", L.name.c_str()); tc_conv_fwd(L.tc, s); }
    else if (pass == SV_PASS_DGRAD) { if (!L.tc.dgrad_ok) return fail(h, SV_ERR_NOT_IMPLEMENTED, "no tensor-core dgrad for %s", L.name.c_str()); tc_conv_dgrad(L.tc, s); }
    else { if (!L.tc.wgrad_ok) return fail(h, SV_ERR_NOT_IMPLEMENTED, "no tensor-core wgrad for %s", L.name.c_str()); tc_conv_wgrad(L.tc, L.g, h->grads, s); }
  } else {
    if (pass == SV_PASS_FWD) ref_conv_fwd(L.g, in, L.in_dt, h->params, bp(h, L.out), L.out_dt, h->round_w, s);
    else if (pass == SV_PASS_DGRAD) {
      if (L.din < 0) return fail(h, SV_ERR_INVALID, "layer %d has no dgrad", i);
      ref_conv_dgrad(L.g, bp(h, L.dout), T, h->params, bp(h, L.din), in, L.in_dt, L.mask_act, h->round_w, s);
    } else ref_conv_wgrad(L.g, in, L.in_dt, bp(h, L.dout), T, h->grads, h->round_w, s);
  }
  return check_launch(h, "sv_debug_run_layer");
}

int32_t sv_debug_halo_trace(uint64_t* out_host, int32_t max_ctas) {
  if (!out_host || max_ctas < 1) return -1;
  cudaDeviceSynchronize();
  return tc_halo_trace_read((unsigned long long*)out_host, max_ctas);
}

sv_status sv_stage_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int32_t B, int32_t H, int32_t W, int32_t p, void* stream) {
  if (!u8 || !perm || !inputs || B < 1 || p < 1 || H % p || W % p) return SV_ERR_INVALID;
  stage_scramble(u8, perm, inputs, B, H, W, p, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? SV_OK : SV_ERR_DEVICE;
}

sv_status sv_stage_resize_scramble(const uint8_t* u8, const int32_t* perm, float* inputs, int32_t B, int32_t Hs, int32_t Ws,
                                   int32_t crop_y, int32_t crop_x, int32_t crop_h, int32_t crop_w, int32_t H, int32_t W, int32_t p, void* stream) {
  if (!u8 || !perm || !inputs || B < 1 || p < 1 || H % p || W % p || crop_y < 0 || crop_x < 0 || crop_h < 1 || crop_w < 1 ||
      crop_y + crop_h > Hs || crop_x + crop_w > Ws)
    return SV_ERR_INVALID;
  stage_resize_scramble(u8, perm, inputs, B, Hs, Ws, crop_y, crop_x, crop_h, crop_w, H, W, p, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? SV_OK : SV_ERR_DEVICE;
}

sv_status sv_draw_permutations(int32_t* perm, int32_t B, int32_t n_patch, uint64_t seed, uint64_t step, void* stream) {
  if (!perm || B < 1 || n_patch < 1 || n_patch > 4096) return SV_ERR_INVALID;
  draw_permutations(perm, B, n_patch, seed, step, (cudaStream_t)stream);
  return cudaPeekAtLastError() == cudaSuccess ? SV_OK : SV_ERR_DEVICE;
}

}  // extern "C"
