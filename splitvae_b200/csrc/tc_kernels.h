// Tensor-core (tcgen05 + TMA) implicit-GEMM path: per-layer plan and launchers (internal header).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sv {

constexpr int kTcMaxMaps = 8;

struct TcGemmPlan {          // one implicit-GEMM launch configuration
  bool ok = false;
  int tile_n = 0, tile_h = 0, tile_w = 0;   // M tile = tile_n images x tile_h rows x tile_w cols = 128 pixels
  int bk = 0;                // K elements per pipeline stage (channels per TMA box)
  int swizzle = 0;           // bytes: 128 / 64 / 32
  int n_pad = 0;             // UMMA N (Cout padded)
  int n_tiles = 1;           // CTAs along N (Cout split)
  int stages = 0;
  int ci_pad = 0;            // packed input channels per tap
  int taps_h = 0, taps_w = 0, pad_t = 0, pad_l = 0;  // effective stride-1 conv on the A tensor
  int splits = 1;            // split-K
  CUtensorMap map_a[4];      // A operand maps (dgrad stride-2: one per output parity class)
  CUtensorMap map_b[4];
};

struct TcLayer {
  bool fwd_ok = false, dgrad_ok = false, wgrad_ok = false;
  int fwd_launches = 0, dgrad_launches = 0, wgrad_launches = 0;
  TcGemmPlan fwd, dgrad, wgrad;
  // packed bf16 weights inside the tensor-core workspace
  void* w_fwd = nullptr;     // [n_pad][taps][ci_pad]
  void* w_dgrad = nullptr;   // [classes][ci_pad_out][taps'][co_pad]
  size_t w_fwd_bytes = 0, w_dgrad_bytes = 0, partial_bytes = 0;
  float* partial = nullptr;  // split-K partial sums
  const void* in = nullptr; void* out = nullptr; void* dout = nullptr; void* din = nullptr;
  int in_dt = 0, out_dt = 0;
};

// plan (no CUDA calls), workspace size, bind (creates TMA descriptors; returns NULL or an error text)
void tc_plan_layer(TcLayer& t, const ConvGeom& g, int in_dt, int out_dt, bool has_internal_input, bool has_dgrad);
size_t tc_workspace_bytes(const TcLayer& t, const ConvGeom& g);
const char* tc_bind_layer(TcLayer& t, const ConvGeom& g, const void* in, void* out, void* dout, void* din, char* ws);
int tc_repack_weights(TcLayer& t, const ConvGeom& g, const float* params, cudaStream_t s);  // returns #launches

void tc_conv_fwd(TcLayer& t, const ConvGeom& g, const float* params, void* out, int out_dt, cudaStream_t s);
void tc_conv_dgrad(TcLayer& t, const ConvGeom& g, const void* mask_src, int mask_act, void* din, cudaStream_t s);
void tc_conv_wgrad(TcLayer& t, const ConvGeom& g, float* grads, cudaStream_t s);

}  // namespace sv
