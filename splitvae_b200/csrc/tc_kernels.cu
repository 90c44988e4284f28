// Tensor-core (tcgen05 + TMA) implicit-GEMM path.  Placeholder: no layer is planned onto the
// tensor cores yet, so every layer runs the reference kernels.
#include "tc_kernels.h"

namespace sv {

void tc_plan_layer(TcLayer& t, const ConvGeom&, int in_dt, int out_dt, bool, bool) {
  t.in_dt = in_dt;
  t.out_dt = out_dt;
}
size_t tc_workspace_bytes(const TcLayer& t, const ConvGeom&) { return t.w_fwd_bytes + t.w_dgrad_bytes + t.partial_bytes; }
const char* tc_bind_layer(TcLayer& t, const ConvGeom&, const void* in, void* out, void* dout, void* din, char*) {
  t.in = in; t.out = out; t.dout = dout; t.din = din;
  return nullptr;
}
int tc_repack_weights(TcLayer&, const ConvGeom&, const float*, cudaStream_t) { return 0; }
void tc_conv_fwd(TcLayer&, const ConvGeom&, const float*, void*, int, cudaStream_t) {}
void tc_conv_dgrad(TcLayer&, const ConvGeom&, const void*, int, void*, cudaStream_t) {}
void tc_conv_wgrad(TcLayer&, const ConvGeom&, float*, cudaStream_t) {}

}  // namespace sv
