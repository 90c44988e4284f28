// Shared device helpers and host-side descriptors for libsplitvae (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sv {

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2, ACT_SOFTPLUS = 3 };
enum DType : int { DT_F32 = 0, DT_BF16 = 1 };

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float round_bf16(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------
// A kernel launched through launch_pdl() may be scheduled while the previous kernel of its stream is still running (the ~3 us
// between two dependent launches - grid drain, launch latency, the next kernel's set-up - overlap): pdl_wait() blocks until that
// kernel has completed and its writes are visible, so everything before it may only touch memory no earlier kernel of the step
// writes; pdl_trigger() lets the NEXT kernel's CTAs be scheduled once this grid's CTAs are all resident.  Rule: every thread (or at
// least one thread per CTA whose exit the CTA waits for) runs pdl_wait() before any dependent access and before the CTA can exit, so
// that completion stays transitive along the stream.  Both are no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }     // kernels without a prologue worth overlapping

// Keras activations (vae/model.py:36-42,50-76,152-156): relu, elu(alpha=1), softplus.
__device__ __forceinline__ float softplus_f(float x) {
  // log(1+e^x), stable for both tails (TF's softplus switches to x / e^x at the extremes too)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_ELU: return x > 0.f ? x : expm1f(x);
    case ACT_SOFTPLUS: return softplus_f(x);
    default: return x;
  }
}
// derivative of the activation expressed through its OUTPUT y
__device__ __forceinline__ float act_grad_from_out(float y, int act) {
  switch (act) {
    case ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case ACT_SOFTPLUS: return 1.f - expf(-y);
    default: return 1.f;
  }
}

// One "layer" = Conv2D(padding='same') or Dense (a 1x1 conv on a 1x1 image).  Up to three Keras
// variables may be fused along the output-channel axis (e.g. e4_mean|e4_sd share their input).
struct ConvGeom {
  int B, Hi, Wi, Ci, Ho, Wo, Co;   // logical (unpadded) sizes; Co = sum(part_n)
  int kh, kw, stride, pt, pl;      // TF 'same' padding: pt/pl = pad before (top/left)
  int in_ld, in_coff;              // channel pitch / first channel of the input tensor
  int out_ld;                      // channel pitch of the output tensor
  int dout_ld;                     // channel pitch of the output-gradient tensor
  int din_ld;                      // channel pitch of the input-gradient tensor
  int nparts;
  int part_n[3];
  int part_act[3];
  long long part_w[3];             // float offsets into the parameter/gradient arenas
  long long part_b[3];
};

__device__ __forceinline__ int part_of(const ConvGeom& g, int co, int& local) {
  int j = 0;
  local = co;
  while (j + 1 < g.nparts && local >= g.part_n[j]) { local -= g.part_n[j]; ++j; }
  return j;
}

// Host side: launch with programmatic stream serialization.  ONLY for kernels that follow the rule above.  SV_PDL=0: plain launches.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace sv
