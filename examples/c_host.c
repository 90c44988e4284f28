/* A non-Python host of libsplitvae.so in plain C99: what a maintainer of another runtime binds.
 *
 *   gcc -std=c99 -Iinclude examples/c_host.c -o c_host -Lsplitvae_b200 -lsplitvae -Wl,-rpath,$PWD/splitvae_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart
 *   ./c_host                 plan only (no GPU needed): the variable inventory of model.trainable_variables + buffer sizes
 *   ./c_host --run [steps]   on a B200: bind caller-owned buffers, capture the whole train step as a CUDA graph, replay it,
 *                            read the step's scalars (vae/trainer.py:125-135) and the optimizer's iteration count
 *
 * The CUDA runtime is only used for the caller's side of the contract: allocating the buffers the library is bound to, the stream,
 * and copies of the caller's own data.  Everything the reference does inside train_step_lg_vae happens behind sv_* calls. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "splitvae.h"

/* the few CUDA runtime entry points a host needs (declared here so the example compiles without cuda_runtime.h) */
typedef struct CUstream_st* cudaStream_t;
extern int cudaMalloc(void** p, size_t n);
extern int cudaFree(void* p);
extern int cudaMemset(void* p, int v, size_t n);
extern int cudaMemcpy(void* dst, const void* src, size_t n, int kind);   /* kind 1 = host to device, 2 = device to host */
extern int cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);    /* flags 1 = non-blocking */
extern int cudaStreamSynchronize(cudaStream_t s);
extern int cudaStreamDestroy(cudaStream_t s);

#define CHECK(call)                                                                              \
  do {                                                                                           \
    sv_status st_ = (call);                                                                      \
    if (st_ != SV_OK) {                                                                          \
      fprintf(stderr, "%s failed (%d): %s\n", #call, (int)st_, sv_last_error(h));                \
      return 1;                                                                                  \
    }                                                                                            \
  } while (0)
#define CUDA(call)                                                                               \
  do {                                                                                           \
    int e_ = (call);                                                                             \
    if (e_ != 0) { fprintf(stderr, "%s failed: cuda error %d\n", #call, e_); return 1; }          \
  } while (0)

static float lcg_uniform(unsigned* s) {   /* deterministic host-side fill, [-1, 1) */
  *s = *s * 1664525u + 1013904223u;
  return (float)((*s >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

int main(int argc, char** argv) {
  const int run = argc > 1 && strcmp(argv[1], "--run") == 0;
  const int steps = argc > 2 ? atoi(argv[2]) : 20;
  sv_handle* h = NULL;
  sv_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.model = SV_MODEL_LGVAE;                 /* vae/main.py --model lgvae --dataset celeba64 --beta 120 -no_label */
  cfg.height = 64; cfg.width = 64; cfg.batch = run ? 32 : 256;
  cfg.global_latent_dims = 128; cfg.local_latent_dims = 128;
  cfg.y_size = 30; cfg.tau = 0.4f; cfg.beta = 120.0f; cfg.alpha = 40.0f; cfg.learning_rate = 1e-4f;
  cfg.world_size = 1;
  cfg.precision = SV_PRECISION_BF16X3;
  cfg.flags = run ? 0 : 1;                    /* 1 = plan only: inventory and sizes without touching CUDA */
  printf("%s\n", sv_version());
  CHECK(sv_create(&cfg, &h));

  const int n = sv_param_count(h);
  const long long arena = (long long)sv_arena_floats(h), ws_bytes = (long long)sv_workspace_bytes(h);
  printf("variables %d arena_floats %lld workspace_bytes %lld segments %d\n", n, arena, ws_bytes, (int)sv_num_segments(h));
  for (int i = 0; i < n; ++i) {
    sv_param_desc d;
    CHECK(sv_param_describe(h, i, &d));
    printf("  %-28s [", d.name);
    for (int k = 0; k < d.ndim; ++k) printf("%s%d", k ? "," : "", (int)d.shape[k]);
    printf("] offset %lld count %lld\n", (long long)d.offset, (long long)d.count);
  }
  if (!run) { sv_destroy(h); return 0; }

  /* ---- caller-owned device buffers (sv_bind) ---- */
  float *params = NULL, *grads = NULL, *m = NULL, *v = NULL, *inputs = NULL;
  void* ws = NULL;
  const size_t abytes = (size_t)arena * sizeof(float);
  const size_t in_floats = (size_t)cfg.batch * cfg.height * cfg.width * 6;
  CUDA(cudaMalloc((void**)&params, abytes)); CUDA(cudaMalloc((void**)&grads, abytes));
  CUDA(cudaMalloc((void**)&m, abytes)); CUDA(cudaMalloc((void**)&v, abytes));
  CUDA(cudaMalloc(&ws, (size_t)ws_bytes)); CUDA(cudaMalloc((void**)&inputs, in_floats * sizeof(float)));
  CUDA(cudaMemset(grads, 0, abytes)); CUDA(cudaMemset(m, 0, abytes)); CUDA(cudaMemset(v, 0, abytes)); CUDA(cudaMemset(ws, 0, (size_t)ws_bytes));
  /* weights: uniform(-limit, limit) per variable (Glorot-like, fan_in + fan_out from the Keras shapes), biases zero */
  float* host = (float*)calloc((size_t)arena, sizeof(float));
  unsigned seed = 5;
  for (int i = 0; i < n; ++i) {
    sv_param_desc d;
    CHECK(sv_param_describe(h, i, &d));
    if (d.ndim < 2) continue;
    double rf = 1.0;
    for (int k = 0; k + 2 < d.ndim; ++k) rf *= d.shape[k];
    const double fan_in = rf * d.shape[d.ndim - 2], fan_out = rf * d.shape[d.ndim - 1];
    double limit = 6.0 / (fan_in + fan_out), x = limit;
    for (int it = 0; it < 30; ++it) x = 0.5 * (x + limit / x);          /* sqrt without libm */
    for (long long j = 0; j < (long long)d.count; ++j) host[d.offset + j] = (float)x * lcg_uniform(&seed);
  }
  CUDA(cudaMemcpy(params, host, abytes, 1));
  float* hin = (float*)malloc(in_floats * sizeof(float));
  for (size_t j = 0; j < in_floats; ++j) hin[j] = lcg_uniform(&seed);
  CUDA(cudaMemcpy(inputs, hin, in_floats * sizeof(float), 1));

  cudaStream_t s;
  CUDA(cudaStreamCreateWithFlags(&s, 1));
  CHECK(sv_bind(h, params, grads, m, v, ws, ws_bytes));
  CHECK(sv_params_updated(h, s));
  /* ---- the captured step: NULL noise pointers = in-kernel Philox draws, fresh at every replay ---- */
  CHECK(sv_capture_graph(h, inputs, NULL, NULL, NULL, s));
  float first[SV_SCALAR_COUNT], last[SV_SCALAR_COUNT];
  void* sc = NULL;
  int64_t cnt = 0;
  CHECK(sv_output_ptr(h, SV_OUT_SCALARS, &sc, &cnt));
  CHECK(sv_replay(h, 1, s));
  CUDA(cudaStreamSynchronize(s));
  CUDA(cudaMemcpy(first, sc, sizeof(first), 2));
  CHECK(sv_replay(h, steps - 1, s));
  CUDA(cudaStreamSynchronize(s));
  CUDA(cudaMemcpy(last, sc, sizeof(last), 2));
  int64_t it = 0;
  CHECK(sv_get_iterations(h, &it, s));
  printf("step 1: total %.3f recon_x %.3f kl_x %.4f | step %d: total %.3f | optimizer.iterations %lld | kernels launched %lld\n",
         first[SV_SCALAR_TOTAL], first[SV_SCALAR_RECON_X], first[SV_SCALAR_KL_X], steps, last[SV_SCALAR_TOTAL], (long long)it,
         (long long)sv_launch_count(h));
  const int ok = it == steps && last[SV_SCALAR_TOTAL] == last[SV_SCALAR_TOTAL] && last[SV_SCALAR_TOTAL] < first[SV_SCALAR_TOTAL];
  printf("%s\n", ok ? "OK: the loss fell over the replayed steps" : "FAILED");
  sv_destroy(h);
  cudaStreamDestroy(s);
  cudaFree(params); cudaFree(grads); cudaFree(m); cudaFree(v); cudaFree(ws); cudaFree(inputs);
  free(host); free(hin);
  return ok ? 0 : 1;
}
