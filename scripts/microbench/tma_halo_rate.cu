// Micro-benchmark: how fast can ONE SM's TMA unit stream row-halo boxes of an NHWC bf16 tensor into a shared-memory ring when nothing
// else happens (no MMAs, no epilogue)?  Persistent grid of 148 CTAs, one producer thread, one consumer thread that only waits for a
// stage and hands it back.  Cases are the halo shapes of nsconv_kernel: the question is whether its measured
// "loads only" floor (SV_NS_DEBUG=3: 1000-2000 cycles per tile, 10-24 B/clk/SM) is a property of the box geometry / ring depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I splitvae_b200/csrc scripts/microbench/tma_halo_rate.cu -o gpurun_out/tma_halo_rate -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "tc_device.cuh"

using namespace sv;

struct Params {
  CUtensorMap map;
  int W, H, n_img;          // image geometry (pixels)
  int rows, step;           // halo rows per tile, output rows per tile (tile t of an image starts at row t*step - 2)
  int tiles_per_img, tiles;
  int stages, stage_bytes;
  int split;                // TMA instructions per stage (the box is `rows / split` rows)
  int box_bytes;
  long long* out;           // per CTA: cycles
};

struct Ctl { uint64_t full[8], empty[8]; };

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + (size_t)P.stages * P.stage_bytes);
  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&P.map);
    for (int i = 0; i < P.stages; ++i) { tc::mbar_init(&ctl->full[i], 1); tc::mbar_init(&ctl->empty[i], 1); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0) {
    if (tc::elect_one()) {
      int i = 0;
      for (int t = blockIdx.x; t < P.tiles; t += gridDim.x, ++i) {
        const int st = i % P.stages, ph = (i / P.stages) & 1;
        const int n = t / P.tiles_per_img, y0 = (t - n * P.tiles_per_img) * P.step - 2;
        tc::mbar_wait(&ctl->empty[st], ph ^ 1);
        tc::mbar_expect_tx(&ctl->full[st], (uint32_t)(P.split * P.box_bytes));
        for (int j = 0; j < P.split; ++j)
          tc::tma_load_4d(smem + (size_t)st * P.stage_bytes + (size_t)j * P.box_bytes, &P.map, &ctl->full[st], 0, 0, y0 + j * (P.rows / P.split), n);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      int i = 0;
      for (int t = blockIdx.x; t < P.tiles; t += gridDim.x, ++i) {
        const int st = i % P.stages, ph = (i / P.stages) & 1;
        tc::mbar_wait(&ctl->full[st], ph);
        tc::mbar_arrive(&ctl->empty[st]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) P.out[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN_encodeTiled enc = (PFN_encodeTiled)fn;
  struct Case { const char* name; int C, W, H, rows, step, stages, split; };
  const Case cases[] = {
      {"d5 fwd   C32 W64 9 rows/4  5 st", 32, 64, 64, 9, 4, 5, 1},
      {"d5 fwd   same, 3 boxes/stage   ", 32, 64, 64, 9, 4, 5, 3},
      {"d5 fwd   same, 9 boxes/stage   ", 32, 64, 64, 9, 4, 5, 9},
      {"d5 fwd   C32 W64 21 rows/16 2st", 32, 64, 64, 21, 16, 2, 1},
      {"d5 fwd   same, 3 boxes/stage   ", 32, 64, 64, 21, 16, 2, 3},
      {"d4 fwd   C64 W32 9 rows/4  2 st", 64, 32, 32, 9, 4, 2, 1},
      {"d4 fwd   same 4 stages         ", 64, 32, 32, 9, 4, 4, 1},
      {"d4 fwd   same 4 st, 3 boxes    ", 64, 32, 32, 9, 4, 4, 3},
      {"d4 fwd   same 4 st, 9 boxes    ", 64, 32, 32, 9, 4, 4, 9},
      {"d4 dgrad C32 W32 9 rows/4  6 st", 32, 32, 32, 9, 4, 6, 1},
      {"d4 dgrad same, 3 boxes         ", 32, 32, 32, 9, 4, 6, 3},
      {"d4 dgrad same, 9 boxes         ", 32, 32, 32, 9, 4, 6, 9},
      {"d3 fwd   C64 W16 11 rows/8 4 st", 64, 16, 16, 11, 8, 4, 1},
      {"d3 fwd   same, 11 boxes        ", 64, 16, 16, 11, 8, 4, 11},
  };
  const int n_img = 256;
  long long* d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  printf("%-34s %9s %9s %9s %9s\n", "case", "us", "GB/s", "clk/tile", "B/clk/SM");
  for (const Case& c : cases) {
    const size_t elems = (size_t)n_img * c.H * c.W * c.C;
    void* d_x;
    cudaMalloc(&d_x, elems * 2);
    cudaMemset(d_x, 0, elems * 2);
    Params P{};
    const int rows_box = c.rows / c.split;
    cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)n_img};
    cuuint64_t strides[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)c.W * c.C * 2, (cuuint64_t)c.H * c.W * c.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)c.C, (cuuint32_t)c.W, (cuuint32_t)rows_box, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapSwizzle sw = c.C * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(&P.map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", c.name, (int)r); continue; }
    P.W = c.W; P.H = c.H; P.n_img = n_img; P.rows = c.rows; P.step = c.step;
    P.tiles_per_img = c.H / c.step; P.tiles = P.tiles_per_img * n_img;
    P.stages = c.stages; P.split = c.split;
    P.box_bytes = rows_box * c.W * c.C * 2;
    P.stage_bytes = (c.rows * c.W * c.C * 2 + 1023) / 1024 * 1024;
    P.out = d_out;
    const size_t smem = (size_t)P.stages * P.stage_bytes + sizeof(Ctl) + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, 128, smem>>>(P);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) k<<<148, 128, smem>>>(P);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(err)); return 1; }
    const double us = ms * 1000.0 / reps;
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (long long v : h) mx = v > mx ? v : mx;
    const double bytes = (double)P.tiles * c.rows * c.W * c.C * 2;
    const double tiles_per_cta = (double)P.tiles / 148.0;
    printf("%-34s %9.1f %9.0f %9.0f %9.1f\n", c.name, us, bytes / us * 1e-3, mx / tiles_per_cta, bytes / 148.0 / mx);
    cudaFree(d_x);
  }
  return 0;
}
