// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M=128) as a function of N, operand major-ness,
// swizzle mode, descriptor strides and start-address alignment.  One thread issues `iters` MMAs from shared memory
// that is never reloaded (contents irrelevant), cycling over `nacc` accumulators; time = clock64 around issue+commit+wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I splitvae_b200/csrc scripts/microbench/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "tc_device.cuh"

using namespace sv;

struct Cfg {
  const char* name;
  int N, a_mn, b_mn, a_lt, b_lt;          // layout types: 2=SW128 4=SW64 6=SW32
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;    // bytes
  uint32_t a_off[8], b_off[8];            // byte offsets cycled per MMA
  int n_off, nacc;
};

struct Params {
  uint32_t idesc;
  uint64_t a_tmpl, b_tmpl;
  uint32_t a_off[8], b_off[8];
  int n_off, nacc, N, iters;
  long long* out;
};

__global__ void __launch_bounds__(128) k(const __grid_constant__ Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(&tmem_base_s, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t sa = tc::smem_u32(smem), sb = sa + 128 * 1024;
    // descriptors precomputed into registers and the loop unrolled x8, so the issuing thread spends ~2 instructions per MMA
    uint64_t da[8], db[8];
    uint32_t acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      da[j] = P.a_tmpl + ((sa + P.a_off[j % P.n_off]) >> 4);
      db[j] = P.b_tmpl + ((sb + P.b_off[j % P.n_off]) >> 4);
      acc[j] = tmem_base + (j % P.nacc) * P.N;
    }
    long long t0 = clock64();
    for (int i = 0; i < P.iters; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) tc::umma_bf16(acc[j], da[j], db[j], P.idesc, 1u);
    }
    long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { P.out[0] = t1 - t0; P.out[1] = t2 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); }
}

static uint64_t tmpl(uint32_t lbo, uint32_t sbo, uint32_t lt) {
  uint64_t d = 0;
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)lt << 61;
  return d;
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 1;
  const int iters = 4096;
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  std::vector<Cfg> cfgs;
  auto add = [&](Cfg c) { cfgs.push_back(c); };
  // ---- K-major, canonical aligned tiles, one accumulator (dependent chain) and 4 accumulators
  for (int N : {16, 32, 64, 128, 256}) {
    add({"Kmaj SW128 aligned nacc1", N, 0, 0, 2, 2, 16, 1024, 16, 1024, {0, 32, 64, 96}, {0, 32, 64, 96}, 4, 1});
    add({"Kmaj SW128 aligned nacc4", N, 0, 0, 2, 2, 16, 1024, 16, 1024, {0, 32, 64, 96}, {0, 32, 64, 96}, 4, N <= 128 ? 4 : 2});
  }
  // ---- K-major SW128, A shifted by pixels (halo views): start offsets multiples of 128 B, SBO = 37 pixels
  for (int N : {16, 32, 64, 128}) {
    add({"Kmaj SW128 A shifted SBO=37px", N, 0, 0, 2, 2, 16, 37 * 128, 16, 1024, {128, 128 + 32, 384 + 64, 640 + 96, 37 * 128, 37 * 128 + 32 + 256}, {0, 32, 64, 96, 0, 32}, 6, 4});
    add({"Kmaj SW128 A aligned SBO=40px", N, 0, 0, 2, 2, 16, 40 * 128, 16, 1024, {0, 32, 64, 96}, {0, 32, 64, 96}, 4, 4});
    add({"Kmaj SW128 A shifted SBO=40px", N, 0, 0, 2, 2, 16, 40 * 128, 16, 1024, {128, 128 + 32, 384 + 64, 640 + 96}, {0, 32, 64, 96}, 4, 4});
  }
  // ---- K-major SW64 / SW32 (narrow channels)
  for (int N : {16, 32}) {
    add({"Kmaj SW64 aligned", N, 0, 0, 4, 4, 16, 512, 16, 512, {0, 32}, {0, 32}, 2, 4});
    add({"Kmaj SW64 A shifted SBO=37px", N, 0, 0, 4, 4, 16, 37 * 64, 16, 512, {64, 64 + 32, 192, 192 + 32}, {0, 32, 0, 32}, 4, 4});
    add({"Kmaj SW32 aligned", N, 0, 0, 6, 6, 16, 256, 16, 256, {0, 0}, {0, 0}, 1, 4});
    add({"Kmaj SW32 A shifted SBO=69px", N, 0, 0, 6, 6, 16, 69 * 32, 16, 256, {32, 96, 160}, {0, 0, 0}, 3, 4});
  }
  // ---- MN-major canonical (A: 2 atoms of 64 ch at LBO = 8 KB; K rows of 128 B; SBO = 1024)
  for (int N : {16, 64, 128, 256}) {
    add({"MNmaj SW128 A canonical, B canonical", N, 1, 1, 2, 2, 8192, 1024, 8192, 1024, {0, 2048, 4096}, {0, 2048, 4096}, 3, N <= 128 ? 4 : 2});
  }
  // ---- MN-major halo views: A blocks shifted by one pixel (LBO = 128 B) / one row
  add({"MNmaj SW128 A LBO=1px  N=32 (B canon)", 32, 1, 1, 2, 4, 128, 1024, 4096, 512, {0, 2048, 37 * 128}, {0, 1024, 2048}, 3, 4});
  add({"MNmaj SW128 A LBO=row  N=192 B LBO=1px SW64", 192, 1, 1, 2, 4, 32 * 128, 1024, 64, 512, {0, 2048, 4096}, {0, 1024, 2048}, 3, 2});
  add({"MNmaj SW128 A canon    N=192 B LBO=1px SW64", 192, 1, 1, 2, 4, 8192, 1024, 64, 512, {0, 2048, 4096}, {0, 1024, 2048}, 3, 2});
  add({"MNmaj SW128 A canon    N=192 B canon SW64", 192, 1, 1, 2, 4, 8192, 1024, 2048, 512, {0, 2048, 4096}, {0, 1024, 2048}, 3, 2});
  add({"MNmaj SW64 A LBO=row   N=96 B LBO=1px SW32", 96, 1, 1, 4, 6, 32 * 64, 512, 32, 256, {0, 1024, 2048}, {0, 512, 1024}, 3, 4});
  add({"MNmaj SW64 A canon     N=96 B canon SW32", 96, 1, 1, 4, 6, 4096, 512, 1024, 256, {0, 1024, 2048}, {0, 512, 1024}, 3, 4});
  add({"MNmaj SW64 A canon     N=16 B canon SW32", 16, 1, 1, 4, 6, 4096, 512, 1024, 256, {0, 1024, 2048}, {0, 512, 1024}, 3, 4});
  add({"MNmaj SW64 A LBO=1px   N=16 B canon SW32", 16, 1, 1, 4, 6, 64, 512, 1024, 256, {0, 1024, 2048}, {0, 512, 1024}, 3, 4});
  add({"MNmaj SW128 A canon    N=256 B LBO=1px SW128", 256, 1, 1, 2, 2, 8192, 1024, 128, 1024, {0, 2048, 4096}, {0, 2048, 4096}, 3, 2});
  // mixed: A K-major, B MN-major and vice versa
  add({"A Kmaj SW128, B MNmaj SW128 N=128", 128, 0, 1, 2, 2, 16, 1024, 8192, 1024, {0, 32, 64, 96}, {0, 2048, 4096, 6144}, 4, 4});
  add({"A MNmaj SW128, B Kmaj SW128 N=128", 128, 1, 0, 2, 2, 8192, 1024, 16, 1024, {0, 2048, 4096, 6144}, {0, 32, 64, 96}, 4, 4});

  printf("%-48s %5s %5s %10s %10s\n", "config", "N", "nacc", "cyc/issue", "cyc/mma");
  for (auto& c : cfgs) {
    Params P{};
    P.idesc = tc::make_idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    P.a_tmpl = tmpl(c.a_lbo, c.a_sbo, c.a_lt);
    P.b_tmpl = tmpl(c.b_lbo, c.b_sbo, c.b_lt);
    for (int i = 0; i < 8; ++i) { P.a_off[i] = c.a_off[i]; P.b_off[i] = c.b_off[i]; }
    P.n_off = c.n_off; P.nacc = c.nacc; P.N = c.N; P.iters = iters; P.out = out;
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      k<<<grid, 128, 195 * 1024>>>(P);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%-48s %5d %5d %10.1f %10.1f\n", c.name, c.N, c.nacc, (double)h[0] / iters, (double)h[1] / iters);
  }
  return 0;
}
