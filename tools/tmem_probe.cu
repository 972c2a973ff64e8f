// Standalone probe #3 (round 2): TMEM read / write throughput and latency seen by epilogue warps.
// Each of `nwarps` warps (lane quadrant = warp % 4) streams tcgen05.ld (32x32b.x32 = 4 KB per warp instruction) over its
// own columns; reports cycles per instruction for a dependent (ld + wait) and a batched (4 x ld + wait) pattern, and the
// same for tcgen05.st.  Decides how many bytes per cycle an accumulator drain can expect (csrc/tower_fwd.cu epilogue).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/tmem_probe tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

#include "../news_recsys_b200/csrc/umma.cuh"

using namespace nrx::umma;

__global__ void __launch_bounds__(512) ldtm_kernel(int mode, int iters, long long* cycles, float* sink) {
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);   // 4 column blocks of 128
  float acc = 0.f;
  uint32_t w[16];
  for (int j = 0; j < 16; ++j) w[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {          // dependent: ld32 + wait
    for (int it = 0; it < iters; ++it) {
      float v[32];
      tmem_ld32(base + (uint32_t)((it & 3) * 32), v);
      tmem_ld_wait();
      acc += v[0] + v[31];
    }
  } else if (mode == 1) {   // 4 loads in flight, one wait
    for (int it = 0; it < iters; it += 4) {
      float v0[32], v1[32], v2[32], v3[32];
      tmem_ld32(base, v0); tmem_ld32(base + 32, v1); tmem_ld32(base + 64, v2); tmem_ld32(base + 96, v3);
      tmem_ld_wait();
      acc += v0[0] + v1[1] + v2[2] + v3[3];
    }
  } else if (mode == 2) {   // st16 (2 KB per warp instruction) + wait
    for (int it = 0; it < iters; ++it) {
      tmem_st16(base + (uint32_t)((it & 7) * 16), w);
      tmem_st_wait();
    }
  } else {                  // 4 stores in flight, one wait
    for (int it = 0; it < iters; it += 4) {
      tmem_st16(base, w); tmem_st16(base + 16, w); tmem_st16(base + 32, w); tmem_st16(base + 48, w);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  long long* dc; float* ds;
  cudaMalloc(&dc, 148 * 16 * 8); cudaMalloc(&ds, 148 * 512 * 4);
  const int iters = 1024;
  for (int nw : {1, 4, 8, 16}) for (int mode = 0; mode < 4; ++mode) {
    ldtm_kernel<<<1, nw * 32>>>(mode, iters, dc, ds);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("TMEM_PROBE error %s\n", cudaGetErrorString(e)); return 1; }
    long long c[16]; cudaMemcpy(c, dc, nw * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < nw; ++i) mx = c[i] > mx ? c[i] : mx;
    const double per = (double)mx / iters;
    const double bytes = (mode < 2 ? 4096.0 : 2048.0) * nw;
    printf("TMEM_PROBE warps=%2d %-28s %.1f cycles per warp instruction -> %.0f B/cycle/SM\n", nw,
           mode == 0 ? "ld.x32 + wait (dependent)" : mode == 1 ? "4 x ld.x32 + wait" : mode == 2 ? "st.x16 + wait (dependent)" : "4 x st.x16 + wait",
           per, bytes / per);
  }
  return 0;
}
