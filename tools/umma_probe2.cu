// Standalone probe #2 (round 2): (1) validates the TMEM-resident A operand of tcgen05.mma (".ts" form: activations
// written with tcgen05.st as packed bf16x2, one matrix row per TMEM lane) against a host reference, and
// (2) measures the issue-to-completion rate of back-to-back 128 x N x 16 MMAs with the A operand in shared memory
// (SS, no-swizzle canonical layout) versus in TMEM (TS).  Used to decide the layout of csrc/tower_fwd.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/umma_probe2 tools/umma_probe2.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../news_recsys_b200/csrc/umma.cuh"

using namespace nrx::umma;

constexpr int M = 128, N = 128, K = 128;

__global__ void __launch_bounds__(128) ts_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem;  // N*K*2
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * (K / 8); i += 128) {
    const int r = i % N, kc = i / N;
    *reinterpret_cast<uint4*>(sB + canon_off(N, r, kc)) = *reinterpret_cast<const uint4*>(B + r * K + kc * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t acc = tmem, act = tmem + 128;  // accumulator: 128 cols; A operand: K/2 = 64 cols
  // every thread writes its own row (lane = row) as packed bf16x2 words
  {
    const int r = warp * 32 + lane;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(A + r * K);
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t w[16];
      for (int j = 0; j < 16; ++j) w[j] = src[c0 + j];
      tmem_st16(act + ((uint32_t)(warp * 32) << 16) + c0, w);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t bd = make_smem_desc(smem_u32(sB) + k * 2 * (N * 16), N * 16, 128);
      mma_bf16_ts(acc, act + k * 8, bd, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// rate: `iters` rounds of K/16 MMAs (N columns each) issued by one thread, one commit per round
__global__ void __launch_bounds__(128) rate_kernel(int mode, int ncols, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                   // 128*128*2 = 32 KB
  uint8_t* sB = smem + 128 * 128 * 2;   // 256*128*2 = 64 KB
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (32 + 64) * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (warp == 0) {
    // the whole warp runs the loop in uniform control flow; one elected lane issues (umma::elect_one_sync)
    const uint32_t leader = elect_one_sync() ? 1u : 0u;
    const uint32_t idesc = make_idesc_bf16(128, ncols);
    const uint64_t ad0 = make_smem_desc(smem_u32(sA), 128 * 16, 128);
    const uint64_t bd0 = make_smem_desc(smem_u32(sB), ncols * 16, 128);
    const uint32_t bstep = (2u * (uint32_t)ncols * 16u) >> 4, astep = (2u * 128u * 16u) >> 4;
    const int nacc = mode >= 2 ? mode : 1;
    const long long t0 = clock64();
    // two rounds in flight: round `it` commits to bar[it & 1]; before re-using a barrier wait for its previous phase
    // (an mbarrier may run at most one phase ahead of its waiter)
    for (int it = 0; it < iters; ++it) {
      if (it >= 2) mbar_wait(&bar[it & 1], ((it - 2) >> 1) & 1);
      uint64_t ad = ad0, bd = bd0;
      uint32_t ta = tmem + 256;
      if (nacc == 1) {
#pragma unroll 2
        for (int k = 0; k < 8; ++k) {
          if (mode == 0) mma_bf16_ss_if(leader, tmem, ad, bd, idesc, k > 0);
          else mma_bf16_ts_if(leader, tmem, ta, bd, idesc, k > 0);
          ad += astep; bd += bstep; ta += 8;
        }
      } else {
        // `nacc` independent accumulators, round-robin per instruction
#pragma unroll 2
        for (int k = 0; k < 8; ++k) {
          mma_bf16_ss_if(leader, tmem + (uint32_t)((k & (nacc - 1)) * ncols), ad, bd, idesc, k >= nacc);
          ad += astep; bd += bstep;
        }
      }
      mma_commit_if(leader, &bar[it & 1]);
    }
    for (int it = iters - 2; it < iters; ++it)
      if (it >= 0) mbar_wait(&bar[it & 1], (it >> 1) & 1);
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  {
    const int smem = N * K * 2;
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(dD, 0, M * N * 4);
    ts_kernel<<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("UMMA_PROBE2 ts: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
    printf("UMMA_PROBE2 A-in-TMEM (lane=row, bf16x2 per column) max_abs_err=%.6g %s\n", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  }
  {
    const int smem = 96 * 1024;
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long* dc; cudaMalloc(&dc, 148 * 8);
    for (int grid : {1, 148}) for (int mode : {0, 1, 2, 4}) for (int ncols : {32, 64, 128, 256}) {
      if (mode >= 2 && mode * ncols > 512) continue;
      const int iters = 512;
      rate_kernel<<<grid, 128, smem>>>(mode, ncols, iters, dc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("UMMA_PROBE2 rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      long long c[148]; cudaMemcpy(c, dc, grid * 8, cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) mx = c[i] > mx ? c[i] : mx;
      const double per = (double)mx / (iters * 8.0);
      printf("UMMA_PROBE2 rate grid=%d %s N=%d: %.1f cycles per 128xNx16 MMA (floor %.0f) -> %.0f%% of the tensor floor\n", grid,
             mode == 0 ? "SS" : mode == 1 ? "TS" : mode == 2 ? "SS-2acc" : "SS-4acc", ncols, per, ncols / 2.0, 100.0 * (ncols / 2.0) / per);
    }
  }
  return 0;
}
