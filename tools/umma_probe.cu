// Standalone probe: validates the tcgen05 descriptor conventions used by csrc/umma.cuh on a real B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_bin/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../news_recsys_b200/csrc/umma.cuh"

using namespace nrx::umma;

constexpr int M = 128, N = 64, K = 64;

__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                 // M*K*2 = 16 KB
  uint8_t* sB = smem + M * K * 2;     // N*K*2 = 8 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill canonical layout
  for (int i = tid; i < M * (K / 8); i += 128) {
    const int r = i % M, kc = i / M;
    *reinterpret_cast<uint4*>(sA + canon_off(M, r, kc)) = *reinterpret_cast<const uint4*>(A + r * K + kc * 8);
  }
  for (int i = tid; i < N * (K / 8); i += 128) {
    const int r = i % N, kc = i / N;
    *reinterpret_cast<uint4*>(sB + canon_off(N, r, kc)) = *reinterpret_cast<const uint4*>(B + r * K + kc * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(M, N);
    for (int k = 0; k < K / 16; ++k) {
      uint64_t ad, bd;
      const uint32_t a0 = smem_u32(sA) + k * 2 * (M * 16), b0 = smem_u32(sB) + k * 2 * (N * 16);
      if (variant == 0) { ad = make_smem_desc(a0, M * 16, 128); bd = make_smem_desc(b0, N * 16, 128); }
      else              { ad = make_smem_desc(a0, 128, M * 16); bd = make_smem_desc(b0, 128, N * 16); }
      mma_bf16_ss(tmem, ad, bd, idesc, k > 0);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  float v[32];
  for (int c0 = 0; c0 < N; c0 += 32) {
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + (tid & 31)) * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__nv_bfloat16> hA(M * K), hB(N * K);
  std::vector<float> fA(M * K), fB(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(x); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * K; ++i) { float x = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(x); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < K; ++k) s += (double)fA[m * K + k] * fB[n * K + k]; ref[m * N + n] = (float)s; }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), N * K * 2, cudaMemcpyHostToDevice);
  const int smem = (M + N) * K * 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dD, 0, M * N * 4);
    probe_kernel<<<1, 128, smem>>>(dA, dB, dD, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs((double)out[i] - ref[i]));
    printf("UMMA_PROBE variant=%d (%s) max_abs_err=%.6g %s\n", variant, variant == 0 ? "LBO=K-chunk stride,SBO=128" : "swapped", maxerr, maxerr < 1e-3 ? "PASS" : "FAIL");
  }
  return 0;
}
