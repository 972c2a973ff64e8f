// Shared device/host helpers for libnrx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/nrx.h"

#define NRX_FULL_MASK 0xffffffffu

namespace nrx {

// ---- error plumbing (thread-local message, no exceptions across the ABI) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> NRX_ELAUNCH

#define NRX_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) {                            \
      ::nrx::set_error(__VA_ARGS__);          \
      return (code);                          \
    }                                         \
  } while (0)

int sm_count();  // cached per process for the current device

// ---- device-side feature descriptors (kernel parameter, < 4 KB) --------------
struct DFeat {
  const float* table;
  const void* idx;
  const float* mask;
  float* inv_den;
  long long rows;
  long long occ_off;  // first occurrence number of this feature in the backward plan
  int dim, stride, L, pool, out_col, idx32, table_id, cstart;  // cstart: first vec column (sparse group)
};

struct DFeats {
  DFeat f[NRX_MAX_FEATS];
  int n;
  int n_sparse, n_array;
  int sparse_ids[NRX_MAX_FEATS];
  int array_ids[NRX_MAX_FEATS];
  int sparse_cols;  // total vec columns of the sparse group
  int vec;          // 4 or 1
  int max_dim;
  int n_tables;
  long long n_occ;
};

// Validates and converts host NrxFeat[] -> DFeats.  `out_ld`/`out` only used for the vector-width decision.
int make_dfeats(const NrxFeat* feats, int n, long long B, const void* out, long long out_ld, DFeats* d);

__device__ __forceinline__ long long load_idx(const void* p, long long i, int idx32) {
  return idx32 ? (long long)__ldg(reinterpret_cast<const int*>(p) + i)
               : __ldg(reinterpret_cast<const long long*>(p) + i);
}

// torch.optim.AdamW update of one element (deep/model.py:55); shared by the dense kernels and K7 so that the
// single-GPU and the peer-memory step round identically.
__device__ __forceinline__ void adamw_update(float& p, float g, float& m, float& v, float lr, float bc1, float bc2s,
                                             float b1, float b2, float eps, float wd) {
  float pi = p * (1.f - lr * wd);
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  pi -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
  p = pi;
}

__device__ __forceinline__ float sigmoid_f(float z) { return 1.f / (1.f + expf(-z)); }

__device__ __forceinline__ void bce_terms(float p, float y, float invB, float* loss, float* dlogit) {
  // F.binary_cross_entropy: log clamped at -100; autograd: (p-y)/max(p(1-p),1e-12) then sigmoid' = p(1-p)
  const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
  if (loss) *loss = -(y * lp + (1.f - y) * l1p);
  if (dlogit) {
    const float pq = p * (1.f - p);
    *dlogit = ((p - y) / fmaxf(pq, 1e-12f)) * pq * invB;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(NRX_FULL_MASK, v, o);
  return v;
}

template <int V>
struct VecT;
template <>
struct VecT<4> {
  using type = float4;
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float4 load(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
  static __device__ __forceinline__ void fma(float4& a, float m, float4 v) {
    a.x = fmaf(m, v.x, a.x); a.y = fmaf(m, v.y, a.y); a.z = fmaf(m, v.z, a.z); a.w = fmaf(m, v.w, a.w);
  }
  static __device__ __forceinline__ float4 shfl_xor(float4 v, int o) {
    v.x = __shfl_xor_sync(NRX_FULL_MASK, v.x, o); v.y = __shfl_xor_sync(NRX_FULL_MASK, v.y, o);
    v.z = __shfl_xor_sync(NRX_FULL_MASK, v.z, o); v.w = __shfl_xor_sync(NRX_FULL_MASK, v.w, o);
    return v;
  }
  static __device__ __forceinline__ void add(float4& a, float4 b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  static __device__ __forceinline__ float4 div(float4 a, float d) { return make_float4(a.x / d, a.y / d, a.z / d, a.w / d); }
};
template <>
struct VecT<1> {
  using type = float;
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float load(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void store(float* p, float v) { *p = v; }
  static __device__ __forceinline__ void fma(float& a, float m, float v) { a = fmaf(m, v, a); }
  static __device__ __forceinline__ float shfl_xor(float v, int o) { return __shfl_xor_sync(NRX_FULL_MASK, v, o); }
  static __device__ __forceinline__ void add(float& a, float b) { a += b; }
  static __device__ __forceinline__ float div(float a, float d) { return a / d; }
};

// bf16 tile image of a [B, Kp] operand of the tower kernels: [tile][Kp/8][128 rows][8] — 16 bytes per
// (row, 8-column chunk).  Stores 4 consecutive columns (col % 4 == 0) of row b.
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
__device__ __forceinline__ void img_store4(uint8_t* __restrict__ img, int Kp, long long b, int col, float4 v) {
  uint8_t* p = img + (size_t)(b >> 7) * Kp * 256 + (size_t)(col >> 3) * 2048 + (size_t)(b & 127) * 16 + (size_t)(col & 7) * 2;
  *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
// zero the rows of the last tile that lie beyond B (the dW GEMMs contract over all 128 rows of a tile)
int img_zero_tail(void* image, int Kp, long long B, cudaStream_t st);

static inline int pow2_ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace nrx
