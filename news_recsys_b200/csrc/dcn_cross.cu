// K5 — DCN-v1 cross stack, fused (reference src/model/sort/dcn/dcn_arch.py:14-30 layer, :53-70 net;
// concat with the raw features for the head: sort/dcn/model.py:29).
//
// The reference materialises x0 * xl^T as a [B, d, d] tensor per layer (3.3 GB at B = 65536, d = 112).
// Algebraically (x0 xl^T) w == x0 * (xl . w), so one warp per row keeps x0 / x_l in registers, does the
// row-dot with a shuffle reduction and writes cat[x, x_L] once: 4 B * (d + 2 d) of HBM traffic per row.
// Backward recomputes the x_l chain in registers; the parameter gradients (sums over the batch) are
// accumulated per warp in a fixed row order, combined per block in shared memory in warp order and summed
// over blocks by a second kernel in block order => bitwise reproducible, no float atomics.
#include "common.cuh"

namespace nrx {

static constexpr int kMaxCrossLayers = 8;
static constexpr int kCrossNC = 8;  // lane owns columns lane + 32*k, k < 8  => d <= 256

struct CrossP {
  const float* w[kMaxCrossLayers];
  const float* b[kMaxCrossLayers];
  int n_layers, d;
};

template <int NC>
__global__ void __launch_bounds__(256)
dcn_cross_fwd_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ CrossP P,
                     float* __restrict__ out, long long old, float* __restrict__ dots) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int d = P.d;
  float x0[NC], xl[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int c = lane + 32 * k;
    x0[k] = c < d ? __ldg(x + row * ld + c) : 0.f;
    xl[k] = x0[k];
  }
  for (int l = 0; l < P.n_layers; ++l) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < d) s = fmaf(xl[k], __ldg(P.w[l] + c), s);
    }
    s = warp_sum(s);
    if (dots != nullptr && lane == 0) dots[(long long)l * B + row] = s;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < d) xl[k] = fmaf(x0[k], s, __ldg(P.b[l] + c) + xl[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int c = lane + 32 * k;
    if (c < d) {
      out[row * old + c] = x0[k];
      out[row * old + d + c] = xl[k];
    }
  }
}

// 128-bit variant: lane owns the float4 at columns 4 * (lane + 32 k).  Emits the fp32 concat (optional) and / or the
// bf16 tile image of cat[x, x_L] that the tower forward streams as its layer-0 operand (nrx_tower_fwd, NRX_TOWER_XIMG).
template <int NV>
__global__ void __launch_bounds__(256)
dcn_cross_fwd_v4_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ CrossP P,
                        float* __restrict__ out, long long old, uint8_t* __restrict__ img, int img_kp) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= B) return;
  const int d = P.d;
  float4 x0[NV], xl[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    x0[k] = c < d ? __ldg(reinterpret_cast<const float4*>(x + row * ld + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    xl[k] = x0[k];
  }
  for (int l = 0; l < P.n_layers; ++l) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(P.w[l] + c));
        s = fmaf(xl[k].x, w.x, s); s = fmaf(xl[k].y, w.y, s); s = fmaf(xl[k].z, w.z, s); s = fmaf(xl[k].w, w.w, s);
      }
    }
    s = warp_sum(s);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(P.b[l] + c));
        xl[k].x = fmaf(x0[k].x, s, b.x + xl[k].x); xl[k].y = fmaf(x0[k].y, s, b.y + xl[k].y);
        xl[k].z = fmaf(x0[k].z, s, b.z + xl[k].z); xl[k].w = fmaf(x0[k].w, s, b.w + xl[k].w);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    if (c < d) {
      if (out != nullptr) {
        *reinterpret_cast<float4*>(out + row * old + c) = x0[k];
        *reinterpret_cast<float4*>(out + row * old + d + c) = xl[k];
      }
      if (img != nullptr) {
        img_store4(img, img_kp, row, c, x0[k]);
        img_store4(img, img_kp, row, d + c, xl[k]);
      }
    }
  }
}

// partials layout: [block][layer][2 (w,b)][d]
template <int NC>
__global__ void __launch_bounds__(256)
dcn_cross_bwd_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ CrossP P,
                     const float* __restrict__ go, long long gold, float* __restrict__ gx, long long gxld,
                     float* __restrict__ partials) {
  extern __shared__ float sm[];  // [8 warps][layers][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.d, Lc = P.n_layers;
  float gw[kMaxCrossLayers][NC], gb[kMaxCrossLayers][NC];
#pragma unroll
  for (int l = 0; l < kMaxCrossLayers; ++l)
#pragma unroll
    for (int k = 0; k < NC; ++k) { gw[l][k] = 0.f; gb[l][k] = 0.f; }

  const long long warps_total = (long long)gridDim.x * 8;
  for (long long row = (long long)blockIdx.x * 8 + warp; row < B; row += warps_total) {
    float x0[NC], xs[kMaxCrossLayers][NC], s[kMaxCrossLayers];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      x0[k] = c < d ? __ldg(x + row * ld + c) : 0.f;
    }
    // recompute the chain, keeping x_l (input of layer l)
#pragma unroll
    for (int l = 0; l < kMaxCrossLayers; ++l) {
      if (l >= Lc) break;
#pragma unroll
      for (int k = 0; k < NC; ++k) xs[l][k] = (l == 0) ? x0[k] : fmaf(x0[k], s[l - 1], __ldg(P.b[l - 1] + min(lane + 32 * k, d - 1)) + xs[l - 1][k]);
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        if (c < d) t = fmaf(xs[l][k], __ldg(P.w[l] + c), t);
      }
      s[l] = warp_sum(t);
    }
    float g[NC], g0[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      g0[k] = c < d ? __ldg(go + row * gold + c) : 0.f;        // d/dx through the concat's first half
      g[k] = c < d ? __ldg(go + row * gold + d + c) : 0.f;     // d/dx_L
    }
#pragma unroll
    for (int l = kMaxCrossLayers - 1; l >= 0; --l) {
      if (l >= Lc) continue;
      float ds = 0.f;
#pragma unroll
      for (int k = 0; k < NC; ++k) ds = fmaf(g[k], x0[k], ds);
      ds = warp_sum(ds);
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        gb[l][k] += g[k];
        gw[l][k] = fmaf(ds, xs[l][k], gw[l][k]);
        g0[k] = fmaf(g[k], s[l], g0[k]);
        if (c < d) g[k] = fmaf(ds, __ldg(P.w[l] + c), g[k]);
      }
    }
    if (gx != nullptr) {
#pragma unroll
      for (int k = 0; k < NC; ++k) {
        const int c = lane + 32 * k;
        if (c < d) gx[row * gxld + c] = g0[k] + g[k];
      }
    }
  }
  // block combine in warp order
#pragma unroll
  for (int l = 0; l < kMaxCrossLayers; ++l) {
    if (l >= Lc) break;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const int c = lane + 32 * k;
      if (c < d) {
        sm[((warp * Lc + l) * 2 + 0) * d + c] = gw[l][k];
        sm[((warp * Lc + l) * 2 + 1) * d + c] = gb[l][k];
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Lc * 2 * d; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sm[w * Lc * 2 * d + i];
    partials[(long long)blockIdx.x * Lc * 2 * d + i] = t;
  }
}

// 128-bit backward: lane owns the float4 at columns 4 * (lane + 32 k); the layer count is a template parameter so that
// the per-lane state (x_l chain, parameter-gradient accumulators) is exactly LC deep — the generic kernel above sizes
// every array for 8 layers, which costs registers and occupancy (135 us at B = 65536, d = 112: 13 % of the HBM
// roofline).  Same arithmetic and the same fixed combination order (rows of a warp in order, warps of a block in
// order, blocks in order) => bitwise reproducible.
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }
__device__ __forceinline__ void axpy4(float4& y, float a, const float4& x) {
  y.x = fmaf(a, x.x, y.x); y.y = fmaf(a, x.y, y.y); y.z = fmaf(a, x.z, y.z); y.w = fmaf(a, x.w, y.w);
}
__device__ __forceinline__ void add4(float4& y, const float4& x) { y.x += x.x; y.y += x.y; y.z += x.z; y.w += x.w; }

template <int NV, int LC>
__global__ void __launch_bounds__(256)
dcn_cross_bwd_v4_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ CrossP P,
                        const float* __restrict__ go, long long gold, float* __restrict__ gx, long long gxld,
                        float* __restrict__ partials) {
  extern __shared__ float sm[];  // [8 warps][LC][2][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = P.d;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 gw[LC][NV], gb[LC][NV], w[LC][NV], bb[LC][NV];
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = 4 * (lane + 32 * k);
      gw[l][k] = z4; gb[l][k] = z4;
      w[l][k] = c < d ? __ldg(reinterpret_cast<const float4*>(P.w[l] + c)) : z4;
      bb[l][k] = c < d ? __ldg(reinterpret_cast<const float4*>(P.b[l] + c)) : z4;
    }
  const long long warps_total = (long long)gridDim.x * 8;
  // software pipeline: the three row loads of the NEXT row are in flight while this row's chain is recomputed (one row at
  // a time left every warp idle for a full memory round trip per row: 1.5 TB/s at B = 65536)
  float4 nx0[NV], ng[NV], ng0[NV];
  long long row = (long long)blockIdx.x * 8 + warp;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c = 4 * (lane + 32 * k);
    const bool ok = c < d && row < B;
    nx0[k] = ok ? __ldg(reinterpret_cast<const float4*>(x + row * ld + c)) : z4;
    ng0[k] = ok ? __ldg(reinterpret_cast<const float4*>(go + row * gold + c)) : z4;       // d/dx through the concat's first half
    ng[k] = ok ? __ldg(reinterpret_cast<const float4*>(go + row * gold + d + c)) : z4;    // d/dx_L
  }
  for (; row < B; row += warps_total) {
    float4 x0[NV], xs[LC][NV], g[NV], g0[NV];
    float s[LC];
    const long long nrow = row + warps_total;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      x0[k] = nx0[k]; g0[k] = ng0[k]; g[k] = ng[k];
      const int c = 4 * (lane + 32 * k);
      const bool ok = c < d && nrow < B;
      nx0[k] = ok ? __ldg(reinterpret_cast<const float4*>(x + nrow * ld + c)) : z4;
      ng0[k] = ok ? __ldg(reinterpret_cast<const float4*>(go + nrow * gold + c)) : z4;
      ng[k] = ok ? __ldg(reinterpret_cast<const float4*>(go + nrow * gold + d + c)) : z4;
    }
    // recompute the chain, keeping x_l (input of layer l) and s_l = x_l . w_l
#pragma unroll
    for (int l = 0; l < LC; ++l) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        if (l == 0) xs[l][k] = x0[k];
        else {
          xs[l][k].x = fmaf(x0[k].x, s[l - 1], bb[l - 1][k].x + xs[l - 1][k].x);
          xs[l][k].y = fmaf(x0[k].y, s[l - 1], bb[l - 1][k].y + xs[l - 1][k].y);
          xs[l][k].z = fmaf(x0[k].z, s[l - 1], bb[l - 1][k].z + xs[l - 1][k].z);
          xs[l][k].w = fmaf(x0[k].w, s[l - 1], bb[l - 1][k].w + xs[l - 1][k].w);
        }
        t += dot4(xs[l][k], w[l][k]);
      }
      s[l] = warp_sum(t);
    }
#pragma unroll
    for (int l = LC - 1; l >= 0; --l) {
      float ds = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) ds += dot4(g[k], x0[k]);
      ds = warp_sum(ds);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        add4(gb[l][k], g[k]);
        axpy4(gw[l][k], ds, xs[l][k]);
        axpy4(g0[k], s[l], g[k]);
        axpy4(g[k], ds, w[l][k]);
      }
    }
    if (gx != nullptr) {
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int c = 4 * (lane + 32 * k);
        if (c < d) *reinterpret_cast<float4*>(gx + row * gxld + c) = make_float4(g0[k].x + g[k].x, g0[k].y + g[k].y, g0[k].z + g[k].z, g0[k].w + g[k].w);
      }
    }
  }
  // block combine in warp order
#pragma unroll
  for (int l = 0; l < LC; ++l)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c = 4 * (lane + 32 * k);
      if (c < d) {
        *reinterpret_cast<float4*>(sm + ((warp * LC + l) * 2 + 0) * d + c) = gw[l][k];
        *reinterpret_cast<float4*>(sm + ((warp * LC + l) * 2 + 1) * d + c) = gb[l][k];
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < LC * 2 * d; i += blockDim.x) {
    float t = 0.f;
    for (int wv = 0; wv < 8; ++wv) t += sm[wv * LC * 2 * d + i];
    partials[(long long)blockIdx.x * LC * 2 * d + i] = t;
  }
}

struct CrossG { float* gw[kMaxCrossLayers]; float* gb[kMaxCrossLayers]; };
__global__ void __launch_bounds__(256)
dcn_cross_reduce_kernel(const float* __restrict__ partials, int n_blocks, int Lc, int d, const __grid_constant__ CrossG G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Lc * 2 * d) return;
  float t = 0.f;
  for (int b = 0; b < n_blocks; ++b) t += partials[(long long)b * Lc * 2 * d + i];
  const int l = i / (2 * d), which = (i / d) % 2, c = i % d;
  float* dst = which == 0 ? G.gw[l] : G.gb[l];
  if (dst) dst[c] = t;
}

static int make_cross(int d, int n_layers, const float* const* w, const float* const* b, CrossP* P) {
  NRX_REQUIRE(d >= 1 && d <= 32 * kCrossNC, NRX_EUNSUPPORTED, "cross width %d outside [1,%d]", d, 32 * kCrossNC);
  NRX_REQUIRE(n_layers >= 1 && n_layers <= kMaxCrossLayers, NRX_EINVAL, "cross layers %d outside [1,%d]", n_layers, kMaxCrossLayers);
  NRX_REQUIRE(w && b, NRX_EINVAL, "null parameter arrays");
  memset(P, 0, sizeof(*P));
  P->d = d;
  P->n_layers = n_layers;
  for (int l = 0; l < n_layers; ++l) {
    NRX_REQUIRE(w[l] && b[l], NRX_EINVAL, "cross layer %d: null w/b", l);
    P->w[l] = w[l];
    P->b[l] = b[l];
  }
  return NRX_OK;
}

static int cross_bwd_blocks(long long B) {
  long long blocks = (B + 63) / 64;  // >= 8 rows per warp
  const long long cap = (long long)sm_count() * 3;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace nrx

using namespace nrx;

extern "C" int nrx_dcn_cross_fwd(const float* x, int64_t ld, int64_t B, int d, int n_layers, const float* const* h_w,
                                 const float* const* h_b, float* out, int64_t out_ld, float* dots, nrx_stream_t stream) {
  CrossP P;
  int rc = make_cross(d, n_layers, h_w, h_b, &P);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE((x && out) || B == 0, NRX_EINVAL, "null x / out");
  NRX_REQUIRE(ld >= d && out_ld >= 2 * d, NRX_EINVAL, "leading dimension too small");
  if (B == 0) return NRX_OK;
  const unsigned blocks = (unsigned)((B + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (d <= 128) dcn_cross_fwd_kernel<4><<<blocks, 256, 0, st>>>(x, ld, B, P, out, out_ld, dots);
  else dcn_cross_fwd_kernel<8><<<blocks, 256, 0, st>>>(x, ld, B, P, out, out_ld, dots);
  return check_launch("dcn_cross_fwd");
}

extern "C" size_t nrx_dcn_cross_workspace_bytes(int64_t B, int d, int n_layers) {
  return (size_t)cross_bwd_blocks(B) * n_layers * 2 * d * sizeof(float);
}

extern "C" int nrx_dcn_cross_bwd(const float* x, int64_t ld, int64_t B, int d, int n_layers, const float* const* h_w,
                                 const float* const* h_b, const float* grad_out, int64_t go_ld, const float* dots,
                                 float* grad_x, int64_t gx_ld, float* const* h_grad_w, float* const* h_grad_b, void* ws,
                                 size_t ws_bytes, nrx_stream_t stream) {
  (void)dots;  // the chain is recomputed in registers
  CrossP P;
  int rc = make_cross(d, n_layers, h_w, h_b, &P);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE((x && grad_out) || B == 0, NRX_EINVAL, "null x / grad_out");
  NRX_REQUIRE(ld >= d && go_ld >= 2 * d && (!grad_x || gx_ld >= d), NRX_EINVAL, "leading dimension too small");
  NRX_REQUIRE(h_grad_w && h_grad_b, NRX_EINVAL, "null gradient pointer arrays");
  const int blocks = cross_bwd_blocks(B);
  const size_t need = (size_t)blocks * n_layers * 2 * d * sizeof(float);
  NRX_REQUIRE(ws && ws_bytes >= need, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)8 * n_layers * 2 * d * sizeof(float);
  bool v4 = d % 4 == 0 && ld % 4 == 0 && go_ld % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)grad_out % 16 == 0) && n_layers <= 4 &&
            (!grad_x || (gx_ld % 4 == 0 && (uintptr_t)grad_x % 16 == 0)) && smem <= 48 * 1024;
  for (int l = 0; l < n_layers; ++l) v4 = v4 && ((uintptr_t)h_w[l] % 16 == 0) && ((uintptr_t)h_b[l] % 16 == 0);
  if (v4) {
#define NRX_CROSS_V4(NV, LC) dcn_cross_bwd_v4_kernel<NV, LC><<<blocks, 256, smem, st>>>(x, ld, B, P, grad_out, go_ld, grad_x, gx_ld, (float*)ws)
    if (d <= 128) {
      switch (n_layers) { case 1: NRX_CROSS_V4(1, 1); break; case 2: NRX_CROSS_V4(1, 2); break; case 3: NRX_CROSS_V4(1, 3); break; default: NRX_CROSS_V4(1, 4); }
    } else {
      switch (n_layers) { case 1: NRX_CROSS_V4(2, 1); break; case 2: NRX_CROSS_V4(2, 2); break; case 3: NRX_CROSS_V4(2, 3); break; default: NRX_CROSS_V4(2, 4); }
    }
#undef NRX_CROSS_V4
  } else if (d <= 128) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(dcn_cross_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dcn_cross_bwd_kernel<4><<<blocks, 256, smem, st>>>(x, ld, B, P, grad_out, go_ld, grad_x, gx_ld, (float*)ws);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(dcn_cross_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dcn_cross_bwd_kernel<8><<<blocks, 256, smem, st>>>(x, ld, B, P, grad_out, go_ld, grad_x, gx_ld, (float*)ws);
  }
  rc = check_launch("dcn_cross_bwd");
  if (rc != NRX_OK) return rc;
  CrossG G;
  memset(&G, 0, sizeof(G));
  for (int l = 0; l < n_layers; ++l) { G.gw[l] = h_grad_w[l]; G.gb[l] = h_grad_b[l]; }
  dcn_cross_reduce_kernel<<<(n_layers * 2 * d + 255) / 256, 256, 0, st>>>((const float*)ws, blocks, n_layers, d, G);
  return check_launch("dcn_cross_reduce");
}

// K5 writing the tower's input operand: `image` = bf16 tile image of cat[x, x_L] (width 2d, needs d % 8 == 0 and
// 16-byte aligned rows / parameters); `out` (fp32 concat) is optional.
extern "C" int nrx_dcn_cross_fwd_img(const float* x, int64_t ld, int64_t B, int d, int n_layers, const float* const* h_w,
                                     const float* const* h_b, float* out, int64_t out_ld, void* image, nrx_stream_t stream) {
  CrossP P;
  int rc = make_cross(d, n_layers, h_w, h_b, &P);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE((x && image) || B == 0, NRX_EINVAL, "null x / image");
  NRX_REQUIRE(ld >= d && (!out || out_ld >= 2 * d), NRX_EINVAL, "leading dimension too small");
  bool ok = d % 8 == 0 && d <= 256 && ld % 4 == 0 && ((uintptr_t)x % 16 == 0) && (!out || (out_ld % 4 == 0 && (uintptr_t)out % 16 == 0));
  for (int l = 0; l < n_layers; ++l) ok = ok && ((uintptr_t)h_w[l] % 16 == 0) && ((uintptr_t)h_b[l] % 16 == 0);
  NRX_REQUIRE(ok, NRX_EUNSUPPORTED, "cross image output needs d %% 8 == 0, d <= 256 and 16-byte aligned operands");
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rc = img_zero_tail(image, 2 * d, B, st);
  if (rc != NRX_OK) return rc;
  const unsigned blocks = (unsigned)((B + 7) / 8);
  if (d <= 128) dcn_cross_fwd_v4_kernel<1><<<blocks, 256, 0, st>>>(x, ld, B, P, out, out_ld, (uint8_t*)image, 2 * d);
  else dcn_cross_fwd_v4_kernel<2><<<blocks, 256, 0, st>>>(x, ld, B, P, out, out_ld, (uint8_t*)image, 2 * d);
  return check_launch("dcn_cross_fwd_img");
}
