// K2 — feature-interaction logits, sigmoid+BCE, deterministic reduction, dense AdamW, L2 normalise.
//
//   nrx_field_logit_*  : FM (fm/model.py:18-25,48-59), wide sum (widedeep/model.py:25,58-65),
//                        LR sum (lr/model.py:24-27) on the concatenated feature buffer.
//   nrx_fm_fused_*     : gather-fused FM for sparse-only equal-width fields (never materialises the concat).
//   nrx_logit_loss_fwd : sigmoid + F.binary_cross_entropy on probabilities (deep/model.py:32-33), with
//                        torch's log clamp (-100) and the exact autograd chain for dL/dlogit.
//   nrx_adamw_dense    : torch.optim.AdamW update (deep/model.py:55).
// All HBM/L2-bandwidth bound, warp-per-sample with lanes on contiguous columns.
#include <math.h>

#include "common.cuh"

namespace nrx {

static constexpr int kMaxFields = 32;
struct Fields {
  int col[kMaxFields];
  int dim[kMaxFields];
  int n;
  int mode;
};

// S_d accumulators: lane owns columns d = lane + 32*k, k < 4 (field width <= 128).
__global__ void __launch_bounds__(256)
field_logit_fwd_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ Fields F,
                       float* __restrict__ logit, int accumulate) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* xr = x + b * ld;
  float first = 0.f, second = 0.f;
  if (F.mode == NRX_FIELD_FM) {
    float S[4] = {0.f, 0.f, 0.f, 0.f};
    float sq = 0.f;
    for (int f0 = 0; f0 < F.n; f0 += 8) {  // 8 fields per round: all loads issued before any use
      float v[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int d = lane + 32 * k;
          v[u][k] = (f0 + u < F.n && d < F.dim[f0 + u]) ? __ldg(xr + F.col[f0 + u] + d) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (lane + 32 * k == 0) first += v[u][k];
          else { S[k] += v[u][k]; sq = fmaf(v[u][k], v[u][k], sq); }
        }
    }
    second = 0.5f * (S[0] * S[0] + S[1] * S[1] + S[2] * S[2] + S[3] * S[3] - sq);
  } else if (F.mode == NRX_FIELD_WIDE) {
    for (int f = lane; f < F.n; f += 32) first += __ldg(xr + F.col[f]);
  } else {
    for (int f = 0; f < F.n; ++f)
      for (int d = lane; d < F.dim[f]; d += 32) first += __ldg(xr + F.col[f] + d);
  }
  const float tot = warp_sum(first + second);
  if (lane == 0) logit[b] = accumulate ? logit[b] + tot : tot;
}

__global__ void __launch_bounds__(256)
field_logit_bwd_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ Fields F,
                       const float* __restrict__ dlogit, float* __restrict__ gx, long long gld, int accumulate) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* xr = x + b * ld;
  float* gr = gx + b * gld;
  const float g = __ldg(dlogit + b);
  if (F.mode == NRX_FIELD_FM) {
    float S[4] = {0.f, 0.f, 0.f, 0.f};
    if (F.n <= 8) {  // common case: keep every field's values in registers (one round of loads)
      float v[8][4], o[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int d = lane + 32 * k;
          const bool ok = u < F.n && d < F.dim[u];
          v[u][k] = ok ? __ldg(xr + F.col[u] + d) : 0.f;
          o[u][k] = (ok && accumulate) ? gr[F.col[u] + d] : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (lane + 32 * k > 0) S[k] += v[u][k];
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int d = lane + 32 * k;
          if (u < F.n && d < F.dim[u]) gr[F.col[u] + d] = o[u][k] + ((d == 0) ? g : g * (S[k] - v[u][k]));
        }
      return;
    }
    for (int f = 0; f < F.n; ++f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = lane + 32 * k;
        if (d > 0 && d < F.dim[f]) S[k] += __ldg(xr + F.col[f] + d);
      }
    }
    for (int f = 0; f < F.n; ++f) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int d = lane + 32 * k;
        if (d < F.dim[f]) {
          const float v = __ldg(xr + F.col[f] + d);
          const float gv = (d == 0) ? g : g * (S[k] - v);
          float* p = gr + F.col[f] + d;
          *p = accumulate ? *p + gv : gv;
        }
      }
    }
  } else if (F.mode == NRX_FIELD_WIDE) {
    for (int f = lane; f < F.n; f += 32) {
      float* p = gr + F.col[f];
      *p = accumulate ? *p + g : g;
    }
  } else {
    for (int f = 0; f < F.n; ++f)
      for (int d = lane; d < F.dim[f]; d += 32) {
        float* p = gr + F.col[f] + d;
        *p = accumulate ? *p + g : g;
      }
  }
}

// FM backward, vector form: equal field widths D (D % 4 == 0), 16-byte aligned rows, n * D / 4 <= 32.  One warp per
// sample, lane = one float4 of one field (coalesced 16-byte loads / stores, one round trip); S_d is summed over the fields
// in field order by shuffles (same order as the scalar kernel).  Element 0 of a field is its first-order weight:
// gradient dlogit, excluded from S (fm/model.py:48-59).
__global__ void __launch_bounds__(256)
field_logit_bwd_fm_v4_kernel(const float* __restrict__ x, long long ld, long long B, const __grid_constant__ Fields F, int q,
                             const float* __restrict__ dlogit, float* __restrict__ gx, long long gld, int accumulate) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int f = lane / q, c = lane - f * q;
  const bool live = f < F.n;
  const int col = live ? F.col[f] + 4 * c : 0;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f), o = v;
  if (live) {
    v = __ldg(reinterpret_cast<const float4*>(x + b * ld + col));
    if (accumulate) o = *reinterpret_cast<const float4*>(gx + b * gld + col);
  }
  const float g = __ldg(dlogit + b);
  float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int u = 0; u < F.n; ++u) {
    const int src = c + q * u;
    S.x += __shfl_sync(NRX_FULL_MASK, v.x, src);
    S.y += __shfl_sync(NRX_FULL_MASK, v.y, src);
    S.z += __shfl_sync(NRX_FULL_MASK, v.z, src);
    S.w += __shfl_sync(NRX_FULL_MASK, v.w, src);
  }
  if (!live) return;
  float4 r;
  r.x = o.x + (c == 0 ? g : g * (S.x - v.x));
  r.y = o.y + g * (S.y - v.y);
  r.z = o.z + g * (S.z - v.z);
  r.w = o.w + g * (S.w - v.w);
  *reinterpret_cast<float4*>(gx + b * gld + col) = r;
}

static int make_fields(const int32_t* cols, const int32_t* dims, int n, int mode, long long ld, Fields* F) {
  NRX_REQUIRE(cols && dims && n > 0 && n <= kMaxFields, NRX_EINVAL, "n_fields=%d outside [1,%d]", n, kMaxFields);
  NRX_REQUIRE(mode == NRX_FIELD_FM || mode == NRX_FIELD_WIDE || mode == NRX_FIELD_SUM, NRX_EINVAL, "bad field mode");
  F->n = n;
  F->mode = mode;
  for (int i = 0; i < n; ++i) {
    NRX_REQUIRE(cols[i] >= 0 && dims[i] >= 1 && cols[i] + dims[i] <= ld, NRX_EINVAL, "field %d outside the row", i);
    if (mode == NRX_FIELD_FM) {
      NRX_REQUIRE(dims[i] == dims[0], NRX_EINVAL, "FM needs equal field widths (fm/model.py:58 torch.stack)");
      NRX_REQUIRE(dims[i] <= 128, NRX_EUNSUPPORTED, "FM field width %d > 128", dims[i]);
    }
    F->col[i] = cols[i];
    F->dim[i] = dims[i];
  }
  return NRX_OK;
}

// ---- gather-fused FM (vec4 path: D % 4 == 0, D/4 a power of two <= 32) ---------------------
template <int SPW, bool BWD>
__global__ void __launch_bounds__(256)
fm_fused_kernel(const __grid_constant__ DFeats P, long long B, int LPF, const float* __restrict__ bias,
                const float* __restrict__ label, long long lstride, float* __restrict__ logit,
                float* __restrict__ prob, float* __restrict__ loss, float* __restrict__ dlogit,
                const float* __restrict__ dlogit_in, float* __restrict__ gx, long long gld,
                int* __restrict__ status) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long b0 = warp * SPW;
  if (b0 >= B) return;
  const int C = P.n * LPF;             // 16-byte columns per sample
  const int off4 = (lane % LPF) * 4;   // same for every pass because 32 % LPF == 0
  float4 S[SPW];
  float first[SPW], sq[SPW];
#pragma unroll
  for (int s = 0; s < SPW; ++s) { S[s] = make_float4(0.f, 0.f, 0.f, 0.f); first[s] = 0.f; sq[s] = 0.f; }
  bool bad = false;
  for (int c = lane; c < ((C + 31) & ~31); c += 32) {
    const bool act = c < C;
    const DFeat& F = P.f[act ? c / LPF : 0];
    long long id[SPW];
#pragma unroll
    for (int s = 0; s < SPW; ++s) id[s] = (act && b0 + s < B) ? load_idx(F.idx, b0 + s, F.idx32) : 0;
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (act && b0 + s < B) {
        if ((unsigned long long)id[s] < (unsigned long long)F.rows) v = __ldg(reinterpret_cast<const float4*>(F.table + id[s] * F.stride + off4));
        else bad = true;
      }
      if (off4 == 0) { first[s] += v.x; v.x = 0.f; }   // column 0 is the first-order weight
      S[s].x += v.x; S[s].y += v.y; S[s].z += v.z; S[s].w += v.w;
      sq[s] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  }
  // sum S over fields: lanes with equal lane % LPF
#pragma unroll
  for (int s = 0; s < SPW; ++s)
    for (int o = LPF; o < 32; o <<= 1) {
      S[s].x += __shfl_xor_sync(NRX_FULL_MASK, S[s].x, o); S[s].y += __shfl_xor_sync(NRX_FULL_MASK, S[s].y, o);
      S[s].z += __shfl_xor_sync(NRX_FULL_MASK, S[s].z, o); S[s].w += __shfl_xor_sync(NRX_FULL_MASK, S[s].w, o);
    }
  if (!BWD) {
    const float bz = bias ? __ldg(bias) : 0.f;
    const float invB = 1.f / (float)B;
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const float ss = (lane < LPF) ? (S[s].x * S[s].x + S[s].y * S[s].y + S[s].z * S[s].z + S[s].w * S[s].w) : 0.f;
      const float z = bz + warp_sum(first[s] + 0.5f * (ss - sq[s]));
      const long long b = b0 + s;
      if (lane == 0 && b < B) {
        if (logit) logit[b] = z;
        const float p = sigmoid_f(z);
        if (prob) prob[b] = p;
        if (label) bce_terms(p, __ldg(label + b * lstride), invB, loss ? loss + b : nullptr, dlogit ? dlogit + b : nullptr);
      }
    }
    if (bad && status) atomicOr(status, 1);
  } else {
    // second pass: re-read the rows (L1/L2 hits) and emit grad_x
    for (int c = lane; c < C; c += 32) {
      const DFeat& F = P.f[c / LPF];
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        const long long b = b0 + s;
        if (b >= B) continue;
        const long long id = load_idx(F.idx, b, F.idx32);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned long long)id < (unsigned long long)F.rows) v = __ldg(reinterpret_cast<const float4*>(F.table + id * F.stride + off4));
        const float g = __ldg(dlogit_in + b);
        float4 o;
        o.x = (off4 == 0) ? g : g * (S[s].x - v.x);
        o.y = g * (S[s].y - v.y); o.z = g * (S[s].z - v.z); o.w = g * (S[s].w - v.w);
        *reinterpret_cast<float4*>(gx + b * gld + F.out_col + off4) = o;
      }
    }
  }
}

// ---- gather-fused FM, one THREAD per sample (field width D = 4 * NQ <= 32) -----------------------------------------
// The warp-cooperative kernel above spends ~200 warp instructions per sample (shuffle trees, divergent descriptor reads,
// 12 of 32 lanes idle at 5 fields x 4 columns): instruction-bound at 0.19 of the HBM peak.  Here a thread keeps S_d of its
// sample in registers: ids are read coalesced across the warp, every row is four independent 16-byte loads (both
// sectors of the 64-byte row used), no shuffles, field descriptors are warp-uniform constant reads: ~7 warp instructions
// per sample.  Backward: the same gather again (L1 / L2 hits), then grad_x rows as 16-byte stores.
template <int NQ, bool BWD>
__global__ void __launch_bounds__(128)
fm_fused_tps_kernel(const __grid_constant__ DFeats P, long long B, const float* __restrict__ bias,
                    const float* __restrict__ label, long long lstride, float* __restrict__ logit,
                    float* __restrict__ prob, float* __restrict__ loss, float* __restrict__ dlogit,
                    const float* __restrict__ dlogit_in, float* __restrict__ gx, long long gld, int* __restrict__ status,
                    int stage_w4, int stage_col0) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = b < B;
  if (!live && !(BWD && stage_w4 > 0)) return;   // staged stores are warp-cooperative: tail threads stay for the copy-out
  float4 S[NQ];
#pragma unroll
  for (int k = 0; k < NQ; ++k) S[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  float first = 0.f, sq = 0.f;
  bool bad = false;
  for (int f = 0; f < P.n; ++f) {
    const DFeat& F = P.f[f];
    const long long id = live ? load_idx(F.idx, b, F.idx32) : 0;
    const bool ok = (unsigned long long)id < (unsigned long long)F.rows;
    bad |= !ok;
    const float4* row = reinterpret_cast<const float4*>(F.table + (ok ? id : 0) * F.stride);
    float4 v[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k) v[k] = ok ? __ldg(row + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (BWD && stage_w4 > 0) {   // keep the rows for the second pass (no second gather)
      extern __shared__ __align__(16) float4 stage[];
      float4* keep = stage + (size_t)threadIdx.x * (stage_w4 + 1) + (F.out_col - stage_col0) / 4;
#pragma unroll
      for (int k = 0; k < NQ; ++k) keep[k] = v[k];
    }
    first += v[0].x;
    v[0].x = 0.f;   // column 0 is the first-order weight
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      S[k].x += v[k].x; S[k].y += v[k].y; S[k].z += v[k].z; S[k].w += v[k].w;
      sq += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
    }
  }
  if (!BWD) {
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NQ; ++k) ss += S[k].x * S[k].x + S[k].y * S[k].y + S[k].z * S[k].z + S[k].w * S[k].w;
    const float z = (bias ? __ldg(bias) : 0.f) + first + 0.5f * (ss - sq);
    if (logit) logit[b] = z;
    const float p = sigmoid_f(z);
    if (prob) prob[b] = p;
    if (label) bce_terms(p, __ldg(label + b * lstride), 1.f / (float)B, loss ? loss + b : nullptr, dlogit ? dlogit + b : nullptr);
    if (bad && status) atomicOr(status, 1);
  } else {
    // grad_x rows: written through shared memory when the fields tile one contiguous column range (stage_w4 > 0), so that
    // a warp stores its 32 rows as whole 16-byte-per-lane coalesced lines instead of 32 scattered half sectors per store
    extern __shared__ __align__(16) float4 stage[];   // [128][stage_w4 + 1]
    const float g = live ? __ldg(dlogit_in + b) : 0.f;
    for (int f = 0; f < P.n; ++f) {
      const DFeat& F = P.f[f];
      float4* dst = stage_w4 > 0 ? stage + (size_t)threadIdx.x * (stage_w4 + 1) + (F.out_col - stage_col0) / 4
                                 : reinterpret_cast<float4*>(gx + b * gld + F.out_col);
      bool ok = false;
      const float4* row = nullptr;
      if (stage_w4 == 0) {   // no staging buffer: gather the row again (L1 / L2 hit)
        const long long id = live ? load_idx(F.idx, b, F.idx32) : 0;
        ok = live && (unsigned long long)id < (unsigned long long)F.rows;
        row = reinterpret_cast<const float4*>(F.table + (ok ? id : 0) * F.stride);
      }
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        float4 v = stage_w4 > 0 ? dst[k] : (ok ? __ldg(row + k) : make_float4(0.f, 0.f, 0.f, 0.f));
        if (k == 0) v.x = 0.f;   // (unused below: column 0 takes g)
        float4 o;
        o.x = (k == 0) ? g : g * (S[k].x - v.x);
        o.y = g * (S[k].y - v.y); o.z = g * (S[k].z - v.z); o.w = g * (S[k].w - v.w);
        if (live || stage_w4 > 0) dst[k] = o;
      }
    }
    if (stage_w4 > 0) {
      __syncwarp();
      const int lane = threadIdx.x & 31, w0 = threadIdx.x & ~31;      // this warp's 32 rows
      const long long bw = (long long)blockIdx.x * blockDim.x + w0;
      for (int i = lane; i < 32 * stage_w4; i += 32) {
        const int r = i / stage_w4, c = i - r * stage_w4;
        if (bw + r < B)
          *reinterpret_cast<float4*>(gx + (bw + r) * gld + stage_col0 + 4 * c) = stage[(size_t)(w0 + r) * (stage_w4 + 1) + c];
      }
    }
  }
}

template <bool BWD>
static bool launch_fm_tps(const DFeats& d, long long B, int LPF, const float* bias, const float* label, long long lstride,
                          float* logit, float* prob, float* loss, float* dlogit, const float* dlogit_in, float* gx, long long gld,
                          int* status, cudaStream_t st) {
  if (getenv("NRX_FM_WARP") != nullptr) return false;   // keep the warp-cooperative kernel reachable for A/B runs
  const unsigned blocks = (unsigned)((B + 127) / 128);
  // backward: do the fields tile ONE contiguous column range of grad_x?  then rows go out through shared memory
  int w4 = 0, col0 = 0;
  size_t smem = 0;
  if (BWD) {
    int lo = d.f[0].out_col, hi = lo;
    bool used[NRX_MAX_FEATS * 32] = {false};
    for (int i = 0; i < d.n; ++i) lo = d.f[i].out_col < lo ? d.f[i].out_col : lo;
    bool tiled = true;
    for (int i = 0; i < d.n && tiled; ++i) {
      const int q = (d.f[i].out_col - lo) / (4 * LPF);
      tiled = (d.f[i].out_col - lo) % (4 * LPF) == 0 && q < d.n && !used[q];
      if (tiled) used[q] = true;
      hi = d.f[i].out_col + 4 * LPF > hi ? d.f[i].out_col + 4 * LPF : hi;
    }
    if (tiled && hi - lo == d.n * 4 * LPF && (size_t)128 * (d.n * LPF + 1) * 16 <= 96 * 1024) {
      w4 = d.n * LPF; col0 = lo;
      smem = (size_t)128 * (w4 + 1) * 16;
    }
  }
#define NRX_FM_TPS(NQ) do { if (smem > 48 * 1024) cudaFuncSetAttribute(fm_fused_tps_kernel<NQ, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    fm_fused_tps_kernel<NQ, BWD><<<blocks, 128, smem, st>>>(d, B, bias, label, lstride, logit, prob, loss, dlogit, dlogit_in, gx, gld, status, w4, col0); } while (0)
  switch (LPF) {
    case 1: NRX_FM_TPS(1); return true;
    case 2: NRX_FM_TPS(2); return true;
    case 4: NRX_FM_TPS(4); return true;
    case 8: NRX_FM_TPS(8); return true;
    default: return false;
  }
#undef NRX_FM_TPS
}

static int fm_fused_check(const DFeats& d, int* LPF) {
  NRX_REQUIRE(d.n_array == 0, NRX_EUNSUPPORTED, "fm_fused: array features go through embed_pool + field_logit");
  const int D = d.f[0].dim;
  for (int i = 0; i < d.n; ++i) {
    NRX_REQUIRE(d.f[i].dim == D, NRX_EINVAL, "FM needs equal field widths (fm/model.py:58)");
    NRX_REQUIRE((uintptr_t)d.f[i].table % 16 == 0 && d.f[i].stride % 4 == 0, NRX_EUNSUPPORTED, "fm_fused: table %d not 16B aligned", i);
  }
  NRX_REQUIRE(D % 4 == 0 && D <= 128 && (D / 4 & (D / 4 - 1)) == 0, NRX_EUNSUPPORTED,
              "fm_fused: width %d (needs D/4 a power of two <= 32)", D);
  *LPF = D / 4;
  return NRX_OK;
}

// ---- sigmoid + BCE ----------------------------------------------------------------------------
struct Terms { const float* t[8]; int n; };
__global__ void __launch_bounds__(256)
logit_loss_kernel(const __grid_constant__ Terms T, const float* __restrict__ bias, long long B,
                  const float* __restrict__ label, long long lstride, float* __restrict__ prob,
                  float* __restrict__ loss, float* __restrict__ dlogit) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float z = 0.f;
  for (int i = 0; i < T.n; ++i) z += __ldg(T.t[i] + b);
  if (bias) z += __ldg(bias);
  const float p = sigmoid_f(z);
  if (prob) prob[b] = p;
  if (label) bce_terms(p, __ldg(label + b * lstride), 1.f / (float)B, loss ? loss + b : nullptr, dlogit ? dlogit + b : nullptr);
}

__global__ void __launch_bounds__(256)
bce_fwd_kernel(const float* __restrict__ prob, const float* __restrict__ label, long long lstride, long long B,
               float* __restrict__ loss) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) bce_terms(__ldg(prob + b), __ldg(label + b * lstride), 0.f, loss + b, nullptr);
}
__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ label, long long lstride, long long B,
               const float* __restrict__ upstream, float* __restrict__ gp) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float p = __ldg(prob + b), y = __ldg(label + b * lstride);
  gp[b] = ((p - y) / fmaxf(p * (1.f - p), 1e-12f)) * (__ldg(upstream) / (float)B);
}
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ gp, long long B, float* __restrict__ gz) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float p = __ldg(prob + b);
  gz[b] = __ldg(gp + b) * (p * (1.f - p));
}

// ---- deterministic reduction (fixed tree: 1 block, 1024 threads) -------------------------------
__global__ void __launch_bounds__(1024)
reduce_kernel(const float* __restrict__ x, long long n, float scale, float* __restrict__ out) {
  __shared__ float sm[32];
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) a += __ldg(x + i);
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = warp_sum(sm[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = a * scale;
  }
}

// ---- dense AdamW -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2s,
             const float* __restrict__ d_hp) {
  if (d_hp) { lr = __ldg(d_hp); bc1 = __ldg(d_hp + 1); bc2s = __ldg(d_hp + 2); }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    adamw_update(pi, g[i], mi, vi, lr, bc1, bc2s, b1, b2, eps, wd);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

// float4 flavour: n4 = n / 4, all four buffers 16-byte aligned (the trainers' flat buffers always are)
__global__ void __launch_bounds__(256)
adamw_kernel_v4(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                long long n4, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2s,
                const float* __restrict__ d_hp) {
  if (d_hp) { lr = __ldg(d_hp); bc1 = __ldg(d_hp + 1); bc2s = __ldg(d_hp + 2); }
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pi = p[i], mi = m[i], vi = v[i];
    const float4 gi = g[i];
    adamw_update(pi.x, gi.x, mi.x, vi.x, lr, bc1, bc2s, b1, b2, eps, wd);
    adamw_update(pi.y, gi.y, mi.y, vi.y, lr, bc1, bc2s, b1, b2, eps, wd);
    adamw_update(pi.z, gi.z, mi.z, vi.z, lr, bc1, bc2s, b1, b2, eps, wd);
    adamw_update(pi.w, gi.w, mi.w, vi.w, lr, bc1, bc2s, b1, b2, eps, wd);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

static void launch_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps,
                         float wd, float bc1, float bc2s, const float* d_hp, cudaStream_t st) {
  const bool v4 = n % 4 == 0 && ((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16 == 0;
  const long long items = v4 ? n / 4 : n;
  long long blocks = (items + 255) / 256;
  const long long cap = 8LL * sm_count();
  if (blocks > cap) blocks = cap;
  if (v4)
    adamw_kernel_v4<<<(unsigned)blocks, 256, 0, st>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, items, lr, b1, b2,
                                                      eps, wd, bc1, bc2s, d_hp);
  else
    adamw_kernel<<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, items, lr, b1, b2, eps, wd, bc1, bc2s, d_hp);
}

// One thread: ++step, then hp = {lr(step-1) per CosinDecayLR (lr_schedule.py:16-28), 1-b1^step, sqrt(1-b2^step)}.
__global__ void hparams_step_kernel(int* __restrict__ step, float* __restrict__ hp, float lr0, float lr1, int m0, int m1,
                                    float b1, float b2) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int s = step[0] + 1;  // 1-based optimizer step
  step[0] = s;
  const int e = s - 1;        // scheduler epoch used for this step
  double lr;
  if (e < m0) lr = lr0;
  else if (e >= m1) lr = lr1;
  else {
    const double t = (double)(e - m0) / (double)max(1, m1 - m0);
    lr = (double)lr1 + ((double)lr0 - (double)lr1) * 0.5 * (1.0 + cos(3.14159265358979323846 * t));
  }
  hp[0] = (float)lr;
  hp[1] = (float)(1.0 - pow((double)b1, (double)s));
  hp[2] = (float)sqrt(1.0 - pow((double)b2, (double)s));
}

__global__ void __launch_bounds__(256)
l2norm_kernel(const float* __restrict__ x, long long ld, long long n, int d, float* __restrict__ y, long long yld) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  float ss = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = __ldg(x + r * ld + c); ss = fmaf(t, t, ss); }
  ss = warp_sum(ss);
  const float den = fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps (DSSM/model.py:69-71)
  for (int c = lane; c < d; c += 32) y[r * yld + c] = __ldg(x + r * ld + c) / den;
}

static inline unsigned warps_grid(long long items, int wpb) { return (unsigned)((items + wpb - 1) / wpb); }

}  // namespace nrx

using namespace nrx;

extern "C" int nrx_field_logit_fwd(const float* x, int64_t ld, int64_t B, const int32_t* h_cols, const int32_t* h_dims,
                                   int n_fields, int mode, float* logit, int accumulate, nrx_stream_t stream) {
  Fields F;
  int rc = make_fields(h_cols, h_dims, n_fields, mode, ld, &F);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(x && logit, NRX_EINVAL, "null pointer");
  if (B == 0) return NRX_OK;
  field_logit_fwd_kernel<<<warps_grid(B, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, B, F, logit, accumulate);
  return check_launch("field_logit_fwd");
}

extern "C" int nrx_field_logit_bwd(const float* x, int64_t ld, int64_t B, const int32_t* h_cols, const int32_t* h_dims,
                                   int n_fields, int mode, const float* dlogit, float* grad_x, int64_t grad_ld,
                                   int accumulate, nrx_stream_t stream) {
  Fields F;
  int rc = make_fields(h_cols, h_dims, n_fields, mode, ld, &F);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(x && dlogit && grad_x, NRX_EINVAL, "null pointer");
  if (B == 0) return NRX_OK;
  if (mode == NRX_FIELD_FM) {
    const int D = F.dim[0];
    bool v4 = D % 4 == 0 && (D / 4) * n_fields <= 32 && ld % 4 == 0 && grad_ld % 4 == 0 &&
              ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(grad_x)) & 15) == 0;
    for (int i = 0; i < n_fields && v4; ++i) v4 = F.col[i] % 4 == 0;
    if (v4) {
      field_logit_bwd_fm_v4_kernel<<<warps_grid(B, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, B, F, D / 4, dlogit, grad_x, grad_ld,
                                                                                       accumulate);
      return check_launch("field_logit_bwd(fm v4)");
    }
  }
  field_logit_bwd_kernel<<<warps_grid(B, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, B, F, dlogit, grad_x, grad_ld, accumulate);
  return check_launch("field_logit_bwd");
}

extern "C" int nrx_fm_fused_fwd(const NrxFeat* h_feats, int n_feats, int64_t B, const float* bias, const float* label,
                                int64_t label_stride, float* logit, float* prob, float* loss_per_sample, float* dlogit,
                                int32_t* status, nrx_stream_t stream) {
  DFeats d;
  int rc = make_dfeats(h_feats, n_feats, B, nullptr, 0, &d);
  if (rc != NRX_OK) return rc;
  int LPF = 0;
  rc = fm_fused_check(d, &LPF);
  if (rc != NRX_OK) return rc;
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_fm_tps<false>(d, B, LPF, bias, label, label_stride, logit, prob, loss_per_sample, dlogit, nullptr, nullptr, 0, status, st))
    return check_launch("fm_fused_fwd");
  if (B < (long long)sm_count() * 64)
    fm_fused_kernel<2, false><<<warps_grid((B + 1) / 2, 8), 256, 0, st>>>(d, B, LPF, bias, label, label_stride, logit, prob,
                                                                        loss_per_sample, dlogit, nullptr, nullptr, 0, status);
  else
    fm_fused_kernel<4, false><<<warps_grid((B + 3) / 4, 8), 256, 0, st>>>(d, B, LPF, bias, label, label_stride, logit, prob,
                                                                        loss_per_sample, dlogit, nullptr, nullptr, 0, status);
  return check_launch("fm_fused_fwd");
}

extern "C" int nrx_fm_fused_bwd(const NrxFeat* h_feats, int n_feats, int64_t B, const float* dlogit, float* grad_x,
                                int64_t grad_ld, nrx_stream_t stream) {
  DFeats d;
  int rc = make_dfeats(h_feats, n_feats, B, grad_x, grad_ld, &d);
  if (rc != NRX_OK) return rc;
  int LPF = 0;
  rc = fm_fused_check(d, &LPF);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(dlogit && grad_x, NRX_EINVAL, "null pointer");
  NRX_REQUIRE(d.vec == 4, NRX_EUNSUPPORTED, "fm_fused_bwd: grad_x must be 16B aligned with ld %% 4 == 0 and out_col %% 4 == 0");
  for (int i = 0; i < d.n; ++i) NRX_REQUIRE(d.f[i].out_col + d.f[i].dim <= grad_ld, NRX_EINVAL, "feature %d overruns grad_ld", i);
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_fm_tps<true>(d, B, LPF, nullptr, nullptr, 0, nullptr, nullptr, nullptr, nullptr, dlogit, grad_x, grad_ld, nullptr, st))
    return check_launch("fm_fused_bwd");
  if (B < (long long)sm_count() * 64)
    fm_fused_kernel<2, true><<<warps_grid((B + 1) / 2, 8), 256, 0, st>>>(d, B, LPF, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                                                                       nullptr, dlogit, grad_x, grad_ld, nullptr);
  else
    fm_fused_kernel<4, true><<<warps_grid((B + 3) / 4, 8), 256, 0, st>>>(d, B, LPF, nullptr, nullptr, 0, nullptr, nullptr, nullptr,
                                                                       nullptr, dlogit, grad_x, grad_ld, nullptr);
  return check_launch("fm_fused_bwd");
}

extern "C" int nrx_logit_loss_fwd(const float* const* h_terms, int n_terms, const float* bias, int64_t B, const float* label,
                                  int64_t label_stride, float* prob, float* loss_per_sample, float* dlogit,
                                  nrx_stream_t stream) {
  NRX_REQUIRE(n_terms >= 0 && n_terms <= 8 && (n_terms == 0 || h_terms), NRX_EINVAL, "n_terms=%d outside [0,8]", n_terms);
  Terms T;
  T.n = n_terms;
  for (int i = 0; i < n_terms; ++i) {
    NRX_REQUIRE(h_terms[i], NRX_EINVAL, "null term %d", i);
    T.t[i] = h_terms[i];
  }
  if (B == 0) return NRX_OK;
  logit_loss_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, bias, B, label, label_stride, prob,
                                                                                  loss_per_sample, dlogit);
  return check_launch("logit_loss_fwd");
}

extern "C" int nrx_bce_fwd(const float* prob, const float* label, int64_t label_stride, int64_t B, float* loss_per_sample,
                           nrx_stream_t stream) {
  NRX_REQUIRE(prob && label && loss_per_sample, NRX_EINVAL, "null pointer");
  if (B == 0) return NRX_OK;
  bce_fwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prob, label, label_stride, B, loss_per_sample);
  return check_launch("bce_fwd");
}

extern "C" int nrx_bce_bwd(const float* prob, const float* label, int64_t label_stride, int64_t B, const float* upstream,
                           float* grad_prob, nrx_stream_t stream) {
  NRX_REQUIRE(prob && label && upstream && grad_prob, NRX_EINVAL, "null pointer");
  if (B == 0) return NRX_OK;
  bce_bwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prob, label, label_stride, B, upstream, grad_prob);
  return check_launch("bce_bwd");
}

extern "C" int nrx_sigmoid_bwd(const float* prob, const float* grad_prob, int64_t B, float* grad_logit, nrx_stream_t stream) {
  NRX_REQUIRE(prob && grad_prob && grad_logit, NRX_EINVAL, "null pointer");
  if (B == 0) return NRX_OK;
  sigmoid_bwd_kernel<<<(unsigned)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(prob, grad_prob, B, grad_logit);
  return check_launch("sigmoid_bwd");
}

__global__ void __launch_bounds__(1024)
reduce2_kernel(const float* __restrict__ x1, long long n1, float s1, float* __restrict__ o1, const float* __restrict__ x2,
               long long n2, float s2, float* __restrict__ o2) {
  __shared__ float sm[32];
  const float* x = blockIdx.x == 0 ? x1 : x2;
  const long long n = blockIdx.x == 0 ? n1 : n2;
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += 1024) a += __ldg(x + i);
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    a = warp_sum(sm[threadIdx.x]);
    if (threadIdx.x == 0) { if (blockIdx.x == 0) o1[0] = a * s1; else o2[0] = a * s2; }
  }
}

extern "C" int nrx_reduce2_f32(const float* x1, int64_t n1, float scale1, float* out1, const float* x2, int64_t n2,
                               float scale2, float* out2, nrx_stream_t stream) {
  NRX_REQUIRE(out1 && out2 && x1 && x2 && n1 >= 0 && n2 >= 0, NRX_EINVAL, "null pointer");
  reduce2_kernel<<<2, 1024, 0, (cudaStream_t)stream>>>(x1, n1, scale1, out1, x2, n2, scale2, out2);
  return check_launch("reduce2_f32");
}

extern "C" int nrx_reduce_f32(const float* x, int64_t n, float scale, float* out, nrx_stream_t stream) {
  NRX_REQUIRE(out && (x || n == 0) && n >= 0, NRX_EINVAL, "null pointer");
  reduce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, scale, out);
  return check_launch("reduce_f32");
}

extern "C" int nrx_adamw_dense(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                               float eps, float weight_decay, int32_t step, nrx_stream_t stream) {
  NRX_REQUIRE(p && g && m && v && n >= 0 && step >= 1, NRX_EINVAL, "bad AdamW arguments");
  if (n == 0) return NRX_OK;
  const float bc1 = (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  launch_adamw(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, nullptr, (cudaStream_t)stream);
  return check_launch("adamw_dense");
}

extern "C" int nrx_adamw_dense_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* d_hparams, float beta1,
                                   float beta2, float eps, float weight_decay, nrx_stream_t stream) {
  NRX_REQUIRE(p && g && m && v && d_hparams && n >= 0, NRX_EINVAL, "bad AdamW arguments");
  if (n == 0) return NRX_OK;
  launch_adamw(p, g, m, v, n, 0.f, beta1, beta2, eps, weight_decay, 1.f, 1.f, d_hparams, (cudaStream_t)stream);
  return check_launch("adamw_dense_dev");
}

extern "C" int nrx_hparams_step(int32_t* d_step, float* d_hparams, float lr, float min_lr, int32_t milestone0,
                                int32_t milestone1, float beta1, float beta2, nrx_stream_t stream) {
  NRX_REQUIRE(d_step && d_hparams, NRX_EINVAL, "null pointer");
  hparams_step_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_step, d_hparams, lr, min_lr, milestone0, milestone1, beta1, beta2);
  return check_launch("hparams_step");
}

extern "C" int nrx_l2_normalize(const float* x, int64_t ld, int64_t n, int d, float* y, int64_t y_ld, nrx_stream_t stream) {
  NRX_REQUIRE(x && y && d > 0 && ld >= d && y_ld >= d, NRX_EINVAL, "bad l2_normalize arguments");
  if (n == 0) return NRX_OK;
  l2norm_kernel<<<warps_grid(n, 8), 256, 0, (cudaStream_t)stream>>>(x, ld, n, d, y, y_ld);
  return check_launch("l2_normalize");
}
