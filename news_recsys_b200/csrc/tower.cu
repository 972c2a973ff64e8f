// K4 — fused bf16 MLP tower on tcgen05 (5th-gen tensor cores, accumulators in TMEM).
//
// Replaces the nn.Linear/ReLU stacks of the reference: `MLP` (src/model/model_utils/utils.py:6-17) used by
// Deep / WideDeep / DCN (sort/deep/model.py:12-21, sort/widedeep/model.py:14-27, sort/dcn/model.py:15-29)
// and the LeakyReLU(0.2) towers of DSSM (recall/DSSM/model.py:26-44), forward and backward.
//
// Design (one CTA per SM, persistent over 128-row tiles):
//   * every layer's weights are packed once per step to bf16 in the canonical no-swizzle K-major UMMA
//     layout (umma.cuh) and stay resident in shared memory (bulk async copies, SASS UBLKCP);
//   * a row tile's activations never leave the SM: layer l's accumulator [128 x N_l] lives in TMEM, the
//     epilogue (tcgen05.ld -> +bias -> ReLU -> bf16) writes the next layer's A operand straight back to
//     shared memory, so HBM sees x once and y once (+ bf16 activation images when training);
//   * a final layer with <= 4 outputs (the logit) is a register dot product in fp32 — no padded MMA;
//   * backward = (1) dX chain, same structure with W^T images, dz_l = (dz_{l+1} W_{l+1}) * act'(a_l);
//     (2) dW_l = dz_l^T a_{l-1} as split-K tcgen05 GEMMs over the batch that read the SAME tile images
//     as MN-major operands (no transposes are ever materialised), db_l rides along as a ones column;
//     (3) fixed-order reduction of the per-CTA partials (bitwise reproducible).
// Tensor-bound when fused (AI >> ridge); flops per sample in DESIGN.md.
#include <math.h>

#include <stdlib.h>

#include "tower.cuh"

namespace nrx {
using namespace umma;

static int r16(int x) { return (x + 15) & ~15; }

int make_tower(const NrxTower* t, long long B, int training, TowerK* k) {
  NRX_REQUIRE(t != nullptr, NRX_EINVAL, "null tower");
  NRX_REQUIRE(t->n_layers >= 1 && t->n_layers <= NRX_MAX_LAYERS, NRX_EINVAL, "n_layers=%d outside [1,%d]", t->n_layers,
              (int)NRX_MAX_LAYERS);
  NRX_REQUIRE(t->act == NRX_ACT_RELU || t->act == NRX_ACT_LEAKY, NRX_EINVAL, "bad activation");
  memset(k, 0, sizeof(*k));
  k->n_layers = t->n_layers;
  k->act = t->act;
  k->slope = t->act == NRX_ACT_LEAKY ? t->negative_slope : 0.f;
  const int last = t->n_layers - 1;
  k->tiny = (t->dims[last + 1] <= kMaxTiny && t->n_layers >= 2) ? 1 : 0;
  k->n_mma = t->n_layers - k->tiny;
  int max_np = 16;
  for (int l = 0; l < t->n_layers; ++l) {
    NRX_REQUIRE(t->dims[l] >= 1 && t->dims[l + 1] >= 1, NRX_EINVAL, "layer %d: bad dims", l);
    NRX_REQUIRE(t->w[l] && t->b[l], NRX_EINVAL, "layer %d: null weight/bias", l);
    k->K[l] = t->dims[l];
    k->N[l] = t->dims[l + 1];
    k->Kp[l] = r16(k->K[l]);
    k->Np[l] = r16(k->N[l]);
    k->w[l] = t->w[l];
    k->bias[l] = t->b[l];
    // the first layer is K-streamed by the pipelined forward (tower_fwd.cu): any input width whose weights fit in SMEM
    NRX_REQUIRE((l == 0 ? k->Kp[l] <= 1024 : k->Kp[l] <= 240) && k->Np[l] <= 240, NRX_EUNSUPPORTED,
                "layer %d: widths up to 240 supported (got %d -> %d)", l, k->K[l], k->N[l]);
    if (l == 0 && k->Kp[l] > 240) k->wide0 = 1;
    NRX_REQUIRE(!training || k->Np[l] <= 128, NRX_EUNSUPPORTED, "layer %d: training supports out widths up to 128 (got %d)", l, k->N[l]);
    if (k->Kp[l] > k->max_kp) k->max_kp = k->Kp[l];
    if (l < k->n_mma && k->Np[l] > max_np) max_np = k->Np[l];
    if (l < k->n_mma && k->Kp[l] > max_np && k->Kp[l] <= 240) max_np = k->Kp[l];  // dX accumulators are Kp wide
  }
  for (int l = 0; l + 1 < t->n_layers; ++l)
    NRX_REQUIRE(k->Np[l] == k->Kp[l + 1], NRX_EINVAL, "layer %d out (%d) != layer %d in (%d)", l, k->N[l], l + 1, k->K[l + 1]);
  k->tmem_cols = max_np <= 32 ? 32 : max_np <= 64 ? 64 : max_np <= 128 ? 128 : 256;
  k->n_tiles = (B + kRows - 1) / kRows;
  // packed weight images (MMA layers only)
  unsigned off = 0;
  for (int l = 0; l < k->n_mma; ++l) { k->w_off[l] = off; off += (unsigned)k->Kp[l] * k->Np[l] * 2; }
  k->w_bytes = off;
  off = 0;
  for (int l = 0; l < k->n_mma; ++l) { k->wt_off[l] = off; off += (unsigned)k->Kp[l] * k->Np[l] * 2; }
  k->wt_bytes = off;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  k->wpack_off = (long long)o;  o += al(k->w_bytes);
  k->wtpack_off = (long long)o; o += al(k->wt_bytes);
  // the input image a_0 exists in inference too: it is what the pipelined forward streams
  k->act_off[0] = (long long)o; o += al((size_t)k->n_tiles * k->Kp[0] * kRows * 2);
  if (training) {
    for (int l = 1; l < k->n_layers; ++l) { k->act_off[l] = (long long)o; o += al((size_t)k->n_tiles * k->Kp[l] * kRows * 2); }
    for (int l = 0; l < k->n_layers; ++l) { k->dz_off[l] = (long long)o; o += al((size_t)k->n_tiles * k->Np[l] * kRows * 2); }
    long long p = 0;
    for (int l = 0; l < k->n_layers; ++l) { k->part_layer_off[l] = p; p += (long long)k->N[l] * (k->Kp[l] + 16); }
    k->part_stride = p;
    k->part_off = (long long)o;
    o += al((size_t)sm_count() * p * sizeof(float));
  }
  k->total_bytes = o;
  return NRX_OK;
}

// ---- weight packing: fp32 [N,K] -> bf16 canonical images of W (B operand of fwd) and W^T (B operand of dX) ----
__global__ void __launch_bounds__(256)
tower_pack_kernel(const __grid_constant__ TowerK T, uint8_t* __restrict__ wpack, uint8_t* __restrict__ wtpack) {
  int grand = 0;
  for (int l = 0; l < T.n_mma; ++l) grand += (T.Kp[l] / 8) * T.Np[l];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < grand; e += gridDim.x * blockDim.x) {
    int l = 0, i = e;  // flat chunk index over all layers -> (layer, chunk)
    while (i >= (T.Kp[l] / 8) * T.Np[l]) { i -= (T.Kp[l] / 8) * T.Np[l]; ++l; }
    const int Kp = T.Kp[l], Np = T.Np[l], K = T.K[l], N = T.N[l];
    const float* W = T.w[l];
    {
      {  // W image: chunk (kc, n) = W[n][kc*8 .. +7]
        const int n = i % Np, kc = i / Np;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k0 = kc * 8 + 2 * j;
          const float a = (n < N && k0 < K) ? W[(long long)n * K + k0] : 0.f;
          const float b = (n < N && k0 + 1 < K) ? W[(long long)n * K + k0 + 1] : 0.f;
          p[j] = pack_bf16(a, b);
        }
        *reinterpret_cast<uint4*>(wpack + T.w_off[l] + (size_t)i * 16) = make_uint4(p[0], p[1], p[2], p[3]);
      }
      {  // W^T image: chunk (nc, k) = W[nc*8 .. +7][k]
        const int k = i % Kp, nc = i / Kp;
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n0 = nc * 8 + 2 * j;
          const float a = (k < K && n0 < N) ? W[(long long)n0 * K + k] : 0.f;
          const float b = (k < K && n0 + 1 < N) ? W[(long long)(n0 + 1) * K + k] : 0.f;
          p[j] = pack_bf16(a, b);
        }
        *reinterpret_cast<uint4*>(wtpack + T.wt_off[l] + (size_t)i * 16) = make_uint4(p[0], p[1], p[2], p[3]);
      }
    }
  }
}


// Issue the K loop of one layer: D[128 x n_cols] = A[128 x Kdim] * B[n_cols x Kdim]^T, both K-major canonical.
__device__ __forceinline__ void issue_layer_mma(uint32_t tmem, uint32_t a_base, uint32_t b_base, int kdim, int n_cols) {
  const uint32_t idesc = make_idesc_bf16(kRows, n_cols);
  for (int k16 = 0; k16 < kdim / 16; ++k16) {
    const uint64_t ad = make_smem_desc(a_base + (uint32_t)k16 * 2u * (kRows * 16u), kRows * 16u, 128u);
    const uint64_t bd = make_smem_desc(b_base + (uint32_t)k16 * 2u * ((uint32_t)n_cols * 16u), (uint32_t)n_cols * 16u, 128u);
    mma_bf16_ss(tmem, ad, bd, idesc, k16 > 0);
  }
}

// Same, called by a WHOLE warp in uniform control flow; `leader` (umma::elect_one_sync) issues.  Keeps the descriptors in
// uniform registers and avoids the per-thread waterfall the compiler builds around UTCHMMA under `if (tid == 0)`.
__device__ __forceinline__ void issue_layer_mma_u(uint32_t leader, uint32_t tmem, uint32_t a_base, uint32_t b_base, int kdim, int n_cols) {
  const uint32_t idesc = make_idesc_bf16(kRows, n_cols);
  uint64_t ad = make_smem_desc(a_base, kRows * 16u, 128u);
  uint64_t bd = make_smem_desc(b_base, (uint32_t)n_cols * 16u, 128u);
  const uint32_t astep = (2u * (kRows * 16u)) >> 4, bstep = (2u * ((uint32_t)n_cols * 16u)) >> 4;
  uint32_t accum = 0u;
#pragma unroll 2
  for (int k16 = 0; k16 < kdim / 16; ++k16) {
    mma_bf16_ss_if(leader, tmem, ad, bd, idesc, accum);
    ad += astep;
    bd += bstep;
    accum = 1u;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFwdThreads, 1)
tower_fwd_kernel(const __grid_constant__ TowerK T, const float* __restrict__ x, long long ldx, long long B,
                 float* __restrict__ y, long long ldy, const uint8_t* __restrict__ wpack, uint8_t* __restrict__ ws,
                 int training) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((T.w_bytes + 1023) & ~1023u);
  float* xch = reinterpret_cast<float*>(sA + (size_t)kRows * T.max_kp * 2);  // [kColSplit][128][kMaxTiny]
  float* sB = xch + kColSplit * kRows * kMaxTiny;                             // [n_mma][kBiasStride] biases (zero padded)
  __shared__ uint64_t wbar, mbar;
  __shared__ uint32_t tmem_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  const int r = q * 32 + lane;

  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, (uint32_t)T.tmem_cols);
  for (int l = 0; l < T.n_mma; ++l)
    for (int i = tid; i < kBiasStride; i += kFwdThreads) sB[l * kBiasStride + i] = i < T.N[l] ? __ldg(T.bias[l] + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid == 0) {  // weights -> smem once per CTA (TMA engine), overlapped with the first input tile
    mbar_expect_tx(&wbar, T.w_bytes);
    for (int l = 0; l < T.n_mma; ++l)
      bulk_g2s(sW + T.w_off[l], wpack + T.w_off[l], (uint32_t)T.Kp[l] * T.Np[l] * 2u, &wbar);
  }
  bool w_ready = false;
  uint32_t phase = 0;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);

  for (long long tile = blockIdx.x; tile < T.n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kRows;
    // ---- input tile: fp32 -> bf16 canonical (and its image for the backward) ----
    {
      const int Kp0 = T.Kp[0], K0 = T.K[0];
      uint8_t* img = training ? ws + T.act_off[0] + (size_t)tile * Kp0 * kRows * 2 : nullptr;
      constexpr int kBatch = 4;  // chunks per thread whose loads are all issued before the first use
      const int nchunk = (Kp0 / 8) * kRows;
      for (int i0 = tid; i0 < nchunk; i0 += kFwdThreads * kBatch) {
        float4 la[kBatch], lb[kBatch];
        const bool fast = vec_ok && (K0 % 8 == 0);
        if (fast) {
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kFwdThreads;
            const int rr = i % kRows, kc = i / kRows;
            const long long row = row0 + rr;
            la[u] = lb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nchunk && row < B && kc * 8 < K0) {
              const float4* src = reinterpret_cast<const float4*>(x + row * ldx + kc * 8);
              la[u] = __ldg(src);
              lb[u] = __ldg(src + 1);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = i0 + u * kFwdThreads;
          if (i >= nchunk) break;
          const int rr = i % kRows, kc = i / kRows;
          const long long row = row0 + rr;
          float f[8];
          if (fast) {
            f[0] = la[u].x; f[1] = la[u].y; f[2] = la[u].z; f[3] = la[u].w;
            f[4] = lb[u].x; f[5] = lb[u].y; f[6] = lb[u].z; f[7] = lb[u].w;
          } else if (row < B) {
            const float* src = x + row * ldx + kc * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (kc * 8 + j < K0) ? __ldg(src + j) : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
          const uint4 pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, rr, kc)) = pk;
          if (img) *reinterpret_cast<uint4*>(img + canon_off(kRows, rr, kc)) = pk;
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (!w_ready) { mbar_wait(&wbar, 0); w_ready = true; }

    float dot[kMaxTiny];
#pragma unroll
    for (int o = 0; o < kMaxTiny; ++o) dot[o] = 0.f;

    for (int l = 0; l < T.n_mma; ++l) {
      const int Np = T.Np[l], N = T.N[l];
      if (tid == 0) {
        tc_fence_after();
        issue_layer_mma(tmem, smem_u32(sA), smem_u32(sW + T.w_off[l]), T.Kp[l], Np);
        mma_commit(&mbar);
      }
      mbar_wait(&mbar, phase);
      phase ^= 1;
      tc_fence_after();
      const bool is_final = (l == T.n_layers - 1);
      const bool feeds_tiny = T.tiny && (l == T.n_mma - 1);
      const long long row = row0 + r;
      uint8_t* img = (training && l + 1 < T.n_layers) ? ws + T.act_off[l + 1] + (size_t)tile * Np * kRows * 2 : nullptr;
      for (int g = h; g < Np / 16; g += kColSplit) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 16u, v);
        tmem_ld_wait();
        // bias as four 16-byte SMEM loads; the bf16 rounding happens once, in pack_bf16 below — only the layer that
        // feeds the register dot product of the tiny last layer needs the rounded VALUES as floats
        const float4* bp4 = reinterpret_cast<const float4*>(sB + l * kBiasStride + g * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = bp4[j4];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float z = v[j4 * 4 + jj] + bb[jj];
            if (!is_final) {
              z = act_fwd(z, T.slope);
              if (feeds_tiny) z = bf16_round(z);
            }
            v[j4 * 4 + jj] = z;
          }
        }
        if (is_final) {
          if (row < B) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (g * 16 + j < N) y[row * ldy + g * 16 + j] = v[j];
          }
        } else {
          const uint4 c0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          const uint4 c1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g)) = c0;
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g + 1)) = c1;
          if (img) {
            *reinterpret_cast<uint4*>(img + canon_off(kRows, r, 2 * g)) = c0;
            *reinterpret_cast<uint4*>(img + canon_off(kRows, r, 2 * g + 1)) = c1;
          }
          if (feeds_tiny) {
            const int Kt = T.K[T.n_layers - 1], Nt = T.N[T.n_layers - 1];
            const float* wt = T.w[T.n_layers - 1];
#pragma unroll
            for (int o = 0; o < kMaxTiny; ++o) {
              if (o < Nt) {
                float s = dot[o];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (g * 16 + j < Kt) s = fmaf(v[j], __ldg(wt + o * Kt + g * 16 + j), s);
                dot[o] = s;
              }
            }
          }
        }
      }
      if (feeds_tiny) {
#pragma unroll
        for (int o = 0; o < kMaxTiny; ++o) xch[(h * kRows + r) * kMaxTiny + o] = dot[o];
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
    if (T.tiny && h == 0) {
      const long long row = row0 + r;
      const int Nt = T.N[T.n_layers - 1];
      if (row < B) {
        for (int o = 0; o < Nt; ++o)
        {
          float t = 0.f;
#pragma unroll
          for (int hh = 0; hh < kColSplit; ++hh) t += xch[(hh * kRows + r) * kMaxTiny + o];
          y[row * ldy + o] = t + __ldg(T.bias[T.n_layers - 1] + o);
        }
      }
    }
    if (T.tiny) __syncthreads();  // xch is reused by the next tile
  }
  if (!w_ready) mbar_wait(&wbar, 0);  // CTA without tiles: do not exit with the bulk copy in flight
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)T.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// Forward, two tile slots per CTA (large batches).  Two groups of 8 warps each run the same per-tile pipeline on
// their own tile, A-operand buffer and TMEM accumulator, sharing one SMEM weight image: while one group is in its
// epilogue (TMEM -> bias/act -> bf16 -> SMEM/HBM, CUDA cores) the other group's MMAs run on the tensor pipe, and
// one group's input-tile loads hide behind the other's compute.  Groups synchronise on named barriers only.
// ------------------------------------------------------------------------------------------------------------
static constexpr int kGrpThreads = 256;
static constexpr int kGrpSplit = kGrpThreads / 128;  // column slices per group

__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(kGrpThreads) : "memory"); }

__global__ void __launch_bounds__(2 * kGrpThreads, 1)
tower_fwd2_kernel(const __grid_constant__ TowerK T, const float* __restrict__ x, long long ldx, long long B,
                  float* __restrict__ y, long long ldy, const uint8_t* __restrict__ wpack, uint8_t* __restrict__ ws,
                  int training) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int grp = threadIdx.x / kGrpThreads, tid = threadIdx.x % kGrpThreads;
  const int warp = tid >> 5, lane = tid & 31;   // warp inside the group; CTA warp id = grp * 8 + warp, same value mod 4
  const int q = warp & 3, h = warp >> 2;
  const int r = q * 32 + lane;
  const size_t a_bytes = (size_t)kRows * T.max_kp * 2;
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((T.w_bytes + 1023) & ~1023u) + grp * a_bytes;
  float* xch = reinterpret_cast<float*>(smem + ((T.w_bytes + 1023) & ~1023u) + 2 * a_bytes) + grp * (kGrpSplit * kRows * kMaxTiny);
  float* sB = reinterpret_cast<float*>(smem + ((T.w_bytes + 1023) & ~1023u) + 2 * a_bytes) + 2 * (kGrpSplit * kRows * kMaxTiny);
  __shared__ uint64_t wbar, mbar2[2];
  __shared__ uint32_t tmem_s;
  uint64_t* mbar = &mbar2[grp];

  if (threadIdx.x == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar2[0], 1);
    mbar_init(&mbar2[1], 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) tmem_alloc(&tmem_s, (uint32_t)(2 * T.tmem_cols));
  for (int l = 0; l < T.n_mma; ++l)
    for (int i = threadIdx.x; i < kBiasStride; i += 2 * kGrpThreads) sB[l * kBiasStride + i] = i < T.N[l] ? __ldg(T.bias[l] + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s + (uint32_t)(grp * T.tmem_cols);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&wbar, T.w_bytes);
    for (int l = 0; l < T.n_mma; ++l)
      bulk_g2s(sW + T.w_off[l], wpack + T.w_off[l], (uint32_t)T.Kp[l] * T.Np[l] * 2u, &wbar);
  }
  bool w_ready = false;
  uint32_t phase = 0;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0);

  for (long long tile = (long long)blockIdx.x * 2 + grp; tile < T.n_tiles; tile += 2LL * gridDim.x) {
    const long long row0 = tile * kRows;
    {
      const int Kp0 = T.Kp[0], K0 = T.K[0];
      uint8_t* img = training ? ws + T.act_off[0] + (size_t)tile * Kp0 * kRows * 2 : nullptr;
      constexpr int kBatch = 4;
      const int nchunk = (Kp0 / 8) * kRows;
      const bool fast = vec_ok && (K0 % 8 == 0);
      for (int i0 = tid; i0 < nchunk; i0 += kGrpThreads * kBatch) {
        float4 la[kBatch], lb[kBatch];
        if (fast) {
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u * kGrpThreads;
            const int rr = i % kRows, kc = i / kRows;
            const long long row = row0 + rr;
            la[u] = lb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < nchunk && row < B && kc * 8 < K0) {
              const float4* src = reinterpret_cast<const float4*>(x + row * ldx + kc * 8);
              la[u] = __ldg(src);
              lb[u] = __ldg(src + 1);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          const int i = i0 + u * kGrpThreads;
          if (i >= nchunk) break;
          const int rr = i % kRows, kc = i / kRows;
          const long long row = row0 + rr;
          float f[8];
          if (fast) {
            f[0] = la[u].x; f[1] = la[u].y; f[2] = la[u].z; f[3] = la[u].w;
            f[4] = lb[u].x; f[5] = lb[u].y; f[6] = lb[u].z; f[7] = lb[u].w;
          } else if (row < B) {
            const float* src = x + row * ldx + kc * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = (kc * 8 + j < K0) ? __ldg(src + j) : 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = 0.f;
          }
          const uint4 pk = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, rr, kc)) = pk;
          if (img) *reinterpret_cast<uint4*>(img + canon_off(kRows, rr, kc)) = pk;
        }
      }
    }
    fence_async_smem();
    group_sync(grp);
    if (!w_ready) { mbar_wait(&wbar, 0); w_ready = true; }

    float dot[kMaxTiny];
#pragma unroll
    for (int o = 0; o < kMaxTiny; ++o) dot[o] = 0.f;

    for (int l = 0; l < T.n_mma; ++l) {
      const int Np = T.Np[l], N = T.N[l];
      if (tid == 0) {
        tc_fence_after();
        issue_layer_mma(tmem, smem_u32(sA), smem_u32(sW + T.w_off[l]), T.Kp[l], Np);
        mma_commit(mbar);
      }
      mbar_wait(mbar, phase);
      phase ^= 1;
      tc_fence_after();
      const bool is_final = (l == T.n_layers - 1);
      const bool feeds_tiny = T.tiny && (l == T.n_mma - 1);
      const long long row = row0 + r;
      uint8_t* img = (training && l + 1 < T.n_layers) ? ws + T.act_off[l + 1] + (size_t)tile * Np * kRows * 2 : nullptr;
      for (int g = h; g < Np / 16; g += kGrpSplit) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g * 16u, v);
        tmem_ld_wait();
        // bias as four 16-byte SMEM loads; the bf16 rounding happens once, in pack_bf16 below — only the layer that
        // feeds the register dot product of the tiny last layer needs the rounded VALUES as floats
        const float4* bp4 = reinterpret_cast<const float4*>(sB + l * kBiasStride + g * 16);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 b4 = bp4[j4];
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            float z = v[j4 * 4 + jj] + bb[jj];
            if (!is_final) {
              z = act_fwd(z, T.slope);
              if (feeds_tiny) z = bf16_round(z);
            }
            v[j4 * 4 + jj] = z;
          }
        }
        if (is_final) {
          if (row < B) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (g * 16 + j < N) y[row * ldy + g * 16 + j] = v[j];
          }
        } else {
          const uint4 c0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          const uint4 c1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g)) = c0;
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g + 1)) = c1;
          if (img) {
            *reinterpret_cast<uint4*>(img + canon_off(kRows, r, 2 * g)) = c0;
            *reinterpret_cast<uint4*>(img + canon_off(kRows, r, 2 * g + 1)) = c1;
          }
          if (feeds_tiny) {
            const int Kt = T.K[T.n_layers - 1], Nt = T.N[T.n_layers - 1];
            const float* wt = T.w[T.n_layers - 1];
#pragma unroll
            for (int o = 0; o < kMaxTiny; ++o) {
              if (o < Nt) {
                float sacc = dot[o];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (g * 16 + j < Kt) sacc = fmaf(v[j], __ldg(wt + o * Kt + g * 16 + j), sacc);
                dot[o] = sacc;
              }
            }
          }
        }
      }
      if (feeds_tiny) {
#pragma unroll
        for (int o = 0; o < kMaxTiny; ++o) xch[(h * kRows + r) * kMaxTiny + o] = dot[o];
      }
      fence_async_smem();
      tc_fence_before();
      group_sync(grp);
    }
    if (T.tiny && h == 0) {
      const long long row = row0 + r;
      const int Nt = T.N[T.n_layers - 1];
      if (row < B) {
        for (int o = 0; o < Nt; ++o) {
          float t = 0.f;
#pragma unroll
          for (int hh = 0; hh < kGrpSplit; ++hh) t += xch[(hh * kRows + r) * kMaxTiny + o];
          y[row * ldy + o] = t + __ldg(T.bias[T.n_layers - 1] + o);
        }
      }
    }
    if (T.tiny) group_sync(grp);  // xch is reused by the group's next tile
  }
  if (!w_ready) mbar_wait(&wbar, 0);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem_s, (uint32_t)(2 * T.tmem_cols));
}

static size_t fwd2_smem_bytes(const TowerK& k) {
  return ((k.w_bytes + 1023) & ~1023u) + 2 * (size_t)kRows * k.max_kp * 2 +
         (2 * kGrpSplit * kRows * kMaxTiny + NRX_MAX_LAYERS * kBiasStride) * sizeof(float);
}

// ------------------------------------------------------------------------------------------------------------
// Backward (1): dX chain.  Writes the bf16 image of dL/dz_l for every layer (M-side operand of the dW GEMMs)
// and grad_x.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& c, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 t = __bfloat1622float2(p[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}

__global__ void __launch_bounds__(kFwdThreads, 1)
tower_bwd_dx_kernel(const __grid_constant__ TowerK T, long long B, const float* __restrict__ gy, long long ldgy,
                    float* __restrict__ gx, long long ldgx, int accumulate_gx, const uint8_t* __restrict__ wtpack,
                    uint8_t* __restrict__ ws) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sA = smem + ((T.wt_bytes + 1023) & ~1023u);
  __shared__ uint64_t wbar, mbar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, h = warp >> 2;
  const int r = q * 32 + lane;

  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, (uint32_t)T.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (tid == 0) {
    mbar_expect_tx(&wbar, T.wt_bytes);
    for (int l = 0; l < T.n_mma; ++l)
      bulk_g2s(sW + T.wt_off[l], wtpack + T.wt_off[l], (uint32_t)T.Kp[l] * T.Np[l] * 2u, &wbar);
  }
  bool w_ready = false;
  uint32_t phase = 0;
  const int L = T.n_layers;

  for (long long tile = blockIdx.x; tile < T.n_tiles; tile += gridDim.x) {
    const long long row0 = tile * kRows;
    const long long row = row0 + r;
    // ---- seed: dz of the last MMA layer into sA (+ images) ----
    {
      const int lm = T.n_mma - 1;           // last MMA layer
      const int Np = T.Np[lm], N = T.N[lm];
      uint8_t* dzimg = ws + T.dz_off[lm] + (size_t)tile * Np * kRows * 2;
      float g[kMaxTiny];
#pragma unroll
      for (int o = 0; o < kMaxTiny; ++o) g[o] = 0.f;
      if (T.tiny) {
        const int Nt = T.N[L - 1];
        if (row < B)
          for (int o = 0; o < Nt; ++o) g[o] = __ldg(gy + row * ldgy + o);
        if (h == 0) {  // image of dz_tiny (= grad_y, padded to 16 columns)
          uint8_t* timg = ws + T.dz_off[L - 1] + (size_t)tile * 16 * kRows * 2;
          *reinterpret_cast<uint4*>(timg + canon_off(kRows, r, 0)) = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), 0u, 0u);
          *reinterpret_cast<uint4*>(timg + canon_off(kRows, r, 1)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      const uint8_t* aimg = T.tiny ? ws + T.act_off[L - 1] + (size_t)tile * Np * kRows * 2 : nullptr;
      for (int g16 = h; g16 < Np / 16; g16 += kColSplit) {
        float v[16];
        if (T.tiny) {  // da = g (x) w_tiny, dz = da * act'(a)
          const int Kt = T.K[L - 1], Nt = T.N[L - 1];
          const float* wt = T.w[L - 1];
          float a[16];
          unpack8(*reinterpret_cast<const uint4*>(aimg + canon_off(kRows, r, 2 * g16)), *reinterpret_cast<float(*)[8]>(&a[0]));
          unpack8(*reinterpret_cast<const uint4*>(aimg + canon_off(kRows, r, 2 * g16 + 1)), *reinterpret_cast<float(*)[8]>(&a[8]));
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = g16 * 16 + j;
            float d = 0.f;
            if (col < Kt)
              for (int o = 0; o < Nt; ++o) d = fmaf(g[o], __ldg(wt + o * Kt + col), d);
            v[j] = d * (a[j] > 0.f ? 1.f : T.slope);
          }
        } else {  // last layer is an MMA layer without activation: dz = grad_y
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = g16 * 16 + j;
            v[j] = (row < B && col < N) ? __ldg(gy + row * ldgy + col) : 0.f;
          }
        }
        const uint4 c0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        const uint4 c1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
        *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g16)) = c0;
        *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g16 + 1)) = c1;
        *reinterpret_cast<uint4*>(dzimg + canon_off(kRows, r, 2 * g16)) = c0;
        *reinterpret_cast<uint4*>(dzimg + canon_off(kRows, r, 2 * g16 + 1)) = c1;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (!w_ready) { mbar_wait(&wbar, 0); w_ready = true; }

    for (int l = T.n_mma - 1; l >= 0; --l) {
      if (l == 0 && gx == nullptr) break;  // nobody needs dL/dx
      const int Kp = T.Kp[l], K = T.K[l];
      if (warp == 0) {   // whole warp, uniform control flow; one elected lane issues
        const uint32_t leader = elect_one_sync() ? 1u : 0u;
        tc_fence_after();
        // D[128 x Kp] = dz_l[128 x Np] * (W^T image: Kp rows, contraction Np)
        issue_layer_mma_u(leader, tmem, smem_u32(sA), smem_u32(sW + T.wt_off[l]), T.Np[l], Kp);
        mma_commit_if(leader, &mbar);
        __syncwarp();
      }
      const uint8_t* aimg = l > 0 ? ws + T.act_off[l] + (size_t)tile * Kp * kRows * 2 : nullptr;
      uint8_t* dzimg = l > 0 ? ws + T.dz_off[l - 1] + (size_t)tile * Kp * kRows * 2 : nullptr;
      // the saved activations (for act') do not depend on the MMA: fetch them while it runs
      constexpr int kMaxG = 4;  // 16-column groups per thread: Kp <= 240 -> 15 groups over kColSplit = 4 slices
      uint4 pa[kMaxG][2];
      if (l > 0) {
#pragma unroll
        for (int u = 0; u < kMaxG; ++u) {
          const int g16 = h + u * kColSplit;
          if (g16 < Kp / 16) {
            pa[u][0] = *reinterpret_cast<const uint4*>(aimg + canon_off(kRows, r, 2 * g16));
            pa[u][1] = *reinterpret_cast<const uint4*>(aimg + canon_off(kRows, r, 2 * g16 + 1));
          }
        }
      }
      mbar_wait(&mbar, phase);
      phase ^= 1;
      tc_fence_after();
#pragma unroll
      for (int u = 0; u < kMaxG; ++u) {
        const int g16 = h + u * kColSplit;
        if (g16 >= Kp / 16) break;
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)g16 * 16u, v);
        tmem_ld_wait();
        if (l > 0) {
          float a[16];
          unpack8(pa[u][0], *reinterpret_cast<float(*)[8]>(&a[0]));
          unpack8(pa[u][1], *reinterpret_cast<float(*)[8]>(&a[8]));
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] *= (a[j] > 0.f ? 1.f : T.slope);
          const uint4 c0 = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          const uint4 c1 = make_uint4(pack_bf16(v[8], v[9]), pack_bf16(v[10], v[11]), pack_bf16(v[12], v[13]), pack_bf16(v[14], v[15]));
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g16)) = c0;
          *reinterpret_cast<uint4*>(sA + canon_off(kRows, r, 2 * g16 + 1)) = c1;
          *reinterpret_cast<uint4*>(dzimg + canon_off(kRows, r, 2 * g16)) = c0;
          *reinterpret_cast<uint4*>(dzimg + canon_off(kRows, r, 2 * g16 + 1)) = c1;
        } else if (row < B) {
          float* p = gx + row * ldgx + g16 * 16;
          if (!accumulate_gx && g16 * 16 + 16 <= K && (ldgx % 4 == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(p + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (g16 * 16 + j < K) p[j] = accumulate_gx ? p[j] + v[j] : v[j];
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }
  }
  if (!w_ready) mbar_wait(&wbar, 0);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)T.tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// Backward (2): dW_l[N_l x K_l] (+ db_l as column Kp_l) = sum over tiles dz_l^T a_l, split over CTAs.
// Operands are the tile images read as MN-major (M = dz columns, N = a columns, K = the 128 batch rows).
// ------------------------------------------------------------------------------------------------------------
static constexpr int kDwThreads = 128;
static constexpr int kDwStages = 2;
static constexpr int kStageM = kRows * 128 * 2;          // dz image zero-extended to 128 columns: 32 KB
static constexpr int kStageN = kRows * (240 + 16) * 2;   // a image + ones chunk pair: 64 KB (wider first layers: stage_n argument)

__global__ void __launch_bounds__(kDwThreads, 1)
tower_bwd_dw_kernel(const __grid_constant__ TowerK T, const uint8_t* __restrict__ ws, float* __restrict__ partials, int stage_n,
                    int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sM[kDwStages];
  uint8_t* sN[kDwStages];
  for (int s = 0; s < kDwStages; ++s) {
    sM[s] = smem + (size_t)s * (kStageM + stage_n);
    sN[s] = sM[s] + kStageM;
  }
  __shared__ uint64_t full[kDwStages], empty[kDwStages], done;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(NRX_FULL_MASK, tid >> 5, 0);   // provably warp-uniform: the role branches below are uniform, so ptxas keeps the MMA descriptors in uniform registers

  if (tid == 0) {
    for (int s = 0; s < kDwStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;

  // contiguous tile range of this CTA
  const long long per = (T.n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = (long long)blockIdx.x * per;
  const long long t1 = t0 + per < T.n_tiles ? t0 + per : T.n_tiles;
  uint32_t full_phase[kDwStages] = {0, 0}, empty_phase[kDwStages] = {0, 0}, done_phase = 0;
  bool stage_used[kDwStages] = {false, false};
  float* mypart = partials + (long long)blockIdx.x * T.part_stride;

  // layers are a grid dimension: CTA (x, y) reduces tile range x of layer y (5x more CTAs in flight than
  // walking the layers inside one CTA)
  for (int l = blockIdx.y; l < T.n_layers; l += gridDim.y) {
    const int Np = T.Np[l], Kp = T.Kp[l], N = T.N[l];
    const int ncols = Kp + 16;
    // ---- per-layer stage setup: zero the unused M columns, place the ones chunk pair after the a image ----
    for (int s = 0; s < kDwStages; ++s) {
      for (int i = tid; i < (128 - Np) / 8 * kRows; i += kDwThreads)
        *reinterpret_cast<uint4*>(sM[s] + (size_t)Np * kRows * 2 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
      for (int i = tid; i < 2 * kRows; i += kDwThreads) {
        const uint32_t one = (i < kRows) ? 0x00003F80u : 0u;  // bf16 1.0 in element 0 of the first chunk only
        *reinterpret_cast<uint4*>(sN[s] + (size_t)Kp * kRows * 2 + (size_t)i * 16) = make_uint4(one, 0u, 0u, 0u);
      }
    }
    fence_async_smem();
    __syncthreads();

    if (t1 > t0) {
      if (warp == 0) {  // producer: TMA bulk copies of the two tile images (uniform role branch, one lane works)
        if (lane == 0) {
          int s = 0;
          for (long long t = t0; t < t1; ++t) {
            if (stage_used[s]) { mbar_wait(&empty[s], empty_phase[s]); empty_phase[s] ^= 1; }
            stage_used[s] = true;
            const uint32_t bm = (uint32_t)Np * kRows * 2, bn = (uint32_t)Kp * kRows * 2;
            mbar_expect_tx(&full[s], bm + bn);
            bulk_g2s(sM[s], ws + T.dz_off[l] + (size_t)t * bm, bm, &full[s]);
            bulk_g2s(sN[s], ws + T.act_off[l] + (size_t)t * bn, bn, &full[s]);
            s ^= 1;
          }
        }
        __syncwarp();
      } else if (warp == 1) {  // MMA issuer: whole warp in uniform control flow, one elected lane issues
        const uint32_t leader = elect_one_sync() ? 1u : 0u;
        // one MMA covers <= 256 output columns: a wider first layer (K + 16 > 256) is issued as two column blocks that
        // alternate per K step (independent accumulators)
        const int n1 = ncols <= 256 ? ncols : (((ncols / 2) + 15) & ~15);
        const int n2 = ncols - n1;
        const uint32_t idesc = make_idesc_bf16(128, n1, 1, 1);
        const uint32_t idesc2 = make_idesc_bf16(128, n2 > 0 ? n2 : 16, 1, 1);
        uint32_t s = 0, fph = full_phase[0] | (full_phase[1] << 1);
        const uint32_t stage_bytes = (uint32_t)(kStageM + stage_n), sm0 = smem_u32(smem);
        for (long long t = t0; t < t1; ++t) {
          mbar_wait(&full[s], (fph >> s) & 1u);
          fph ^= 1u << s;
          tc_fence_after();
          // MN-major: 8-element MN groups are kRows*16 B apart (SBO), 8-row K groups 128 B apart (LBO)
          uint64_t ad = make_smem_desc(sm0 + s * stage_bytes, 128u, kRows * 16u);
          uint64_t bd = make_smem_desc(sm0 + s * stage_bytes + (uint32_t)kStageM, 128u, kRows * 16u);
          const uint64_t b2off = (uint64_t)(((uint32_t)n1 * 256u) >> 4);   // n1 columns further on in the MN-major operand
          uint32_t accum = t > t0 ? 1u : 0u;
#pragma unroll
          for (int k16 = 0; k16 < kRows / 16; ++k16) {
            mma_bf16_ss_if(leader, tmem, ad, bd, idesc, accum);
            if (n2 > 0) mma_bf16_ss_if(leader, tmem + (uint32_t)n1, ad, bd + b2off, idesc2, accum);
            ad += 256u >> 4;
            bd += 256u >> 4;
            accum = 1u;
          }
          mma_commit_if(leader, &empty[s]);
          s ^= 1u;
        }
        full_phase[0] = fph & 1u; full_phase[1] = (fph >> 1) & 1u;
        mma_commit_if(leader, &done);
        __syncwarp();
      }
      mbar_wait(&done, done_phase);
      done_phase ^= 1;
      tc_fence_after();
      // drain: thread n owns accumulator row n (= output feature n of layer l)
      const int n = warp * 32 + lane;
      float* dst = mypart + T.part_layer_off[l] + (long long)n * ncols;
      for (int g = 0; g < ncols / 16; ++g) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)g * 16u, v);
        tmem_ld_wait();
        if (n < N) {
#pragma unroll
          for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + g * 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// Backward (3): fixed-order sum of the per-CTA partials -> dW [N,K] (nn.Linear layout) and db [N].
__device__ __forceinline__ void
tower_bwd_reduce_body(const TowerK& T, const float* __restrict__ partials, int n_parts, long long tiles_per_part,
                      float* const* gw, float* const* gb) {
  // one flat index space over all layers: every thread owns one output element, so the (latency-bound)
  // partial loads of all layers are in flight together instead of layer after layer
  long long grand = 0;
  for (int l = 0; l < T.n_layers; ++l) grand += (long long)T.N[l] * (T.K[l] + 1);
  // CTAs past the last tile wrote nothing; the sum order over the live ones is fixed (c ascending)
  const int live = (int)min((long long)n_parts, (T.n_tiles + tiles_per_part - 1) / tiles_per_part);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < grand; e += (long long)gridDim.x * blockDim.x) {
    int l = 0;
    long long i = e;
    while (i >= (long long)T.N[l] * (T.K[l] + 1)) { i -= (long long)T.N[l] * (T.K[l] + 1); ++l; }
    const int ncols = T.Kp[l] + 16, K = T.K[l];
    const int n = (int)(i / (K + 1)), k = (int)(i % (K + 1));
    const int col = (k < K) ? k : T.Kp[l];
    const float* p = partials + T.part_layer_off[l] + (long long)n * ncols + col;
    float s = 0.f;
    int c = 0;
    for (; c + 16 <= live; c += 16) {  // 16 independent loads in flight, added in order
      float v[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) v[u] = __ldg(p + (long long)(c + u) * T.part_stride);
#pragma unroll
      for (int u = 0; u < 16; ++u) s += v[u];
    }
    for (; c < live; ++c) s += __ldg(p + (long long)c * T.part_stride);
    if (k < K) { if (gw[l]) gw[l][(long long)n * K + k] = s; }
    else if (gb[l]) gb[l][n] = s;
  }
}

struct PtrPack { float* gw[NRX_MAX_LAYERS]; float* gb[NRX_MAX_LAYERS]; };
__global__ void __launch_bounds__(256)
tower_bwd_reduce_entry(const __grid_constant__ TowerK T, const float* __restrict__ partials, int n_parts, long long tiles_per_part,
                       const __grid_constant__ PtrPack P) {
  tower_bwd_reduce_body(T, partials, n_parts, tiles_per_part, P.gw, P.gb);
}

static size_t fwd_smem_bytes(const TowerK& k, bool bwd) {
  const unsigned wb = bwd ? k.wt_bytes : k.w_bytes;
  return ((wb + 1023) & ~1023u) + (size_t)kRows * k.max_kp * 2 +
         (bwd ? 0 : (kColSplit * kRows * kMaxTiny + NRX_MAX_LAYERS * kBiasStride) * sizeof(float));
}

}  // namespace nrx

using namespace nrx;

extern "C" size_t nrx_tower_workspace_bytes(const NrxTower* h_tower, int64_t B, int training) {
  TowerK k;
  if (make_tower(h_tower, B, training, &k) != NRX_OK) return 0;
  return k.total_bytes;
}

static int tower_pack(const TowerK& k, uint8_t* ws, cudaStream_t st) {
  tower_pack_kernel<<<sm_count(), 256, 0, st>>>(k, ws + k.wpack_off, ws + k.wtpack_off);
  return check_launch("tower_pack");
}

extern "C" int nrx_tower_pack(const NrxTower* h_tower, int64_t B, int training, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  TowerK k;
  int rc = make_tower(h_tower, B, training & 1, &k);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws && ws_bytes >= k.total_bytes, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, k.total_bytes);
  return tower_pack(k, (uint8_t*)ws, (cudaStream_t)stream);
}

static int tower_fwd_any(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B, float* y, int64_t y_ld, int flags,
                         const NrxTowerHead* head, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  const int prepacked = flags & NRX_TOWER_PREPACKED;
  const int ximg = flags & NRX_TOWER_XIMG;
  const int training = flags & 1;
  TowerK k;
  int rc = make_tower(h_tower, B, training, &k);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws && ws_bytes >= k.total_bytes, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, k.total_bytes);
  NRX_REQUIRE(((x || ximg) && (y || head)) || B == 0, NRX_EINVAL, "null x / y");
  NRX_REQUIRE((ximg || x_ld >= k.K[0]) && (!y || y_ld >= k.N[k.n_layers - 1]), NRX_EINVAL, "leading dimension too small");
  NRX_REQUIRE(y || k.tiny, NRX_EINVAL, "y may only be omitted when the tower ends in the fused head");
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!prepacked) {
    rc = tower_pack(k, (uint8_t*)ws, st);
    if (rc != NRX_OK) return rc;
  }
  const char* env3 = getenv("NRX_TOWER_V3");
  const bool want3 = env3 ? atoi(env3) != 0 : true;
  if ((want3 || ximg || head || k.wide0) && tower_fwd3_eligible(k))
    return tower_fwd3_launch(k, ximg ? nullptr : x, x_ld, B, y, y_ld, (uint8_t*)ws, training, head, st);
  NRX_REQUIRE(!ximg && !head && !k.wide0, NRX_EUNSUPPORTED, "tower shape outside the pipelined forward (image input / fused head / wide first layer need it)");
  const size_t smem = fwd_smem_bytes(k, false);
  NRX_REQUIRE(smem <= 227 * 1024, NRX_EUNSUPPORTED, "tower needs %zu B of shared memory (> 227 KB)", smem);
  {
    // large batches: two tile slots per CTA (needs >= 3 tiles per SM to pay, 2 A buffers + 2 accumulators to fit)
    const size_t smem2 = fwd2_smem_bytes(k);
    const char* env = getenv("NRX_TOWER_V2");
    const bool want = env ? atoi(env) != 0 : true;
    if (want && k.n_tiles >= 3LL * sm_count() && smem2 <= 227 * 1024 && 2 * k.tmem_cols <= 512) {
      cudaError_t e2 = cudaFuncSetAttribute(tower_fwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      NRX_REQUIRE(e2 == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e2));
      tower_fwd2_kernel<<<(unsigned)sm_count(), 2 * kGrpThreads, smem2, st>>>(k, x, x_ld, B, y, y_ld, (const uint8_t*)ws + k.wpack_off,
                                                                            (uint8_t*)ws, training);
      return check_launch("tower_fwd2");
    }
  }
  cudaError_t e = cudaFuncSetAttribute(tower_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
  const long long grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  tower_fwd_kernel<<<(unsigned)grid, kFwdThreads, smem, st>>>(k, x, x_ld, B, y, y_ld, (const uint8_t*)ws + k.wpack_off, (uint8_t*)ws,
                                                             training);
  return check_launch("tower_fwd");
}

extern "C" int nrx_tower_fwd(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B, float* y, int64_t y_ld,
                             int training, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  return tower_fwd_any(h_tower, x, x_ld, B, y, y_ld, training, nullptr, ws, ws_bytes, stream);
}

extern "C" int nrx_tower_fwd_head(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B, const NrxTowerHead* h_head,
                                  int flags, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  NRX_REQUIRE(h_head != nullptr, NRX_EINVAL, "null head");
  return tower_fwd_any(h_tower, x, x_ld, B, nullptr, 1, flags, h_head, ws, ws_bytes, stream);
}

extern "C" int nrx_tower_bwd_dx(const NrxTower* h_tower, int64_t B, const float* grad_y, int64_t gy_ld, float* grad_x,
                                int64_t gx_ld, int accumulate_gx, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  TowerK k;
  int rc = make_tower(h_tower, B, 1, &k);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws && ws_bytes >= k.total_bytes, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, k.total_bytes);
  NRX_REQUIRE(grad_y || B == 0, NRX_EINVAL, "null grad_y");
  NRX_REQUIRE(!k.wide0 || tower_dx3_eligible(k, grad_x != nullptr), NRX_EUNSUPPORTED,
              "tower backward: a first layer wider than 240 columns (got %d) needs the pipelined dX kernel", k.K[0]);
  NRX_REQUIRE(gy_ld >= k.N[k.n_layers - 1] && (!grad_x || gx_ld >= k.K[0]), NRX_EINVAL, "leading dimension too small");
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)ws;
  {
    const char* env3 = getenv("NRX_TOWER_DX3");
    if ((env3 ? atoi(env3) != 0 : true) && tower_dx3_eligible(k, grad_x != nullptr))
      return tower_dx3_launch(k, B, grad_y, gy_ld, grad_x, gx_ld, accumulate_gx, w, st);
    const size_t smem = fwd_smem_bytes(k, true);
    NRX_REQUIRE(smem <= 227 * 1024, NRX_EUNSUPPORTED, "tower backward needs %zu B of shared memory (> 227 KB)", smem);
    cudaError_t e = cudaFuncSetAttribute(tower_bwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
    const long long grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
    tower_bwd_dx_kernel<<<(unsigned)grid, kFwdThreads, smem, st>>>(k, B, grad_y, gy_ld, grad_x, gx_ld, accumulate_gx,
                                                                  w + k.wtpack_off, w);
    rc = check_launch("tower_bwd_dx");
  }
  return rc;
}

extern "C" int nrx_tower_bwd_dw(const NrxTower* h_tower, int64_t B, float* const* h_grad_w, float* const* h_grad_b, void* ws,
                                size_t ws_bytes, nrx_stream_t stream) {
  TowerK k;
  int rc = make_tower(h_tower, B, 1, &k);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws && ws_bytes >= k.total_bytes, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, k.total_bytes);
  NRX_REQUIRE(h_grad_w && h_grad_b, NRX_EINVAL, "null gradient pointer arrays");
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)ws;
  {
    const int stage_n = k.max_kp <= 240 ? kStageN : kRows * (k.max_kp + 16) * 2;
    const size_t smem = (size_t)kDwStages * (kStageM + stage_n);
    NRX_REQUIRE(smem <= 227 * 1024 && k.max_kp + 16 <= 496, NRX_EUNSUPPORTED, "tower dW: first layer of %d columns is too wide", k.K[0]);
    const int dw_tmem = k.max_kp + 16 <= 256 ? 256 : 512;
    cudaError_t e = cudaFuncSetAttribute(tower_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
    // >= 4 tiles per CTA so the partial-sum traffic stays small next to the GEMM work
    // one wave of (tile range x layer) CTAs: parts * n_layers ~= SM count keeps every SM busy with the longest
    // possible accumulation chains and the fewest partial buffers
    long long parts = sm_count() / k.n_layers;
    if (parts > k.n_tiles) parts = k.n_tiles;
    if (parts < 1) parts = 1;
    const long long per = (k.n_tiles + parts - 1) / parts;
    float* partials = (float*)(w + k.part_off);
    tower_bwd_dw_kernel<<<dim3((unsigned)parts, (unsigned)k.n_layers), kDwThreads, smem, st>>>(k, w, partials, stage_n, dw_tmem);
    rc = check_launch("tower_bwd_dw");
    if (rc != NRX_OK) return rc;
    PtrPack P;
    for (int l = 0; l < NRX_MAX_LAYERS; ++l) {
      P.gw[l] = l < k.n_layers ? h_grad_w[l] : nullptr;
      P.gb[l] = l < k.n_layers ? h_grad_b[l] : nullptr;
    }
    tower_bwd_reduce_entry<<<sm_count() * 2, 256, 0, st>>>(k, partials, (int)parts, per, P);
    rc = check_launch("tower_bwd_reduce");
  }
  return rc;
}

extern "C" int nrx_tower_bwd(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B, const float* grad_y,
                             int64_t gy_ld, float* grad_x, int64_t gx_ld, int accumulate_gx, float* const* h_grad_w,
                             float* const* h_grad_b, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  (void)x; (void)x_ld;  // the forward saved the bf16 image of x in ws
  int rc = nrx_tower_bwd_dx(h_tower, B, grad_y, gy_ld, grad_x, gx_ld, accumulate_gx, ws, ws_bytes, stream);
  if (rc != NRX_OK) return rc;
  return nrx_tower_bwd_dw(h_tower, B, h_grad_w, h_grad_b, ws, ws_bytes, stream);
}

// Byte offsets (inside the training workspace) and widths of the saved tile images, for tests and tooling:
// image of layer l's input a_l (width Kp_l) and of dL/dz_l (width Np_l); each is [tile][width/8][128][8] bf16.
extern "C" int nrx_tower_image_layout(const NrxTower* h_tower, int64_t B, int64_t* act_off, int32_t* act_width,
                                      int64_t* dz_off, int32_t* dz_width) {
  TowerK k;
  int rc = make_tower(h_tower, B, 1, &k);
  if (rc != NRX_OK) return rc;
  for (int l = 0; l < k.n_layers; ++l) {
    if (act_off) act_off[l] = k.act_off[l];
    if (act_width) act_width[l] = k.Kp[l];
    if (dz_off) dz_off[l] = k.dz_off[l];
    if (dz_width) dz_width[l] = k.Np[l];
  }
  return NRX_OK;
}
