// K4 forward, warp-specialised and pipelined (round 2).  Same contract as the one-tile kernel in tower.cu
// (MLP utils.py:6-17 of the reference behind deep/model.py:12-21, widedeep/model.py:14-27, dcn/model.py:15-29 and the
// DSSM towers recall/DSSM/model.py:26-44), different machine mapping:
//
//   warp 0      producer : bulk async copies (TMA engine, SASS UBLKCP) of the NEXT tiles' bf16 input images into a
//                          ring of shared-memory stages, one stage = 128 rows x <= 128 input columns (the first layer
//                          is K-streamed, so any input width works and a stage is free again after its K chunk);
//   warp 1      issuer   : one thread issues every tcgen05.mma.  Layer 0 reads its A operand from the stage ring
//                          (SS form); layers >= 1 read their A operand from TENSOR MEMORY (TS form): activations never
//                          touch shared memory, so weights (SMEM) are the only operand stream on the SMEM port;
//   warps 2..17 epilogue : two groups of 8 warps, one per TILE SLOT.  A slot owns a TMEM accumulator (128 columns)
//                          and a TMEM activation region (64 columns = 128 bf16).  Per layer: tcgen05.ld accumulator ->
//                          + bias -> activation -> bf16 pack (FADD2 + F2FP.RELU) -> tcgen05.st into the slot's
//                          activation region -> mbarrier arrive; the packed registers also go to HBM as the saved
//                          activation image when training.
//   Two slots in flight: the issuer alternates slot 0 / slot 1 layer by layer, so the tensor pipe runs slot 1's layer
//   while slot 0's epilogue runs on the CUDA cores, and vice versa.  The 64 -> 1 logit layer stays a register dot
//   product, now followed by the fused head (FM / wide term + bias + sigmoid + BCE + dL/dlogit, NrxTowerHead).
//
// Input is the bf16 tile image [tile][Kp0/8][128][8] in the workspace: written by tower_ximg_kernel below from fp32
// rows, or directly by the producing kernel (K1 / K5) — in training it doubles as the saved a_0 image.
#include <math.h>

#include "tower.cuh"

namespace nrx {
using namespace umma;

static constexpr int kF3Threads = 64 + 512;  // producer warp, issuer warp, 2 x 8 epilogue warps
static constexpr int kF3MaxStages = 4;
static constexpr int kF3MaxChunks = 8;
static constexpr int kSlotCols = 256;  // TMEM columns per slot: accumulator [0,128) + activations [128,192)
static constexpr int kSmemLimit = 232448 - 1024;  // 227 KB minus static barriers / slack

struct Fwd3Geom {
  int n_chunks;
  int chunk_col[kF3MaxChunks + 1];  // first input column of chunk c (multiples of 16); chunk_col[n_chunks] = Kp0
  int stage_bytes, n_stages;
  unsigned off_stage, off_bias, off_wt, off_xch, smem_bytes;
};

static bool make_fwd3_geom(const TowerK& k, Fwd3Geom* g) {
  memset(g, 0, sizeof(*g));
  if (k.n_mma < 1) return false;
  for (int l = 0; l < k.n_mma; ++l) {
    if (k.Np[l] > 128) return false;
    if (l > 0 && k.Kp[l] > 128) return false;
  }
  if (k.tiny && (k.K[k.n_layers - 1] > 128)) return false;
  const int Kp0 = k.Kp[0];
  g->n_chunks = (Kp0 + 127) / 128;
  if (g->n_chunks > kF3MaxChunks) return false;
  int per = (((Kp0 + g->n_chunks - 1) / g->n_chunks) + 15) & ~15;
  int c = 0, mx = 0;
  for (int i = 0; i < g->n_chunks; ++i) {
    g->chunk_col[i] = c;
    const int w = (Kp0 - c) < per ? (Kp0 - c) : per;
    if (w > mx) mx = w;
    c += w;
  }
  g->chunk_col[g->n_chunks] = Kp0;
  g->stage_bytes = mx * kRows * 2;
  unsigned o = (k.w_bytes + 1023) & ~1023u;
  g->off_stage = o;
  const unsigned nt = k.tiny ? (unsigned)k.N[k.n_layers - 1] : 0u;
  const unsigned fixed = (unsigned)k.n_mma * 128 * 4 + nt * 128 * 4 + 2u * kRows * nt * 4;
  int ns = ((int)kSmemLimit - (int)o - (int)fixed) / g->stage_bytes;
  if (ns > kF3MaxStages) ns = kF3MaxStages;
  if (ns < 2) return false;
  g->n_stages = ns;
  o += (unsigned)ns * g->stage_bytes;
  g->off_bias = o; o += (unsigned)k.n_mma * 128 * 4;
  g->off_wt = o;   o += nt * 128 * 4;
  g->off_xch = o;  o += 2u * kRows * nt * 4;
  g->smem_bytes = o;
  return true;
}

bool tower_fwd3_eligible(const TowerK& k) {
  Fwd3Geom g;
  return make_fwd3_geom(k, &g);
}

// ---- fp32 rows -> bf16 tile image --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tower_ximg_kernel(const float* __restrict__ x, long long ldx, long long B, int K, int Kp, uint8_t* __restrict__ img,
                  long long n_tiles) {
  const long long nchunk = n_tiles * (Kp / 8) * kRows;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (ldx % 4 == 0) && (K % 8 == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nchunk; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % kRows);
    const long long rest = i / kRows;
    const int kc = (int)(rest % (Kp / 8));
    const long long tile = rest / (Kp / 8);
    const long long row = tile * kRows + r;
    float f[8];
    if (row < B && vec_ok && kc * 8 < K) {
      const float4* src = reinterpret_cast<const float4*>(x + row * ldx + kc * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (row < B && kc * 8 + j < K) ? __ldg(x + row * ldx + kc * 8 + j) : 0.f;
    }
    *reinterpret_cast<uint4*>(img + (size_t)tile * Kp * kRows * 2 + canon_off(kRows, r, kc)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// ---- epilogue arithmetic -----------------------------------------------------------------------------------
// (z0 + b0, z1 + b1) -> activation -> one bf16x2 word (element 0 in the low half)
template <bool RELU>
__device__ __forceinline__ uint32_t bias_act_pack(float z0, float z1, float b0, float b1, float slope) {
  uint64_t rz, rb, rs;
  asm("mov.b64 %0, {%1, %2};" : "=l"(rz) : "f"(z0), "f"(z1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rs) : "l"(rz), "l"(rb));
  float s0, s1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(rs));
  uint32_t d;
  if (RELU) {
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(s1), "f"(s0));
  } else {
    s0 = act_fwd(s0, slope);
    s1 = act_fwd(s1, slope);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(s1), "f"(s0));
  }
  return d;
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// Epilogue arithmetic of one pass over NG 16-column groups held in v[]: fully unrolled, no per-group guards.
template <bool RELU, int NG>
__device__ __forceinline__ void hidden_pack(const float (&v)[32], uint32_t (&pk)[16], const float* __restrict__ bias, float slope) {
  const float4* bp4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int j4 = 0; j4 < 4 * NG; ++j4) {
    const float4 b4 = bp4[j4];
    pk[j4 * 2] = bias_act_pack<RELU>(v[j4 * 4], v[j4 * 4 + 1], b4.x, b4.y, slope);
    pk[j4 * 2 + 1] = bias_act_pack<RELU>(v[j4 * 4 + 2], v[j4 * 4 + 3], b4.z, b4.w, slope);
  }
}
// 2 * NG 16-byte chunks (8 columns each) of this thread's row into the saved tile image
template <int NG>
__device__ __forceinline__ void image_store(uint8_t* img, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q4 = 0; q4 < 2 * NG; ++q4)
    *reinterpret_cast<uint4*>(img + (size_t)q4 * (kRows * 16)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
}
// register dot product of the <= 4-wide last layer on the bf16-rounded activations (weights [Nt][128] in shared memory)
template <int NG>
__device__ __forceinline__ void tiny_dot(const uint32_t (&pk)[16], const float* __restrict__ wt, int Nt, float (&dot)[kMaxTiny]) {
#pragma unroll
  for (int o = 0; o < kMaxTiny; ++o) {
    if (o < Nt) {
      const float4* wp4 = reinterpret_cast<const float4*>(wt + o * 128);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < 4 * NG; ++j4) {
        const float4 w4 = wp4[j4];
        a0 = fmaf(bf16_lo(pk[j4 * 2]), w4.x, a0);
        a1 = fmaf(bf16_hi(pk[j4 * 2]), w4.y, a1);
        a0 = fmaf(bf16_lo(pk[j4 * 2 + 1]), w4.z, a0);
        a1 = fmaf(bf16_hi(pk[j4 * 2 + 1]), w4.w, a1);
      }
      dot[o] += a0 + a1;
    }
  }
}

struct HeadK {
  const float* terms[4];
  int n_terms;
  const float* bias;
  const float* label;
  long long lstride;
  float* logit;
  float* prob;
  float* loss;
  float* dlogit;
  int on;
};


template <bool RELU>
__global__ void __launch_bounds__(kF3Threads, 1)
tower_fwd3_kernel(const __grid_constant__ TowerK T, const __grid_constant__ Fwd3Geom G, const __grid_constant__ HeadK H,
                  long long B, float* __restrict__ y, long long ldy, const uint8_t* __restrict__ wpack,
                  uint8_t* __restrict__ ws, int training) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  uint8_t* sStage = smem + G.off_stage;
  float* sBias = reinterpret_cast<float*>(smem + G.off_bias);  // [n_mma][128]
  float* sWt = reinterpret_cast<float*>(smem + G.off_wt);      // [Nt][128] weights of the register-dot layer
  float* sXch = reinterpret_cast<float*>(smem + G.off_xch);    // [slot][128][Nt]
  __shared__ uint64_t wbar[NRX_MAX_LAYERS], full[kF3MaxStages], empty[kF3MaxStages], acc_full[2], epi_done[2];
  __shared__ uint32_t tmem_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(NRX_FULL_MASK, tid >> 5, 0);   // provably warp-uniform: the role branches below are uniform, so ptxas keeps the MMA descriptors in uniform registers
  const int L = T.n_layers, nm = T.n_mma;
  const long long n_my = T.n_tiles > blockIdx.x ? (T.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (tid == 0) {
    for (int l = 0; l < NRX_MAX_LAYERS; ++l) mbar_init(&wbar[l], 1);
    for (int s = 0; s < kF3MaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&epi_done[s], 8); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, 512u);
  for (int i = tid; i < nm * 128; i += kF3Threads) {
    const int l = i >> 7, c = i & 127;
    sBias[i] = c < T.N[l] ? __ldg(T.bias[l] + c) : 0.f;
  }
  if (T.tiny) {
    const int Kt = T.K[L - 1], Nt = T.N[L - 1];
    for (int i = tid; i < Nt * 128; i += kF3Threads) {
      const int o = i >> 7, c = i & 127;
      sWt[i] = c < Kt ? __ldg(T.w[L - 1] + (long long)o * Kt + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      for (int l = 0; l < nm; ++l) {   // one barrier per layer: layer 0 starts as soon as ITS weights have landed
        const uint32_t wb = (uint32_t)T.Kp[l] * T.Np[l] * 2u;
        mbar_expect_tx(&wbar[l], wb);
        bulk_g2s(sW + T.w_off[l], wpack + T.w_off[l], wb, &wbar[l]);
      }
      const uint8_t* ximg = ws + T.act_off[0];
      const size_t tile_bytes = (size_t)T.Kp[0] * kRows * 2;
      int st = 0;
      uint32_t eph = 0;   // bit s = parity to wait for on empty[s]
      uint32_t used = 0;
      for (long long j = 0; j < n_my; ++j) {
        const long long tile = blockIdx.x + j * gridDim.x;
        for (int c = 0; c < G.n_chunks; ++c) {
          if (used & (1u << st)) { mbar_wait(&empty[st], (eph >> st) & 1u); eph ^= 1u << st; }
          used |= 1u << st;
          const uint32_t bytes = (uint32_t)(G.chunk_col[c + 1] - G.chunk_col[c]) * kRows * 2u;
          mbar_expect_tx(&full[st], bytes);
          bulk_g2s(sStage + (size_t)st * G.stage_bytes, ximg + (size_t)tile * tile_bytes + (size_t)G.chunk_col[c] * kRows * 2, bytes,
                   &full[st]);
          st = (st + 1 == G.n_stages) ? 0 : st + 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the schedule in uniform control flow; one elected lane issues (see umma::elect_one_sync).
    const uint32_t leader = elect_one_sync() ? 1u : 0u;
    int st = 0;
    uint32_t fph = 0, dph = 0, slot_used = 0;
    for (long long j0 = 0; j0 < n_my; j0 += 2) {
      for (int l = 0; l < nm; ++l) {
        if (j0 == 0) mbar_wait(&wbar[l], 0);   // first use of this layer's weights
        const int Np = T.Np[l];
        const uint32_t idesc = make_idesc_bf16(kRows, Np);
        const uint32_t bstep = 2u * (uint32_t)Np * 16u;   // bytes between two K = 16 steps of the weight image
        const uint64_t bd0 = make_smem_desc(smem_u32(sW + T.w_off[l]), (uint32_t)Np * 16u, 128u);
        for (int s = 0; s < 2; ++s) {
          if (j0 + s >= n_my) continue;
          // the slot's previous epilogue has drained the accumulator and (l >= 1) published this layer's A operand
          if (slot_used & (1u << s)) { mbar_wait(&epi_done[s], (dph >> s) & 1u); dph ^= 1u << s; }
          slot_used |= 1u << s;
          const uint32_t acc = tmem + (uint32_t)s * kSlotCols;
          if (l == 0) {
            for (int c = 0; c < G.n_chunks; ++c) {
              mbar_wait(&full[st], (fph >> st) & 1u);
              fph ^= 1u << st;
              tc_fence_after();
              const uint64_t ad0 = make_smem_desc(smem_u32(sStage + (size_t)st * G.stage_bytes), kRows * 16u, 128u);
              const uint64_t bdc = desc_advance(bd0, (uint32_t)(G.chunk_col[c] / 8) * ((uint32_t)Np * 16u));
              const int n16 = (G.chunk_col[c + 1] - G.chunk_col[c]) / 16;
              uint64_t ad = ad0, bd = bdc;
              uint32_t accum = c > 0 ? 1u : 0u;
#pragma unroll 2
              for (int k16 = 0; k16 < n16; ++k16) {
                mma_bf16_ss_if(leader, acc, ad, bd, idesc, accum);
                ad += (uint64_t)((2u * (kRows * 16u)) >> 4);
                bd += (uint64_t)(bstep >> 4);
                accum = 1u;
              }
              mma_commit_if(leader, &empty[st]);
              st = (st + 1 == G.n_stages) ? 0 : st + 1;
            }
          } else {
            tc_fence_after();
            const uint32_t act = acc + 128u;
            const int n16 = T.Kp[l] / 16;
            uint64_t bd = bd0;
            uint32_t a = act, accum = 0u;
#pragma unroll 2
            for (int k16 = 0; k16 < n16; ++k16) {
              mma_bf16_ts_if(leader, acc, a, bd, idesc, accum);
              bd += (uint64_t)(bstep >> 4);
              a += 8u;
              accum = 1u;
            }
          }
          mma_commit_if(leader, &acc_full[s]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    // Two groups of 8 warps, one per tile slot; inside a group warp = (TMEM lane quadrant, column half), a thread owns
    // one row and up to 64 columns of the layer, drained in two passes of 32 (tcgen05.ld.x32 -> 16 FADD2 + 16 F2FP ->
    // tcgen05.st.x16).  Fewer, fatter warps keep the per-warp bookkeeping (barrier wait, addresses, flags) off the
    // issue slots: the epilogue of a layer has to fit inside the OTHER slot's MMA time for the tensor pipe to stay busy.
    const int e = warp - 2;
    const int s = e >> 3;                 // tile slot of this warp group
    const int half = (e >> 2) & 1;        // column half of the layer
    const int qd = warp & 3;              // TMEM lane quadrant this warp may touch
    const int r = qd * 32 + lane;
    const uint32_t acc = tmem + (uint32_t)s * kSlotCols + ((uint32_t)(qd * 32) << 16);
    const uint32_t act = acc + 128u;
    uint64_t* const done_bar = &epi_done[s];
    uint64_t* const full_bar = &acc_full[s];
    uint32_t aph = 0;
    const int Nt = T.tiny ? T.N[L - 1] : 0;
    const int n_my_i = (int)n_my;
    for (int j = s; j < n_my_i; j += 2) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const long long row = (long long)tile * kRows + r;
      for (int l = 0; l < nm; ++l) {
        const int Np = T.Np[l];
        const int csplit = ((Np >> 1) + 15) & ~15;
        const int c0 = half ? csplit : 0;
        const int n16 = ((half ? Np : csplit) - c0) >> 4;     // 16-column groups of this thread: 0..4
        const bool is_final = (l == L - 1);
        const bool feeds_tiny = T.tiny && (l == nm - 1);
        const bool to_tmem = (l + 1 < nm);
        const float* bias_l = sBias + l * 128 + c0;
        mbar_wait(full_bar, aph);
        aph ^= 1u;
        tc_fence_after();
        if (is_final) {
          // last Linear (no activation) wider than the register-dot limit: fp32 rows out, 16 columns at a time
          for (int g = 0; g < n16; ++g) {
            float v[16];
            tmem_ld16(acc + (uint32_t)(c0 + g * 16), v);
            tmem_ld_wait();
            if (row < B) {
              const float4* bp4 = reinterpret_cast<const float4*>(bias_l + g * 16);
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                const float4 b4 = bp4[j4];
                const int col = c0 + g * 16 + j4 * 4;
                const float o0 = v[j4 * 4] + b4.x, o1 = v[j4 * 4 + 1] + b4.y, o2 = v[j4 * 4 + 2] + b4.z, o3 = v[j4 * 4 + 3] + b4.w;
                float* dst = y + row * ldy + col;
                if (col + 4 <= T.N[l] && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
                  *reinterpret_cast<float4*>(dst) = make_float4(o0, o1, o2, o3);
                } else {
                  if (col < T.N[l]) dst[0] = o0;
                  if (col + 1 < T.N[l]) dst[1] = o1;
                  if (col + 2 < T.N[l]) dst[2] = o2;
                  if (col + 3 < T.N[l]) dst[3] = o3;
                }
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(done_bar);
          continue;
        }
        // hidden layer: two passes of <= 32 columns
        uint8_t* img = (training && l + 1 < L) ? ws + T.act_off[l + 1] + (size_t)tile * Np * kRows * 2 + canon_off(kRows, r, c0 >> 3)
                                               : nullptr;
        float dot[kMaxTiny];
#pragma unroll
        for (int o = 0; o < kMaxTiny; ++o) dot[o] = 0.f;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const int np = n16 - 2 * pass;       // groups of this pass: <= 0, 1 or >= 2
          if (np <= 0) continue;
          float v[32];
          uint32_t pk[16];
          const int cp = pass * 32;
          if (np >= 2) {
            tmem_ld32(acc + (uint32_t)(c0 + cp), v);
            tmem_ld_wait();
            hidden_pack<RELU, 2>(v, pk, bias_l + cp, T.slope);
            if (to_tmem) tmem_st16(act + (uint32_t)((c0 + cp) >> 1), pk);
            if (img != nullptr) image_store<2>(img + (size_t)(cp >> 3) * (kRows * 16), pk);
            if (feeds_tiny) tiny_dot<2>(pk, sWt + c0 + cp, Nt, dot);
          } else {
            tmem_ld16(acc + (uint32_t)(c0 + cp), *reinterpret_cast<float(*)[16]>(&v[0]));
            tmem_ld_wait();
            hidden_pack<RELU, 1>(v, pk, bias_l + cp, T.slope);
            if (to_tmem) tmem_st8(act + (uint32_t)((c0 + cp) >> 1), *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
            if (img != nullptr) image_store<1>(img + (size_t)(cp >> 3) * (kRows * 16), pk);
            if (feeds_tiny) tiny_dot<1>(pk, sWt + c0 + cp, Nt, dot);
          }
        }
        if (to_tmem) tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(done_bar);   // accumulator drained, next layer's A operand published
        if (feeds_tiny) {
          // the two column halves of a row meet in shared memory; half 0 finishes the row.  One buffer per slot is enough:
          // the slot's next write comes nm layer epilogues later, each of which every warp of the group (half 0 included,
          // after its reads below) must have passed for the issuer to move on.
          float* xc = sXch + (size_t)s * kRows * Nt;   // [row][Nt]
          if (half == 1) {
#pragma unroll
            for (int o = 0; o < kMaxTiny; ++o)
              if (o < Nt) xc[r * Nt + o] = dot[o];
          }
          asm volatile("bar.sync %0, 256;" ::"r"(s + 1) : "memory");
          if (half == 0 && row < B) {
#pragma unroll
            for (int o = 0; o < kMaxTiny; ++o) {
              if (o < Nt) {
                const float t = dot[o] + xc[r * Nt + o] + __ldg(T.bias[L - 1] + o);
                if (y) y[row * ldy + o] = t;
                if (H.on && o == 0) {
                  if (H.logit) H.logit[row] = t;
                  float z = 0.f;
                  for (int i = 0; i < H.n_terms; ++i) z += __ldg(H.terms[i] + row);
                  z += t;
                  if (H.bias) z += __ldg(H.bias);
                  const float p = sigmoid_f(z);
                  if (H.prob) H.prob[row] = p;
                  if (H.label)
                    bce_terms(p, __ldg(H.label + row * H.lstride), 1.f / (float)B, H.loss ? H.loss + row : nullptr,
                              H.dlogit ? H.dlogit + row : nullptr);
                }
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512u);
}

int tower_fwd3_launch(const TowerK& k, const float* x, long long ldx, long long B, float* y, long long ldy, uint8_t* ws,
                      int training, const NrxTowerHead* head, cudaStream_t st) {
  Fwd3Geom g;
  NRX_REQUIRE(make_fwd3_geom(k, &g), NRX_EUNSUPPORTED, "tower shape outside the pipelined forward");
  if (x != nullptr) {
    const long long nchunk = k.n_tiles * (k.Kp[0] / 8) * kRows;
    long long blocks = (nchunk + 255) / 256;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    tower_ximg_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, B, k.K[0], k.Kp[0], ws + k.act_off[0], k.n_tiles);
    int rc = check_launch("tower_ximg");
    if (rc != NRX_OK) return rc;
  }
  HeadK H;
  memset(&H, 0, sizeof(H));
  if (head != nullptr) {
    NRX_REQUIRE(k.tiny && k.N[k.n_layers - 1] == 1, NRX_EINVAL, "fused head needs a tower that ends in one logit");
    NRX_REQUIRE(head->n_terms >= 0 && head->n_terms <= 4, NRX_EINVAL, "head: n_terms outside [0,4]");
    H.on = 1;
    H.n_terms = head->n_terms;
    for (int i = 0; i < head->n_terms; ++i) {
      NRX_REQUIRE(head->terms[i] != nullptr, NRX_EINVAL, "head: null term %d", i);
      H.terms[i] = head->terms[i];
    }
    H.bias = head->bias; H.label = head->label; H.lstride = head->label_stride;
    H.logit = head->logit; H.prob = head->prob; H.loss = head->loss_per_sample; H.dlogit = head->dlogit;
  }
  const long long grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  if (k.act == NRX_ACT_RELU) {
    cudaError_t e = cudaFuncSetAttribute(tower_fwd3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
    tower_fwd3_kernel<true><<<(unsigned)grid, kF3Threads, g.smem_bytes, st>>>(k, g, H, B, y, ldy, ws + k.wpack_off, ws, training);
  } else {
    cudaError_t e = cudaFuncSetAttribute(tower_fwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
    tower_fwd3_kernel<false><<<(unsigned)grid, kF3Threads, g.smem_bytes, st>>>(k, g, H, B, y, ldy, ws + k.wpack_off, ws, training);
  }
  return check_launch("tower_fwd3");
}

}  // namespace nrx

using namespace nrx;

extern "C" int nrx_tower_image_from_rows(const float* x, int64_t x_ld, int64_t B, int width, void* image, nrx_stream_t stream) {
  NRX_REQUIRE(width >= 1 && x_ld >= width && B >= 0, NRX_EINVAL, "bad image arguments");
  if (B == 0) return NRX_OK;
  NRX_REQUIRE(x && image, NRX_EINVAL, "null pointer");
  const int Kp = (width + 15) & ~15;
  const long long n_tiles = (B + kRows - 1) / kRows;
  const long long nchunk = n_tiles * (Kp / 8) * kRows;
  long long blocks = (nchunk + 255) / 256;
  if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
  tower_ximg_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, x_ld, B, width, Kp, (uint8_t*)image, n_tiles);
  return check_launch("tower_ximg");
}
