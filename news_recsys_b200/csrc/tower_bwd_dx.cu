// K4 backward (1), warp-specialised and pipelined (round 2): the dX chain  dz_{l-1} = (dz_l W_l) * act'(a_l)  of the MLP /
// DSSM towers (reference: autograd of utils.py:6-17, recall/DSSM/model.py:26-44), same machine mapping as the forward
// in tower_fwd.cu:
//   warp 0      : loads the W^T operand images of every layer once (bulk async copies), one mbarrier per layer;
//   warp 1      : issues every tcgen05.mma in uniform control flow (umma::elect_one_sync).  The A operand dz_l lives in
//                 TENSOR MEMORY (written by the previous epilogue with tcgen05.st), the B operand is the layer's W^T image
//                 in shared memory.  Layers whose input is wider than 128 columns (the first layer of DCN: 224, with
//                 user_history 288) are issued as column blocks of <= 128, one accumulator pass each;
//   warps 2..17 : two groups of 8 epilogue warps, one per TILE SLOT (accumulator 128 TMEM columns + dz operand 64).
//                 Per pass: the saved activation chunks a_l (act' gate) are fetched from the HBM image while the MMA runs,
//                 then tcgen05.ld -> * gate -> bf16 pack -> tcgen05.st (next A operand) + dz image store (operand of the dW
//                 GEMMs); the last layer writes grad_x in fp32.  The chain starts with a "seed" epilogue that builds
//                 dz of the last MMA layer from grad_y (and the <= 4-wide register-dot layer's weights).
// Two slots alternate layer by layer, so one slot's MMAs run under the other's epilogue.
#include <math.h>

#include "tower.cuh"

namespace nrx {
using namespace umma;

static constexpr int kDx3Threads = 64 + 512;
static constexpr int kDxSlotCols = 256;      // TMEM columns per slot: accumulator [0,128) + dz operand [128,192)
static constexpr int kDxMaxItems = 16;       // MMA passes per tile (layers x column blocks)

struct Dx3Geom {
  int n_items;
  int layer[kDxMaxItems];   // MMA layer of the pass (descending)
  int col0[kDxMaxItems];    // first output column of the pass (multiple of 16)
  int ncol[kDxMaxItems];    // columns of the pass (multiple of 16, <= 128)
  unsigned smem_bytes, off_wt;
};

static bool make_dx3_geom(const TowerK& k, bool need_gx, Dx3Geom* g) {
  memset(g, 0, sizeof(*g));
  for (int l = 0; l < k.n_mma; ++l)
    if (k.Np[l] > 128) return false;
  if (k.tiny && k.K[k.n_layers - 1] > 128) return false;
  int n = 0;
  for (int l = k.n_mma - 1; l >= 0; --l) {
    if (l == 0 && !need_gx) break;
    const int Kp = k.Kp[l];
    const int nb = (Kp + 127) / 128;
    const int per = (((Kp + nb - 1) / nb) + 15) & ~15;
    for (int c = 0; c < Kp; c += per) {
      if (n >= kDxMaxItems) return false;
      g->layer[n] = l;
      g->col0[n] = c;
      g->ncol[n] = (Kp - c) < per ? (Kp - c) : per;
      ++n;
    }
  }
  g->n_items = n;
  const unsigned nt = k.tiny ? (unsigned)k.N[k.n_layers - 1] : 0u;
  g->off_wt = (k.wt_bytes + 1023) & ~1023u;
  g->smem_bytes = g->off_wt + nt * 128 * 4;
  return g->smem_bytes <= 232448 - 1024;
}

bool tower_dx3_eligible(const TowerK& k, bool need_gx) {
  Dx3Geom g;
  return make_dx3_geom(k, need_gx, &g);
}

__device__ __forceinline__ void unpack8f(const uint4& c, float (&f)[8]) {
  const uint32_t w[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = __uint_as_float(w[j] << 16);
    f[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// dz = v * act'(a) for NG 16-column groups, packed to bf16x2 words
template <int NG>
__device__ __forceinline__ void gate_pack(const float (&v)[32], const uint4 (&a)[4], float slope, uint32_t (&pk)[16]) {
#pragma unroll
  for (int q4 = 0; q4 < 2 * NG; ++q4) {
    float f[8];
    unpack8f(a[q4], f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g0 = f[2 * j] > 0.f ? 1.f : slope, g1 = f[2 * j + 1] > 0.f ? 1.f : slope;
      pk[q4 * 4 + j] = pack_bf16(v[q4 * 8 + 2 * j] * g0, v[q4 * 8 + 2 * j + 1] * g1);
    }
  }
}

template <int NG>
__device__ __forceinline__ void store_chunks(uint8_t* img, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int q4 = 0; q4 < 2 * NG; ++q4)
    *reinterpret_cast<uint4*>(img + (size_t)q4 * (kRows * 16)) = make_uint4(pk[q4 * 4], pk[q4 * 4 + 1], pk[q4 * 4 + 2], pk[q4 * 4 + 3]);
}

__global__ void __launch_bounds__(kDx3Threads, 1)
tower_bwd_dx3_kernel(const __grid_constant__ TowerK T, const __grid_constant__ Dx3Geom G, long long B, const float* __restrict__ gy,
                     long long ldgy, float* __restrict__ gx, long long ldgx, int accumulate_gx, const uint8_t* __restrict__ wtpack,
                     uint8_t* __restrict__ ws) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem;
  float* sWt = reinterpret_cast<float*>(smem + G.off_wt);   // [Nt][128] weights of the register-dot layer
  __shared__ uint64_t wbar[NRX_MAX_LAYERS], acc_full[2], epi_done[2], seed_done[2];
  __shared__ uint32_t tmem_s;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(NRX_FULL_MASK, tid >> 5, 0);   // provably warp-uniform: the role branches below are uniform, so ptxas keeps the MMA descriptors in uniform registers
  const int L = T.n_layers, nm = T.n_mma;
  const int n_my = (int)(T.n_tiles > blockIdx.x ? (T.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);

  if (tid == 0) {
    for (int l = 0; l < NRX_MAX_LAYERS; ++l) mbar_init(&wbar[l], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&epi_done[s], 8); mbar_init(&seed_done[s], 8); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, 512u);
  if (T.tiny) {
    const int Kt = T.K[L - 1], Nt = T.N[L - 1];
    for (int i = tid; i < Nt * 128; i += kDx3Threads) {
      const int o = i >> 7, c = i & 127;
      sWt[i] = c < Kt ? __ldg(T.w[L - 1] + (long long)o * Kt + c) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;

  if (warp == 0) {
    if (lane == 0) {   // W^T images, the layer the chain needs first goes first
      for (int l = nm - 1; l >= 0; --l) {
        const uint32_t wb = (uint32_t)T.Kp[l] * T.Np[l] * 2u;
        mbar_expect_tx(&wbar[l], wb);
        bulk_g2s(sW + T.wt_off[l], wtpack + T.wt_off[l], wb, &wbar[l]);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (uniform control flow, elected lane issues) =====================
    const uint32_t leader = elect_one_sync() ? 1u : 0u;
    uint32_t dph = 0, sph = 0;
    // Every completed phase of a barrier is consumed by exactly one wait, in order (an mbarrier may run at most one
    // phase ahead of its waiter): per tile and slot the epilogue arrives once on seed_done (the seed) and once per pass on
    // epi_done (the drain); the issuer waits seed_done before the first pass, epi_done before every later pass, and the
    // last drain of a tile before the first pass of the slot's NEXT tile.
    for (int j0 = 0; j0 < n_my; j0 += 2) {
      for (int it = 0; it < G.n_items; ++it) {
        const int l = G.layer[it];
        if (j0 == 0 && (it == 0 || G.layer[it - 1] != l)) mbar_wait(&wbar[l], 0);
        const int Kp = T.Kp[l], ncol = G.ncol[it];
        const uint32_t idesc = make_idesc_bf16(kRows, ncol);
        // B operand: rows [col0, col0 + ncol) of the W^T image (Kp rows, contraction Np), K-major canonical
        const uint64_t bd0 = make_smem_desc(smem_u32(sW + T.wt_off[l]) + (uint32_t)G.col0[it] * 16u, (uint32_t)Kp * 16u, 128u);
        const uint32_t bstep = (2u * (uint32_t)Kp * 16u) >> 4;
        const int n16 = T.Np[l] / 16;
        for (int s = 0; s < 2; ++s) {
          if (j0 + s >= n_my) continue;
          if (it == 0) {
            if (j0 > 0) { mbar_wait(&epi_done[s], (dph >> s) & 1u); dph ^= 1u << s; }   // last drain of the slot's previous tile
            mbar_wait(&seed_done[s], (sph >> s) & 1u);
            sph ^= 1u << s;
          } else {
            mbar_wait(&epi_done[s], (dph >> s) & 1u);
            dph ^= 1u << s;
          }
          tc_fence_after();
          const uint32_t acc = tmem + (uint32_t)s * kDxSlotCols;
          uint64_t bd = bd0;
          uint32_t a = acc + 128u, accum = 0u;
#pragma unroll 2
          for (int k16 = 0; k16 < n16; ++k16) {
            mma_bf16_ts_if(leader, acc, a, bd, idesc, accum);
            bd += bstep;
            a += 8u;
            accum = 1u;
          }
          mma_commit_if(leader, &acc_full[s]);
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int e = warp - 2;
    const int s = e >> 3;                 // tile slot of this warp group
    const int half = (e >> 2) & 1;        // column half
    const int qd = warp & 3;
    const int r = qd * 32 + lane;
    const uint32_t acc = tmem + (uint32_t)s * kDxSlotCols + ((uint32_t)(qd * 32) << 16);
    const uint32_t act = acc + 128u;
    uint64_t* const done_bar = &epi_done[s];
    uint64_t* const full_bar = &acc_full[s];
    uint32_t aph = 0;
    const int lm = nm - 1;                // last MMA layer
    const float slope = T.slope;
    for (int j = s; j < n_my; j += 2) {
      const int tile = (int)blockIdx.x + j * (int)gridDim.x;
      const long long row = (long long)tile * kRows + r;
      // ---- seed: dz of the last MMA layer -> TMEM operand + image --------------------------------------------------
      {
        const int Np = T.Np[lm], N = T.N[lm];
        const int csplit = ((Np >> 1) + 15) & ~15;
        const int c0 = half ? csplit : 0;
        const int n16 = ((half ? Np : csplit) - c0) >> 4;
        uint8_t* dzimg = ws + T.dz_off[lm] + (size_t)tile * Np * kRows * 2 + canon_off(kRows, r, c0 >> 3);
        float g[kMaxTiny];
#pragma unroll
        for (int o = 0; o < kMaxTiny; ++o) g[o] = 0.f;
        if (T.tiny) {
          const int Nt = T.N[L - 1];
          if (row < B)
            for (int o = 0; o < Nt; ++o) g[o] = __ldg(gy + row * ldgy + o);
          if (half == 0 && qd >= 0) {  // image of dz_tiny (= grad_y, padded to 16 columns)
            uint8_t* timg = ws + T.dz_off[L - 1] + (size_t)tile * 16 * kRows * 2;
            *reinterpret_cast<uint4*>(timg + canon_off(kRows, r, 0)) = make_uint4(pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]), 0u, 0u);
            *reinterpret_cast<uint4*>(timg + canon_off(kRows, r, 1)) = make_uint4(0u, 0u, 0u, 0u);
          }
        }
        const uint8_t* aimg = T.tiny ? ws + T.act_off[L - 1] + (size_t)tile * Np * kRows * 2 + canon_off(kRows, r, c0 >> 3) : nullptr;
        for (int g16 = 0; g16 < n16; ++g16) {
          float v[16];
          if (T.tiny) {   // da = g (x) w_tiny, dz = da * act'(a)
            const int Nt = T.N[L - 1];
            float a[16];
            unpack8f(*reinterpret_cast<const uint4*>(aimg + (size_t)(2 * g16) * (kRows * 16)), *reinterpret_cast<float(*)[8]>(&a[0]));
            unpack8f(*reinterpret_cast<const uint4*>(aimg + (size_t)(2 * g16 + 1) * (kRows * 16)), *reinterpret_cast<float(*)[8]>(&a[8]));
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int col = c0 + g16 * 16 + jj;
              float d = 0.f;
#pragma unroll
              for (int o = 0; o < kMaxTiny; ++o)
                if (o < Nt) d = fmaf(g[o], sWt[o * 128 + col], d);
              v[jj] = d * (a[jj] > 0.f ? 1.f : slope);
            }
          } else {        // the last layer is an MMA layer without activation: dz = grad_y
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
              const int col = c0 + g16 * 16 + jj;
              v[jj] = (row < B && col < N) ? __ldg(gy + row * ldgy + col) : 0.f;
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) pk[jj] = pack_bf16(v[2 * jj], v[2 * jj + 1]);
          tmem_st8(act + (uint32_t)((c0 + g16 * 16) >> 1), pk);
          *reinterpret_cast<uint4*>(dzimg + (size_t)(2 * g16) * (kRows * 16)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(dzimg + (size_t)(2 * g16 + 1) * (kRows * 16)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&seed_done[s]);
      }
      // ---- the chain ---------------------------------------------------------------------------------------------------
      for (int it = 0; it < G.n_items; ++it) {
        const int l = G.layer[it];
        const int ncol = G.ncol[it], cb = G.col0[it];
        const int csplit = ((ncol >> 1) + 15) & ~15;
        const int c0 = half ? csplit : 0;                      // inside the pass
        const int n16 = ((half ? ncol : csplit) - c0) >> 4;    // 16-column groups of this thread: 0..4
        const int Kp = T.Kp[l];
        // the act' gate of this thread's columns: chunks of the saved a_l image, in flight while the MMA runs
        uint4 ga[2][4];
        if (l > 0) {
          const uint8_t* aimg = ws + T.act_off[l] + (size_t)tile * Kp * kRows * 2 + canon_off(kRows, r, (cb + c0) >> 3);
#pragma unroll
          for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4)
              if (p * 4 + q4 < 2 * n16) ga[p][q4] = *reinterpret_cast<const uint4*>(aimg + (size_t)(p * 4 + q4) * (kRows * 16));
        }
        mbar_wait(full_bar, aph);
        aph ^= 1u;
        tc_fence_after();
        if (l > 0) {
          uint8_t* dzimg = ws + T.dz_off[l - 1] + (size_t)tile * Kp * kRows * 2 + canon_off(kRows, r, (cb + c0) >> 3);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            const int np = n16 - 2 * pass;
            if (np <= 0) continue;
            float v[32];
            uint32_t pk[16];
            const int cp = pass * 32;
            if (np >= 2) {
              tmem_ld32(acc + (uint32_t)(c0 + cp), v);
              tmem_ld_wait();
              gate_pack<2>(v, ga[pass], slope, pk);
              tmem_st16(act + (uint32_t)((cb + c0 + cp) >> 1), pk);
              store_chunks<2>(dzimg + (size_t)(cp >> 3) * (kRows * 16), pk);
            } else {
              tmem_ld16(acc + (uint32_t)(c0 + cp), *reinterpret_cast<float(*)[16]>(&v[0]));
              tmem_ld_wait();
              gate_pack<1>(v, ga[pass], slope, pk);
              tmem_st8(act + (uint32_t)((cb + c0 + cp) >> 1), *reinterpret_cast<uint32_t(*)[8]>(&pk[0]));
              store_chunks<1>(dzimg + (size_t)(cp >> 3) * (kRows * 16), pk);
            }
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          // the operand of the next layer is complete only after the LAST column block of this layer (l >= 1 has one block)
          if (lane == 0) mbar_arrive(done_bar);
        } else {
          // layer 0: grad_x in fp32, 16 columns at a time
          const int K = T.K[0];
          const bool vec_ok = !accumulate_gx && (ldgx % 4 == 0) && ((reinterpret_cast<uintptr_t>(gx) & 15) == 0);
          for (int g16 = 0; g16 < n16; ++g16) {
            float v[16];
            tmem_ld16(acc + (uint32_t)(c0 + g16 * 16), v);
            tmem_ld_wait();
            if (row < B) {
              const int col = cb + c0 + g16 * 16;
              float* p = gx + row * ldgx + col;
              if (vec_ok && col + 16 <= K) {
#pragma unroll
                for (int jj = 0; jj < 16; jj += 4) *reinterpret_cast<float4*>(p + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
              } else {
#pragma unroll
                for (int jj = 0; jj < 16; ++jj)
                  if (col + jj < K) p[jj] = accumulate_gx ? p[jj] + v[jj] : v[jj];
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(done_bar);   // accumulator drained (the dz_0 operand is still needed by the next block)
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512u);
}

int tower_dx3_launch(const TowerK& k, long long B, const float* gy, long long ldgy, float* gx, long long ldgx, int accumulate_gx,
                     uint8_t* ws, cudaStream_t st) {
  Dx3Geom g;
  NRX_REQUIRE(make_dx3_geom(k, gx != nullptr, &g), NRX_EUNSUPPORTED, "tower shape outside the pipelined dX kernel");
  cudaError_t e = cudaFuncSetAttribute(tower_bwd_dx3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
  const long long grid = k.n_tiles < sm_count() ? k.n_tiles : sm_count();
  tower_bwd_dx3_kernel<<<(unsigned)grid, kDx3Threads, g.smem_bytes, st>>>(k, g, B, gy, ldgy, gx, ldgx, accumulate_gx, ws + k.wtpack_off, ws);
  return check_launch("tower_bwd_dx3");
}

}  // namespace nrx
