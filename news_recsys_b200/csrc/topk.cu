// K6 — exact inner-product top-k over a full corpus (replaces faiss.IndexFlatIP.add/search as used at
// reference src/model/recall/DSSM/model.py:209,249-251 and src/model/model_utils/TopKSearcher.py:34-47,73-77).
//
// Result contract: for every query the k corpus rows with the largest inner product, ordered by
// (inner product desc, id asc), where the ordering key is the inner product of the fp32 inputs accumulated in
// fp64 (so near-ties that an fp32 BLAS would order arbitrarily are pinned); returned scores are that value
// rounded to fp32 (optionally also the fp64 value, which is what a sharded merge must order by); ids are corpus
// positions + id_base; -1 / -FLT_MAX pad when k > N.
//
// Fast path (tensor cores) with a proof of completeness per query:
//   index   : corpus packed once to bf16 128-row tile images in the canonical UMMA layout (+ max row norm);
//   sample  : S~ = Q C^T on tcgen05 over every `stride`-th corpus tile (1/8 of the corpus), epilogue keeps the max of
//             each 32-row group; theta = the ceil(k'/stride)+8-th largest group max (k' = 2k + 64): an ESTIMATE of the
//             k'-th best approximate score of the whole corpus — its quality only decides how many candidates pass;
//   scan    : the same GEMM over the WHOLE corpus, once; the epilogue appends rows with S~ >= theta to the query's
//             candidate list (a few hundred rows out of N);
//   final   : candidates re-scored exactly (fp64), sorted by (score desc, id asc).  Every row outside the list
//             has S~ < theta, hence true score < theta + eps (eps = bf16 rounding bound 2^-7.5 |q| max|c|).
//             If the k-th exact score is >= theta + eps the list provably contains the true top-k.  Otherwise
//             (too few candidates, or list overflow) the query goes on a device-side list for the
//   fallback: an exact fp64 scan of the corpus with a block-level streaming top-k, several blocks per listed query
//             (corpus slices) + a merge; with an empty list its fixed grid exits at once.
// One CTA scans a corpus slice for up to FOUR 128-query tiles: a corpus tile is loaded once per 512 queries (the L2
// stream drops 4x), and the issuer interleaves the K-steps of two query tiles so that consecutive tcgen05.mma never
// accumulate into the same TMEM accumulator (back-to-back dependent MMAs are spaced ~105 cycles apart, two
// independent chains run at ~72 cycles per 128x128x16 MMA: tools/umma_probe2.cu).
// Tensor-bound for Q >= ~256 (2 N D flops per query), HBM-bound (corpus stream) below.
#include <float.h>

#include "common.cuh"
#include "umma.cuh"

namespace nrx {
using namespace umma;

static constexpr int kTR = 128;        // corpus rows per tile == UMMA N; queries per tile == UMMA M
static constexpr int kCap = 2048;      // candidate list capacity per query (24 KB of shared memory in the final kernel)
static constexpr int kScanThreads = 64 + 512;  // TMA warp + MMA warp + 16 epilogue warps
static constexpr int kMaxQT = 4;       // query tiles per CTA (4 x 128 TMEM columns)
static constexpr int kHdrBytes = 256;
static constexpr int kScanSmem = 232448 - 2048;

struct TopkGeom {
  long long N, Q, n_tiles, n_qtiles, Qp;
  int D, Dp, k, kprime, stages;
  int nq;             // query tiles per CTA
  long long n_qgroups;
  long long slices;   // corpus slices per query group (grid.x of the scan)
  int stride;         // the sample pass visits every stride-th corpus tile
  long long n_stiles; // tiles of the sample pass
  int kprime_s;       // rank of theta among the sample's group maxima
  int cap_s;          // candidate capacity per (query, slice, column quarter): private region, no atomics
  int fb_grid, fb_items;   // fallback: fixed grid, capacity of the (query, corpus slice) work list
  size_t tile_bytes, index_bytes;
  // workspace offsets
  size_t gmax, theta, eps, count, cand, flag, flist, fpart, qimg, pre_id, pre_sc, total;
  int split_final;    // rescoring as its own wide kernel (few queries: the per-query final blocks alone leave the machine idle)
};

static int make_geom(long long Q, long long N, int D, int k, TopkGeom* g, int kprime_override = 0) {
  NRX_REQUIRE(N >= 0 && Q >= 0 && D >= 1 && k >= 1, NRX_EINVAL, "bad top-k sizes");
  NRX_REQUIRE(D <= 256, NRX_EUNSUPPORTED, "top-k supports D <= 256 (got %d)", D);
  NRX_REQUIRE(k <= 1024, NRX_EUNSUPPORTED, "top-k supports k <= 1024 (got %d)", k);
  memset(g, 0, sizeof(*g));
  g->N = N; g->Q = Q; g->D = D; g->k = k;
  g->Dp = (D + 15) & ~15;
  g->n_tiles = (N + kTR - 1) / kTR;
  g->n_qtiles = (Q + kTR - 1) / kTR;
  g->Qp = g->n_qtiles * kTR;
  g->kprime = kprime_override > 0 ? kprime_override : 2 * k + 64;
  g->tile_bytes = (size_t)kTR * g->Dp * 2;
  g->index_bytes = kHdrBytes + (size_t)g->n_tiles * g->tile_bytes;
  // query tiles per CTA: as many as fit beside >= 2 corpus stages
  int nq = (int)(kScanSmem / g->tile_bytes) - 2;
  if (nq > kMaxQT) nq = kMaxQT;
  if (nq < 1) nq = 1;
  if (nq > g->n_qtiles) nq = g->n_qtiles > 0 ? (int)g->n_qtiles : 1;
  g->nq = nq;
  g->stages = (int)(kScanSmem / g->tile_bytes) - nq;
  if (g->stages > 4) g->stages = 4;
  g->n_qgroups = (g->n_qtiles + nq - 1) / nq;
  g->slices = g->n_qgroups > 0 ? sm_count() / g->n_qgroups : 1;
  if (g->slices < 1) g->slices = 1;
  if (g->slices > g->n_tiles) g->slices = g->n_tiles > 0 ? g->n_tiles : 1;
  g->stride = (int)(g->n_tiles / 128);
  if (g->stride > 8) g->stride = 8;
  if (g->stride < 1) g->stride = 1;
  g->n_stiles = (g->n_tiles + g->stride - 1) / g->stride;
  g->kprime_s = g->stride == 1 ? g->kprime : (g->kprime + g->stride - 1) / g->stride + 8;
  g->cap_s = (int)(kCap / (4 * g->slices));
  if (g->cap_s < 12) g->cap_s = 12;
  if (g->cap_s > 512) g->cap_s = 512;
  g->fb_grid = 2 * sm_count();
  g->fb_items = (int)(Q > g->fb_grid ? Q : g->fb_grid);
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  g->gmax = o;  o += al((size_t)g->n_stiles * 4 * g->Qp * 4);
  g->theta = o; o += al((size_t)g->Qp * 4);
  g->eps = o;   o += al((size_t)g->Qp * 4);
  g->count = o; o += al((size_t)g->Qp * g->slices * 4 * 4);
  g->flag = o;  o += al((size_t)g->Qp * 4);
  g->flist = o; o += al((size_t)(g->Qp + 4) * 4);              // [0] = number of listed queries, then their ids
  g->fpart = o; o += al((size_t)g->fb_items * k * 12);         // fallback partial lists: fp64 score + u32 id
  g->qimg = o;  o += al((size_t)g->n_qtiles * g->tile_bytes);   // bf16 tile images of the queries (packed once per search)
  g->cand = o;  o += al((size_t)g->Qp * g->slices * 4 * g->cap_s * 4);
  // measured (B200, N = 1M, D = 128, k = 100): Q = 1: 97 vs 122 us, Q = 64: 106 vs 129 us, Q = 256: equal, Q = 1024: 397 vs
  // 344 us — with many queries the per-query blocks already fill the machine and the 16 extra blocks per query only
  // repeat the region prefix.  NRX_TOPK_FINAL_SPLIT=0/1 forces the choice.
  const char* fs = getenv("NRX_TOPK_FINAL_SPLIT");
  g->split_final = fs != nullptr ? (fs[0] == '1') : (Q <= 128);
  if (g->split_final) {
    g->pre_id = o; o += al((size_t)g->Qp * kCap * 4);
    g->pre_sc = o; o += al((size_t)g->Qp * kCap * 8);
  }
  g->total = o;
  return NRX_OK;
}

// ---- index build: fp32 corpus -> bf16 tile images + max row norm ------------------------------------------
__global__ void __launch_bounds__(256)
topk_pack_kernel(const float* __restrict__ c, long long ld, long long N, int D, int Dp, uint8_t* __restrict__ img,
                 unsigned* __restrict__ max_norm_bits) {
  const long long nchunks = ((N + kTR - 1) / kTR) * (Dp / 8) * kTR;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nchunks; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % kTR);
    const long long rest = i / kTR;
    const int kc = (int)(rest % (Dp / 8));
    const long long tile = rest / (Dp / 8);
    const long long row = tile * kTR + r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = kc * 8 + j;
      f[j] = (row < N && d < D) ? __ldg(c + row * ld + d) : 0.f;
    }
    *reinterpret_cast<uint4*>(img + (size_t)tile * kTR * Dp * 2 + canon_off(kTR, r, kc)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
  // max row norm (warp per row)
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  float m = 0.f;
  for (long long row = warp; row < N; row += nw) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) { const float t = __ldg(c + row * ld + d); s = fmaf(t, t, s); }
    s = warp_sum(s);
    m = fmaxf(m, s);
  }
  if (lane == 0 && m > 0.f) atomicMax(max_norm_bits, __float_as_uint(sqrtf(m) * 1.0001f));
}

// ---- queries: fp32 -> bf16 tile images, once per search (both scans bulk-load them) ---------------------------
__global__ void __launch_bounds__(256)
topk_qpack_kernel(const float* __restrict__ q, long long ld, long long Q, int D, int Dp, uint8_t* __restrict__ img, long long n_qtiles) {
  const long long nchunks = n_qtiles * (Dp / 8) * kTR;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(q) & 15) == 0) && (ld % 4 == 0) && (D % 8 == 0);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nchunks; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i % kTR);
    const long long rest = i / kTR;
    const int kc = (int)(rest % (Dp / 8));
    const long long tile = rest / (Dp / 8);
    const long long row = tile * kTR + r;
    float f[8];
    if (row < Q && vec_ok && kc * 8 < D) {
      const float4* src = reinterpret_cast<const float4*>(q + row * ld + kc * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (row < Q && kc * 8 + j < D) ? __ldg(q + row * ld + kc * 8 + j) : 0.f;
    }
    *reinterpret_cast<uint4*>(img + (size_t)tile * kTR * Dp * 2 + canon_off(kTR, r, kc)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

// ---- the scan (MODE 0: group maxima of the sample tiles, MODE 1: candidate filter over every tile) ------------
template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
topk_scan_kernel(const uint8_t* __restrict__ img, long long N, long long n_tiles, int tstride, int Dp,
                 const uint8_t* __restrict__ qimg, long long Qp, int nq_max, int stages, float* __restrict__ gmax,
                 long long n_groups, const float* __restrict__ theta, unsigned* __restrict__ count, unsigned* __restrict__ cand,
                 int cap_s) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const size_t tile_bytes = (size_t)kTR * Dp * 2;
  uint8_t* sQ = smem;                                  // [nq][tile image]
  uint8_t* sC = smem + (size_t)nq_max * tile_bytes;    // [stages][tile image]
  __shared__ uint64_t full[4], empty[4], tfull[kMaxQT], tempty[kMaxQT], qbar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(NRX_FULL_MASK, tid >> 5, 0);   // provably warp-uniform: the role branches below are uniform, so ptxas keeps the MMA descriptors in uniform registers
  const long long qt0 = (long long)blockIdx.y * nq_max;             // first query tile of this CTA
  const long long n_qtiles = Qp / kTR;
  const int nq = (int)(n_qtiles - qt0 < nq_max ? n_qtiles - qt0 : nq_max);
  // corpus slice of this CTA, in units of visited tiles (tile = visited index * tstride)
  const long long n_vis = (n_tiles + tstride - 1) / tstride;
  const long long per = (n_vis + gridDim.x - 1) / gridDim.x;
  const long long t0 = (long long)blockIdx.x * per;
  const long long t1 = min(t0 + per, n_vis);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < kMaxQT; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 16); }
    mbar_init(&qbar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, 512u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;

  if (t1 > t0) {
    if (warp == 0) {
      if (lane == 0) {  // TMA producer: the query tile images first (packed once per search), then the corpus stream
        mbar_expect_tx(&qbar, (uint32_t)(nq * tile_bytes));
        for (int a = 0; a < nq; ++a) bulk_g2s(sQ + (size_t)a * tile_bytes, qimg + (size_t)(qt0 + a) * tile_bytes, (uint32_t)tile_bytes, &qbar);
        uint32_t ph = 0, used = 0;
        int s = 0;
        for (long long t = t0; t < t1; ++t) {
          if (used & (1u << s)) { mbar_wait(&empty[s], (ph >> s) & 1u); ph ^= 1u << s; }
          used |= 1u << s;
          mbar_expect_tx(&full[s], (uint32_t)tile_bytes);
          bulk_g2s(sC + (size_t)s * tile_bytes, img + (size_t)(t * tstride) * tile_bytes, (uint32_t)tile_bytes, &full[s]);
          s = (s + 1 == stages) ? 0 : s + 1;
        }
      }
    } else if (warp == 1) {
      // MMA issuer: the whole warp walks the loop in uniform control flow, one elected lane issues (umma.cuh).
      // Query tiles are issued in pairs whose K-steps alternate between the two accumulators.
      const uint32_t leader = elect_one_sync() ? 1u : 0u;
      const uint32_t idesc = make_idesc_bf16(kTR, kTR);
      uint32_t fph = 0, eph = 0, aused = 0;
      int s = 0;
      const uint32_t kstep = (2u * (kTR * 16u)) >> 4;
      const uint32_t qstep = (uint32_t)(tile_bytes >> 4);
      const uint64_t qd0 = make_smem_desc(smem_u32(sQ), kTR * 16u, 128u);
      mbar_wait(&qbar, 0);
      for (long long t = t0; t < t1; ++t) {
        mbar_wait(&full[s], (fph >> s) & 1u); fph ^= 1u << s;
        const uint64_t cd0 = make_smem_desc(smem_u32(sC + (size_t)s * tile_bytes), kTR * 16u, 128u);
        for (int a0 = 0; a0 < nq; a0 += 2) {
          const bool two = a0 + 1 < nq;
          if (aused & (1u << a0)) { mbar_wait(&tempty[a0], (eph >> a0) & 1u); eph ^= 1u << a0; }
          if (two && (aused & (2u << a0))) { mbar_wait(&tempty[a0 + 1], (eph >> (a0 + 1)) & 1u); eph ^= 2u << a0; }
          aused |= (two ? 3u : 1u) << a0;
          tc_fence_after();
          uint64_t ad = qd0 + (uint64_t)a0 * qstep, bd = cd0;
          uint32_t accum = 0u;
          const uint32_t acc0 = tmem + (uint32_t)a0 * kTR;
          if (two) {
#pragma unroll 2
            for (int k16 = 0; k16 < Dp / 16; ++k16) {
              mma_bf16_ss_if(leader, acc0, ad, bd, idesc, accum);
              mma_bf16_ss_if(leader, acc0 + kTR, ad + qstep, bd, idesc, accum);
              ad += kstep; bd += kstep; accum = 1u;
            }
            mma_commit_if(leader, &tfull[a0]);
            mma_commit_if(leader, &tfull[a0 + 1]);
          } else {
#pragma unroll 2
            for (int k16 = 0; k16 < Dp / 16; ++k16) {
              mma_bf16_ss_if(leader, acc0, ad, bd, idesc, accum);
              ad += kstep; bd += kstep; accum = 1u;
            }
            mma_commit_if(leader, &tfull[a0]);
          }
        }
        mma_commit_if(leader, &empty[s]);
        __syncwarp();
        s = (s + 1 == stages) ? 0 : s + 1;
      }
    } else {  // epilogue: 16 warps = (TMEM lane quadrant, 32-column quarter of the tile); thread == query row
      const int qd = warp & 3;
      const int cq = (warp - 2) >> 2;
      const int r = qd * 32 + lane;
      uint32_t tph = 0;
      float th[kMaxQT];
      unsigned mycnt[kMaxQT];
#pragma unroll
      for (int a = 0; a < kMaxQT; ++a) {
        mycnt[a] = 0;
        th[a] = (MODE == 1 && a < nq) ? __ldg(theta + (qt0 + a) * kTR + r) : 0.f;   // theta is allocated for Qp rows
      }
      for (long long t = t0; t < t1; ++t) {
        const long long base = (t * tstride) * kTR + cq * 32;
        const bool tail = base + 32 > N;
#pragma unroll
        for (int a = 0; a < kMaxQT; ++a) {
          if (a < nq) {
            mbar_wait(&tfull[a], (tph >> a) & 1u); tph ^= 1u << a;
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a * kTR + cq * 32), v);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[a]);  // values are in registers: hand the accumulator back before the scan
            if (tail) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (base + j >= N) v[j] = __int_as_float(0xff800000);  // -inf: never admitted, even by theta = -FLT_MAX
            }
            float m4[4] = {v[0], v[1], v[2], v[3]};  // four independent max chains
#pragma unroll
            for (int j = 4; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
            const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            const long long qrow = (qt0 + a) * kTR + r;
            if (MODE == 0) {
              gmax[qrow * n_groups + t * 4 + cq] = m;   // groups of 32 corpus rows
            } else if (m >= th[a]) {  // rare: a few hundred rows out of N pass the threshold
              // this thread is the only writer of its (query, slice, quarter) region: plain stores, register counter.
              // Kept compact (hit mask + bit loop instead of 32 unrolled tests): the unrolled form made the epilogue
              // larger than the instruction cache (28 % of the stall samples were instruction fetch).
              unsigned hits = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) hits |= (v[j] >= th[a] ? 1u : 0u) << j;
              unsigned* mycand = cand + (((size_t)qrow * gridDim.x + blockIdx.x) * 4 + cq) * cap_s;
              while (hits) {
                const int j = __ffs(hits) - 1;
                hits &= hits - 1;
                if (mycnt[a] < (unsigned)cap_s) mycand[mycnt[a]] = (unsigned)(base + j);
                ++mycnt[a];
              }
            }
          }
        }
      }
      if (MODE == 1) {
#pragma unroll
        for (int a = 0; a < kMaxQT; ++a)
          if (a < nq) count[(((size_t)((qt0 + a) * kTR + r)) * gridDim.x + blockIdx.x) * 4 + cq] = mycnt[a];
      }
    }
  } else if (MODE == 1 && warp >= 2) {  // slice without tiles
    for (int a = 0; a < nq; ++a)
      count[(((size_t)((qt0 + a) * kTR + (warp & 3) * 32 + lane)) * gridDim.x + blockIdx.x) * 4 + ((warp - 2) >> 2)] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512u);
}

// ---- theta: k'-th largest group max per query (radix select on order-preserving uint keys) ------------------
__device__ __forceinline__ unsigned f2key(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

static constexpr int kFinalThreads = 128;   // 8 blocks per SM: Q = 1024 query blocks are ONE wave (256 threads: 4 per SM, two waves)
static constexpr int kThetaStage = 10240;   // group maxima staged in shared memory (40 KB)
__global__ void __launch_bounds__(256)
topk_theta_kernel(const float* __restrict__ gmax, long long n_tiles /* groups per query */, long long Qp, long long Q, int kprime,
                  const float* __restrict__ q, long long qld, int D, const unsigned* __restrict__ max_norm_bits,
                  float* __restrict__ theta, float* __restrict__ eps, unsigned* __restrict__ flist, int* __restrict__ flag) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining;
  __shared__ float s_norm[8];
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) flag[qi] = 0;
  if (qi == 0 && tid == 0) flist[0] = 0u;   // the fallback list of this search starts empty
  if (qi >= Q) { if (tid == 0) { theta[qi] = FLT_MAX; eps[qi] = 0.f; } return; }
  // |q|
  float ss = 0.f;
  for (int d = tid; d < D; d += 256) { const float t = __ldg(q + qi * qld + d); ss = fmaf(t, t, ss); }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) s_norm[tid >> 5] = ss;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_norm[w];
    const float qn = sqrtf(t) * 1.0001f;
    eps[qi] = 0.0055243f /* 2^-7.5 */ * qn * __uint_as_float(*max_norm_bits) + 1e-30f;
  }
  if (n_tiles < kprime) {  // not enough groups for a bound: admit everything (small corpus)
    if (tid == 0) theta[qi] = -FLT_MAX;
    return;
  }
  // the query's group maxima as order-preserving keys, staged once in shared memory when they fit (4 radix passes)
  extern __shared__ unsigned s_keys[];
  const bool staged = n_tiles <= kThetaStage;
  if (staged) {
    for (long long t = tid; t < n_tiles; t += 256) s_keys[t] = f2key(__ldg(gmax + qi * n_tiles + t));
    __syncthreads();
  }
  // bytes on which every key agrees need no pass (scores of one query share sign, exponent and often more: the first
  // pass would hammer ONE histogram bin with every key)
  unsigned kmin = 0xffffffffu, kmax = 0u;
  for (long long t = tid; t < n_tiles; t += 256) {
    const unsigned key = staged ? s_keys[t] : f2key(__ldg(gmax + qi * n_tiles + t));
    kmin = min(kmin, key);
    kmax = max(kmax, key);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    kmin = min(kmin, __shfl_xor_sync(NRX_FULL_MASK, kmin, o));
    kmax = max(kmax, __shfl_xor_sync(NRX_FULL_MASK, kmax, o));
  }
  __shared__ unsigned s_mm[16];
  if ((tid & 31) == 0) { s_mm[tid >> 5] = kmin; s_mm[8 + (tid >> 5)] = kmax; }
  __syncthreads();
  for (int w = 0; w < 8; ++w) { kmin = min(kmin, s_mm[w]); kmax = max(kmax, s_mm[8 + w]); }
  unsigned prefix = 0, mask = 0, remaining = (unsigned)kprime;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (((kmin ^ kmax) >> shift) == 0u) {   // all keys agree down to this byte
      prefix |= kmin & (255u << shift);
      mask |= 255u << shift;
      continue;
    }
    hist[tid] = 0;
    __syncthreads();
    for (long long t = tid; t < n_tiles; t += 256) {
      const unsigned key = staged ? s_keys[t] : f2key(__ldg(gmax + qi * n_tiles + t));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // bin of the `remaining`-th largest key, bins walked from 255 down: lane L owns bins 255 - 8L .. 248 - 8L
      unsigned c[8], tot = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[255 - 8 * tid - j]; tot += c[j]; }
      unsigned inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(NRX_FULL_MASK, inc, o); if (tid >= o) inc += v; }
      const unsigned before = inc - tot;                         // keys in the bins above this lane's
      const bool here = before < remaining && remaining <= inc;  // exactly one lane (remaining <= number of keys)
      if (here) {
        unsigned acc = before;
        int b = 255 - 8 * tid;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (acc + c[j] >= remaining) { b = 255 - 8 * tid - j; break; }
          acc += c[j];
        }
        s_prefix = prefix | ((unsigned)b << shift);
        s_remaining = remaining - acc;
      }
    }
    __syncthreads();
    prefix = s_prefix;
    remaining = s_remaining;
    mask |= 255u << shift;
    __syncthreads();
  }
  if (tid == 0) theta[qi] = key2f(prefix);
}

// ---- exact scoring + ordering -------------------------------------------------------------------------------
// The ONE exact scoring function (final re-scoring and the fallback scan both use it, so equal rows always get bit-equal
// scores): a warp scores one row — lane l takes elements 4l .. 4l+3 of every 128-element block (coalesced 16-byte
// loads), accumulates in fp64, and the lane partials are combined by a fixed xor-shuffle tree.  Every lane returns
// the same value.  `qs` is the query in shared memory.
__device__ __forceinline__ double dot64_warp(const float* __restrict__ qs, const float* __restrict__ row, int D, int lane, bool vec) {
  double s = 0.0;
  if (vec) {
    for (int d = lane * 4; d < D; d += 128) {
      const float4 c = __ldg(reinterpret_cast<const float4*>(row + d));
      const float4 q = *reinterpret_cast<const float4*>(qs + d);
      s = fma((double)q.x, (double)c.x, s);
      s = fma((double)q.y, (double)c.y, s);
      s = fma((double)q.z, (double)c.z, s);
      s = fma((double)q.w, (double)c.w, s);
    }
  } else {
    for (int d0 = lane * 4; d0 < D; d0 += 128) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (d0 + j < D) s = fma((double)qs[d0 + j], (double)__ldg(row + d0 + j), s);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(NRX_FULL_MASK, s, o);
  return s;
}
// Four rows at once: the four 16-byte loads of a 128-element block are issued before the first FMA (one memory
// latency per block of four candidates instead of four).  Same per-row arithmetic, bit for bit, as dot64_warp: the
// cross-lane sum is the same xor-butterfly tree (levels 16, 8, 4, 2, 1; fp addition commutes), but the first two levels
// halve the payload instead of keeping all four rows on every lane — 7 instead of 20 fp64 shuffles per call.
// Returns the score of row dot64_row_of_lane(lane) (every lane holds one of the four rows).
__device__ __forceinline__ int dot64_row_of_lane(int lane) { return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); }
__device__ __forceinline__ int dot64_lane_of_row(int u) { return ((u >> 1) << 4) | ((u & 1) << 3); }
__device__ __forceinline__ double dot64_warp4(const float* __restrict__ qs, const float* const (&row)[4], int D, int lane, bool vec) {
  if (!vec) {
    double out[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) out[u] = dot64_warp(qs, row[u], D, lane, false);
    const int u = dot64_row_of_lane(lane);
    return u == 0 ? out[0] : u == 1 ? out[1] : u == 2 ? out[2] : out[3];
  }
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  for (int d = lane * 4; d < D; d += 128) {
    float4 c[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = __ldg(reinterpret_cast<const float4*>(row[u] + d));
    const float4 q = *reinterpret_cast<const float4*>(qs + d);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = fma((double)q.x, (double)c[u].x, s[u]);
      s[u] = fma((double)q.y, (double)c[u].y, s[u]);
      s[u] = fma((double)q.z, (double)c[u].z, s[u]);
      s[u] = fma((double)q.w, (double)c[u].w, s[u]);
    }
  }
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  // level 16: lanes 0-15 keep rows 0, 1 and hand rows 2, 3 to their partner (and the other way round)
  const double k0 = hi16 ? s[2] : s[0], k1 = hi16 ? s[3] : s[1];
  const double g0 = hi16 ? s[0] : s[2], g1 = hi16 ? s[1] : s[3];
  const double a0 = k0 + __shfl_xor_sync(NRX_FULL_MASK, g0, 16);
  const double a1 = k1 + __shfl_xor_sync(NRX_FULL_MASK, g1, 16);
  // level 8: keep one of the two
  double b = (hi8 ? a1 : a0) + __shfl_xor_sync(NRX_FULL_MASK, hi8 ? a0 : a1, 8);
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) b += __shfl_xor_sync(NRX_FULL_MASK, b, o);
  return b;
}
__device__ __forceinline__ bool rows_vec_ok(const float* c, long long cld, int D) {
  return ((reinterpret_cast<uintptr_t>(c) & 15) == 0) && (cld % 4 == 0) && (D % 4 == 0);
}

// (score desc, id asc): returns true if a must come before b
template <typename IdT>
__device__ __forceinline__ bool before(double sa, IdT ia, double sb, IdT ib) { return sa > sb || (sa == sb && ia < ib); }

// One compare-exchange of the bitonic network done with shuffles: element i = (sv, iv) against element i ^ j (j < 32: the
// partner sits in the same warp).  Position i keeps the element that sorts first iff (i is the lower index) == (ascending
// block); equal elements (padding) are interchangeable.
template <typename IdT>
__device__ __forceinline__ void bitonic_shfl_step(double& sv, IdT& iv, int i, int j, int k2) {
  const double ps = __shfl_xor_sync(NRX_FULL_MASK, sv, j);
  const IdT pi = __shfl_xor_sync(NRX_FULL_MASK, iv, j);
  const bool want_first = ((i & j) == 0) == ((i & k2) == 0);
  const bool mine_first = before(sv, iv, ps, pi);
  if (want_first != mine_first) { sv = ps; iv = pi; }
}

// Block-wide bitonic sort of n (power of two, >= 32) pairs in shared memory by (score desc, id asc).  Exchange distances
// below 32 stay inside a warp and run on registers + shuffles with no block barrier (all of k2 <= 32 in ONE pass); only the
// distances >= 32 go through shared memory.  n = 128: 3 barrier stages instead of 28.  nthreads: a multiple of 32; the
// caller has synchronised its writes to s / id; the result is visible to the whole block on return.
template <typename IdT>
__device__ void bitonic_sort(double* s, IdT* id, int n, int tid, int nthreads) {
  for (int i = tid; i < n; i += nthreads) {          // k2 = 2 .. 32 entirely in registers
    double sv = s[i];
    IdT iv = id[i];
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
      for (int j = k2 >> 1; j > 0; j >>= 1) bitonic_shfl_step(sv, iv, i, j, k2);
    s[i] = sv; id[i] = iv;
  }
  __syncthreads();
  for (int k2 = 64; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j >= 32; j >>= 1) {
      for (int i = tid; i < n; i += nthreads) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = ((i & k2) == 0);
          const bool sw = up ? before(s[p], id[p], s[i], id[i]) : before(s[i], id[i], s[p], id[p]);
          if (sw) {
            const double ts = s[i]; s[i] = s[p]; s[p] = ts;
            const IdT ti = id[i]; id[i] = id[p]; id[p] = ti;
          }
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += nthreads) {
      double sv = s[i];
      IdT iv = id[i];
#pragma unroll
      for (int j = 16; j > 0; j >>= 1) bitonic_shfl_step(sv, iv, i, j, k2);
      s[i] = sv; id[i] = iv;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void flag_query(long long qi, int* flag, unsigned* flist) {
  flag[qi] = 1;
  flist[1 + atomicAdd(flist, 1u)] = (unsigned)qi;
}

// Exact re-scoring as its own WIDE kernel: grid (chunks, Q).  Inside the per-query final kernel the ~350 candidate rows of
// a query were fetched by 4 warps in ~22 dependent rounds (one DRAM round trip each); here 16 blocks x 8 warps per query
// have every row of every query in flight at once (random 512-byte rows: DRAM-bound).  Every block recomputes the query's
// region offsets (a few hundred counts) and re-scores candidates [chunk * 32 + m * 32 * gridDim.x, +32); the final kernel
// then only sorts and proves.  Same dot64_warp4 arithmetic as the fused path, bit for bit.  Used for small query batches
// (make_geom: Q <= 128), where Q blocks of the final kernel cannot fill 148 SMs.
static constexpr int kRescoreChunks = 16;
__global__ void __launch_bounds__(256)
topk_rescore_kernel(const float* __restrict__ c, long long cld, int D, const float* __restrict__ q, long long qld,
                    const unsigned* __restrict__ count, const unsigned* __restrict__ cand, int n_slices, int cap_s,
                    unsigned* __restrict__ pre_id, double* __restrict__ pre_sc) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  float* qs = reinterpret_cast<float*>(sm_raw);                           // [D]
  unsigned* s_off = reinterpret_cast<unsigned*>(qs + ((D + 3) & ~3));     // [n_slices]
  const long long qi = blockIdx.y;
  const int tid = threadIdx.x;
  __shared__ unsigned s_total, s_over, s_warp[8];
  unsigned over = 0, run = 0;
  for (int b0 = 0; b0 < n_slices; b0 += 256) {
    const int sl = b0 + tid;
    unsigned cs = sl < n_slices ? count[qi * n_slices + sl] : 0u;
    if (cs > (unsigned)cap_s) { over = 1; cs = (unsigned)cap_s; }
    unsigned inc = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(NRX_FULL_MASK, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < (tid >> 5); ++w) wbase += s_warp[w];
    if (sl < n_slices) s_off[sl] = run + wbase + inc - cs;
    unsigned tot = 0;
    for (int w = 0; w < 8; ++w) tot += s_warp[w];
    run += tot;
    __syncthreads();
  }
  over = __syncthreads_or(over);
  if (tid == 0) { s_total = run; s_over = over; }
  for (int d = tid; d < D; d += 256) qs[d] = __ldg(q + qi * qld + d);
  __syncthreads();
  const unsigned cnt = s_total;
  if (s_over || cnt > (unsigned)kCap) return;   // the final kernel sends the query to the exact scan
  const int warp = tid >> 5, lane = tid & 31;
  const bool vec = rows_vec_ok(c, cld, D);
  for (unsigned base = blockIdx.x * 32u; base < cnt; base += gridDim.x * 32u) {
    const unsigned i0 = base + (unsigned)warp * 4;
    if (i0 >= cnt) continue;
    unsigned myid = 0;
    if (lane < 4) {   // lanes 0..3 look up the four candidate rows of this warp
      const unsigned i = min(i0 + (unsigned)lane, cnt - 1);   // past the end: re-score the last one
      int lo = 0, hi = n_slices - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
      }
      myid = cand[((size_t)qi * n_slices + lo) * cap_s + (i - s_off[lo])];
    }
    const float* rowp[4];
    unsigned ids[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { ids[u] = __shfl_sync(NRX_FULL_MASK, myid, u); rowp[u] = c + (long long)ids[u] * cld; }
    const double v = dot64_warp4(qs, rowp, D, lane, vec);
    const int u = dot64_row_of_lane(lane);
    if ((lane & 7) == 0 && i0 + u < cnt) {
      pre_sc[(size_t)qi * kCap + i0 + u] = v;
      pre_id[(size_t)qi * kCap + i0 + u] = u == 0 ? ids[0] : u == 1 ? ids[1] : u == 2 ? ids[2] : ids[3];
    }
  }
}

__global__ void __launch_bounds__(kFinalThreads, 8)
topk_final_kernel(const float* __restrict__ c, long long cld, long long N, int D, const float* __restrict__ q, long long qld,
                  int k, long long id_base, const float* __restrict__ theta, const float* __restrict__ eps,
                  const unsigned* __restrict__ count, const unsigned* __restrict__ cand, int n_slices, int cap_s,
                  int* __restrict__ flag, unsigned* __restrict__ flist, float* __restrict__ out_s, double* __restrict__ out_s64,
                  long long* __restrict__ out_i, const unsigned* __restrict__ pre_id, const double* __restrict__ pre_sc) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* s = reinterpret_cast<double*>(sm_raw);              // [kCap]
  unsigned* id = reinterpret_cast<unsigned*>(s + kCap);       // [kCap]
  float* qs = reinterpret_cast<float*>(id + kCap);            // [D]
  unsigned* s_off = reinterpret_cast<unsigned*>(qs + ((D + 3) & ~3));   // [n_slices]
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  __shared__ unsigned s_total, s_over, s_warp[kFinalThreads / 32];
  const long long kk = k < N ? k : N;
  // exclusive scan of the per-region counts (n_slices = corpus slices x 4 column quarters)
  unsigned over = 0, run = 0;
  for (int b0 = 0; b0 < n_slices; b0 += kFinalThreads) {
    const int sl = b0 + tid;
    unsigned cs = sl < n_slices ? count[qi * n_slices + sl] : 0u;
    if (cs > (unsigned)cap_s) { over = 1; cs = (unsigned)cap_s; }
    unsigned inc = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(NRX_FULL_MASK, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < (tid >> 5); ++w) wbase += s_warp[w];
    if (sl < n_slices) s_off[sl] = run + wbase + inc - cs;
    unsigned tot = 0;
    for (int w = 0; w < kFinalThreads / 32; ++w) tot += s_warp[w];
    run += tot;
    __syncthreads();
  }
  over = __syncthreads_or(over);
  if (tid == 0) { s_total = run; s_over = over; }
  for (int d = tid; d < D; d += kFinalThreads) qs[d] = __ldg(q + qi * qld + d);
  __syncthreads();
  const unsigned cnt = s_total;
  if (s_over || cnt > (unsigned)kCap || (long long)cnt < kk) {
    if (tid == 0) flag_query(qi, flag, flist);
    return;
  }
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < (int)cnt) n2 <<= 1;
  for (int i = (int)cnt + tid; i < n2; i += kFinalThreads) { id[i] = 0xffffffffu; s[i] = -DBL_MAX; }
  if (pre_id != nullptr) {   // re-scored by topk_rescore_kernel: only fetch (id, fp64 score)
    for (unsigned i = tid; i < cnt; i += kFinalThreads) { id[i] = pre_id[(size_t)qi * kCap + i]; s[i] = pre_sc[(size_t)qi * kCap + i]; }
  } else {
    for (unsigned i = tid; i < cnt; i += kFinalThreads) {  // candidate row of list position i: region by binary search
      int lo = 0, hi = n_slices - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
      }
      id[i] = cand[((size_t)qi * n_slices + lo) * cap_s + (i - s_off[lo])];
    }
    __syncthreads();
    {  // exact scores: one warp per candidate, 4 rows in flight per warp
      const int warp = tid >> 5, lane = tid & 31;
      const bool vec = rows_vec_ok(c, cld, D);
      for (unsigned i0 = (unsigned)warp * 4; i0 < cnt; i0 += 4 * (kFinalThreads / 32)) {
        const float* rowp[4];
  #pragma unroll
        for (int u = 0; u < 4; ++u) rowp[u] = c + (long long)id[min(i0 + u, cnt - 1)] * cld;   // past the end: re-score the last one
        const double v = dot64_warp4(qs, rowp, D, lane, vec);
        if ((lane & 7) == 0 && i0 + dot64_row_of_lane(lane) < cnt) s[i0 + dot64_row_of_lane(lane)] = v;
      }
    }
  }
  __syncthreads();
  bitonic_sort(s, id, n2, tid, kFinalThreads);
  // completeness proof: every row outside the list scores < theta + eps
  const bool ok = (theta[qi] == -FLT_MAX) || (kk == 0) || (s[kk - 1] >= (double)theta[qi] + (double)eps[qi]);
  if (!ok) {
    if (tid == 0) flag_query(qi, flag, flist);
    return;
  }
  for (int i = tid; i < k; i += kFinalThreads) {
    if (i < kk) {
      out_s[qi * k + i] = (float)s[i];
      if (out_s64) out_s64[qi * k + i] = s[i];
      out_i[qi * k + i] = (long long)id[i] + id_base;
    } else {
      out_s[qi * k + i] = -FLT_MAX;
      if (out_s64) out_s64[qi * k + i] = -DBL_MAX;
      out_i[qi * k + i] = -1;
    }
  }
}

// Exact fallback: fp64 scan with a block-level streaming top-k (buffer + periodic bitonic prune).  Fixed grid; the work
// list is (listed query, corpus slice): S = grid / n_listed slices per query (>= 1), every item leaves its slice's top-k
// (fp64 score, row) in `part`; topk_exact_merge_kernel combines a query's S lists.  An empty list costs two empty launches.
static constexpr int kFbCap = 2048;
__host__ __device__ __forceinline__ int fb_max_slices(int k) { const int m = 4096 / k; return m < 1 ? 1 : (m > 64 ? 64 : m); }
__device__ __forceinline__ int fb_slices(unsigned n_listed, int grid, int max_items, int k) {
  if (n_listed == 0) return 1;
  long long S = grid / (long long)n_listed;
  if (S < 1) S = 1;
  if (S > fb_max_slices(k)) S = fb_max_slices(k);
  while (S > 1 && S * n_listed > (unsigned)max_items) --S;
  return (int)S;
}

// The corpus as up to NRX_MAX_PEERS row segments (one locally; in the sharded peer search: every rank's shard, read over
// NVLink by the owner of a listed query).  Global row = segment base + local row.
struct CorpusSegs {
  const float* c[NRX_MAX_PEERS];
  long long base[NRX_MAX_PEERS + 1];   // prefix sums of the segment sizes; base[n_seg] = N
  int n_seg;
};
__device__ __forceinline__ const float* seg_row(const CorpusSegs& G, long long row, long long cld) {
  int g = 0;
  while (g + 1 < G.n_seg && row >= G.base[g + 1]) ++g;
  return G.c[g] + (row - G.base[g]) * cld;
}

__global__ void __launch_bounds__(256)
topk_exact_kernel(const __grid_constant__ CorpusSegs CS, long long cld, long long N, int D, const float* __restrict__ q, long long qld,
                  int k, const unsigned* __restrict__ flist, int force_Q, int max_items, double* __restrict__ part_s,
                  unsigned* __restrict__ part_i) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* s = reinterpret_cast<double*>(sm_raw);
  unsigned* id = reinterpret_cast<unsigned*>(s + kFbCap);
  float* qs = reinterpret_cast<float*>(id + kFbCap);
  __shared__ unsigned s_cnt;
  __shared__ double s_th;
  const int tid = threadIdx.x;
  const unsigned n_listed = force_Q > 0 ? (unsigned)force_Q : flist[0];
  if (n_listed == 0) return;
  const int S = fb_slices(n_listed, gridDim.x, max_items, k);
  const long long n_items = (long long)n_listed * S;
  const long long kk = k < N ? k : N;
  for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
    const long long qi = force_Q > 0 ? item / S : (long long)flist[1 + item / S];
    const int sl = (int)(item % S);
    const long long r_lo = N * sl / S, r_hi = N * (sl + 1) / S;
    __syncthreads();
    for (int d = tid; d < D; d += 256) qs[d] = __ldg(q + qi * qld + d);
    if (tid == 0) { s_cnt = 0; s_th = -DBL_MAX; }
    __syncthreads();
    const int warp = tid >> 5, lane = tid & 31;
    bool vec = (cld % 4 == 0) && (D % 4 == 0);
    for (int g = 0; g < CS.n_seg; ++g) vec = vec && ((reinterpret_cast<uintptr_t>(CS.c[g]) & 15) == 0);
    for (long long r0 = r_lo; r0 < r_hi; r0 += 256) {
      // 256 rows per round: every warp scores 32 rows (four at a time, coalesced), lane u keeps row u's score
      double mine = 0.0;
      for (int u = 0; u < 32; u += 4) {
        const long long rb = r0 + warp * 32 + u;
        if (rb >= r_hi) break;
        const float* rowp[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) rowp[j] = seg_row(CS, rb + j < r_hi ? rb + j : r_hi - 1, cld);
        const double v = dot64_warp4(qs, rowp, D, lane, vec);
        const double vj = __shfl_sync(NRX_FULL_MASK, v, dot64_lane_of_row(lane & 3));   // lane u + j takes row j
        if ((lane & ~3) == u) mine = vj;
      }
      const long long row = r0 + tid;
      if (row < r_hi && mine >= s_th) {
        const unsigned slot = atomicAdd(&s_cnt, 1u);
        s[slot] = mine;           // slot < kFbCap: pruned whenever fewer than 256 free slots remain
        id[slot] = (unsigned)row;
      }
      __syncthreads();
      if (s_cnt > (unsigned)(kFbCap - 256) || r0 + 256 >= r_hi) {
        const unsigned cnt = s_cnt;
        for (int i = tid; i < kFbCap; i += 256)
          if (i >= (int)cnt) { s[i] = -DBL_MAX; id[i] = 0xffffffffu; }
        __syncthreads();
        bitonic_sort(s, id, kFbCap, tid, 256);
        if (tid == 0) {
          const unsigned keep = cnt < (unsigned)kk ? cnt : (unsigned)kk;
          s_cnt = keep;
          s_th = (keep == (unsigned)kk && kk > 0) ? s[kk - 1] : -DBL_MAX;
        }
        __syncthreads();
      }
    }
    if (r_hi <= r_lo) {   // empty slice
      for (int i = tid; i < kFbCap; i += 256) { s[i] = -DBL_MAX; id[i] = 0xffffffffu; }
      __syncthreads();
    }
    for (int i = tid; i < k; i += 256) {
      part_s[item * k + i] = s[i];         // sorted; entries past the slice's row count are (-DBL_MAX, 0xffffffff)
      part_i[item * k + i] = id[i];
    }
  }
}

struct PeerOuts {   // the result buffers of every rank (peer-mapped); n == 0: write the local out_s / out_i only
  float* s[NRX_MAX_PEERS];
  long long* i[NRX_MAX_PEERS];
  int n;
};

__global__ void __launch_bounds__(256)
topk_exact_merge_kernel(long long N, int k, long long id_base, const unsigned* __restrict__ flist, int force_Q, int exact_grid,
                        int max_items, const double* __restrict__ part_s, const unsigned* __restrict__ part_i,
                        float* __restrict__ out_s, double* __restrict__ out_s64, long long* __restrict__ out_i,
                        int* __restrict__ status, const __grid_constant__ PeerOuts PO) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int tid = threadIdx.x;
  const unsigned n_listed = force_Q > 0 ? (unsigned)force_Q : flist[0];
  if (n_listed == 0) return;
  const int S = fb_slices(n_listed, exact_grid, max_items, k);
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < S * k) n2 <<= 1;
  double* s = reinterpret_cast<double*>(sm_raw);          // [n2]
  unsigned* id = reinterpret_cast<unsigned*>(s + n2);     // [n2]
  const long long kk = k < N ? k : N;
  for (long long e = blockIdx.x; e < n_listed; e += gridDim.x) {
    const long long qi = force_Q > 0 ? e : (long long)flist[1 + e];
    __syncthreads();
    for (int i = tid; i < n2; i += 256) {
      if (i < S * k) { s[i] = part_s[e * S * k + i]; id[i] = part_i[e * S * k + i]; }
      else { s[i] = -DBL_MAX; id[i] = 0xffffffffu; }
    }
    __syncthreads();
    if (S > 1) bitonic_sort(s, id, n2, tid, 256);
    for (int i = tid; i < k; i += 256) {
      const float fs = i < kk ? (float)s[i] : -FLT_MAX;
      const long long fi = i < kk ? (long long)id[i] + id_base : -1;
      if (PO.n > 0) {
        for (int g = 0; g < PO.n; ++g) { PO.s[g][qi * k + i] = fs; PO.i[g][qi * k + i] = fi; }
      } else {
        out_s[qi * k + i] = fs;
        if (out_s64) out_s64[qi * k + i] = i < kk ? s[i] : -DBL_MAX;
        out_i[qi * k + i] = fi;
      }
    }
    if (status != nullptr && tid == 0 && force_Q == 0) status[qi] = 1;
  }
}


// ---- sharded search over peer memory (2-D: every rank scans ITS corpus shard for ALL queries, the owner of a query —
// rank q / ceil(Q / G) — merges the shards' lists and proves completeness) -------------------------------------------
// Inbox of an owner: [Qown][G][k] entries {fp64 score, global id}, then counts int32 [Qown][G] (-1: the shard's candidate
// list overflowed), then bounds float [Qown][G] (theta_g + eps_g: every row of shard g outside its list scores below it).
struct TopkEntry { double s; long long id; };
struct PeerInbox {
  uint8_t* box[NRX_MAX_PEERS];
  int rank, world;
  long long q_own;     // queries per owner
};
__host__ __device__ __forceinline__ size_t inbox_bytes(long long q_own, int world, int k) {
  return (size_t)q_own * world * ((size_t)k * sizeof(TopkEntry) + 8);
}
__device__ __forceinline__ TopkEntry* inbox_entries(uint8_t* box, long long ql, int g, int world, int k) {
  return reinterpret_cast<TopkEntry*>(box) + (ql * world + g) * k;
}
__device__ __forceinline__ int* inbox_cnt(uint8_t* box, long long q_own, int world, int k) {
  return reinterpret_cast<int*>(box + (size_t)q_own * world * k * sizeof(TopkEntry));
}
__device__ __forceinline__ float* inbox_bound(uint8_t* box, long long q_own, int world, int k) {
  return reinterpret_cast<float*>(box + (size_t)q_own * world * ((size_t)k * sizeof(TopkEntry) + 4));
}

// Shard side: exact re-scoring + sort of this shard's candidates of query qi (as topk_final_kernel), then the best
// min(count, k) go straight into the owner's inbox (peer stores) with the shard's completeness bound.
__global__ void __launch_bounds__(kFinalThreads, 8)
topk_final_peer_kernel(const float* __restrict__ c, long long cld, long long N, int D, const float* __restrict__ q, long long qld,
                       int k, long long id_base, const float* __restrict__ theta, const float* __restrict__ eps,
                       const unsigned* __restrict__ count, const unsigned* __restrict__ cand, int n_slices, int cap_s,
                       const __grid_constant__ PeerInbox PB, const unsigned* __restrict__ pre_id,
                       const double* __restrict__ pre_sc) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* s = reinterpret_cast<double*>(sm_raw);              // [kCap]
  unsigned* id = reinterpret_cast<unsigned*>(s + kCap);       // [kCap]
  float* qs = reinterpret_cast<float*>(id + kCap);            // [D]
  unsigned* s_off = reinterpret_cast<unsigned*>(qs + ((D + 3) & ~3));   // [n_slices]
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  __shared__ unsigned s_total, s_over, s_warp[kFinalThreads / 32];
  const int owner = (int)(qi / PB.q_own);
  const long long ql = qi % PB.q_own;
  uint8_t* box = PB.box[owner];
  unsigned over = 0, run = 0;
  for (int b0 = 0; b0 < n_slices; b0 += kFinalThreads) {
    const int sl = b0 + tid;
    unsigned cs = sl < n_slices ? count[qi * n_slices + sl] : 0u;
    if (cs > (unsigned)cap_s) { over = 1; cs = (unsigned)cap_s; }
    unsigned inc = cs;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(NRX_FULL_MASK, inc, o); if ((tid & 31) >= o) inc += t; }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    unsigned wbase = 0;
    for (int w = 0; w < (tid >> 5); ++w) wbase += s_warp[w];
    if (sl < n_slices) s_off[sl] = run + wbase + inc - cs;
    unsigned tot = 0;
    for (int w = 0; w < kFinalThreads / 32; ++w) tot += s_warp[w];
    run += tot;
    __syncthreads();
  }
  over = __syncthreads_or(over);
  if (tid == 0) { s_total = run; s_over = over; }
  for (int d = tid; d < D; d += kFinalThreads) qs[d] = __ldg(q + qi * qld + d);
  __syncthreads();
  const unsigned cnt = s_total;
  if (s_over || cnt > (unsigned)kCap) {   // the shard cannot vouch for its list: the owner sends the query to the exact scan
    if (tid == 0) { inbox_cnt(box, PB.q_own, PB.world, k)[ql * PB.world + PB.rank] = -1; inbox_bound(box, PB.q_own, PB.world, k)[ql * PB.world + PB.rank] = 0.f; }
    return;
  }
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < (int)cnt) n2 <<= 1;
  for (int i = (int)cnt + tid; i < n2; i += kFinalThreads) { id[i] = 0xffffffffu; s[i] = -DBL_MAX; }
  if (pre_id != nullptr) {   // re-scored by topk_rescore_kernel: only fetch (id, fp64 score)
    for (unsigned i = tid; i < cnt; i += kFinalThreads) { id[i] = pre_id[(size_t)qi * kCap + i]; s[i] = pre_sc[(size_t)qi * kCap + i]; }
  } else {
    for (unsigned i = tid; i < cnt; i += kFinalThreads) {
      int lo = 0, hi = n_slices - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
      }
      id[i] = cand[((size_t)qi * n_slices + lo) * cap_s + (i - s_off[lo])];
    }
    __syncthreads();
    {
      const int warp = tid >> 5, lane = tid & 31;
      const bool vec = rows_vec_ok(c, cld, D);
      for (unsigned i0 = (unsigned)warp * 4; i0 < cnt; i0 += 4 * (kFinalThreads / 32)) {
        const float* rowp[4];
  #pragma unroll
        for (int u = 0; u < 4; ++u) rowp[u] = c + (long long)id[min(i0 + u, cnt - 1)] * cld;
        const double v = dot64_warp4(qs, rowp, D, lane, vec);
        if ((lane & 7) == 0 && i0 + dot64_row_of_lane(lane) < cnt) s[i0 + dot64_row_of_lane(lane)] = v;
      }
    }
  }
  __syncthreads();
  bitonic_sort(s, id, n2, tid, kFinalThreads);
  const int n_send = (int)cnt < k ? (int)cnt : k;
  TopkEntry* dst = inbox_entries(box, ql, PB.rank, PB.world, k);
  for (int i = tid; i < n_send; i += kFinalThreads) { TopkEntry e; e.s = s[i]; e.id = (long long)id[i] + id_base; dst[i] = e; }
  if (tid == 0) {
    inbox_cnt(box, PB.q_own, PB.world, k)[ql * PB.world + PB.rank] = n_send;
    // rows of this shard outside the list score < theta + eps, or are beaten by k listed rows of the same shard
    inbox_bound(box, PB.q_own, PB.world, k)[ql * PB.world + PB.rank] =
        theta[qi] == -FLT_MAX ? -FLT_MAX : (float)((double)theta[qi] + (double)eps[qi]) ;
  }
}

// Owner side: merge the G lists of one owned query, prove completeness, publish the result to EVERY rank (peer stores);
// a query that cannot be proven goes on the owner's fallback list (exact scan over all shards through peer memory).
// The lists arrive sorted, so the merge is a RANK merge, not a sort: the final position of an element is its index in its
// own list plus, for every other list, the number of elements that sort before it (one binary search each; ids are unique
// across shards, so the order is strict).  8 lists x 100: 23.5 us as a 1024-element bitonic sort, a few us this way.
__device__ __forceinline__ int count_before(const double* ls, const long long* li, int n, double s, long long id) {
  int lo = 0, hi = n;   // first index whose element does NOT sort before (s, id)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (before(ls[mid], li[mid], s, id)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
topk_owner_merge_kernel(long long Q, long long N_total, int k, const __grid_constant__ PeerInbox PB, unsigned* __restrict__ flist,
                        int* __restrict__ status, const __grid_constant__ PeerOuts PO) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const long long ql = blockIdx.x;
  const long long qi = (long long)PB.rank * PB.q_own + ql;
  const int tid = threadIdx.x;
  if (qi >= Q) return;
  uint8_t* box = PB.box[PB.rank];
  const int G = PB.world;
  double* s = reinterpret_cast<double*>(sm_raw);                    // [G][k]
  long long* id = reinterpret_cast<long long*>(s + (size_t)G * k);   // [G][k]
  double* os = reinterpret_cast<double*>(id + (size_t)G * k);       // [k] merged
  long long* oi = reinterpret_cast<long long*>(os + k);             // [k]
  __shared__ int s_cnt[NRX_MAX_PEERS];
  __shared__ int s_bad, s_total;
  __shared__ float s_bound;
  if (tid == 0) {
    int bad = 0, total = 0;
    float bound = -FLT_MAX;
    for (int g = 0; g < G; ++g) {
      const int cg = inbox_cnt(box, PB.q_own, G, k)[ql * G + g];
      if (cg < 0) bad = 1;
      s_cnt[g] = cg < 0 ? 0 : cg;
      total += s_cnt[g];
      bound = fmaxf(bound, inbox_bound(box, PB.q_own, G, k)[ql * G + g]);
    }
    s_bad = bad;
    s_total = total;
    s_bound = bound;
  }
  __syncthreads();
  const long long kk = k < N_total ? k : N_total;
  bool ok = !s_bad && (long long)s_total >= kk;
  if (ok) {
    for (int i = tid; i < G * k; i += 256) {
      const int g = i / k, j = i - g * k;
      if (j < s_cnt[g]) { const TopkEntry e = inbox_entries(box, ql, g, G, k)[j]; s[i] = e.s; id[i] = e.id; }
    }
    __syncthreads();
    for (int i = tid; i < G * k; i += 256) {
      const int g = i / k, j = i - g * k;
      if (j >= s_cnt[g]) continue;
      const double es = s[i];
      const long long ei = id[i];
      int rank = j;
      for (int g2 = 0; g2 < G && rank < kk; ++g2)
        if (g2 != g) rank += count_before(s + (size_t)g2 * k, id + (size_t)g2 * k, s_cnt[g2], es, ei);
      if (rank < kk) { os[rank] = es; oi[rank] = ei; }
    }
    __syncthreads();
    // complete iff the k-th merged score clears every shard's bound
    ok = kk == 0 || s_bound == -FLT_MAX || os[kk - 1] >= (double)s_bound;
  }
  if (!ok) {
    if (tid == 0) { flist[1 + atomicAdd(flist, 1u)] = (unsigned)qi; if (status) status[qi] = 1; }
    return;
  }
  for (int i = tid; i < k; i += 256) {
    const float fs = i < kk ? (float)os[i] : -FLT_MAX;
    const long long fi = i < kk ? oi[i] : -1;
    for (int g = 0; g < PO.n; ++g) { PO.s[g][qi * k + i] = fs; PO.i[g][qi * k + i] = fi; }
  }
}

// ---- merge of per-shard lists ([n_lists][Q][k]) ----------------------------------------------------------------
// ST = float: lists carry fp32-rounded scores (two rows whose fp64 scores differ by less than one fp32 ulp then
// merge in id order); ST = double: the per-shard fp64 ordering keys, the merge equals one index bit for bit.
template <typename ST>
__global__ void __launch_bounds__(256)
topk_merge_kernel(const ST* __restrict__ sc, const long long* __restrict__ ids, int n_lists, long long Q, int k,
                  float* __restrict__ out_s, long long* __restrict__ out_i) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int total = n_lists * k;
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < total) n2 <<= 1;
  long long* id = reinterpret_cast<long long*>(sm_raw);
  ST* s = reinterpret_cast<ST*>(id + n2);
  const ST lowest = (ST)(sizeof(ST) == 8 ? -DBL_MAX : -FLT_MAX);
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < n2; i += 256) {
    if (i < total) {
      const int l = i / k, j = i % k;
      s[i] = sc[((long long)l * Q + qi) * k + j];
      id[i] = ids[((long long)l * Q + qi) * k + j];
      if (id[i] < 0) { s[i] = lowest; id[i] = 0x7fffffffffffffffll; }
    } else { s[i] = lowest; id[i] = 0x7fffffffffffffffll; }
  }
  __syncthreads();
  for (int k2 = 2; k2 <= n2; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += 256) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = ((i & k2) == 0);
          const bool a_first = s[p] > s[i] || (s[p] == s[i] && id[p] < id[i]);   // p before i
          const bool b_first = s[i] > s[p] || (s[i] == s[p] && id[i] < id[p]);
          if (up ? a_first : b_first) {
            const ST ts = s[i]; s[i] = s[p]; s[p] = ts;
            const long long ti = id[i]; id[i] = id[p]; id[p] = ti;
          }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < k; i += 256) {
    const bool pad = i >= total || id[i] == 0x7fffffffffffffffll;
    out_s[qi * k + i] = pad ? -FLT_MAX : (float)s[i];
    out_i[qi * k + i] = pad ? -1 : id[i];
  }
}

}  // namespace nrx

using namespace nrx;

extern "C" size_t nrx_topk_index_bytes(int64_t N, int D) {
  TopkGeom g;
  if (make_geom(1, N, D, 1, &g) != NRX_OK) return 0;
  return g.index_bytes;
}

extern "C" int nrx_topk_index_build(const float* corpus, int64_t c_ld, int64_t N, int D, void* index, size_t index_bytes,
                                    nrx_stream_t stream) {
  TopkGeom g;
  int rc = make_geom(1, N, D, 1, &g);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(index && index_bytes >= g.index_bytes, NRX_EWORKSPACE, "index buffer %zu < %zu", index_bytes, g.index_bytes);
  NRX_REQUIRE((corpus && c_ld >= D) || N == 0, NRX_EINVAL, "bad corpus");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(index, 0, kHdrBytes, st);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  if (N == 0) return NRX_OK;
  topk_pack_kernel<<<sm_count() * 4, 256, 0, st>>>(corpus, c_ld, N, D, g.Dp, (uint8_t*)index + kHdrBytes, (unsigned*)index);
  return check_launch("topk_pack");
}

extern "C" size_t nrx_topk_search_workspace_bytes(int64_t Q, int64_t N, int D, int k) {
  TopkGeom g;
  if (make_geom(Q, N, D, k, &g) != NRX_OK) return 0;
  return g.total;
}

extern "C" int nrx_topk_search64(const void* index, const float* corpus, int64_t c_ld, int64_t N, int D, const float* queries,
                                 int64_t q_ld, int64_t Q, int k, int64_t id_base, float* out_scores, double* out_scores64,
                                 int64_t* out_ids, int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  TopkGeom g;
  int rc = make_geom(Q, N, D, k, &g);
  if (rc != NRX_OK) return rc;
  if (Q == 0) return NRX_OK;
  NRX_REQUIRE(index && queries && out_scores && out_ids && q_ld >= D, NRX_EINVAL, "null / bad argument");
  NRX_REQUIRE((corpus && c_ld >= D) || N == 0, NRX_EINVAL, "bad corpus");
  NRX_REQUIRE(ws && ws_bytes >= g.total, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, g.total);
  NRX_REQUIRE(N < (1ll << 32) - 1, NRX_EUNSUPPORTED, "corpus too large for 32-bit row ids");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)ws;
  float* gmax = (float*)(w + g.gmax);
  float* theta = (float*)(w + g.theta);
  float* eps = (float*)(w + g.eps);
  unsigned* count = (unsigned*)(w + g.count);
  int* flag = (int*)(w + g.flag);
  unsigned* flist = (unsigned*)(w + g.flist);
  double* part_s = (double*)(w + g.fpart);
  unsigned* part_i = (unsigned*)(w + g.fpart + (size_t)g.fb_items * k * 8);
  unsigned* cand = (unsigned*)(w + g.cand);
  const uint8_t* img = (const uint8_t*)index + kHdrBytes;
  if (status != nullptr) {
    cudaError_t e = cudaMemsetAsync(status, 0, (size_t)Q * sizeof(int32_t), st);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  }
  const size_t fb_smem = (size_t)kFbCap * 12 + (size_t)D * 4;
  const bool fast = g.n_tiles >= 4;  // tiny corpora go straight to the exact kernel
  if (fast) {
    const size_t smem = (size_t)(g.nq + g.stages) * g.tile_bytes;
    cudaFuncSetAttribute(topk_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(topk_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NRX_REQUIRE(g.slices <= 1024, NRX_EUNSUPPORTED, "more than 1024 corpus slices");
    const long long n_groups = g.n_stiles * 4;
    uint8_t* qimg = w + g.qimg;
    {
      const long long nchunks = g.n_qtiles * (g.Dp / 8) * kTR;
      long long blocks = (nchunks + 255) / 256;
      if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
      topk_qpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(queries, q_ld, Q, D, g.Dp, qimg, g.n_qtiles);
      rc = check_launch("topk_qpack");
      if (rc != NRX_OK) return rc;
    }
    // sample: every stride-th tile -> group maxima -> theta
    long long s_slices = g.slices < g.n_stiles ? g.slices : g.n_stiles;
    topk_scan_kernel<0><<<dim3((unsigned)s_slices, (unsigned)g.n_qgroups), kScanThreads, smem, st>>>(
        img, N, g.n_tiles, g.stride, g.Dp, qimg, g.Qp, g.nq, g.stages, gmax, n_groups, nullptr, nullptr, nullptr, g.cap_s);
    rc = check_launch("topk_scan<sample>");
    if (rc != NRX_OK) return rc;
    const size_t tsm = n_groups <= kThetaStage ? (size_t)n_groups * 4 : 0;
    topk_theta_kernel<<<(unsigned)g.Qp, 256, tsm, st>>>(gmax, n_groups, g.Qp, Q, g.kprime_s, queries, q_ld, D, (const unsigned*)index, theta,
                                                       eps, flist, flag);
    rc = check_launch("topk_theta");
    if (rc != NRX_OK) return rc;
    topk_scan_kernel<1><<<dim3((unsigned)g.slices, (unsigned)g.n_qgroups), kScanThreads, smem, st>>>(
        img, N, g.n_tiles, 1, g.Dp, qimg, g.Qp, g.nq, g.stages, nullptr, 0, theta, count, cand, g.cap_s);
    rc = check_launch("topk_scan<filter>");
    if (rc != NRX_OK) return rc;
    const int n_regions = (int)(4 * g.slices);
    const size_t fsm = (size_t)kCap * 12 + (size_t)((D + 3) & ~3) * 4 + (size_t)n_regions * 4;
    unsigned* pre_id = g.split_final ? (unsigned*)(w + g.pre_id) : nullptr;
    double* pre_sc = g.split_final ? (double*)(w + g.pre_sc) : nullptr;
    if (g.split_final) {
      const size_t rsm = (size_t)((D + 3) & ~3) * 4 + (size_t)n_regions * 4;
      topk_rescore_kernel<<<dim3(kRescoreChunks, (unsigned)Q), 256, rsm, st>>>(corpus, c_ld, D, queries, q_ld, count, cand, n_regions, g.cap_s,
                                                                               pre_id, pre_sc);
      rc = check_launch("topk_rescore");
      if (rc != NRX_OK) return rc;
    }
    cudaFuncSetAttribute(topk_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    cudaFuncSetAttribute(topk_final_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    topk_final_kernel<<<(unsigned)Q, kFinalThreads, fsm, st>>>(corpus, c_ld, N, D, queries, q_ld, k, id_base, theta, eps, count, cand, n_regions,
                                                    g.cap_s, flag, flist, out_scores, out_scores64, (long long*)out_ids, pre_id, pre_sc);
    rc = check_launch("topk_final");
    if (rc != NRX_OK) return rc;
  }
  // exact fallback for the listed queries (every query when the corpus is tiny): fixed grid, exits at once on an empty list
  CorpusSegs CS;
  memset(&CS, 0, sizeof(CS));
  CS.n_seg = 1; CS.c[0] = corpus; CS.base[0] = 0; CS.base[1] = N;
  PeerOuts PO;
  memset(&PO, 0, sizeof(PO));
  cudaFuncSetAttribute(topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fb_smem);
  topk_exact_kernel<<<(unsigned)g.fb_grid, 256, fb_smem, st>>>(CS, c_ld, N, D, queries, q_ld, k, flist, fast ? 0 : (int)Q, g.fb_items,
                                                             part_s, part_i);
  rc = check_launch("topk_exact");
  if (rc != NRX_OK) return rc;
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < fb_max_slices(k) * k) n2 <<= 1;
  const size_t msm = (size_t)n2 * 12;
  cudaFuncSetAttribute(topk_exact_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
  topk_exact_merge_kernel<<<(unsigned)g.fb_grid, 256, msm, st>>>(N, k, id_base, flist, fast ? 0 : (int)Q, g.fb_grid, g.fb_items, part_s,
                                                               part_i, out_scores, out_scores64, (long long*)out_ids, status, PO);
  return check_launch("topk_exact_merge");
}

extern "C" int nrx_topk_search(const void* index, const float* corpus, int64_t c_ld, int64_t N, int D, const float* queries,
                               int64_t q_ld, int64_t Q, int k, int64_t id_base, float* out_scores, int64_t* out_ids,
                               int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  return nrx_topk_search64(index, corpus, c_ld, N, D, queries, q_ld, Q, k, id_base, out_scores, nullptr, out_ids, status, ws, ws_bytes,
                           stream);
}


// ---- sharded search over peer memory --------------------------------------------------------------------------------
extern "C" size_t nrx_topk_peer_inbox_bytes(int64_t Q, int world, int k) {
  if (Q < 0 || world < 1 || k < 1) return 0;
  const long long q_own = (Q + world - 1) / world;
  return (inbox_bytes(q_own > 0 ? q_own : 1, world, k) + 255) & ~(size_t)255;
}

extern "C" int nrx_topk_search_peer(const void* index, int64_t N_local, int D, const float* queries, int64_t q_ld, int64_t Q, int k,
                                    const NrxTopkPeer* h_peer, int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  NRX_REQUIRE(h_peer != nullptr, NRX_EINVAL, "null peer descriptor");
  const int G = h_peer->world, R = h_peer->rank;
  NRX_REQUIRE(G >= 1 && G <= NRX_MAX_PEERS && R >= 0 && R < G, NRX_EINVAL, "bad rank / world");
  NRX_REQUIRE((long long)G * k <= 8192, NRX_EUNSUPPORTED, "peer search supports world * k <= 8192");
  long long N_total = 0, id_base = 0;
  for (int j = 0; j < G; ++j) {
    NRX_REQUIRE(h_peer->corpus[j] && h_peer->inbox[j] && h_peer->out_scores[j] && h_peer->out_ids[j] && h_peer->sig[j], NRX_EINVAL,
                "rank %d: null peer buffer", j);
    if (j < R) id_base += h_peer->n_rows[j];
    N_total += h_peer->n_rows[j];
  }
  NRX_REQUIRE(h_peer->n_rows[R] == N_local, NRX_EINVAL, "n_rows[rank] != N_local");
  // per-shard threshold rank: the shards share the k' = 2k + 64 budget (each keeps a slack of its own)
  TopkGeom g;
  const int kp = h_peer->kprime > 0 ? (int)h_peer->kprime : (2 * k + 64 + G - 1) / G + 24;
  int rc = make_geom(Q, N_local, D, k, &g, kp);
  if (rc != NRX_OK) return rc;
  if (Q == 0) return NRX_OK;
  NRX_REQUIRE(index && queries && q_ld >= D, NRX_EINVAL, "null / bad argument");
  NRX_REQUIRE(ws && ws_bytes >= g.total, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, g.total);
  NRX_REQUIRE(N_total < (1ll << 32) - 1, NRX_EUNSUPPORTED, "corpus too large for 32-bit row ids");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = (uint8_t*)ws;
  float* gmax = (float*)(w + g.gmax);
  float* theta = (float*)(w + g.theta);
  float* eps = (float*)(w + g.eps);
  unsigned* count = (unsigned*)(w + g.count);
  int* flag = (int*)(w + g.flag);
  unsigned* flist = (unsigned*)(w + g.flist);
  double* part_s = (double*)(w + g.fpart);
  unsigned* part_i = (unsigned*)(w + g.fpart + (size_t)g.fb_items * k * 8);
  unsigned* cand = (unsigned*)(w + g.cand);
  uint8_t* qimg = w + g.qimg;
  const uint8_t* img = (const uint8_t*)index + kHdrBytes;
  const float* corpus = h_peer->corpus[R];
  if (status != nullptr) {
    cudaError_t e = cudaMemsetAsync(status, 0, (size_t)Q * sizeof(int32_t), st);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  }
  PeerInbox PB;
  memset(&PB, 0, sizeof(PB));
  PB.rank = R; PB.world = G; PB.q_own = (Q + G - 1) / G;
  PeerOuts PO;
  memset(&PO, 0, sizeof(PO));
  PO.n = G;
  for (int j = 0; j < G; ++j) { PB.box[j] = (uint8_t*)h_peer->inbox[j]; PO.s[j] = h_peer->out_scores[j]; PO.i[j] = (long long*)h_peer->out_ids[j]; }
  NrxPeerStep bar;
  memset(&bar, 0, sizeof(bar));
  bar.rank = R; bar.world = G; bar.status = h_peer->status; bar.timeout_ms = h_peer->timeout_ms;
  for (int j = 0; j < G; ++j) bar.sig[j] = h_peer->sig[j];

  const bool fast = g.n_tiles >= 4;
  const int n_regions = (int)(4 * g.slices);
  if (fast) {
    const size_t smem = (size_t)(g.nq + g.stages) * g.tile_bytes;
    cudaFuncSetAttribute(topk_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(topk_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NRX_REQUIRE(g.slices <= 1024, NRX_EUNSUPPORTED, "more than 1024 corpus slices");
    const long long n_groups = g.n_stiles * 4;
    const long long nchunks = g.n_qtiles * (g.Dp / 8) * kTR;
    long long blocks = (nchunks + 255) / 256;
    if (blocks > 4LL * sm_count()) blocks = 4LL * sm_count();
    topk_qpack_kernel<<<(unsigned)blocks, 256, 0, st>>>(queries, q_ld, Q, D, g.Dp, qimg, g.n_qtiles);
    long long s_slices = g.slices < g.n_stiles ? g.slices : g.n_stiles;
    topk_scan_kernel<0><<<dim3((unsigned)s_slices, (unsigned)g.n_qgroups), kScanThreads, smem, st>>>(
        img, N_local, g.n_tiles, g.stride, g.Dp, qimg, g.Qp, g.nq, g.stages, gmax, n_groups, nullptr, nullptr, nullptr, g.cap_s);
    const size_t tsm = n_groups <= kThetaStage ? (size_t)n_groups * 4 : 0;
    topk_theta_kernel<<<(unsigned)g.Qp, 256, tsm, st>>>(gmax, n_groups, g.Qp, Q, g.kprime_s, queries, q_ld, D, (const unsigned*)index, theta,
                                                       eps, flist, flag);
    topk_scan_kernel<1><<<dim3((unsigned)g.slices, (unsigned)g.n_qgroups), kScanThreads, smem, st>>>(
        img, N_local, g.n_tiles, 1, g.Dp, qimg, g.Qp, g.nq, g.stages, nullptr, 0, theta, count, cand, g.cap_s);
    rc = check_launch("topk_scan(peer)");
    if (rc != NRX_OK) return rc;
  } else {
    // tiny shard: no tensor-core filter — every row is a candidate region-less; tell the owner to scan exactly
    cudaError_t e = cudaMemsetAsync(count, 0xff, (size_t)g.Qp * g.slices * 4 * 4, st);   // counts > cap_s: "overflow" on purpose
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
    e = cudaMemsetAsync(flist, 0, 16, st);
    NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  }
  const size_t fsm = (size_t)kCap * 12 + (size_t)((D + 3) & ~3) * 4 + (size_t)n_regions * 4;
  cudaFuncSetAttribute(topk_final_peer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
  cudaFuncSetAttribute(topk_final_peer_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  unsigned* pre_id = (g.split_final && fast) ? (unsigned*)(w + g.pre_id) : nullptr;
  double* pre_sc = (g.split_final && fast) ? (double*)(w + g.pre_sc) : nullptr;
  if (pre_id != nullptr) {
    const size_t rsm = (size_t)((D + 3) & ~3) * 4 + (size_t)n_regions * 4;
    topk_rescore_kernel<<<dim3(kRescoreChunks, (unsigned)Q), 256, rsm, st>>>(corpus, D, D, queries, q_ld, count, cand, n_regions, g.cap_s,
                                                                             pre_id, pre_sc);
    rc = check_launch("topk_rescore(peer)");
    if (rc != NRX_OK) return rc;
  }
  topk_final_peer_kernel<<<(unsigned)Q, kFinalThreads, fsm, st>>>(corpus, D, N_local, D, queries, q_ld, k, id_base, theta, eps, count, cand, n_regions,
                                                        g.cap_s, PB, pre_id, pre_sc);
  rc = check_launch("topk_final_peer");
  if (rc != NRX_OK) return rc;
  rc = nrx_peer_barrier(&bar, stream);     // every shard's lists have landed in the owners' inboxes
  if (rc != NRX_OK) return rc;
  const size_t osm = ((size_t)G * k + k) * 16;   // the G lists + the merged list, (fp64 score, int64 id) each
  cudaFuncSetAttribute(topk_owner_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)osm);
  topk_owner_merge_kernel<<<(unsigned)PB.q_own, 256, osm, st>>>(Q, N_total, k, PB, flist, status, PO);
  rc = check_launch("topk_owner_merge");
  if (rc != NRX_OK) return rc;
  // listed queries (proof failed / a shard overflowed / tiny shards): exact scan of EVERY shard by the owner, through peer memory
  CorpusSegs CS;
  memset(&CS, 0, sizeof(CS));
  CS.n_seg = G;
  for (int j = 0; j < G; ++j) { CS.c[j] = h_peer->corpus[j]; CS.base[j + 1] = CS.base[j] + h_peer->n_rows[j]; }
  const size_t fb_smem = (size_t)kFbCap * 12 + (size_t)D * 4;
  cudaFuncSetAttribute(topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fb_smem);
  topk_exact_kernel<<<(unsigned)g.fb_grid, 256, fb_smem, st>>>(CS, D, N_total, D, queries, q_ld, k, flist, 0, g.fb_items, part_s, part_i);
  rc = check_launch("topk_exact(peer)");
  if (rc != NRX_OK) return rc;
  int m2 = 32;   // the sort network needs >= one warp of elements
  while (m2 < fb_max_slices(k) * k) m2 <<= 1;
  const size_t msm = (size_t)m2 * 12;
  cudaFuncSetAttribute(topk_exact_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
  topk_exact_merge_kernel<<<(unsigned)g.fb_grid, 256, msm, st>>>(N_total, k, 0, flist, 0, g.fb_grid, g.fb_items, part_s, part_i, nullptr,
                                                               nullptr, nullptr, nullptr, PO);
  rc = check_launch("topk_exact_merge(peer)");
  if (rc != NRX_OK) return rc;
  return nrx_peer_barrier(&bar, stream);   // every owner's results are in every rank's output buffers
}

extern "C" size_t nrx_topk_ip_workspace_bytes(int64_t Q, int64_t N, int D, int k) {
  TopkGeom g;
  if (make_geom(Q, N, D, k, &g) != NRX_OK) return 0;
  return ((g.index_bytes + 255) & ~(size_t)255) + g.total;
}

extern "C" int nrx_topk_ip(const float* queries, int64_t q_ld, const float* corpus, int64_t c_ld, int64_t Q, int64_t N, int D,
                           int k, int64_t id_base, float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes,
                           nrx_stream_t stream) {
  TopkGeom g;
  int rc = make_geom(Q, N, D, k, &g);
  if (rc != NRX_OK) return rc;
  const size_t ib = (g.index_bytes + 255) & ~(size_t)255;
  NRX_REQUIRE(ws && ws_bytes >= ib + g.total, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, ib + g.total);
  rc = nrx_topk_index_build(corpus, c_ld, N, D, ws, ib, stream);
  if (rc != NRX_OK) return rc;
  return nrx_topk_search(ws, corpus, c_ld, N, D, queries, q_ld, Q, k, id_base, out_scores, out_ids, nullptr, (uint8_t*)ws + ib,
                         ws_bytes - ib, stream);
}

template <typename ST>
static int merge_launch(const ST* scores, const int64_t* ids, int n_lists, int64_t Q, int k, float* out_scores, int64_t* out_ids,
                        nrx_stream_t stream) {
  NRX_REQUIRE(scores && ids && out_scores && out_ids && n_lists >= 1 && k >= 1, NRX_EINVAL, "bad merge arguments");
  NRX_REQUIRE((long long)n_lists * k <= 8192, NRX_EUNSUPPORTED, "merge supports n_lists*k <= 8192");
  if (Q == 0) return NRX_OK;
  int n2 = 32;   // the sort network needs >= one warp of elements
  while (n2 < n_lists * k) n2 <<= 1;
  const size_t smem = (size_t)n2 * (8 + sizeof(ST));
  cudaFuncSetAttribute(topk_merge_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  topk_merge_kernel<ST><<<(unsigned)Q, 256, smem, (cudaStream_t)stream>>>(scores, (const long long*)ids, n_lists, Q, k, out_scores,
                                                                          (long long*)out_ids);
  return check_launch("topk_merge");
}

extern "C" int nrx_topk_merge(const float* scores, const int64_t* ids, int n_lists, int64_t Q, int k, float* out_scores,
                              int64_t* out_ids, nrx_stream_t stream) {
  return merge_launch<float>(scores, ids, n_lists, Q, k, out_scores, out_ids, stream);
}

extern "C" int nrx_topk_merge64(const double* scores64, const int64_t* ids, int n_lists, int64_t Q, int k, float* out_scores,
                                int64_t* out_ids, nrx_stream_t stream) {
  return merge_launch<double>(scores64, ids, n_lists, Q, k, out_scores, out_ids, stream);
}
