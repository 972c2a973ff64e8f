// K6 — exact inner-product top-k over a full corpus (replaces faiss.IndexFlatIP.add/search as used at
// reference src/model/recall/DSSM/model.py:209,249-251 and src/model/model_utils/TopKSearcher.py:34-47,73-77).
//
// Result contract: for every query the k corpus rows with the largest inner product, ordered by
// (inner product desc, id asc), where the ordering key is the inner product of the fp32 inputs accumulated in
// fp64 (so near-ties that an fp32 BLAS would order arbitrarily are pinned); returned scores are that value
// rounded to fp32; ids are corpus positions + id_base; -1 / -FLT_MAX pad when k > N.
//
// Fast path (tensor cores) with a proof of completeness per query:
//   index   : corpus packed once to bf16 128-row tile images in the canonical UMMA layout (+ max row norm);
//   pass A  : S~ = Q C^T on tcgen05 (bf16 in, fp32 accumulate in TMEM), epilogue keeps only the max of each
//             128-row group -> gmax[tile][query];
//   theta   : per query the k'-th largest group max (k' > k).  k' distinct rows score >= theta, so theta is a
//             lower bound of the k'-th best approximate score;
//   pass B  : the same GEMM, epilogue appends rows with S~ >= theta to the query's candidate list (a few
//             hundred rows out of N);
//   final   : candidates re-scored exactly (fp64), sorted by (score desc, id asc).  Every row outside the list
//             has S~ < theta, hence true score < theta + eps (eps = bf16 rounding bound 2^-7.5 |q| max|c|).
//             If the k-th exact score is >= theta + eps the list provably contains the true top-k.  Otherwise
//             (or on list overflow) the query is flagged and served by
//   fallback: an exact fp64 scan of the whole corpus with a block-level streaming top-k.
// Tensor-bound for Q >= ~256 (2 N D flops per query per pass), HBM-bound (corpus stream) below.
#include <float.h>

#include "common.cuh"
#include "umma.cuh"

namespace nrx {
using namespace umma;

static constexpr int kTR = 128;        // corpus rows per tile == UMMA N; queries per tile == UMMA M
static constexpr int kCap = 2048;      // candidate list capacity per query
static constexpr int kScanThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
static constexpr int kHdrBytes = 256;

struct TopkGeom {
  long long N, Q, n_tiles, n_qtiles, Qp;
  int D, Dp, k, kprime, stages;
  long long slices;   // corpus slices per query tile (grid.x of the scan)
  int cap_s;          // candidate capacity per (query, slice): private region, no atomics
  size_t tile_bytes, index_bytes;
  // workspace offsets
  size_t gmax, theta, eps, count, cand, flag, total;
};

static int make_geom(long long Q, long long N, int D, int k, TopkGeom* g) {
  NRX_REQUIRE(N >= 0 && Q >= 0 && D >= 1 && k >= 1, NRX_EINVAL, "bad top-k sizes");
  NRX_REQUIRE(D <= 256, NRX_EUNSUPPORTED, "top-k supports D <= 256 (got %d)", D);
  NRX_REQUIRE(k <= 1024, NRX_EUNSUPPORTED, "top-k supports k <= 1024 (got %d)", k);
  memset(g, 0, sizeof(*g));
  g->N = N; g->Q = Q; g->D = D; g->k = k;
  g->Dp = (D + 15) & ~15;
  g->n_tiles = (N + kTR - 1) / kTR;
  g->n_qtiles = (Q + kTR - 1) / kTR;
  g->Qp = g->n_qtiles * kTR;
  g->kprime = 2 * k + 64;
  g->tile_bytes = (size_t)kTR * g->Dp * 2;
  g->index_bytes = kHdrBytes + (size_t)g->n_tiles * g->tile_bytes;
  g->stages = g->Dp <= 128 ? 4 : 2;
  g->slices = g->n_qtiles > 0 ? sm_count() / g->n_qtiles : 1;
  if (g->slices < 1) g->slices = 1;
  if (g->slices > g->n_tiles) g->slices = g->n_tiles > 0 ? g->n_tiles : 1;
  g->cap_s = (int)(kCap / (2 * g->slices));  // regions are per (query, slice, 64-column half)
  if (g->cap_s < 16) g->cap_s = 16;
  if (g->cap_s > 512) g->cap_s = 512;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o = 0;
  g->gmax = o;  o += al((size_t)g->n_tiles * 2 * g->Qp * 4);
  g->theta = o; o += al((size_t)g->Qp * 4);
  g->eps = o;   o += al((size_t)g->Qp * 4);
  g->count = o; o += al((size_t)g->Qp * g->slices * 2 * 4);
  g->flag = o;  o += al((size_t)g->Qp * 4);
  g->cand = o;  o += al((size_t)g->Qp * g->slices * 2 * g->cap_s * 4);
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

// ---- the scan (pass A: MODE 0 group maxima, pass B: MODE 1 candidate filter) -------------------------------
template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
topk_scan_kernel(const uint8_t* __restrict__ img, long long N, long long n_tiles, int Dp, const float* __restrict__ q,
                 long long qld, long long Q, int D, long long Qp, int stages, float* __restrict__ gmax,
                 const float* __restrict__ theta, unsigned* __restrict__ count, unsigned* __restrict__ cand, int cap_s) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const size_t tile_bytes = (size_t)kTR * Dp * 2;
  uint8_t* sQ = smem;
  uint8_t* sC = smem + tile_bytes;
  __shared__ uint64_t full[4], empty[4], tfull[2], tempty[2];
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q0 = (long long)blockIdx.y * kTR;
  // corpus slice of this CTA
  const long long per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = (long long)blockIdx.x * per;
  const long long t1 = min(t0 + per, n_tiles);

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 256); }
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_s, 256u);
  // query tile: fp32 -> bf16 canonical (rows >= Q are zero)
  for (int i = tid; i < (Dp / 8) * kTR; i += kScanThreads) {
    const int r = i % kTR, kc = i / kTR;
    const long long row = q0 + r;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int d = kc * 8 + j;
      f[j] = (row < Q && d < D) ? __ldg(q + row * qld + d) : 0.f;
    }
    *reinterpret_cast<uint4*>(sQ + canon_off(kTR, r, kc)) =
        make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;

  if (t1 > t0) {
    if (warp == 0) {
      if (lane == 0) {  // TMA producer
        uint32_t ph[4] = {0, 0, 0, 0};
        bool used[4] = {false, false, false, false};
        int s = 0;
        for (long long t = t0; t < t1; ++t) {
          if (used[s]) { mbar_wait(&empty[s], ph[s]); ph[s] ^= 1; }
          used[s] = true;
          mbar_expect_tx(&full[s], (uint32_t)tile_bytes);
          bulk_g2s(sC + (size_t)s * tile_bytes, img + (size_t)t * tile_bytes, (uint32_t)tile_bytes, &full[s]);
          s = (s + 1 == stages) ? 0 : s + 1;
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {  // MMA issuer
        const uint32_t idesc = make_idesc_bf16(kTR, kTR);
        uint32_t fph[4] = {0, 0, 0, 0}, eph[2] = {0, 0};
        bool aused[2] = {false, false};
        int s = 0, a = 0;
        for (long long t = t0; t < t1; ++t) {
          mbar_wait(&full[s], fph[s]); fph[s] ^= 1;
          if (aused[a]) { mbar_wait(&tempty[a], eph[a]); eph[a] ^= 1; }
          aused[a] = true;
          tc_fence_after();
          const uint32_t cb = smem_u32(sC + (size_t)s * tile_bytes);
          for (int k16 = 0; k16 < Dp / 16; ++k16) {
            const uint64_t ad = make_smem_desc(smem_u32(sQ) + (uint32_t)k16 * 2u * (kTR * 16u), kTR * 16u, 128u);
            const uint64_t bd = make_smem_desc(cb + (uint32_t)k16 * 2u * (kTR * 16u), kTR * 16u, 128u);
            mma_bf16_ss(tmem + (uint32_t)a * kTR, ad, bd, idesc, k16 > 0);
          }
          mma_commit(&empty[s]);
          mma_commit(&tfull[a]);
          s = (s + 1 == stages) ? 0 : s + 1;
          a ^= 1;
        }
      }
    } else {  // epilogue: 8 warps = 2 per TMEM lane quadrant; thread == (query row, 64-column half of the tile)
      const int qd = warp & 3;
      const int half = (warp - 2) >> 2;
      const int r = qd * 32 + lane;
      const long long qrow = q0 + r;
      uint32_t tph[2] = {0, 0};
      float th = 0.f;
      if (MODE == 1) th = __ldg(theta + qrow);  // theta is allocated for Qp rows
      // this thread is the only writer of its (query, slice, half) candidate region: plain stores, register counter
      unsigned mycnt = 0;
      const size_t region = ((size_t)qrow * gridDim.x + blockIdx.x) * 2 + half;
      unsigned* mycand = (MODE == 1) ? cand + region * cap_s : nullptr;
      int a = 0;
      for (long long t = t0; t < t1; ++t) {
        mbar_wait(&tfull[a], tph[a]); tph[a] ^= 1;
        tc_fence_after();
        const long long base = t * kTR + half * 64;
        const bool tail = base + 64 > N;
        float v[64];
        {  // both TMEM loads in flight before the single wait
          float (&v0)[32] = *reinterpret_cast<float(*)[32]>(&v[0]);
          float (&v1)[32] = *reinterpret_cast<float(*)[32]>(&v[32]);
          const uint32_t ta = tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(a * kTR + half * 64);
          tmem_ld32(ta, v0);
          tmem_ld32(ta + 32u, v1);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(&tempty[a]);  // values are in registers: hand the accumulator back before the scan
        if (tail) {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (base + j >= N) v[j] = __int_as_float(0xff800000);  // -inf: never admitted, even by theta = -FLT_MAX
        }
        float m4[4] = {v[0], v[1], v[2], v[3]};  // four independent max chains
#pragma unroll
        for (int j = 4; j < 64; ++j) m4[j & 3] = fmaxf(m4[j & 3], v[j]);
        const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        if (MODE == 0) {
          gmax[(qrow * n_tiles + t) * 2 + half] = m;   // groups of 64 corpus rows
        } else if (m >= th) {  // rare: a few hundred rows out of N pass the threshold
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            if (v[j] >= th) {
              if (mycnt < (unsigned)cap_s) mycand[mycnt] = (unsigned)(base + j);
              ++mycnt;
            }
          }
        }
        a ^= 1;
      }
      if (MODE == 1) count[region] = mycnt;
    }
  } else if (MODE == 1 && warp >= 2) {  // slice without tiles
    count[((size_t)(q0 + (warp & 3) * 32 + lane) * gridDim.x + blockIdx.x) * 2 + ((warp - 2) >> 2)] = 0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256u);
}

// ---- theta: k'-th largest group max per query (radix select on order-preserving uint keys) ------------------
__device__ __forceinline__ unsigned f2key(float f) { unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float key2f(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(256)
topk_theta_kernel(const float* __restrict__ gmax, long long n_tiles, long long Qp, long long Q, int kprime,
                  const float* __restrict__ q, long long qld, int D, const unsigned* __restrict__ max_norm_bits,
                  float* __restrict__ theta, float* __restrict__ eps, unsigned* __restrict__ count, int* __restrict__ flag) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_remaining;
  __shared__ float s_norm[8];
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid == 0) flag[qi] = 0;
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
  unsigned prefix = 0, mask = 0, remaining = (unsigned)kprime;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    for (long long t = tid; t < n_tiles; t += 256) {
      const unsigned key = f2key(__ldg(gmax + qi * n_tiles + t));
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned acc = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (acc + hist[b] >= remaining) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((unsigned)b << shift);
      s_remaining = remaining - acc;
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
__device__ __forceinline__ double dot64(const float* __restrict__ qs, const float* __restrict__ row, int D) {
  // four independent fp64 chains (shorter dependency chain), combined in a fixed order; every path that
  // scores a row uses this one function, so equal rows always get bit-equal scores
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int d = 0;
  for (; d + 4 <= D; d += 4) {
    s0 = fma((double)qs[d], (double)__ldg(row + d), s0);
    s1 = fma((double)qs[d + 1], (double)__ldg(row + d + 1), s1);
    s2 = fma((double)qs[d + 2], (double)__ldg(row + d + 2), s2);
    s3 = fma((double)qs[d + 3], (double)__ldg(row + d + 3), s3);
  }
  for (; d < D; ++d) s0 = fma((double)qs[d], (double)__ldg(row + d), s0);
  return (s0 + s1) + (s2 + s3);
}

// (score desc, id asc): returns true if a must come before b
__device__ __forceinline__ bool before(double sa, unsigned ia, double sb, unsigned ib) { return sa > sb || (sa == sb && ia < ib); }

__device__ void bitonic_sort(double* s, unsigned* id, int n, int tid, int nthreads) {  // n power of two
  for (int k2 = 2; k2 <= n; k2 <<= 1) {
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += nthreads) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = ((i & k2) == 0);
          const bool sw = up ? before(s[p], id[p], s[i], id[i]) : before(s[i], id[i], s[p], id[p]);
          if (sw) {
            const double ts = s[i]; s[i] = s[p]; s[p] = ts;
            const unsigned ti = id[i]; id[i] = id[p]; id[p] = ti;
          }
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256)
topk_final_kernel(const float* __restrict__ c, long long cld, long long N, int D, const float* __restrict__ q, long long qld,
                  int k, long long id_base, const float* __restrict__ theta, const float* __restrict__ eps,
                  const unsigned* __restrict__ count, const unsigned* __restrict__ cand, int n_slices, int cap_s,
                  int* __restrict__ flag, float* __restrict__ out_s, long long* __restrict__ out_i) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* s = reinterpret_cast<double*>(sm_raw);              // [kCap]
  unsigned* id = reinterpret_cast<unsigned*>(s + kCap);       // [kCap]
  float* qs = reinterpret_cast<float*>(id + kCap);            // [D]
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  __shared__ unsigned s_off[320];
  __shared__ unsigned s_total, s_over;
  const long long kk = k < N ? k : N;
  if (tid == 0) {  // exclusive scan of the per-slice counts (n_slices <= SM count)
    unsigned tot = 0, over = 0;
    for (int sl = 0; sl < n_slices; ++sl) {
      const unsigned cs = count[qi * n_slices + sl];
      if (cs > (unsigned)cap_s) over = 1;
      s_off[sl] = tot;
      tot += cs < (unsigned)cap_s ? cs : (unsigned)cap_s;
    }
    s_total = tot;
    s_over = over;
  }
  for (int d = tid; d < D; d += 256) qs[d] = __ldg(q + qi * qld + d);
  __syncthreads();
  const unsigned cnt = s_total;
  if (s_over || cnt > (unsigned)kCap || (long long)cnt < kk) {
    if (tid == 0) flag[qi] = 1;
    return;
  }
  int n2 = 1;
  while (n2 < (int)cnt) n2 <<= 1;
  for (int i = (int)cnt + tid; i < n2; i += 256) { id[i] = 0xffffffffu; s[i] = -DBL_MAX; }
  for (unsigned i = tid; i < cnt; i += 256) {  // all candidates in flight together: slice by binary search
    int lo = 0, hi = n_slices - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    const unsigned row = cand[((size_t)qi * n_slices + lo) * cap_s + (i - s_off[lo])];
    id[i] = row;
    s[i] = dot64(qs, c + (long long)row * cld, D);
  }
  __syncthreads();
  bitonic_sort(s, id, n2, tid, 256);
  // completeness proof: every row outside the list scores < theta + eps
  const bool ok = (theta[qi] == -FLT_MAX) || (kk == 0) || (s[kk - 1] >= (double)theta[qi] + (double)eps[qi]);
  if (!ok) {
    if (tid == 0) flag[qi] = 1;
    return;
  }
  for (int i = tid; i < k; i += 256) {
    if (i < kk) { out_s[qi * k + i] = (float)s[i]; out_i[qi * k + i] = (long long)id[i] + id_base; }
    else { out_s[qi * k + i] = -FLT_MAX; out_i[qi * k + i] = -1; }
  }
}

// Exact fallback: fp64 scan of the whole corpus, block-level streaming top-k (buffer + periodic bitonic prune).
static constexpr int kFbCap = 2048;
__global__ void __launch_bounds__(256)
topk_exact_kernel(const float* __restrict__ c, long long cld, long long N, int D, const float* __restrict__ q, long long qld,
                  int k, long long id_base, const int* __restrict__ flag, int force, float* __restrict__ out_s,
                  long long* __restrict__ out_i, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  double* s = reinterpret_cast<double*>(sm_raw);
  unsigned* id = reinterpret_cast<unsigned*>(s + kFbCap);
  float* qs = reinterpret_cast<float*>(id + kFbCap);
  __shared__ unsigned s_cnt;
  __shared__ double s_th;
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  const bool mine = force || flag[qi] != 0;
  if (status != nullptr && tid == 0) status[qi] = mine && !force ? 1 : 0;
  if (!mine) return;
  for (int d = tid; d < D; d += 256) qs[d] = __ldg(q + qi * qld + d);
  if (tid == 0) { s_cnt = 0; s_th = -DBL_MAX; }
  __syncthreads();
  const long long kk = k < N ? k : N;
  for (long long r0 = 0; r0 < N; r0 += 256) {
    const long long row = r0 + tid;
    if (row < N) {
      const double v = dot64(qs, c + row * cld, D);
      if (v >= s_th) {
        const unsigned slot = atomicAdd(&s_cnt, 1u);
        s[slot] = v;           // slot < kFbCap: pruned whenever fewer than 256 free slots remain
        id[slot] = (unsigned)row;
      }
    }
    __syncthreads();
    if (s_cnt > (unsigned)(kFbCap - 256) || r0 + 256 >= N) {
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
  for (int i = tid; i < k; i += 256) {
    if (i < kk) { out_s[qi * k + i] = (float)s[i]; out_i[qi * k + i] = (long long)id[i] + id_base; }
    else { out_s[qi * k + i] = -FLT_MAX; out_i[qi * k + i] = -1; }
  }
}

// ---- merge of per-shard lists ([n_lists][Q][k]) ----------------------------------------------------------------
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ sc, const long long* __restrict__ ids, int n_lists, long long Q, int k,
                  float* __restrict__ out_s, long long* __restrict__ out_i) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int total = n_lists * k;
  int n2 = 1;
  while (n2 < total) n2 <<= 1;
  float* s = reinterpret_cast<float*>(sm_raw);
  long long* id = reinterpret_cast<long long*>(s + n2 + (n2 & 1));
  const long long qi = blockIdx.x;
  const int tid = threadIdx.x;
  for (int i = tid; i < n2; i += 256) {
    if (i < total) {
      const int l = i / k, j = i % k;
      s[i] = sc[((long long)l * Q + qi) * k + j];
      id[i] = ids[((long long)l * Q + qi) * k + j];
      if (id[i] < 0) { s[i] = -FLT_MAX; id[i] = 0x7fffffffffffffffll; }
    } else { s[i] = -FLT_MAX; id[i] = 0x7fffffffffffffffll; }
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
            const float ts = s[i]; s[i] = s[p]; s[p] = ts;
            const long long ti = id[i]; id[i] = id[p]; id[p] = ti;
          }
        }
      }
      __syncthreads();
    }
  for (int i = tid; i < k; i += 256) {
    const bool pad = i >= total || id[i] == 0x7fffffffffffffffll;
    out_s[qi * k + i] = pad ? -FLT_MAX : s[i];
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

extern "C" int nrx_topk_search(const void* index, const float* corpus, int64_t c_ld, int64_t N, int D, const float* queries,
                               int64_t q_ld, int64_t Q, int k, int64_t id_base, float* out_scores, int64_t* out_ids,
                               int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream) {
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
  unsigned* cand = (unsigned*)(w + g.cand);
  const uint8_t* img = (const uint8_t*)index + kHdrBytes;
  const size_t fb_smem = (size_t)kFbCap * 12 + (size_t)D * 4;
  const bool fast = g.n_tiles >= 4;  // tiny corpora go straight to the exact kernel
  if (fast) {
    const size_t smem = (size_t)(1 + g.stages) * g.tile_bytes;
    cudaFuncSetAttribute(topk_scan_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(topk_scan_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    NRX_REQUIRE(g.slices <= 160, NRX_EUNSUPPORTED, "more than 160 corpus slices");
    dim3 grid((unsigned)g.slices, (unsigned)g.n_qtiles);
    topk_scan_kernel<0><<<grid, kScanThreads, smem, st>>>(img, N, g.n_tiles, g.Dp, queries, q_ld, Q, D, g.Qp, g.stages, gmax,
                                                          nullptr, nullptr, nullptr, g.cap_s);
    rc = check_launch("topk_scan<A>");
    if (rc != NRX_OK) return rc;
    topk_theta_kernel<<<(unsigned)g.Qp, 256, 0, st>>>(gmax, 2 * g.n_tiles /* groups of 64 rows */, g.Qp, Q, g.kprime, queries, q_ld, D, (const unsigned*)index,
                                                     theta, eps, count, flag);
    rc = check_launch("topk_theta");
    if (rc != NRX_OK) return rc;
    topk_scan_kernel<1><<<grid, kScanThreads, smem, st>>>(img, N, g.n_tiles, g.Dp, queries, q_ld, Q, D, g.Qp, g.stages, nullptr,
                                                          theta, count, cand, g.cap_s);
    rc = check_launch("topk_scan<B>");
    if (rc != NRX_OK) return rc;
    const size_t fsm = (size_t)kCap * 12 + (size_t)D * 4;
    cudaFuncSetAttribute(topk_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm);
    topk_final_kernel<<<(unsigned)Q, 256, fsm, st>>>(corpus, c_ld, N, D, queries, q_ld, k, id_base, theta, eps, count, cand,
                                                    (int)(2 * g.slices), g.cap_s, flag,
                                                    out_scores, (long long*)out_ids);
    rc = check_launch("topk_final");
    if (rc != NRX_OK) return rc;
  }
  cudaFuncSetAttribute(topk_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fb_smem);
  topk_exact_kernel<<<(unsigned)Q, 256, fb_smem, st>>>(corpus, c_ld, N, D, queries, q_ld, k, id_base, flag, fast ? 0 : 1, out_scores,
                                                      (long long*)out_ids, status);
  return check_launch("topk_exact");
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

extern "C" int nrx_topk_merge(const float* scores, const int64_t* ids, int n_lists, int64_t Q, int k, float* out_scores,
                              int64_t* out_ids, nrx_stream_t stream) {
  NRX_REQUIRE(scores && ids && out_scores && out_ids && n_lists >= 1 && k >= 1, NRX_EINVAL, "bad merge arguments");
  NRX_REQUIRE((long long)n_lists * k <= 8192, NRX_EUNSUPPORTED, "merge supports n_lists*k <= 8192");
  if (Q == 0) return NRX_OK;
  int n2 = 1;
  while (n2 < n_lists * k) n2 <<= 1;
  const size_t smem = (size_t)(n2 + 1) * 4 + (size_t)n2 * 8 + 16;
  cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  topk_merge_kernel<<<(unsigned)Q, 256, smem, (cudaStream_t)stream>>>(scores, (const long long*)ids, n_lists, Q, k, out_scores,
                                                                      (long long*)out_ids);
  return check_launch("topk_merge");
}
