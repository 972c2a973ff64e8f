// K3 — deterministic sorted-index segment-reduce backward of the embedding gather/pool,
// with optional fused sparse row update (SGD / AdamW).
//
// Replaces aten::embedding_dense_backward behind nn.Embedding(..., padding_idx=0)
// (reference src/model/BaseModel/base_model.py:164,271) and the autograd of the masked
// mean pooling (base_model.py:278-282).
//
//   plan : one key per (feature, sample, position) occurrence, key = table_id << row_bits | row,
//          invalid occurrences (row 0 = padding_idx, mask == 0, out-of-range id) get a sentinel
//          key that sorts last; stable LSD radix sort (cub::DeviceRadixSort) of (key, occurrence).
//   apply: phase A, one warp per tile of sorted occurrences: lanes own gradient columns, the
//          warp walks its tile in sorted order and sums grad_out rows per run of equal keys.
//          Runs inside a tile are final; runs crossing tile borders leave a partial per tile.
//          phase B, one warp per run that crosses a tile border: adds the partials in tile order.
//          Every unique row is finalised by exactly one warp in a fixed order => bitwise
//          reproducible, no float atomics.
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include "common.cuh"

namespace nrx {

static constexpr int kTile = 32;  // sorted occurrences per warp in phase A (lane = occurrence)

struct DTables {
  float* g[NRX_MAX_TABLES];   // dense grads
  float* w[NRX_MAX_TABLES];   // tables (row update modes)
  float* m[NRX_MAX_TABLES];
  float* v[NRX_MAX_TABLES];
  int dim[NRX_MAX_TABLES];
  int stride[NRX_MAX_TABLES];
  int mode;
  int row_bits;
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt;
  const float* d_hp;  // optional device [lr, bc1, bc2_sqrt]: overrides the host scalars (CUDA-graph replay)
};

struct PlanLayout {
  size_t keys_in, vals_in, keys_out, vals_out, head, tail, cub, total;
  size_t cub_bytes;
  long long n_tiles;
  int row_bits, key_bits, nc, dp;  // dp: floats per tile partial (max dim rounded up to 4)
};

static int plan_layout(const DFeats& d, PlanLayout* L) {
  long long max_rows = 1;
  for (int i = 0; i < d.n; ++i) max_rows = d.f[i].rows > max_rows ? d.f[i].rows : max_rows;
  int row_bits = 1;
  while ((1ll << row_bits) < max_rows) ++row_bits;
  int tbits = 0;
  while ((1 << tbits) < d.n_tables) ++tbits;
  L->row_bits = row_bits;
  L->key_bits = row_bits + tbits;
  NRX_REQUIRE(L->key_bits + 1 <= 32, NRX_EUNSUPPORTED, "table_bits+row_bits=%d does not fit 31-bit keys", L->key_bits);
  NRX_REQUIRE(d.n_occ < (1ll << 31), NRX_EUNSUPPORTED, "too many occurrences (%lld)", d.n_occ);
  NRX_REQUIRE(d.max_dim <= 128, NRX_EUNSUPPORTED, "backward supports dim <= 128 (got %d)", d.max_dim);
  L->nc = (d.max_dim + 31) / 32;
  if (L->nc == 3) L->nc = 4;
  const size_t n = (size_t)d.n_occ;
  L->n_tiles = (d.n_occ + kTile - 1) / kTile;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  L->keys_in = off;  off += al(n * 4);
  L->vals_in = off;  off += al(n * 4);
  L->keys_out = off; off += al(n * 4);
  L->vals_out = off; off += al(n * 4);
  L->dp = (d.max_dim + 3) & ~3;
  const size_t part = al((size_t)L->n_tiles * L->dp * 4);
  L->head = off; off += part;
  L->tail = off; off += part;
  size_t cub_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(n ? n : 1), 0, L->key_bits + 1);
  L->cub_bytes = cub_bytes;
  L->cub = off; off += al(cub_bytes);
  L->total = off;
  return NRX_OK;
}

__global__ void __launch_bounds__(256)
build_keys_kernel(const __grid_constant__ DFeats P, int row_bits, uint32_t sentinel,
                  uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n_occ) return;
  int f = 0;
  for (int i = 1; i < P.n; ++i)
    if (p >= P.f[i].occ_off) f = i;
  const DFeat& F = P.f[f];
  const long long r = p - F.occ_off;  // == b * L + l
  const long long id = load_idx(F.idx, r, F.idx32);
  bool valid = id > 0 && id < F.rows;  // row 0 is padding_idx: never receives gradient
  if (valid && F.pool == NRX_POOL_MASKED_MEAN) valid = __ldg(F.mask + r) != 0.f;
  keys[p] = valid ? (((uint32_t)F.table_id << row_bits) | (uint32_t)id) : sentinel;
  vals[p] = (uint32_t)p;
}

// ---- small-batch plan: chunk sort + rank merge, two launches spread over the whole GPU ------------------------
// Segment s holds every occurrence of table s (features in descriptor order); it occupies positions
// [seg_off[s], seg_off[s+1]) of the sorted arrays.  Inside a segment valid keys come first (ascending row,
// stable), then the segment's sentinels; equal keys stay contiguous, which is all the apply kernels need.
//   1. plan_chunk_sort_kernel: the segment is cut into chunks of 1024 occurrences; one 256-thread CTA builds the
//      keys of a chunk (row id, or 1 << bits for an invalid occurrence) and sorts them stably on exactly the bits
//      this table needs (cub::BlockRadixSort in shared memory).
//   2. plan_merge_kernel: one CTA per chunk again; the chunk-sorted keys of the WHOLE segment are staged in shared
//      memory and every element's final position is its rank: position inside its own chunk + #(keys <= it) in
//      earlier chunks + #(keys < it) in later chunks (two binary searches' worth per chunk) — a stable multiway
//      merge without any inter-CTA dependency.
// (A single-CTA-per-table block sort measured 67.6 us at B = 16384: one SM busy, 147 idle, on the step's
//  critical path — profiles/r1_ncu_launches_deepfm_final.csv.)
static constexpr int kChunkThreads = 256;
static constexpr int kChunkItems = 4;
static constexpr int kChunk = kChunkThreads * kChunkItems;  // 1024 occurrences
static constexpr int kMaxChunks = 40;                        // per table: 40960 occurrences, 160 KB of keys in SMEM (56 chunks were tried for cfg1's 52224: the rank merge grows with the chunk count, 60 us vs 68 us for the device radix sort, and the step got slower)
using ChunkSort = cub::BlockRadixSort<uint32_t, kChunkThreads, kChunkItems, uint32_t>;

struct SmallPlan {
  int n_seg;
  int total_chunks;
  int seg_table[NRX_MAX_TABLES];
  int seg_bits[NRX_MAX_TABLES];       // bits of (rows - 1): valid ids are < 1 << bits, the invalid marker is 1 << bits
  int seg_off[NRX_MAX_TABLES + 1];    // occurrence offsets
  int chunk_off[NRX_MAX_TABLES + 1];  // chunk offsets (blockIdx.x space)
};

static bool make_small_plan(const DFeats& d, SmallPlan* sp) {
  memset(sp, 0, sizeof(*sp));
  long long off = 0;
  int chunks = 0;
  for (int t = 0; t < d.n_tables; ++t) {
    long long cnt = 0, rows = 1;
    for (int i = 0; i < d.n; ++i)
      if (d.f[i].table_id == t) {
        cnt += (i + 1 < d.n ? d.f[i + 1].occ_off : d.n_occ) - d.f[i].occ_off;
        rows = d.f[i].rows;
      }
    if (cnt == 0) continue;
    if (cnt > (long long)kChunk * kMaxChunks) return false;
    int bits = 1;
    while ((1ll << bits) < rows) ++bits;
    if (bits > 30) return false;
    sp->seg_table[sp->n_seg] = t;
    sp->seg_bits[sp->n_seg] = bits;
    sp->seg_off[sp->n_seg] = (int)off;
    sp->chunk_off[sp->n_seg] = chunks;
    off += cnt;
    chunks += (int)((cnt + kChunk - 1) / kChunk);
    ++sp->n_seg;
  }
  sp->seg_off[sp->n_seg] = (int)off;
  sp->chunk_off[sp->n_seg] = chunks;
  sp->total_chunks = chunks;
  return sp->n_seg > 0 && off == d.n_occ;
}

__device__ __forceinline__ int chunk_segment(const SmallPlan& S, int b) {
  int s = 0;
  for (int i = 1; i < S.n_seg; ++i)
    if (b >= S.chunk_off[i]) s = i;
  return s;
}

__global__ void __launch_bounds__(kChunkThreads)
plan_chunk_sort_kernel(const __grid_constant__ DFeats P, const __grid_constant__ SmallPlan S,
                       uint32_t* __restrict__ keys_mid, uint32_t* __restrict__ vals_mid) {
  __shared__ typename ChunkSort::TempStorage temp;
  const int seg = chunk_segment(S, blockIdx.x);
  const int t = S.seg_table[seg];
  const int n = S.seg_off[seg + 1] - S.seg_off[seg];
  const int base = (blockIdx.x - S.chunk_off[seg]) * kChunk;
  const int bits = S.seg_bits[seg];
  uint32_t k[kChunkItems], v[kChunkItems];
#pragma unroll
  for (int j = 0; j < kChunkItems; ++j) {
    const int e = base + threadIdx.x * kChunkItems + j;  // blocked arrangement: position inside the segment
    k[j] = 0xffffffffu;                                  // beyond the segment: sorts last, never written
    v[j] = 0;
    if (e < n) {
      // e-th occurrence of table t: walk the features of this table in descriptor order
      int rem = e;
      for (int i = 0; i < P.n; ++i) {
        if (P.f[i].table_id != t) continue;
        const long long cnt = (i + 1 < P.n ? P.f[i + 1].occ_off : P.n_occ) - P.f[i].occ_off;
        if (rem < cnt) {
          const DFeat& F = P.f[i];
          const long long id = load_idx(F.idx, rem, F.idx32);
          bool valid = id > 0 && id < F.rows;
          if (valid && F.pool == NRX_POOL_MASKED_MEAN) valid = __ldg(F.mask + rem) != 0.f;
          k[j] = valid ? (uint32_t)id : (1u << bits);
          v[j] = (uint32_t)(F.occ_off + rem);
          break;
        }
        rem -= (int)cnt;
      }
    }
  }
  ChunkSort(temp).Sort(k, v, 0, bits + 1);
  uint32_t* ko = keys_mid + S.seg_off[seg] + base;
  uint32_t* vo = vals_mid + S.seg_off[seg] + base;
  const int len = min(kChunk, n - base);
  if (threadIdx.x * kChunkItems + kChunkItems <= len) {  // whole 16-byte group inside the chunk
    if ((((uintptr_t)ko) & 15) == 0) {
      reinterpret_cast<uint4*>(ko)[threadIdx.x] = make_uint4(k[0], k[1], k[2], k[3]);
      reinterpret_cast<uint4*>(vo)[threadIdx.x] = make_uint4(v[0], v[1], v[2], v[3]);
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < kChunkItems; ++j) {
    const int e = threadIdx.x * kChunkItems + j;
    if (e < len) { ko[e] = k[j]; vo[e] = v[j]; }
  }
}

// #(x in a[0, len) with x < key) and #(x <= key), branch-free binary searches over a sorted SMEM run (len <= 1024)
__device__ __forceinline__ int count_less(const uint32_t* a, int len, uint32_t key) {
  int lo = 0;
#pragma unroll
  for (int step = kChunk; step >= 1; step >>= 1) {
    const int mid = lo + step;
    if (mid <= len && a[mid - 1] < key) lo = mid;
  }
  return lo;
}
__device__ __forceinline__ int count_less_equal(const uint32_t* a, int len, uint32_t key) {
  int lo = 0;
#pragma unroll
  for (int step = kChunk; step >= 1; step >>= 1) {
    const int mid = lo + step;
    if (mid <= len && a[mid - 1] <= key) lo = mid;
  }
  return lo;
}

// grid = chunks x kChunkItems: CTA (chunk c, quarter q) ranks 256 elements of chunk c (one per thread), so a CTA lives
// for one SMEM fill + (n_chunks - 1) binary searches and the whole merge is a single wave.
__global__ void __launch_bounds__(kChunkThreads)
plan_merge_kernel(const __grid_constant__ SmallPlan S, int row_bits, uint32_t sentinel,
                  const uint32_t* __restrict__ keys_mid, const uint32_t* __restrict__ vals_mid,
                  uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  extern __shared__ __align__(16) uint32_t sk[];  // the segment's chunk-sorted keys
  const int chunk = blockIdx.x / kChunkItems, q = blockIdx.x % kChunkItems;
  const int seg = chunk_segment(S, chunk);
  const int t = S.seg_table[seg];
  const int n = S.seg_off[seg + 1] - S.seg_off[seg];
  const int c = chunk - S.chunk_off[seg];
  const int n_chunks = S.chunk_off[seg + 1] - S.chunk_off[seg];
  const int bits = S.seg_bits[seg];
  const int base = c * kChunk;
  const int len = min(kChunk, n - base);
  if (q * kChunkThreads >= len) return;          // ragged last chunk: nothing in this quarter
  const uint32_t* km = keys_mid + S.seg_off[seg];
  if ((((uintptr_t)km) & 15) == 0) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += kChunkThreads)
      reinterpret_cast<uint4*>(sk)[i] = __ldg(reinterpret_cast<const uint4*>(km) + i);
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += kChunkThreads) sk[i] = __ldg(km + i);
  } else {
    for (int i = threadIdx.x; i < n; i += kChunkThreads) sk[i] = __ldg(km + i);
  }
  __syncthreads();
  const int j = q * kChunkThreads + threadIdx.x;
  if (j >= len) return;
  const uint32_t key = sk[base + j];
  const uint32_t val = __ldg(vals_mid + S.seg_off[seg] + base + j);
  int pos = j;
  for (int c2 = 0; c2 < c; ++c2) pos += count_less_equal(sk + c2 * kChunk, kChunk, key);   // earlier chunks are full
  for (int c2 = c + 1; c2 < n_chunks; ++c2) pos += count_less(sk + c2 * kChunk, min(kChunk, n - c2 * kChunk), key);
  keys_out[S.seg_off[seg] + pos] = (key >> bits) ? sentinel : (((uint32_t)t << row_bits) | key);
  vals_out[S.seg_off[seg] + pos] = val;
}

// ---- finalise one unique row (lanes own columns c = lane + 32*k) -------------------------
template <int NC>
__device__ __forceinline__ void finalize_row(const DTables& T, uint32_t key, const float (&acc)[NC], int lane) {
  const int t = (int)(key >> T.row_bits);
  const long long row = (long long)(key & ((1u << T.row_bits) - 1u));
  const int dim = T.dim[t];
  const long long base = row * T.stride[t];
  const float lr = T.d_hp ? __ldg(T.d_hp) : T.lr;
  const float bc1 = T.d_hp ? __ldg(T.d_hp + 1) : T.bc1;
  const float bc2s = T.d_hp ? __ldg(T.d_hp + 2) : T.bc2_sqrt;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const int c = lane + 32 * k;
    if (c >= dim) continue;
    const float g = acc[k];
    if (T.mode == NRX_BWD_DENSE) {
      T.g[t][base + c] = g;
    } else if (T.mode == NRX_BWD_SGD) {
      float p = T.w[t][base + c];
      T.w[t][base + c] = p - lr * (g + T.wd * p);
    } else {  // AdamW on the touched row (torch.optim.AdamW update rule)
      float p = T.w[t][base + c];
      float m = T.m[t][base + c], v = T.v[t][base + c];
      adamw_update(p, g, m, v, lr, bc1, bc2s, T.beta1, T.beta2, T.eps, T.wd);
      T.w[t][base + c] = p;
      T.m[t][base + c] = m;
      T.v[t][base + c] = v;
    }
  }
}

// Lane-per-run finalisation of CW consecutive columns starting at c0 (values in registers).
template <int CW>
__device__ __forceinline__ void finalize_cols(const DTables& T, uint32_t key, const float (&g)[CW], int c0) {
  const int t = (int)(key >> T.row_bits);
  const long long row = (long long)(key & ((1u << T.row_bits) - 1u));
  const int dim = T.dim[t];
  if (c0 >= dim) return;
  const long long base = row * T.stride[t] + c0;
  const float lr = T.d_hp ? __ldg(T.d_hp) : T.lr;
  const float bc1 = T.d_hp ? __ldg(T.d_hp + 1) : T.bc1;
  const float bc2s = T.d_hp ? __ldg(T.d_hp + 2) : T.bc2_sqrt;
  if (T.mode == NRX_BWD_DENSE) {
    float* dst = T.g[t] + base;
#pragma unroll
    for (int j = 0; j < CW; ++j)
      if (c0 + j < dim) dst[j] = g[j];
    return;
  }
  float* pw = T.w[t] + base;
  if (T.mode == NRX_BWD_SGD) {
    float p[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) p[j] = (c0 + j < dim) ? pw[j] : 0.f;
#pragma unroll
    for (int j = 0; j < CW; ++j)
      if (c0 + j < dim) pw[j] = p[j] - lr * (g[j] + T.wd * p[j]);
    return;
  }
  float* pm = T.m[t] + base;
  float* pv = T.v[t] + base;
  float p[CW], m[CW], v[CW];
#pragma unroll
  for (int j = 0; j < CW; ++j) {  // all loads first: one round trip per run instead of one per column
    const bool ok = c0 + j < dim;
    p[j] = ok ? pw[j] : 0.f;
    m[j] = ok ? pm[j] : 0.f;
    v[j] = ok ? pv[j] : 0.f;
  }
#pragma unroll
  for (int j = 0; j < CW; ++j) {
    if (c0 + j < dim) {  // torch.optim.AdamW update rule on the touched row
      float pj = p[j] * (1.f - lr * T.wd);
      const float mj = T.beta1 * m[j] + (1.f - T.beta1) * g[j];
      const float vj = T.beta2 * v[j] + (1.f - T.beta2) * g[j] * g[j];
      pj -= (lr / bc1) * (mj / (sqrtf(vj) / bc2s + T.eps));
      pw[j] = pj; pm[j] = mj; pv[j] = vj;
    }
  }
}

// Phase A: one warp per tile of 32 sorted occurrences, LANE = OCCURRENCE.  Each lane loads its occurrence's
// gradient row (CW columns per pass) into registers, a segmented inclusive scan over the lanes (fixed
// shuffle tree => reproducible) leaves each run's total in the run's last lane, and all run-ends of the
// tile finalise their rows concurrently.  Runs touching a tile border leave a head / tail partial.
template <int CW>
__global__ void __launch_bounds__(128)
segment_reduce_kernel(const __grid_constant__ DFeats P, const __grid_constant__ DTables T,
                      const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                      const float* __restrict__ grad, long long gld, long long n_occ, uint32_t sentinel,
                      float* __restrict__ head, float* __restrict__ tail, int DP, int max_dim) {
  const long long tile = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long ts = tile * kTile;
  if (ts >= n_occ) return;
  const long long te = min(ts + (long long)kTile, n_occ);
  const int n = (int)(te - ts);
  const bool valid = lane < n;
  const uint32_t my_key = valid ? __ldg(keys + ts + lane) : 0xffffffffu;
  const uint32_t prev_key = ts > 0 ? __ldg(keys + ts - 1) : 0xffffffffu;
  const uint32_t next_key = te < n_occ ? __ldg(keys + te) : 0xfffffffeu;
  uint32_t k_up = __shfl_up_sync(NRX_FULL_MASK, my_key, 1);
  uint32_t k_dn = __shfl_down_sync(NRX_FULL_MASK, my_key, 1);
  const bool is_head = (lane == 0) || (my_key != k_up);
  const bool is_end = valid && (lane == n - 1 || k_dn != my_key);
  const unsigned heads = __ballot_sync(NRX_FULL_MASK, is_head);
  const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));  // first lane of my run
  const uint32_t k0 = __shfl_sync(NRX_FULL_MASK, my_key, 0);
  const bool from_prev = (start == 0) && (k0 == prev_key);
  const bool to_next = is_end && (lane == n - 1) && (my_key == next_key);
  const bool live = valid && my_key != sentinel;

  long long off = 0;
  float scale = 0.f;
  int dim = 0;
  bool vec4 = false;
  if (live) {
    const long long p = __ldg(vals + ts + lane);
    int f = 0;
    for (int q = 1; q < P.n; ++q)
      if (p >= P.f[q].occ_off) f = q;
    const DFeat& F = P.f[f];
    const long long r = p - F.occ_off;
    const long long b = (F.L == 1) ? r : r / F.L;
    off = b * gld + F.out_col;
    dim = F.dim;
    vec4 = ((gld | F.out_col) & 3) == 0 && (reinterpret_cast<uintptr_t>(grad) & 15) == 0;
    if (F.pool == NRX_POOL_NONE) scale = 1.f;
    else if (F.pool == NRX_POOL_MEAN) scale = 1.f / (float)F.L;
    else scale = __ldg(F.mask + r) * __ldg(F.inv_den + b);
  }
  for (int c0 = 0; c0 < max_dim; c0 += CW) {
    float v[CW];
    if (vec4 && live && c0 + CW <= dim) {   // whole 16-byte quads of this occurrence's gradient row
#pragma unroll
      for (int j = 0; j < CW; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(grad + off + c0 + j));
        v[j] = scale * g4.x; v[j + 1] = scale * g4.y; v[j + 2] = scale * g4.z; v[j + 3] = scale * g4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) v[j] = (live && c0 + j < dim) ? scale * __ldg(grad + off + c0 + j) : 0.f;
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const bool take = lane - d >= start;
#pragma unroll
      for (int j = 0; j < CW; ++j) {
        const float up = __shfl_up_sync(NRX_FULL_MASK, v[j], d);
        if (take) v[j] += up;
      }
    }
    if (is_end && live) {
      if (!from_prev && !to_next) {
        finalize_cols<CW>(T, my_key, v, c0);
      } else {
        float* dst = (from_prev ? head : tail) + tile * DP + c0;
#pragma unroll
        for (int j = 0; j < CW; ++j)
          if (c0 + j < DP) dst[j] = v[j];
      }
    }
  }
}

// Phase B: tile `t` owns a run iff its last run starts in t and continues into t+1.  The owner warp first
// finds how many following tiles the run covers (lanes test 32 tile-leading keys per round trip), then
// sums the partials in tile order with lanes on columns.
template <int NC>
__global__ void __launch_bounds__(128)
segment_fixup_kernel(const __grid_constant__ DTables T, const uint32_t* __restrict__ keys, long long n_occ,
                     long long n_tiles, uint32_t sentinel, const float* __restrict__ head,
                     const float* __restrict__ tail, int DP) {
  const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (t + 1 >= n_tiles) return;
  const long long ts = t * kTile, te = ts + kTile;  // full tile (not the last one)
  const uint32_t K = __ldg(keys + te - 1);
  if (K == sentinel || __ldg(keys + te) != K) return;               // does not continue
  const bool started_here = (__ldg(keys + ts) != K) || ts == 0 || __ldg(keys + ts - 1) != K;
  if (!started_here) return;                                        // an earlier tile owns it
  long long cont = 0;  // number of following tiles whose first key is K
  for (long long u0 = t + 1; u0 < n_tiles; u0 += 32) {
    const long long u = u0 + lane;
    const bool match = (u < n_tiles) && (__ldg(keys + u * kTile) == K);
    const unsigned mm = __ballot_sync(NRX_FULL_MASK, match);
    if (mm == 0xffffffffu) { cont += 32; continue; }
    cont += __ffs(~mm) - 1;
    break;
  }
  float acc[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) { const int c = lane + 32 * k; acc[k] = c < DP ? tail[t * DP + c] : 0.f; }
  long long u = t + 1;
  const long long uend = t + 1 + cont;
  for (; u + 8 <= uend; u += 8) {
    float h[8][NC];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int k = 0; k < NC; ++k) { const int c = lane + 32 * k; h[q][k] = c < DP ? __ldg(head + (u + q) * DP + c) : 0.f; }
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
      for (int k = 0; k < NC; ++k) acc[k] += h[q][k];
  }
  for (; u < uend; ++u)
#pragma unroll
    for (int k = 0; k < NC; ++k) { const int c = lane + 32 * k; if (c < DP) acc[k] += __ldg(head + u * DP + c); }
  finalize_row<NC>(T, K, acc, lane);
}

template <int NC>
static int launch_apply(const DFeats& d, const DTables& T, const PlanLayout& L, char* ws, const float* grad,
                        long long gld, cudaStream_t st) {
  const uint32_t sentinel = 1u << L.key_bits;
  const uint32_t* keys = (const uint32_t*)(ws + L.keys_out);
  const uint32_t* vals = (const uint32_t*)(ws + L.vals_out);
  float* head = (float*)(ws + L.head);
  float* tail = (float*)(ws + L.tail);
  const int wpb = 4;
  const unsigned blocks = (unsigned)((L.n_tiles + wpb - 1) / wpb);
  // Wide rows (D = 32): two passes of 16 columns when the gradient rows can be gathered as 16-byte quads (cfg3: the
  // 32-column instantiation needs 168+ registers, 16 % of the warp slots resident, every warp waiting on its gather:
  // 66 -> 54 us); with unaligned rows (cfg5: row width 115) the single 32-column pass stays faster (0.168 vs 0.182 ms/step)
  bool quads = (gld % 4 == 0) && (reinterpret_cast<uintptr_t>(grad) % 16 == 0);
  for (int i = 0; i < d.n && quads; ++i) quads = d.f[i].out_col % 4 == 0;
  if (d.max_dim <= 16 || (quads && getenv("NRX_K3_CW32") == nullptr))
    segment_reduce_kernel<16><<<blocks, wpb * 32, 0, st>>>(d, T, keys, vals, grad, gld, d.n_occ, sentinel, head, tail, L.dp, d.max_dim);
  else
    segment_reduce_kernel<32><<<blocks, wpb * 32, 0, st>>>(d, T, keys, vals, grad, gld, d.n_occ, sentinel, head, tail, L.dp, d.max_dim);
  int rc = check_launch("segment_reduce");
  if (rc != NRX_OK) return rc;
  if (L.n_tiles > 1) {
    segment_fixup_kernel<NC><<<blocks, wpb * 32, 0, st>>>(T, keys, d.n_occ, L.n_tiles, sentinel, head, tail, L.dp);
    rc = check_launch("segment_fixup");
  }
  return rc;
}

// ---- dense-AdamW semantics at sparse cost: the zero-gradient step on every row the batch does NOT touch ----------
struct SweepArgs {
  float* w[NRX_MAX_TABLES];
  float* m[NRX_MAX_TABLES];
  float* v[NRX_MAX_TABLES];
  long long unit_off[NRX_MAX_TABLES + 1];   // prefix of rows * units_per_row
  long long touched_off[NRX_MAX_TABLES];    // byte offset of the table's row map
  int upr[NRX_MAX_TABLES];                  // units (float4 or float) per row
  int stride[NRX_MAX_TABLES];
  int n_tables, vec;
  float beta1, beta2, eps, wd;
  const float* d_hp;
};

// one thread per occurrence, straight from the raw ids (no dependence on the sort plan): idempotent byte stores
__global__ void __launch_bounds__(256)
mark_rows_kernel(const __grid_constant__ DFeats P, const __grid_constant__ SweepArgs A, unsigned char* __restrict__ touched) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.n_occ) return;
  int f = 0;
  for (int i = 1; i < P.n; ++i)
    if (p >= P.f[i].occ_off) f = i;
  const DFeat& F = P.f[f];
  const long long r = p - F.occ_off;
  const long long id = load_idx(F.idx, r, F.idx32);
  bool valid = id > 0 && id < F.rows;   // same validity rule as the plan's keys
  if (valid && F.pool == NRX_POOL_MASKED_MEAN) valid = __ldg(F.mask + r) != 0.f;
  if (valid) touched[A.touched_off[F.table_id] + id] = 1;
}

// A deliberately SMALL grid of FULL-SM CTAs (1024 threads): it runs beside the forward/backward chain for most of
// the step, and a CTA that needs a whole SM can only land on one the persistent tower kernels leave free (128 tiles
// on 148 SMs at B = 16384) instead of squatting on theirs.  Two independent rows' worth of loads in flight per thread.
constexpr int kSweepThreads = 1024;
template <int V>
__global__ void __launch_bounds__(kSweepThreads, 1)
sweep_untouched_kernel(const __grid_constant__ SweepArgs A, const unsigned char* __restrict__ touched) {
  constexpr int U = 2;
  const float lr = __ldg(A.d_hp), bc1 = __ldg(A.d_hp + 1), bc2s = __ldg(A.d_hp + 2);
  const long long total = A.unit_off[A.n_tables];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long u0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; u0 < total; u0 += stride * U) {
    long long e[U];
    int tb[U];
    bool live[U];
    unsigned char hit[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const long long u = u0 + k * stride;
      live[k] = u < total;
      int t = 0;
      for (int i = 1; i < A.n_tables; ++i)
        if (u >= A.unit_off[i]) t = i;
      const long long local = live[k] ? u - A.unit_off[t] : 0;
      const long long row = (long long)((unsigned long long)local / (unsigned)A.upr[t]);
      tb[k] = t;
      e[k] = row * A.stride[t] + (local - row * A.upr[t]) * V;
      hit[k] = live[k] ? __ldg(touched + A.touched_off[t] + row) : (unsigned char)1;
    }
    // the row map and the row data are fetched in ONE round trip (loads for the few touched rows are wasted)
    if (V == 4) {
      float4 p[U], m[U], v[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (live[k]) {
          p[k] = *reinterpret_cast<const float4*>(A.w[tb[k]] + e[k]);
          m[k] = *reinterpret_cast<const float4*>(A.m[tb[k]] + e[k]);
          v[k] = *reinterpret_cast<const float4*>(A.v[tb[k]] + e[k]);
        }
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (!hit[k]) {   // touched rows belong to K3's fused update
          adamw_update(p[k].x, 0.f, m[k].x, v[k].x, lr, bc1, bc2s, A.beta1, A.beta2, A.eps, A.wd);
          adamw_update(p[k].y, 0.f, m[k].y, v[k].y, lr, bc1, bc2s, A.beta1, A.beta2, A.eps, A.wd);
          adamw_update(p[k].z, 0.f, m[k].z, v[k].z, lr, bc1, bc2s, A.beta1, A.beta2, A.eps, A.wd);
          adamw_update(p[k].w, 0.f, m[k].w, v[k].w, lr, bc1, bc2s, A.beta1, A.beta2, A.eps, A.wd);
          *reinterpret_cast<float4*>(A.w[tb[k]] + e[k]) = p[k];
          *reinterpret_cast<float4*>(A.m[tb[k]] + e[k]) = m[k];
          *reinterpret_cast<float4*>(A.v[tb[k]] + e[k]) = v[k];
        }
    } else {
      float p[U], m[U], v[U];
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (live[k]) { p[k] = A.w[tb[k]][e[k]]; m[k] = A.m[tb[k]][e[k]]; v[k] = A.v[tb[k]][e[k]]; }
#pragma unroll
      for (int k = 0; k < U; ++k)
        if (!hit[k]) {
          adamw_update(p[k], 0.f, m[k], v[k], lr, bc1, bc2s, A.beta1, A.beta2, A.eps, A.wd);
          A.w[tb[k]][e[k]] = p[k]; A.m[tb[k]][e[k]] = m[k]; A.v[tb[k]][e[k]] = v[k];
        }
    }
  }
}

}  // namespace nrx

extern "C" size_t nrx_embed_bwd_workspace_bytes(const NrxFeat* h_feats, int n_feats, int64_t B) {
  using namespace nrx;
  DFeats d;
  if (make_dfeats(h_feats, n_feats, B, nullptr, 0, &d) != NRX_OK) return 0;
  PlanLayout L;
  if (plan_layout(d, &L) != NRX_OK) return 0;
  return L.total;
}

extern "C" int nrx_embed_bwd_plan(const NrxFeat* h_feats, int n_feats, int64_t B, void* ws, size_t ws_bytes,
                                  nrx_stream_t stream) {
  return nrx_embed_bwd_plan_stage(h_feats, n_feats, B, ws, ws_bytes, NRX_PLAN_ALL, stream);
}

extern "C" int nrx_embed_bwd_plan_is_staged(const NrxFeat* h_feats, int n_feats, int64_t B) {
  using namespace nrx;
  DFeats d;
  if (make_dfeats(h_feats, n_feats, B, nullptr, 0, &d) != NRX_OK || d.n_occ == 0) return 0;
  SmallPlan sp;
  return make_small_plan(d, &sp) ? 1 : 0;
}

extern "C" int nrx_embed_bwd_plan_stage(const NrxFeat* h_feats, int n_feats, int64_t B, void* ws, size_t ws_bytes, int stage,
                                        nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(stage == NRX_PLAN_ALL || stage == NRX_PLAN_SORT || stage == NRX_PLAN_MERGE, NRX_EINVAL, "bad plan stage %d", stage);
  DFeats d;
  int rc = make_dfeats(h_feats, n_feats, B, nullptr, 0, &d);
  if (rc != NRX_OK) return rc;
  PlanLayout L;
  rc = plan_layout(d, &L);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws != nullptr && ws_bytes >= L.total, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, L.total);
  if (d.n_occ == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)ws;
  const uint32_t sentinel = 1u << L.key_bits;
  {
    // Small-batch fast path (every table <= 40960 occurrences): chunk sort + rank merge, see above.
    SmallPlan sp;
    if (make_small_plan(d, &sp)) {
      int max_n = 0;
      for (int s = 0; s < sp.n_seg; ++s) max_n = max(max_n, sp.seg_off[s + 1] - sp.seg_off[s]);
      const size_t smem = (size_t)max_n * sizeof(uint32_t);
      cudaError_t e = cudaFuncSetAttribute(plan_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "smem opt-in: %s", cudaGetErrorString(e));
      uint32_t* km = (uint32_t*)(w + L.keys_in);
      uint32_t* vm = (uint32_t*)(w + L.vals_in);
      if (stage != NRX_PLAN_MERGE) {
        plan_chunk_sort_kernel<<<sp.total_chunks, kChunkThreads, 0, st>>>(d, sp, km, vm);
        rc = check_launch("plan_chunk_sort");
        if (rc != NRX_OK) return rc;
      }
      if (stage == NRX_PLAN_SORT) return NRX_OK;
      plan_merge_kernel<<<sp.total_chunks * kChunkItems, kChunkThreads, smem, st>>>(sp, L.row_bits, sentinel, km, vm,
                                                                     (uint32_t*)(w + L.keys_out), (uint32_t*)(w + L.vals_out));
      return check_launch("plan_merge");
    }
  }
  if (stage == NRX_PLAN_MERGE) return NRX_OK;   // large batches: the device radix sort of the SORT stage is the whole plan
  const unsigned blocks = (unsigned)((d.n_occ + 255) / 256);
  build_keys_kernel<<<blocks, 256, 0, st>>>(d, L.row_bits, sentinel, (uint32_t*)(w + L.keys_in), (uint32_t*)(w + L.vals_in));
  rc = check_launch("build_keys");
  if (rc != NRX_OK) return rc;
  size_t cub_bytes = L.cub_bytes;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(w + L.cub, cub_bytes, (const uint32_t*)(w + L.keys_in),
                                                  (uint32_t*)(w + L.keys_out), (const uint32_t*)(w + L.vals_in),
                                                  (uint32_t*)(w + L.vals_out), (int)d.n_occ, 0, L.key_bits + 1, st);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "radix sort: %s", cudaGetErrorString(e));
  return NRX_OK;
}

extern "C" int nrx_embed_bwd_apply(const NrxFeat* h_feats, int n_feats, int64_t B, const float* grad_out,
                                   int64_t grad_ld, int mode, float* const* h_grads, float* const* h_tables,
                                   const NrxRowOpt* h_opt, void* ws, size_t ws_bytes, nrx_stream_t stream) {
  using namespace nrx;
  DFeats d;
  int rc = make_dfeats(h_feats, n_feats, B, nullptr, 0, &d);
  if (rc != NRX_OK) return rc;
  PlanLayout L;
  rc = plan_layout(d, &L);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(ws != nullptr && ws_bytes >= L.total, NRX_EWORKSPACE, "workspace %zu < %zu", ws_bytes, L.total);
  const bool no_zero = (mode & NRX_BWD_NO_ZERO) != 0;
  mode &= ~NRX_BWD_NO_ZERO;
  NRX_REQUIRE(mode == NRX_BWD_DENSE || mode == NRX_BWD_SGD || mode == NRX_BWD_ADAMW, NRX_EINVAL, "bad mode %d", mode);
  NRX_REQUIRE(!no_zero || mode == NRX_BWD_DENSE, NRX_EINVAL, "NRX_BWD_NO_ZERO only applies to NRX_BWD_DENSE");
  NRX_REQUIRE(grad_out != nullptr || B == 0, NRX_EINVAL, "null grad_out");
  cudaStream_t st = (cudaStream_t)stream;

  DTables T;
  memset(&T, 0, sizeof(T));
  T.mode = mode;
  T.row_bits = L.row_bits;
  long long trows[NRX_MAX_TABLES] = {0};
  for (int i = 0; i < d.n; ++i) {
    const DFeat& F = d.f[i];
    const int t = F.table_id;
    if (T.dim[t] != 0)
      NRX_REQUIRE(T.dim[t] == F.dim && T.stride[t] == F.stride && trows[t] == F.rows, NRX_EINVAL,
                  "features sharing table %d disagree on its shape", t);
    T.dim[t] = F.dim; T.stride[t] = F.stride; trows[t] = F.rows;
    NRX_REQUIRE(F.pool != NRX_POOL_MASKED_MEAN || F.inv_den != nullptr, NRX_EINVAL,
                "feature %d: masked-mean backward needs inv_den from the forward", i);
    NRX_REQUIRE(F.out_col + F.dim <= grad_ld, NRX_EINVAL, "feature %d overruns grad_ld", i);
  }
  for (int t = 0; t < d.n_tables; ++t) {
    if (T.dim[t] == 0) continue;
    if (mode == NRX_BWD_DENSE) {
      NRX_REQUIRE(h_grads && h_grads[t], NRX_EINVAL, "dense mode: missing grad buffer for table %d", t);
      T.g[t] = h_grads[t];
      if (!no_zero) {
        cudaError_t e = cudaMemsetAsync(T.g[t], 0, (size_t)trows[t] * T.stride[t] * sizeof(float), st);
        NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
      }
    } else {
      NRX_REQUIRE(h_tables && h_tables[t] && h_opt, NRX_EINVAL, "row-update mode: missing table %d / options", t);
      T.w[t] = h_tables[t];
      if (mode == NRX_BWD_ADAMW) {
        NRX_REQUIRE(h_opt->m[t] && h_opt->v[t], NRX_EINVAL, "AdamW: missing moments for table %d", t);
        T.m[t] = h_opt->m[t]; T.v[t] = h_opt->v[t];
      }
    }
  }
  if (mode != NRX_BWD_DENSE) {
    T.lr = h_opt->lr; T.beta1 = h_opt->beta1; T.beta2 = h_opt->beta2; T.eps = h_opt->eps; T.wd = h_opt->weight_decay;
    const int step = h_opt->step > 0 ? h_opt->step : 1;
    T.bc1 = (float)(1.0 - pow((double)T.beta1, (double)step));
    T.bc2_sqrt = (float)sqrt(1.0 - pow((double)T.beta2, (double)step));
    T.d_hp = h_opt->d_hparams;
  }
  if (d.n_occ == 0) return NRX_OK;
  char* w = (char*)ws;
  switch (L.nc) {
    case 1: return launch_apply<1>(d, T, L, w, grad_out, grad_ld, st);
    case 2: return launch_apply<2>(d, T, L, w, grad_out, grad_ld, st);
    default: return launch_apply<4>(d, T, L, w, grad_out, grad_ld, st);
  }
}

extern "C" size_t nrx_adamw_untouched_rows_scratch_bytes(const NrxFeat* h_feats, int n_feats) {
  using namespace nrx;
  DFeats d;
  if (make_dfeats(h_feats, n_feats, 0, nullptr, 0, &d) != NRX_OK) return 0;
  long long rows[NRX_MAX_TABLES] = {0};
  for (int i = 0; i < d.n; ++i) rows[d.f[i].table_id] = d.f[i].rows;
  size_t total = 0;
  for (int t = 0; t < d.n_tables; ++t) total += ((size_t)rows[t] + 15) & ~(size_t)15;
  return total ? total : 16;
}

extern "C" int nrx_adamw_untouched_rows(const NrxFeat* h_feats, int n_feats, int64_t B, float* const* h_tables,
                                        const NrxRowOpt* h_opt, void* scratch, size_t scratch_bytes, nrx_stream_t stream) {
  using namespace nrx;
  DFeats d;
  int rc = make_dfeats(h_feats, n_feats, B, nullptr, 0, &d);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(h_tables && h_opt && h_opt->d_hparams && scratch, NRX_EINVAL, "null tables / options / device hparams / scratch");
  SweepArgs A;
  memset(&A, 0, sizeof(A));
  long long rows[NRX_MAX_TABLES] = {0};
  int dim[NRX_MAX_TABLES] = {0};
  bool vec4 = true;
  for (int i = 0; i < d.n; ++i) {
    const DFeat& F = d.f[i];
    rows[F.table_id] = F.rows; dim[F.table_id] = F.dim; A.stride[F.table_id] = F.stride;
  }
  for (int t = 0; t < d.n_tables; ++t) {
    if (dim[t] == 0) continue;
    NRX_REQUIRE(h_tables[t] && h_opt->m[t] && h_opt->v[t], NRX_EINVAL, "missing table / moments for table %d", t);
    A.w[t] = h_tables[t]; A.m[t] = h_opt->m[t]; A.v[t] = h_opt->v[t];
    if (dim[t] % 4 || A.stride[t] % 4 || ((uintptr_t)A.w[t] | (uintptr_t)A.m[t] | (uintptr_t)A.v[t]) % 16) vec4 = false;
  }
  A.vec = vec4 ? 4 : 1;
  A.n_tables = d.n_tables;
  size_t toff = 0;
  for (int t = 0; t < d.n_tables; ++t) {
    A.upr[t] = dim[t] ? dim[t] / A.vec : 1;
    A.unit_off[t + 1] = A.unit_off[t] + (dim[t] ? rows[t] * A.upr[t] : 0);
    A.touched_off[t] = (long long)toff;
    toff += ((size_t)rows[t] + 15) & ~(size_t)15;
  }
  NRX_REQUIRE(scratch_bytes >= toff, NRX_EWORKSPACE, "row-map scratch %zu < %zu", scratch_bytes, toff);
  A.beta1 = h_opt->beta1; A.beta2 = h_opt->beta2; A.eps = h_opt->eps; A.wd = h_opt->weight_decay;
  A.d_hp = h_opt->d_hparams;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(scratch, 0, toff, st);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "memset: %s", cudaGetErrorString(e));
  if (d.n_occ > 0) {
    mark_rows_kernel<<<(unsigned)((d.n_occ + 255) / 256), 256, 0, st>>>(d, A, (unsigned char*)scratch);
    rc = check_launch("mark_rows");
    if (rc != NRX_OK) return rc;
  }
  const long long total = A.unit_off[d.n_tables];
  if (total == 0) return NRX_OK;
  long long blocks = (total + 2 * kSweepThreads - 1) / (2 * kSweepThreads);
  long long cap = sm_count() / 8 > 16 ? sm_count() / 8 : 16;   // ~1/8 of the SMs
  if (const char* e = getenv("NRX_SWEEP_CTAS")) { const long long v = atoll(e); if (v > 0) cap = v; }   // tuning knob
  if (blocks > cap) blocks = cap;
  if (vec4) sweep_untouched_kernel<4><<<(unsigned)blocks, kSweepThreads, 0, st>>>(A, (const unsigned char*)scratch);
  else sweep_untouched_kernel<1><<<(unsigned)blocks, kSweepThreads, 0, st>>>(A, (const unsigned char*)scratch);
  return check_launch("sweep_untouched");
}
