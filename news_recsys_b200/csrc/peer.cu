// K7 — gradient all-reduce fused with the dense AdamW over NVLink peer memory (include/nrx.h, "K7").
//
// One kernel per rank per step.  Slice r of the flat buffers (float4 granularity) belongs to rank r:
//   A  every rank tells every peer "my gradients are written" (release store into the peer's signal pad) and
//      waits until all peers said the same (acquire loads of its own pad);
//   1  the owner loads its slice of g from every rank (peer loads over NVLink, all `world` loads in flight,
//      summed in rank order, x 1/world), updates p/m/v with the torch.optim.AdamW rule and stores the new
//      parameters into every rank's parameter buffer (peer stores);
//   B  the last block of a rank to finish signals "my stores are out" to every peer and waits for theirs, so
//      the kernel retires only when this rank's parameters are complete and nobody still reads its gradients.
// Signal values are a launch counter kept in the pad itself (monotone; never reset by optimizer-state restores).
// Pad layout (u32 words): [0,16) A flags by source rank, [64,80) B flags, 128 launch counter, 129 block counter,
// 130 timed-out flag, [160,176) flags of the stand-alone barrier (nrx_peer_barrier), 192 its launch counter.
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"

namespace nrx {

constexpr int kPeerThreads = 256;
constexpr int kSigA = 0, kSigB = 64, kSigEpoch = 128, kSigBlocks = 129, kSigC = 160, kSigEpochC = 192;
constexpr unsigned kDefaultTimeoutMs = 20000;  // a peer that never arrives must not hang the GPU for ever

struct PeerArgs {
  int rank, world;
  float4* p[NRX_MAX_PEERS];
  const float4* g[NRX_MAX_PEERS];
  uint32_t* sig[NRX_MAX_PEERS];
  float4* m;
  float4* v;
  long long lo4, hi4;  // owned slice in float4 units
  const float* d_hp;
  float b1, b2, eps, wd;
  int* status;                 // host-visible trainer status word: bit 1 set when this rank gave up on a peer
  unsigned long long spin_ns;  // spin limit of one flag wait
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {  // written by another GPU before barrier A
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Spin until *flag >= want (wrap-safe signed distance); false on time-out or when the sticky error word of this
// rank's pad is set (a peer that timed out raises it on every rank, so nobody waits out the full limit).
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t want, const uint32_t* err, unsigned long long limit_ns) {
  if ((int32_t)(ld_acquire_sys(flag) - want) >= 0) return true;
  const unsigned long long t0 = now_ns();
  for (unsigned it = 0;; ++it) {
    if ((int32_t)(ld_acquire_sys(flag) - want) >= 0) return true;
    if ((it & 63u) == 63u && (now_ns() - t0 > limit_ns || *(const volatile uint32_t*)err != 0u)) return false;
    __nanosleep(32);
  }
}

// A time-out is FATAL and sticky: the error word is raised on this rank and on every peer, the trainer's status word
// gets bit 1 (the host raises at its next read-back), and every later launch on any rank returns at entry without
// touching parameters, moments, flags or epochs — ranks can no longer pair stale flags with new gradients.
__device__ __forceinline__ void raise_fatal(const PeerArgs& a) {
  for (int j = 0; j < a.world; ++j) st_relaxed_sys(a.sig[j] + NRX_PEER_SIG_ERR, 1u);
  if (a.status != nullptr) atomicOr(a.status, 2);
}

// W = compile-time bound on the world size (loads of all ranks in flight), U = float4 per thread per trip.
template <int W, int U>
__global__ void __launch_bounds__(kPeerThreads)
adamw_allreduce_peer_kernel(const __grid_constant__ PeerArgs a) {
  __shared__ int s_flag;
  uint32_t* sig = a.sig[a.rank];
  const int tid = threadIdx.x;
  if (*(volatile uint32_t*)(sig + NRX_PEER_SIG_ERR) != 0u) {   // dead exchange (sticky): do nothing, keep telling the host
    if (blockIdx.x == 0 && tid == 0 && a.status != nullptr) atomicOr(a.status, 2);
    return;
  }
  const uint32_t epoch = *(volatile uint32_t*)(sig + kSigEpoch) + 1u;  // bumped by the last block, after all read it
  if (tid == 0) s_flag = 0;
  __syncthreads();
  // ---- barrier A ------------------------------------------------------------------------------------
  // This rank's gradients were written by EARLIER kernels of the stream.  The flag is a RELEASE store at system scope
  // (fence.acq_rel.sys + store): together with the peers' acquire loads it orders those writes before the peers'
  // gradient loads under the PTX memory model, instead of relying on what a kernel boundary happens to flush.  At
  // kernel entry this rank has no outstanding stores of its own, so the fence is cheap (the costly fences of the first
  // version sat after the peer stores).
  if (blockIdx.x == 0 && tid < a.world) st_release_sys(a.sig[tid] + kSigA + a.rank, epoch);
  if (tid < a.world && !wait_flag(sig + kSigA + tid, epoch, sig + NRX_PEER_SIG_ERR, a.spin_ns)) s_flag = 1;
  __syncthreads();
  if (s_flag) {  // a peer never arrived: fatal, parameters untouched
    if (tid == 0) raise_fatal(a);
    return;
  }
  // ---- reduce + AdamW + broadcast over the owned slice ---------------------------------------------
  const float lr = __ldg(a.d_hp), bc1 = __ldg(a.d_hp + 1), bc2s = __ldg(a.d_hp + 2);
  const float inv_world = 1.f / (float)a.world;
  const long long stride = (long long)gridDim.x * kPeerThreads;
  float4* __restrict__ p_loc = a.p[a.rank];
  for (long long i0 = a.lo4 + (long long)blockIdx.x * kPeerThreads + tid; i0 < a.hi4; i0 += stride * U) {
    float4 gs[U][W], p[U], m[U], v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i < a.hi4) {
#pragma unroll
        for (int j = 0; j < W; ++j)
          if (j < a.world) gs[u][j] = ld_peer(a.g[j] + i);
        p[u] = p_loc[i]; m[u] = a.m[i]; v[u] = a.v[i];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * stride;
      if (i >= a.hi4) continue;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < a.world) { g.x += gs[u][j].x; g.y += gs[u][j].y; g.z += gs[u][j].z; g.w += gs[u][j].w; }
      g.x *= inv_world; g.y *= inv_world; g.z *= inv_world; g.w *= inv_world;
      adamw_update(p[u].x, g.x, m[u].x, v[u].x, lr, bc1, bc2s, a.b1, a.b2, a.eps, a.wd);
      adamw_update(p[u].y, g.y, m[u].y, v[u].y, lr, bc1, bc2s, a.b1, a.b2, a.eps, a.wd);
      adamw_update(p[u].z, g.z, m[u].z, v[u].z, lr, bc1, bc2s, a.b1, a.b2, a.eps, a.wd);
      adamw_update(p[u].w, g.w, m[u].w, v[u].w, lr, bc1, bc2s, a.b1, a.b2, a.eps, a.wd);
      a.m[i] = m[u];
      a.v[i] = v[u];
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (j < a.world) a.p[j][i] = p[u];
    }
  }
  // ---- barrier B: last block of this rank signals, then waits for every peer's stores ----------------
  // one system-scope fence per CTA: bar.sync orders every thread's stores before thread 0's fence (fences are cumulative)
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();                                   // release this CTA's peer stores (the one sys fence per CTA)
    s_flag = (atomicAdd(sig + kSigBlocks, 1u) == gridDim.x - 1) ? 2 : 0;
    if (s_flag == 2) __threadfence();                         // last CTA: acquire the other CTAs' counter increments
  }
  __syncthreads();
  if (s_flag != 2) return;
  if (tid < a.world) {
    st_release_sys(a.sig[tid] + kSigB + a.rank, epoch);
    if (!wait_flag(sig + kSigB + tid, epoch, sig + NRX_PEER_SIG_ERR, a.spin_ns)) raise_fatal(a);
  }
  __syncthreads();
  if (tid == 0) {
    sig[kSigBlocks] = 0u;
    *(volatile uint32_t*)(sig + kSigEpoch) = epoch;
  }
}


// ---- stand-alone barrier over the signal pads -------------------------------------------------------------------
// "Every rank's earlier work on this stream (peer stores included) is complete and visible before any rank's later
// work starts."  One CTA: fence.sys, release-store the launch counter into every peer's pad, acquire-wait for theirs.
__global__ void __launch_bounds__(32)
peer_barrier_kernel(const __grid_constant__ PeerArgs a) {
  uint32_t* sig = a.sig[a.rank];
  const int tid = threadIdx.x;
  if (*(volatile uint32_t*)(sig + NRX_PEER_SIG_ERR) != 0u) {
    if (tid == 0 && a.status != nullptr) atomicOr(a.status, 2);
    return;
  }
  const uint32_t epoch = *(volatile uint32_t*)(sig + kSigEpochC) + 1u;
  __syncwarp();
  bool ok = true;
  if (tid < a.world) {
    __threadfence_system();
    st_release_sys(a.sig[tid] + kSigC + a.rank, epoch);
    ok = wait_flag(sig + kSigC + tid, epoch, sig + NRX_PEER_SIG_ERR, a.spin_ns);
  }
  ok = __all_sync(NRX_FULL_MASK, ok);
  if (!ok) { if (tid == 0) raise_fatal(a); return; }
  if (tid == 0) *(volatile uint32_t*)(sig + kSigEpochC) = epoch;
}

// ---- row-sharded embedding tables: owner-side gather pushed straight into the requesters' feature rows -------------
// Rank `rank` owns rows [lo, hi) of every sharded table (stored behind a zero row 0).  For every sample of EVERY rank
// (ids all-gathered, [world * B]) whose id it owns, it reads the row and stores it into that rank's feature matrix
// x_r[b, out_col : out_col + dim] over NVLink (peer stores) — the "ids -> owner gather -> vectors" exchange of SURVEY
// §8e without a collective, a compaction or host-known sizes: each rank receives exactly B * sum(dim) * 4 bytes.
struct ShardFeat {
  const float* table;     // local shard: [hi - lo + 1, stride], row 0 = zeros
  const void* ids;        // global ids of all ranks [world * B]
  long long lo, hi;
  int dim, stride, out_col, idx32;
};
struct ShardArgs {
  ShardFeat f[NRX_MAX_FEATS];
  int n, rank, world;
  long long B;
  float* x[NRX_MAX_PEERS];        // push: feature matrices of every rank; pull: gradient matrices of every rank
  long long ld;
  float* g_global;                // pull: local [world * B, ld] gradient buffer (rows of rank r at r * B)
};

__global__ void __launch_bounds__(256)
shard_push_kernel(const __grid_constant__ ShardArgs a) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)a.world * a.B;
  if (warp >= total) return;
  const int r = (int)(warp / a.B);
  const long long b = warp % a.B;
  float* xr = a.x[r] + b * a.ld;
  for (int i = 0; i < a.n; ++i) {
    const ShardFeat& F = a.f[i];
    const long long id = load_idx(F.ids, warp, F.idx32);
    if (id < F.lo || id >= F.hi) continue;                       // another rank's row (or outside the table: nobody pushes)
    const float* row = F.table + (id - F.lo + 1) * F.stride;     // global padding id 0 -> shard 0's zeroed row
    for (int c = lane; c < F.dim; c += 32) xr[F.out_col + c] = __ldg(row + c);
  }
}

// Backward: the owner pulls, from every rank's gradient matrix, the rows of the samples whose id it owns into its local
// [world * B, ld] buffer (peer loads) — the operand of ONE K3 over the global batch with ids masked to the owned rows.
// `all_cols` features (replicated tables: every rank updates them identically) are pulled for every sample.
__global__ void __launch_bounds__(256)
shard_pull_kernel(const __grid_constant__ ShardArgs a, int n_all, const __grid_constant__ ShardArgs rep) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)a.world * a.B;
  if (warp >= total) return;
  const int r = (int)(warp / a.B);
  const long long b = warp % a.B;
  const float* gr = a.x[r] + b * a.ld;
  float* dst = a.g_global + warp * a.ld;
  for (int i = 0; i < a.n; ++i) {
    const ShardFeat& F = a.f[i];
    const long long id = load_idx(F.ids, warp, F.idx32);
    if (id < F.lo || id >= F.hi) continue;
    for (int c = lane; c < F.dim; c += 32) dst[F.out_col + c] = ld_peer_f32(gr + F.out_col + c);
  }
  for (int i = 0; i < n_all; ++i) {
    const ShardFeat& F = rep.f[i];
    for (int c = lane; c < F.dim; c += 32) dst[F.out_col + c] = ld_peer_f32(gr + F.out_col + c);
  }
}

}  // namespace nrx

extern "C" int nrx_peer_alloc(size_t bytes, void** ptr) {
  using namespace nrx;
  NRX_REQUIRE(ptr != nullptr && bytes > 0, NRX_EINVAL, "bad peer_alloc arguments");
  cudaError_t e = cudaMalloc(ptr, bytes);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  e = cudaMemset(*ptr, 0, bytes);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaMemset: %s", cudaGetErrorString(e));
  // the zero fill must have LANDED before the handle can reach a peer: a peer's first flag store must not be overwritten
  e = cudaDeviceSynchronize();
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaDeviceSynchronize: %s", cudaGetErrorString(e));
  return NRX_OK;
}

extern "C" int nrx_peer_free(void* ptr) {
  using namespace nrx;
  cudaError_t e = cudaFree(ptr);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaFree: %s", cudaGetErrorString(e));
  return NRX_OK;
}

extern "C" int nrx_peer_export(void* ptr, unsigned char handle[64]) {
  using namespace nrx;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  NRX_REQUIRE(ptr != nullptr && handle != nullptr, NRX_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  memcpy(handle, &h, 64);
  return NRX_OK;
}

extern "C" int nrx_peer_open(const unsigned char handle[64], void** ptr) {
  using namespace nrx;
  NRX_REQUIRE(ptr != nullptr && handle != nullptr, NRX_EINVAL, "null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return NRX_OK;
}

extern "C" int nrx_peer_close(void* ptr) {
  using namespace nrx;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "cudaIpcCloseMemHandle: %s", cudaGetErrorString(e));
  return NRX_OK;
}

extern "C" int nrx_adamw_allreduce_peer(const NrxPeerStep* s, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(s != nullptr, NRX_EINVAL, "null step");
  NRX_REQUIRE(s->world >= 1 && s->world <= NRX_MAX_PEERS && s->rank >= 0 && s->rank < s->world, NRX_EINVAL,
              "rank %d / world %d outside [1,%d]", s->rank, s->world, (int)NRX_MAX_PEERS);
  NRX_REQUIRE(s->n >= 0 && s->n % 4 == 0, NRX_EINVAL, "n=%lld must be a non-negative multiple of 4", (long long)s->n);
  NRX_REQUIRE(s->m && s->v && s->d_hparams, NRX_EINVAL, "null moments / hparams");
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = s->rank;
  a.world = s->world;
  for (int j = 0; j < s->world; ++j) {
    NRX_REQUIRE(s->p[j] && s->g[j] && s->sig[j], NRX_EINVAL, "rank %d: null peer buffer", j);
    NRX_REQUIRE((uintptr_t)s->p[j] % 16 == 0 && (uintptr_t)s->g[j] % 16 == 0, NRX_EINVAL, "rank %d: buffers must be 16-byte aligned", j);
    a.p[j] = (float4*)s->p[j];
    a.g[j] = (const float4*)s->g[j];
    a.sig[j] = s->sig[j];
  }
  NRX_REQUIRE((uintptr_t)s->m % 16 == 0 && (uintptr_t)s->v % 16 == 0, NRX_EINVAL, "moments must be 16-byte aligned");
  a.m = (float4*)s->m;
  a.v = (float4*)s->v;
  const long long n4 = s->n / 4, base = n4 / s->world, rem = n4 % s->world;  // same split as parallel.shard_range
  a.lo4 = s->rank * base + (s->rank < rem ? s->rank : rem);
  a.hi4 = a.lo4 + base + (s->rank < rem ? 1 : 0);
  a.d_hp = s->d_hparams;
  a.b1 = s->beta1; a.b2 = s->beta2; a.eps = s->eps; a.wd = s->weight_decay;
  a.status = s->status;
  unsigned ms = s->timeout_ms ? s->timeout_ms : kDefaultTimeoutMs;
  if (const char* e = getenv("NRX_PEER_TIMEOUT_MS")) { const long v = atol(e); if (v > 0) ms = (unsigned)v; }
  a.spin_ns = (unsigned long long)ms * 1000000ull;
  // Blocks that are not resident yet simply find barrier A already satisfied when they start: the grid needs
  // no co-residency on ITS device, only that every rank's kernel eventually starts on its own device.
  const int U = s->world <= 2 ? 4 : (s->world <= 4 ? 2 : 1);
  long long blocks = (a.hi4 - a.lo4 + (long long)kPeerThreads * U - 1) / ((long long)kPeerThreads * U);
  long long cap = 2LL * sm_count();   // one wave at 2 CTAs/SM (76-118 registers/thread); the loop strides
  if (const char* e = getenv("NRX_K7_CAP")) { const long long v = atoll(e); if (v > 0) cap = v; }   // tuning knob
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (s->world <= 2) adamw_allreduce_peer_kernel<2, 4><<<(unsigned)blocks, kPeerThreads, 0, st>>>(a);
  else if (s->world <= 4) adamw_allreduce_peer_kernel<4, 2><<<(unsigned)blocks, kPeerThreads, 0, st>>>(a);
  else if (s->world <= 8) adamw_allreduce_peer_kernel<8, 1><<<(unsigned)blocks, kPeerThreads, 0, st>>>(a);
  else adamw_allreduce_peer_kernel<16, 1><<<(unsigned)blocks, kPeerThreads, 0, st>>>(a);
  return check_launch("adamw_allreduce_peer");
}


extern "C" int nrx_peer_barrier(const NrxPeerStep* s, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(s != nullptr, NRX_EINVAL, "null step");
  NRX_REQUIRE(s->world >= 1 && s->world <= NRX_MAX_PEERS && s->rank >= 0 && s->rank < s->world, NRX_EINVAL, "bad rank / world");
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  a.rank = s->rank;
  a.world = s->world;
  for (int j = 0; j < s->world; ++j) {
    NRX_REQUIRE(s->sig[j], NRX_EINVAL, "rank %d: null signal pad", j);
    a.sig[j] = s->sig[j];
  }
  a.status = s->status;
  unsigned ms = s->timeout_ms ? s->timeout_ms : kDefaultTimeoutMs;
  if (const char* e = getenv("NRX_PEER_TIMEOUT_MS")) { const long v = atol(e); if (v > 0) ms = (unsigned)v; }
  a.spin_ns = (unsigned long long)ms * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a);
  return check_launch("peer_barrier");
}

static int make_shard_args(const NrxShardFeat* feats, int n, int rank, int world, int64_t B, float* const* h_mats, int64_t ld,
                           nrx::ShardArgs* a) {
  using namespace nrx;
  NRX_REQUIRE(n >= 0 && n <= NRX_MAX_FEATS && (feats || n == 0), NRX_EINVAL, "bad feature list");
  NRX_REQUIRE(world >= 1 && world <= NRX_MAX_PEERS && rank >= 0 && rank < world && B >= 0, NRX_EINVAL, "bad rank / world / B");
  memset(a, 0, sizeof(*a));
  a->n = n; a->rank = rank; a->world = world; a->B = B; a->ld = ld;
  for (int i = 0; i < n; ++i) {
    const NrxShardFeat& s = feats[i];
    NRX_REQUIRE(s.table && s.ids && s.dim >= 1 && s.row_stride >= s.dim && s.out_col >= 0 && s.out_col + s.dim <= ld && s.lo <= s.hi,
                NRX_EINVAL, "sharded feature %d: bad descriptor", i);
    a->f[i].table = s.table; a->f[i].ids = s.ids; a->f[i].lo = s.lo; a->f[i].hi = s.hi; a->f[i].dim = s.dim;
    a->f[i].stride = s.row_stride; a->f[i].out_col = s.out_col; a->f[i].idx32 = s.idx_dtype == NRX_IDX_I32;
  }
  if (h_mats != nullptr)
    for (int j = 0; j < world; ++j) {
      NRX_REQUIRE(h_mats[j], NRX_EINVAL, "rank %d: null matrix", j);
      a->x[j] = h_mats[j];
    }
  return NRX_OK;
}

extern "C" int nrx_shard_push(const NrxShardFeat* h_feats, int n_feats, int rank, int world, int64_t B, float* const* h_x, int64_t ld,
                              nrx_stream_t stream) {
  using namespace nrx;
  ShardArgs a;
  int rc = make_shard_args(h_feats, n_feats, rank, world, B, h_x, ld, &a);
  if (rc != NRX_OK) return rc;
  if (B == 0 || n_feats == 0) return NRX_OK;
  const long long warps = (long long)world * B;
  shard_push_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("shard_push");
}

extern "C" int nrx_shard_pull(const NrxShardFeat* h_feats, int n_feats, const NrxShardFeat* h_rep, int n_rep, int rank, int world,
                              int64_t B, float* const* h_g, int64_t ld, float* g_global, nrx_stream_t stream) {
  using namespace nrx;
  ShardArgs a, rep;
  int rc = make_shard_args(h_feats, n_feats, rank, world, B, h_g, ld, &a);
  if (rc != NRX_OK) return rc;
  rc = make_shard_args(h_rep, n_rep, rank, world, B, nullptr, ld, &rep);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(g_global != nullptr || B == 0, NRX_EINVAL, "null gradient buffer");
  a.g_global = g_global;
  if (B == 0) return NRX_OK;
  const long long warps = (long long)world * B;
  shard_pull_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a, n_rep, rep);
  return check_launch("shard_pull");
}

extern "C" int nrx_peer_status(const uint32_t* sig, int32_t* timed_out, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(sig && timed_out, NRX_EINVAL, "null pointer");
  uint32_t v = 0;
  cudaError_t e = cudaMemcpyAsync(&v, sig + NRX_PEER_SIG_ERR, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
  NRX_REQUIRE(e == cudaSuccess, NRX_ELAUNCH, "peer_status: %s", cudaGetErrorString(e));
  *timed_out = (int32_t)v;
  return NRX_OK;
}
