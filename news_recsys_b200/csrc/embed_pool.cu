// K1 — fused embedding gather + masked-mean pooling + concat.
//
// Replaces BaseModel.get_feature_embedding / array_feature_pooling /
// get_embeddings_from_batch (reference src/model/BaseModel/base_model.py:262-308):
// one launch writes out[b, out_col_f : out_col_f + D_f] for every feature instead of
// one aten::embedding + 4 elementwise/reduce kernels + a cat per feature.
//
// Work split (one launch):
//   * "sparse group" warps: SPW consecutive samples per warp; lane c owns one 16-byte
//     (or 4-byte on the scalar path) column of the concatenated sparse features, so the
//     row reads are 128-bit and the output row is written fully coalesced.
//   * "bag" warps: one warp per (sample, array feature).  32 ids + mask values are read
//     coalesced per step and broadcast by shuffle; 32/LPR rows are in flight per step
//     (LPR lanes cover one row); masked-out positions are never fetched.
// HBM-bound (gather): algorithmic bytes per sample in DESIGN.md §Kernels.
#include "common.cuh"

namespace nrx {

template <int V, int SPW>
__global__ void __launch_bounds__(256)
embed_pool_fwd_kernel(const __grid_constant__ DFeats P, long long B, float* __restrict__ out, long long ld,
                      int* __restrict__ status, long long n_sparse_warps, uint8_t* __restrict__ img, int img_kp) {
  using VT = VecT<V>;
  using vec_t = typename VT::type;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  bool bad = false;

  if (warp < n_sparse_warps) {
    const long long b0 = warp * SPW;
    for (int c = lane; c < P.sparse_cols; c += 32) {
      int f = P.sparse_ids[0];
      for (int k = 1; k < P.n_sparse; ++k) {
        const int cand = P.sparse_ids[k];
        if (c >= P.f[cand].cstart) f = cand;
      }
      const DFeat& F = P.f[f];
      const int off = (c - F.cstart) * V;
      long long id[SPW];
#pragma unroll
      for (int s = 0; s < SPW; ++s) id[s] = (b0 + s < B) ? load_idx(F.idx, b0 + s, F.idx32) : 0;
      vec_t val[SPW];
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        if ((unsigned long long)id[s] < (unsigned long long)F.rows) {
          val[s] = VT::load(F.table + id[s] * F.stride + off);
        } else {
          val[s] = VT::zero();
          bad = true;
        }
      }
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        if (b0 + s < B) {
          if (out != nullptr) VT::store(out + (b0 + s) * ld + F.out_col + off, val[s]);
          if constexpr (V == 4) { if (img != nullptr) img_store4(img, img_kp, b0 + s, F.out_col + off, val[s]); }
        }
      }
    }
  } else {
    const long long item = warp - n_sparse_warps;
    if (P.n_array == 0 || item >= B * P.n_array) return;
    const long long b = item / P.n_array;
    const DFeat& F = P.f[P.array_ids[(int)(item % P.n_array)]];
    const int nvec = F.dim / V;
    int LPR = 1;
    while (LPR < nvec && LPR < 32) LPR <<= 1;   // lanes per row (power of two)
    const int R = 32 / LPR;                     // rows in flight per step
    const int r = lane / LPR, c = lane % LPR;
    const int L = F.L;
    const long long base = b * (long long)L;
    const bool masked = (F.pool == NRX_POOL_MASKED_MEAN);
    float msum_lane = 0.f;
    float den = 1.f;
    for (int cbase = 0; cbase < nvec; cbase += LPR) {
      const bool cact = (cbase + c) < nvec;
      vec_t acc = VT::zero();
      for (int l0 = 0; l0 < L; l0 += 32) {
        const int my_l = l0 + lane;
        long long id_l = 0;
        float m_l = 0.f;
        if (my_l < L) {
          id_l = load_idx(F.idx, base + my_l, F.idx32);
          m_l = masked ? __ldg(F.mask + base + my_l) : 1.f;
          if ((unsigned long long)id_l >= (unsigned long long)F.rows) { bad = true; id_l = 0; m_l = masked ? m_l : 1.f; }
        }
        if (cbase == 0) msum_lane += m_l;
        const int n = min(32, L - l0);
#pragma unroll 4
        for (int k = 0; k * R < n; ++k) {
          const int j = k * R + r;
          const long long id = __shfl_sync(NRX_FULL_MASK, id_l, j & 31);
          const float m = __shfl_sync(NRX_FULL_MASK, m_l, j & 31);
          if (j < n && cact && m != 0.f) VT::fma(acc, m, VT::load(F.table + id * F.stride + (cbase + c) * V));
        }
      }
      for (int o = LPR; o < 32; o <<= 1) VT::add(acc, VT::shfl_xor(acc, o));
      if (cbase == 0) {
        const float msum = warp_sum(msum_lane);
        den = masked ? (msum + 1e-8f) : (float)L;   // base_model.py:281 / :276
      }
      if (r == 0 && cact) {
        const vec_t pooled = VT::div(acc, den);
        if (out != nullptr) VT::store(out + b * ld + F.out_col + (cbase + c) * V, pooled);
        if constexpr (V == 4) { if (img != nullptr) img_store4(img, img_kp, b, F.out_col + (cbase + c) * V, pooled); }
      }
    }
    if (F.inv_den != nullptr && lane == 0) F.inv_den[b] = 1.f / den;
  }
  if (bad && status != nullptr) atomicOr(status, 1);
}

template <int V, int SPW>
static int launch_fwd(const DFeats& d, long long B, float* out, long long ld, int* status, cudaStream_t st,
                      uint8_t* img = nullptr, int img_kp = 0) {
  const long long n_sparse_warps = d.n_sparse ? (B + SPW - 1) / SPW : 0;
  const long long warps = n_sparse_warps + B * d.n_array;
  if (warps == 0) return NRX_OK;
  const int wpb = 8;
  const long long blocks = (warps + wpb - 1) / wpb;
  NRX_REQUIRE(blocks < (1ll << 31), NRX_EUNSUPPORTED, "batch too large for one launch");
  embed_pool_fwd_kernel<V, SPW><<<(unsigned)blocks, wpb * 32, 0, st>>>(d, B, out, ld, status, n_sparse_warps, img, img_kp);
  return check_launch("embed_pool_fwd");
}

}  // namespace nrx

extern "C" int nrx_embed_pool_fwd(const NrxFeat* h_feats, int n_feats, int64_t B, float* out, int64_t out_ld,
                                  int32_t* status, nrx_stream_t stream) {
  using namespace nrx;
  DFeats d;
  NRX_REQUIRE(out != nullptr || B == 0, NRX_EINVAL, "null output");
  int rc = make_dfeats(h_feats, n_feats, B, out, out_ld, &d);
  if (rc != NRX_OK) return rc;
  for (int i = 0; i < d.n; ++i)
    NRX_REQUIRE(d.f[i].out_col + d.f[i].dim <= out_ld, NRX_EINVAL, "feature %d overruns out_ld", i);
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // Few samples per warp while the grid is small (latency-bound), 4 once the GPU is full.
  const bool small = B < (long long)sm_count() * 64;
  if (d.vec == 4) return small ? launch_fwd<4, 2>(d, B, out, out_ld, status, st) : launch_fwd<4, 4>(d, B, out, out_ld, status, st);
  return small ? launch_fwd<1, 2>(d, B, out, out_ld, status, st) : launch_fwd<1, 4>(d, B, out, out_ld, status, st);
}

// K1 writing the tower's input operand directly: besides (or instead of, out == NULL) the fp32 rows, the bf16 tile
// image [tile][width/8][128][8] that nrx_tower_fwd(..., NRX_TOWER_XIMG) streams — saves the fp32 -> bf16 pass and,
// for Deep (nobody else reads the fp32 concat in inference), the 4*width B/sample write.  Needs the 128-bit path
// (every dim / out_col a multiple of 4, 16-byte aligned tables) and width == the concat width, a multiple of 16.
extern "C" int nrx_embed_pool_fwd_img(const NrxFeat* h_feats, int n_feats, int64_t B, float* out, int64_t out_ld,
                                      void* image, int image_width, int32_t* status, nrx_stream_t stream) {
  using namespace nrx;
  DFeats d;
  NRX_REQUIRE(image != nullptr || B == 0, NRX_EINVAL, "null image");
  NRX_REQUIRE(image_width > 0 && image_width % 16 == 0, NRX_EUNSUPPORTED, "image width %d is not a multiple of 16", image_width);
  int rc = make_dfeats(h_feats, n_feats, B, out != nullptr ? (const void*)out : image, out != nullptr ? out_ld : 4, &d);
  if (rc != NRX_OK) return rc;
  NRX_REQUIRE(d.vec == 4, NRX_EUNSUPPORTED, "image output needs the 128-bit path (dims / columns multiples of 4, aligned tables)");
  int total = 0;
  for (int i = 0; i < d.n; ++i) {
    NRX_REQUIRE(d.f[i].out_col + d.f[i].dim <= image_width && (out == nullptr || d.f[i].out_col + d.f[i].dim <= out_ld), NRX_EINVAL,
                "feature %d overruns the output width", i);
    total += d.f[i].dim;
  }
  NRX_REQUIRE(total == image_width, NRX_EUNSUPPORTED, "features cover %d of %d image columns (pad columns would stay undefined)", total,
              image_width);
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rc = img_zero_tail(image, image_width, B, st);
  if (rc != NRX_OK) return rc;
  const bool small = B < (long long)sm_count() * 64;
  return small ? launch_fwd<4, 2>(d, B, out, out_ld, status, st, (uint8_t*)image, image_width)
               : launch_fwd<4, 4>(d, B, out, out_ld, status, st, (uint8_t*)image, image_width);
}
