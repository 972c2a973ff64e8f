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
                      int* __restrict__ status, long long n_sparse_warps, uint8_t* __restrict__ img, int img_kp,
                      float* __restrict__ fm_logit, int fm_lpf) {
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
      if constexpr (V == 4) {
        // FM logit in the pooling epilogue (fm/model.py:18-25 on the w / v split of :48-59): every field is fm_lpf
        // lanes wide (column 0 of its first lane = first-order weight), all fields fit one pass of the warp.
        //   logit = sum_f w_f + 1/2 sum_d [ (sum_f v_fd)^2 - sum_f v_fd^2 ]
        if (fm_logit != nullptr) {   // warp-uniform; host guarantees sparse_cols <= 32 (single pass) and equal widths
          const bool live = c < P.sparse_cols;
          const bool head = live && (off == 0);            // this lane holds columns 0..3 of its field
#pragma unroll
          for (int s = 0; s < SPW; ++s) {
            float4 v = live ? val[s] : make_float4(0.f, 0.f, 0.f, 0.f);
            const float w = head ? v.x : 0.f;
            if (head) v.x = 0.f;
            float q = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;           // sum_f v^2 (lane part)
            float4 S = v;                                                      // sum over fields of the same 4 dims
            for (int o = fm_lpf; o < 32; o <<= 1) {
              S.x += __shfl_xor_sync(NRX_FULL_MASK, S.x, o); S.y += __shfl_xor_sync(NRX_FULL_MASK, S.y, o);
              S.z += __shfl_xor_sync(NRX_FULL_MASK, S.z, o); S.w += __shfl_xor_sync(NRX_FULL_MASK, S.w, o);
            }
            // lanes 0 .. fm_lpf-1 hold every dimension's field sum exactly once
            float t = (lane < fm_lpf) ? 0.5f * (S.x * S.x + S.y * S.y + S.z * S.z + S.w * S.w) : 0.f;
            t += w - 0.5f * q;
            t = warp_sum(t);
            if (lane == 0 && b0 + s < B) fm_logit[b0 + s] = t;
          }
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
                      uint8_t* img = nullptr, int img_kp = 0, float* fm_logit = nullptr, int fm_lpf = 0) {
  const long long n_sparse_warps = d.n_sparse ? (B + SPW - 1) / SPW : 0;
  const long long warps = n_sparse_warps + B * d.n_array;
  if (warps == 0) return NRX_OK;
  const int wpb = 8;
  const long long blocks = (warps + wpb - 1) / wpb;
  NRX_REQUIRE(blocks < (1ll << 31), NRX_EUNSUPPORTED, "batch too large for one launch");
  embed_pool_fwd_kernel<V, SPW><<<(unsigned)blocks, wpb * 32, 0, st>>>(d, B, out, ld, status, n_sparse_warps, img, img_kp, fm_logit, fm_lpf);
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
// for Deep (nobody else reads the fp32 concat in inference), the 4*width B/sample write.  `fm_logit` (optional, [B]):
// the FM first + second order logit over all features, computed in the same epilogue (DeepFM: north_star item 2).  Needs the 128-bit path
// (every dim / out_col a multiple of 4, 16-byte aligned tables) and width == the concat width, a multiple of 16.
extern "C" int nrx_embed_pool_fwd_img(const NrxFeat* h_feats, int n_feats, int64_t B, float* out, int64_t out_ld,
                                      void* image, int image_width, float* fm_logit, int32_t* status, nrx_stream_t stream) {
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
  int fm_lpf = 0;
  if (fm_logit != nullptr) {   // FM over ALL features: sparse only, equal widths, one pass of the warp
    NRX_REQUIRE(d.n_array == 0 && d.n_sparse >= 1, NRX_EUNSUPPORTED, "fused FM logit: sparse (single-id) features only");
    const int D = d.f[d.sparse_ids[0]].dim;
    for (int k = 0; k < d.n_sparse; ++k)
      NRX_REQUIRE(d.f[d.sparse_ids[k]].dim == D, NRX_EUNSUPPORTED, "fused FM logit: fields must have equal widths");
    fm_lpf = D / 4;
    NRX_REQUIRE(fm_lpf >= 1 && (fm_lpf & (fm_lpf - 1)) == 0 && d.sparse_cols <= 32, NRX_EUNSUPPORTED,
                "fused FM logit: width/4 must be a power of two and all fields must fit 32 lanes (got %d x %d)", d.n_sparse, D);
  }
  if (B == 0) return NRX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rc = img_zero_tail(image, image_width, B, st);
  if (rc != NRX_OK) return rc;
  const bool small = B < (long long)sm_count() * 64;
  return small ? launch_fwd<4, 2>(d, B, out, out_ld, status, st, (uint8_t*)image, image_width, fm_logit, fm_lpf)
               : launch_fwd<4, 4>(d, B, out, out_ld, status, st, (uint8_t*)image, image_width, fm_logit, fm_lpf);
}
