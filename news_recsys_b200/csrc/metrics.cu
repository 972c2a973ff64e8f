// Validation metrics: per-user AUC / NDCG@k / HR@k / MRR@k on the device.
//
// Replaces the Python loop of BaseModel.on_validation_epoch_end (src/model/BaseModel/base_model.py:352-437): for every
// user a list of (score, label) tuples is sorted, roc_auc_score is called and the top-k walked in Python.  Here the
// samples arrive sorted by (user, score descending, arrival order) — two stable device sorts by the caller — and one
// thread walks one user's segment in fp64:
//   AUC  : Mann-Whitney with ties counted one half (what sklearn's trapezoidal ROC area equals), valid when the
//          user has both classes (:376);  label == 1 is "positive", anything else "negative" (:390);
//   top-k: the first min(k, n) entries of the segment (stable descending order == Python's sorted(reverse=True), :387);
//          users without positives score 0 and still count (:393-401).
// Not a hot path of the training step: thread-per-user keeps the arithmetic order of the reference.
#include "common.cuh"

namespace nrx {

__global__ void __launch_bounds__(128)
grouped_rank_metrics_kernel(const float* __restrict__ score, const float* __restrict__ label, const long long* __restrict__ seg,
                            long long n_users, int k, double* __restrict__ out, int* __restrict__ flags) {
  const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_users) return;
  const long long lo = seg[u], hi = seg[u + 1];
  long long n_pos = 0;
  for (long long i = lo; i < hi; ++i) n_pos += (__ldg(label + i) == 1.f);
  const long long n_neg = (hi - lo) - n_pos;
  double auc = 0.0;
  if (n_pos > 0 && n_neg > 0) {
    double acc = 0.0;
    long long neg_seen = 0;
    for (long long i = lo; i < hi;) {
      const float s = __ldg(score + i);
      long long p = 0, q = 0, j = i;
      for (; j < hi && __ldg(score + j) == s; ++j) {
        if (__ldg(label + j) == 1.f) ++p; else ++q;
      }
      acc += (double)p * ((double)(n_neg - neg_seen - q) + 0.5 * (double)q);
      neg_seen += q;
      i = j;
    }
    auc = acc / ((double)n_pos * (double)n_neg);
  }
  double hr = 0.0, dcg = 0.0, mrr = 0.0, ndcg = 0.0;
  if (n_pos > 0) {
    const long long top = (hi - lo) < k ? (hi - lo) : k;
    for (long long r = 1; r <= top; ++r) {
      if (__ldg(label + lo + r - 1) == 1.f) {
        hr = 1.0;
        dcg += 1.0 / log2((double)(r + 1));
        if (mrr == 0.0) mrr = 1.0 / (double)r;
      }
    }
    double idcg = 0.0;
    const long long ideal = n_pos < k ? n_pos : k;
    for (long long r = 1; r <= ideal; ++r) idcg += 1.0 / log2((double)(r + 1));
    ndcg = idcg > 0.0 ? dcg / idcg : 0.0;
  }
  out[u * 4 + 0] = auc; out[u * 4 + 1] = ndcg; out[u * 4 + 2] = hr; out[u * 4 + 3] = mrr;
  flags[u] = ((n_pos > 0 && n_neg > 0) ? 1 : 0) | (n_pos > 0 ? 2 : 0);
}

}  // namespace nrx

extern "C" int nrx_grouped_rank_metrics(const float* d_scores, const float* d_labels, const int64_t* d_seg_off, int64_t n_users,
                                        int32_t k, double* d_out, int32_t* d_flags, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(n_users >= 0 && k >= 1, NRX_EINVAL, "bad sizes (n_users=%lld, k=%d)", (long long)n_users, k);
  if (n_users == 0) return NRX_OK;
  NRX_REQUIRE(d_scores && d_labels && d_seg_off && d_out && d_flags, NRX_EINVAL, "null pointer");
  const long long blocks = (n_users + 127) / 128;
  grouped_rank_metrics_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(d_scores, d_labels, (const long long*)d_seg_off,
                                                                                 n_users, k, d_out, d_flags);
  return check_launch("grouped_rank_metrics");
}
