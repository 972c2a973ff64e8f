// Batch ingestion, host side: assemble one training batch straight into the pinned batch blob from a
// columnar feature file (int32 id columns, CSR offsets + values for array features).
//
// Replaces the per-sample Python of DataReader.__getitem__ (src/dataset/DataReader/data_reader.py:54-114: split(':')
// per item, list padding, one torch.tensor per feature) followed by torch's default collate.  Output layout is what
// that pair produces: ids [B] / [B, L] right-padded with 0 and truncated to the first L, mask [B, L] of 1/0 floats.
// Pure host code (no CUDA): it lives in libnrx.so so that one library is the whole boundary.
#include <stdint.h>
#include <stdlib.h>

#include <thread>
#include <vector>

#include "common.cuh"

namespace nrx {

// Row ranges of a batch are independent: large batches CAN be split over host threads (NRX_INGEST_THREADS, default 1);
// shuffled rows are random reads of memory-mapped columns, so each loop prefetches a few rows ahead.
static int ingest_threads() {
  static int n = [] {
    if (const char* e = getenv("NRX_INGEST_THREADS")) { const int v = atoi(e); if (v >= 1) return v > 64 ? 64 : v; }
    return 1;   // measured: spawning threads per call costs more than a 16 384-row batch takes (0.7 ms); opt in for huge B
  }();
  return n;
}

template <typename F>
static void parallel_rows(int64_t B, F fn) {
  const int nt = (B >= 8192) ? ingest_threads() : 1;
  if (nt <= 1) { fn((int64_t)0, B); return; }
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int t = 0; t < nt; ++t) {
    const int64_t b0 = B * t / nt, b1 = B * (t + 1) / nt;
    if (b0 < b1) th.emplace_back([=] { fn(b0, b1); });
  }
  for (auto& x : th) x.join();
}

static constexpr int64_t kAhead = 8;  // rows of look-ahead for the software prefetch

template <typename T>
static void gather_rows(const int32_t* col, const int64_t* rows, int64_t row0, int64_t b0, int64_t b1, T* out) {
  if (rows) {
    for (int64_t b = b0; b < b1; ++b) {
      if (b + kAhead < b1) __builtin_prefetch(col + rows[b + kAhead], 0, 0);
      out[b] = (T)col[rows[b]];
    }
  } else {
    for (int64_t b = b0; b < b1; ++b) out[b] = (T)col[row0 + b];
  }
}

template <typename T>
static void csr_expand(const int64_t* off, const int32_t* val, const int64_t* rows, int64_t row0, int64_t b0, int64_t b1, int L,
                       T* ids, float* mask) {
  for (int64_t b = b0; b < b1; ++b) {
    if (rows && b + kAhead < b1) {
      __builtin_prefetch(off + rows[b + kAhead], 0, 0);
      if (b + kAhead / 2 < b1) __builtin_prefetch(val + off[rows[b + kAhead / 2]], 0, 0);   // its offsets line arrived by now
    }
    const int64_t r = rows ? rows[b] : row0 + b;
    const int64_t lo = off[r];
    int64_t n = off[r + 1] - lo;
    if (n > L) n = L;  // keep the first L (data_reader.py:103-105)
    T* o = ids + b * L;
    float* m = mask ? mask + b * L : nullptr;
    for (int64_t l = 0; l < n; ++l) o[l] = (T)val[lo + l];
    for (int64_t l = n; l < L; ++l) o[l] = 0;
    if (m) {
      for (int64_t l = 0; l < n; ++l) m[l] = 1.f;
      for (int64_t l = n; l < L; ++l) m[l] = 0.f;
    }
  }
}

}  // namespace nrx

extern "C" int nrx_ingest_gather_ids(const int32_t* column, int64_t n_rows, const int64_t* rows, int64_t row0, int64_t B,
                                     void* out, int idx_dtype) {
  using namespace nrx;
  NRX_REQUIRE(column && out && B >= 0 && n_rows >= 0, NRX_EINVAL, "bad gather_ids arguments");
  NRX_REQUIRE(idx_dtype == NRX_IDX_I64 || idx_dtype == NRX_IDX_I32, NRX_EINVAL, "bad idx dtype %d", idx_dtype);
  if (rows) {
    for (int64_t b = 0; b < B; ++b) NRX_REQUIRE(rows[b] >= 0 && rows[b] < n_rows, NRX_EINVAL, "row %lld outside [0,%lld)", (long long)rows[b], (long long)n_rows);
  } else {
    NRX_REQUIRE(row0 >= 0 && row0 + B <= n_rows, NRX_EINVAL, "rows [%lld,%lld) outside the file", (long long)row0, (long long)(row0 + B));
  }
  parallel_rows(B, [=](int64_t b0, int64_t b1) {
    if (idx_dtype == NRX_IDX_I64) gather_rows<int64_t>(column, rows, row0, b0, b1, (int64_t*)out);
    else gather_rows<int32_t>(column, rows, row0, b0, b1, (int32_t*)out);
  });
  return NRX_OK;
}

extern "C" int nrx_ingest_csr_expand(const int64_t* offsets, const int32_t* values, int64_t n_rows, const int64_t* rows,
                                     int64_t row0, int64_t B, int32_t L, void* out_ids, int idx_dtype, float* out_mask) {
  using namespace nrx;
  NRX_REQUIRE(offsets && out_ids && B >= 0 && L >= 1 && n_rows >= 0, NRX_EINVAL, "bad csr_expand arguments");
  NRX_REQUIRE(values || offsets[n_rows] == 0, NRX_EINVAL, "null values");
  NRX_REQUIRE(idx_dtype == NRX_IDX_I64 || idx_dtype == NRX_IDX_I32, NRX_EINVAL, "bad idx dtype %d", idx_dtype);
  if (rows) {
    for (int64_t b = 0; b < B; ++b) NRX_REQUIRE(rows[b] >= 0 && rows[b] < n_rows, NRX_EINVAL, "row %lld outside [0,%lld)", (long long)rows[b], (long long)n_rows);
  } else {
    NRX_REQUIRE(row0 >= 0 && row0 + B <= n_rows, NRX_EINVAL, "rows [%lld,%lld) outside the file", (long long)row0, (long long)(row0 + B));
  }
  parallel_rows(B, [=](int64_t b0, int64_t b1) {
    if (idx_dtype == NRX_IDX_I64) csr_expand<int64_t>(offsets, values, rows, row0, b0, b1, L, (int64_t*)out_ids, out_mask);
    else csr_expand<int32_t>(offsets, values, rows, row0, b0, b1, L, (int32_t*)out_ids, out_mask);
  });
  return NRX_OK;
}

extern "C" int nrx_ingest_gather_labels(const float* labels, int64_t n_rows, int32_t n_labels, const int64_t* rows, int64_t row0,
                                        int64_t B, float* out, int32_t out_ld) {
  using namespace nrx;
  NRX_REQUIRE(labels && out && B >= 0 && n_labels >= 1 && out_ld >= 1, NRX_EINVAL, "bad gather_labels arguments");
  const int nc = n_labels < out_ld ? n_labels : out_ld;
  for (int64_t b = 0; b < B; ++b) {
    const int64_t r = rows ? rows[b] : row0 + b;
    NRX_REQUIRE(r >= 0 && r < n_rows, NRX_EINVAL, "row %lld outside [0,%lld)", (long long)r, (long long)n_rows);
    for (int c = 0; c < nc; ++c) out[b * out_ld + c] = labels[r * n_labels + c];
    for (int c = nc; c < out_ld; ++c) out[b * out_ld + c] = 0.f;
  }
  return NRX_OK;
}

// ---- device-resident feature file: assemble the batch blob on the GPU ----------------------------------------------
// The columnar file of a MIND-scale click log is tens of MB to a few GB: it fits HBM many times over, so the batch can
// be gathered where it is consumed.  One warp per sample: sparse ids by lane 0, array features by lanes striding over
// L (ids right-padded with 0 + mask), labels by the first lanes.  `rows` (device, nullable) selects shuffled rows.
namespace nrx {

struct IngestCols {
  const int32_t* data[NRX_MAX_FEATS];
  const int64_t* off[NRX_MAX_FEATS];   // null: sparse column
  void* out_ids[NRX_MAX_FEATS];
  float* out_mask[NRX_MAX_FEATS];
  int L[NRX_MAX_FEATS];
  int i32[NRX_MAX_FEATS];
  int n;
};

__global__ void __launch_bounds__(256)
ingest_assemble_kernel(const __grid_constant__ IngestCols C, const float* __restrict__ labels, int n_labels,
                       float* __restrict__ out_labels, int out_ld, long long n_rows, const long long* __restrict__ rows,
                       long long row0, long long B, int* __restrict__ status) {
  const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const long long r = rows ? __ldg(rows + b) : row0 + b;
  const bool ok = r >= 0 && r < n_rows;   // a row outside the file becomes an all-padding sample and raises status bit 2
  if (!ok && lane == 0 && status != nullptr) atomicOr(status, 4);
  for (int c = 0; c < C.n; ++c) {
    if (C.off[c] == nullptr) {
      if (lane == 0) {
        const int v = ok ? __ldg(C.data[c] + r) : 0;
        if (C.i32[c]) reinterpret_cast<int*>(C.out_ids[c])[b] = v;
        else reinterpret_cast<long long*>(C.out_ids[c])[b] = v;
      }
    } else {
      const int L = C.L[c];
      long long lo = 0, n = 0;
      if (ok) { lo = __ldg(C.off[c] + r); n = __ldg(C.off[c] + r + 1) - lo; }
      if (n > L) n = L;
      for (int l = lane; l < L; l += 32) {
        const int v = l < n ? __ldg(C.data[c] + lo + l) : 0;
        if (C.i32[c]) reinterpret_cast<int*>(C.out_ids[c])[b * L + l] = v;
        else reinterpret_cast<long long*>(C.out_ids[c])[b * L + l] = v;
        if (C.out_mask[c]) C.out_mask[c][b * L + l] = l < n ? 1.f : 0.f;
      }
    }
  }
  if (out_labels)
    for (int j = lane; j < out_ld; j += 32) out_labels[b * out_ld + j] = (ok && j < n_labels) ? __ldg(labels + r * n_labels + j) : 0.f;
}

}  // namespace nrx

extern "C" int nrx_ingest_assemble_device(const NrxIngestCol* h_cols, int n_cols, const float* labels, int32_t n_labels,
                                          float* out_labels, int32_t out_ld, int64_t n_rows, const int64_t* d_rows, int64_t row0,
                                          int64_t B, int32_t* status, nrx_stream_t stream) {
  using namespace nrx;
  NRX_REQUIRE(h_cols && n_cols >= 1 && n_cols <= NRX_MAX_FEATS, NRX_EINVAL, "n_cols=%d outside [1,%d]", n_cols, (int)NRX_MAX_FEATS);
  NRX_REQUIRE(B >= 0 && n_rows >= 0, NRX_EINVAL, "negative sizes");
  NRX_REQUIRE(!out_labels || (labels && n_labels >= 1 && out_ld >= 1), NRX_EINVAL, "bad label arguments");
  IngestCols C;
  memset(&C, 0, sizeof(C));
  C.n = n_cols;
  for (int c = 0; c < n_cols; ++c) {
    const NrxIngestCol& s = h_cols[c];
    NRX_REQUIRE(s.data && s.out_ids, NRX_EINVAL, "column %d: null data / output", c);
    NRX_REQUIRE(s.idx_dtype == NRX_IDX_I64 || s.idx_dtype == NRX_IDX_I32, NRX_EINVAL, "column %d: bad idx dtype", c);
    NRX_REQUIRE(s.offsets == nullptr || s.L >= 1, NRX_EINVAL, "column %d: array column needs L >= 1", c);
    C.data[c] = s.data; C.off[c] = s.offsets; C.out_ids[c] = s.out_ids; C.out_mask[c] = s.out_mask;
    C.L[c] = s.L; C.i32[c] = s.idx_dtype == NRX_IDX_I32;
  }
  if (B == 0) return NRX_OK;
  const long long blocks = (B * 32 + 255) / 256;
  ingest_assemble_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(C, labels, n_labels, out_labels, out_ld, n_rows,
                                                                           (const long long*)d_rows, row0, B, status);
  return check_launch("ingest_assemble");
}
