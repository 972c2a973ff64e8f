// Error plumbing, descriptor validation and small shared host helpers.
#include <stdarg.h>

#include "common.cuh"

namespace nrx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NRX_ELAUNCH;
  }
  return NRX_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

int img_zero_tail(void* image, int Kp, long long B, cudaStream_t st) {
  if (B % 128 == 0) return NRX_OK;
  const long long tile = B / 128;
  cudaError_t e = cudaMemsetAsync((uint8_t*)image + (size_t)tile * Kp * 256, 0, (size_t)Kp * 256, st);
  if (e != cudaSuccess) { set_error("image tail memset: %s", cudaGetErrorString(e)); return NRX_ELAUNCH; }
  return NRX_OK;
}

int make_dfeats(const NrxFeat* feats, int n, long long B, const void* out, long long out_ld, DFeats* d) {
  NRX_REQUIRE(feats != nullptr && n > 0 && n <= NRX_MAX_FEATS, NRX_EINVAL,
              "n_feats=%d outside [1,%d]", n, (int)NRX_MAX_FEATS);
  NRX_REQUIRE(B >= 0, NRX_EINVAL, "negative batch");
  memset(d, 0, sizeof(*d));
  d->n = n;
  bool vec4 = ((uintptr_t)out % 16 == 0) && (out_ld % 4 == 0);
  int max_table = -1;
  for (int i = 0; i < n; ++i) {
    const NrxFeat& s = feats[i];
    NRX_REQUIRE(s.table && (s.idx || B == 0), NRX_EINVAL, "feature %d: null table/idx", i);  // empty batch: null idx ok
    NRX_REQUIRE(s.rows > 0 && s.dim > 0 && s.row_stride >= s.dim, NRX_EINVAL, "feature %d: bad rows/dim/stride", i);
    NRX_REQUIRE(s.L >= 1, NRX_EINVAL, "feature %d: L must be >= 1", i);
    NRX_REQUIRE(s.pool == NRX_POOL_NONE || s.pool == NRX_POOL_MASKED_MEAN || s.pool == NRX_POOL_MEAN, NRX_EINVAL,
                "feature %d: bad pool mode %d", i, s.pool);
    NRX_REQUIRE(s.pool != NRX_POOL_NONE || s.L == 1, NRX_EINVAL, "feature %d: unpooled feature needs L == 1", i);
    NRX_REQUIRE(s.pool != NRX_POOL_MASKED_MEAN || s.mask, NRX_EINVAL, "feature %d: masked mean without mask", i);
    NRX_REQUIRE(s.idx_dtype == NRX_IDX_I64 || s.idx_dtype == NRX_IDX_I32, NRX_EINVAL, "feature %d: bad idx dtype", i);
    NRX_REQUIRE(s.table_id >= 0 && s.table_id < NRX_MAX_TABLES, NRX_EINVAL, "feature %d: bad table_id", i);
    NRX_REQUIRE(s.out_col >= 0, NRX_EINVAL, "feature %d: negative out_col", i);
    NRX_REQUIRE(s.dim <= 512, NRX_EUNSUPPORTED, "feature %d: dim %d > 512", i, s.dim);
    DFeat& t = d->f[i];
    t.table = s.table; t.idx = s.idx; t.mask = s.mask; t.inv_den = s.inv_den;
    t.rows = s.rows; t.dim = s.dim; t.stride = s.row_stride; t.L = s.L; t.pool = s.pool;
    t.out_col = s.out_col; t.idx32 = (s.idx_dtype == NRX_IDX_I32); t.table_id = s.table_id;
    t.occ_off = d->n_occ;
    d->n_occ += B * (long long)s.L;
    if (s.dim > d->max_dim) d->max_dim = s.dim;
    if (s.table_id > max_table) max_table = s.table_id;
    if ((uintptr_t)s.table % 16 != 0 || s.row_stride % 4 != 0 || s.dim % 4 != 0 || s.out_col % 4 != 0) vec4 = false;
    if (s.pool == NRX_POOL_NONE) d->sparse_ids[d->n_sparse++] = i;
    else d->array_ids[d->n_array++] = i;
  }
  d->n_tables = max_table + 1;
  d->vec = vec4 ? 4 : 1;
  int c = 0;
  for (int k = 0; k < d->n_sparse; ++k) {
    DFeat& t = d->f[d->sparse_ids[k]];
    t.cstart = c;
    c += t.dim / d->vec;
  }
  d->sparse_cols = c;
  return NRX_OK;
}

}  // namespace nrx

extern "C" int nrx_version(void) { return NRX_VERSION; }
extern "C" const char* nrx_last_error(void) { return nrx::g_err; }
