/*
 * nrx.h — C ABI of libnrx.so, the B200 (sm_100a) implementation of the
 * News_Recsys embedding-and-interaction hot path.
 *
 * The reference (ZhangHaoyang493/News_Recsys) has no FFI seam of its own: the
 * boundary is its Python object model.  Each entry point below
 * names the reference function(s) it replaces (paths relative to the reference
 * root) and is what a maintainer would bind from `src/model/BaseModel/base_model.py`
 * and the `src/model/sort/<model>/model.py` heads (see INTEGRATION.md for the
 * ctypes stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer valid on `stream`, unless the parameter
 *     name starts with `h_` (host);
 *   - return 0 on success, a negative NRX_E* code otherwise; nrx_last_error()
 *     returns a thread-local message for the last failure on the calling thread;
 *   - no call allocates or frees caller memory, synchronises the device or
 *     changes the current device; scratch comes from the caller (`ws`, sized by
 *     the matching *_workspace_bytes query);
 *   - no C++ exception crosses the boundary; the library is re-entrant.
 */
#ifndef NRX_H_
#define NRX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRX_VERSION 100 /* major*100 + minor */

typedef void* nrx_stream_t; /* cudaStream_t */

enum {
  NRX_OK = 0,
  NRX_EINVAL = -1,      /* bad argument */
  NRX_EUNSUPPORTED = -2,/* shape / dtype outside what the kernels handle */
  NRX_ELAUNCH = -3,     /* CUDA launch or runtime failure */
  NRX_EWORKSPACE = -4   /* workspace too small */
};

enum { NRX_POOL_NONE = 0, NRX_POOL_MASKED_MEAN = 1, NRX_POOL_MEAN = 2 };
enum { NRX_IDX_I64 = 0, NRX_IDX_I32 = 1 };
enum { NRX_MAX_FEATS = 16, NRX_MAX_TABLES = 16, NRX_MAX_LAYERS = 8 };

/* One feature of get_embeddings_from_batch (base_model.py:284-308). */
typedef struct NrxFeat {
  const float* table;   /* nn.Embedding weight [rows, row_stride] fp32 (base_model.py:164) */
  int64_t rows;
  int32_t dim;          /* logical D */
  int32_t row_stride;   /* elements between rows, >= dim */
  int32_t table_id;     /* 0..n_tables-1; features aliased by share_emb_table_features share an id */
  int32_t idx_dtype;    /* NRX_IDX_* (the reference casts with .long(), base_model.py:271) */
  const void* idx;      /* [B] (L == 1) or [B, L] row ids, 0 = padding row */
  int32_t L;            /* 1 for sparse features, array_max_length for array features */
  int32_t pool;         /* NRX_POOL_*: array_feature_pooling (base_model.py:273-282) */
  const float* mask;    /* [B, L] fp32 or NULL (NRX_POOL_MEAN / NONE) */
  float* inv_den;       /* optional out [B]: 1/(sum(mask)+1e-8), reused by the backward; may be NULL */
  int32_t out_col;      /* first column of this feature in the concatenated output */
  int32_t reserved;
} NrxFeat;

int nrx_version(void);
const char* nrx_last_error(void);

/* ---- K1: fused gather + masked-mean pooling + concat ----------------------
 * Replaces BaseModel.get_feature_embedding / array_feature_pooling /
 * get_embeddings_from_batch (base_model.py:262-308): out[b, out_col_f : +dim_f].
 * `status` (optional int32[1]) is set non-zero if any id is outside [0, rows). */
int nrx_embed_pool_fwd(const NrxFeat* h_feats, int n_feats, int64_t B,
                       float* out, int64_t out_ld, int32_t* status, nrx_stream_t stream);

/* Same, writing the tower's input operand as well: `image` = bf16 tile image [tile][image_width/8][128][8] of the
 * concatenated row (the layout nrx_tower_fwd streams with NRX_TOWER_XIMG; its slot inside the tower workspace comes
 * from nrx_tower_image_layout).  `out` may be NULL when nothing else reads the fp32 concat.  128-bit path only.
 * `fm_logit` (optional, [B]): the FM logit of fm/model.py:18-25 over ALL the features (w = column 0, v = the rest,
 * :48-59), computed in the pooling epilogue — sparse features of equal width only (width/4 a power of two, all fields
 * within 32 lanes); NRX_EUNSUPPORTED otherwise (use nrx_field_logit_fwd). */
int nrx_embed_pool_fwd_img(const NrxFeat* h_feats, int n_feats, int64_t B, float* out, int64_t out_ld,
                           void* image, int image_width, float* fm_logit, int32_t* status, nrx_stream_t stream);

/* ---- K3: deterministic sorted-index segment-reduce backward ----------------
 * Replaces aten::embedding_dense_backward behind base_model.py:271 (padding_idx=0
 * rows receive no gradient) plus the backward of the pooling at :278-282.
 *
 * nrx_embed_bwd_plan sorts the (table,row) keys of every (feature, sample,
 * position) occurrence once per batch; nrx_embed_bwd_apply then reduces
 * grad_out rows per unique table row in a fixed order (ascending occurrence)
 * and either writes dense per-table gradients (mode DENSE, `grads[t]` is
 * [rows_t, stride_t], zeroed by this call) or applies a fused sparse row update
 * in place (SGD / AdamW on touched rows only). */
enum { NRX_BWD_DENSE = 0, NRX_BWD_SGD = 1, NRX_BWD_ADAMW = 2,
       /* OR-ed onto NRX_BWD_DENSE: the caller has already zeroed grads[t] (e.g. with ONE memset over a flat gradient buffer,
        * early in the step and off its critical path); the call then only writes the touched rows */
       NRX_BWD_NO_ZERO = 0x100 };

typedef struct NrxRowOpt {
  float lr, beta1, beta2, eps, weight_decay;
  int32_t step;               /* 1-based, for Adam bias correction */
  float* m[NRX_MAX_TABLES];   /* AdamW first moment per table (same shape as the table) */
  float* v[NRX_MAX_TABLES];   /* AdamW second moment per table */
  const float* d_hparams;     /* optional DEVICE float[3] = {lr, 1-beta1^step, sqrt(1-beta2^step)}: read at run time
                                 instead of lr/step above, so a captured CUDA graph can be replayed across steps */
} NrxRowOpt;

size_t nrx_embed_bwd_workspace_bytes(const NrxFeat* h_feats, int n_feats, int64_t B);
int nrx_embed_bwd_plan(const NrxFeat* h_feats, int n_feats, int64_t B,
                       void* ws, size_t ws_bytes, nrx_stream_t stream);
/* The plan in two enqueues, for callers that schedule its halves around other kernels (the fused trainer runs the chunk
 * sort beside the pooling kernel and the merge beside the dW GEMMs, so that neither holds SMs the persistent tower
 * kernels need): SORT then MERGE on the same stream == ALL.  For batches on the device-radix-sort path SORT does
 * everything and MERGE is a no-op. */
enum { NRX_PLAN_ALL = 0, NRX_PLAN_SORT = 1, NRX_PLAN_MERGE = 2 };
int nrx_embed_bwd_plan_stage(const NrxFeat* h_feats, int n_feats, int64_t B, void* ws, size_t ws_bytes, int stage,
                             nrx_stream_t stream);
/* 1 when the batch takes the two-kernel path (chunk sort + merge), i.e. when NRX_PLAN_MERGE enqueues a kernel. */
int nrx_embed_bwd_plan_is_staged(const NrxFeat* h_feats, int n_feats, int64_t B);
int nrx_embed_bwd_apply(const NrxFeat* h_feats, int n_feats, int64_t B,
                        const float* grad_out, int64_t grad_ld,
                        int mode, float* const* h_grads /* [n_tables] device ptrs, DENSE */,
                        float* const* h_tables /* [n_tables] device ptrs, SGD/ADAMW */,
                        const NrxRowOpt* h_opt,
                        void* ws, size_t ws_bytes, nrx_stream_t stream);

/* ---- K2: field logits on the concatenated features --------------------------
 * FM    : fm/model.py:48-59 (w = col 0, v = cols 1..) + :18-25  -> first + second order
 * WIDE  : widedeep/model.py:58-65 + :25                          -> sum of col 0 of the listed fields
 * SUM   : lr/model.py:24-27                                      -> sum of every column of the listed fields
 * logit[b] (+)= term; backward adds into grad_x. */
enum { NRX_FIELD_FM = 0, NRX_FIELD_WIDE = 1, NRX_FIELD_SUM = 2 };
int nrx_field_logit_fwd(const float* x, int64_t ld, int64_t B, const int32_t* h_cols,
                        const int32_t* h_dims, int n_fields, int mode,
                        float* logit, int accumulate, nrx_stream_t stream);
int nrx_field_logit_bwd(const float* x, int64_t ld, int64_t B, const int32_t* h_cols,
                        const int32_t* h_dims, int n_fields, int mode,
                        const float* dlogit, float* grad_x, int64_t grad_ld, int accumulate,
                        nrx_stream_t stream);

/* Gather-fused FM for sparse-only fields of equal width (BASELINE config 2):
 * gathers the rows, never materialises the concat, and emits prob (+ per-sample
 * BCE and dL/dlogit when label != NULL).  fm/model.py:18-26,43-59 + bceLoss :39-40. */
int nrx_fm_fused_fwd(const NrxFeat* h_feats, int n_feats, int64_t B, const float* bias,
                     const float* label, int64_t label_stride,
                     float* logit, float* prob, float* loss_per_sample, float* dlogit,
                     int32_t* status, nrx_stream_t stream);
/* grad_x[b, out_col_f + d] = dlogit[b] * (d == 0 ? 1 : S_d - v_fd)  (S re-gathered). */
int nrx_fm_fused_bwd(const NrxFeat* h_feats, int n_feats, int64_t B, const float* dlogit,
                     float* grad_x, int64_t grad_ld, nrx_stream_t stream);

/* ---- sigmoid + BCE on probabilities (bceLoss, deep/model.py:32-33) ----------
 * prob = sigmoid(sum_t terms[t][b] + bias[0]); loss_b = -[y log p + (1-y) log(1-p)]
 * with log clamped at -100; dlogit_b = (p-y)/max(p(1-p),1e-12) * p(1-p) / B. */
int nrx_logit_loss_fwd(const float* const* h_terms, int n_terms, const float* bias, int64_t B,
                       const float* label, int64_t label_stride,
                       float* prob, float* loss_per_sample, float* dlogit, nrx_stream_t stream);
/* Un-fused pieces of the same chain, for the autograd route (loss.backward()):
 *   loss_b = BCE(prob_b, y_b)                          (bceLoss forward)
 *   grad_prob_b = (p-y)/max(p(1-p),1e-12) * upstream[0] / B   (its autograd)
 *   grad_logit_b = grad_prob_b * p(1-p)               (torch.sigmoid autograd) */
int nrx_bce_fwd(const float* prob, const float* label, int64_t label_stride, int64_t B,
                float* loss_per_sample, nrx_stream_t stream);
int nrx_bce_bwd(const float* prob, const float* label, int64_t label_stride, int64_t B,
                const float* upstream, float* grad_prob, nrx_stream_t stream);
int nrx_sigmoid_bwd(const float* prob, const float* grad_prob, int64_t B, float* grad_logit,
                    nrx_stream_t stream);
/* Two independent reductions in one launch: out1[0] = scale1 * sum(x1[0..n1)), out2[0] = scale2 * sum(x2[0..n2)). */
int nrx_reduce2_f32(const float* x1, int64_t n1, float scale1, float* out1,
                    const float* x2, int64_t n2, float scale2, float* out2, nrx_stream_t stream);
/* Deterministic mean/sum of n floats into out[0] (fixed reduction tree). */
int nrx_reduce_f32(const float* x, int64_t n, float scale, float* out, nrx_stream_t stream);

/* ---- dense AdamW exactly as torch.optim.AdamW (deep/model.py:55) ------------ */
int nrx_adamw_dense(float* p, const float* g, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, float weight_decay,
                    int32_t step, nrx_stream_t stream);

/* Same update with {lr, 1-beta1^step, sqrt(1-beta2^step)} read from DEVICE memory (CUDA-graph replay). */
int nrx_adamw_dense_dev(float* p, const float* g, float* m, float* v, int64_t n,
                        const float* d_hparams, float beta1, float beta2, float eps, float weight_decay,
                        nrx_stream_t stream);

/* Device-side optimizer clock: ++d_step[0]; d_hparams = {lr per CosinDecayLR (model_utils/lr_schedule.py:16-28)
 * at scheduler step d_step-1, 1-beta1^d_step, sqrt(1-beta2^d_step)}.  Lets a whole training step replay as a
 * CUDA graph with no host-written scalars. */
int nrx_hparams_step(int32_t* d_step, float* d_hparams, float lr, float min_lr, int32_t milestone0,
                     int32_t milestone1, float beta1, float beta2, nrx_stream_t stream);

/* ---- batch ingestion, HOST side ---------------------------------------------------------------
 * Replaces DataReader.__getitem__ + default collate (src/dataset/DataReader/data_reader.py:54-114): the per-sample
 * split(':') / list padding / torch.tensor of the reference becomes three copies per batch out of a columnar feature
 * file (int32 id columns; array features as CSR offsets + values, already truncated to max_len; float labels),
 * written STRAIGHT into the pinned batch blob.  All pointers here are HOST pointers; no CUDA call is made.
 * `rows`: B row numbers (shuffled batches) or NULL for the contiguous rows [row0, row0 + B).
 * Output = what the reference collates: ids right-padded with 0, mask 1/0, first L ids kept. */
int nrx_ingest_gather_ids(const int32_t* column, int64_t n_rows, const int64_t* rows, int64_t row0, int64_t B,
                          void* out, int idx_dtype);
int nrx_ingest_csr_expand(const int64_t* offsets, const int32_t* values, int64_t n_rows, const int64_t* rows,
                          int64_t row0, int64_t B, int32_t L, void* out_ids, int idx_dtype, float* out_mask /* nullable */);
int nrx_ingest_gather_labels(const float* labels, int64_t n_rows, int32_t n_labels, const int64_t* rows, int64_t row0,
                             int64_t B, float* out, int32_t out_ld);

/* Device-resident feature file: the same batch assembled ON THE GPU from device copies of the columns (the whole
 * columnar click log fits HBM), one warp per sample; `d_rows` (device int64[B], nullable) selects shuffled rows, a row
 * outside the file yields an all-padding sample and sets bit 2 of `status` (the trainer raises at its next status read).  All pointers are DEVICE pointers except `h_cols`. */
typedef struct NrxIngestCol {
  const int32_t* data;      /* sparse: ids [n_rows]; array: CSR values */
  const int64_t* offsets;   /* array: CSR offsets [n_rows + 1]; NULL for a sparse column */
  int32_t L;                /* array: padded length of the output */
  int32_t idx_dtype;        /* dtype of out_ids: NRX_IDX_I64 / NRX_IDX_I32 */
  void* out_ids;            /* [B] or [B, L] */
  float* out_mask;          /* [B, L] or NULL */
} NrxIngestCol;
int nrx_ingest_assemble_device(const NrxIngestCol* h_cols, int n_cols, const float* labels, int32_t n_labels,
                               float* out_labels, int32_t out_ld, int64_t n_rows, const int64_t* d_rows, int64_t row0,
                               int64_t B, int32_t* status /* nullable device word: bit 2 (value 4) is OR-ed in when a row
                               index lies outside the file */, nrx_stream_t stream);

/* ---- validation metrics ------------------------------------------------------------------------
 * Per-user AUC / NDCG@k / HR@k / MRR@k, replacing the Python loop of BaseModel.on_validation_epoch_end
 * (src/model/BaseModel/base_model.py:352-437).  Samples sorted by (user, score descending, arrival order);
 * d_seg_off [n_users + 1] delimits the users.  d_out [n_users, 4] = {auc, ndcg, hr, mrr} in fp64,
 * d_flags [n_users]: bit 0 = both classes present (AUC counts towards GAUC), bit 1 = has a positive. */
int nrx_grouped_rank_metrics(const float* d_scores, const float* d_labels, const int64_t* d_seg_off, int64_t n_users,
                             int32_t k, double* d_out, int32_t* d_flags, nrx_stream_t stream);

/* ---- dense-AdamW semantics at sparse cost ---------------------------------------------------------------
 * The reference's torch.optim.AdamW (deep/model.py:55) sweeps whole tables every step.  For a row no occurrence of
 * the batch touches the gradient is exactly zero, so its update does not depend on the backward pass:
 *   nrx_adamw_untouched_rows   — as soon as the ids are on the device: marks the rows the batch touches (same
 *                                validity rule as the plan: id 0, masked and out-of-range ids touch nothing) and
 *                                applies the g = 0 AdamW step to every OTHER row — off the critical path, beside
 *                                forward/backward, on a few SMs; it writes rows the batch never reads;
 *   nrx_embed_bwd_apply(ADAMW) — the fused row update of the touched rows.
 * Together they equal one dense AdamW step over tables + moments.  `scratch`: row map, sized by *_scratch_bytes.
 * `opt->d_hparams` (device {lr, 1-beta1^t, sqrt(1-beta2^t)}) is required. */
size_t nrx_adamw_untouched_rows_scratch_bytes(const NrxFeat* h_feats, int n_feats);
int nrx_adamw_untouched_rows(const NrxFeat* h_feats, int n_feats, int64_t B, float* const* h_tables,
                             const NrxRowOpt* opt, void* scratch, size_t scratch_bytes, nrx_stream_t stream);

/* ---- K7: gradient all-reduce FUSED with the dense AdamW over NVLink peer memory (multi-GPU).
 * The reference trains on one GPU (`devices=1`, sort/deep/train.py:41-42); this replaces what
 * DistributedDataParallel's NCCL all-reduce + torch.optim.AdamW.step() would be for its models.
 * Every rank owns one contiguous slice of the flat parameter buffer.  ONE kernel per rank per step:
 *   barrier (all ranks finished writing their gradients)  ->  load the owned slice of g from EVERY rank's
 *   buffer (peer loads, summed in rank order, x 1/world)  ->  AdamW on the slice with the local moments  ->
 *   store the new parameters into EVERY rank's parameter buffer (peer stores)  ->  barrier.
 * Replicas stay bitwise identical (each element is computed once, by its owner) and the optimizer state of a
 * slice is live on its owner only.  Buffers come from nrx_peer_alloc (cudaMalloc'd, IPC-exportable); the 64-byte
 * handles travel through the host's own control plane (e.g. torch.distributed all_gather_object). */
enum { NRX_MAX_PEERS = 16, NRX_PEER_SIG_WORDS = 256 };

int nrx_peer_alloc(size_t bytes, void** ptr);                        /* zero-filled device memory */
int nrx_peer_free(void* ptr);
int nrx_peer_export(void* ptr, unsigned char handle[64]);            /* cudaIpcGetMemHandle */
int nrx_peer_open(const unsigned char handle[64], void** ptr);       /* map a peer's buffer (same node) */
int nrx_peer_close(void* ptr);

typedef struct NrxPeerStep {
  int32_t rank, world;
  float* p[NRX_MAX_PEERS];        /* flat parameter buffer of every rank; p[rank] is the local one */
  const float* g[NRX_MAX_PEERS];  /* flat gradient buffer of every rank */
  uint32_t* sig[NRX_MAX_PEERS];   /* signal pad of every rank: NRX_PEER_SIG_WORDS u32, zero at start */
  float* m;                       /* local AdamW moments, n floats (only the owned slice is touched) */
  float* v;
  int64_t n;                      /* floats in the flat buffers; multiple of 4, buffers 16-byte aligned */
  const float* d_hparams;         /* {lr, 1-beta1^t, sqrt(1-beta2^t)} on the device (nrx_hparams_step) */
  float beta1, beta2, eps, weight_decay;
  int32_t* status;                /* optional DEVICE int32: bit 1 (value 2) is OR-ed in when the exchange is dead */
  uint32_t timeout_ms;            /* spin limit of one flag wait; 0 = 20 s (env NRX_PEER_TIMEOUT_MS overrides) */
  uint32_t reserved;
} NrxPeerStep;

/* Returns NRX_OK after the launch.  A peer that does not arrive within the limit is FATAL and sticky: the kernel
 * raises sig[.][NRX_PEER_SIG_ERR] on every rank and bit 1 of `status`, leaves parameters / moments / epochs untouched,
 * and every later launch on any rank returns at entry (no rank can pair stale flags with new gradients).  The host
 * must treat it as the end of the run (nrx_peer_status, or the status word it reads back with the loss). */
enum { NRX_PEER_SIG_ERR = 130 };
int nrx_adamw_allreduce_peer(const NrxPeerStep* step, nrx_stream_t stream);
int nrx_peer_status(const uint32_t* sig, int32_t* timed_out, nrx_stream_t stream);  /* synchronises the stream */

/* Stand-alone barrier over the same signal pads (only rank / world / sig / status / timeout_ms of `step` are read):
 * every rank's earlier work on its stream, peer stores included, is complete and visible before any rank continues.
 * Same fatal / sticky time-out rule as K7. */
int nrx_peer_barrier(const NrxPeerStep* step, nrx_stream_t stream);

/* ---- row-sharded embedding tables over peer memory (ids -> owner gather -> pooled vectors) ------------
 * The reference keeps every nn.Embedding whole on one GPU (base_model.py:141-166); BASELINE config 5 shards the big
 * tables by contiguous row ranges.  Rank `rank` stores rows [lo, hi) of a sharded table behind a zero row 0.
 *   nrx_shard_push — forward: for every sample of EVERY rank (ids all-gathered: [world * B]) whose id this rank owns,
 *     gather the row and store it straight into that rank's feature matrix h_x[r][b, out_col : +dim] (peer stores):
 *     each rank receives exactly B * sum(dim) * 4 bytes, no collective, no compaction, no host-known sizes;
 *   nrx_shard_pull — backward: copy, from every rank's gradient matrix h_g[r], the columns of the samples whose id this
 *     rank owns (and, for the `h_rep` features of replicated tables, of every sample) into the local [world * B, ld]
 *     buffer `g_global` (peer loads) — the operand of one nrx_embed_bwd_apply over the global batch.
 * A barrier (nrx_peer_barrier, or any collective on the stream) must separate a push from the readers of x. */
typedef struct NrxShardFeat {
  const float* table;   /* push: local shard [hi - lo + 1, row_stride] fp32, row 0 zero; pull: may equal any non-NULL pointer */
  const void* ids;      /* global ids of all ranks, [world * B] */
  int64_t lo, hi;       /* owned row range */
  int32_t dim, row_stride, out_col, idx_dtype;
} NrxShardFeat;
int nrx_shard_push(const NrxShardFeat* h_feats, int n_feats, int rank, int world, int64_t B,
                   float* const* h_x /* [world] peer-mapped feature matrices */, int64_t ld, nrx_stream_t stream);
int nrx_shard_pull(const NrxShardFeat* h_feats, int n_feats, const NrxShardFeat* h_rep, int n_rep, int rank, int world,
                   int64_t B, float* const* h_g /* [world] peer-mapped gradient matrices */, int64_t ld,
                   float* g_global, nrx_stream_t stream);

/* ---- K4/K5: fused bf16 tower on tcgen05 (MLP utils.py:6-17, DSSM towers
 * recall/DSSM/model.py:26-44, DCN cross dcn_arch.py:14-30,53-70) --------------- */
enum { NRX_ACT_RELU = 0, NRX_ACT_LEAKY = 1 };
typedef struct NrxTower {
  int32_t n_layers;                   /* Linear layers */
  int32_t dims[NRX_MAX_LAYERS + 1];   /* in, hidden..., out */
  const float* w[NRX_MAX_LAYERS];     /* [out, in] fp32 (nn.Linear layout) */
  const float* b[NRX_MAX_LAYERS];     /* [out] */
  int32_t act;                        /* NRX_ACT_* between layers, none after the last */
  float negative_slope;
} NrxTower;

size_t nrx_tower_workspace_bytes(const NrxTower* h_tower, int64_t B, int training);
/* Packs the fp32 weights into the bf16 operand images kept in `ws` (same B / training as the forward that
 * follows).  nrx_tower_fwd does this itself unless `training` carries NRX_TOWER_PREPACKED — a trainer can
 * issue the pack on another stream as soon as the optimizer has written the weights. */
enum { NRX_TOWER_TRAINING = 1, NRX_TOWER_PREPACKED = 2 };
int nrx_tower_pack(const NrxTower* h_tower, int64_t B, int training, void* ws, size_t ws_bytes, nrx_stream_t stream);
/* y[B, dims[n]] = tower(x[B, dims[0]]); with training != 0 the bf16 activations of
 * every layer are kept in `ws` for nrx_tower_bwd. */
int nrx_tower_fwd(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B,
                  float* y, int64_t y_ld, int training, void* ws, size_t ws_bytes,
                  nrx_stream_t stream);
/* NRX_TOWER_XIMG: the bf16 tile image of the input (layer 0's operand, see nrx_tower_image_layout) is already in `ws`
 * — written there by the producing kernel (nrx_embed_pool_fwd_img / nrx_dcn_cross_fwd with an image) — `x` may be NULL. */
enum { NRX_TOWER_XIMG = 4 };

/* Fused head for towers that end in ONE logit (dims[n_layers] == 1): the last epilogue of the forward computes
 *   z = sum_t terms[t][b] + tower(x)[b] + bias[0];  prob = sigmoid(z)
 * and, with a label, the per-sample BCE on probabilities (log clamped at -100: bceLoss, deep/model.py:32-33) and
 * dL/dlogit of the MEAN loss — the arithmetic of nrx_logit_loss_fwd without the extra launch and the [B] round trip.
 * Replaces `torch.sigmoid(wide + deep)` (widedeep/model.py:27), `sigmoid(score_fc(x))` (deep/model.py:21,
 * dcn/model.py:29) followed by F.binary_cross_entropy.  Any output pointer may be NULL. */
typedef struct NrxTowerHead {
  const float* terms[4];   /* extra per-sample logit terms [B] (FM / wide logit), added before the tower's */
  int32_t n_terms;
  int32_t reserved;
  const float* bias;       /* [1] or NULL */
  const float* label;      /* [B * label_stride] or NULL */
  int64_t label_stride;
  float* logit;            /* [B] tower logit alone (what y would hold) */
  float* prob;             /* [B] */
  float* loss_per_sample;  /* [B] */
  float* dlogit;           /* [B] */
} NrxTowerHead;
int nrx_tower_fwd_head(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B,
                       const NrxTowerHead* h_head, int flags, void* ws, size_t ws_bytes, nrx_stream_t stream);

/* fp32 rows [B, width] -> bf16 tile image [tile][width16/8][128][8] (the operand layout of the tower kernels). */
int nrx_tower_image_from_rows(const float* x, int64_t x_ld, int64_t B, int width, void* image, nrx_stream_t stream);

/* grad_x (+)= d/dx, grad_w[l] / grad_b[l] = d/dW_l, d/db_l (overwritten). */
int nrx_tower_bwd(const NrxTower* h_tower, const float* x, int64_t x_ld, int64_t B,
                  const float* grad_y, int64_t gy_ld,
                  float* grad_x, int64_t gx_ld, int accumulate_gx,
                  float* const* h_grad_w, float* const* h_grad_b,
                  void* ws, size_t ws_bytes, nrx_stream_t stream);

/* The two halves of nrx_tower_bwd as separate calls: dX chain (writes grad_x and the dz images into `ws`) and
 * dW / db (consumes the images).  A trainer can overlap other work that only needs grad_x with the dW half. */
int nrx_tower_bwd_dx(const NrxTower* h_tower, int64_t B, const float* grad_y, int64_t gy_ld,
                     float* grad_x, int64_t gx_ld, int accumulate_gx, void* ws, size_t ws_bytes,
                     nrx_stream_t stream);
int nrx_tower_bwd_dw(const NrxTower* h_tower, int64_t B, float* const* h_grad_w, float* const* h_grad_b,
                     void* ws, size_t ws_bytes, nrx_stream_t stream);

/* Layout of the bf16 tile images the training forward/backward keep in `ws` (tests, tooling):
 * image of layer l's input (width act_width[l]) and of dL/dz_l (width dz_width[l]), each stored as
 * [tile][width/8][128 rows][8] bf16 — the UMMA operand layout, see DESIGN.md. */
int nrx_tower_image_layout(const NrxTower* h_tower, int64_t B, int64_t* act_off, int32_t* act_width,
                           int64_t* dz_off, int32_t* dz_width);

/* DCN-v1 cross stack: x_{l+1} = x0 * (x_l . w_l) + b_l + x_l  (dcn_arch.py:14-30),
 * writes out[B, 2d] = cat[x, x_L] (dcn/model.py:29).  dots[L, B] keeps x_l . w_l. */
int nrx_dcn_cross_fwd(const float* x, int64_t ld, int64_t B, int d, int n_layers,
                      const float* const* h_w, const float* const* h_b,
                      float* out, int64_t out_ld, float* dots, nrx_stream_t stream);
/* Same forward, emitting the bf16 tile image of cat[x, x_L] (width 2d) for the tower; `out` may be NULL. */
int nrx_dcn_cross_fwd_img(const float* x, int64_t ld, int64_t B, int d, int n_layers,
                          const float* const* h_w, const float* const* h_b,
                          float* out, int64_t out_ld, void* image, nrx_stream_t stream);
int nrx_dcn_cross_bwd(const float* x, int64_t ld, int64_t B, int d, int n_layers,
                      const float* const* h_w, const float* const* h_b,
                      const float* grad_out, int64_t go_ld, const float* dots,
                      float* grad_x, int64_t gx_ld,
                      float* const* h_grad_w, float* const* h_grad_b,
                      void* ws, size_t ws_bytes, nrx_stream_t stream);
size_t nrx_dcn_cross_workspace_bytes(int64_t B, int d, int n_layers);

/* ---- K6: exact inner-product top-k -----------------------------------------
 * Replaces faiss.IndexFlatIP.add/search (recall/DSSM/model.py:209,249-251;
 * model_utils/TopKSearcher.py:34-47,73-77): out ordered by (score desc, id asc),
 * ids are corpus positions + id_base, padding (k > N) is id -1 / score -FLT_MAX. */
/* Index = the corpus packed once to bf16 tile images for the tensor-core scan (IndexFlatIP.add). */
size_t nrx_topk_index_bytes(int64_t N, int D);
int nrx_topk_index_build(const float* corpus, int64_t c_ld, int64_t N, int D, void* index, size_t index_bytes,
                         nrx_stream_t stream);
/* IndexFlatIP.search.  `corpus` (the same fp32 rows the index was built from) is read for the exact fp64
 * re-scoring of the candidates; status[q] (optional) = 1 when query q was served by the exact fallback scan. */
size_t nrx_topk_search_workspace_bytes(int64_t Q, int64_t N, int D, int k);
int nrx_topk_search(const void* index, const float* corpus, int64_t c_ld, int64_t N, int D,
                    const float* queries, int64_t q_ld, int64_t Q, int k, int64_t id_base,
                    float* out_scores, int64_t* out_ids, int32_t* status, void* ws, size_t ws_bytes,
                    nrx_stream_t stream);
/* Same search, additionally returning the fp64 ordering keys (out_scores64 [Q, k], nullable): what a merge of per-shard
 * lists must order by to equal one index bit for bit (two rows may differ by less than one fp32 ulp). */
int nrx_topk_search64(const void* index, const float* corpus, int64_t c_ld, int64_t N, int D,
                      const float* queries, int64_t q_ld, int64_t Q, int k, int64_t id_base,
                      float* out_scores, double* out_scores64, int64_t* out_ids, int32_t* status,
                      void* ws, size_t ws_bytes, nrx_stream_t stream);
/* Sharded search over NVLink peer memory (BASELINE config 4: corpus row-sharded across the GPUs of one node; the
 * reference holds the whole corpus in ONE faiss.IndexFlatIP and searches it on the CPU, recall/DSSM/model.py:249-251,
 * model_utils/TopKSearcher.py:34-47,73-77).  2-D:
 * every rank scans ITS shard for ALL queries (sample -> theta -> one filter scan, the threshold budget shared between the
 * shards) and ships, per query, its exactly re-scored candidates (<= k, fp64 score + global id) into the inbox of the
 * query's owner — rank q / ceil(Q / world) — by peer stores; after a flag barrier the owner merges the `world` lists of
 * its Q / world queries, proves completeness against every shard's bound and writes the result into EVERY rank's output
 * buffers; a query that cannot be proven is re-scanned exactly by its owner over all shards through peer memory.  No NCCL,
 * the per-query fixed work (final sort, merge) divides by `world`.  All buffers of NrxTopkPeer come from nrx_peer_alloc
 * and are mapped with nrx_peer_open; corpus shards are contiguous [n_rows, D] fp32.  Results: out_scores[rank] / out_ids[rank]. */
typedef struct NrxTopkPeer {
  int32_t rank, world;
  const float* corpus[NRX_MAX_PEERS];   /* fp32 shard of every rank */
  int64_t n_rows[NRX_MAX_PEERS];        /* rows per shard; global id = sum of the preceding shards + local row */
  void* inbox[NRX_MAX_PEERS];           /* nrx_topk_peer_inbox_bytes(Q, world, k) bytes on every rank */
  float* out_scores[NRX_MAX_PEERS];     /* [Q, k] on every rank */
  int64_t* out_ids[NRX_MAX_PEERS];      /* [Q, k] on every rank */
  uint32_t* sig[NRX_MAX_PEERS];         /* signal pads (NRX_PEER_SIG_WORDS u32), as for K7 */
  int32_t* status;                      /* optional trainer-style status word (bit 1: dead exchange) */
  uint32_t timeout_ms;
  uint32_t kprime;                      /* per-shard threshold rank; 0 = ceil((2k + 64) / world) + 24 */
} NrxTopkPeer;
size_t nrx_topk_peer_inbox_bytes(int64_t Q, int world, int k);
int nrx_topk_search_peer(const void* index, int64_t N_local, int D, const float* queries, int64_t q_ld, int64_t Q, int k,
                         const NrxTopkPeer* h_peer, int32_t* status /* [Q]: 1 = served by the exact scan (owner only) */,
                         void* ws, size_t ws_bytes, nrx_stream_t stream);
/* One-shot build + search (index lives in `ws`). */
size_t nrx_topk_ip_workspace_bytes(int64_t Q, int64_t N, int D, int k);
int nrx_topk_ip(const float* queries, int64_t q_ld, const float* corpus, int64_t c_ld,
                int64_t Q, int64_t N, int D, int k, int64_t id_base,
                float* out_scores, int64_t* out_ids, void* ws, size_t ws_bytes,
                nrx_stream_t stream);
/* Merge `n_lists` per-shard lists [n_lists][Q][k] into the global top-k. */
int nrx_topk_merge(const float* scores, const int64_t* ids, int n_lists, int64_t Q, int k,
                   float* out_scores, int64_t* out_ids, nrx_stream_t stream);
/* Merge of per-shard lists that carry the fp64 ordering keys of nrx_topk_search64. */
int nrx_topk_merge64(const double* scores64, const int64_t* ids, int n_lists, int64_t Q, int k,
                     float* out_scores, int64_t* out_ids, nrx_stream_t stream);
/* Row-wise L2 normalisation (faiss.normalize_L2 / F.normalize, DSSM/model.py:69-71). */
int nrx_l2_normalize(const float* x, int64_t ld, int64_t n, int d, float* y, int64_t y_ld,
                     nrx_stream_t stream);

/* DSSM training tail, fused (recall/DSSM/model.py:51-73 forward + :92-110 infoNCE_loss): from the RAW tower outputs
 * user [B, d] and item [B, d] and the negative-sampling permutations (h_perms: host array of n_neg DEVICE pointers to
 * int64 [B]; negative j of sample b is item perms[j][b]; every list must be a permutation of 0..B-1, as torch.randperm
 * yields) it computes  u^ = user / max(|user|, 1e-12), likewise the items,  logits [u^.p^, u^.n^_1 ..] / temperature,
 * loss_per_sample[b] = mask[b] * cross_entropy(logits_b, 0)  (mask nullable = ones; read at mask[b * mask_stride])
 * and — when the pointers are given — the gradients of  mean_b loss_per_sample[b]  with respect to the RAW user and item
 * rows (through the normalisations and the gather of the negatives; the item gradient is a deterministic gather over the
 * inverse permutations, no float atomics).  status (nullable): bit 0 is set when a list was not a permutation.
 * Two launches.  d <= 256, n_neg <= 7. */
size_t nrx_dssm_infonce_workspace_bytes(int64_t B, int n_neg);
int nrx_dssm_infonce(const float* user, int64_t u_ld, const float* item, int64_t i_ld, int64_t B, int d,
                     const int64_t* const* h_perms, int n_neg, const float* mask, int64_t mask_stride,
                     float temperature, float* loss_per_sample, float* grad_user, int64_t gu_ld, float* grad_item,
                     int64_t gi_ld, int32_t* status, void* ws, size_t ws_bytes, nrx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NRX_H_ */
