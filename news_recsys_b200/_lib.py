"""ctypes binding of libnrx.so (the C ABI declared in include/nrx.h).

There is no fallback: if the library is missing or a call fails, this raises.
torch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnrx.so")

NRX_MAX_FEATS = 16
NRX_MAX_TABLES = 16
NRX_MAX_LAYERS = 8
NRX_MAX_PEERS = 16
NRX_PEER_SIG_WORDS = 256

POOL_NONE, POOL_MASKED_MEAN, POOL_MEAN = 0, 1, 2
IDX_I64, IDX_I32 = 0, 1
BWD_DENSE, BWD_SGD, BWD_ADAMW = 0, 1, 2
BWD_NO_ZERO = 0x100
PLAN_ALL, PLAN_SORT, PLAN_MERGE = 0, 1, 2
FIELD_FM, FIELD_WIDE, FIELD_SUM = 0, 1, 2
ACT_RELU, ACT_LEAKY = 0, 1


class NrxError(RuntimeError):
    pass


class NrxFeat(C.Structure):
    _fields_ = [
        ("table", C.c_void_p), ("rows", C.c_int64), ("dim", C.c_int32), ("row_stride", C.c_int32),
        ("table_id", C.c_int32), ("idx_dtype", C.c_int32), ("idx", C.c_void_p), ("L", C.c_int32),
        ("pool", C.c_int32), ("mask", C.c_void_p), ("inv_den", C.c_void_p), ("out_col", C.c_int32),
        ("reserved", C.c_int32),
    ]


class NrxRowOpt(C.Structure):
    _fields_ = [
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("weight_decay", C.c_float), ("step", C.c_int32),
        ("m", C.c_void_p * NRX_MAX_TABLES), ("v", C.c_void_p * NRX_MAX_TABLES),
        ("d_hparams", C.c_void_p),
    ]


class NrxPeerStep(C.Structure):
    _fields_ = [
        ("rank", C.c_int32), ("world", C.c_int32),
        ("p", C.c_void_p * NRX_MAX_PEERS), ("g", C.c_void_p * NRX_MAX_PEERS), ("sig", C.c_void_p * NRX_MAX_PEERS),
        ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64), ("d_hparams", C.c_void_p),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("status", C.c_void_p), ("timeout_ms", C.c_uint32), ("reserved", C.c_uint32),
    ]


class NrxShardFeat(C.Structure):
    _fields_ = [("table", C.c_void_p), ("ids", C.c_void_p), ("lo", C.c_int64), ("hi", C.c_int64),
                ("dim", C.c_int32), ("row_stride", C.c_int32), ("out_col", C.c_int32), ("idx_dtype", C.c_int32)]


class NrxTopkPeer(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32),
                ("corpus", C.c_void_p * NRX_MAX_PEERS), ("n_rows", C.c_int64 * NRX_MAX_PEERS), ("inbox", C.c_void_p * NRX_MAX_PEERS),
                ("out_scores", C.c_void_p * NRX_MAX_PEERS), ("out_ids", C.c_void_p * NRX_MAX_PEERS), ("sig", C.c_void_p * NRX_MAX_PEERS),
                ("status", C.c_void_p), ("timeout_ms", C.c_uint32), ("kprime", C.c_uint32)]


class NrxIngestCol(C.Structure):
    _fields_ = [("data", C.c_void_p), ("offsets", C.c_void_p), ("L", C.c_int32), ("idx_dtype", C.c_int32),
                ("out_ids", C.c_void_p), ("out_mask", C.c_void_p)]


class NrxTower(C.Structure):
    _fields_ = [
        ("n_layers", C.c_int32), ("dims", C.c_int32 * (NRX_MAX_LAYERS + 1)),
        ("w", C.c_void_p * NRX_MAX_LAYERS), ("b", C.c_void_p * NRX_MAX_LAYERS),
        ("act", C.c_int32), ("negative_slope", C.c_float),
    ]


class NrxTowerHead(C.Structure):
    _fields_ = [
        ("terms", C.c_void_p * 4), ("n_terms", C.c_int32), ("reserved", C.c_int32),
        ("bias", C.c_void_p), ("label", C.c_void_p), ("label_stride", C.c_int64),
        ("logit", C.c_void_p), ("prob", C.c_void_p), ("loss_per_sample", C.c_void_p), ("dlogit", C.c_void_p),
    ]


TOWER_TRAINING, TOWER_PREPACKED, TOWER_XIMG = 1, 2, 4

_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); mirrors include/nrx.h one to one
SIGNATURES = {
    "nrx_version": (C.c_int, []),
    "nrx_last_error": (C.c_char_p, []),
    "nrx_embed_pool_fwd": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _I64, _P, _P]),
    "nrx_embed_pool_fwd_img": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _I64, _P, C.c_int, _P, _P, _P]),
    "nrx_embed_bwd_workspace_bytes": (_SZ, [C.POINTER(NrxFeat), C.c_int, _I64]),
    "nrx_embed_bwd_plan": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _SZ, _P]),
    "nrx_embed_bwd_plan_is_staged": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64]),
    "nrx_embed_bwd_plan_stage": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _SZ, C.c_int, _P]),
    "nrx_embed_bwd_apply": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _I64, C.c_int,
                                      C.POINTER(_P), C.POINTER(_P), C.POINTER(NrxRowOpt), _P, _SZ, _P]),
    "nrx_ingest_gather_ids": (C.c_int, [_P, _I64, _P, _I64, _I64, _P, C.c_int]),
    "nrx_ingest_csr_expand": (C.c_int, [_P, _P, _I64, _P, _I64, _I64, _I32, _P, C.c_int, _P]),
    "nrx_ingest_gather_labels": (C.c_int, [_P, _I64, _I32, _P, _I64, _I64, _P, _I32]),
    "nrx_ingest_assemble_device": (C.c_int, [C.POINTER(NrxIngestCol), C.c_int, _P, _I32, _P, _I32, _I64, _P, _I64, _I64, _P, _P]),
    "nrx_grouped_rank_metrics": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, _P]),
    "nrx_adamw_untouched_rows_scratch_bytes": (_SZ, [C.POINTER(NrxFeat), C.c_int]),
    "nrx_adamw_untouched_rows": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, C.POINTER(_P), C.POINTER(NrxRowOpt), _P, _SZ, _P]),
    "nrx_field_logit_fwd": (C.c_int, [_P, _I64, _I64, C.POINTER(_I32), C.POINTER(_I32), C.c_int, C.c_int, _P, C.c_int, _P]),
    "nrx_field_logit_bwd": (C.c_int, [_P, _I64, _I64, C.POINTER(_I32), C.POINTER(_I32), C.c_int, C.c_int, _P, _P, _I64, C.c_int, _P]),
    "nrx_fm_fused_fwd": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _P]),
    "nrx_fm_fused_bwd": (C.c_int, [C.POINTER(NrxFeat), C.c_int, _I64, _P, _P, _I64, _P]),
    "nrx_logit_loss_fwd": (C.c_int, [C.POINTER(_P), C.c_int, _P, _I64, _P, _I64, _P, _P, _P, _P]),
    "nrx_bce_fwd": (C.c_int, [_P, _P, _I64, _I64, _P, _P]),
    "nrx_bce_bwd": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P]),
    "nrx_sigmoid_bwd": (C.c_int, [_P, _P, _I64, _P, _P]),
    "nrx_reduce2_f32": (C.c_int, [_P, _I64, _F, _P, _P, _I64, _F, _P, _P]),
    "nrx_reduce_f32": (C.c_int, [_P, _I64, _F, _P, _P]),
    "nrx_adamw_dense": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I32, _P]),
    "nrx_adamw_dense_dev": (C.c_int, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _P]),
    "nrx_hparams_step": (C.c_int, [_P, _P, _F, _F, _I32, _I32, _F, _F, _P]),
    "nrx_peer_alloc": (C.c_int, [_SZ, C.POINTER(_P)]),
    "nrx_peer_free": (C.c_int, [_P]),
    "nrx_peer_export": (C.c_int, [_P, C.c_char_p]),
    "nrx_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "nrx_peer_close": (C.c_int, [_P]),
    "nrx_adamw_allreduce_peer": (C.c_int, [C.POINTER(NrxPeerStep), _P]),
    "nrx_peer_barrier": (C.c_int, [C.POINTER(NrxPeerStep), _P]),
    "nrx_shard_push": (C.c_int, [C.POINTER(NrxShardFeat), C.c_int, C.c_int, C.c_int, _I64, C.POINTER(_P), _I64, _P]),
    "nrx_shard_pull": (C.c_int, [C.POINTER(NrxShardFeat), C.c_int, C.POINTER(NrxShardFeat), C.c_int, C.c_int, C.c_int, _I64,
                                 C.POINTER(_P), _I64, _P, _P]),
    "nrx_peer_status": (C.c_int, [_P, C.POINTER(C.c_int32), _P]),
    "nrx_tower_workspace_bytes": (_SZ, [C.POINTER(NrxTower), _I64, C.c_int]),
    "nrx_tower_pack": (C.c_int, [C.POINTER(NrxTower), _I64, C.c_int, _P, _SZ, _P]),
    "nrx_tower_fwd": (C.c_int, [C.POINTER(NrxTower), _P, _I64, _I64, _P, _I64, C.c_int, _P, _SZ, _P]),
    "nrx_tower_fwd_head": (C.c_int, [C.POINTER(NrxTower), _P, _I64, _I64, C.POINTER(NrxTowerHead), C.c_int, _P, _SZ, _P]),
    "nrx_tower_image_from_rows": (C.c_int, [_P, _I64, _I64, C.c_int, _P, _P]),
    "nrx_tower_bwd": (C.c_int, [C.POINTER(NrxTower), _P, _I64, _I64, _P, _I64, _P, _I64, C.c_int,
                                C.POINTER(_P), C.POINTER(_P), _P, _SZ, _P]),
    "nrx_tower_bwd_dx": (C.c_int, [C.POINTER(NrxTower), _I64, _P, _I64, _P, _I64, C.c_int, _P, _SZ, _P]),
    "nrx_tower_bwd_dw": (C.c_int, [C.POINTER(NrxTower), _I64, C.POINTER(_P), C.POINTER(_P), _P, _SZ, _P]),
    "nrx_tower_image_layout": (C.c_int, [C.POINTER(NrxTower), _I64, C.POINTER(_I64), C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I32)]),
    "nrx_dcn_cross_fwd": (C.c_int, [_P, _I64, _I64, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P), _P, _I64, _P, _P]),
    "nrx_dcn_cross_fwd_img": (C.c_int, [_P, _I64, _I64, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P), _P, _I64, _P, _P]),
    "nrx_dcn_cross_bwd": (C.c_int, [_P, _I64, _I64, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P), _P, _I64, _P,
                                    _P, _I64, C.POINTER(_P), C.POINTER(_P), _P, _SZ, _P]),
    "nrx_dcn_cross_workspace_bytes": (_SZ, [_I64, C.c_int, C.c_int]),
    "nrx_topk_index_bytes": (_SZ, [_I64, C.c_int]),
    "nrx_topk_index_build": (C.c_int, [_P, _I64, _I64, C.c_int, _P, _SZ, _P]),
    "nrx_topk_search_workspace_bytes": (_SZ, [_I64, _I64, C.c_int, C.c_int]),
    "nrx_topk_search": (C.c_int, [_P, _P, _I64, _I64, C.c_int, _P, _I64, _I64, C.c_int, _I64, _P, _P, _P, _P, _SZ, _P]),
    "nrx_topk_search64": (C.c_int, [_P, _P, _I64, _I64, C.c_int, _P, _I64, _I64, C.c_int, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "nrx_topk_peer_inbox_bytes": (_SZ, [_I64, C.c_int, C.c_int]),
    "nrx_topk_search_peer": (C.c_int, [_P, _I64, C.c_int, _P, _I64, _I64, C.c_int, C.POINTER(NrxTopkPeer), _P, _P, _SZ, _P]),
    "nrx_topk_merge64": (C.c_int, [_P, _P, C.c_int, _I64, C.c_int, _P, _P, _P]),
    "nrx_topk_ip_workspace_bytes": (_SZ, [_I64, _I64, C.c_int, C.c_int]),
    "nrx_topk_ip": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, C.c_int, C.c_int, _I64, _P, _P, _P, _SZ, _P]),
    "nrx_topk_merge": (C.c_int, [_P, _P, C.c_int, _I64, C.c_int, _P, _P, _P]),
    "nrx_dssm_infonce_workspace_bytes": (_SZ, [_I64, C.c_int]),
    "nrx_dssm_infonce": (C.c_int, [_P, _I64, _P, _I64, _I64, C.c_int, _P, C.c_int, _P, _I64, C.c_float, _P, _P, _I64, _P, _I64, _P, _P, _SZ, _P]),
    "nrx_l2_normalize": (C.c_int, [_P, _I64, _I64, C.c_int, _P, _I64, _P]),
}

_lock = threading.Lock()
_lib = None
MISSING = []  # header symbols the loaded library does not export


def load() -> C.CDLL:
    """Load libnrx.so (once).  Raises NrxError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NrxError(
                f"{LIB_PATH} is missing — build it with `python -m news_recsys_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError:
                MISSING.append(name)  # tests/test_abi.py requires this list to be empty
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


# kernels of OURS launched per successful API call (library kernels such as the CUB radix sort are not counted)
KERNELS_PER_CALL = {"nrx_tower_fwd": 3, "nrx_tower_fwd(prepacked)": 2, "nrx_tower_fwd_head": 3, "nrx_tower_fwd_head(prepacked)": 2,
                    "nrx_tower_fwd(prepacked,ximg)": 1, "nrx_tower_fwd_head(prepacked,ximg)": 1, "nrx_tower_bwd": 3, "nrx_tower_bwd_dw": 2, "nrx_embed_bwd_apply": 2, "nrx_embed_bwd_plan": 2, "nrx_embed_bwd_plan(sort)": 1, "nrx_embed_bwd_plan(merge)": 1, "nrx_embed_bwd_plan_is_staged": 0, "nrx_dssm_infonce": 2, "nrx_dssm_infonce_workspace_bytes": 0, "nrx_adamw_untouched_rows": 2, "nrx_adamw_untouched_rows_scratch_bytes": 0, "nrx_dcn_cross_bwd": 2,
                    "nrx_embed_bwd_apply(dense)": 2, "nrx_embed_bwd_apply(rowopt)": 2, "nrx_topk_ip": 2,
                    "nrx_topk_search": 7, "nrx_topk_search64": 7, "nrx_topk_search_peer": 11, "nrx_topk_peer_inbox_bytes": 0, "nrx_topk_index_build": 1,
                    "nrx_peer_alloc": 0, "nrx_peer_free": 0, "nrx_peer_export": 0, "nrx_peer_open": 0, "nrx_peer_close": 0,
                    "nrx_peer_status": 0, "nrx_ingest_gather_ids": 0, "nrx_ingest_csr_expand": 0, "nrx_ingest_gather_labels": 0,
                    "nrx_tower_workspace_bytes": 0, "nrx_embed_bwd_workspace_bytes": 0, "nrx_tower_image_layout": 0}
launch_count = 0  # running total, read by bench.py


def check(rc: int, what: str) -> None:
    global launch_count
    launch_count += KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = load().nrx_last_error()
        raise NrxError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def ptr_array(ts, n=None):
    """Host array of device pointers (void*[n])."""
    n = len(ts) if n is None else n
    arr = (_P * n)()
    for i, t in enumerate(ts):
        arr[i] = 0 if t is None else (t if isinstance(t, int) else t.data_ptr())
    return arr


def i32_array(xs):
    return (_I32 * len(xs))(*[int(x) for x in xs])
