"""ctypes view of include/epp_engine.h (the C ABI of libepp_engine.so).  No torch, no numpy types in the ABI:
plain pointers and sizes.  Loading fails loudly when the library has not been built; creating an engine fails
loudly (EPP_ERR_NO_DEVICE) when there is no CUDA device -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libepp_engine.so")

EPP_MAX_SCORERS = 8
EPP_NO_ENDPOINT = 0xFFFFFFFF
EPP_BATCH_DEVICE_PTRS = 1
EPP_BATCH_ASYNC = 2
EPP_BATCH_LENGTHS_EXCEED_ROWS = 4

EPP_OK, EPP_ERR_INVALID, EPP_ERR_CUDA, EPP_ERR_NO_DEVICE, EPP_ERR_CAPACITY, EPP_ERR_STATE, EPP_ERR_NCCL = \
    0, -1, -2, -3, -4, -5, -6
(SCORER_PREFIX, SCORER_KV_UTIL, SCORER_QUEUE, SCORER_LOAD_AWARE, SCORER_EXTERNAL, SCORER_RUNNING, SCORER_TOKEN_LOAD,
 SCORER_ACTIVE_REQUEST, SCORER_LORA_AFFINITY) = range(9)
(ROLE_NONE, ROLE_DECODE, ROLE_PREFILL, ROLE_PREFILL_DECODE, ROLE_BOTH, ROLE_ENCODE, ROLE_ENCODE_PREFILL,
 ROLE_ENCODE_PREFILL_DECODE, ROLE_OTHER) = range(9)
FILTER_NONE, FILTER_DECODE, FILTER_PREFILL, FILTER_ENCODE = range(4)
HANDLER_SINGLE, HANDLER_DISAGG = 0, 1


class ScorerCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("column", C.c_int32), ("weight", C.c_double), ("param", C.c_double),
                ("param2", C.c_double)]


class ProfileCfg(C.Structure):
    _fields_ = [("filter", C.c_int32), ("n_scorers", C.c_int32), ("scorers", ScorerCfg * EPP_MAX_SCORERS),
                ("affinity_threshold", C.c_double), ("exploration_probability", C.c_double),
                ("max_ttft_penalty_ms", C.c_double), ("ttft_column", C.c_int32), ("reserved", C.c_int32)]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("max_endpoints", C.c_int32),
                ("block_size_tokens", C.c_int32), ("max_prefix_blocks", C.c_int32),
                ("lru_capacity_per_server", C.c_int32), ("handler", C.c_int32), ("always_disagg", C.c_int32),
                ("non_cached_tokens", C.c_int64), ("n_ext_cols", C.c_int32), ("pick_k", C.c_int32),
                ("primary", ProfileCfg), ("prefill", ProfileCfg), ("encode", ProfileCfg), ("tie_seed", C.c_uint64),
                ("encode_enabled", C.c_int32), ("index_commit_interval_us", C.c_int32), ("reserved1", C.c_uint64 * 2)]


class Decision(C.Structure):
    _fields_ = [("status", C.c_int32), ("pick", C.c_uint32), ("score", C.c_double), ("prefill_pick", C.c_uint32),
                ("tie_count", C.c_uint32), ("total_blocks", C.c_int32), ("match_blocks", C.c_int32)]


class DecisionDetail(C.Structure):
    _fields_ = [("prefill_score", C.c_double), ("prefill_tie_count", C.c_uint32), ("prefill_ran", C.c_uint32),
                ("encode_score", C.c_double), ("encode_pick", C.c_uint32), ("encode_tie_count", C.c_uint32),
                ("encode_ran", C.c_uint32), ("reserved", C.c_uint32)]


class Batch(C.Structure):
    _fields_ = [("n_requests", C.c_int64), ("data", C.c_void_p), ("offsets", C.c_void_p),
                ("uniform_len", C.c_uint64), ("model_ids", C.c_void_p), ("flags", C.c_uint32),
                ("reserved", C.c_uint32), ("lengths", C.c_void_p), ("multimodal", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("n_batches", C.c_uint64), ("n_decisions", C.c_uint64), ("index_pairs", C.c_uint64),
                ("index_hashes", C.c_uint64), ("index_slots", C.c_uint64), ("last_h2d_ms", C.c_double),
                ("last_kernels_ms", C.c_double), ("last_d2h_ms", C.c_double), ("last_hash_ms", C.c_double),
                ("last_match_pick_ms", C.c_double), ("last_kernel_launches", C.c_uint64),
                ("device_bytes", C.c_uint64), ("last_kernel_ms", C.c_double * 8), ("last_probes", C.c_uint64),
                ("last_postings", C.c_uint64), ("last_index_apply_ms", C.c_double), ("last_index_build_ms", C.c_double),
                ("last_index_items", C.c_uint64), ("last_index_launches", C.c_uint64), ("last_index_patched", C.c_uint64)]


class TopkOut(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("k", C.c_int32), ("primary", C.c_void_p), ("primary_scores", C.c_void_p),
                ("prefill", C.c_void_p), ("encode", C.c_void_p)]


class BatcherCfg(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("max_batch", C.c_int32), ("max_delay_us", C.c_int32),
                ("index_picks", C.c_int32)]


class BatcherStats(C.Structure):
    _fields_ = [("n_flushes", C.c_uint64), ("n_requests", C.c_uint64), ("n_full_flushes", C.c_uint64),
                ("n_pending", C.c_uint64), ("n_index_errors", C.c_uint64)]


class ShardBest(C.Structure):
    _fields_ = [("score", C.c_double), ("pick", C.c_uint32), ("tie_count", C.c_uint32),
                ("match_blocks", C.c_int32), ("status", C.c_int32)]


assert C.sizeof(Decision) == 32 and C.sizeof(DecisionDetail) == 40 and C.sizeof(ShardBest) == 24

# name -> (restype, argtypes): every symbol include/epp_engine.h declares
SIGNATURES = {
    "epp_abi_version": (C.c_int32, []),
    "epp_last_error": (C.c_char_p, []),
    "epp_engine_create": (C.c_int32, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "epp_engine_destroy": (C.c_int32, [C.c_void_p]),
    "epp_config_default": (None, [C.POINTER(Config)]),
    "epp_host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "epp_host_free": (C.c_int32, [C.c_void_p]),
    "epp_model_register": (C.c_int32, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                       C.POINTER(C.c_uint32)]),
    "epp_model_seed": (C.c_int32, [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint64)]),
    "epp_pool_set": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]),
    "epp_index_add": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_int32, C.c_void_p, C.c_int32]),
    "epp_index_remove_endpoint": (C.c_int32, [C.c_void_p, C.c_uint32]),
    "epp_index_retain_endpoints": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "epp_index_load_snapshot": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]),
    "epp_index_commit": (C.c_int32, [C.c_void_p]),
    "epp_index_get": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]),
    "epp_hash_prompts": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    "epp_prefix_match": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p]),
    "epp_pool_set_lora": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "epp_score": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                              C.c_uint32]),
    "epp_schedule": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_int32]),
    "epp_schedule_with_match": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_void_p, C.c_void_p, C.c_uint32]),
    "epp_schedule_topk": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(TopkOut)]),
    "epp_schedule_with_match_topk": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(TopkOut)]),
    "epp_index_add_picked": (C.c_int32, [C.c_void_p]),
    "epp_get_stats": (C.c_int32, [C.c_void_p, C.POINTER(Stats)]),
    "epp_synchronize": (C.c_int32, [C.c_void_p]),
    "epp_event_record": (C.c_int32, [C.c_void_p, C.c_int32]),
    "epp_event_elapsed_ms": (C.c_int32, [C.c_void_p, C.POINTER(C.c_double)]),
    "epp_shard_set": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_uint32]),
    "epp_shard_probe": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    "epp_shard_pick": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "epp_shard_merge": (C.c_int32, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "epp_shard_p2p_export": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_uint64)]),
    "epp_shard_p2p_connect": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32]),
    "epp_shard_schedule_p2p": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p]),
    "epp_shard_p2p_phase": (C.c_int32, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_int32]),
    "epp_get_config": (C.c_int32, [C.c_void_p, C.POINTER(Config)]),
    "epp_batcher_create": (C.c_int32, [C.c_void_p, C.POINTER(BatcherCfg), C.POINTER(C.c_void_p)]),
    "epp_batcher_destroy": (C.c_int32, [C.c_void_p]),
    "epp_submit": (C.c_int32, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "epp_wait": (C.c_int32, [C.c_void_p, C.c_uint64, C.POINTER(Decision), C.POINTER(DecisionDetail)]),
    "epp_wait_topk": (C.c_int32, [C.c_void_p, C.c_uint64, C.POINTER(Decision), C.POINTER(DecisionDetail), C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "epp_batcher_stats": (C.c_int32, [C.c_void_p, C.POINTER(BatcherStats)]),
    "epp_batcher_last_error": (C.c_char_p, []),
}

_lib = None


def load():
    """dlopen libepp_engine.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is not built: run `python llm-d-inference-scheduler_b200/build.py` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
