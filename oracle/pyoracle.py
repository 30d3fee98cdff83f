"""ctypes binding for the CPU ORACLE (oracle/epp_oracle.c).

TEST INFRASTRUCTURE ONLY -- may be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (llm-d-inference-scheduler_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libepp_oracle.so")

MAX_SCORERS = 8
(SCORER_PREFIX, SCORER_KV_UTIL, SCORER_QUEUE, SCORER_LOAD_AWARE, SCORER_EXTERNAL, SCORER_RUNNING, SCORER_TOKEN_LOAD,
 SCORER_ACTIVE_REQUEST, SCORER_LORA_AFFINITY) = range(9)
ROLE_NONE, ROLE_DECODE, ROLE_PREFILL, ROLE_PREFILL_DECODE, ROLE_BOTH, ROLE_ENCODE, ROLE_ENCODE_PREFILL, \
    ROLE_ENCODE_PREFILL_DECODE, ROLE_OTHER = range(9)
ROLE_ABSENT = 0xFF
FILTER_NONE, FILTER_DECODE, FILTER_PREFILL, FILTER_ENCODE = range(4)


class Scorer(C.Structure):
    _fields_ = [("kind", C.c_int32), ("column", C.c_int32), ("weight", C.c_double), ("param", C.c_double),
                ("param2", C.c_double)]


class Profile(C.Structure):
    _fields_ = [("filter", C.c_int32), ("n_scorers", C.c_int32), ("scorers", Scorer * MAX_SCORERS),
                ("affinity_threshold", C.c_double), ("exploration_probability", C.c_double),
                ("max_ttft_penalty_ms", C.c_double), ("ttft_column", C.c_int32), ("_pad", C.c_int32)]


class Pool(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_ext_cols", C.c_int32), ("role", C.c_void_p), ("kv_usage", C.c_void_p),
                ("waiting", C.c_void_p), ("running", C.c_void_p), ("ext", C.c_void_p), ("lora_state", C.c_void_p),
                ("lora_max", C.c_void_p), ("lora_loaded", C.c_void_p)]


class Decision(C.Structure):
    _fields_ = [("status", C.c_int32), ("pick", C.c_int32), ("tie_count", C.c_int32),
                ("prefill_pick", C.c_int32), ("prefill_tie_count", C.c_int32), ("prefill_ran", C.c_int32),
                ("score", C.c_double), ("prefill_score", C.c_double), ("encode_pick", C.c_int32),
                ("encode_tie_count", C.c_int32), ("encode_ran", C.c_int32), ("_pad", C.c_int32),
                ("encode_score", C.c_double)]


class CycleCfg(C.Structure):
    _fields_ = [("block_size_tokens", C.c_int32), ("max_prefix_blocks", C.c_int32),
                ("non_cached_tokens", C.c_int64), ("always_disagg", C.c_int32), ("_pad", C.c_int32),
                ("model", C.c_char_p), ("model_len", C.c_size_t), ("tie_seed", C.c_uint64), ("tie_base", C.c_uint64),
                ("encode", C.POINTER(Profile)), ("multimodal", C.c_void_p), ("topk", C.c_int32), ("_pad2", C.c_int32),
                ("topk_primary", C.c_void_p), ("topk_primary_scores", C.c_void_p), ("topk_prefill", C.c_void_p),
                ("topk_encode", C.c_void_p)]


DECISION_DTYPE = np.dtype([("status", "<i4"), ("pick", "<i4"), ("tie_count", "<i4"), ("prefill_pick", "<i4"),
                           ("prefill_tie_count", "<i4"), ("prefill_ran", "<i4"), ("score", "<f8"),
                           ("prefill_score", "<f8"), ("encode_pick", "<i4"), ("encode_tie_count", "<i4"),
                           ("encode_ran", "<i4"), ("_pad", "<i4"), ("encode_score", "<f8")])
assert DECISION_DTYPE.itemsize == C.sizeof(Decision)


def build(force: bool = False) -> str:
    """Compile oracle/epp_oracle.c with gcc (Makefile in this directory)."""
    src = os.path.join(_HERE, "epp_oracle.c")
    hdr = os.path.join(_HERE, "epp_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_xxh64.restype = C.c_uint64
        L.orc_xxh64.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64]
        L.orc_hash_prompt.restype = C.c_int
        L.orc_hash_prompt.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_indexer_new.restype = C.c_void_p
        L.orc_indexer_new.argtypes = [C.c_int]
        L.orc_indexer_free.argtypes = [C.c_void_p]
        L.orc_indexer_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_int]
        L.orc_indexer_get.restype = C.c_int
        L.orc_indexer_get.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        L.orc_indexer_remove_pod.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_indexer_pods.restype = C.c_int
        L.orc_indexer_pods.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.orc_indexer_lru_len.restype = C.c_int
        L.orc_indexer_lru_len.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_indexer_export.restype = C.c_size_t
        L.orc_indexer_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_indexer_load_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_match_longest_prefix.restype = C.c_int
        L.orc_match_longest_prefix.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_role_filter_keeps.restype = C.c_int
        L.orc_role_filter_keeps.argtypes = [C.c_int, C.c_int]
        L.orc_profile_run.restype = C.c_int
        L.orc_profile_run.argtypes = [C.POINTER(Profile), C.POINTER(Pool), C.c_void_p, C.c_int32, C.c_void_p,
                                      C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p]
        L.orc_score_column.argtypes = [C.POINTER(Scorer), C.POINTER(Pool), C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_void_p]
        L.orc_pd_decide.restype = C.c_int
        L.orc_pd_decide.argtypes = [C.c_int64, C.c_int64, C.c_int32, C.c_int32]
        L.orc_schedule.argtypes = [C.POINTER(Profile), C.POINTER(Profile), C.POINTER(Pool), C.c_void_p, C.c_int32,
                                   C.c_int32, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.POINTER(Decision)]
        L.orc_tie_rank.restype = C.c_uint32
        L.orc_tie_rank.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.orc_profile_run_topk.restype = C.c_int
        L.orc_profile_run_topk.argtypes = [C.POINTER(Profile), C.POINTER(Pool), C.c_void_p, C.c_int32, C.c_uint64,
                                           C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]
        L.orc_explore_u.restype = C.c_double
        L.orc_explore_u.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_cycle_batch.argtypes = [C.POINTER(CycleCfg), C.c_void_p, C.POINTER(Profile), C.POINTER(Profile),
                                      C.POINTER(Pool), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                      C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def xxh64(data: bytes, seed: int = 0) -> int:
    buf = C.create_string_buffer(data, len(data)) if data else None
    return int(lib().orc_xxh64(buf, len(data), seed))


def hash_prompt(data: bytes, model: bytes, block_size_tokens: int, max_prefix_blocks: int,
                salt: bytes = b"") -> list[int]:
    """hashPrompt (approximateprefix/hashing.go:35-99)."""
    cap = max(1, len(data) // max(1, block_size_tokens * 4) + 2)
    out = np.zeros(cap, dtype=np.uint64)
    d = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, np.uint8)
    n = lib().orc_hash_prompt(_ptr(d), len(data), model, len(model), salt, len(salt),
                              block_size_tokens, max_prefix_blocks, _ptr(out), cap)
    return [int(x) for x in out[:n]]


def make_profile(filter_kind: int, scorers: list[tuple], affinity: tuple | None = None) -> Profile:
    """affinity = (affinityThreshold, explorationProbability, maxTTFTPenaltyMs[, ttft ext column]) of a
    prefix-cache-affinity-filter placed after the role filter, or None."""
    p = Profile()
    p.filter = filter_kind
    p.n_scorers = len(scorers)
    p.ttft_column = -1
    if affinity is not None:
        p.affinity_threshold, p.exploration_probability, p.max_ttft_penalty_ms = affinity[:3]
        p.ttft_column = int(affinity[3]) if len(affinity) > 3 else -1
    for i, sc in enumerate(scorers):              # (kind, weight, param[, column[, param2]])
        kind, weight, param = sc[:3]
        p.scorers[i].kind = kind
        p.scorers[i].weight = weight
        p.scorers[i].param = param
        p.scorers[i].column = int(sc[3]) if len(sc) > 3 else 0
        p.scorers[i].param2 = float(sc[4]) if len(sc) > 4 else 0.0
    return p


class PoolState:
    """Struct-of-arrays pool snapshot; keeps the numpy arrays alive for the ctypes struct."""

    def __init__(self, role, kv_usage, waiting, running=None, ext=None):
        self.role = np.ascontiguousarray(role, dtype=np.uint8)
        n = self.role.shape[0]
        self.kv_usage = np.ascontiguousarray(kv_usage, dtype=np.float64)
        self.waiting = np.ascontiguousarray(waiting, dtype=np.int32)
        self.running = np.ascontiguousarray(running if running is not None else np.zeros(n), dtype=np.int32)
        self.ext = None if ext is None else np.ascontiguousarray(ext, dtype=np.float64).reshape(-1, n)
        self.n = n
        self.c = Pool()
        self.c.n = n
        self.c.n_ext_cols = 0 if self.ext is None else self.ext.shape[0]
        self.c.role = self.role.ctypes.data
        self.c.kv_usage = self.kv_usage.ctypes.data
        self.c.waiting = self.waiting.ctypes.data
        self.c.running = self.running.ctypes.data
        self.c.ext = 0 if self.ext is None else self.ext.ctypes.data
        self.c.lora_state = self.c.lora_max = self.c.lora_loaded = 0

    def set_lora(self, state, max_active, loaded):
        """LoRA residency as seen by a request for ONE adapter: state[e] in {0, 1 active, 2 waiting}."""
        self.lora_state = np.ascontiguousarray(state, dtype=np.uint8)
        self.lora_max = np.ascontiguousarray(max_active, dtype=np.int32)
        self.lora_loaded = np.ascontiguousarray(loaded, dtype=np.int32)
        assert self.lora_state.shape[0] == self.lora_max.shape[0] == self.lora_loaded.shape[0] == self.n
        self.c.lora_state = self.lora_state.ctypes.data
        self.c.lora_max = self.lora_max.ctypes.data
        self.c.lora_loaded = self.lora_loaded.ctypes.data
        return self


class Indexer:
    """indexer (approximateprefix/indexer.go)."""

    def __init__(self, default_lru_size: int = 31250):
        self.h = lib().orc_indexer_new(default_lru_size)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_indexer_free(self.h)
            self.h = None

    def add(self, hashes, server: int, num_gpu_blocks: int = 0):
        a = np.ascontiguousarray(hashes, dtype=np.uint64)
        lib().orc_indexer_add(self.h, _ptr(a), a.shape[0], server, num_gpu_blocks)

    def get(self, h: int) -> set[int]:
        out = np.zeros(64, dtype=np.uint32)
        n = lib().orc_indexer_get(self.h, h, _ptr(out), out.shape[0])
        if n > out.shape[0]:
            out = np.zeros(n, dtype=np.uint32)
            n = lib().orc_indexer_get(self.h, h, _ptr(out), out.shape[0])
        return set(int(x) for x in out[:n])

    def remove_pod(self, server: int):
        lib().orc_indexer_remove_pod(self.h, server)

    def pods(self) -> list[int]:
        out = np.zeros(4096, dtype=np.uint32)
        n = lib().orc_indexer_pods(self.h, _ptr(out), out.shape[0])
        return sorted(int(x) for x in out[:n])

    def lru_len(self, server: int) -> int:
        return lib().orc_indexer_lru_len(self.h, server)

    def export(self):
        n = lib().orc_indexer_export(self.h, None, None, 0)
        hs = np.zeros(max(n, 1), dtype=np.uint64)
        sv = np.zeros(max(n, 1), dtype=np.uint32)
        lib().orc_indexer_export(self.h, _ptr(hs), _ptr(sv), n)
        return hs[:n], sv[:n]

    def load_pairs(self, hashes, servers):
        hs = np.ascontiguousarray(hashes, dtype=np.uint64)
        sv = np.ascontiguousarray(servers, dtype=np.uint32)
        assert hs.shape == sv.shape
        lib().orc_indexer_load_pairs(self.h, _ptr(hs), _ptr(sv), hs.shape[0])

    def match_longest_prefix(self, hashes, n_servers: int):
        """matchLongestPrefix (plugin.go:214-230) -> (dense counts[n_servers], blocks walked)."""
        a = np.ascontiguousarray(hashes, dtype=np.uint64)
        counts = np.zeros(max(n_servers, 1), dtype=np.int32)
        walked = lib().orc_match_longest_prefix(self.h, _ptr(a), a.shape[0], _ptr(counts), n_servers)
        return counts[:n_servers], walked


def profile_run(profile: Profile, pool: PoolState, match, total: int):
    """SchedulerProfile.Run -> (scores[E] (-1 = filtered out), max, lowest-index pick, argmax set)."""
    m = np.ascontiguousarray(match, dtype=np.int32)
    scores = np.zeros(pool.n, dtype=np.float64)
    amax = np.zeros(max(pool.n, 1), dtype=np.int32)
    mx = C.c_double(0)
    pick = C.c_int32(-1)
    cnt = lib().orc_profile_run(C.byref(profile), C.byref(pool.c), _ptr(m), total, _ptr(scores), C.byref(mx),
                                C.byref(pick), _ptr(amax))
    return scores, mx.value, pick.value, [int(x) for x in amax[:cnt]]


def score_column(scorer: tuple, pool: PoolState, cand, match, total: int):
    s = Scorer()
    s.kind, s.weight, s.param = scorer[:3]
    s.column = int(scorer[3]) if len(scorer) > 3 else 0
    s.param2 = float(scorer[4]) if len(scorer) > 4 else 0.0
    c = np.ascontiguousarray(cand, dtype=np.uint8)
    m = np.ascontiguousarray(match, dtype=np.int32)
    out = np.zeros(pool.n, dtype=np.float64)
    lib().orc_score_column(C.byref(s), C.byref(pool.c), _ptr(c), _ptr(m), total, _ptr(out))
    return out


def profile_run_topk(profile: Profile, pool: PoolState, match, total: int, k: int, tie_seed: int = 0, tie_key: int = 0):
    """SchedulerProfile.Run with the build's reproducible draws -> (picks list (<= k), their scores, arg-max set size,
    scores[E] with -1 for endpoints the role filter dropped)."""
    m = np.ascontiguousarray(match, dtype=np.int32)
    picks = np.full(max(k, 1), -1, dtype=np.int32)
    sc = np.zeros(max(k, 1), dtype=np.float64)
    scores = np.zeros(pool.n, dtype=np.float64)
    n = C.c_int32(0)
    ties = lib().orc_profile_run_topk(C.byref(profile), C.byref(pool.c), _ptr(m), total, tie_seed, tie_key, k, _ptr(picks),
                                      _ptr(sc), C.byref(n), _ptr(scores))
    return [int(x) for x in picks[:n.value]], [float(x) for x in sc[:n.value]], ties, scores


def explore_u(seed: int, key: int) -> float:
    return float(lib().orc_explore_u(seed, key))


def tie_rank(seed: int, key: int, n: int) -> int:
    """Rank of the arg-max-set member picked under the build's reproducible tie rule (oracle/epp_oracle.h)."""
    return int(lib().orc_tie_rank(seed, key, n))


def pd_decide(nct: int, input_len_bytes: int, match_blocks: int, block_size_tokens: int) -> bool:
    return bool(lib().orc_pd_decide(nct, input_len_bytes, match_blocks, block_size_tokens))


def schedule(primary: Profile, prefill: Profile | None, pool: PoolState, match, total: int,
             block_size_tokens: int, input_len_bytes: int, nct: int, always_disagg: bool = False) -> Decision:
    m = np.ascontiguousarray(match, dtype=np.int32)
    scratch = np.zeros(max(pool.n, 1), dtype=np.float64)
    d = Decision()
    lib().orc_schedule(C.byref(primary), C.byref(prefill) if prefill is not None else None, C.byref(pool.c),
                       _ptr(m), total, block_size_tokens, input_len_bytes, nct, int(always_disagg), _ptr(scratch),
                       C.byref(d))
    return d


def cycle_batch(model: bytes, block_size_tokens: int, max_prefix_blocks: int, nct: int, always_disagg: bool,
                indexer: Indexer, primary: Profile, prefill: Profile | None, pool: PoolState,
                data: np.ndarray, offsets: np.ndarray, n_threads: int = 1, tie_seed: int = 0, tie_base: int = 0,
                encode: Profile | None = None, multimodal=None, topk: int = 0):
    """Whole cycle (hash -> match -> schedule) for R prompts against a frozen index.
    Returns (decisions structured array [R], totals int32[R]); with topk > 1 also a dict of the pickers' first-k lists
    {"primary": int32[R][k] (-1 padded), "primary_scores": f64[R][k], "prefill": ..., "encode": ...}."""
    cfg = CycleCfg()
    cfg.block_size_tokens = block_size_tokens
    cfg.max_prefix_blocks = max_prefix_blocks
    cfg.non_cached_tokens = nct
    cfg.always_disagg = int(always_disagg)
    cfg.model = model
    cfg.model_len = len(model)
    cfg.tie_seed = tie_seed
    cfg.tie_base = tie_base
    mm = None
    if encode is not None:
        cfg.encode = C.pointer(encode)
        if multimodal is not None:
            mm = np.ascontiguousarray(multimodal, dtype=np.uint8)
            cfg.multimodal = mm.ctypes.data
    data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    R = offsets.shape[0] - 1
    out = np.zeros(R, dtype=DECISION_DTYPE)
    totals = np.zeros(R, dtype=np.int32)
    lists = None
    if topk > 1:
        cfg.topk = topk
        lists = {"primary": np.full((R, topk), -1, np.int32), "primary_scores": np.zeros((R, topk), np.float64),
                 "prefill": np.full((R, topk), -1, np.int32), "encode": np.full((R, topk), -1, np.int32)}
        cfg.topk_primary = lists["primary"].ctypes.data
        cfg.topk_primary_scores = lists["primary_scores"].ctypes.data
        cfg.topk_prefill = lists["prefill"].ctypes.data
        cfg.topk_encode = lists["encode"].ctypes.data
    lib().orc_cycle_batch(C.byref(cfg), indexer.h, C.byref(primary),
                          C.byref(prefill) if prefill is not None else None, C.byref(pool.c), _ptr(data),
                          _ptr(offsets), R, n_threads, _ptr(out), _ptr(totals))
    if lists is not None:
        return out, totals, lists
    return out, totals
