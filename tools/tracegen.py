"""Python wrapper of tools/tracegen.c: the seeded synthetic workload of SURVEY.md 8(d), shared by tests and bench.

It generates INPUTS only (tokens, pool metrics, which endpoint caches which family to which depth).  The
(hash, endpoint) pairs of the index snapshot need the family prompts' block hashes; the caller supplies them
from whichever hasher it is exercising (the CUDA engine in bench.py's GPU arm, the oracle in its CPU arm).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libtracegen.so")
SRC = os.path.join(HERE, "tracegen.c")
MODEL = b"synthetic-model"
BASE_SEED = 0x5EEDE990


class _Cfg(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("E", C.c_int32), ("T", C.c_int32), ("bst", C.c_int32), ("G", C.c_int32),
                ("R", C.c_int64), ("n_prefill", C.c_int32), ("_pad", C.c_int32)]


def build(force: bool = False) -> str:
    if force or not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", SO, SRC, "-lm"])
    return SO


_lib = None


def _L():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.tg_families.restype = C.c_int32
        L.tg_families.argtypes = [C.POINTER(_Cfg)]
        L.tg_family_tokens.argtypes = [C.POINTER(_Cfg), C.c_void_p]
        L.tg_requests.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int]
        L.tg_pool.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tg_index_plan.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


# scorer tuples: (kind, weight, param) with kinds of include/epp_engine.h
PREFIX, KV_UTIL, QUEUE, LOAD_AWARE = 0, 1, 2, 3
FILTER_NONE, FILTER_DECODE, FILTER_PREFILL = 0, 1, 2


@dataclass
class Workload:
    """One BASELINE.json config (or a scaled-down copy of it for tests)."""
    name: str
    config_index: int
    E: int
    T: int
    R: int
    block_size_tokens: int = 16
    max_prefix_blocks: int = 256
    primary_filter: int = FILTER_DECODE
    primary_scorers: tuple = ((QUEUE, 2.0, 0.0), (KV_UTIL, 2.0, 0.0), (PREFIX, 3.0, 0.0))
    prefill_scorers: tuple | None = None           # set => P/D disagg (prefill-filter)
    non_cached_tokens: int = 0
    n_prefill: int = 0
    G: int = 0

    @property
    def seed(self) -> int:
        return BASE_SEED + self.config_index

    @property
    def blocks(self) -> int:
        return min(self.T // self.block_size_tokens, self.max_prefix_blocks)

    @property
    def prompt_bytes(self) -> int:
        return self.T * 4

    def scaled(self, E=None, R=None, T=None, name=None) -> "Workload":
        import dataclasses
        w = dataclasses.replace(self, E=E or self.E, R=R or self.R, T=T or self.T, name=name or self.name + "-scaled")
        if self.n_prefill:
            w.n_prefill = max(1, int(round(self.n_prefill * w.E / self.E)))
        w.max_prefix_blocks = max(1, min(self.max_prefix_blocks, w.T // w.block_size_tokens)) if T else self.max_prefix_blocks
        return w


def baseline_configs() -> dict:
    """BASELINE.json configs 1-5."""
    return {
        # (1) deploy/config/epp-estimate-prefix-cache-config.yaml: decode-filter + prefix w1 + load-aware w1
        "config1": Workload("config1", 1, E=64, T=256, R=4096, primary_scorers=((PREFIX, 1.0, 0.0), (LOAD_AWARE, 1.0, 128.0))),
        # (2) prefix-cache-affinity scorer only
        "config2": Workload("config2", 2, E=1024, T=2048, R=8192, primary_scorers=((PREFIX, 1.0, 0.0),)),
        # (3) reference default order/weights queue 2, kv 2, prefix 3 (config/loader/defaults.go:47-49, 78-87)
        "config3": Workload("config3", 3, E=4096, T=4096, R=65536),
        # (4) P/D: deploy/config/pd-epp-config.yaml: both profiles prefix w2 + queue w1, nonCachedTokens 16
        "config4": Workload("config4", 4, E=2560, T=8192, R=65536, max_prefix_blocks=512, n_prefill=512,
                            primary_scorers=((PREFIX, 2.0, 0.0), (QUEUE, 1.0, 0.0)),
                            prefill_scorers=((PREFIX, 2.0, 0.0), (QUEUE, 1.0, 0.0)), non_cached_tokens=16),
        # (5) 8 GPUs, 32768 endpoints (4096 per GPU), config-3 scorers
        "config5": Workload("config5", 5, E=32768, T=4096, R=65536),
    }


class Trace:
    def __init__(self, w: Workload, n_threads: int | None = None):
        self.w = w
        self.c = _Cfg()
        self.c.seed = w.seed
        self.c.E, self.c.T, self.c.bst, self.c.G, self.c.R, self.c.n_prefill = w.E, w.T, w.block_size_tokens, w.G, w.R, w.n_prefill
        self.G = int(_L().tg_families(C.byref(self.c)))
        self.n_threads = n_threads or max(1, min(32, os.cpu_count() or 1))
        self._fam = None

    def family_tokens(self) -> np.ndarray:
        if self._fam is None:
            out = np.empty((self.G, self.w.T), dtype=np.uint32)
            _L().tg_family_tokens(C.byref(self.c), out.ctypes.data)
            self._fam = out
        return self._fam

    def requests(self, r0: int = 0, n: int | None = None, out: np.ndarray | None = None):
        """tokens [n, T] u32 (+ family id [-1 = cold], shared blocks L_r)."""
        n = self.w.R - r0 if n is None else n
        fam = self.family_tokens()
        if out is None:
            out = np.empty((n, self.w.T), dtype=np.uint32)
        assert out.dtype == np.uint32 and out.size == n * self.w.T and out.flags["C_CONTIGUOUS"]
        fam_of = np.empty(n, dtype=np.int32)
        shared = np.empty(n, dtype=np.int32)
        _L().tg_requests(C.byref(self.c), fam.ctypes.data, r0, n, out.ctypes.data, fam_of.ctypes.data,
                         shared.ctypes.data, self.n_threads)
        return out, fam_of, shared

    def pool(self):
        E = self.w.E
        kv = np.empty(E, dtype=np.float64)
        waiting = np.empty(E, dtype=np.int32)
        running = np.empty(E, dtype=np.int32)
        role = np.empty(E, dtype=np.uint8)
        _L().tg_pool(C.byref(self.c), kv.ctypes.data, waiting.ctypes.data, running.ctypes.data, role.ctypes.data)
        return role, kv, waiting, running

    def index_plan(self):
        n = np.zeros(self.G, dtype=np.int32)
        ep = np.zeros((self.G, 8), dtype=np.uint32)
        depth = np.zeros((self.G, 8), dtype=np.int32)
        hole = np.zeros((self.G, 8), dtype=np.int32)
        _L().tg_index_plan(C.byref(self.c), n.ctypes.data, ep.ctypes.data, depth.ctypes.data, hole.ctypes.data)
        return n, ep, depth, hole

    def index_pairs(self, family_hashes: np.ndarray):
        """family_hashes [G, >=B] u64 (block hashes of the canonical family prompts) -> (hashes, endpoints)."""
        n, ep, depth, hole = self.index_plan()
        B = self.w.blocks
        fh = np.asarray(family_hashes).astype(np.uint64, copy=False)
        hs, es = [], []
        for g in range(self.G):
            for k in range(int(n[g])):
                d = min(int(depth[g, k]), B)
                h0 = min(int(hole[g, k]), d)
                if d > h0:
                    hs.append(fh[g, h0:d])
                    es.append(np.full(d - h0, ep[g, k], dtype=np.uint32))
        if not hs:
            return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
        return np.concatenate(hs), np.concatenate(es)
