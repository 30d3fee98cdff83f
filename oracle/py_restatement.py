"""Independent pure-Python restatement of the hot path (TEST INFRASTRUCTURE ONLY).

Second, deliberately naive restatement of the same reference functions as oracle/epp_oracle.c, written
with Python dicts/sets so that it shares no code (and no data-structure choices) with the C oracle.
Used only to cross-check the C oracle on small cases.  XXH64 comes from python-xxhash (an independent
implementation of the public XXH64 spec) when importable, else from the bundled pure-Python version.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

M64 = (1 << 64) - 1
P1, P2, P3, P4, P5 = (0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9, 0x85EBCA77C2B2AE63,
                      0x27D4EB2F165667C5)


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _round(acc, x):
    return (_rotl((acc + x * P2) & M64, 31) * P1) & M64


def _merge(h, v):
    return (((h ^ _round(0, v)) * P1) + P4) & M64


def xxh64_pure(data: bytes, seed: int = 0) -> int:
    """Public XXH64 spec (what cespare/xxhash/v2 implements; go.mod:10)."""
    n = len(data)
    p = 0
    if n >= 32:
        v1, v2, v3, v4 = (seed + P1 + P2) & M64, (seed + P2) & M64, seed & M64, (seed - P1) & M64
        while p + 32 <= n:
            a, b, c, d = struct.unpack_from("<4Q", data, p)
            v1, v2, v3, v4 = _round(v1, a), _round(v2, b), _round(v3, c), _round(v4, d)
            p += 32
        h = (_rotl(v1, 1) + _rotl(v2, 7) + _rotl(v3, 12) + _rotl(v4, 18)) & M64
        for v in (v1, v2, v3, v4):
            h = _merge(h, v)
    else:
        h = (seed + P5) & M64
    h = (h + n) & M64
    while p + 8 <= n:
        (k,) = struct.unpack_from("<Q", data, p)
        h ^= _round(0, k)
        h = (_rotl(h, 27) * P1 + P4) & M64
        p += 8
    if p + 4 <= n:
        (k,) = struct.unpack_from("<I", data, p)
        h ^= (k * P1) & M64
        h = (_rotl(h, 23) * P2 + P3) & M64
        p += 4
    while p < n:
        h ^= (data[p] * P5) & M64
        h = (_rotl(h, 11) * P1) & M64
        p += 1
    h ^= h >> 33
    h = (h * P2) & M64
    h ^= h >> 29
    h = (h * P3) & M64
    h ^= h >> 32
    return h


try:  # independent implementation of the same spec
    import xxhash as _xxhash

    def xxh64(data: bytes, seed: int = 0) -> int:
        return _xxhash.xxh64(data, seed=seed).intdigest()
except Exception:  # pragma: no cover
    xxh64 = xxh64_pure


def hash_prompt(data: bytes, model: bytes, block_size_tokens: int, max_prefix_blocks: int, salt: bytes = b""):
    """hashPrompt: approximateprefix/hashing.go:35-99."""
    bs = block_size_tokens * 4                                   # :49
    if bs <= 0 or len(data) < bs:                                # :51-61
        return []
    if len(data) > bs * max_prefix_blocks:                       # :63-66
        data = data[: max(0, bs * max_prefix_blocks)]
    prev = xxh64(model + salt)                                   # :71-78
    res = []
    i = 0
    while i + bs <= len(data):                                   # :80-87
        prev = xxh64(data[i:i + bs] + struct.pack("<Q", prev))
        res.append(prev)
        i += bs
    if i < len(data):                                            # :90-96
        prev = xxh64(data[i:] + struct.pack("<Q", prev))
        res.append(prev)
    return res


class Indexer:
    """approximateprefix/indexer.go:32-182 with golang-lru semantics (OrderedDict: last = most recent)."""

    def __init__(self, default_lru_size=31250):
        self.hash_to_pods: dict[int, set[int]] = {}
        self.pod_to_lru: dict[int, tuple[int, OrderedDict]] = {}
        self.default = default_lru_size

    def _evict(self, h, server):                                 # :105-115
        s = self.hash_to_pods.get(h)
        if s is not None:
            s.discard(server)
            if not s:
                del self.hash_to_pods[h]

    def add(self, hashes, server, num_gpu_blocks=0):             # :52-83
        if server not in self.pod_to_lru:
            size = num_gpu_blocks if num_gpu_blocks > 0 else self.default
            self.pod_to_lru[server] = (size, OrderedDict())
        size, lru = self.pod_to_lru[server]
        for h in hashes:
            if h in lru:
                lru.move_to_end(h)
            else:
                lru[h] = None
                if len(lru) > size:
                    old, _ = lru.popitem(last=False)
                    self._evict(old, server)
        for h in hashes:
            self.hash_to_pods.setdefault(h, set()).add(server)

    def get(self, h):                                            # :86-102
        return set(self.hash_to_pods.get(h, ()))

    def remove_pod(self, server):                                # :167-182
        if server not in self.pod_to_lru:
            return
        _, lru = self.pod_to_lru[server]
        for h in list(lru.keys()):
            self._evict(h, server)
        del self.pod_to_lru[server]

    def match_longest_prefix(self, hashes):                      # plugin.go:214-230
        res: dict[int, int] = {}
        for h in hashes:
            servers = self.get(h)
            if not servers:
                break
            for s in servers:
                res[s] = res.get(s, 0) + 1
        return res


DECODE_KEEP = {"decode", "prefill-decode", "both", "encode-prefill-decode"}            # roles.go:46-48
PREFILL_KEEP = {"prefill", "encode-prefill", "prefill-decode", "both", "encode-prefill-decode"}  # :56-58
ENCODE_KEEP = {"encode", "encode-prefill", "encode-prefill-decode"}                    # :68-70


def role_filter(kind: str, label):
    """filter.go:104-117; label None = no llm-d.ai/role label."""
    if kind == "none":
        return True
    if kind == "decode":
        return label is None or label in DECODE_KEEP
    if kind == "prefill":
        return label in PREFILL_KEEP
    if kind == "encode":
        return label in ENCODE_KEEP
    raise ValueError(kind)


def clamp01(s):                                                  # scheduler_profile.go:194-202
    if s < 0:
        return 0.0
    if s > 1:
        return 1.0
    return s


def score(kind, param, endpoints, match, total):
    """endpoints: list of dicts {kv, waiting, running, ext:[...]}, candidates only."""
    if kind == "prefix":
        return [(match[i] / total) if total != 0 else 0.0 for i, _ in enumerate(endpoints)]
    if kind == "kv":
        return [1 - e["kv"] for e in endpoints]
    if kind in ("queue", "running"):
        key = "waiting" if kind == "queue" else "running"
        mn = min(e[key] for e in endpoints)
        mx = max(e[key] for e in endpoints)
        return [1.0 if mx == mn else float(mx - e[key]) / float(mx - mn) for e in endpoints]
    if kind == "load":
        thr = float(param) if param > 0 else 128.0
        out = []
        for e in endpoints:
            w = float(e["waiting"])
            out.append(0.5 if w == 0 else 0.5 * (1.0 - (min(w, thr) / thr)))
        return out
    if kind == "ext":
        return [e["ext"][int(param)] for e in endpoints]
    if kind == "token_load":                                     # scorer/tokenload/token_load.go:84-112
        thr = float(param) if param > 0 else 4194304.0           # :27-28, :60-62
        out = []
        for e in endpoints:
            load = float(e.get("tokens", 0))
            out.append(1.0 if load <= 0 else 1.0 - (min(load, thr) / thr))
        return out
    if kind == "active_request":                                 # scorer/activerequest/active_request.go:140-173
        max_busy, idle = param                                   # NewActiveRequest :83-93
        idle = int(idle) if idle >= 0 else 0
        max_busy = max_busy if 0 <= max_busy <= 1.0 else 1.0
        counts = [int(e.get("requests", 0)) for e in endpoints]
        mx = max([0] + counts)
        return [1.0 if c <= idle else float(mx - c) / float(mx) * max_busy for c in counts]
    if kind == "lora":                                           # scorer/loraaffinity/lora_affinity.go:76-100
        target = param
        out = []
        for e in endpoints:
            active, waiting_m, cap = e.get("active", set()), e.get("waiting_models", set()), e.get("max_active", 0)
            if target in active:
                out.append(1.0)
            elif len(active) + len(waiting_m) < cap:
                out.append(0.8)
            elif target in waiting_m:
                out.append(0.6)
            else:
                out.append(0.0)
        return out
    raise ValueError(kind)


def affinity_filter(cfg, cand, endpoints, match, total, explore_draw):
    """Plugin.Filter, filter/prefixcacheaffinity/plugin.go:105-151.  cfg = (affinityThreshold, explorationProbability,
    maxTTFTPenaltyMs); endpoints[i].get("ttft") = predicted TTFT or None (attribute absent); explore_draw = the value
    rand.Float64() returned.  Returns the surviving candidate indices."""
    thr, eps, max_penalty = cfg
    if len(cand) <= 1 or thr <= 0:                                # :108-110
        return cand
    if explore_draw < eps:                                        # :113-117
        return cand

    def prefix_cache_score(i):                                    # :160-171
        return match[i] / total if total > 0 else 0.0

    def best_ttft(idx):                                           # :173-184
        best = 1.7976931348623157e308
        for i in idx:
            t = endpoints[i].get("ttft")
            if t is not None and t < best:
                best = t
        return best

    sticky = [i for i in cand if prefix_cache_score(i) >= thr]   # :120-127
    non_sticky = [i for i in cand if not prefix_cache_score(i) >= thr]
    if not sticky:                                                # :130-134
        return cand
    if max_penalty > 0 and non_sticky and best_ttft(sticky) - best_ttft(non_sticky) > max_penalty:   # :137-146
        return cand
    return sticky


def pick_first_k(acc, cand, k, shuffled):
    """MaxScorePicker.Pick, picker/maxscore/picker.go:87-115: `shuffled` is the candidate list after
    ShuffleScoredEndpoints (any permutation of cand); stable sort by score descending; first k."""
    assert sorted(shuffled) == sorted(cand)
    return sorted(shuffled, key=lambda i: -acc[i])[:k]           # Python's sort is stable, like slices.SortStableFunc


def profile_run(filter_kind, scorers, endpoints, match, total, affinity=None, explore_draw=1.0):
    """scheduler_profile.go:117-192; returns (dict idx->score, max, argmax set) or None if no candidates."""
    cand = [i for i, e in enumerate(endpoints) if role_filter(filter_kind, e.get("role"))]
    if not cand:
        return None
    if affinity is not None:
        cand = affinity_filter(affinity, cand, endpoints, match, total, explore_draw)
    acc = {i: 0.0 for i in cand}
    sub = [endpoints[i] for i in cand]
    subm = [match[i] for i in cand]
    for kind, weight, param in scorers:
        col = score(kind, param, sub, subm, total)
        for k, i in enumerate(cand):
            acc[i] = acc[i] + clamp01(col[k]) * weight
    mx = max(acc.values())
    return acc, mx, [i for i in cand if acc[i] == mx]


def pd_decide(nct, input_len_bytes, match_blocks, block_size_tokens):   # prefix_based_pd_decider.go:99-149
    if nct == 0:
        return False
    tokens = input_len_bytes // 4
    if tokens < nct:
        return False
    return (tokens - match_blocks * block_size_tokens) >= nct
