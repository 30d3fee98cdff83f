"""CPU check of the ALGORITHM behind the sparse match/pick kernel (csrc/match_sparse.cu).

The reference scores EVERY candidate endpoint for every request (runScorerPlugins, scheduler_profile.go:151-174) and
picks a member of the arg-max set.  The kernel instead evaluates only the endpoints that hold part of the prompt's
prefix and takes everyone else from a per-pool precomputation:

  * an endpoint with zero matched blocks has exactly the ordered weighted sum `base[e]` = the same left-to-right
    accumulation with every prefix term replaced by clamp(0) * w -- request independent, bit for bit;
  * candidates sorted once by (base desc, slot asc) with the size of every equal-base group; the best unmatched
    endpoint is the first unmatched entry of that order, and the number of unmatched endpoints tied with it is its
    group size minus the matched endpoints of the same base.

`sparse_pick` restates that in Python floats (IEEE double, no fusion) and is compared with the oracle's full scan:
max score bits, lowest slot id of the arg-max set and the size of the set.  No GPU involved."""
from __future__ import annotations

import numpy as np
import pytest


def clamp01(x):
    return 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)


def pool_precompute(orc, scorers, pool, cand):
    """k_pool_terms / k_pool_sort / k_pool_groups: contrib[s][e], base[e], order, group sizes."""
    n = pool.n
    zeros = np.zeros(n, np.int32)
    contrib = []
    for sc in scorers:
        if sc[0] == orc.SCORER_PREFIX:
            contrib.append(None)
        else:
            col = orc.score_column(sc, pool, cand, zeros, 0)
            contrib.append([clamp01(float(col[e])) * sc[1] for e in range(n)])
    base = []
    for e in range(n):
        acc = 0.0
        for s, sc in enumerate(scorers):
            acc = acc + (clamp01(0.0) * sc[1] if contrib[s] is None else contrib[s][e])
        base.append(acc)
    order = sorted((e for e in range(n) if cand[e]), key=lambda e: (-base[e], e))
    grp = {}
    for e in order:
        grp[base[e]] = grp.get(base[e], 0) + 1
    return contrib, base, order, grp


def sparse_pick(scorers, contrib, base, order, grp, cand, matched: dict, total: int):
    """eval_profile_lanes: matched = {endpoint: blocks}.  -> (max score, lowest slot id, tie count) or None."""
    best = None

    def add(val, e, ties):
        nonlocal best
        if best is None or val > best[0]:
            best = (val, e, ties)
        elif val == best[0]:
            best = (val, min(e, best[1]), best[2] + ties)

    for e, m in matched.items():
        if not cand[e]:
            continue
        acc = 0.0
        for s, sc in enumerate(scorers):
            if contrib[s] is None:
                raw = 0.0 if total == 0 else float(m) / float(total)
                acc = acc + clamp01(raw) * sc[1]
            else:
                acc = acc + contrib[s][e]
        add(acc, e, 1)
    for e in order:
        if e in matched:
            continue
        same = sum(1 for me in matched if cand[me] and base[me] == base[e])
        add(base[e], e, grp[base[e]] - same)
        break
    return best


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_sparse_pick_equals_full_scan(orc, seed):
    rng = np.random.default_rng(seed)
    n = 61
    for trial in range(30):
        role = rng.choice([0, 1, 2, 3, 4, 8], n).astype(np.uint8)
        # coarse metric values on purpose: many endpoints share a base, so the tie bookkeeping is exercised
        kv = rng.choice([0.0, 0.25, 0.5, 0.75, 1.0, 1.25], n)
        waiting = rng.choice([0, 0, 0, 3, 7, 200], n).astype(np.int32)
        running = rng.choice([0, 1, 5], n).astype(np.int32)
        ext = np.stack([rng.choice([-0.5, 0.0, 0.4, 1.0, 1.7], n), rng.choice([0, 100, 6000], n).astype(np.float64)])
        pool = orc.PoolState(role, kv, waiting, running, ext)
        scorers = [(orc.SCORER_QUEUE, 2.0, 0.0), (orc.SCORER_KV_UTIL, 2.0, 0.0), (orc.SCORER_PREFIX, 3.0, 0.0),
                   (orc.SCORER_LOAD_AWARE, 1.0, 10.0), (orc.SCORER_EXTERNAL, 0.5, 0), (orc.SCORER_TOKEN_LOAD, 1.0, 5000.0, 1)]
        k = int(rng.integers(1, len(scorers) + 1))
        scorers = [scorers[i] for i in rng.permutation(len(scorers))[:k]]
        filt = int(rng.choice([orc.FILTER_NONE, orc.FILTER_DECODE, orc.FILTER_PREFILL]))
        prof = orc.make_profile(filt, scorers)
        cand = np.array([orc.lib().orc_role_filter_keeps(filt, int(r)) for r in role], np.uint8)
        contrib, base, order, grp = pool_precompute(orc, scorers, pool, cand)
        for _ in range(8):
            total = int(rng.choice([0, 1, 16, 256]))
            match = np.zeros(n, np.int32)
            if total:
                idx = rng.choice(n, int(rng.integers(0, 9)), replace=False)
                match[idx] = rng.integers(1, total + 1, len(idx))
            matched = {int(e): int(match[e]) for e in np.flatnonzero(match)}
            scores, mx, pick, argmax = orc.profile_run(prof, pool, match, total)
            got = sparse_pick(scorers, contrib, base, order, grp, cand, matched, total)
            if not argmax:
                assert got is None
                continue
            assert got is not None
            assert np.float64(got[0]).view(np.uint64) == np.float64(mx).view(np.uint64), (seed, trial, scorers)
            assert got[1] == pick and got[2] == len(argmax), (seed, trial, scorers, got, pick, len(argmax))
