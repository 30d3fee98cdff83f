"""Shared workload set-up (bench.py, tools, parity tests): builds the same synthetic workload for the CUDA engine and
for the CPU oracle."""
from __future__ import annotations

import numpy as np

import epp_b200 as epp
from tools import tracegen as tg


def make_engine(w: tg.Workload, **kw) -> epp.Engine:
    primary = epp.ProfileSpec(w.primary_filter, [epp.ScorerSpec(*s) for s in w.primary_scorers])
    prefill = None
    if w.prefill_scorers is not None:
        prefill = epp.ProfileSpec(tg.FILTER_PREFILL, [epp.ScorerSpec(*s) for s in w.prefill_scorers])
    return epp.Engine(w.E, primary, prefill, block_size_tokens=w.block_size_tokens,
                      max_prefix_blocks=w.max_prefix_blocks, non_cached_tokens=w.non_cached_tokens, **kw)


def filler_pairs(E: int, per_endpoint: int, seed: int = 0x0F111E5):
    """`per_endpoint` extra (hash, endpoint) pairs per endpoint whose hashes no prompt of the trace produces: they fill
    the index to its production size (31 250 blocks per endpoint = the reference's default LRU capacity) without
    changing a single decision."""
    rng = np.random.default_rng(seed)
    n = E * per_endpoint
    hs = rng.integers(1, 2**63, size=n, dtype=np.int64).view(np.uint64) | np.uint64(1 << 63)
    es = np.repeat(np.arange(E, dtype=np.uint32), per_endpoint)
    return hs, es


def setup_engine(eng: epp.Engine, w: tg.Workload, trace: tg.Trace, filler_per_endpoint: int = 0):
    """model + pool snapshot + index snapshot (family hashes come from the ENGINE's own hash kernel)."""
    mid = eng.register_model(tg.MODEL)
    role, kv, waiting, running = trace.pool()
    eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
    fam = trace.family_tokens()
    fh, nb = eng.hash_prompts(fam, uniform_len=w.prompt_bytes)
    hs, es = trace.index_pairs(fh)
    if filler_per_endpoint > 0:
        fhs, fes = filler_pairs(w.E, filler_per_endpoint)
        hs, es = np.concatenate([hs, fhs]), np.concatenate([es, fes])
    eng.index_load_snapshot(hs, es)
    return mid, (hs, es)


def setup_oracle(orc, w: tg.Workload, trace: tg.Trace):
    role, kv, waiting, running = trace.pool()
    pool = orc.PoolState(role, kv, waiting, running)
    fam = trace.family_tokens()
    fh = np.zeros((trace.G, w.blocks), dtype=np.uint64)
    for g in range(trace.G):
        h = orc.hash_prompt(fam[g].tobytes(), tg.MODEL, w.block_size_tokens, w.max_prefix_blocks)
        fh[g, : len(h)] = h
    hs, es = trace.index_pairs(fh)
    ix = orc.Indexer()
    ix.load_pairs(hs, es)
    primary = orc.make_profile(w.primary_filter, list(w.primary_scorers))
    prefill = orc.make_profile(tg.FILTER_PREFILL, list(w.prefill_scorers)) if w.prefill_scorers is not None else None
    return pool, ix, primary, prefill, (hs, es)


def oracle_decisions(orc, w, pool, ix, primary, prefill, tokens: np.ndarray, n_threads: int = 4, **kw):
    """kw: tie_seed / tie_base (the reproducible random tie rule), encode / multimodal (the encode stage)."""
    R = tokens.shape[0]
    offs = np.arange(R + 1, dtype=np.uint64) * np.uint64(w.prompt_bytes)
    return orc.cycle_batch(tg.MODEL, w.block_size_tokens, w.max_prefix_blocks, w.non_cached_tokens, False, ix, primary,
                           prefill, pool, tokens, offs, n_threads, **kw)


def assert_decisions_equal(dec, det, odec, ototal, *, where=""):
    """Engine decisions vs oracle decisions: bit-exact scores, identical picks under the lowest-index tie rule."""
    dec = np.asarray(dec)
    assert dec.shape[0] == odec.shape[0]
    ok = odec["status"] == 0
    np.testing.assert_array_equal(dec["status"], odec["status"], err_msg=where + " status")
    np.testing.assert_array_equal(dec["total_blocks"], ototal, err_msg=where + " total_blocks")
    np.testing.assert_array_equal(dec["pick"][ok].astype(np.int64), odec["pick"][ok].astype(np.int64), err_msg=where + " pick")
    np.testing.assert_array_equal(dec["score"][ok].view(np.uint64), odec["score"][ok].view(np.uint64), err_msg=where + " score bits")
    np.testing.assert_array_equal(dec["tie_count"][ok].astype(np.int64), odec["tie_count"][ok].astype(np.int64), err_msg=where + " tie_count")
    eng_pf = dec["prefill_pick"].astype(np.int64)
    eng_pf[eng_pf == 0xFFFFFFFF] = -1
    np.testing.assert_array_equal(eng_pf[ok], odec["prefill_pick"][ok].astype(np.int64), err_msg=where + " prefill_pick")
    if det is not None:
        det = np.asarray(det)
        np.testing.assert_array_equal(det["prefill_ran"][ok].astype(np.int64), odec["prefill_ran"][ok].astype(np.int64), err_msg=where + " prefill_ran")
        has = ok & (odec["prefill_pick"] >= 0)
        np.testing.assert_array_equal(det["prefill_score"][has].view(np.uint64), odec["prefill_score"][has].view(np.uint64), err_msg=where + " prefill score bits")
        np.testing.assert_array_equal(det["prefill_tie_count"][has].astype(np.int64), odec["prefill_tie_count"][has].astype(np.int64), err_msg=where + " prefill ties")
        if "encode_ran" in odec.dtype.names:
            np.testing.assert_array_equal(det["encode_ran"][ok].astype(np.int64), odec["encode_ran"][ok].astype(np.int64), err_msg=where + " encode_ran")
            eng_en = det["encode_pick"].astype(np.int64)
            eng_en[eng_en == 0xFFFFFFFF] = -1
            np.testing.assert_array_equal(eng_en[ok], odec["encode_pick"][ok].astype(np.int64), err_msg=where + " encode_pick")
            hase = ok & (odec["encode_pick"] >= 0)
            np.testing.assert_array_equal(det["encode_score"][hase].view(np.uint64), odec["encode_score"][hase].view(np.uint64), err_msg=where + " encode score bits")
            np.testing.assert_array_equal(det["encode_tie_count"][hase].astype(np.int64), odec["encode_tie_count"][hase].astype(np.int64), err_msg=where + " encode ties")
