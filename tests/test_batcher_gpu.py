"""The in-library micro-batcher (csrc/batcher.cu: epp_submit / epp_wait) -- the shape the reference's director needs:
Scheduler.Schedule(ctx, *InferenceRequest, []Endpoint) for ONE request, called concurrently from one goroutine per
in-flight request (requestcontrol/director.go:69-71, 243; handlers/server.go:168)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def epp():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import epp_b200
    epp_b200.build.build()
    return epp_b200


@pytest.fixture(scope="module")
def tg():
    from tools import tracegen
    tracegen.build()
    return tracegen


def test_64_threads_one_request_each_vs_oracle(epp, orc, tg):
    """64 threads call schedule() for single requests against a frozen index: every decision must be the oracle's for
    that prompt, whatever batches the flusher happened to form (scaled BASELINE config 4: P/D two-stage pick)."""
    import helpers
    w = tg.baseline_configs()["config4"].scaled(E=160, R=64 * 40, T=512, name="config4")
    w.non_cached_tokens = 64
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, prefill, tokens)
    with helpers.make_engine(w) as eng:
        helpers.setup_engine(eng, w, trace)
        got = np.zeros(w.R, dtype=epp.DECISION_DTYPE)
        gdet = np.zeros(w.R, dtype=epp.DETAIL_DTYPE)
        errors = []
        with epp.Batcher(eng, max_batch=48, max_delay_us=300) as bt:
            def worker(t):
                try:
                    for r in range(t, w.R, 64):
                        d, dd = bt.schedule(tokens[r])
                        got[r], gdet[r] = d, dd
                except Exception as e:      # noqa: BLE001
                    errors.append(e)
            th = [threading.Thread(target=worker, args=(t,)) for t in range(64)]
            for x in th:
                x.start()
            for x in th:
                x.join()
            st = bt.stats()
        assert not errors, errors[:1]
        helpers.assert_decisions_equal(got, gdet, odec, ototal, where="batcher, 64 threads")
        assert st["n_requests"] == w.R and st["n_pending"] == 0
        assert st["n_flushes"] < w.R            # requests really were batched
        assert (gdet["prefill_ran"] == 1).any() and (got["match_blocks"] > 0).any()


def test_submit_many_then_wait_and_ticket_rules(epp, orc, tg):
    """submit() returns at once; tickets of one thread come back in order; full batches flush without waiting for the
    delay; a ticket can be waited for again; prompts longer than the hashed prefix keep their true length for the P/D
    decider (EPP_BATCH_LENGTHS_EXCEED_ROWS)."""
    import helpers
    w = tg.baseline_configs()["config4"].scaled(E=96, R=512, T=256, name="config4")
    w.non_cached_tokens = 32
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    # prompts twice as long as the engine hashes (max_prefix_blocks covers T tokens): the tail is never read
    long_tokens = np.concatenate([tokens, np.flip(tokens, axis=1)], axis=1)
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    offs = np.arange(w.R + 1, dtype=np.uint64) * np.uint64(2 * w.prompt_bytes)
    odec, ototal = orc.cycle_batch(tg.MODEL, w.block_size_tokens, w.max_prefix_blocks, w.non_cached_tokens, False, ix,
                                   primary, prefill, pool, long_tokens, offs, 2)
    with helpers.make_engine(w) as eng:
        helpers.setup_engine(eng, w, trace)
        with epp.Batcher(eng, max_batch=64, max_delay_us=2_000_000) as bt:      # only FULL batches flush
            tickets = [bt.submit(long_tokens[r]) for r in range(w.R)]
            got = np.zeros(w.R, dtype=epp.DECISION_DTYPE)
            gdet = np.zeros(w.R, dtype=epp.DETAIL_DTYPE)
            for r in reversed(range(w.R)):               # any order
                got[r], gdet[r] = bt.wait(tickets[r])
            again, _ = bt.wait(tickets[0])               # a ticket stays good until 4096 later batches were flushed
            assert again == got[0]
            with pytest.raises(epp.EngineError):
                bt.wait(tickets[-1] + 7)                 # never issued
            st = bt.stats()
            assert st["n_flushes"] == w.R // 64 and st["n_full_flushes"] == w.R // 64
        helpers.assert_decisions_equal(got, gdet, odec, ototal, where="batcher, long prompts")


def test_batch_of_one_with_prerequest_is_the_reference_order(epp, orc, tg):
    """max_batch = 1 + index_picks: schedule, PreRequest, schedule, ... -- exactly the reference's per-request order
    (director.go:243-253: Schedule then the PreRequest plugins), so the evolving index must match the oracle's."""
    import helpers
    w = tg.baseline_configs()["config3"].scaled(E=24, R=120, T=256, name="config3")
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    role, kv, waiting, running = trace.pool()
    pool = orc.PoolState(role, kv, waiting, running)
    primary = orc.make_profile(w.primary_filter, list(w.primary_scorers))
    ix = orc.Indexer(40)
    with helpers.make_engine(w, lru_capacity_per_server=40) as eng:
        eng.register_model(tg.MODEL)
        eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        with epp.Batcher(eng, max_batch=1, max_delay_us=0, index_picks=True) as bt:
            for r in range(w.R):
                d, dd = bt.schedule(tokens[r])
                od, ot = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens[r:r + 1], 1)
                helpers.assert_decisions_equal(np.array([d]), np.array([dd]), od, ot, where=f"request {r}")
                if od["status"][0] == 0:
                    ix.add(orc.hash_prompt(tokens[r].tobytes(), tg.MODEL, w.block_size_tokens, w.max_prefix_blocks), int(od["pick"][0]))
        eng.index_commit()
        assert eng.stats()["index_pairs"] == len(ix.export()[0])


def test_batcher_returns_the_first_k_lists(epp, orc, tg):
    """pick_k = 3 through epp_submit / epp_wait_topk from 16 threads: the lists are the oracle's first-k of the request's
    profiles (deterministic tie mode, so the answer does not depend on how the flusher grouped the requests)."""
    import helpers
    K = 3
    w = tg.baseline_configs()["config4"].scaled(E=96, R=16 * 20, T=256, name="config4")
    w.non_cached_tokens = 32
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, ototal, olists = helpers.oracle_decisions(orc, w, pool, ix, primary, prefill, tokens, topk=K)
    with helpers.make_engine(w, pick_k=K) as eng:
        helpers.setup_engine(eng, w, trace)
        got = np.zeros(w.R, dtype=epp.DECISION_DTYPE)
        gdet = np.zeros(w.R, dtype=epp.DETAIL_DTYPE)
        glists = [None] * w.R
        errors = []
        with epp.Batcher(eng, max_batch=24, max_delay_us=200) as bt:
            def worker(t):
                try:
                    for r in range(t, w.R, 16):
                        got[r], gdet[r], glists[r] = bt.wait_topk(bt.submit(tokens[r]))
                except Exception as e:      # noqa: BLE001
                    errors.append(e)
            th = [threading.Thread(target=worker, args=(t,)) for t in range(16)]
            for x in th:
                x.start()
            for x in th:
                x.join()
        assert not errors, errors[:1]
        helpers.assert_decisions_equal(got, gdet, odec, ototal, where="batcher top-k")
        for r in range(w.R):
            if odec["status"][r] != 0:
                continue
            for name in ("primary", "prefill"):
                assert glists[r][name] == [int(x) for x in olists[name][r] if x >= 0], (r, name)
            assert glists[r]["encode"] == []
        assert any(len(l["prefill"]) == K for l in glists if l)
