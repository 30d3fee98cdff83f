"""Known-answer tests taken from the reference's OWN unit tests, re-encoded against
the CPU oracle.  Each test cites the reference test file:line that holds the expected value."""
import numpy as np
import pytest

from oracle import py_restatement as pr

EXT = 4  # orc.SCORER_EXTERNAL


def _pool(orc, role=None, kv=None, waiting=None, running=None, ext=None, n=None):
    n = n or len(role if role is not None else kv if kv is not None else waiting)
    role = role if role is not None else [orc.ROLE_NONE] * n
    kv = kv if kv is not None else [0.0] * n
    waiting = waiting if waiting is not None else [0] * n
    return orc.PoolState(role, kv, waiting, running, ext)


# ---- B.1 #1-#3: scheduling/scheduler_profile_test.go:64-108, :172 -------------------------------------
def test_weighted_sum_same_weight(orc):
    # two test scorers return 0.3 and 0.8 for every endpoint; filters keep pod1, pod2 of 3
    pool = _pool(orc, role=[orc.ROLE_NONE, orc.ROLE_NONE, orc.ROLE_ABSENT], ext=[[0.3] * 3, [0.8] * 3])
    prof = orc.make_profile(orc.FILTER_NONE, [(EXT, 1.0, 0), (EXT, 1.0, 1)])
    scores, mx, pick, aset = orc.profile_run(prof, pool, [0, 0, 0], 0)
    assert mx == 1.1                       # exact `!=` check at scheduler_profile_test.go:172
    assert aset == [0, 1] and pick == 0
    assert scores[2] == -1.0               # filtered out


def test_weighted_sum_different_weights(orc):
    pool = _pool(orc, role=[orc.ROLE_NONE, orc.ROLE_NONE, orc.ROLE_ABSENT], ext=[[0.3] * 3, [0.8] * 3])
    prof = orc.make_profile(orc.FILTER_NONE, [(EXT, 60.0, 0), (EXT, 40.0, 1)])
    _, mx, _, _ = orc.profile_run(prof, pool, [0, 0, 0], 0)
    assert mx == 50                        # scheduler_profile_test.go:80-94


def test_filter_all_is_error(orc):
    pool = _pool(orc, role=[orc.ROLE_DECODE] * 3)
    prof = orc.make_profile(orc.FILTER_PREFILL, [(EXT, 1.0, 0)])
    _, _, pick, aset = orc.profile_run(prof, pool, [0, 0, 0], 0)
    assert aset == [] and pick == -1       # scheduler_profile_test.go:96-108 -> "no endpoints available"
    d = orc.schedule(prof, None, pool, [0, 0, 0], 0, 16, 0, 0)
    assert d.status == -1


# ---- B.1 #4: clamp, scheduler_profile_test.go:350-413 --------------------------------------------------
def test_enforce_score_range(orc):
    vals = [-0.5, 1.5, 0.0, 1.0, 0.5]
    pool = _pool(orc, n=5, ext=[vals])
    prof = orc.make_profile(orc.FILTER_NONE, [(EXT, 1.0, 0)])
    scores, _, _, _ = orc.profile_run(prof, pool, [0] * 5, 0)
    assert list(scores) == [0.0, 1.0, 0.0, 1.0, 0.5]
    # Run with (-0.5, 1.5) -> total 1.0
    pool = _pool(orc, n=1, ext=[[-0.5], [1.5]])
    prof = orc.make_profile(orc.FILTER_NONE, [(EXT, 1.0, 0), (EXT, 1.0, 1)])
    _, mx, _, _ = orc.profile_run(prof, pool, [0], 0)
    assert mx == 1.0


# ---- B.1 #6/#7: scheduling/scheduler_test.go:67-141 ----------------------------------------------------
def test_default_scheduler_kat(orc):
    # pods (q,kv): (0,0.2) (0,0.2)+critical active (10,0.8); lora-affinity tiers (lora_affinity.go:76-100):
    # pod1: foo,bar active, max 2 -> no slot, not waiting -> 0.0 ; pod2: active -> 1.0 ; pod3: 1 of 2 -> 0.8
    pool = _pool(orc, kv=[0.2, 0.2, 0.8], waiting=[0, 0, 10], ext=[[0.0, 1.0, 0.8]])
    prof = orc.make_profile(orc.FILTER_NONE, [(orc.SCORER_KV_UTIL, 1.0, 0), (orc.SCORER_QUEUE, 1.0, 0),
                                              (orc.SCORER_PREFIX, 1.0, 0), (EXT, 1.0, 0)])
    scores, mx, pick, aset = orc.profile_run(prof, pool, [0, 0, 0], 0)
    assert pick == 1 and aset == [1]
    assert mx == 2.8                       # scheduler_test.go:125 Score: 2.8
    d = orc.schedule(prof, None, pool, [0, 0, 0], 0, 16, 0, 0)
    assert (d.status, d.pick, d.score) == (0, 1, 2.8)


def test_no_candidates_is_error(orc):
    pool = orc.PoolState(np.zeros(0, np.uint8), np.zeros(0), np.zeros(0, np.int32))
    prof = orc.make_profile(orc.FILTER_NONE, [(orc.SCORER_KV_UTIL, 1.0, 0)])
    d = orc.schedule(prof, None, pool, np.zeros(0, np.int32), 0, 16, 0, 0)
    assert d.status == -1                  # scheduler_test.go:67-75


# ---- B.1 #8-#11: scorer unit tests --------------------------------------------------------------------
def test_prefix_scorer(orc):
    pool = _pool(orc, n=2)
    out = orc.score_column((orc.SCORER_PREFIX, 1.0, 0), pool, [1, 1], [5, 2], 10)
    assert list(out) == [0.5, 0.2]         # scorer/prefix/plugin_test.go:31-44 (exact assert.Equal)
    assert list(orc.score_column((orc.SCORER_PREFIX, 1.0, 0), pool, [1, 1], [5, 2], 0)) == [0.0, 0.0]


def test_kv_util_scorer(orc):
    kv = [0.8, 0.5, 0.0, 0.6, 1.0]
    pool = _pool(orc, kv=kv)
    out = orc.score_column((orc.SCORER_KV_UTIL, 1.0, 0), pool, [1] * 5, [0] * 5, 0)
    assert out == pytest.approx([0.2, 0.5, 1.0, 0.4, 0.0], abs=1e-4)   # kvcache_utilization_test.go:36-79
    assert list(out) == [1 - x for x in kv]


def test_queue_scorer(orc):
    pool = _pool(orc, waiting=[10, 5, 0])
    out = orc.score_column((orc.SCORER_QUEUE, 1.0, 0), pool, [1, 1, 1], [0] * 3, 0)
    assert list(out) == [0.0, 0.5, 1.0]    # queuedepth/queue_test.go:36-67
    pool = _pool(orc, waiting=[7, 7, 7])
    assert list(orc.score_column((orc.SCORER_QUEUE, 1.0, 0), pool, [1, 1, 1], [0] * 3, 0)) == [1.0] * 3
    # min/max are over the CANDIDATES only (queue.go:79-91)
    pool = _pool(orc, waiting=[10, 5, 0])
    out = orc.score_column((orc.SCORER_QUEUE, 1.0, 0), pool, [1, 1, 0], [0] * 3, 0)
    assert list(out) == [0.0, 1.0, 0.0]


def test_load_aware_scorer(orc):
    pool = _pool(orc, waiting=[2, 0, 15])
    out = orc.score_column((orc.SCORER_LOAD_AWARE, 1.0, 10), pool, [1, 1, 1], [0] * 3, 0)
    assert list(out) == [0.4, 0.5, 0.0]    # loadaware/load_aware_test.go:38-60 (exact cmp.Diff)
    out = orc.score_column((orc.SCORER_LOAD_AWARE, 1.0, 0), pool, [1, 1, 1], [0] * 3, 0)   # <=0 -> 128
    assert list(out) == [0.5 * (1.0 - 2 / 128.0), 0.5, 0.5 * (1.0 - 15 / 128.0)]


def test_token_load_scorer(orc):
    # tokenload/token_load_test.go:32-61 (threshold 1000; 0 / 500 / 1000 in-flight tokens)
    pool = _pool(orc, n=3, kv=[0.0] * 3, ext=[[0, 500, 1000]])
    out = orc.score_column((orc.SCORER_TOKEN_LOAD, 1.0, 1000.0, 0), pool, [1, 1, 1], [0] * 3, 0)
    assert list(out) == [1.0, 0.5, 0.0]
    out = orc.score_column((orc.SCORER_TOKEN_LOAD, 1.0, 0.0, 0), pool, [1, 1, 1], [0] * 3, 0)   # <= 0 -> 4194304
    assert list(out) == [1.0, 1.0 - 500 / 4194304.0, 1.0 - 1000 / 4194304.0]


def test_active_request_scorer(orc):
    AR = orc.SCORER_ACTIVE_REQUEST

    def run(counts, max_busy=1.0, idle=0.0, cand=None):
        pool = _pool(orc, n=len(counts), kv=[0.0] * len(counts), ext=[counts])
        return list(orc.score_column((AR, 1.0, max_busy, 0, idle), pool, cand or [1] * len(counts), [0] * len(counts), 0))

    # activerequest/active_request_test.go:36-92
    assert run([0, 0, 0]) == [1.0, 1.0, 1.0]
    assert run([3, 0, 6]) == [0.5, 1.0, 0.0]
    assert run([4, 0, 1]) == [0.0, 1.0, 0.75]
    # :172-216 idleThreshold / maxBusyScore
    assert run([0, 0], 0.0, 0) == [1.0, 1.0]
    assert run([1, 0], 0.0, 0) == [0.0, 1.0]
    assert run([1, 2, 0], 0.5, 1) == [1.0, 0.0, 1.0]
    # NewActiveRequest :83-93: out-of-range parameters fall back to the defaults; max is over the candidates only
    assert run([2, 4], 7.0, -3) == [0.5, 0.0]
    assert run([2, 4, 8], cand=[1, 1, 0]) == [0.5, 0.0, 0.0]


def test_lora_affinity_scorer(orc):
    # loraaffinity/lora_affinity_test.go:30-141, TargetModel "active-model-1"; per pod (state, MaxActiveModels, loaded)
    LA = (orc.SCORER_LORA_AFFINITY, 1.0, 0.0)

    def run(pods):
        st, mx, ld = zip(*pods)
        pool = _pool(orc, n=len(pods), kv=[0.0] * len(pods)).set_lora(st, mx, ld)
        return list(orc.score_column(LA, pool, [1] * len(pods), [0] * len(pods), 0))

    assert run([(1, 5, 1)]) == [1.0]                       # "Target model is active"
    assert run([(2, 2, 2)]) == [0.6]                       # "Target model is waiting" (active-model-2 + waiting = 2 of 2)
    assert run([(0, 2, 2), (0, 0, 0)]) == [0.0, 0.0]       # "Endpoints have no space for new model"
    # "Multiple endpoints with mixed active and waiting models"
    assert run([(1, 5, 1), (2, 5, 2), (0, 2, 1), (2, 2, 2), (0, 2, 2)]) == [1.0, 0.8, 0.8, 0.6, 0.0]
    # no adapter information at all: nothing is resident, nobody has room
    pool = _pool(orc, n=2, kv=[0.0, 0.0])
    assert list(orc.score_column(LA, pool, [1, 1], [0, 0], 0)) == [0.0, 0.0]


# ---- B.1 #12: picker/maxscore/picker_test.go:43-110 -- ties are a SET ----------------------------------
def test_max_score_picker_tie_set(orc):
    pool = _pool(orc, n=4, ext=[[0.5, 0.9, 0.9, 0.1]])
    prof = orc.make_profile(orc.FILTER_NONE, [(EXT, 1.0, 0)])
    _, mx, pick, aset = orc.profile_run(prof, pool, [0] * 4, 0)
    assert mx == 0.9 and aset == [1, 2] and pick == 1


# ---- B.1 #13-#15, #20: prefix producer + indexer -------------------------------------------------------
def test_produce_empty_index(orc):
    ix = orc.Indexer()
    hashes = orc.hash_prompt(b"aaaabbbb", b"test-model1", 1, 256)
    counts, walked = ix.match_longest_prefix(hashes, 2)
    assert len(hashes) == 2 and list(counts) == [0, 0] and walked == 0    # plugin_test.go:37-82


def test_prerequest_then_partial_match(orc):
    """plugin_test.go:174-226: index 'aaaaaa' for pod1 (primary) and pod3 (prefill); 'aaaabbbb' then
    matches 1/2 on pod1 and pod3, 0/2 on pod2."""
    ix = orc.Indexer()
    h1 = orc.hash_prompt(b"aaaaaa", b"test-model1", 1, 256)
    ix.add(h1, 0)
    ix.add(h1, 2)
    h3 = orc.hash_prompt(b"aaaabbbb", b"test-model1", 1, 256)
    counts, walked = ix.match_longest_prefix(h3, 3)
    assert list(counts) == [1, 0, 1] and walked == 1 and len(h3) == 2


def test_indexer_add_and_get(orc):
    ix = orc.Indexer(3)                    # indexer_test.go:27-55: server limit 2 beats default 3
    ix.add([1], 7, 2)
    assert ix.lru_len(7) == 1 and ix.get(1) == {7}
    ix.add([2], 7, 2)
    assert ix.lru_len(7) == 2
    ix.add([3], 7, 2)
    assert ix.lru_len(7) == 2
    assert ix.get(4) == set()
    assert ix.get(1) == set()              # evicted (oldest)
    assert ix.get(2) == {7} and ix.get(3) == {7}


def test_indexer_remove_pod_and_eviction(orc):
    size = 10                              # indexer_test.go:57-113
    ix = orc.Indexer(size)
    for j in range(size):
        ix.add([j], 1)
        ix.add([j], 2)
    assert ix.lru_len(1) == size and ix.lru_len(2) == size
    for j in range(size):
        assert ix.get(j) == {1, 2}
    ix.add([size], 1)                      # evicts hash 0 from server 1
    assert ix.lru_len(1) == size
    assert ix.get(0) == {2}
    ix.remove_pod(2)
    assert ix.get(0) == set()
    hs, sv = ix.export()
    assert len(hs) == size and set(sv.tolist()) == {1}
    assert ix.pods() == [1]


def test_indexer_lru_refresh_and_overflow(orc):
    ix = orc.Indexer(3)
    ix.add([1, 2, 3], 0)
    ix.add([1], 0)                         # refresh 1 -> order (old->new) 2,3,1
    ix.add([4], 0)                         # evicts 2
    assert ix.get(2) == set() and ix.get(1) == {0} and ix.get(3) == {0} and ix.get(4) == {0}
    # len(hashes) > capacity: the second loop re-inserts hashes the LRU just evicted (indexer.go:76-83)
    ix2 = orc.Indexer(2)
    ix2.add([10, 11, 12], 5)
    assert ix2.lru_len(5) == 2
    assert ix2.get(10) == {5} and ix2.get(11) == {5} and ix2.get(12) == {5}


def test_indexer_random_vs_python_restatement(orc):
    import random
    rng = random.Random(3)
    ix = orc.Indexer(5)
    py = pr.Indexer(5)
    for step in range(3000):
        op = rng.random()
        srv = rng.randrange(6)
        if op < 0.8:
            hs = [rng.randrange(40) for _ in range(rng.randint(1, 8))]
            cap = rng.choice([0, 3, 7])
            ix.add(hs, srv, cap)
            py.add(hs, srv, cap)
        elif op < 0.9:
            ix.remove_pod(srv)
            py.remove_pod(srv)
        if step % 50 == 0:
            for h in range(40):
                assert ix.get(h) == py.get(h), (step, h)
            q = [rng.randrange(40) for _ in range(6)]
            counts, _ = ix.match_longest_prefix(q, 6)
            want = py.match_longest_prefix(q)
            assert {i: int(c) for i, c in enumerate(counts) if c} == want


def test_match_global_stop_rule(orc):
    """the walk stops at the first block NO server holds; a server missing an earlier block
    still counts later ones (plugin.go:214-230)."""
    ix = orc.Indexer()
    ix.load_pairs([100, 101, 102, 101, 102, 104], [0, 0, 0, 1, 1, 1])
    counts, walked = ix.match_longest_prefix([100, 101, 102, 103, 104], 2)
    assert list(counts) == [3, 2] and walked == 3
    # servers outside the candidate range keep the walk alive (App. C.5)
    ix.load_pairs([103], [9])
    counts, walked = ix.match_longest_prefix([100, 101, 102, 103, 104], 2)
    assert list(counts) == [3, 3] and walked == 5


# ---- B.1 #21: role filters, filter/bylabel/roles.go:46-70 ----------------------------------------------
def test_role_filters(orc):
    L = orc.lib()
    names = {orc.ROLE_NONE: None, orc.ROLE_DECODE: "decode", orc.ROLE_PREFILL: "prefill",
             orc.ROLE_PREFILL_DECODE: "prefill-decode", orc.ROLE_BOTH: "both", orc.ROLE_ENCODE: "encode",
             orc.ROLE_ENCODE_PREFILL: "encode-prefill", orc.ROLE_ENCODE_PREFILL_DECODE: "encode-prefill-decode",
             orc.ROLE_OTHER: "something-else"}
    for fk, fname in ((orc.FILTER_DECODE, "decode"), (orc.FILTER_PREFILL, "prefill"), (orc.FILTER_ENCODE, "encode"),
                      (orc.FILTER_NONE, "none")):
        for role, label in names.items():
            assert bool(L.orc_role_filter_keeps(fk, role)) == pr.role_filter(fname, label), (fname, label)
        assert not L.orc_role_filter_keeps(fk, orc.ROLE_ABSENT)
    assert L.orc_role_filter_keeps(orc.FILTER_DECODE, orc.ROLE_NONE)         # unlabeled pod serves decode
    assert not L.orc_role_filter_keeps(orc.FILTER_PREFILL, orc.ROLE_NONE)
    assert not L.orc_role_filter_keeps(orc.FILTER_DECODE, orc.ROLE_PREFILL)


# ---- B.1 #22: disagg/prefix_based_pd_decider_test.go:203-265 -------------------------------------------
@pytest.mark.parametrize("nct,tokens,match,want", [
    (0, 10, 5, False),     # threshold zero disables
    (20, 10, 0, False),    # input shorter than threshold
    (5, 10, 8, False),     # non-cached suffix below threshold
    (5, 10, 5, True),      # suffix equals threshold
    (3, 10, 2, True),      # suffix above threshold
    (1, 10, 10, False),    # fully cached
    (5, 10, 0, True),      # no hit
])
def test_pd_decider(orc, nct, tokens, match, want):
    # makeRequestWithTokens(n): prompt of n*4 chars; makeTestEndpoint(m): matchBlocks=m, blockSizeTokens=1
    assert orc.pd_decide(nct, tokens * 4, match, 1) is want
    assert pr.pd_decide(nct, tokens * 4, match, 1) is want


# ---- B.1 #23: TestPDSchedule, disagg/scheduler_test.go:34-297 ------------------------------------------
def _pd_profiles(orc):
    prefill = orc.make_profile(orc.FILTER_PREFILL, [(orc.SCORER_PREFIX, 50.0, 0)])
    decode = orc.make_profile(orc.FILTER_DECODE, [(orc.SCORER_LOAD_AWARE, 1.0, 128), (orc.SCORER_PREFIX, 0.0, 0)])
    return decode, prefill


def _pd(orc, roles, waiting, prompt, cached):
    decode, prefill = _pd_profiles(orc)
    pool = _pool(orc, role=roles, waiting=waiting)
    tokens = len(prompt) // 4
    match = [tokens if cached else 0] * len(roles)
    return orc.schedule(decode, prefill, pool, match, tokens, 1, len(prompt), 2)


def test_pd_schedule(orc):
    P, D, N = orc.ROLE_PREFILL, orc.ROLE_DECODE, orc.ROLE_NONE
    # one decode endpoint, long prompt -> decode only? no prefill candidates: prefill run fails, tolerated
    d = _pd(orc, [D], [0], "12345678901", False)
    assert (d.status, d.pick, d.prefill_pick) == (0, 0, -1)
    # one prefill endpoint -> error (no decode workers)
    d = _pd(orc, [P], [0], "12345678901", False)
    assert d.status == -1
    # 1P1D long prompt -> decode=ep2(idx1), prefill=ep1(idx0); second call fully cached -> decode only
    d = _pd(orc, [P, D], [0, 0], "12345678906", False)
    assert (d.status, d.pick, d.prefill_pick) == (0, 1, 0)
    d = _pd(orc, [P, D], [0, 0], "12345678906", True)
    assert (d.status, d.pick, d.prefill_pick, d.prefill_ran) == (0, 1, -1, 0)
    # 1P1D short prompt ("12345" -> 1 token < NCT 2) -> decode only
    d = _pd(orc, [P, D], [0, 0], "12345", False)
    assert (d.status, d.pick, d.prefill_pick) == (0, 1, -1)
    # TestRolesWithNoDecode: unlabeled pod serves decode, ep1 prefill
    d = _pd(orc, [P, N], [0, 2], "12345678901", False)
    assert (d.status, d.pick, d.prefill_pick) == (0, 1, 0)
    # 1P2D: ep2 (q0 -> 0.5) beats unlabeled (q2 -> 0.4921875)
    d = _pd(orc, [P, D, N], [0, 0, 2], "1234567890123456789012345678901234567890", False)
    assert (d.status, d.pick, d.prefill_pick) == (0, 1, 0)
    assert d.score == 0.5
    d = _pd(orc, [P, D, N], [0, 0, 2], "1234567890123456789012345678901234567890", True)
    assert (d.status, d.pick, d.prefill_pick) == (0, 1, -1)
    scores, _, _, _ = orc.profile_run(_pd_profiles(orc)[0], _pool(orc, role=[P, D, N], waiting=[0, 0, 2]), [0] * 3, 10)
    assert scores[2] == 0.4921875


# ---- B.1 #25: ext-proc end-to-end pick, test/integration/epp/common_tests.go:287-300 -------------------
def test_integration_select_lower_queue_and_kv(orc):
    # pods (q,kv) (3,0.2) (0,0.1) (10,0.2); default config queue 2 / kv 2 / prefix 3 (defaults.go:47-49)
    pool = _pool(orc, kv=[0.2, 0.1, 0.2], waiting=[3, 0, 10])
    prof = orc.make_profile(orc.FILTER_NONE, [(orc.SCORER_QUEUE, 2.0, 0), (orc.SCORER_KV_UTIL, 2.0, 0),
                                              (orc.SCORER_PREFIX, 3.0, 0)])
    _, _, pick, aset = orc.profile_run(prof, pool, [0, 0, 0], 0)
    assert pick == 1 and aset == [1]


# ---- cross-check profile_run against the independent Python restatement --------------------------------
def test_profile_run_random_vs_python_restatement(orc):
    import random
    rng = random.Random(11)
    kinds = {orc.SCORER_PREFIX: "prefix", orc.SCORER_KV_UTIL: "kv", orc.SCORER_QUEUE: "queue",
             orc.SCORER_LOAD_AWARE: "load", orc.SCORER_RUNNING: "running"}
    labels = {orc.ROLE_NONE: None, orc.ROLE_DECODE: "decode", orc.ROLE_PREFILL: "prefill",
              orc.ROLE_BOTH: "both", orc.ROLE_ENCODE: "encode"}
    fnames = {orc.FILTER_NONE: "none", orc.FILTER_DECODE: "decode", orc.FILTER_PREFILL: "prefill"}
    for _ in range(200):
        n = rng.randint(1, 40)
        roles = [rng.choice(list(labels)) for _ in range(n)]
        kv = [rng.choice([0.0, 1.0, rng.random(), rng.randrange(1001) / 1000.0]) for _ in range(n)]
        waiting = [rng.choice([0, 0, rng.randint(1, 200)]) for _ in range(n)]
        running = [rng.randint(0, 50) for _ in range(n)]
        total = rng.choice([0, 1, 7, 256])
        match = [rng.randint(0, total) for _ in range(n)]
        sc = [(rng.choice(list(kinds)), rng.choice([0.0, 1.0, 2.0, 3.0, 0.37, 50.0]), rng.choice([0, 10, 128]))
              for _ in range(rng.randint(1, 5))]
        fk = rng.choice(list(fnames))
        pool = _pool(orc, role=roles, kv=kv, waiting=waiting, running=running)
        scores, mx, pick, aset = orc.profile_run(orc.make_profile(fk, sc), pool, match, total)
        eps = [{"role": labels[r], "kv": kv[i], "waiting": waiting[i], "running": running[i]} for i, r in enumerate(roles)]
        want = pr.profile_run(fnames[fk], [(kinds[k], w, p) for k, w, p in sc], eps, match, total)
        if want is None:
            assert aset == []
            continue
        acc, wmx, wset = want
        assert mx == wmx and aset == wset and pick == wset[0]
        for i in range(n):
            assert scores[i] == (acc[i] if i in acc else -1.0)


def test_new_scorers_random_vs_python_restatement(orc):
    """token-load, active-request and lora-affinity of the C oracle against the independent Python restatement (dicts
    and sets, straight from the Go switch statements), inside whole profiles with the other scorers."""
    import random
    rng = random.Random(23)
    labels = {orc.ROLE_NONE: None, orc.ROLE_DECODE: "decode", orc.ROLE_PREFILL: "prefill", orc.ROLE_BOTH: "both"}
    fnames = {orc.FILTER_NONE: "none", orc.FILTER_DECODE: "decode", orc.FILTER_PREFILL: "prefill"}
    adapters = ["a", "b", "c", "d"]
    for _ in range(200):
        n = rng.randint(1, 30)
        roles = [rng.choice(list(labels)) for _ in range(n)]
        kv = [rng.randrange(1001) / 1000.0 for _ in range(n)]
        waiting = [rng.choice([0, 0, rng.randint(1, 200)]) for _ in range(n)]
        tokens = [rng.choice([0, 0, rng.randint(1, 9000), 5000000]) for _ in range(n)]
        requests = [rng.choice([0, 0, 1, rng.randint(2, 40)]) for _ in range(n)]
        active = [set(rng.sample(adapters, rng.randint(0, 3))) for _ in range(n)]
        waiting_m = [set(rng.sample(adapters, rng.randint(0, 2))) - active[i] for i in range(n)]
        max_active = [rng.randint(0, 5) for _ in range(n)]
        target = rng.choice(adapters)
        total = rng.choice([0, 4, 64])
        match = [rng.randint(0, total) for _ in range(n)]
        thr = rng.choice([0.0, 1000.0, 5000.0])
        max_busy, idle = rng.choice([1.0, 0.5, 0.0, 7.0]), rng.choice([0.0, 1.0, 3.0, -2.0])
        pool = _pool(orc, role=roles, kv=kv, waiting=waiting, ext=[tokens, requests])
        state = [1 if target in active[i] else (2 if target in waiting_m[i] else 0) for i in range(n)]
        pool.set_lora(state, max_active, [len(active[i]) + len(waiting_m[i]) for i in range(n)])
        c_sc = [(orc.SCORER_TOKEN_LOAD, 1.5, thr, 0), (orc.SCORER_ACTIVE_REQUEST, 0.7, max_busy, 1, idle),
                (orc.SCORER_LORA_AFFINITY, 2.0, 0.0), (orc.SCORER_PREFIX, 1.0, 0.0), (orc.SCORER_QUEUE, 1.0, 0.0)]
        p_sc = [("token_load", 1.5, thr), ("active_request", 0.7, (max_busy, idle)), ("lora", 2.0, target),
                ("prefix", 1.0, 0), ("queue", 1.0, 0)]
        order = list(range(len(c_sc)))
        rng.shuffle(order)
        k = rng.randint(1, len(order))
        c_sc, p_sc = [c_sc[i] for i in order[:k]], [p_sc[i] for i in order[:k]]
        fk = rng.choice(list(fnames))
        scores, mx, pick, aset = orc.profile_run(orc.make_profile(fk, c_sc), pool, match, total)
        eps = [{"role": labels[r], "kv": kv[i], "waiting": waiting[i], "tokens": tokens[i], "requests": requests[i],
                "active": active[i], "waiting_models": waiting_m[i], "max_active": max_active[i]} for i, r in enumerate(roles)]
        want = pr.profile_run(fnames[fk], p_sc, eps, match, total)
        if want is None:
            assert aset == []
            continue
        acc, wmx, wset = want
        assert mx == wmx and aset == wset and pick == wset[0]
        for i in range(n):
            assert scores[i] == (acc[i] if i in acc else -1.0)


# ------------------------------------------------------------------------------------------------
# max-score picker, first k (picker/maxscore/picker_test.go:30-128) and prefix-cache-affinity-filter
# (filter/prefixcacheaffinity/plugin_test.go:50-118)
# ------------------------------------------------------------------------------------------------
PICKER_KATS = [
    ("Single max score", 1, [10, 25, 15], [1], 0),
    ("Multiple max scores, all are equally scored", 2, [50, 50, 30], [0, 1], 2),
    ("Multiple results sorted by highest score, more pods than needed", 2, [20, 25, 30], [2, 1], 0),
    ("Multiple results sorted by highest score, less pods than needed", 4, [20, 25, 30], [2, 1, 0], 0),
    ("Multiple results sorted by highest score, num of pods exactly needed", 3, [30, 25, 30], [0, 2, 1], 2),
]


@pytest.mark.parametrize("seed", [0, 9, 10, 11])
@pytest.mark.parametrize("name,k,scores,want,n_tie", PICKER_KATS, ids=[c[0] for c in PICKER_KATS])
def test_reference_TestPickMaxScorePicker(orc, name, k, scores, want, n_tie, seed):
    pool = _pool(orc, role=[0, 0, 0], kv=[0, 0, 0], waiting=[0, 0, 0], ext=[[x / 100.0 for x in scores]])
    prof = orc.make_profile(orc.FILTER_NONE, [(orc.SCORER_EXTERNAL, 1.0, 0)])
    picks, sc, ties, _ = orc.profile_run_topk(prof, pool, [0, 0, 0], 0, k, tie_seed=seed, tie_key=12)
    assert len(picks) == len(want) and sc == [scores[e] / 100.0 for e in picks]
    assert sorted(picks[:n_tie]) == sorted(want[:n_tie]) and picks[n_tie:] == want[n_tie:]
    assert ties == max(n_tie, 1)


AFFINITY_KATS = [
    ("AffinityThresholdDisabled", (0.0, 0.0, 0.0), [(0, 10), (90, 20)], 2),
    ("SingleEndpoint", (0.80, 0.0, 0.0), [(90, 10)], 1),
    ("NoStickyEndpoints", (0.80, 0.0, 0.0), [(10, 10), (20, 20), (50, 30)], 3),
    ("NarrowToSticky", (0.80, 0.0, 5000.0), [(90, 100), (85, 120), (10, 50)], 2),
    ("TTFTPenaltyBreaksStickiness", (0.80, 0.0, 100.0), [(90, 500), (10, 50)], 2),
    ("ExplorationProbability", (0.80, 1.0, 0.0), [(90, 100), (10, 50)], 2),
]


@pytest.mark.parametrize("name,cfg,eps,kept", AFFINITY_KATS, ids=[c[0] for c in AFFINITY_KATS])
def test_reference_affinity_filter(orc, name, cfg, eps, kept):
    n = len(eps)
    pool = _pool(orc, role=[0] * n, kv=[0] * n, waiting=[0] * n, ext=[[t for _, t in eps]])
    prof = orc.make_profile(orc.FILTER_NONE, [], affinity=cfg + (0,))
    _, _, ties, _ = orc.profile_run_topk(prof, pool, [m for m, _ in eps], 100, 1, tie_seed=3, tie_key=0)
    assert ties == kept                      # no scorer: every surviving candidate scores 0.0
    got = pr.affinity_filter(cfg, list(range(n)), [{"ttft": t} for _, t in eps], [m for m, _ in eps], 100,
                             explore_draw=0.5)
    assert len(got) == kept


def test_affinity_and_topk_random_vs_python_restatement(orc):
    """C oracle vs the independent restatement: filter outcome, renormalised scores, and the first k endpoints -- the
    oracle's reproducible order must be what the reference's shuffle + stable sort yields for SOME shuffle, namely the
    one that lists the endpoints in the oracle's own order."""
    import random
    rng = random.Random(31)
    labels = {orc.ROLE_NONE: None, orc.ROLE_DECODE: "decode", orc.ROLE_PREFILL: "prefill", orc.ROLE_BOTH: "both"}
    fnames = {orc.FILTER_NONE: "none", orc.FILTER_DECODE: "decode", orc.FILTER_PREFILL: "prefill"}
    kinds = {orc.SCORER_PREFIX: "prefix", orc.SCORER_KV_UTIL: "kv", orc.SCORER_QUEUE: "queue", orc.SCORER_RUNNING: "running"}
    narrowed = explored = 0
    for it in range(300):
        n = rng.randint(1, 30)
        roles = [rng.choice(list(labels)) for _ in range(n)]
        kv = [rng.randrange(4) / 4.0 for _ in range(n)]
        waiting = [rng.choice([0, 0, 3, rng.randint(1, 9)]) for _ in range(n)]
        running = [rng.randint(0, 5) for _ in range(n)]
        ttft = [float(rng.randint(0, 100)) for _ in range(n)]
        requests = [rng.choice([0, 1, rng.randint(2, 9)]) for _ in range(n)]
        total = rng.choice([0, 4, 10])
        match = [rng.choice([0, 0, rng.randint(0, total)]) for _ in range(n)]
        sc = [(rng.choice(list(kinds)), rng.choice([1.0, 2.0, 0.5]), 0) for _ in range(rng.randint(0, 4))]
        use_ar = rng.random() < 0.5
        c_sc = list(sc) + ([(orc.SCORER_ACTIVE_REQUEST, 1.0, 1.0, 1, 1.0)] if use_ar else [])
        p_sc = [(kinds[k], w, p) for k, w, p in sc] + ([("active_request", 1.0, (1.0, 1.0))] if use_ar else [])
        aff = (rng.choice([0.0, 0.3, 0.5, 1.0]), rng.choice([0.0, 0.2, 1.0]), rng.choice([0.0, 20.0, 5000.0]))
        with_ttft = rng.random() < 0.7
        fk = rng.choice(list(fnames))
        seed, key, k = rng.choice([0, 1, 0xFEED]), rng.randrange(1 << 40), rng.randint(1, 6)
        pool = _pool(orc, role=roles, kv=kv, waiting=waiting, running=running, ext=[ttft, requests])
        prof = orc.make_profile(fk, c_sc, affinity=aff + ((0,) if with_ttft else ()))
        picks, pscores, ties, scores = orc.profile_run_topk(prof, pool, match, total, k, tie_seed=seed, tie_key=key)
        draw = orc.explore_u(seed, key) if seed else 1.0         # deterministic mode never explores
        eps = [{"role": labels[r], "kv": kv[i], "waiting": waiting[i], "running": running[i], "requests": requests[i],
                "ttft": ttft[i] if with_ttft else None} for i, r in enumerate(roles)]
        want = pr.profile_run(fnames[fk], p_sc, eps, match, total, affinity=aff, explore_draw=draw)
        if want is None:
            assert ties == 0 and picks == []
            continue
        acc, wmx, wset = want
        plain = pr.profile_run(fnames[fk], p_sc, eps, match, total)
        narrowed += len(acc) < len(plain[0])
        explored += seed != 0 and draw < aff[1]
        assert ties == len(wset) and pscores[0] == wmx and picks[0] in wset
        assert len(picks) == min(k, len(acc)) and len(set(picks)) == len(picks)
        for e, v in zip(picks, pscores):
            assert acc[e] == v
        rest = [i for i in acc if i not in picks]
        assert picks == pr.pick_first_k(acc, list(acc), k, picks + rest)      # a shuffle the reference could have drawn
        assert all(acc[i] <= pscores[-1] for i in rest)
        if seed == 0:                                             # deterministic mode: ascending slots inside a group
            assert picks == sorted(picks, key=lambda i: (-acc[i], i))
    assert narrowed > 30 and explored > 10
