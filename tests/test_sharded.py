"""Endpoint-sharded mode.

CPU (not gpu): world-size-2 gloo run of the exchange protocol (mask all-gather + OR, record all-gather + merge) with
the oracle standing in for each rank's kernels -- proves that the two exchanges reproduce the unsharded decision,
including non-prefix-closed index states where a single all-reduce of best scores would be wrong.
GPU: the same protocol through epp_shard_* on one GPU (two engines = two shards), and on 2 GPUs over NCCL when the
box has them."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _sharded():
    return importlib.import_module("llm-d-inference-scheduler_b200.sharded")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload(E=96, R=160, T=512):
    from tools import tracegen as tg
    return tg.baseline_configs()["config5"].scaled(E=E, R=R, T=T, name="config5-tiny")


def _oracle_rank_outputs(orc, w, trace, tokens, lo, hi):
    """What rank [lo, hi) computes locally, restated with the oracle: presence masks of ITS postings."""
    import helpers
    from tools import tracegen as tg
    pool, ix_full, primary, _, (hs, es) = helpers.setup_oracle(orc, w, trace)
    keep = (es >= lo) & (es < hi)
    ix = orc.Indexer()
    ix.load_pairs(hs[keep], es[keep])
    W = (w.max_prefix_blocks + 31) // 32
    masks = np.zeros((tokens.shape[0], W), dtype=np.uint32)
    hashes = []
    for r in range(tokens.shape[0]):
        h = orc.hash_prompt(tokens[r].tobytes(), tg.MODEL, w.block_size_tokens, w.max_prefix_blocks)
        hashes.append(h)
        for i, x in enumerate(h):
            if ix.get(x):
                masks[r, i // 32] |= np.uint32(1 << (i % 32))
    return pool, ix, primary, masks, hashes


def _oracle_rank_pick(orc, sh, w, pool, ix, primary, hashes, gmasks, lo, hi):
    """Phase 2 restated: global stop from the OR-ed masks, local counts, best among this shard's endpoints."""
    R = len(hashes)
    best = np.zeros(R, dtype=sh.SHARD_BEST_DTYPE)
    for r in range(R):
        h = hashes[r]
        stop = len(h)
        for i in range(len(h)):
            if not (gmasks[r, i // 32] >> (i % 32)) & 1:
                stop = i
                break
        counts = np.zeros(w.E, dtype=np.int32)
        for i in range(stop):
            for s in ix.get(h[i]):
                counts[s] += 1
        # ... but queue min/max span the whole pool (queue.go:79-91): score with the full pool, then mask
        scores, _, _, _ = orc.profile_run(primary, pool, counts, len(h))
        scores = np.where((np.arange(w.E) >= lo) & (np.arange(w.E) < hi), scores, -1.0)
        if (scores >= 0).any():
            mx = scores.max()
            idx = np.flatnonzero(scores == mx)
            best[r] = (mx, idx[0], len(idx), counts[idx[0]], 0)
        else:
            best[r] = (0.0, sh.NO_ENDPOINT, 0, 0, -1)
    return best


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as orc
        from tools import tracegen as tg
        sh = _sharded()
        w = _workload()
        trace = tg.Trace(w, n_threads=1)
        tokens, _, _ = trace.requests()
        lo, hi = sh.shard_range(rank, world, w.E)
        pool, ix, primary, masks, hashes = _oracle_rank_outputs(orc, w, trace, tokens, lo, hi)
        gm = sh.or_allgather(torch.from_numpy(masks.view(np.int32)), dist).numpy().view(np.uint32)
        best = _oracle_rank_pick(orc, sh, w, pool, ix, primary, hashes, gm, lo, hi)
        allb = sh.allgather_records(torch.from_numpy(best.view(np.uint8).reshape(-1, 24)), dist).numpy()
        merged = sh.merge_records_host(allb.reshape(world, -1).view(sh.SHARD_BEST_DTYPE).reshape(world, -1))
        q.put((rank, merged.tobytes(), masks.tobytes(), gm.tobytes()))
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_gloo_world2(orc):
    import torch.multiprocessing as mp
    import helpers
    from tools import tracegen as tg
    sh = _sharded()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, merged, masks, gm = q.get(timeout=240)
        res[rank] = (np.frombuffer(merged, dtype=sh.SHARD_BEST_DTYPE), np.frombuffer(masks, np.uint32), np.frombuffer(gm, np.uint32))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks agree, and agree with the UNSHARDED oracle
    np.testing.assert_array_equal(res[0][0], res[1][0])
    np.testing.assert_array_equal(res[0][2], res[1][2])
    assert (res[0][1] != res[1][1]).any()                # the shards really hold different blocks
    w = _workload()
    trace = tg.Trace(w, n_threads=1)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, _ = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens, 1)
    m = res[0][0]
    np.testing.assert_array_equal(m["status"], odec["status"])
    np.testing.assert_array_equal(m["pick"].astype(np.int64), odec["pick"].astype(np.int64))
    np.testing.assert_array_equal(m["score"].view(np.uint64), odec["score"].view(np.uint64))
    np.testing.assert_array_equal(m["tie_count"].astype(np.int64), odec["tie_count"].astype(np.int64))
    assert (m["match_blocks"] > 0).any()


def test_merge_records_host_rules():
    sh = _sharded()
    a = np.zeros((3, 4), dtype=sh.SHARD_BEST_DTYPE)
    a["status"] = -1
    a[0, 0] = (1.5, 7, 2, 3, 0); a[1, 0] = (1.5, 4, 1, 9, 0); a[2, 0] = (1.0, 1, 5, 0, 0)   # tie: lowest id, ties summed
    a[1, 1] = (0.25, 9, 1, 0, 0)                                                               # single candidate shard
    a[0, 2] = (2.0, 3, 1, 1, 0); a[2, 2] = (3.0, 8, 4, 2, 0)                                   # strictly better later
    m = sh.merge_records_host(a)
    assert tuple(m[0]) == (1.5, 4, 3, 9, 0)
    assert tuple(m[1]) == (0.25, 9, 1, 0, 0)
    assert tuple(m[2]) == (3.0, 8, 4, 2, 0)
    assert m[3]["status"] == -1 and m[3]["pick"] == sh.NO_ENDPOINT


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
def _gpu_sharded_decisions(epp, w, trace, tokens, world, dist=None, rank=None, device=0):
    """Runs the epp_shard_* protocol.  dist None: all `world` shards live on ONE GPU (sequential engines)."""
    import torch
    import helpers
    sh = _sharded()
    role, kv, waiting, running = trace.pool()
    fam = trace.family_tokens()
    dt = torch.from_numpy(tokens.view(np.int32)).cuda(device)
    ranks = [rank] if dist is not None else list(range(world))
    engines, masks = {}, {}
    W = (w.max_prefix_blocks + 31) // 32
    for g in ranks:
        eng = helpers.make_engine(w, device=device)
        eng.register_model(b"synthetic-model")
        lo, hi = sh.shard_range(g, world, w.E)
        eng.shard_set(lo, hi)
        eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        fh, _ = eng.hash_prompts(fam, uniform_len=w.prompt_bytes)
        hs, es = trace.index_pairs(fh)
        keep = (es >= lo) & (es < hi)
        eng.index_load_snapshot(hs[keep], es[keep])        # this rank holds ITS postings only
        m = torch.empty((tokens.shape[0], W), dtype=torch.int32, device=dt.device)
        eng.shard_probe(dt, m, uniform_len=w.prompt_bytes)
        engines[g], masks[g] = eng, m
    if dist is not None:
        gm = sh.or_allgather(masks[rank], dist)
    else:
        gm = masks[0].clone()
        for g in ranks[1:]:
            gm |= masks[g]
    torch.cuda.synchronize()           # torch-produced masks must be complete before the engine reads them
    bests = {}
    for g in ranks:
        b = torch.empty((tokens.shape[0], 24), dtype=torch.uint8, device=dt.device)
        engines[g].shard_pick(tokens.shape[0], gm, b)
        bests[g] = b
    allb = sh.allgather_records(bests[rank], dist) if dist is not None else torch.stack([bests[g] for g in ranks])
    dec = torch.empty((tokens.shape[0], 32), dtype=torch.uint8, device=dt.device)
    torch.cuda.synchronize()
    engines[ranks[0]].shard_merge(tokens.shape[0], allb.shape[0], allb, dec)
    torch.cuda.synchronize()
    out = epp.decisions_from_torch(dec)
    for e in engines.values():
        e.close()
    return out, [m.cpu().numpy() for m in masks.values()]


@pytest.mark.gpu
def test_sharded_two_shards_one_gpu(orc):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    w = _workload(E=512, R=384, T=1024)
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens)
    for world in (2, 4):
        dec, masks = _gpu_sharded_decisions(epp, w, trace, tokens, world)
        helpers.assert_decisions_equal(dec, None, odec, ototal, where=f"sharded x{world}")
        assert any((masks[0] != m).any() for m in masks[1:])


def _p2p_engines_one_gpu(epp, w, trace, world, R_max):
    """`world` shard engines on ONE GPU wired to each other's exchange buffers by raw pointers (same process)."""
    import helpers
    sh = _sharded()
    role, kv, waiting, running = trace.pool()
    fam = trace.family_tokens()
    engines, ptrs = [], []
    for g in range(world):
        eng = helpers.make_engine(w)
        eng.register_model(b"synthetic-model")
        lo, hi = sh.shard_range(g, world, w.E)
        eng.shard_set(lo, hi)
        eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        fh, _ = eng.hash_prompts(fam, uniform_len=w.prompt_bytes)
        hs, es = trace.index_pairs(fh)
        keep = (es >= lo) & (es < hi)
        eng.index_load_snapshot(hs[keep], es[keep])
        _, ptr = eng.shard_p2p_export(R_max)
        engines.append(eng)
        ptrs.append(ptr)
    for g, eng in enumerate(engines):
        eng.shard_p2p_connect(world, g, np.array(ptrs, dtype=np.uint64), ipc_handles=False)
    return engines


@pytest.mark.gpu
def test_sharded_p2p_exchange_one_gpu(orc, monkeypatch):
    """The peer-memory exchange (flags, OR of the masks, gather + merge of the records) with THREE shard engines on one
    GPU, stepped phase by phase from one host thread (epp_shard_p2p_phase: ranks that share a GPU must not spin-wait
    for each other on the device): three consecutive batches (buffer reuse across epochs), every rank's decisions
    identical and equal to the oracle's.  Same kernels, same buffers, same flags as epp_shard_schedule_p2p."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    w = _workload(E=510, R=384, T=1024)
    trace = tg.Trace(w)
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    world = 3
    engines = _p2p_engines_one_gpu(epp, w, trace, world, R_max=512)
    try:
        for b in range(3):
            tokens, _, _ = trace.requests(b * w.R, w.R)
            odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens)
            dt = torch.from_numpy(tokens.view(np.int32)).cuda()
            outs = [torch.zeros((w.R, 32), dtype=torch.uint8, device="cuda") for _ in range(world)]
            torch.cuda.synchronize()
            for phase in range(3):
                for g in range(world):
                    engines[g].shard_p2p_phase(dt, outs[g], phase, uniform_len=w.prompt_bytes)
            decs = [epp.decisions_from_torch(o) for o in outs]
            for d in decs[1:]:
                np.testing.assert_array_equal(d, decs[0])
            helpers.assert_decisions_equal(decs[0], None, odec, ototal, where=f"p2p sharded x{world}, batch {b}")
            assert (decs[0]["match_blocks"] > 0).any()
    finally:
        for e in engines:
            e.close()


@pytest.mark.gpu
def test_sharded_p2p_missing_peer_breaks_the_exchange(orc, monkeypatch):
    """A peer that never shows up: the waiting rank gets EPP_ERR_NCCL after the time-out instead of hanging, poisons its
    flags, and every later call on it -- and on a peer that meets the poison -- fails until the buffers are exported and
    connected again; a rank without a shard range is rejected."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    monkeypatch.setenv("EPP_P2P_TIMEOUT_MS", "200")
    w = _workload(E=128, R=64, T=256)
    trace = tg.Trace(w)
    engines = _p2p_engines_one_gpu(epp, w, trace, 2, R_max=64)
    try:
        tokens, _, _ = trace.requests(0, w.R)
        dt = torch.from_numpy(tokens.view(np.int32)).cuda()
        out = torch.zeros((w.R, 32), dtype=torch.uint8, device="cuda")
        with pytest.raises(epp.EngineError, match="did not reach batch"):
            engines[0].shard_schedule_p2p(dt, out, uniform_len=w.prompt_bytes)      # rank 1 never calls
        with pytest.raises(epp.EngineError, match="broken"):
            engines[0].shard_schedule_p2p(dt, out, uniform_len=w.prompt_bytes)      # sticky
        with pytest.raises(epp.EngineError, match="broken sharded exchange|did not reach"):
            engines[1].shard_schedule_p2p(dt, out, uniform_len=w.prompt_bytes)      # meets rank 0's poisoned flags
    finally:
        for e in engines:
            e.close()
    with helpers.make_engine(w) as eng:
        eng.register_model(b"synthetic-model")
        role, kv, waiting, running = trace.pool()
        eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        _, ptr = eng.shard_p2p_export(64)
        eng.shard_p2p_connect(1, 0, np.array([ptr], dtype=np.uint64), ipc_handles=False)
        tokens, _, _ = trace.requests(0, w.R)
        with pytest.raises(epp.EngineError, match="epp_shard_set"):
            eng.shard_schedule_p2p(torch.from_numpy(tokens.view(np.int32)).cuda(), torch.zeros((w.R, 32), dtype=torch.uint8, device="cuda"),
                                   uniform_len=w.prompt_bytes)


def _p2p_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import epp_b200 as epp
        import helpers
        from tools import tracegen as tg
        sh = _sharded()
        w = _workload(E=512, R=384, T=1024)
        trace = tg.Trace(w, n_threads=2)
        role, kv, waiting, running = trace.pool()
        eng = helpers.make_engine(w, device=rank)
        eng.register_model(b"synthetic-model")
        lo, hi = sh.shard_range(rank, world, w.E)
        eng.shard_set(lo, hi)
        eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        fh, _ = eng.hash_prompts(trace.family_tokens(), uniform_len=w.prompt_bytes)
        hs, es = trace.index_pairs(fh)
        keep = (es >= lo) & (es < hi)
        eng.index_load_snapshot(hs[keep], es[keep])
        sh.connect_p2p(eng, 512, dist)                     # CUDA IPC handles, the only collective
        res = []
        for b in range(2):
            tokens, _, _ = trace.requests(b * w.R, w.R)
            dt = torch.from_numpy(tokens.view(np.int32)).cuda(rank)
            dec = sh.schedule_sharded_p2p(eng, dt, w.prompt_bytes)
            torch.cuda.synchronize()
            res.append(epp.decisions_from_torch(dec).tobytes())
        q.put((rank, res))
        dist.barrier()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_two_gpus_p2p(orc):
    """Two ranks on two GPUs: exchange buffers opened through CUDA IPC, masks and records read over NVLink."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = _workload(E=512, R=384, T=1024)
    trace = tg.Trace(w)
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    for b in range(2):
        tokens, _, _ = trace.requests(b * w.R, w.R)
        odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens)
        d0 = np.frombuffer(res[0][b], dtype=epp.DECISION_DTYPE)
        d1 = np.frombuffer(res[1][b], dtype=epp.DECISION_DTYPE)
        np.testing.assert_array_equal(d0, d1)
        helpers.assert_decisions_equal(d0, None, odec, ototal, where=f"sharded p2p x2, batch {b}")


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import epp_b200 as epp
        from tools import tracegen as tg
        w = _workload(E=512, R=384, T=1024)
        trace = tg.Trace(w, n_threads=2)
        tokens, _, _ = trace.requests()
        dec, _ = _gpu_sharded_decisions(epp, w, trace, tokens, world, dist=dist, rank=rank, device=rank)
        q.put((rank, dec.tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_two_gpus_nccl(orc):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    w = _workload(E=512, R=384, T=1024)
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, None, tokens)
    d0 = np.frombuffer(res[0], dtype=epp.DECISION_DTYPE)
    d1 = np.frombuffer(res[1], dtype=epp.DECISION_DTYPE)
    np.testing.assert_array_equal(d0, d1)
    helpers.assert_decisions_equal(d0, None, odec, ototal, where="sharded nccl x2")


@pytest.mark.gpu
def test_two_engines_on_two_devices_in_one_process(orc):
    """Replicas inside ONE process (what the Go shim does: one engine per GPU): per-device state of the launchers
    (occupancy caches, opt-in shared-memory attributes of the dense-counter and the small-batch kernel) must not leak
    from device 0 to device 1.  Host batches (single-launch path), device batches (throughput kernels) and the
    dense-counter kernel (Produce rows) on both devices, decisions equal to the oracle's."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import epp_b200 as epp
    import helpers
    from tools import tracegen as tg
    tg.build()
    w = tg.baseline_configs()["config3"].scaled(E=4096, R=2048, T=1024, name="config3")
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    pool, ix, primary, prefill, _ = helpers.setup_oracle(orc, w, trace)
    odec, ototal = helpers.oracle_decisions(orc, w, pool, ix, primary, prefill, tokens)
    engines = []
    try:
        for dev in (0, 1):
            eng = helpers.make_engine(w, device=dev)
            helpers.setup_engine(eng, w, trace)
            engines.append(eng)
        for dev, eng in enumerate(engines):
            dec, det = eng.schedule(tokens[:512], uniform_len=w.prompt_bytes)                 # one launch: k_cycle_small
            helpers.assert_decisions_equal(dec, det, odec[:512], ototal[:512], where=f"device {dev}, small host batch")
            dec, det = eng.schedule(tokens, uniform_len=w.prompt_bytes)                       # chunked host path
            helpers.assert_decisions_equal(dec, det, odec, ototal, where=f"device {dev}, host batch")
            with torch.cuda.device(dev):
                dt = torch.from_numpy(tokens.view(np.int32)).to(f"cuda:{dev}")
                ddec, ddet = eng.schedule(dt, uniform_len=w.prompt_bytes)
                torch.cuda.synchronize(dev)
                helpers.assert_decisions_equal(epp.decisions_from_torch(ddec), None, odec, ototal, where=f"device {dev}, device batch")
            match, total = eng.prefix_match(tokens[:64], uniform_len=w.prompt_bytes)          # dense-counter kernel, 69 KiB of smem
            np.testing.assert_array_equal(total, ototal[:64])
            assert (match.max(axis=1) >= 0).all()
    finally:
        for e in engines:
            e.close()
