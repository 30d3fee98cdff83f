"""Endpoint-sharded mode: host-side protocol around the epp_shard_* entry points.

Rank g holds the postings of endpoints [begin_g, end_g) and the FULL pool state.  Because the reference's walk stops
at the first block NO server holds (approximateprefix/plugin.go:219-223), a shard cannot finish alone:

    phase 1  epp_shard_probe   local block-presence masks            [R][W] u32,  W = ceil(max_prefix_blocks / 32)
    exchange all-gather + bitwise OR of the masks (NCCL has no OR reduction)
    phase 2  epp_shard_pick    global stop applied to local counts -> local best record   [R] x 24 B
    exchange all-gather of the records
    phase 3  epp_shard_merge   max score, lowest slot id among equals, ties summed

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the Go shim uses ncclAllGather directly.
"""
from __future__ import annotations

import numpy as np

SHARD_BEST_DTYPE = np.dtype([("score", "<f8"), ("pick", "<u4"), ("tie_count", "<u4"), ("match_blocks", "<i4"),
                             ("status", "<i4")])
NO_ENDPOINT = 0xFFFFFFFF


def shard_range(rank: int, world: int, n_endpoints: int) -> tuple[int, int]:
    per = (n_endpoints + world - 1) // world
    return min(n_endpoints, rank * per), min(n_endpoints, (rank + 1) * per)


def or_allgather(local_masks, dist=None):
    """All-gather the per-rank masks and OR them.  local_masks: torch int32 tensor [R, W] (CPU or CUDA)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_masks.clone()
    world = dist.get_world_size()
    gathered = torch.empty((world,) + tuple(local_masks.shape), dtype=local_masks.dtype, device=local_masks.device)
    dist.all_gather_into_tensor(gathered.view(-1), local_masks.contiguous().view(-1))
    out = gathered[0].clone()
    for g in range(1, world):
        out |= gathered[g]
    return out


def allgather_records(local_best, dist=None):
    """All-gather the 24-byte best records.  local_best: torch uint8 tensor [R, 24] -> [world, R, 24]."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local_best.unsqueeze(0).clone()
    world = dist.get_world_size()
    gathered = torch.empty((world,) + tuple(local_best.shape), dtype=local_best.dtype, device=local_best.device)
    dist.all_gather_into_tensor(gathered.view(-1), local_best.contiguous().view(-1))
    return gathered


def merge_records_host(all_best: np.ndarray) -> np.ndarray:
    """Reference semantics of epp_shard_merge on the host (numpy), used by the CPU protocol tests only:
    all_best [world, R] of SHARD_BEST_DTYPE -> (status, pick, score, tie_count, match_blocks) arrays."""
    world, R = all_best.shape
    out = np.zeros(R, dtype=SHARD_BEST_DTYPE)
    out["status"] = -1
    out["pick"] = NO_ENDPOINT
    for g in range(world):
        rec = all_best[g]
        ok = rec["status"] == 0
        have = out["status"] == 0
        better = ok & (~have | (rec["score"] > out["score"]))
        equal = ok & have & (rec["score"] == out["score"])
        # ties: sum counts, keep the lowest slot id (and its match count)
        lower = equal & (rec["pick"] < out["pick"])
        out["tie_count"] = np.where(better, rec["tie_count"], np.where(equal, out["tie_count"] + rec["tie_count"], out["tie_count"]))
        take = better | lower
        out["pick"] = np.where(take, rec["pick"], out["pick"])
        out["match_blocks"] = np.where(take, rec["match_blocks"], out["match_blocks"])
        out["score"] = np.where(better, rec["score"], out["score"])
        out["status"] = np.where(ok, 0, out["status"])
    return out


def schedule_sharded(engine, tokens_dev, uniform_len: int, dist=None):
    """One batch through the sharded protocol on this rank's engine.  tokens_dev: torch CUDA tensor [R, T].
    Returns a [R, 32] uint8 CUDA tensor of epp_decision records (identical on every rank)."""
    import torch
    R = tokens_dev.shape[0]
    W = (engine.B + 31) // 32
    dev = tokens_dev.device
    masks = torch.empty((R, W), dtype=torch.int32, device=dev)
    engine.shard_probe(tokens_dev, masks, uniform_len=uniform_len)
    gmasks = or_allgather(masks, dist)
    best = torch.empty((R, 24), dtype=torch.uint8, device=dev)
    # The engine runs on its own streams and only guarantees that ITS outputs are complete on return; tensors produced
    # by torch / NCCL on the caller's stream must be complete before they are handed over (epp_engine.h, "streams").
    torch.cuda.current_stream(dev).synchronize()
    engine.shard_pick(R, gmasks, best)
    allb = allgather_records(best, dist)
    dec = torch.empty((R, 32), dtype=torch.uint8, device=dev)
    torch.cuda.current_stream(dev).synchronize()
    engine.shard_merge(R, allb.shape[0], allb, dec)
    return dec


# ---- the same protocol with the exchanges over NVLink peer memory (csrc/shard_p2p.cu): no NCCL on the data path ----
def connect_p2p(engine, max_requests: int, dist=None):
    """Export this rank's exchange buffer, all-gather the 64-byte CUDA IPC handles (the only collective, once) and open
    every peer's buffer."""
    import torch
    handle, _ = engine.shard_p2p_export(max_requests)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        engine.shard_p2p_connect(1, 0, handle.reshape(1, 64), ipc_handles=True)
        return
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    mine = torch.from_numpy(handle).to(dev)
    allh = torch.empty((world, 64), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(allh.view(-1), mine)
    engine.shard_p2p_connect(world, rank, allh.cpu().numpy(), ipc_handles=True)


def schedule_sharded_p2p(engine, tokens_dev, uniform_len: int, out=None):
    """One batch; every rank calls it with the same tokens.  Returns a [R, 32] uint8 CUDA tensor of epp_decision."""
    import torch
    R = tokens_dev.shape[0]
    dec = out if out is not None else torch.empty((R, 32), dtype=torch.uint8, device=tokens_dev.device)
    engine.shard_schedule_p2p(tokens_dev, dec, uniform_len=uniform_len)
    return dec
