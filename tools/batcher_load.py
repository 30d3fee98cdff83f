#!/usr/bin/env python
"""Closed-loop load test of the micro-batcher (csrc/batcher.cu) with NATIVE threads (tools/loadgen.c): N callers each
schedule one request at a time through epp_submit + epp_wait -- the reference's call shape, one goroutine per in-flight
request -- against the config-3 pool (4 096 endpoints, 4 096-token prompts).  Prints one JSON line per setting:
decisions/s, latency percentiles, mean batch size; every decision is compared with the oracle-checked batch path.

    python tools/batcher_load.py [--threads 1,8,64,256] [--seconds 2] [--max-batch 256] [--delay-us 50] [--index-picks]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libloadgen.so")
SRC = os.path.join(HERE, "loadgen.c")


def build():
    if not os.path.exists(SO) or os.path.getmtime(SRC) > os.path.getmtime(SO):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-pthread", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.loadgen_run.restype = C.c_int64
    L.loadgen_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int64, C.c_int, C.c_double,
                              C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="1,8,64,256")
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--max-batch", type=int, default=256)
    ap.add_argument("--delay-us", type=int, default=50)
    ap.add_argument("--prompts", type=int, default=8192)
    ap.add_argument("--index-picks", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true")
    ap.add_argument("--commit-us", type=int, default=0, help="epp_config.index_commit_interval_us")
    ap.add_argument("--tie-seed", type=int, default=0, help="epp_config.tie_seed (non-zero: random member of the arg-max set)")
    args = ap.parse_args()
    import bench
    import epp_b200 as epp
    from epp_b200 import capi
    from tools import tracegen as tg
    from tools import workload_setup as helpers
    numa_node = None if args.no_numa_bind else bench._bind_to_gpu_numa_node(0)[0]
    lg = build()
    tg.build()
    lib = capi.load()
    w = tg.baseline_configs()["config3"].scaled(R=args.prompts, name="config3")
    trace = tg.Trace(w)
    tokens, _, _ = trace.requests()
    prompts = np.ascontiguousarray(tokens).view(np.uint8).reshape(-1)
    with helpers.make_engine(w, index_commit_interval_us=args.commit_us, tie_seed=args.tie_seed) as eng:
        if args.index_picks:                      # PreRequest needs an incrementally built index: start empty, the picks fill it
            eng.register_model(tg.MODEL)
            role, kv, waiting, running = trace.pool()
            eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
        else:
            helpers.setup_engine(eng, w, trace)
        want, _ = eng.schedule(tokens, uniform_len=w.prompt_bytes, detail=False)      # the batch path (parity-tested)
        submit = C.cast(lib.epp_submit, C.c_void_p)
        wait = C.cast(lib.epp_wait, C.c_void_p)
        for nt in [int(x) for x in args.threads.split(",")]:
            with epp.Batcher(eng, max_batch=args.max_batch, max_delay_us=args.delay_us, index_picks=args.index_picks) as bt:
                cap = int(args.seconds * 60000) + 1000
                lat = np.zeros((nt, cap), dtype=np.float32)
                done = np.zeros(nt, dtype=np.int64)
                picks = np.zeros(w.R, dtype=epp.DECISION_DTYPE)
                err = lg.loadgen_run(submit, wait, bt._b, prompts.ctypes.data, w.prompt_bytes, w.R, nt, args.seconds,
                                     lat.ctypes.data, cap, done.ctypes.data, picks.ctypes.data)
                st = bt.stats()
            l = np.concatenate([lat[t, : min(done[t], cap)] for t in range(nt)])
            if err or l.size == 0:
                print(json.dumps({"threads": nt, "errors": int(err), "message": (lib.epp_batcher_last_error() or b"").decode(),
                                  "engine_message": (lib.epp_last_error() or b"").decode()}), flush=True)
                continue
            seen = picks["total_blocks"] > 0
            same = bool((picks[seen] == want[seen]).all()) if not (args.index_picks or args.tie_seed) else None
            print(json.dumps({"threads": nt, "seconds": args.seconds, "max_batch": args.max_batch, "max_delay_us": args.delay_us,
                              "index_picks": bool(args.index_picks), "index_commit_interval_us": args.commit_us, "tie_seed": args.tie_seed, "errors": int(err),
                              "index_errors": int(st["n_index_errors"]),
                              "decisions_per_s": float(done.sum() / args.seconds),
                              "latency_us": {"p50": float(np.percentile(l, 50)), "p90": float(np.percentile(l, 90)),
                                             "p99": float(np.percentile(l, 99)), "mean": float(l.mean())},
                              "mean_batch": float(st["n_requests"] / max(1, st["n_flushes"])), "flushes": int(st["n_flushes"]),
                              "decisions_equal_batch_path": same, "prompts_seen": int(seen.sum()), "numa_node": numa_node}), flush=True)


if __name__ == "__main__":
    main()
