"""Cost of PreRequest (epp_index_add_picked) + the commit it causes for SMALL batches: config-3 pool, batches of 32 / 256
requests through epp_schedule(keep_hashes) + epp_index_add_picked + the next schedule.  EPP_STORE_TIMING=1 prints the
store's per-kernel times."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import epp_b200 as epp
from tools import tracegen as tg, workload_setup as helpers
tg.build()
w = tg.baseline_configs()["config3"].scaled(R=8192, name="config3")
trace = tg.Trace(w)
tokens, _, _ = trace.requests()
buf = epp.PinnedBuffer(tokens.nbytes)
buf.array[:] = tokens.view(np.uint8).reshape(-1)
ptok = buf.array.view(np.uint32).reshape(tokens.shape)
with helpers.make_engine(w, tie_seed=int(os.environ.get("TIE_SEED", "0"))) as eng:
    eng.register_model(tg.MODEL)
    role, kv, waiting, running = trace.pool()
    eng.pool_set(np.arange(w.E, dtype=np.uint32), role, kv, waiting, running)
    for n in (32, 256):
        ts = {"schedule": [], "add_picked": [], "commit": []}
        for it in range(12):
            rows = ptok[(it * n) % 8192: (it * n) % 8192 + n]
            t0 = time.perf_counter(); eng.schedule(rows, uniform_len=w.prompt_bytes, keep_hashes=True, detail=False)
            t1 = time.perf_counter(); eng.index_add_picked()
            t2 = time.perf_counter(); eng.index_commit()
            t3 = time.perf_counter()
            if it >= 4:
                ts["schedule"].append(t1 - t0); ts["add_picked"].append(t2 - t1); ts["commit"].append(t3 - t2)
        st = eng.stats()
        print(n, {k: round(float(np.median(v)) * 1e6, 1) for k, v in ts.items()}, "us; patched:", st["last_index_patched"], "pairs:", st["index_pairs"], file=sys.stderr)
