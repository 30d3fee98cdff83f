"""Where the time of one endpoint-sharded step goes: two shard engines on ONE GPU, stepped phase by phase
(epp_shard_p2p_phase), each phase of rank 0 timed with CUDA events (rank 1 only keeps the protocol moving).
    python tools/experiments/shard_phase_times.py [R]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import epp_b200 as epp
from tools import tracegen as tg, workload_setup as helpers
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import test_sharded as ts

R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
tg.build()
w = tg.baseline_configs()["config5"].scaled(E=8192, R=R, T=4096, name="config5")
trace = tg.Trace(w)
engines = ts._p2p_engines_one_gpu(epp, w, trace, 2, R_max=R)
tokens, _, _ = trace.requests(0, R)
dt = torch.from_numpy(tokens.view(np.int32)).cuda()
outs = [torch.zeros((R, 32), dtype=torch.uint8, device="cuda") for _ in range(2)]
res = {0: [], 1: [], 2: []}
for it in range(6):
    for phase in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        engines[0].shard_p2p_phase(dt, outs[0], phase, uniform_len=w.prompt_bytes)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        engines[1].shard_p2p_phase(dt, outs[1], phase, uniform_len=w.prompt_bytes)
        torch.cuda.synchronize()
        if it >= 2:
            res[phase].append((t1 - t0) * 1e3)
for ph, name in enumerate(("hash + presence masks + signal", "wait + OR + match over the shard + signal", "wait + gather + merge + check")):
    print(f"phase {ph} ({name}): {np.median(res[ph]):.3f} ms (host clock incl. launch + sync)")
st = engines[0].stats()
print("kernel ms of the last hash pass:", [round(x, 4) for x in st["last_kernel_ms"][:4]])
for e in engines:
    e.close()
