"""Latency of ONE synchronous epp_schedule call as a function of the batch size (BASELINE config 3 shape: 4 096
endpoints, 4 096-token prompts) -- what the micro-batcher of go/eppcuda/batcher.go trades against throughput.
Host = pinned host prompts through the C ABI (H2D + kernels + D2H), device = prompts already in HBM."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import epp_b200 as epp
    from tools import workload_setup as helpers
    from tools import tracegen as tg
    w = tg.baseline_configs()["config3"].scaled(R=16384, name="config3")
    trace = tg.Trace(w)
    import ctypes as C
    lib = epp.capi.load()
    ptr = C.c_void_p()
    nbytes = w.R * w.prompt_bytes
    assert lib.epp_host_alloc(nbytes, C.byref(ptr)) == 0            # pinned, like the shim's staging buffer
    pinned = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
    tokens = pinned.view(np.uint32).reshape(w.R, w.T)
    trace.requests(0, w.R, out=tokens)
    out = {}
    with helpers.make_engine(w) as eng:
        helpers.setup_engine(eng, w, trace)
        dev = torch.from_numpy(tokens.view(np.int32)).cuda()
        for R in (1, 16, 64, 256, 1024, 4096, 16384):
            row = {}
            for name, data in (("host", tokens[:R]), ("device", dev[:R])):
                for _ in range(5):
                    eng.schedule(data, uniform_len=w.prompt_bytes, detail=False)
                ts = []
                for _ in range(40):
                    t0 = time.perf_counter()
                    eng.schedule(data, uniform_len=w.prompt_bytes, detail=False)
                    if name == "device":
                        torch.cuda.synchronize()
                    ts.append(time.perf_counter() - t0)
                row[name + "_us_median"] = float(np.median(ts) * 1e6)
                row[name + "_decisions_per_s"] = R / float(np.median(ts))
            out[R] = row
    print(json.dumps({"workload": "config3 shape, one synchronous epp_schedule call per batch", "latency": out}))


if __name__ == "__main__":
    main()
