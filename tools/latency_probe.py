"""Latency of ONE synchronous epp_schedule call as a function of the batch size (BASELINE config 3 shape: 4 096
endpoints, 4 096-token prompts) -- what the micro-batcher (csrc/batcher.cu) trades against throughput.

The calls are made on the raw C ABI (prebuilt epp_batch struct, pinned input and output buffers), so the numbers are
the library's, not the Python wrapper's:
  host    pinned host prompts in, decisions out on the host (H2D + kernels + D2H inside the call)
  device  prompts and decisions resident in HBM (kernels only)
  kernels the engine's own CUDA-event time around the GPU work of the last host call

    python tools/latency_probe.py [--sizes 1,16,64,256,1024,4096] [--reps 200] > profiles/rN_latency_by_batch_size.json
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,4,16,64,256,1024,4096")
    ap.add_argument("--reps", type=int, default=200)
    args = ap.parse_args()
    sizes = [int(x) for x in args.sizes.split(",")]
    import torch
    import bench
    numa_node, _ = bench._bind_to_gpu_numa_node(0)       # pinned buffers first-touched next to the GPU, like the bench
    import epp_b200 as epp
    from epp_b200 import capi
    from tools import workload_setup as helpers
    from tools import tracegen as tg
    Rmax = max(sizes)
    w = tg.baseline_configs()["config3"].scaled(R=Rmax, name="config3")
    trace = tg.Trace(w)
    lib = capi.load()
    ptr = C.c_void_p()
    nbytes = w.R * w.prompt_bytes
    assert lib.epp_host_alloc(nbytes, C.byref(ptr)) == 0            # pinned, like the batcher's staging buffer
    pinned = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))
    tokens = pinned.view(np.uint32).reshape(w.R, w.T)
    trace.requests(0, w.R, out=tokens)
    optr = C.c_void_p()
    assert lib.epp_host_alloc(32 * Rmax, C.byref(optr)) == 0
    out = {}
    with helpers.make_engine(w) as eng:
        helpers.setup_engine(eng, w, trace)
        dev = torch.from_numpy(tokens.view(np.int32)).cuda()
        ddec = torch.empty((Rmax, 32), dtype=torch.uint8, device="cuda")
        st = capi.Stats()
        for R in sizes:
            row = {}
            for name in ("host", "device"):
                b = capi.Batch()
                b.n_requests = R
                b.uniform_len = w.prompt_bytes
                if name == "host":
                    b.data, dst = ptr.value, optr
                else:
                    b.data, dst, b.flags = dev.data_ptr(), C.c_void_p(ddec.data_ptr()), capi.EPP_BATCH_DEVICE_PTRS
                call = lambda: lib.epp_schedule(eng._h, C.byref(b), dst, None, 0)
                for _ in range(10):
                    assert call() == 0, lib.epp_last_error()
                ts, ks = [], []
                for _ in range(args.reps):
                    t0 = time.perf_counter()
                    call()
                    ts.append(time.perf_counter() - t0)
                    lib.epp_get_stats(eng._h, C.byref(st))
                    ks.append(st.last_kernels_ms)
                ts = np.array(ts) * 1e6
                row[name + "_us_median"] = float(np.median(ts))
                row[name + "_us_p99"] = float(np.percentile(ts, 99))
                row[name + "_gpu_us_median"] = float(np.median(ks) * 1e3)
                row[name + "_launches"] = int(st.last_kernel_launches)
                row[name + "_decisions_per_s"] = R / float(np.median(ts) * 1e-6)
            out[R] = row
    print(json.dumps({"workload": "config3 shape (4096 endpoints, 16 KiB prompts), one synchronous epp_schedule call per "
                      "batch on the raw C ABI, %d calls per size" % args.reps, "numa_node": numa_node, "latency": out}))


if __name__ == "__main__":
    main()
