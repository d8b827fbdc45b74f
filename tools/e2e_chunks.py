#!/usr/bin/env python
"""vvb200_step_host (pinned host buffers in and out, one step) at BASELINE config 5 for several VVB200_HOST_CHUNKS:
    python tools/e2e_chunks.py [chunks ...]          ms per step, best of 3 runs of 5 steps"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
import bench  # noqa: E402

vv = entry.load_package()
import torch  # noqa: E402

spec = vv.make_bulk_ionic_liquid(442368)
params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
host = vv.make_state(spec, "mixed", force_sigma=1.0)
pst = bench.pinned_state(vv, host)
for chunks in [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]:
    os.environ["VVB200_HOST_CHUNKS"] = str(chunks)
    plan = vv.Plan(spec, params, "mixed").upload()
    for _ in range(2):
        plan.step_host(pst, steps=1)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            plan.step_host(pst, steps=1)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / 5)
    print(f"VVB200_HOST_CHUNKS={chunks}: {1e3 * best:.2f} ms per step", flush=True)
    del plan
