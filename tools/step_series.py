#!/usr/bin/env python
"""ms per step of vvb200_step_middle over successive windows of K steps (does the step time drift with run length?).
    [VVB200_LIB=...] python tools/step_series.py [ion pairs] [windows ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
vv = entry.load_package()
import torch
n_ip = int(sys.argv[1]) if len(sys.argv) > 1 else 442368
windows = [int(a) for a in sys.argv[2:]] or [20, 20, 100, 100, 300, 20]
spec = vv.make_bulk_ionic_liquid(n_ip)
params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
host = vv.make_state(spec, "mixed", force_sigma=1.0)
plan = vv.Plan(spec, params, "mixed").upload()
b = vv.DeviceBuffers(host)
for _ in range(5): plan.step_middle(b)
torch.cuda.synchronize()
out = []
for K in windows:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): plan.step_middle(b)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{K}: {1e3 * e0.elapsed_time(e1) / K:.1f}")
print(os.environ.get("VVB200_LIB", "default"), spec.n, "us/step per window ->", "  ".join(out), flush=True)
# per-kernel times now (late state), then after reloading the initial state (early state)
def prof(tag, k=20):
    plan.profile_enable(k)
    for _ in range(k): plan.step_middle(b)
    torch.cuda.synchronize()
    a, bb, n = plan.profile_read(); plan.profile_enable(0)
    print(f"   {tag}: pass A {1e3 * a / n:.1f} us, pass B {1e3 * bb / n:.1f} us", flush=True)
prof("late state ")
b.load(host); torch.cuda.synchronize()
prof("reloaded   ")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b.load(host); torch.cuda.synchronize()
for _ in range(5): plan.step_middle(b)
e0.record()
for _ in range(20): plan.step_middle(b)
e1.record(); torch.cuda.synchronize()
print(f"   reloaded, 20 steps: {1e3 * e0.elapsed_time(e1) / 20:.1f} us/step", flush=True)
