"""Diagnostic (not collected): small-system step latency, resident single-launch kernel vs the two streaming passes,
eager and replayed from a CUDA graph."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
vv = entry.load_package()
st = torch.cuda.current_stream()
sizes = [int(x) for x in (sys.argv[1:] or [250, 1250, 2500, 4000, 6000])]
for n_ip in sizes:
    spec = vv.make_bulk_ionic_liquid(n_ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    row = [f"N={spec.n:7d}"]
    for mode in (1, 0):
        plan = vv.Plan(spec, params, "mixed").upload()
        plan.set_resident_mode(mode)
        b = vv.DeviceBuffers(host)
        for _ in range(5): plan.step_middle(b)
        torch.cuda.synchronize()
        l0 = plan.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 300
        e0.record(st)
        for _ in range(K): plan.step_middle(b)
        e1.record(st); torch.cuda.synchronize()
        eager = 1e3 * e0.elapsed_time(e1) / K
        lps = (plan.launch_count - l0) / K
        side = torch.cuda.Stream(); side.wait_stream(st)
        with torch.cuda.stream(side):
            plan.step_middle(b)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20): plan.step_middle(b)
            for _ in range(3): g.replay()
            side.synchronize()
            e0.record(side)
            for _ in range(15): g.replay()
            e1.record(side); side.synchronize()
        graph = 1e3 * e0.elapsed_time(e1) / 300
        row.append(f"{'resident' if mode else 'streaming'}: {eager:6.1f} us eager, {graph:6.1f} us graph, {lps:.0f} launch/step (resident steps {plan.resident_launch_count})")
    print(" | ".join(row), flush=True)
