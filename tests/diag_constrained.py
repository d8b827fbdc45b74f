"""Diagnostic (not collected): per-call timing of the constrained flow at 16.4M particles."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
vv = entry.load_package()
spec = vv.make_bulk_ionic_liquid(442368)
params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
host = vv.make_state(spec, "mixed", force_sigma=1.0)
plan = vv.Plan(spec, params, "mixed").upload()
b = vv.DeviceBuffers(host, with_pos_delta=True)
st = torch.cuda.current_stream()
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = spec.n
for name, fn, bytes_ in (("kick (kickOnly)", lambda: plan.middle_kick(b), 92),
                         ("thermostat_delta (reduce + scale_delta)", lambda: plan.middle_thermostat_delta(b), 36 + 132),
                         ("finish + hardwall", lambda: plan.middle_finish(b), 192 + 23),
                         ("delta(0)", lambda: plan.middle_delta(b, 0), 96),
                         ("delta(1)", lambda: plan.middle_delta(b, 1), 160),
                         ("thermostat (reduce + scale)", lambda: plan.thermostat(b), 36 + 68),
                         ("vv_positions + hardwall", lambda: plan.vv_positions(b), 160 + 23),
                         ("measure_temperatures (reduce only)", lambda: plan.measure_temperatures(b), 36),
                         ("step_vv_first (reduce + scale/kick/drift)", lambda: plan.step_vv_first(b), 36 + 156),
                         ("step_vv_second (kick+reduce + scale)", lambda: plan.step_vv_second(b), 92 + 68),
                         ("step_middle fused", lambda: plan.step_middle(b), 224)):
    if len(sys.argv) > 1 and not any(a in name for a in sys.argv[1:]):
        continue
    ms = timed(fn)
    print(f"{name:42s} {ms*1e3:8.1f} us  {bytes_ * N / ms / 1e6:8.0f} GB/s (model {bytes_} B/particle)")
