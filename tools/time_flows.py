#!/usr/bin/env python
"""Per-kernel timing of every flow at several sizes (GPU box):   python tools/time_flows.py [n_ion_pairs ...]
For each size: the fused middle step, the velocity-Verlet halves and the constraint-bearing split calls, with the
library's own CUDA events around each pass-A / pass-B launch (vvb200_profile_*), as microseconds and as a fraction of
the measured HBM copy peak for that call's algorithmic bytes."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

vv = entry.load_package()
import torch  # noqa: E402

PEAK = 6555.5
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

# algorithmic bytes per particle (mixed), SURVEY 8(d)
CALLS = [("step_middle A (kick+reduce)", "step_middle", 0, 88), ("step_middle B (scale+drift)", "step_middle", 1, 128),
         ("middle_kick", "middle_kick", 0, 88), ("thermostat A (reduce only)", "thermostat", 0, 32), ("thermostat B (scale only)", "thermostat", 1, 64),
         ("thermostat_delta A (reduce only)", "middle_thermostat_delta", 0, 32), ("thermostat_delta B (scale+deltas)", "middle_thermostat_delta", 1, 128),
         ("middle_finish", "middle_finish", 1, 192),
         ("vv_first A (reduce only)", "step_vv_first", 0, 32), ("vv_first B (scale+kick+drift)", "step_vv_first", 1, 152),
         ("vv_second A (kick+reduce)", "step_vv_second", 0, 88), ("vv_second B (scale only)", "step_vv_second", 1, 64)]


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [27648, 110592, 442368]
    only = os.environ.get("FLOWS")            # comma-separated entry-point names, e.g. FLOWS=thermostat,step_middle
    reps = 20
    for n_ip in sizes:
        spec = vv.make_bulk_ionic_liquid(n_ip)
        params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
        host = vv.make_state(spec, "mixed", force_sigma=1.0)
        plan = vv.Plan(spec, params, "mixed").upload()
        bufs = vv.DeviceBuffers(host, with_pos_delta=True)
        print(f"--- {spec.n} particles ---", flush=True)
        done = {}
        for label, fn, which, bpp in CALLS:
            if only and fn not in only.split(","):
                continue
            if fn not in done:
                call = getattr(plan, fn)
                if fn == "middle_finish":
                    plan.middle_thermostat_delta(bufs)
                for _ in range(3):
                    call(bufs)
                torch.cuda.synchronize()
                plan.profile_enable(reps)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    call(bufs)
                e1.record()
                torch.cuda.synchronize()
                a, b, k = plan.profile_read()
                plan.profile_enable(0)
                # the same call again WITHOUT the per-kernel events (an event between two launches forbids their PDL overlap)
                e0.record()
                for _ in range(5 * reps):
                    call(bufs)
                e1.record()
                torch.cuda.synchronize()
                done[fn] = (1e3 * a / max(k, 1), 1e3 * b / max(k, 1), 1e3 * e0.elapsed_time(e1) / (5 * reps))
            us = done[fn][which]
            gbs = bpp * spec.n / (us * 1e-6) / 1e9 if us > 0 else 0
            print(f"  {label:36s} {us:8.1f} us  {gbs:7.0f} GB/s  {gbs / PEAK:5.3f} of measured peak   (whole call, no events: {done[fn][2]:.1f} us)", flush=True)
        del plan, bufs


if __name__ == "__main__":
    main()
