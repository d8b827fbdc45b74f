"""compute-sanitizer driver (not collected by pytest): runs every kernel family once on small systems.
    compute-sanitizer --tool memcheck python tests/sanitize_gpu.py
    compute-sanitizer --tool racecheck python tests/sanitize_gpu.py"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

vv = entry.load_package()
vo = entry.load_oracle()        # the constraint stand-in (test infrastructure) between the split entry points
EV = 1.60217662e-22
import torch  # noqa: E402

P = vv.Params
bulk = vv.make_bulk_ionic_liquid(40, hbond_constraints=True)
edl = vv.make_edl(n_ion_pairs=12, n_electrode=120, electrode_molecules=3, hbond_constraints=True)
poly = vv.make_polymer(1, 600, 10)
poly_adj = vv.make_polymer(2, 500, 10, adjacent=True)       # molecules cut across tiles on the fused path
cases = [(bulk, P(max_drude_distance=0.02), {}), (bulk, P(max_drude_distance=0.02, cos_acceleration=0.02), dict(cos=True)),
         (edl, P(max_drude_distance=0.02, mirror_location=1.2, electric_field=0.25 * EV), dict(n_random=4 * 122 * 4, mirror=1.2)),
         (poly, P(max_drude_distance=0.02), {}), (poly_adj, P(max_drude_distance=0.02), {}),
         (poly_adj, P(max_drude_distance=0.02, cos_acceleration=0.02), dict(cos=True))]
for spec, params, kw in cases:
    for middle in (True, False):
        for mode in ("mixed", "single", "double"):
            if mode != "mixed" and spec.image_pairs.size:
                continue
            kw2 = dict(kw)
            cos = kw2.pop("cos", False)
            par = dataclasses.replace(params.resolved_for(spec), use_middle_scheme=middle)
            host = vv.make_state(spec, mode, **kw2)
            for resident in (1, 0):          # single-launch resident kernel, then the two streaming passes
                plan = vv.Plan(spec, par, mode).upload()
                plan.set_resident_mode(resident)
                bufs = vv.DeviceBuffers(host, with_pos_delta=True)
                plan.step(bufs, steps=2, inv_box_z=1.0 / host.box[2] if cos else 0.0)
                # the constraint-bearing flow (both schemes, Langevin / image systems included) with a live stand-in
                ibz = 1.0 / host.box[2] if cos else 0.0
                solver = vo.DeviceStandin(vo.ConstraintStandin(spec, host), mode)
                ri = plan.step_constrained(bufs, solver, steps=2, random_index=2 * plan.random_request, inv_box_z=ibz)
                if middle:
                    plan.middle_kick(bufs, random_index=ri, inv_box_z=ibz); solver.apply_velocity_constraints(bufs)
                    plan.middle_delta(bufs, 0); plan.thermostat(bufs, inv_box_z=ibz); plan.middle_delta(bufs, 1)
                    solver.apply_constraints(bufs); plan.middle_finish(bufs)
                torch.cuda.synchronize()
                assert bool(torch.isfinite(bufs.velm).all()), "non-finite velocities"
                print("ok", spec.name, mode, "middle" if middle else "vv", "tiled" if plan.tiled else "general",
                      f"resident launches {plan.resident_launch_count}", flush=True)
print("sanitize driver done")
