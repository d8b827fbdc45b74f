#!/usr/bin/env python
"""bench_configs.py -- the other BASELINE.json configs (2: Drude bulk ~9k/46k TGNH; 3: EDL with images + field + Langevin;
4: cosine perturbation; 1's topology: non-polarizable NH box; plus the VV scheme, the constraint-bearing split path and
single/double precision), integrator only, on ONE GPU.  These systems are small: the quantity of interest is
microseconds per step and launches per step, ours vs the reference's own CUDA kernels (oracle/_ref, same GPU, same
inputs, OpenMM's launch geometry and its blocking D2H/H2D per thermostat call).  Also times ours replayed from a CUDA
graph (no per-launch CPU cost -- what a 46k-particle production run is bound by).

    python bench_configs.py [--steps 300] > profiles/configs_rNN.json          (bench.py is the contract benchmark)"""
import argparse
import dataclasses
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

EV = 1.60217662e-22


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--big", action="store_true", help="also run 1M / 4M particle boxes in every precision")
    args = ap.parse_args()
    import torch
    vv, vo = entry.load_package(), entry.load_oracle()
    P = vv.Params
    bulk9k, bulk46k = vv.make_bulk_ionic_liquid(250), vv.make_bulk_ionic_liquid(1250)
    edl = vv.make_edl(n_ion_pairs=511, n_electrode=2496, electrode_molecules=4)
    box4k = vv.make_nonpolar_box(512, 8)
    cases = [
        ("config2 bulk_Im21 9,250 TGNH middle", bulk9k, P(max_drude_distance=0.02), "mixed", {}),
        ("config2 bulk 46,250 TGNH middle", bulk46k, P(max_drude_distance=0.02), "mixed", {}),
        ("config2 bulk 46,250 TGNH velocity-Verlet", bulk46k, dataclasses.replace(P(max_drude_distance=0.02), use_middle_scheme=False), "mixed", {}),
        ("config3 EDL 40,310 Langevin+field+images", edl, P(max_drude_distance=0.02, mirror_location=8.0, electric_field=0.25 * EV), "mixed",
         dict(mirror=8.0, n_random=2500 * 400)),
        ("config4 bulk 46,250 cosine perturbation middle", bulk46k, P(max_drude_distance=0.02, cos_acceleration=0.02), "mixed", dict(cos=True)),
        ("config1 topology: non-polarizable 4,096 NH", box4k, P(), "mixed", {}),
        ("config2 bulk 46,250 single precision", bulk46k, P(max_drude_distance=0.02), "single", {}),
        ("config2 bulk 46,250 double precision", bulk46k, P(max_drude_distance=0.02), "double", {}),
    ]
    if args.big:
        for n_ip in (27648, 110592):
            s = vv.make_bulk_ionic_liquid(n_ip)
            for mode in ("mixed", "single", "double"):
                cases.append((f"bulk {s.n} TGNH middle {mode}", s, P(max_drude_distance=0.02), mode, dict(force_sigma=1.0)))
    out = []
    stream = torch.cuda.current_stream()
    for name, spec, params, mode, kw in cases:
        params = params.resolved_for(spec)
        cos = kw.pop("cos", False)
        kw.setdefault("force_sigma", 1.0)
        host = vv.make_state(spec, mode, **kw)
        inv_box_z = 1.0 / host.box[2] if cos else 0.0
        steps = args.steps if spec.n < 500000 else max(20, args.steps // 10)
        req_steps = steps

        plan = vv.Plan(spec, params, mode).upload()
        bufs = vv.DeviceBuffers(host)
        if plan.random_request:
            req_steps = min(steps, host.random.shape[0] // plan.random_request - 8)

        def run(k, start=0):
            ri = start
            for _ in range(k):
                ri = plan.step(bufs, steps=1, random_index=ri, inv_box_z=inv_box_z)
            return ri
        ri = run(5)
        torch.cuda.synchronize()
        l0 = plan.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(req_steps - 5, ri)
        e1.record(stream)
        torch.cuda.synchronize()
        ours_us = 1e3 * e0.elapsed_time(e1) / (req_steps - 5)
        launches = (plan.launch_count - l0) / (req_steps - 5)

        # the same step replayed from a CUDA graph (fixed random index: the arithmetic cost is identical)
        graph_us = None
        try:
            side = torch.cuda.Stream()
            side.wait_stream(stream)
            with torch.cuda.stream(side):
                run(2)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    run(10)
                for _ in range(3):
                    g.replay()
                side.synchronize()
                e0.record(side)
                for _ in range(max(1, req_steps // 10)):
                    g.replay()
                e1.record(side)
                side.synchronize()
            graph_us = 1e3 * e0.elapsed_time(e1) / (10 * max(1, req_steps // 10))
        except Exception as e:  # noqa: BLE001
            graph_us = f"capture failed: {e}"

        # the flow used when OpenMM constraints sit between the sub-steps (both example scripts use HBonds):
        # kick | thermostat_delta | finish = 440 B/particle (mixed); OpenMM's own solver launches are not included
        split_us = split_graph_us = None
        if params.use_middle_scheme:
            pb = vv.Plan(spec, params, mode).upload()
            sb = vv.DeviceBuffers(host, with_pos_delta=True)

            split_ri = [0]

            def run_split(k):
                for _ in range(k):
                    pb.middle_kick(sb, inv_box_z=inv_box_z, random_index=split_ri[0])
                    pb.middle_thermostat_delta(sb, inv_box_z=inv_box_z)
                    pb.middle_finish(sb)
                    if pb.random_request:      # wrap inside the injected stream (timing only)
                        split_ri[0] = (split_ri[0] + pb.random_request) % max(1, host.random.shape[0] - 2 * pb.random_request)
            run_split(5)
            torch.cuda.synchronize()
            e0.record(stream)
            run_split(req_steps - 5)
            e1.record(stream)
            torch.cuda.synchronize()
            split_us = 1e3 * e0.elapsed_time(e1) / (req_steps - 5)
            # ... and replayed from a CUDA graph: three Python / ctypes calls per step cost the host more than the three
            # kernels cost the GPU at these sizes, so the eager figure above is largely the harness's launch rate
            try:
                side = torch.cuda.Stream()
                side.wait_stream(stream)
                with torch.cuda.stream(side):
                    run_split(2)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        run_split(10)
                    for _ in range(3):
                        g.replay()
                    side.synchronize()
                    e0.record(side)
                    for _ in range(max(1, req_steps // 10)):
                        g.replay()
                    e1.record(side)
                    side.synchronize()
                split_graph_us = 1e3 * e0.elapsed_time(e1) / (10 * max(1, req_steps // 10))
                del g
            except Exception as e:  # noqa: BLE001
                split_graph_us = f"capture failed: {e}"
            del pb, sb

        ref_us = ref_launches = None
        if vo.ref_available(mode, gpu=True):
            oracle = vo.Oracle(spec, params, mode, literal=False)
            ref = vo.Reference(oracle, gpu=True)
            rb = vv.DeviceBuffers(host)
            ref.step(rb, steps=5, inv_box_z=inv_box_z)
            torch.cuda.synchronize()
            rsteps = req_steps - 5 if spec.n < 500000 else 5
            e0.record(stream)
            nl = ref.step(rb, steps=rsteps, inv_box_z=inv_box_z)
            e1.record(stream)
            torch.cuda.synchronize()
            ref_us = 1e3 * e0.elapsed_time(e1) / rsteps
            ref_launches = nl / rsteps
        finite = bool(np.isfinite(bufs.to_host().velm).all())
        row = {"config": name, "particles": spec.n, "precision": mode, "ours_us_per_step": ours_us,
               "ours_launches_per_step": launches, "ours_cuda_graph_us_per_step": graph_us,
               "ours_constrained_flow_us_per_step": split_us,
               "ours_constrained_flow_cuda_graph_us_per_step": split_graph_us,
               "reference_kernels_us_per_step": ref_us, "reference_launches_per_step": ref_launches,
               "speedup_vs_reference_kernels": (ref_us / ours_us) if ref_us else None,
               "particle_updates_per_s": spec.n / (ours_us * 1e-6), "finite": finite}
        out.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"gpu": torch.cuda.get_device_name(0), "steps": args.steps, "rows": out}, indent=1))


if __name__ == "__main__":
    main()
