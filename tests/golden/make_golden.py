#!/usr/bin/env python
"""Generates tests/golden/*.npz: small input/output vectors of the integration path produced by the REFERENCE'S
OWN kernel sources (platforms/cuda/src/kernels/*.cu under /root/reference) compiled for the host by oracle/Makefile
(oracle/_ref/libvvref_cpu_<mode>.so) and driven by the restated host schedule of oracle/ref_harness.cpp.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures are committed; tests/test_golden.py checks the CPU oracle (CPU suite) and the CUDA path (-m gpu)
against them.  Everything a consumer needs is inside each file: topology arrays, parameters, the state before,
the state after `steps` steps and the thermostat state."""
import dataclasses
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

vv = entry.load_package()
vo = entry.load_oracle()
EV = 1.60217662e-22


def cases():
    bulk = vv.make_bulk_ionic_liquid(10)
    P = vv.Params
    yield "bulk_tgnh_middle_mixed", bulk, P(max_drude_distance=0.02).resolved_for(bulk), "mixed", 4, {}
    yield "bulk_tgnh_vv_double", bulk, dataclasses.replace(P(max_drude_distance=0.02).resolved_for(bulk),
                                                           use_middle_scheme=False), "double", 3, {}
    yield "bulk_tgnh_middle_single", bulk, P().resolved_for(bulk), "single", 2, {}
    yield "bulk_hardwall_fires_mixed", bulk, P(max_drude_distance=0.02).resolved_for(bulk), "mixed", 2, dict(drude_spread=0.015)
    yield "bulk_cosine_middle_mixed", bulk, P(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(bulk), "mixed", 3, dict(cos=True)
    box = vv.make_nonpolar_box(32, 8)
    yield "nonpolar_nh_mixed", box, P().resolved_for(box), "mixed", 4, {}
    edl = vv.make_edl(n_ion_pairs=6, n_electrode=60, electrode_molecules=3)
    yield "edl_langevin_field_image_mixed", edl, P(max_drude_distance=0.02, mirror_location=1.0,
                                                   electric_field=0.25 * EV).resolved_for(edl), "mixed", 3, dict(mirror=1.0, n_random=4 * 62)
    # added later (python tests/golden/make_golden.py polymer edl_vv bulk_tgnh_middle_double): a thermostat molecule
    # longer than a tile (the fused path cuts it and sums the fragments), the EDL system under the VV scheme, double
    poly = vv.make_polymer(1, 330, 6, has_cmm=True, adjacent=True)
    yield "polymer_cut_molecule_mixed", poly, P(max_drude_distance=0.02).resolved_for(poly), "mixed", 3, {}
    yield "polymer_cut_molecule_cosine_mixed", poly, P(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(poly), "mixed", 3, dict(cos=True)
    yield "edl_vv_mixed", edl, dataclasses.replace(P(max_drude_distance=0.02, mirror_location=1.0, electric_field=0.25 * EV).resolved_for(edl),
                                                   use_middle_scheme=False), "mixed", 3, dict(mirror=1.0, n_random=4 * 62)
    yield "bulk_tgnh_middle_double", bulk, P(max_drude_distance=0.02).resolved_for(bulk), "double", 3, {}
    rag = vv.make_ragged(seed=1, n_molecules=24, max_size=20, scattered_molecules=0)
    yield "ragged_mixed", rag, P(max_drude_distance=0.02, mirror_location=1.0, electric_field=1e-22).resolved_for(rag), "mixed", 3, dict(mirror=1.0, n_random=8 * rag.n)
    # round 2 (python tests/golden/make_golden.py constrained): the constraint-bearing flow.  OpenMM's solvers between the
    # sub-steps are the stand-in of oracle/constraint_standin.h (cluster tables stored in the fixture), applied by the
    # harness where the reference calls integration.applyVelocityConstraints / applyConstraints.
    cbulk = vv.make_bulk_ionic_liquid(10, hbond_constraints=True)
    yield "constrained_bulk_middle_mixed", cbulk, P(max_drude_distance=0.02).resolved_for(cbulk), "mixed", 3, dict(constrained=True)
    yield "constrained_bulk_vv_mixed", cbulk, dataclasses.replace(P(max_drude_distance=0.02).resolved_for(cbulk), use_middle_scheme=False), "mixed", 3, dict(constrained=True)
    yield "constrained_bulk_hardwall_fires_mixed", cbulk, P(max_drude_distance=0.02).resolved_for(cbulk), "mixed", 2, dict(constrained=True, drude_spread=0.015)
    yield "constrained_bulk_cosine_middle_mixed", cbulk, P(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(cbulk), "mixed", 3, dict(constrained=True, cos=True)
    yield "constrained_bulk_middle_double", cbulk, P(max_drude_distance=0.02).resolved_for(cbulk), "double", 3, dict(constrained=True)
    cedl = vv.make_edl(n_ion_pairs=6, n_electrode=60, electrode_molecules=3, hbond_constraints=True)
    yield "constrained_edl_middle_mixed", cedl, P(max_drude_distance=0.02, mirror_location=1.0, electric_field=0.25 * EV).resolved_for(cedl), "mixed", 3, dict(constrained=True, mirror=1.0, n_random=4 * 62)
    yield "constrained_edl_vv_mixed", cedl, dataclasses.replace(P(max_drude_distance=0.02, mirror_location=1.0, electric_field=0.25 * EV).resolved_for(cedl),
                                                                use_middle_scheme=False), "mixed", 3, dict(constrained=True, mirror=1.0, n_random=4 * 62)


def main():
    only = sys.argv[1:]                 # name prefixes; existing fixtures are left untouched unless named
    for name, spec, params, mode, steps, kw in cases():
        if only and not any(name.startswith(o) for o in only):
            continue
        kw = dict(kw)
        cos = kw.pop("cos", False)
        constrained = kw.pop("constrained", False)
        host = vv.make_state(spec, mode, **kw)
        inv_box_z = 1.0 / host.box[2] if cos else 0.0
        oracle = vo.Oracle(spec, params, mode, literal=True)          # only supplies the index arrays
        ref = vo.Reference(oracle, gpu=False)
        cons = vo.ConstraintStandin(spec, host) if constrained else None
        if cons is not None:
            ref.set_constraints(cons)
        after = host.copy()
        ref.step(after, steps=steps, inv_box_z=inv_box_z)
        st = ref.thermostat_state()
        out = dict(
            mode=mode, steps=steps, inv_box_z=inv_box_z, box=np.array(host.box),
            params=np.array([getattr(params, f.name) for f in dataclasses.fields(params)], dtype=np.float64),
            param_names=np.array([f.name for f in dataclasses.fields(params)]),
            n=spec.n, masses=spec.masses, mol_id=spec.mol_id, drude_pairs=spec.drude_pairs, constraints=spec.constraints,
            has_cmm=spec.has_cmm, langevin=spec.langevin, image_pairs=spec.image_pairs, electrolyte=spec.electrolyte,
            posq0=host.posq, velm0=host.velm, force=host.force, random=host.random,
            posq1=after.posq, velm1=after.velm,
            ke2=st["ke2"], vscale=st["vscale"], velocity_bias=st["velocity_bias"], eta_dot=st["eta_dot"], eta=st["eta"])
        if host.corr is not None:
            out["corr0"], out["corr1"] = host.corr, after.corr
        if cons is not None:
            out.update(cons_offset=cons.offset, cons_atoms=cons.atoms, cons_distance=cons.distance,
                       cons_iterations=cons.iterations)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: N={spec.n} {mode} {steps} steps -> {os.path.getsize(path)} bytes")


if __name__ == "__main__":
    assert vo.ref_available("mixed"), "oracle/_ref not built: run `make -C oracle ref` with /root/reference present"
    main()
