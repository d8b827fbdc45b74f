"""The OpenMM glue LINKED AND RUN: csrc/glue/CudaVVKernelsB200.cpp (the seven Cuda*Kernel classes forwarding to
libvvb200.so) registered by the reference's unchanged CudaVVKernelFactory.cpp and driven by the reference's unchanged
VVIntegrator::step() -- its own stepMiddle / stepVV issue the virtual calls -- under the mini-OpenMM of oracle/mini_openmm
(device-memory flavour, oracle/_ref/libvvplugin_glue_cuda_<mode>.so).  Real OpenMM is not in the image; this is the closest
executable statement of "drop-in under OpenMM's Context" available here.  Checked:
  - glue == the C ABI called directly (bitwise): the glue adds no arithmetic of its own;
  - glue vs THE REFERENCE PLUGIN ITSELF under the same mini-OpenMM on the same GPU (libvvplugin_ref_cuda: reference host
    code + reference kernels): <= 1e-6 relative, BASELINE.json's bar, for bulk / EDL / cosine, both schemes, with and
    without constraints (the stand-in of oracle/constraint_standin.h plays OpenMM's solvers on both sides);
  - launch counts, the stand-in's call counts, ForceInfo registered before initializeContexts, step-size bookkeeping."""
import dataclasses

import numpy as np
import pytest

from conftest import TOL, rel_err

pytestmark = pytest.mark.gpu
EV = 1.60217662e-22


def need(vo, *flavours):
    for f in flavours:
        if not vo.plugin_available(f, "mixed"):
            pytest.skip(f"oracle/_ref/libvvplugin_{f} not built")


def systems(vv):
    P = vv.Params
    return {
        "bulk": (vv.make_bulk_ionic_liquid(250), P(max_drude_distance=0.02), {}),
        "bulk_constrained": (vv.make_bulk_ionic_liquid(250, hbond_constraints=True, has_cmm=True), P(max_drude_distance=0.02), dict(drude_spread=0.012)),
        "cosine": (vv.make_bulk_ionic_liquid(100), P(max_drude_distance=0.02, cos_acceleration=0.02), {}),
        "cosine_constrained": (vv.make_bulk_ionic_liquid(100, hbond_constraints=True), P(max_drude_distance=0.02, cos_acceleration=0.02), {}),
        "edl": (vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4), P(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * EV),
                dict(n_random=8 * 626, mirror=2.0)),
        "edl_constrained": (vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4, hbond_constraints=True),
                            P(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * EV), dict(n_random=8 * 626, mirror=2.0)),
        "nonpolar": (vv.make_nonpolar_box(512, 8), P(), {}),
    }


@pytest.mark.parametrize("middle", [True, False], ids=["middle", "vv"])
@pytest.mark.parametrize("name", ["bulk", "bulk_constrained", "cosine", "cosine_constrained", "edl", "edl_constrained", "nonpolar"])
def test_glue_under_mini_openmm(vv, vo, name, middle, step_path):
    need(vo, "glue_cuda", "ref_cuda")
    import torch
    mode, steps = "mixed", 3
    spec, params, kw = systems(vv)[name]
    params = dataclasses.replace(params.resolved_for(spec), use_middle_scheme=middle)
    host = vv.make_state(spec, mode, **kw)
    constrained = spec.constraints.shape[0] > 0
    cons = vo.ConstraintStandin(spec, host) if constrained else None
    cos = params.cos_acceleration != 0
    inv_box_z = 1.0 / host.box[2] if cos else 0.0

    # ---- VVIntegrator::step() through the glue ----
    glue = vo.MiniContext(spec, params, mode, "glue_cuda").set_state(host).set_constraints(cons)
    glue.step(steps)
    got = glue.get_state()
    c = glue.counters()
    assert c["step_count"] == steps and c["reference_kernel_launches"] == 0 and c["force_info_before_init"] == 1
    # OpenMM's solvers are only called where they can act: the unconstrained middle step is ONE fused call
    solver_calls = steps if constrained else 0
    assert c["constraint_calls"] == solver_calls and c["velocity_constraint_calls"] == solver_calls and c["reorder_calls"] == steps
    per_step = c["vvb200_launches"] / steps
    if not constrained and step_path == "resident" and not (spec.langevin.size and not middle):
        assert per_step == (1 if middle else 2), f"{per_step} launches per step"     # the whole (half) step is one launch
    print(f"{name} {'middle' if middle else 'vv'} {step_path}: {per_step:.1f} vvb200 launches / step through the glue")

    # ---- the same calls made directly on the C ABI: bitwise ----
    plan = vv.Plan(spec, params, mode).upload()
    bufs = vv.DeviceBuffers(host, with_pos_delta=True)
    if constrained:
        plan.step_constrained(bufs, vo.DeviceStandin(cons, mode), steps=steps, inv_box_z=inv_box_z)
    else:
        plan.step(bufs, steps=steps, inv_box_z=inv_box_z)
    direct = bufs.to_host()
    assert np.array_equal(got.velm, direct.velm) and np.array_equal(got.posq, direct.posq)
    assert np.array_equal(got.corr, direct.corr)
    if spec.n - spec.langevin.size - spec.image_pairs.shape[0] > 0:
        a, b = glue.thermostat_state(), plan.thermostat_state()
        assert np.array_equal(a["vscale"], b["vscale"][: a["num_temp_groups"]])

    # ---- the reference plugin itself (its host code + its kernels) under the same mini-OpenMM, same GPU ----
    ref = vo.MiniContext(spec, params, mode, "ref_cuda").set_state(host).set_constraints(cons)
    ref.step(steps)
    want = ref.get_state()
    torch.cuda.synchronize()
    n = spec.n
    ev, ex = rel_err(got.velm[:n, :3], want.velm[:n, :3]), rel_err(got.positions()[:n], want.positions()[:n])
    rc = ref.counters()
    print(f"    vs the reference plugin: v {ev:.2e} x {ex:.2e}; reference launches / step {rc['reference_kernel_launches'] / steps:.0f}")
    assert ex <= TOL[mode] and ev <= (5e-6 if "edl" in name else TOL[mode])
    assert np.array_equal(got.velm[:, 3], want.velm[:, 3]) and np.array_equal(got.posq[:, 3], want.posq[:, 3])
    if spec.n - spec.langevin.size - spec.image_pairs.shape[0] > 0:
        a, b = glue.thermostat_state(), ref.thermostat_state()
        tk = 1e-10 if cos else 1e-11
        assert rel_err(a["ke2"], b["ke2"]) <= tk and rel_err(a["vscale"], b["vscale"]) <= tk
        assert rel_err(a["eta_dot"], b["eta_dot"]) <= 1e-9
    if cos:
        v1, i1 = glue.viscosity()
        v2, i2 = ref.viscosity()
        assert abs(v1 - v2) <= 1e-10 * max(abs(v2), 1e-3) and abs(i1 - i2) <= 1e-10 * max(abs(i2), 1e-3)


def test_glue_force_info_keeps_plugin_roles_apart(vv, vo, step_path):
    """VVB200ForceInfo (registered BEFORE initializeContexts, where CudaContext::findMoleculeGroups would read it): two
    particles are interchangeable for reorderAtoms only if they carry the same plugin roles"""
    need(vo, "glue_cuda")
    spec = vv.make_edl(n_ion_pairs=4, n_electrode=12, electrode_molecules=2)
    params = vv.Params(mirror_location=1.0, electric_field=0.25 * EV).resolved_for(spec)
    ctx = vo.MiniContext(spec, params, "mixed", "glue_cuda")
    assert ctx.counters()["force_info_before_init"] == 1
    electrode, ion, image = 0, 12, int(spec.image_pairs[0, 0])
    assert ctx.particles_identical(electrode, 1) and ctx.particles_identical(ion, ion + 37)
    assert not ctx.particles_identical(electrode, ion) and not ctx.particles_identical(ion, image)


def test_glue_picks_up_a_step_size_change(vv, vo, step_path):
    need(vo, "glue_cuda", "ref_cuda")
    spec = vv.make_bulk_ionic_liquid(60)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    out = []
    for flavour in ("glue_cuda", "ref_cuda"):
        ctx = vo.MiniContext(spec, params, "mixed", flavour).set_state(host)
        ctx.step(1)
        ctx.set_step_size(0.0005)
        ctx.step(2)
        out.append(ctx.get_state())
    n = spec.n
    assert rel_err(out[0].velm[:n, :3], out[1].velm[:n, :3]) <= TOL["mixed"]
    assert rel_err(out[0].positions()[:n], out[1].positions()[:n]) <= TOL["mixed"]
