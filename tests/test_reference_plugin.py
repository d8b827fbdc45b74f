"""The reference plugin's OWN host code, compiled unchanged and RUN here: openmmapi/src/VVIntegrator.cpp,
platforms/cuda/src/CudaVVKernels.cpp and CudaVVKernelFactory.cpp from /root/reference, linked against the mini-OpenMM of
oracle/mini_openmm (host flavour: oracle/_ref/libvvplugin_ref_cpu_<mode>.so; the prebuilt files travel to the GPU box).

This is what pins the INTEGER WORK of the path to reference-compiled code rather than to a
restatement: every index array the reference's initialize() methods upload is captured by CudaArray name and compared
bit for bit with what libvvb200's O(N) builders (csrc/vvb200_plan.cpp) and the C oracle produce; DOFs, chain masses, NkbT
and 1/M_total are compared as raw fp64 bytes.  It also pins the SCHEDULE: trajectories here come from the reference's own
VVIntegrator::stepMiddle / stepVV issuing the virtual calls, its own scaleVelocity doing the D2H / propagateNHChain / H2D.
"""
import dataclasses

import numpy as np
import pytest

from conftest import rel_err

EV = 1.60217662e-22
INT_ARRAYS = ("drudePairs", "particlesNH", "moleculesNH", "normalNH", "pairsNH", "particleMolId", "particlesInMolecules",
              "sortedByMol", "normalLD", "pairsLD", "imagePairs", "electrolyte")


def need(vo, flavour="ref_cpu", mode="mixed"):
    if not vo.plugin_available(flavour, mode):
        pytest.skip("oracle/_ref/libvvplugin not built (needs /root/reference at build time)")


def systems(vv):
    P = vv.Params
    out = {
        "bulk": (vv.make_bulk_ionic_liquid(20), P(max_drude_distance=0.02)),
        "bulk_constrained_cmm": (vv.make_bulk_ionic_liquid(20, hbond_constraints=True, has_cmm=True), P(max_drude_distance=0.02)),
        "nonpolar": (vv.make_nonpolar_box(40, 8), P()),
        "edl": (vv.make_edl(n_ion_pairs=8, n_electrode=90, electrode_molecules=3, hbond_constraints=True),
                P(max_drude_distance=0.02, mirror_location=1.1, electric_field=0.25 * EV)),
        "cosine": (vv.make_bulk_ionic_liquid(12), P(max_drude_distance=0.02, cos_acceleration=0.02)),
        "polymer": (vv.make_polymer(2, 60, 5, has_cmm=True, adjacent=True), P(max_drude_distance=0.02)),
        "polymer_far_partners": (vv.make_polymer(2, 40, 4), P()),
    }
    for seed in range(4):
        rag = vv.make_ragged(seed=seed)
        out[f"ragged{seed}"] = (rag, P(max_drude_distance=0.02, mirror_location=1.0, electric_field=1e-22))
    return out


NAMES = ["bulk", "bulk_constrained_cmm", "nonpolar", "edl", "cosine", "polymer", "polymer_far_partners",
         "ragged0", "ragged1", "ragged2", "ragged3"]


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("middle", [True, False], ids=["middle", "vv"])
def test_index_builders_match_reference_compiled_code(vv, vo, name, middle):
    """bit-exact integer arrays and fp64 DOFs / chain masses: libvvb200's plan builders and the C oracle (both its literal
    O(N^2) and its O(N) variant) against the arrays the reference's own initialize() methods uploaded"""
    need(vo)
    spec, params = systems(vv)[name]
    params = dataclasses.replace(params.resolved_for(spec), use_middle_scheme=middle)
    ref = vo.MiniContext(spec, params, "mixed", "ref_cpu")
    plan = vv.Plan(spec, params, "mixed")
    oracles = [vo.Oracle(spec, params, "mixed", literal=True), vo.Oracle(spec, params, "mixed", literal=False)]
    for arr in INT_ARRAYS:
        want = ref.upload(arr)
        if want is None:                       # the reference never uploads an empty vector (the array stays zeroed)
            want = np.zeros(0, np.int32)
        got = plan.int_array(arr).reshape(-1)
        assert np.array_equal(got, want), f"plan {arr}"
        for o in oracles:
            assert np.array_equal(o.array(arr).reshape(-1), want), f"oracle {arr}"
    # VVIntegrator members (VVIntegrator.cpp:123-145)
    for arr in ("particlesNH", "moleculesNH", "particleMolId"):
        assert np.array_equal(plan.int_array(arr), ref.int_list(arr)), arr
    assert np.array_equal(spec.mol_id, ref.int_list("particleMolId")), "ContextImpl::getMolecules labelling"
    for arr in ("moleculeMasses", "moleculeInvMasses"):
        assert plan.f64_array(arr).tobytes() == ref.f64(arr).tobytes(), arr
    if ref.int_list("particlesNH").size:
        ng = int(ref.f64("settings")[3])
        assert plan.num_temp_groups == ng and oracles[0].num_temp_groups == ng
        assert plan.f64_array("dof").tobytes() == ref.f64("dof").tobytes()
        assert plan.f64_array("etaMass").tobytes() == ref.f64("etaMass").tobytes()
        assert plan.f64_array("NkbT").tobytes() == ref.f64("NkbT").tobytes()
        for o in oracles:
            assert o.array("dof").tobytes() == ref.f64("dof").tobytes()
            assert o.array("etaMass")[: ng * params.num_nh_chains].tobytes() == ref.f64("etaMass").tobytes()
    if params.cos_acceleration != 0:
        assert plan.f64_array("invMassTotal").tobytes() == ref.f64("invMassTotal").tobytes()


@pytest.mark.parametrize("name", ["bulk", "nonpolar", "edl", "polymer"])
def test_auto_rules_of_initialize(vv, vo, name):
    """VVIntegrator::initialize's auto rules (COM group and friction 5/ps with a DrudeForce, off and 1/ps without;
    VVIntegrator.cpp:106-121) as `Params.resolved_for` restates them for libvvb200's callers"""
    need(vo)
    spec, params = systems(vv)[name]
    ref = vo.MiniContext(spec, params, "mixed", "ref_cpu", auto=True)
    friction, drude_friction, use_com, _ = ref.f64("settings")
    resolved = params.resolved_for(spec)
    assert (friction, drude_friction, bool(use_com)) == (resolved.friction, resolved.drude_friction, bool(resolved.use_com_temp_group))


def test_configuration_errors_have_the_reference_texts(vv, vo):
    """the OpenMMExceptions of VVIntegrator::initialize / Cuda*Kernel::initialize (VVIntegrator.cpp:103,149,155;
    CudaVVKernels.cpp:519,537) thrown by reference-compiled code == the VVB200_ERR_CONFLICT messages of vvb200_plan_create"""
    need(vo)
    P = vv.Params

    def both(spec, params, **kw):
        with pytest.raises(vo.PluginError) as r:
            vo.MiniContext(spec, params, "mixed", "ref_cpu", **kw)
        with pytest.raises(vv.VVB200Error) as p:
            vv.Plan(spec, params, "mixed")
        assert p.value.code == 2 and p.value.message == str(r.value)
        return str(r.value)
    bulk = vv.make_bulk_ionic_liquid(4)
    # a Langevin particle inside a Nose-Hoover molecule
    s = dataclasses.replace(bulk, langevin=np.array([3], np.int32))
    assert "NH and Langevin" in both(s, P().resolved_for(s))
    # Langevin + cosine acceleration
    s = dataclasses.replace(bulk, langevin=np.arange(37, dtype=np.int32))
    assert "shouldn't be used together" in both(s, P(cos_acceleration=0.02).resolved_for(s))
    # a Drude particle whose parent sits in the other thermostat: molecule 0 = particles 0..26; put parent 0 under
    # Langevin together with everything of the molecule except its Drude (1) ... which is then an NH particle of an
    # otherwise-Langevin molecule -> the molecule rule fires first, like in the reference
    s = dataclasses.replace(bulk, langevin=np.array([i for i in range(27) if i != 1], np.int32))
    assert "same molecule" in both(s, P().resolved_for(s))
    # constraint across thermostats: a constraint between an electrode (Langevin) atom and an ion (NH)
    edl = vv.make_edl(n_ion_pairs=2, n_electrode=6, electrode_molecules=2)
    # (a constraint also ties its two particles into ONE molecule, so the molecule rule is what fires -- on both sides)
    s = dataclasses.replace(edl, constraints=np.array([[0, 6]], np.int32)).finalize()
    msg = both(s, P().resolved_for(s))
    assert "same molecule" in msg
    # two DrudeForces in the System (no counterpart in vvb200_system: the glue sees one DrudeForce pointer)
    with pytest.raises(vo.PluginError, match="multiple DrudeForces"):
        vo.MiniContext(bulk, P().resolved_for(bulk), "mixed", "ref_cpu", num_drude_forces=2)


def run_plugin(vv, vo, spec, params, mode, steps, flavour="ref_cpu", constrained=False, cos=False, **kw):
    host = vv.make_state(spec, mode, **kw)
    ctx = vo.MiniContext(spec, params, mode, flavour).set_state(host)
    cons = vo.ConstraintStandin(spec, host) if constrained else None
    if cons is not None:
        ctx.set_constraints(cons)
    ctx.step(steps)
    return host, cons, ctx, ctx.get_state()


CASES = [("bulk", False, {}), ("bulk_constrained_cmm", True, {}), ("bulk_constrained_cmm", True, dict(drude_spread=0.015)),
         ("nonpolar", False, {}), ("edl", False, dict(n_random=4 * 92, mirror=1.1)), ("edl", True, dict(n_random=4 * 92, mirror=1.1)),
         ("cosine", False, dict(cos=True)), ("polymer", False, {}), ("ragged1", False, dict(n_random=2000, mirror=1.0))]


@pytest.mark.parametrize("mode", ["mixed", "double", "single"])
@pytest.mark.parametrize("middle", [True, False], ids=["middle", "vv"])
@pytest.mark.parametrize("case", CASES, ids=[f"{c[0]}{'+constraints' if c[1] else ''}{'+wall' if 'drude_spread' in c[2] else ''}" for c in CASES])
def test_oracle_schedule_matches_the_reference_integrator(vv, vo, case, middle, mode):
    """VVIntegrator::step() of the reference (its own stepMiddle / stepVV, firstIntegrate / secondIntegrate, scaleVelocity
    with the host NH chain) against the oracle's restated schedule (vvo_step): element-wise kernels are the same arithmetic,
    so positions and velocities agree to the last bits; only the two single-block sums are reassociated"""
    need(vo, "ref_cpu", mode)
    name, constrained, kw = case
    kw = dict(kw)
    cos = kw.pop("cos", False)
    spec, params = systems(vv)[name]
    if mode != "mixed" and spec.image_pairs.size:
        pytest.skip("image charges only work in mixed mode in the reference (SURVEY Appendix C-3)")
    params = dataclasses.replace(params.resolved_for(spec), use_middle_scheme=middle)
    host = vv.make_state(spec, mode, **kw)
    box = host.box
    host, cons, ctx, got = run_plugin(vv, vo, spec, params, mode, 3, constrained=constrained, **kw)
    oracle = vo.Oracle(spec, params, mode, literal=True).set_constraints(cons)
    want = host.copy()
    oracle.step(want, steps=3, inv_box_z=1.0 / box[2] if cos else 0.0)
    n = spec.n
    tol = 2e-5 if mode == "single" else 1e-12
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= tol and rel_err(got.positions()[:n], want.positions()[:n]) <= tol
    assert np.array_equal(got.velm[:, 3], want.velm[:, 3]) and np.array_equal(got.posq[:, 3], want.posq[:, 3])
    if ctx.int_list("particlesNH").size:
        a, b = ctx.thermostat_state(), oracle.thermostat_state()
        tk = 1e-4 if mode == "single" else 1e-12
        assert rel_err(a["ke2"], b["ke2"]) <= tk and rel_err(a["vscale"], b["vscale"]) <= tk
        assert rel_err(a["eta_dot"], b["eta_dot"]) <= max(tk, 1e-9) and rel_err(a["eta"], b["eta"]) <= max(tk, 1e-9)
    c = ctx.counters()
    assert c["step_count"] == 3 and c["force_evaluations"] == (3 if middle else 4)      # stepVV: one extra evaluation up front
    assert c["constraint_calls"] == 3 and c["velocity_constraint_calls"] == 3 and c["reorder_calls"] == 3
    if cos:
        v1, i1 = ctx.viscosity()
        v2, i2 = oracle.viscosity(box)
        assert abs(v1 - v2) <= 1e-12 * max(abs(v2), 1e-3) and abs(i1 - i2) <= 1e-12 * max(abs(i2), 1e-3)


@pytest.mark.parametrize("middle", [True, False], ids=["middle", "vv"])
def test_restated_harness_schedule_matches_the_reference_integrator(vv, vo, middle):
    """oracle/ref_harness.cpp drives the reference's kernels with a RESTATED host schedule (it is what the golden fixtures
    and the on-GPU reference baseline use): same result as the reference's own VVIntegrator::step driving them"""
    need(vo)
    if not vo.ref_available("mixed", gpu=False):
        pytest.skip("oracle/_ref/libvvref_cpu not built")
    spec, params = systems(vv)["edl"]
    params = dataclasses.replace(params.resolved_for(spec), use_middle_scheme=middle)
    host, cons, ctx, got = run_plugin(vv, vo, spec, params, "mixed", 3, constrained=True, n_random=4 * 92, mirror=1.1)
    oracle = vo.Oracle(spec, params, "mixed", literal=True)
    ref = vo.Reference(oracle, gpu=False).set_constraints(cons)
    want = host.copy()
    ref.step(want, steps=3)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= 1e-14 and rel_err(got.positions()[:n], want.positions()[:n]) <= 1e-14
    a, b = ctx.thermostat_state(), ref.thermostat_state()
    assert rel_err(a["vscale"], b["vscale"]) <= 1e-14


def test_step_size_change_between_steps(vv, vo):
    """the reference re-reads getStepSize() every step (CudaVVKernels.cpp:137-141)"""
    need(vo)
    spec, params = systems(vv)["bulk"]
    params = params.resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    ctx = vo.MiniContext(spec, params, "mixed", "ref_cpu").set_state(host)
    ctx.step(1)
    ctx.set_step_size(0.0005)
    ctx.step(2)
    got = ctx.get_state()
    want = host.copy()
    vo.Oracle(spec, params, "mixed").step(want, steps=1)
    o2 = vo.Oracle(spec, dataclasses.replace(params, step_size=0.0005), "mixed")
    # carry the chain state over: second oracle continues the first one's thermostat
    o1 = vo.Oracle(spec, params, "mixed")
    w = host.copy()
    o1.step(w, steps=1)
    st = o1.thermostat_state()
    o2.lib.vvo_set_nhc_state(o2.h, vo._ptr(np.ascontiguousarray(st["eta"])), vo._ptr(np.ascontiguousarray(st["eta_dot"])),
                             vo._ptr(np.ascontiguousarray(st["eta_dotdot"])))
    o2.step(w, steps=2)
    n = spec.n
    assert rel_err(got.velm[:n, :3], w.velm[:n, :3]) <= 1e-11 and rel_err(got.positions()[:n], w.positions()[:n]) <= 1e-12
