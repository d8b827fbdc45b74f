"""GPU parity: the sm_100a path (through the C ABI) against the CPU oracle on identical inputs.

Tolerances (BASELINE.json north_star / SURVEY.md section 7): positions and velocities after frozen-force
steps within 1e-6 relative in mixed precision (1e-12 double, 1e-5 single); group kinetic energies and
NH scale factors within 1e-12 (mixed/double); integer work bit-exact (tests/test_plan_builders.py)."""
import dataclasses

import numpy as np
import pytest

from conftest import TIGHT, TIGHT_HARDWALL, TOL, TOL_KE, rel_err, rms_err

pytestmark = pytest.mark.gpu


def run_both(vv, vo, spec, params, precision, steps, inv_box_z=0.0, n_random=0, seed=12345, **state_kw):
    import torch
    assert torch.cuda.is_available()
    host = vv.make_state(spec, precision, n_random=n_random, seed=seed, **state_kw)
    plan = vv.Plan(spec, params, precision).upload()
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=steps, inv_box_z=inv_box_z)
    got = bufs.to_host()
    oracle = vo.Oracle(spec, params, precision, literal=False)
    want = host.copy()
    oracle.step(want, steps=steps, inv_box_z=inv_box_z)
    plan._bufs = bufs      # kept for check_com_velocities
    return plan, oracle, got, want


def check_com_velocities(spec, plan, got, precision):
    """calcCOMVelocities (drudeNoseHoover.cu:5-30) of the velocities the run ended with: the reduce pass behind
    vvb200_measure_temperatures leaves V_mol = sum(m v) / sum(m) of every thermostat molecule in comV"""
    plan.measure_temperatures(plan._bufs)
    com = plan.com_velocities()
    n = spec.n
    mols = plan.int_array("moleculesNH")
    mid = np.asarray(spec.mol_id[:n])
    w = got.velm[:n, 3].astype(np.float64)
    m = np.where(w != 0, 1.0 / np.where(w != 0, w, 1.0), 0.0)
    msum = np.bincount(mid, weights=m, minlength=spec.n_mol)
    want = np.stack([np.bincount(mid, weights=m * got.velm[:n, k].astype(np.float64), minlength=spec.n_mol) for k in range(3)], axis=1)
    want = want[mols] / msum[mols, None]
    err = rms_err if precision == "single" else rel_err
    assert err(com[mols, :3], want) <= (1e-5 if precision == "single" else 1e-12)
    assert err(com[mols, 3], 1.0 / msum[mols]) <= (1e-6 if precision == "single" else 1e-14)


def check_state(spec, got, want, precision, hardwall=True):
    n = spec.n
    err = rms_err if precision == "single" else rel_err
    ev = err(got.velm[:n, :3], want.velm[:n, :3])
    ex = err(got.positions()[:n], want.positions()[:n])
    assert np.array_equal(got.velm[:n, 3], want.velm[:n, 3]), "inverse masses must not change"
    assert np.array_equal(got.posq[:n, 3], want.posq[:n, 3]), "charges must not change"
    tol = (TIGHT_HARDWALL if hardwall else TIGHT)[precision]
    assert tol <= TOL[precision] or precision == "single"
    assert ev <= tol, f"velocity rel err {ev}"
    assert ex <= tol, f"position rel err {ex}"
    return ev, ex


def check_thermostat(plan, oracle, precision):
    a, b = plan.thermostat_state(), oracle.thermostat_state()
    ng = b["num_temp_groups"]
    assert a["num_temp_groups"] == ng
    assert rel_err(a["ke2"][:ng], b["ke2"]) <= TOL_KE[precision]
    assert rel_err(a["vscale"][:ng], b["vscale"]) <= TOL_KE[precision]
    if precision != "single":
        assert rel_err(a["eta_dot"], b["eta_dot"]) <= 1e-9
        assert rel_err(a["eta"], b["eta"]) <= 1e-9


@pytest.mark.parametrize("precision", ["mixed", "double", "single"])
@pytest.mark.parametrize("hardwall", [0.0, 0.02])
def test_bulk_tgnh_middle(vv, vo, precision, hardwall):
    """BASELINE config 2: Drude ionic-liquid bulk, TGNH (atom/COM/Drude groups), middle scheme."""
    spec = vv.make_bulk_ionic_liquid(250)     # 9,250 particles = examples/models/bulk_Im21
    params = vv.Params(max_drude_distance=hardwall).resolved_for(spec)
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=3)
    check_state(spec, got, want, precision, hardwall=hardwall > 0)
    check_thermostat(plan, oracle, precision)
    if params.use_com_temp_group:
        check_com_velocities(spec, plan, got, precision)


@pytest.mark.parametrize("precision", ["mixed", "double", "single"])
def test_nonpolar_nh(vv, vo, precision):
    """BASELINE config 1 topology: non-polarizable box, single NH group, COM group off, CMMotionRemover."""
    spec = vv.make_nonpolar_box(512, 8)
    params = vv.Params().resolved_for(spec)
    assert not params.use_com_temp_group
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=4)
    assert plan.num_temp_groups == 1
    check_state(spec, got, want, precision, hardwall=False)
    check_thermostat(plan, oracle, precision)


@pytest.mark.parametrize("precision", ["mixed"])
def test_edl_langevin_field_images(vv, vo, precision):
    """BASELINE config 3: Langevin electrode + NH electrolyte + external field + image charges, with the
    same injected random stream on both sides."""
    spec = vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=2.0,
                       electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=3, n_random=4 * (624 + 2), mirror=2.0)
    assert plan.random_request == 624 + 2      # padded request: SURVEY Appendix C-7
    check_state(spec, got, want, precision)
    check_thermostat(plan, oracle, precision)
    n = spec.n
    im, pa = spec.image_pairs[:, 0], spec.image_pairs[:, 1]
    # image x,y are bit-identical copies of the parent's (imageCharge.cu:14-17)
    assert np.array_equal(got.posq[im, :2], got.posq[pa, :2])
    assert np.array_equal(got.corr[im, :2], got.corr[pa, :2])
    z = got.positions()
    assert np.max(np.abs(z[im, 2] + z[pa, 2] - 2 * 2.0)) < 1e-6


@pytest.mark.parametrize("precision", ["mixed", "double", "single"])
@pytest.mark.parametrize("middle", [True, False])
def test_cosine_perturbation(vv, vo, precision, middle):
    """BASELINE config 4: periodic-perturbation viscosity run (cosine acceleration + velocity-profile
    removal around the thermostat)."""
    spec = vv.make_bulk_ionic_liquid(100)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(spec),
                                 use_middle_scheme=middle)
    host0 = vv.make_state(spec, precision)
    inv_box_z = 1.0 / host0.box[2]
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=3, inv_box_z=inv_box_z)
    # the moment expansion of the bias reassociates fp64 sums: same tolerance class
    check_state(spec, got, want, precision)
    a, b = plan.thermostat_state(), oracle.thermostat_state()
    tk = 1e-10 if precision != "single" else 1e-4
    assert rel_err(a["ke2"], b["ke2"]) <= tk
    assert rel_err(a["vscale"], b["vscale"]) <= tk
    assert abs(a["velocity_bias"] - b["velocity_bias"]) <= tk * max(abs(b["velocity_bias"]), 1e-3)
    v1, i1 = plan.viscosity(host0.box)
    v2, i2 = oracle.viscosity(host0.box)
    assert abs(v1 - v2) <= tk * max(abs(v2), 1e-3) and abs(i1 - i2) <= tk * max(abs(i2), 1e-3)


@pytest.mark.parametrize("precision", ["mixed", "double", "single"])
def test_bulk_vv_scheme(vv, vo, precision):
    """classic velocity-Verlet schedule (VVIntegrator::stepVV): two thermostat half steps per step."""
    spec = vv.make_bulk_ionic_liquid(100)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02).resolved_for(spec), use_middle_scheme=False)
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=3)
    check_state(spec, got, want, precision)
    check_thermostat(plan, oracle, precision)


def test_vv_scheme_edl(vv, vo):
    spec = vv.make_edl(n_ion_pairs=32, n_electrode=300, electrode_molecules=3)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02, mirror_location=1.5,
                                           electric_field=0.25 * 1.60217662e-22).resolved_for(spec),
                                 use_middle_scheme=False)
    plan, oracle, got, want = run_both(vv, vo, spec, params, "mixed", steps=3, n_random=4 * 302, mirror=1.5)
    check_state(spec, got, want, "mixed")
    check_thermostat(plan, oracle, "mixed")


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5, 6, 7])
@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_ragged_topologies(vv, vo, seed, precision):
    """ragged molecule sizes, non-adjacent Drude partners, massless sites, Langevin molecules with Drude
    pairs, duplicated electrolyte entries, images, CMMotionRemover."""
    spec = vv.make_ragged(seed=seed, scattered_molecules=0)
    params = vv.Params(max_drude_distance=0.02, mirror_location=1.0, electric_field=1e-22).resolved_for(spec)
    plan, oracle, got, want = run_both(vv, vo, spec, params, precision, steps=3, n_random=8 * spec.n, mirror=1.0)
    assert plan.tiled
    check_state(spec, got, want, precision)
    check_thermostat(plan, oracle, precision)


def test_split_entry_points_match_fused(vv, vo):
    """the VVKernels.h-shaped entry points (kick / thermostat / delta / finish), as the OpenMM glue calls
    them around constraints, reproduce the fused step when there are no constraints."""
    import torch
    spec = vv.make_bulk_ionic_liquid(100)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    plan_a = vv.Plan(spec, params, "mixed").upload()
    plan_b = vv.Plan(spec, params, "mixed").upload()
    a = vv.DeviceBuffers(host)
    b = vv.DeviceBuffers(host, with_pos_delta=True)
    for _ in range(3):
        plan_a.step_middle(a)
        plan_b.middle_kick(b)
        plan_b.middle_delta(b, 0)
        plan_b.thermostat(b)
        plan_b.middle_delta(b, 1)
        plan_b.middle_finish(b)
    torch.cuda.synchronize()
    ha, hb = a.to_host(), b.to_host()
    n = spec.n
    # no FMA contraction anywhere: the split kernels round exactly like the fused ones
    assert np.array_equal(hb.velm, ha.velm) and np.array_equal(hb.posq, ha.posq) and np.array_equal(hb.corr, ha.corr)
    sa, sb = plan_a.thermostat_state(), plan_b.thermostat_state()
    assert np.array_equal(sb["vscale"], sa["vscale"])


def test_multi_gpu_split_equals_fused(vv, vo):
    """kick_reduce + (all-reduce stand-in: identity on one rank) + nhc_scale_drift == step_middle"""
    import torch
    spec = vv.make_bulk_ionic_liquid(64)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    p1, p2 = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host)
    for _ in range(2):
        p1.step_middle(a)
        p2.middle_kick_reduce(b)
        p2.middle_nhc_scale_drift(b)
    ha, hb = a.to_host(), b.to_host()
    assert np.array_equal(ha.velm, hb.velm) and np.array_equal(ha.posq, hb.posq) and np.array_equal(ha.corr, hb.corr)


def test_determinism(vv, vo):
    """fixed-order reductions: two runs give bitwise identical trajectories"""
    spec = vv.make_bulk_ionic_liquid(300)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    outs = []
    for _ in range(2):
        plan = vv.Plan(spec, params, "mixed").upload()
        bufs = vv.DeviceBuffers(host)
        plan.step(bufs, steps=5)
        outs.append(bufs.to_host())
    assert np.array_equal(outs[0].velm, outs[1].velm)
    assert np.array_equal(outs[0].posq, outs[1].posq)
    assert np.array_equal(outs[0].corr, outs[1].corr)


def test_step_host_roundtrip(vv, vo):
    """vvb200_step_host (host buffers in/out) equals the device entry points"""
    spec = vv.make_bulk_ionic_liquid(50)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    p1, p2 = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    p1.step(bufs, steps=2)
    want = bufs.to_host()
    got = host.copy()
    p2.step_host(got, steps=2)
    assert np.array_equal(got.velm, want.velm) and np.array_equal(got.posq, want.posq) and np.array_equal(got.corr, want.corr)


@pytest.mark.parametrize("cos", [False, True])
@pytest.mark.parametrize("precision", ["mixed", "single"])
def test_step_host_pipelined(vv, vo, monkeypatch, cos, precision):
    """one step through host buffers with copy-in / pass A / pass B / copy-out overlapped chunk by chunk (what
    vvb200_step_host does for >= 1M particles; forced here at 15k with 5 chunks) equals the device-resident step up to
    the association order of the group sums"""
    monkeypatch.setenv("VVB200_HOST_PIPELINE_MIN", "1000")
    monkeypatch.setenv("VVB200_HOST_CHUNKS", "5")
    monkeypatch.setenv("VVB200_SMALL_TILES", "0")
    spec = vv.make_bulk_ionic_liquid(400)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02 if cos else 0.0).resolved_for(spec)
    host = vv.make_state(spec, precision)
    inv_box_z = 1.0 / host.box[2] if cos else 0.0
    p1, p2 = vv.Plan(spec, params, precision).upload(), vv.Plan(spec, params, precision).upload()
    p1.set_resident_mode(0)
    bufs = vv.DeviceBuffers(host)
    got = host.copy()
    for _ in range(3):                        # three single-step calls: state round-trips through the host every step
        p1.step(bufs, steps=1, inv_box_z=inv_box_z)
        p2.step_host(got, steps=1, inv_box_z=inv_box_z)
    want = bufs.to_host()
    assert p2.launch_count == 3 * 10          # 5 tile ranges x (pass A + pass B) per step
    n = spec.n
    tol = 1e-11 if precision == "mixed" else 2e-5
    err = rel_err if precision == "mixed" else rms_err
    assert err(got.velm[:n, :3], want.velm[:n, :3]) <= tol and err(got.positions()[:n], want.positions()[:n]) <= tol
    assert np.array_equal(got.velm[n:], want.velm[n:]) and np.array_equal(got.posq[n:], want.posq[n:])   # padding untouched
    a, b = p1.thermostat_state(), p2.thermostat_state()
    assert rel_err(b["ke2"], a["ke2"]) <= (1e-12 if precision == "mixed" else 1e-5)


def test_large_system_properties(vv, vo):
    """>=1M particles (beyond what the scalar oracle checks every element of in seconds): size-independent
    properties -- group orthogonality: after a thermostat application with factors
    s_g, a fresh evaluation gives s_g^2 * KE2_g; and agreement with the (OpenMP) oracle on a sample."""
    import torch
    spec = vv.make_bulk_ionic_liquid(27648)          # 1,022,976 particles
    params = vv.Params().resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=0.0)
    plan = vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    plan.thermostat(bufs)
    s1 = plan.thermostat_state()
    # second evaluation sees the scaled velocities
    plan.thermostat(bufs)
    s2 = plan.thermostat_state()
    assert rel_err(s2["ke2"], s1["ke2"] * s1["vscale"] ** 2) <= 1e-10
    # total kinetic energy decomposes into the three groups
    v = host.velm[: spec.n, :3]
    m = spec.masses
    total = float(np.sum(m[:, None] * v * v))
    assert abs(s1["ke2"].sum() - total) <= 1e-10 * total
    # full step against the oracle on every element (oracle uses all host threads here)
    oracle = vo.Oracle(spec, params, "mixed", literal=False, threads=0)
    want = host.copy()
    oracle.scale_velocity(want)
    oracle.scale_velocity(want)
    got = bufs.to_host()
    assert rel_err(got.velm[: spec.n, :3], want.velm[: spec.n, :3]) <= 1e-6


def test_reduce_only_pass_with_several_tiles_per_block(vv, vo):
    """1M particles are 2,000 tiles on at most 444 blocks: every block of the reduce-only pass takes several tiles and its
    molecule lanes rotate from one tile to the next (vvb200_stream.cuh, reduce_kernel).  Its group energies against the
    oracle's and against pass A's sums of the same velocities (zero forces: the kick adds exact zeros)."""
    spec = vv.make_bulk_ionic_liquid(27648)          # 1,022,976 particles
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=0.0)
    red = vv.Plan(spec, params, "mixed").upload()
    b1 = vv.DeviceBuffers(host)
    red.thermostat(b1)
    ke_red = red.thermostat_state()["ke2"]
    fused = vv.Plan(spec, params, "mixed").upload()
    b2 = vv.DeviceBuffers(host)
    fused.step_middle(b2)
    ke_a = fused.thermostat_state()["ke2"]
    oracle = vo.Oracle(spec, params, "mixed", literal=False, threads=0)
    want = host.copy()
    oracle.scale_velocity(want)
    ke_o = oracle.thermostat_state()["ke2"]
    assert rel_err(ke_red, ke_a) <= 1e-13 and rel_err(ke_red, ke_o[: len(ke_red)]) <= TOL_KE["mixed"]
    assert rel_err(red.thermostat_state()["vscale"], fused.thermostat_state()["vscale"]) <= 1e-13
    # the molecular velocities both passes leave behind
    assert rel_err(red.com_velocities(), fused.com_velocities()) <= 1e-13


def test_vv_split_entry_points_match_fused(vv, vo):
    """velocity-Verlet scheme through the VVKernels.h-shaped calls (thermostat / half kick + posDelta / positions /
    half kick / thermostat, as the OpenMM glue issues them around constraints) == the fused vv_first + vv_second"""
    import torch
    spec = vv.make_edl(n_ion_pairs=32, n_electrode=300, electrode_molecules=3)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02, mirror_location=1.5,
                                           electric_field=0.25 * 1.60217662e-22).resolved_for(spec),
                                 use_middle_scheme=False)
    host = vv.make_state(spec, "mixed", n_random=4 * 302, mirror=1.5)
    pa, pb = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host, with_pos_delta=True)
    ri = 0
    for _ in range(3):
        pa.step_vv_first(a, random_index=ri)
        pa.step_vv_second(a, random_index=ri)
        pb.thermostat(b)
        pb.vv_kick(b, second_half=False, update_pos_delta=True, random_index=ri)
        pb.vv_positions(b)
        pb.update_image_positions(b)
        pb.vv_kick(b, second_half=True, update_pos_delta=False, random_index=ri)
        pb.thermostat(b)
        ri += pa.random_request
    torch.cuda.synchronize()
    ha, hb = a.to_host(), b.to_host()
    assert np.array_equal(hb.velm, ha.velm) and np.array_equal(hb.posq, ha.posq) and np.array_equal(hb.corr, ha.corr)


def test_cuda_graph_capture_replays_the_step(vv, vo):
    """no allocation, no synchronisation, no host round trip inside a step: ten steps captured into a CUDA graph and
    replayed give bitwise the trajectory of ten eager steps"""
    import torch
    spec = vv.make_bulk_ionic_liquid(250)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    eager_plan, graph_plan = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host)
    eager_plan.step(a, steps=12)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        graph_plan.step(b, steps=2)                 # warm-up: function attributes are set on first launch
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            graph_plan.step(b, steps=5)
        g.replay()
        g.replay()
        side.synchronize()
    ha, hb = a.to_host(), b.to_host()
    assert np.array_equal(ha.velm, hb.velm) and np.array_equal(ha.posq, hb.posq) and np.array_equal(ha.corr, hb.corr)
    sa, sb = eager_plan.thermostat_state(), graph_plan.thermostat_state()
    assert np.array_equal(sa["eta_dot"], sb["eta_dot"])


@pytest.mark.parametrize("cos", [False, True])
def test_constrained_flow_matches_fused(vv, vo, cos, step_path):
    """the 440 B/particle flow the glue uses when OpenMM constraints exist -- kick | thermostat_delta | finish, with
    OpenMM's solvers between them -- is bitwise the fused step when nothing is constrained"""
    import torch
    spec = vv.make_bulk_ionic_liquid(120)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02 if cos else 0.0).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    inv_box_z = 1.0 / host.box[2] if cos else 0.0
    pa, pb = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host, with_pos_delta=True)
    for _ in range(3):
        pa.step_middle(a, inv_box_z=inv_box_z)
        pb.middle_kick(b, inv_box_z=inv_box_z)
        pb.middle_thermostat_delta(b, inv_box_z=inv_box_z)
        pb.middle_finish(b)
    torch.cuda.synchronize()
    ha, hb = a.to_host(), b.to_host()
    assert np.array_equal(hb.velm, ha.velm) and np.array_equal(hb.posq, ha.posq) and np.array_equal(hb.corr, ha.corr)
    # kick | reduce + scale+delta (one launch when the system is resident) | finish + hard wall (one launch)
    assert pb.launch_count == 3 * (3 if step_path == "resident" else 4)
