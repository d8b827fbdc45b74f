"""Edge cases of the CUDA path through the C ABI: sizes that are not multiples of anything, one-particle and
one-pair systems, systems without a thermostat, without Drude particles, massless sites, step-size changes,
thermostat-state save / restore (the reference loses its chain state on resume), error returns."""
import dataclasses

import numpy as np
import pytest

from conftest import TIGHT_HARDWALL, rel_err

pytestmark = pytest.mark.gpu


def both(vv, vo, spec, params, precision="mixed", steps=3, **kw):
    host = vv.make_state(spec, precision, **kw)
    plan = vv.Plan(spec, params, precision).upload()
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=steps)
    got = bufs.to_host()
    oracle = vo.Oracle(spec, params, precision, literal=False)
    want = host.copy()
    oracle.step(want, steps=steps)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= TIGHT_HARDWALL[precision]
    assert rel_err(got.positions()[:n], want.positions()[:n]) <= TIGHT_HARDWALL[precision]
    return plan, oracle, got, want


def tiny(vv, masses, pairs=(), mol=None, **kw):
    n = len(masses)
    bonds = [(p, d) for d, p in pairs]
    spec = vv.SystemSpec(n=n, masses=np.array(masses, float), bonds=np.array(bonds, np.int32).reshape(-1, 2),
                         drude_pairs=np.array(pairs, np.int32).reshape(-1, 2), **kw)
    return spec.finalize(mol_id=None if mol is None else np.array(mol, np.int32))


@pytest.mark.parametrize("precision", ["mixed", "double", "single"])
def test_single_particle(vv, vo, precision):
    spec = tiny(vv, [12.0])
    params = vv.Params().resolved_for(spec)
    if precision == "single":
        host = vv.make_state(spec, precision)
        plan = vv.Plan(spec, params, precision).upload()
        bufs = vv.DeviceBuffers(host)
        plan.step(bufs, steps=2)
        assert np.isfinite(bufs.to_host().velm).all()
    else:
        both(vv, vo, spec, params, precision)


def test_single_drude_pair(vv, vo):
    spec = tiny(vv, [11.6, 0.4], pairs=[(1, 0)])
    both(vv, vo, spec, vv.Params(max_drude_distance=0.02).resolved_for(spec))


@pytest.mark.parametrize("n_ip", [1, 3, 7, 13, 14, 28])
def test_sizes_around_tile_boundaries(vv, vo, n_ip):
    """37 n particles: tiles end at odd indices, bulk copies start at unaligned slots"""
    spec = vv.make_bulk_ionic_liquid(n_ip)
    plan, *_ = both(vv, vo, spec, vv.Params(max_drude_distance=0.02).resolved_for(spec))
    ts = plan.int_array("tileStart")
    assert ts[-1] == 37 * n_ip


def test_odd_monatomic_box(vv, vo):
    """1,001 monatomic ions: 128 molecules per tile cap, every molecule a single particle (atom group has 0 DOF)"""
    spec = tiny(vv, [35.45] * 1001)
    params = dataclasses.replace(vv.Params(), use_com_temp_group=True)
    plan, *_ = both(vv, vo, spec, params)
    assert np.all(np.diff(plan.int_array("tileStart")) <= 128)


def test_no_thermostat_all_langevin(vv, vo):
    spec = vv.make_nonpolar_box(50, 5, has_cmm=False)
    spec = dataclasses.replace(spec, langevin=np.arange(spec.n, dtype=np.int32)).finalize(mol_id=spec.mol_id)
    params = dataclasses.replace(vv.Params(), use_com_temp_group=False)
    plan, *_ = both(vv, vo, spec, params, n_random=4 * (spec.n + 2))
    assert plan.int_array("particlesNH").size == 0


def test_massless_sites_and_virtual_like_particles(vv, vo):
    masses = [15.999, 1.008, 1.008, 0.0] * 40            # TIP4P-like: a massless site in every molecule
    mol = np.repeat(np.arange(40), 4)
    spec = tiny(vv, masses, mol=mol)
    both(vv, vo, spec, dataclasses.replace(vv.Params(), use_com_temp_group=True))


def test_step_size_change_is_picked_up(vv, vo):
    spec = vv.make_bulk_ionic_liquid(20)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    plan = vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=2)
    plan.set_step_size(0.0005)
    plan.step(bufs, steps=2)
    got = bufs.to_host()
    want = host.copy()
    o1 = vo.Oracle(spec, params, "mixed", literal=False)
    o1.step(want, steps=2)
    st = o1.thermostat_state()
    p2 = dataclasses.replace(params, step_size=0.0005)
    o2 = vo.Oracle(spec, p2, "mixed", literal=False)
    o2.lib.vvo_set_nhc_state(o2.h, st["eta"].ctypes.data, st["eta_dot"].ctypes.data, st["eta_dotdot"].ctypes.data)
    o2.step(want, steps=2)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= 1e-8


def test_thermostat_state_save_restore(vv, vo):
    """4 steps == 2 steps, save chain state, fresh plan, restore, 2 steps (bitwise)"""
    spec = vv.make_bulk_ionic_liquid(40)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    a = vv.DeviceBuffers(host)
    pa = vv.Plan(spec, params, "mixed").upload()
    pa.step(a, steps=4)
    b = vv.DeviceBuffers(host)
    pb = vv.Plan(spec, params, "mixed").upload()
    pb.step(b, steps=2)
    st = pb.thermostat_state()
    mid = b.to_host()
    pc = vv.Plan(spec, params, "mixed").upload()
    pc.set_thermostat_state(st["eta"], st["eta_dot"], st["eta_dotdot"])
    c = vv.DeviceBuffers(mid)
    pc.step(c, steps=2)
    ha, hc = a.to_host(), c.to_host()
    assert np.array_equal(ha.velm, hc.velm) and np.array_equal(ha.posq, hc.posq) and np.array_equal(ha.corr, hc.corr)


def test_error_returns(vv):
    import torch
    spec = vv.make_edl(n_ion_pairs=4, n_electrode=30, electrode_molecules=3)
    params = vv.Params(mirror_location=1.0).resolved_for(spec)
    host = vv.make_state(spec, "mixed", n_random=200, mirror=1.0)
    plan = vv.Plan(spec, params, "mixed")
    bufs = vv.DeviceBuffers(host)
    with pytest.raises(vv.VVB200Error) as e:
        plan.step_middle(bufs)                              # not uploaded
    assert e.value.code == 5
    plan.upload()
    bufs.corr = None
    with pytest.raises(vv.VVB200Error) as e:
        plan.step_middle(bufs)                              # mixed mode needs posqCorrection
    assert e.value.code == 1
    bufs = vv.DeviceBuffers(host)
    bufs.random = None
    with pytest.raises(vv.VVB200Error) as e:
        plan.step_middle(bufs)                              # Langevin particles but no random buffer
    assert e.value.code == 1 and "random" in e.value.message
    bufs = vv.DeviceBuffers(host)
    with pytest.raises(vv.VVB200Error):
        plan.middle_finish(bufs)                            # no posDelta
    with pytest.raises(vv.VVB200Error):
        plan.step_host(host, steps=1)                       # host entry point refuses Langevin systems
    torch.cuda.synchronize()


def test_viscosity_and_launch_count(vv, vo, step_path):
    spec = vv.make_bulk_ionic_liquid(30)
    params = vv.Params(cos_acceleration=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    plan = vv.Plan(spec, params, "mixed").upload()
    v0, iv0 = plan.viscosity(host.box)
    assert v0 == 0.0 and iv0 == 0.0                         # defined before the first step (the reference's is not)
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=2, inv_box_z=1.0 / host.box[2])
    # one launch per step when the system is resident in shared memory, else two; never a host sync
    assert plan.launch_count == (2 if step_path == "resident" else 4)
    assert plan.resident_launch_count == (2 if step_path == "resident" else 0)
    v, iv = plan.viscosity(host.box)
    assert v != 0.0 and np.isfinite(iv)


def test_resident_kernel_several_tiles_per_block(vv, vo, step_path, monkeypatch):
    """above 2 x 148 tiles the resident kernel keeps several tiles per block (one block per SM); not the default at
    this size (the streaming passes are faster there) but it must stay correct: same results as the two passes"""
    if step_path != "resident":
        pytest.skip("resident-only case")
    monkeypatch.setenv("VVB200_RESIDENT_MAX_PARTICLES", "400000")
    spec = vv.make_bulk_ionic_liquid(4500)                  # 166,500 particles -> 326 tiles -> 3 tiles per block
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    one, two = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    two.set_resident_mode(0)
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host)
    one.step(a, steps=3)
    two.step(b, steps=3)
    assert one.resident_launch_count == 3 and one.launch_count == 3
    assert two.resident_launch_count == 0 and two.launch_count == 6
    ha, hb = a.to_host(), b.to_host()
    n = spec.n
    assert rel_err(ha.velm[:n, :3], hb.velm[:n, :3]) <= 1e-10 and rel_err(ha.positions()[:n], hb.positions()[:n]) <= 1e-12
    sa, sb = one.thermostat_state(), two.thermostat_state()
    assert rel_err(sa["ke2"], sb["ke2"]) <= 1e-13 and rel_err(sa["vscale"], sb["vscale"]) <= 1e-13


def _reporter_temperatures(spec, velm, use_cmm):
    """independent numpy statement of what the reference's Drude-temperature reporter prints
    (examples/ommhelper/reporter/drudetemperaturereporter.py:98-129): COM / atom / Drude kinetic energies x2 and DOFs"""
    kB = 1.380649e-23 * 6.02214076e23 / 1000.0
    n = spec.n
    v = velm[:n, :3].astype(np.float64).copy()
    m = spec.masses.astype(np.float64).copy()
    mol = spec.mol_id
    mass_mol = np.bincount(mol, weights=m, minlength=spec.n_mol)
    vmol = np.stack([np.bincount(mol, weights=m * v[:, d], minlength=spec.n_mol) for d in range(3)], axis=1)
    vmol = vmol / np.where(mass_mol > 0, mass_mol, 1.0)[:, None]
    ke2_com = float(np.sum(mass_mol * np.sum(vmol ** 2, axis=1)))
    dof_com = 3 * np.count_nonzero(mass_mol) - (3 if use_cmm else 0)
    nd = spec.drude_pairs.shape[0]
    dof_atom = 3 * np.count_nonzero(m) - 3 * np.count_nonzero(mass_mol) - spec.constraints.shape[0] - 3 * nd
    v -= vmol[mol]
    is_drude = np.zeros(n, bool)
    for d, c in spec.drude_pairs:
        md, mc = m[d], m[c]
        vd, vc = v[d].copy(), v[c].copy()
        v[d], v[c] = vd - vc, (md * vd + mc * vc) / (md + mc)
        m[d], m[c] = md * mc / (md + mc), md + mc
        is_drude[d] = True
    mvv = m * np.sum(v ** 2, axis=1)
    ke2_atom, ke2_drude = float(mvv[~is_drude].sum()), float(mvv[is_drude].sum())
    return (np.array([ke2_atom, ke2_com, ke2_drude]), np.array([dof_atom, dof_com, 3 * nd]),
            np.array([ke2_atom / (dof_atom * kB), ke2_com / (dof_com * kB), ke2_drude / (3 * nd * kB)]))


@pytest.mark.parametrize("general", [False, True])
def test_measure_temperatures_matches_the_reporter(vv, vo, general, monkeypatch):
    """vvb200_measure_temperatures: one reduction launch, no stepping, thermostat state untouched; the numbers are the
    reference reporter's (all particles thermostatted => same energies and DOFs)"""
    if general:
        monkeypatch.setenv("VVB200_FORCE_GENERAL", "1")
    spec = vv.make_bulk_ionic_liquid(200, has_cmm=True)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)       # gentle frozen forces: two steps barely heat the box
    plan = vv.Plan(spec, params, "mixed").upload()
    assert plan.tiled != general
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=2)
    before = plan.thermostat_state()
    state = bufs.to_host()
    t = plan.measure_temperatures(bufs)
    after = plan.thermostat_state()
    for k in ("eta", "eta_dot", "eta_dotdot", "vscale", "ke2"):
        assert np.array_equal(before[k], after[k]), k                  # measuring is not stepping
    assert np.array_equal(bufs.to_host().velm, state.velm)
    ke2, dof, temp = _reporter_temperatures(spec, state.velm, use_cmm=True)
    assert t["num_temp_groups"] == 3
    assert rel_err(t["dof"], dof) <= 1e-13      # the reference accumulates fractional per-particle DOFs in fp64 (13200.000000000036)
    assert rel_err(t["ke2"], ke2) <= 1e-11 and rel_err(t["temperature"], temp) <= 1e-11
    assert 150 < t["temperature"][0] < 600 and t["temperature"][2] < 50      # sane: ~333 K atoms, cold Drudes


def test_checkpoint_resumes_bitwise(vv, vo):
    """save after 3 steps, keep going 4 more; a fresh plan + loaded checkpoint + the saved arrays reproduces them bitwise
    (the reference restarts its NH chains from zero on resume)"""
    spec = vv.make_bulk_ionic_liquid(150)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    p1 = vv.Plan(spec, params, "mixed").upload()
    b1 = vv.DeviceBuffers(host)
    p1.step(b1, steps=3)
    blob, saved = p1.checkpoint_save(), b1.to_host()
    p1.step(b1, steps=4)
    want, want_state = b1.to_host(), p1.thermostat_state()

    p2 = vv.Plan(spec, params, "mixed").upload()
    p2.checkpoint_load(blob)
    b2 = vv.DeviceBuffers(saved)
    p2.step(b2, steps=4)
    got, got_state = b2.to_host(), p2.thermostat_state()
    assert np.array_equal(got.velm, want.velm) and np.array_equal(got.posq, want.posq) and np.array_equal(got.corr, want.corr)
    for k in ("eta", "eta_dot", "eta_dotdot", "vscale", "ke2"):
        assert np.array_equal(got_state[k], want_state[k]), k
    # without the checkpoint the chains restart cold and the trajectory differs
    p3 = vv.Plan(spec, params, "mixed").upload()
    b3 = vv.DeviceBuffers(saved)
    p3.step(b3, steps=4)
    assert not np.array_equal(b3.to_host().velm, want.velm)
    # refused: truncated blob, foreign bytes, another chain length
    with pytest.raises(vv.VVB200Error):
        p2.checkpoint_load(blob[:40])
    with pytest.raises(vv.VVB200Error):
        p2.checkpoint_load(bytes(len(blob)))
    other = vv.Plan(spec, dataclasses.replace(params, num_nh_chains=5), "mixed").upload()
    with pytest.raises(vv.VVB200Error) as e:
        other.checkpoint_load(blob)
    assert e.value.code == 2


def test_edl_step_is_one_launch(vv, vo, step_path):
    """BASELINE config 3 (electrode + electrolyte: Langevin subset, external field, image charges): the Langevin force is
    evaluated inside the kick and the image particles are mirrored by their parents' threads, so the whole step is one
    launch when resident (two streaming passes otherwise); the reference needs 14"""
    spec = vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    host = vv.make_state(spec, "mixed", n_random=8 * 626, mirror=2.0)
    plan = vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    ri = plan.step(bufs, steps=4)
    assert ri == 4 * plan.random_request
    assert plan.launch_count == (4 if step_path == "resident" else 8)
    got, want = bufs.to_host(), host.copy()
    vo.Oracle(spec, params, "mixed", literal=False).step(want, steps=4)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= TIGHT_HARDWALL["mixed"]
    assert rel_err(got.positions()[:n], want.positions()[:n]) <= TIGHT_HARDWALL["mixed"]
    img = spec.image_pairs[:, 0]
    assert np.array_equal(got.posq[img, 3], host.posq[img, 3])            # the images keep their own charges
    assert np.array_equal(got.posq[img, :2], got.posq[spec.image_pairs[:, 1], :2])   # x, y copied bit for bit


def test_launch_overlap_changes_nothing(vv, vo, monkeypatch):
    """programmatic dependent launch (each streaming kernel is scheduled while its predecessor drains and waits at
    griddepcontrol.wait before touching data): 1,500 steps of a 220k-particle box with it on and off end bitwise
    identical -- a kernel that started reading early would show up here.  Same for the early hand-over (pass B does not
    wait for the whole of pass A but for a device word its last block sets: VVB200_HANDOVER), in every combination."""
    spec = vv.make_bulk_ionic_liquid(6000)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    monkeypatch.setenv("VVB200_RESIDENT", "0")
    outs = []
    for pdl, hand in (("1", "1"), ("0", "0"), ("1", "0"), ("0", "1")):
        monkeypatch.setenv("VVB200_PDL", pdl)
        monkeypatch.setenv("VVB200_HANDOVER", hand)
        plan = vv.Plan(spec, params, "mixed").upload()
        bufs = vv.DeviceBuffers(host)
        plan.step(bufs, steps=1500)
        assert plan.launch_count == 3000
        outs.append((bufs.to_host(), plan.thermostat_state()))
    (a, sa) = outs[0]
    for b, sb in outs[1:]:
        assert np.array_equal(a.velm, b.velm) and np.array_equal(a.posq, b.posq) and np.array_equal(a.corr, b.corr)
        assert np.array_equal(sa["eta_dot"], sb["eta_dot"]) and np.array_equal(sa["eta"], sb["eta"])
    assert np.isfinite(a.velm).all()


@pytest.mark.gpu
@pytest.mark.parametrize("flow", ["vv", "constrained", "thermostat", "cosine", "polymer"])
def test_hand_over_changes_nothing_in_any_flow(vv, monkeypatch, flow):
    """the early hand-over in every entry point that uses it (velocity-Verlet halves, thermostat + deltas, thermostat
    alone), with the cosine moments and with molecules cut across tiles (whose centre of mass the LAST block finishes
    before pass B may load it): bitwise equal to waiting for the whole grid"""
    import dataclasses
    monkeypatch.setenv("VVB200_RESIDENT", "0")
    if flow == "polymer":
        spec = vv.make_polymer(n_chains=6, chain_len=3000, n_solvent=200, adjacent=True)
    else:
        spec = vv.make_bulk_ionic_liquid(5000)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02 if flow == "cosine" else 0.0).resolved_for(spec)
    if flow == "vv":
        params = dataclasses.replace(params, use_middle_scheme=False)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    inv_box_z = 1.0 / host.box[2] if flow == "cosine" else 0.0
    outs = []
    for hand in ("1", "0"):
        monkeypatch.setenv("VVB200_HANDOVER", hand)
        plan = vv.Plan(spec, params, "mixed").upload()
        bufs = vv.DeviceBuffers(host, with_pos_delta=True)
        for _ in range(200):
            if flow == "constrained":
                plan.middle_kick(bufs)
                plan.middle_thermostat_delta(bufs)
                plan.middle_finish(bufs)
            elif flow == "thermostat":
                plan.middle_kick(bufs)
                plan.thermostat(bufs)
            else:
                plan.step(bufs, steps=1, inv_box_z=inv_box_z)
        outs.append((bufs.to_host(), plan.thermostat_state(), plan.com_velocities()))
    (a, sa, ca), (b, sb, cb) = outs
    assert np.array_equal(a.velm, b.velm) and np.array_equal(a.posq, b.posq) and np.array_equal(a.corr, b.corr)
    assert np.array_equal(sa["eta_dot"], sb["eta_dot"]) and np.array_equal(sa["vscale"], sb["vscale"]) and np.array_equal(ca, cb)
    assert np.isfinite(a.velm).all()
