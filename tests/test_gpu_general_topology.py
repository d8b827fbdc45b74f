"""The any-topology path (csrc/vvb200_general.cuh): systems whose thermostat molecules or Drude pairs do not fit a
512-slot tile, and -- through the VVB200_FORCE_GENERAL hook -- every ordinary system, against the CPU oracle and
against the tiled path."""
import dataclasses
import os

import numpy as np
import pytest

from conftest import TIGHT_HARDWALL, TOL_KE, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def force_general():
    os.environ["VVB200_FORCE_GENERAL"] = "1"
    yield
    os.environ.pop("VVB200_FORCE_GENERAL", None)


def run(vv, vo, spec, params, precision, steps, inv_box_z=0.0, expect_tiled=False, **kw):
    host = vv.make_state(spec, precision, **kw)
    plan = vv.Plan(spec, params, precision)
    assert plan.tiled == expect_tiled
    plan.upload()
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=steps, inv_box_z=inv_box_z)
    got = bufs.to_host()
    oracle = vo.Oracle(spec, params, precision, literal=False)
    want = host.copy()
    oracle.step(want, steps=steps, inv_box_z=inv_box_z)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= TIGHT_HARDWALL[precision]
    assert rel_err(got.positions()[:n], want.positions()[:n]) <= TIGHT_HARDWALL[precision]
    a, b = plan.thermostat_state(), oracle.thermostat_state()
    ng = b["num_temp_groups"]
    tk = TOL_KE[precision] if inv_box_z == 0 else 1e-10
    assert rel_err(a["ke2"][:ng], b["ke2"]) <= tk and rel_err(a["vscale"][:ng], b["vscale"]) <= tk
    assert rel_err(a["eta_dot"], b["eta_dot"]) <= 1e-9
    return plan, got


@pytest.mark.parametrize("middle", [True, False])
@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_polymer_is_not_tileable_and_matches_oracle(vv, vo, precision, middle):
    spec = vv.make_polymer(3, 700, 40, has_cmm=True)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02).resolved_for(spec), use_middle_scheme=middle)
    plan, _ = run(vv, vo, spec, params, precision, 3)
    assert plan.num_temp_groups == 3


@pytest.mark.parametrize("middle", [True, False])
def test_forced_general_bulk(vv, vo, force_general, middle):
    spec = vv.make_bulk_ionic_liquid(250)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02).resolved_for(spec), use_middle_scheme=middle)
    run(vv, vo, spec, params, "mixed", 3)


def test_forced_general_cosine(vv, vo, force_general):
    spec = vv.make_bulk_ionic_liquid(100)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(spec)
    host0 = vv.make_state(spec, "mixed")
    run(vv, vo, spec, params, "mixed", 3, inv_box_z=1.0 / host0.box[2])


def test_forced_general_edl(vv, vo, force_general):
    spec = vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    run(vv, vo, spec, params, "mixed", 3, n_random=4 * 626, mirror=2.0)


def test_forced_general_nonpolar_single_group(vv, vo, force_general):
    spec = vv.make_nonpolar_box(512, 8)
    run(vv, vo, spec, vv.Params().resolved_for(spec), "mixed", 4)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_ragged_with_scattered_molecules(vv, vo, seed):
    """image / extra massless sites bonded into far-away molecules: whichever path the plan picks must match"""
    spec = vv.make_ragged(seed=seed, scattered_molecules=2)
    params = vv.Params(max_drude_distance=0.02, mirror_location=1.0, electric_field=1e-22).resolved_for(spec)
    host = vv.make_state(spec, "mixed", n_random=8 * spec.n, mirror=1.0)
    plan = vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=3)
    got = bufs.to_host()
    oracle = vo.Oracle(spec, params, "mixed", literal=False)
    want = host.copy()
    oracle.step(want, steps=3)
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= TIGHT_HARDWALL["mixed"]
    assert rel_err(got.positions()[:n], want.positions()[:n]) <= TIGHT_HARDWALL["mixed"]


def test_general_equals_tiled(vv, vo):
    """same system through both paths: group sums are reassociated, everything else is the same arithmetic"""
    spec = vv.make_bulk_ionic_liquid(300)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    tiled = vv.Plan(spec, params, "mixed").upload()
    os.environ["VVB200_FORCE_GENERAL"] = "1"
    try:
        general = vv.Plan(spec, params, "mixed").upload()
    finally:
        os.environ.pop("VVB200_FORCE_GENERAL", None)
    assert tiled.tiled and not general.tiled
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host)
    tiled.step(a, steps=4)
    general.step(b, steps=4)
    ha, hb = a.to_host(), b.to_host()
    n = spec.n
    assert rel_err(hb.velm[:n, :3], ha.velm[:n, :3]) <= 1e-10 and rel_err(hb.positions()[:n], ha.positions()[:n]) <= 1e-12
    assert general.launch_count > tiled.launch_count


# ---- thermostat molecules longer than a tile, Drudes next to their parents: the FUSED path cuts the molecule, every tile
#      sums its fragment and the last block of pass A finishes the centre of mass (csrc/vvb200_stream.cuh) --------------
@pytest.mark.parametrize("middle", [True, False])
@pytest.mark.parametrize("precision", ["mixed", "double"])
def test_long_molecules_run_on_the_fused_path(vv, vo, precision, middle):
    spec = vv.make_polymer(3, 700, 40, has_cmm=True, adjacent=True)      # 3 chains of 1,750 particles + 40 waters
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02).resolved_for(spec), use_middle_scheme=middle)
    plan, _ = run(vv, vo, spec, params, precision, 3, expect_tiled=True)
    assert plan.num_temp_groups == 3
    assert np.max(np.diff(plan.int_array("tileStart"))) <= 512 < 1750


def test_long_molecules_cosine_and_single(vv, vo):
    spec = vv.make_polymer(2, 900, 30, adjacent=True)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(spec)
    host0 = vv.make_state(spec, "mixed")
    run(vv, vo, spec, params, "mixed", 3, inv_box_z=1.0 / host0.box[2], expect_tiled=True)
    # single precision: tolerance of the single-precision parity tests
    from conftest import rms_err
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "single")
    plan = vv.Plan(spec, params, "single").upload()
    assert plan.tiled
    bufs = vv.DeviceBuffers(host)
    plan.step(bufs, steps=3)
    got, want = bufs.to_host(), host.copy()
    vo.Oracle(spec, params, "single", literal=False).step(want, steps=3)
    n = spec.n
    assert rms_err(got.velm[:n, :3], want.velm[:n, :3]) <= 1e-4 and rms_err(got.positions()[:n], want.positions()[:n]) <= 1e-4


def test_long_molecules_fused_equals_any_topology_path(vv, vo, monkeypatch):
    """a polymer melt of 203k particles (streaming kernels; 40 chains of 5,000 particles cut into ~10 fragments each):
    the fused path with cross-tile centres of mass against the gather kernels of the any-topology path"""
    import torch
    spec = vv.make_polymer(40, 2000, 1000, adjacent=True)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    fused = vv.Plan(spec, params, "mixed").upload()
    monkeypatch.setenv("VVB200_SPLIT_MOLECULES", "0")
    general = vv.Plan(spec, params, "mixed").upload()
    assert fused.tiled and not general.tiled
    a, b = vv.DeviceBuffers(host), vv.DeviceBuffers(host)
    fused.step(a, steps=4)
    general.step(b, steps=4)
    ha, hb = a.to_host(), b.to_host()
    n = spec.n
    assert rel_err(ha.velm[:n, :3], hb.velm[:n, :3]) <= 1e-10 and rel_err(ha.positions()[:n], hb.positions()[:n]) <= 1e-12
    sa, sb = fused.thermostat_state(), general.thermostat_state()
    assert rel_err(sa["ke2"], sb["ke2"]) <= 1e-12 and rel_err(sa["vscale"], sb["vscale"]) <= 1e-12
    assert rel_err(fused.com_velocities()[:, :3], general.com_velocities()[:, :3]) <= 1e-9
    assert fused.launch_count < general.launch_count
    # and it is what makes the difference for such systems: time both
    st = torch.cuda.current_stream()
    out = []
    for plan, bufs in ((fused, a), (general, b)):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        plan.step(bufs, steps=50)
        e1.record(st)
        torch.cuda.synchronize()
        out.append(1e3 * e0.elapsed_time(e1) / 50)
    print(f"polymer melt {n} particles: fused {out[0]:.1f} us/step, any-topology path {out[1]:.1f} us/step")


@pytest.mark.parametrize("cos", [False, True])
def test_long_molecules_through_the_chunked_host_pipeline(vv, vo, monkeypatch, cos):
    """vvb200_step_host cuts a step of a large system into tile ranges (copy-in / pass A / pass B / copy-out overlapped).
    With thermostat molecules longer than a tile the fragments of one molecule can land in different ranges: the
    centre-of-mass finish of the cut molecules (and the M|V|^2 transfer between the atom and molecular groups) must run
    ONCE, in the launch that covers the last tiles -- not once per range on partly written fragment sums."""
    monkeypatch.setenv("VVB200_HOST_PIPELINE_MIN", "1000")
    monkeypatch.setenv("VVB200_HOST_CHUNKS", "7")
    spec = vv.make_polymer(4, 900, 30, adjacent=True)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02 if cos else 0.0).resolved_for(spec)
    host = vv.make_state(spec, "mixed")
    inv_box_z = 1.0 / host.box[2] if cos else 0.0
    p1, p2 = vv.Plan(spec, params, "mixed").upload(), vv.Plan(spec, params, "mixed").upload()
    assert p1.tiled and p1.int_array("splitMolId").size > 0
    p1.set_resident_mode(0)
    bufs = vv.DeviceBuffers(host)
    got = host.copy()
    for _ in range(3):
        p1.step(bufs, steps=1, inv_box_z=inv_box_z)
        p2.step_host(got, steps=1, inv_box_z=inv_box_z)
    want = bufs.to_host()
    assert p2.launch_count > 3 * 2            # really chunked
    n = spec.n
    assert rel_err(got.velm[:n, :3], want.velm[:n, :3]) <= 1e-11 and rel_err(got.positions()[:n], want.positions()[:n]) <= 1e-11
    a, b = p1.thermostat_state(), p2.thermostat_state()
    assert rel_err(b["ke2"], a["ke2"]) <= 1e-12 and rel_err(b["vscale"], a["vscale"]) <= 1e-12
