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
