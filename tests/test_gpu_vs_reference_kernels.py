"""GPU parity against THE REFERENCE'S OWN CUDA KERNELS on the same B200 and the same inputs
(oracle/_ref/libvvref_cuda_<mode>.so: the reference's platforms/cuda/src/kernels/*.cu compiled unmodified for
sm_100a by oracle/Makefile, driven with OpenMM's launch geometry and the reference's blocking D2H/H2D around
the host Nose-Hoover chain; see oracle/ref_harness.cpp).  This is the comparison BASELINE.json's north_star
names first; tolerance = its 1e-6 relative for mixed precision (the reference build contracts FMAs, ours does
not, and Langevin forces are rounded to fp32 in mixed mode, so agreement is ~1e-13 .. 1e-7)."""
import dataclasses

import numpy as np
import pytest

from conftest import TOL, rel_err, rms_err

pytestmark = pytest.mark.gpu


def run(vv, vo, spec, params, mode, steps, inv_box_z=0.0, **kw):
    if not vo.ref_available(mode, gpu=True):
        pytest.skip("oracle/_ref/libvvref_cuda not built")
    host = vv.make_state(spec, mode, **kw)
    plan = vv.Plan(spec, params, mode).upload()
    a = vv.DeviceBuffers(host)
    plan.step(a, steps=steps, inv_box_z=inv_box_z)
    oracle = vo.Oracle(spec, params, mode, literal=False)         # index arrays only
    ref = vo.Reference(oracle, gpu=True)
    b = vv.DeviceBuffers(host)
    launches = ref.step(b, steps=steps, inv_box_z=inv_box_z)
    got, want = a.to_host(), b.to_host()
    n = spec.n
    err = rms_err if mode == "single" else rel_err
    ev, ex = err(got.velm[:n, :3], want.velm[:n, :3]), err(got.positions()[:n], want.positions()[:n])
    sa, sb = plan.thermostat_state(), ref.thermostat_state()
    ng = sb["num_temp_groups"]
    ek, es = rel_err(sa["ke2"][:ng], sb["ke2"]), rel_err(sa["vscale"][:ng], sb["vscale"])
    print(f"{spec.name} {mode}: v {ev:.2e} x {ex:.2e} ke2 {ek:.2e} vscale {es:.2e}; reference launches/step "
          f"{launches / steps:.0f}, ours {plan.launch_count / steps:.0f}")
    return ev, ex, ek, es


@pytest.mark.parametrize("mode", ["mixed", "double", "single"])
def test_bulk_tgnh(vv, vo, mode):
    spec = vv.make_bulk_ionic_liquid(250)
    ev, ex, ek, es = run(vv, vo, spec, vv.Params(max_drude_distance=0.02).resolved_for(spec), mode, 3)
    bar = {"mixed": TOL["mixed"], "double": TOL["double"], "single": 1e-4}[mode]
    assert max(ev, ex) <= bar
    assert max(ek, es) <= (1e-12 if mode != "single" else 1e-5)


def test_bulk_vv_scheme(vv, vo):
    spec = vv.make_bulk_ionic_liquid(100)
    params = dataclasses.replace(vv.Params(max_drude_distance=0.02).resolved_for(spec), use_middle_scheme=False)
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 3)
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-12


def test_cosine(vv, vo):
    spec = vv.make_bulk_ionic_liquid(100)
    params = vv.Params(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(spec)
    host0 = vv.make_state(spec, "mixed")
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 3, inv_box_z=1.0 / host0.box[2])
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-10


def test_edl(vv, vo):
    spec = vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 3, n_random=4 * 626, mirror=2.0)
    # Langevin forces are fp32 (real3) in mixed mode: an FMA-contracted and a non-contracted build differ in their
    # last float bit, i.e. ~1e-7 of the force => a few 1e-7 relative in the electrode velocities
    assert ex <= TOL["mixed"] and ev <= 5e-6 and max(ek, es) <= 1e-12


def test_nonpolar(vv, vo):
    spec = vv.make_nonpolar_box(512, 8)
    ev, ex, ek, es = run(vv, vo, spec, vv.Params().resolved_for(spec), "mixed", 4)
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-12


def test_million_particles(vv, vo):
    """1M-particle box: every element against the reference kernels (the reference's uninitialised KE-buffer
    tail, SURVEY Appendix C-2, is zeroed once by the harness)"""
    spec = vv.make_bulk_ionic_liquid(27648)
    ev, ex, ek, es = run(vv, vo, spec, vv.Params(max_drude_distance=0.02).resolved_for(spec), "mixed", 2, force_sigma=10.0)
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-11


# ---- the bar as BASELINE.json states it: ONE integrator step ("single substep") against the reference's own CUDA build ----
def test_edl_single_step_meets_the_stated_bar(vv, vo):
    """north_star: positions and velocities within 1e-6 relative after a single step, mixed precision.  The 3-step EDL
    test above allows 5e-6 in the velocities (fp32 Langevin forces, FMA-contracted in the reference build, amplified by
    the hard wall over three steps); the stated bar is for one step, and one step meets it with room."""
    spec = vv.make_edl(n_ion_pairs=64, n_electrode=624, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=2.0, electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 1, n_random=4 * 626, mirror=2.0)
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-12


def test_config3_at_its_full_size(vv, vo):
    """BASELINE configs[2] / SURVEY 8(d) C3 at the size bench.py times: 40,310 particles (511 ion pairs + 2,496 electrode
    atoms with Langevin, field on the electrolyte, image charges, hard wall) -- one step at the stated bar, three steps at
    the loosened velocity bar of test_edl"""
    spec = vv.make_edl(n_ion_pairs=511, n_electrode=2496, electrode_molecules=4)
    assert spec.n == 40310
    params = vv.Params(max_drude_distance=0.02, mirror_location=8.0, electric_field=0.25 * 1.60217662e-22).resolved_for(spec)
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 1, n_random=4 * 2500, mirror=8.0)
    assert max(ev, ex) <= TOL["mixed"] and max(ek, es) <= 1e-12
    ev, ex, ek, es = run(vv, vo, spec, params, "mixed", 3, n_random=4 * 2500, mirror=8.0)
    assert ex <= TOL["mixed"] and ev <= 5e-6 and max(ek, es) <= 1e-12


def test_single_precision_single_step(vv, vo):
    """single precision, one step: 1e-5 in the rms metric -- the 3-step test above allows 1e-4 once the
    hard wall's float cancellation has been through three steps"""
    spec = vv.make_bulk_ionic_liquid(250)
    ev, ex, ek, es = run(vv, vo, spec, vv.Params(max_drude_distance=0.02).resolved_for(spec), "single", 1)
    assert max(ev, ex) <= TOL["single"] and max(ek, es) <= 1e-5
