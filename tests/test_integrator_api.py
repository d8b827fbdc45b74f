"""The host-side mirror of the reference's user-facing API (openmm-velocityverlet_b200/integrator.py): names,
defaults, auto-default rules, list builders and exceptions of OpenMM::VVIntegrator / its SWIG surface.  CPU tests
cover the host logic; the GPU test steps a Context like examples/run-bulk.py does."""
import numpy as np
import pytest


def bulk_system(vv, n_ip=10, **kw):
    spec = vv.make_bulk_ionic_liquid(n_ip, **kw)
    return spec, vv.System.from_spec(spec)


def test_constructor_defaults_match_reference(vv):
    it = vv.VVIntegrator(300.0, 10.0, 1.0, 40.0, 0.001)
    # VVIntegrator.cpp:49-69
    assert it.getNumNHChains() == 3 and it.getLoopsPerStep() == 1
    assert it.getConstraintTolerance() == 1e-5 and it.getMaxDrudeDistance() == 0
    assert it.getFriction() == 5 and it.getDrudeFriction() == 20 and it.getRandomNumberSeed() == 0
    assert it.getMirrorLocation() == 0.0 and it.getElectricField() == 0.0 and it.getCosAcceleration() == 0.0
    assert it.getUseMiddleScheme() is True and it.getUseCOMTempGroup() is False and it.getDebugEnabled() is False
    # SWIG quirks (velocityverletplugin.i:110-113): frictions are truncated to int on the Python side
    it.setFriction(2.7)
    assert it.getFriction() == 2
    it.setDrudeFriction(9.9)
    assert it.getDrudeFriction() == 9
    assert it.addParticleLangevin(5) == 1 and it.addParticleLangevin(7) == 2
    assert it.addImagePair(10, 2) == 1
    assert it.addParticleElectrolyte(3) is None


def test_auto_defaults(vv):
    """VVIntegrator.cpp:106-121 and SURVEY Appendix C-12/C-16"""
    spec, system = bulk_system(vv)
    it = vv.VVIntegrator(300.0, 10.0, 1.0, 40.0, 0.001)
    vv.Context(system, it)
    assert it.getUseCOMTempGroup() is True and it.getFriction() == 5          # Drude model
    box = vv.make_nonpolar_box(8, 4)
    it2 = vv.VVIntegrator(300.0, 10.0, 1.0, 40.0, 0.001)
    vv.Context(vv.System.from_spec(box), it2)
    assert it2.getUseCOMTempGroup() is False and it2.getFriction() == 1       # non-polarizable
    it3 = vv.VVIntegrator(300.0, 10.0, 1.0, 40.0, 0.001)
    it3.setUseCOMTempGroup(True)                                              # user choice wins
    it3.setDrudeFriction(30)                                                  # also disables the friction auto-default
    vv.Context(vv.System.from_spec(box), it3)
    assert it3.getUseCOMTempGroup() is True and it3.getFriction() == 5


def test_lists_and_exceptions(vv):
    spec = vv.make_edl(n_ion_pairs=4, n_electrode=30, electrode_molecules=3)
    system = vv.System.from_spec(spec)
    it = vv.VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)
    for i in spec.langevin:
        it.addParticleLangevin(int(i))
    for im, pa in spec.image_pairs:
        it.addImagePair(int(im), int(pa))
    for i in spec.electrolyte:
        it.addParticleElectrolyte(int(i))
    vv.Context(system, it)
    nh = it.getParticlesNH()
    assert nh == sorted(nh) and len(nh) == 4 * 37 and min(nh) == 30
    assert it.isParticleLD(0) and not it.isParticleNH(0) and it.isParticleImage(int(spec.image_pairs[0, 0]))
    assert it.getNumMolecules() == spec.n_mol and it.getParticleMolId(0) == 0
    assert abs(it.getMoleculeInvMass(0) - 1.0 / spec.masses[spec.mol_id == 0].sum()) < 1e-18
    # a second context cannot bind the same integrator
    with pytest.raises(vv.OpenMMException, match="already bound"):
        vv.Context(system, it)
    # the reference's configuration errors surface with its texts
    it2 = vv.VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)
    it2.addParticleLangevin(int(nh[0]))
    with pytest.raises(vv.OpenMMException, match="same molecule"):
        vv.Context(system, it2)
    it3 = vv.VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)
    for i in spec.langevin:
        it3.addParticleLangevin(int(i))
    it3.setCosAcceleration(0.01)
    with pytest.raises(vv.OpenMMException, match="periodic perturbation"):
        vv.Context(system, it3)
    with pytest.raises(vv.OpenMMException, match="multiple DrudeForces"):
        vv.Context(vv.System(spec.masses, spec.bonds, spec.drude_pairs, num_drude_forces=2), vv.VVIntegrator(1, 1, 1, 1, 1))
    with pytest.raises(vv.OpenMMException, match="not bound"):
        vv.VVIntegrator(1, 1, 1, 1, 1).step(1)


def test_public_propagate_nh_chain(vv):
    it = vv.VVIntegrator(300.0, 10.0, 1.0, 40.0, 0.001)
    kT = 1.380649e-23 * 6.02214076e23 / 1000.0 * 300.0
    q = np.array([1000 * kT / 100, kT / 100, kT / 100])
    eta, ed, edd = np.zeros(3), np.zeros(4), np.zeros(3)
    f = it.propagateNHChain(eta, ed, edd, q, 1.1 * 1000 * kT, 1000 * kT, 300.0)
    assert abs(f - 0.9999987500007812) < 2e-16


@pytest.mark.gpu
def test_context_steps_like_run_bulk(vv, vo):
    """examples/run-bulk.py in miniature: Drude bulk, TGNH, middle scheme, hard wall 0.02 nm, frozen forces"""
    spec, system = bulk_system(vv, 100)
    it = vv.VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)
    it.setMaxDrudeDistance(0.02)
    ctx = vv.Context(system, it, precision="mixed")
    host = vv.make_state(spec, "mixed")
    ctx.setState(host)
    it.step(3)
    got = ctx.getState()
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    want = host.copy()
    vo.Oracle(spec, params, "mixed", literal=False).step(want, steps=3)
    from conftest import rel_err
    assert rel_err(got.velm[: spec.n, :3], want.velm[: spec.n, :3]) <= 1e-8
    assert rel_err(got.positions()[: spec.n], want.positions()[: spec.n]) <= 1e-8
    it.setStepSize(0.002)                       # picked up at the next step, like the reference
    it.step(1)
    assert np.isfinite(ctx.getState().velm).all()


@pytest.mark.gpu
def test_context_checkpoint_and_group_temperatures(vv, vo):
    """Context.createCheckpoint / loadCheckpoint carry the integrator's NH-chain state (the reference drops it), and
    getGroupTemperatures reports the three group temperatures from the device"""
    spec, system = bulk_system(vv, 60)
    it = vv.VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)
    it.setMaxDrudeDistance(0.02)
    ctx = vv.Context(system, it, precision="mixed")
    ctx.setState(vv.make_state(spec, "mixed"))
    it.step(3)
    blob = ctx.createCheckpoint()
    it.step(2)
    want = ctx.getState()
    ctx.loadCheckpoint(blob)
    it.step(2)
    got = ctx.getState()
    assert np.array_equal(got.velm, want.velm) and np.array_equal(got.posq, want.posq) and np.array_equal(got.corr, want.corr)
    t = it.getGroupTemperatures()
    assert t["num_temp_groups"] == 3 and np.all(t["temperature"] > 0) and np.all(t["dof"] > 0)
