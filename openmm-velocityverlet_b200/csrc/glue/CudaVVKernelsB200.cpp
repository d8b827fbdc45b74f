// CudaVVKernelsB200.cpp -- the plugin's seven CUDA kernel classes as thin forwards to libvvb200.
//
// Replaces platforms/cuda/src/CudaVVKernels.cpp of the reference (1134 lines of array set-up, NVRTC modules and
// launches).  VVIntegrator.cpp keeps calling the same virtuals in the same order (VVIntegrator.cpp:232-338); each
// one maps onto the C ABI as follows.
//
// Middle scheme, no OpenMM constraints / virtual sites in the System (the fused fast path: ONE launch per step while
// the system fits in shared memory, ~100k particles; two streaming passes beyond that):
//   resetExtraForce / applyElectricForce / applyCosineForce      no-ops: extra forces are evaluated inside the kick
//   applyLangevinForce                                            prepareRandomNumbers (same request size) only
//   firstIntegrate                                                no-op (nothing of OpenMM's runs before secondIntegrate)
//   calc/remove/restoreVelocityBias, scaleVelocity                no-ops: folded into the step
//   secondIntegrate                                               vvb200_step_middle  (kick + reductions + NH chains +
//                                                                 scale + drift + hard wall + image mirror)
//   updateImagePositions                                          no-op (done by whichever call wrote the positions)
// With constraints or virtual sites OpenMM's solvers must run between the sub-steps, so the split entry points are
// used: kick -> applyVelocityConstraints | (no-op) | thermostat_delta -> applyConstraints -> finish  (440 B/particle).
// Velocity-Verlet scheme without constraints: firstIntegrate = vvb200_step_vv_first (thermostat half step + half kick +
// drift), secondIntegrate = vvb200_step_vv_second (half kick + thermostat half step); both scaleVelocity calls are no-ops.
// With constraints: thermostat | vv_kick(+posDelta) -> applyConstraints -> vv_positions | vv_kick -> applyVelocityConstraints |
// thermostat.
#include "CudaVVKernelsB200.h"

#include <iostream>
#include <map>
#include <mutex>

#include "openmm/CMMotionRemover.h"
#include "openmm/OpenMMException.h"
#include "openmm/common/ContextSelector.h"
#include "openmm/internal/ContextImpl.h"
#include "CudaIntegrationUtilities.h"

using namespace OpenMM;
using namespace std;

#define VVB200_CHECK(call)                                                      \
    do {                                                                        \
        if ((call) != VVB200_OK) throw OpenMMException(vvb200_last_error());    \
    } while (0)

// ---- shared state -----------------------------------------------------------------------------------------
static mutex g_lock;
static map<CudaContext *, weak_ptr<VVB200Shared> > g_shared;

shared_ptr<VVB200Shared> VVB200Shared::get(CudaContext &cu, bool create) {
    lock_guard<mutex> guard(g_lock);
    shared_ptr<VVB200Shared> sp = g_shared[&cu].lock();
    if (!sp && create) {
        sp = make_shared<VVB200Shared>();
        g_shared[&cu] = sp;
    }
    if (!sp)
        throw OpenMMException("VVB200: the step kernel must be initialized before the modifier kernels");
    return sp;
}

static vvb200_buffers deviceBuffers(CudaContext &cu) {
    CudaIntegrationUtilities &integration = cu.getIntegrationUtilities();
    vvb200_buffers b;
    b.posq = (void *) cu.getPosq().getDevicePointer();
    b.posq_correction = cu.getUseMixedPrecision() ? (void *) cu.getPosqCorrection().getDevicePointer() : nullptr;
    b.velm = (void *) cu.getVelm().getDevicePointer();
    b.force = (const long long *) cu.getForce().getDevicePointer();
    b.pos_delta = (void *) integration.getPosDelta().getDevicePointer();
    b.random = (const void *) integration.getRandom().getDevicePointer();
    return b;
}

static vvb200_step_args stepArgs(CudaContext &cu, const VVB200Shared &sh, const VVIntegrator &integrator) {
    vvb200_step_args a;
    a.random_index = sh.randomIndex;
    a.inv_box_z = integrator.getCosAcceleration() != 0 ? 1.0 / cu.getPeriodicBoxSize().z : 0.0;
    return a;
}

static void syncStepSize(CudaContext &cu, VVB200Shared &sh, const VVIntegrator &integrator) {
    // The reference re-reads the step size every step and keeps OpenMM's own copy current: setNextStepSize in the middle
    // scheme (CudaVVKernels.cpp:137-141), an upload of (0, dt) into integration.getStepSize() in the velocity-Verlet
    // scheme (:309-319).  Other OpenMM code reads that array (time-shifted kinetic energy, constraint kernels), so it is
    // kept in step here too, although the vvb200 kernels take dt as a parameter.
    const double stepSize = integrator.getStepSize();
    if (stepSize == sh.stepSize && sh.stepSizePublished)
        return;
    CudaIntegrationUtilities &integration = cu.getIntegrationUtilities();
    if (integrator.getUseMiddleScheme()) {
        integration.setNextStepSize(stepSize);
    } else if (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) {
        double2 ss = make_double2(0, stepSize);
        integration.getStepSize().upload(&ss);
    } else {
        float2 ss = make_float2(0, (float) stepSize);
        integration.getStepSize().upload(&ss);
    }
    if (stepSize != sh.stepSize)
        VVB200_CHECK(vvb200_set_step_size(sh.plan, stepSize));
    sh.stepSize = stepSize;
    sh.stepSizePublished = true;
}

// What makes two particles different for THIS plugin.  CudaContext::reorderAtoms only ever swaps molecules all of whose
// ForceInfos call them identical [OMM-mem]; the reference registers none for its particle sets, which is why its
// README (README.md:189-194) asks users to put whole classes of identical molecules into the Langevin / electrolyte /
// image sets.  With this ForceInfo a molecule whose members carry different plugin roles is simply never swapped with
// one that does not, so any subset is safe.  (The per-slot tables of the plan describe array slots; a swap of two
// molecules that are identical in this sense leaves them valid.)
class VVB200ForceInfo : public CudaForceInfo {
public:
    VVB200ForceInfo(int numParticles, const VVIntegrator &integrator) : role(numParticles, 0) {
        for (int p : integrator.getParticlesLD()) role[p] |= 1;
        for (const pair<int, int> &ip : integrator.getImagePairs()) {
            role[ip.first] |= 2;       // image
            role[ip.second] |= 4;      // has an image
        }
        for (int p : integrator.getParticlesElectrolyte()) role[p] += 8;     // multiplicity matters (field is added per entry)
    }
    bool areParticlesIdentical(int particle1, int particle2) override { return role[particle1] == role[particle2]; }
private:
    vector<int> role;
};

// Builds the plan from what VVIntegrator::initialize and the reference's Cuda*Kernel::initialize methods read
// (VVIntegrator.cpp:123-151; CudaVVKernels.cpp:66-77, 483-594, 775-804, 884-891, 954-957, 1028-1031).
static void createPlan(VVB200Shared &sh, CudaContext &cu, const System &system, const VVIntegrator &integrator,
                       const DrudeForce *force) {
    ContextSelector selector(cu);
    // BEFORE initializeContexts: that call is what runs CudaContext::initialize() -> findMoleculeGroups(), the only
    // consumer of ForceInfos.  Registered afterwards it would never be consulted and molecules with different plugin
    // roles could still be swapped by reorderAtoms.  (It needs the System and the integrator's lists, not the plan.)
    const int n = system.getNumParticles();
    cu.addForce(new VVB200ForceInfo(n, integrator));     // owned by the context, like every ForceInfo
    cu.getPlatformData().initializeContexts(system);
    cu.getIntegrationUtilities().initRandomNumberGenerator((unsigned int) integrator.getRandomNumberSeed());

    vector<double> masses(n);
    vector<int32_t> molId(n);
    bool virtualSites = false;
    for (int i = 0; i < n; i++) {
        masses[i] = system.getParticleMass(i);
        molId[i] = integrator.getParticleMolId(i);
        virtualSites = virtualSites || system.isVirtualSite(i);
    }
    vector<int32_t> drude, cons, images;
    if (force != NULL)
        for (int i = 0; i < force->getNumParticles(); i++) {
            int p, p1, p2, p3, p4;
            double charge, polarizability, aniso12, aniso34;
            force->getParticleParameters(i, p, p1, p2, p3, p4, charge, polarizability, aniso12, aniso34);
            drude.push_back(p);
            drude.push_back(p1);
        }
    for (int i = 0; i < system.getNumConstraints(); i++) {
        int p1, p2;
        double distance;
        system.getConstraintParameters(i, p1, p2, distance);
        cons.push_back(p1);
        cons.push_back(p2);
    }
    for (const pair<int, int> &ip : integrator.getImagePairs()) {
        images.push_back(ip.first);
        images.push_back(ip.second);
    }
    bool cmm = false;
    for (int i = 0; i < system.getNumForces(); i++)
        cmm = cmm || dynamic_cast<const CMMotionRemover *>(&system.getForce(i)) != NULL;
    vector<int32_t> ld(integrator.getParticlesLD().begin(), integrator.getParticlesLD().end());
    vector<int32_t> electrolyte(integrator.getParticlesElectrolyte().begin(), integrator.getParticlesElectrolyte().end());

    vvb200_system s;
    s.num_particles = n;
    s.padded_num_atoms = cu.getPaddedNumAtoms();
    s.num_molecules = integrator.getNumMolecules();
    s.masses = masses.data();
    s.particle_mol_id = molId.data();
    s.num_drude = (int32_t) drude.size() / 2;
    s.drude_pairs = drude.data();
    s.num_constraints = (int32_t) cons.size() / 2;
    s.constraints = cons.data();
    s.has_cm_motion_remover = cmm;
    s.num_langevin = (int32_t) ld.size();
    s.particles_langevin = ld.data();
    s.num_image_pairs = (int32_t) images.size() / 2;
    s.image_pairs = images.data();
    s.num_electrolyte = (int32_t) electrolyte.size();
    s.particles_electrolyte = electrolyte.data();

    vvb200_params par;
    par.temperature = integrator.getTemperature();
    par.frequency = integrator.getFrequency();
    par.drude_temperature = integrator.getDrudeTemperature();
    par.drude_frequency = integrator.getDrudeFrequency();
    par.step_size = integrator.getStepSize();
    par.num_nh_chains = integrator.getNumNHChains();
    par.loops_per_step = integrator.getLoopsPerStep();
    par.use_com_temp_group = integrator.getUseCOMTempGroup();
    par.use_middle_scheme = integrator.getUseMiddleScheme();
    par.max_drude_distance = integrator.getMaxDrudeDistance();
    par.friction = integrator.getFriction();
    par.drude_friction = integrator.getDrudeFriction();
    par.mirror_location = integrator.getMirrorLocation();
    par.electric_field = integrator.getElectricField();
    par.cos_acceleration = integrator.getCosAcceleration();

    const int precision = cu.getUseDoublePrecision() ? VVB200_DOUBLE : cu.getUseMixedPrecision() ? VVB200_MIXED : VVB200_SINGLE;
    VVB200_CHECK(vvb200_plan_create(&s, &par, precision, &sh.plan));     // reference's exception texts on conflicts
    VVB200_CHECK(vvb200_plan_upload(sh.plan, cu.getCurrentStream()));
    sh.constrained = system.getNumConstraints() > 0 || virtualSites;
    sh.hasNH = !integrator.getParticlesNH().empty();
    sh.stepSize = par.step_size;
    cerr << "vvb200 (sm_100a) integration path created\n"
         << "    NUM_ATOMS: " << n << ", PADDED_NUM_ATOMS: " << s.padded_num_atoms << "\n"
         << "    Num Drude pairs: " << s.num_drude << ", Drude hardwall distance: " << par.max_drude_distance << " nm\n"
         << "    Num temperature groups: " << vvb200_plan_num_temp_groups(sh.plan)
         << ", fused two-pass path: " << (sh.constrained ? "no (constraints present)" : "yes") << "\n" << flush;
}

static void finishStep(CudaContext &cu, double stepSize) {
    cu.setTime(cu.getTime() + stepSize);
    cu.setStepCount(cu.getStepCount() + 1);
}

// ---- middle scheme ---------------------------------------------------------------------------------------------
void CudaIntegrateMiddleStepKernel::initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force) {
    sh = VVB200Shared::get(cu, true);
    createPlan(*sh, cu, system, integrator, force);
}

void CudaIntegrateMiddleStepKernel::resetExtraForce(ContextImpl &, const VVIntegrator &) {
    // nothing to reset: there is no forceExtra array (extra forces are computed inside the kick)
}

void CudaIntegrateMiddleStepKernel::firstIntegrate(ContextImpl &, const VVIntegrator &integrator) {
    ContextSelector selector(cu);
    syncStepSize(cu, *sh, integrator);
    vvb200_buffers b = deviceBuffers(cu);
    vvb200_step_args a = stepArgs(cu, *sh, integrator);
    if (!sh->constrained)
        return;        // the whole step runs in secondIntegrate (vvb200_step_middle): no OpenMM solver sits in between
    VVB200_CHECK(vvb200_middle_kick(sh->plan, &b, &a, cu.getCurrentStream()));
    cu.getIntegrationUtilities().applyVelocityConstraints(integrator.getConstraintTolerance());
}

void CudaIntegrateMiddleStepKernel::secondIntegrate(ContextImpl &, const VVIntegrator &integrator) {
    ContextSelector selector(cu);
    vvb200_buffers b = deviceBuffers(cu);
    vvb200_step_args a = stepArgs(cu, *sh, integrator);
    CudaIntegrationUtilities &integration = cu.getIntegrationUtilities();
    if (!sh->constrained) {
        VVB200_CHECK(vvb200_step_middle(sh->plan, &b, &a, cu.getCurrentStream()));
    } else {
        // thermostat (bias remove / restore included) + both half drifts into posDelta / oldDelta, fused
        VVB200_CHECK(vvb200_middle_thermostat_delta(sh->plan, &b, &a, cu.getCurrentStream()));
        integration.applyConstraints(integrator.getConstraintTolerance());
        VVB200_CHECK(vvb200_middle_finish(sh->plan, &b, cu.getCurrentStream()));
    }
    integration.computeVirtualSites();
    cu.reorderAtoms();
    finishStep(cu, integrator.getStepSize());
}

double CudaIntegrateMiddleStepKernel::computeKineticEnergy(ContextImpl &, const VVIntegrator &) {
    return cu.getIntegrationUtilities().computeKineticEnergy(0);
}

// ---- velocity-Verlet scheme ---------------------------------------------------------------------------------------
void CudaIntegrateVVStepKernel::initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force) {
    sh = VVB200Shared::get(cu, true);
    createPlan(*sh, cu, system, integrator, force);
}

void CudaIntegrateVVStepKernel::resetExtraForce(ContextImpl &, const VVIntegrator &) {
}

void CudaIntegrateVVStepKernel::firstIntegrate(ContextImpl &, const VVIntegrator &integrator) {
    ContextSelector selector(cu);
    syncStepSize(cu, *sh, integrator);
    vvb200_buffers b = deviceBuffers(cu);
    vvb200_step_args a = stepArgs(cu, *sh, integrator);
    CudaIntegrationUtilities &integration = cu.getIntegrationUtilities();
    if (!sh->constrained) {
        // no OpenMM solver between the sub-steps: thermostat half step (the scaleVelocity call that preceded this one was
        // deferred to here) + half kick + drift + hard wall + image mirror, fused (1 launch resident, 2 streaming)
        VVB200_CHECK(vvb200_step_vv_first(sh->plan, &b, &a, cu.getCurrentStream()));
    } else {
        VVB200_CHECK(vvb200_vv_kick(sh->plan, &b, &a, 0, 1, cu.getCurrentStream()));
        integration.applyConstraints(integrator.getConstraintTolerance());
        VVB200_CHECK(vvb200_vv_positions(sh->plan, &b, cu.getCurrentStream()));
    }
    integration.computeVirtualSites();
    cu.reorderAtoms();          // after the first half, like the reference (CudaVVKernels.cpp:376-381)
}

void CudaIntegrateVVStepKernel::secondIntegrate(ContextImpl &, const VVIntegrator &integrator) {
    ContextSelector selector(cu);
    vvb200_buffers b = deviceBuffers(cu);
    vvb200_step_args a = stepArgs(cu, *sh, integrator);
    if (!sh->constrained) {
        // extra forces + half kick + the thermostat half step that follows (the next scaleVelocity call is a no-op), fused
        VVB200_CHECK(vvb200_step_vv_second(sh->plan, &b, &a, cu.getCurrentStream()));
    } else {
        VVB200_CHECK(vvb200_vv_kick(sh->plan, &b, &a, 1, 0, cu.getCurrentStream()));
        cu.getIntegrationUtilities().applyVelocityConstraints(integrator.getConstraintTolerance());
    }
    finishStep(cu, integrator.getStepSize());
}

double CudaIntegrateVVStepKernel::computeKineticEnergy(ContextImpl &, const VVIntegrator &) {
    return cu.getIntegrationUtilities().computeKineticEnergy(0);
}

// ---- thermostat ------------------------------------------------------------------------------------------------------
void CudaModifyDrudeNoseKernel::initialize(const System &, const VVIntegrator &, const DrudeForce *) {
    sh = VVB200Shared::get(cu, false);      // index arrays, DOFs and chain masses live in the plan
}

void CudaModifyDrudeNoseKernel::scaleVelocity(ContextImpl &, const VVIntegrator &integrator) {
    if (integrator.getUseMiddleScheme() || !sh->constrained)
        return;                             // fused into the calls first/secondIntegrate make (see the table at the top)
    ContextSelector selector(cu);
    vvb200_buffers b = deviceBuffers(cu);
    vvb200_step_args a = stepArgs(cu, *sh, integrator);
    VVB200_CHECK(vvb200_thermostat(sh->plan, &b, &a, cu.getCurrentStream()));   // bias remove/restore included
}

// ---- Langevin -----------------------------------------------------------------------------------------------------------
void CudaModifyDrudeLangevinKernel::initialize(const System &, const VVIntegrator &, const DrudeForce *, Kernel &) {
    sh = VVB200Shared::get(cu, false);
}

void CudaModifyDrudeLangevinKernel::applyLangevinForce(ContextImpl &, const VVIntegrator &) {
    // consume OpenMM's random stream exactly like the reference (padded request, CudaVVKernels.cpp:863); the force
    // itself is evaluated by the kick that follows, from the same random numbers
    sh->randomIndex = (unsigned int) cu.getIntegrationUtilities().prepareRandomNumbers((int) vvb200_plan_random_request(sh->plan));
}

// ---- image charges ---------------------------------------------------------------------------------------------------
void CudaModifyImageChargeKernel::initialize(const System &, const VVIntegrator &) {
    sh = VVB200Shared::get(cu, false);
}

void CudaModifyImageChargeKernel::updateImagePositions(ContextImpl &, const VVIntegrator &) {
    // nothing left to do: every call that writes positions -- vvb200_step_middle, vvb200_middle_finish,
    // vvb200_vv_positions -- already mirrored the images of the particles it moved (include/vvb200.h)
}

// ---- electric field / cosine acceleration: forces are evaluated inside the kick ----------------------------------------
void CudaModifyElectricFieldKernel::initialize(const System &, const VVIntegrator &, Kernel &) {
}

void CudaModifyElectricFieldKernel::applyElectricForce(ContextImpl &, const VVIntegrator &) {
}

void CudaModifyCosineAccelerateKernel::initialize(const System &, const VVIntegrator &, Kernel &) {
    sh = VVB200Shared::get(cu, false);
}

void CudaModifyCosineAccelerateKernel::applyCosineForce(ContextImpl &, const VVIntegrator &) {
}

void CudaModifyCosineAccelerateKernel::calcVelocityBias(ContextImpl &, const VVIntegrator &) {
}

void CudaModifyCosineAccelerateKernel::removeVelocityBias(ContextImpl &, const VVIntegrator &) {
}

void CudaModifyCosineAccelerateKernel::restoreVelocityBias(ContextImpl &, const VVIntegrator &) {
}

void CudaModifyCosineAccelerateKernel::calcViscosity(ContextImpl &, const VVIntegrator &, double &vMax, double &invVis) {
    ContextSelector selector(cu);
    const double4 box = cu.getPeriodicBoxSize();
    VVB200_CHECK(vvb200_calc_viscosity(sh->plan, box.x, box.y, box.z, &vMax, &invVis, cu.getCurrentStream()));
}
