/*
 * mini_api.cpp -- TEST INFRASTRUCTURE ONLY: a C interface (for ctypes) around "a Context with a VVIntegrator" under the
 * mini-OpenMM.  The integrator is the REFERENCE'S OWN VVIntegrator (openmmapi/src/VVIntegrator.cpp, compiled unchanged);
 * the kernels behind it are whichever set of Cuda*Kernel classes this library was linked with:
 *     libvvplugin_ref_{cpu,cuda}_<mode>.so    the reference's platforms/cuda/src/CudaVVKernels.cpp (+ its kernel sources)
 *     libvvplugin_glue_cuda_<mode>.so         csrc/glue/CudaVVKernelsB200.cpp forwarding to libvvb200.so   (-DMOMM_GLUE)
 * both registered by the reference's unchanged CudaVVKernelFactory.cpp.  momm_step() calls VVIntegrator::step(), so the
 * reference's own stepMiddle / stepVV issue the virtual calls.
 *
 * `private` members of the reference classes are read (never written) to report what initialize() computed -- DOFs, chain
 * masses, NkbT, the thermostat state: the class layout is untouched, only access control is waived for this TU.
 */
#include "mini_openmm.h"

#include <algorithm>
#include <cmath>
#include <iostream>
#include <set>

#define private public
#define protected public
#include "openmm/VVIntegrator.h"
#include "openmm/VVKernels.h"
#ifdef MOMM_GLUE
#include "CudaVVKernelsB200.h"
#else
#include "CudaVVKernels.h"
#endif
#undef private
#undef protected

using namespace OpenMM;

extern "C" void registerCudaVVKernelFactories();      /* the reference's CudaVVKernelFactory.cpp:57-65 */

extern "C" {

typedef struct {
    int32_t numParticles; const double *masses;
    int32_t numBonds; const int32_t *bonds;
    int32_t numDrude; const int32_t *drudePairs;                 /* (Drude, parent) in DrudeForce order */
    int32_t numConstraints; const int32_t *constraints; const double *constraintDistances;
    int32_t hasCMMotionRemover;
    int32_t numDrudeForces;                                      /* -1: one iff numDrude > 0 */
    int32_t numLD; const int32_t *particlesLD;
    int32_t numImagePairs; const int32_t *imagePairs;
    int32_t numElectrolyte; const int32_t *particlesElectrolyte;
} momm_system;

typedef struct {
    double temperature, frequency, drudeTemperature, drudeFrequency, stepSize;
    int32_t numNHChains, loopsPerStep;
    int32_t useCOMTempGroup;      /* -1: leave it to VVIntegrator::initialize's auto rule */
    int32_t useMiddleScheme;
    double maxDrudeDistance;
    double friction, drudeFriction;   /* < 0: leave the constructor / auto values */
    double mirrorLocation, electricField, cosAcceleration;
    int32_t randomNumberSeed, debug;
} momm_params;

}   // extern "C"

struct momm_ctx {
    System system;
    VVIntegrator *integrator = nullptr;
    Context *context = nullptr;
    std::vector<int32_t> scratchInt;
    std::vector<int32_t> sOffset, sAtoms;      /* host copies; the device copies live in standin */
    std::vector<double> sDistance;
    void *dOffset = nullptr, *dAtoms = nullptr, *dDistance = nullptr;
    ~momm_ctx() {
        delete context;
        delete integrator;
        miniomm::deviceFree(dOffset); miniomm::deviceFree(dAtoms); miniomm::deviceFree(dDistance);
    }
};

static thread_local std::string g_error;

extern "C" {

const char *momm_last_error(void) { return g_error.c_str(); }
int momm_is_cuda(void) { return miniomm::isCuda() ? 1 : 0; }
int momm_is_glue(void) {
#ifdef MOMM_GLUE
    return 1;
#else
    return 0;
#endif
}
int momm_precision_mode(void) { return MINIOMM_MODE; }

momm_ctx *momm_create(const momm_system *s, const momm_params *p, void *stream) {
    momm_ctx *c = new momm_ctx();
    try {
        registerCudaVVKernelFactories();
        for (int i = 0; i < s->numParticles; i++)
            c->system.addParticle(s->masses[i]);
        if (s->numBonds > 0) {
            BondListForce *bonds = new BondListForce();
            for (int i = 0; i < s->numBonds; i++)
                bonds->addBond(s->bonds[2 * i], s->bonds[2 * i + 1]);
            c->system.addForce(bonds);
        }
        const int nDrudeForces = s->numDrudeForces >= 0 ? s->numDrudeForces : (s->numDrude > 0 ? 1 : 0);
        for (int f = 0; f < nDrudeForces; f++) {
            DrudeForce *drude = new DrudeForce();
            if (f == 0)
                for (int i = 0; i < s->numDrude; i++)
                    drude->addParticle(s->drudePairs[2 * i], s->drudePairs[2 * i + 1], -1, -1, -1, -1.0, 0.001, 1.0, 1.0);
            c->system.addForce(drude);
        }
        for (int i = 0; i < s->numConstraints; i++)
            c->system.addConstraint(s->constraints[2 * i], s->constraints[2 * i + 1], s->constraintDistances ? s->constraintDistances[i] : 0.1);
        if (s->hasCMMotionRemover)
            c->system.addForce(new CMMotionRemover());

        VVIntegrator *vv = new VVIntegrator(p->temperature, p->frequency, p->drudeTemperature, p->drudeFrequency, p->stepSize,
                                            p->numNHChains, p->loopsPerStep);
        c->integrator = vv;
        vv->setMaxDrudeDistance(p->maxDrudeDistance);
        vv->setUseMiddleScheme(p->useMiddleScheme != 0);
        if (p->useCOMTempGroup >= 0) vv->setUseCOMTempGroup(p->useCOMTempGroup != 0);
        if (p->friction >= 0) vv->setFriction(p->friction);
        if (p->drudeFriction >= 0) {
            const bool autoFriction = vv->autoSetFriction;     /* setDrudeFriction clears it */
            vv->setDrudeFriction(p->drudeFriction);
            if (p->friction < 0) vv->autoSetFriction = autoFriction;
        }
        vv->setMirrorLocation(p->mirrorLocation);
        vv->setElectricField(p->electricField);
        vv->setCosAcceleration(p->cosAcceleration);
        vv->setRandomNumberSeed(p->randomNumberSeed);
        vv->setDebugEnabled(p->debug != 0);
        for (int i = 0; i < s->numLD; i++) vv->addParticleLangevin(s->particlesLD[i]);
        for (int i = 0; i < s->numImagePairs; i++) vv->addImagePair(s->imagePairs[2 * i], s->imagePairs[2 * i + 1]);
        for (int i = 0; i < s->numElectrolyte; i++) vv->addParticleElectrolyte(s->particlesElectrolyte[i]);

        c->context = new Context(c->system, *vv, Platform::getPlatformByName("CUDA"), stream);    /* -> VVIntegrator::initialize */
    } catch (const std::exception &e) {
        g_error = e.what();
        delete c;
        return nullptr;
    }
    return c;
}

void momm_destroy(momm_ctx *c) { delete c; }

static CudaContext &cuOf(momm_ctx *c) { return c->context->getImpl().cuda(); }

/* host arrays in OpenMM's device layouts -> the context's arrays; random: float4[nRandom] injected N(0,1) stream */
int momm_set_state(momm_ctx *c, const void *posq, const void *corr, const void *velm, const long long *force,
                   const float *random, int64_t nRandom, const double *box) {
    try {
        CudaContext &cu = cuOf(c);
        cu.getPosq().upload(posq);
        if (cu.getUseMixedPrecision() && corr) cu.getPosqCorrection().upload(corr);
        cu.getVelm().upload(velm);
        cu.getForce().upload(force);
        cu.getIntegrationUtilities().setRandomStream(random, (size_t) nRandom);
        if (box) cu.setPeriodicBoxSize(box[0], box[1], box[2]);
        c->integrator->stateChanged(State::Positions);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

int momm_get_state(momm_ctx *c, void *posq, void *corr, void *velm) {
    try {
        CudaContext &cu = cuOf(c);
        if (posq) cu.getPosq().download(posq);
        if (corr && cu.getUseMixedPrecision()) cu.getPosqCorrection().download(corr);
        if (velm) cu.getVelm().download(velm);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

/* device (or, host flavour, host) addresses of posq, posqCorrection, velm, force, posDelta, random */
void momm_array_pointers(momm_ctx *c, void **out6) {
    CudaContext &cu = cuOf(c);
    out6[0] = (void *) cu.getPosq().getDevicePointer();
    out6[1] = cu.getUseMixedPrecision() ? (void *) cu.getPosqCorrection().getDevicePointer() : nullptr;
    out6[2] = (void *) cu.getVelm().getDevicePointer();
    out6[3] = (void *) cu.getForce().getDevicePointer();
    out6[4] = (void *) cu.getIntegrationUtilities().getPosDelta().getDevicePointer();
    out6[5] = (void *) cu.getIntegrationUtilities().getRandom().getDevicePointer();
}

void momm_set_constraint_standin(momm_ctx *c, int numClusters, const int32_t *clusterOffset, const int32_t *atoms,
                                 const double *distance, int iterations) {
    CudaContext &cu = cuOf(c);
    miniomm::Standin &s = cu.getIntegrationUtilities().standin;
    miniomm::deviceFree(c->dOffset); miniomm::deviceFree(c->dAtoms); miniomm::deviceFree(c->dDistance);
    c->dOffset = c->dAtoms = c->dDistance = nullptr;
    s = miniomm::Standin();
    if (numClusters <= 0)
        return;
    const int nCons = clusterOffset[numClusters];
    c->dOffset = miniomm::deviceAlloc((numClusters + 1) * sizeof(int32_t));
    c->dAtoms = miniomm::deviceAlloc((2 * (size_t) nCons + 1) * sizeof(int32_t));
    c->dDistance = miniomm::deviceAlloc(((size_t) nCons + 1) * sizeof(double));
    miniomm::copyToDevice(c->dOffset, clusterOffset, (numClusters + 1) * sizeof(int32_t), cu.stream);
    miniomm::copyToDevice(c->dAtoms, atoms, 2 * (size_t) nCons * sizeof(int32_t), cu.stream);
    miniomm::copyToDevice(c->dDistance, distance, (size_t) nCons * sizeof(double), cu.stream);
    s.numClusters = numClusters;
    s.iterations = iterations;
    s.offset = (int32_t *) c->dOffset;
    s.atoms = (int32_t *) c->dAtoms;
    s.distance = (double *) c->dDistance;
}

int momm_step(momm_ctx *c, int steps) {
    try {
        c->integrator->step(steps);
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

void momm_set_step_size(momm_ctx *c, double dt) { c->integrator->setStepSize(dt); }

/* the bytes of the LAST upload() to the CudaArray called `name` ("vvDrudePairs", "particlesSortedByMolId", ...) */
int momm_get_upload(momm_ctx *c, const char *name, const void **ptr, int64_t *bytes, int32_t *elementSize) {
    CudaContext &cu = cuOf(c);
    std::map<std::string, std::vector<unsigned char> >::const_iterator it = cu.uploads.find(name);
    if (it == cu.uploads.end())
        return -1;
    *ptr = it->second.data();
    *bytes = (int64_t) it->second.size();
    *elementSize = cu.uploadElementSize[name];
    return 0;
}

/* what VVIntegrator::initialize built: 0 particlesNH, 1 moleculesNH, 2 particleMolId, 3 particlesLD, 4 particlesElectrolyte */
int momm_get_int_list(momm_ctx *c, int which, const int32_t **ptr, int64_t *n) {
    const VVIntegrator &vv = *c->integrator;
    const std::vector<int> *v = nullptr;
    switch (which) {
    case 0: v = &vv.getParticlesNH(); break;
    case 1: v = &vv.getMoleculesNH(); break;
    case 2: v = &vv.particleMolId; break;
    case 3: v = &vv.getParticlesLD(); break;
    case 4: v = &vv.getParticlesElectrolyte(); break;
    default: return -1;
    }
    c->scratchInt.assign(v->begin(), v->end());
    *ptr = c->scratchInt.data();
    *n = (int64_t) c->scratchInt.size();
    return 0;
}

enum { MOMM_F64_MOLECULE_MASSES = 0, MOMM_F64_MOLECULE_INV_MASSES, MOMM_F64_DOF, MOMM_F64_ETA_MASS, MOMM_F64_NKBT,
       MOMM_F64_KE2, MOMM_F64_VSCALE, MOMM_F64_ETA, MOMM_F64_ETA_DOT, MOMM_F64_ETA_DOTDOT, MOMM_F64_INV_MASS_TOTAL,
       MOMM_F64_SETTINGS /* friction, drudeFriction, useCOMTempGroup, numTempGroup */ };

#ifdef MOMM_GLUE
static vvb200_plan *planOf(momm_ctx *c) { return VVB200Shared::get(cuOf(c), false)->plan; }
#endif

/* returns the number of doubles written (<= cap), -1 if unavailable */
int momm_get_f64(momm_ctx *c, int which, double *out, int cap) {
    VVIntegrator &vv = *c->integrator;
    std::vector<double> v;
    try {
        if (which == MOMM_F64_MOLECULE_MASSES) v = vv.moleculeMasses;
        else if (which == MOMM_F64_MOLECULE_INV_MASSES) v = vv.moleculeInvMasses;
        else if (which == MOMM_F64_SETTINGS) {
            v.push_back(vv.friction); v.push_back(vv.drudeFriction); v.push_back(vv.getUseCOMTempGroup() ? 1 : 0);
#ifdef MOMM_GLUE
            v.push_back(vv.getParticlesNH().empty() ? 0 : vvb200_plan_num_temp_groups(planOf(c)));
#else
            v.push_back(vv.getParticlesNH().empty() ? 0 : vv.nhKernel.getAs<CudaModifyDrudeNoseKernel>().numTempGroup);
#endif
        } else {
#ifdef MOMM_GLUE
            vvb200_plan *plan = planOf(c);
            const int ng = vvb200_plan_num_temp_groups(plan), nc = vv.getNumNHChains();
            const double *ptr = nullptr;
            int64_t len = 0;
            vvb200_thermostat_state st;
            switch (which) {
            case MOMM_F64_DOF: vvb200_plan_get_f64_array(plan, VVB200_F64_DOF, &ptr, &len); v.assign(ptr, ptr + len); break;
            case MOMM_F64_ETA_MASS: vvb200_plan_get_f64_array(plan, VVB200_F64_ETA_MASS, &ptr, &len); v.assign(ptr, ptr + len); break;
            case MOMM_F64_NKBT: vvb200_plan_get_f64_array(plan, VVB200_F64_NKBT, &ptr, &len); v.assign(ptr, ptr + len); break;
            case MOMM_F64_INV_MASS_TOTAL: vvb200_plan_get_f64_array(plan, VVB200_F64_INV_MASS_TOTAL, &ptr, &len); v.assign(ptr, ptr + len); break;
            default:
                if (vvb200_get_thermostat_state(plan, &st, cuOf(c).getCurrentStream()) != VVB200_OK)
                    throw OpenMMException(vvb200_last_error());
                if (which == MOMM_F64_KE2) v.assign(st.ke2, st.ke2 + ng);
                else if (which == MOMM_F64_VSCALE) v.assign(st.vscale, st.vscale + ng);
                else if (which == MOMM_F64_ETA) v.assign(st.eta, st.eta + ng * nc);
                else if (which == MOMM_F64_ETA_DOT) v.assign(st.eta_dot, st.eta_dot + ng * (nc + 1));
                else if (which == MOMM_F64_ETA_DOTDOT) v.assign(st.eta_dotdot, st.eta_dotdot + ng * nc);
                else return -1;
            }
#else
            if (which == MOMM_F64_INV_MASS_TOTAL) {
                if (vv.getCosAcceleration() == 0) return -1;
                v.push_back(vv.ppKernel.getAs<CudaModifyCosineAccelerateKernel>().invMassTotal);
            } else {
                if (vv.getParticlesNH().empty()) return -1;
                CudaModifyDrudeNoseKernel &nh = vv.nhKernel.getAs<CudaModifyDrudeNoseKernel>();
                auto flat = [&](const std::vector<std::vector<double> > &m) { for (size_t g = 0; g < m.size(); g++) v.insert(v.end(), m[g].begin(), m[g].end()); };
                switch (which) {
                case MOMM_F64_DOF: v = nh.tempGroupDof; break;
                case MOMM_F64_ETA_MASS: flat(nh.etaMass); break;
                case MOMM_F64_NKBT: v = nh.tempGroupNkbT; break;
                case MOMM_F64_KE2: v = nh.kineticEnergiesNHVec; break;
                case MOMM_F64_VSCALE: v = nh.vscaleFactorsNHVec; break;
                case MOMM_F64_ETA: flat(nh.eta); break;
                case MOMM_F64_ETA_DOT: flat(nh.etaDot); break;
                case MOMM_F64_ETA_DOTDOT: flat(nh.etaDotDot); break;
                default: return -1;
                }
            }
#endif
        }
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    const int n = std::min((int) v.size(), cap);
    for (int i = 0; i < n; i++) out[i] = v[i];
    return n;
}

int momm_get_viscosity(momm_ctx *c, double *out2) {
    try {
        std::vector<double> r = c->integrator->getViscosity();
        out2[0] = r[0]; out2[1] = r[1];
    } catch (const std::exception &e) {
        g_error = e.what();
        return -1;
    }
    return 0;
}

/* kernel launches issued through CudaContext::executeKernel (the reference's kernels; 0 for the glue, whose launches are
 * counted by vvb200_launch_count), force evaluations, stand-in constraint calls (positions, velocities), reorderAtoms calls,
 * CudaContext step count, whether a ForceInfo was registered before initializeContexts, vvb200 launches (glue) */
void momm_counters(momm_ctx *c, long long *out8) {
    CudaContext &cu = cuOf(c);
    out8[0] = cu.kernelLaunches;
    out8[1] = c->context->getImpl().forceEvaluations;
    out8[2] = cu.getIntegrationUtilities().constraintCalls;
    out8[3] = cu.getIntegrationUtilities().velocityConstraintCalls;
    out8[4] = cu.reorderCalls;
    out8[5] = cu.getStepCount();
    out8[6] = cu.getForceInfos().empty() ? -1 : (cu.forceInfoAddedBeforeInit ? 1 : 0);
#ifdef MOMM_GLUE
    out8[7] = vvb200_launch_count(planOf(c));
#else
    out8[7] = 0;
#endif
}

/* areParticlesIdentical(i, j) of every registered ForceInfo ANDed (what CudaContext::findMoleculeGroups asks) */
int momm_particles_identical(momm_ctx *c, int i, int j) {
    std::vector<ComputeForceInfo *> &infos = cuOf(c).getForceInfos();
    for (size_t k = 0; k < infos.size(); k++)
        if (!infos[k]->areParticlesIdentical(i, j)) return 0;
    return 1;
}

}   // extern "C"
