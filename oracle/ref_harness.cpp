/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the REFERENCE'S OWN kernel sources -- #included from where they lie under
 * /root/reference/platforms/cuda/src/kernels (never copied into this repository) -- with a
 * restatement of the reference's host schedule.  Built by oracle/Makefile into oracle/_ref/
 * (git-ignored) in two flavours per precision mode:
 *
 *   libvvref_cpu_<mode>.so   g++ : the kernels are compiled as host C++; a tiny SIMT shim
 *                            (blockIdx/threadIdx as thread-locals, __syncthreads a no-op) runs
 *                            every (block, thread) of OpenMM's launch geometry, blocks spread
 *                            over OpenMP threads.  Runs without a GPU; used to pin the plain-C
 *                            oracle (tests/test_oracle_vs_ref.py) and as the CPU baseline.
 *   libvvref_cuda_<mode>.so  nvcc -gencode arch=compute_100a,code=sm_100a : the same kernels
 *                            on the GPU with OpenMM's launch geometry (block 64, grid <=
 *                            numThreadBlocks; the two sums 1 x 512) and the reference's blocking
 *                            D2H / H2D around the host NH-chain.  Used as the on-GPU parity
 *                            target and the measured "reference kernels on B200" baseline.
 *
 * What is restated here (because it needs OpenMM to compile in the reference):
 *   - the prelude CudaContext::createModule prepends (typedefs real/mixed, make_*, SQRT, RECIP,
 *     USE_*_PRECISION) [OMM-mem];
 *   - per-module #defines (NUM_ATOMS, ...) -> variables (constant memory on the GPU);
 *   - CudaContext::executeKernel's geometry [OMM-mem];
 *   - the host methods of CudaVVKernels.cpp (:119-231, 296-431, 670-754, 826-872, 904-934,
 *     971-992, 1037-1134) and the schedules VVIntegrator.cpp:232-338 and NHC :340-376.
 * Index arrays are INPUTS (built by the caller): this harness pins kernel arithmetic only.
 *
 * Deviation on the CPU flavour: the two single-block reductions (sumNormalizedKineticEnergies,
 * sumV) use __syncthreads() inside a tree, which a sequential SIMT shim cannot interleave; they
 * are run with blockDim = 1 (then the kernel's own first loop sums everything and the tree loop
 * is empty).  The sum is the same up to floating-point reassociation.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ref_kernels.inc.h"

/* ---- stand-in for OpenMM's constraint solvers between the sub-steps (oracle/constraint_standin.h;
 *      NOT reference code): applied where the reference calls integration.applyConstraints /
 *      applyVelocityConstraints (CudaVVKernels.cpp:151,176,351,427) -------------------------------- */
#define VVC_REAL4 real4
#define VVC_MIXED4 mixed4
#define VVC_MIXED mixed
#if VVREF_GPU
#define VVC_FN __host__ __device__ inline
#else
#define VVC_FN static inline
#endif
#include "constraint_standin.h"
#if VVREF_GPU
__global__ void standinPositionsKernel(vvc_constraints cs, const real4 *posq, const real4 *corr, const mixed4 *velm, mixed4 *posDelta) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)
        vvc_cluster_positions(cs, c, posq, corr, velm, posDelta);
}
__global__ void standinVelocitiesKernel(vvc_constraints cs, const real4 *posq, const real4 *corr, mixed4 *velm) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)
        vvc_cluster_velocities(cs, c, posq, corr, velm);
}
#endif

/* ---------------------------------------------------------------------------------------- */
static const double VVREF_BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0;
static const double VVREF_AVOGADRO = 6.02214076e23;
#define VVREF_MAX_CHAINS 16

extern "C" {

typedef struct {
    int32_t numAtoms, paddedNumAtoms, numMolecules;
    int32_t numDrude; const int32_t *drudePairs;
    int32_t numParticlesNH; const int32_t *particlesNH;
    int32_t numMoleculesNH; const int32_t *moleculesNH;
    int32_t numNormalNH; const int32_t *normalParticlesNH;
    int32_t numPairsNH; const int32_t *pairParticlesNH;
    const int32_t *particleMolId;            /* [numAtoms] */
    const int32_t *particlesInMolecules;     /* int2[numMolecules] (count,start) */
    const int32_t *particlesSortedByMolId;   /* [numAtoms] */
    int32_t numTempGroup;
    const double *etaMass;                   /* [numTempGroup*numNHChains] */
    const double *tempGroupNkbT;             /* [numTempGroup] */
    int32_t numParticlesLD;                  /* integrator.getParticlesLD().size() */
    int32_t numNormalLD; const int32_t *normalParticlesLD;
    int32_t numPairsLD; const int32_t *pairParticlesLD;
    int32_t numImagePairs; const int32_t *imagePairs;
    int32_t numElectrolyte; const int32_t *particlesElectrolyte;
    double invMassTotal;
} vvref_indices;

typedef struct {
    double temperature, frequency, drudeTemperature, drudeFrequency, stepSize;
    int32_t numNHChains, loopsPerStep;
    int32_t useCOMTempGroup, useMiddleScheme;
    double maxDrudeDistance, friction, drudeFriction;
    double mirrorLocation, electricField, cosAcceleration;
} vvref_params;

typedef struct {
    void *posq, *posqCorrection, *velm;
    long long *force;
    void *posDelta;
    const float *random;
} vvref_buffers;

} /* extern "C" */

template <class T>
struct Arr {
    T *p = nullptr;
    size_t n = 0;
    void alloc(size_t count) {
        n = std::max<size_t>(count, 1);
#if VVREF_GPU
        cudaMalloc((void **) &p, n * sizeof(T));
        cudaMemset(p, 0, n * sizeof(T));
#else
        p = (T *) calloc(n, sizeof(T));
#endif
    }
    void upload(const T *src, size_t count) {
        if (count == 0) return;
#if VVREF_GPU
        cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
#else
        memcpy(p, src, count * sizeof(T));
#endif
    }
    void release() {
#if VVREF_GPU
        if (p) cudaFree(p);
#else
        free(p);
#endif
        p = nullptr;
    }
};

struct vvref_ctx {
    vvref_params par;
    int numAtoms, paddedNumAtoms, numMolecules, numThreadBlocks;
    int nDrude, nNH, nMolNH, nNormalNH, nPairsNH, numTG, nLDall, nNormalLD, nPairsLD, nImg, nEl;
    double invMassTotal;
    Arr<int2> drudePairs, pairParticlesNH, particlesInMolecules, pairParticlesLD, imagePairs;
    Arr<int> particlesNH, moleculesNH, normalParticlesNH, particleMolId, particlesSortedByMolId;
    Arr<int> normalParticlesLD, particlesElectrolyte;
    Arr<real3> forceExtra;
    Arr<mixed4> oldDelta, comVelm;
    Arr<mixed> kineticEnergyBufferNH, kineticEnergiesNH, vscaleFactorsNH, vMaxBuffer;
    Arr<mixed2> stepSize;
    double etaMass[3][VVREF_MAX_CHAINS], eta[3][VVREF_MAX_CHAINS], etaDot[3][VVREF_MAX_CHAINS + 1],
        etaDotDot[3][VVREF_MAX_CHAINS], NkbT[3];
    double ke2[3], vscale[3];
    long launches;
    /* constraint stand-in tables (device memory on the GPU flavour) */
    vvc_constraints cons;
    Arr<int> consOffset, consAtoms;
    Arr<double> consDistance;
#if VVREF_GPU
    cudaStream_t stream;
#endif
};

/* ---- CudaContext::executeKernel geometry [OMM-mem]: block 64 unless given, grid =
 *      min(ceil(workUnits/block), numThreadBlocks) ------------------------------------------ */
#if VVREF_GPU
#define VVREF_LAUNCH(c, kernel, workUnits, blockSize, shmem, ...)                              \
    do {                                                                                       \
        int bs_ = (blockSize);                                                                 \
        int grid_ = std::min(((workUnits) + bs_ - 1) / bs_, (c)->numThreadBlocks);             \
        if (grid_ > 0) {                                                                       \
            kernel<<<grid_, bs_, (shmem), (c)->stream>>>(__VA_ARGS__);                         \
            (c)->launches++;                                                                   \
        }                                                                                      \
    } while (0)
#else
#define VVREF_LAUNCH(c, kernel, workUnits, blockSize, shmem, ...)                              \
    do {                                                                                       \
        int bs_ = (blockSize);                                                                 \
        int grid_ = std::min(((workUnits) + bs_ - 1) / bs_, (c)->numThreadBlocks);             \
        if ((shmem) > 0) { bs_ = 1; grid_ = 1; } /* single-block reductions: see header */     \
        if (grid_ > 0) {                                                                       \
            (c)->launches++;                                                                   \
            _Pragma("omp parallel for schedule(static)")                                       \
            for (int b_ = 0; b_ < grid_; b_++) {                                               \
                blockDim.x = bs_; blockDim.y = blockDim.z = 1;                                 \
                gridDim.x = grid_; gridDim.y = gridDim.z = 1;                                  \
                blockIdx.x = b_; blockIdx.y = blockIdx.z = 0;                                  \
                threadIdx.y = threadIdx.z = 0;                                                 \
                for (int t_ = 0; t_ < bs_; t_++) {                                             \
                    threadIdx.x = t_;                                                          \
                    kernel(__VA_ARGS__);                                                       \
                }                                                                              \
            }                                                                                  \
        }                                                                                      \
    } while (0)
#endif

static void bindConstants(const vvref_ctx *c) {
#if VVREF_GPU
#define SETC(sym, val) do { int v_ = (val); cudaMemcpyToSymbolAsync(sym, &v_, sizeof(int), 0, cudaMemcpyHostToDevice, c->stream); } while (0)
#else
#define SETC(sym, val) sym = (val)
#endif
    SETC(vr_NUM_ATOMS, c->numAtoms);
    SETC(vr_PADDED_NUM_ATOMS, c->paddedNumAtoms);
    SETC(vr_NUM_DRUDE_PAIRS, c->nDrude);
    SETC(vr_NUM_PARTICLES_NH, c->nNH);
    SETC(vr_NUM_MOLECULES_NH, c->nMolNH);
    SETC(vr_NUM_NORMAL_PARTICLES_NH, c->nNormalNH);
    SETC(vr_NUM_PAIRS_NH, c->nPairsNH);
    SETC(vr_NUM_NORMAL_PARTICLES_LD, c->nNormalLD);
    SETC(vr_NUM_PAIRS_LD, c->nPairsLD);
    SETC(vr_NUM_IMAGES, c->nImg);
    SETC(vr_NUM_PARTICLES_ELECTROLYTE, c->nEl);
#undef SETC
}

/* VVIntegrator::propagateNHChain restated -- VVIntegrator.cpp:340-376 */
static void propagateNHChain(const vvref_params &par, double *eta, double *eta_dot, double *eta_dotdot,
                             const double *eta_mass, double ke2, double ke2_target, double t_target,
                             double &factor) {
    double expfac = 0;
    double dt2 = par.stepSize / par.loopsPerStep / 2;
    double dt4 = dt2 / 2;
    double dt8 = dt4 / 2;
    factor = 1.0;
    eta_dotdot[0] = (ke2 - ke2_target) / eta_mass[0];
    for (int iloop = 0; iloop < par.loopsPerStep; iloop++) {
        for (int ich = par.numNHChains - 1; ich >= 0; ich--) {
            expfac = exp(-dt8 * eta_dot[ich + 1]);
            eta_dot[ich] *= expfac;
            eta_dot[ich] += eta_dotdot[ich] * dt4;
            eta_dot[ich] *= expfac;
        }
        factor *= exp(-dt2 * eta_dot[0]);
        for (int ich = 0; ich < par.numNHChains; ich++)
            eta[ich] += dt2 * eta_dot[ich];
        eta_dotdot[0] = (ke2 * factor * factor - ke2_target) / eta_mass[0];
        eta_dot[0] *= expfac;
        eta_dot[0] += eta_dotdot[0] * dt4;
        eta_dot[0] *= expfac;
        for (int ich = 1; ich < par.numNHChains; ich++) {
            expfac = exp(-dt8 * eta_dot[ich + 1]);
            eta_dot[ich] *= expfac;
            eta_dotdot[ich] = (eta_mass[ich - 1] * eta_dot[ich - 1] * eta_dot[ich - 1]
                               - VVREF_BOLTZ * t_target) / eta_mass[ich];
            eta_dot[ich] += eta_dotdot[ich] * dt4;
            eta_dot[ich] *= expfac;
        }
    }
}

#define POSQ(b) ((real4 *) (b)->posq)
#define CORR(b) ((real4 *) (b)->posqCorrection)
#define VELM(b) ((mixed4 *) (b)->velm)
#define PDELTA(b) ((mixed4 *) (b)->posDelta)

/* ---- host methods restated ---------------------------------------------------------------- */

/* CudaIntegrate{Middle,VV}StepKernel::resetExtraForce -- CudaVVKernels.cpp:119-127, 384-393 */
static void resetExtra(vvref_ctx *c) {
    VVREF_LAUNCH(c, resetExtraForce, c->numAtoms, 64, 0, c->forceExtra.p);
}

/* CudaModifyDrudeLangevinKernel::applyLangevinForce -- CudaVVKernels.cpp:826-872 */
static void applyLangevin(vvref_ctx *c, const vvref_buffers *b, unsigned *randomPos) {
    double stepSize = c->par.stepSize;
    double dragFactor = c->par.friction;
    double randFactor = sqrt(2.0 * VVREF_BOLTZ * c->par.temperature * dragFactor / stepSize);
    double dragFactorDrude = c->par.drudeFriction;
    double randFactorDrude = sqrt(2.0 * VVREF_BOLTZ * c->par.drudeTemperature * dragFactorDrude / stepSize);
    /* prepareRandomNumbers(normalParticlesLD->getSize() + 2*pairParticlesLD->getSize()): padded sizes */
    unsigned randomIndex = *randomPos;
    *randomPos += (unsigned) (std::max(c->nNormalLD, 1) + 2 * std::max(c->nPairsLD, 1));
    VVREF_LAUNCH(c, addExtraForceDrudeLangevin, c->nLDall, 64, 0,
                 VELM(b), c->forceExtra.p, c->normalParticlesLD.p, c->pairParticlesLD.p,
                 (mixed) dragFactor, (mixed) randFactor, (mixed) dragFactorDrude, (mixed) randFactorDrude,
                 (const float4 *) b->random, randomIndex);
}

/* CudaModifyElectricFieldKernel::applyElectricForce -- CudaVVKernels.cpp:971-992 */
static void applyField(vvref_ctx *c, const vvref_buffers *b) {
    double efscale = c->par.electricField * VVREF_AVOGADRO;
    VVREF_LAUNCH(c, addExtraForceElectricField, std::max(c->nEl, 1), 64, 0,
                 POSQ(b), c->forceExtra.p, c->particlesElectrolyte.p, (real) efscale);
}

static real4 invBox(double invBoxZ) {
    return make_real4(0, 0, (real) invBoxZ, 0);
}

/* CudaModifyCosineAccelerateKernel -- CudaVVKernels.cpp:1037-1110 */
static void applyCosine(vvref_ctx *c, const vvref_buffers *b, double invBoxZ) {
    VVREF_LAUNCH(c, addCosAcceleration, c->numAtoms, 64, 0, POSQ(b), VELM(b), c->forceExtra.p,
                 (real) c->par.cosAcceleration, invBox(invBoxZ));
}
static void calcVelocityBias(vvref_ctx *c, const vvref_buffers *b, double invBoxZ) {
    VVREF_LAUNCH(c, calcPeriodicVelocityBias, c->numAtoms, 64, 0, POSQ(b), VELM(b), c->vMaxBuffer.p, invBox(invBoxZ));
    int bufferSize = c->numAtoms;
    VVREF_LAUNCH(c, sumV, 512, 512, 512 * sizeof(mixed), c->vMaxBuffer.p, c->invMassTotal, bufferSize);
}
static void removeBias(vvref_ctx *c, const vvref_buffers *b, double invBoxZ) {
    VVREF_LAUNCH(c, removePeriodicVelocityBias, c->numAtoms, 64, 0, POSQ(b), VELM(b), c->vMaxBuffer.p, invBox(invBoxZ));
}
static void restoreBias(vvref_ctx *c, const vvref_buffers *b, double invBoxZ) {
    VVREF_LAUNCH(c, restorePeriodicVelocityBias, c->numAtoms, 64, 0, POSQ(b), VELM(b), c->vMaxBuffer.p, invBox(invBoxZ));
}

/* CudaModifyDrudeNoseKernel::scaleVelocity -- CudaVVKernels.cpp:670-754 */
#define NH_DISPATCH(c, name, workUnits, bs, shmem, ...)                                       \
    do {                                                                                      \
        if ((c)->numTG == 1) VVREF_LAUNCH(c, name##_tg1, workUnits, bs, shmem, __VA_ARGS__);  \
        else if ((c)->numTG == 2) VVREF_LAUNCH(c, name##_tg2, workUnits, bs, shmem, __VA_ARGS__); \
        else VVREF_LAUNCH(c, name##_tg3, workUnits, bs, shmem, __VA_ARGS__);                  \
    } while (0)

static void scaleVelocityHost(vvref_ctx *c, const vvref_buffers *b) {
    if (c->par.useCOMTempGroup) {
        NH_DISPATCH(c, calcCOMVelocities, c->nMolNH, 64, 0, VELM(b), c->comVelm.p, c->particlesInMolecules.p,
                    c->particlesSortedByMolId.p, c->moleculesNH.p);
        NH_DISPATCH(c, normalizeVelocities, c->nNH, 64, 0, VELM(b), c->comVelm.p, c->particleMolId.p, c->particlesNH.p);
    }
    int bufferSize = c->nNH * c->numTG;
    NH_DISPATCH(c, computeNormalizedKineticEnergies, c->nNH, 64, 0, VELM(b), c->comVelm.p, c->normalParticlesNH.p,
                c->pairParticlesNH.p, c->kineticEnergyBufferNH.p, c->moleculesNH.p, bufferSize);
    NH_DISPATCH(c, sumNormalizedKineticEnergies, 512, 512, 512 * c->numTG * sizeof(mixed),
                c->kineticEnergyBufferNH.p, c->kineticEnergiesNH.p, bufferSize);
    mixed keHost[3] = {0, 0, 0};
#if VVREF_GPU
    /* the reference's blocking download, :709-716 */
    cudaMemcpyAsync(keHost, c->kineticEnergiesNH.p, c->numTG * sizeof(mixed), cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
#else
    memcpy(keHost, c->kineticEnergiesNH.p, c->numTG * sizeof(mixed));
#endif
    for (int g = 0; g < 3; g++) {
        c->ke2[g] = g < c->numTG ? (double) keHost[g] : 0.0;
        c->vscale[g] = 1.0;
    }
    for (int itg = 0; itg < c->numTG; itg++) {
        const double T = itg == TG_DRUDE ? c->par.drudeTemperature : c->par.temperature;
        if (c->etaMass[itg][0] > 0)
            propagateNHChain(c->par, c->eta[itg], c->etaDot[itg], c->etaDotDot[itg], c->etaMass[itg],
                             c->ke2[itg], c->NkbT[itg], T, c->vscale[itg]);
    }
    /* upload of numTempGroup factors, :741-746.  The kernel reads three;
     * the array here always has room for three and the tail holds 1. */
    mixed vs[3] = {(mixed) c->vscale[0], (mixed) c->vscale[1], (mixed) c->vscale[2]};
#if VVREF_GPU
    cudaMemcpyAsync(c->vscaleFactorsNH.p, vs, 3 * sizeof(mixed), cudaMemcpyHostToDevice, c->stream);
#else
    memcpy(c->vscaleFactorsNH.p, vs, 3 * sizeof(mixed));
#endif
    NH_DISPATCH(c, scaleVelocity, c->nNH, 64, 0, VELM(b), c->comVelm.p, c->particleMolId.p, c->normalParticlesNH.p,
                c->pairParticlesNH.p, c->vscaleFactorsNH.p);
}

/* CudaModifyImageChargeKernel::updateImagePositions -- CudaVVKernels.cpp:904-934 */
static void updateImages(vvref_ctx *c, const vvref_buffers *b) {
    VVREF_LAUNCH(c, updateImagePositions, c->nImg, 64, 0, POSQ(b), CORR(b), c->imagePairs.p,
                 (mixed) c->par.mirrorLocation);
}

static void hardWall(vvref_ctx *c, const vvref_buffers *b, bool middle) {
    double maxDrudeDistance = c->par.maxDrudeDistance;
    double hardwallScaleDrude = sqrt(VVREF_BOLTZ * c->par.drudeTemperature);
    if (maxDrudeDistance > 0 && c->nDrude > 0) {
        if (middle)
            VVREF_LAUNCH(c, applyHardWallConstraints, std::max(c->nDrude, 1), 64, 0, POSQ(b), CORR(b), VELM(b),
                         c->drudePairs.p, c->stepSize.p, (mixed) maxDrudeDistance, (mixed) hardwallScaleDrude);
        else
            VVREF_LAUNCH(c, applyHardWallConstraints_vv, std::max(c->nDrude, 1), 64, 0, POSQ(b), CORR(b), VELM(b),
                         c->drudePairs.p, c->stepSize.p, (mixed) maxDrudeDistance, (mixed) hardwallScaleDrude);
    }
}

/* integration.applyConstraints(tol) / applyVelocityConstraints(tol): the stand-in, see above */
static void applyConstraintsStandin(vvref_ctx *c, const vvref_buffers *b) {
    if (c->cons.numClusters <= 0) return;
#if VVREF_GPU
    standinPositionsKernel<<<(c->cons.numClusters + 127) / 128, 128, 0, c->stream>>>(c->cons, POSQ(b), CORR(b), VELM(b), PDELTA(b));
#else
    _Pragma("omp parallel for schedule(static)")
    for (int k = 0; k < c->cons.numClusters; k++)
        vvc_cluster_positions(c->cons, k, POSQ(b), CORR(b), VELM(b), PDELTA(b));
#endif
}
static void applyVelocityConstraintsStandin(vvref_ctx *c, const vvref_buffers *b) {
    if (c->cons.numClusters <= 0) return;
#if VVREF_GPU
    standinVelocitiesKernel<<<(c->cons.numClusters + 127) / 128, 128, 0, c->stream>>>(c->cons, POSQ(b), CORR(b), VELM(b));
#else
    _Pragma("omp parallel for schedule(static)")
    for (int k = 0; k < c->cons.numClusters; k++)
        vvc_cluster_velocities(c->cons, k, POSQ(b), CORR(b), VELM(b));
#endif
}

static void nhHalf(vvref_ctx *c, const vvref_buffers *b, double invBoxZ) {
    if (c->nNH > 0) {
        if (c->par.cosAcceleration != 0) {
            calcVelocityBias(c, b, invBoxZ);
            removeBias(c, b, invBoxZ);
        }
        scaleVelocityHost(c, b);
        if (c->par.cosAcceleration != 0)
            restoreBias(c, b, invBoxZ);
    }
}

static void extraForces(vvref_ctx *c, const vvref_buffers *b, double invBoxZ, unsigned *randomPos) {
    if (c->nLDall > 0 || c->nEl > 0 || c->par.cosAcceleration != 0)
        resetExtra(c);
    if (c->nLDall > 0)
        applyLangevin(c, b, randomPos);
    if (c->nEl > 0)
        applyField(c, b);
    if (c->par.cosAcceleration != 0)
        applyCosine(c, b, invBoxZ);
}

extern "C" {

int vvref_is_gpu(void) { return VVREF_GPU; }
int vvref_precision_mode(void) { return VVREF_MODE; }
int vvref_max_threads(void) {
#if !VVREF_GPU && defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vvref_set_num_threads(int n) {
#if !VVREF_GPU && defined(_OPENMP)
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());   /* n <= 0: all host cores */
#else
    (void) n;
#endif
}

vvref_ctx *vvref_create(const vvref_indices *ix, const vvref_params *par, int numThreadBlocks, void *stream) {
    vvref_ctx *c = new vvref_ctx();
    c->par = *par;
    c->numAtoms = ix->numAtoms;
    c->paddedNumAtoms = ix->paddedNumAtoms;
    c->numMolecules = ix->numMolecules;
    c->numThreadBlocks = numThreadBlocks;
    c->nDrude = ix->numDrude; c->nNH = ix->numParticlesNH; c->nMolNH = ix->numMoleculesNH;
    c->nNormalNH = ix->numNormalNH; c->nPairsNH = ix->numPairsNH; c->numTG = ix->numTempGroup;
    c->nLDall = ix->numParticlesLD; c->nNormalLD = ix->numNormalLD; c->nPairsLD = ix->numPairsLD;
    c->nImg = ix->numImagePairs; c->nEl = ix->numElectrolyte;
    c->invMassTotal = ix->invMassTotal;
    c->launches = 0;
#if VVREF_GPU
    c->stream = (cudaStream_t) stream;
#else
    (void) stream;
#endif
    c->drudePairs.alloc(c->nDrude); c->drudePairs.upload((const int2 *) ix->drudePairs, c->nDrude);
    c->particlesNH.alloc(c->nNH); c->particlesNH.upload(ix->particlesNH, c->nNH);
    c->moleculesNH.alloc(c->nMolNH); c->moleculesNH.upload(ix->moleculesNH, c->nMolNH);
    c->normalParticlesNH.alloc(c->nNormalNH); c->normalParticlesNH.upload(ix->normalParticlesNH, c->nNormalNH);
    c->pairParticlesNH.alloc(c->nPairsNH); c->pairParticlesNH.upload((const int2 *) ix->pairParticlesNH, c->nPairsNH);
    c->particleMolId.alloc(c->numAtoms); c->particleMolId.upload(ix->particleMolId, c->numAtoms);
    c->particlesInMolecules.alloc(c->numMolecules);
    c->particlesInMolecules.upload((const int2 *) ix->particlesInMolecules, c->numMolecules);
    c->particlesSortedByMolId.alloc(c->numAtoms); c->particlesSortedByMolId.upload(ix->particlesSortedByMolId, c->numAtoms);
    c->normalParticlesLD.alloc(c->nNormalLD); c->normalParticlesLD.upload(ix->normalParticlesLD, c->nNormalLD);
    c->pairParticlesLD.alloc(c->nPairsLD); c->pairParticlesLD.upload((const int2 *) ix->pairParticlesLD, c->nPairsLD);
    c->imagePairs.alloc(c->nImg); c->imagePairs.upload((const int2 *) ix->imagePairs, c->nImg);
    c->particlesElectrolyte.alloc(c->nEl); c->particlesElectrolyte.upload(ix->particlesElectrolyte, c->nEl);
    c->forceExtra.alloc(c->numAtoms);   /* zero-initialised, CudaVVKernels.cpp:79-89 */
    c->oldDelta.alloc(c->numAtoms);
    c->comVelm.alloc(c->numMolecules);  /* zero-initialised, :606-617 */
    /* zeroed once: SURVEY Appendix C-2 (the reference relies on cuMemAlloc returning zeros) */
    c->kineticEnergyBufferNH.alloc((size_t) c->nNH * c->numTG);
    c->kineticEnergiesNH.alloc(3);
    c->vscaleFactorsNH.alloc(3);
    c->vMaxBuffer.alloc(c->numAtoms);
    c->stepSize.alloc(1);
    mixed2 ss = make_mixed2(0, (mixed) par->stepSize);   /* :309-318 */
    c->stepSize.upload(&ss, 1);
    memset(c->eta, 0, sizeof c->eta); memset(c->etaDot, 0, sizeof c->etaDot);
    memset(c->etaDotDot, 0, sizeof c->etaDotDot); memset(c->etaMass, 0, sizeof c->etaMass);
    for (int g = 0; g < c->numTG; g++) {
        c->NkbT[g] = ix->tempGroupNkbT[g];
        for (int k = 0; k < par->numNHChains; k++)
            c->etaMass[g][k] = ix->etaMass[g * par->numNHChains + k];
    }
    for (int g = 0; g < 3; g++) { c->ke2[g] = 0; c->vscale[g] = 1; }
    memset(&c->cons, 0, sizeof c->cons);
    return c;
}

/* cluster tables of the constraint stand-in (host pointers; copied).  numClusters == 0: off. */
void vvref_set_constraint_standin(vvref_ctx *c, int numClusters, const int32_t *clusterOffset, const int32_t *atoms,
                                  const double *distance, int iterations) {
    c->consOffset.release(); c->consAtoms.release(); c->consDistance.release();
    memset(&c->cons, 0, sizeof c->cons);
    if (numClusters <= 0) return;
    const int nCons = clusterOffset[numClusters];
    c->consOffset.alloc(numClusters + 1); c->consOffset.upload(clusterOffset, numClusters + 1);
    c->consAtoms.alloc(2 * (size_t) nCons); c->consAtoms.upload(atoms, 2 * (size_t) nCons);
    c->consDistance.alloc(nCons); c->consDistance.upload(distance, nCons);
    c->cons.numClusters = numClusters;
    c->cons.iterations = iterations;
    c->cons.clusterOffset = c->consOffset.p;
    c->cons.atoms = c->consAtoms.p;
    c->cons.distance = c->consDistance.p;
}

void vvref_destroy(vvref_ctx *c) {
    if (!c) return;
    c->drudePairs.release(); c->pairParticlesNH.release(); c->particlesInMolecules.release();
    c->pairParticlesLD.release(); c->imagePairs.release(); c->particlesNH.release(); c->moleculesNH.release();
    c->normalParticlesNH.release(); c->particleMolId.release(); c->particlesSortedByMolId.release();
    c->normalParticlesLD.release(); c->particlesElectrolyte.release(); c->forceExtra.release();
    c->oldDelta.release(); c->comVelm.release(); c->kineticEnergyBufferNH.release();
    c->kineticEnergiesNH.release(); c->vscaleFactorsNH.release(); c->vMaxBuffer.release(); c->stepSize.release();
    c->consOffset.release(); c->consAtoms.release(); c->consDistance.release();
    delete c;
}

/* steps of VVIntegrator::stepMiddle (VVIntegrator.cpp:232-270) or stepVV (:272-338) with the
 * forces in b->force held fixed.  Returns the number of kernel launches issued. */
long vvref_step(vvref_ctx *c, const vvref_buffers *b, int steps, double invBoxZ, unsigned *randomPos) {
    bindConstants(c);
    long before = c->launches;
    for (int s = 0; s < steps; s++) {
        if (c->par.useMiddleScheme) {
            extraForces(c, b, invBoxZ, randomPos);
            /* firstIntegrate -- CudaVVKernels.cpp:129-159 */
            VVREF_LAUNCH(c, integrateMiddleVel, c->numAtoms, 64, 0, VELM(b), b->force, c->forceExtra.p, c->stepSize.p);
            applyVelocityConstraintsStandin(c, b);      /* integration.applyVelocityConstraints, :151 */
            VVREF_LAUNCH(c, integrateMiddlePos1, c->numAtoms, 64, 0, VELM(b), PDELTA(b), c->oldDelta.p, c->stepSize.p);
            nhHalf(c, b, invBoxZ);
            /* secondIntegrate -- :161-220 */
            VVREF_LAUNCH(c, integrateMiddlePos2, c->numAtoms, 64, 0, VELM(b), PDELTA(b), c->oldDelta.p, c->stepSize.p);
            applyConstraintsStandin(c, b);              /* integration.applyConstraints, :176 */
            VVREF_LAUNCH(c, integrateMiddlePos3, c->numAtoms, 64, 0, POSQ(b), CORR(b), PDELTA(b), c->oldDelta.p,
                         VELM(b), c->stepSize.p);
            hardWall(c, b, true);
            if (c->nImg > 0)
                updateImages(c, b);
        } else {
            nhHalf(c, b, invBoxZ);
            /* firstIntegrate -- :296-382 */
            double fscale = 0.5 * c->par.stepSize / (double) 0x100000000;
            VVREF_LAUNCH(c, velocityVerletIntegrateVelocities, c->numAtoms, 64, 0, VELM(b), b->force, c->forceExtra.p,
                         PDELTA(b), c->stepSize.p, (mixed) fscale, true);
            applyConstraintsStandin(c, b);              /* integration.applyConstraints, :351 */
            VVREF_LAUNCH(c, velocityVerletIntegratePositions, c->numAtoms, 64, 0, POSQ(b), CORR(b), PDELTA(b), VELM(b),
                         c->stepSize.p);
            hardWall(c, b, false);
            if (c->nImg > 0)
                updateImages(c, b);
            extraForces(c, b, invBoxZ, randomPos);
            /* secondIntegrate -- :395-431 */
            VVREF_LAUNCH(c, velocityVerletIntegrateVelocities, c->numAtoms, 64, 0, VELM(b), b->force, c->forceExtra.p,
                         PDELTA(b), c->stepSize.p, (mixed) fscale, false);
            applyVelocityConstraintsStandin(c, b);      /* integration.applyVelocityConstraints, :427 */
            nhHalf(c, b, invBoxZ);
        }
    }
    return c->launches - before;
}

/* one thermostat application only (CudaModifyDrudeNoseKernel::scaleVelocity) */
void vvref_scale_velocity(vvref_ctx *c, const vvref_buffers *b) {
    bindConstants(c);
    scaleVelocityHost(c, b);
}

void vvref_get_state(vvref_ctx *c, double *ke2, double *vscale, double *eta, double *etaDot, double *etaDotDot,
                     double *vBias) {
    const int nc = c->par.numNHChains;
    for (int g = 0; g < 3; g++) { ke2[g] = c->ke2[g]; vscale[g] = c->vscale[g]; }
    for (int g = 0; g < c->numTG; g++) {
        for (int k = 0; k < nc; k++) { eta[g * nc + k] = c->eta[g][k]; etaDotDot[g * nc + k] = c->etaDotDot[g][k]; }
        for (int k = 0; k < nc + 1; k++) etaDot[g * (nc + 1) + k] = c->etaDot[g][k];
    }
    mixed v = 0;
#if VVREF_GPU
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(&v, c->vMaxBuffer.p, sizeof(mixed), cudaMemcpyDeviceToHost);
#else
    v = c->vMaxBuffer.p[0];
#endif
    *vBias = (double) v;
}

/* copy of comVelm (mixed4 per molecule) for inspection */
void vvref_get_com_velm(vvref_ctx *c, void *out) {
#if VVREF_GPU
    cudaStreamSynchronize(c->stream);
    cudaMemcpy(out, c->comVelm.p, c->numMolecules * sizeof(mixed4), cudaMemcpyDeviceToHost);
#else
    memcpy(out, c->comVelm.p, c->numMolecules * sizeof(mixed4));
#endif
}

int vvref_last_cuda_error(void) {
#if VVREF_GPU
    return (int) cudaGetLastError();
#else
    return 0;
#endif
}

} /* extern "C" */
