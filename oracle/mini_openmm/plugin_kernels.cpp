/*
 * plugin_kernels.cpp -- TEST INFRASTRUCTURE ONLY: the kernel side of the mini-OpenMM (mini_openmm.h).
 *
 * OpenMM JIT-compiles the plugin's kernel sources with NVRTC at Context creation; here the same sources -- the
 * reference's platforms/cuda/src/kernels/[name].cu, #included from where they lie by oracle/ref_kernels.inc.h -- are
 * compiled AHEAD OF TIME (g++ for the host flavour, nvcc sm_100a for -DMINIOMM_CUDA) and looked up by (module, name)
 * when the reference's host code calls cu.getKernel().  Also: device memory, copies, and the constraint stand-in on this
 * flavour's memory.  Compile as C++ (host) or with `nvcc -x cu` (CUDA).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../ref_kernels.inc.h"
#include "miniomm_backend.h"

#if defined(MINIOMM_CUDA) != VVREF_GPU
#error "MINIOMM_CUDA must be set exactly when this file is compiled by nvcc"
#endif

#define VVC_REAL4 real4
#define VVC_MIXED4 mixed4
#define VVC_MIXED mixed
#if VVREF_GPU
#define VVC_FN __host__ __device__ inline
#else
#define VVC_FN static inline
#endif
#include "../constraint_standin.h"

namespace miniomm {

struct KernelEntry {
    std::string module, name;
    int numTG;                                         /* 0: any */
    const void *fn;
    void (*hostInvoke)(const void *fn, void **args);   /* host flavour: unpack `args` and call */
};
static std::vector<KernelEntry> &registry() { static std::vector<KernelEntry> r; return r; }

template <class... A> struct Invoker {
    template <size_t... I> static void impl(void (*f)(A...), void **args, std::index_sequence<I...>) {
        f(*reinterpret_cast<typename std::remove_cv<typename std::remove_reference<A>::type>::type *>(args[I])...);
    }
    static void call(const void *fn, void **args) { impl((void (*)(A...)) fn, args, std::index_sequence_for<A...>{}); }
};
template <class... A> static void reg(const char *module, const char *name, int numTG, void (*fn)(A...)) {
    KernelEntry e;
    e.module = module; e.name = name; e.numTG = numTG; e.fn = (const void *) fn;
    e.hostInvoke = &Invoker<A...>::call;
    registry().push_back(e);
}

static void registerAll() {
    if (!registry().empty())
        return;
    reg("middle", "integrateMiddleVel", 0, integrateMiddleVel);
    reg("middle", "integrateMiddlePos1", 0, integrateMiddlePos1);
    reg("middle", "integrateMiddlePos2", 0, integrateMiddlePos2);
    reg("middle", "integrateMiddlePos3", 0, integrateMiddlePos3);
    reg("middle", "applyHardWallConstraints", 0, applyHardWallConstraints);
    reg("middle", "resetExtraForce", 0, resetExtraForce);
    reg("velocityVerlet", "velocityVerletIntegrateVelocities", 0, velocityVerletIntegrateVelocities);
    reg("velocityVerlet", "velocityVerletIntegratePositions", 0, velocityVerletIntegratePositions);
    reg("velocityVerlet", "applyHardWallConstraints", 0, applyHardWallConstraints_vv);
    reg("velocityVerlet", "resetExtraForce", 0, resetExtraForce_vv);
    reg("drudeLangevin", "addExtraForceDrudeLangevin", 0, addExtraForceDrudeLangevin);
    reg("cosineAccelerate", "addCosAcceleration", 0, addCosAcceleration);
    reg("cosineAccelerate", "calcPeriodicVelocityBias", 0, calcPeriodicVelocityBias);
    reg("cosineAccelerate", "sumV", 0, sumV);
    reg("cosineAccelerate", "removePeriodicVelocityBias", 0, removePeriodicVelocityBias);
    reg("cosineAccelerate", "restorePeriodicVelocityBias", 0, restorePeriodicVelocityBias);
    reg("electricField", "addExtraForceElectricField", 0, addExtraForceElectricField);
    reg("imageCharge", "updateImagePositions", 0, updateImagePositions);
#define REG_NH(n)                                                                                         \
    reg("drudeNoseHoover", "calcCOMVelocities", n, calcCOMVelocities_tg##n);                              \
    reg("drudeNoseHoover", "normalizeVelocities", n, normalizeVelocities_tg##n);                          \
    reg("drudeNoseHoover", "computeNormalizedKineticEnergies", n, computeNormalizedKineticEnergies_tg##n); \
    reg("drudeNoseHoover", "sumNormalizedKineticEnergies", n, sumNormalizedKineticEnergies_tg##n);        \
    reg("drudeNoseHoover", "scaleVelocity", n, scaleVelocity_tg##n);
    REG_NH(1) REG_NH(2) REG_NH(3)
#undef REG_NH
}

KernelEntry *findKernel(const std::string &module, const std::string &name, int numTG) {
    registerAll();
    for (size_t i = 0; i < registry().size(); i++) {
        KernelEntry &e = registry()[i];
        if (e.name == name && e.module == module && (e.numTG == 0 || e.numTG == numTG))
            return &e;
    }
    return nullptr;
}

void setDefine(const std::string &name, long value, void *stream) {
    const int v = (int) value;
    (void) stream;
#if VVREF_GPU
#define SETC(sym) cudaMemcpyToSymbolAsync(sym, &v, sizeof(int), 0, cudaMemcpyHostToDevice, (cudaStream_t) stream)
#else
#define SETC(sym) sym = v
#endif
    if (name == "NUM_ATOMS") SETC(vr_NUM_ATOMS);
    else if (name == "PADDED_NUM_ATOMS") SETC(vr_PADDED_NUM_ATOMS);
    else if (name == "NUM_DRUDE_PAIRS") SETC(vr_NUM_DRUDE_PAIRS);
    else if (name == "NUM_PARTICLES_NH") SETC(vr_NUM_PARTICLES_NH);
    else if (name == "NUM_MOLECULES_NH") SETC(vr_NUM_MOLECULES_NH);
    else if (name == "NUM_NORMAL_PARTICLES_NH") SETC(vr_NUM_NORMAL_PARTICLES_NH);
    else if (name == "NUM_PAIRS_NH") SETC(vr_NUM_PAIRS_NH);
    else if (name == "NUM_NORMAL_PARTICLES_LD") SETC(vr_NUM_NORMAL_PARTICLES_LD);
    else if (name == "NUM_PAIRS_LD") SETC(vr_NUM_PAIRS_LD);
    else if (name == "NUM_IMAGES") SETC(vr_NUM_IMAGES);
    else if (name == "NUM_PARTICLES_ELECTROLYTE") SETC(vr_NUM_PARTICLES_ELECTROLYTE);
    /* TG_ATOM / TG_COM / TG_DRUDE are fixed (0, 1, 2) in ref_kernels.inc.h like in CudaVVKernels.cpp:49 */
#undef SETC
}

void launchKernel(KernelEntry *k, void **args, int grid, int block, unsigned sharedSize, void *stream) {
#if VVREF_GPU
    cudaLaunchKernel(k->fn, dim3(grid), dim3(block), args, sharedSize, (cudaStream_t) stream);
#else
    (void) stream;
    if (sharedSize > 0) { block = 1; grid = 1; }      /* the two single-block tree reductions: see ref_harness.cpp header */
    _Pragma("omp parallel for schedule(static)")
    for (int b = 0; b < grid; b++) {
        blockDim.x = block; blockDim.y = blockDim.z = 1;
        gridDim.x = grid; gridDim.y = gridDim.z = 1;
        blockIdx.x = b; blockIdx.y = blockIdx.z = 0;
        threadIdx.y = threadIdx.z = 0;
        for (int t = 0; t < block; t++) {
            threadIdx.x = t;
            k->hostInvoke(k->fn, args);
        }
    }
#endif
}

void *deviceAlloc(size_t bytes) {
#if VVREF_GPU
    void *p = nullptr;
    cudaMalloc(&p, bytes);
    cudaMemset(p, 0, bytes);
    return p;
#else
    return calloc(bytes, 1);
#endif
}
void deviceFree(void *p) {
#if VVREF_GPU
    if (p) cudaFree(p);
#else
    free(p);
#endif
}
void copyToDevice(void *dst, const void *src, size_t bytes, void *stream) {
#if VVREF_GPU
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t) stream);
    cudaStreamSynchronize((cudaStream_t) stream);      /* `src` may be a temporary of the caller */
#else
    (void) stream;
    memcpy(dst, src, bytes);
#endif
}
void copyToHost(void *dst, const void *src, size_t bytes, void *stream) {
#if VVREF_GPU
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t) stream);
    cudaStreamSynchronize((cudaStream_t) stream);
#else
    (void) stream;
    memcpy(dst, src, bytes);
#endif
}
int numThreadBlocks() {
#if VVREF_GPU
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 4 * sms;
#else
    return 4 * 148;
#endif
}
bool isCuda() { return VVREF_GPU != 0; }

#if VVREF_GPU
__global__ void miniStandinPositions(vvc_constraints cs, const real4 *posq, const real4 *corr, const mixed4 *velm, mixed4 *posDelta) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)
        vvc_cluster_positions(cs, c, posq, corr, velm, posDelta);
}
__global__ void miniStandinVelocities(vvc_constraints cs, const real4 *posq, const real4 *corr, mixed4 *velm) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)
        vvc_cluster_velocities(cs, c, posq, corr, velm);
}
#endif

static vvc_constraints asConstraints(const Standin &s) {
    vvc_constraints cs;
    cs.numClusters = s.numClusters; cs.iterations = s.iterations;
    cs.clusterOffset = s.offset; cs.atoms = s.atoms; cs.distance = s.distance;
    return cs;
}

void standinPositions(const Standin &s, void *posq, void *corr, void *velm, void *posDelta, void *stream) {
    if (s.numClusters <= 0) return;
    const vvc_constraints cs = asConstraints(s);
#if VVREF_GPU
    miniStandinPositions<<<(s.numClusters + 127) / 128, 128, 0, (cudaStream_t) stream>>>(cs, (const real4 *) posq, (const real4 *) corr,
                                                                                       (const mixed4 *) velm, (mixed4 *) posDelta);
#else
    (void) stream;
    _Pragma("omp parallel for schedule(static)")
    for (int c = 0; c < cs.numClusters; c++)
        vvc_cluster_positions(cs, c, (const real4 *) posq, (const real4 *) corr, (const mixed4 *) velm, (mixed4 *) posDelta);
#endif
}

void standinVelocities(const Standin &s, void *posq, void *corr, void *velm, void *stream) {
    if (s.numClusters <= 0) return;
    const vvc_constraints cs = asConstraints(s);
#if VVREF_GPU
    miniStandinVelocities<<<(s.numClusters + 127) / 128, 128, 0, (cudaStream_t) stream>>>(cs, (const real4 *) posq, (const real4 *) corr,
                                                                                        (mixed4 *) velm);
#else
    (void) stream;
    _Pragma("omp parallel for schedule(static)")
    for (int c = 0; c < cs.numClusters; c++)
        vvc_cluster_velocities(cs, c, (const real4 *) posq, (const real4 *) corr, (mixed4 *) velm);
#endif
}

}   // namespace miniomm
