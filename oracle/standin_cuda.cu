/*
 * standin_cuda.cu -- TEST INFRASTRUCTURE ONLY.
 *
 * The constraint stand-in of oracle/constraint_standin.h on DEVICE buffers, for the GPU tests that drive the product's
 * split entry points the way the OpenMM glue does:
 *     vvb200_middle_kick | <applyVelocityConstraints> | vvb200_middle_thermostat_delta | <applyConstraints> | vvb200_middle_finish
 *     vvb200_vv_kick(posDelta) | <applyConstraints> | vvb200_vv_positions ... vvb200_vv_kick | <applyVelocityConstraints>
 * Built by `make -C oracle oracle` into oracle/build/libvvstandin_cuda.so (no reference source involved).  Launches go
 * to the stream passed in; nothing synchronises.
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

struct SF4 { float x, y, z, w; };
struct SD4 { double x, y, z, w; };

template <int MODE> struct Types;      // OpenMM's CudaPrecision modes: 0 single, 1 mixed, 2 double
template <> struct Types<0> { typedef SF4 real4; typedef SF4 mixed4; typedef float mixed; };
template <> struct Types<1> { typedef SF4 real4; typedef SD4 mixed4; typedef double mixed; };
template <> struct Types<2> { typedef SD4 real4; typedef SD4 mixed4; typedef double mixed; };

#define VVC_FN __host__ __device__ inline
namespace m0 {
#define VVC_REAL4 Types<0>::real4
#define VVC_MIXED4 Types<0>::mixed4
#define VVC_MIXED Types<0>::mixed
#include "constraint_standin.h"
#undef VVC_REAL4
#undef VVC_MIXED4
#undef VVC_MIXED
}
#undef VVC_CONSTRAINT_STANDIN_H_
namespace m1 {
#define VVC_REAL4 Types<1>::real4
#define VVC_MIXED4 Types<1>::mixed4
#define VVC_MIXED Types<1>::mixed
#include "constraint_standin.h"
#undef VVC_REAL4
#undef VVC_MIXED4
#undef VVC_MIXED
}
#undef VVC_CONSTRAINT_STANDIN_H_
namespace m2 {
#define VVC_REAL4 Types<2>::real4
#define VVC_MIXED4 Types<2>::mixed4
#define VVC_MIXED Types<2>::mixed
#include "constraint_standin.h"
#undef VVC_REAL4
#undef VVC_MIXED4
#undef VVC_MIXED
}

#define KERNELS(NS, M)                                                                                                          \
    __global__ void positions_##NS(NS::vvc_constraints cs, const Types<M>::real4 *posq, const Types<M>::real4 *corr,            \
                                   const Types<M>::mixed4 *velm, Types<M>::mixed4 *posDelta) {                                  \
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)                    \
            NS::vvc_cluster_positions(cs, c, posq, corr, velm, posDelta);                                                       \
    }                                                                                                                           \
    __global__ void velocities_##NS(NS::vvc_constraints cs, const Types<M>::real4 *posq, const Types<M>::real4 *corr,           \
                                    Types<M>::mixed4 *velm) {                                                                   \
        for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cs.numClusters; c += blockDim.x * gridDim.x)                    \
            NS::vvc_cluster_velocities(cs, c, posq, corr, velm);                                                                \
    }
KERNELS(m0, 0)
KERNELS(m1, 1)
KERNELS(m2, 2)

struct vvstandin {
    int precision;
    int numClusters, iterations;
    int32_t *offset, *atoms;
    double *distance;
};

extern "C" {

vvstandin *vvstandin_create(int precision, int numClusters, const int32_t *clusterOffset, const int32_t *atoms,
                            const double *distance, int iterations) {
    if (precision < 0 || precision > 2 || numClusters < 0)
        return nullptr;
    vvstandin *s = new vvstandin();
    s->precision = precision;
    s->numClusters = numClusters;
    s->iterations = iterations;
    const int nCons = numClusters > 0 ? clusterOffset[numClusters] : 0;
    cudaMalloc((void **) &s->offset, (numClusters + 1) * sizeof(int32_t));
    cudaMalloc((void **) &s->atoms, (2 * (size_t) nCons + 1) * sizeof(int32_t));
    cudaMalloc((void **) &s->distance, ((size_t) nCons + 1) * sizeof(double));
    if (numClusters > 0) {
        cudaMemcpy(s->offset, clusterOffset, (numClusters + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
        cudaMemcpy(s->atoms, atoms, 2 * (size_t) nCons * sizeof(int32_t), cudaMemcpyHostToDevice);
        cudaMemcpy(s->distance, distance, (size_t) nCons * sizeof(double), cudaMemcpyHostToDevice);
    }
    return s;
}

void vvstandin_destroy(vvstandin *s) {
    if (!s) return;
    cudaFree(s->offset); cudaFree(s->atoms); cudaFree(s->distance);
    delete s;
}

#define FILL(NS) NS::vvc_constraints cs; cs.numClusters = s->numClusters; cs.iterations = s->iterations; \
                 cs.clusterOffset = s->offset; cs.atoms = s->atoms; cs.distance = s->distance

/* integration.applyConstraints stand-in: posDelta rewritten */
int vvstandin_apply_constraints(vvstandin *s, const void *posq, const void *corr, const void *velm, void *posDelta, void *stream) {
    if (!s || s->numClusters == 0) return 0;
    const int grid = (s->numClusters + 127) / 128;
    cudaStream_t st = (cudaStream_t) stream;
    if (s->precision == 0) { FILL(m0); positions_m0<<<grid, 128, 0, st>>>(cs, (const SF4 *) posq, (const SF4 *) corr, (const SF4 *) velm, (SF4 *) posDelta); }
    else if (s->precision == 1) { FILL(m1); positions_m1<<<grid, 128, 0, st>>>(cs, (const SF4 *) posq, (const SF4 *) corr, (const SD4 *) velm, (SD4 *) posDelta); }
    else { FILL(m2); positions_m2<<<grid, 128, 0, st>>>(cs, (const SD4 *) posq, (const SD4 *) corr, (const SD4 *) velm, (SD4 *) posDelta); }
    return (int) cudaGetLastError();
}

/* integration.applyVelocityConstraints stand-in: velm rewritten */
int vvstandin_apply_velocity_constraints(vvstandin *s, const void *posq, const void *corr, void *velm, void *stream) {
    if (!s || s->numClusters == 0) return 0;
    const int grid = (s->numClusters + 127) / 128;
    cudaStream_t st = (cudaStream_t) stream;
    if (s->precision == 0) { FILL(m0); velocities_m0<<<grid, 128, 0, st>>>(cs, (const SF4 *) posq, (const SF4 *) corr, (SF4 *) velm); }
    else if (s->precision == 1) { FILL(m1); velocities_m1<<<grid, 128, 0, st>>>(cs, (const SF4 *) posq, (const SF4 *) corr, (SD4 *) velm); }
    else { FILL(m2); velocities_m2<<<grid, 128, 0, st>>>(cs, (const SD4 *) posq, (const SD4 *) corr, (SD4 *) velm); }
    return (int) cudaGetLastError();
}

}   // extern "C"
