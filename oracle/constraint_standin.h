/*
 * constraint_standin.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A deterministic stand-in for the two OpenMM calls that sit BETWEEN the plugin's sub-steps when
 * the System has constraints (both shipped example scripts use HBonds):
 *     CudaIntegrationUtilities::applyVelocityConstraints(tol)   CudaVVKernels.cpp:151, 427
 *     CudaIntegrationUtilities::applyConstraints(tol)           CudaVVKernels.cpp:176, 351
 * OpenMM is not in the image, so its SHAKE / SETTLE / CCMA kernels cannot run here.  What the
 * integration path needs from them is only their CONTRACT: applyConstraints rewrites posDelta (the
 * pending displacement) using posq (+posqCorrection) as the reference geometry and velm.w as the
 * inverse masses; applyVelocityConstraints rewrites velm.  This header implements that contract as
 * a fixed number of SHAKE / RATTLE sweeps over clusters of distance constraints, one cluster per
 * thread, constraints inside a cluster visited in list order, every product rounded on its own
 * (VVC_MUL: never contracted into an FMA, whatever the compiler flags of the including file) -- so
 * every consumer (the C oracle, the reference-kernel harness on host and GPU, the mini-OpenMM the
 * reference plugin and the glue are linked against, and the GPU tests that call the product's
 * split entry points) applies bit-for-bit the same operator to its own buffers.  It is NOT OpenMM's solver and no physical
 * claim is made; it exists so that posDelta != oldDelta and velocities change between sub-steps.
 *
 * The includer defines, before including:
 *     VVC_REAL4, VVC_MIXED4   4-vectors with .x .y .z .w (posq / posqCorrection; velm / posDelta)
 *     VVC_MIXED               scalar type of velm / posDelta
 *     VVC_FN                  function qualifiers (e.g. `static inline`, `__host__ __device__ inline`)
 */
#ifndef VVC_CONSTRAINT_STANDIN_H_
#define VVC_CONSTRAINT_STANDIN_H_

#include <stdint.h>

/* a product that is never fused with a following add: nvcc contracts by default (-fmad=true), the host builds of this
 * repository are compiled with -ffp-contract=off */
#if defined(__CUDACC__)
__host__ __device__ inline double vvc_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__host__ __device__ inline float vvc_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
#define VVC_MUL(a, b) vvc_mul((a), (b))
#else
#define VVC_MUL(a, b) ((a) * (b))
#endif

typedef struct {
    int32_t numClusters;
    int32_t iterations;             /* sweeps over every cluster */
    const int32_t *clusterOffset;   /* [numClusters + 1] prefix into the constraint list */
    const int32_t *atoms;           /* [2 * numConstraints] (i, j), cluster-major */
    const double *distance;         /* [numConstraints] target |x_i - x_j| */
} vvc_constraints;

/* position of particle i as OpenMM reconstructs it: posq (+ posqCorrection in mixed mode) */
VVC_FN void vvc_load_pos(const VVC_REAL4 *posq, const VVC_REAL4 *corr, int i, VVC_MIXED *x) {
    x[0] = (VVC_MIXED) posq[i].x; x[1] = (VVC_MIXED) posq[i].y; x[2] = (VVC_MIXED) posq[i].z;
    if (corr) {
        x[0] += (VVC_MIXED) corr[i].x; x[1] += (VVC_MIXED) corr[i].y; x[2] += (VVC_MIXED) corr[i].z;
    }
}

/* applyConstraints stand-in for one cluster: SHAKE on posDelta along the old bond vectors */
VVC_FN void vvc_cluster_positions(const vvc_constraints cs, int c, const VVC_REAL4 *posq, const VVC_REAL4 *corr,
                                  const VVC_MIXED4 *velm, VVC_MIXED4 *posDelta) {
    for (int sweep = 0; sweep < cs.iterations; sweep++)
        for (int k = cs.clusterOffset[c]; k < cs.clusterOffset[c + 1]; k++) {
            const int i = cs.atoms[2 * k], j = cs.atoms[2 * k + 1];
            const VVC_MIXED wi = velm[i].w, wj = velm[j].w;
            if (wi + wj == 0)
                continue;
            VVC_MIXED xi[3], xj[3];
            vvc_load_pos(posq, corr, i, xi);
            vvc_load_pos(posq, corr, j, xj);
            const VVC_MIXED r0x = xi[0] - xj[0], r0y = xi[1] - xj[1], r0z = xi[2] - xj[2];
            const VVC_MIXED rx = r0x + (posDelta[i].x - posDelta[j].x);
            const VVC_MIXED ry = r0y + (posDelta[i].y - posDelta[j].y);
            const VVC_MIXED rz = r0z + (posDelta[i].z - posDelta[j].z);
            const VVC_MIXED d0 = (VVC_MIXED) cs.distance[k];
            const VVC_MIXED diff = VVC_MUL(d0, d0) - (VVC_MUL(rx, rx) + VVC_MUL(ry, ry) + VVC_MUL(rz, rz));
            const VVC_MIXED rr0 = VVC_MUL(rx, r0x) + VVC_MUL(ry, r0y) + VVC_MUL(rz, r0z);
            const VVC_MIXED g = diff / VVC_MUL(VVC_MUL((VVC_MIXED) 2, wi + wj), rr0);
            const VVC_MIXED gi = VVC_MUL(g, wi), gj = VVC_MUL(g, wj);
            posDelta[i].x += VVC_MUL(gi, r0x); posDelta[i].y += VVC_MUL(gi, r0y); posDelta[i].z += VVC_MUL(gi, r0z);
            posDelta[j].x -= VVC_MUL(gj, r0x); posDelta[j].y -= VVC_MUL(gj, r0y); posDelta[j].z -= VVC_MUL(gj, r0z);
        }
}

/* applyVelocityConstraints stand-in for one cluster: RATTLE, relative velocity along each bond removed */
VVC_FN void vvc_cluster_velocities(const vvc_constraints cs, int c, const VVC_REAL4 *posq, const VVC_REAL4 *corr,
                                   VVC_MIXED4 *velm) {
    for (int sweep = 0; sweep < cs.iterations; sweep++)
        for (int k = cs.clusterOffset[c]; k < cs.clusterOffset[c + 1]; k++) {
            const int i = cs.atoms[2 * k], j = cs.atoms[2 * k + 1];
            const VVC_MIXED wi = velm[i].w, wj = velm[j].w;
            if (wi + wj == 0)
                continue;
            VVC_MIXED xi[3], xj[3];
            vvc_load_pos(posq, corr, i, xi);
            vvc_load_pos(posq, corr, j, xj);
            const VVC_MIXED r0x = xi[0] - xj[0], r0y = xi[1] - xj[1], r0z = xi[2] - xj[2];
            const VVC_MIXED vx = velm[i].x - velm[j].x, vy = velm[i].y - velm[j].y, vz = velm[i].z - velm[j].z;
            const VVC_MIXED vr = VVC_MUL(vx, r0x) + VVC_MUL(vy, r0y) + VVC_MUL(vz, r0z);
            const VVC_MIXED r2 = VVC_MUL(r0x, r0x) + VVC_MUL(r0y, r0y) + VVC_MUL(r0z, r0z);
            const VVC_MIXED k2 = vr / VVC_MUL(wi + wj, r2);
            const VVC_MIXED ki = VVC_MUL(k2, wi), kj = VVC_MUL(k2, wj);
            velm[i].x -= VVC_MUL(ki, r0x); velm[i].y -= VVC_MUL(ki, r0y); velm[i].z -= VVC_MUL(ki, r0z);
            velm[j].x += VVC_MUL(kj, r0x); velm[j].y += VVC_MUL(kj, r0y); velm[j].z += VVC_MUL(kj, r0z);
        }
}

#endif /* VVC_CONSTRAINT_STANDIN_H_ */
