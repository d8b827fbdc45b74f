// vvb200_device.cu -- device-side plan state, the small kernels, launch configuration and the C-ABI step calls.
//
// Design: the reference's 10-16 launches per step collapse into two
// streaming passes over the particle arrays, both working on molecule-aligned tiles (vvb200_stream.cuh):
//
//   pass A  kick_reduce_kernel   velm += dt*w*(F + Fextra)   [integrateMiddleVel, middle.cu:6; extra forces computed
//           inline: drudeLangevin.cu:2 (via the compact ldForce array), electricField.cu:2, cosineAccelerate.cu:2];
//           per-molecule COM velocity [calcCOMVelocities, drudeNoseHoover.cu:5]; group kinetic energies
//           [normalizeVelocities :37 + computeNormalizedKineticEnergies :55] and the velocity-bias moments
//           [calcPeriodicVelocityBias, cosineAccelerate.cu:16]; deterministic two-level reduction [replaces
//           sumNormalizedKineticEnergies :121 and sumV :34]; the LAST block advances the Nose-Hoover chains on the
//           device [VVIntegrator::propagateNHChain, VVIntegrator.cpp:340-376] -- no D2H/H2D round trip
//           (CudaVVKernels.cpp:709-746).
//   pass B  scale_drift_kernel   thermostat scaling [scaleVelocity, drudeNoseHoover.cu:157], bias remove/restore
//           [cosineAccelerate.cu:63,76], both half drifts and the double-float position write
//           [integrateMiddlePos1/2/3, middle.cu:29-100], Drude hard wall [applyHardWallConstraints, middle.cu:106].
//
// Systems that cannot be tiled run the thermostat through the gather kernels of vvb200_general.cuh; multi-GPU runs
// exchange the reduction vector inside nhc_peer_kernel (below).  All sums are fp64 (`mixed`) in a fixed order:
// results are bitwise reproducible run to run.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vvb200_internal.h"

#define BOLTZ_D (1.380649e-23 * 6.02214076e23 / 1000.0)
#define AVOGADRO_D 6.02214076e23

#define THREADS 256                       // block size of the small element-wise kernels
#define VVB200_MAX_BLOCKS_PER_SM 8

// ------------------------------------------------------------------------------------------------
// precision traits: OpenMM's CudaPrecision modes
// ------------------------------------------------------------------------------------------------
struct alignas(16) F4 { float x, y, z, w; };
struct alignas(32) D4 { double x, y, z, w; };
struct F3 { float x, y, z; };
struct D3 { double x, y, z; };

template <int MODE> struct Prec;
template <> struct Prec<VVB200_SINGLE> {
    typedef float real; typedef float mixed; typedef F4 real4; typedef F4 mixed4; typedef F3 real3;
    static constexpr bool kMixed = false;
};
template <> struct Prec<VVB200_MIXED> {
    typedef float real; typedef double mixed; typedef F4 real4; typedef D4 mixed4; typedef F3 real3;
    static constexpr bool kMixed = true;
};
template <> struct Prec<VVB200_DOUBLE> {
    typedef double real; typedef double mixed; typedef D4 real4; typedef D4 mixed4; typedef D3 real3;
    static constexpr bool kMixed = false;
};

// SQRT / RECIP as OpenMM defines them for the JIT'ed kernels [OMM-mem]: sqrtf and 1.0f/(x) unless
// CudaPrecision=double.  (1.0f/(double) is still a double division.)
template <int MODE, class T> __device__ __forceinline__ T vv_sqrt(T x) {
    if (MODE == VVB200_DOUBLE) return (T) sqrt((double) x);
    return (T) sqrtf((float) x);
}
__device__ __forceinline__ float vv_recip(float x) { return 1.0f / x; }
__device__ __forceinline__ double vv_recip(double x) { return 1.0 / x; }
// Reciprocal for the MASSES that enter the kinetic-energy / centre-of-mass reductions (m = 1/velm.w, mu = m1 m2/(m1+m2)):
// hardware seed (upper 20 mantissa bits) + two Newton steps, 5 instructions instead of the ~14 of the IEEE division
// routine, within 1 ulp of it.  The reductions are the only consumers: the sums change at the 1e-16 level, nothing that
// is written back per particle goes through here.
__device__ __forceinline__ double rcpMass(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return r;
}
__device__ __forceinline__ float rcpMass(float x) { return 1.0f / x; }

// ---- streaming loads / stores ------------------------------------------------------------------
__device__ __forceinline__ D4 ld_stream(const D4 *p) {
    D4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ F4 ld_stream(const F4 *p) {
    F4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(D4 *p, const D4 &v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(F4 *p, const F4 &v) {
    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ long long ld_force(const long long *p) {
    long long v;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// ------------------------------------------------------------------------------------------------
// device-resident thermostat state
// ------------------------------------------------------------------------------------------------
struct NhcDevice {
    double eta[3][VVB200_MAX_CHAINS];
    double etaDot[3][VVB200_MAX_CHAINS + 1];
    double etaDotDot[3][VVB200_MAX_CHAINS];
    double etaMass[3][VVB200_MAX_CHAINS];
    double NkbT[3];
    double tTarget[3];
    double ke2[3];
    double vscale[3];
    double red[VVB200_NRED];     // this rank's (or, after an all-reduce, the global) sums
    double vBias;
    double invMassTotal;
    // the part of the NEXT chain update that does not depend on the next kinetic energy (nhcPre), run at the end of the
    // previous one: etaDot[k >= 1] after the first sweep and the exponential that multiplies etaDot[0]; valid for the step
    // size preDt only (0 = not valid: the state was set from the host, or advanced by the unsplit routine)
    double preEtaDot[3][VVB200_MAX_CHAINS + 1];
    double preE0[3];
    double preDt;
    int numTG, nc, loops;
};

// ---- all-reduce of the reduction vector over NVLink peer memory ---------------------------------------------------
// Multi-GPU runs (one process per GPU, particles partitioned by whole molecules) need ONE exchange per step: the sum
// over ranks of <= 10 doubles.  No collective launch: the block that already holds this rank's final sums -- the LAST
// block of pass A (lastBlockFinish), or the single block of nhc_peer_kernel on the paths that have no pass A --
// stores the vector into a slot of every peer's exchange buffer (cudaIpc-mapped, plain st.global over NVLink),
// publishes a sequence number with release semantics, waits for the other ranks' numbers with acquire loads, sums the
// slots in rank order (=> bitwise identical on every rank) and goes on to the NH chains.  Slots are double-buffered
// by step parity; a rank cannot run two steps ahead because it needs every peer's flag of the step in between.  The
// sequence number lives in device memory and is advanced by the exchanging block itself, so a captured CUDA graph
// replays correctly.  The wait is bounded (VVB200_PEER_TIMEOUT_S, default 30 s): on expiry the block raises a flag in
// mapped host memory -- the next host call on the plan returns VVB200_ERR_CUDA -- and poisons the sums with NaN so that
// no silently wrong trajectory continues; the GPU is never left hanging on a dead peer.
#define VVB200_MAX_RANKS 8
struct PeerSlots {
    double data[2][VVB200_MAX_RANKS][16];
    unsigned long long flag[2][VVB200_MAX_RANKS];
};
struct PeerCtx {
    PeerSlots *buf[VVB200_MAX_RANKS];    // buf[r] = rank r's exchange buffer as mapped in this process
    int rank, world;
    unsigned long long timeoutNs;
    unsigned long long *seq;             // device: exchanges done so far
    volatile unsigned int *timedOut;     // mapped host memory: set when a wait expired
};

// One 16-byte record per scale factor: the value and a tag, written with ONE 128-bit store and read with one 128-bit load
// (a 16-byte aligned access is a single transaction), so a reader that sees the tag has the value -- no fence between
// payload and flag, nothing to fetch after the flag.
struct __align__(16) FactorRec {
    double value;
    unsigned long long tag;
};
__device__ __forceinline__ void factorPublish(FactorRec *r, double value, unsigned long long tag) {
    asm volatile("st.global.v2.u64 [%0], {%1, %2};" ::"l"(r), "l"(__double_as_longlong(value)), "l"(tag) : "memory");
}
__device__ __forceinline__ bool factorPoll(const FactorRec *r, unsigned long long tag, double *value) {
    unsigned long long v, t;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(r) : "memory");
    *value = __longlong_as_double((long long) v);
    return t == tag;
}

struct KParams {
    int N, paddedN, numTiles;
    const int4 *tileDesc;                      // 2 x int4 per tile, see vvb200_stream.cuh
    const int32_t *tileMolList, *tileMolInfo;
    const uint32_t *slotMeta;
    const int32_t *ldSlot;
    const int32_t *sortedByMol, *particlesInMolecules;
    // image update fused into pass B: image index per parent, mirror plane
    const int32_t *imageOf;
    double mirror;
    int imageFused;
    // Langevin force evaluated inside the kick (middle scheme): OpenMM's N(0,1) buffer, the index of this step's first
    // number, the coefficients of CudaVVKernels.cpp:835-839 and the number of unpaired Langevin particles
    const float4 *random;
    unsigned int randomIndex;
    int ldInline, nNormalLD;
    double ldDrag, ldRand, ldDragDrude, ldRandDrude;
    // thermostat molecules cut across tiles (longer than a tile): fragment index per tile-local molecule, the cut
    // molecules with their fragment lists, and the per-fragment sums (sum m v (3), sum m, sum m c) of this step
    const int32_t *tileMolFrag, *splitMolId, *splitFragOffset, *splitFragList;
    double *fragPartials;
    int numSplit;
    void *posq, *corr, *velm;
    void *posDelta, *oldDelta;   // VAR_SCALE_DELTA only
    const long long *force;
    const void *ldForce;
    void *comV;            // mixed4 per molecule
    void *comCbar;         // mixed per molecule (cosine runs)
    double *partials;      // [gridDim.x][NRED]
    NhcDevice *nhc;
    unsigned int *counter;
    double dt;
    double efscale, accel, invBoxZ, maxDrudeDistance, hardwallScale;
    int useCOM, hasLD, hasField, hardwall, extraForces, fuseNHC, cosine, kickOnly;
    double maxD2safeD;                         // conservative pre-test of the hard wall: maxDrudeDistance^2 (1 - 1e-4), and as
    float maxD2safeF;                          // the float the single / mixed kernels compare with
    int stagesA, stagesB;
    // single-launch resident kernel (vvb200_resident.cuh)
    int tilesPerBlock, doReduce;
    // streaming kernels: the tile range of this launch (the whole system unless the host pipeline splits a step) and
    // whether its sums are added to those already in nhc->red
    int tileBegin, tileEnd, accumulateRed;
    int pdl;     // launch with programmatic stream serialization (host side only)
    unsigned int *gridGen;   // generation word of its grid barrier
    // pass B synchronised with pass A's last block through a device word instead of waiting for the whole grid (see
    // "early hand-over" in vvb200_stream.cuh): 0 = reset by pass A, 1 = every tile's velocities are stored and visible,
    // 2 = this step's scale factors are in nhc->vscale
    unsigned int *syncFlag;
    unsigned int *tileTicket;      // pass B under the hand-over takes its tiles from this counter (zeroed by pass A's block 0)
    FactorRec *factorRec;          // [4]: this step's three scale factors and the velocity bias, each with its own "valid" tag
    int flagSync;
    int tileReverse;               // ... starting from the last tile
    // multi-GPU: pass A's last block exchanges the reduction vector with the peers before it advances the chains
    int peerOn;
    PeerCtx peer;
};

enum { KICK_NONE = 0, KICK_MIDDLE = 1, KICK_VV = 2 };
enum { VAR_MIDDLE = 0, VAR_VV_FIRST = 1, VAR_SCALE_ONLY = 2, VAR_SCALE_DELTA = 3, VAR_FINISH = 4, VAR_VV_POSITIONS = 5 };

// tileMolInfo word: bits 0-10 first slot in tile, bits 11-21 count, bit 31 = not contiguous
#define MOLINFO_FIRST(w) ((w) & 0x7FF)
#define MOLINFO_COUNT(w) (((w) >> 11) & 0x7FF)
#define MOLINFO_SCATTERED(w) (((w) >> 31) & 1)
#define MOLINFO_FRAGMENT(w) (((w) >> 30) & 1)     // part of a molecule that continues in a neighbouring tile

__device__ __forceinline__ double cosPhase(double z, double invBoxZ) {
    // the reference's 8-digit pi literal and double-precision cos (cosineAccelerate.cu:9,26,70,83)
    return cos(2 * 3.1415926 * z * invBoxZ);
}

// VVIntegrator::propagateNHChain on the device (VVIntegrator.cpp:340-376).  The chain state is pulled into
// thread-local arrays first (independent loads, one L2 round trip) and written back at the end: working on the
// global arrays in place serialises ~50 dependent L2 accesses (~0.13 us each) and was most of a small system's step.
// NC > 0: chain length known at compile time, loops unrolled, state in registers (the common lengths 1..4);
// NC == 0: any length up to VVB200_MAX_CHAINS from local memory.  Same operation order either way.
template <int NC>
__device__ __forceinline__ double nhcPropagateT(NhcDevice *s, int g, double dt, double ke2) {
    constexpr int CAP = NC > 0 ? NC : VVB200_MAX_CHAINS;
    const int nc = NC > 0 ? NC : s->nc;
    const int loops = s->loops;
    double eta[CAP], etaDot[CAP + 1], etaDotDot[CAP], Q[CAP];
#pragma unroll
    for (int k = 0; k < CAP; k++) {
        if (k < nc) {
            eta[k] = s->eta[g][k];
            etaDot[k] = s->etaDot[g][k];
            etaDotDot[k] = s->etaDotDot[g][k];
            Q[k] = s->etaMass[g][k];
        }
    }
    etaDot[nc] = s->etaDot[g][nc];
    const double target = s->NkbT[g];
    const double kT = BOLTZ_D * s->tTarget[g];
    const double h2 = dt / loops / 2, h4 = h2 / 2, h8 = h4 / 2;
    double factor = 1.0, e = 0.0;
    etaDotDot[0] = (ke2 - target) / Q[0];
    for (int l = 0; l < loops; l++) {
#pragma unroll
        for (int k = CAP - 1; k >= 0; k--) {
            if (k < nc) {
                e = exp(-h8 * etaDot[k + 1]);
                etaDot[k] *= e;
                etaDot[k] += etaDotDot[k] * h4;
                etaDot[k] *= e;
            }
        }
        factor *= exp(-h2 * etaDot[0]);
#pragma unroll
        for (int k = 0; k < CAP; k++)
            if (k < nc) eta[k] += h2 * etaDot[k];
        etaDotDot[0] = (ke2 * factor * factor - target) / Q[0];
        etaDot[0] *= e;
        etaDot[0] += etaDotDot[0] * h4;
        etaDot[0] *= e;
#pragma unroll
        for (int k = 1; k < CAP; k++) {
            if (k < nc) {
                e = exp(-h8 * etaDot[k + 1]);
                etaDot[k] *= e;
                etaDotDot[k] = (Q[k - 1] * etaDot[k - 1] * etaDot[k - 1] - kT) / Q[k];
                etaDot[k] += etaDotDot[k] * h4;
                etaDot[k] *= e;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < CAP; k++) {
        if (k < nc) {
            s->eta[g][k] = eta[k];
            s->etaDot[g][k] = etaDot[k];
            s->etaDotDot[g][k] = etaDotDot[k];
        }
    }
    return factor;
}

// not inlined: one copy for all kernels, and its registers / local arrays stay out of the streaming kernels' budget
__device__ __noinline__ double nhcPropagate(NhcDevice *s, int g, double dt, double ke2) {
    switch (s->nc) {
    case 1: return nhcPropagateT<1>(s, g, dt, ke2);
    case 2: return nhcPropagateT<2>(s, g, dt, ke2);
    case 3: return nhcPropagateT<3>(s, g, dt, ke2);
    case 4: return nhcPropagateT<4>(s, g, dt, ke2);
    default: return nhcPropagateT<0>(s, g, dt, ke2);
    }
}

// From the reduced sums to bias, group energies and scale factors (CudaVVKernels.cpp:709-746 moved
// onto the device).  Called by threads 0..2 of one block; `red` must already be final.
template <bool COS>
__device__ void nhcFinish(NhcDevice *s, double dt, int g, const double *red = nullptr) {
    if (!red) red = s->red;      // callers that just summed the vector pass their shared-memory copy
    double V = 0.0;
    if (COS)
        V = red[3] * s->invMassTotal;
    double ke2 = red[g];
    if (COS)
        ke2 = red[g] - 2.0 * V * red[4 + g] + V * V * red[7 + g];
    double scale = 1.0;
    if (g < s->numTG) {
        if (s->etaMass[g][0] > 0)
            scale = nhcPropagate(s, g, dt, ke2);
    } else {
        ke2 = 0.0;
    }
    s->ke2[g] = ke2;
    s->vscale[g] = scale;
    if (g == 0) {
        s->vBias = V;
        s->preDt = 0.0;      // the unsplit routine does not look ahead
    }
}

// ---- the chain update cut where the scale factor is known ---------------------------------------------------------
// propagateNHChain (VVIntegrator.cpp:340-376) with loopsPerStep == 1 (the constructor's default) touches this step's
// kinetic energy in exactly one place before the factor is final: etaDotDot[0].  Everything above it in the first sweep
// (the chains ich = nc-1 .. 1 and the exponential that multiplies etaDot[0]) depends on the previous step only, and
// everything after `factor` only prepares the next step.  So the block that finishes the reduction runs
//     nhcCrit  ke2 -> etaDotDot[0] -> etaDot[0] -> factor = exp(-dt/2 etaDot[0]): one division and ONE exp on the
//              critical path instead of six dependent exps and four divisions; the factor is published right away
//     nhcPost  eta, the second sweep, the state for the next step -- and then nhcPre for the NEXT step, whose results
//              (preEtaDot, preE0) travel in the thermostat state: all of it off the critical path
// The operations, their operands and their order are those of nhcPropagateT (bit-identical results); other loop counts
// and groups without a thermostat mass keep the unsplit routine inside nhcCrit.  preDt says which step size the
// look-ahead is valid for; anything that sets the state from outside zeroes it and nhcCrit then runs nhcPre itself.
__device__ __forceinline__ bool nhcSplittable(const NhcDevice *s, int g) {
    return s->loops == 1 && g < s->numTG && s->etaMass[g][0] > 0;
}

__device__ __noinline__ void nhcPre(NhcDevice *s, int g, double dt) {
    if (!nhcSplittable(s, g))
        return;
    const int nc = s->nc;
    const double h2 = dt / s->loops / 2, h4 = h2 / 2, h8 = h4 / 2;
    double next = s->etaDot[g][nc];           // etaDot[k + 1] as the sweep sees it
    for (int k = nc - 1; k >= 1; k--) {
        const double e = exp(-h8 * next);
        double x = s->etaDot[g][k];
        x *= e;
        x += s->etaDotDot[g][k] * h4;
        x *= e;
        s->preEtaDot[g][k] = x;
        next = x;
    }
    s->preE0[g] = exp(-h8 * next);
}

// threads 0..2 of the block that holds the final sums; ke2 / vscale / vBias land in *s
template <bool COS>
__device__ __noinline__ void nhcCrit(NhcDevice *s, double dt, int g, const double *red) {
    double V = 0.0;
    if (COS)
        V = red[3] * s->invMassTotal;
    double ke2 = red[g];
    if (COS)
        ke2 = red[g] - 2.0 * V * red[4 + g] + V * V * red[7 + g];
    double scale = 1.0;
    if (g < s->numTG) {
        if (nhcSplittable(s, g)) {
            if (s->preDt != dt)
                nhcPre(s, g, dt);       // no valid look-ahead (first step, state set from the host, step size changed)
            for (int k = 1; k < s->nc; k++)
                s->etaDot[g][k] = s->preEtaDot[g][k];
            const double e0 = s->preE0[g];
            const double h2 = dt / s->loops / 2, h4 = h2 / 2;
            const double dd = (ke2 - s->NkbT[g]) / s->etaMass[g][0];
            double x = s->etaDot[g][0];
            x *= e0;
            x += dd * h4;
            x *= e0;
            s->etaDot[g][0] = x;
            s->etaDotDot[g][0] = dd;
            scale = exp(-h2 * x);       // factor = 1.0 * exp(...): the product with 1.0 is exact
        } else if (s->etaMass[g][0] > 0) {
            scale = nhcPropagate(s, g, dt, ke2);
        }
    } else {
        ke2 = 0.0;
    }
    s->ke2[g] = ke2;
    s->vscale[g] = scale;
    if (g == 0)
        s->vBias = V;
}

// after nhcCrit, by the same three threads; `s` must not be read by anybody else until they are done
__device__ __noinline__ void nhcPost(NhcDevice *s, double dt, int g) {
    if (!nhcSplittable(s, g)) {
        if (g == 0) s->preDt = 0.0;
        return;
    }
    const int nc = s->nc;
    const double h2 = dt / s->loops / 2, h4 = h2 / 2, h8 = h4 / 2;
    const double kT = BOLTZ_D * s->tTarget[g];
    const double ke2 = s->ke2[g], factor = s->vscale[g], e0 = s->preE0[g];
    for (int k = 0; k < nc; k++)
        s->eta[g][k] += h2 * s->etaDot[g][k];
    const double dd = (ke2 * factor * factor - s->NkbT[g]) / s->etaMass[g][0];
    s->etaDotDot[g][0] = dd;
    double x = s->etaDot[g][0];
    x *= e0;
    x += dd * h4;
    x *= e0;
    s->etaDot[g][0] = x;
    for (int k = 1; k < nc; k++) {
        const double e = exp(-h8 * s->etaDot[g][k + 1]);
        double y = s->etaDot[g][k];
        y *= e;
        const double ddk = (s->etaMass[g][k - 1] * s->etaDot[g][k - 1] * s->etaDot[g][k - 1] - kT) / s->etaMass[g][k];
        s->etaDotDot[g][k] = ddk;
        y += ddk * h4;
        y *= e;
        s->etaDot[g][k] = y;
    }
    nhcPre(s, g, dt);               // the look-ahead for the next step
    if (g == 0) s->preDt = dt;      // (every group of a plan is splittable or none is: loops is shared)
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded wait for *f >= want (a producer that never comes must not hang the device): false when it expired (~2 s)
__device__ __forceinline__ bool waitFlagAtLeast(const unsigned int *f, unsigned int want) {
    const long long c0 = clock64();
    while (ld_acquire_gpu(f) < want) {
        if (clock64() - c0 > 4000000000LL)
            return false;
        __nanosleep(40);
    }
    return true;
}

template <bool COS>
__global__ void nhc_kernel(NhcDevice *s, double dt) {
    if (threadIdx.x < 3)
        nhcFinish<COS>(s, dt, threadIdx.x);
}

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// First half of the exchange, by thread t < world of the exchanging block: my vector (`mine`, VVB200_NRED doubles in
// shared or global memory) into rank t's buffer, then wait for rank t's vector in my own buffer.  Returns false when the
// wait expired.  `seq` = number of this exchange (1, 2, ...), the same on every rank.
__device__ __forceinline__ bool peerPublishAndWait(const PeerCtx &ctx, unsigned long long seq, const double *mine, int t) {
    const int par = (int) (seq & 1);
    double *dst = ctx.buf[t]->data[par][ctx.rank];
    for (int k = 0; k < VVB200_NRED; k++)
        dst[k] = mine[k];
    __threadfence_system();
    st_release_sys(&ctx.buf[t]->flag[par][ctx.rank], seq);
    const unsigned long long *f = &ctx.buf[ctx.rank]->flag[par][t];
    const unsigned long long t0 = globalTimerNs();
    while (ld_acquire_sys(f) != seq) {
        if (globalTimerNs() - t0 > ctx.timeoutNs) {      // a peer died or stalled: do not hang the device
            *ctx.timedOut = 1u;
            return false;
        }
        __nanosleep(100);
    }
    return true;
}
// Second half, by thread t < VVB200_NRED after a block barrier: the sum over ranks in rank order
__device__ __forceinline__ double peerSum(const PeerCtx &ctx, unsigned long long seq, int t, bool expired) {
    const int par = (int) (seq & 1);
    double v = 0;
    for (int r = 0; r < ctx.world; r++)
        v += __ldcg(&ctx.buf[ctx.rank]->data[par][r][t]);
    return expired ? __longlong_as_double(0x7ff8000000000000LL) : v;
}

// Exchange + NH chains as a kernel of its own: the paths whose sums do not come out of pass A's last block (the
// any-topology gather kernels, the chunked host pipeline)
template <bool COS>
__global__ void nhc_peer_kernel(NhcDevice *s, const PeerCtx ctx, double dt) {
    __shared__ int expired;
    __shared__ unsigned long long seqS;
    const int t = threadIdx.x;
    if (t == 0) {
        expired = 0;
        seqS = ++*ctx.seq;
    }
    __syncthreads();
    const unsigned long long seq = seqS;
    if (t < ctx.world && !peerPublishAndWait(ctx, seq, s->red, t))
        expired = 1;
    __syncthreads();
    if (t < VVB200_NRED)
        s->red[t] = peerSum(ctx, seq, t, expired != 0);
    __syncthreads();
    if (t < 3)
        nhcFinish<COS>(s, dt, t);
}

#include "vvb200_stream.cuh"
#include "vvb200_resident.cuh"
#include "vvb200_general.cuh"

// ------------------------------------------------------------------------------------------------
// small gather kernels
// ------------------------------------------------------------------------------------------------

// Langevin force into the compact ldForce array (drudeLangevin.cu:2-59): value = what the reference
// holds in forceExtra right after resetExtraForce + addExtraForceDrudeLangevin.
template <int MODE>
__global__ void langevin_force_kernel(const void *velmRaw, void *ldForceRaw, const int32_t *normalLD, int nNormal,
                                      const int2 *pairsLD, int nPairs, double dragD, double randD, double dragDrudeD,
                                      double randDrudeD, const float4 *random, unsigned int randomIndex) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::mixed4 mixed4;
    typedef typename P::real3 real3;
    const mixed4 *velm = reinterpret_cast<const mixed4 *>(velmRaw);
    real3 *out = reinterpret_cast<real3 *>(ldForceRaw);
    const mixed dragFactor = (mixed) dragD, randFactor = (mixed) randD;
    const mixed dragFactorDrude = (mixed) dragDrudeD, randFactorDrude = (mixed) randDrudeD;
    const int stride = blockDim.x * gridDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nNormal; i += stride) {
        const mixed4 v = velm[normalLD[i]];
        real3 f; f.x = 0; f.y = 0; f.z = 0;
        if (v.w != 0) {
            const mixed mass = vv_recip(v.w);
            const mixed sqrtMass = vv_sqrt<MODE, mixed>(mass);
            const float4 r = random[randomIndex + i];
            f.x += (-dragFactor * mass * v.x + randFactor * sqrtMass * r.x);
            f.y += (-dragFactor * mass * v.y + randFactor * sqrtMass * r.y);
            f.z += (-dragFactor * mass * v.z + randFactor * sqrtMass * r.z);
        }
        out[i] = f;
    }
    randomIndex += nNormal;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nPairs; i += stride) {
        const int2 pr = pairsLD[i];
        const mixed4 v1 = velm[pr.x], v2 = velm[pr.y];
        const mixed mass1 = vv_recip(v1.w), mass2 = vv_recip(v2.w);
        const mixed totMass = mass1 + mass2;
        const mixed sqrtTotMass = vv_sqrt<MODE, mixed>(totMass);
        const mixed redMass = vv_recip((mass1 + mass2) * v1.w * v2.w);
        const mixed sqrtRedMass = vv_sqrt<MODE, mixed>(redMass);
        const mixed invTotMass = vv_recip(totMass);
        const mixed m1f = invTotMass * mass1, m2f = invTotMass * mass2;
        const float4 r1 = random[randomIndex + 2 * i], r2 = random[randomIndex + 2 * i + 1];
        const mixed cm[3] = {v1.x * m1f + v2.x * m2f, v1.y * m1f + v2.y * m2f, v1.z * m1f + v2.z * m2f};
        const mixed rel[3] = {v2.x - v1.x, v2.y - v1.y, v2.z - v1.z};
        const float ra[3] = {r1.x, r1.y, r1.z}, rb[3] = {r2.x, r2.y, r2.z};
        real a[3], b[3];
        const real f1 = (real) m1f, f2 = (real) m2f;   // real3 * mixed resolves to operator*(real3, real)
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const real cmForce = (real) (-dragFactor * totMass * cm[d] + randFactor * sqrtTotMass * ra[d]);
            const real relForce = (real) (-dragFactorDrude * redMass * rel[d] + randFactorDrude * sqrtRedMass * rb[d]);
            a[d] = (real) 0 + (f1 * cmForce - relForce);
            b[d] = (real) 0 + (f2 * cmForce + relForce);
        }
        real3 fa, fb;
        fa.x = a[0]; fa.y = a[1]; fa.z = a[2];
        fb.x = b[0]; fb.y = b[1]; fb.z = b[2];
        out[nNormal + 2 * i] = fa;
        out[nNormal + 2 * i + 1] = fb;
    }
}

// updateImagePositions (imageCharge.cu:2-27)
template <int MODE>
__global__ void image_kernel(void *posqRaw, void *corrRaw, const int2 *imagePairs, int nImages, double mirrorD) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    real4 *posq = reinterpret_cast<real4 *>(posqRaw);
    real4 *corr = reinterpret_cast<real4 *>(corrRaw);
    const mixed mirror = (mixed) mirrorD;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nImages; i += blockDim.x * gridDim.x) {
        const int2 pr = imagePairs[i];
        const real4 pp = posq[pr.y];
        real4 pi = posq[pr.x];
        pi.x = pp.x;
        pi.y = pp.y;
        if (P::kMixed) {
            const real4 pc = corr[pr.y];
            real4 ic = corr[pr.x];
            ic.x = pc.x;
            ic.y = pc.y;
            mixed z = (mixed) pp.z + (mixed) pc.z;
            z = mirror * 2 - z;
            pi.z = (real) z;
            ic.z = (real) (z - (real) z);
            corr[pr.x] = ic;
        } else {
            pi.z = 2 * mirror - pp.z;
        }
        posq[pr.x] = pi;
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise kernels of the constraint-bearing path (OpenMM's constraint kernels run between them)
// ------------------------------------------------------------------------------------------------

// integrateMiddlePos1 / integrateMiddlePos2 (middle.cu:29-61)
template <int MODE>
__global__ void __launch_bounds__(THREADS) delta_kernel(const void *velmRaw, void *posDeltaRaw, void *oldDeltaRaw, int N,
                                                        double dt, int accumulate) {
    typedef typename Prec<MODE>::mixed mixed;
    typedef typename Prec<MODE>::mixed4 mixed4;
    const mixed4 *velm = reinterpret_cast<const mixed4 *>(velmRaw);
    mixed4 *posDelta = reinterpret_cast<mixed4 *>(posDeltaRaw);
    mixed4 *oldDelta = reinterpret_cast<mixed4 *>(oldDeltaRaw);
    const mixed halfdt = 0.5f * (mixed) dt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) {
        const mixed4 v = ld_stream(velm + i);
        if (v.w != 0) {
            mixed4 d;
            d.x = halfdt * v.x; d.y = halfdt * v.y; d.z = halfdt * v.z; d.w = 0;
            if (accumulate) {
                mixed4 a = ld_stream(posDelta + i), b = ld_stream(oldDelta + i);
                a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
                b.x += d.x; b.y += d.y; b.z += d.z; b.w += d.w;
                st_stream(posDelta + i, a);
                st_stream(oldDelta + i, b);
            } else {
                st_stream(posDelta + i, d);
                st_stream(oldDelta + i, d);
            }
        }
    }
}

// integrateMiddlePos3 (middle.cu:66-100)
template <int MODE>
__global__ void __launch_bounds__(THREADS) finish_kernel(void *posqRaw, void *corrRaw, const void *posDeltaRaw,
                                                         const void *oldDeltaRaw, void *velmRaw, int N, double dt) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    real4 *posq = reinterpret_cast<real4 *>(posqRaw);
    real4 *corr = reinterpret_cast<real4 *>(corrRaw);
    mixed4 *velm = reinterpret_cast<mixed4 *>(velmRaw);
    const mixed4 *posDelta = reinterpret_cast<const mixed4 *>(posDeltaRaw);
    const mixed4 *oldDelta = reinterpret_cast<const mixed4 *>(oldDeltaRaw);
    const mixed invDt = 1 / (mixed) dt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) {
        mixed4 v = ld_stream(velm + i);
        if (v.w != 0) {
            const mixed4 d = ld_stream(posDelta + i), o = ld_stream(oldDelta + i);
            v.x += (d.x - o.x) * invDt;
            v.y += (d.y - o.y) * invDt;
            v.z += (d.z - o.z) * invDt;
            st_stream(velm + i, v);
            real4 pq = ld_stream(posq + i);
            if (P::kMixed) {
                const real4 c = ld_stream(corr + i);
                mixed x = pq.x + (mixed) c.x, y = pq.y + (mixed) c.y, z = pq.z + (mixed) c.z;
                x += d.x; y += d.y; z += d.z;
                real4 oc;
                splitPos<MODE>(x, pq.x, oc.x);
                splitPos<MODE>(y, pq.y, oc.y);
                splitPos<MODE>(z, pq.z, oc.z);
                oc.w = 0;
                st_stream(posq + i, pq);
                st_stream(corr + i, oc);
            } else {
                pq.x += d.x; pq.y += d.y; pq.z += d.z;
                st_stream(posq + i, pq);
            }
        }
    }
}

// posDelta = dt * v after the first half kick (velocityVerlet.cu:24-26)
template <int MODE>
__global__ void __launch_bounds__(THREADS) vv_delta_kernel(const void *velmRaw, void *posDeltaRaw, int N, double dt) {
    typedef typename Prec<MODE>::mixed mixed;
    typedef typename Prec<MODE>::mixed4 mixed4;
    const mixed4 *velm = reinterpret_cast<const mixed4 *>(velmRaw);
    mixed4 *posDelta = reinterpret_cast<mixed4 *>(posDeltaRaw);
    const mixed stepSize = (mixed) dt;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) {
        const mixed4 v = ld_stream(velm + i);
        if (v.w != 0) {
            mixed4 d;
            d.x = stepSize * v.x; d.y = stepSize * v.y; d.z = stepSize * v.z; d.w = 0;
            st_stream(posDelta + i, d);
        }
    }
}

// velocityVerletIntegratePositions (velocityVerlet.cu:35-68)
template <int MODE>
__global__ void __launch_bounds__(THREADS) vv_positions_kernel(void *posqRaw, void *corrRaw, const void *posDeltaRaw,
                                                               void *velmRaw, int N, double dt) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    real4 *posq = reinterpret_cast<real4 *>(posqRaw);
    real4 *corr = reinterpret_cast<real4 *>(corrRaw);
    mixed4 *velm = reinterpret_cast<mixed4 *>(velmRaw);
    const mixed4 *posDelta = reinterpret_cast<const mixed4 *>(posDeltaRaw);
    const mixed invStepSize = (mixed) (1.0 / (mixed) dt);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += blockDim.x * gridDim.x) {
        mixed4 v = ld_stream(velm + i);
        if (v.w != 0) {
            const mixed4 d = ld_stream(posDelta + i);
            real4 pq = ld_stream(posq + i);
            v.x = (mixed) (invStepSize * d.x);
            v.y = (mixed) (invStepSize * d.y);
            v.z = (mixed) (invStepSize * d.z);
            if (P::kMixed) {
                const real4 c = ld_stream(corr + i);
                mixed x = pq.x + (mixed) c.x, y = pq.y + (mixed) c.y, z = pq.z + (mixed) c.z;
                x += d.x; y += d.y; z += d.z;
                real4 oc;
                splitPos<MODE>(x, pq.x, oc.x);
                splitPos<MODE>(y, pq.y, oc.y);
                splitPos<MODE>(z, pq.z, oc.z);
                oc.w = 0;
                st_stream(posq + i, pq);
                st_stream(corr + i, oc);
            } else {
                pq.x += d.x; pq.y += d.y; pq.z += d.z;
                st_stream(posq + i, pq);
            }
            st_stream(velm + i, v);
        }
    }
}

// applyHardWallConstraints as a pair gather kernel (middle.cu:106-221), used after OpenMM's position
// constraints where the fused pass B cannot be.
template <int MODE>
__global__ void hardwall_pairs_kernel(void *posqRaw, void *corrRaw, void *velmRaw, const int2 *pairs, int nPairs, double dt,
                                      double maxDrudeDistance, double hardwallScale) {
    typedef Prec<MODE> P;
    typedef typename P::real real;
    typedef typename P::mixed mixed;
    typedef typename P::real4 real4;
    typedef typename P::mixed4 mixed4;
    real4 *posq = reinterpret_cast<real4 *>(posqRaw);
    real4 *corr = reinterpret_cast<real4 *>(corrRaw);
    mixed4 *velm = reinterpret_cast<mixed4 *>(velmRaw);
    const mixed stepSize = (mixed) dt, maxD = (mixed) maxDrudeDistance, hwScale = (mixed) hardwallScale;
    const real maxD2safe = (real) (maxDrudeDistance * maxDrudeDistance * (1.0 - 1e-4));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nPairs; i += blockDim.x * gridDim.x) {
        const int2 pr = pairs[i];
        real4 q1 = posq[pr.x], q2 = posq[pr.y];
        real4 c1, c2;
        c1.x = c1.y = c1.z = c1.w = 0;
        c2 = c1;
        if (P::kMixed) { c1 = corr[pr.x]; c2 = corr[pr.y]; }
        // conservative pre-test in `real` arithmetic (see pass B): most pairs are nowhere near the wall
        const real sx = (q1.x - q2.x) + (c1.x - c2.x), sy = (q1.y - q2.y) + (c1.y - c2.y), sz = (q1.z - q2.z) + (c1.z - c2.z);
        if (sx * sx + sy * sy + sz * sz < maxD2safe)
            continue;
        mixed pos1[3] = {(mixed) q1.x, (mixed) q1.y, (mixed) q1.z}, pos2[3] = {(mixed) q2.x, (mixed) q2.y, (mixed) q2.z};
        if (P::kMixed) {
            pos1[0] += (mixed) c1.x; pos1[1] += (mixed) c1.y; pos1[2] += (mixed) c1.z;
            pos2[0] += (mixed) c2.x; pos2[1] += (mixed) c2.y; pos2[2] += (mixed) c2.z;
        }
        const mixed dx = pos1[0] - pos2[0], dy = pos1[1] - pos2[1], dz = pos1[2] - pos2[2];
        const mixed r = vv_sqrt<MODE, mixed>(dx * dx + dy * dy + dz * dz);
        const mixed rInv = vv_recip(r);
        if (!(rInv * maxD < 1))
            continue;
        const mixed bond[3] = {dx * rInv, dy * rInv, dz * rInv};
        mixed4 v1 = velm[pr.x], v2 = velm[pr.y];
        mixed vel1[3] = {v1.x, v1.y, v1.z}, vel2[3] = {v2.x, v2.y, v2.z};
        const mixed mass1 = vv_recip(v1.w), mass2 = vv_recip(v2.w);
        const mixed deltaR = r - maxD;
        mixed deltaT = stepSize;
        mixed dotvr1 = vel1[0] * bond[0] + vel1[1] * bond[1] + vel1[2] * bond[2];
        mixed vp1[3];
        for (int d = 0; d < 3; d++) vp1[d] = vel1[d] - bond[d] * dotvr1;
        const bool both = v2.w != 0;
        if (!both) {
            if (dotvr1 != 0) deltaT = deltaR / fabs(dotvr1);
            if (deltaT > stepSize) deltaT = stepSize;
            dotvr1 = -dotvr1 * hwScale / (fabs(dotvr1) * vv_sqrt<MODE, mixed>(mass1));
            const mixed dr = -deltaR + deltaT * dotvr1;
            for (int d = 0; d < 3; d++) {
                pos1[d] += bond[d] * dr;
                vel1[d] = vp1[d] + bond[d] * dotvr1;
            }
        } else {
            const mixed invTotalMass = vv_recip(mass1 + mass2);
            mixed dotvr2 = vel2[0] * bond[0] + vel2[1] * bond[1] + vel2[2] * bond[2];
            mixed vp2[3];
            for (int d = 0; d < 3; d++) vp2[d] = vel2[d] - bond[d] * dotvr2;
            const mixed vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
            dotvr1 -= vbCMass;
            dotvr2 -= vbCMass;
            if (dotvr1 != dotvr2) deltaT = deltaR / fabs(dotvr1 - dotvr2);
            if (deltaT > stepSize) deltaT = stepSize;
            const mixed vBond = hwScale / vv_sqrt<MODE, mixed>(mass1);
            dotvr1 = -dotvr1 * vBond * mass2 * invTotalMass / fabs(dotvr1);
            dotvr2 = -dotvr2 * vBond * mass1 * invTotalMass / fabs(dotvr2);
            const mixed dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
            const mixed dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
            dotvr1 += vbCMass;
            dotvr2 += vbCMass;
            for (int d = 0; d < 3; d++) {
                pos1[d] += bond[d] * dr1;
                pos2[d] += bond[d] * dr2;
                vel1[d] = vp1[d] + bond[d] * dotvr1;
                vel2[d] = vp2[d] + bond[d] * dotvr2;
            }
        }
        for (int who = 0; who < (both ? 2 : 1); who++) {
            const int idx = who == 0 ? pr.x : pr.y;
            const mixed *x = who == 0 ? pos1 : pos2;
            const mixed *v = who == 0 ? vel1 : vel2;
            real4 o = who == 0 ? q1 : q2;
            if (P::kMixed) {
                real4 oc;
                splitPos<MODE>(x[0], o.x, oc.x);
                splitPos<MODE>(x[1], o.y, oc.y);
                splitPos<MODE>(x[2], o.z, oc.z);
                oc.w = 0;
                posq[idx] = o;
                corr[idx] = oc;
            } else {
                o.x = (real) x[0]; o.y = (real) x[1]; o.z = (real) x[2];
                posq[idx] = o;
            }
            mixed4 ov;
            ov.x = v[0]; ov.y = v[1]; ov.z = v[2]; ov.w = who == 0 ? v1.w : v2.w;
            velm[idx] = ov;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: device state
// ------------------------------------------------------------------------------------------------
struct vvb200_device_state {
    int device = 0;
    int numSM = 148;
    int numTiles = 0;
    int4 *tileDesc = nullptr;
    int32_t *tileMolList = nullptr, *tileMolInfo = nullptr;
    int stagesA = 0, stagesB = 0, blocksPerSM = 0;   // 0: chosen per kernel from the shared-memory budget
    uint32_t *slotMeta = nullptr;
    int32_t *ldSlot = nullptr, *normalLD = nullptr, *sortedByMol = nullptr, *particlesInMolecules = nullptr;
    int32_t *imageOf = nullptr;
    int32_t *tileMolFrag = nullptr, *splitMolId = nullptr, *splitFragOffset = nullptr, *splitFragList = nullptr;
    double *fragPartials = nullptr;
    int2 *pairsLD = nullptr, *imagePairs = nullptr, *drudePairs = nullptr;
    // any-topology path
    int32_t *moleculesNH = nullptr, *normalNH = nullptr, *particleMolId = nullptr;
    int2 *pairsNH = nullptr;
    void *ownPosDelta = nullptr;
    // NVLink peer exchange (multi-GPU)
    PeerSlots *peerLocal = nullptr;
    PeerCtx peer{};
    bool peerAttached = false;
    unsigned int *peerTimedOutHost = nullptr;     // mapped pinned host word the exchanging block raises on expiry
    std::vector<void *> peerMapped;
    void *oldDelta = nullptr;   // mixed4[N], plugin-owned like the reference's (CudaVVKernels.cpp:90-96)
    void *comV = nullptr, *comCbar = nullptr, *ldForce = nullptr;
    double *partials = nullptr;
    NhcDevice *nhc = nullptr;
    unsigned int *counter = nullptr;   // [0] arrival counter of the last-block reductions, [32] grid-barrier generation, [64] pass A -> pass B hand-over word, [96] pass B's tile ticket, [128..143] the four factor records
    int64_t residentLaunches = 0;
    int pdl = 1;                       // VVB200_PDL at upload time
    int handOver = 1;                  // VVB200_HANDOVER at upload time: pass B takes over from pass A through the hand-over word
    int tileReverse = 1;               // VVB200_B_REVERSE at upload time
    int residentMode = -1;
    int residentMaxParticles = -1;     // -1: from the environment (VVB200_RESIDENT_MAX_PARTICLES, default 120000)             // -1: from the environment (VVB200_RESIDENT, default on), 0 off, 1 on
    bool extraForcesValid = false;   // VV scheme: forceExtra is zero until the first second half
    // optional per-kernel timing (vvb200_profile_*): events recorded on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> profEvents;      // 4 per step: before/after pass A, before/after pass B
    size_t profUsed = 0;
    // host staging for vvb200_step_host
    void *hPosq = nullptr, *hCorr = nullptr, *hVelm = nullptr;
    long long *hForce = nullptr;
    size_t stagedN = 0;
    // vvb200_step_host pipeline: copy-in / copy-out streams and per-chunk events
    cudaStream_t sIn = nullptr, sOut = nullptr;
    std::vector<cudaEvent_t> pipeEvents;
    std::vector<void *> allocations;
};

#define CUDA_TRY(expr)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            vvb200_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return VVB200_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)

template <class T>
static int uploadVec(vvb200_device_state *d, T **dst, const void *src, size_t count, cudaStream_t st) {
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CUDA_TRY(cudaMalloc((void **) dst, bytes));
    d->allocations.push_back(*dst);
    CUDA_TRY(cudaMemsetAsync(*dst, 0, bytes, st));
    if (count && src)
        CUDA_TRY(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, st));
    return VVB200_OK;
}

static size_t mixedSize(int precision) { return precision == VVB200_SINGLE ? 4 : 8; }
static size_t realSize(int precision) { return precision == VVB200_DOUBLE ? 8 : 4; }

extern "C" int vvb200_has_device_code(void) { return 1; }

void vvb200_device_free(vvb200_plan *plan) {
    vvb200_device_state *d = plan->dev;
    if (!d)
        return;
    for (void *ptr : d->allocations)
        cudaFree(ptr);
    for (cudaEvent_t e : d->profEvents)
        cudaEventDestroy(e);
    for (cudaEvent_t e : d->pipeEvents)
        cudaEventDestroy(e);
    if (d->sIn) cudaStreamDestroy(d->sIn);
    if (d->sOut) cudaStreamDestroy(d->sOut);
    for (void *m : d->peerMapped)
        cudaIpcCloseMemHandle(m);
    if (d->peerTimedOutHost) cudaFreeHost(d->peerTimedOutHost);
    delete d;
    plan->dev = nullptr;
}

static void fillNhcHost(const vvb200_plan *p, NhcDevice &h) {
    memset(&h, 0, sizeof h);
    const int nc = p->par.num_nh_chains;
    h.numTG = p->numTempGroup;
    h.nc = nc;
    h.loops = p->par.loops_per_step;
    const double realKbT = BOLTZ_D * p->par.temperature, drudeKbT = BOLTZ_D * p->par.drude_temperature;
    for (int g = 0; g < 3; g++) {
        h.vscale[g] = 1.0;
        h.tTarget[g] = g == VVB200_TG_DRUDE ? p->par.drude_temperature : p->par.temperature;
        if (g >= p->numTempGroup)
            continue;
        // same expressions as CudaVVKernels.cpp:583-594, on the (possibly global) DOFs
        const double kT = g == VVB200_TG_DRUDE ? drudeKbT : realKbT;
        const double q = g == VVB200_TG_DRUDE ? drudeKbT / std::pow(p->par.drude_frequency, 2)
                                              : realKbT / std::pow(p->par.frequency, 2);
        h.NkbT[g] = p->dofGlobal[g] * kT;
        h.etaMass[g][0] = p->dofGlobal[g] * q;
        for (int k = 1; k < nc; k++)
            h.etaMass[g][k] = q;
    }
    h.invMassTotal = 1.0 / p->totalMassGlobal;
}

static int envInt(const char *name, int dflt);

extern "C" int vvb200_plan_upload(vvb200_plan *p, void *stream) {
    if (!p) {
        vvb200_set_error("vvb200_plan_upload: null plan");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    vvb200_device_free(p);
    vvb200_device_state *d = new vvb200_device_state();
    p->dev = d;
    CUDA_TRY(cudaGetDevice(&d->device));
    CUDA_TRY(cudaDeviceGetAttribute(&d->numSM, cudaDevAttrMultiProcessorCount, d->device));
    if (d->numSM != p->tileSM && !vvb200_build_tiles(p, d->numSM)) {     // tile sizes follow the device actually used
        vvb200_set_error("vvb200_plan_upload: %s", p->tiledWhyNot.c_str());
        return VVB200_ERR_UNSUPPORTED_TOPOLOGY;
    }
    d->numTiles = (int) p->tileStart.size() - 1;
    d->pdl = envInt("VVB200_PDL", 1) ? 1 : 0;
    d->handOver = envInt("VVB200_HANDOVER", 1) ? 1 : 0;
    d->tileReverse = envInt("VVB200_B_REVERSE", 1) ? 1 : 0;

    // per tile-local molecule: first slot, count, contiguity
    std::vector<int32_t> molInfo(p->tileMolList.size(), 0);
    {
        std::vector<int32_t> first(p->M, -1), last(p->M, -1), cnt(p->M, 0);
        for (int t = 0; t < d->numTiles; t++) {
            const int a = p->tileStart[t], b = p->tileStart[t + 1];
            for (int i = a; i < b; i++) {
                const uint32_t lm = p->slotMeta[i] & VVB200_META_MOL_MASK;
                if (lm == VVB200_META_MOL_NONE) continue;
                const int m = p->particleMolId[i];
                if (first[m] < 0) first[m] = i - a;
                last[m] = i - a;
                cnt[m]++;
            }
            for (int k = p->tileMolOffset[t]; k < p->tileMolOffset[t + 1]; k++) {
                const int m = p->tileMolList[k];
                uint32_t w = (uint32_t) first[m] | ((uint32_t) cnt[m] << 11);
                if (last[m] - first[m] + 1 != cnt[m]) w |= 1u << 31;
                if (!p->tileMolFrag.empty() && p->tileMolFrag[k] >= 0) w |= 1u << 30;
                molInfo[k] = (int32_t) w;
                first[m] = last[m] = -1;
                cnt[m] = 0;
            }
        }
    }
    // tile descriptors: (t0, t1, m0, nMol), (molFirst or -1, 0, 0, 0)
    std::vector<int32_t> desc((size_t) d->numTiles * 8, 0);
    for (int t = 0; t < d->numTiles; t++) {
        const int m0 = p->tileMolOffset[t], nMol = p->tileMolOffset[t + 1] - m0;
        int molFirst = nMol > 0 ? p->tileMolList[m0] : 0;
        for (int j = 0; j < nMol; j++)
            if (p->tileMolList[m0 + j] != molFirst + j) molFirst = -1;
        int32_t *e = desc.data() + (size_t) t * 8;
        e[0] = p->tileStart[t]; e[1] = p->tileStart[t + 1]; e[2] = m0; e[3] = nMol; e[4] = molFirst;
    }
    molInfo.resize(molInfo.size() + 8, 0);                       // bulk copies read rounded-up ranges
    std::vector<uint32_t> metaPadded(p->slotMeta);
    metaPadded.resize((size_t) p->paddedN + 8, VVB200_META_MOL_NONE);
    int rc;
    if ((rc = uploadVec(d, &d->tileDesc, desc.data(), (size_t) d->numTiles * 2, st))) return rc;
    if ((rc = uploadVec(d, &d->tileMolList, p->tileMolList.data(), p->tileMolList.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->tileMolInfo, molInfo.data(), molInfo.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->slotMeta, metaPadded.data(), metaPadded.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->sortedByMol, p->sortedByMol.data(), p->sortedByMol.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->particlesInMolecules, p->particlesInMolecules.data(), p->particlesInMolecules.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->imageOf, p->imageOf.data(), p->imageOf.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->tileMolFrag, p->tileMolFrag.data(), p->tileMolFrag.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->splitMolId, p->splitMolId.data(), p->splitMolId.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->splitFragOffset, p->splitFragOffset.data(), p->splitFragOffset.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->splitFragList, p->splitFragList.data(), p->splitFragList.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->fragPartials, nullptr, p->splitFragList.size() * 5, st))) return rc;
    if ((rc = uploadVec(d, &d->ldSlot, p->ldSlot.data(), p->ldSlot.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->normalLD, p->normalLD.data(), p->normalLD.size(), st))) return rc;
    if ((rc = uploadVec(d, &d->pairsLD, p->pairsLD.data(), p->pairsLD.size() / 2, st))) return rc;
    if ((rc = uploadVec(d, &d->imagePairs, p->imagePairs.data(), p->imagePairs.size() / 2, st))) return rc;
    if ((rc = uploadVec(d, &d->drudePairs, p->drudePairs.data(), p->drudePairs.size() / 2, st))) return rc;
    if (!p->tiled) {
        if ((rc = uploadVec(d, &d->moleculesNH, p->moleculesNH.data(), p->moleculesNH.size(), st))) return rc;
        if ((rc = uploadVec(d, &d->normalNH, p->normalNH.data(), p->normalNH.size(), st))) return rc;
        if ((rc = uploadVec(d, &d->pairsNH, p->pairsNH.data(), p->pairsNH.size() / 2, st))) return rc;
        if ((rc = uploadVec(d, &d->particleMolId, p->particleMolId.data(), p->particleMolId.size(), st))) return rc;
    }

    const size_t ms = mixedSize(p->precision), rs = realSize(p->precision);
    unsigned char *raw = nullptr;
    if ((rc = uploadVec(d, &raw, nullptr, (size_t) p->M * 4 * ms, st))) return rc;   // zero-initialised, :606-617
    d->comV = raw;
    if ((rc = uploadVec(d, &raw, nullptr, ((size_t) p->M + 8) * ms, st))) return rc;
    d->comCbar = raw;
    const size_t nLDslots = p->normalLD.size() + p->pairsLD.size();
    if ((rc = uploadVec(d, &raw, nullptr, std::max<size_t>(nLDslots, 1) * 3 * rs, st))) return rc;
    d->ldForce = raw;

    // per-block partial sums: sized for the largest persistent grid any instantiation may use
    if ((rc = uploadVec(d, &d->partials, nullptr, (size_t) d->numSM * VVB200_MAX_BLOCKS_PER_SM * VVB200_NRED, st))) return rc;
    if ((rc = uploadVec(d, &d->counter, nullptr, 192, st))) return rc;      // four words and the factor records, each on a 128-byte line of its own
    NhcDevice h;
    fillNhcHost(p, h);
    if ((rc = uploadVec(d, &d->nhc, &h, 1, st))) return rc;
    CUDA_TRY(cudaStreamSynchronize(st));
    return VVB200_OK;
}

extern "C" int vvb200_set_global_thermostat(vvb200_plan *p, const double *dof3, double totalMass) {
    if (!p || !dof3 || !(totalMass > 0)) {
        vvb200_set_error("vvb200_set_global_thermostat: invalid argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    for (int g = 0; g < 3; g++) p->dofGlobal[g] = dof3[g];
    p->totalMassGlobal = totalMass;
    p->partitioned = true;
    // temperature-group count follows the global DOFs (CudaVVKernels.cpp:567-573)
    p->numTempGroup = 3;
    if (p->dofGlobal[VVB200_TG_DRUDE] == 0) {
        p->numTempGroup = 2;
        if (p->dofGlobal[VVB200_TG_COM] == 0) p->numTempGroup = 1;
    }
    if (p->dev) {
        NhcDevice h;
        fillNhcHost(p, h);
        CUDA_TRY(cudaMemcpy(p->dev->nhc, &h, sizeof h, cudaMemcpyHostToDevice));
    }
    return VVB200_OK;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static int checkStepArgs(const vvb200_plan *p, const vvb200_buffers *b, const char *who, bool needPos, bool needForce) {
    if (!p || !b) {
        vvb200_set_error("%s: null argument", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (!p->dev) {
        vvb200_set_error("%s: plan not uploaded (call vvb200_plan_upload)", who);
        return VVB200_ERR_NOT_UPLOADED;
    }
    if (p->dev->peerTimedOutHost && *(volatile unsigned int *) p->dev->peerTimedOutHost) {
        vvb200_set_error("%s: a multi-GPU peer exchange timed out in an earlier step (a rank died or stalled for longer than "
                         "VVB200_PEER_TIMEOUT_S); velocities since then are NaN", who);
        return VVB200_ERR_CUDA;
    }
    if (!b->velm || (needPos && !b->posq) || (needForce && !b->force) ||
        (needPos && p->precision == VVB200_MIXED && !b->posq_correction)) {
        vvb200_set_error("%s: missing device buffer", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (!p->particlesLD.empty() && needForce && !b->random) {
        vvb200_set_error("%s: Langevin particles present but no random buffer", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if ((p->par.cos_acceleration != 0 || !p->particlesElectrolyte.empty()) && !b->posq) {
        vvb200_set_error("%s: posq required", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    return VVB200_OK;
}

static int envInt(const char *name, int dflt);

static KParams makeParams(const vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a) {
    const vvb200_device_state *d = p->dev;
    KParams k;
    memset(&k, 0, sizeof k);
    k.N = p->N; k.paddedN = p->paddedN; k.numTiles = d->numTiles;
    k.tileBegin = 0; k.tileEnd = d->numTiles;
    k.pdl = d->pdl;
    k.peerOn = 0;                 // set by the multi-GPU step calls when the peer exchange is attached
    k.peer = d->peer;
    k.tileDesc = d->tileDesc; k.tileMolList = d->tileMolList;
    k.tileMolInfo = d->tileMolInfo; k.slotMeta = d->slotMeta; k.ldSlot = d->ldSlot;
    k.sortedByMol = d->sortedByMol; k.particlesInMolecules = d->particlesInMolecules;
    k.tileMolFrag = d->tileMolFrag; k.splitMolId = d->splitMolId; k.splitFragOffset = d->splitFragOffset;
    k.splitFragList = d->splitFragList; k.fragPartials = d->fragPartials; k.numSplit = (int) p->splitMolId.size();
    k.posq = b->posq; k.corr = b->posq_correction; k.velm = b->velm; k.force = b->force;
    k.ldForce = d->ldForce; k.comV = d->comV; k.comCbar = d->comCbar;
    k.partials = d->partials; k.nhc = d->nhc; k.counter = d->counter; k.gridGen = d->counter + 32; k.syncFlag = d->counter + 64; k.tileTicket = d->counter + 96; k.tileReverse = d->tileReverse;
    k.factorRec = reinterpret_cast<FactorRec *>(d->counter + 128);
    k.dt = p->par.step_size;
    k.efscale = p->par.electric_field * AVOGADRO_D;          // CudaVVKernels.cpp:978
    k.accel = p->par.cos_acceleration;                       // :1044
    k.invBoxZ = a ? a->inv_box_z : 0.0;
    k.maxDrudeDistance = p->par.max_drude_distance;          // :189
    k.hardwallScale = std::sqrt(BOLTZ_D * p->par.drude_temperature);   // :190
    k.maxD2safeD = p->par.max_drude_distance * p->par.max_drude_distance * (1.0 - 1e-4);
    k.maxD2safeF = (float) k.maxD2safeD;
    k.useCOM = p->par.use_com_temp_group != 0 && !p->moleculesNH.empty();
    k.hasLD = !p->particlesLD.empty();
    k.hasField = !p->particlesElectrolyte.empty();
    k.hardwall = p->par.max_drude_distance > 0 && !p->drudePairs.empty();
    k.extraForces = 1;
    k.fuseNHC = 1;
    k.cosine = p->par.cos_acceleration != 0;
    k.kickOnly = p->tiled ? 0 : 1;
    static const int imageFusedEnv = envInt("VVB200_IMAGE_FUSED", 1);
    k.imageFused = imageFusedEnv && p->tiled && !p->imageOf.empty();
    k.imageOf = d->imageOf;
    k.mirror = p->par.mirror_location;
    // Langevin force inside the kick: middle scheme on the tiled path (pairs are never cut there)
    static const int ldInlineEnv = envInt("VVB200_LD_INLINE", 1);
    k.ldInline = ldInlineEnv && p->tiled && p->par.use_middle_scheme && !p->particlesLD.empty() && b->random != nullptr;
    k.random = (const float4 *) b->random;
    k.randomIndex = a ? a->random_index : 0;
    k.nNormalLD = (int) p->normalLD.size();
    k.ldDrag = p->par.friction;                                                           // CudaVVKernels.cpp:835-839
    k.ldDragDrude = p->par.drude_friction;
    k.ldRand = std::sqrt(2.0 * BOLTZ_D * p->par.temperature * p->par.friction / p->par.step_size);
    k.ldRandDrude = std::sqrt(2.0 * BOLTZ_D * p->par.drude_temperature * p->par.drude_friction / p->par.step_size);
    return k;
}

// Persistent launch geometry: blocksPerSM co-resident blocks per SM (148 SMs), each with a ring of `stages`
// shared-memory stages, striding over the molecule-aligned tiles.  Defaults: MINBLOCKS_A / MINBLOCKS_B blocks per SM
// (what the kernels' __launch_bounds__ were compiled for) and as many stages as then fit the 227 KB of shared
// memory; VVB200_STAGES_A/B and VVB200_BLOCKS_A/B override (tuning).
static int envInt(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

struct LaunchCfg { int stages, perSM; size_t smem; };

template <class F>
static LaunchCfg configure(F kernel, size_t (*smemBytes)(int), const char *stagesEnv, const char *blocksEnv, int dfltBlocks,
                           int blockThreads, int maxStages) {
    int dev = 0, maxOptin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    LaunchCfg c;
    c.perSM = std::max(1, std::min(envInt(blocksEnv, dfltBlocks), VVB200_MAX_BLOCKS_PER_SM));
    const size_t perSMBudget = 227 * 1024;
    int stages = envInt(stagesEnv, 0);
    if (stages <= 0) {
        stages = 1;
        while (stages < maxStages && (smemBytes(stages + 1) + 1024) * c.perSM <= perSMBudget) stages++;
    }
    while (stages > 1 && smemBytes(stages) > (size_t) maxOptin) stages--;
    c.stages = stages;
    c.smem = smemBytes(stages);
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) c.smem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, blockThreads, c.smem);
    c.perSM = std::max(1, std::min(c.perSM, occ));
    return c;
}

// The two streaming kernels are launched with programmatic stream serialization (PDL): each may be scheduled while
// its predecessor drains (last-block reduction, NH chains, store tail) and waits at its own gridDepWait() before
// touching data -- the ~2-3 us of launch latency and prologue per kernel leave the critical path (VVB200_PDL=0: off).
template <class... Args>
static cudaError_t launchStreaming(void (*kernel)(Args...), int grid, int blockThreads, size_t smem, cudaStream_t st, const KParams &k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(blockThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = k.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, k);
}

// Launch configurations are cached per kernel instantiation AND per device: cudaFuncSetAttribute / occupancy belong to
// the device that is current, and one process may hold OpenMM Contexts on several GPUs.
#define VVB200_MAX_DEVICES 64
struct CfgCache {
    LaunchCfg cfg[VVB200_MAX_DEVICES];
    bool valid[VVB200_MAX_DEVICES] = {};
};
static int currentDeviceSlot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < VVB200_MAX_DEVICES ? dev : 0;
}

template <int MODE, int KICK, bool EXTRA>
static cudaError_t launchA(KParams k, int numSM, cudaStream_t st) {
    static CfgCache cache;
    const int slot = currentDeviceSlot();
    if (!cache.valid[slot]) {
        cache.cfg[slot] = configure(kick_reduce_kernel<MODE, KICK, EXTRA>, smemBytesA<MODE, EXTRA, KICK != KICK_NONE>, "VVB200_STAGES_A", "VVB200_BLOCKS_A", passABlocks(KICK), BTHREADS, 8);
        cache.valid[slot] = true;
    }
    const LaunchCfg &cfg = cache.cfg[slot];
    k.stagesA = cfg.stages;
    const int grid = std::max(1, std::min(k.tileEnd - k.tileBegin, numSM * cfg.perSM));
    return launchStreaming(kick_reduce_kernel<MODE, KICK, EXTRA>, grid, BTHREADS, cfg.smem, st, k);
}

// the reduce-only pass without the cosine perturbation: its own latency-organised kernel (vvb200_stream.cuh)
template <int MODE>
static cudaError_t launchRed(KParams k, int numSM, cudaStream_t st) {
    static CfgCache cache;
    const int slot = currentDeviceSlot();
    if (!cache.valid[slot]) {
        cache.cfg[slot] = configure(reduce_kernel<MODE>, smemBytesRed<MODE>, "VVB200_STAGES_R", "VVB200_BLOCKS_R", MINBLOCKS_RED, BTHREADS, 3);
        cache.valid[slot] = true;
    }
    const LaunchCfg &cfg = cache.cfg[slot];
    k.stagesA = cfg.stages;
    const int grid = std::max(1, std::min(k.tileEnd - k.tileBegin, numSM * cfg.perSM));
    return launchStreaming(reduce_kernel<MODE>, grid, BTHREADS, cfg.smem, st, k);
}

template <int MODE, int VARIANT, bool EXTRA>
static cudaError_t launchB(KParams k, int numSM, cudaStream_t st) {
    // the scale-only variant stages 36 B/particle instead of 68: it needs a deeper ring for the same bytes in flight
    constexpr int maxStages = VARIANT == VAR_SCALE_ONLY ? 8 : MAXSTAGES_B;
    constexpr int blockThreads = passBConsumers(VARIANT) + 32;
    static CfgCache cache;
    const int slot = currentDeviceSlot();
    if (!cache.valid[slot]) {
        cache.cfg[slot] = configure(scale_drift_kernel<MODE, VARIANT, EXTRA>, smemBytesB<MODE, VARIANT, EXTRA>, "VVB200_STAGES_B", "VVB200_BLOCKS_B", passBMinBlocks(VARIANT), blockThreads, maxStages);
        cache.valid[slot] = true;
    }
    const LaunchCfg &cfg = cache.cfg[slot];
    k.stagesB = cfg.stages;
    const int grid = std::max(1, std::min(k.tileEnd - k.tileBegin, numSM * cfg.perSM));
    return launchStreaming(scale_drift_kernel<MODE, VARIANT, EXTRA>, grid, blockThreads, cfg.smem, st, k);
}

// Pass B of a call that launches both passes itself takes over from pass A through the hand-over word instead of waiting
// for the whole grid (vvb200_stream.cuh, "early hand-over"); VVB200_HANDOVER=0 restores griddepcontrol.wait (tests).
static int handOverFor(const vvb200_plan *p, const KParams &k) {
    return p->dev->handOver && k.fuseNHC && !k.kickOnly ? 1 : 0;
}

// EXTRA kernels stage posq as well: needed by the external field (charge), the cosine acceleration (z) and, for
// uniformity, Langevin systems (which are never large)
static bool needsExtra(const KParams &k) { return k.cosine || k.hasField || k.hasLD; }

template <int KICK>
static cudaError_t dispatchA(int precision, bool, const KParams &k, int numSM, cudaStream_t st) {
    static const int dedicatedReduce = envInt("VVB200_REDUCE_KERNEL", 1);
    if (KICK == KICK_NONE && !k.cosine && !k.kickOnly && dedicatedReduce) {
        switch (precision) {
        case VVB200_SINGLE: return launchRed<VVB200_SINGLE>(k, numSM, st);
        case VVB200_MIXED: return launchRed<VVB200_MIXED>(k, numSM, st);
        default: return launchRed<VVB200_DOUBLE>(k, numSM, st);
        }
    }
    switch (precision * 2 + (needsExtra(k) ? 1 : 0)) {
    case 0: return launchA<VVB200_SINGLE, KICK, false>(k, numSM, st);
    case 1: return launchA<VVB200_SINGLE, KICK, true>(k, numSM, st);
    case 2: return launchA<VVB200_MIXED, KICK, false>(k, numSM, st);
    case 3: return launchA<VVB200_MIXED, KICK, true>(k, numSM, st);
    case 4: return launchA<VVB200_DOUBLE, KICK, false>(k, numSM, st);
    default: return launchA<VVB200_DOUBLE, KICK, true>(k, numSM, st);
    }
}

template <int VARIANT>
static cudaError_t dispatchB(int precision, bool, const KParams &k, int numSM, cudaStream_t st) {
    switch (precision * 2 + (needsExtra(k) ? 1 : 0)) {
    case 0: return launchB<VVB200_SINGLE, VARIANT, false>(k, numSM, st);
    case 1: return launchB<VVB200_SINGLE, VARIANT, true>(k, numSM, st);
    case 2: return launchB<VVB200_MIXED, VARIANT, false>(k, numSM, st);
    case 3: return launchB<VVB200_MIXED, VARIANT, true>(k, numSM, st);
    case 4: return launchB<VVB200_DOUBLE, VARIANT, false>(k, numSM, st);
    default: return launchB<VVB200_DOUBLE, VARIANT, true>(k, numSM, st);
    }
}

// ---- single-launch resident step (vvb200_resident.cuh) ------------------------------------------------------------
// Returns 1 when launched, 0 when the system does not fit on chip (the caller falls back to the streaming kernels),
// -1 on a CUDA error.  Geometry: one tile per block while all blocks are co-resident (2 per SM); beyond that several
// tiles per block with one block per SM.  The grid never exceeds occupancy x SMs: the kernel has a grid barrier.
template <int MODE, int KICK, int VARIANT, bool EXTRA>
static int launchResident(KParams k, int numSM, cudaStream_t st) {
    constexpr int MAXT = 8;
    auto kernel = resident_step_kernel<MODE, KICK, VARIANT, EXTRA>;
    // per device, like the streaming kernels' configurations
    struct ResidentCache { int maxOptin = -1; int occ[MAXT + 1]; };
    static ResidentCache caches[VVB200_MAX_DEVICES];
    ResidentCache &rc = caches[currentDeviceSlot()];
    if (rc.maxOptin < 0) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kernel) == cudaSuccess)
            v -= (int) ((fa.sharedSizeBytes + 127) / 128 * 128);   // static shared memory counts against the same limit
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v) != cudaSuccess) {
            cudaGetLastError();
            v = 48 * 1024;
        }
        rc.maxOptin = v;
        for (int t = 0; t <= MAXT; t++) rc.occ[t] = -1;
    }
    const int maxOptin = rc.maxOptin;
    int *occCache = rc.occ;
    auto smem = [](int T) { return smemBytesR<MODE, KICK, VARIANT, EXTRA>(T); };
    auto occ = [&](int T) {
        if (occCache[T] < 0) {
            int o = 0;
            if (smem(T) <= (size_t) maxOptin)
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kernel, CTHREADS, smem(T));
            occCache[T] = o;
        }
        return occCache[T];
    };
    if (k.numTiles < 1)
        return 0;
    int T = 1, grid = k.numTiles;
    if (k.numTiles > occ(1) * numSM) {
        T = (k.numTiles + numSM - 1) / numSM;
        if (T > MAXT || occ(T) < 1)
            return 0;
        grid = (k.numTiles + T - 1) / T;
    }
    k.tilesPerBlock = T;
    // Cooperative launch: the runtime places the whole grid or nothing, so two contexts stepping on different streams
    // of one GPU can never hold half of each other's blocks while both wait at their grid barriers.
    static const int coop = envInt("VVB200_COOP", 1);      // an environment switch, not device state
    if (coop) {
        void *args[] = {&k};
        return cudaLaunchCooperativeKernel((const void *) kernel, dim3(grid), dim3(CTHREADS), args, smem(T), st) == cudaSuccess ? 1 : -1;
    }
    kernel<<<grid, CTHREADS, smem(T), st>>>(k);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

template <int KICK, int VARIANT>
static int dispatchResident(int precision, const KParams &k, int numSM, cudaStream_t st) {
    switch (precision * 2 + (needsExtra(k) ? 1 : 0)) {
    case 0: return launchResident<VVB200_SINGLE, KICK, VARIANT, false>(k, numSM, st);
    case 1: return launchResident<VVB200_SINGLE, KICK, VARIANT, true>(k, numSM, st);
    case 2: return launchResident<VVB200_MIXED, KICK, VARIANT, false>(k, numSM, st);
    case 3: return launchResident<VVB200_MIXED, KICK, VARIANT, true>(k, numSM, st);
    case 4: return launchResident<VVB200_DOUBLE, KICK, VARIANT, false>(k, numSM, st);
    default: return launchResident<VVB200_DOUBLE, KICK, VARIANT, true>(k, numSM, st);
    }
}

static bool residentEnabled(const vvb200_plan *p) {
    vvb200_device_state *d = p->dev;
    if (d->residentMode < 0)
        d->residentMode = envInt("VVB200_RESIDENT", 1) ? 1 : 0;
    // measured on B200 (tests/diag_small.py): one launch wins up to ~100k particles; beyond that the streaming
    // kernels' deeper load pipeline does (148k: 18.4 vs 21.5 us per step)
    if (d->residentMaxParticles < 0)
        d->residentMaxParticles = envInt("VVB200_RESIDENT_MAX_PARTICLES", 120000);
    return d->residentMode == 1 && p->tiled && p->N <= d->residentMaxParticles;
}

// Tries the resident kernel for the pass-A<KICK> + pass-B<VARIANT> pair; *launched tells whether it ran.
template <int KICK, int VARIANT>
static int tryResident(vvb200_plan *p, KParams k, bool reduce, cudaStream_t st, bool *launched) {
    *launched = false;
    if (!residentEnabled(p))
        return VVB200_OK;
    k.doReduce = reduce ? 1 : 0;
    k.fuseNHC = reduce ? 1 : 0;
    const int r = dispatchResident<KICK, VARIANT>(p->precision, k, p->dev->numSM, st);
    if (r < 0) {
        vvb200_set_error("resident step kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        return VVB200_ERR_CUDA;
    }
    if (r == 1) {
        *launched = true;
        p->launches++;
        p->dev->residentLaunches++;
    }
    return VVB200_OK;
}

static int launchLangevin(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, cudaStream_t st) {
    vvb200_device_state *d = p->dev;
    if (makeParams(p, b, a).ldInline)
        return VVB200_OK;        // middle scheme on the tiled path: evaluated inside the kick (langevinForceInline)
    const int nNormal = (int) p->normalLD.size(), nPairs = (int) p->pairsLD.size() / 2;
    const int work = std::max(nNormal, nPairs);
    if (work == 0)
        return VVB200_OK;
    // CudaVVKernels.cpp:835-839
    const double dt = p->par.step_size;
    const double drag = p->par.friction, dragDrude = p->par.drude_friction;
    const double randF = std::sqrt(2.0 * BOLTZ_D * p->par.temperature * drag / dt);
    const double randFD = std::sqrt(2.0 * BOLTZ_D * p->par.drude_temperature * dragDrude / dt);
    const int grid = std::min((work + 127) / 128, d->numSM * 8);
    const unsigned int ri = a ? a->random_index : 0;
    switch (p->precision) {
    case VVB200_SINGLE:
        langevin_force_kernel<VVB200_SINGLE><<<grid, 128, 0, st>>>(b->velm, d->ldForce, d->normalLD, nNormal, d->pairsLD, nPairs,
                                                                  drag, randF, dragDrude, randFD, (const float4 *) b->random, ri);
        break;
    case VVB200_MIXED:
        langevin_force_kernel<VVB200_MIXED><<<grid, 128, 0, st>>>(b->velm, d->ldForce, d->normalLD, nNormal, d->pairsLD, nPairs,
                                                                 drag, randF, dragDrude, randFD, (const float4 *) b->random, ri);
        break;
    default:
        langevin_force_kernel<VVB200_DOUBLE><<<grid, 128, 0, st>>>(b->velm, d->ldForce, d->normalLD, nNormal, d->pairsLD, nPairs,
                                                                  drag, randF, dragDrude, randFD, (const float4 *) b->random, ri);
    }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return VVB200_OK;
}

extern "C" int vvb200_update_image_positions(vvb200_plan *p, const vvb200_buffers *b, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_update_image_positions", true, false);
    if (rc) return rc;
    const int n = (int) p->imagePairs.size() / 2;
    if (n == 0)
        return VVB200_OK;
    cudaStream_t st = (cudaStream_t) stream;
    const int grid = std::min((n + 127) / 128, p->dev->numSM * 8);
    switch (p->precision) {
    case VVB200_SINGLE: image_kernel<VVB200_SINGLE><<<grid, 128, 0, st>>>(b->posq, b->posq_correction, p->dev->imagePairs, n, p->par.mirror_location); break;
    case VVB200_MIXED: image_kernel<VVB200_MIXED><<<grid, 128, 0, st>>>(b->posq, b->posq_correction, p->dev->imagePairs, n, p->par.mirror_location); break;
    default: image_kernel<VVB200_DOUBLE><<<grid, 128, 0, st>>>(b->posq, b->posq_correction, p->dev->imagePairs, n, p->par.mirror_location);
    }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return VVB200_OK;
}

static bool hasNH(const vvb200_plan *p) { return !p->particlesNH.empty(); }

// per-kernel timing: which = 0 before pass A, 1 after it, 2 before pass B, 3 after it
static void profMark(vvb200_device_state *d, int which, cudaStream_t st) {
    if (!d->profiling)
        return;
    if (which == 0) {
        if (d->profUsed + 4 > d->profEvents.size())
            return;   // all slots used: later steps are not sampled
    } else if (d->profUsed % 4 != (size_t) which) {
        return;
    }
    cudaEventRecord(d->profEvents[d->profUsed++], st);
}

extern "C" int vvb200_profile_enable(vvb200_plan *p, int maxSteps) {
    if (!p || !p->dev || maxSteps < 0) {
        vvb200_set_error("vvb200_profile_enable: plan not uploaded or invalid argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    vvb200_device_state *d = p->dev;
    while (d->profEvents.size() < (size_t) maxSteps * 4) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        d->profEvents.push_back(e);
    }
    d->profUsed = 0;
    d->profiling = maxSteps > 0;
    return VVB200_OK;
}

extern "C" int vvb200_profile_read(vvb200_plan *p, double *msPassA, double *msPassB, int32_t *steps) {
    if (!p || !p->dev || !msPassA || !msPassB || !steps) {
        vvb200_set_error("vvb200_profile_read: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    vvb200_device_state *d = p->dev;
    const size_t n = d->profUsed / 4;
    double a = 0, b = 0;
    if (n)
        CUDA_TRY(cudaEventSynchronize(d->profEvents[4 * n - 1]));
    for (size_t i = 0; i < n; i++) {
        float ta = 0, tb = 0;
        CUDA_TRY(cudaEventElapsedTime(&ta, d->profEvents[4 * i], d->profEvents[4 * i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&tb, d->profEvents[4 * i + 2], d->profEvents[4 * i + 3]));
        a += ta;
        b += tb;
    }
    *msPassA = a;
    *msPassB = b;
    *steps = (int32_t) n;
    d->profUsed = 0;
    return VVB200_OK;
}


// ---- any-topology path (vvb200_general.cuh) --------------------------------------------------------------------
static GParams makeGParams(const vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a) {
    const vvb200_device_state *d = p->dev;
    GParams g;
    memset(&g, 0, sizeof g);
    g.N = p->N;
    g.nMolNH = (int) p->moleculesNH.size();
    g.nNormal = (int) p->normalNH.size();
    g.nPairs = (int) p->pairsNH.size() / 2;
    g.moleculesNH = d->moleculesNH; g.normalNH = d->normalNH; g.particleMolId = d->particleMolId;
    g.sortedByMol = d->sortedByMol; g.particlesInMolecules = d->particlesInMolecules; g.pairsNH = d->pairsNH;
    g.velm = b->velm; g.posq = b->posq; g.comV = d->comV; g.comCbar = d->comCbar;
    g.partials = d->partials; g.nhc = d->nhc; g.counter = d->counter;
    g.dt = p->par.step_size;
    g.invBoxZ = a ? a->inv_box_z : 0.0;
    g.useCOM = p->par.use_com_temp_group != 0 && !p->moleculesNH.empty();
    g.cosine = p->par.cos_acceleration != 0;
    g.fuseNHC = 1;
    return g;
}

template <int MODE>
static void launchGeneralThermostat(const GParams &g, int numSM, bool reduce, bool scale, cudaStream_t st) {
    const int work = std::max(std::max(g.nNormal, g.nPairs), std::max(g.nMolNH, g.cosine ? g.N : 0));
    const int grid = std::max(1, std::min((work + 255) / 256, numSM * 8));
    if (reduce) {
        if (g.useCOM) {
            const int gridCom = std::max(1, std::min((g.nMolNH + 7) / 8, numSM * 8));
            general_com_kernel<MODE><<<gridCom, 256, 0, st>>>(g);
        }
        general_ke_kernel<MODE><<<grid, 256, 0, st>>>(g);
    }
    if (scale)
        general_scale_kernel<MODE><<<grid, 256, 0, st>>>(g);
}

// reduce: COM velocities + group energies (+ NH chains when fuse); scale: in-place velocity scaling
static int generalThermostat(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, bool reduce, bool fuse,
                             bool scale, cudaStream_t st) {
    if (p->par.cos_acceleration != 0 && !b->posq) {
        vvb200_set_error("vvb200: posq required for the cosine perturbation");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    GParams g = makeGParams(p, b, a);
    g.fuseNHC = fuse ? 1 : 0;
    switch (p->precision) {
    case VVB200_SINGLE: launchGeneralThermostat<VVB200_SINGLE>(g, p->dev->numSM, reduce, scale, st); break;
    case VVB200_MIXED: launchGeneralThermostat<VVB200_MIXED>(g, p->dev->numSM, reduce, scale, st); break;
    default: launchGeneralThermostat<VVB200_DOUBLE>(g, p->dev->numSM, reduce, scale, st);
    }
    p->launches += (reduce ? (g.useCOM ? 2 : 1) : 0) + (scale ? 1 : 0);
    CUDA_TRY(cudaGetLastError());
    return VVB200_OK;
}

// the general path needs posDelta; standalone callers may not own one
static int withPosDelta(vvb200_plan *p, const vvb200_buffers *b, vvb200_buffers *out, cudaStream_t st) {
    *out = *b;
    if (out->pos_delta)
        return VVB200_OK;
    vvb200_device_state *d = p->dev;
    if (!d->ownPosDelta) {
        const size_t bytes = (size_t) p->paddedN * 4 * (p->precision == VVB200_SINGLE ? 4 : 8);
        CUDA_TRY(cudaMalloc(&d->ownPosDelta, bytes));
        d->allocations.push_back(d->ownPosDelta);
        CUDA_TRY(cudaMemsetAsync(d->ownPosDelta, 0, bytes, st));
    }
    out->pos_delta = d->ownPosDelta;
    return VVB200_OK;
}

extern "C" int vvb200_middle_delta(vvb200_plan *p, const vvb200_buffers *b, int accumulate, void *stream);
extern "C" int vvb200_middle_finish(vvb200_plan *p, const vvb200_buffers *b, void *stream);
extern "C" int vvb200_vv_positions(vvb200_plan *p, const vvb200_buffers *b, void *stream);

static int generalKick(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, int kick, bool extraForces,
                       cudaStream_t st) {
    KParams k = makeParams(p, b, a);
    k.fuseNHC = 0;
    k.extraForces = extraForces ? 1 : 0;
    if (kick == KICK_MIDDLE) CUDA_TRY((dispatchA<KICK_MIDDLE>(p->precision, k.cosine, k, p->dev->numSM, st)));
    else CUDA_TRY((dispatchA<KICK_VV>(p->precision, k.cosine, k, p->dev->numSM, st)));
    p->launches++;
    return VVB200_OK;
}

extern "C" int vvb200_middle_kick_reduce(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_middle_kick_reduce", false, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    if (!p->particlesLD.empty() && (rc = launchLangevin(p, b, a, st))) return rc;
    if (!p->tiled) {
        vvb200_buffers bb;
        if ((rc = withPosDelta(p, b, &bb, st))) return rc;
        if ((rc = generalKick(p, &bb, a, KICK_MIDDLE, true, st))) return rc;
        if ((rc = vvb200_middle_delta(p, &bb, 0, stream))) return rc;
        return hasNH(p) ? generalThermostat(p, &bb, a, true, false, false, st) : VVB200_OK;
    }
    KParams k = makeParams(p, b, a);
    // peers attached: the last block of this launch exchanges the sums over NVLink and advances the chains itself, and
    // vvb200_middle_nhc_scale_drift is pass B alone.  Otherwise the sums stay in vvb200_partials_ptr() for the caller's
    // all-reduce.
    k.peerOn = p->dev->peerAttached ? 1 : 0;
    k.fuseNHC = k.peerOn;
    profMark(p->dev, 0, st);
    CUDA_TRY((dispatchA<KICK_MIDDLE>(p->precision, p->par.cos_acceleration != 0, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 1, st);
    return VVB200_OK;
}

static int launchNhc(vvb200_plan *p, cudaStream_t st) {
    vvb200_device_state *d = p->dev;
    if (d->peerAttached) {
        // all-reduce over peer memory + NH chains in one single-block kernel (no NCCL launch)
        if (p->par.cos_acceleration != 0)
            nhc_peer_kernel<true><<<1, 32, 0, st>>>(d->nhc, d->peer, p->par.step_size);
        else
            nhc_peer_kernel<false><<<1, 32, 0, st>>>(d->nhc, d->peer, p->par.step_size);
        p->launches++;
        CUDA_TRY(cudaGetLastError());
        return VVB200_OK;
    }
    if (p->par.cos_acceleration != 0)
        nhc_kernel<true><<<1, 32, 0, st>>>(p->dev->nhc, p->par.step_size);
    else
        nhc_kernel<false><<<1, 32, 0, st>>>(p->dev->nhc, p->par.step_size);
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return VVB200_OK;
}

extern "C" int vvb200_middle_nhc_scale_drift(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_middle_nhc_scale_drift", true, false);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    // The exchange (when peers are attached) and the chains already ran inside pass A on the tiled path.  A partition
    // without thermostat particles still takes part in the exchange: its peers wait for its (zero) vector.
    const bool exchangeHere = !(p->tiled && p->dev->peerAttached);
    if (exchangeHere && (hasNH(p) || p->partitioned) && (rc = launchNhc(p, st))) return rc;
    if (!p->tiled) {
        vvb200_buffers bb;
        if ((rc = withPosDelta(p, b, &bb, st))) return rc;
        if (hasNH(p) && (rc = generalThermostat(p, &bb, a, false, false, true, st))) return rc;
        if ((rc = vvb200_middle_delta(p, &bb, 1, stream))) return rc;
        return vvb200_middle_finish(p, &bb, stream);      // position write + hard wall + images
    }
    KParams k = makeParams(p, b, a);
    // the pass-B interval of the split path starts here (the all-reduce and the NHC block are not in it)
    profMark(p->dev, 2, st);
    CUDA_TRY((dispatchB<VAR_MIDDLE>(p->precision, p->par.cos_acceleration != 0, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 3, st);
    return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);   // fused: pass B mirrored them
}

extern "C" int vvb200_peer_export(vvb200_plan *p, void *handleOut64) {
    if (!p || !p->dev || !handleOut64) {
        vvb200_set_error("vvb200_peer_export: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    vvb200_device_state *d = p->dev;
    if (!d->peerLocal) {
        CUDA_TRY(cudaMalloc((void **) &d->peerLocal, sizeof(PeerSlots)));
        d->allocations.push_back(d->peerLocal);
        CUDA_TRY(cudaMemset(d->peerLocal, 0, sizeof(PeerSlots)));
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, d->peerLocal));
    memcpy(handleOut64, &h, sizeof h);
    return VVB200_OK;
}

extern "C" int vvb200_peer_attach(vvb200_plan *p, int rank, int world, const void *handles) {
    if (!p || !p->dev || !handles || world < 1 || world > VVB200_MAX_RANKS || rank < 0 || rank >= world || !p->dev->peerLocal) {
        vvb200_set_error("vvb200_peer_attach: invalid argument (world <= %d, vvb200_peer_export first)", VVB200_MAX_RANKS);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    vvb200_device_state *d = p->dev;
    unsigned long long *keepSeq = d->peer.seq;
    memset(&d->peer, 0, sizeof d->peer);
    d->peer.seq = keepSeq;
    d->peer.rank = rank;
    d->peer.world = world;
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            d->peer.buf[r] = d->peerLocal;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + 64 * r, sizeof h);
        void *mapped = nullptr;
        CUDA_TRY(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
        d->peerMapped.push_back(mapped);
        d->peer.buf[r] = (PeerSlots *) mapped;
    }
    if (!d->peer.seq) {
        unsigned long long *seq = nullptr;
        CUDA_TRY(cudaMalloc((void **) &seq, sizeof *seq));
        d->allocations.push_back(seq);
        d->peer.seq = seq;
    }
    CUDA_TRY(cudaMemset(d->peer.seq, 0, sizeof(unsigned long long)));
    if (!d->peerTimedOutHost)
        CUDA_TRY(cudaHostAlloc((void **) &d->peerTimedOutHost, sizeof(unsigned int), cudaHostAllocMapped));
    *d->peerTimedOutHost = 0;
    void *flagDev = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer(&flagDev, d->peerTimedOutHost, 0));
    d->peer.timedOut = (volatile unsigned int *) flagDev;
    d->peer.timeoutNs = (unsigned long long) std::max(1, envInt("VVB200_PEER_TIMEOUT_S", 30)) * 1000000000ull;
    d->peerAttached = world > 1;
    return VVB200_OK;
}

extern "C" int vvb200_partials_ptr(vvb200_plan *p, void **ptr, int32_t *n) {
    if (!p || !p->dev || !ptr || !n) {
        vvb200_set_error("vvb200_partials_ptr: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    *ptr = p->dev->nhc->red;   // address arithmetic only; never dereferenced on the host
    *n = VVB200_NRED;
    return VVB200_OK;
}

extern "C" int vvb200_middle_nhc_scale_drift(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream);

extern "C" int vvb200_step_middle(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_step_middle", true, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    const bool cosine = p->par.cos_acceleration != 0;
    if (!p->tiled && p->dev->peerAttached) {      // any topology on several GPUs: the split calls carry the exchange
        if ((rc = vvb200_middle_kick_reduce(p, b, a, stream))) return rc;
        return vvb200_middle_nhc_scale_drift(p, b, a, stream);
    }
    if (!p->particlesLD.empty() && (rc = launchLangevin(p, b, a, st))) return rc;
    if (!p->tiled) {
        // any topology: kick | posDelta = dt/2 v | thermostat (gather kernels, NH chains on the device) |
        // posDelta += dt/2 v' | position write + hard wall | images
        vvb200_buffers bb;
        if ((rc = withPosDelta(p, b, &bb, st))) return rc;
        if ((rc = generalKick(p, &bb, a, KICK_MIDDLE, true, st))) return rc;
        if ((rc = vvb200_middle_delta(p, &bb, 0, stream))) return rc;
        if (hasNH(p) && (rc = generalThermostat(p, &bb, a, true, true, true, st))) return rc;
        if ((rc = vvb200_middle_delta(p, &bb, 1, stream))) return rc;
        return vvb200_middle_finish(p, &bb, stream);      // position write + hard wall + images
    }
    KParams k = makeParams(p, b, a);
    k.fuseNHC = hasNH(p);
    if (p->dev->peerAttached) {      // molecule-partitioned multi-GPU run: see vvb200_middle_kick_reduce
        k.peerOn = 1;
        k.fuseNHC = 1;
    }
    k.flagSync = handOverFor(p, k);
    profMark(p->dev, 0, st);
    bool resident = false;
    if (!k.peerOn && (rc = tryResident<KICK_MIDDLE, VAR_MIDDLE>(p, k, hasNH(p), st, &resident))) return rc;
    if (resident) {   // small system: the whole step ran as one launch; reported as "pass A", pass B = 0
        profMark(p->dev, 1, st);
        profMark(p->dev, 2, st);
        profMark(p->dev, 3, st);
        return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);   // fused: pass B mirrored them
    }
    CUDA_TRY((dispatchA<KICK_MIDDLE>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    CUDA_TRY((dispatchB<VAR_MIDDLE>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 3, st);
    return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);   // fused: pass B mirrored them
}

extern "C" int vvb200_step_vv_first(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_step_vv_first", true, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    const bool cosine = p->par.cos_acceleration != 0;
    if (!p->tiled) {
        vvb200_buffers bb;
        if ((rc = withPosDelta(p, b, &bb, st))) return rc;
        if (hasNH(p) && (rc = generalThermostat(p, &bb, a, true, true, true, st))) return rc;
        if ((rc = generalKick(p, &bb, a, KICK_VV, p->dev->extraForcesValid, st))) return rc;
        const int grid = std::max(1, std::min((p->N + THREADS - 1) / THREADS, p->dev->numSM * 8));
        switch (p->precision) {
        case VVB200_SINGLE: vv_delta_kernel<VVB200_SINGLE><<<grid, THREADS, 0, st>>>(bb.velm, bb.pos_delta, p->N, p->par.step_size); break;
        case VVB200_MIXED: vv_delta_kernel<VVB200_MIXED><<<grid, THREADS, 0, st>>>(bb.velm, bb.pos_delta, p->N, p->par.step_size); break;
        default: vv_delta_kernel<VVB200_DOUBLE><<<grid, THREADS, 0, st>>>(bb.velm, bb.pos_delta, p->N, p->par.step_size);
        }
        p->launches++;
        return vvb200_vv_positions(p, &bb, stream);       // position write + hard wall + images
    }
    KParams k = makeParams(p, b, a);
    k.extraForces = p->dev->extraForcesValid ? 1 : 0;
    bool resident = false;
    if ((rc = tryResident<KICK_NONE, VAR_VV_FIRST>(p, k, hasNH(p), st, &resident))) return rc;
    if (resident)
        return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);   // fused: pass B mirrored them
    k.fuseNHC = hasNH(p);
    k.flagSync = handOverFor(p, k);
    profMark(p->dev, 0, st);
    if (hasNH(p)) {
        CUDA_TRY((dispatchA<KICK_NONE>(p->precision, cosine, k, p->dev->numSM, st)));
        p->launches++;
    }
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    CUDA_TRY((dispatchB<VAR_VV_FIRST>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 3, st);
    return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);   // fused: pass B mirrored them
}

extern "C" int vvb200_step_vv_second(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_step_vv_second", true, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    const bool cosine = p->par.cos_acceleration != 0;
    if (!p->particlesLD.empty() && (rc = launchLangevin(p, b, a, st))) return rc;
    p->dev->extraForcesValid = true;
    if (!p->tiled) {
        if ((rc = generalKick(p, b, a, KICK_VV, true, st))) return rc;
        return hasNH(p) ? generalThermostat(p, b, a, true, true, true, st) : VVB200_OK;
    }
    KParams k = makeParams(p, b, a);
    k.fuseNHC = hasNH(p);
    if (hasNH(p)) {      // without a thermostat the second half is the kick alone: one streaming launch already
        bool resident = false;
        if ((rc = tryResident<KICK_VV, VAR_SCALE_ONLY>(p, k, true, st, &resident))) return rc;
        if (resident)
            return VVB200_OK;
    }
    k.flagSync = handOverFor(p, k);
    profMark(p->dev, 0, st);
    CUDA_TRY((dispatchA<KICK_VV>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    if (hasNH(p)) {
        CUDA_TRY((dispatchB<VAR_SCALE_ONLY>(p->precision, cosine, k, p->dev->numSM, st)));
        p->launches++;
    }
    profMark(p->dev, 3, st);
    return VVB200_OK;
}

// ---- the VVKernels.h interfaces one by one (constraint-bearing path) ----------------------------
extern "C" int vvb200_middle_kick(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    // extra forces + integrateMiddleVel; the reductions that ride along are discarded because OpenMM's
    // applyVelocityConstraints runs next (CudaVVKernels.cpp:144-151): the molecule / pair phases are skipped
    int rc = checkStepArgs(p, b, "vvb200_middle_kick", false, true);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t) stream;
    if (!p->particlesLD.empty() && (rc = launchLangevin(p, b, a, st))) return rc;
    KParams k = makeParams(p, b, a);
    k.fuseNHC = 0;
    k.kickOnly = 1;
    profMark(p->dev, 0, st);
    CUDA_TRY((dispatchA<KICK_MIDDLE>(p->precision, k.cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    profMark(p->dev, 3, st);
    return VVB200_OK;
}

extern "C" int vvb200_thermostat(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_thermostat", false, false);
    if (rc) return rc;
    if (!hasNH(p))
        return VVB200_OK;
    cudaStream_t st = (cudaStream_t) stream;
    const bool cosine = p->par.cos_acceleration != 0;
    if (!p->tiled)
        return generalThermostat(p, b, a, true, true, true, st);
    KParams k = makeParams(p, b, a);
    bool resident = false;
    if ((rc = tryResident<KICK_NONE, VAR_SCALE_ONLY>(p, k, true, st, &resident))) return rc;
    if (resident)
        return VVB200_OK;
    k.flagSync = handOverFor(p, k);
    profMark(p->dev, 0, st);
    CUDA_TRY((dispatchA<KICK_NONE>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    CUDA_TRY((dispatchB<VAR_SCALE_ONLY>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 3, st);
    return VVB200_OK;
}

extern "C" int vvb200_middle_delta(vvb200_plan *p, const vvb200_buffers *b, int accumulate, void *stream);

static int ensureOldDelta(vvb200_plan *p, cudaStream_t st) {
    vvb200_device_state *d = p->dev;
    if (!d->oldDelta) {
        const size_t bytes = (size_t) p->paddedN * 4 * mixedSize(p->precision);
        CUDA_TRY(cudaMalloc(&d->oldDelta, bytes));
        d->allocations.push_back(d->oldDelta);
        CUDA_TRY(cudaMemsetAsync(d->oldDelta, 0, bytes, st));
    }
    return VVB200_OK;
}

// scaleVelocity + integrateMiddlePos1 + integrateMiddlePos2 in one go for systems with OpenMM constraints
// (CudaVVKernels.cpp:154-173, 670-754): thermostat, then posDelta = oldDelta = dt/2 v + dt/2 v'
extern "C" int vvb200_middle_thermostat_delta(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_middle_thermostat_delta", false, false);
    if (rc) return rc;
    if (!b->pos_delta) {
        vvb200_set_error("vvb200_middle_thermostat_delta: pos_delta buffer required");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    if ((rc = ensureOldDelta(p, st))) return rc;
    if (!p->tiled) {
        if ((rc = vvb200_middle_delta(p, b, 0, stream))) return rc;
        if (hasNH(p) && (rc = generalThermostat(p, b, a, true, true, true, st))) return rc;
        return vvb200_middle_delta(p, b, 1, stream);
    }
    const bool cosine = p->par.cos_acceleration != 0;
    KParams k = makeParams(p, b, a);
    k.posDelta = b->pos_delta;
    k.oldDelta = p->dev->oldDelta;
    profMark(p->dev, 0, st);
    if (hasNH(p)) {
        bool resident = false;
        if ((rc = tryResident<KICK_NONE, VAR_SCALE_DELTA>(p, k, true, st, &resident))) return rc;
        if (resident) {
            profMark(p->dev, 1, st);
            profMark(p->dev, 2, st);
            profMark(p->dev, 3, st);
            return VVB200_OK;
        }
        k.flagSync = handOverFor(p, k);
        CUDA_TRY((dispatchA<KICK_NONE>(p->precision, cosine, k, p->dev->numSM, st)));
        p->launches++;
    }
    profMark(p->dev, 1, st);
    profMark(p->dev, 2, st);
    CUDA_TRY((dispatchB<VAR_SCALE_DELTA>(p->precision, cosine, k, p->dev->numSM, st)));
    p->launches++;
    profMark(p->dev, 3, st);
    return VVB200_OK;
}

static int elementwiseGrid(const vvb200_plan *p, int n) {
    return std::max(1, std::min((n + THREADS - 1) / THREADS, p->dev->numSM * 8));
}

extern "C" int vvb200_middle_delta(vvb200_plan *p, const vvb200_buffers *b, int accumulate, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_middle_delta", false, false);
    if (rc) return rc;
    if (!b->pos_delta) {
        vvb200_set_error("vvb200_middle_delta: pos_delta buffer required");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    vvb200_device_state *d = p->dev;
    if ((rc = ensureOldDelta(p, st))) return rc;
    const int grid = elementwiseGrid(p, p->N);
    switch (p->precision) {
    case VVB200_SINGLE: delta_kernel<VVB200_SINGLE><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, d->oldDelta, p->N, p->par.step_size, accumulate); break;
    case VVB200_MIXED: delta_kernel<VVB200_MIXED><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, d->oldDelta, p->N, p->par.step_size, accumulate); break;
    default: delta_kernel<VVB200_DOUBLE><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, d->oldDelta, p->N, p->par.step_size, accumulate);
    }
    p->launches++;
    CUDA_TRY(cudaGetLastError());
    return VVB200_OK;
}

extern "C" int vvb200_middle_finish(vvb200_plan *p, const vvb200_buffers *b, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_middle_finish", true, false);
    if (rc) return rc;
    vvb200_device_state *d = p->dev;
    if (!b->pos_delta || !d->oldDelta) {
        vvb200_set_error("vvb200_middle_finish: pos_delta missing or vvb200_middle_delta not called yet");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (p->tiled && envInt("VVB200_FUSED_FINISH", 1)) {
        // integrateMiddlePos3 + applyHardWallConstraints (+ the image mirror) in one streaming launch: pass B's tile
        // machinery with posDelta / oldDelta staged next to velm / posq
        KParams k = makeParams(p, b, nullptr);
        k.posDelta = b->pos_delta;
        k.oldDelta = d->oldDelta;
        profMark(d, 0, st);
        profMark(d, 1, st);
        profMark(d, 2, st);
        CUDA_TRY((dispatchB<VAR_FINISH>(p->precision, false, k, d->numSM, st)));
        p->launches++;
        profMark(d, 3, st);
        return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);
    }
    const int grid = elementwiseGrid(p, p->N);
    const int nPairs = (int) p->drudePairs.size() / 2;
    const bool hw = p->par.max_drude_distance > 0 && nPairs > 0;
    const int gridHW = std::max(1, std::min((nPairs + 127) / 128, d->numSM * 8));
    const double hwScale = std::sqrt(BOLTZ_D * p->par.drude_temperature);
    switch (p->precision) {
    case VVB200_SINGLE:
        finish_kernel<VVB200_SINGLE><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, d->oldDelta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_SINGLE><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
        break;
    case VVB200_MIXED:
        finish_kernel<VVB200_MIXED><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, d->oldDelta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_MIXED><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
        break;
    default:
        finish_kernel<VVB200_DOUBLE><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, d->oldDelta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_DOUBLE><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
    }
    p->launches += hw ? 2 : 1;
    CUDA_TRY(cudaGetLastError());
    return vvb200_update_image_positions(p, b, stream);      // images follow their parents in every variant of this call
}

// velocity-Verlet scheme around OpenMM's constraint kernels (CudaVVKernels.cpp:296-431)
extern "C" int vvb200_vv_kick(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a, int secondHalf,
                              int updatePosDelta, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_vv_kick", false, true);
    if (rc) return rc;
    if (updatePosDelta && !b->pos_delta) {
        vvb200_set_error("vvb200_vv_kick: pos_delta buffer required");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    if (secondHalf) {
        // second half: the extra forces of this step (VVIntegrator.cpp:316-325)
        if (!p->particlesLD.empty() && (rc = launchLangevin(p, b, a, st))) return rc;
        p->dev->extraForcesValid = true;
    }
    KParams k = makeParams(p, b, a);
    k.extraForces = p->dev->extraForcesValid ? 1 : 0;
    k.fuseNHC = 0;
    CUDA_TRY((dispatchA<KICK_VV>(p->precision, k.cosine, k, p->dev->numSM, st)));
    p->launches++;
    if (updatePosDelta) {
        const int grid = elementwiseGrid(p, p->N);
        switch (p->precision) {
        case VVB200_SINGLE: vv_delta_kernel<VVB200_SINGLE><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, p->N, p->par.step_size); break;
        case VVB200_MIXED: vv_delta_kernel<VVB200_MIXED><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, p->N, p->par.step_size); break;
        default: vv_delta_kernel<VVB200_DOUBLE><<<grid, THREADS, 0, st>>>(b->velm, b->pos_delta, p->N, p->par.step_size);
        }
        p->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return VVB200_OK;
}

extern "C" int vvb200_vv_positions(vvb200_plan *p, const vvb200_buffers *b, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_vv_positions", true, false);
    if (rc) return rc;
    if (!b->pos_delta) {
        vvb200_set_error("vvb200_vv_positions: pos_delta buffer required");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    vvb200_device_state *d = p->dev;
    if (p->tiled && envInt("VVB200_FUSED_FINISH", 1)) {
        // velocityVerletIntegratePositions + applyHardWallConstraints (+ the image mirror) in one streaming launch
        KParams k = makeParams(p, b, nullptr);
        k.posDelta = b->pos_delta;
        CUDA_TRY((dispatchB<VAR_VV_POSITIONS>(p->precision, false, k, d->numSM, st)));
        p->launches++;
        return k.imageFused ? VVB200_OK : vvb200_update_image_positions(p, b, stream);
    }
    const int grid = elementwiseGrid(p, p->N);
    const int nPairs = (int) p->drudePairs.size() / 2;
    const bool hw = p->par.max_drude_distance > 0 && nPairs > 0;
    const int gridHW = std::max(1, std::min((nPairs + 127) / 128, d->numSM * 8));
    const double hwScale = std::sqrt(BOLTZ_D * p->par.drude_temperature);
    switch (p->precision) {
    case VVB200_SINGLE:
        vv_positions_kernel<VVB200_SINGLE><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_SINGLE><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
        break;
    case VVB200_MIXED:
        vv_positions_kernel<VVB200_MIXED><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_MIXED><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
        break;
    default:
        vv_positions_kernel<VVB200_DOUBLE><<<grid, THREADS, 0, st>>>(b->posq, b->posq_correction, b->pos_delta, b->velm, p->N, p->par.step_size);
        if (hw) hardwall_pairs_kernel<VVB200_DOUBLE><<<gridHW, 128, 0, st>>>(b->posq, b->posq_correction, b->velm, d->drudePairs, nPairs, p->par.step_size, p->par.max_drude_distance, hwScale);
    }
    p->launches += hw ? 2 : 1;
    CUDA_TRY(cudaGetLastError());
    return vvb200_update_image_positions(p, b, stream);      // images follow their parents in every variant of this call
}

// ---- state ------------------------------------------------------------------------------------
extern "C" int vvb200_get_thermostat_state(vvb200_plan *p, vvb200_thermostat_state *out, void *stream) {
    if (!p || !p->dev || !out) {
        vvb200_set_error("vvb200_get_thermostat_state: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    cudaStream_t st = (cudaStream_t) stream;
    NhcDevice h;
    CUDA_TRY(cudaMemcpyAsync(&h, p->dev->nhc, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    memset(out, 0, sizeof *out);
    const int nc = p->par.num_nh_chains;
    out->num_temp_groups = p->numTempGroup;
    out->velocity_bias = h.vBias;
    for (int g = 0; g < 3; g++) {
        out->ke2[g] = h.ke2[g];
        out->vscale[g] = h.vscale[g];
    }
    for (int g = 0; g < p->numTempGroup; g++) {
        for (int k = 0; k < nc; k++) {
            out->eta[g * nc + k] = h.eta[g][k];
            out->eta_dotdot[g * nc + k] = h.etaDotDot[g][k];
        }
        for (int k = 0; k < nc + 1; k++)
            out->eta_dot[g * (nc + 1) + k] = h.etaDot[g][k];
    }
    return VVB200_OK;
}

extern "C" int vvb200_set_thermostat_state(vvb200_plan *p, const vvb200_thermostat_state *in, void *stream) {
    if (!p || !p->dev || !in) {
        vvb200_set_error("vvb200_set_thermostat_state: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    cudaStream_t st = (cudaStream_t) stream;
    NhcDevice h;
    CUDA_TRY(cudaMemcpyAsync(&h, p->dev->nhc, sizeof h, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int nc = p->par.num_nh_chains;
    for (int g = 0; g < p->numTempGroup; g++) {
        for (int k = 0; k < nc; k++) {
            h.eta[g][k] = in->eta[g * nc + k];
            h.etaDotDot[g][k] = in->eta_dotdot[g * nc + k];
        }
        for (int k = 0; k < nc + 1; k++)
            h.etaDot[g][k] = in->eta_dot[g * (nc + 1) + k];
    }
    h.preDt = 0.0;       // the look-ahead of the chain update belongs to the old state
    CUDA_TRY(cudaMemcpyAsync(p->dev->nhc, &h, sizeof h, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return VVB200_OK;
}

extern "C" int vvb200_calc_viscosity(vvb200_plan *p, double bx, double by, double bz, double *vMax, double *invVis,
                                     void *stream) {
    if (!p || !p->dev || !vMax || !invVis) {
        vvb200_set_error("vvb200_calc_viscosity: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    cudaStream_t st = (cudaStream_t) stream;
    double v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, &p->dev->nhc->vBias, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (p->precision == VVB200_SINGLE)
        v = (double) (float) v;
    *vMax = v;
    // CudaVVKernels.cpp:1129-1133
    const double vol = bx * by * bz;
    *invVis = v * vol * (1.0 / p->totalMassGlobal) / p->par.cos_acceleration * (2 * 3.1415926 / bz) * (2 * 3.1415926 / bz);
    return VVB200_OK;
}

extern "C" int vvb200_get_com_velocities(vvb200_plan *p, void *hostOut, void *stream) {
    if (!p || !p->dev || !hostOut) {
        vvb200_set_error("vvb200_get_com_velocities: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    cudaStream_t st = (cudaStream_t) stream;
    CUDA_TRY(cudaMemcpyAsync(hostOut, p->dev->comV, (size_t) p->M * 4 * mixedSize(p->precision), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return VVB200_OK;
}

// ---- checkpoint ---------------------------------------------------------------------------------
struct CheckpointBlob {
    uint32_t magic, version;
    int32_t numTG, nc, extraForcesValid, reserved;
    vvb200_thermostat_state state;
};
static const uint32_t kCheckpointMagic = 0x32425656u;   // "VVB2"

extern "C" int vvb200_checkpoint_size(const vvb200_plan *p, int64_t *bytes) {
    if (!p || !bytes) {
        vvb200_set_error("vvb200_checkpoint_size: null argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    *bytes = (int64_t) sizeof(CheckpointBlob);
    return VVB200_OK;
}

extern "C" int vvb200_checkpoint_save(vvb200_plan *p, void *out, int64_t capacity, void *stream) {
    if (!p || !p->dev || !out) {
        vvb200_set_error("vvb200_checkpoint_save: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    if (capacity < (int64_t) sizeof(CheckpointBlob)) {
        vvb200_set_error("vvb200_checkpoint_save: buffer of %lld bytes, %zu needed", (long long) capacity, sizeof(CheckpointBlob));
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    CheckpointBlob b;
    memset(&b, 0, sizeof b);
    b.magic = kCheckpointMagic;
    b.version = 1;
    b.numTG = p->numTempGroup;
    b.nc = p->par.num_nh_chains;
    b.extraForcesValid = p->dev->extraForcesValid ? 1 : 0;
    int rc = vvb200_get_thermostat_state(p, &b.state, stream);
    if (rc) return rc;
    memcpy(out, &b, sizeof b);
    return VVB200_OK;
}

extern "C" int vvb200_checkpoint_load(vvb200_plan *p, const void *in, int64_t bytes, void *stream) {
    if (!p || !p->dev || !in) {
        vvb200_set_error("vvb200_checkpoint_load: plan not uploaded or null argument");
        return VVB200_ERR_NOT_UPLOADED;
    }
    CheckpointBlob b;
    if (bytes < (int64_t) sizeof b) {
        vvb200_set_error("vvb200_checkpoint_load: truncated checkpoint (%lld of %zu bytes)", (long long) bytes, sizeof b);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    memcpy(&b, in, sizeof b);
    if (b.magic != kCheckpointMagic || b.version != 1) {
        vvb200_set_error("vvb200_checkpoint_load: not a vvb200 checkpoint (magic %08x version %u)", b.magic, b.version);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (b.numTG != p->numTempGroup || b.nc != p->par.num_nh_chains) {
        vvb200_set_error("vvb200_checkpoint_load: checkpoint has %d temperature groups x %d chains, this integrator %d x %d",
                         b.numTG, b.nc, p->numTempGroup, p->par.num_nh_chains);
        return VVB200_ERR_CONFLICT;
    }
    int rc = vvb200_set_thermostat_state(p, &b.state, stream);
    if (rc) return rc;
    // scale factors, energies and bias of the last step (what getViscosity / the temperature getters report)
    cudaStream_t st = (cudaStream_t) stream;
    NhcDevice *n = p->dev->nhc;
    CUDA_TRY(cudaMemcpyAsync(n->ke2, b.state.ke2, sizeof n->ke2, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(n->vscale, b.state.vscale, sizeof n->vscale, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(&n->vBias, &b.state.velocity_bias, sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->dev->extraForcesValid = b.extraForcesValid != 0;
    return VVB200_OK;
}

// ---- group temperatures of the current velocities (reporter path) ----------------------------------------
extern "C" int vvb200_measure_temperatures(vvb200_plan *p, const vvb200_buffers *b, const vvb200_step_args *a,
                                           vvb200_temperatures *out, void *stream) {
    int rc = checkStepArgs(p, b, "vvb200_measure_temperatures", false, false);
    if (rc) return rc;
    if (!out) {
        vvb200_set_error("vvb200_measure_temperatures: null output");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    memset(out, 0, sizeof *out);
    out->num_temp_groups = p->numTempGroup;
    // this plan's own particles: its DOFs and its mass (a molecule-partitioned run reports per partition; the caller
    // all-reduces ke2 and dof for the whole box)
    for (int g = 0; g < 3; g++) out->dof[g] = p->dof[g];
    if (!hasNH(p))
        return VVB200_OK;
    cudaStream_t st = (cudaStream_t) stream;
    const bool cosine = p->par.cos_acceleration != 0;
    if (p->tiled) {
        KParams k = makeParams(p, b, a);
        k.fuseNHC = 0;                   // sums only: the chains and scale factors stay as they are
        CUDA_TRY((dispatchA<KICK_NONE>(p->precision, cosine, k, p->dev->numSM, st)));
        p->launches++;
    } else if ((rc = generalThermostat(p, b, a, true, false, false, st))) {
        return rc;
    }
    double red[VVB200_NRED];
    CUDA_TRY(cudaMemcpyAsync(red, p->dev->nhc->red, sizeof red, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const double V = cosine ? red[3] * p->invMassTotal : 0.0;              // nhcFinish's expression on this plan's mass
    out->velocity_bias = V;
    for (int g = 0; g < p->numTempGroup; g++) {
        double ke2 = red[g];
        if (cosine) ke2 = red[g] - 2.0 * V * red[4 + g] + V * V * red[7 + g];
        out->ke2[g] = ke2;
        out->temperature[g] = out->dof[g] > 0 ? ke2 / (out->dof[g] * BOLTZ_D) : 0.0;
    }
    return VVB200_OK;
}

#ifdef VVB200_TRACE
extern "C" int vvb200_debug_trace(unsigned long long *out, int n) {
    return (int) cudaMemcpyFromSymbol(out, g_vvb200Trace, sizeof(unsigned long long) * n);
}
extern "C" int vvb200_debug_trace_streaming(unsigned long long *out, int n) {
    return (int) cudaMemcpyFromSymbol(out, g_vvb200TraceS, sizeof(unsigned long long) * n);
}
#endif

extern "C" int vvb200_set_resident_mode(vvb200_plan *p, int mode) {
    if (!p || !p->dev || mode < -1 || mode > 1) {
        vvb200_set_error("vvb200_set_resident_mode: plan not uploaded or mode outside -1..1");
        return p && p->dev ? VVB200_ERR_INVALID_ARGUMENT : VVB200_ERR_NOT_UPLOADED;
    }
    p->dev->residentMode = mode;
    return VVB200_OK;
}

extern "C" int64_t vvb200_resident_launch_count(const vvb200_plan *p) { return p && p->dev ? p->dev->residentLaunches : 0; }

// ---- host-buffer entry point ---------------------------------------------------------------------
// One middle-scheme step of a large system with the transfers overlapped (PCIe is full duplex, and pass A only needs
// velocities and forces):
//   copy-in stream :  velm+force chunk 0..C-1 | posq+corr chunk 0..C-1
//   compute stream :      pass A chunk 0..C-1 (sums accumulate; the last chunk advances the NH chains)
//                                             |  pass B chunk c as soon as its positions are in
//   copy-out stream:                          |     velm+posq+corr chunk c back as soon as pass B chunk c is done
// Chunks are contiguous tile ranges, so no molecule or Drude pair is ever split.  Results equal the unsplit step up to
// the association order of the group sums.
// phase: 0 = the whole step (single GPU); 1 = everything up to the group sums (copy-in of all arrays is queued, pass A
// runs chunk by chunk, the NH chains are NOT advanced); 2 = NH chains (with the peer exchange when attached), pass B
// and copy-out.  Multi-GPU callers run 1, all-reduce the reduction vector if they use NCCL, then 2.
static int stepHostPipelined(vvb200_plan *p, const vvb200_buffers *hb, const vvb200_step_args *a, cudaStream_t st, int phase) {
    vvb200_device_state *d = p->dev;
    const int numTiles = d->numTiles;
    const int C = std::max(1, std::min(std::min(envInt("VVB200_HOST_CHUNKS", 8), 64), numTiles));
    const size_t P = p->paddedN;
    const size_t ms = mixedSize(p->precision) * 4, rs = realSize(p->precision) * 4;
    const bool mixedMode = p->precision == VVB200_MIXED;
    if (!d->sIn) {
        CUDA_TRY(cudaStreamCreateWithFlags(&d->sIn, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&d->sOut, cudaStreamNonBlocking));
    }
    while (d->pipeEvents.size() < (size_t) 3 * C + 2) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        d->pipeEvents.push_back(e);
    }
    cudaEvent_t *evA = d->pipeEvents.data(), *evB = evA + C, *evC = evB + C, evStart = evC[C], evDone = evC[C + 1];
    if (phase != 2) {
        CUDA_TRY(cudaEventRecord(evStart, st));
        CUDA_TRY(cudaStreamWaitEvent(d->sIn, evStart, 0));
        CUDA_TRY(cudaStreamWaitEvent(d->sOut, evStart, 0));
    }
    auto tileLo = [&](int c) { return (int) ((long long) numTiles * c / C); };
    auto partLo = [&](int c) { return c >= C ? P : (size_t) p->tileStart[tileLo(c)]; };   // the last chunk takes the padding
    char *dVelm = (char *) d->hVelm, *dPosq = (char *) d->hPosq, *dCorr = (char *) d->hCorr;
    vvb200_buffers db;
    memset(&db, 0, sizeof db);
    db.posq = d->hPosq;
    db.posq_correction = mixedMode ? d->hCorr : nullptr;
    db.velm = d->hVelm;
    db.force = d->hForce;
    const bool cosine = p->par.cos_acceleration != 0;
    const bool extra = cosine || !p->particlesElectrolyte.empty();   // pass A then reads posq too
    // ---- velocities + forces in, pass A chunk by chunk ----
    for (int c = 0; c < C && phase != 2; c++) {
        const size_t lo = partLo(c), n = partLo(c + 1) - lo;
        CUDA_TRY(cudaMemcpyAsync(dVelm + lo * ms, (const char *) hb->velm + lo * ms, n * ms, cudaMemcpyHostToDevice, d->sIn));
        for (int k = 0; k < 3; k++)
            CUDA_TRY(cudaMemcpyAsync(d->hForce + k * P + lo, hb->force + k * P + lo, n * sizeof(long long), cudaMemcpyHostToDevice, d->sIn));
        if (extra) {
            CUDA_TRY(cudaMemcpyAsync(dPosq + lo * rs, (const char *) hb->posq + lo * rs, n * rs, cudaMemcpyHostToDevice, d->sIn));
            if (mixedMode)
                CUDA_TRY(cudaMemcpyAsync(dCorr + lo * rs, (const char *) hb->posq_correction + lo * rs, n * rs, cudaMemcpyHostToDevice, d->sIn));
        }
        CUDA_TRY(cudaEventRecord(evA[c], d->sIn));
        CUDA_TRY(cudaStreamWaitEvent(st, evA[c], 0));
        KParams k = makeParams(p, &db, a);
        k.tileBegin = tileLo(c);
        k.tileEnd = tileLo(c + 1);
        k.accumulateRed = c > 0;
        k.fuseNHC = hasNH(p) && c == C - 1 && phase == 0;
        CUDA_TRY((dispatchA<KICK_MIDDLE>(p->precision, cosine, k, d->numSM, st)));
        p->launches++;
    }
    // ---- positions in (queued behind the velocities: the copy engine keeps going while the sums are exchanged) ----
    for (int c = 0; c < C && phase != 2 && !extra; c++) {
        const size_t lo = partLo(c), n = partLo(c + 1) - lo;
        CUDA_TRY(cudaMemcpyAsync(dPosq + lo * rs, (const char *) hb->posq + lo * rs, n * rs, cudaMemcpyHostToDevice, d->sIn));
        if (mixedMode)
            CUDA_TRY(cudaMemcpyAsync(dCorr + lo * rs, (const char *) hb->posq_correction + lo * rs, n * rs, cudaMemcpyHostToDevice, d->sIn));
        CUDA_TRY(cudaEventRecord(evB[c], d->sIn));
    }
    if (phase == 1)
        return VVB200_OK;
    if (phase == 2 && (hasNH(p) || p->partitioned)) {
        int rc = launchNhc(p, st);      // all-reduce over NVLink peer memory (when attached) + NH chains
        if (rc) return rc;
    }
    // ---- pass B, results out ----
    for (int c = 0; c < C; c++) {
        const size_t lo = partLo(c), n = partLo(c + 1) - lo;
        if (!extra)
            CUDA_TRY(cudaStreamWaitEvent(st, evB[c], 0));
        KParams k = makeParams(p, &db, a);
        k.tileBegin = tileLo(c);
        k.tileEnd = tileLo(c + 1);
        CUDA_TRY((dispatchB<VAR_MIDDLE>(p->precision, cosine, k, d->numSM, st)));
        p->launches++;
        CUDA_TRY(cudaEventRecord(evC[c], st));
        CUDA_TRY(cudaStreamWaitEvent(d->sOut, evC[c], 0));
        CUDA_TRY(cudaMemcpyAsync((char *) hb->velm + lo * ms, dVelm + lo * ms, n * ms, cudaMemcpyDeviceToHost, d->sOut));
        CUDA_TRY(cudaMemcpyAsync((char *) hb->posq + lo * rs, dPosq + lo * rs, n * rs, cudaMemcpyDeviceToHost, d->sOut));
        if (mixedMode)
            CUDA_TRY(cudaMemcpyAsync((char *) hb->posq_correction + lo * rs, dCorr + lo * rs, n * rs, cudaMemcpyDeviceToHost, d->sOut));
    }
    CUDA_TRY(cudaEventRecord(evDone, d->sOut));
    CUDA_TRY(cudaStreamWaitEvent(st, evDone, 0));
    CUDA_TRY(cudaStreamSynchronize(st));
    return VVB200_OK;
}

static int ensureStaging(vvb200_plan *p) {
    vvb200_device_state *d = p->dev;
    const size_t P = p->paddedN;
    const size_t ms = mixedSize(p->precision), rs = realSize(p->precision);
    if (d->stagedN != P) {
        CUDA_TRY(cudaMalloc(&d->hPosq, P * 4 * rs)); d->allocations.push_back(d->hPosq);
        CUDA_TRY(cudaMalloc(&d->hCorr, P * 4 * rs)); d->allocations.push_back(d->hCorr);
        CUDA_TRY(cudaMalloc(&d->hVelm, P * 4 * ms)); d->allocations.push_back(d->hVelm);
        CUDA_TRY(cudaMalloc((void **) &d->hForce, P * 3 * sizeof(long long))); d->allocations.push_back(d->hForce);
        d->stagedN = P;
    }
    return VVB200_OK;
}

static bool pipelineEligible(const vvb200_plan *p) {
    return p->tiled && p->par.use_middle_scheme && p->imagePairs.empty() && p->particlesLD.empty();
}

static int checkHostSplit(vvb200_plan *p, const vvb200_buffers *hb, const char *who) {
    if (!p || !hb) {
        vvb200_set_error("%s: null argument", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (!p->dev) {
        vvb200_set_error("%s: plan not uploaded (call vvb200_plan_upload)", who);
        return VVB200_ERR_NOT_UPLOADED;
    }
    if (!pipelineEligible(p)) {
        vvb200_set_error("%s: needs the tiled middle-scheme path without Langevin or image particles", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (!hb->velm || !hb->posq || !hb->force || (p->precision == VVB200_MIXED && !hb->posq_correction)) {
        vvb200_set_error("%s: missing host buffer", who);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    return VVB200_OK;
}

extern "C" int vvb200_step_host_begin(vvb200_plan *p, const vvb200_buffers *hb, const vvb200_step_args *a, void *stream) {
    int rc = checkHostSplit(p, hb, "vvb200_step_host_begin");
    if (rc) return rc;
    if ((rc = ensureStaging(p))) return rc;
    return stepHostPipelined(p, hb, a, (cudaStream_t) stream, 1);
}

extern "C" int vvb200_step_host_finish(vvb200_plan *p, const vvb200_buffers *hb, const vvb200_step_args *a, void *stream) {
    int rc = checkHostSplit(p, hb, "vvb200_step_host_finish");
    if (rc) return rc;
    if (p->dev->stagedN != (size_t) p->paddedN) {
        vvb200_set_error("vvb200_step_host_finish: call vvb200_step_host_begin first");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    return stepHostPipelined(p, hb, a, (cudaStream_t) stream, 2);
}

extern "C" int vvb200_step_host(vvb200_plan *p, const vvb200_buffers *hb, const vvb200_step_args *a, int steps, void *stream) {
    if (!p || !hb || steps < 0) {
        vvb200_set_error("vvb200_step_host: invalid argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    if (!p->dev) {
        vvb200_set_error("vvb200_step_host: plan not uploaded (call vvb200_plan_upload)");
        return VVB200_ERR_NOT_UPLOADED;
    }
    if (!p->particlesLD.empty()) {
        vvb200_set_error("vvb200_step_host: Langevin systems need the device random buffer; use the device entry points");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    cudaStream_t st = (cudaStream_t) stream;
    vvb200_device_state *d = p->dev;
    const size_t P = p->paddedN;
    const size_t ms = mixedSize(p->precision), rs = realSize(p->precision);
    int rc = ensureStaging(p);
    if (rc) return rc;
    const bool mixedMode = p->precision == VVB200_MIXED;
    // one step of a large tiled system: overlap copy-in, the two passes and copy-out (VVB200_HOST_PIPELINE=0 disables)
    if (steps == 1 && pipelineEligible(p) && p->N >= envInt("VVB200_HOST_PIPELINE_MIN", 1000000) &&
        envInt("VVB200_HOST_PIPELINE", 1))
        return stepHostPipelined(p, hb, a, st, 0);
    CUDA_TRY(cudaMemcpyAsync(d->hVelm, hb->velm, P * 4 * ms, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->hForce, hb->force, P * 3 * sizeof(long long), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d->hPosq, hb->posq, P * 4 * rs, cudaMemcpyHostToDevice, st));
    if (mixedMode)
        CUDA_TRY(cudaMemcpyAsync(d->hCorr, hb->posq_correction, P * 4 * rs, cudaMemcpyHostToDevice, st));
    vvb200_buffers db;
    memset(&db, 0, sizeof db);
    db.posq = d->hPosq;
    db.posq_correction = mixedMode ? d->hCorr : nullptr;
    db.velm = d->hVelm;
    db.force = d->hForce;
    for (int s = 0; s < steps; s++) {
        int rc;
        if (p->par.use_middle_scheme) {
            if ((rc = vvb200_step_middle(p, &db, a, stream))) return rc;
        } else {
            if ((rc = vvb200_step_vv_first(p, &db, a, stream))) return rc;
            if ((rc = vvb200_step_vv_second(p, &db, a, stream))) return rc;
        }
    }
    CUDA_TRY(cudaMemcpyAsync(hb->posq, d->hPosq, P * 4 * rs, cudaMemcpyDeviceToHost, st));
    if (mixedMode)
        CUDA_TRY(cudaMemcpyAsync(hb->posq_correction, d->hCorr, P * 4 * rs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(hb->velm, d->hVelm, P * 4 * ms, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return VVB200_OK;
}
