/*
 * vv_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C) of the reference
 * plugin's per-step integration path.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it, and only as the checker / the reported CPU baseline.
 *
 * PARITY STATUS: the reference ships no tests, golden vectors or fixtures for this path
 * (platforms/cuda/tests/ holds only a CMakeLists.txt) => "parity unpinned" by the reference's
 * own tests.  This restatement is instead pinned against oracle/_ref: the reference's own
 * kernel sources (the .cu files under platforms/cuda/src/kernels) compiled for the host from where they lie
 * under /root/reference (see oracle/Makefile, oracle/ref_harness.cpp) and run on identical
 * inputs (tests/test_oracle_vs_ref.py), plus the analytic known-answer tests of
 * tests/test_oracle_kat.py.
 *
 * The same source is compiled three times (-DVVO_SINGLE / -DVVO_MIXED / -DVVO_DOUBLE) into
 * libvvoracle_{single,mixed,double}.so; `real` / `mixed` follow OpenMM's CudaPrecision modes.
 *
 * All citations "file:line" are relative to /root/reference.
 */
#ifndef VV_ORACLE_H_
#define VV_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(VVO_SINGLE)
typedef float vvo_real;
typedef float vvo_mixed;
#elif defined(VVO_DOUBLE)
typedef double vvo_real;
typedef double vvo_mixed;
#else /* VVO_MIXED (default) */
#ifndef VVO_MIXED
#define VVO_MIXED 1
#endif
typedef float vvo_real;
typedef double vvo_mixed;
#endif

typedef struct { vvo_real x, y, z, w; } vvo_real4;
typedef struct { vvo_real x, y, z; } vvo_real3;
typedef struct { vvo_mixed x, y, z, w; } vvo_mixed4;
typedef struct { float x, y, z, w; } vvo_float4;
typedef struct { int32_t x, y; } vvo_int2;

/* What VVIntegrator::initialize and the Cuda*Kernel::initialize methods read from
 * OpenMM's System / ContextImpl / DrudeForce (VVIntegrator.cpp:96-151,
 * CudaVVKernels.cpp:66-77, 483-594, 775-804, 884-891, 954-957, 1028-1031). */
typedef struct {
    int32_t numParticles;          /* System::getNumParticles() == cu.getNumAtoms() */
    int32_t paddedNumAtoms;        /* cu.getPaddedNumAtoms() */
    int32_t numMolecules;          /* ContextImpl::getMolecules().size() */
    const double *masses;          /* System::getParticleMass(i) */
    const int32_t *particleMolId;  /* molecule index of each particle */
    int32_t numDrude;              /* DrudeForce::getNumParticles() (0 if no DrudeForce) */
    const int32_t *drudePairs;     /* (p, p1) per DrudeForce entry, in DrudeForce order */
    int32_t numConstraints;
    const int32_t *constraints;    /* (p, p1) per System constraint */
    int32_t hasCMMotionRemover;
    int32_t numLD;
    const int32_t *particlesLD;    /* addParticleLangevin order */
    int32_t numImagePairs;
    const int32_t *imagePairs;     /* (image, parent) in addImagePair order */
    int32_t numElectrolyte;
    const int32_t *particlesElectrolyte; /* addParticleElectrolyte order, duplicates kept */
} vvo_system;

/* VVIntegrator state (VVIntegrator.h:62-431), after the auto-defaults of
 * VVIntegrator.cpp:106-121 have been applied by the caller. */
typedef struct {
    double temperature, frequency, drudeTemperature, drudeFrequency, stepSize;
    int32_t numNHChains, loopsPerStep;
    int32_t useCOMTempGroup, useMiddleScheme;
    double maxDrudeDistance, friction, drudeFriction;
    double mirrorLocation, electricField, cosAcceleration;
} vvo_params;

/* OpenMM-owned device arrays, here plain host arrays in the same layouts
 * (SURVEY.md Appendix D). */
typedef struct {
    vvo_real4 *posq;              /* [padded] x,y,z,q */
    vvo_real4 *posqCorrection;    /* [padded], mixed mode only (else NULL) */
    vvo_mixed4 *velm;             /* [padded] vx,vy,vz,1/m */
    long long *force;             /* [3*padded] fixed point 2^32, component-major */
    vvo_mixed4 *posDelta;         /* [padded] */
    const vvo_float4 *random;     /* injected N(0,1) stream */
} vvo_buffers;

typedef struct vvo_ctx vvo_ctx;

enum { VVO_TG_ATOM = 0, VVO_TG_COM = 1, VVO_TG_DRUDE = 2, VVO_NUM_TG_MAX = 3 };
#define VVO_MAX_CHAINS 16

/* integer / fp64 arrays produced by the index builders; ids for vvo_get_array */
enum {
    VVO_ARR_PARTICLES_NH = 0, VVO_ARR_MOLECULES_NH, VVO_ARR_PARTICLE_MOL_ID,
    VVO_ARR_DRUDE_PAIRS, VVO_ARR_SORTED_BY_MOL, VVO_ARR_PARTICLES_IN_MOLECULES,
    VVO_ARR_NORMAL_NH, VVO_ARR_PAIRS_NH, VVO_ARR_NORMAL_LD, VVO_ARR_PAIRS_LD,
    VVO_ARR_IMAGE_PAIRS, VVO_ARR_ELECTROLYTE,
    VVO_ARR_MOLECULE_MASSES = 100, VVO_ARR_MOLECULE_INV_MASSES, VVO_ARR_DOF,
    VVO_ARR_ETA_MASS, VVO_ARR_NKBT, VVO_ARR_INV_MASS_TOTAL,
    VVO_ARR_ETA, VVO_ARR_ETA_DOT, VVO_ARR_ETA_DOTDOT, VVO_ARR_KE2, VVO_ARR_VSCALE,
    VVO_ARR_VBIAS
};

const char *vvo_last_error(void);
int vvo_precision_mode(void);          /* 0 single, 1 mixed, 2 double */
void vvo_set_num_threads(int n);       /* OpenMP threads for the element-wise loops */
int vvo_get_max_threads(void);

/* OpenMM ContextImpl::getMolecules restated [OMM-mem]: connected components over bonds,
 * numbered by ascending first atom. Returns the number of molecules. */
int vvo_find_molecules(int numParticles, int numBonds, const int32_t *bonds, int32_t *molIdOut);

/* literal != 0 -> builders follow the reference's O(N^2) loops statement by statement;
 * literal == 0 -> same results through O(N) lookups (needed for the >=1M CPU baseline). */
vvo_ctx *vvo_create(const vvo_system *sys, const vvo_params *par, int literal);
void vvo_destroy(vvo_ctx *c);
int vvo_num_temp_groups(const vvo_ctx *c);
/* returns element count (int32 elements for integer ids, doubles for fp ids) */
int64_t vvo_get_array(const vvo_ctx *c, int id, const void **ptr);
void vvo_set_nhc_state(vvo_ctx *c, const double *eta, const double *etaDot, const double *etaDotDot);

/* VVIntegrator::propagateNHChain, VVIntegrator.cpp:340-376 */
void vvo_propagate_nh_chain(double stepSize, int loopsPerStep, int numNHChains,
                            double *eta, double *etaDot, double *etaDotDot, const double *etaMass,
                            double ke2, double ke2Target, double tTarget, double *factor);

/* Individual kernels (one call == one reference launch over the full index range). */
void vvo_reset_extra_force(vvo_ctx *c);                                   /* middle.cu:227 */
void vvo_langevin_force(vvo_ctx *c, const vvo_buffers *b, unsigned randomIndex); /* drudeLangevin.cu:2 */
void vvo_electric_force(vvo_ctx *c, const vvo_buffers *b);                /* electricField.cu:2 */
void vvo_cosine_force(vvo_ctx *c, const vvo_buffers *b, double invBoxZ);  /* cosineAccelerate.cu:2 */
void vvo_middle_vel(vvo_ctx *c, const vvo_buffers *b);                    /* middle.cu:6 */
void vvo_middle_pos1(vvo_ctx *c, const vvo_buffers *b);                   /* middle.cu:29 */
void vvo_middle_pos2(vvo_ctx *c, const vvo_buffers *b);                   /* middle.cu:47 */
void vvo_middle_pos3(vvo_ctx *c, const vvo_buffers *b);                   /* middle.cu:66 */
void vvo_hard_wall(vvo_ctx *c, const vvo_buffers *b);                     /* middle.cu:106 */
void vvo_vv_velocities(vvo_ctx *c, const vvo_buffers *b, int updatePosDelta); /* velocityVerlet.cu:6 */
void vvo_vv_positions(vvo_ctx *c, const vvo_buffers *b);                  /* velocityVerlet.cu:35 */
void vvo_calc_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ); /* cosineAccelerate.cu:16,34 */
void vvo_remove_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ); /* :63 */
void vvo_restore_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ); /* :76 */
void vvo_scale_velocity(vvo_ctx *c, const vvo_buffers *b);                /* CudaVVKernels.cpp:670-754 */
void vvo_update_images(vvo_ctx *c, const vvo_buffers *b);                 /* imageCharge.cu:2 */
void vvo_calc_viscosity(vvo_ctx *c, double boxX, double boxY, double boxZ, double *vMax, double *invVis);

/* Whole steps with frozen forces: VVIntegrator::stepMiddle (:232-270) / stepVV (:272-338).
 * randomIndex is advanced per step as prepareRandomNumbers would (CudaVVKernels.cpp:863).
 * For stepVV, forcesValid follows VVIntegrator's forcesAreValid flag (frozen forces: no-op). */
void vvo_step(vvo_ctx *c, const vvo_buffers *b, int steps, double invBoxZ, unsigned *randomIndex);

/* Stand-in for OpenMM's constraint solvers between the sub-steps (oracle/constraint_standin.h): NOT from the
 * reference and NOT OpenMM's algorithm -- a deterministic operator with the same contract (applyConstraints rewrites
 * posDelta, applyVelocityConstraints rewrites velm; CudaVVKernels.cpp:151,176,351,427) so that the constraint-bearing
 * flow is exercised with posDelta != oldDelta.  vvo_step applies it where the reference calls OpenMM. */
void vvo_set_constraint_standin(vvo_ctx *c, int numClusters, const int32_t *clusterOffset, const int32_t *atoms,
                                const double *distance, int iterations);
void vvo_apply_constraints(vvo_ctx *c, const vvo_buffers *b);
void vvo_apply_velocity_constraints(vvo_ctx *c, const vvo_buffers *b);

/* Toy force field for the 10^4-step statistical tests (NOT from the reference; shared
 * definition with tests/toyforce): F_i = -k_t (x_i - x0_i) for massive non-Drude particles,
 * plus a Drude spring -k_d (x_drude - x_parent) on each pair. Writes fixed-point forces. */
void vvo_toy_forces(vvo_ctx *c, const vvo_buffers *b, const double *x0, double kTether, double kDrude);

#ifdef __cplusplus
}
#endif
#endif
