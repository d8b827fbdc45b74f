/*
 * ref_kernels.inc.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The reference's kernel sources, #included from where they lie under /root/reference/platforms/cuda/src/kernels
 * (-I in oracle/Makefile; never copied into this repository), behind a restatement of what OpenMM's
 * CudaContext::createModule prepends to every module [OMM-mem]: the real/mixed typedefs, make_*, SQRT, RECIP,
 * USE_*_PRECISION -- and with the per-module #defines (NUM_ATOMS, ...) turned into variables (constant memory on the
 * GPU).  Compiled by g++ (host flavour: a tiny SIMT shim supplies blockIdx/threadIdx) or nvcc (sm_100a flavour).
 * Included by oracle/ref_harness.cpp (restated host schedule) and oracle/mini_openmm/plugin_kernels.cpp (the
 * reference's OWN host code driving them through a stand-in CudaContext).  Select the precision mode with
 * -DVVREF_SINGLE / -DVVREF_MIXED (default) / -DVVREF_DOUBLE.
 */
#ifndef VVREF_KERNELS_INC_H_
#define VVREF_KERNELS_INC_H_
#if defined(__CUDACC__)
#define VVREF_GPU 1
#include <cuda_runtime.h>
#else
#define VVREF_GPU 0
#include <vector_types.h>
#include <vector_functions.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#undef __global__
#undef __device__
#undef __shared__
#undef __restrict__
#define __global__
#define __device__
#define __shared__
#define __restrict__
struct VVRefDim3 { unsigned x, y, z; };
static thread_local VVRefDim3 blockIdx, threadIdx, blockDim, gridDim;
static inline void __syncthreads() {}
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
using std::fabs;
#endif

/* ---- prelude as CudaContext::createModule would prepend [OMM-mem] ---------------------- */
#if defined(VVREF_DOUBLE)
#define USE_DOUBLE_PRECISION 1
typedef double real; typedef double2 real2; typedef double3 real3; typedef double4 real4;
typedef double mixed; typedef double2 mixed2; typedef double3 mixed3; typedef double4 mixed4;
#define make_real2 make_double2
#define make_real3 make_double3
#define make_real4 make_double4
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrt
#define RECIP(x) (1.0/(x))
#define VVREF_MODE 2
#elif defined(VVREF_SINGLE)
typedef float real; typedef float2 real2; typedef float3 real3; typedef float4 real4;
typedef float mixed; typedef float2 mixed2; typedef float3 mixed3; typedef float4 mixed4;
#define make_real2 make_float2
#define make_real3 make_float3
#define make_real4 make_float4
#define make_mixed2 make_float2
#define make_mixed3 make_float3
#define make_mixed4 make_float4
#define SQRT sqrtf
#define RECIP(x) (1.0f/(x))
#define VVREF_MODE 0
#else
#define USE_MIXED_PRECISION 1
typedef float real; typedef float2 real2; typedef float3 real3; typedef float4 real4;
typedef double mixed; typedef double2 mixed2; typedef double3 mixed3; typedef double4 mixed4;
#define make_real2 make_float2
#define make_real3 make_float3
#define make_real4 make_float4
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrtf
#define RECIP(x) (1.0f/(x))
#define VVREF_MODE 1
#endif

/* ---- per-module #defines become variables ---------------------------------------------- */
#if VVREF_GPU
#define VVREF_VAR __constant__
#else
#define VVREF_VAR static
#endif
VVREF_VAR int vr_NUM_ATOMS, vr_PADDED_NUM_ATOMS, vr_NUM_DRUDE_PAIRS;
VVREF_VAR int vr_NUM_PARTICLES_NH, vr_NUM_MOLECULES_NH, vr_NUM_NORMAL_PARTICLES_NH, vr_NUM_PAIRS_NH;
VVREF_VAR int vr_NUM_NORMAL_PARTICLES_LD, vr_NUM_PAIRS_LD, vr_NUM_IMAGES, vr_NUM_PARTICLES_ELECTROLYTE;
#define NUM_ATOMS vr_NUM_ATOMS
#define PADDED_NUM_ATOMS vr_PADDED_NUM_ATOMS
#define NUM_DRUDE_PAIRS vr_NUM_DRUDE_PAIRS
#define NUM_PARTICLES_NH vr_NUM_PARTICLES_NH
#define NUM_MOLECULES_NH vr_NUM_MOLECULES_NH
#define NUM_NORMAL_PARTICLES_NH vr_NUM_NORMAL_PARTICLES_NH
#define NUM_PAIRS_NH vr_NUM_PAIRS_NH
#define NUM_NORMAL_PARTICLES_LD vr_NUM_NORMAL_PARTICLES_LD
#define NUM_PAIRS_LD vr_NUM_PAIRS_LD
#define NUM_IMAGES vr_NUM_IMAGES
#define NUM_PARTICLES_ELECTROLYTE vr_NUM_PARTICLES_ELECTROLYTE
#define TG_ATOM 0
#define TG_COM 1
#define TG_DRUDE 2

#if !VVREF_GPU
/* `extern __shared__ mixed temp[];` inside the two sum kernels binds to this array */
mixed temp[16];
#endif

/* ---- the reference's kernel sources, included from where they lie (-I in the Makefile) --- */
#include "vectorOps.cu"
#include "middle.cu"
#define applyHardWallConstraints applyHardWallConstraints_vv   /* duplicate definitions in */
#define resetExtraForce resetExtraForce_vv                     /* velocityVerlet.cu:74,195 */
#include "velocityVerlet.cu"
#undef applyHardWallConstraints
#undef resetExtraForce
#include "drudeLangevin.cu"
#include "cosineAccelerate.cu"
#include "electricField.cu"
#include "imageCharge.cu"
/* drudeNoseHoover.cu uses `#if NUM_TG > TG_COM`, so NUM_TG must be a literal: one copy each */
#define NUM_TG 1
#define calcCOMVelocities calcCOMVelocities_tg1
#define normalizeVelocities normalizeVelocities_tg1
#define computeNormalizedKineticEnergies computeNormalizedKineticEnergies_tg1
#define sumNormalizedKineticEnergies sumNormalizedKineticEnergies_tg1
#define scaleVelocity scaleVelocity_tg1
#include "drudeNoseHoover.cu"
#undef NUM_TG
#undef calcCOMVelocities
#undef normalizeVelocities
#undef computeNormalizedKineticEnergies
#undef sumNormalizedKineticEnergies
#undef scaleVelocity
#define NUM_TG 2
#define calcCOMVelocities calcCOMVelocities_tg2
#define normalizeVelocities normalizeVelocities_tg2
#define computeNormalizedKineticEnergies computeNormalizedKineticEnergies_tg2
#define sumNormalizedKineticEnergies sumNormalizedKineticEnergies_tg2
#define scaleVelocity scaleVelocity_tg2
#include "drudeNoseHoover.cu"
#undef NUM_TG
#undef calcCOMVelocities
#undef normalizeVelocities
#undef computeNormalizedKineticEnergies
#undef sumNormalizedKineticEnergies
#undef scaleVelocity
#define NUM_TG 3
#define calcCOMVelocities calcCOMVelocities_tg3
#define normalizeVelocities normalizeVelocities_tg3
#define computeNormalizedKineticEnergies computeNormalizedKineticEnergies_tg3
#define sumNormalizedKineticEnergies sumNormalizedKineticEnergies_tg3
#define scaleVelocity scaleVelocity_tg3
#include "drudeNoseHoover.cu"
#undef NUM_TG
#undef calcCOMVelocities
#undef normalizeVelocities
#undef computeNormalizedKineticEnergies
#undef sumNormalizedKineticEnergies
#undef scaleVelocity

#endif /* VVREF_KERNELS_INC_H_ */
