/*
 * vvb200.h -- C ABI of libvvb200.so: the B200-native (sm_100a) integration path of the OpenMM
 * velocity-Verlet plugin (z-gong/openmm-velocityVerlet).
 *
 * This is the drop-in boundary.  Everything above it (the VVIntegrator class, the seven
 * Cuda*Kernel classes and the KernelFactory registration, see INTEGRATION.md and
 * openmm-velocityverlet_b200/csrc/glue/) is a thin layer that forwards OpenMM's own device
 * arrays and stream to these entry points.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a VVB200_ERR_* code; vvb200_last_error() gives the
 *     message (thread-local).  Nothing throws, nothing allocates inside step calls.
 *   - "device pointer" arguments are OpenMM's arrays used IN PLACE, in OpenMM's layouts:
 *       posq            real4  [paddedNumAtoms]   (x, y, z, charge)
 *       posqCorrection  real4  [paddedNumAtoms]   mixed mode only, else NULL
 *       velm            mixed4 [paddedNumAtoms]   (vx, vy, vz, 1/mass)
 *       force           int64  [3*paddedNumAtoms] fixed point 2^32, component-major
 *       posDelta        mixed4 [paddedNumAtoms]   (only the constraint-bearing entry points)
 *       random          float4 [...]              OpenMM's N(0,1) buffer (Langevin only)
 *     real = float (single, mixed) | double (double); mixed = float (single) | double.
 *   - all launches go to the cudaStream_t passed as `stream` (void* here), i.e. the OpenMM
 *     context's own stream; no call synchronises the stream unless documented.
 *
 * Citations "file:line" name the reference interface each entry point replaces, relative to
 * the reference repository root.
 */
#ifndef VVB200_H_
#define VVB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VVB200_VERSION 100

enum {
    VVB200_OK = 0,
    VVB200_ERR_INVALID_ARGUMENT = 1,
    VVB200_ERR_CONFLICT = 2,             /* an OpenMMException of the reference's initialize() */
    VVB200_ERR_UNSUPPORTED_TOPOLOGY = 3,
    VVB200_ERR_CUDA = 4,
    VVB200_ERR_NOT_UPLOADED = 5,
    VVB200_ERR_NO_DEVICE_CODE = 6
};

enum { VVB200_SINGLE = 0, VVB200_MIXED = 1, VVB200_DOUBLE = 2 };          /* CudaPrecision */
enum { VVB200_TG_ATOM = 0, VVB200_TG_COM = 1, VVB200_TG_DRUDE = 2 };      /* CudaVVKernels.cpp:49 */
#define VVB200_MAX_CHAINS 16

/* What VVIntegrator::initialize (VVIntegrator.cpp:92-188) and the Cuda*Kernel::initialize
 * methods (CudaVVKernels.cpp:56-117, 242-294, 462-667, 761-824, 878-902, 940-969, 998-1035) read
 * from System / ContextImpl / DrudeForce / the integrator's particle lists. Host pointers. */
typedef struct {
    int32_t num_particles;            /* System::getNumParticles() == cu.getNumAtoms() */
    int32_t padded_num_atoms;         /* cu.getPaddedNumAtoms() */
    int32_t num_molecules;            /* ContextImpl::getMolecules().size() */
    const double *masses;             /* [N] System::getParticleMass */
    const int32_t *particle_mol_id;   /* [N] molecule of each particle (VVIntegrator.cpp:123-128) */
    int32_t num_drude;                /* DrudeForce::getNumParticles(), 0 without a DrudeForce */
    const int32_t *drude_pairs;       /* [2*num_drude] (p, p1) = (Drude, parent), DrudeForce order */
    int32_t num_constraints;
    const int32_t *constraints;       /* [2*num_constraints] (p, p1) */
    int32_t has_cm_motion_remover;    /* CudaVVKernels.cpp:550-558 */
    int32_t num_langevin;
    const int32_t *particles_langevin;    /* VVIntegrator::addParticleLangevin order (VVIntegrator.h:199) */
    int32_t num_image_pairs;
    const int32_t *image_pairs;           /* [2*n] (image, parent), addImagePair order (VVIntegrator.cpp:76) */
    int32_t num_electrolyte;
    const int32_t *particles_electrolyte; /* addParticleElectrolyte order, duplicates kept (VVIntegrator.h:302) */
} vvb200_system;

/* VVIntegrator parameters (VVIntegrator.h:62-431) with the auto-defaults of
 * VVIntegrator.cpp:106-121 already resolved by the caller. step_size may be changed later
 * through vvb200_set_step_size (the reference re-reads it every step, CudaVVKernels.cpp:137). */
typedef struct {
    double temperature, frequency, drude_temperature, drude_frequency, step_size;
    int32_t num_nh_chains, loops_per_step;
    int32_t use_com_temp_group, use_middle_scheme;
    double max_drude_distance, friction, drude_friction;
    double mirror_location, electric_field, cos_acceleration;
} vvb200_params;

typedef struct {
    void *posq;
    void *posq_correction;
    void *velm;
    const long long *force;
    void *pos_delta;
    const void *random;
} vvb200_buffers;

/* per-call scalars */
typedef struct {
    uint32_t random_index;   /* what integration.prepareRandomNumbers(n) returned (CudaVVKernels.cpp:863) */
    double inv_box_z;        /* cu.getInvPeriodicBoxSizePointer()->z (cosine runs only) */
} vvb200_step_args;

typedef struct vvb200_plan vvb200_plan;

const char *vvb200_last_error(void);
int vvb200_version(void);
/* 1 when the library carries sm_100a device code (always, for this build) */
int vvb200_has_device_code(void);

/* ContextImpl::getMolecules restated for standalone callers [OMM-mem]: connected components
 * over `bonds` ([2*num_bonds]), numbered by ascending first atom. Writes mol_id[N], returns the
 * number of molecules in *num_molecules. (The OpenMM glue passes ContextImpl's own result.) */
int vvb200_find_molecules(int32_t num_particles, int32_t num_bonds, const int32_t *bonds,
                          int32_t *mol_id, int32_t *num_molecules);

/* ---- plan: host-side index builders (no CUDA call) ------------------------------------------
 * Replaces the builders of VVIntegrator::initialize (VVIntegrator.cpp:123-155) and of the
 * Cuda*Kernel::initialize methods, in O(N).  Integer arrays and fp64 DOFs / eta masses are
 * bit-identical to the reference's.  The reference's exceptions map to VVB200_ERR_CONFLICT with
 * the reference's message text. */
int vvb200_plan_create(const vvb200_system *sys, const vvb200_params *par, int precision, vvb200_plan **out);
void vvb200_plan_destroy(vvb200_plan *plan);

/* read-only views of the built arrays (host memory owned by the plan) */
enum {
    VVB200_ARR_PARTICLES_NH = 0,        /* VVIntegrator::getParticlesNH */
    VVB200_ARR_MOLECULES_NH,            /* VVIntegrator::getMoleculesNH */
    VVB200_ARR_PARTICLE_MOL_ID,         /* VVIntegrator::getParticleMolId */
    VVB200_ARR_DRUDE_PAIRS,             /* drudePairsVec, CudaVVKernels.cpp:67-74 */
    VVB200_ARR_SORTED_BY_MOL,           /* particlesSortedByMolIdVec, :483-494 */
    VVB200_ARR_PARTICLES_IN_MOLECULES,  /* particlesInMoleculesVec (count,start) */
    VVB200_ARR_NORMAL_NH,               /* normalParticlesNHVec, :529 */
    VVB200_ARR_PAIRS_NH,                /* pairParticlesNHVec, :523 */
    VVB200_ARR_NORMAL_LD,               /* normalParticlesLDVec, :804 */
    VVB200_ARR_PAIRS_LD,                /* pairParticlesLDVec, :791 */
    VVB200_ARR_IMAGE_PAIRS,             /* imagePairsVec, :885-887 */
    VVB200_ARR_ELECTROLYTE,             /* particlesElectrolyteVec, :954 */
    VVB200_ARR_TILE_START,              /* new: molecule-aligned tile boundaries of the fused path */
    VVB200_ARR_SLOT_META,               /* new: packed per-slot topology word of the fused path */
    VVB200_ARR_TILE_MOL_OFFSET,         /* new: [tiles+1] prefix into TILE_MOL_LIST */
    VVB200_ARR_TILE_MOL_LIST,           /* new: thermostat molecules of each tile, in order of first appearance */
    VVB200_ARR_TILE_MOL_FRAG,           /* new: per entry of TILE_MOL_LIST: fragment index, or -1 for a whole molecule */
    VVB200_ARR_SPLIT_MOL_ID,            /* new: thermostat molecules longer than a tile (cut into fragments) */
    VVB200_ARR_SPLIT_FRAG_OFFSET,       /* new: [cut molecules+1] prefix into SPLIT_FRAG_LIST */
    VVB200_ARR_SPLIT_FRAG_LIST,         /* new: fragment indices of each cut molecule, in tile order */
    VVB200_ARR_IMAGE_OF                 /* new: [N] image particle mirrored by each particle's thread or -1; empty when
                                           the image update is not fused into pass B */
};
int vvb200_plan_get_int_array(const vvb200_plan *plan, int which, const int32_t **ptr, int64_t *len);

enum {
    VVB200_F64_MOLECULE_MASSES = 0,     /* VVIntegrator.cpp:130-132 */
    VVB200_F64_MOLECULE_INV_MASSES,     /* VVIntegrator::getMoleculeInvMass */
    VVB200_F64_DOF,                     /* tempGroupDof[3], CudaVVKernels.cpp:497-564 */
    VVB200_F64_ETA_MASS,                /* etaMass[numTempGroup][numNHChains], :583-594 */
    VVB200_F64_NKBT,                    /* tempGroupNkbT[numTempGroup] */
    VVB200_F64_INV_MASS_TOTAL,          /* invMassTotal, :1028-1031 */
    VVB200_F64_DOF_GLOBAL               /* new: whole-box DOFs[3] set by vvb200_set_global_thermostat */
};
int vvb200_plan_get_f64_array(const vvb200_plan *plan, int which, const double **ptr, int64_t *len);
int vvb200_plan_num_temp_groups(const vvb200_plan *plan);
/* 1: every thermostat molecule and Drude pair fits one tile -> fused two-pass kernels;
 * 0: general gather kernels (any topology). */
int vvb200_plan_uses_tiled_path(const vvb200_plan *plan);
/* random numbers one Langevin application consumes (the padded request of CudaVVKernels.cpp:863) */
uint32_t vvb200_plan_random_request(const vvb200_plan *plan);
int vvb200_set_step_size(vvb200_plan *plan, double step_size);

/* VVIntegrator::propagateNHChain (VVIntegrator.cpp:340-376; public API, host, fp64).
 * eta[nc], eta_dot[nc+1], eta_dotdot[nc], eta_mass[nc]. */
int vvb200_propagate_nh_chain(double step_size, int loops_per_step, int num_nh_chains,
                              double *eta, double *eta_dot, double *eta_dotdot, const double *eta_mass,
                              double ke2, double ke2_target, double t_target, double *factor);

/* ---- device side --------------------------------------------------------------------------- */
/* Uploads the plan's tables to the current CUDA device and allocates the plugin-private scratch
 * (what the Cuda*Kernel::initialize methods allocate through CudaArray::create). */
int vvb200_plan_upload(vvb200_plan *plan, void *stream);

/* One whole integrator step with the forces currently in buf->force, for systems without OpenMM
 * constraints / virtual sites between the sub-steps.
 *   middle scheme: VVIntegrator::stepMiddle body after calcForcesAndEnergy (VVIntegrator.cpp:237-268):
 *       resetExtraForce, Langevin/field/cosine forces, firstIntegrate, [bias remove], scaleVelocity,
 *       [bias restore], secondIntegrate (incl. hard wall), updateImagePositions.
 *   Fused into: pass A (extra forces incl. the Langevin force + kick + COM / group-KE / bias
 *   reduction, NH chains advanced by the last block on the device) and pass B (thermostat scaling +
 *   both half drifts + position write + hard wall + image mirror) -- or, for systems of up to
 *   ~120k particles, ONE launch that keeps the state in shared memory across a grid barrier
 *   (vvb200_set_resident_mode).  No host sync.  Thermostat molecules longer than a tile (polymers)
 *   are cut and their centre of mass finished by the last block of pass A. */
int vvb200_step_middle(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);

/* Velocity-Verlet scheme (VVIntegrator::stepVV, VVIntegrator.cpp:272-338), split where OpenMM
 * recomputes forces:
 *   first  = NH half step + firstIntegrate (half kick, drift, hard wall) + updateImagePositions
 *            (:295-310; forces in buf->force are those of the previous positions);
 *   second = extra forces + secondIntegrate (half kick) + NH half step (:316-336; buf->force holds
 *            the forces at the new positions). */
int vvb200_step_vv_first(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);
int vvb200_step_vv_second(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);

/* Multi-GPU split of the middle step (particles partitioned by whole molecules across ranks):
 *   kick_reduce : pass A, leaves this rank's partial sums in a device vector (vvb200_partials_ptr)
 *   -- caller all-reduces that vector (NCCL sum, fp64) --
 *   nhc_scale_drift : NH chains from the reduced sums (redundantly on every rank) + pass B. */
int vvb200_middle_kick_reduce(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);
int vvb200_middle_nhc_scale_drift(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);
/* Alternative to the caller's all-reduce, for one-process-per-GPU runs on an NVLink box: the ranks exchange the
 * reduction vector through cudaIpc-mapped peer memory INSIDE the single-block kernel that advances the NH chains
 * (vvb200_middle_nhc_scale_drift then needs no collective in front of it).  Every rank calls vvb200_peer_export
 * (64-byte cudaIpcMemHandle_t out), the handles are gathered by any means (rank-major, 64 bytes each), every rank
 * calls vvb200_peer_attach.  world <= 8.  All ranks must then step in lockstep (same number of calls). */
int vvb200_peer_export(vvb200_plan *plan, void *handle_out_64_bytes);
int vvb200_peer_attach(vvb200_plan *plan, int rank, int world, const void *handles_rank_major);
/* device pointer to the fp64 reduction vector and its length (<= 16 doubles) */
int vvb200_partials_ptr(vvb200_plan *plan, void **device_ptr, int32_t *num_doubles);
/* Thermostat degrees of freedom / NkbT / eta masses / total mass of the WHOLE system when this
 * plan only holds one rank's partition (defaults: the plan's own). */
int vvb200_set_global_thermostat(vvb200_plan *plan, const double *dof3, double total_mass);

/* The kernel interfaces of VVKernels.h one by one, for the OpenMM glue when constraints or virtual
 * sites sit between the sub-steps (CudaVVKernels.cpp:151,176,214,351,374,427).  Each is one or two
 * launches on `stream`; together they reproduce the fused step exactly.
 *   extra forces: IntegrateMiddleStepKernel::resetExtraForce + ModifyDrudeLangevinKernel::
 *                 applyLangevinForce + ModifyElectricFieldKernel::applyElectricForce +
 *                 ModifyCosineAccelerateKernel::applyCosineForce are folded into the kick. */
int vvb200_middle_kick(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);      /* integrateMiddleVel, CudaVVKernels.cpp:144-148 */
int vvb200_thermostat(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);       /* calc/removeVelocityBias + scaleVelocity + restoreVelocityBias */
/* thermostat + integrateMiddlePos1 + integrateMiddlePos2 fused (one reduction pass + one scaling pass that also
 * writes posDelta = oldDelta = dt/2 v + dt/2 v'): what the glue calls between OpenMM's velocity and position
 * constraints.  kick (88 B) + this (32 + 128 B) + finish (192 B) = 440 B/particle, the constrained-path minimum. */
int vvb200_middle_thermostat_delta(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, void *stream);
/* accumulate == 0: integrateMiddlePos1 (posDelta = oldDelta = dt/2 v, :154-158); != 0: integrateMiddlePos2 (+=, :169-173) */
int vvb200_middle_delta(vvb200_plan *plan, const vvb200_buffers *buf, int accumulate, void *stream);
/* integrateMiddlePos3 + applyHardWallConstraints (CudaVVKernels.cpp:179-212) AND ModifyImageChargeKernel::updateImagePositions
 * (:904-934): images follow their parents inside this call (one launch on the tiled path), so the glue's
 * updateImagePositions is a no-op after it.  vvb200_vv_positions below does the same. */
int vvb200_middle_finish(vvb200_plan *plan, const vvb200_buffers *buf, void *stream);
/* The same for the velocity-Verlet scheme (CudaIntegrateVVStepKernel::firstIntegrate / secondIntegrate,
 * CudaVVKernels.cpp:296-431): velocityVerletIntegrateVelocities (second_half != 0: this step's Langevin force is
 * computed first, like VVIntegrator.cpp:316-325; update_pos_delta != 0: posDelta = dt v) and
 * velocityVerletIntegratePositions + applyHardWallConstraints. */
int vvb200_vv_kick(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args, int second_half,
                   int update_pos_delta, void *stream);
int vvb200_vv_positions(vvb200_plan *plan, const vvb200_buffers *buf, void *stream);
int vvb200_update_image_positions(vvb200_plan *plan, const vvb200_buffers *buf, void *stream);                        /* ModifyImageChargeKernel::updateImagePositions, :904-934 */

/* ---- state / observables (each synchronises `stream`) --------------------------------------- */
typedef struct {
    int32_t num_temp_groups;
    double ke2[3];           /* sum m v^2 per temperature group at the last thermostat call */
    double vscale[3];        /* NH scale factors applied at the last thermostat call */
    double velocity_bias;    /* V of the cosine profile (vMaxBuffer[0]) */
    double eta[3 * VVB200_MAX_CHAINS];
    double eta_dot[3 * (VVB200_MAX_CHAINS + 1)];
    double eta_dotdot[3 * VVB200_MAX_CHAINS];
} vvb200_thermostat_state;
int vvb200_get_thermostat_state(vvb200_plan *plan, vvb200_thermostat_state *out, void *stream);
int vvb200_set_thermostat_state(vvb200_plan *plan, const vvb200_thermostat_state *in, void *stream);
/* ModifyCosineAccelerateKernel::calcViscosity, CudaVVKernels.cpp:1112-1134 (one scalar read back
 * instead of the reference's N-element download) */
int vvb200_calc_viscosity(vvb200_plan *plan, double box_x, double box_y, double box_z,
                          double *v_max, double *inv_vis, void *stream);
/* device copy of comVelm (mixed4 per molecule, CudaVVKernels.cpp:606-617) into host memory */
int vvb200_get_com_velocities(vvb200_plan *plan, void *host_out, void *stream);
/* kernel launches issued by this plan since creation (bench.py's gpu_launches) */
int64_t vvb200_launch_count(const vvb200_plan *plan);

/* ---- checkpoint ------------------------------------------------------------------
 * What Integrator::createCheckpoint / loadCheckpoint need from this path: the Nose-Hoover chain
 * state (the reference keeps eta / eta_dot / eta_dotdot as VVIntegrator members and LOSES them on
 * resume: a restarted run re-equilibrates its thermostat), the last scale factors and velocity
 * bias, and whether the VV scheme's extra forces are valid.  A flat little-endian blob with a
 * magic/version header; load refuses blobs written for another group count or chain length. */
int vvb200_checkpoint_size(const vvb200_plan *plan, int64_t *bytes);
int vvb200_checkpoint_save(vvb200_plan *plan, void *host_out, int64_t capacity, void *stream);
int vvb200_checkpoint_load(vvb200_plan *plan, const void *host_in, int64_t bytes, void *stream);

/* ---- group temperatures on demand ------------------------------------------------
 * Temperatures of the atom / molecular-COM / Drude groups of the CURRENT velocities, without
 * stepping and without advancing the Nose-Hoover chains or changing the scale factors (the scratch
 * the thermostat recomputes every step -- reduction vector, comVelm -- is overwritten): one
 * reduction launch and an 80-byte read back.  A plan that holds one rank's partition reports that
 * partition (its own DOFs and mass); all-reduce ke2 and dof for the whole box.  Replaces the host path of examples/ommhelper/reporter/drudetemperaturereporter.py:98-129
 * (download of N velocities + numpy) with the thermostat's own definitions
 * (drudeNoseHoover.cu:55-114, DOFs of CudaVVKernels.cpp:516-573).  ke2 = sum m v^2 per group
 * (kJ/mol), temperature = ke2 / (dof kB).  Must not be called between vvb200_middle_kick_reduce
 * and vvb200_middle_nhc_scale_drift (it reuses the reduction vector).  Synchronises the stream. */
typedef struct {
    int32_t num_temp_groups;
    double ke2[3], dof[3], temperature[3];
    double velocity_bias;    /* cosine runs: amplitude of the velocity profile, removed before the sums */
} vvb200_temperatures;
int vvb200_measure_temperatures(vvb200_plan *plan, const vvb200_buffers *buf, const vvb200_step_args *args,
                                vvb200_temperatures *out, void *stream);

/* Small systems (all tiles co-resident in shared memory; used up to VVB200_RESIDENT_MAX_PARTICLES =
 * 120,000 particles by default, beyond which the streaming kernels are faster) run
 * the whole thermostatted step -- what CudaVVKernels.cpp:144-185 + 670-754 do in 9-10 launches and
 * a blocking host round trip -- as ONE launch with a grid barrier (csrc/vvb200_resident.cuh).
 * mode: -1 = default (on unless the environment says VVB200_RESIDENT=0), 0 = always use the two
 * streaming passes, 1 = on.  vvb200_resident_launch_count: how many steps took the resident path. */
int vvb200_set_resident_mode(vvb200_plan *plan, int mode);
int64_t vvb200_resident_launch_count(const vvb200_plan *plan);

/* Launch behaviour fixed per plan at vvb200_plan_upload (environment, for tests and tuning; defaults are what ships):
 *   VVB200_PDL=0        no programmatic dependent launch between the two streaming passes
 *   VVB200_HANDOVER=0   pass B waits for the whole of pass A (griddepcontrol.wait) instead of taking over through the
 *                       hand-over word and the factor records its last block publishes (vvb200_stream.cuh); applies to the
 *                       calls that launch both passes: vvb200_step_middle, vvb200_thermostat,
 *                       vvb200_middle_thermostat_delta, vvb200_step_vv_first / _second
 *   VVB200_B_REVERSE=0  under the hand-over pass B draws its tiles first-to-last instead of last-to-first
 * Results are bitwise identical in every combination (tests/test_gpu_edge_cases.py). */

/* Optional per-kernel timing for bench.py's roofline: CUDA events are recorded on the launching stream
 * around the pass-A kernel (kick and / or reductions) and the pass-B kernel (scale / drift / position
 * write) of every step-like call -- vvb200_step_middle, vvb200_step_vv_first / _second, and the split
 * calls vvb200_middle_kick, vvb200_thermostat, vvb200_middle_thermostat_delta, vvb200_middle_finish --
 * for up to max_steps such calls (0 disables).  vvb200_profile_read synchronises on the last event, returns the
 * summed durations in milliseconds and the number of steps they cover, and restarts the sampling. */
int vvb200_profile_enable(vvb200_plan *plan, int max_steps);
int vvb200_profile_read(vvb200_plan *plan, double *ms_pass_a, double *ms_pass_b, int32_t *steps);

/* ---- host-buffer convenience (what Context.setPositions/setVelocities + step + getState do):
 * H2D of posq/posqCorrection/velm/force, `steps` integrator steps with those forces held
 * fixed, D2H of posq/posqCorrection/velm.  Pointers are HOST memory (pinned for full speed).
 * Device staging buffers are owned by the plan. Synchronises. */
int vvb200_step_host(vvb200_plan *plan, const vvb200_buffers *host_buf, const vvb200_step_args *args,
                     int steps, void *stream);
/* For one step of a large system (>= 1M particles) vvb200_step_host overlaps the transfers with the two
 * passes chunk by chunk (copy-in, compute and copy-out streams; PCIe is full duplex).  Molecule-
 * partitioned multi-GPU runs use the same pipeline in two calls around their exchange of the
 * reduction vector: _begin queues all copies in and runs pass A (returns without synchronising);
 * the caller all-reduces vvb200_partials_ptr() unless the peer exchange is attached; _finish
 * advances the NH chains, runs pass B, copies the results out and synchronises. */
int vvb200_step_host_begin(vvb200_plan *plan, const vvb200_buffers *host_buf, const vvb200_step_args *args, void *stream);
int vvb200_step_host_finish(vvb200_plan *plan, const vvb200_buffers *host_buf, const vvb200_step_args *args, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VVB200_H_ */
