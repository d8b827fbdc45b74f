/*
 * vv_oracle.c -- TEST INFRASTRUCTURE ONLY (see vv_oracle.h for the parity status).
 *
 * Plain-C restatement of the reference plugin's integration path.  Every function names the
 * reference file:line it follows (paths relative to /root/reference).  Arithmetic is written
 * in the reference's types (`real`, `mixed`) and operation order so that, up to FMA
 * contraction and reduction order, it reproduces what the reference kernels compute.
 *
 * OpenMM-side facts taken from memory of OpenMM 8.1.2 ([OMM-mem] in SURVEY.md):
 *   BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000, AVOGADRO = 6.02214076e23
 *   SQRT == sqrtf and RECIP(x) == 1.0f/(x) unless CudaPrecision=double (then sqrt, 1.0/(x)).
 */
#include "vv_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef vvo_real real;
typedef vvo_mixed mixed;
typedef vvo_real3 real3;
typedef vvo_real4 real4;
typedef vvo_mixed4 mixed4;

#define VVO_BOLTZ (1.380649e-23 * 6.02214076e23 / 1000.0)
#define VVO_AVOGADRO 6.02214076e23

#if defined(VVO_DOUBLE)
#define SQRT(x) sqrt(x)
#define RECIP(x) (1.0 / (x))
#else
#define SQRT(x) sqrtf(x)
#define RECIP(x) (1.0f / (x))
#endif

/* the constraint stand-in (oracle/constraint_standin.h) in this build's types */
#define VVC_REAL4 vvo_real4
#define VVC_MIXED4 vvo_mixed4
#define VVC_MIXED vvo_mixed
#define VVC_FN static inline
#include "constraint_standin.h"

static char g_err[512];
const char *vvo_last_error(void) { return g_err; }

int vvo_precision_mode(void) {
#if defined(VVO_SINGLE)
    return 0;
#elif defined(VVO_DOUBLE)
    return 2;
#else
    return 1;
#endif
}

void vvo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n > 0 ? n : omp_get_num_procs());   /* n <= 0: all host cores */
#else
    (void) n;
#endif
}

int vvo_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

struct vvo_ctx {
    vvo_params par;
    int numAtoms, paddedNumAtoms, numMolecules;
    int literal;
    double *masses;
    /* VVIntegrator members (VVIntegrator.h:467-491) */
    int32_t *particleMolId;
    double *moleculeMasses, *moleculeInvMasses;
    int nNH; int32_t *particlesNH;
    int nMolNH; int32_t *moleculesNH;
    int nLD; int32_t *particlesLD;
    int nImg; vvo_int2 *imagePairs; int32_t *particlesImage;
    int nEl; int32_t *particlesElectrolyte;
    /* lookup marks for the non-literal path */
    unsigned char *markNH, *markLD, *markImage;
    /* step kernel (CudaVVKernels.cpp:66-96) */
    int nDrude; vvo_int2 *drudePairs;
    real3 *forceExtra;
    mixed4 *oldDelta;
    /* NH kernel (CudaVVKernels.cpp:462-667) */
    int32_t *particlesSortedByMolId; vvo_int2 *particlesInMolecules;
    int nNormalNH; int32_t *normalParticlesNH;
    int nPairsNH; vvo_int2 *pairParticlesNH;
    double tempGroupDof[VVO_NUM_TG_MAX];
    int numTempGroup;
    double etaMass[VVO_NUM_TG_MAX][VVO_MAX_CHAINS];
    double eta[VVO_NUM_TG_MAX][VVO_MAX_CHAINS];
    double etaDot[VVO_NUM_TG_MAX][VVO_MAX_CHAINS + 1];
    double etaDotDot[VVO_NUM_TG_MAX][VVO_MAX_CHAINS];
    double tempGroupNkbT[VVO_NUM_TG_MAX];
    mixed4 *comVelm;
    double ke2[VVO_NUM_TG_MAX];      /* kineticEnergiesNHVec */
    double vscale[VVO_NUM_TG_MAX];   /* vscaleFactorsNHVec */
    /* LD kernel */
    int nNormalLD; int32_t *normalParticlesLD;
    int nPairsLD; vvo_int2 *pairParticlesLD;
    /* cosine kernel */
    double invMassTotal;
    double vBias;                    /* vMaxBuffer[0] after sumV */
    /* flat copies for vvo_get_array */
    double flatEtaMass[VVO_NUM_TG_MAX * VVO_MAX_CHAINS];
    double flatEta[VVO_NUM_TG_MAX * VVO_MAX_CHAINS];
    double flatEtaDot[VVO_NUM_TG_MAX * (VVO_MAX_CHAINS + 1)];
    double flatEtaDotDot[VVO_NUM_TG_MAX * VVO_MAX_CHAINS];
    /* stand-in for OpenMM's constraint solvers (owning copies of the cluster tables) */
    vvc_constraints cons;
    int32_t *consOffset, *consAtoms;
    double *consDistance;
};

/* ------------------------------------------------------------------------------------ */
/* Molecules: OpenMM ContextImpl::getMolecules / findMolecules / tagParticlesInMolecule   */
/* [OMM-mem].  Particles are visited ascending; an untagged particle opens molecule       */
/* numMolecules++ and tags everything reachable over bonds.  Hence molecule ids ascend    */
/* with the first atom and atoms inside a molecule are listed ascending.                  */
/* ------------------------------------------------------------------------------------ */
int vvo_find_molecules(int n, int numBonds, const int32_t *bonds, int32_t *molId) {
    int *deg = (int *) calloc((size_t) n + 1, sizeof(int));
    for (int b = 0; b < numBonds; b++) {
        deg[bonds[2 * b]]++;
        deg[bonds[2 * b + 1]]++;
    }
    int *start = (int *) malloc(((size_t) n + 1) * sizeof(int));
    start[0] = 0;
    for (int i = 0; i < n; i++)
        start[i + 1] = start[i] + deg[i];
    int *fill = (int *) calloc((size_t) n + 1, sizeof(int));
    int *adj = (int *) malloc(((size_t) 2 * numBonds + 1) * sizeof(int));
    for (int b = 0; b < numBonds; b++) {
        int p = bonds[2 * b], q = bonds[2 * b + 1];
        adj[start[p] + fill[p]++] = q;
        adj[start[q] + fill[q]++] = p;
    }
    for (int i = 0; i < n; i++)
        molId[i] = -1;
    int *stack = (int *) malloc(((size_t) n + 1) * sizeof(int));
    int numMolecules = 0;
    for (int i = 0; i < n; i++) {
        if (molId[i] != -1)
            continue;
        int sp = 0;
        stack[sp++] = i;
        molId[i] = numMolecules;
        while (sp > 0) {
            int p = stack[--sp];
            for (int k = start[p]; k < start[p + 1]; k++) {
                int q = adj[k];
                if (molId[q] == -1) {
                    molId[q] = numMolecules;
                    stack[sp++] = q;
                }
            }
        }
        numMolecules++;
    }
    free(deg); free(start); free(fill); free(adj); free(stack);
    return numMolecules;
}

/* ------------------------------------------------------------------------------------ */
/* Membership predicates: VVIntegrator.h:326-343 are std::find over the vectors.          */
/* ------------------------------------------------------------------------------------ */
static int find_int(const int32_t *v, int n, int x) {
    for (int i = 0; i < n; i++)
        if (v[i] == x)
            return 1;
    return 0;
}
static int isParticleLD(const vvo_ctx *c, int i) {
    return c->literal ? find_int(c->particlesLD, c->nLD, i) : c->markLD[i];
}
static int isParticleImage(const vvo_ctx *c, int i) {
    return c->literal ? find_int(c->particlesImage, c->nImg, i) : c->markImage[i];
}
static int isParticleNH(const vvo_ctx *c, int i) {
    return c->literal ? find_int(c->particlesNH, c->nNH, i) : c->markNH[i];
}

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n > 0 ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "vv_oracle: out of memory\n");
        abort();
    }
    return p;
}

void vvo_destroy(vvo_ctx *c) {
    if (!c)
        return;
    free(c->masses); free(c->particleMolId); free(c->moleculeMasses); free(c->moleculeInvMasses);
    free(c->particlesNH); free(c->moleculesNH); free(c->particlesLD); free(c->imagePairs);
    free(c->particlesImage); free(c->particlesElectrolyte); free(c->markNH); free(c->markLD);
    free(c->markImage); free(c->drudePairs); free(c->forceExtra); free(c->oldDelta);
    free(c->particlesSortedByMolId); free(c->particlesInMolecules); free(c->normalParticlesNH);
    free(c->pairParticlesNH); free(c->comVelm); free(c->normalParticlesLD); free(c->pairParticlesLD);
    free(c->consOffset); free(c->consAtoms); free(c->consDistance);
    free(c);
}

vvo_ctx *vvo_create(const vvo_system *sys, const vvo_params *par, int literal) {
    g_err[0] = 0;
    if (par->numNHChains > VVO_MAX_CHAINS || par->numNHChains < 1) {
        snprintf(g_err, sizeof g_err, "numNHChains out of range");
        return NULL;
    }
    vvo_ctx *c = (vvo_ctx *) xcalloc(1, sizeof(vvo_ctx));
    const int N = sys->numParticles;
    c->par = *par;
    c->literal = literal;
    c->numAtoms = N;
    c->paddedNumAtoms = sys->paddedNumAtoms;
    c->numMolecules = sys->numMolecules;
    c->masses = (double *) xcalloc(N, sizeof(double));
    memcpy(c->masses, sys->masses, (size_t) N * sizeof(double));

    /* user-populated lists (VVIntegrator.h:199-202, 302-305; VVIntegrator.cpp:76-80) */
    c->nLD = sys->numLD;
    c->particlesLD = (int32_t *) xcalloc(c->nLD, sizeof(int32_t));
    if (c->nLD) memcpy(c->particlesLD, sys->particlesLD, (size_t) c->nLD * sizeof(int32_t));
    c->nImg = sys->numImagePairs;
    c->imagePairs = (vvo_int2 *) xcalloc(c->nImg, sizeof(vvo_int2));
    c->particlesImage = (int32_t *) xcalloc(c->nImg, sizeof(int32_t));
    for (int k = 0; k < c->nImg; k++) {
        c->imagePairs[k].x = sys->imagePairs[2 * k];
        c->imagePairs[k].y = sys->imagePairs[2 * k + 1];
        c->particlesImage[k] = sys->imagePairs[2 * k];
    }
    c->nEl = sys->numElectrolyte;
    c->particlesElectrolyte = (int32_t *) xcalloc(c->nEl, sizeof(int32_t));
    if (c->nEl) memcpy(c->particlesElectrolyte, sys->particlesElectrolyte, (size_t) c->nEl * sizeof(int32_t));

    c->markLD = (unsigned char *) xcalloc(N, 1);
    c->markImage = (unsigned char *) xcalloc(N, 1);
    c->markNH = (unsigned char *) xcalloc(N, 1);
    for (int k = 0; k < c->nLD; k++) c->markLD[c->particlesLD[k]] = 1;
    for (int k = 0; k < c->nImg; k++) c->markImage[c->particlesImage[k]] = 1;

    /* VVIntegrator.cpp:123-135: particleMolId, moleculeMasses (particle order), inverses */
    c->particleMolId = (int32_t *) xcalloc(N, sizeof(int32_t));
    memcpy(c->particleMolId, sys->particleMolId, (size_t) N * sizeof(int32_t));
    const int M = sys->numMolecules;
    c->moleculeMasses = (double *) xcalloc(M, sizeof(double));
    c->moleculeInvMasses = (double *) xcalloc(M, sizeof(double));
    for (int i = 0; i < N; i++)
        c->moleculeMasses[c->particleMolId[i]] += sys->masses[i];
    for (int i = 0; i < M; i++)
        c->moleculeInvMasses[i] = 1.0 / c->moleculeMasses[i];

    /* VVIntegrator.cpp:138-145: particlesNH ascending, moleculesNH in first-seen order */
    c->particlesNH = (int32_t *) xcalloc(N, sizeof(int32_t));
    c->moleculesNH = (int32_t *) xcalloc(M, sizeof(int32_t));
    unsigned char *molSeen = (unsigned char *) xcalloc(M, 1);
    for (int i = 0; i < N; i++) {
        if (!isParticleLD(c, i) && !isParticleImage(c, i)) {
            c->particlesNH[c->nNH++] = i;
            c->markNH[i] = 1;
            int mol = c->particleMolId[i];
            int seen = literal ? find_int(c->moleculesNH, c->nMolNH, mol) : molSeen[mol];
            if (!seen) {
                c->moleculesNH[c->nMolNH++] = mol;
                molSeen[mol] = 1;
            }
        }
    }
    /* VVIntegrator.cpp:146-151 */
    for (int i = 0; i < N; i++) {
        if (isParticleLD(c, i)) {
            int mol = c->particleMolId[i];
            int inNH = literal ? find_int(c->moleculesNH, c->nMolNH, mol) : molSeen[mol];
            if (inNH) {
                snprintf(g_err, sizeof g_err, "NH and Langevin thermostat cannot be applied on the same molecule");
                free(molSeen);
                vvo_destroy(c);
                return NULL;
            }
        }
    }
    free(molSeen);
    /* VVIntegrator.cpp:154-155 */
    if (c->nLD > 0 && par->cosAcceleration != 0) {
        snprintf(g_err, sizeof g_err, "Langevin thermostat and periodic perturbation shouldn't be used together");
        vvo_destroy(c);
        return NULL;
    }

    /* Step kernel: CudaVVKernels.cpp:67-96 */
    c->nDrude = sys->numDrude;
    c->drudePairs = (vvo_int2 *) xcalloc(c->nDrude, sizeof(vvo_int2));
    for (int k = 0; k < c->nDrude; k++) {
        c->drudePairs[k].x = sys->drudePairs[2 * k];
        c->drudePairs[k].y = sys->drudePairs[2 * k + 1];
    }
    c->forceExtra = (real3 *) xcalloc(N, sizeof(real3));
    c->oldDelta = (mixed4 *) xcalloc(N, sizeof(mixed4));

    /* NH kernel: CudaVVKernels.cpp:483-494 (particlesSortedByMolId, particlesInMolecules) */
    c->particlesSortedByMolId = (int32_t *) xcalloc(N, sizeof(int32_t));
    c->particlesInMolecules = (vvo_int2 *) xcalloc(M, sizeof(vvo_int2));
    if (literal) {
        int id_start = 0, fill = 0;
        for (int id_mol = 0; id_mol < M; id_mol++) {
            int n_in_mol = 0;
            for (int i = 0; i < N; i++) {
                if (c->particleMolId[i] == id_mol) {
                    n_in_mol++;
                    c->particlesSortedByMolId[fill++] = i;
                }
            }
            c->particlesInMolecules[id_mol].x = n_in_mol;
            c->particlesInMolecules[id_mol].y = id_start;
            id_start += n_in_mol;
        }
    } else {
        for (int i = 0; i < N; i++)
            c->particlesInMolecules[c->particleMolId[i]].x++;
        int id_start = 0;
        for (int m = 0; m < M; m++) {
            c->particlesInMolecules[m].y = id_start;
            id_start += c->particlesInMolecules[m].x;
        }
        int *cursor = (int *) xcalloc(M, sizeof(int));
        for (int i = 0; i < N; i++) {
            int m = c->particleMolId[i];
            c->particlesSortedByMolId[c->particlesInMolecules[m].y + cursor[m]++] = i;
        }
        free(cursor);
    }

    /* CudaVVKernels.cpp:496-511: particlesNHSet + atomic DOFs in particle order */
    for (int g = 0; g < VVO_NUM_TG_MAX; g++)
        c->tempGroupDof[g] = 0.0;
    unsigned char *inSet = (unsigned char *) xcalloc(N, 1);
    for (int i = 0; i < N; i++) {
        if (isParticleNH(c, i))
            inSet[i] = 1;
        int id_mol = c->particleMolId[i];
        double mass = sys->masses[i];
        double molInvMass = c->moleculeInvMasses[id_mol];
        if (isParticleNH(c, i) && mass != 0.0) {
            c->tempGroupDof[VVO_TG_ATOM] += 3;
            if (par->useCOMTempGroup)
                c->tempGroupDof[VVO_TG_ATOM] -= 3 * mass * molInvMass;
        }
    }
    /* CudaVVKernels.cpp:513-529 */
    c->pairParticlesNH = (vvo_int2 *) xcalloc(c->nDrude, sizeof(vvo_int2));
    for (int k = 0; k < c->nDrude; k++) {
        int p = c->drudePairs[k].x, p1 = c->drudePairs[k].y;
        if (isParticleNH(c, p) != isParticleNH(c, p1)) {
            snprintf(g_err, sizeof g_err, "Drude particle and its parent atom should be in the same thermostat");
            free(inSet);
            vvo_destroy(c);
            return NULL;
        }
        if (isParticleNH(c, p)) {
            inSet[p] = 0;
            inSet[p1] = 0;
            c->pairParticlesNH[c->nPairsNH].x = p;
            c->pairParticlesNH[c->nPairsNH].y = p1;
            c->nPairsNH++;
            c->tempGroupDof[VVO_TG_ATOM] -= 3;
            c->tempGroupDof[VVO_TG_DRUDE] += 3;
        }
    }
    c->normalParticlesNH = (int32_t *) xcalloc(N, sizeof(int32_t));
    for (int i = 0; i < N; i++)
        if (inSet[i])
            c->normalParticlesNH[c->nNormalNH++] = i;
    /* CudaVVKernels.cpp:531-541 */
    for (int k = 0; k < sys->numConstraints; k++) {
        int p = sys->constraints[2 * k], p1 = sys->constraints[2 * k + 1];
        if (isParticleNH(c, p) != isParticleNH(c, p1)) {
            snprintf(g_err, sizeof g_err, "Constrained particle pair should be in the same thermostat");
            free(inSet);
            vvo_destroy(c);
            return NULL;
        }
        if (isParticleNH(c, p))
            c->tempGroupDof[VVO_TG_ATOM] -= 1;
    }
    /* CudaVVKernels.cpp:547-564 */
    if (par->useCOMTempGroup)
        c->tempGroupDof[VVO_TG_COM] = 3 * c->nMolNH;
    if (sys->hasCMMotionRemover) {
        if (par->useCOMTempGroup)
            c->tempGroupDof[VVO_TG_COM] -= 3;
        else
            c->tempGroupDof[VVO_TG_ATOM] -= 3;
    }
    for (int g = 0; g < VVO_NUM_TG_MAX; g++)
        c->tempGroupDof[g] = c->tempGroupDof[g] > 0 ? c->tempGroupDof[g] : 0;
    /* CudaVVKernels.cpp:567-573 */
    c->numTempGroup = 3;
    if (c->tempGroupDof[VVO_TG_DRUDE] == 0) {
        c->numTempGroup = 2;
        if (c->tempGroupDof[VVO_TG_COM] == 0)
            c->numTempGroup = 1;
    }
    /* CudaVVKernels.cpp:577-594 */
    double realKbT = VVO_BOLTZ * par->temperature;
    double drudeKbT = VVO_BOLTZ * par->drudeTemperature;
    for (int g = 0; g < c->numTempGroup; g++) {
        double tgKbT = g == VVO_TG_DRUDE ? drudeKbT : realKbT;
        double tgMass = g == VVO_TG_DRUDE ? drudeKbT / pow(par->drudeFrequency, 2)
                                          : realKbT / pow(par->frequency, 2);
        c->tempGroupNkbT[g] = c->tempGroupDof[g] * tgKbT;
        c->etaMass[g][0] = c->tempGroupDof[g] * tgMass;
        for (int ich = 1; ich < par->numNHChains; ich++)
            c->etaMass[g][ich] = tgMass;
    }
    /* CudaVVKernels.cpp:606-617: comVelm zero-initialised */
    c->comVelm = (mixed4 *) xcalloc(M, sizeof(mixed4));
    for (int g = 0; g < VVO_NUM_TG_MAX; g++)
        c->vscale[g] = 1.0;

    /* LD kernel: CudaVVKernels.cpp:775-804 */
    memset(inSet, 0, (size_t) N);
    if (c->nLD > 0) {
        for (int i = 0; i < N; i++)
            if (isParticleLD(c, i))
                inSet[i] = 1;
        c->pairParticlesLD = (vvo_int2 *) xcalloc(c->nDrude, sizeof(vvo_int2));
        for (int k = 0; k < c->nDrude; k++) {
            int p = c->drudePairs[k].x, p1 = c->drudePairs[k].y;
            if (isParticleLD(c, p) != isParticleLD(c, p1)) {
                snprintf(g_err, sizeof g_err, "Drude particle and its parent atom should be in the same thermostat");
                free(inSet);
                vvo_destroy(c);
                return NULL;
            }
            if (isParticleLD(c, p)) {
                inSet[p] = 0;
                inSet[p1] = 0;
                c->pairParticlesLD[c->nPairsLD].x = p;
                c->pairParticlesLD[c->nPairsLD].y = p1;
                c->nPairsLD++;
            }
        }
        for (int k = 0; k < sys->numConstraints; k++) {
            int p = sys->constraints[2 * k], p1 = sys->constraints[2 * k + 1];
            if (isParticleLD(c, p) != isParticleLD(c, p1)) {
                snprintf(g_err, sizeof g_err, "Constrained particle pair should be in the same thermostat");
                free(inSet);
                vvo_destroy(c);
                return NULL;
            }
        }
        c->normalParticlesLD = (int32_t *) xcalloc(N, sizeof(int32_t));
        for (int i = 0; i < N; i++)
            if (inSet[i])
                c->normalParticlesLD[c->nNormalLD++] = i;
    }
    free(inSet);

    /* cosine kernel: CudaVVKernels.cpp:1028-1031 */
    double massTotal = 0;
    for (int i = 0; i < N; i++)
        massTotal += sys->masses[i];
    c->invMassTotal = 1.0 / massTotal;
    return c;
}

int vvo_num_temp_groups(const vvo_ctx *c) { return c->numTempGroup; }

int64_t vvo_get_array(const vvo_ctx *cc, int id, const void **ptr) {
    vvo_ctx *c = (vvo_ctx *) cc;
    const int nc = c->par.numNHChains;
    switch (id) {
    case VVO_ARR_PARTICLES_NH: *ptr = c->particlesNH; return c->nNH;
    case VVO_ARR_MOLECULES_NH: *ptr = c->moleculesNH; return c->nMolNH;
    case VVO_ARR_PARTICLE_MOL_ID: *ptr = c->particleMolId; return c->numAtoms;
    case VVO_ARR_DRUDE_PAIRS: *ptr = c->drudePairs; return 2 * (int64_t) c->nDrude;
    case VVO_ARR_SORTED_BY_MOL: *ptr = c->particlesSortedByMolId; return c->numAtoms;
    case VVO_ARR_PARTICLES_IN_MOLECULES: *ptr = c->particlesInMolecules; return 2 * (int64_t) c->numMolecules;
    case VVO_ARR_NORMAL_NH: *ptr = c->normalParticlesNH; return c->nNormalNH;
    case VVO_ARR_PAIRS_NH: *ptr = c->pairParticlesNH; return 2 * (int64_t) c->nPairsNH;
    case VVO_ARR_NORMAL_LD: *ptr = c->normalParticlesLD; return c->nNormalLD;
    case VVO_ARR_PAIRS_LD: *ptr = c->pairParticlesLD; return 2 * (int64_t) c->nPairsLD;
    case VVO_ARR_IMAGE_PAIRS: *ptr = c->imagePairs; return 2 * (int64_t) c->nImg;
    case VVO_ARR_ELECTROLYTE: *ptr = c->particlesElectrolyte; return c->nEl;
    case VVO_ARR_MOLECULE_MASSES: *ptr = c->moleculeMasses; return c->numMolecules;
    case VVO_ARR_MOLECULE_INV_MASSES: *ptr = c->moleculeInvMasses; return c->numMolecules;
    case VVO_ARR_DOF: *ptr = c->tempGroupDof; return VVO_NUM_TG_MAX;
    case VVO_ARR_NKBT: *ptr = c->tempGroupNkbT; return c->numTempGroup;
    case VVO_ARR_INV_MASS_TOTAL: *ptr = &c->invMassTotal; return 1;
    case VVO_ARR_KE2: *ptr = c->ke2; return c->numTempGroup;
    case VVO_ARR_VSCALE: *ptr = c->vscale; return c->numTempGroup;
    case VVO_ARR_VBIAS: *ptr = &c->vBias; return 1;
    case VVO_ARR_ETA_MASS:
        for (int g = 0; g < c->numTempGroup; g++)
            for (int k = 0; k < nc; k++) c->flatEtaMass[g * nc + k] = c->etaMass[g][k];
        *ptr = c->flatEtaMass; return (int64_t) c->numTempGroup * nc;
    case VVO_ARR_ETA:
        for (int g = 0; g < c->numTempGroup; g++)
            for (int k = 0; k < nc; k++) c->flatEta[g * nc + k] = c->eta[g][k];
        *ptr = c->flatEta; return (int64_t) c->numTempGroup * nc;
    case VVO_ARR_ETA_DOT:
        for (int g = 0; g < c->numTempGroup; g++)
            for (int k = 0; k < nc + 1; k++) c->flatEtaDot[g * (nc + 1) + k] = c->etaDot[g][k];
        *ptr = c->flatEtaDot; return (int64_t) c->numTempGroup * (nc + 1);
    case VVO_ARR_ETA_DOTDOT:
        for (int g = 0; g < c->numTempGroup; g++)
            for (int k = 0; k < nc; k++) c->flatEtaDotDot[g * nc + k] = c->etaDotDot[g][k];
        *ptr = c->flatEtaDotDot; return (int64_t) c->numTempGroup * nc;
    default: *ptr = NULL; return -1;
    }
}

void vvo_set_nhc_state(vvo_ctx *c, const double *eta, const double *etaDot, const double *etaDotDot) {
    const int nc = c->par.numNHChains;
    for (int g = 0; g < c->numTempGroup; g++) {
        for (int k = 0; k < nc; k++) {
            c->eta[g][k] = eta[g * nc + k];
            c->etaDotDot[g][k] = etaDotDot[g * nc + k];
        }
        for (int k = 0; k < nc + 1; k++)
            c->etaDot[g][k] = etaDot[g * (nc + 1) + k];
    }
}

/* ------------------------------------------------------------------------------------ */
/* VVIntegrator::propagateNHChain -- VVIntegrator.cpp:340-376                              */
/* ------------------------------------------------------------------------------------ */
void vvo_propagate_nh_chain(double stepSize, int loopsPerStep, int numNHChains,
                            double *eta, double *eta_dot, double *eta_dotdot, const double *eta_mass,
                            double ke2, double ke2_target, double t_target, double *factorOut) {
    double expfac = 0;
    double dt2 = stepSize / loopsPerStep / 2;
    double dt4 = dt2 / 2;
    double dt8 = dt4 / 2;
    double factor = 1.0;
    eta_dotdot[0] = (ke2 - ke2_target) / eta_mass[0];
    for (int iloop = 0; iloop < loopsPerStep; iloop++) {
        for (int ich = numNHChains - 1; ich >= 0; ich--) {
            expfac = exp(-dt8 * eta_dot[ich + 1]);
            eta_dot[ich] *= expfac;
            eta_dot[ich] += eta_dotdot[ich] * dt4;
            eta_dot[ich] *= expfac;
        }
        factor *= exp(-dt2 * eta_dot[0]);
        for (int ich = 0; ich < numNHChains; ich++)
            eta[ich] += dt2 * eta_dot[ich];
        eta_dotdot[0] = (ke2 * factor * factor - ke2_target) / eta_mass[0];
        eta_dot[0] *= expfac;              /* reuses the last expfac, as in the reference */
        eta_dot[0] += eta_dotdot[0] * dt4;
        eta_dot[0] *= expfac;
        for (int ich = 1; ich < numNHChains; ich++) {
            expfac = exp(-dt8 * eta_dot[ich + 1]);
            eta_dot[ich] *= expfac;
            eta_dotdot[ich] = (eta_mass[ich - 1] * eta_dot[ich - 1] * eta_dot[ich - 1]
                               - VVO_BOLTZ * t_target) / eta_mass[ich];
            eta_dot[ich] += eta_dotdot[ich] * dt4;
            eta_dot[ich] *= expfac;
        }
    }
    *factorOut = factor;
}

/* ------------------------------------------------------------------------------------ */
/* Extra forces                                                                           */
/* ------------------------------------------------------------------------------------ */

/* resetExtraForce -- middle.cu:227-231 / velocityVerlet.cu:195-199 */
void vvo_reset_extra_force(vvo_ctx *c) {
    const int N = c->numAtoms;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; i++) {
        c->forceExtra[i].x = 0;
        c->forceExtra[i].y = 0;
        c->forceExtra[i].z = 0;
    }
}

/* addExtraForceDrudeLangevin -- drudeLangevin.cu:2-59, host factors CudaVVKernels.cpp:835-859 */
void vvo_langevin_force(vvo_ctx *c, const vvo_buffers *b, unsigned randomIndex) {
    const double stepSize = c->par.stepSize;
    const double dragD = c->par.friction;
    const double randD = sqrt(2.0 * VVO_BOLTZ * c->par.temperature * dragD / stepSize);
    const double dragDrudeD = c->par.drudeFriction;
    const double randDrudeD = sqrt(2.0 * VVO_BOLTZ * c->par.drudeTemperature * dragDrudeD / stepSize);
    const mixed dragFactor = (mixed) dragD, randFactor = (mixed) randD;
    const mixed dragFactorDrude = (mixed) dragDrudeD, randFactorDrude = (mixed) randDrudeD;
    const mixed4 *velm = b->velm;
    real3 *forceExtra = c->forceExtra;
    const vvo_float4 *random = b->random;

#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->nNormalLD; i++) {
        int index = c->normalParticlesLD[i];
        mixed4 velocity = velm[index];
        if (velocity.w != 0) {
            mixed mass = RECIP(velocity.w);
            mixed sqrtMass = SQRT(mass);
            vvo_float4 rand = random[randomIndex + i];
            forceExtra[index].x += (-dragFactor * mass * velocity.x + randFactor * sqrtMass * rand.x);
            forceExtra[index].y += (-dragFactor * mass * velocity.y + randFactor * sqrtMass * rand.y);
            forceExtra[index].z += (-dragFactor * mass * velocity.z + randFactor * sqrtMass * rand.z);
        }
    }
    randomIndex += c->nNormalLD;
    /* pairs touch two slots each; a particle belongs to at most one pair, so this is race-free */
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->nPairsLD; i++) {
        vvo_int2 particles = c->pairParticlesLD[i];
        mixed4 velocity1 = velm[particles.x];
        mixed4 velocity2 = velm[particles.y];
        mixed mass1 = RECIP(velocity1.w);
        mixed mass2 = RECIP(velocity2.w);
        mixed totMass = mass1 + mass2;
        mixed sqrtTotMass = SQRT(totMass);
        mixed redMass = RECIP((mass1 + mass2) * velocity1.w * velocity2.w);
        mixed sqrtRedMass = SQRT(redMass);
        mixed invTotMass = RECIP(totMass);
        mixed mass1fract = invTotMass * mass1;
        mixed mass2fract = invTotMass * mass2;
        mixed cmx = velocity1.x * mass1fract + velocity2.x * mass2fract;
        mixed cmy = velocity1.y * mass1fract + velocity2.y * mass2fract;
        mixed cmz = velocity1.z * mass1fract + velocity2.z * mass2fract;
        mixed relx = velocity2.x - velocity1.x;
        mixed rely = velocity2.y - velocity1.y;
        mixed relz = velocity2.z - velocity1.z;
        vvo_float4 rand1 = random[randomIndex + 2 * i];
        vvo_float4 rand2 = random[randomIndex + 2 * i + 1];
        real3 cmForce, relForce;   /* real3: rounded to float unless double mode */
        cmForce.x = (real) (-dragFactor * totMass * cmx + randFactor * sqrtTotMass * rand1.x);
        cmForce.y = (real) (-dragFactor * totMass * cmy + randFactor * sqrtTotMass * rand1.y);
        cmForce.z = (real) (-dragFactor * totMass * cmz + randFactor * sqrtTotMass * rand1.z);
        relForce.x = (real) (-dragFactorDrude * redMass * relx + randFactorDrude * sqrtRedMass * rand2.x);
        relForce.y = (real) (-dragFactorDrude * redMass * rely + randFactorDrude * sqrtRedMass * rand2.y);
        relForce.z = (real) (-dragFactorDrude * redMass * relz + randFactorDrude * sqrtRedMass * rand2.z);
        /* vectorOps: real3 * mixed -> operator*(float3,float)/(double3,double): the scalar is
         * converted to `real` at the call (float3*double picks operator*(float3,float)). */
        real f1 = (real) mass1fract, f2 = (real) mass2fract;
        real3 a, d;
        a.x = f1 * cmForce.x - relForce.x; a.y = f1 * cmForce.y - relForce.y; a.z = f1 * cmForce.z - relForce.z;
        d.x = f2 * cmForce.x + relForce.x; d.y = f2 * cmForce.y + relForce.y; d.z = f2 * cmForce.z + relForce.z;
        forceExtra[particles.x].x += a.x; forceExtra[particles.x].y += a.y; forceExtra[particles.x].z += a.z;
        forceExtra[particles.y].x += d.x; forceExtra[particles.y].y += d.y; forceExtra[particles.y].z += d.z;
    }
}

/* addExtraForceElectricField -- electricField.cu:2-11; efscale CudaVVKernels.cpp:978-985
 * (a `real`: float unless double mode).  Duplicated electrolyte entries add twice, so the
 * loop stays serial in list order. */
void vvo_electric_force(vvo_ctx *c, const vvo_buffers *b) {
    double efscaleD = c->par.electricField * VVO_AVOGADRO;
    real efscale = (real) efscaleD;
    for (int i = 0; i < c->nEl; i++) {
        int index = c->particlesElectrolyte[i];
        real charge = b->posq[index].w;
        c->forceExtra[index].z += efscale * charge;
    }
}

/* cos(2*3.1415926*z*invBox.z): the literal and the double-precision cos are the reference's
 * (cosineAccelerate.cu:9,26,70,83). */
static inline double cosPhase(real z, real invBoxZ) {
    return cos(2 * 3.1415926 * z * invBoxZ);
}

/* addCosAcceleration -- cosineAccelerate.cu:2-14; acceleration is a `real` (CudaVVKernels.cpp:1044-1051) */
void vvo_cosine_force(vvo_ctx *c, const vvo_buffers *b, double invBoxZ) {
    const real acceleration = (real) c->par.cosAcceleration;
    const real ibz = (real) invBoxZ;
    const int N = c->numAtoms;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++)
        c->forceExtra[index].x += acceleration * cosPhase(b->posq[index].z, ibz) * RECIP(b->velm[index].w);
}

/* ------------------------------------------------------------------------------------ */
/* Middle scheme kernels -- middle.cu                                                      */
/* ------------------------------------------------------------------------------------ */

/* integrateMiddleVel -- middle.cu:6-23 */
void vvo_middle_vel(vvo_ctx *c, const vvo_buffers *b) {
    const int N = c->numAtoms, P = c->paddedNumAtoms;
    const mixed stepSize = (mixed) c->par.stepSize;
    const mixed fscale = stepSize / (mixed) 0x100000000;
    mixed4 *velm = b->velm;
    const long long *force = b->force;
    const real3 *forceExtra = c->forceExtra;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 velocity = velm[index];
        if (velocity.w != 0) {
            velocity.x += stepSize * velocity.w * forceExtra[index].x + fscale * velocity.w * force[index];
            velocity.y += stepSize * velocity.w * forceExtra[index].y + fscale * velocity.w * force[index + P];
            velocity.z += stepSize * velocity.w * forceExtra[index].z + fscale * velocity.w * force[index + P * 2];
            velm[index] = velocity;
        }
    }
}

/* integrateMiddlePos1 -- middle.cu:29-43 */
void vvo_middle_pos1(vvo_ctx *c, const vvo_buffers *b) {
    const int N = c->numAtoms;
    const mixed halfdt = 0.5f * (mixed) c->par.stepSize;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 velocity = b->velm[index];
        if (velocity.w != 0) {
            mixed4 delta = { halfdt * velocity.x, halfdt * velocity.y, halfdt * velocity.z, 0 };
            b->posDelta[index] = delta;
            c->oldDelta[index] = delta;
        }
    }
}

/* integrateMiddlePos2 -- middle.cu:47-61 */
void vvo_middle_pos2(vvo_ctx *c, const vvo_buffers *b) {
    const int N = c->numAtoms;
    const mixed halfdt = 0.5f * (mixed) c->par.stepSize;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 velocity = b->velm[index];
        if (velocity.w != 0) {
            mixed4 delta = { halfdt * velocity.x, halfdt * velocity.y, halfdt * velocity.z, 0 };
            b->posDelta[index].x += delta.x; b->posDelta[index].y += delta.y;
            b->posDelta[index].z += delta.z; b->posDelta[index].w += delta.w;
            c->oldDelta[index].x += delta.x; c->oldDelta[index].y += delta.y;
            c->oldDelta[index].z += delta.z; c->oldDelta[index].w += delta.w;
        }
    }
}

/* integrateMiddlePos3 -- middle.cu:66-100 */
void vvo_middle_pos3(vvo_ctx *c, const vvo_buffers *b) {
    const int N = c->numAtoms;
    const mixed invDt = 1 / (mixed) c->par.stepSize;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 velocity = b->velm[index];
        if (velocity.w != 0.0) {
            mixed4 delta = b->posDelta[index];
            velocity.x += (delta.x - c->oldDelta[index].x) * invDt;
            velocity.y += (delta.y - c->oldDelta[index].y) * invDt;
            velocity.z += (delta.z - c->oldDelta[index].z) * invDt;
            b->velm[index] = velocity;
#if defined(VVO_MIXED)
            real4 pos1 = b->posq[index];
            real4 pos2 = b->posqCorrection[index];
            mixed4 pos = { pos1.x + (mixed) pos2.x, pos1.y + (mixed) pos2.y, pos1.z + (mixed) pos2.z, pos1.w };
#else
            real4 pos = b->posq[index];
#endif
            pos.x += delta.x;
            pos.y += delta.y;
            pos.z += delta.z;
#if defined(VVO_MIXED)
            real4 o = { (real) pos.x, (real) pos.y, (real) pos.z, (real) pos.w };
            real4 oc = { (real) (pos.x - (real) pos.x), (real) (pos.y - (real) pos.y), (real) (pos.z - (real) pos.z), 0 };
            b->posq[index] = o;
            b->posqCorrection[index] = oc;
#else
            b->posq[index] = pos;
#endif
        }
    }
}

/* applyHardWallConstraints -- middle.cu:106-221 (== velocityVerlet.cu:74-190);
 * host scalars CudaVVKernels.cpp:189-200 */
void vvo_hard_wall(vvo_ctx *c, const vvo_buffers *b) {
    if (!(c->par.maxDrudeDistance > 0) || c->nDrude == 0)
        return;
    const mixed stepSize = (mixed) c->par.stepSize;
    const mixed maxDrudeDistance = (mixed) c->par.maxDrudeDistance;
    const mixed hardwallscaleDrude = (mixed) sqrt(VVO_BOLTZ * c->par.drudeTemperature);
    real4 *posq = b->posq;
    mixed4 *velm = b->velm;
#if defined(VVO_MIXED)
    real4 *posqCorrection = b->posqCorrection;
#endif
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->nDrude; i++) {
        vvo_int2 particles = c->drudePairs[i];
#if defined(VVO_MIXED)
        real4 posReal1 = posq[particles.x];
        real4 posReal2 = posq[particles.y];
        real4 posCorr1 = posqCorrection[particles.x];
        real4 posCorr2 = posqCorrection[particles.y];
        mixed4 pos1 = { posReal1.x + (mixed) posCorr1.x, posReal1.y + (mixed) posCorr1.y, posReal1.z + (mixed) posCorr1.z, posReal1.w };
        mixed4 pos2 = { posReal2.x + (mixed) posCorr2.x, posReal2.y + (mixed) posCorr2.y, posReal2.z + (mixed) posCorr2.z, posReal2.w };
#else
        mixed4 pos1 = { posq[particles.x].x, posq[particles.x].y, posq[particles.x].z, posq[particles.x].w };
        mixed4 pos2 = { posq[particles.y].x, posq[particles.y].y, posq[particles.y].z, posq[particles.y].w };
#endif
        mixed4 delta = { pos1.x - pos2.x, pos1.y - pos2.y, pos1.z - pos2.z, pos1.w - pos2.w };
        mixed r = SQRT(delta.x * delta.x + delta.y * delta.y + delta.z * delta.z);
        mixed rInv = RECIP(r);
        if (rInv * maxDrudeDistance < 1) {
            mixed4 bondDir = { delta.x * rInv, delta.y * rInv, delta.z * rInv, delta.w * rInv };
            mixed4 vel1 = velm[particles.x];
            mixed4 vel2 = velm[particles.y];
            mixed mass1 = RECIP(vel1.w);
            mixed mass2 = RECIP(vel2.w);
            mixed deltaR = r - maxDrudeDistance;
            mixed deltaT = stepSize;
            mixed dotvr1 = vel1.x * bondDir.x + vel1.y * bondDir.y + vel1.z * bondDir.z;
            mixed4 vb1 = { bondDir.x * dotvr1, bondDir.y * dotvr1, bondDir.z * dotvr1, 0 };
            mixed4 vp1 = { vel1.x - vb1.x, vel1.y - vb1.y, vel1.z - vb1.z, 0 };
            if (vel2.w == 0) {
                if (dotvr1 != 0)
                    deltaT = deltaR / fabs(dotvr1);
                if (deltaT > stepSize)
                    deltaT = stepSize;
                dotvr1 = -dotvr1 * hardwallscaleDrude / (fabs(dotvr1) * SQRT(mass1));
                mixed dr = -deltaR + deltaT * dotvr1;
                pos1.x += bondDir.x * dr;
                pos1.y += bondDir.y * dr;
                pos1.z += bondDir.z * dr;
#if defined(VVO_MIXED)
                real4 o = { (real) pos1.x, (real) pos1.y, (real) pos1.z, (real) pos1.w };
                real4 oc = { (real) (pos1.x - (real) pos1.x), (real) (pos1.y - (real) pos1.y), (real) (pos1.z - (real) pos1.z), 0 };
                posq[particles.x] = o;
                posqCorrection[particles.x] = oc;
#else
                real4 o = { (real) pos1.x, (real) pos1.y, (real) pos1.z, (real) pos1.w };
                posq[particles.x] = o;
#endif
                vel1.x = vp1.x + bondDir.x * dotvr1;
                vel1.y = vp1.y + bondDir.y * dotvr1;
                vel1.z = vp1.z + bondDir.z * dotvr1;
                velm[particles.x] = vel1;
            } else {
                mixed invTotalMass = RECIP(mass1 + mass2);
                mixed dotvr2 = vel2.x * bondDir.x + vel2.y * bondDir.y + vel2.z * bondDir.z;
                mixed4 vb2 = { bondDir.x * dotvr2, bondDir.y * dotvr2, bondDir.z * dotvr2, 0 };
                mixed4 vp2 = { vel2.x - vb2.x, vel2.y - vb2.y, vel2.z - vb2.z, 0 };
                mixed vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
                dotvr1 -= vbCMass;
                dotvr2 -= vbCMass;
                if (dotvr1 != dotvr2)
                    deltaT = deltaR / fabs(dotvr1 - dotvr2);
                if (deltaT > stepSize)
                    deltaT = stepSize;
                mixed vBond = hardwallscaleDrude / SQRT(mass1);
                dotvr1 = -dotvr1 * vBond * mass2 * invTotalMass / fabs(dotvr1);
                dotvr2 = -dotvr2 * vBond * mass1 * invTotalMass / fabs(dotvr2);
                mixed dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
                mixed dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
                dotvr1 += vbCMass;
                dotvr2 += vbCMass;
                pos1.x += bondDir.x * dr1;
                pos1.y += bondDir.y * dr1;
                pos1.z += bondDir.z * dr1;
                pos2.x += bondDir.x * dr2;
                pos2.y += bondDir.y * dr2;
                pos2.z += bondDir.z * dr2;
#if defined(VVO_MIXED)
                real4 o1 = { (real) pos1.x, (real) pos1.y, (real) pos1.z, (real) pos1.w };
                real4 o2 = { (real) pos2.x, (real) pos2.y, (real) pos2.z, (real) pos2.w };
                real4 c1 = { (real) (pos1.x - (real) pos1.x), (real) (pos1.y - (real) pos1.y), (real) (pos1.z - (real) pos1.z), 0 };
                real4 c2 = { (real) (pos2.x - (real) pos2.x), (real) (pos2.y - (real) pos2.y), (real) (pos2.z - (real) pos2.z), 0 };
                posq[particles.x] = o1;
                posq[particles.y] = o2;
                posqCorrection[particles.x] = c1;
                posqCorrection[particles.y] = c2;
#else
                real4 o1 = { (real) pos1.x, (real) pos1.y, (real) pos1.z, (real) pos1.w };
                real4 o2 = { (real) pos2.x, (real) pos2.y, (real) pos2.z, (real) pos2.w };
                posq[particles.x] = o1;
                posq[particles.y] = o2;
#endif
                vel1.x = vp1.x + bondDir.x * dotvr1;
                vel1.y = vp1.y + bondDir.y * dotvr1;
                vel1.z = vp1.z + bondDir.z * dotvr1;
                vel2.x = vp2.x + bondDir.x * dotvr2;
                vel2.y = vp2.y + bondDir.y * dotvr2;
                vel2.z = vp2.z + bondDir.z * dotvr2;
                velm[particles.x] = vel1;
                velm[particles.y] = vel2;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Velocity-Verlet kernels -- velocityVerlet.cu                                            */
/* ------------------------------------------------------------------------------------ */

/* velocityVerletIntegrateVelocities -- velocityVerlet.cu:6-29; fscale from host
 * CudaVVKernels.cpp:306,405 (double 0.5*dt/2^32, cast to float in single mode) */
void vvo_vv_velocities(vvo_ctx *c, const vvo_buffers *b, int updatePosDelta) {
    const int N = c->numAtoms, P = c->paddedNumAtoms;
    const mixed stepSize = (mixed) c->par.stepSize;
    const double fscaleD = 0.5 * c->par.stepSize / (double) 0x100000000;
    const mixed fscale = (mixed) fscaleD;
    mixed4 *velm = b->velm;
    const long long *force = b->force;
    const real3 *forceExtra = c->forceExtra;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 velocity = velm[index];
        if (velocity.w != 0) {
            /* `0.5 * stepSize * ...` : the double literal promotes the product to double */
            velocity.x += 0.5 * stepSize * velocity.w * forceExtra[index].x + fscale * velocity.w * force[index];
            velocity.y += 0.5 * stepSize * velocity.w * forceExtra[index].y + fscale * velocity.w * force[index + P];
            velocity.z += 0.5 * stepSize * velocity.w * forceExtra[index].z + fscale * velocity.w * force[index + P * 2];
            velm[index] = velocity;
            if (updatePosDelta) {
                mixed4 d = { stepSize * velocity.x, stepSize * velocity.y, stepSize * velocity.z, 0 };
                b->posDelta[index] = d;
            }
        }
    }
}

/* velocityVerletIntegratePositions -- velocityVerlet.cu:35-68 */
void vvo_vv_positions(vvo_ctx *c, const vvo_buffers *b) {
    const int N = c->numAtoms;
    const mixed invStepSize = 1.0 / (mixed) c->par.stepSize;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++) {
        mixed4 vel = b->velm[index];
        if (vel.w != 0) {
#if defined(VVO_MIXED)
            real4 pos1 = b->posq[index];
            real4 pos2 = b->posqCorrection[index];
            mixed4 pos = { pos1.x + (mixed) pos2.x, pos1.y + (mixed) pos2.y, pos1.z + (mixed) pos2.z, pos1.w };
#else
            real4 pos = b->posq[index];
#endif
            mixed4 delta = b->posDelta[index];
            pos.x += delta.x;
            pos.y += delta.y;
            pos.z += delta.z;
            vel.x = (mixed) (invStepSize * delta.x);
            vel.y = (mixed) (invStepSize * delta.y);
            vel.z = (mixed) (invStepSize * delta.z);
#if defined(VVO_MIXED)
            real4 o = { (real) pos.x, (real) pos.y, (real) pos.z, (real) pos.w };
            real4 oc = { (real) (pos.x - (real) pos.x), (real) (pos.y - (real) pos.y), (real) (pos.z - (real) pos.z), 0 };
            b->posq[index] = o;
            b->posqCorrection[index] = oc;
#else
            b->posq[index] = pos;
#endif
            b->velm[index] = vel;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Cosine velocity bias -- cosineAccelerate.cu:16-84                                       */
/* ------------------------------------------------------------------------------------ */

/* calcPeriodicVelocityBias (:16-32) + sumV (:34-61).  The sum runs in particle order in
 * `mixed`; the reference's 512-thread tree gives the same value up to reassociation. */
void vvo_calc_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ) {
    const int N = c->numAtoms;
    const real ibz = (real) invBoxZ;
    mixed total = 0;
    for (int index = 0; index < N; index++) {
        mixed v;
        if (b->velm[index].w == 0)
            v = 0;
        else
            v = (mixed) (RECIP(b->velm[index].w) * b->velm[index].x * 2 * cosPhase(b->posq[index].z, ibz));
        total += v;
    }
    c->vBias = (double) (mixed) (total * c->invMassTotal);
}

/* removePeriodicVelocityBias -- cosineAccelerate.cu:63-74 */
void vvo_remove_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ) {
    const int N = c->numAtoms;
    const real ibz = (real) invBoxZ;
    const mixed V = (mixed) c->vBias;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++)
        b->velm[index].x -= V * cosPhase(b->posq[index].z, ibz);
}

/* restorePeriodicVelocityBias -- cosineAccelerate.cu:76-85 */
void vvo_restore_velocity_bias(vvo_ctx *c, const vvo_buffers *b, double invBoxZ) {
    const int N = c->numAtoms;
    const real ibz = (real) invBoxZ;
    const mixed V = (mixed) c->vBias;
#pragma omp parallel for schedule(static)
    for (int index = 0; index < N; index++)
        b->velm[index].x += V * cosPhase(b->posq[index].z, ibz);
}

/* calcViscosity -- CudaVVKernels.cpp:1112-1134 */
void vvo_calc_viscosity(vvo_ctx *c, double boxX, double boxY, double boxZ, double *vMax, double *invVis) {
    *vMax = c->vBias;
    double vol = boxX * boxY * boxZ;
    *invVis = *vMax * vol * c->invMassTotal / c->par.cosAcceleration
              * (2 * 3.1415926 / boxZ) * (2 * 3.1415926 / boxZ);
}

/* ------------------------------------------------------------------------------------ */
/* Thermostat -- CudaModifyDrudeNoseKernel::scaleVelocity, CudaVVKernels.cpp:670-754,      */
/* kernels drudeNoseHoover.cu                                                              */
/* ------------------------------------------------------------------------------------ */
void vvo_scale_velocity(vvo_ctx *c, const vvo_buffers *b) {
    mixed4 *velm = b->velm;
    mixed4 *comVelm = c->comVelm;
    const int numTG = c->numTempGroup;

    if (c->par.useCOMTempGroup) {
        /* calcCOMVelocities -- drudeNoseHoover.cu:5-31 */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < c->nMolNH; i++) {
            int id_mol = c->moleculesNH[i];
            mixed4 acc = { 0, 0, 0, 0 };
            mixed comMass = 0.0;
            for (int j = 0; j < c->particlesInMolecules[id_mol].x; j++) {
                int index = c->particlesSortedByMolId[c->particlesInMolecules[id_mol].y + j];
                mixed4 velocity = velm[index];
                if (velocity.w != 0) {
                    mixed mass = RECIP(velocity.w);
                    acc.x += velocity.x * mass;
                    acc.y += velocity.y * mass;
                    acc.z += velocity.z * mass;
                    comMass += mass;
                }
            }
            acc.w = RECIP(comMass);
            acc.x *= acc.w;
            acc.y *= acc.w;
            acc.z *= acc.w;
            comVelm[id_mol] = acc;
        }
        /* normalizeVelocities -- drudeNoseHoover.cu:37-49 */
#pragma omp parallel for schedule(static)
        for (int i = 0; i < c->nNH; i++) {
            int index = c->particlesNH[i];
            int id_mol = c->particleMolId[index];
            velm[index].x -= comVelm[id_mol].x;
            velm[index].y -= comVelm[id_mol].y;
            velm[index].z -= comVelm[id_mol].z;
        }
    }

    /* computeNormalizedKineticEnergies (:55-115) + sumNormalizedKineticEnergies (:121-151).
     * Accumulated in `mixed` in list order (the reference's per-thread partials + tree sum
     * give the same values up to reassociation). */
    mixed ke[VVO_NUM_TG_MAX] = { 0, 0, 0 };
    for (int i = 0; i < c->nNormalNH; i++) {
        int index = c->normalParticlesNH[i];
        mixed4 velocity = velm[index];
        if (velocity.w != 0)
            ke[VVO_TG_ATOM] += (velocity.x * velocity.x + velocity.y * velocity.y + velocity.z * velocity.z) / velocity.w;
    }
    if (numTG > VVO_TG_COM) {
        for (int i = 0; i < c->nMolNH; i++) {
            int id_mol = c->moleculesNH[i];
            mixed4 velocity = comVelm[id_mol];
            if (velocity.w != 0)
                ke[VVO_TG_COM] += (velocity.x * velocity.x + velocity.y * velocity.y + velocity.z * velocity.z) / velocity.w;
        }
    }
    for (int i = 0; i < c->nPairsNH; i++) {
        vvo_int2 pair = c->pairParticlesNH[i];
        mixed4 velocity1 = velm[pair.x];
        mixed4 velocity2 = velm[pair.y];
        mixed mass1 = RECIP(velocity1.w);
        mixed mass2 = RECIP(velocity2.w);
        mixed invTotalMass = RECIP(mass1 + mass2);
        mixed invReducedMass = (mass1 + mass2) * velocity1.w * velocity2.w;
        mixed mass1fract = invTotalMass * mass1;
        mixed mass2fract = invTotalMass * mass2;
        mixed cmx = velocity1.x * mass1fract + velocity2.x * mass2fract;
        mixed cmy = velocity1.y * mass1fract + velocity2.y * mass2fract;
        mixed cmz = velocity1.z * mass1fract + velocity2.z * mass2fract;
        mixed relx = velocity1.x - velocity2.x;
        mixed rely = velocity1.y - velocity2.y;
        mixed relz = velocity1.z - velocity2.z;
        ke[VVO_TG_ATOM] += (cmx * cmx + cmy * cmy + cmz * cmz) * (mass1 + mass2);
        ke[VVO_TG_DRUDE] += (relx * relx + rely * rely + relz * relz) / invReducedMass;
    }
    for (int g = 0; g < numTG; g++)
        c->ke2[g] = (double) ke[g];

    /* host NHC loop -- CudaVVKernels.cpp:726-733 */
    for (int g = 0; g < VVO_NUM_TG_MAX; g++)
        c->vscale[g] = 1.0;
    for (int itg = 0; itg < numTG; itg++) {
        const double T = itg == VVO_TG_DRUDE ? c->par.drudeTemperature : c->par.temperature;
        if (c->etaMass[itg][0] > 0)
            vvo_propagate_nh_chain(c->par.stepSize, c->par.loopsPerStep, c->par.numNHChains,
                                   c->eta[itg], c->etaDot[itg], c->etaDotDot[itg], c->etaMass[itg],
                                   c->ke2[itg], c->tempGroupNkbT[itg], T, &c->vscale[itg]);
    }

    /* scaleVelocity -- drudeNoseHoover.cu:157-208.  Factors beyond numTempGroup are read out
     * of bounds by the reference and only ever multiply zeros; 1 here. */
    const mixed vscaleAtom = (mixed) c->vscale[0];
    const mixed vscaleCOM = (mixed) c->vscale[1];
    const mixed vscaleDrude = (mixed) c->vscale[2];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->nNormalNH; i++) {
        int index = c->normalParticlesNH[i];
        int id_mol = c->particleMolId[index];
        mixed4 velCOM = comVelm[id_mol];
        if (velm[index].w != 0) {
            velm[index].x = vscaleAtom * velm[index].x + vscaleCOM * velCOM.x;
            velm[index].y = vscaleAtom * velm[index].y + vscaleCOM * velCOM.y;
            velm[index].z = vscaleAtom * velm[index].z + vscaleCOM * velCOM.z;
        }
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < c->nPairsNH; i++) {
        vvo_int2 particles = c->pairParticlesNH[i];
        int id_mol = c->particleMolId[particles.x];
        mixed4 velAtom1 = velm[particles.x];
        mixed4 velAtom2 = velm[particles.y];
        mixed4 velCOM = comVelm[id_mol];
        mixed mass1 = RECIP(velAtom1.w);
        mixed mass2 = RECIP(velAtom2.w);
        mixed invTotalMass = RECIP(mass1 + mass2);
        mixed mass1fract = invTotalMass * mass1;
        mixed mass2fract = invTotalMass * mass2;
        mixed cmx = velAtom1.x * mass1fract + velAtom2.x * mass2fract;
        mixed cmy = velAtom1.y * mass1fract + velAtom2.y * mass2fract;
        mixed cmz = velAtom1.z * mass1fract + velAtom2.z * mass2fract;
        mixed relx = velAtom2.x - velAtom1.x;
        mixed rely = velAtom2.y - velAtom1.y;
        mixed relz = velAtom2.z - velAtom1.z;
        cmx = vscaleAtom * cmx; cmy = vscaleAtom * cmy; cmz = vscaleAtom * cmz;
        relx = vscaleDrude * relx; rely = vscaleDrude * rely; relz = vscaleDrude * relz;
        velAtom1.x = cmx - relx * mass2fract + vscaleCOM * velCOM.x;
        velAtom1.y = cmy - rely * mass2fract + vscaleCOM * velCOM.y;
        velAtom1.z = cmz - relz * mass2fract + vscaleCOM * velCOM.z;
        velAtom2.x = cmx + relx * mass1fract + vscaleCOM * velCOM.x;
        velAtom2.y = cmy + rely * mass1fract + vscaleCOM * velCOM.y;
        velAtom2.z = cmz + relz * mass1fract + vscaleCOM * velCOM.z;
        velm[particles.x] = velAtom1;
        velm[particles.y] = velAtom2;
    }
}

/* updateImagePositions -- imageCharge.cu:2-27; mirror CudaVVKernels.cpp:918-926.
 * Only the mixed mode is well defined in the reference (the other modes
 * dereference a null posqCorrection); there we apply the non-mixed branch to posq only. */
void vvo_update_images(vvo_ctx *c, const vvo_buffers *b) {
    const mixed mirror = (mixed) c->par.mirrorLocation;
    for (int i = 0; i < c->nImg; i++) {
        int index_img = c->imagePairs[i].x;
        int index_par = c->imagePairs[i].y;
        b->posq[index_img].x = b->posq[index_par].x;
        b->posq[index_img].y = b->posq[index_par].y;
#if defined(VVO_MIXED)
        b->posqCorrection[index_img].x = b->posqCorrection[index_par].x;
        b->posqCorrection[index_img].y = b->posqCorrection[index_par].y;
        mixed z = (mixed) b->posq[index_par].z + (mixed) b->posqCorrection[index_par].z;
        z = mirror * 2 - z;
        b->posq[index_img].z = (real) z;
        b->posqCorrection[index_img].z = (real) (z - (real) z);
#else
        b->posq[index_img].z = 2 * mirror - b->posq[index_par].z;
#endif
    }
}

/* ------------------------------------------------------------------------------------ */
/* Step schedules                                                                         */
/* ------------------------------------------------------------------------------------ */

static unsigned prepareRandomNumbers(const vvo_ctx *c, unsigned *randomIndex) {
    /* CudaVVKernels.cpp:863: request uses the padded (>=1) array sizes */
    unsigned request = (unsigned) ((c->nNormalLD > 1 ? c->nNormalLD : 1) + 2 * (c->nPairsLD > 1 ? c->nPairsLD : 1));
    unsigned old = *randomIndex;
    *randomIndex += request;
    return old;
}

static void nhHalf(vvo_ctx *c, const vvo_buffers *b, double invBoxZ) {
    /* VVIntegrator.cpp:251-260 (== :295-304, :327-336) */
    if (c->nNH > 0) {
        if (c->par.cosAcceleration != 0) {
            vvo_calc_velocity_bias(c, b, invBoxZ);
            vvo_remove_velocity_bias(c, b, invBoxZ);
        }
        vvo_scale_velocity(c, b);
        if (c->par.cosAcceleration != 0)
            vvo_restore_velocity_bias(c, b, invBoxZ);
    }
}

static void extraForces(vvo_ctx *c, const vvo_buffers *b, double invBoxZ, unsigned *randomIndex) {
    /* VVIntegrator.cpp:238-245 (== :316-323) */
    if (c->nLD > 0 || c->nEl > 0 || c->par.cosAcceleration != 0)
        vvo_reset_extra_force(c);
    if (c->nLD > 0)
        vvo_langevin_force(c, b, prepareRandomNumbers(c, randomIndex));
    if (c->nEl > 0)
        vvo_electric_force(c, b);
    if (c->par.cosAcceleration != 0)
        vvo_cosine_force(c, b, invBoxZ);
}

/* Stand-in for OpenMM's constraint solvers between the sub-steps (oracle/constraint_standin.h): cluster tables built
 * by the caller.  numClusters == 0 switches it off (then the calls below are no-ops, like OpenMM's for a System
 * without constraints). */
void vvo_set_constraint_standin(vvo_ctx *c, int numClusters, const int32_t *clusterOffset, const int32_t *atoms,
                                const double *distance, int iterations) {
    free(c->consOffset); free(c->consAtoms); free(c->consDistance);
    c->consOffset = c->consAtoms = NULL;
    c->consDistance = NULL;
    memset(&c->cons, 0, sizeof c->cons);
    if (numClusters <= 0)
        return;
    const int nCons = clusterOffset[numClusters];
    c->consOffset = (int32_t *) xcalloc((size_t) numClusters + 1, sizeof(int32_t));
    c->consAtoms = (int32_t *) xcalloc((size_t) 2 * nCons + 1, sizeof(int32_t));
    c->consDistance = (double *) xcalloc((size_t) nCons + 1, sizeof(double));
    memcpy(c->consOffset, clusterOffset, ((size_t) numClusters + 1) * sizeof(int32_t));
    memcpy(c->consAtoms, atoms, (size_t) 2 * nCons * sizeof(int32_t));
    memcpy(c->consDistance, distance, (size_t) nCons * sizeof(double));
    c->cons.numClusters = numClusters;
    c->cons.iterations = iterations;
    c->cons.clusterOffset = c->consOffset;
    c->cons.atoms = c->consAtoms;
    c->cons.distance = c->consDistance;
}

/* integration.applyConstraints(tol), CudaVVKernels.cpp:176, 351 */
void vvo_apply_constraints(vvo_ctx *c, const vvo_buffers *b) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < c->cons.numClusters; k++)
        vvc_cluster_positions(c->cons, k, b->posq, b->posqCorrection, b->velm, b->posDelta);
}

/* integration.applyVelocityConstraints(tol), CudaVVKernels.cpp:151, 427 */
void vvo_apply_velocity_constraints(vvo_ctx *c, const vvo_buffers *b) {
#pragma omp parallel for schedule(static)
    for (int k = 0; k < c->cons.numClusters; k++)
        vvc_cluster_velocities(c->cons, k, b->posq, b->posqCorrection, b->velm);
}

void vvo_step(vvo_ctx *c, const vvo_buffers *b, int steps, double invBoxZ, unsigned *randomIndex) {
    for (int s = 0; s < steps; s++) {
        if (c->par.useMiddleScheme) {
            /* VVIntegrator::stepMiddle, VVIntegrator.cpp:232-270; forces are frozen (b->force) */
            extraForces(c, b, invBoxZ, randomIndex);
            /* firstIntegrate, CudaVVKernels.cpp:129-159 */
            vvo_middle_vel(c, b);
            vvo_apply_velocity_constraints(c, b);
            vvo_middle_pos1(c, b);
            nhHalf(c, b, invBoxZ);
            /* secondIntegrate, CudaVVKernels.cpp:161-220 */
            vvo_middle_pos2(c, b);
            vvo_apply_constraints(c, b);
            vvo_middle_pos3(c, b);
            vvo_hard_wall(c, b);
            if (c->nImg > 0)
                vvo_update_images(c, b);
        } else {
            /* VVIntegrator::stepVV, VVIntegrator.cpp:272-338 */
            nhHalf(c, b, invBoxZ);
            /* firstIntegrate, CudaVVKernels.cpp:296-382 */
            vvo_vv_velocities(c, b, 1);
            vvo_apply_constraints(c, b);
            vvo_vv_positions(c, b);
            vvo_hard_wall(c, b);
            if (c->nImg > 0)
                vvo_update_images(c, b);
            /* forces at the new positions: frozen here */
            extraForces(c, b, invBoxZ, randomIndex);
            /* secondIntegrate, CudaVVKernels.cpp:395-431 */
            vvo_vv_velocities(c, b, 0);
            vvo_apply_velocity_constraints(c, b);
            nhHalf(c, b, invBoxZ);
        }
    }
}

/* Toy force field for long statistical runs (test-only definition, see header). Each
 * contribution is converted to 2^32 fixed point and added as an integer, like OpenMM's
 * force accumulation, so the result does not depend on summation order. */
void vvo_toy_forces(vvo_ctx *c, const vvo_buffers *b, const double *x0, double kTether, double kDrude) {
    const int N = c->numAtoms, P = c->paddedNumAtoms;
    unsigned char *isDrude = (unsigned char *) xcalloc(N, 1);
    for (int k = 0; k < c->nDrude; k++)
        isDrude[c->drudePairs[k].x] = 1;
    for (int i = 0; i < N; i++) {
        double f[3] = { 0, 0, 0 };
        if (b->velm[i].w != 0 && !isDrude[i]) {
            double x[3];
#if defined(VVO_MIXED)
            x[0] = (double) b->posq[i].x + (double) b->posqCorrection[i].x;
            x[1] = (double) b->posq[i].y + (double) b->posqCorrection[i].y;
            x[2] = (double) b->posq[i].z + (double) b->posqCorrection[i].z;
#else
            x[0] = b->posq[i].x; x[1] = b->posq[i].y; x[2] = b->posq[i].z;
#endif
            for (int d = 0; d < 3; d++)
                f[d] = -kTether * (x[d] - x0[3 * i + d]);
        }
        for (int d = 0; d < 3; d++)
            b->force[i + d * (long long) P] = (long long) (f[d] * 4294967296.0);
    }
    for (int k = 0; k < c->nDrude; k++) {
        int p = c->drudePairs[k].x, q = c->drudePairs[k].y;
        double xp[3], xq[3];
#if defined(VVO_MIXED)
        xp[0] = (double) b->posq[p].x + (double) b->posqCorrection[p].x;
        xp[1] = (double) b->posq[p].y + (double) b->posqCorrection[p].y;
        xp[2] = (double) b->posq[p].z + (double) b->posqCorrection[p].z;
        xq[0] = (double) b->posq[q].x + (double) b->posqCorrection[q].x;
        xq[1] = (double) b->posq[q].y + (double) b->posqCorrection[q].y;
        xq[2] = (double) b->posq[q].z + (double) b->posqCorrection[q].z;
#else
        xp[0] = b->posq[p].x; xp[1] = b->posq[p].y; xp[2] = b->posq[p].z;
        xq[0] = b->posq[q].x; xq[1] = b->posq[q].y; xq[2] = b->posq[q].z;
#endif
        for (int d = 0; d < 3; d++) {
            long long fd = (long long) (-kDrude * (xp[d] - xq[d]) * 4294967296.0);
            b->force[p + d * (long long) P] += fd;
            b->force[q + d * (long long) P] -= fd;
        }
    }
    free(isDrude);
}
