// vvb200_plan.cpp -- host-side index builders of the integration path, O(N).
//
// Produces, bit for bit, the arrays the reference builds with O(N*M) / O(N*N_LD) std::find loops:
//   VVIntegrator::initialize                    VVIntegrator.cpp:123-155
//   CudaIntegrate{Middle,VV}StepKernel::initialize   CudaVVKernels.cpp:66-77, 253-260
//   CudaModifyDrudeNoseKernel::initialize       CudaVVKernels.cpp:483-594
//   CudaModifyDrudeLangevinKernel::initialize   CudaVVKernels.cpp:775-804
//   CudaModifyImageChargeKernel::initialize     CudaVVKernels.cpp:884-891
//   CudaModifyElectricFieldKernel::initialize   CudaVVKernels.cpp:954-957
//   CudaModifyCosineAccelerateKernel::initialize CudaVVKernels.cpp:1028-1031
// plus the tables only the fused sm_100a kernels need (molecule-aligned tiles, the packed slot
// word).  No CUDA call in this file: it is testable on a machine without a GPU.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "vvb200_internal.h"

static thread_local char g_error[1024] = "";

void vvb200_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}

static const double BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0;   // SimTKOpenMMRealType.h [OMM-mem]

extern "C" const char *vvb200_last_error(void) { return g_error; }
extern "C" int vvb200_version(void) { return VVB200_VERSION; }

extern "C" int vvb200_find_molecules(int32_t n, int32_t numBonds, const int32_t *bonds, int32_t *molId,
                                     int32_t *numMolecules) {
    if (n < 0 || numBonds < 0 || (numBonds > 0 && !bonds) || (n > 0 && !molId)) {
        vvb200_set_error("vvb200_find_molecules: invalid argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    // union-find with "smallest index is the root", then number roots in ascending order:
    // the same labelling OpenMM's depth-first tagging yields (molecules ordered by first atom).
    std::vector<int32_t> root(n);
    std::iota(root.begin(), root.end(), 0);
    auto find = [&](int32_t x) {
        while (root[x] != x) {
            root[x] = root[root[x]];
            x = root[x];
        }
        return x;
    };
    for (int32_t b = 0; b < numBonds; b++) {
        int32_t p = bonds[2 * b], q = bonds[2 * b + 1];
        if (p < 0 || p >= n || q < 0 || q >= n) {
            vvb200_set_error("vvb200_find_molecules: bond %d references particle out of range", b);
            return VVB200_ERR_INVALID_ARGUMENT;
        }
        int32_t rp = find(p), rq = find(q);
        if (rp < rq) root[rq] = rp;
        else if (rq < rp) root[rp] = rq;
    }
    int32_t count = 0;
    for (int32_t i = 0; i < n; i++) {
        int32_t r = find(i);
        molId[i] = (r == i) ? count++ : molId[r];   // r < i, already labelled
    }
    if (numMolecules) *numMolecules = count;
    return VVB200_OK;
}

extern "C" int vvb200_propagate_nh_chain(double stepSize, int loops, int nc, double *eta, double *etaDot,
                                         double *etaDotDot, const double *etaMass, double ke2, double ke2Target,
                                         double tTarget, double *factorOut) {
    if (nc < 1 || loops < 1 || !eta || !etaDot || !etaDotDot || !etaMass || !factorOut) {
        vvb200_set_error("vvb200_propagate_nh_chain: invalid argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    // Same recurrences as VVIntegrator::propagateNHChain (VVIntegrator.cpp:340-376), including
    // the reuse of the innermost exp factor for thermostat 0 after the scale update.
    const double h2 = stepSize / loops / 2, h4 = h2 / 2, h8 = h4 / 2;
    double s = 1.0, e = 0.0;
    etaDotDot[0] = (ke2 - ke2Target) / etaMass[0];
    for (int l = 0; l < loops; l++) {
        for (int k = nc - 1; k >= 0; k--) {
            e = std::exp(-h8 * etaDot[k + 1]);
            etaDot[k] *= e;
            etaDot[k] += etaDotDot[k] * h4;
            etaDot[k] *= e;
        }
        s *= std::exp(-h2 * etaDot[0]);
        for (int k = 0; k < nc; k++)
            eta[k] += h2 * etaDot[k];
        etaDotDot[0] = (ke2 * s * s - ke2Target) / etaMass[0];
        etaDot[0] *= e;
        etaDot[0] += etaDotDot[0] * h4;
        etaDot[0] *= e;
        for (int k = 1; k < nc; k++) {
            e = std::exp(-h8 * etaDot[k + 1]);
            etaDot[k] *= e;
            etaDotDot[k] = (etaMass[k - 1] * etaDot[k - 1] * etaDot[k - 1] - BOLTZ * tTarget) / etaMass[k];
            etaDot[k] += etaDotDot[k] * h4;
            etaDot[k] *= e;
        }
    }
    *factorOut = s;
    return VVB200_OK;
}

static int conflict(vvb200_plan *p, const char *msg) {
    vvb200_set_error("%s", msg);
    delete p;
    return VVB200_ERR_CONFLICT;
}

static bool buildTiles(vvb200_plan *p);
static bool buildPlainTiles(vvb200_plan *p);

extern "C" int vvb200_plan_create(const vvb200_system *sys, const vvb200_params *par, int precision,
                                  vvb200_plan **out) {
    if (!sys || !par || !out) {
        vvb200_set_error("vvb200_plan_create: null argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    *out = nullptr;
    const int N = sys->num_particles, M = sys->num_molecules;
    if (N <= 0 || M <= 0 || sys->padded_num_atoms < N || sys->padded_num_atoms % 4 != 0 || !sys->masses || !sys->particle_mol_id ||
        precision < VVB200_SINGLE || precision > VVB200_DOUBLE || par->num_nh_chains < 1 ||
        par->num_nh_chains > VVB200_MAX_CHAINS || par->loops_per_step < 1 ||
        (sys->num_drude > 0 && !sys->drude_pairs) || (sys->num_constraints > 0 && !sys->constraints) ||
        (sys->num_langevin > 0 && !sys->particles_langevin) || (sys->num_image_pairs > 0 && !sys->image_pairs) ||
        (sys->num_electrolyte > 0 && !sys->particles_electrolyte)) {
        vvb200_set_error("vvb200_plan_create: invalid system or parameters");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    auto inRange = [N](int32_t i) { return i >= 0 && i < N; };
    for (int i = 0; i < N; i++)
        if (sys->particle_mol_id[i] < 0 || sys->particle_mol_id[i] >= M) {
            vvb200_set_error("vvb200_plan_create: particle_mol_id[%d] out of range", i);
            return VVB200_ERR_INVALID_ARGUMENT;
        }
    for (int k = 0; k < 2 * sys->num_drude; k++)
        if (!inRange(sys->drude_pairs[k])) { vvb200_set_error("drude_pairs out of range"); return VVB200_ERR_INVALID_ARGUMENT; }
    for (int k = 0; k < 2 * sys->num_constraints; k++)
        if (!inRange(sys->constraints[k])) { vvb200_set_error("constraints out of range"); return VVB200_ERR_INVALID_ARGUMENT; }
    for (int k = 0; k < sys->num_langevin; k++)
        if (!inRange(sys->particles_langevin[k])) { vvb200_set_error("particles_langevin out of range"); return VVB200_ERR_INVALID_ARGUMENT; }
    for (int k = 0; k < 2 * sys->num_image_pairs; k++)
        if (!inRange(sys->image_pairs[k])) { vvb200_set_error("image_pairs out of range"); return VVB200_ERR_INVALID_ARGUMENT; }
    for (int k = 0; k < sys->num_electrolyte; k++)
        if (!inRange(sys->particles_electrolyte[k])) { vvb200_set_error("particles_electrolyte out of range"); return VVB200_ERR_INVALID_ARGUMENT; }

    vvb200_plan *p = new vvb200_plan();
    p->precision = precision;
    p->par = *par;
    p->N = N;
    p->paddedN = sys->padded_num_atoms;
    p->M = M;
    p->hasCMMotionRemover = sys->has_cm_motion_remover != 0;
    p->masses.assign(sys->masses, sys->masses + N);
    p->particleMolId.assign(sys->particle_mol_id, sys->particle_mol_id + N);
    p->particlesLD.assign(sys->particles_langevin, sys->particles_langevin + sys->num_langevin);
    p->imagePairs.assign(sys->image_pairs, sys->image_pairs + 2 * (size_t) sys->num_image_pairs);
    p->particlesElectrolyte.assign(sys->particles_electrolyte, sys->particles_electrolyte + sys->num_electrolyte);
    p->drudePairs.assign(sys->drude_pairs, sys->drude_pairs + 2 * (size_t) sys->num_drude);

    // membership marks replace the std::find predicates of VVIntegrator.h:326-343
    p->isLD.assign(N, 0);
    p->isImage.assign(N, 0);
    p->isNH.assign(N, 0);
    for (int32_t i : p->particlesLD) p->isLD[i] = 1;
    for (size_t k = 0; k < p->imagePairs.size(); k += 2) p->isImage[p->imagePairs[k]] = 1;

    // molecule masses accumulate in ascending particle order (VVIntegrator.cpp:130-135)
    p->moleculeMasses.assign(M, 0.0);
    for (int i = 0; i < N; i++)
        p->moleculeMasses[p->particleMolId[i]] += p->masses[i];
    p->moleculeInvMasses.resize(M);
    for (int m = 0; m < M; m++)
        p->moleculeInvMasses[m] = 1.0 / p->moleculeMasses[m];

    // Nose-Hoover particles ascending; their molecules in first-seen order (VVIntegrator.cpp:138-145)
    std::vector<unsigned char> molIsNH(M, 0);
    for (int i = 0; i < N; i++) {
        if (p->isLD[i] || p->isImage[i])
            continue;
        p->isNH[i] = 1;
        p->particlesNH.push_back(i);
        int32_t mol = p->particleMolId[i];
        if (!molIsNH[mol]) {
            molIsNH[mol] = 1;
            p->moleculesNH.push_back(mol);
        }
    }
    for (int32_t i : p->particlesLD)   // VVIntegrator.cpp:146-151
        if (molIsNH[p->particleMolId[i]])
            return conflict(p, "NH and Langevin thermostat cannot be applied on the same molecule");
    if (!p->particlesLD.empty() && par->cos_acceleration != 0)   // VVIntegrator.cpp:154-155
        return conflict(p, "Langevin thermostat and periodic perturbation shouldn't be used together");

    // particles grouped by molecule: a counting sort gives the reference's molecule-major,
    // particle-ascending order (CudaVVKernels.cpp:483-494)
    p->particlesInMolecules.assign(2 * (size_t) M, 0);
    for (int i = 0; i < N; i++)
        p->particlesInMolecules[2 * (size_t) p->particleMolId[i]]++;
    {
        int32_t start = 0;
        for (int m = 0; m < M; m++) {
            p->particlesInMolecules[2 * (size_t) m + 1] = start;
            start += p->particlesInMolecules[2 * (size_t) m];
        }
        std::vector<int32_t> cursor(M, 0);
        p->sortedByMol.resize(N);
        for (int i = 0; i < N; i++) {
            int32_t m = p->particleMolId[i];
            p->sortedByMol[p->particlesInMolecules[2 * (size_t) m + 1] + cursor[m]++] = i;
        }
    }

    // degrees of freedom: identical operation order to CudaVVKernels.cpp:497-564
    const bool useCOM = par->use_com_temp_group != 0;
    double dofAtom = 0.0, dofCOM = 0.0, dofDrude = 0.0;
    for (int i = 0; i < N; i++) {
        const double mass = p->masses[i];
        if (p->isNH[i] && mass != 0.0) {
            dofAtom += 3;
            if (useCOM)
                dofAtom -= 3 * mass * p->moleculeInvMasses[p->particleMolId[i]];
        }
    }
    std::vector<unsigned char> inPair(N, 0);
    for (size_t k = 0; k < p->drudePairs.size(); k += 2) {
        const int32_t d = p->drudePairs[k], parent = p->drudePairs[k + 1];
        if (p->isNH[d] != p->isNH[parent])
            return conflict(p, "Drude particle and its parent atom should be in the same thermostat");
        if (p->isNH[d]) {
            inPair[d] = inPair[parent] = 1;
            p->pairsNH.push_back(d);
            p->pairsNH.push_back(parent);
            dofAtom -= 3;
            dofDrude += 3;
        }
    }
    for (int i = 0; i < N; i++)
        if (p->isNH[i] && !inPair[i])
            p->normalNH.push_back(i);
    for (int k = 0; k < sys->num_constraints; k++) {
        const int32_t a = sys->constraints[2 * k], b = sys->constraints[2 * k + 1];
        if (p->isNH[a] != p->isNH[b])
            return conflict(p, "Constrained particle pair should be in the same thermostat");
        if (p->isNH[a])
            dofAtom -= 1;
    }
    if (useCOM)
        dofCOM = 3 * (double) p->moleculesNH.size();
    if (p->hasCMMotionRemover) {
        if (useCOM) dofCOM -= 3;
        else dofAtom -= 3;
    }
    p->dof[VVB200_TG_ATOM] = std::max(dofAtom, 0.0);
    p->dof[VVB200_TG_COM] = std::max(dofCOM, 0.0);
    p->dof[VVB200_TG_DRUDE] = std::max(dofDrude, 0.0);
    p->numTempGroup = 3;
    if (p->dof[VVB200_TG_DRUDE] == 0) {
        p->numTempGroup = 2;
        if (p->dof[VVB200_TG_COM] == 0)
            p->numTempGroup = 1;
    }
    // chain masses and targets (CudaVVKernels.cpp:583-594)
    const int nc = par->num_nh_chains;
    const double realKbT = BOLTZ * par->temperature, drudeKbT = BOLTZ * par->drude_temperature;
    p->etaMass.assign((size_t) p->numTempGroup * nc, 0.0);
    p->NkbT.assign(p->numTempGroup, 0.0);
    for (int g = 0; g < p->numTempGroup; g++) {
        const double kT = g == VVB200_TG_DRUDE ? drudeKbT : realKbT;
        const double q = g == VVB200_TG_DRUDE ? drudeKbT / std::pow(par->drude_frequency, 2)
                                              : realKbT / std::pow(par->frequency, 2);
        p->NkbT[g] = p->dof[g] * kT;
        p->etaMass[(size_t) g * nc] = p->dof[g] * q;
        for (int k = 1; k < nc; k++)
            p->etaMass[(size_t) g * nc + k] = q;
    }
    for (int g = 0; g < 3; g++)
        p->dofGlobal[g] = p->dof[g];

    // Langevin sets (CudaVVKernels.cpp:775-804); built only when the LD kernel would exist
    if (!p->particlesLD.empty()) {
        std::fill(inPair.begin(), inPair.end(), 0);
        for (size_t k = 0; k < p->drudePairs.size(); k += 2) {
            const int32_t d = p->drudePairs[k], parent = p->drudePairs[k + 1];
            if (p->isLD[d] != p->isLD[parent])
                return conflict(p, "Drude particle and its parent atom should be in the same thermostat");
            if (p->isLD[d]) {
                inPair[d] = inPair[parent] = 1;
                p->pairsLD.push_back(d);
                p->pairsLD.push_back(parent);
            }
        }
        for (int k = 0; k < sys->num_constraints; k++)
            if (p->isLD[sys->constraints[2 * k]] != p->isLD[sys->constraints[2 * k + 1]])
                return conflict(p, "Constrained particle pair should be in the same thermostat");
        for (int i = 0; i < N; i++)
            if (p->isLD[i] && !inPair[i])
                p->normalLD.push_back(i);
    }

    // total mass in particle order (CudaVVKernels.cpp:1028-1031)
    double massTotal = 0;
    for (int i = 0; i < N; i++)
        massTotal += p->masses[i];
    p->invMassTotal = 1.0 / massTotal;
    p->totalMassGlobal = massTotal;

    if (!vvb200_build_tiles(p, p->tileSM)) {
        vvb200_set_error("vvb200_plan_create: %s", p->tiledWhyNot.c_str());
        delete p;
        return VVB200_ERR_UNSUPPORTED_TOPOLOGY;
    }
    *out = p;
    return VVB200_OK;
}

// (Re)builds the tile tables for a GPU with `numSM` multiprocessors: plan_create assumes a B200 (148); plan_upload
// calls this again when the device it uploads to reports another count (tile sizes of small systems follow the number of
// co-resident blocks).  False: no tiling of either kind exists (tiledWhyNot says why).
bool vvb200_build_tiles(vvb200_plan *p, int numSM) {
    p->tileSM = numSM > 0 ? numSM : 148;
    p->tiled = buildTiles(p);
    if (p->tiled && getenv("VVB200_FORCE_GENERAL")) {     // testing hook: exercise the any-topology path
        p->tiled = false;
        p->tiledWhyNot = "forced by VVB200_FORCE_GENERAL";
    }
    return p->tiled || buildPlainTiles(p);
}

// Tables of the fused two-pass kernels.  A tile is a contiguous particle range of at most
// VVB200_TILE_CAP slots that never separates (a) the massive-or-thermostatted particles of one
// thermostat molecule when the COM temperature group is on, or (b) the two particles of a Drude
// pair.  Returns false (with tiledWhyNot set) if the topology cannot be tiled that way.
// Nearly equal tiles whose count is a multiple of 888 (see buildTiles).  cover[i] != 0: no cut between slots i-1 and i;
// molBefore: prefix count of thermostat molecules.  Returns false when no such partition respects the tile limits.
static bool buildBalancedTiles(vvb200_plan *p, const std::vector<int32_t> &cover, const std::vector<int32_t> &molBefore) {
    const int N = p->N;
    const int unit = 888;
    for (int T = (N / VVB200_TILE_CAP / unit + 1) * unit; (double) N / T >= 192.0; T += unit) {
        p->tileStart.assign(1, 0);
        bool ok = true;
        for (int i = 1; i <= T && ok; i++) {
            int32_t e = i == T ? N : (int32_t) ((long long) N * i / T);
            while (e > p->tileStart.back() && e < N && cover[e] != 0)
                e--;
            const int32_t s = p->tileStart.back();
            if (e <= s || e - s > VVB200_TILE_CAP || molBefore[e] - molBefore[s] > VVB200_TILE_MAX_MOLS)
                ok = false;
            else
                p->tileStart.push_back(e);
        }
        if (ok)
            return true;
    }
    p->tileStart.clear();
    return false;
}

static bool buildTiles(vvb200_plan *p) {
    const int N = p->N, M = p->M;
    const bool useCOM = p->par.use_com_temp_group != 0;
    std::vector<unsigned char> molIsNH(M, 0);
    for (int32_t m : p->moleculesNH) molIsNH[m] = 1;

    // a particle takes part in a molecular COM if its molecule is thermostatted and it either
    // carries mass (contributes) or is a NH particle (gets normalised)
    auto comMember = [&](int i) {
        return useCOM && molIsNH[p->particleMolId[i]] && (p->masses[i] != 0.0 || p->isNH[i]);
    };

    // the fused kernels take sum m|v - V|^2 as sum m|v|^2 - M|V|^2, which needs every massive member of a thermostat
    // molecule to be a thermostat particle itself (a massive image particle would not be)
    if (useCOM)
        for (int i = 0; i < N; i++)
            if (molIsNH[p->particleMolId[i]] && p->masses[i] != 0.0 && !p->isNH[i]) {
                p->tiledWhyNot = "a massive particle outside the thermostat belongs to a thermostat molecule";
                return false;
            }
    std::vector<int32_t> lo(M, N), hi(M, -1);
    for (int i = 0; i < N; i++)
        if (comMember(i)) {
            int32_t m = p->particleMolId[i];
            lo[m] = std::min(lo[m], i);
            hi[m] = std::max(hi[m], i);
        }
    // cover[i] > 0  <=>  a cut between slots i-1 and i would split some unit
    std::vector<int32_t> cover(N + 2, 0);
    auto protect = [&](int32_t a, int32_t b) {
        if (b > a) {
            cover[a + 1]++;
            cover[b + 1]--;
        }
    };
    std::vector<int32_t> partner(N, -1);
    std::vector<unsigned char> role(N, VVB200_ROLE_NONE);
    for (size_t k = 0; k < p->drudePairs.size(); k += 2) {
        const int32_t d = p->drudePairs[k], parent = p->drudePairs[k + 1];
        if (d == parent || partner[d] != -1 || partner[parent] != -1) {
            p->tiledWhyNot = "a particle belongs to more than one Drude pair";
            return false;
        }
        partner[d] = parent;
        partner[parent] = d;
        role[d] = VVB200_ROLE_DRUDE;
        role[parent] = VVB200_ROLE_PARENT;
        protect(std::min(d, parent), std::max(d, parent));
        if (std::abs(d - parent) >= VVB200_META_PARTNER_BIAS) {
            p->tiledWhyNot = "Drude pair partners are further apart than a tile";
            return false;
        }
    }
    // Drude pairs are never cut; thermostat molecules are not cut either unless they are longer than a tile
    // (splitLongerThan): those are summed fragment by fragment and finished by the last block of pass A
    const std::vector<int32_t> pairCover = cover;
    std::vector<unsigned char> isSplit(M, 0);
    bool splitting = false;
    auto buildCover = [&](int splitLongerThan) {
        cover = pairCover;
        std::fill(isSplit.begin(), isSplit.end(), 0);
        for (int m = 0; m < M; m++)
            if (hi[m] >= 0) {
                if (hi[m] - lo[m] + 1 > splitLongerThan) isSplit[m] = 1;
                else protect(lo[m], hi[m]);
            }
        for (int i = 1; i <= N; i++)
            cover[i] += cover[i - 1];
    };
    buildCover(N + 1);
    // molBefore[i] = thermostat molecules whose first COM member lies below slot i (a tile owns the
    // molecules that start inside it, plus at most one that continues from the tile before)
    std::vector<int32_t> molBefore(N + 1, 0);
    for (int m = 0; m < M; m++)
        if (hi[m] >= 0)
            molBefore[lo[m] + 1]++;
    for (int i = 1; i <= N; i++)
        molBefore[i] += molBefore[i - 1];

    // Tile size.  Large systems: VVB200_TILE_CAP (what the streaming kernels' stages hold).  Small systems are bound
    // by the latency of each thread's dependent fp64 chain, not by bytes, so they are cut into more, smaller tiles --
    // about one per co-resident block of the GPU (2 per SM; 148 SMs on a B200) -- which spreads the same work over more SMs.
    int cap = VVB200_TILE_CAP;
    {
        const char *env = getenv("VVB200_SMALL_TILES");
        if (!env || atoi(env) != 0) {
            const int perBlock = (N + 2 * p->tileSM - 1) / (2 * p->tileSM);
            cap = std::min(VVB200_TILE_CAP, std::max(128, (perBlock + 31) / 32 * 32));
        }
        if (const char *t = getenv("VVB200_TILE_PARTICLES"))       // tuning: a fixed tile size
            if (atoi(t) >= 64) cap = std::min(VVB200_TILE_CAP, atoi(t) / 32 * 32);
    }
    // Experiment kept behind VVB200_BALANCED_TILES=1 (off): the persistent grids of pass A (3 x 148 blocks) and pass B
    // (2 x 148) assign tiles round-robin, so mid-size systems leave part of the machine idle for a tile at the end
    // (1M particles: 4.5 tiles per pass-A block).  Cutting into a multiple of lcm(444, 296) = 888 equal tiles removes
    // that -- and measured SLOWER on B200 (1M: 68.4 vs 65.0 us per step, 4M: 184.5 vs 179.9): the smaller tiles' fixed
    // per-tile cost (barriers, pipeline hand-offs) outweighs the imbalance.
    bool balanced = false;
    if (cap == VVB200_TILE_CAP && N <= 12000000) {
        const char *env = getenv("VVB200_BALANCED_TILES");
        balanced = env && atoi(env) != 0 && buildBalancedTiles(p, cover, molBefore);
    }
    while (!balanced) {
        p->tileStart.clear();
        p->tileStart.push_back(0);
        int32_t s = 0;
        bool ok = true;
        while (s < N) {
            int32_t e = std::min(N, s + cap);
            while (e > s + 1 && molBefore[e] - molBefore[s] > VVB200_TILE_MAX_MOLS - (splitting ? 1 : 0))
                e--;
            while (e > s && e < N && cover[e] != 0)
                e--;
            if (e == s) {
                ok = false;
                break;
            }
            p->tileStart.push_back(e);
            s = e;
        }
        // small tiles only pay while every tile still gets its own co-resident block (units that may not be split
        // leave tiles partly empty, so the count can exceed N / cap): otherwise grow the tile and cut again
        if (ok && cap < VVB200_TILE_CAP && (int) p->tileStart.size() - 1 > 2 * p->tileSM && !getenv("VVB200_TILE_PARTICLES")) {
            cap += 32;
            continue;
        }
        if (ok)
            break;
        if (cap < VVB200_TILE_CAP) {     // a molecule longer than the small tile: fall back to full-size tiles
            cap = VVB200_TILE_CAP;
            continue;
        }
        if (!splitting && useCOM && !(getenv("VVB200_SPLIT_MOLECULES") && atoi(getenv("VVB200_SPLIT_MOLECULES")) == 0)) {
            splitting = true;            // polymers, proteins: cut the molecules that cannot fit a tile
            buildCover(VVB200_TILE_CAP);
            continue;
        }
        p->tiledWhyNot = "a thermostat molecule or Drude pair spans more than one tile";
        p->tileStart.clear();
        return false;
    }
    const int numTiles = (int) p->tileStart.size() - 1;

    // electrolyte multiplicities
    std::vector<int32_t> elecCount(N, 0);
    for (int32_t i : p->particlesElectrolyte)
        if (++elecCount[i] > (int) VVB200_META_ELEC_MASK) {
            p->tiledWhyNot = "a particle appears more than 7 times in the electrolyte list";
            p->tileStart.clear();
            return false;
        }

    // tile-local molecule numbering in order of first appearance
    p->slotMeta.assign(N, 0);
    p->tileMolOffset.assign(numTiles + 1, 0);
    p->tileMolList.clear();
    std::vector<int32_t> localOf(M, -1);
    for (int t = 0; t < numTiles; t++) {
        const int32_t a = p->tileStart[t], b = p->tileStart[t + 1];
        const size_t base = p->tileMolList.size();
        for (int32_t i = a; i < b; i++) {
            uint32_t word = VVB200_META_MOL_NONE;
            if (comMember(i)) {
                const int32_t m = p->particleMolId[i];
                if (localOf[m] < 0) {
                    localOf[m] = (int32_t) (p->tileMolList.size() - base);
                    p->tileMolList.push_back(m);
                }
                word = (uint32_t) localOf[m];
            }
            word |= (uint32_t) elecCount[i] << VVB200_META_ELEC_SHIFT;
            if (p->isNH[i]) word |= VVB200_META_NH;
            if (p->isLD[i]) word |= VVB200_META_LD;
            word |= (uint32_t) role[i] << VVB200_META_ROLE_SHIFT;
            const int32_t off = role[i] != VVB200_ROLE_NONE ? partner[i] - i : 0;
            word |= (uint32_t) (off + VVB200_META_PARTNER_BIAS) << VVB200_META_PARTNER_SHIFT;
            p->slotMeta[i] = word;
        }
        for (size_t k = base; k < p->tileMolList.size(); k++)
            localOf[p->tileMolList[k]] = -1;
        p->tileMolOffset[t + 1] = (int32_t) p->tileMolList.size();
    }

    // Image update fused into pass B: possible when every parent has exactly one image, no image is itself a parent and
    // images are massless particles outside both thermostats (their own thread then writes nothing)
    p->imageOf.clear();
    if (!p->imagePairs.empty()) {
        std::vector<int32_t> imageOf(N, -1);
        bool ok = true;
        for (size_t k = 0; k < p->imagePairs.size() && ok; k += 2) {
            const int32_t img = p->imagePairs[k], parent = p->imagePairs[k + 1];
            ok = imageOf[parent] < 0 && !p->isImage[parent] && p->masses[img] == 0.0 && !p->isNH[img] && !p->isLD[img] &&
                 role[img] == VVB200_ROLE_NONE;
            imageOf[parent] = img;
        }
        if (ok) {
            p->imageOf.swap(imageOf);
            for (int i = 0; i < N; i++)
                if (p->imageOf[i] >= 0) p->slotMeta[i] |= VVB200_META_HAS_IMAGE;
        }
    }

    // fragments of the cut molecules, numbered in tile order
    p->tileMolFrag.assign(p->tileMolList.size(), -1);
    p->splitMolId.clear();
    p->splitFragOffset.assign(1, 0);
    p->splitFragList.clear();
    if (splitting) {
        std::vector<std::vector<int32_t>> frags(M);
        int32_t numFrag = 0;
        for (size_t k = 0; k < p->tileMolList.size(); k++)
            if (isSplit[p->tileMolList[k]]) {
                p->tileMolFrag[k] = numFrag;
                frags[p->tileMolList[k]].push_back(numFrag++);
            }
        for (int m = 0; m < M; m++)
            if (isSplit[m] && !frags[m].empty()) {
                p->splitMolId.push_back(m);
                p->splitFragList.insert(p->splitFragList.end(), frags[m].begin(), frags[m].end());
                p->splitFragOffset.push_back((int32_t) p->splitFragList.size());
            }
    }

    // compact Langevin-force slots: normal i -> i, pair k -> nNormal + 2k (Drude), +1 (parent):
    // the same positions as the reference's random-number indices (drudeLangevin.cu:20,45-46)
    if (!p->particlesLD.empty()) {
        p->ldSlot.assign(N, -1);
        for (size_t i = 0; i < p->normalLD.size(); i++)
            p->ldSlot[p->normalLD[i]] = (int32_t) i;
        const int32_t base = (int32_t) p->normalLD.size();
        for (size_t k = 0; k < p->pairsLD.size(); k += 2) {
            p->ldSlot[p->pairsLD[k]] = base + (int32_t) k;
            p->ldSlot[p->pairsLD[k + 1]] = base + (int32_t) k + 1;
        }
    }
    return true;
}

// Tables of the any-topology path: fixed 512-slot tiles that only the element-wise kick uses (pass A with its
// molecule and pair phases switched off); the thermostat runs through gather kernels over the reference's own index
// arrays (normalNH, pairsNH, sortedByMol, ...), so nothing has to be tile-local.
static bool buildPlainTiles(vvb200_plan *p) {
    const int N = p->N;
    std::vector<int32_t> elecCount(N, 0);
    for (int32_t i : p->particlesElectrolyte)
        if (++elecCount[i] > (int) VVB200_META_ELEC_MASK) {
            p->tiledWhyNot = "a particle appears more than 7 times in the electrolyte list";
            return false;
        }
    p->tileStart.clear();
    for (int32_t s = 0; s < N; s += VVB200_TILE_CAP)
        p->tileStart.push_back(s);
    p->tileStart.push_back(N);
    const int numTiles = (int) p->tileStart.size() - 1;
    p->tileMolOffset.assign(numTiles + 1, 0);
    p->tileMolList.clear();
    p->slotMeta.assign(N, 0);
    for (int i = 0; i < N; i++) {
        uint32_t word = VVB200_META_MOL_NONE | ((uint32_t) elecCount[i] << VVB200_META_ELEC_SHIFT);
        if (p->isNH[i]) word |= VVB200_META_NH;
        if (p->isLD[i]) word |= VVB200_META_LD;
        word |= (uint32_t) VVB200_META_PARTNER_BIAS << VVB200_META_PARTNER_SHIFT;
        p->slotMeta[i] = word;
    }
    if (!p->particlesLD.empty()) {
        p->ldSlot.assign(N, -1);
        for (size_t i = 0; i < p->normalLD.size(); i++)
            p->ldSlot[p->normalLD[i]] = (int32_t) i;
        const int32_t base = (int32_t) p->normalLD.size();
        for (size_t k = 0; k < p->pairsLD.size(); k += 2) {
            p->ldSlot[p->pairsLD[k]] = base + (int32_t) k;
            p->ldSlot[p->pairsLD[k + 1]] = base + (int32_t) k + 1;
        }
    }
    return true;
}

extern "C" void vvb200_plan_destroy(vvb200_plan *plan) {
    if (!plan)
        return;
    vvb200_device_free(plan);
    delete plan;
}

extern "C" int vvb200_plan_get_int_array(const vvb200_plan *p, int which, const int32_t **ptr, int64_t *len) {
    if (!p || !ptr || !len) {
        vvb200_set_error("vvb200_plan_get_int_array: null argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    const std::vector<int32_t> *v = nullptr;
    switch (which) {
    case VVB200_ARR_PARTICLES_NH: v = &p->particlesNH; break;
    case VVB200_ARR_MOLECULES_NH: v = &p->moleculesNH; break;
    case VVB200_ARR_PARTICLE_MOL_ID: v = &p->particleMolId; break;
    case VVB200_ARR_DRUDE_PAIRS: v = &p->drudePairs; break;
    case VVB200_ARR_SORTED_BY_MOL: v = &p->sortedByMol; break;
    case VVB200_ARR_PARTICLES_IN_MOLECULES: v = &p->particlesInMolecules; break;
    case VVB200_ARR_NORMAL_NH: v = &p->normalNH; break;
    case VVB200_ARR_PAIRS_NH: v = &p->pairsNH; break;
    case VVB200_ARR_NORMAL_LD: v = &p->normalLD; break;
    case VVB200_ARR_PAIRS_LD: v = &p->pairsLD; break;
    case VVB200_ARR_IMAGE_PAIRS: v = &p->imagePairs; break;
    case VVB200_ARR_ELECTROLYTE: v = &p->particlesElectrolyte; break;
    case VVB200_ARR_TILE_START: v = &p->tileStart; break;
    case VVB200_ARR_TILE_MOL_OFFSET: v = &p->tileMolOffset; break;
    case VVB200_ARR_TILE_MOL_LIST: v = &p->tileMolList; break;
    case VVB200_ARR_TILE_MOL_FRAG: v = &p->tileMolFrag; break;
    case VVB200_ARR_SPLIT_MOL_ID: v = &p->splitMolId; break;
    case VVB200_ARR_SPLIT_FRAG_OFFSET: v = &p->splitFragOffset; break;
    case VVB200_ARR_SPLIT_FRAG_LIST: v = &p->splitFragList; break;
    case VVB200_ARR_IMAGE_OF: v = &p->imageOf; break;
    case VVB200_ARR_SLOT_META:
        *ptr = reinterpret_cast<const int32_t *>(p->slotMeta.data());
        *len = (int64_t) p->slotMeta.size();
        return VVB200_OK;
    default:
        vvb200_set_error("vvb200_plan_get_int_array: unknown array id %d", which);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    *ptr = v->data();
    *len = (int64_t) v->size();
    return VVB200_OK;
}

extern "C" int vvb200_plan_get_f64_array(const vvb200_plan *p, int which, const double **ptr, int64_t *len) {
    if (!p || !ptr || !len) {
        vvb200_set_error("vvb200_plan_get_f64_array: null argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    switch (which) {
    case VVB200_F64_MOLECULE_MASSES: *ptr = p->moleculeMasses.data(); *len = p->M; break;
    case VVB200_F64_MOLECULE_INV_MASSES: *ptr = p->moleculeInvMasses.data(); *len = p->M; break;
    case VVB200_F64_DOF: *ptr = p->dof; *len = 3; break;
    case VVB200_F64_ETA_MASS: *ptr = p->etaMass.data(); *len = (int64_t) p->etaMass.size(); break;
    case VVB200_F64_NKBT: *ptr = p->NkbT.data(); *len = (int64_t) p->NkbT.size(); break;
    case VVB200_F64_INV_MASS_TOTAL: *ptr = &p->invMassTotal; *len = 1; break;
    case VVB200_F64_DOF_GLOBAL: *ptr = p->dofGlobal; *len = 3; break;
    default:
        vvb200_set_error("vvb200_plan_get_f64_array: unknown array id %d", which);
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    return VVB200_OK;
}

extern "C" int vvb200_plan_num_temp_groups(const vvb200_plan *p) { return p ? p->numTempGroup : 0; }
extern "C" int vvb200_plan_uses_tiled_path(const vvb200_plan *p) { return p && p->tiled ? 1 : 0; }

extern "C" uint32_t vvb200_plan_random_request(const vvb200_plan *p) {
    if (!p || p->particlesLD.empty())
        return 0;
    // the reference sizes both device arrays max(size,1) and requests from those sizes
    const size_t nNormal = std::max<size_t>(p->normalLD.size(), 1);
    const size_t nPairs = std::max<size_t>(p->pairsLD.size() / 2, 1);
    return (uint32_t) (nNormal + 2 * nPairs);
}

extern "C" int vvb200_set_step_size(vvb200_plan *p, double stepSize) {
    if (!p || !(stepSize > 0)) {
        vvb200_set_error("vvb200_set_step_size: invalid argument");
        return VVB200_ERR_INVALID_ARGUMENT;
    }
    p->par.step_size = stepSize;
    return VVB200_OK;
}

extern "C" int64_t vvb200_launch_count(const vvb200_plan *p) { return p ? p->launches : 0; }
