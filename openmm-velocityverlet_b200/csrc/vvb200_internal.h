// vvb200_internal.h -- plan object shared by the host builders (vvb200_plan.cpp) and the CUDA
// side (vvb200_device.cu).  Not part of the public ABI (include/vvb200.h is).
#ifndef VVB200_INTERNAL_H_
#define VVB200_INTERNAL_H_

#include <cstdint>
#include <string>
#include <vector>

#include "vvb200.h"

// ---- packed per-slot topology word read by the fused kernels -------------------------------
//  bits  0-10  tile-local index of the particle's thermostat molecule (0x7FF = takes no part in
//              a molecular centre of mass)
//  bits 11-13  how many times the particle appears in the electrolyte list (duplicates add twice,
//              CudaVVKernels.cpp:954-957 / electricField.cu:7-11)
//  bit  14     has an image particle that pass B mirrors along (fused image update)
//  bit  16     member of the Nose-Hoover set (VVIntegrator::isParticleNH)
//  bits 17-18  Drude pair role: 0 none, 1 Drude (pair.x), 2 parent (pair.y)
//  bit  19     Langevin particle (has an entry in ldSlot[])
//  bits 20-31  partner slot minus own slot, biased by 2048 (pair roles only)
#define VVB200_META_MOL_MASK 0x7FFu
#define VVB200_META_MOL_NONE 0x7FFu
#define VVB200_META_ELEC_SHIFT 11
#define VVB200_META_ELEC_MASK 0x7u
#define VVB200_META_HAS_IMAGE (1u << 14)    // parent of an image particle (imageOf[] holds the image's index)
#define VVB200_META_NH (1u << 16)
#define VVB200_META_ROLE_SHIFT 17
#define VVB200_META_ROLE_MASK 0x3u
#define VVB200_META_LD (1u << 19)
#define VVB200_META_PARTNER_SHIFT 20
#define VVB200_META_PARTNER_BIAS 2048
#define VVB200_ROLE_NONE 0u
#define VVB200_ROLE_DRUDE 1u
#define VVB200_ROLE_PARENT 2u

// particles per tile of the fused kernels (threads x items); must be <= 1024 (11-bit local ids)
#ifndef VVB200_TILE_CAP
#define VVB200_TILE_CAP 512
#endif
// thermostat molecules per tile (bounds the per-stage molecule tables of the fused kernels)
#define VVB200_TILE_MAX_MOLS 128

// reduction vector layout (fp64), per block partial and final:
//  [0..2]  A_g : sum m u^2 per temperature group (u = velocity relative to molecule COM, bias-free)
//  [3]     S   : sum 2 m vx c          (cosine runs)   -> V = S / M_total
//  [4..6]  B_g : first moment of the bias in group g   (cosine runs)
//  [7..9]  C_g : second moment of the bias in group g  (cosine runs)
#define VVB200_NRED 10

struct vvb200_device_state;   // defined in vvb200_device.cu

struct vvb200_plan {
    int precision = VVB200_MIXED;
    vvb200_params par{};
    int N = 0, paddedN = 0, M = 0;
    bool hasCMMotionRemover = false;

    std::vector<double> masses;
    // VVIntegrator members
    std::vector<int32_t> particleMolId, particlesNH, moleculesNH, particlesLD, particlesElectrolyte;
    std::vector<int32_t> imagePairs;            // flattened (image,parent)
    std::vector<double> moleculeMasses, moleculeInvMasses;
    // kernel-object arrays
    std::vector<int32_t> drudePairs;            // flattened (p,p1)
    std::vector<int32_t> sortedByMol, particlesInMolecules /* (count,start) */;
    std::vector<int32_t> normalNH, pairsNH /* flattened */, normalLD, pairsLD /* flattened */;
    double dof[3] = {0, 0, 0};
    int numTempGroup = 1;
    std::vector<double> etaMass, NkbT;          // [numTG*nc], [numTG]
    double invMassTotal = 0;
    // global-thermostat overrides for molecule-partitioned multi-GPU runs
    double dofGlobal[3] = {0, 0, 0};
    double totalMassGlobal = 0;
    bool partitioned = false;                   // vvb200_set_global_thermostat was called: one rank of a multi-GPU run

    // fused-path tables
    int tileSM = 148;                           // multiprocessor count the tile sizes were chosen for
    bool tiled = false;
    std::string tiledWhyNot;
    std::vector<int32_t> tileStart;             // [numTiles+1]
    std::vector<int32_t> tileMolOffset;         // [numTiles+1] prefix into tileMolList
    std::vector<int32_t> tileMolList;           // global molecule id of each tile-local molecule
    std::vector<uint32_t> slotMeta;             // [N]
    std::vector<int32_t> ldSlot;                // [N] compact Langevin-force slot or -1 (only if LD)
    // thermostat molecules longer than a tile are cut: each tile sums its fragment, the last block of pass A adds the
    // fragments up (tile order) and finishes the molecule's centre of mass
    std::vector<int32_t> imageOf;               // [N] index of the particle's image or -1; empty unless the image update is fused
    std::vector<int32_t> tileMolFrag;           // per tile-local molecule: fragment index or -1
    std::vector<int32_t> splitMolId;            // molecules that are cut, ascending
    std::vector<int32_t> splitFragOffset;       // [numSplit+1] prefix into splitFragList
    std::vector<int32_t> splitFragList;         // fragment indices of each cut molecule, in tile order
    std::vector<unsigned char> isNH, isLD, isImage;

    int64_t launches = 0;
    vvb200_device_state *dev = nullptr;
};

void vvb200_set_error(const char *fmt, ...);
bool vvb200_build_tiles(vvb200_plan *plan, int numSM);
void vvb200_device_free(vvb200_plan *plan);

#endif
