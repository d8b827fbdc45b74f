/*
 * mini_openmm.cpp -- TEST INFRASTRUCTURE ONLY: the out-of-line half of mini_openmm.h (plain host C++; everything that
 * touches kernels or device memory goes through the miniomm:: functions of plugin_kernels.cpp).
 */
#include "mini_openmm.h"

#include <algorithm>
#include <cmath>

using namespace OpenMM;

namespace {
struct MiniModule {
    std::string tag;      /* which kernel file the module was built from: "middle", "drudeNoseHoover", ... */
    int numTG;
};
/* the #defines live in process-wide variables (constant memory on the GPU): whoever launches rebinds them when another
 * context launched last */
CudaContext *g_boundContext = NULL;
const size_t realSize = MINIOMM_MODE == 2 ? 8 : 4;
const size_t mixedSize = MINIOMM_MODE == 0 ? 4 : 8;
}   // namespace

/* ---- CudaArray ------------------------------------------------------------------------------------------------- */
CudaArray::CudaArray(CudaContext &cu, int size, int elementSize, const std::string &name)
    : cu(cu), size(size), elementSize(elementSize), name(name), pointer(0) {
    if (size < 0)
        throw OpenMMException("Error creating array " + name + ": negative size");
    /* cuMemAlloc does not clear memory; this stand-in does (the reference's kinetic-energy buffer
     * tail is only ever written by the threads that exist, and summed in full) */
    pointer = (CUdeviceptr) miniomm::deviceAlloc((size_t) std::max(size, 1) * elementSize);
}

void CudaArray::upload(const void *data, bool) {
    const size_t bytes = (size_t) size * elementSize;
    miniomm::copyToDevice((void *) pointer, data, bytes, cu.stream);
    std::vector<unsigned char> &rec = cu.uploads[name];
    rec.assign((const unsigned char *) data, (const unsigned char *) data + bytes);
    cu.uploadElementSize[name] = elementSize;
}

void CudaArray::download(void *data, bool) const {
    miniomm::copyToHost(data, (const void *) pointer, (size_t) size * elementSize, cu.stream);
}

/* ---- CudaIntegrationUtilities ------------------------------------------------------------------------------------ */
CudaIntegrationUtilities::CudaIntegrationUtilities(CudaContext &cu)
    : constraintCalls(0), velocityConstraintCalls(0), cu(cu), posDelta(NULL), stepSize(NULL), random(NULL), randomCount(0),
      randomPos(0), randomSeed(0), lastStepSize(-1.0) {
    posDelta = new CudaArray(cu, cu.getPaddedNumAtoms(), (int) (4 * mixedSize), "posDelta");
    stepSize = new CudaArray(cu, 1, (int) (2 * mixedSize), "stepSize");
    random = new CudaArray(cu, 1, 16, "random");
}

void CudaIntegrationUtilities::setRandomStream(const float *values4, size_t count) {
    delete random;
    random = new CudaArray(cu, (int) std::max<size_t>(count, 1), 16, "random");
    if (count)
        miniomm::copyToDevice((void *) random->getDevicePointer(), values4, count * 16, cu.stream);
    randomCount = count;
    randomPos = 0;
}

void CudaIntegrationUtilities::setNextStepSize(double size) {
    if (size == lastStepSize)
        return;
    lastStepSize = size;
    if (MINIOMM_MODE == 0) {
        float2 ss = make_float2(0, (float) size);
        stepSize->upload(&ss);
    } else {
        double2 ss = make_double2(0, size);
        stepSize->upload(&ss);
    }
}

void CudaIntegrationUtilities::applyConstraints(double) {
    constraintCalls++;
    miniomm::standinPositions(standin, (void *) cu.getPosq().getDevicePointer(),
                              MINIOMM_MODE == 1 ? (void *) cu.getPosqCorrection().getDevicePointer() : NULL,
                              (void *) cu.getVelm().getDevicePointer(), (void *) posDelta->getDevicePointer(), cu.stream);
}

void CudaIntegrationUtilities::applyVelocityConstraints(double) {
    velocityConstraintCalls++;
    miniomm::standinVelocities(standin, (void *) cu.getPosq().getDevicePointer(),
                               MINIOMM_MODE == 1 ? (void *) cu.getPosqCorrection().getDevicePointer() : NULL,
                               (void *) cu.getVelm().getDevicePointer(), cu.stream);
}

double CudaIntegrationUtilities::computeKineticEnergy(double) {
    /* 1/2 sum m v^2 from a download (the real one runs a reduction kernel; not on the path under test) */
    const int n = cu.getNumAtoms();
    std::vector<unsigned char> raw((size_t) cu.getPaddedNumAtoms() * 4 * mixedSize);
    cu.getVelm().download(raw.data());
    double ke = 0;
    for (int i = 0; i < n; i++) {
        double v[4];
        for (int k = 0; k < 4; k++)
            v[k] = mixedSize == 8 ? ((const double *) raw.data())[4 * i + k] : (double) ((const float *) raw.data())[4 * i + k];
        if (v[3] != 0)
            ke += (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / v[3];
    }
    return 0.5 * ke;
}

/* ---- CudaContext -------------------------------------------------------------------------------------------------- */
CudaContext::CudaContext(const System &system, CudaPlatform::PlatformData &data, void *stream)
    : stream(stream), kernelLaunches(0), reorderCalls(0), forceInfoAddedBeforeInit(false), platformData(data), posq(NULL),
      posqCorrection(NULL), velm(NULL), force(NULL), integration(NULL), time(0), stepCount(0) {
    numAtoms = system.getNumParticles();
    paddedNumAtoms = (numAtoms + 31) / 32 * 32;             /* TileSize = 32 [OMM-mem] */
    numThreadBlocks = miniomm::numThreadBlocks();
    posq = new CudaArray(*this, paddedNumAtoms, (int) (4 * realSize), "posq");
    posqCorrection = MINIOMM_MODE == 1 ? new CudaArray(*this, paddedNumAtoms, 16, "posqCorrection") : new CudaArray(*this);
    velm = new CudaArray(*this, paddedNumAtoms, (int) (4 * mixedSize), "velm");
    force = new CudaArray(*this, 3 * paddedNumAtoms, 8, "force");
    integration = new CudaIntegrationUtilities(*this);
    setPeriodicBoxSize(1, 1, 1);
}

CudaContext::~CudaContext() {
    if (g_boundContext == this)
        g_boundContext = NULL;
    delete integration;
    delete posq; delete posqCorrection; delete velm; delete force;
    for (size_t i = 0; i < forceInfos.size(); i++) delete forceInfos[i];
    for (size_t i = 0; i < modules.size(); i++) delete (MiniModule *) modules[i];
}

void CudaContext::setPeriodicBoxSize(double x, double y, double z) {
    box = make_double4(x, y, z, 0);
    invBoxDouble = make_double4(1.0 / x, 1.0 / y, 1.0 / z, 0);
    invBoxFloat = make_float4((float) (1.0 / x), (float) (1.0 / y), (float) (1.0 / z), 0);
}

CUmodule CudaContext::createModule(const std::string &source, const std::map<std::string, std::string> &defs, const char *) {
    /* `source` is CudaVVKernelSources::vectorOps + CudaVVKernelSources::<file>; the stand-in CudaVVKernelSources.h (generated
     * by oracle/Makefile) holds "[<file>]" tags instead of the text, since the kernels are compiled ahead of time */
    MiniModule *m = new MiniModule();
    m->numTG = 0;
    const size_t close = source.rfind(']'), open = source.rfind('[');
    if (open == std::string::npos || close == std::string::npos || close < open)
        throw OpenMMException("mini-OpenMM: createModule got a source without a [file] tag");
    m->tag = source.substr(open + 1, close - open - 1);
    for (std::map<std::string, std::string>::const_iterator it = defs.begin(); it != defs.end(); ++it) {
        const long value = atol(it->second.c_str());
        if (it->first == "NUM_TG")
            m->numTG = (int) value;
        else
            defines[it->first] = value;
    }
    if (g_boundContext == this)
        g_boundContext = NULL;         /* rebind on the next launch */
    modules.push_back((CUmodule) m);
    return (CUmodule) m;
}

CUfunction CudaContext::getKernel(CUmodule module, const std::string &name) {
    const MiniModule *m = (const MiniModule *) module;
    miniomm::KernelEntry *k = miniomm::findKernel(m->tag, name, m->numTG);
    if (!k)
        throw OpenMMException("Error creating kernel " + name + ": no such kernel in module " + m->tag);
    return (CUfunction) k;
}

void CudaContext::executeKernel(CUfunction kernel, void **arguments, int workUnits, int blockSize, unsigned int sharedSize) {
    if (blockSize == -1)
        blockSize = ThreadBlockSize;
    const int gridSize = std::min((workUnits + blockSize - 1) / blockSize, numThreadBlocks);
    if (gridSize <= 0)
        return;      /* cuLaunchKernel rejects an empty grid; the reference never asks for one with data behind it */
    if (g_boundContext != this) {
        for (std::map<std::string, long>::const_iterator it = defines.begin(); it != defines.end(); ++it)
            miniomm::setDefine(it->first, it->second, stream);
        g_boundContext = this;
    }
    miniomm::launchKernel((miniomm::KernelEntry *) kernel, arguments, gridSize, blockSize, sharedSize, stream);
    kernelLaunches++;
}

void CudaPlatform::PlatformData::initializeContexts(const System &) {
    /* [OMM-mem] first call: CudaContext::initialize() -> findMoleculeGroups() consults the ForceInfos registered so far */
    contextsInitialized = true;
}

/* ---- ContextImpl ---------------------------------------------------------------------------------------------------- */
ContextImpl::ContextImpl(Context &owner, const System &system, Integrator &integrator, Platform &platform, void *stream)
    : forceEvaluations(0), owner(owner), system(system), integrator(integrator), platform(platform), cu(NULL), hasMolecules(false) {
    cu = new CudaContext(system, platformData, stream);
    platformData.contexts.push_back(cu);
}

void ContextImpl::initializeIntegrator() {
    integrator.initialize(*this);
}

ContextImpl::~ContextImpl() {
    integrator.cleanup();
    delete cu;
}

const std::vector<std::vector<int> > &ContextImpl::getMolecules() const {
    if (hasMolecules)
        return molecules;
    const int n = system.getNumParticles();
    std::vector<std::vector<int> > adjacent(n);
    for (int f = 0; f < system.getNumForces(); f++) {
        const std::vector<std::pair<int, int> > bonds = system.getForce(f).getBondedParticles();
        for (size_t b = 0; b < bonds.size(); b++) {
            adjacent[bonds[b].first].push_back(bonds[b].second);
            adjacent[bonds[b].second].push_back(bonds[b].first);
        }
    }
    for (int c = 0; c < system.getNumConstraints(); c++) {
        int p1, p2;
        double d;
        system.getConstraintParameters(c, p1, p2, d);
        adjacent[p1].push_back(p2);
        adjacent[p2].push_back(p1);
    }
    std::vector<int> tag(n, -1);
    int count = 0;
    std::vector<int> stack;
    for (int i = 0; i < n; i++) {
        if (tag[i] != -1)
            continue;
        tag[i] = count;
        stack.assign(1, i);
        while (!stack.empty()) {
            const int a = stack.back();
            stack.pop_back();
            for (size_t k = 0; k < adjacent[a].size(); k++)
                if (tag[adjacent[a][k]] == -1) {
                    tag[adjacent[a][k]] = count;
                    stack.push_back(adjacent[a][k]);
                }
        }
        count++;
    }
    molecules.assign(count, std::vector<int>());
    for (int i = 0; i < n; i++)
        molecules[tag[i]].push_back(i);
    hasMolecules = true;
    return molecules;
}
