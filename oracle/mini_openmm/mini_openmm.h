/*
 * mini_openmm.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A small WORKING stand-in for the slice of OpenMM 8.1.2 the velocity-Verlet plugin touches, so that the reference's
 * own host code -- openmmapi/src/VVIntegrator.cpp, platforms/cuda/src/CudaVVKernels.cpp and CudaVVKernelFactory.cpp,
 * compiled UNCHANGED from where they lie under /root/reference -- and the B200 glue (csrc/glue/CudaVVKernelsB200.cpp)
 * can be linked and RUN here, where OpenMM itself (headers, libOpenMM, libOpenMMCUDA, libOpenMMDrude) does not exist.
 *
 * What it is: System / Force / DrudeForce / CMMotionRemover with the getters the plugin calls; Platform + KernelFactory
 * registry + Kernel handles; Integrator base class; Context / ContextImpl (molecules from bonds, Drude pairs and
 * constraints; calcForcesAndEnergy is a no-op: forces are whatever the test put into the force array, i.e. frozen);
 * CudaContext owning posq / posqCorrection / velm / force in OpenMM's layouts; CudaArray with create / upload /
 * download / getDevicePointer (every upload is also recorded by array name, which is how the tests read the index
 * arrays the reference's initialize() methods build); CudaContext::createModule / getKernel / executeKernel with
 * OpenMM's launch geometry, backed by the reference kernels compiled ahead of time (plugin_kernels.cpp) instead of NVRTC;
 * CudaIntegrationUtilities with posDelta, stepSize, an injected random stream with prepareRandomNumbers' index
 * bookkeeping, and applyConstraints / applyVelocityConstraints implemented by oracle/constraint_standin.h.
 * Semantics marked [OMM-mem] are from memory of the OpenMM 8.1.2 sources.
 *
 * What it is not: OpenMM.  No force field, no neighbour lists, no atom reordering (reorderAtoms is a no-op), no virtual
 * sites, no real SHAKE/SETTLE/CCMA, no NVRTC.  Two flavours, chosen at compile time: host memory (default; kernels run
 * through a SIMT shim) and -DMINIOMM_CUDA (device memory, real launches on a cudaStream_t).
 * Precision mode: -DVVREF_SINGLE / -DVVREF_MIXED (default) / -DVVREF_DOUBLE, like oracle/ref_kernels.inc.h.
 */
#ifndef MINI_OPENMM_H_
#define MINI_OPENMM_H_

#include <math.h>      /* SimTKOpenMMRealType.h pulls <cmath> into the reference sources */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include <vector_functions.h>     /* int2, double4, make_int2 ... (host-usable CUDA headers) */
#include <vector_types.h>

#define OPENMM_EXPORT
#define OPENMM_EXPORT_DRUDE

/* driver-API handle types as the reference's headers spell them; never passed to a real driver call */
typedef unsigned long long CUdeviceptr;
typedef struct MiniModule_st *CUmodule;
typedef struct MiniKernel_st *CUfunction;
typedef struct CUstream_st *CUstream;

#if defined(VVREF_DOUBLE)
#define MINIOMM_MODE 2
#elif defined(VVREF_SINGLE)
#define MINIOMM_MODE 0
#else
#define MINIOMM_MODE 1
#endif

#include "miniomm_backend.h"

namespace OpenMM {

class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string &m) : msg(m) {}
    ~OpenMMException() throw() {}
    const char *what() const throw() { return msg.c_str(); }
private:
    std::string msg;
};

/* openmm/internal/AssertionUtilities.h */
#define ASSERT_VALID_INDEX(index, vector) { if ((index) < 0 || (index) >= (int) (vector).size()) throw OpenMM::OpenMMException("Assertion failure: Index out of range"); }

/* openmm/reference/SimTKOpenMMRealType.h [OMM-mem] */
#ifndef BOLTZ
#define ANGSTROM (1e-10)
#define KILO (1e3)
#define NANO (1e-9)
#define PICO (1e-12)
#define A2NM (ANGSTROM / NANO)
#define NM2A (NANO / ANGSTROM)
#define RAD2DEG (180.0 / M_PI)
#define CAL2JOULE (4.184)
#define E_CHARGE (1.602176634e-19)
#define AMU (1.66053906660e-27)
#define BOLTZMANN (1.380649e-23)
#define AVOGADRO (6.02214076e23)
#define RGAS (BOLTZMANN * AVOGADRO)
#define BOLTZ (RGAS / KILO)
#endif

class Vec3 {
public:
    Vec3() { v[0] = v[1] = v[2] = 0; }
    Vec3(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double operator[](int i) const { return v[i]; }
    double &operator[](int i) { return v[i]; }
private:
    double v[3];
};

class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 };
};

class Force {
public:
    virtual ~Force() {}
    /* bonded particle pairs this force contributes to ContextImpl::getMolecules (ForceImpl::getBondedParticles) */
    virtual std::vector<std::pair<int, int> > getBondedParticles() const { return std::vector<std::pair<int, int> >(); }
};

class CMMotionRemover : public Force {
public:
    explicit CMMotionRemover(int frequency = 1) : frequency(frequency) {}
    int getFrequency() const { return frequency; }
private:
    int frequency;
};

/* any bonded force of a real System, reduced to what matters here: which particles it ties into one molecule */
class BondListForce : public Force {
public:
    void addBond(int a, int b) { bonds.push_back(std::make_pair(a, b)); }
    std::vector<std::pair<int, int> > getBondedParticles() const { return bonds; }
private:
    std::vector<std::pair<int, int> > bonds;
};

class DrudeForce : public Force {
public:
    int addParticle(int particle, int particle1, int particle2, int particle3, int particle4, double charge,
                    double polarizability, double aniso12, double aniso34) {
        Entry e = {particle, particle1, particle2, particle3, particle4, charge, polarizability, aniso12, aniso34};
        entries.push_back(e);
        return (int) entries.size() - 1;
    }
    int getNumParticles() const { return (int) entries.size(); }
    void getParticleParameters(int index, int &particle, int &particle1, int &particle2, int &particle3, int &particle4,
                               double &charge, double &polarizability, double &aniso12, double &aniso34) const {
        ASSERT_VALID_INDEX(index, entries);
        const Entry &e = entries[index];
        particle = e.p; particle1 = e.p1; particle2 = e.p2; particle3 = e.p3; particle4 = e.p4;
        charge = e.charge; polarizability = e.pol; aniso12 = e.a12; aniso34 = e.a34;
    }
    std::vector<std::pair<int, int> > getBondedParticles() const {       /* DrudeForceImpl::getBondedParticles [OMM-mem] */
        std::vector<std::pair<int, int> > b;
        for (size_t i = 0; i < entries.size(); i++) b.push_back(std::make_pair(entries[i].p, entries[i].p1));
        return b;
    }
private:
    struct Entry { int p, p1, p2, p3, p4; double charge, pol, a12, a34; };
    std::vector<Entry> entries;
};

class System {
public:
    ~System() { for (size_t i = 0; i < forces.size(); i++) delete forces[i]; }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size() - 1; }
    int getNumParticles() const { return (int) masses.size(); }
    double getParticleMass(int index) const { ASSERT_VALID_INDEX(index, masses); return masses[index]; }
    int addConstraint(int p1, int p2, double distance) {
        Constraint c = {p1, p2, distance};
        constraints.push_back(c);
        return (int) constraints.size() - 1;
    }
    int getNumConstraints() const { return (int) constraints.size(); }
    void getConstraintParameters(int index, int &particle1, int &particle2, double &distance) const {
        ASSERT_VALID_INDEX(index, constraints);
        particle1 = constraints[index].p1; particle2 = constraints[index].p2; distance = constraints[index].d;
    }
    int addForce(Force *force) { forces.push_back(force); return (int) forces.size() - 1; }      /* takes ownership */
    int getNumForces() const { return (int) forces.size(); }
    const Force &getForce(int index) const { ASSERT_VALID_INDEX(index, forces); return *forces[index]; }
    bool isVirtualSite(int) const { return false; }
private:
    struct Constraint { int p1, p2; double d; };
    std::vector<double> masses;
    std::vector<Constraint> constraints;
    std::vector<Force *> forces;
};

class Platform;
class ContextImpl;
class Context;

class KernelImpl {
public:
    KernelImpl(std::string name, const Platform &platform) : name(name), platform(&platform) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform &getPlatform() { return *platform; }
private:
    std::string name;
    const Platform *platform;
};

class Kernel {
public:
    Kernel() {}
    explicit Kernel(KernelImpl *impl) : impl(impl) {}
    KernelImpl &getImpl() { if (!impl) throw OpenMMException("Kernel has no implementation"); return *impl; }
    const KernelImpl &getImpl() const { return *impl; }
    template <class T> T &getAs() { return dynamic_cast<T &>(getImpl()); }
    template <class T> const T &getAs() const { return dynamic_cast<const T &>(*impl); }
private:
    std::shared_ptr<KernelImpl> impl;
};

class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl *createKernelImpl(std::string name, const Platform &platform, ContextImpl &context) const = 0;
};

class Platform {
public:
    virtual ~Platform() {}
    virtual const std::string &getName() const = 0;
    void registerKernelFactory(const std::string &name, KernelFactory *factory) { factories[name] = factory; }
    Kernel createKernel(const std::string &name, ContextImpl &context) const {
        std::map<std::string, KernelFactory *>::const_iterator it = factories.find(name);
        if (it == factories.end())
            throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel: " + name);
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }
    static std::vector<Platform *> &all() { static std::vector<Platform *> v; return v; }
    static void registerPlatform(Platform *platform) { all().push_back(platform); }
    static Platform &getPlatformByName(const std::string &name) {
        for (size_t i = 0; i < all().size(); i++)
            if (all()[i]->getName() == name) return *all()[i];
        throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    }
private:
    std::map<std::string, KernelFactory *> factories;
};

class Integrator {
public:
    Integrator() : stepSize(0), constraintTol(1e-5), context(NULL), owner(NULL) {}
    virtual ~Integrator() {}
    virtual double getStepSize() const { return stepSize; }
    virtual void setStepSize(double size) { stepSize = size; }
    virtual double getConstraintTolerance() const { return constraintTol; }
    virtual void setConstraintTolerance(double tol) { constraintTol = tol; }
    virtual void step(int steps) = 0;
protected:
    friend class ContextImpl;
    friend class Context;
    virtual void initialize(ContextImpl &context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual void stateChanged(State::DataType) {}
    virtual double computeKineticEnergy() = 0;
    virtual bool kineticEnergyRequiresForce() const { return true; }
    ContextImpl *context;
    Context *owner;
private:
    double stepSize, constraintTol;
};

/* ---- CUDA platform ------------------------------------------------------------------------------------------- */
class CudaContext;

class CudaArray {
public:
    template <class T> static CudaArray *create(CudaContext &cu, int size, const std::string &name) {
        return new CudaArray(cu, size, (int) sizeof(T), name);
    }
    CudaArray(CudaContext &cu, int size, int elementSize, const std::string &name);
    explicit CudaArray(CudaContext &cu) : cu(cu), size(0), elementSize(1), name("uninitialized"), pointer(0) {}
    ~CudaArray() { if (pointer) miniomm::deviceFree((void *) pointer); }
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string &getName() const { return name; }
    CUdeviceptr &getDevicePointer() { return pointer; }
    template <class T> void upload(const std::vector<T> &data, bool convert = false) {
        (void) convert;
        if ((int) sizeof(T) != elementSize || (int) data.size() != size)
            throw OpenMMException("Error uploading array " + name + ": The specified vector does not match the size of the array");
        upload(data.data());
    }
    void upload(const void *data, bool blocking = true);
    template <class T> void download(std::vector<T> &data) const {
        if ((int) sizeof(T) != elementSize)
            throw OpenMMException("Error downloading array " + name + ": element size mismatch");
        if ((int) data.size() != size) data.resize(size);
        download(data.data());
    }
    void download(void *data, bool blocking = true) const;
private:
    CudaContext &cu;
    int size, elementSize;
    std::string name;
    CUdeviceptr pointer;
};

class CudaIntegrationUtilities {
public:
    explicit CudaIntegrationUtilities(CudaContext &cu);
    ~CudaIntegrationUtilities() { delete posDelta; delete stepSize; delete random; }
    void initRandomNumberGenerator(unsigned int seed) { randomSeed = seed; }
    /* [OMM-mem] returns the index of the first of `numValues` fresh float4s in getRandom() and advances the cursor; the
     * stream here is injected by the test (setRandomStream) and sized so that it never has to be regenerated */
    int prepareRandomNumbers(int numValues) {
        if (randomPos + numValues > (int) randomCount)
            throw OpenMMException("mini-OpenMM: injected random stream exhausted");
        const int old = randomPos;
        randomPos += numValues;
        return old;
    }
    void setRandomStream(const float *values4, size_t count);
    CudaArray &getRandom() { return *random; }
    CudaArray &getPosDelta() { return *posDelta; }
    CudaArray &getStepSize() { return *stepSize; }
    void setNextStepSize(double size);           /* [OMM-mem] uploads (0, size) as mixed2 when it changed */
    void applyConstraints(double tol);           /* oracle/constraint_standin.h on posDelta */
    void applyVelocityConstraints(double tol);   /* ... on velm */
    void computeVirtualSites() {}
    double computeKineticEnergy(double timeShift);
    miniomm::Standin standin;
    long constraintCalls, velocityConstraintCalls;
private:
    CudaContext &cu;
    CudaArray *posDelta, *stepSize, *random;
    size_t randomCount;
    int randomPos;
    unsigned int randomSeed;
    double lastStepSize;
};

class ComputeForceInfo {
public:
    virtual ~ComputeForceInfo() {}
    virtual bool areParticlesIdentical(int, int) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int, std::vector<int> &) {}
    virtual bool areGroupsIdentical(int, int) { return true; }
};
class CudaForceInfo : public ComputeForceInfo {};

class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        PlatformData() : contextsInitialized(false) {}
        void initializeContexts(const System &system);
        std::vector<CudaContext *> contexts;
        bool contextsInitialized;
    };
    const std::string &getName() const { static const std::string n = "CUDA"; return n; }
};

class CudaContext {
public:
    static const int ThreadBlockSize = 64;
    CudaContext(const System &system, CudaPlatform::PlatformData &data, void *stream);
    ~CudaContext();
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    bool getUseDoublePrecision() const { return MINIOMM_MODE == 2; }
    bool getUseMixedPrecision() const { return MINIOMM_MODE == 1; }
    int getNumThreadBlocks() const { return numThreadBlocks; }
    CudaArray &getPosq() { return *posq; }
    /* only allocated in mixed precision; otherwise an uninitialised array whose device pointer is 0 [OMM-mem] -- the
     * reference's middle-scheme kernels are handed it unconditionally (CudaVVKernels.cpp:180,205) */
    CudaArray &getPosqCorrection() { return *posqCorrection; }
    CudaArray &getVelm() { return *velm; }
    CudaArray &getForce() { return *force; }
    CUstream getCurrentStream() { return (CUstream) stream; }
    double4 getPeriodicBoxSize() const { return box; }
    void setPeriodicBoxSize(double x, double y, double z);
    void *getInvPeriodicBoxSizePointer() { return MINIOMM_MODE == 2 ? (void *) &invBoxDouble : (void *) &invBoxFloat; }
    CudaIntegrationUtilities &getIntegrationUtilities() { return *integration; }
    CudaPlatform::PlatformData &getPlatformData() { return platformData; }
    void setAsCurrent() {}
    void reorderAtoms() { reorderCalls++; }
    void addForce(ComputeForceInfo *info) { forceInfos.push_back(info); forceInfoAddedBeforeInit = forceInfoAddedBeforeInit || !platformData.contextsInitialized; }
    std::vector<ComputeForceInfo *> &getForceInfos() { return forceInfos; }
    double getTime() { return time; }
    void setTime(double t) { time = t; }
    long long getStepCount() { return stepCount; }
    void setStepCount(long long c) { stepCount = c; }
    std::string intToString(int value) const { std::ostringstream s; s << value; return s.str(); }
    std::string doubleToString(double value) const { std::ostringstream s; s.precision(16); s << value; return s.str(); }
    /* modules: the source text is ignored (the kernels were compiled ahead of time), the #defines become variables */
    CUmodule createModule(const std::string &source, const std::map<std::string, std::string> &defines, const char *optimizationFlags = NULL);
    CUmodule createModule(const std::string &source, const char *optimizationFlags = NULL) { return createModule(source, std::map<std::string, std::string>(), optimizationFlags); }
    CUfunction getKernel(CUmodule module, const std::string &name);
    /* [OMM-mem] CudaContext::executeKernel: block = ThreadBlockSize unless given, grid = min(ceil(work/block), numThreadBlocks) */
    void executeKernel(CUfunction kernel, void **arguments, int workUnits, int blockSize = -1, unsigned int sharedSize = 0);
    void *stream;
    long kernelLaunches, reorderCalls;
    bool forceInfoAddedBeforeInit;
    /* every CudaArray::upload, by array name (last upload wins): how tests read what initialize() built */
    std::map<std::string, std::vector<unsigned char> > uploads;
    std::map<std::string, int> uploadElementSize;
    std::map<std::string, long> defines;      /* union of this context's module #defines (consistent across modules) */
private:
    int numAtoms, paddedNumAtoms, numThreadBlocks;
    CudaPlatform::PlatformData &platformData;
    CudaArray *posq, *posqCorrection, *velm, *force;
    CudaIntegrationUtilities *integration;
    double4 box, invBoxDouble;
    float4 invBoxFloat;
    double time;
    long long stepCount;
    std::vector<ComputeForceInfo *> forceInfos;
    std::vector<CUmodule> modules;
};

class ContextSelector {
public:
    explicit ContextSelector(CudaContext &) {}
};

/* ---- Context -------------------------------------------------------------------------------------------------- */
class ContextImpl {
public:
    ContextImpl(Context &owner, const System &system, Integrator &integrator, Platform &platform, void *stream);
    ~ContextImpl();
    Context &getOwner() { return owner; }
    const System &getSystem() const { return system; }
    Integrator &getIntegrator() { return integrator; }
    Platform &getPlatform() { return platform; }
    void *getPlatformData() { return &platformData; }
    /* [OMM-mem] ContextImpl::getMolecules: connected components over every force's bonded pairs and the constraints,
     * molecules numbered by their lowest particle, particles ascending inside a molecule */
    const std::vector<std::vector<int> > &getMolecules() const;
    bool updateContextState() { return false; }                       /* no barostat / CMMotionRemover kernel here */
    double calcForcesAndEnergy(bool, bool, int = 0xFFFFFFFF) { forceEvaluations++; return 0.0; }    /* frozen forces */
    CudaContext &cuda() { return *cu; }
    void initializeIntegrator();
    long forceEvaluations;
private:
    Context &owner;
    const System &system;
    Integrator &integrator;
    Platform &platform;
    CudaPlatform::PlatformData platformData;
    CudaContext *cu;
    mutable std::vector<std::vector<int> > molecules;
    mutable bool hasMolecules;
};

class Context {
public:
    Context(const System &system, Integrator &integrator, Platform &platform, void *stream = NULL)
        : impl(new ContextImpl(*this, system, integrator, platform, stream)) { impl->initializeIntegrator(); }
    ~Context() { delete impl; }
    ContextImpl &getImpl() { return *impl; }
private:
    ContextImpl *impl;
};

}   // namespace OpenMM
#endif /* MINI_OPENMM_H_ */
