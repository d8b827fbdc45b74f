// openmm_shim_all.h -- COMPILE-CHECK STAND-IN for the OpenMM 8.1.2 headers the glue includes.
//
// OpenMM is not installed in the build image, so CudaVVKernelsB200.cpp could otherwise not even be parsed.  This
// file declares, from memory of the OpenMM 8.1.2 API [OMM-mem], just the classes and members the glue and the
// reference's own VVIntegrator.h / VVKernels.h touch, with no implementations.  It is only ever used by
// `make -C csrc/glue check` (g++ -fsyntax-only); a real build uses the real headers.
#ifndef VVB200_OPENMM_SHIM_ALL_H_
#define VVB200_OPENMM_SHIM_ALL_H_
#include <exception>
#include <map>
#include <string>
#include <utility>
#include <vector>

#define OPENMM_EXPORT
#define OPENMM_EXPORT_DRUDE
typedef unsigned long long CUdeviceptr;
typedef struct CUstream_st *CUstream;
struct double4 { double x, y, z, w; };
struct float4 { float x, y, z, w; };

namespace OpenMM {
class Vec3 { public: double v[3]; double operator[](int i) const { return v[i]; } };
class OpenMMException : public std::exception {
public:
    explicit OpenMMException(const std::string &m) : msg(m) {}
    const char *what() const noexcept override { return msg.c_str(); }
private:
    std::string msg;
};
class State { public: enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 }; };
class Force { public: virtual ~Force() {} };
class CMMotionRemover : public Force {};
class DrudeForce : public Force {
public:
    int getNumParticles() const;
    void getParticleParameters(int index, int &particle, int &particle1, int &particle2, int &particle3, int &particle4,
                               double &charge, double &polarizability, double &aniso12, double &aniso34) const;
};
class System {
public:
    int getNumParticles() const;
    double getParticleMass(int index) const;
    int getNumConstraints() const;
    void getConstraintParameters(int index, int &particle1, int &particle2, double &distance) const;
    int getNumForces() const;
    const Force &getForce(int index) const;
    bool isVirtualSite(int index) const;
};
class Platform {
public:
    static Platform &getPlatformByName(const std::string &name);
    static void registerPlatform(Platform *platform);
    void registerKernelFactory(const std::string &name, class KernelFactory *factory);
};
class ContextImpl {
public:
    const std::vector<std::vector<int> > &getMolecules() const;
    void *getPlatformData();
    const System &getSystem() const;
};
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform &platform);
    virtual ~KernelImpl() {}
};
class Kernel {
public:
    Kernel();
    KernelImpl &getImpl();
    template <class T> T &getAs() { return dynamic_cast<T &>(getImpl()); }
};
class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl *createKernelImpl(std::string name, const Platform &platform, ContextImpl &context) const = 0;
};
class Integrator {
public:
    virtual ~Integrator() {}
    double getStepSize() const;
    void setStepSize(double size);
    double getConstraintTolerance() const;
    void setConstraintTolerance(double tol);
    virtual void step(int steps) = 0;
protected:
    virtual void initialize(ContextImpl &context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual void stateChanged(State::DataType changed) {}
    virtual double computeKineticEnergy() = 0;
    virtual bool kineticEnergyRequiresForce() const { return true; }
    ContextImpl *context;
    class Context *owner;
};

// ---- CUDA platform ----
class CudaArray {
public:
    CUdeviceptr &getDevicePointer();
};
class CudaIntegrationUtilities {
public:
    void initRandomNumberGenerator(unsigned int seed);
    int prepareRandomNumbers(int numValues);
    CudaArray &getRandom();
    CudaArray &getPosDelta();
    CudaArray &getStepSize();
    void setNextStepSize(double size);
    void applyConstraints(double tol);
    void applyVelocityConstraints(double tol);
    void computeVirtualSites();
    double computeKineticEnergy(double timeShift);
};
class CudaPlatform : public Platform {
public:
    class PlatformData {
    public:
        void initializeContexts(const System &system);
        std::vector<class CudaContext *> contexts;
    };
};
// openmm/common/ComputeForceInfo.h + CudaForceInfo.h [OMM-mem]: what CudaContext::findMoleculeGroups consults before
// it lets reorderAtoms swap two molecules
class ComputeForceInfo {
public:
    virtual ~ComputeForceInfo() {}
    virtual bool areParticlesIdentical(int particle1, int particle2) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int index, std::vector<int> &particles) {}
    virtual bool areGroupsIdentical(int group1, int group2) { return true; }
};
class CudaForceInfo : public ComputeForceInfo {};
class CudaContext {
public:
    static const int ThreadBlockSize = 64;
    void addForce(ComputeForceInfo *force);
    int getNumAtoms() const;
    int getPaddedNumAtoms() const;
    bool getUseDoublePrecision() const;
    bool getUseMixedPrecision() const;
    int getNumThreadBlocks() const;
    CudaArray &getPosq();
    CudaArray &getPosqCorrection();
    CudaArray &getVelm();
    CudaArray &getForce();
    CUstream getCurrentStream();
    double4 getPeriodicBoxSize() const;
    void *getInvPeriodicBoxSizePointer();
    CudaIntegrationUtilities &getIntegrationUtilities();
    CudaPlatform::PlatformData &getPlatformData();
    void setAsCurrent();
    void reorderAtoms();
    double getTime();
    void setTime(double t);
    long long getStepCount();
    void setStepCount(long long c);
};
class ContextSelector {
public:
    explicit ContextSelector(CudaContext &cu);
};
}  // namespace OpenMM
#endif
