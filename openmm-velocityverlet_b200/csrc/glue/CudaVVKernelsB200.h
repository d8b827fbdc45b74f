// CudaVVKernelsB200.h -- the seven CUDA-platform kernel classes of the velocity-Verlet plugin, re-implemented as
// thin forwards to libvvb200 (include/vvb200.h).  Drop-in for platforms/cuda/include/CudaVVKernels.h of the
// reference: same class names, same constructors, same virtual interface (openmmapi/include/openmm/VVKernels.h is
// used unchanged), so CudaVVKernelFactory.cpp compiles against it as is.
#ifndef CUDA_VV_KERNELS_B200_H_
#define CUDA_VV_KERNELS_B200_H_

#include <memory>

#include "openmm/VVKernels.h"
#include "CudaContext.h"
#include "CudaArray.h"
#include "vvb200.h"

namespace OpenMM {

// One per CudaContext: the plan and the per-step bookkeeping shared by the seven kernel objects (the reference shares
// `forceExtra` between them the same way, CudaVVKernels.h:86,146).
struct VVB200Shared {
    vvb200_plan *plan = nullptr;
    bool constrained = false;     // OpenMM constraints / virtual sites sit between the sub-steps
    bool hasNH = false;
    unsigned int randomIndex = 0; // what prepareRandomNumbers returned this step
    double stepSize = -1.0;
    bool stepSizePublished = false;   // OpenMM's own step-size array holds stepSize
    ~VVB200Shared() { vvb200_plan_destroy(plan); }
    static std::shared_ptr<VVB200Shared> get(CudaContext &cu, bool create);
};

class CudaIntegrateMiddleStepKernel : public IntegrateMiddleStepKernel {
public:
    CudaIntegrateMiddleStepKernel(std::string name, const Platform &platform, CudaContext &cu)
        : IntegrateMiddleStepKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force);
    void firstIntegrate(ContextImpl &context, const VVIntegrator &integrator);
    void resetExtraForce(ContextImpl &context, const VVIntegrator &integrator);
    void secondIntegrate(ContextImpl &context, const VVIntegrator &integrator);
    double computeKineticEnergy(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

class CudaIntegrateVVStepKernel : public IntegrateVVStepKernel {
public:
    CudaIntegrateVVStepKernel(std::string name, const Platform &platform, CudaContext &cu)
        : IntegrateVVStepKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force);
    void firstIntegrate(ContextImpl &context, const VVIntegrator &integrator);
    void resetExtraForce(ContextImpl &context, const VVIntegrator &integrator);
    void secondIntegrate(ContextImpl &context, const VVIntegrator &integrator);
    double computeKineticEnergy(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

class CudaModifyDrudeNoseKernel : public ModifyDrudeNoseKernel {
public:
    CudaModifyDrudeNoseKernel(std::string name, const Platform &platform, CudaContext &cu)
        : ModifyDrudeNoseKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force);
    void scaleVelocity(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

class CudaModifyDrudeLangevinKernel : public ModifyDrudeLangevinKernel {
public:
    CudaModifyDrudeLangevinKernel(std::string name, const Platform &platform, CudaContext &cu)
        : ModifyDrudeLangevinKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, const DrudeForce *force, Kernel &vvKernel);
    void applyLangevinForce(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

class CudaModifyImageChargeKernel : public ModifyImageChargeKernel {
public:
    CudaModifyImageChargeKernel(std::string name, const Platform &platform, CudaContext &cu)
        : ModifyImageChargeKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator);
    void updateImagePositions(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

class CudaModifyElectricFieldKernel : public ModifyElectricFieldKernel {
public:
    CudaModifyElectricFieldKernel(std::string name, const Platform &platform, CudaContext &cu)
        : ModifyElectricFieldKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, Kernel &vvKernel);
    void applyElectricForce(ContextImpl &context, const VVIntegrator &integrator);
private:
    CudaContext &cu;
};

class CudaModifyCosineAccelerateKernel : public ModifyCosineAccelerateKernel {
public:
    CudaModifyCosineAccelerateKernel(std::string name, const Platform &platform, CudaContext &cu)
        : ModifyCosineAccelerateKernel(name, platform), cu(cu) {}
    void initialize(const System &system, const VVIntegrator &integrator, Kernel &vvKernel);
    void applyCosineForce(ContextImpl &context, const VVIntegrator &integrator);
    void calcVelocityBias(ContextImpl &context, const VVIntegrator &integrator);
    void removeVelocityBias(ContextImpl &context, const VVIntegrator &integrator);
    void restoreVelocityBias(ContextImpl &context, const VVIntegrator &integrator);
    void calcViscosity(ContextImpl &context, const VVIntegrator &integrator, double &vMax, double &invVis);
private:
    CudaContext &cu;
    std::shared_ptr<VVB200Shared> sh;
};

}  // namespace OpenMM
#endif
