"""Host-side mirror of the reference's user-facing interface, for hosts without OpenMM.

`VVIntegrator` keeps the names, argument meaning, defaults and error behaviour of OpenMM::VVIntegrator
(openmmapi/include/openmm/VVIntegrator.h:49-507, openmmapi/src/VVIntegrator.cpp) and of its SWIG surface
(python/velocityverletplugin.i:83-129, including the `int getFriction()` / `setDrudeFriction(int)` quirks the Python
user actually sees); `System` and `Context` stand in for the three OpenMM objects the integrator reads (particle
masses, bonds/constraints, DrudeForce pairs, CMMotionRemover) and for the Context that owns the device arrays.

    system = System(masses, bonds=..., drude_pairs=..., constraints=..., cm_motion_remover=True)
    integrator = VVIntegrator(333.0, 10.0, 1.0, 40.0, 0.001)            # T, freq, T_drude, freq_drude, dt
    integrator.setMaxDrudeDistance(0.02)
    context = Context(system, integrator, precision="mixed")            # -> VVIntegrator.initialize
    context.setState(host_state)                                        # posq / velm / force in OpenMM's layouts
    integrator.step(100)

Everything numerical happens in libvvb200.so (no compute here); forces are whatever `context.setForces` last set
(force evaluation is OpenMM's and out of scope)."""
import numpy as np

from ._cabi import Params, Plan, VVB200Error, propagate_nh_chain
from .buffers import DeviceBuffers
from .system import SystemSpec


class OpenMMException(RuntimeError):
    """what the reference throws (surfaced through SWIG as a Python exception)"""


class System:
    """The parts of OpenMM::System / DrudeForce / ContextImpl::getMolecules the integrator reads."""

    def __init__(self, masses, bonds=(), drude_pairs=(), constraints=(), cm_motion_remover=False, num_drude_forces=None):
        self.masses = np.asarray(masses, dtype=np.float64)
        self.bonds = np.asarray(bonds, dtype=np.int32).reshape(-1, 2)
        self.drude_pairs = np.asarray(drude_pairs, dtype=np.int32).reshape(-1, 2)       # (Drude, parent)
        self.constraints = np.asarray(constraints, dtype=np.int32).reshape(-1, 2)
        self.cm_motion_remover = bool(cm_motion_remover)
        self.num_drude_forces = (1 if self.drude_pairs.size else 0) if num_drude_forces is None else num_drude_forces

    def getNumParticles(self):
        return int(self.masses.size)

    @classmethod
    def from_spec(cls, spec):
        return cls(spec.masses, spec.bonds, spec.drude_pairs, spec.constraints, spec.has_cmm)


class VVIntegrator:
    def __init__(self, temperature, frequency, drudeTemperature, drudeFrequency, stepSize, numNHChains=3, loopsPerStep=1):
        # VVIntegrator.cpp:45-69
        self.setTemperature(temperature)
        self.setFrequency(frequency)
        self.setDrudeTemperature(drudeTemperature)
        self.setDrudeFrequency(drudeFrequency)
        self.setStepSize(stepSize)
        self.setNumNHChains(numNHChains)
        self.setLoopsPerStep(loopsPerStep)
        self.setConstraintTolerance(1e-5)
        self.setMaxDrudeDistance(0)
        self.setFriction(5.0)
        self.setDrudeFriction(20.0)
        self.setRandomNumberSeed(0)
        self.setMirrorLocation(0.0)
        self.setElectricField(0.0)
        self.setCosAcceleration(0.0)
        self.setUseCOMTempGroup(False)
        self.setUseMiddleScheme(True)
        self.setDebugEnabled(False)
        self._autoSetCOMTempGroup = True
        self._autoSetFriction = True
        self._particlesLD, self._imagePairs, self._particlesElectrolyte = [], [], []
        self._particlesNH, self._moleculesNH = [], []
        self._context = None

    # ---- plain accessors (VVIntegrator.h:62-431) ----
    def getTemperature(self): return self._temperature
    def setTemperature(self, temp): self._temperature = float(temp)
    def getFrequency(self): return self._frequency
    def setFrequency(self, tau): self._frequency = float(tau)
    def getDrudeTemperature(self): return self._drudeTemperature
    def setDrudeTemperature(self, temp): self._drudeTemperature = float(temp)
    def getDrudeFrequency(self): return self._drudeFrequency
    def setDrudeFrequency(self, tau): self._drudeFrequency = float(tau)
    def getStepSize(self): return self._stepSize
    def setStepSize(self, size): self._stepSize = float(size)
    def getConstraintTolerance(self): return self._constraintTolerance
    def setConstraintTolerance(self, tol): self._constraintTolerance = float(tol)
    def getNumNHChains(self): return self._numNHChains
    def setNumNHChains(self, numChains): self._numNHChains = int(numChains)
    def getLoopsPerStep(self): return self._loopsPerStep
    def setLoopsPerStep(self, loops): self._loopsPerStep = int(loops)
    def getMaxDrudeDistance(self): return self._maxDrudeDistance
    def setMaxDrudeDistance(self, distance): self._maxDrudeDistance = float(distance)
    def getRandomNumberSeed(self): return self._seed
    def setRandomNumberSeed(self, seed): self._seed = int(seed)
    def getMirrorLocation(self): return self._mirror
    def setMirrorLocation(self, z): self._mirror = float(z)
    def getElectricField(self): return self._efield
    def setElectricField(self, field): self._efield = float(field)      # kJ/(nm e); 1 V/nm = 1.60217662e-22
    def getCosAcceleration(self): return self._cosAcceleration
    def setCosAcceleration(self, a): self._cosAcceleration = float(a)
    def getUseMiddleScheme(self): return self._useMiddleScheme
    def setUseMiddleScheme(self, use): self._useMiddleScheme = bool(use)
    def getDebugEnabled(self): return self._debug
    def setDebugEnabled(self, enabled): self._debug = bool(enabled)

    def getUseCOMTempGroup(self): return self._useCOMTempGroup

    def setUseCOMTempGroup(self, use):                                   # VVIntegrator.h:176-179
        self._useCOMTempGroup = bool(use)
        self._autoSetCOMTempGroup = False

    def getFriction(self): return int(self._friction)                    # `int getFriction()` in the SWIG file (:110)

    def setFriction(self, fric):                                         # VVIntegrator.h:214-217
        self._friction = float(fric)
        self._autoSetFriction = False

    def getDrudeFriction(self): return int(self._drudeFriction)          # :112

    def setDrudeFriction(self, fric):                                    # `setDrudeFriction(int)` in the SWIG file (:113)
        self._drudeFriction = float(int(fric))
        self._autoSetFriction = False

    def addParticleLangevin(self, particle):                             # VVIntegrator.h:199-202
        self._particlesLD.append(int(particle))
        return len(self._particlesLD)

    def addImagePair(self, image, parent):                               # VVIntegrator.cpp:75-79
        self._imagePairs.append((int(image), int(parent)))
        return len(self._imagePairs)

    def addParticleElectrolyte(self, particle):                          # VVIntegrator.h:302-305 (void in SWIG)
        self._particlesElectrolyte.append(int(particle))

    # ---- what initialize() derives (VVIntegrator.h:317-375) ----
    def getParticlesNH(self): return list(self._particlesNH)
    def getParticlesLD(self): return list(self._particlesLD)
    def getMoleculesNH(self): return list(self._moleculesNH)
    def getImagePairs(self): return list(self._imagePairs)
    def getParticlesElectrolyte(self): return list(self._particlesElectrolyte)
    def isParticleNH(self, i): return i in set(self._particlesNH)
    def isParticleLD(self, i): return i in set(self._particlesLD)
    def isParticleImage(self, i): return i in {im for im, _ in self._imagePairs}
    def getNumMolecules(self): return self._spec.n_mol
    def getParticleMolId(self, particle): return int(self._spec.mol_id[particle])
    def getMoleculeInvMass(self, molid): return float(self._plan.f64_array("moleculeInvMasses")[molid])

    def propagateNHChain(self, eta, etaDot, etaDotDot, etaMass, ke2, ke2Target, tTarget):
        """VVIntegrator::propagateNHChain (public, VVIntegrator.cpp:340-376): arrays updated in place, returns the factor"""
        return propagate_nh_chain(self._stepSize, self._loopsPerStep, eta, etaDot, etaDotDot, etaMass, ke2, ke2Target, tTarget)

    # ---- Context hooks ----
    def _initialize(self, context):
        """VVIntegrator::initialize (VVIntegrator.cpp:92-188)"""
        if self._context is not None and self._context is not context:
            raise OpenMMException("This Integrator is already bound to a context")
        system = context.system
        if system.num_drude_forces > 1:
            raise OpenMMException("The System contains multiple DrudeForces")
        has_drude = system.drude_pairs.shape[0] > 0
        # auto-defaults go through the public setters, which clear the auto flags
        if self._autoSetCOMTempGroup:
            self.setUseCOMTempGroup(has_drude)
        if self._autoSetFriction:
            self.setFriction(5.0 if has_drude else 1.0)
        spec = SystemSpec(n=system.getNumParticles(), masses=system.masses, bonds=system.bonds,
                          drude_pairs=system.drude_pairs, constraints=system.constraints, has_cmm=system.cm_motion_remover,
                          langevin=np.array(self._particlesLD, np.int32),
                          image_pairs=np.array(self._imagePairs, np.int32).reshape(-1, 2),
                          electrolyte=np.array(self._particlesElectrolyte, np.int32)).finalize()
        params = Params(temperature=self._temperature, frequency=self._frequency, drude_temperature=self._drudeTemperature,
                        drude_frequency=self._drudeFrequency, step_size=self._stepSize, num_nh_chains=self._numNHChains,
                        loops_per_step=self._loopsPerStep, use_com_temp_group=self._useCOMTempGroup,
                        use_middle_scheme=self._useMiddleScheme, max_drude_distance=self._maxDrudeDistance,
                        friction=self._friction, drude_friction=self._drudeFriction, mirror_location=self._mirror,
                        electric_field=self._efield, cos_acceleration=self._cosAcceleration)
        try:
            plan = Plan(spec, params, context.precision)
        except VVB200Error as e:
            if e.code == 2:                       # the reference's configuration errors, same texts
                raise OpenMMException(e.message) from None
            raise
        self._spec, self._params, self._plan, self._context = spec, params, plan, context
        self._particlesNH = plan.int_array("particlesNH").tolist()
        self._moleculesNH = plan.int_array("moleculesNH").tolist()
        self._randomIndex = 0
        return spec

    def step(self, steps):
        """VVIntegrator::step (VVIntegrator.cpp:223-230): `steps` steps with the forces currently set in the context"""
        ctx = self._context
        if ctx is None:
            raise OpenMMException("This Integrator is not bound to a context!")
        if self._stepSize != self._params.step_size:      # the reference re-reads the step size every step
            self._plan.set_step_size(self._stepSize)
            self._params.step_size = self._stepSize
        ctx._require_device()
        inv_box_z = 1.0 / ctx.box[2] if self._cosAcceleration != 0 else 0.0
        self._randomIndex = self._plan.step(ctx.buffers, steps=int(steps), random_index=self._randomIndex, inv_box_z=inv_box_z)

    def getGroupTemperatures(self):
        """{"temperature": [T_atom, T_COM, T_Drude][:numGroups], "ke2", "dof"} of the current velocities, measured on the
        device (what examples/ommhelper/reporter/drudetemperaturereporter.py:98-129 computes on the host with numpy)"""
        ctx = self._context
        ctx._require_device()
        inv_box_z = 1.0 / ctx.box[2] if self._cosAcceleration != 0 else 0.0
        return self._plan.measure_temperatures(ctx.buffers, inv_box_z=inv_box_z)

    def getViscosity(self):
        """[vMax, 1/viscosity] (VVIntegrator.cpp:378-383)"""
        v, inv = self._plan.viscosity(self._context.box)
        return [v, inv]


class Context:
    """Owns the device copies of OpenMM's arrays (posq, posqCorrection, velm, force, random) and binds the integrator."""

    def __init__(self, system, integrator, precision="mixed"):
        self.system, self.integrator, self.precision = system, integrator, precision
        self.buffers, self.box = None, (1.0, 1.0, 1.0)
        self.spec = integrator._initialize(self)

    def setState(self, host_state):
        """host arrays in OpenMM's layouts (system.HostState) -> device"""
        self.box = host_state.box
        self.buffers = DeviceBuffers(host_state)
        self.integrator._plan.upload()

    def setForces(self, force_fixed_point):
        import torch
        self.buffers.force.copy_(torch.from_numpy(np.ascontiguousarray(force_fixed_point)))

    def getState(self):
        return self.buffers.to_host()

    # ---- OpenMM's Context.createCheckpoint / loadCheckpoint: positions, velocities, box AND the integrator's own state
    #      (Integrator::createCheckpoint hook; the reference does not implement it and restarts its NH chains cold) ----
    def createCheckpoint(self):
        import io
        st = self.getState()
        out = io.BytesIO()
        np.savez(out, posq=st.posq, corr=st.corr if st.corr is not None else np.zeros(0), velm=st.velm,
                 box=np.array(self.box), random_index=np.array([self.integrator._randomIndex]),
                 integrator=np.frombuffer(self.integrator._plan.checkpoint_save(), dtype=np.uint8))
        return out.getvalue()

    def loadCheckpoint(self, blob):
        import io
        import torch
        self._require_device()
        z = np.load(io.BytesIO(blob))
        self.buffers.posq.copy_(torch.from_numpy(z["posq"]))
        if self.buffers.corr is not None:
            self.buffers.corr.copy_(torch.from_numpy(z["corr"]))
        self.buffers.velm.copy_(torch.from_numpy(z["velm"]))
        self.box = tuple(float(x) for x in z["box"])
        self.integrator._randomIndex = int(z["random_index"][0])
        self.integrator._plan.checkpoint_load(z["integrator"].tobytes())

    def _require_device(self):
        if self.buffers is None:
            raise OpenMMException("Particle positions have not been set")
