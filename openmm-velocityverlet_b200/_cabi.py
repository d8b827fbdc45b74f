"""ctypes bindings for include/vvb200.h.  No compute happens here."""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvvb200.so")
PRECISIONS = {"single": 0, "mixed": 1, "double": 2}
MAX_CHAINS = 16

# vvb200_plan_get_int_array ids
INT_ARRAYS = {
    "particlesNH": 0, "moleculesNH": 1, "particleMolId": 2, "drudePairs": 3, "sortedByMol": 4,
    "particlesInMolecules": 5, "normalNH": 6, "pairsNH": 7, "normalLD": 8, "pairsLD": 9,
    "imagePairs": 10, "electrolyte": 11, "tileStart": 12, "slotMeta": 13,
    "tileMolOffset": 14, "tileMolList": 15, "tileMolFrag": 16, "splitMolId": 17, "splitFragOffset": 18,
    "splitFragList": 19, "imageOf": 20,
}
F64_ARRAYS = {"moleculeMasses": 0, "moleculeInvMasses": 1, "dof": 2, "etaMass": 3, "NkbT": 4, "invMassTotal": 5, "dofGlobal": 6}


class VVB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"vvb200 error {code}: {message}")
        self.code = code
        self.message = message


class _System(C.Structure):
    _fields_ = [
        ("num_particles", C.c_int32), ("padded_num_atoms", C.c_int32), ("num_molecules", C.c_int32),
        ("masses", C.c_void_p), ("particle_mol_id", C.c_void_p),
        ("num_drude", C.c_int32), ("drude_pairs", C.c_void_p),
        ("num_constraints", C.c_int32), ("constraints", C.c_void_p),
        ("has_cm_motion_remover", C.c_int32),
        ("num_langevin", C.c_int32), ("particles_langevin", C.c_void_p),
        ("num_image_pairs", C.c_int32), ("image_pairs", C.c_void_p),
        ("num_electrolyte", C.c_int32), ("particles_electrolyte", C.c_void_p),
    ]


class _Params(C.Structure):
    _fields_ = [
        ("temperature", C.c_double), ("frequency", C.c_double), ("drude_temperature", C.c_double),
        ("drude_frequency", C.c_double), ("step_size", C.c_double),
        ("num_nh_chains", C.c_int32), ("loops_per_step", C.c_int32),
        ("use_com_temp_group", C.c_int32), ("use_middle_scheme", C.c_int32),
        ("max_drude_distance", C.c_double), ("friction", C.c_double), ("drude_friction", C.c_double),
        ("mirror_location", C.c_double), ("electric_field", C.c_double), ("cos_acceleration", C.c_double),
    ]


class _Buffers(C.Structure):
    _fields_ = [("posq", C.c_void_p), ("posq_correction", C.c_void_p), ("velm", C.c_void_p),
                ("force", C.c_void_p), ("pos_delta", C.c_void_p), ("random", C.c_void_p)]


class _StepArgs(C.Structure):
    _fields_ = [("random_index", C.c_uint32), ("inv_box_z", C.c_double)]


class _ThermostatState(C.Structure):
    _fields_ = [("num_temp_groups", C.c_int32), ("ke2", C.c_double * 3), ("vscale", C.c_double * 3),
                ("velocity_bias", C.c_double), ("eta", C.c_double * (3 * MAX_CHAINS)),
                ("eta_dot", C.c_double * (3 * (MAX_CHAINS + 1))), ("eta_dotdot", C.c_double * (3 * MAX_CHAINS))]


class _Temperatures(C.Structure):
    _fields_ = [("num_temp_groups", C.c_int32), ("ke2", C.c_double * 3), ("dof", C.c_double * 3),
                ("temperature", C.c_double * 3), ("velocity_bias", C.c_double)]


@dataclass
class Params:
    """VVIntegrator's parameters (VVIntegrator.h:62-431); defaults are the constructor's
    (VVIntegrator.cpp:49-69) with the run-bulk.py thermostat settings."""
    temperature: float = 333.0
    frequency: float = 10.0
    drude_temperature: float = 1.0
    drude_frequency: float = 40.0
    step_size: float = 0.001
    num_nh_chains: int = 3
    loops_per_step: int = 1
    use_com_temp_group: bool = True
    use_middle_scheme: bool = True
    max_drude_distance: float = 0.0
    friction: float = 5.0
    drude_friction: float = 20.0
    mirror_location: float = 0.0
    electric_field: float = 0.0
    cos_acceleration: float = 0.0

    def resolved_for(self, spec, auto_com=True, auto_friction=True):
        """Apply VVIntegrator::initialize's auto-defaults (VVIntegrator.cpp:106-121)."""
        import dataclasses
        p = dataclasses.replace(self)
        has_drude = spec.drude_pairs.shape[0] > 0
        if auto_com:
            p.use_com_temp_group = has_drude
        if auto_friction:
            p.friction = 5.0 if has_drude else 1.0
        return p

    def to_c(self):
        return _Params(self.temperature, self.frequency, self.drude_temperature, self.drude_frequency,
                       self.step_size, self.num_nh_chains, self.loops_per_step,
                       int(self.use_com_temp_group), int(self.use_middle_scheme), self.max_drude_distance,
                       self.friction, self.drude_friction, self.mirror_location, self.electric_field,
                       self.cos_acceleration)


_lib = None


def load_library(path=None):
    """dlopen libvvb200.so and declare every prototype of include/vvb200.h. Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or os.environ.get("VVB200_LIB") or LIB_PATH      # VVB200_LIB: tuning builds of the same ABI
    if not os.path.exists(path):
        raise VVB200Error(6, f"{path} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             f"(there is no CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.vvb200_last_error.restype = C.c_char_p
    lib.vvb200_last_error.argtypes = []
    lib.vvb200_version.restype = C.c_int
    lib.vvb200_has_device_code.restype = C.c_int
    lib.vvb200_find_molecules.argtypes = [i32, i32, vp, vp, P(i32)]
    lib.vvb200_plan_create.argtypes = [P(_System), P(_Params), C.c_int, P(vp)]
    lib.vvb200_plan_destroy.argtypes = [vp]
    lib.vvb200_plan_destroy.restype = None
    lib.vvb200_plan_get_int_array.argtypes = [vp, C.c_int, P(vp), P(i64)]
    lib.vvb200_plan_get_f64_array.argtypes = [vp, C.c_int, P(vp), P(i64)]
    lib.vvb200_plan_num_temp_groups.argtypes = [vp]
    lib.vvb200_plan_uses_tiled_path.argtypes = [vp]
    lib.vvb200_plan_random_request.argtypes = [vp]
    lib.vvb200_plan_random_request.restype = C.c_uint32
    lib.vvb200_set_step_size.argtypes = [vp, dbl]
    lib.vvb200_propagate_nh_chain.argtypes = [dbl, C.c_int, C.c_int, vp, vp, vp, vp, dbl, dbl, dbl, P(dbl)]
    lib.vvb200_plan_upload.argtypes = [vp, vp]
    for name in ("vvb200_step_middle", "vvb200_step_vv_first", "vvb200_step_vv_second",
                 "vvb200_middle_kick_reduce", "vvb200_middle_nhc_scale_drift", "vvb200_middle_kick",
                 "vvb200_thermostat", "vvb200_middle_thermostat_delta"):
        getattr(lib, name).argtypes = [vp, P(_Buffers), P(_StepArgs), vp]
    lib.vvb200_middle_delta.argtypes = [vp, P(_Buffers), C.c_int, vp]
    lib.vvb200_middle_finish.argtypes = [vp, P(_Buffers), vp]
    lib.vvb200_update_image_positions.argtypes = [vp, P(_Buffers), vp]
    lib.vvb200_vv_kick.argtypes = [vp, P(_Buffers), P(_StepArgs), C.c_int, C.c_int, vp]
    lib.vvb200_vv_positions.argtypes = [vp, P(_Buffers), vp]
    lib.vvb200_partials_ptr.argtypes = [vp, P(vp), P(i32)]
    lib.vvb200_peer_export.argtypes = [vp, vp]
    lib.vvb200_peer_attach.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.vvb200_set_global_thermostat.argtypes = [vp, vp, dbl]
    lib.vvb200_get_thermostat_state.argtypes = [vp, P(_ThermostatState), vp]
    lib.vvb200_set_thermostat_state.argtypes = [vp, P(_ThermostatState), vp]
    lib.vvb200_calc_viscosity.argtypes = [vp, dbl, dbl, dbl, P(dbl), P(dbl), vp]
    lib.vvb200_get_com_velocities.argtypes = [vp, vp, vp]
    lib.vvb200_launch_count.argtypes = [vp]
    lib.vvb200_launch_count.restype = i64
    lib.vvb200_step_host.argtypes = [vp, P(_Buffers), P(_StepArgs), C.c_int, vp]
    lib.vvb200_step_host_begin.argtypes = [vp, P(_Buffers), P(_StepArgs), vp]
    lib.vvb200_step_host_finish.argtypes = [vp, P(_Buffers), P(_StepArgs), vp]
    lib.vvb200_checkpoint_size.argtypes = [vp, P(i64)]
    lib.vvb200_checkpoint_save.argtypes = [vp, vp, i64, vp]
    lib.vvb200_checkpoint_load.argtypes = [vp, vp, i64, vp]
    lib.vvb200_measure_temperatures.argtypes = [vp, P(_Buffers), P(_StepArgs), P(_Temperatures), vp]
    lib.vvb200_set_resident_mode.argtypes = [vp, C.c_int]
    lib.vvb200_resident_launch_count.argtypes = [vp]
    lib.vvb200_resident_launch_count.restype = i64
    lib.vvb200_profile_enable.argtypes = [vp, C.c_int]
    lib.vvb200_profile_read.argtypes = [vp, P(dbl), P(dbl), P(i32)]
    if path in (LIB_PATH, os.environ.get("VVB200_LIB")):
        _lib = lib
    return lib


def _check(lib, rc):
    if rc != 0:
        raise VVB200Error(rc, lib.vvb200_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def find_molecules(num_particles, bonds):
    lib = load_library()
    bonds = np.ascontiguousarray(bonds, dtype=np.int32).reshape(-1, 2)
    mol = np.empty(num_particles, dtype=np.int32)
    n = C.c_int32(0)
    _check(lib, lib.vvb200_find_molecules(num_particles, bonds.shape[0], _ptr(bonds), _ptr(mol), C.byref(n)))
    return mol, n.value


def propagate_nh_chain(step_size, loops, eta, eta_dot, eta_dotdot, eta_mass, ke2, ke2_target, t_target):
    """VVIntegrator::propagateNHChain through the C ABI; arrays are updated in place."""
    lib = load_library()
    f = C.c_double(0)
    _check(lib, lib.vvb200_propagate_nh_chain(step_size, loops, len(eta), _ptr(eta), _ptr(eta_dot),
                                              _ptr(eta_dotdot), _ptr(eta_mass), ke2, ke2_target, t_target,
                                              C.byref(f)))
    return f.value


class Plan:
    """Owns a vvb200_plan*. Host-side builders run in the constructor (no GPU needed);
    upload() / step_*() need a CUDA device."""

    def __init__(self, spec, params, precision="mixed"):
        self.lib = load_library()
        self.spec = spec
        self.params = params
        self.precision = precision
        self._keep = spec.c_arrays()
        k = self._keep
        sysc = _System(spec.n, spec.padded_n, spec.n_mol, _ptr(k["masses"]), _ptr(k["mol_id"]),
                       spec.drude_pairs.shape[0], _ptr(k["drude_pairs"]),
                       spec.constraints.shape[0], _ptr(k["constraints"]), int(spec.has_cmm),
                       spec.langevin.shape[0], _ptr(k["langevin"]),
                       spec.image_pairs.shape[0], _ptr(k["image_pairs"]),
                       spec.electrolyte.shape[0], _ptr(k["electrolyte"]))
        parc = params.to_c()
        h = C.c_void_p()
        _check(self.lib, self.lib.vvb200_plan_create(C.byref(sysc), C.byref(parc), PRECISIONS[precision], C.byref(h)))
        self.h = h
        self.uploaded = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.vvb200_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-side views ----
    def int_array(self, name):
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib, self.lib.vvb200_plan_get_int_array(self.h, INT_ARRAYS[name], C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.int32)
        dtype = np.uint32 if name == "slotMeta" else np.int32
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(n.value,)).view(dtype).copy()

    def f64_array(self, name):
        p, n = C.c_void_p(), C.c_int64()
        _check(self.lib, self.lib.vvb200_plan_get_f64_array(self.h, F64_ARRAYS[name], C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_double)), shape=(n.value,)).copy()

    @property
    def num_temp_groups(self):
        return self.lib.vvb200_plan_num_temp_groups(self.h)

    @property
    def tiled(self):
        return bool(self.lib.vvb200_plan_uses_tiled_path(self.h))

    @property
    def random_request(self):
        return int(self.lib.vvb200_plan_random_request(self.h))

    @property
    def launch_count(self):
        return int(self.lib.vvb200_launch_count(self.h))

    @property
    def resident_launch_count(self):
        """steps that ran as ONE launch with the state resident in shared memory (small systems)"""
        return int(self.lib.vvb200_resident_launch_count(self.h))

    def set_resident_mode(self, mode):
        """-1 default (on unless VVB200_RESIDENT=0), 0 streaming passes only, 1 on"""
        _check(self.lib, self.lib.vvb200_set_resident_mode(self.h, int(mode)))

    def set_step_size(self, dt):
        _check(self.lib, self.lib.vvb200_set_step_size(self.h, dt))

    # ---- device side ----
    @staticmethod
    def _stream(stream):
        if stream is None:
            import torch
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)
        return C.c_void_p(int(stream))

    def upload(self, stream=None):
        _check(self.lib, self.lib.vvb200_plan_upload(self.h, self._stream(stream)))
        self.uploaded = True
        return self

    def _call(self, fn, bufs, random_index=0, inv_box_z=0.0, stream=None):
        b = bufs.c_struct()
        a = _StepArgs(random_index, inv_box_z)
        _check(self.lib, fn(self.h, C.byref(b), C.byref(a), self._stream(stream)))

    def step_middle(self, bufs, **kw):
        self._call(self.lib.vvb200_step_middle, bufs, **kw)

    def step_vv_first(self, bufs, **kw):
        self._call(self.lib.vvb200_step_vv_first, bufs, **kw)

    def step_vv_second(self, bufs, **kw):
        self._call(self.lib.vvb200_step_vv_second, bufs, **kw)

    def middle_kick_reduce(self, bufs, **kw):
        self._call(self.lib.vvb200_middle_kick_reduce, bufs, **kw)

    def middle_nhc_scale_drift(self, bufs, **kw):
        self._call(self.lib.vvb200_middle_nhc_scale_drift, bufs, **kw)

    def middle_kick(self, bufs, **kw):
        self._call(self.lib.vvb200_middle_kick, bufs, **kw)

    def middle_thermostat_delta(self, bufs, **kw):
        self._call(self.lib.vvb200_middle_thermostat_delta, bufs, **kw)

    def thermostat(self, bufs, **kw):
        self._call(self.lib.vvb200_thermostat, bufs, **kw)

    def middle_delta(self, bufs, accumulate, stream=None):
        b = bufs.c_struct()
        _check(self.lib, self.lib.vvb200_middle_delta(self.h, C.byref(b), int(accumulate), self._stream(stream)))

    def middle_finish(self, bufs, stream=None):
        b = bufs.c_struct()
        _check(self.lib, self.lib.vvb200_middle_finish(self.h, C.byref(b), self._stream(stream)))

    def vv_kick(self, bufs, second_half, update_pos_delta, random_index=0, inv_box_z=0.0, stream=None):
        b = bufs.c_struct()
        a = _StepArgs(random_index, inv_box_z)
        _check(self.lib, self.lib.vvb200_vv_kick(self.h, C.byref(b), C.byref(a), int(second_half), int(update_pos_delta),
                                                 self._stream(stream)))

    def vv_positions(self, bufs, stream=None):
        b = bufs.c_struct()
        _check(self.lib, self.lib.vvb200_vv_positions(self.h, C.byref(b), self._stream(stream)))

    def update_image_positions(self, bufs, stream=None):
        b = bufs.c_struct()
        _check(self.lib, self.lib.vvb200_update_image_positions(self.h, C.byref(b), self._stream(stream)))

    def step_constrained(self, bufs, solver, steps=1, random_index=0, inv_box_z=0.0, stream=None):
        """`steps` steps of the constraint-bearing flow, exactly as the OpenMM glue issues it (csrc/glue/
        CudaVVKernelsB200.cpp): `solver` plays CudaIntegrationUtilities -- solver.apply_velocity_constraints(bufs) and
        solver.apply_constraints(bufs) run between the split entry points (CudaVVKernels.cpp:151,176,351,427).
        bufs needs pos_delta.  Middle scheme: kick | vel. constraints | thermostat + both half drifts into posDelta |
        pos. constraints | finish (+ hard wall + images).  Velocity-Verlet: thermostat | half kick + posDelta |
        pos. constraints | positions (+ hard wall + images) | half kick | vel. constraints | thermostat."""
        req = self.random_request
        kw = dict(random_index=random_index, inv_box_z=inv_box_z, stream=stream)
        for _ in range(steps):
            kw["random_index"] = random_index
            if self.params.use_middle_scheme:
                self.middle_kick(bufs, **kw)
                solver.apply_velocity_constraints(bufs)
                self.middle_thermostat_delta(bufs, **kw)
                solver.apply_constraints(bufs)
                self.middle_finish(bufs, stream=stream)
            else:
                self.thermostat(bufs, **kw)
                self.vv_kick(bufs, False, True, **kw)
                solver.apply_constraints(bufs)
                self.vv_positions(bufs, stream=stream)
                self.vv_kick(bufs, True, False, **kw)
                solver.apply_velocity_constraints(bufs)
                self.thermostat(bufs, **kw)
            random_index += req
        return random_index

    def step(self, bufs, steps=1, random_index=0, inv_box_z=0.0, stream=None):
        """`steps` whole integrator steps with the forces in bufs held fixed; returns the
        random index after the last step (advanced like prepareRandomNumbers)."""
        req = self.random_request
        for _ in range(steps):
            if self.params.use_middle_scheme:
                self.step_middle(bufs, random_index=random_index, inv_box_z=inv_box_z, stream=stream)
                random_index += req
            else:
                self.step_vv_first(bufs, random_index=random_index, inv_box_z=inv_box_z, stream=stream)
                self.step_vv_second(bufs, random_index=random_index, inv_box_z=inv_box_z, stream=stream)
                random_index += req
        return random_index

    def partials(self):
        """(device pointer, count) of the fp64 reduction vector for the multi-GPU all-reduce."""
        p, n = C.c_void_p(), C.c_int32()
        _check(self.lib, self.lib.vvb200_partials_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def peer_export(self):
        """64-byte cudaIpcMemHandle_t of this rank's exchange buffer"""
        h = np.zeros(64, dtype=np.uint8)
        _check(self.lib, self.lib.vvb200_peer_export(self.h, _ptr(h)))
        return h

    def peer_attach(self, rank, world, handles):
        handles = np.ascontiguousarray(handles, dtype=np.uint8).reshape(world, 64)
        _check(self.lib, self.lib.vvb200_peer_attach(self.h, int(rank), int(world), _ptr(handles)))

    def set_global_thermostat(self, dof3, total_mass):
        dof3 = np.ascontiguousarray(dof3, dtype=np.float64)
        _check(self.lib, self.lib.vvb200_set_global_thermostat(self.h, _ptr(dof3), float(total_mass)))

    def thermostat_state(self, stream=None):
        s = _ThermostatState()
        _check(self.lib, self.lib.vvb200_get_thermostat_state(self.h, C.byref(s), self._stream(stream)))
        nc, ng = self.params.num_nh_chains, s.num_temp_groups
        return {
            "num_temp_groups": ng,
            "ke2": np.array(s.ke2[:]), "vscale": np.array(s.vscale[:]), "velocity_bias": s.velocity_bias,
            "eta": np.array(s.eta[: ng * nc]).reshape(ng, nc),
            "eta_dot": np.array(s.eta_dot[: ng * (nc + 1)]).reshape(ng, nc + 1),
            "eta_dotdot": np.array(s.eta_dotdot[: ng * nc]).reshape(ng, nc),
        }

    def set_thermostat_state(self, eta, eta_dot, eta_dotdot, stream=None):
        s = _ThermostatState()
        for dst, src in ((s.eta, eta), (s.eta_dot, eta_dot), (s.eta_dotdot, eta_dotdot)):
            flat = np.asarray(src, dtype=np.float64).ravel()
            for i, v in enumerate(flat):
                dst[i] = v
        _check(self.lib, self.lib.vvb200_set_thermostat_state(self.h, C.byref(s), self._stream(stream)))

    def checkpoint_save(self, stream=None):
        """the integrator's part of a checkpoint (NH chain state, last scale factors, VV extra-force flag) as bytes"""
        n = C.c_int64()
        _check(self.lib, self.lib.vvb200_checkpoint_size(self.h, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        _check(self.lib, self.lib.vvb200_checkpoint_save(self.h, _ptr(buf), n.value, self._stream(stream)))
        return buf.tobytes()

    def checkpoint_load(self, blob, stream=None):
        buf = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
        _check(self.lib, self.lib.vvb200_checkpoint_load(self.h, _ptr(buf), buf.size, self._stream(stream)))

    def measure_temperatures(self, bufs, inv_box_z=0.0, stream=None):
        """group temperatures of the current velocities, no stepping (the Drude-temperature reporter's numbers)"""
        b = bufs.c_struct()
        a = _StepArgs(0, inv_box_z)
        t = _Temperatures()
        _check(self.lib, self.lib.vvb200_measure_temperatures(self.h, C.byref(b), C.byref(a), C.byref(t), self._stream(stream)))
        ng = t.num_temp_groups
        return {"num_temp_groups": ng, "ke2": np.array(t.ke2[:ng]), "dof": np.array(t.dof[:ng]),
                "temperature": np.array(t.temperature[:ng]), "velocity_bias": t.velocity_bias}

    def viscosity(self, box, stream=None):
        v, iv = C.c_double(), C.c_double()
        _check(self.lib, self.lib.vvb200_calc_viscosity(self.h, box[0], box[1], box[2], C.byref(v), C.byref(iv),
                                                        self._stream(stream)))
        return v.value, iv.value

    def com_velocities(self, stream=None):
        mixed = np.float32 if self.precision == "single" else np.float64
        out = np.zeros((self.spec.n_mol, 4), dtype=mixed)
        _check(self.lib, self.lib.vvb200_get_com_velocities(self.h, _ptr(out), self._stream(stream)))
        return out

    def profile_enable(self, max_steps):
        _check(self.lib, self.lib.vvb200_profile_enable(self.h, int(max_steps)))

    def profile_read(self):
        """(ms in pass A, ms in pass B, steps covered) since the last read"""
        a, b, n = C.c_double(), C.c_double(), C.c_int32()
        _check(self.lib, self.lib.vvb200_profile_read(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def _host_buffers(self, host_state):
        return _Buffers(_ptr(host_state.posq), _ptr(host_state.corr) if host_state.corr is not None else None,
                        _ptr(host_state.velm), _ptr(host_state.force), None, None)

    def step_host_begin(self, host_state, inv_box_z=0.0, stream=None):
        """first half of a pipelined host-buffer step: all copies in are queued, pass A runs; does not synchronise"""
        b, a = self._host_buffers(host_state), _StepArgs(0, inv_box_z)
        _check(self.lib, self.lib.vvb200_step_host_begin(self.h, C.byref(b), C.byref(a), self._stream(stream)))

    def step_host_finish(self, host_state, inv_box_z=0.0, stream=None):
        """second half: NH chains (+ peer exchange), pass B, copies out; synchronises"""
        b, a = self._host_buffers(host_state), _StepArgs(0, inv_box_z)
        _check(self.lib, self.lib.vvb200_step_host_finish(self.h, C.byref(b), C.byref(a), self._stream(stream)))

    def step_host(self, host_state, steps=1, inv_box_z=0.0, stream=None):
        """vvb200_step_host: host arrays in, host arrays out (H2D + steps + D2H inside)."""
        b = _Buffers(_ptr(host_state.posq), _ptr(host_state.corr) if host_state.corr is not None else None,
                     _ptr(host_state.velm), _ptr(host_state.force), None, None)
        a = _StepArgs(0, inv_box_z)
        _check(self.lib, self.lib.vvb200_step_host(self.h, C.byref(b), C.byref(a), steps, self._stream(stream)))
