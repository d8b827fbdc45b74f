"""TEST INFRASTRUCTURE ONLY -- ctypes wrappers for the CPU oracle (oracle/vv_oracle.c) and for
the reference's own kernels compiled under oracle/_ref (oracle/ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (openmm-velocityverlet_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODES = ("single", "mixed", "double")

ARR = {"particlesNH": 0, "moleculesNH": 1, "particleMolId": 2, "drudePairs": 3, "sortedByMol": 4,
       "particlesInMolecules": 5, "normalNH": 6, "pairsNH": 7, "normalLD": 8, "pairsLD": 9,
       "imagePairs": 10, "electrolyte": 11,
       "moleculeMasses": 100, "moleculeInvMasses": 101, "dof": 102, "etaMass": 103, "NkbT": 104,
       "invMassTotal": 105, "eta": 106, "etaDot": 107, "etaDotDot": 108, "ke2": 109, "vscale": 110,
       "vBias": 111}


def build(target="all"):
    """make -C oracle <target>; `ref` is a no-op when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)


class _System(C.Structure):
    _fields_ = [("numParticles", C.c_int32), ("paddedNumAtoms", C.c_int32), ("numMolecules", C.c_int32),
                ("masses", C.c_void_p), ("particleMolId", C.c_void_p),
                ("numDrude", C.c_int32), ("drudePairs", C.c_void_p),
                ("numConstraints", C.c_int32), ("constraints", C.c_void_p),
                ("hasCMMotionRemover", C.c_int32),
                ("numLD", C.c_int32), ("particlesLD", C.c_void_p),
                ("numImagePairs", C.c_int32), ("imagePairs", C.c_void_p),
                ("numElectrolyte", C.c_int32), ("particlesElectrolyte", C.c_void_p)]


class _Params(C.Structure):
    _fields_ = [("temperature", C.c_double), ("frequency", C.c_double), ("drudeTemperature", C.c_double),
                ("drudeFrequency", C.c_double), ("stepSize", C.c_double),
                ("numNHChains", C.c_int32), ("loopsPerStep", C.c_int32),
                ("useCOMTempGroup", C.c_int32), ("useMiddleScheme", C.c_int32),
                ("maxDrudeDistance", C.c_double), ("friction", C.c_double), ("drudeFriction", C.c_double),
                ("mirrorLocation", C.c_double), ("electricField", C.c_double), ("cosAcceleration", C.c_double)]


class _Buffers(C.Structure):
    _fields_ = [("posq", C.c_void_p), ("posqCorrection", C.c_void_p), ("velm", C.c_void_p),
                ("force", C.c_void_p), ("posDelta", C.c_void_p), ("random", C.c_void_p)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None


def _params_c(p):
    return _Params(p.temperature, p.frequency, p.drude_temperature, p.drude_frequency, p.step_size,
                   p.num_nh_chains, p.loops_per_step, int(p.use_com_temp_group), int(p.use_middle_scheme),
                   p.max_drude_distance, p.friction, p.drude_friction, p.mirror_location, p.electric_field,
                   p.cos_acceleration)


_oracle_libs = {}


def oracle_lib(mode):
    if mode not in _oracle_libs:
        path = os.path.join(HERE, "build", f"libvvoracle_{mode}.so")
        if not os.path.exists(path):
            build("oracle")
        lib = C.CDLL(path)
        lib.vvo_last_error.restype = C.c_char_p
        lib.vvo_create.restype = C.c_void_p
        lib.vvo_create.argtypes = [C.POINTER(_System), C.POINTER(_Params), C.c_int]
        lib.vvo_destroy.argtypes = [C.c_void_p]
        lib.vvo_num_temp_groups.argtypes = [C.c_void_p]
        lib.vvo_get_array.restype = C.c_int64
        lib.vvo_get_array.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        lib.vvo_set_nhc_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.vvo_find_molecules.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.vvo_propagate_nh_chain.argtypes = [C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double)]
        lib.vvo_step.argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_int, C.c_double, C.POINTER(C.c_uint)]
        for name in ("vvo_reset_extra_force",):
            getattr(lib, name).argtypes = [C.c_void_p]
        for name in ("vvo_electric_force", "vvo_middle_vel", "vvo_middle_pos1", "vvo_middle_pos2", "vvo_middle_pos3",
                     "vvo_hard_wall", "vvo_vv_positions", "vvo_scale_velocity", "vvo_update_images"):
            getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(_Buffers)]
        lib.vvo_langevin_force.argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_uint]
        lib.vvo_vv_velocities.argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_int]
        for name in ("vvo_cosine_force", "vvo_calc_velocity_bias", "vvo_remove_velocity_bias", "vvo_restore_velocity_bias"):
            getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_double]
        lib.vvo_calc_viscosity.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.vvo_toy_forces.argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_void_p, C.c_double, C.c_double]
        lib.vvo_set_num_threads.argtypes = [C.c_int]
        lib.vvo_set_constraint_standin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for name in ("vvo_apply_constraints", "vvo_apply_velocity_constraints"):
            getattr(lib, name).argtypes = [C.c_void_p, C.POINTER(_Buffers)]
        _oracle_libs[mode] = lib
    return _oracle_libs[mode]


def find_molecules(n, bonds):
    lib = oracle_lib("mixed")
    bonds = np.ascontiguousarray(bonds, dtype=np.int32).reshape(-1, 2)
    out = np.empty(n, dtype=np.int32)
    count = lib.vvo_find_molecules(n, bonds.shape[0], _ptr(bonds), _ptr(out))
    return out, count


def propagate_nh_chain(step_size, loops, eta, eta_dot, eta_dotdot, eta_mass, ke2, ke2_target, t_target):
    lib = oracle_lib("mixed")
    f = C.c_double()
    lib.vvo_propagate_nh_chain(step_size, loops, len(eta), _ptr(eta), _ptr(eta_dot), _ptr(eta_dotdot), _ptr(eta_mass),
                               ke2, ke2_target, t_target, C.byref(f))
    return f.value


class OracleError(RuntimeError):
    pass


# ---------------------------------------------------------------------------------------------
# stand-in for OpenMM's constraint solvers (oracle/constraint_standin.h)
# ---------------------------------------------------------------------------------------------
class ConstraintStandin:
    """Cluster tables for the constraint stand-in: connected components of the System's constraint graph (one thread
    each, like OpenMM's SHAKE clusters), constraints inside a cluster in System order, target distances = the
    separations in `state` (the geometry the run starts from), `iterations` sweeps.  The same tables go to the C
    oracle, the reference-kernel harness, the mini-OpenMM and the device helper, so all of them apply one operator."""

    def __init__(self, spec, state, iterations=3):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        cons = np.ascontiguousarray(spec.constraints, dtype=np.int32).reshape(-1, 2)
        self.iterations = int(iterations)
        if cons.shape[0] == 0:
            self.n_clusters = 0
            self.offset = np.zeros(1, np.int32)
            self.atoms = np.zeros((0, 2), np.int32)
            self.distance = np.zeros(0, np.float64)
            return
        n = spec.n
        g = coo_matrix((np.ones(len(cons), np.int8), (cons[:, 0], cons[:, 1])), shape=(n, n))
        _, lab = connected_components(g, directed=False)
        cl = lab[cons[:, 0]]
        # clusters numbered by their first constraint; constraints keep System order inside a cluster
        _, first = np.unique(cl, return_index=True)
        rank = np.empty(lab.max() + 1, dtype=np.int64)
        rank[cl[np.sort(first)]] = np.arange(len(first))
        order = np.argsort(rank[cl], kind="stable")
        self.atoms = np.ascontiguousarray(cons[order])
        counts = np.bincount(rank[cl], minlength=len(first))
        self.offset = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.n_clusters = int(len(first))
        x = state.positions()
        self.distance = np.ascontiguousarray(np.linalg.norm(x[self.atoms[:, 0]] - x[self.atoms[:, 1]], axis=1), dtype=np.float64)

    def c_args(self):
        return (self.n_clusters, _ptr(self.offset), _ptr(self.atoms), _ptr(self.distance), self.iterations)


_standin_lib = None


def standin_lib():
    global _standin_lib
    if _standin_lib is None:
        path = os.path.join(HERE, "build", "libvvstandin_cuda.so")
        if not os.path.exists(path):
            build("oracle")
        lib = C.CDLL(path)
        lib.vvstandin_create.restype = C.c_void_p
        lib.vvstandin_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.vvstandin_destroy.argtypes = [C.c_void_p]
        lib.vvstandin_apply_constraints.argtypes = [C.c_void_p] * 6
        lib.vvstandin_apply_velocity_constraints.argtypes = [C.c_void_p] * 5
        _standin_lib = lib
    return _standin_lib


class DeviceStandin:
    """The stand-in on device buffers (oracle/standin_cuda.cu): what a GPU test calls between the product's split entry
    points, where the OpenMM glue calls integration.applyVelocityConstraints / applyConstraints."""

    def __init__(self, standin, precision):
        self.lib = standin_lib()
        self._keep = standin
        self.h = self.lib.vvstandin_create(MODES.index(precision), *standin.c_args())

    def __del__(self):
        try:
            if self.h:
                self.lib.vvstandin_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def apply_constraints(self, bufs, stream=None):
        import torch
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        rc = self.lib.vvstandin_apply_constraints(self.h, self._p(bufs.posq), self._p(bufs.corr), self._p(bufs.velm),
                                                  self._p(bufs.pos_delta), st)
        assert rc == 0, f"stand-in launch failed: cuda error {rc}"

    def apply_velocity_constraints(self, bufs, stream=None):
        import torch
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        rc = self.lib.vvstandin_apply_velocity_constraints(self.h, self._p(bufs.posq), self._p(bufs.corr), self._p(bufs.velm), st)
        assert rc == 0, f"stand-in launch failed: cuda error {rc}"


def _bufs(state, pos_delta=None):
    return _Buffers(_ptr(state.posq), _ptr(state.corr) if state.corr is not None else None, _ptr(state.velm),
                    _ptr(state.force), _ptr(pos_delta) if pos_delta is not None else None, _ptr(state.random))


class Oracle:
    """CPU restatement of the reference path for one system (oracle/vv_oracle.c)."""

    def __init__(self, spec, params, precision="mixed", literal=True, threads=1):
        self.lib = oracle_lib(precision)
        self.lib.vvo_set_num_threads(threads)
        self.spec, self.params, self.precision = spec, params, precision
        self._keep = spec.c_arrays()
        k = self._keep
        s = _System(spec.n, spec.padded_n, spec.n_mol, _ptr(k["masses"]), _ptr(k["mol_id"]),
                    spec.drude_pairs.shape[0], _ptr(k["drude_pairs"]), spec.constraints.shape[0], _ptr(k["constraints"]),
                    int(spec.has_cmm), spec.langevin.shape[0], _ptr(k["langevin"]),
                    spec.image_pairs.shape[0], _ptr(k["image_pairs"]), spec.electrolyte.shape[0], _ptr(k["electrolyte"]))
        p = _params_c(params)
        self.h = self.lib.vvo_create(C.byref(s), C.byref(p), int(literal))
        if not self.h:
            raise OracleError(self.lib.vvo_last_error().decode())
        self.random_index = C.c_uint(0)
        self._pos_delta = None

    def __del__(self):
        try:
            if self.h:
                self.lib.vvo_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def array(self, name):
        p = C.c_void_p()
        n = self.lib.vvo_get_array(self.h, ARR[name], C.byref(p))
        if n < 0:
            raise KeyError(name)
        if n == 0:
            return np.zeros(0, dtype=np.int32 if ARR[name] < 100 else np.float64)
        ct = C.c_int32 if ARR[name] < 100 else C.c_double
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n,)).copy()

    @property
    def num_temp_groups(self):
        return self.lib.vvo_num_temp_groups(self.h)

    def _b(self, state):
        if self._pos_delta is None or self._pos_delta.shape != state.velm.shape or self._pos_delta.dtype != state.velm.dtype:
            self._pos_delta = np.zeros_like(state.velm)
        return _bufs(state, self._pos_delta)

    def step(self, state, steps=1, inv_box_z=0.0):
        """advance `state` (HostState, in place) by whole integrator steps with frozen forces"""
        b = self._b(state)
        self.lib.vvo_step(self.h, C.byref(b), steps, inv_box_z, C.byref(self.random_index))

    def scale_velocity(self, state):
        b = self._b(state)
        self.lib.vvo_scale_velocity(self.h, C.byref(b))

    def set_constraints(self, standin):
        """switch the constraint stand-in on (ConstraintStandin) or off (None) for vvo_step"""
        if standin is None:
            self.lib.vvo_set_constraint_standin(self.h, 0, None, None, None, 0)
        else:
            self.lib.vvo_set_constraint_standin(self.h, *standin.c_args())
        return self

    def call(self, kernel, state, *extra):
        b = self._b(state)
        getattr(self.lib, "vvo_" + kernel)(self.h, C.byref(b), *extra)

    def toy_forces(self, state, x0, k_tether, k_drude):
        b = self._b(state)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.lib.vvo_toy_forces(self.h, C.byref(b), _ptr(x0), k_tether, k_drude)

    def thermostat_state(self):
        ng, nc = self.num_temp_groups, self.params.num_nh_chains
        return {"num_temp_groups": ng, "ke2": self.array("ke2"), "vscale": self.array("vscale"),
                "velocity_bias": float(self.array("vBias")[0]),
                "eta": self.array("eta").reshape(ng, nc), "eta_dot": self.array("etaDot").reshape(ng, nc + 1),
                "eta_dotdot": self.array("etaDotDot").reshape(ng, nc)}

    def viscosity(self, box):
        v, iv = C.c_double(), C.c_double()
        self.lib.vvo_calc_viscosity(self.h, box[0], box[1], box[2], C.byref(v), C.byref(iv))
        return v.value, iv.value


# ---------------------------------------------------------------------------------------------
# the reference's own kernels (oracle/_ref)
# ---------------------------------------------------------------------------------------------
class _RefIndices(C.Structure):
    _fields_ = [("numAtoms", C.c_int32), ("paddedNumAtoms", C.c_int32), ("numMolecules", C.c_int32),
                ("numDrude", C.c_int32), ("drudePairs", C.c_void_p),
                ("numParticlesNH", C.c_int32), ("particlesNH", C.c_void_p),
                ("numMoleculesNH", C.c_int32), ("moleculesNH", C.c_void_p),
                ("numNormalNH", C.c_int32), ("normalParticlesNH", C.c_void_p),
                ("numPairsNH", C.c_int32), ("pairParticlesNH", C.c_void_p),
                ("particleMolId", C.c_void_p), ("particlesInMolecules", C.c_void_p),
                ("particlesSortedByMolId", C.c_void_p),
                ("numTempGroup", C.c_int32), ("etaMass", C.c_void_p), ("tempGroupNkbT", C.c_void_p),
                ("numParticlesLD", C.c_int32),
                ("numNormalLD", C.c_int32), ("normalParticlesLD", C.c_void_p),
                ("numPairsLD", C.c_int32), ("pairParticlesLD", C.c_void_p),
                ("numImagePairs", C.c_int32), ("imagePairs", C.c_void_p),
                ("numElectrolyte", C.c_int32), ("particlesElectrolyte", C.c_void_p),
                ("invMassTotal", C.c_double)]


def ref_lib_path(mode, gpu=False):
    return os.path.join(HERE, "_ref", f"libvvref_{'cuda' if gpu else 'cpu'}_{mode}.so")


def ref_available(mode="mixed", gpu=False):
    return os.path.exists(ref_lib_path(mode, gpu))


_ref_libs = {}


def ref_lib(mode, gpu=False):
    key = (mode, gpu)
    if key not in _ref_libs:
        lib = C.CDLL(ref_lib_path(mode, gpu))
        lib.vvref_create.restype = C.c_void_p
        lib.vvref_create.argtypes = [C.POINTER(_RefIndices), C.POINTER(_Params), C.c_int, C.c_void_p]
        lib.vvref_destroy.argtypes = [C.c_void_p]
        lib.vvref_step.restype = C.c_long
        lib.vvref_step.argtypes = [C.c_void_p, C.POINTER(_Buffers), C.c_int, C.c_double, C.POINTER(C.c_uint)]
        lib.vvref_scale_velocity.argtypes = [C.c_void_p, C.POINTER(_Buffers)]
        lib.vvref_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.POINTER(C.c_double)]
        lib.vvref_get_com_velm.argtypes = [C.c_void_p, C.c_void_p]
        lib.vvref_set_num_threads.argtypes = [C.c_int]
        lib.vvref_set_constraint_standin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _ref_libs[key] = lib
    return _ref_libs[key]


class Reference:
    """The reference plugin's own kernels + restated host schedule (oracle/ref_harness.cpp).
    `indices` come from an Oracle (or any object with .array(name)) built for the same system.
    gpu=False: host build (SIMT shim). gpu=True: sm_100a build, buffers are device pointers."""

    def __init__(self, oracle, gpu=False, num_thread_blocks=4 * 148, threads=1, stream=0):
        spec, params = oracle.spec, oracle.params
        self.precision = oracle.precision
        self.gpu = gpu
        self.lib = ref_lib(self.precision, gpu)
        self.lib.vvref_set_num_threads(threads)
        self.params = params
        names = ("drudePairs", "particlesNH", "moleculesNH", "normalNH", "pairsNH", "particleMolId",
                 "particlesInMolecules", "sortedByMol", "normalLD", "pairsLD", "imagePairs", "electrolyte",
                 "etaMass", "NkbT")
        a = self._keep = {k: np.ascontiguousarray(oracle.array(k)) for k in names}
        self.num_temp_groups = oracle.num_temp_groups
        ix = _RefIndices(spec.n, spec.padded_n, spec.n_mol,
                         a["drudePairs"].size // 2, _ptr(a["drudePairs"]),
                         a["particlesNH"].size, _ptr(a["particlesNH"]),
                         a["moleculesNH"].size, _ptr(a["moleculesNH"]),
                         a["normalNH"].size, _ptr(a["normalNH"]),
                         a["pairsNH"].size // 2, _ptr(a["pairsNH"]),
                         _ptr(a["particleMolId"]), _ptr(a["particlesInMolecules"]), _ptr(a["sortedByMol"]),
                         self.num_temp_groups, _ptr(a["etaMass"]), _ptr(a["NkbT"]),
                         spec.langevin.size,
                         a["normalLD"].size, _ptr(a["normalLD"]), a["pairsLD"].size // 2, _ptr(a["pairsLD"]),
                         a["imagePairs"].size // 2, _ptr(a["imagePairs"]),
                         a["electrolyte"].size, _ptr(a["electrolyte"]),
                         float(oracle.array("invMassTotal")[0]))
        p = _params_c(params)
        self.h = self.lib.vvref_create(C.byref(ix), C.byref(p), num_thread_blocks, C.c_void_p(stream))
        self.random_index = C.c_uint(0)
        self._pos_delta = None
        self.n_mol = spec.n_mol

    def __del__(self):
        try:
            if self.h:
                self.lib.vvref_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _b_host(self, state):
        if self._pos_delta is None or self._pos_delta.shape != state.velm.shape or self._pos_delta.dtype != state.velm.dtype:
            self._pos_delta = np.zeros_like(state.velm)
        return _bufs(state, self._pos_delta)

    def _b_dev(self, bufs):
        import torch
        if self._pos_delta is None:
            self._pos_delta = torch.zeros_like(bufs.velm)
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        return _Buffers(p(bufs.posq), p(bufs.corr), p(bufs.velm), p(bufs.force), p(self._pos_delta), p(bufs.random))

    def step(self, state_or_bufs, steps=1, inv_box_z=0.0):
        b = self._b_dev(state_or_bufs) if self.gpu else self._b_host(state_or_bufs)
        return self.lib.vvref_step(self.h, C.byref(b), steps, inv_box_z, C.byref(self.random_index))

    def scale_velocity(self, state_or_bufs):
        b = self._b_dev(state_or_bufs) if self.gpu else self._b_host(state_or_bufs)
        self.lib.vvref_scale_velocity(self.h, C.byref(b))

    def set_constraints(self, standin):
        if standin is None:
            self.lib.vvref_set_constraint_standin(self.h, 0, None, None, None, 0)
        else:
            self.lib.vvref_set_constraint_standin(self.h, *standin.c_args())
        return self

    def thermostat_state(self):
        ng, nc = self.num_temp_groups, self.params.num_nh_chains
        ke2, vs = np.zeros(3), np.zeros(3)
        eta, ed, edd = np.zeros(3 * nc), np.zeros(3 * (nc + 1)), np.zeros(3 * nc)
        vb = C.c_double()
        self.lib.vvref_get_state(self.h, _ptr(ke2), _ptr(vs), _ptr(eta), _ptr(ed), _ptr(edd), C.byref(vb))
        return {"num_temp_groups": ng, "ke2": ke2[:ng], "vscale": vs[:ng], "velocity_bias": vb.value,
                "eta": eta[: ng * nc].reshape(ng, nc), "eta_dot": ed[: ng * (nc + 1)].reshape(ng, nc + 1),
                "eta_dotdot": edd[: ng * nc].reshape(ng, nc)}

    def com_velm(self):
        mixed = np.float32 if self.precision == "single" else np.float64
        out = np.zeros((self.n_mol, 4), dtype=mixed)
        self.lib.vvref_get_com_velm(self.h, _ptr(out))
        return out


# ---------------------------------------------------------------------------------------------
# the reference plugin's OWN host code (and the B200 glue) running under the mini-OpenMM
# (oracle/mini_openmm; libraries oracle/_ref/libvvplugin_{ref_cpu,ref_cuda,glue_cuda}_<mode>.so)
# ---------------------------------------------------------------------------------------------
class _MommSystem(C.Structure):
    _fields_ = [("numParticles", C.c_int32), ("masses", C.c_void_p),
                ("numBonds", C.c_int32), ("bonds", C.c_void_p),
                ("numDrude", C.c_int32), ("drudePairs", C.c_void_p),
                ("numConstraints", C.c_int32), ("constraints", C.c_void_p), ("constraintDistances", C.c_void_p),
                ("hasCMMotionRemover", C.c_int32), ("numDrudeForces", C.c_int32),
                ("numLD", C.c_int32), ("particlesLD", C.c_void_p),
                ("numImagePairs", C.c_int32), ("imagePairs", C.c_void_p),
                ("numElectrolyte", C.c_int32), ("particlesElectrolyte", C.c_void_p)]


class _MommParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("frequency", C.c_double), ("drudeTemperature", C.c_double),
                ("drudeFrequency", C.c_double), ("stepSize", C.c_double),
                ("numNHChains", C.c_int32), ("loopsPerStep", C.c_int32),
                ("useCOMTempGroup", C.c_int32), ("useMiddleScheme", C.c_int32),
                ("maxDrudeDistance", C.c_double), ("friction", C.c_double), ("drudeFriction", C.c_double),
                ("mirrorLocation", C.c_double), ("electricField", C.c_double), ("cosAcceleration", C.c_double),
                ("randomNumberSeed", C.c_int32), ("debug", C.c_int32)]


def plugin_lib_path(flavour, mode):
    return os.path.join(HERE, "_ref", f"libvvplugin_{flavour}_{mode}.so")


def plugin_available(flavour="ref_cpu", mode="mixed"):
    return os.path.exists(plugin_lib_path(flavour, mode))


_plugin_libs = {}


def plugin_lib(flavour, mode):
    key = (flavour, mode)
    if key not in _plugin_libs:
        lib = C.CDLL(plugin_lib_path(flavour, mode))      # RTLD_LOCAL: every flavour keeps its own Platform registry
        lib.momm_last_error.restype = C.c_char_p
        lib.momm_create.restype = C.c_void_p
        lib.momm_create.argtypes = [C.POINTER(_MommSystem), C.POINTER(_MommParams), C.c_void_p]
        lib.momm_destroy.argtypes = [C.c_void_p]
        lib.momm_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        lib.momm_get_state.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.momm_array_pointers.argtypes = [C.c_void_p, C.c_void_p]
        lib.momm_set_constraint_standin.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.momm_step.argtypes = [C.c_void_p, C.c_int]
        lib.momm_set_step_size.argtypes = [C.c_void_p, C.c_double]
        lib.momm_get_upload.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        lib.momm_get_int_list.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        lib.momm_get_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        lib.momm_get_viscosity.argtypes = [C.c_void_p, C.c_void_p]
        lib.momm_counters.argtypes = [C.c_void_p, C.c_void_p]
        lib.momm_particles_identical.argtypes = [C.c_void_p, C.c_int, C.c_int]
        _plugin_libs[key] = lib
    return _plugin_libs[key]


class PluginError(RuntimeError):
    """an OpenMMException thrown by the reference's (or the glue's) code"""


class MiniContext:
    """An OpenMM::Context holding the reference's OWN VVIntegrator (openmmapi/src/VVIntegrator.cpp, compiled unchanged)
    under the mini-OpenMM.  flavour: "ref_cpu" / "ref_cuda" = the reference's own CudaVVKernels.cpp + kernel sources
    (host memory + SIMT shim, or device memory + real launches); "glue_cuda" = csrc/glue/CudaVVKernelsB200.cpp over
    libvvb200.so.  step() is VVIntegrator::step(): the reference's stepMiddle / stepVV make the virtual calls.
    auto=True leaves useCOMTempGroup / friction to VVIntegrator::initialize's own auto rules."""

    F64 = {"moleculeMasses": 0, "moleculeInvMasses": 1, "dof": 2, "etaMass": 3, "NkbT": 4, "ke2": 5, "vscale": 6,
           "eta": 7, "etaDot": 8, "etaDotDot": 9, "invMassTotal": 10, "settings": 11}
    INT = {"particlesNH": 0, "moleculesNH": 1, "particleMolId": 2, "particlesLD": 3, "electrolyte": 4}
    # CudaArray names in the reference's initialize() methods (CudaVVKernels.cpp:76, 598-604, 806-807, 889, 955)
    UPLOADS = {"drudePairs": "vvDrudePairs", "particlesNH": "particlesNH", "moleculesNH": "moleculesNH",
               "normalNH": "normalParticlesNH", "pairsNH": "pairParticlesNH", "particleMolId": "particleMolId",
               "particlesInMolecules": "particlesInMolecules", "sortedByMol": "particlesSortedByMolId",
               "normalLD": "normalParticlesLD", "pairsLD": "drudePairParticlesLD", "imagePairs": "imagePairs",
               "electrolyte": "particlesElectrolyte"}

    def __init__(self, spec, params, precision="mixed", flavour="ref_cpu", auto=False, stream=0, num_drude_forces=-1,
                 constraint_distances=None):
        self.lib = plugin_lib(flavour, precision)
        self.spec, self.params, self.precision, self.flavour = spec, params, precision, flavour
        k = self._keep = spec.c_arrays()
        bonds = np.ascontiguousarray(spec.bonds, dtype=np.int32)
        dist = None if constraint_distances is None else np.ascontiguousarray(constraint_distances, dtype=np.float64)
        self._keep2 = (bonds, dist)
        s = _MommSystem(spec.n, _ptr(k["masses"]), bonds.shape[0], _ptr(bonds), spec.drude_pairs.shape[0], _ptr(k["drude_pairs"]),
                        spec.constraints.shape[0], _ptr(k["constraints"]), _ptr(dist), int(spec.has_cmm), num_drude_forces,
                        spec.langevin.shape[0], _ptr(k["langevin"]), spec.image_pairs.shape[0], _ptr(k["image_pairs"]),
                        spec.electrolyte.shape[0], _ptr(k["electrolyte"]))
        p = _MommParams(params.temperature, params.frequency, params.drude_temperature, params.drude_frequency, params.step_size,
                        params.num_nh_chains, params.loops_per_step, -1 if auto else int(params.use_com_temp_group),
                        int(params.use_middle_scheme), params.max_drude_distance, -1.0 if auto else params.friction,
                        -1.0 if auto else params.drude_friction, params.mirror_location, params.electric_field,
                        params.cos_acceleration, 0, 0)
        self.h = self.lib.momm_create(C.byref(s), C.byref(p), C.c_void_p(stream))
        if not self.h:
            raise PluginError(self.lib.momm_last_error().decode())
        self._shape = None

    def __del__(self):
        try:
            if self.h:
                self.lib.momm_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _ok(self, rc):
        if rc != 0:
            raise PluginError(self.lib.momm_last_error().decode())

    def set_state(self, state):
        self._state0 = state
        box = np.array(state.box, dtype=np.float64)
        self._ok(self.lib.momm_set_state(self.h, _ptr(state.posq), _ptr(state.corr) if state.corr is not None else None,
                                         _ptr(state.velm), _ptr(state.force), _ptr(state.random), state.random.shape[0],
                                         _ptr(box)))
        return self

    def get_state(self):
        s = self._state0.copy()
        self._ok(self.lib.momm_get_state(self.h, _ptr(s.posq), _ptr(s.corr) if s.corr is not None else None, _ptr(s.velm)))
        return s

    def set_constraints(self, standin):
        if standin is None:
            self.lib.momm_set_constraint_standin(self.h, 0, None, None, None, 0)
        else:
            self._standin = standin
            self.lib.momm_set_constraint_standin(self.h, *standin.c_args())
        return self

    def step(self, steps=1):
        self._ok(self.lib.momm_step(self.h, steps))

    def set_step_size(self, dt):
        self.lib.momm_set_step_size(self.h, dt)

    def upload(self, name):
        """int32 view of the last upload to the reference CudaArray behind `name` (None if it was never uploaded)"""
        p, n, e = C.c_void_p(), C.c_int64(), C.c_int32()
        if self.lib.momm_get_upload(self.h, self.UPLOADS.get(name, name).encode(), C.byref(p), C.byref(n), C.byref(e)) != 0:
            return None
        return np.frombuffer(C.string_at(p, n.value), dtype=np.int32).copy()

    def int_list(self, name):
        p, n = C.c_void_p(), C.c_int64()
        assert self.lib.momm_get_int_list(self.h, self.INT[name], C.byref(p), C.byref(n)) == 0
        return np.frombuffer(C.string_at(p, 4 * n.value), dtype=np.int32).copy()

    def f64(self, name, cap=1 << 22):
        cap = min(cap, max(64, self.spec.n_mol + 8))
        out = np.zeros(cap, dtype=np.float64)
        n = self.lib.momm_get_f64(self.h, self.F64[name], _ptr(out), cap)
        return None if n < 0 else out[:n].copy()

    def thermostat_state(self):
        ng, nc = int(self.f64("settings")[3]), self.params.num_nh_chains
        return {"num_temp_groups": ng, "ke2": self.f64("ke2"), "vscale": self.f64("vscale"),
                "eta": self.f64("eta").reshape(ng, nc), "eta_dot": self.f64("etaDot").reshape(ng, nc + 1),
                "eta_dotdot": self.f64("etaDotDot").reshape(ng, nc)}

    def viscosity(self):
        out = np.zeros(2)
        self._ok(self.lib.momm_get_viscosity(self.h, _ptr(out)))
        return float(out[0]), float(out[1])

    def counters(self):
        out = np.zeros(8, dtype=np.int64)
        self.lib.momm_counters(self.h, _ptr(out))
        names = ("reference_kernel_launches", "force_evaluations", "constraint_calls", "velocity_constraint_calls",
                 "reorder_calls", "step_count", "force_info_before_init", "vvb200_launches")
        return dict(zip(names, (int(x) for x in out)))

    def particles_identical(self, i, j):
        return bool(self.lib.momm_particles_identical(self.h, int(i), int(j)))

    def array_pointers(self):
        out = (C.c_void_p * 6)()
        self.lib.momm_array_pointers(self.h, out)
        return dict(zip(("posq", "corr", "velm", "force", "pos_delta", "random"), (int(x or 0) for x in out)))
