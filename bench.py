#!/usr/bin/env python
"""bench.py -- integrator-only throughput of the B200-native VVIntegrator path (BASELINE.json metric:
particle-updates/s per integrator step + HBM GB/s vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                          the reference's own kernels on the host cores
  python bench.py --scaling strong --total-ion-pairs 1769472      headline = a FIXED box partitioned over the N ranks

Workload (config.workload): BASELINE config 5, the synthetic Drude-polarizable ionic-liquid box of
SURVEY.md section 8(d) -- 442,368 ion pairs = 16,367,616 particles PER GPU (weak scaling: a box of
N x 16.4M particles partitioned by whole molecules), mixed precision, temperature-grouped Nose-Hoover
(atom / molecular-COM / Drude groups), middle scheme, Drude hard wall on, frozen synthetic forces
(force evaluation is OpenMM's and out of scope).  A "step" is one whole integrator step = pass A + pass B.

Beyond the headline every line also carries: `flows` (velocity-Verlet scheme, the constraint-bearing flow and the
reduce-only kernel at the same 16.4M particles), `config3_edl` (BASELINE config 3 at 40,310 particles),
`force_sigma_1000` (the survey's force width, with the hard-wall fire fraction of both widths), `parity` (N=1: against
the reference's own CUDA kernels on the same inputs; N>1: against ONE GPU stepping the whole box), `strong` (a FIXED
box of 16.4M / 65.5M particles partitioned over the N ranks), `e2e_resident` (state resident across K steps) and
`full_step`.  JSON keys beyond the base contract are documented in DESIGN.md (section "Measurement").
"""
import argparse
import dataclasses
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

METRIC = "integrator_particle_updates_per_s"
UNIT = "particle-updates/s"
# algorithmic bytes per particle: mixed precision
BYTES_PASS_A = 88      # R velm 32 + R force 24 + W velm 32
BYTES_PASS_B = 128     # R velm 32 + W velm 32 + R posq 16 + R corr 16 + W posq 16 + W corr 16
BYTES_REDUCE = 32      # R velm 32 (nothing is written)
BYTES_VV_STEP = 336    # first half 32 + 152, second half 88 + 64
BYTES_CONSTRAINED_STEP = 440    # kick 88 | reduce 32 + scale & deltas 128 | finish + hard wall 192
FALLBACK_HBM_GBS = 6650.0
# frozen forces are kept small so that 50+ steps without a force field stay physical (a frozen force heats
# linearly); the kernels read and convert the whole fixed-point force array regardless of its values.  The survey's
# N(0, 1000) is timed as well (`force_sigma_1000`): only the hard-wall fire rate depends on it.
FORCE_SIGMA = 1.0
SURVEY_FORCE_SIGMA = 1000.0
EV = 1.60217662e-22


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vvb200", choices=["vvb200", "reference"])
    ap.add_argument("--ion-pairs", type=int, default=442368, help="ion pairs per GPU (37 particles each), weak scaling")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (the driver's contract): --ion-pairs per GPU.  strong: --total-ion-pairs split over the ranks")
    ap.add_argument("--total-ion-pairs", type=int, default=1769472, help="the fixed box of --scaling strong")
    ap.add_argument("--strong-totals", default="442368,1769472",
                    help="fixed boxes (ion pairs) of the `strong` block every line carries; empty = skip")
    ap.add_argument("--precision", default="mixed", choices=["single", "mixed", "double"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-flows", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-full-step", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU exchange of the 10-double reduction vector: NVLink peer memory inside pass A's last block, or NCCL")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="budget of the whole --impl reference run")
    return ap.parse_args()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def per_gpu_ion_pairs(args, world):
    return args.ion_pairs if args.scaling == "weak" else args.total_ion_pairs // world


def dtype_label(precision):
    return {"mixed": "mixed(f32 pos / f64 vel)", "single": "f32", "double": "f64"}[precision]


def config_dict(args, world):
    """identical in both arms (the driver compares them)"""
    ip = per_gpu_ion_pairs(args, world)
    n = ip * 37
    return {"workload": (f"synthetic Drude ionic-liquid box (SURVEY 8d / BASELINE config 5): {ip} ion pairs = {n} "
                         f"particles per GPU x {world} GPU(s), TGNH 3 groups, middle scheme, hard wall 0.02 nm, dt 1 fs, "
                         f"frozen forces N(0,{FORCE_SIGMA:g}) kJ/mol/nm"),
            "precision": args.precision, "force_sigma": FORCE_SIGMA,
            "particles_per_gpu": n, "particles_total": n * world,
            "l2": "inputs larger than L2 (1.4 GB of state per GPU vs 126 MB; CPU arm: sample larger than the host caches)",
            "parallelism": f"molecule-partitioned x{world}" if world > 1 else "single GPU"}


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        return FALLBACK_HBM_GBS, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


# ---------------------------------------------------------------------------------------------
# clocks (NVML) sampled during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own kernel sources compiled for the host
# (oracle/_ref/libvvref_cpu_<mode>.so, SIMT shim + OpenMP over blocks); falls back to the plain-C port
# ---------------------------------------------------------------------------------------------
def cpu_arm(vv, vo, args, ion_pairs, steps, warmup, seconds_budget):
    """Times the reference CPU implementation of the same step on a bounded sample of the workload.
    Returns (value, ms_per_step, cpu_baseline dict)."""
    threads = host_threads()
    params = vv.Params(max_drude_distance=0.02)
    use_ref = vo.ref_available(args.precision, gpu=False)

    def make(n_ip):
        spec = vv.make_bulk_ionic_liquid(n_ip)
        par = params.resolved_for(spec)
        host = vv.make_state(spec, args.precision, force_sigma=FORCE_SIGMA)
        oracle = vo.Oracle(spec, par, args.precision, literal=False, threads=threads)
        runner = vo.Reference(oracle, gpu=False, threads=threads) if use_ref else oracle
        return spec, host, runner

    # calibrate on ~0.5M particles, then size the sample so that (warmup + steps) fit the budget
    spec, host, runner = make(13824)
    runner.step(host, steps=1)
    t0 = time.perf_counter()
    runner.step(host, steps=2)
    rate = 2 * spec.n / (time.perf_counter() - t0)
    per_step_budget = seconds_budget / max(1, steps + warmup)
    n_ip = int(min(ion_pairs, max(1024, rate * per_step_budget / 37)))
    if n_ip != 13824:
        del runner, host
        spec, host, runner = make(n_ip)
    runner.step(host, steps=max(1, warmup))
    t0 = time.perf_counter()
    runner.step(host, steps=steps)
    dt = time.perf_counter() - t0
    value = spec.n * steps / dt
    info = {"value": value, "unit": UNIT, "cores": threads,
            "kind": "reference" if use_ref else "port",
            "sample": (f"first {n_ip} of {ion_pairs} ion pairs ({spec.n} particles), {steps} steps after "
                       f"{max(1, warmup)} warm-up, "
                       + ("reference kernel sources compiled for the host (oracle/_ref, OpenMP over CUDA blocks)"
                          if use_ref else "plain-C oracle port (oracle/vv_oracle.c, OpenMP)")),
            "not_measured": "OpenMM's CPU platform (BASELINE.md row C): OpenMM is not in this image, and the plugin itself "
                            "ships no CPU kernels"}
    return value, 1e3 * dt / steps, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    vv = entry.load_package()
    vo = entry.load_oracle()
    world = max(world, args.gpus)
    value, ms, info = cpu_arm(vv, vo, args, per_gpu_ion_pairs(args, world), args.steps, args.warmup, seconds_budget=args.ref_seconds)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": dtype_label(args.precision), "data": "synthetic",
            "config": config_dict(args, world),
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# helpers of our arm
# ---------------------------------------------------------------------------------------------
def pinned_state(vv, host):
    """copy a HostState into page-locked memory (torch is the allocator: cudaHostAlloc)"""
    import torch

    def pin(a):
        if a is None:
            return None
        t = torch.from_numpy(a).pin_memory()
        return t.numpy(), t
    keep = []
    out = []
    for a in (host.posq, host.corr, host.velm, host.force):
        r = pin(a)
        if r is None:
            out.append(None)
        else:
            out.append(r[0])
            keep.append(r[1])
    st = vv.HostState(host.precision, out[0], out[1], out[2], out[3], host.random, host.box)
    st._keep = keep
    return st


def device_state(vv, torch, spec, precision, seed, force_sigma, temperature=333.0, drude_temperature=1.0,
                 density=158.0, drude_spread=0.005):
    """The synthetic state of system.make_state generated ON the device (same distributions, torch's generator): used
    where only device-resident stepping is timed (strong-scaling boxes, the second force width) -- numpy needs 8 s per
    16M particles.  Returns DeviceBuffers in OpenMM's layouts."""
    from vvb200.system import BOLTZ
    real = torch.float64 if precision == "double" else torch.float32
    mixed = torch.float32 if precision == "single" else torch.float64
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    n, P = spec.n, spec.padded_n
    L = (n / density) ** (1.0 / 3.0)
    masses = torch.from_numpy(np.ascontiguousarray(spec.masses, dtype=np.float64)).to(dev)
    massive = masses > 0
    safe_m = torch.where(massive, masses, torch.ones_like(masses))
    x = torch.rand((n, 3), generator=g, device=dev, dtype=torch.float64) * L
    v = torch.randn((n, 3), generator=g, device=dev, dtype=torch.float64)
    v *= torch.where(massive, torch.sqrt(BOLTZ * temperature / safe_m), torch.zeros_like(masses))[:, None]
    q = torch.randn(n, generator=g, device=dev, dtype=torch.float64) * 0.5
    if spec.drude_pairs.size:
        d = torch.from_numpy(spec.drude_pairs[:, 0].astype(np.int64)).to(dev)
        p = torch.from_numpy(spec.drude_pairs[:, 1].astype(np.int64)).to(dev)
        x[d] = x[p] + torch.randn((d.numel(), 3), generator=g, device=dev, dtype=torch.float64) * drude_spread
        mu = masses[d] * masses[p] / (masses[d] + masses[p])
        v[d] = v[p] + torch.randn((d.numel(), 3), generator=g, device=dev, dtype=torch.float64) * torch.sqrt(BOLTZ * drude_temperature / mu)[:, None]
        q[d] = -(1.0 + torch.randn(d.numel(), generator=g, device=dev, dtype=torch.float64).abs() * 0.3)
        del d, p, mu
    posq = torch.zeros((P, 4), dtype=real, device=dev)
    corr = None
    if precision == "mixed":
        hi = x.float()
        posq[:n, :3] = hi
        corr = torch.zeros((P, 4), dtype=torch.float32, device=dev)
        corr[:n, :3] = (x - hi.double()).float()
        del hi
    else:
        posq[:n, :3] = x.to(real)
    posq[:n, 3] = q.to(real)
    del x, q
    velm = torch.zeros((P, 4), dtype=mixed, device=dev)
    velm[:n, :3] = v.to(mixed)
    velm[:n, 3] = torch.where(massive, 1.0 / safe_m, torch.zeros_like(masses)).to(mixed)
    del v
    force = torch.zeros((3, P), dtype=torch.int64, device=dev)
    f = torch.randn((3, n), generator=g, device=dev, dtype=torch.float64) * force_sigma
    f *= massive[None, :]
    force[:, :n] = (f * 4294967296.0).to(torch.int64)          # truncation toward zero, like (long long)
    del f
    return vv.DeviceBuffers.from_tensors(precision, posq, corr, velm, force, box=(L, L, L))


def clone_buffers(vv, b, with_pos_delta=False):
    import torch
    c = lambda t: t.clone() if t is not None else None
    return vv.DeviceBuffers.from_tensors(b.precision, c(b.posq), c(b.corr), c(b.velm), c(b.force), random=c(b.random), box=b.box,
                                         pos_delta=torch.zeros_like(b.velm) if with_pos_delta else None)


def positions64(b, n):
    x = b.posq[:n, :3].double()
    if b.corr is not None:
        x = x + b.corr[:n, :3].double()
    return x


def rel_err_t(torch, a, b):
    """tests/conftest.py::rel_err on the device: max |a-b| / max(|b|, 1e-3 rms(b))"""
    if a.numel() == 0:
        return 0.0
    a, b = a.double(), b.double()
    scale = torch.clamp(b.abs(), min=float(1e-3 * torch.sqrt(torch.mean(b * b)).item() + 1e-300))
    return float(torch.max((a - b).abs() / scale).item())


def reset_thermostat(plan):
    st = plan.thermostat_state()
    plan.set_thermostat_state(np.zeros_like(st["eta"]), np.zeros_like(st["eta_dot"]), np.zeros_like(st["eta_dotdot"]))


def time_calls(torch, fn, steps, warmup=3, stream=None):
    """microseconds per call of fn(), CUDA events on the launching stream, synchronised on both sides"""
    stream = stream or torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / steps


def small_system_leg(vv, torch, precision, n_ip=1250, steps=400, graph=True):
    """BASELINE configs[1]: 1,250 ion pairs = 46,250 particles, TGNH, middle scheme, hard wall.  The whole step is ONE
    launch (csrc/vvb200_resident.cuh); timed eagerly through the C ABI and replayed from a CUDA graph."""
    spec = vv.make_bulk_ionic_liquid(n_ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, precision, force_sigma=FORCE_SIGMA)
    plan = vv.Plan(spec, params, precision).upload()
    b = vv.DeviceBuffers(host)
    st = torch.cuda.current_stream()
    for _ in range(10):
        plan.step_middle(b)
    torch.cuda.synchronize()
    l0, r0 = plan.launch_count, plan.resident_launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        plan.step_middle(b)
    e1.record(st)
    torch.cuda.synchronize()
    eager_us = 1e3 * e0.elapsed_time(e1) / steps
    launches = (plan.launch_count - l0) / steps
    resident = (plan.resident_launch_count - r0) / steps
    graph_us = None
    try:
        if not graph:
            raise RuntimeError("not requested")
        side = torch.cuda.Stream()
        side.wait_stream(st)
        with torch.cuda.stream(side):
            plan.step_middle(b)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20):
                    plan.step_middle(b)
            for _ in range(3):
                g.replay()
            side.synchronize()
            e0.record(side)
            for _ in range(steps // 20):
                g.replay()
            e1.record(side)
            side.synchronize()
        graph_us = 1e3 * e0.elapsed_time(e1) / (20 * (steps // 20))
    except Exception as e:  # noqa: BLE001
        graph_us = f"capture failed: {e}"
    return {"workload": f"BASELINE configs[1]: Drude ionic-liquid bulk, {n_ip} ion pairs = {spec.n} particles, TGNH 3 groups, "
                        f"middle scheme, hard wall, {precision}",
            "particles": spec.n, "us_per_step": eager_us, "cuda_graph_us_per_step": graph_us, "launches_per_step": launches,
            "single_launch_resident_steps_per_step": resident,
            "value": spec.n / (eager_us * 1e-6), "unit": UNIT,
            "note": "latency-bound (1.5 MB of state): no roofline fraction; the reference's kernels need 10 launches + a "
                    "blocking host round trip for the same step (profiles/configs_r02.json)"}


def sweep_leg(vv, torch, precision, n_ip, peak, warmup=5, steps=20, repeats=5, long_steps=200):
    """Whole-step time at another size under the HEADLINE's protocol: `steps` steps after `warmup` warm-up steps from the
    Maxwell-Boltzmann start state (reloaded, thermostat chains zeroed, for each of `repeats` measurements; the median is
    reported).  The protocol matters: frozen forces pump the Drude oscillators, so the hard wall -- the path's one
    data-dependent branch -- fires more and more often over hundreds of steps (`us_per_step_after_200_steps` shows it);
    a simulation with a force field stays at the start state's fire rate."""
    spec = vv.make_bulk_ionic_liquid(n_ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, precision, force_sigma=FORCE_SIGMA)
    plan = vv.Plan(spec, params, precision).upload()
    b = vv.DeviceBuffers(host)
    times = []
    l0 = plan.launch_count
    for _ in range(repeats):
        b.load(host)
        reset_thermostat(plan)
        times.append(time_calls(torch, lambda: plan.step_middle(b), steps, warmup=warmup))
    launches = (plan.launch_count - l0) / (repeats * (steps + warmup))
    us = float(np.median(times))
    for _ in range(long_steps):
        plan.step_middle(b)
    late = time_calls(torch, lambda: plan.step_middle(b), steps, warmup=0)
    gbs = (BYTES_PASS_A + BYTES_PASS_B) * spec.n / (us * 1e-6) / 1e9
    return {"particles": spec.n, "us_per_step": us, "us_per_step_all_repeats": times, "launches_per_step": launches,
            "step_gbs": gbs, "step_frac": gbs / peak,
            "protocol": f"{steps} steps after {warmup} warm-up steps from the start state, median of {repeats} (the headline's protocol)",
            "us_per_step_after_200_steps": late}


def flows_leg(vv, torch, args, spec, params, plan, bufs, peak, steps):
    """The other flows of the path at the headline size (single-GPU runs), whole calls timed with CUDA events and no
    per-kernel events in between; `frac` = algorithmic bytes / time / measured copy peak."""
    out = []
    n = spec.n

    def row(name, us, bpp, launches, extra=None):
        gbs = bpp * n / (us * 1e-6) / 1e9
        r = {"flow": name, "particles": n, "us_per_step": us, "bytes_per_particle": bpp, "gbs": gbs, "frac": gbs / peak,
             "launches_per_step": launches}
        if extra:
            r.update(extra)
        out.append(r)

    # (1) velocity-Verlet scheme (VVIntegrator.cpp:272-338): first half | forces | second half
    pv = dataclasses.replace(params, use_middle_scheme=False)
    plan_vv = vv.Plan(spec, pv, args.precision).upload()
    b = clone_buffers(vv, bufs)

    def vv_step():
        plan_vv.step_vv_first(b)
        plan_vv.step_vv_second(b)
    l0 = plan_vv.launch_count
    us = time_calls(torch, vv_step, steps)
    row("velocity-Verlet scheme: step_vv_first + step_vv_second", us, BYTES_VV_STEP, (plan_vv.launch_count - l0) / (steps + 3))
    del plan_vv

    # (2) the constraint-bearing flow of both example scripts (CudaVVKernels.cpp:144-220): kick | thermostat + deltas |
    #     finish + hard wall; OpenMM's own constraint launches are not included
    b.pos_delta = torch.zeros_like(b.velm)

    def constrained_step():
        plan.middle_kick(b)
        plan.middle_thermostat_delta(b)
        plan.middle_finish(b)
    l0 = plan.launch_count
    us = time_calls(torch, constrained_step, steps)
    lc = (plan.launch_count - l0) / (steps + 3)
    parts = {}
    for name, fn in (("middle_kick", lambda: plan.middle_kick(b)), ("middle_thermostat_delta", lambda: plan.middle_thermostat_delta(b)),
                     ("middle_finish", lambda: plan.middle_finish(b))):
        parts[name + "_us"] = time_calls(torch, fn, max(5, steps // 2))
    row("constraint-bearing middle flow: middle_kick | middle_thermostat_delta | middle_finish", us, BYTES_CONSTRAINED_STEP, lc, parts)

    # (3) the reduce-only kernel alone (gates vvb200_thermostat, the constrained flow and the VV first half)
    plan.profile_enable(steps)
    for _ in range(steps):
        plan.thermostat(b)
    torch.cuda.synchronize()
    a_ms, b_ms, k = plan.profile_read()
    plan.profile_enable(0)
    if k:
        row("reduce-only pass (reduce_only_kernel inside vvb200_thermostat)", 1e3 * a_ms / k, BYTES_REDUCE, 1)
        row("scale-only pass (inside vvb200_thermostat)", 1e3 * b_ms / k, 64, 1)
    del b
    torch.cuda.empty_cache()
    return out


def config3_leg(vv, torch, precision, steps=300):
    """BASELINE configs[2] / SURVEY 8(d) C3: EDL box of 40,310 particles -- Langevin electrodes, field on the electrolyte,
    image charges, hard wall: fused step and the constraint-bearing flow (run-edl.py uses HBonds)."""
    spec = vv.make_edl(n_ion_pairs=511, n_electrode=2496, electrode_molecules=4)
    params = vv.Params(max_drude_distance=0.02, mirror_location=8.0, electric_field=0.25 * EV).resolved_for(spec)
    host = vv.make_state(spec, precision, force_sigma=FORCE_SIGMA, mirror=8.0, n_random=2500 * 64)
    plan = vv.Plan(spec, params, precision).upload()
    b = vv.DeviceBuffers(host, with_pos_delta=True)
    req = plan.random_request
    wrap = max(1, host.random.shape[0] - 2 * req)
    ri = [0]

    def fused():
        plan.step_middle(b, random_index=ri[0])
        ri[0] = (ri[0] + req) % wrap

    def constrained():
        plan.middle_kick(b, random_index=ri[0])
        plan.middle_thermostat_delta(b)
        plan.middle_finish(b)
        ri[0] = (ri[0] + req) % wrap
    l0 = plan.launch_count
    us = time_calls(torch, fused, steps, warmup=5)
    lf = (plan.launch_count - l0) / (steps + 5)
    l0 = plan.launch_count
    us_c = time_calls(torch, constrained, steps, warmup=5)
    lc = (plan.launch_count - l0) / (steps + 5)
    # the same three launches replayed from a CUDA graph: eagerly, three Python / ctypes calls per step cost the host more
    # than the kernels cost the GPU, so the eager figure is largely the harness's launch rate (fixed random index: the
    # arithmetic cost is identical)
    try:
        st = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(st)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            constrained()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                for _ in range(10):
                    constrained()
            for _ in range(3):
                g.replay()
            side.synchronize()
            e0.record(side)
            for _ in range(max(1, steps // 10)):
                g.replay()
            e1.record(side)
            side.synchronize()
        us_cg = 1e3 * e0.elapsed_time(e1) / (10 * max(1, steps // 10))
        del g
    except Exception as e:  # noqa: BLE001
        us_cg = f"capture failed: {e}"
    return {"workload": f"BASELINE configs[2] (SURVEY 8d C3): EDL, {spec.n} particles, Langevin subset + electric field + image charges + "
                        f"hard wall, middle scheme, {precision}",
            "particles": spec.n, "us_per_step": us, "launches_per_step": lf,
            "constrained_flow_us_per_step": us_c, "constrained_flow_launches_per_step": lc,
            "constrained_flow_cuda_graph_us_per_step": us_cg,
            "value": spec.n / (us * 1e-6), "unit": UNIT, "note": "latency-bound: microseconds and launches, no roofline fraction"}


def hardwall_fire_fraction(vv, torch, args, spec, params, bufs_state):
    """Fraction of Drude pairs the hard wall acts on in ONE step from `bufs_state`: the wall fires exactly where the
    drifted pair distance exceeds maxDrudeDistance (middle.cu:129-141), i.e. where the same step WITHOUT the wall ends."""
    if not spec.drude_pairs.size:
        return 0.0
    nowall = vv.Plan(spec, dataclasses.replace(params, max_drude_distance=0.0), args.precision).upload()
    c = clone_buffers(vv, bufs_state)
    nowall.step_middle(c)
    torch.cuda.synchronize()
    x = positions64(c, spec.n)
    d = torch.from_numpy(spec.drude_pairs[:, 0].astype(np.int64)).cuda()
    p = torch.from_numpy(spec.drude_pairs[:, 1].astype(np.int64)).cuda()
    r = torch.linalg.norm(x[d] - x[p], dim=1)
    frac = float((r > params.max_drude_distance).double().mean().item())
    del nowall, c, x, d, p, r
    torch.cuda.empty_cache()
    return frac


def sigma1000_leg(vv, torch, args, spec, params, steps):
    """the headline step with the survey's force width N(0,1000): same bytes, a different hard-wall fire rate"""
    plan = vv.Plan(spec, params, args.precision).upload()
    b = device_state(vv, torch, spec, args.precision, seed=777, force_sigma=SURVEY_FORCE_SIGMA)
    for _ in range(3):
        plan.step_middle(b)
    fire = hardwall_fire_fraction(vv, torch, args, spec, params, b)
    us = time_calls(torch, lambda: plan.step_middle(b), steps, warmup=2)
    finite = bool(torch.isfinite(b.velm).all().item())
    del plan, b
    torch.cuda.empty_cache()
    return {"force_sigma": SURVEY_FORCE_SIGMA, "ms_per_step": us * 1e-3, "value": spec.n / (us * 1e-6), "unit": UNIT,
            "step_frac_of_peak": None,
            "hardwall_fire_fraction_per_step": fire, "finite_after_run": finite,
            "note": "frozen N(0,1000) kJ/mol/nm forces drive the Drude pairs to the wall within a few steps (no force field holds "
                    "them); fire fraction measured on the 4th step as the share of pairs whose wall-free drift ends beyond 0.02 nm"}


def full_step_leg(vv, torch, precision, steps=300):
    """north_star's "full-step ns/day": needs OpenMM's force evaluation, which is not in the image.  What CAN be run is
    a toy force field (harmonic tethers + Drude springs, tests/test_long_run_statistics.py) evaluated by torch ops on
    the device each step around our integrator step -- labelled as such; it says nothing about OpenMM's nonbonded cost."""
    spec = vv.make_bulk_ionic_liquid(1250)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, precision, force_sigma=0.0, drude_spread=0.0005)
    plan = vv.Plan(spec, params, precision).upload()
    b = vv.DeviceBuffers(host)
    n = spec.n
    dev = b.posq.device
    d_idx = torch.as_tensor(spec.drude_pairs[:, 0].astype(np.int64), device=dev)
    p_idx = torch.as_tensor(spec.drude_pairs[:, 1].astype(np.int64), device=dev)
    tether = torch.as_tensor(spec.masses > 0, device=dev)
    tether[d_idx] = False
    x0 = positions64(b, n).clone()
    k_tether, k_drude = 5000.0, 4.184e5          # kJ/mol/nm^2, as tests/test_long_run_statistics.py

    def full():
        x = positions64(b, n)
        f = torch.where(tether[:, None], -k_tether * (x - x0), torch.zeros_like(x))
        fi = (f * 4294967296.0).to(torch.int64)
        fd = (-k_drude * (x[d_idx] - x[p_idx]) * 4294967296.0).to(torch.int64)
        fi.index_add_(0, d_idx, fd)
        fi.index_add_(0, p_idx, -fd)
        b.force[:, :n] = fi.t()
        plan.step_middle(b)
    us_full = time_calls(torch, full, steps, warmup=10)
    us_int = time_calls(torch, lambda: plan.step_middle(b), steps, warmup=10)
    return {"status": "blocked: full-step ns/day needs OpenMM's force evaluation (nonbonded, PME, bonded, Drude); OpenMM is not "
                      "in this image and cannot be installed offline",
            "toy_force_full_step": {"particles": n, "us_per_step": us_full, "integrator_us_per_step": us_int,
                                    "ns_per_day": 86400.0 / (us_full * 1e-6) * params.step_size * 1e-3,
                                    "forces": "TOY: harmonic tether on every atom + Drude spring per pair, evaluated by ~15 torch "
                                              "element-wise launches per step (not OpenMM forces, not part of this library)"}}


def parity_single_gpu(vv, torch, vo, args, spec, params, host, steps=3):
    """ours vs the reference's own CUDA kernels (oracle/_ref/libvvref_cuda, unmodified platforms/cuda/src/kernels/*.cu)
    from the same initial state, element-wise, at the headline size.  Returns (parity dict, reference timing dict)."""
    stream = torch.cuda.current_stream()
    oracle = vo.Oracle(spec, params, args.precision, literal=False)     # supplies the index arrays only
    ref = vo.Reference(oracle, gpu=True)
    rb = vv.DeviceBuffers(host)
    ref.step(rb, steps=steps)
    torch.cuda.synchronize()
    plan = vv.Plan(spec, params, args.precision).upload()
    ob = vv.DeviceBuffers(host)
    plan.step(ob, steps=steps)
    torch.cuda.synchronize()
    n = spec.n
    ev = rel_err_t(torch, ob.velm[:n, :3], rb.velm[:n, :3])
    ex = rel_err_t(torch, positions64(ob, n), positions64(rb, n))
    parity = {"against": "the reference's own CUDA kernels compiled for sm_100a (oracle/_ref/libvvref_cuda), same inputs, same GPU",
              "particles": n, "steps": steps, "max_rel_v": ev, "max_rel_x": ex, "max_rel": max(ev, ex),
              "metric": "max |a-b| / max(|b|, 1e-3 rms(b)) (tests/conftest.py::rel_err)", "bar": 1e-6}
    del plan, ob
    # timing of the reference kernels (reported, not the headline)
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ksteps = 5
    r0.record(stream)
    nl = ref.step(rb, steps=ksteps)
    r1.record(stream)
    torch.cuda.synchronize()
    rms = r0.elapsed_time(r1) / ksteps
    ref_gpu = {"what": "reference kernels (platforms/cuda/src/kernels/*.cu, unmodified) compiled for sm_100a, "
                       "OpenMM launch geometry, blocking D2H/H2D around the host NH chain (oracle/_ref)",
               "ms_per_step": rms, "value": n / (rms * 1e-3), "unit": UNIT, "launches_per_step": nl / ksteps}
    del ref, rb, oracle
    torch.cuda.empty_cache()
    return parity, ref_gpu


def parity_multi_gpu(vv, torch, dist, args, spec, params, host, dplan, rank, world, steps=2):
    """N ranks vs ONE GPU stepping the whole box: every rank restarts from its initial state and zeroed NH chains, steps
    `steps` times through the distributed path; rank 0 gathers every partition's initial and final arrays, steps the whole
    box (world x particles) on its own GPU with a single plan, and compares element-wise.  The scale factors must be
    BITWISE equal on all ranks (each sums the same slots in rank order)."""
    n = spec.n
    bufs = vv.DeviceBuffers(host)
    reset_thermostat(dplan.plan)
    mixed = args.precision == "mixed"
    init = [bufs.posq[:n].clone(), bufs.corr[:n].clone() if mixed else None, bufs.velm[:n].clone(), bufs.force[:, :n].t().contiguous()]
    for _ in range(steps):
        dplan.step_middle(bufs)
    torch.cuda.synchronize()
    final = [bufs.posq[:n], bufs.corr[:n] if mixed else None, bufs.velm[:n]]
    st = dplan.plan.thermostat_state()
    vs = torch.tensor(np.asarray(st["vscale"][:3], dtype=np.float64), device="cuda").view(torch.int64)
    all_vs = [torch.zeros_like(vs) for _ in range(world)]
    dist.all_gather(all_vs, vs)
    vs_bitwise = all(bool(torch.equal(all_vs[0], v)) for v in all_vs)

    def gather(t):
        if t is None:
            return None
        lst = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
        dist.gather(t.contiguous(), lst, dst=0)
        return lst
    g_init = [gather(t) for t in init]
    g_final = [gather(t) for t in final]
    out = None
    if rank == 0:
        whole = vv.make_bulk_ionic_liquid(world * (n // 37))
        N, P = whole.n, whole.padded_n
        posq = torch.zeros((P, 4), dtype=init[0].dtype, device="cuda")
        posq[:N] = torch.cat(g_init[0])
        corr = None
        if mixed:
            corr = torch.zeros((P, 4), dtype=torch.float32, device="cuda")
            corr[:N] = torch.cat(g_init[1])
        velm = torch.zeros((P, 4), dtype=init[2].dtype, device="cuda")
        velm[:N] = torch.cat(g_init[2])
        force = torch.zeros((3, P), dtype=torch.int64, device="cuda")
        force[:, :N] = torch.cat(g_init[3]).t()
        del g_init
        wb = vv.DeviceBuffers.from_tensors(args.precision, posq, corr, velm, force, box=host.box)
        wplan = vv.Plan(whole, params.resolved_for(whole), args.precision).upload()
        for _ in range(steps):
            wplan.step_middle(wb)
        torch.cuda.synchronize()
        ev = ex = 0.0
        per_rank = []
        for r in range(world):
            lo, hi = r * n, (r + 1) * n
            xr = g_final[0][r][:, :3].double() + (g_final[1][r][:, :3].double() if mixed else 0.0)
            xw = wb.posq[lo:hi, :3].double() + (wb.corr[lo:hi, :3].double() if mixed else 0.0)
            e_v = rel_err_t(torch, g_final[2][r][:, :3], wb.velm[lo:hi, :3])
            e_x = rel_err_t(torch, xr, xw)
            per_rank.append(max(e_v, e_x))
            ev, ex = max(ev, e_v), max(ex, e_x)
        wst = wplan.thermostat_state()
        a, b = np.asarray(st["vscale"][:3]), np.asarray(wst["vscale"][:3])
        out = {"against": f"ONE GPU stepping the whole box ({N} particles, a single plan) from the same initial state",
               "particles": N, "ranks": world, "steps": steps, "max_rel_v": ev, "max_rel_x": ex, "max_rel": max(ev, ex),
               "max_rel_per_rank_partition": per_rank, "vscale_bitwise_equal_across_ranks": vs_bitwise,
               "vscale_rel_vs_single_gpu": float(np.max(np.abs(a - b) / np.abs(b))),
               "metric": "max |a-b| / max(|b|, 1e-3 rms(b)) (tests/conftest.py::rel_err)", "bar": 1e-9}
        del wplan, wb, posq, corr, velm, force
    del g_final, bufs
    torch.cuda.empty_cache()
    dist.barrier()
    return out


def strong_leg(vv, torch, dist, args, total_ion_pairs, rank, world, steps, warmup, exchange):
    """A FIXED box of `total_ion_pairs` partitioned by whole molecules over the `world` ranks (strong scaling).  Device-
    generated state; timed like the headline (barrier + synchronize on both sides, max over ranks)."""
    ip = total_ion_pairs // world
    spec = vv.make_bulk_ionic_liquid(ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    dplan = vv.DistributedPlan(spec, params, args.precision).upload(peer=exchange)
    b = device_state(vv, torch, spec, args.precision, seed=4242 + rank, force_sigma=FORCE_SIGMA)
    stream = torch.cuda.current_stream()

    def one():
        if world == 1:
            dplan.plan.step_middle(b)
        else:
            dplan.step_middle(b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(warmup):
        one()
    barrier()
    l0 = dplan.plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        one()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = (dplan.plan.launch_count - l0) / steps
    res = {"total_ion_pairs": ip * world, "particles_total": spec.n * world, "particles_per_gpu": spec.n, "n_gpus": world,
           "steps": steps, "ms_per_step": ms / steps, "value": spec.n * world * steps / (ms * 1e-3), "unit": UNIT,
           "launches_per_step_per_rank": launches,
           "exchange": ("NVLink peer memory inside pass A's last block" if dplan.peer else "NCCL all-reduce of 10 doubles") if world > 1 else "none"}
    del dplan, b
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    vv = entry.load_package()
    K, W = args.steps, max(args.warmup, 3)
    ion_pairs = per_gpu_ion_pairs(args, world)
    exchange = {"auto": None, "nccl": False, "peer": True}[args.exchange]
    spec = vv.make_bulk_ionic_liquid(ion_pairs)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, args.precision, seed=12345 + 100 * rank, force_sigma=FORCE_SIGMA)
    # this rank holds one whole-molecule partition of a box `world` times larger: DistributedPlan all-reduces the
    # thermostat DOFs and the total mass of the whole box at set-up (world == 1: the plan's own)
    dplan = vv.DistributedPlan(spec, params, args.precision).upload(peer=exchange)
    plan = dplan.plan
    bufs = vv.DeviceBuffers(host)
    n_local = spec.n
    stream = torch.cuda.current_stream()

    def one_step():
        if world == 1:
            plan.step_middle(bufs)                     # pass A (NH chains in its last block) + pass B
        else:
            dplan.step_middle(bufs)                    # pass A (+ exchange of <= 10 doubles over NVLink in its last block) + pass B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        one_step()
    barrier()

    # ---- device-resident timing: EXACTLY K steps between two events on the launching stream.  The first K/5 of them
    #      also carry an event pair around every launch (vvb200_profile_*) for the per-kernel roofline; only a sample,
    #      because an event between two launches forbids the overlap of pass B's launch with pass A's tail
    #      (programmatic dependent launch): ~2 % of a 16M-particle step, ~25 % of a 1M-particle one. --------------
    prof_n = max(1, K // 5)
    plan.profile_enable(prof_n)
    launches0 = plan.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record(stream)
        for _ in range(K):
            one_step()
        e1.record(stream)
        barrier()
    ms_total = e0.elapsed_time(e1)
    launches = plan.launch_count - launches0 + (K if world > 1 and not dplan.peer else 0)   # + NCCL's all-reduce kernel
    ms_a, ms_b, prof_steps = plan.profile_read()
    plan.profile_enable(0)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / K
    n_global = n_local * world
    value = n_global * K / (ms_total * 1e-3)
    st_head = plan.thermostat_state()

    # ---- roofline of the dominant kernel (this rank) ----------------------------------------------
    peak, peak_src = peak_hbm()
    ka = {"kernel": "kick_reduce_kernel (pass A)", "bytes_per_particle": BYTES_PASS_A,
          "ms": ms_a / max(prof_steps, 1)}
    kb = {"kernel": "scale_drift_kernel (pass B)", "bytes_per_particle": BYTES_PASS_B,
          "ms": ms_b / max(prof_steps, 1)}
    for k in (ka, kb):
        k["achieved_gbs"] = k["bytes_per_particle"] * n_local / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
        k["frac"] = k["achieved_gbs"] / peak if k["achieved_gbs"] else None
    dom = kb if kb["ms"] >= ka["ms"] else ka
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": (dom["achieved_gbs"] / peak) if dom["achieved_gbs"] else None,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["bytes_per_particle"] * n_local,
                "kernels": [ka, kb],
                "kernel_timing": f"CUDA events recorded by the library on the launching stream around every launch of the first "
                                 f"{prof_steps} of the {K} timed steps (events between launches forbid launch overlap)",
                "step_gbs": (BYTES_PASS_A + BYTES_PASS_B) * n_local / (ms_per_step * 1e-3) / 1e9,
                "step_frac": (BYTES_PASS_A + BYTES_PASS_B) * n_local / (ms_per_step * 1e-3) / 1e9 / peak}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            with open(traffic_file) as f:
                tr = json.load(f)
            roofline["traffic"] = tr.get(dom["kernel"].split()[0])
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    # ---- end to end through the host-buffer entry point ------------------------------------------------
    e2e = e2e_resident = None
    if not args.no_e2e:
        ke = args.e2e_steps or min(K, 10)
        pst = pinned_state(vv, host)
        real_b = 4 if args.precision != "double" else 8
        mixed_b = 4 if args.precision == "single" else 8
        P = spec.padded_n
        h2d = P * (4 * real_b * (2 if args.precision == "mixed" else 1) + 4 * mixed_b + 24)
        d2h = P * (4 * real_b * (2 if args.precision == "mixed" else 1) + 4 * mixed_b)
        if world == 1:
            def e2e_step():
                plan.step_host(pst, steps=1)           # vvb200_step_host: H2D + step + D2H, synchronises
        else:
            def e2e_step():
                dplan.step_host(pst)                   # vvb200_step_host_begin | exchange of 10 doubles | _finish; synchronises
        for _ in range(3):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": n_global * ke / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": ke, "ms_per_step": 1e3 * dt / ke,
               "api": "vvb200_step_host (pinned host buffers)" if world == 1 else
                      "vvb200_step_host_begin + exchange of the reduction vector + vvb200_step_host_finish (pinned host buffers)"}
        # secondary: the state stays resident across K steps between one copy-in and one copy-out -- what an OpenMM Context
        # does between two getState() calls (single-GPU entry point)
        if world == 1:
            kr = 20
            plan.step_host(pst, steps=kr)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            plan.step_host(pst, steps=kr)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            e2e_resident = {"value": n_global * kr / dt, "unit": UNIT, "steps_per_call": kr, "ms_per_step": 1e3 * dt / kr,
                            "h2d_bytes_per_call": int(h2d), "d2h_bytes_per_call": int(d2h),
                            "api": f"vvb200_step_host(steps={kr}): one copy-in, {kr} steps, one copy-out"}
        del pst

    # ---- parity at the headline size --------------------------------------------------------------------
    parity = ref_gpu = None
    if not args.no_parity:
        if world == 1 and not args.no_ref_gpu:
            vo = entry.load_oracle()
            if vo.ref_available(args.precision, gpu=True):
                parity, ref_gpu = parity_single_gpu(vv, torch, vo, args, spec, params, host)
                ref_gpu["speedup_device_resident"] = ref_gpu["ms_per_step"] / ms_per_step
            else:
                parity = {"status": "oracle/_ref/libvvref_cuda not built"}
        elif world > 1:
            parity = parity_multi_gpu(vv, torch, dist, args, spec, params, host, dplan, rank, world)

    # ---- the other flows at this size, the survey's force width, config 3 (single-GPU runs) -----------
    flows = sigma1000 = config3 = None
    if world == 1 and not args.no_flows:
        flows = flows_leg(vv, torch, args, spec, params, plan, bufs, peak, steps=max(5, K // 2))
        fire1 = hardwall_fire_fraction(vv, torch, args, spec, params, bufs)
        sigma1000 = sigma1000_leg(vv, torch, args, spec, params, steps=max(5, K // 2))
        sigma1000["step_frac_of_peak"] = (BYTES_PASS_A + BYTES_PASS_B) * n_local / (sigma1000["ms_per_step"] * 1e-3) / 1e9 / peak
        sigma1000["hardwall_fire_fraction_per_step_at_sigma_1"] = fire1
        config3 = config3_leg(vv, torch, args.precision)

    # ---- BASELINE configs[1] (run-bulk.py-sized Drude bulk, ~50k particles): latency-bound, so microseconds and
    #      launches per step instead of a roofline fraction; rank 0, single-GPU runs ---------------
    config2 = None
    if rank == 0 and world == 1 and not args.no_config2:
        config2 = small_system_leg(vv, torch, args.precision)

    # ---- the >= 1M-particle target of BASELINE.json at other sizes (rank 0, single-GPU runs): whole-step time and
    #      the fraction of the HBM peak the 216 algorithmic bytes per particle amount to -------------------------
    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = [sweep_leg(vv, torch, args.precision, n_ip, peak, warmup=W, steps=K) for n_ip in (27648, 110592)]

    full_step = None
    if rank == 0 and world == 1 and not args.no_full_step:
        full_step = full_step_leg(vv, torch, args.precision)
    elif rank == 0:
        full_step = {"status": "blocked: full-step ns/day needs OpenMM's force evaluation; OpenMM is not in this image"}

    # ---- strong scaling: fixed boxes partitioned over the ranks (all ranks take part) ---------------------------
    del bufs
    torch.cuda.empty_cache()
    strong = []
    for tok in [s for s in args.strong_totals.split(",") if s.strip()]:
        total = int(tok)
        if args.scaling == "weak" and world == 1 and total == args.ion_pairs:
            strong.append({"total_ion_pairs": total, "particles_total": n_global, "particles_per_gpu": n_local, "n_gpus": 1,
                           "steps": K, "ms_per_step": ms_per_step, "value": value, "unit": UNIT,
                           "launches_per_step_per_rank": launches / K, "exchange": "none", "note": "= the headline run"})
            continue
        if total // world < 64:
            continue
        strong.append(strong_leg(vv, torch, dist, args, total, rank, world, steps=K, warmup=W, exchange=exchange))

    # ---- CPU baseline (rank 0, single-GPU runs only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vo = entry.load_oracle()                 # bench.py's cpu_baseline leg: the one place the product bench runs oracle/
        _, _, cpu = cpu_arm(vv, vo, args, ion_pairs, steps=5, warmup=1, seconds_budget=args.cpu_seconds)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
                "dtype": dtype_label(args.precision), "data": "synthetic",
                "config": config_dict(args, world),
                "exchange": ("NVLink peer memory, fused into pass A's last block" if dplan.peer else
                             "NCCL all-reduce of 10 doubles") if world > 1 else "none",
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": int(launches),
                "parity": parity, "flows": flows, "config3_edl": config3, "force_sigma_1000": sigma1000,
                "strong": strong, "full_step": full_step,
                "reference_kernels_on_gpu": ref_gpu,
                "config2_small_system": config2, "size_sweep": sweep,
                "clocks": clocks.summary(),
                "integrator_only_ns_per_day": 86400.0 / (ms_per_step * 1e-3) * params.step_size * 1e-3,
                "thermostat": {"ke2": [float(x) for x in st_head["ke2"]], "vscale": [float(x) for x in st_head["vscale"]]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
