#!/usr/bin/env python
"""bench.py -- integrator-only throughput of the B200-native VVIntegrator path (BASELINE.json metric:
particle-updates/s per integrator step + HBM GB/s vs roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference [...]                          the reference's own kernels on the host cores

Workload (config.workload): BASELINE config 5, the synthetic Drude-polarizable ionic-liquid box of
SURVEY.md section 8(d) -- 442,368 ion pairs = 16,367,616 particles PER GPU (weak scaling: a box of
N x 16.4M particles partitioned by whole molecules), mixed precision, temperature-grouped Nose-Hoover
(atom / molecular-COM / Drude groups), middle scheme, Drude hard wall on, frozen synthetic forces
(force evaluation is OpenMM's and out of scope).  A "step" is one whole integrator step = pass A + pass B.

JSON keys beyond the base contract are documented in DESIGN.md (section "Measurement").
"""
import argparse
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
# algorithmic bytes per particle: mixed precision, middle scheme, no constraints
BYTES_PASS_A = 88      # R velm 32 + R force 24 + W velm 32
BYTES_PASS_B = 128     # R velm 32 + W velm 32 + R posq 16 + R corr 16 + W posq 16 + W corr 16
FALLBACK_HBM_GBS = 6650.0
# frozen forces are kept small so that 50+ steps without a force field stay physical (a frozen force heats
# linearly); the kernels read and convert the whole fixed-point force array regardless of its values
FORCE_SIGMA = 1.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vvb200", choices=["vvb200", "reference"])
    ap.add_argument("--ion-pairs", type=int, default=442368, help="ion pairs per GPU (37 particles each)")
    ap.add_argument("--precision", default="mixed", choices=["single", "mixed", "double"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 10))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--no-config2", action="store_true")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl", "peer"],
                    help="multi-GPU exchange of the 10-double reduction vector: NVLink peer memory inside the NHC kernel, or NCCL")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="budget of the whole --impl reference run")
    return ap.parse_args()


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def workload_name(args, world):
    n = args.ion_pairs * 37
    return (f"synthetic Drude ionic-liquid box (SURVEY 8d / BASELINE config 5): {args.ion_pairs} ion pairs = {n} "
            f"particles per GPU x {world} GPU(s), TGNH 3 groups, middle scheme, hard wall 0.02 nm, dt 1 fs, frozen forces N(0,{FORCE_SIGMA:g}) kJ/mol/nm")


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
def cpu_arm(vv, vo, args, steps, warmup, seconds_budget):
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
    n_ip = int(min(args.ion_pairs, max(1024, rate * per_step_budget / 37)))
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
            "sample": (f"first {n_ip} of {args.ion_pairs} ion pairs ({spec.n} particles), {steps} steps after "
                       f"{max(1, warmup)} warm-up, "
                       + ("reference kernel sources compiled for the host (oracle/_ref, OpenMP over CUDA blocks)"
                          if use_ref else "plain-C oracle port (oracle/vv_oracle.c, OpenMP)"))}
    return value, 1e3 * dt / steps, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vv = entry.load_package()
    vo = entry.load_oracle()
    value, ms, info = cpu_arm(vv, vo, args, args.steps, args.warmup, seconds_budget=args.ref_seconds)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, args.gpus), "precision": args.precision,
                       "l2": "sample larger than the host caches"},
            "cpu_baseline": info,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def pinned_state(vv, host):
    """copy a HostState into page-locked memory (torch is the allocator)"""
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
                    "blocking host round trip for the same step (profiles/configs_r01.json)"}


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
    spec = vv.make_bulk_ionic_liquid(args.ion_pairs)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, args.precision, seed=12345 + 100 * rank, force_sigma=FORCE_SIGMA)
    # this rank holds one whole-molecule partition of a box `world` times larger: DistributedPlan all-reduces the
    # thermostat DOFs and the total mass of the whole box at set-up (world == 1: the plan's own)
    dplan = vv.DistributedPlan(spec, params, args.precision).upload(peer={"auto": None, "nccl": False, "peer": True}[args.exchange])
    plan = dplan.plan
    bufs = vv.DeviceBuffers(host)
    n_local = spec.n
    stream = torch.cuda.current_stream()

    def one_step():
        if world == 1:
            plan.step_middle(bufs)                     # pass A (NH chains in its last block) + pass B
        else:
            dplan.step_middle(bufs)                    # pass A, all-reduce of <= 10 doubles over NVLink, NHC + pass B

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

    # ---- roofline of the dominant kernel (this rank) ----------------------------------------------
    peak, peak_src = peak_hbm()
    ka = {"kernel": "kick_reduce_kernel (pass A)", "bytes_per_particle": BYTES_PASS_A,
          "ms": ms_a / max(prof_steps, 1)}
    kb = {"kernel": "scale_drift_kernel (pass B)", "bytes_per_particle": BYTES_PASS_B,
          "ms": ms_b / max(prof_steps, 1)}
    for k in (ka, kb):
        k["achieved_gbs"] = k["bytes_per_particle"] * n_local / (k["ms"] * 1e-3) / 1e9 if k["ms"] > 0 else None
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
    e2e = None
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

    # ---- the reference's own CUDA kernels on this GPU (rank 0, single-GPU runs; reported, not the headline) ------
    ref_gpu = None
    if rank == 0 and world == 1 and not args.no_ref_gpu:
        vo = entry.load_oracle()
        if vo.ref_available(args.precision, gpu=True):
            oracle = vo.Oracle(spec, params, args.precision, literal=False)     # supplies the index arrays only
            ref = vo.Reference(oracle, gpu=True)
            rb = vv.DeviceBuffers(host)
            ref.step(rb, steps=2)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ksteps = 5
            r0.record(stream)
            nl = ref.step(rb, steps=ksteps)
            r1.record(stream)
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / ksteps
            ref_gpu = {"what": "reference kernels (platforms/cuda/src/kernels/*.cu, unmodified) compiled for sm_100a, "
                               "OpenMM launch geometry, blocking D2H/H2D around the host NH chain (oracle/_ref)",
                       "ms_per_step": rms, "value": n_local / (rms * 1e-3), "unit": UNIT,
                       "launches_per_step": nl / ksteps, "speedup_device_resident": rms / ms_per_step}
            del ref, rb, oracle

    # ---- BASELINE configs[1] (run-bulk.py-sized Drude bulk, ~50k particles): latency-bound, so microseconds and
    #      launches per step instead of a roofline fraction; rank 0, single-GPU runs ---------------
    config2 = None
    if rank == 0 and world == 1 and not args.no_config2:
        config2 = small_system_leg(vv, torch, args.precision)

    # ---- the >= 1M-particle target of BASELINE.json at other sizes (rank 0, single-GPU runs): whole-step time and
    #      the fraction of the HBM peak the 216 algorithmic bytes per particle amount to -------------------------
    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = []
        for n_ip in (27648, 110592):
            r = small_system_leg(vv, torch, args.precision, n_ip=n_ip, steps=200, graph=False)
            gbs = (BYTES_PASS_A + BYTES_PASS_B) * r["particles"] / (r["us_per_step"] * 1e-6) / 1e9
            sweep.append({"particles": r["particles"], "us_per_step": r["us_per_step"], "launches_per_step": r["launches_per_step"],
                          "step_gbs": gbs, "step_frac": gbs / peak})

    # ---- CPU baseline (rank 0, single-GPU runs only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        vo = entry.load_oracle()                 # bench.py's cpu_baseline leg: the one place the product bench runs oracle/
        _, _, cpu = cpu_arm(vv, vo, args, steps=5, warmup=1, seconds_budget=args.cpu_seconds)

    if rank == 0:
        st = plan.thermostat_state()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args, world), "precision": args.precision,
                           "particles_per_gpu": n_local, "particles_total": n_global,
                           "l2": "inputs larger than L2 (1.4 GB of state per GPU vs 126 MB)",
                           "parallelism": f"molecule-partitioned x{world}" if world > 1 else "single GPU",
                           "exchange": ("NVLink peer memory, fused into the NH-chain kernel" if dplan.peer else
                                        "NCCL all-reduce of 10 doubles") if world > 1 else "none"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "reference_kernels_on_gpu": ref_gpu,
                "config2_small_system": config2, "size_sweep": sweep,
                "clocks": clocks.summary(),
                "integrator_only_ns_per_day": 86400.0 / (ms_per_step * 1e-3) * params.step_size * 1e-3,
                "thermostat": {"ke2": [float(x) for x in st["ke2"]], "vscale": [float(x) for x in st["vscale"]]}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
