#!/usr/bin/env python
"""Host<->device copy bandwidth per rank, alone and with all ranks copying at once (why the host-buffer step of
vvb200_step_host scales the way it does on N GPUs).   torchrun --nproc-per-node N tools/pcie_probe.py
Prints one JSON line from rank 0: per-rank GB/s for H2D, D2H and both directions at once, (a) one rank at a time,
(b) all ranks concurrently; plus the CPU affinity / NUMA node the process and its pinned buffers live on."""
import json, os, subprocess, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
BYTES = 1 << 30
h_in = torch.empty(BYTES, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
d_a = torch.empty(BYTES, dtype=torch.uint8, device="cuda"); d_b = torch.ones(BYTES, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def run(kind, reps=4):
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return BYTES * reps * (2 if kind == "both" else 1) / dt / 1e9

res = {"alone": {}, "concurrent": {}}
for kind in ("h2d", "d2h", "both"):
    run(kind, 1)
    # one rank at a time
    mine = 0.0
    for r in range(world):
        if r == rank: mine = run_alone = None
        barrier()
        if r == rank:
            t0 = time.perf_counter()
            for _ in range(4):
                if kind in ("h2d", "both"):
                    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
                if kind in ("d2h", "both"):
                    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
            torch.cuda.synchronize()
            mine = BYTES * 4 * (2 if kind == "both" else 1) / (time.perf_counter() - t0) / 1e9
        barrier()
    conc = run(kind)
    vals = torch.tensor([mine, conc], dtype=torch.float64, device="cuda")
    if world > 1:
        g = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(g, vals)
    else:
        g = [vals]
    res["alone"][kind] = [round(float(v[0]), 1) for v in g]
    res["concurrent"][kind] = [round(float(v[1]), 1) for v in g]

def sh(cmd):
    try: return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e: return f"failed: {e}"
if rank == 0:
    res["world"] = world
    res["concurrent_sum"] = {k: round(sum(v), 1) for k, v in res["concurrent"].items()}
    res["cpu_affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
    res["numa_nodes"] = sh("ls -d /sys/devices/system/node/node* 2>/dev/null | wc -l")
    res["lscpu"] = sh("lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)'")
    res["gpu_numa"] = sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/class 2>/dev/null)\" = 0x030200 ]; then echo $(basename $d):$(cat $d/numa_node); fi; done | tr '\\n' ' '")
    res["pcie_link"] = sh("nvidia-smi --query-gpu=index,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader | tr '\\n' ';'")
    res["mem"] = sh("grep -E 'MemTotal|MemAvailable' /proc/meminfo | tr '\\n' ' '")
    print(json.dumps(res), flush=True)
if world > 1: dist.destroy_process_group()
