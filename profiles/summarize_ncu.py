#!/usr/bin/env python
"""Reads an `ncu --set full` report (here, no GPU needed) and writes the summary the roofline numbers cite:
    python profiles/summarize_ncu.py gpurun_out/prof_r01.ncu-rep profiles/ncu_r01_summary.txt [profiles/traffic.json]
Per captured launch: duration, DRAM bytes read/written (the `traffic` of bench.py's roofline), DRAM / issue / pipe
utilisation, occupancy limits, the warp-stall breakdown and the SASS opcode mix (from the source page)."""
import collections
import csv
import io
import json
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
       "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
       "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct"]


def ncu(*args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "raw", "--csv"))))
    hdr, units = rows[0], rows[1]
    lines, traffic = [], {"source": rep}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"=== {name}   grid {r[hdr.index('Grid Size')]} block {r[hdr.index('Block Size')]}")
        vals = {}
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                vals[m] = (r[i], units[i])
                lines.append(f"  {m:70s} {r[i]:>16s} {units[i]}")
        if "dram__bytes_read.sum" in vals:
            rd = to_bytes(*vals["dram__bytes_read.sum"])
            wr = to_bytes(*vals["dram__bytes_write.sum"])
            dur = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
            dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(vals["gpu__time_duration.sum"][1], 1e-6)
            lines.append(f"  DRAM traffic per launch: {rd + wr:.4e} B (read {rd:.4e}, write {wr:.4e}) -> {(rd + wr) / dur_s / 1e9:.0f} GB/s under ncu (cold, serialised)")
            key = name.split("<")[0].replace("void ", "").strip()
            traffic.setdefault(key, rd + wr)
        stall = {h.split("_per_issue_active")[0].replace("smsp__average_warps_issue_stalled_", ""): float(r[i] or 0)
                 for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
        top = sorted(stall.items(), key=lambda kv: -kv[1])[:6]
        lines.append("  warp stalls per issue: " + ", ".join(f"{k} {v:.2f}" for k, v in top))
        lines.append("")
    # opcode mix per kernel from the source page
    src = list(csv.reader(io.StringIO(ncu("-i", rep, "--page", "source", "--csv"))))
    kern, hdr2, seen = None, None, set()
    mix = collections.defaultdict(collections.Counter)
    for r in src:
        if r and r[0] == "Kernel Name":
            kern = r[1]
            seen = set()
            continue
        if r and r[0] == "Address":
            hdr2 = r
            continue
        if hdr2 is None or not r or not r[0].startswith("0x") or r[0] in seen:
            continue
        seen.add(r[0])
        toks = r[1].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        try:
            mix[kern][".".join(op.split(".")[:2])] += int(r[hdr2.index("Instructions Executed")])
        except ValueError:
            pass
    for k, c in mix.items():
        tot = sum(c.values())
        lines.append(f"=== SASS opcode mix (warp instructions, all captured launches of this kernel): {k}   total {tot}")
        lines.append("  " + ", ".join(f"{op} {100 * n / tot:.1f}%" for op, n in c.most_common(18)))
        lines.append("  Blackwell evidence: UBLKCP (cp.async.bulk / TMA) %d, SYNCS (mbarrier) %d"
                     % (sum(n for op, n in c.items() if op.startswith("UBLKCP")), sum(n for op, n in c.items() if op.startswith("SYNCS"))))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    if traffic_path:
        json.dump(traffic, open(traffic_path, "w"), indent=1)
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
