#!/usr/bin/env python
"""ptxas -v summary: registers / spills / smem per kernel of vvb200_device.cu (cross-compiled, no GPU needed).
    python tools/ptxas_report.py [out.txt]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "openmm-velocityverlet_b200", "csrc")
cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
       f"-I{ROOT}/include", f"-I{src}", "-Xptxas", "-v", "-cubin", "-o", "/tmp/vvb200_report.cubin",
       os.path.join(src, "vvb200_device.cu")]
log = subprocess.run(cmd, capture_output=True, text=True).stderr
rows, name = [], None
for line in log.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*\)$", "", name).replace("void ", "")
        spill = None
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        spill = (int(m.group(1)), int(m.group(2)), int(m.group(3)))
    m = re.search(r"Used (\d+) registers", line)
    if m and name:
        rows.append((name, int(m.group(1)), spill))
out = "\n".join(f"{n:70s} regs {r:3d}  stack/spill {s}" for n, r, s in sorted(rows))
print(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(out + "\n")
