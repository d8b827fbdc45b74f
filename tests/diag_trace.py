"""Diagnostic (not collected): where the time goes inside the resident kernel.  Needs a -DVVB200_TRACE build:
    make -C openmm-velocityverlet_b200/csrc EXTRA_NVFLAGS=-DVVB200_TRACE OUT=/root/repo/gpurun_out/libvvb200_trace.so
    VVB200_LIB=gpurun_out/libvvb200_trace.so python tests/diag_trace.py"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np, torch
vv = entry.load_package()
lib = vv.load_library()
for n_ip in (250, 1250):
    spec = vv.make_bulk_ionic_liquid(n_ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    plan = vv.Plan(spec, params, "mixed").upload()
    b = vv.DeviceBuffers(host)
    for _ in range(20): plan.step_middle(b)
    torch.cuda.synchronize()
    nb = min(1024, (spec.n + 400) // 400)
    t = np.zeros(1024 * 8, dtype=np.uint64)
    lib.vvb200_debug_trace(t.ctypes.data_as(C.c_void_p), t.size)
    t = t.reshape(1024, 8).astype(np.int64)
    used = t[:, 0] > 0
    tt = t[used]
    base = tt[:, 0].min()
    rel = (tt[:, :7] - base) / 1e3
    last = int(np.argmax(tt[:, 7]))
    print(f"  last block: ticket {rel[last, 4]:.2f}  sums done {(tt[last, 7] - base) / 1e3:.2f}  released {rel[last, 5]:.2f}")
    names = ["entry", "loads issued", "data arrived", "pass A done", "ticket taken", "barrier passed", "pass B done"]
    print(f"N={spec.n} blocks={used.sum()}  (us since the first block's entry; min / median / max over blocks)")
    for i, nm in enumerate(names):
        print(f"  {nm:15s} {rel[:, i].min():7.2f} {np.median(rel[:, i]):7.2f} {rel[:, i].max():7.2f}")
