"""Diagnostic (not collected): where the time goes in the two streaming passes of one step.  Needs a -DVVB200_TRACE build:
    make -C openmm-velocityverlet_b200/csrc EXTRA_NVFLAGS=-DVVB200_TRACE OUT=/root/repo/gpurun_out/libvvb200_trace.so
    VVB200_LIB=gpurun_out/libvvb200_trace.so python tests/diag_trace_streaming.py [ion pairs ...]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import numpy as np, torch
vv = entry.load_package()
lib = vv.load_library()
for n_ip in [int(a) for a in sys.argv[1:]] or [27648, 110592]:
    spec = vv.make_bulk_ionic_liquid(n_ip)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=1.0)
    plan = vv.Plan(spec, params, "mixed").upload()
    b = vv.DeviceBuffers(host)
    for _ in range(20): plan.step_middle(b)
    torch.cuda.synchronize()
    t = np.zeros(2048 * 8, dtype=np.uint64)
    lib.vvb200_debug_trace_streaming(t.ctypes.data_as(C.c_void_p), t.size)
    t = t.reshape(2, 1024, 8).astype(np.int64)
    A, B = t[0][t[0][:, 0] > 0], t[1][t[1][:, 0] > 0]
    base = A[:, 0].min()
    print(f"N={spec.n}: pass A blocks {len(A)}, pass B blocks {len(B)}   (us since the first pass-A block's entry; min / median / max)")
    for name, arr, col in (("A entry", A, 0), ("A past griddepcontrol.wait", A, 1), ("A first tile in smem", A, 2), ("A tiles done", A, 3),
                           ("A ticket taken", A, 4), ("B entry", B, 0), ("B producer saw velocities", B, 1), ("B consumers released", B, 2),
                           ("B first tile in smem", B, 3), ("B tiles done", B, 4)):
        v = (arr[:, col] - base) / 1e3
        v = v[arr[:, col] > 0]
        if v.size:
            print(f"  {name:28s} {v.min():8.2f} {np.median(v):8.2f} {v.max():8.2f}")
    last = A[np.argmax(A[:, 5])]      # rows persist across steps: the block that published most recently
    r = lambda c: (last[c] - base) / 1e3
    print(f"  last block of A: tiles done {r(3) if False else 0:.0f} ticket {r(4):.2f} | fence + word=1 {r(2):.2f} | partials summed {r(3):.2f} | block sums {r(7):.2f} | "
          f"factors published {r(5):.2f} | state stored {r(6):.2f}")
    # next step's pass A relative to this one cannot be seen here (rows are overwritten each step): the step period is the bench's number
