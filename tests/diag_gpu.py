"""Diagnostic (not collected by pytest): prints GPU-vs-oracle error figures for the parity cases.
usage: python tests/diag_gpu.py [libpath]"""
import dataclasses
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry  # noqa: E402
from conftest import rel_err  # noqa: E402

vv = entry.load_package()
vo = entry.load_oracle()
if len(sys.argv) > 1:
    import vvb200._cabi as cabi
    cabi._lib = cabi.load_library(sys.argv[1])
    print("using", sys.argv[1])


def worst(a, b, spec, label):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = np.maximum(np.abs(b), 1e-3 * np.sqrt(np.mean(b * b)) + 1e-300)
    e = np.abs(a - b) / scale
    i = np.unravel_index(np.argmax(e), e.shape)
    p = int(i[0])
    role = "none"
    if spec.drude_pairs.size:
        if p in spec.drude_pairs[:, 0]: role = "drude"
        elif p in spec.drude_pairs[:, 1]: role = "parent"
    nbad = int(np.sum(e.max(axis=1) > 1e-7))
    print(f"   {label}: max rel {e[i]:.3e} at {i} role={role} mass={spec.masses[p]} got={a[i]:.9g} want={b[i]:.9g}  (#>1e-7: {nbad})")


def case(name, spec, params, precision, steps, inv_box_z=0.0, n_random=0, **kw):
    host = vv.make_state(spec, precision, n_random=n_random, **kw)
    plan = vv.Plan(spec, params, precision).upload()
    bufs = vv.DeviceBuffers(host)
    oracle = vo.Oracle(spec, params, precision, literal=False)
    want = host.copy()
    print(f"== {name} [{precision}] N={spec.n}")
    for s in range(steps):
        plan.step(bufs, steps=1, inv_box_z=inv_box_z, random_index=s * plan.random_request)
        oracle.step(want, steps=1, inv_box_z=inv_box_z)
        got = bufs.to_host()
        n = spec.n
        print(f" step {s}")
        worst(got.velm[:n, :3], want.velm[:n, :3], spec, "v")
        worst(got.positions()[:n], want.positions()[:n], spec, "x")
        a, b = plan.thermostat_state(), oracle.thermostat_state()
        ng = b["num_temp_groups"]
        print(f"   ke2 rel {rel_err(a['ke2'][:ng], b['ke2']):.3e} vscale rel {rel_err(a['vscale'][:ng], b['vscale']):.3e}"
              f" etadot rel {rel_err(a['eta_dot'], b['eta_dot']):.3e}")


P = vv.Params
bulk = vv.make_bulk_ionic_liquid(250)
for prec in ("single", "double", "mixed"):
    for hw in (0.0, 0.02):
        case(f"bulk hw={hw}", bulk, P(max_drude_distance=hw).resolved_for(bulk), prec, 3)
small = vv.make_bulk_ionic_liquid(100)
hb = vv.make_state(small, "single")
for middle in (True, False):
    case(f"cos middle={middle}", small,
         dataclasses.replace(P(max_drude_distance=0.02, cos_acceleration=0.02).resolved_for(small), use_middle_scheme=middle),
         "single", 3, inv_box_z=1.0 / hb.box[2])
case("vv", small, dataclasses.replace(P(max_drude_distance=0.02).resolved_for(small), use_middle_scheme=False), "single", 3)
rag = vv.make_ragged(seed=2, scattered_molecules=0)
case("ragged2", rag, P(max_drude_distance=0.02, mirror_location=1.0, electric_field=1e-22).resolved_for(rag), "mixed", 3,
     n_random=8 * rag.n, mirror=1.0)
