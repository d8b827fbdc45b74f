"""The C-ABI library loads on a machine without a GPU and exports every function include/vvb200.h declares
(no compute calls here).  Device entry points must fail loudly -- never fall back -- when no CUDA device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "vvb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(vvb200_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_every_declared_symbol_is_exported(vv):
    lib = C.CDLL(vv.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_device_code(vv):
    lib = vv.load_library()
    assert lib.vvb200_version() == 100
    assert lib.vvb200_has_device_code() == 1


def test_library_carries_sm100a_tma_kernels(vv):
    """SASS evidence without a GPU: the fat binary holds sm_100a code with bulk-copy (UBLKCP) instructions"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", vv.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_Z18kick_reduce_kernelILi1ELi1ELb0EEv7KParams", vv.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def test_no_cpu_fallback(vv):
    """without a CUDA device the device entry points return an error code; with one, stepping before upload does"""
    import torch
    spec = vv.make_bulk_ionic_liquid(2)
    plan = vv.Plan(spec, vv.Params().resolved_for(spec))
    host = vv.make_state(spec, "mixed")
    if not torch.cuda.is_available():
        with pytest.raises(vv.VVB200Error) as e:
            plan.lib.vvb200_plan_upload.restype = C.c_int
            vv._cabi._check(plan.lib, plan.lib.vvb200_plan_upload(plan.h, None))
        assert e.value.code == 4            # VVB200_ERR_CUDA
        with pytest.raises(RuntimeError):
            vv.DeviceBuffers(host)
    with pytest.raises(vv.VVB200Error) as e:
        plan.step_host(host, steps=1, stream=0)
    assert e.value.code in (4, 5)           # not uploaded / CUDA error -- never a silent CPU path


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "openmm-velocityverlet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "vvoracle" not in text and "vv_oracle" not in text and "oracle/" not in text, f


def test_every_script_compiles():
    """scripts that only ever run on the GPU box (multi-rank checks, diagnostics, tools, the benches) must at least be
    valid Python here: a syntax error in tests/multigpu_check.py would otherwise first show up on an 8-GPU box"""
    import glob
    import py_compile
    files = [f for pat in ("tests/*.py", "tools/*.py", "*.py", "openmm-velocityverlet_b200/*.py", "oracle/*.py", "profiles/*.py")
             for f in glob.glob(os.path.join(ROOT, pat))]
    assert len(files) > 30
    for f in files:
        py_compile.compile(f, doraise=True)
