"""N-rank == 1-rank on real GPUs (driver-visible): spawns tests/multigpu_check.py under torch.distributed.run on every
GPU count the box offers (2, 4, 8), for both exchanges -- the NVLink peer-memory exchange inside pass A's last block and
the NCCL all-reduce -- and keeps the log (gpurun_out/multigpu_check_<N>gpu_<exchange>.log; copies are committed under
profiles/).  Each run checks, for a Drude bulk box with a CMMotionRemover, a cosine-perturbation run and a
non-polarizable box: positions / velocities of rank 0's partition against the fused single-GPU step of the WHOLE box
(<= 1e-9 relative), group energies and scale factors (<= 1e-12), scale factors bitwise identical on every rank, and the
host-buffer pipeline against the device-resident steps."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_equals_one_rank(world, exchange, step_path):
    if step_path != "streaming":
        pytest.skip("multi-GPU runs never take the single-launch resident kernel: one parametrisation is enough")
    if _gpus() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_gpus()}")
    env = dict(os.environ, VVB200_EXCHANGE=exchange, OMP_NUM_THREADS="1")
    env.pop("VVB200_RESIDENT", None)
    port = 29500 + 7 * world + (3 if exchange == "peer" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, f"multigpu_check_{world}gpu_{exchange}.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    print(r.stdout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multigpu_check: all OK" in r.stdout
