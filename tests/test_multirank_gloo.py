"""World-size-2 CPU test (gloo) of the molecule-partitioned multi-GPU path's host logic: partitioning by whole
molecules, whole-box thermostat DOFs from per-rank builders, and the per-step exchange -- each rank's partial
group sums all-reduced equal the single-process sums, and the redundantly propagated NH chains agree bitwise.
(The device kernels of the same path are covered on GPUs by tests/multigpu_check.py under `gpurun --gpus 2`.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    vv, vo = entry.load_package(), entry.load_oracle()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = vv.make_bulk_ionic_liquid(41, hbond_constraints=True, has_cmm=True)
        params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
        parts = vv.partition_by_molecules(spec, world)
        a, b = parts[rank]
        local = spec.subset_molecules(a, b)
        dp = vv.DistributedPlan(local, params, "mixed", device="cpu")

        # 1. whole-box DOFs / masses equal the single-process builders'
        full = vv.Plan(spec, params, "mixed")
        dof_full = full.f64_array("dof")
        assert np.allclose(dp.dof, dof_full, rtol=1e-14, atol=0), (dp.dof, dof_full)
        assert abs(dp.total_mass - 1.0 / full.f64_array("invMassTotal")[0]) <= 1e-12 * dp.total_mass
        assert dp.plan.num_temp_groups == full.num_temp_groups == 3
        assert np.array_equal(dp.plan.f64_array("dofGlobal"), dp.dof)

        # 2. the exchange: partial group sums of each partition, all-reduced, equal the full-system sums
        host = vv.make_state(spec, "mixed", force_sigma=0.0)
        o_full = vo.Oracle(spec, params, "mixed", literal=False)
        ref_state = host.copy()
        o_full.scale_velocity(ref_state)
        ke2_full = o_full.thermostat_state()["ke2"]

        P = local.padded_n
        def cut(x, rows=True):
            out_ = np.zeros((P,) + x.shape[1:], x.dtype)
            out_[: b - a] = x[a:b]
            return out_
        lf = np.zeros((3, P), np.int64)
        lf[:, : b - a] = host.force[:, a:b]
        lstate = vv.HostState("mixed", cut(host.posq), cut(host.corr), cut(host.velm), lf, host.random, host.box)
        o_loc = vo.Oracle(local, params, "mixed", literal=False)
        o_loc.scale_velocity(lstate.copy())
        part = torch.tensor(o_loc.thermostat_state()["ke2"], dtype=torch.float64)
        dist.all_reduce(part)
        assert np.allclose(part.numpy(), ke2_full, rtol=1e-13, atol=0)

        # 3. NH chains propagated redundantly from the reduced sums are bitwise identical on every rank
        nc = params.num_nh_chains
        kT = 1.380649e-23 * 6.02214076e23 / 1000.0 * params.temperature
        q = np.array([dp.dof[0] * kT / params.frequency ** 2] + [kT / params.frequency ** 2] * (nc - 1))
        eta, eta_dot, eta_dd = np.zeros(nc), np.zeros(nc + 1), np.zeros(nc)
        f = vv.propagate_nh_chain(params.step_size, 1, eta, eta_dot, eta_dd, q, float(part[0]), dp.dof[0] * kT,
                                  params.temperature)
        mine = torch.tensor([f] + list(eta_dot), dtype=torch.float64)
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        for g in gathered:
            assert torch.equal(g, gathered[0])
        out.put((rank, "ok", parts))
    except Exception as e:  # noqa: BLE001
        import traceback
        out.put((rank, "fail: " + traceback.format_exc(), None))
    finally:
        dist.destroy_process_group()


def test_two_rank_partition_and_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, parts in results:
        assert status == "ok", f"rank {rank}: {status}"
    parts = results[0][2]
    assert parts[0][0] == 0 and parts[0][1] == parts[1][0] and parts[1][1] == 41 * 37


def test_partition_properties(vv):
    spec = vv.make_bulk_ionic_liquid(1000)
    for world in (2, 4, 8):
        parts = vv.partition_by_molecules(spec, world)
        assert parts[0][0] == 0 and parts[-1][1] == spec.n
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 2 * 27                     # balanced to within a molecule or two
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c
            assert spec.mol_id[b - 1] != spec.mol_id[b]               # cuts fall between molecules
        for a, b in parts:
            sub = spec.subset_molecules(a, b)
            assert sub.n == b - a and sub.drude_pairs.min() >= 0 and sub.drude_pairs.max() < sub.n
    edl = vv.make_edl(8, 60, 3)
    with pytest.raises(ValueError):
        vv.partition_by_molecules(edl, 2)                               # images bonded to far parents
