"""Molecule-partitioned multi-GPU stepping: one process per GPU, each rank owns a contiguous
range of WHOLE molecules (so Drude pairs, constraints and molecular centres of mass never cross ranks) with its
slice of posq / velm / force.  The only exchange per step is one all-reduce (sum, fp64) of the <= 10-element
reduction vector between the two passes -- done by the last block of pass A itself over NVLink peer memory, or by one
NCCL all-reduce where the ranks cannot map each other's memory; the Nose-Hoover chains are then advanced redundantly
on every rank (deterministic fp64 => identical scale factors everywhere).

torch.distributed is only the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import dataclasses

import numpy as np

from ._cabi import Plan


def partition_by_molecules(spec, world):
    """[(first, last)) particle ranges: contiguous, whole molecules, balanced by particle count.
    Requires molecule ids to be non-decreasing in particle order (true for OpenMM systems built molecule by
    molecule; images bonded to far-away parents break it and must be co-located first)."""
    mol = spec.mol_id
    if np.any(np.diff(mol) < 0):
        raise ValueError("molecules are not contiguous in particle order: cannot partition by particle ranges")
    starts = np.flatnonzero(np.diff(mol, prepend=-1) != 0)           # first particle of each molecule
    bounds = [0]
    for r in range(1, world):
        target = spec.n * r / world
        k = int(np.searchsorted(starts, target))
        cands = [starts[min(k, len(starts) - 1)]]
        if k > 0:
            cands.append(starts[k - 1])
        cut = int(min(cands, key=lambda c: abs(c - target)))
        bounds.append(max(cut, bounds[-1]))
    bounds.append(spec.n)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def global_thermostat(local_plan, has_cmm, use_com, group=None, device="cpu"):
    """Whole-box DOFs and total mass from per-rank plans that were created WITHOUT a CMMotionRemover: sum the raw
    per-rank DOFs, then apply the remover's -3 once and clamp, exactly where the reference does
    (CudaVVKernels.cpp:550-564)."""
    import torch
    import torch.distributed as dist
    dof = local_plan.f64_array("dof")
    mass = 1.0 / local_plan.f64_array("invMassTotal")[0]
    t = torch.tensor([dof[0], dof[1], dof[2], mass], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, group=group)
    d = t[:3].cpu().numpy().copy()
    if has_cmm:
        d[1 if use_com else 0] -= 3
    d = np.maximum(d, 0.0)
    return d, float(t[3].item())


class DistributedPlan:
    """A rank's plan over its whole-molecule partition plus the per-step exchange."""

    def __init__(self, local_spec, params, precision="mixed", group=None, device=None):
        import torch
        self.group = group
        self.has_cmm = bool(local_spec.has_cmm)
        spec = dataclasses.replace(local_spec, has_cmm=False) if local_spec.has_cmm else local_spec
        self.plan = Plan(spec, params, precision)
        self.device = device or ("cuda" if torch.cuda.is_available() else "cpu")
        self.dof, self.total_mass = global_thermostat(self.plan, self.has_cmm, bool(params.use_com_temp_group),
                                                      group, self.device)
        self.plan.set_global_thermostat(self.dof, self.total_mass)
        self._red = None

    def upload(self, stream=None, peer=None):
        """peer=None: use the NVLink peer-memory exchange when the ranks can map each other's buffers (one node, cudaIpc),
        else NCCL; peer=False forces NCCL; peer=True raises if peer mapping fails."""
        import torch
        import torch.distributed as dist
        self.plan.upload(stream)
        self.peer = False
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world > 1 and peer is not False and world <= 8:
            ok = 1
            try:
                mine = torch.from_numpy(self.plan.peer_export()).cuda()
                gathered = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(gathered, mine, group=self.group)
                handles = torch.stack(gathered).cpu().numpy()
                self.plan.peer_attach(dist.get_rank(self.group), world, handles)
            except Exception:
                if peer:
                    raise
                ok = 0
            flag = torch.tensor([ok], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)      # all ranks or none
            self.peer = bool(flag.item())
            if not self.peer and ok:
                self.plan.peer_attach(0, 1, self.plan.peer_export())           # detach: back to the NCCL path
        ptr, cnt = self.plan.partials()

        class _DevPtr:
            __cuda_array_interface__ = {"shape": (cnt,), "typestr": "<f8", "data": (int(ptr), False), "version": 3,
                                        "strides": None}
        self._red = torch.as_tensor(_DevPtr(), device="cuda")      # zero-copy view of the plan's reduction vector
        return self

    def step_middle(self, bufs, **kw):
        import torch.distributed as dist
        if self.peer:
            # the LAST BLOCK OF PASS A exchanges the sums over cudaIpc-mapped NVLink memory and advances the chains; pass B
            # takes over through the hand-over word -- the same two launches per step, the same call, as on one GPU
            self.plan.step_middle(bufs, **kw)
            return
        self.plan.middle_kick_reduce(bufs, **kw)
        if dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self._red, group=self.group)           # NCCL: the only exchange, <= 10 doubles
        self.plan.middle_nhc_scale_drift(bufs, **kw)

    def step_host(self, host_state, **kw):
        """one step through HOST buffers (pinned for full speed): copy-in, pass A, exchange, pass B and copy-out overlapped
        chunk by chunk (vvb200_step_host_begin / _finish)"""
        import torch.distributed as dist
        self.plan.step_host_begin(host_state, **kw)
        if not self.peer and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self._red, group=self.group)
        self.plan.step_host_finish(host_state, **kw)
