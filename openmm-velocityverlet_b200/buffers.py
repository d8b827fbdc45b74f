"""Device-side copies of OpenMM's arrays as torch CUDA tensors (torch is only the allocator)."""
import ctypes as C

import numpy as np

from ._cabi import _Buffers
from .system import HostState


class DeviceBuffers:
    """posq / posqCorrection / velm / force / posDelta / random on the current CUDA device, in
    OpenMM's layouts. The library works on them in place, exactly as it does on CudaContext's."""

    def __init__(self, host: HostState, device="cuda", with_pos_delta=False):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("vvb200 needs a CUDA device (no CPU fallback)")
        self.precision = host.precision
        self.box = host.box
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.posq = t(host.posq)
        self.corr = t(host.corr) if host.corr is not None else None
        self.velm = t(host.velm)
        self.force = t(host.force)
        self.random = t(host.random)
        self.pos_delta = torch.zeros_like(self.velm) if with_pos_delta else None

    @classmethod
    def from_tensors(cls, precision, posq, corr, velm, force, random=None, box=(1.0, 1.0, 1.0), pos_delta=None):
        """wrap CUDA tensors that already hold OpenMM's layouts (e.g. a state generated on the device)"""
        import torch
        self = cls.__new__(cls)
        self.precision, self.box = precision, box
        self.posq, self.corr, self.velm, self.force = posq, corr, velm, force
        self.random = random if random is not None else torch.zeros((1, 4), dtype=torch.float32, device=velm.device)
        self.pos_delta = pos_delta
        return self

    def c_struct(self):
        p = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        return _Buffers(p(self.posq), p(self.corr), p(self.velm), p(self.force), p(self.pos_delta), p(self.random))

    def load(self, host: HostState):
        """overwrite device contents from a host state (shapes must match)"""
        import torch
        self.posq.copy_(torch.from_numpy(host.posq))
        if self.corr is not None:
            self.corr.copy_(torch.from_numpy(host.corr))
        self.velm.copy_(torch.from_numpy(host.velm))
        self.force.copy_(torch.from_numpy(host.force))

    def to_host(self) -> HostState:
        import torch
        torch.cuda.synchronize()
        c = lambda x: x.cpu().numpy() if x is not None else None
        return HostState(self.precision, c(self.posq), c(self.corr), c(self.velm), c(self.force), c(self.random), self.box)
