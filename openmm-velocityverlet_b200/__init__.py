"""vvb200 -- Python harness over libvvb200.so (the C ABI in include/vvb200.h).

The product is the shared library: hand-written sm_100a CUDA behind a C ABI that the OpenMM
glue (csrc/glue) forwards to.  This package is only plumbing for tests and
bench.py: ctypes bindings, synthetic systems in OpenMM's array layouts, torch tensors as
device memory.  There is no CPU fallback: a missing library or missing CUDA device raises.

The directory name contains a hyphen, so import it through ``__graft_entry__.load_package()``
(which registers it as module ``vvb200``).
"""
from ._cabi import (  # noqa: F401
    LIB_PATH,
    VVB200Error,
    Params,
    Plan,
    PRECISIONS,
    find_molecules,
    load_library,
    propagate_nh_chain,
)
from .system import (  # noqa: F401
    SystemSpec,
    HostState,
    make_bulk_ionic_liquid,
    make_edl,
    make_nonpolar_box,
    make_polymer,
    make_ragged,
    make_state,
    np_dtypes,
)
from .buffers import DeviceBuffers  # noqa: F401
from .integrator import Context, OpenMMException, System, VVIntegrator  # noqa: F401
from .multigpu import DistributedPlan, global_thermostat, partition_by_molecules  # noqa: F401
