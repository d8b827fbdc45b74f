import os, sys
sys.path.insert(0, os.getcwd())
import __graft_entry__ as entry
vv = entry.load_package()
import torch
spec = vv.make_bulk_ionic_liquid(442368)
params = vv.Params(max_drude_distance=0.02, use_middle_scheme=False).resolved_for(spec)
host = vv.make_state(spec, "mixed", force_sigma=1.0)
plan = vv.Plan(spec, params, "mixed").upload()
bufs = vv.DeviceBuffers(host)
for _ in range(4):
    plan.step_vv_first(bufs); plan.step_vv_second(bufs)
torch.cuda.synchronize()
