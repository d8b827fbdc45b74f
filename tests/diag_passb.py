"""Diagnostic (not collected): pass A / pass B times at 16.4M particles for the launch geometry in the environment
(VVB200_BLOCKS_A/B, VVB200_STAGES_A/B) and with the hard wall on / off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
vv = entry.load_package()
n_ip = int(os.environ.get("DIAG_IP", "442368"))
spec = vv.make_bulk_ionic_liquid(n_ip)
host = vv.make_state(spec, "mixed", force_sigma=1.0)
out = []
for hw in (0.02, 0.0):
    params = vv.Params(max_drude_distance=hw).resolved_for(spec)
    plan = vv.Plan(spec, params, "mixed").upload()
    b = vv.DeviceBuffers(host)
    for _ in range(5): plan.step_middle(b)
    torch.cuda.synchronize()
    plan.profile_enable(20)
    for _ in range(20): plan.step_middle(b)
    a_ms, b_ms, n = plan.profile_read()
    out.append(f"hardwall {hw}: A {1e3*a_ms/n:6.1f} us  B {1e3*b_ms/n:6.1f} us")
    del plan, b
env = {k: v for k, v in os.environ.items() if k.startswith("VVB200_")}
print(env, " | ".join(out), flush=True)
