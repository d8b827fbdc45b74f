"""10^4-step run (BASELINE.json north_star: "over 10^4 steps, group temperatures and NH-chain energy conservation
must match within stated tolerances").  Frozen forces heat without bound, so both paths are driven by the same cheap
deterministic toy force field -- a harmonic tether on every massive non-Drude particle plus a Drude spring on every
pair, converted to OpenMM's 2^32 fixed point (test-only definition: vvo_toy_forces in oracle/vv_oracle.c; here the
same formula with torch ops on the GPU) -- recomputed every step, like OpenMM would.

Stated tolerances (mixed precision, middle scheme, 1,110-particle Drude ionic-liquid box, dt = 1 fs):
  * trajectory level: the two paths are the same arithmetic up to fp64 reassociation, and this system is only weakly
    chaotic, so after 10^4 steps positions/velocities still agree to 1e-6 relative (the north-star bar for ONE step);
  * group temperatures (atom / COM / Drude), sampled every 100 steps: block averages over the second half agree to
    1e-6 relative, and sit within 3 % (atom, COM) and 10 % (1 K Drude group) of their targets;
  * extended (Nose-Hoover) energy  H = KE + PE + sum_g [ sum_k Q_gk etadot_gk^2 / 2 + dof_g kT_g eta_g0 +
    kT_g sum_{k>=1} eta_gk ]  sampled every 100 steps: the two series agree to 1e-8 of |H|, i.e. whatever drift the
    reference arithmetic has, ours has the same."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu
BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0
K_TETHER, K_DRUDE = 5000.0, 4.184e5          # kJ/mol/nm^2


def extended_energy(spec, params, dof, x, v, x0, st):
    m = spec.masses
    ke = 0.5 * float(np.sum(m[:, None] * v * v))
    d, p = spec.drude_pairs[:, 0], spec.drude_pairs[:, 1]
    tether = (m > 0)
    tether[d] = False
    pe = 0.5 * K_TETHER * float(np.sum((x[tether] - x0[tether]) ** 2)) + 0.5 * K_DRUDE * float(np.sum((x[d] - x[p]) ** 2))
    h = ke + pe
    for g in range(st["num_temp_groups"]):
        T = params.drude_temperature if g == 2 else params.temperature
        f = params.drude_frequency if g == 2 else params.frequency
        kT = BOLTZ * T
        q = np.full(params.num_nh_chains, kT / f ** 2)
        q[0] *= dof[g]
        h += 0.5 * float(np.sum(q * st["eta_dot"][g][:-1] ** 2)) + dof[g] * kT * st["eta"][g][0] + kT * float(np.sum(st["eta"][g][1:]))
    return h, ke, pe


def test_ten_thousand_steps_match_the_oracle(vv, vo):
    import torch
    steps, every = 10000, 100
    spec = vv.make_bulk_ionic_liquid(30)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    host = vv.make_state(spec, "mixed", force_sigma=0.0, drude_spread=0.0005)
    n, P = spec.n, spec.padded_n
    x0 = host.positions()[:n].copy()

    # ---- GPU path ----
    plan = vv.Plan(spec, params, "mixed").upload()
    bufs = vv.DeviceBuffers(host)
    dev = bufs.posq.device
    d_idx = torch.as_tensor(spec.drude_pairs[:, 0].astype(np.int64), device=dev)
    p_idx = torch.as_tensor(spec.drude_pairs[:, 1].astype(np.int64), device=dev)
    tether = torch.as_tensor(spec.masses > 0, device=dev)
    tether[d_idx] = False
    x0_t = torch.as_tensor(x0, device=dev)

    def toy_forces_gpu():
        x = bufs.posq[:n, :3].double() + bufs.corr[:n, :3].double()
        f = torch.where(tether[:, None], -K_TETHER * (x - x0_t), torch.zeros_like(x))
        fi = (f * 4294967296.0).to(torch.int64)                       # truncation toward zero, like (long long)
        fd = (-K_DRUDE * (x[d_idx] - x[p_idx]) * 4294967296.0).to(torch.int64)
        fi.index_add_(0, d_idx, fd)
        fi.index_add_(0, p_idx, -fd)
        bufs.force[:, :n] = fi.t()

    oracle = vo.Oracle(spec, params, "mixed", literal=False)
    want = host.copy()
    dof = plan.f64_array("dof")
    series = {"gpu": [], "cpu": []}
    temps = {"gpu": [], "cpu": []}
    for s in range(steps):
        toy_forces_gpu()
        plan.step_middle(bufs)
        oracle.toy_forces(want, x0, K_TETHER, K_DRUDE)
        oracle.step(want, steps=1)
        if (s + 1) % every == 0:
            got = bufs.to_host()
            for key, state, st in (("gpu", got, plan.thermostat_state()), ("cpu", want, oracle.thermostat_state())):
                h, ke, pe = extended_energy(spec, params, dof, state.positions()[:n], state.velm[:n, :3].astype(np.float64), x0, st)
                series[key].append(h)
                temps[key].append(st["ke2"][:3] / (dof * BOLTZ))
    got = bufs.to_host()

    # trajectory level
    ev = rel_err(got.velm[:n, :3], want.velm[:n, :3])
    ex = rel_err(got.positions()[:n], want.positions()[:n])
    hg, hc = np.array(series["gpu"]), np.array(series["cpu"])
    tg, tc = np.array(temps["gpu"]), np.array(temps["cpu"])
    half = len(hg) // 2
    drift_g, drift_c = hg[-1] - hg[0], hc[-1] - hc[0]
    print(f"10^4 steps: v {ev:.2e} x {ex:.2e}; <T> gpu {tg[half:].mean(axis=0)} cpu {tc[half:].mean(axis=0)}; "
          f"H drift gpu {drift_g:.3e} cpu {drift_c:.3e} of |H| {abs(hc).mean():.3e}; max|Hg-Hc| {np.max(np.abs(hg - hc)):.3e}")
    assert max(ev, ex) <= 1e-6
    # group temperatures
    assert rel_err(tg[half:].mean(axis=0), tc[half:].mean(axis=0)) <= 1e-6
    target = np.array([params.temperature, params.temperature, params.drude_temperature])
    dev_rel = np.abs(tg[half:].mean(axis=0) - target) / target
    assert dev_rel[0] <= 0.03 and dev_rel[1] <= 0.03 and dev_rel[2] <= 0.10, dev_rel
    # extended energy: same series => same conservation
    assert np.max(np.abs(hg - hc)) <= 1e-8 * np.mean(np.abs(hc))
    assert abs(drift_g - drift_c) <= 1e-8 * np.mean(np.abs(hc))
