"""Replays one fuzz configuration (tests/test_gpu_fuzz.py::test_random_configuration_matches_oracle, same random draws) with
different launch options and reports bit differences (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import test_gpu_fuzz as F, cases, sdde_b200 as S, sdde_oracle as O
seed = int(sys.argv[1])
rng = np.random.default_rng(9000 + seed)
c = F._random_case(rng)
n = int(rng.choice([97, 256, 1000, 1537]))
tst = sorted(set(float(np.round(x, 2)) for x in rng.uniform(c.tspan[0], c.tspan[1], size=int(rng.integers(0, 3)))))
tst = [x for x in tst if c.tspan[0] < x < c.tspan[1]]
if c.alg in (2, 3, 4, 5, 6): tst = []
dd = [(float(np.round(x, 3)), int(rng.integers(0, 2))) for x in rng.uniform(c.tspan[0], c.tspan[1], size=int(rng.integers(0, 3)))]
dd = [d for d in dd if c.tspan[0] < d[0] < c.tspan[1]] if c.alg in (0, 1) else []
sv, ss, se = O.expand_saveat(c.saveat, *c.tspan)
sch = O.build_schedule(c.tspan[0], c.tspan[1], c.dt, c.lags, c.alg, order_discontinuity_t0=c.order0, constant_lags_declared=c.declared,
                       saveat_times=sv, save_start=ss, save_end=se, save_everystep=c.saveat is None, tstops=tst, d_discontinuities=dd)
nchan = c.model.n_chan
dW = cases.increments(sch, nchan, n, key=555 + seed)
ptraj = bool(rng.integers(0, 2))
pb = ub = None
if ptraj:
    pb = np.repeat(np.asarray(c.p, float)[:, None], n, axis=1) * (1.0 + 0.05 * rng.random((len(c.p), n)))
    ub = np.repeat(np.asarray(c.u0, float)[:, None], n, axis=1) * (1.0 + 0.05 * rng.random((len(c.u0), n)))
ref, st = O.solve_ensemble(c.omodel, pb if ptraj else c.p, ub if ptraj else c.u0, c.h0, c.h1, c.lags, sch, c.alg, dW)
grid = not bool(rng.integers(0, 2))
wave = int(rng.choice([0, 0, 128, 512])); chunk = int(rng.choice([128, 8, 24, 1024]))
lean, balance = int(rng.integers(0, 4) > 0), int(rng.integers(0, 4) > 0)
print(c.name, "alg", c.alg, c.lags, c.dt, c.tspan, "order0", c.order0, c.saveat, "n", n, "ptraj", ptraj, "grid", grid, "wave", wave, "chunk", chunk,
      "lean", lean, "balance", balance, "steps", sch.n_steps)
kw = dict(dt=c.dt, saveat=c.saveat, tstops=tst, d_discontinuities=dd)
if c.alg in (2, 3, 5, 6): kw["adaptive"] = False
ctx = S.get_context(0)
base = None
for (ln, ch, bal, wv, out) in [(lean, chunk, balance, wave, "host"), (1, 128, 1, 0, "device"), (1, 128, 1, 0, "host"), (lean, 128, 1, 0, "host"),
                               (1, chunk, 1, 0, "host"), (1, 128, balance, 0, "host"), (1, 128, 1, wave, "host"), (0, 128, 1, 0, "host"),
                               (1, 8, 1, 0, "host"), (1, 1024, 1, 0, "host"), (1, 128, 1, 128, "host")]:
    ctx.set_option("lean", ln); ctx.set_option("chunk_steps", ch); ctx.set_option("balance", bal)
    sim = S.solve(S.EnsembleProblem(c.problem(), p_batch=pb, u0_batch=ub), cases.ALG_MAP[c.alg],
                  S.EnsembleB200(device=0, output=out, wave_traj=wv or None), trajectories=n, noise_increments=dW, **kw)
    o = sim.out.cpu().numpy() if hasattr(sim.out, "cpu") else sim.out
    if base is None: base = o
    d = np.abs(o - base)
    k = np.unravel_index(d.argmax(), d.shape)
    print(f"lean={ln} chunk={ch:4d} balance={bal} wave={wv:4d} out={out:6s} err vs oracle {cases.scaled_err(o, ref):.2e}  max |diff vs first| {d.max():.3e} at {k}  {sim.timing['kernel_name']}")
ctx.set_option("lean", 1); ctx.set_option("chunk_steps", 128); ctx.set_option("balance", 1)
# ---- where do lean = 0 and lean = 1 part?  dense output, step by step ----
kw2 = dict(kw); kw2["saveat"] = None
outs = {}
for ln in (0, 1):
    ctx.set_option("lean", ln)
    sim = S.solve(S.EnsembleProblem(c.problem(), p_batch=pb, u0_batch=ub), cases.ALG_MAP[c.alg], S.EnsembleB200(device=0), trajectories=n,
                  noise_increments=dW, **kw2)
    outs[ln] = sim.out
ctx.set_option("lean", 1)
d = np.abs(outs[0] - outs[1])
steps = np.nonzero(d.max(axis=(1, 2)) > 0)[0]
sv2, ss2, se2 = O.expand_saveat(None, *c.tspan)
sch2 = O.build_schedule(c.tspan[0], c.tspan[1], c.dt, c.lags, c.alg, order_discontinuity_t0=c.order0, constant_lags_declared=c.declared,
                        saveat_times=sv2, save_start=ss2, save_end=se2, save_everystep=True, tstops=tst, d_discontinuities=dd)
print("dense: first differing save rows", steps[:10], "of", d.shape[0], "; trajectories at the first:", np.nonzero(d[steps[0]].max(axis=0) > 0)[0][:10] if len(steps) else None)
if len(steps):
    k0 = int(steps[0]) - 1          # save row r = node r; produced by step r-1
    for k in range(max(0, k0 - 4), min(sch2.n_steps, k0 + 3)):
        print("step", k, "t", sch2.t[k], "dt", sch2.dt[k], "modes", sch2.lk_mode[k], "idx", sch2.lk_idx[k], "theta", sch2.lk_theta[k])
