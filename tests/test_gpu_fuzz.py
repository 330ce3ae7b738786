"""-m gpu: randomised parity sweep.  Random combinations of model, algorithm, lag(s), step size, time span, discontinuity
order, saveat form, user tstops, user d_discontinuities, affine history, per-trajectory parameters / initial states, noise source (increments or
NoiseGrid), forced wave sizes, schedule-chunk sizes, lean batch bodies and balanced launch geometry on / off -- every one compared path by path with the oracle under shared
noise (1e-12).  Exercises the table packing (carried nodes, ring size, clipped steps, records at chunk boundaries, the
next-batch prefetch), the step-by-step fallback for short lags, and the wave loop, in combinations the named tests miss."""
import numpy as np
import pytest

import cases
import sdde_b200 as S
import sdde_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-12
# soak runs: SDDE_FUZZ_SEEDS="first:count" widens / shifts the seed ranges (default: the committed 64 / 16 / 16 cases)
import os as _os
_F0, _FN = ([int(x) for x in _os.environ["SDDE_FUZZ_SEEDS"].split(":")] if _os.environ.get("SDDE_FUZZ_SEEDS") else [0, 0])


def _seeds(n):
    return list(range(_F0, _F0 + (_FN or n)))


def _random_case(rng):
    kind = rng.choice(["linear1", "mg", "hayes", "scalar_add", "c4", "sir"], p=[0.3, 0.25, 0.15, 0.1, 0.1, 0.1])
    dt = float(rng.choice([0.01, 0.02, 0.05, 2.0 ** -5, 0.013]))
    T = float(rng.choice([1.5, 4.0, 7.3]))
    order0 = int(rng.choice([0, 0, 1, 2]))
    if kind == "linear1":
        lag = float(rng.choice([0.05, 0.3, 1.0, 0.777, 2.5 * dt, 1.0 * dt]))
        c = cases.c1(T=T).with_(lags=[lag], dt=dt, alg=int(rng.choice([0, 1, 2, 3, 4, 5, 6])), order0=order0,
                                h0=[float(rng.uniform(0.5, 1.5))], h1=[float(rng.choice([0.0, 0.3]))])
    elif kind == "mg":
        lag = float(rng.choice([0.5, 1.7, 17.0 / 8, 3.01]))
        c = cases.c2(T=T, tau=lag).with_(dt=dt, alg=int(rng.choice([0, 1, 2, 3, 4, 5])), order0=order0)
    elif kind == "hayes":
        c = cases.hayes(T=min(T, 3.0), dt=dt, alg=int(rng.choice([0, 2, 3, 4])), beta=float(rng.choice([0.0, -1.2])))
        c = c.with_(order0=order0)
    elif kind == "scalar_add":
        c = cases.scalar_add(T=T).with_(dt=dt, alg=int(rng.choice([0, 1, 4])), order0=order0)
    elif kind == "c4":
        c = cases.c4(T=min(T, 4.0)).with_(dt=dt, alg=int(rng.choice([0, 1, 4, 5, 6])), order0=order0,
                                         lags=[float(rng.choice([0.5, 1.0])), float(rng.choice([1.3, 2.5]))])
    else:
        c = cases.sir(T=min(T, 4.0), dt=0.02).with_(alg=int(rng.choice([0, 1, 4])))
    if c.alg in (2, 3, 4, 5, 6) and c.dt > min(c.lags):
        c = c.with_(dt=min(c.lags) / 2)                       # src/solve.jl:581-585 guard
    t0, tf = c.tspan
    form = rng.choice(["number", "list", "everystep"])
    if form == "number":
        c = c.with_(saveat=float(rng.choice([0.1, 0.25, 0.5])))
    elif form == "list":
        pts = sorted(set(float(np.round(x, 3)) for x in rng.uniform(t0, tf, size=int(rng.integers(1, 6)))))
        c = c.with_(saveat=([t0] if rng.integers(0, 2) else []) + pts + ([tf] if rng.integers(0, 2) else []))
    else:
        c = c.with_(saveat=None)
    return c


@pytest.mark.parametrize("seed", _seeds(64))
def test_random_configuration_matches_oracle(seed):
    rng = np.random.default_rng(9000 + seed)
    c = _random_case(rng)
    n = int(rng.choice([97, 256, 1000, 1537]))
    tst = sorted(set(float(np.round(x, 2)) for x in rng.uniform(c.tspan[0], c.tspan[1], size=int(rng.integers(0, 3)))))
    tst = [x for x in tst if c.tspan[0] < x < c.tspan[1]]
    if c.alg in (2, 3, 4, 5, 6):
        tst = []                                              # extra stops could make a step longer than the smallest lag
    # user-declared discontinuities (src/solve.jl:20-23): times with orders 0 / 1, propagated by the lags like the one at t0
    dd = [(float(np.round(x, 3)), int(rng.integers(0, 2))) for x in rng.uniform(c.tspan[0], c.tspan[1], size=int(rng.integers(0, 3)))]
    dd = [d for d in dd if c.tspan[0] < d[0] < c.tspan[1]] if c.alg in (0, 1) else []
    sv, ss, se = O.expand_saveat(c.saveat, *c.tspan)
    sch = O.build_schedule(c.tspan[0], c.tspan[1], c.dt, c.lags, c.alg, order_discontinuity_t0=c.order0,
                           constant_lags_declared=c.declared, saveat_times=sv, save_start=ss, save_end=se,
                           save_everystep=c.saveat is None, tstops=tst, d_discontinuities=dd)
    nchan = c.model.n_chan
    dW = cases.increments(sch, nchan, n, key=555 + seed)
    ptraj = bool(rng.integers(0, 2))
    pb = ub = None
    if ptraj:
        pb = np.repeat(np.asarray(c.p, float)[:, None], n, axis=1) * (1.0 + 0.05 * rng.random((len(c.p), n)))
        ub = np.repeat(np.asarray(c.u0, float)[:, None], n, axis=1) * (1.0 + 0.05 * rng.random((len(c.u0), n)))
    ref, st = O.solve_ensemble(c.omodel, pb if ptraj else c.p, ub if ptraj else c.u0, c.h0, c.h1, c.lags, sch, c.alg, dW)
    kw = dict(dt=c.dt, saveat=c.saveat, tstops=tst, d_discontinuities=dd)
    if c.alg in (2, 3, 5, 6):
        kw["adaptive"] = False
    noise = {}
    if rng.integers(0, 2):
        noise["noise_increments"] = dW
    else:
        noise["noise_grid"] = np.concatenate([np.zeros((1, nchan, n)), np.cumsum(dW, axis=0)], axis=0)
        cur = np.zeros((nchan, n)); dWg = np.empty_like(dW)
        for k in range(sch.n_steps):
            dWg[k] = noise["noise_grid"][k + 1] - cur
            cur = cur + dWg[k]
        ref, st = O.solve_ensemble(c.omodel, pb if ptraj else c.p, ub if ptraj else c.u0, c.h0, c.h1, c.lags, sch, c.alg, dWg)
    wave = int(rng.choice([0, 0, 128, 512]))
    ctx = S.get_context(0)
    ctx.set_option("chunk_steps", int(rng.choice([128, 8, 24, 1024])))
    lean, balance = int(rng.integers(0, 4) > 0), int(rng.integers(0, 4) > 0)          # mostly on (the defaults), sometimes off
    ctx.set_option("lean", lean); ctx.set_option("balance", balance)
    try:
        sim = S.solve(S.EnsembleProblem(c.problem(), p_batch=pb, u0_batch=ub), cases.ALG_MAP[c.alg],
                      S.EnsembleB200(device=0, wave_traj=wave or None), trajectories=n, **kw, **noise)
    finally:
        ctx.set_option("chunk_steps", 128); ctx.set_option("lean", 1); ctx.set_option("balance", 1)
    what = (c.name, c.alg, c.lags, c.dt, c.tspan, c.order0, c.saveat if not isinstance(c.saveat, list) else "list", tst, dd, n, ptraj, wave, lean, balance,
            list(noise)[0], sim.timing["kernel_name"])
    assert sim.out.shape == ref.shape, what
    assert np.array_equal(sim.t, sch.save_t), what
    ok = np.isfinite(ref).all(axis=(0, 1))
    assert ok.any(), what
    assert cases.scaled_err(sim.out[:, :, ok], ref[:, :, ok]) < TOL, what
    assert np.array_equal(sim.status, st), what
    # the same solve with a random subset of saved states, device-resident output and ensemble moments
    ns = ref.shape[1]
    idxs = sorted(rng.choice(ns, size=int(rng.integers(1, ns + 1)), replace=False).tolist())
    sim2 = S.solve(S.EnsembleProblem(c.problem(), p_batch=pb, u0_batch=ub), cases.ALG_MAP[c.alg],
                   S.EnsembleB200(device=0, output="device"), trajectories=n, save_idxs=idxs, moments=True, **kw, **noise)
    out2 = sim2.out.cpu().numpy()
    assert np.array_equal(out2, sim.out[:, idxs, :]), what
    if ok.all():
        om = O.moments(ref[:, idxs, :])
        assert np.allclose(sim2.moments[..., 0], n) and np.allclose(sim2.moments[..., 1], om[..., 1], rtol=1e-11, atol=1e-13), what
        assert np.allclose(sim2.moments[..., 2], om[..., 2], rtol=1e-8, atol=1e-13), what
        assert sim2.n_unstable == 0


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_configuration_philox_paths_do_not_depend_on_the_launch_geometry(seed):
    """In-kernel Philox noise: the paths of a random configuration are the same bit for bit whatever the wave size, the
    schedule-chunk size and the output residency (counter = global trajectory id, key = seed)."""
    rng = np.random.default_rng(7000 + seed)
    c = _random_case(rng)
    n = int(rng.choice([300, 1000, 2049]))
    kw = dict(dt=c.dt, saveat=c.saveat, trajectories=n, seed=1234 + seed)
    if c.alg in (2, 3, 5, 6):
        kw["adaptive"] = False
    prob = S.EnsembleProblem(c.problem())
    base = S.solve(prob, cases.ALG_MAP[c.alg], S.EnsembleB200(device=0), **kw)
    ctx = S.get_context(0)
    ctx.set_option("chunk_steps", int(rng.choice([8, 16, 1024])))
    try:
        other = S.solve(prob, cases.ALG_MAP[c.alg], S.EnsembleB200(device=0, wave_traj=int(rng.choice([128, 256, 1024])), output="device"), **kw)
    finally:
        ctx.set_option("chunk_steps", 128)
    what = (c.name, c.alg, c.lags, c.dt, c.tspan, n, base.timing["kernel_name"], other.timing["waves"])
    assert np.array_equal(base.out, other.out.cpu().numpy(), equal_nan=True), what
    assert np.array_equal(base.status, other.status), what


@pytest.mark.parametrize("seed", _seeds(16))
def test_random_sra_configuration_matches_oracle(seed):
    """The SRA family on random additive problems: Hayes with / without the delayed term in g, the scalar additive model, lags
    on and off the step grid and equal to dt, saveat forms, per-trajectory parameters, forced waves, chunk sizes, increments or
    NoiseGrid(t, W, Z) -- per path against the oracle (1e-12)."""
    rng = np.random.default_rng(13000 + seed)
    alg = int(rng.choice([13, 14, 15]))
    dt = float(rng.choice([0.01, 0.02, 0.05, 2.0 ** -5, 0.013]))
    T = float(rng.choice([1.5, 3.0]))
    if rng.integers(0, 3) < 2:
        lag = float(rng.choice([1.0, 0.33, 0.5, dt, 2.5 * dt]))
        c = cases.hayes(T=T, dt=dt, beta=0.0).with_(alg=alg, lags=[lag], order0=int(rng.choice([0, 0, 1, 3])))
        c.p = np.array([lag, -4.0, -2.0, 10.0, 0.0, float(rng.choice([0.0, 0.6])), 0.1])
    else:
        c = cases.scalar_add(T=T).with_(dt=dt, alg=alg, order0=int(rng.choice([0, 1])))
    if c.dt > min(c.lags):
        c = c.with_(dt=min(c.lags) / 2)
    form = rng.choice(["number", "list", "everystep"])
    if form == "number":
        c = c.with_(saveat=float(rng.choice([0.1, 0.25])))
    elif form == "list":
        c = c.with_(saveat=sorted(set(float(np.round(x, 3)) for x in rng.uniform(0.0, T, size=3))))
    else:
        c = c.with_(saveat=None)
    n = int(rng.choice([97, 256, 1000]))
    sch = c.oracle_schedule()
    nz = cases.increments(sch, 2, n, key=777 + seed)
    ptraj = bool(rng.integers(0, 2))
    pb = np.repeat(np.asarray(c.p, float)[:, None], n, axis=1) * (1.0 + 0.05 * rng.random((len(c.p), n))) if ptraj else None
    if ptraj and c.name == "hayes":
        pb[0] = c.p[0]                                       # the lag is a schedule quantity, not a per-trajectory one
        pb[4] = 0.0
    noise = {}
    dWo = nz
    if rng.integers(0, 2):
        noise["noise_increments"] = nz
    else:
        noise["noise_grid"] = np.concatenate([np.zeros((1, 2, n)), np.cumsum(nz, axis=0)], axis=0)
        dWo = O.noise_grid_increments(noise["noise_grid"])
    ref, st = O.solve_ensemble(c.omodel, pb if ptraj else c.p, c.u0, c.h0, c.h1, c.lags, sch, alg, dWo)
    ctx = S.get_context(0)
    ctx.set_option("chunk_steps", int(rng.choice([128, 8, 24])))
    try:
        sim = S.solve(S.EnsembleProblem(c.problem(), p_batch=pb), cases.ALG_MAP[alg],
                      S.EnsembleB200(device=0, wave_traj=int(rng.choice([0, 128])) or None), trajectories=n, dt=c.dt, saveat=c.saveat,
                      adaptive=False, **noise)
    finally:
        ctx.set_option("chunk_steps", 128)
    what = (c.name, alg, c.lags, c.dt, T, c.order0, form, n, ptraj, list(noise)[0])
    assert np.array_equal(sim.t, sch.save_t), what
    assert cases.scaled_err(sim.out, ref) < TOL, what
    assert np.array_equal(sim.status, st), what


@pytest.mark.parametrize("seed", _seeds(8))
def test_random_noise_grid_on_its_own_time_grid(seed):
    """NoiseGrid(t, W[, Z]) on a random grid unrelated to the steps (random spacing, some shared nodes, the last node possibly a
    few ulps off tf), random model / algorithm (incl. SRA with its Z path) / waves / residency: W is interpolated on the device
    exactly as the oracle restates DiffEqNoiseProcess's rule."""
    rng = np.random.default_rng(17000 + seed)
    sra = rng.integers(0, 4) == 0
    if sra:
        alg = int(rng.choice([13, 14, 15]))
        c = cases.hayes(T=float(rng.choice([1.5, 3.0])), dt=float(rng.choice([0.02, 0.05])), beta=0.0).with_(alg=alg, lags=[0.33])
        c.p = np.array([0.33, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1])
        c = c.with_(saveat=0.25)
    else:
        c = _random_case(rng)
        if c.alg >= 5 and c.name != "C4":
            c = c.with_(alg=int(rng.choice([0, 2])))
    n = int(rng.choice([65, 300, 1000]))
    sch = c.oracle_schedule()
    t0, tf = c.tspan
    nch = c.model.n_chan * (2 if c.alg in (13, 14, 15) else 1)
    m = int(rng.integers(5, 400))
    ts = np.sort(rng.uniform(t0, tf, size=m))
    ts = np.unique(np.concatenate([[t0], ts, sch.t[rng.integers(0, sch.t.size, size=3)], [tf * (1 + float(rng.choice([0.0, 3e-16]))) if tf != 0 else tf]]))
    Wg = np.cumsum(np.sqrt(np.diff(ts, prepend=ts[0]))[:, None, None] * rng.standard_normal((ts.size, nch, n)), axis=0)
    Wn = O.noise_grid_on_nodes(ts, Wg, sch.t)
    ref, st = O.solve_ensemble(c.omodel, c.p, c.u0, c.h0, c.h1, c.lags, sch, c.alg, O.noise_grid_increments(Wn))
    W_in = Wg
    if rng.integers(0, 2):
        import torch
        W_in = torch.from_numpy(Wg).cuda()
    kw = dict(dt=c.dt, saveat=c.saveat)
    if c.alg in (2, 3, 5, 6, 13, 14, 15):
        kw["adaptive"] = False
    sim = S.solve(S.EnsembleProblem(c.problem()), cases.ALG_MAP[c.alg], S.EnsembleB200(device=0, wave_traj=int(rng.choice([0, 128])) or None),
                  trajectories=n, noise_grid=(ts, W_in), **kw)
    S.get_context(0).set_option("wave_traj", 0)
    ok = np.isfinite(ref).all(axis=(0, 1))
    what = (c.name, c.alg, c.lags, c.dt, c.tspan, m, n)
    assert ok.any() and cases.scaled_err(sim.out[:, :, ok], ref[:, :, ok]) < TOL, what
