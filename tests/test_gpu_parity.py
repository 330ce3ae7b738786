"""-m gpu: per-path parity of the CUDA path against the oracle under shared (supplied) Brownian increments.

Tolerance: 1e-12 relative in float64 (BASELINE.json north_star).  The CUDA kernels contract a*b+c into FMA like
Julia's @muladd does in the reference's step functions (bit identity with the un-fused oracle is not expected)."""
import numpy as np
import pytest

import cases
import sdde_b200 as S

pytestmark = pytest.mark.gpu
TOL = 1e-12


def run_case(case, n_traj, **extra):
    sch = case.oracle_schedule()
    nchan = case.model.n_chan
    dW = cases.increments(sch, nchan, n_traj)
    sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0),
                  trajectories=n_traj, noise_increments=dW, moments=True, **case.solve_kwargs(), **extra)
    ref, st = case.oracle(dW)
    assert sim.out.shape == ref.shape
    assert np.array_equal(sim.t, sch.save_t)
    return sim, ref, st


@pytest.mark.parametrize("alg", [0, 1, 2, 3])
def test_c1_linear_all_algs(alg):
    case = cases.c1().with_(alg=alg)
    sim, ref, st = run_case(case, 1000)
    assert cases.scaled_err(sim.out, ref) < TOL
    assert np.array_equal(sim.status, st)


def test_c1_save_everystep():
    case = cases.c1(saveat=None)
    sim, ref, _ = run_case(case, 300)
    assert sim.out.shape[0] == 1002
    assert cases.scaled_err(sim.out, ref) < TOL


@pytest.mark.parametrize("alg", [0, 1])
def test_c2_mackey_glass_em(alg):
    case = cases.c2(alg=alg)
    sim, ref, _ = run_case(case, 4096)
    assert sim.timing["n_steps"] == 2001
    assert cases.relerr(sim.out, ref) < TOL


@pytest.mark.parametrize("order0", [0, 2])
@pytest.mark.parametrize("alg", [2, 3])
def test_c3_nongrid_lag_rkmil(order0, alg):
    case = cases.c3(order0=order0, alg=alg)
    sim, ref, _ = run_case(case, 4096)
    assert sim.timing["n_steps"] == (2002 if order0 == 0 else 2001)
    assert cases.relerr(sim.out, ref) < TOL


def test_c4_four_state_two_lags_nondiagonal():
    case = cases.c4()
    sim, ref, _ = run_case(case, 2048)
    assert sim.timing["n_steps"] == 2003
    assert cases.scaled_err(sim.out, ref) < TOL


def test_hayes_time_dependent_history():
    case = cases.hayes(T=3.0, dt=2.0 ** -6, beta=-1.2, saveat=0.25)   # test_prob_sol.jl g with the delayed term
    sim, ref, _ = run_case(case, 512)
    assert cases.scaled_err(sim.out, ref) < TOL


def test_sir_sparse_noise_undeclared_lag():
    case = cases.sir(T=20.0, dt=0.02)    # without the reference test's callback EM stays finite on this span only
    sim, ref, _ = run_case(case, 512)
    assert sim.timing["n_steps"] == 1001
    assert cases.scaled_err(sim.out, ref) < 1e-11    # sqrt(max(0,x)) near 0 amplifies rounding differences


def test_scalar_additive_noise():
    case = cases.scalar_add()
    sim, ref, _ = run_case(case, 512)
    assert cases.scaled_err(sim.out, ref) < TOL


def test_moments_match_numpy():
    import sdde_oracle as O
    case = cases.c2(T=40.0)
    sim, ref, _ = run_case(case, 3001)      # odd count: exercises the scalar-store tail
    om = O.moments(ref)
    assert np.allclose(sim.moments[..., 0], 3001)
    assert np.allclose(sim.moments[..., 1], om[..., 1], rtol=1e-12, atol=1e-14)
    assert np.allclose(sim.moments[..., 2], om[..., 2], rtol=1e-9, atol=1e-14)


def test_waves_and_device_output_agree():
    import torch
    case = cases.c2(T=30.0)
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, 1500)
    kw = dict(trajectories=1500, noise_increments=dW, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, wave_traj=512), moments=True, **kw)
    c = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device", wave_traj=512), **kw)
    S.get_context(0).set_option("wave_traj", 0)
    assert b.timing["waves"] == 3
    assert np.array_equal(a.out, b.out)
    assert np.array_equal(a.out, c.out.cpu().numpy())
    dWd = torch.from_numpy(dW).cuda()
    d = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=1500,
                noise_increments=dWd, **case.solve_kwargs())
    assert np.array_equal(a.out, d.out)


def test_noise_grid_mode():
    case = cases.c1(T=4.0)
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, 256)
    W = np.concatenate([np.zeros((1, 1, 256)), np.cumsum(dW, axis=0)], axis=0)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=256, noise_grid=W,
                  **case.solve_kwargs())
    # the increments the kernel reconstructs: dW_n = W[n+1] - curW, curW += dW_n
    cur = np.zeros((1, 256)); dWg = np.empty_like(dW)
    for n in range(sch.n_steps):
        dWg[n] = W[n + 1] - cur
        cur = cur + dWg[n]
    ref, _ = case.oracle(dWg)
    assert cases.scaled_err(sim.out, ref) < TOL


@pytest.mark.parametrize("on_device", [False, True])
def test_noise_grid_on_its_own_time_grid(on_device):
    """NoiseGrid(t, W) whose grid is unrelated to the step grid (irregular spacing, a few shared nodes, last node a few
    ulps off tf): W is interpolated at the step ends on the device; C3-shaped problem (clipped steps at the lag multiples),
    two waves."""
    import sdde_oracle as O
    case = cases.c3(T=40.0)
    sch = case.oracle_schedule()
    n = 300
    rng = np.random.default_rng(77)
    gaps = rng.uniform(0.004, 0.09, size=1200)
    ts = np.concatenate([[0.0], np.cumsum(gaps)])
    ts = ts[ts < 40.0]
    ts = np.unique(np.concatenate([ts, [0.05, 0.1, 17.02, 20.0], [40.0 * (1 + 3e-16)]]))
    Wg = np.cumsum(np.sqrt(np.diff(ts, prepend=ts[0]))[:, None, None] * rng.standard_normal((ts.size, 1, n)), axis=0)
    Wn = O.noise_grid_on_nodes(ts, Wg, sch.t)
    dW = O.noise_grid_increments(Wn)
    ref, _ = case.oracle(dW)
    W_in = Wg
    if on_device:
        import torch
        W_in = torch.from_numpy(Wg).cuda()
    sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0, wave_traj=256), trajectories=n,
                  noise_grid=(ts, W_in), **case.solve_kwargs())
    assert sim.timing["waves"] == 2
    assert cases.scaled_err(sim.out, ref) < TOL
    short = ts[ts < 30.0]
    with pytest.raises(S.SddeError):
        S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0), trajectories=n,
                noise_grid=(short, Wg[:short.size]), **case.solve_kwargs())


def test_page_locked_host_output_is_written_by_the_kernel_itself():
    """Host `out` in page-locked memory: the kernel stores the paths straight into it (one wave, no staging copy);
    pageable `out` and the option switched off take the staged path; all three agree bit for bit."""
    case = cases.c2(T=40.0)
    n = 70_000
    kw = dict(trajectories=n, seed=21, **case.solve_kwargs())
    prob = S.EnsembleProblem(case.problem())
    n_save = len(case.oracle_schedule().save_t)
    pinned = S.pinned_empty((n_save, 1, n))
    a = S.solve(prob, S.EM(), S.EnsembleB200(device=0), out=pinned, **kw)
    b = S.solve(prob, S.EM(), S.EnsembleB200(device=0), **kw)                      # pageable numpy output
    ctx = S.get_context(0)
    ctx.set_option("zero_copy", 0)
    try:
        c = S.solve(prob, S.EM(), S.EnsembleB200(device=0), out=S.pinned_empty((n_save, 1, n)), **kw)
    finally:
        ctx.set_option("zero_copy", -1)
    assert "mapped host out" in a.timing["kernel_name"]
    assert "mapped host out" not in b.timing["kernel_name"] and "mapped host out" not in c.timing["kernel_name"]
    assert a.timing["d2h_bytes"] >= n_save * n * 8 and c.timing["d2h_bytes"] >= n_save * n * 8
    assert np.array_equal(a.out, b.out) and np.array_equal(a.out, c.out) and np.array_equal(a.status, b.status)


def test_per_trajectory_params_and_u0():
    case = cases.c1(T=5.0)
    n = 700
    rng = np.random.default_rng(5)
    pb = np.stack([-1.0 + 0.2 * rng.random(n), 0.5 + 0.1 * rng.random(n), 0.2 * np.ones(n)])
    ub = 1.0 + 0.1 * rng.random((1, n))
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, n)
    sim = S.solve(S.EnsembleProblem(case.problem(), p_batch=pb, u0_batch=ub), S.EM(), S.EnsembleB200(device=0),
                  trajectories=n, noise_increments=dW, **case.solve_kwargs())
    ref, _ = case.oracle(dW, params=pb, u0=ub)
    assert cases.scaled_err(sim.out, ref) < TOL
    # prob_func form (the reference's EnsembleProblem interface)
    ep = S.EnsembleProblem(case.problem(), prob_func=lambda prob, i, rep: prob.remake(p=pb[:, i - 1], u0=ub[:, i - 1]))
    sim2 = S.solve(ep, S.EM(), S.EnsembleB200(device=0), trajectories=n, noise_increments=dW, **case.solve_kwargs())
    assert np.array_equal(sim.out, sim2.out)


def test_unstable_trajectories_flagged():
    case = cases.c1(T=10.0).with_(p=np.array([300.0, 0.5, 0.2]))   # explicit Euler blows up -> Inf/NaN
    sim, ref, st = run_case(case, 64)
    assert np.all(sim.status == 1) and np.all(st == 1)
    assert sim.retcodes[0] == "Unstable" and sim.n_unstable == 64
    # a moments-only run carries no per-trajectory array, the count comes from the device; mixed stable/unstable ensemble
    n = 3000
    pb = np.tile(np.array([[-1.0], [0.5], [0.2]]), (1, n))
    pb[0, ::3] = 300.0
    dW = cases.increments(case.oracle_schedule(), 1, n)
    good = cases.c1(T=10.0)
    m = S.solve(S.EnsembleProblem(good.problem(), p_batch=pb), S.EM(), S.EnsembleB200(device=0, wave_traj=1024), trajectories=n,
                noise_increments=dW, paths=False, moments=True, **good.solve_kwargs())
    assert m.status is None and m.n_unstable == 1000 and m.timing["waves"] == 3
    stable = S.solve(S.EnsembleProblem(good.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=1, **good.solve_kwargs())
    assert stable.n_unstable == 0 and int(stable.status.sum()) == 0


def test_c4_cooperative_lanes_kernel_matches_oracle_and_thread_kernel():
    """BASELINE.json configs[3]: 4 lanes of a warp per trajectory (sdde_coop.cuh) vs thread-per-trajectory."""
    case = cases.c4()
    sch = case.oracle_schedule()
    n = 1000
    dW = cases.increments(sch, 2, n)
    kw = dict(trajectories=n, noise_increments=dW, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, lanes_per_traj=4), moments=True, **kw)
    assert "coop" in b.timing["kernel_name"] and "coop" not in a.timing["kernel_name"]
    ref, st = case.oracle(dW)
    assert cases.scaled_err(b.out, ref) < TOL
    assert cases.scaled_err(b.out, a.out) < 1e-13
    assert np.array_equal(b.status, st)
    # Philox mode: identical streams in both mappings
    kw = dict(trajectories=n, seed=5, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(False), S.EnsembleB200(device=0), **kw)
    b = S.solve(S.EnsembleProblem(case.problem()), S.EM(False), S.EnsembleB200(device=0, lanes_per_traj=4, wave_traj=320), **kw)
    S.get_context(0).set_option("wave_traj", 0)
    assert cases.scaled_err(b.out, a.out) < 1e-13


# ---- edge cases the reference's loop handles implicitly (empty / ragged inputs, degenerate schedules) ----
@pytest.mark.parametrize("n", [1, 2, 3, 255, 257])
def test_ragged_trajectory_counts(n):
    case = cases.c2(T=25.0)
    sim, ref, _ = run_case(case, n)
    assert cases.relerr(sim.out, ref) < TOL


def test_empty_ensemble():
    case = cases.c1(T=1.0)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=0, moments=True,
                  **case.solve_kwargs())
    assert sim.out.shape == (11, 1, 0) and np.all(sim.moments == 0)


def test_span_shorter_than_lag_reads_only_prehistory():
    # the reference's own smoke test has tspan (0, 0.1) < tau = 1 (test/test_prob_sol.jl:15)
    case = cases.hayes(T=0.1, dt=0.01, beta=-1.2, saveat=None)
    sim, ref, _ = run_case(case, 64)
    assert sim.out.shape[0] == 11 and np.all(sim.out[-1] != 0.0)        # "sol.u[end] != zeros(1)"
    assert cases.scaled_err(sim.out, ref) < TOL


def test_dt_larger_than_lag_with_em():
    # allowed for EM only (src/solve.jl:581-585): the lookup lands inside the step just taken
    case = cases.c1(T=6.0).with_(dt=0.3, lags=[0.1], saveat=0.5)
    sim, ref, _ = run_case(case, 300)
    assert cases.scaled_err(sim.out, ref) < TOL


def test_many_saves_per_step_and_user_tstops():
    case = cases.c1(T=2.0, saveat=0.003)
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, 200)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=200,
                  noise_increments=dW, **case.solve_kwargs())
    ref, _ = case.oracle(dW)
    assert sim.out.shape[0] == len(sch.save_t) == 667 and cases.scaled_err(sim.out, ref) < TOL
    # user tstops shorten steps; saveat as an explicit list without the end points
    import sdde_oracle as O
    case = cases.c1(T=3.0)
    kw = dict(dt=0.01, saveat=[0.5, 1.25, 2.9], tstops=[0.333, 2.0])
    osch = O.build_schedule(0.0, 3.0, 0.01, [1.0], 0, saveat_times=[0.5, 1.25, 2.9], save_start=False, save_end=False,
                            tstops=[0.333, 2.0])
    dW = cases.increments(osch, 1, 100)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=100,
                  noise_increments=dW, **kw)
    ref, _ = O.solve_ensemble(case.omodel, case.p, case.u0, case.h0, case.h1, case.lags, osch, 0, dW)
    assert list(sim.t) == [0.5, 1.25, 2.9] and cases.scaled_err(sim.out, ref) < TOL


def test_golden_fixtures_on_gpu():
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))
    for name, tau, alg, order0 in [("c2_em", 17.0, 0, 0), ("c3_rkmil_o0", 17.02, 2, 0), ("c3_rkmil_o2", 17.02, 2, 2)]:
        case = cases.c2(T=40.0, tau=tau, alg=alg, order0=order0)
        sch = case.oracle_schedule()
        dW = g[name + "_xi"] * sch.sqdt[:, None, None]
        sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[alg], S.EnsembleB200(device=0), trajectories=16,
                      noise_increments=dW, **case.solve_kwargs())
        assert cases.relerr(sim.out, g[name + "_out"]) < TOL
    case = cases.c4(T=6.0)
    sch = case.oracle_schedule()
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=16,
                  noise_increments=g["c4_xi"] * sch.sqdt[:, None, None], **case.solve_kwargs())
    assert cases.scaled_err(sim.out, g["c4_out"]) < TOL
    from test_gpu_rng import device_normals
    z = device_normals(0xB2000001, 0, 8, 6, 3)
    assert np.max(np.abs(z - g["philox_z"][:, :, :8])) < 4e-15 * 8


def test_full_size_properties_c2():
    """BASELINE.json configs[1] at full size (10^6 paths, 10 000 steps): size-independent checks -- a 4096-path slice
    solved separately with the matching traj_offset reproduces its rows bit for bit (partition invariance), moments
    equal the two-pass numpy moments of the returned paths, all paths finite."""
    import torch
    case = cases.c2(T=500.0)
    n = 1_000_000
    ea = S.EnsembleB200(device=0, output="device")
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), ea, trajectories=n, seed=20261019, moments=True,
                  **case.solve_kwargs())
    assert sim.timing["n_steps"] == 10000 and tuple(sim.out.shape) == (501, 1, n)
    assert bool(torch.isfinite(sim.out).all()) and int(sim.status.sum()) == 0
    sl = sim.out[:, :, 500_000:504_096].cpu().numpy()
    from sdde_b200.ensemble import _marshal  # noqa: F401  (public solve has no traj_offset kw: shard via distributed-free API)
    # slice via the C ABI directly: same seed, traj_offset = 500000
    import ctypes as C
    from sdde_b200 import _capi as capi
    m = _marshal(case.problem(), case.model, S.EM(), dt=case.dt, saveat=1.0, save_everystep=None, save_start=None,
                 save_end=None, tstops=(), adaptive=None, n_traj=4096, traj_offset=500_000, seed=20261019, p_batch=None,
                 u0_batch=None, noise=None, noise_mode=capi.NOISE_PHILOX, noise_mem=0, out_mem=capi.MEM_HOST, want_paths=1,
                 want_moments=0)
    out = np.empty((501, 1, 4096))
    capi.check(capi.lib().sdde_b200_solve(S.get_context(0).handle, C.byref(m.desc), capi.np_ptr(out), None, None, None))
    assert np.array_equal(out, sl)
    # the same rows against the ORACLE fed with its own restatement of the in-kernel Philox normal stream: per-path parity
    # with in-kernel noise over the full 10 000 steps of the bench configuration (512 paths of the middle of the ensemble)
    import sdde_oracle as O
    sch = case.oracle_schedule()
    ids = np.arange(500_000, 500_512, dtype=np.uint64)
    z = O.philox_normals(ids, sch.n_steps, 1, 20261019)
    ref, _ = case.oracle(z * sch.sqdt[:, None, None])
    assert cases.relerr(sl[:, :, :512], ref) < 1e-11
    mean = sim.out.mean(dim=2).cpu().numpy()
    assert np.allclose(sim.moments[..., 1], mean, rtol=1e-12)
    var = sim.out.var(dim=2, unbiased=True).cpu().numpy()
    assert np.allclose(sim.moments[..., 2] / (n - 1), var, rtol=1e-9)


def test_save_idxs_subset_and_order():
    """save_idxs keyword (src/solve.jl:355-356, interface.jl:386-391): only the listed states are saved, in list order."""
    case = cases.c4(T=5.0)
    sch = case.oracle_schedule()
    n = 300
    dW = cases.increments(sch, 2, n)
    ref, _ = case.oracle(dW)
    for lanes in (0, 4):
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, lanes_per_traj=lanes),
                      trajectories=n, noise_increments=dW, save_idxs=[2, 0], moments=True, **case.solve_kwargs())
        assert sim.out.shape == (len(sch.save_t), 2, n)
        assert cases.scaled_err(sim.out, ref[:, [2, 0], :]) < TOL
        assert np.allclose(sim.moments[..., 1], ref[:, [2, 0], :].mean(axis=2), rtol=1e-12, atol=1e-15)
    with pytest.raises(ValueError):
        S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=4, save_idxs=[1, 1],
                **case.solve_kwargs())


# ---- EulerHeun: a second set of delayed-state lookups per step, at t + dt ----
@pytest.mark.parametrize("mk,n", [
    (lambda: cases.c1().with_(alg=4), 500), (lambda: cases.c2(T=100.0, alg=4), 2048), (lambda: cases.c3(0, T=60.0, alg=4), 1024),
    (lambda: cases.c4(T=8.0).with_(alg=4), 512), (lambda: cases.sir(T=10.0, dt=0.02).with_(alg=4), 256),
    (lambda: cases.hayes(T=3.0, beta=-1.2, saveat=0.25, alg=4), 256), (lambda: cases.scalar_add().with_(alg=4), 128)])
def test_euler_heun_matches_oracle(mk, n):
    case = mk()
    sim, ref, st = run_case(case, n)
    assert "alg4" in sim.timing["kernel_name"]
    assert cases.scaled_err(sim.out, ref) < (1e-11 if case.name == "sir" else TOL)
    assert np.array_equal(sim.status, st)


def test_euler_heun_step_equal_to_lag_guard():
    # src/solve.jl:581-585 applies to every algorithm but EM
    case = cases.c1(T=5.0).with_(dt=1.5, alg=4)
    with pytest.raises(S.SddeError, match="dt > minimum"):
        S.solve(S.EnsembleProblem(case.problem()), S.EulerHeun(), S.EnsembleB200(device=0), trajectories=4, **case.solve_kwargs())


def test_cooperative_kernel_with_per_trajectory_initial_states():
    case = cases.c4(T=4.0)
    sch = case.oracle_schedule()
    n = 257
    rng = np.random.default_rng(3)
    ub = np.asarray(case.u0)[:, None] * (1.0 + 0.1 * rng.random((4, n)))
    dW = cases.increments(sch, 2, n)
    ref, _ = case.oracle(dW, u0=ub)
    for lanes in (0, 4):
        sim = S.solve(S.EnsembleProblem(case.problem(), u0_batch=ub), S.EM(), S.EnsembleB200(device=0, lanes_per_traj=lanes),
                      trajectories=n, noise_increments=dW, **case.solve_kwargs())
        assert cases.scaled_err(sim.out, ref) < TOL


def test_rkmil_with_per_trajectory_parameters():
    case = cases.c3(0, T=40.0)
    n = 600
    rng = np.random.default_rng(8)
    pb = np.stack([0.2 + 0.02 * rng.random(n), 0.1 * np.ones(n), 10.0 * np.ones(n), 0.05 + 0.1 * rng.random(n)])
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, n)
    sim = S.solve(S.EnsembleProblem(case.problem(), p_batch=pb), S.RKMil(), S.EnsembleB200(device=0), trajectories=n,
                  noise_increments=dW, **case.solve_kwargs())
    ref, _ = case.oracle(dW, params=pb)
    assert "mackey_glass," in sim.timing["kernel_name"]          # generic-exponent kernel (parameters vary per path)
    assert cases.relerr(sim.out, ref) < TOL


def test_memory_limit_forces_waves_same_result():
    """The wave size follows the memory budget (C5's 10^8 paths never need an 82 GB ring): with a 16 MB budget the
    same ensemble runs in many waves and returns identical paths and moments (up to the merge order)."""
    case = cases.c2(T=30.0)
    n = 20000
    kw = dict(trajectories=n, seed=31, moments=True, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    ctx = S.get_context(0)
    ctx.set_option("mem_limit", 16 << 20)
    try:
        b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    finally:
        ctx.set_option("mem_limit", 0)
    assert b.timing["waves"] > a.timing["waves"] >= 1
    assert np.array_equal(a.out, b.out)
    assert np.allclose(a.moments, b.moments, rtol=1e-12, atol=1e-15)


def test_two_contexts_on_one_gpu_are_independent():
    from sdde_b200 import _capi as capi
    case = cases.c1(T=2.0)
    c2 = capi.Context(0)
    try:
        import threading
        res = {}

        def run(tag, seed):
            # each thread drives its own context (the default one / a second one) concurrently
            if tag == "b":
                capi._contexts[0], saved = c2, capi._contexts[0]
            res[tag] = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=3000,
                               seed=seed, **case.solve_kwargs()).out
        run("a", 5)
        saved_ctx = capi._contexts[0]
        capi._contexts[0] = c2
        res["b"] = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=3000, seed=5,
                           **case.solve_kwargs()).out
        capi._contexts[0] = saved_ctx
        assert np.array_equal(res["a"], res["b"])
    finally:
        c2.close()


def test_pinned_output_buffer_reused_across_solves():
    case = cases.c2(T=20.0)
    n = 4096
    sch = case.oracle_schedule()
    out = S.pinned_empty((len(sch.save_t), 1, n))
    ref = None
    for seed in (1, 2):
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=seed,
                      out=out, **case.solve_kwargs())
        assert sim.out is out or np.shares_memory(sim.out, out)
        plain = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=seed,
                        **case.solve_kwargs())
        assert np.array_equal(plain.out, out)


def test_linsys_nvrtc_constant_specialisation_equals_generic_kernel():
    """The linear delay system's matrices are baked into an NVRTC-compiled kernel (zeros vanish); the generic
    constant-bank kernel stays available (option jit_linsys = 0) and both give the same paths."""
    case = cases.c4(T=5.0)
    sch = case.oracle_schedule()
    n = 500
    dW = cases.increments(sch, 2, n)
    kw = dict(trajectories=n, noise_increments=dW, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    ctx = S.get_context(0)
    ctx.set_option("jit_linsys", 0)
    try:
        b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    finally:
        ctx.set_option("jit_linsys", 1)
    assert "nvrtc_constants" in a.timing["kernel_name"] and "nvrtc" not in b.timing["kernel_name"]
    assert cases.scaled_err(a.out, b.out) < 1e-13
    ref, _ = case.oracle(dW)
    assert cases.scaled_err(a.out, ref) < TOL and cases.scaled_err(b.out, ref) < TOL


@pytest.mark.parametrize("chunk", [8, 40, 1024])
def test_schedule_chunk_sizes_give_identical_paths(chunk):
    """The schedule records are staged in shared memory chunk by chunk; any chunk size (down to one batch, up to the
    1024-record maximum, which needs the > 48 KB shared-memory opt-in) gives the same paths, for a built-in model, a
    two-stage algorithm on the 4-state system (176-byte records) and an NVRTC user model."""
    ctx = S.get_context(0)
    runs = []
    c2 = cases.c2(T=100.0)
    c4 = cases.c4(T=12.0).with_(alg=4)
    for case, n in ((c2, 512), (c4, 300)):
        sch = case.oracle_schedule()
        dW = cases.increments(sch, case.model.n_chan, n)
        kw = dict(trajectories=n, noise_increments=dW, **case.solve_kwargs())
        ref = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0), **kw)
        ctx.set_option("chunk_steps", chunk)
        try:
            got = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0), **kw)
        finally:
            ctx.set_option("chunk_steps", 128)
        runs.append((case.name, np.array_equal(ref.out, got.out)))
    assert all(ok for _, ok in runs), runs


def test_d_discontinuities_keyword_paths_match_oracle():
    """User-declared discontinuities (src/solve.jl:20-23): extra clipped steps at the declared times and at their
    lag-propagated images; per-path parity under shared increments, EM and RKMil."""
    import sdde_oracle as O
    dd = [(0.37, 0), (2.2137, 1), 4.05]
    for algid in (0, 2):
        case = cases.c1(T=6.0, saveat=0.5).with_(alg=algid)
        sv, ss, se = O.expand_saveat(case.saveat, *case.tspan)
        sch = O.build_schedule(0.0, 6.0, case.dt, case.lags, algid, saveat_times=sv, save_start=ss, save_end=se, d_discontinuities=dd)
        dW = cases.increments(sch, 1, 700)
        sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[algid], S.EnsembleB200(device=0), trajectories=700,
                      noise_increments=dW, d_discontinuities=dd, **case.solve_kwargs())
        ref, _ = O.solve_ensemble(case.omodel, case.p, case.u0, case.h0, case.h1, case.lags, sch, algid, dW)
        assert sim.timing["n_steps"] == sch.n_steps > cases.c1(T=6.0, saveat=0.5).oracle_schedule().n_steps
        assert cases.scaled_err(sim.out, ref) < TOL


def test_out_rows_guard_rejects_undersized_buffers():
    """sdde_desc.out_rows: a caller that sized `out` for 2 save points but asks for save_everystep gets SDDE_ERR_BAD_DESC,
    not a heap overflow (the failure mode of a binding that computes n_save itself)."""
    import ctypes as C
    from sdde_b200 import _capi as capi
    from sdde_b200.ensemble import _marshal
    case = cases.c1(T=1.0, saveat=None)
    m = _marshal(case.problem(), S.linear1, S.EM(), dt=0.01, saveat=None, save_everystep=True, save_start=None, save_end=None, tstops=(),
                 adaptive=None, n_traj=64, traj_offset=0, seed=1, p_batch=None, u0_batch=None, noise=None,
                 noise_mode=capi.NOISE_PHILOX, noise_mem=0, out_mem=capi.MEM_HOST, want_paths=1, want_moments=0)
    out = np.zeros((2, 1, 64))
    m.desc.out_rows = 2
    rc = capi.lib().sdde_b200_solve(S.get_context(0).handle, C.byref(m.desc), capi.np_ptr(out), None, None, None)
    assert rc == -1 and b"out_rows" in capi.lib().sdde_b200_last_error()
    assert np.all(out == 0.0)
    m.desc.out_rows = 100                      # one row short of the 101 points save_everystep produces on [0, 1] with dt = 0.01
    big = np.zeros((101, 1, 64))
    assert capi.lib().sdde_b200_solve(S.get_context(0).handle, C.byref(m.desc), capi.np_ptr(big), None, None, None) == -1
    m.desc.out_rows = 101
    assert capi.lib().sdde_b200_solve(S.get_context(0).handle, C.byref(m.desc), capi.np_ptr(big), None, None, None) == 0
    assert np.all(big[0] == 1.0) and np.all(np.isfinite(big)) and np.all(big[-1] != 0.0)


def test_torch_out_must_be_float64_on_the_solve_device():
    import torch
    case = cases.c1(T=1.0, saveat=0.5)
    kw = dict(trajectories=128, seed=1, **case.solve_kwargs())
    with pytest.raises(ValueError, match="float64"):
        S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"),
                out=torch.empty((3, 1, 128), dtype=torch.float32, device="cuda:0"), **kw)
    with pytest.raises(ValueError, match="float64"):
        S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0),
                out=torch.empty((3, 1, 128), dtype=torch.float16).pin_memory(), **kw)


@pytest.mark.parametrize("hist", [1.0e3, 30.0, 1.0e-3])
@pytest.mark.parametrize("alg", [2, 3])
def test_near_cancellation_division_and_inverse_sqrt(hist, alg):
    """The two non-IEEE shortcuts of the parity path at their worst cases: div_pos (MUFU.RCP64H + cubic step instead of
    '/') with denominators 1 + x^10 from 1 + 1e-30 to 1e30, and RKMil's difference quotient (g(u~) - g(u)) * (1/sqrt dt)
    instead of / sqrt dt on steps of length 1e-14 and 3e-9 (user tstops), where g(u~) - g(u) cancels all but ~9 digits.
    Same 1e-12 bound against the oracle, which divides."""
    case = cases.c2(T=3.0, alg=alg).with_(h0=[hist], u0=[1.2])
    t1, t2 = 0.6200000000000001, 1.37
    ts = [t1, t1 + 1e-14, t2, t2 + 3e-9]
    import sdde_oracle as O
    sv, ss, se = O.expand_saveat(case.saveat, *case.tspan)
    sch = O.build_schedule(0.0, 3.0, case.dt, case.lags, alg, saveat_times=sv, save_start=ss, save_end=se, tstops=ts)
    assert np.min(sch.dt) < 2e-14 and np.sum(sch.dt < 1e-8) >= 2
    dW = cases.increments(sch, 1, 512)
    sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[alg], S.EnsembleB200(device=0), trajectories=512,
                  noise_increments=dW, tstops=ts, save_everystep=True, **{k: v for k, v in case.solve_kwargs().items() if k != "saveat"})
    sch_all = O.build_schedule(0.0, 3.0, case.dt, case.lags, alg, saveat_times=[], save_everystep=True, tstops=ts)
    ref, _ = O.solve_ensemble(case.omodel, case.p, case.u0, case.h0, case.h1, case.lags, sch_all, alg, dW)
    assert sim.out.shape == ref.shape and cases.relerr(sim.out, ref) < TOL


def test_noise_grid_regrid_more_than_65535_rows():
    """NoiseGrid on its own grid with more schedule nodes than gridDim.y allows (the regrid kernel indexes rows on grid.x)."""
    import sdde_oracle as O
    case = cases.c1(T=7.0, saveat=1.0).with_(dt=1e-4)
    sch = case.oracle_schedule()
    assert sch.n_steps > 65535
    n = 64
    rng = np.random.default_rng(9)
    ts = np.linspace(0.0, 7.0, 3001)
    Wg = np.cumsum(np.sqrt(np.diff(ts, prepend=0.0))[:, None, None] * rng.standard_normal((ts.size, 1, n)), axis=0)
    dW = O.noise_grid_increments(O.noise_grid_on_nodes(ts, Wg, sch.t))
    ref, _ = case.oracle(dW)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, noise_grid=(ts, Wg),
                  **case.solve_kwargs())
    assert cases.scaled_err(sim.out, ref) < TOL


def test_balanced_tail_round_gives_identical_paths():
    """The launch geometry (whole rounds of full CTAs + one round of smaller tail CTAs, option "balance") must not change a
    single bit: Philox mode (streams keyed by global id) on an ensemble with full AND tail CTAs, against full CTAs only."""
    case = cases.c2(T=10.0)
    n = 148 * 1024 + 3 * 1024 + 130          # one full round on 148 SMs + a tail round
    kw = dict(trajectories=n, seed=99, **case.solve_kwargs())
    ctx = S.get_context(0)
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), **kw)
    ctx.set_option("balance", 0)
    try:
        b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), **kw)
    finally:
        ctx.set_option("balance", 1)
    import torch
    assert torch.equal(a.out, b.out) and bool(torch.isfinite(a.out).all())
    ctx.set_option("lean", 0)
    try:
        c = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), **kw)
    finally:
        ctx.set_option("lean", 1)
    assert torch.equal(a.out, c.out)          # the lean batch bodies compute exactly what the general body computes


@pytest.mark.parametrize("alg", [5, 6])
def test_rkmil_commute_on_a_noise_matrix_matches_oracle(alg):
    """RKMilCommute (test/test_prob_sol.jl:44-49) on C4's 4 x 2 noise matrix (G_ij = M_ij u_i: commutative columns): one
    more G evaluation per noise column and step; per-path parity under shared increments."""
    case = cases.c4(T=8.0).with_(alg=alg)
    sim, ref, _ = run_case(case, 1536)
    assert "alg%d" % alg in sim.timing["kernel_name"]
    assert cases.scaled_err(sim.out, ref) < TOL


def test_rkmil_commute_on_diagonal_noise_is_the_rkmil_kernel():
    case = cases.c3(order0=0, T=30.0)
    sch = case.oracle_schedule()
    dW = cases.increments(sch, 1, 600)
    kw = dict(trajectories=600, noise_increments=dW, **case.solve_kwargs())
    for interp in ("Ito", "Stratonovich"):
        a = S.solve(S.EnsembleProblem(case.problem()), S.RKMil(interp), S.EnsembleB200(device=0), **kw)
        b = S.solve(S.EnsembleProblem(case.problem()), S.RKMilCommute(interp), S.EnsembleB200(device=0), **kw)
        assert np.array_equal(a.out, b.out)
    with pytest.raises(S.SddeError, match="diagonal"):          # plain RKMil still refuses a noise matrix, like the reference
        S.solve(S.EnsembleProblem(cases.c4(T=1.0).problem()), S.RKMil(), S.EnsembleB200(device=0), trajectories=64, dt=0.01,
                adaptive=False)


def test_rkmil_commute_strong_order_against_exact_solution_on_device():
    """du_i = a_i u_i dt + sum_j M_ij u_i dW_j has the exact solution u_i(0) exp((a_i - sum_j M_ij^2 / 2) t + sum_j M_ij W_j):
    the device RKMilCommute converges to it with strong order 1.0 +- 0.3 (the reference's band for Milstein-type solvers,
    test/analyticless_convergence_tests.jl:62-91)."""
    M = np.array([[0.5, 0.3], [-0.4, 0.6]])
    a = (-0.3, 0.2)
    model = S.linsys(2, 1, 2)
    p = S.linsys_params(np.diag(a), [np.zeros((2, 2))], np.zeros(2), M, np.zeros((2, 2)))
    n, T, fine = 600, 1.0, 2.0 ** -9
    rng = np.random.default_rng(12)
    dWf = rng.standard_normal((int(T / fine), 2, n)) * np.sqrt(fine)
    u0 = [1.0, 0.7]
    exact = np.array(u0)[:, None] * np.exp((np.array(a) - 0.5 * (M ** 2).sum(axis=1))[:, None] * T + M @ dWf.sum(axis=0))
    prob = S.SDDEProblem(model.f, model.g, u0, S.AffineHistory(tuple(u0), (0.0, 0.0)), (0.0, T), p, constant_lags=(1.0,))
    dts = 2.0 ** -np.arange(7, 2, -1)
    errs = []
    for dt in dts:
        k = int(round(dt / fine))
        dW = dWf.reshape(-1, k, 2, n).sum(axis=1)
        sim = S.solve(S.EnsembleProblem(prob), S.RKMilCommute(), S.EnsembleB200(device=0), trajectories=n, dt=dt, saveat=[T],
                      save_start=False, adaptive=False, noise_increments=dW)
        errs.append(np.mean(np.abs(sim.out[-1] - exact)))
    slope = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert abs(slope - 1.0) < 0.3, slope


def test_output_func_and_reduction_forms():
    """EnsembleProblem(output_func = FinalState() / SaveIdxs(...), reduction = MomentsReduction()): the forms of SciML's host
    closures this path can express map onto saveat = [tf], save_idxs and device moments."""
    case = cases.c4(T=2.0)
    base = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=512, seed=5, moments=True,
                   **case.solve_kwargs())
    fin = S.solve(S.EnsembleProblem(case.problem(), output_func=S.FinalState()), S.EM(), S.EnsembleB200(device=0), trajectories=512,
                  seed=5, dt=case.dt)
    assert fin.out.shape == (1, 4, 512) and np.array_equal(fin.out[0], base.out[-1]) and fin.t.tolist() == [2.0]
    sub = S.solve(S.EnsembleProblem(case.problem(), output_func=S.SaveIdxs([3, 1])), S.EM(), S.EnsembleB200(device=0), trajectories=512,
                  seed=5, **case.solve_kwargs())
    assert np.array_equal(sub.out, base.out[:, [3, 1], :])
    red = S.solve(S.EnsembleProblem(case.problem(), reduction=S.MomentsReduction()), S.EM(), S.EnsembleB200(device=0),
                  trajectories=512, seed=5, **case.solve_kwargs())
    assert red.out is None and np.allclose(red.moments, base.moments, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("alg", [7, 8, 9, 10, 11, 12])
def test_wangli3smil_paths_match_oracle(alg):
    """WangLi3SMil_A .. _F (test/test_prob_sol.jl:50-61), one device kernel with the stage order as an argument: per-path
    parity with the oracle's six orders under shared increments on the reference's Hayes problem (multiplicative noise with
    the delayed term) and on Mackey-Glass off the grid."""
    for case in (cases.hayes(T=3.0, dt=2.0 ** -6, beta=-1.2, saveat=0.25).with_(alg=alg), cases.c3(0, T=30.0).with_(alg=alg)):
        sch = case.oracle_schedule()
        dW = cases.increments(sch, 1, 700)
        sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[alg], S.EnsembleB200(device=0), trajectories=700,
                      noise_increments=dW, **case.solve_kwargs())
        ref, _ = case.oracle(dW)
        assert "alg%d" % alg in sim.timing["kernel_name"]
        assert cases.scaled_err(sim.out, ref) < TOL, case.name
    with pytest.raises(S.SddeError, match="diagonal"):
        S.solve(S.EnsembleProblem(cases.c4(T=1.0).problem()), cases.ALG_MAP[alg], S.EnsembleB200(device=0), trajectories=64, dt=0.01)


# ----------------------------------------------------------------------------------------------
# SRA1 / SRA3 / SOSRA / SRA(tableau): Roessler's methods for additive noise, fixed step (test/test_prob_sol.jl:74-87)
# ----------------------------------------------------------------------------------------------
def _sra_noise(sch, nc, n, key=13):
    """[n_steps, 2 nc, n]: dW rows, then dZ rows, both N(0, dt_n)."""
    return cases.increments(sch, 2 * nc, n, key=key)


SRA_CUSTOM = dict(c0=(0.0, 0.5), c1=(0.25, 1.0), A0=((0.0, 0.0), (0.5, 0.0)), B0=((0.0, 0.0), (1.0, 0.0)),
                  alpha=(0.25, 0.75), beta1=(0.5, 0.5), beta2=(1.0, -1.0))


@pytest.mark.parametrize("alg", [13, 14, 15, 16])
def test_sra_paths_match_oracle(alg):
    """Per-path parity under shared (dW, dZ): the reference's additive Hayes problem with a delayed term in g (beta != 0: g
    is evaluated at t + c1 dt with its own history lookup), lag on the grid, off the grid (every stage interpolates) and
    equal to dt (the t + dt stages read the node being left); and the additive scalar model of test/events.jl."""
    import sdde_oracle as O
    palg = S.SRA(tableau=SRA_CUSTOM) if alg == 16 else cases.ALG_MAP[alg]
    otab = O.SraTableau(**SRA_CUSTOM) if alg == 16 else None
    n = 700
    for lag, dt, T in ((1.0, 2.0 ** -6, 3.0), (0.33, 0.05, 2.0), (0.05, 0.05, 1.0)):
        case = cases.hayes(T=T, dt=dt, beta=0.6, saveat=0.25).with_(alg=alg, lags=[lag])
        case.p = np.array([lag, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1])
        sv, ss, se = O.expand_saveat(0.25, 0.0, T)
        sch = O.build_schedule(0.0, T, dt, [lag], alg, saveat_times=sv, save_start=ss, save_end=se, tableau=otab)
        nz = _sra_noise(sch, 1, n)
        sim = S.solve(S.EnsembleProblem(case.problem()), palg, S.EnsembleB200(device=0), trajectories=n,
                      noise_increments=nz, dt=dt, saveat=0.25, adaptive=False)
        ref, _ = O.solve_ensemble(case.omodel, case.p, case.u0, case.h0, case.h1, [lag], sch, alg, nz, tableau=otab)
        assert np.array_equal(sim.t, sch.save_t)
        assert "alg%d" % alg in sim.timing["kernel_name"]
        assert cases.scaled_err(sim.out, ref) < TOL, (alg, lag)
    case = cases.scalar_add(T=4.0).with_(alg=alg)
    sv, ss, se = O.expand_saveat(0.5, 0.0, 4.0)
    sch = O.build_schedule(0.0, 4.0, 0.01, case.lags, alg, saveat_times=sv, save_start=ss, save_end=se, tableau=otab)
    nz = _sra_noise(sch, 1, 257)
    sim = S.solve(S.EnsembleProblem(case.problem()), palg, S.EnsembleB200(device=0), trajectories=257, noise_increments=nz,
                  dt=0.01, saveat=0.5, adaptive=False)
    ref, _ = O.solve_ensemble(case.omodel, case.p, case.u0, case.h0, case.h1, case.lags, sch, alg, nz, tableau=otab)
    assert cases.scaled_err(sim.out, ref) < TOL


def test_sra_noise_grid_with_z_and_in_kernel_philox():
    """NoiseGrid(t, W, Z) on the schedule's nodes reproduces the increment run; with in-kernel Philox the two increments of
    step n are normals 2n and 2n + 1 of the trajectory's stream (the oracle draws the same stream)."""
    import sdde_oracle as O
    alg, n, T, dt = 13, 512, 2.0, 0.05
    case = cases.hayes(T=T, dt=dt, beta=0.6, saveat=0.5).with_(alg=alg)
    case.p = np.array([1.0, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1])
    sch = case.oracle_schedule()
    nz = _sra_noise(sch, 1, n)
    inc = S.solve(S.EnsembleProblem(case.problem()), S.SRA1(), S.EnsembleB200(device=0), trajectories=n, noise_increments=nz,
                  **case.solve_kwargs())
    WZ = np.concatenate([np.zeros((1, 2, n)), np.cumsum(nz, axis=0)], axis=0)
    grid = S.solve(S.EnsembleProblem(case.problem()), S.SRA1(), S.EnsembleB200(device=0), trajectories=n, noise_grid=WZ,
                   **case.solve_kwargs())
    assert cases.scaled_err(grid.out, inc.out) < 1e-11           # W[n+1] - W[n] re-derives dW to rounding
    seed = 0xB2000002
    phx = S.solve(S.EnsembleProblem(case.problem()), S.SRA1(), S.EnsembleB200(device=0), trajectories=n, seed=seed,
                  **case.solve_kwargs())
    z = O.philox_normals(np.arange(n, dtype=np.uint64), sch.n_steps, 2, seed)
    ref, _ = case.oracle(z * sch.sqdt[:, None, None])
    assert cases.scaled_err(phx.out, ref) < 1e-11
    with pytest.raises(ValueError, match=r"\[n, 2, "):
        S.solve(S.EnsembleProblem(case.problem()), S.SRA1(), S.EnsembleB200(device=0), trajectories=n, noise_increments=nz[:, :1],
                **case.solve_kwargs())
    with pytest.raises(S.SddeError, match="additive"):
        S.solve(S.EnsembleProblem(cases.c2(T=5.0).problem()), S.SRA1(), S.EnsembleB200(device=0), trajectories=64, dt=0.05, adaptive=False)


def test_large_host_output_picks_the_faster_of_staged_and_zero_copy():
    """Default (zero_copy = -1, auto): a large page-locked host output (>= 64 MB) runs staged on the first solve of its size class,
    zero-copy on the second and the faster of the two (device-timed) from then on; the three are bit-identical."""
    case = cases.c1(T=10.0, saveat=0.1)
    n = 120_000                                               # 101 rows x 120 000 x 8 B = 97 MB
    ctx = S.get_context(0)
    ctx.set_option("zero_copy", -1)                           # also forgets earlier trials
    prob = S.EnsembleProblem(case.problem())
    out = S.pinned_empty((101, 1, n))
    sims, how = [], []
    for i in range(4):
        sim = S.solve(prob, S.EM(), S.EnsembleB200(device=0), out=out, trajectories=n, seed=3, **case.solve_kwargs())
        sims.append(sim.out.copy()); how.append("mapped host out" in sim.timing["kernel_name"])
    assert how[0] is False and how[1] is True and how[2] == how[3]
    assert all(np.array_equal(sims[0], s) for s in sims[1:])


def test_golden_v2_fixtures_on_gpu():
    """The committed round-2 vectors (tests/golden/golden_v2.npz: EM(false), RKMil Stratonovich, EulerHeun, RKMilCommute on a noise
    matrix, WangLi3SMil_A / _F, SRA1 / SRA3 / SOSRA) against the device, 1e-12."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v2.npz"))
    mgc = lambda tau, alg: cases.c2(T=30.0, tau=tau, alg=alg)
    hay = lambda p, lag, dt, alg: cases.hayes(T=3.0, dt=dt, saveat=0.25).with_(alg=alg, lags=[lag], p=np.array(p))
    pmul, padd = [1.0, -4.0, -2.0, 10.0, -1.3, -1.2, 1.1], [0.33, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]
    table = {"mg_em_nosplit": mgc(17.0, 1), "mg_rkmil_strat": mgc(17.02, 3), "mg_euler_heun": mgc(17.0, 4),
             "hayes_wangli_a": hay(pmul, 1.0, 2.0 ** -6, 7), "hayes_wangli_f": hay(pmul, 1.0, 2.0 ** -6, 12),
             "hayes_sra1": hay(padd, 0.33, 0.05, 13), "hayes_sra3": hay(padd, 0.33, 0.05, 14), "hayes_sosra": hay(padd, 0.33, 0.05, 15),
             "c4_rkmil_commute": cases.c4(T=4.0).with_(alg=5)}
    for name, case in table.items():
        xi = g[name + "_xi"]
        sch = case.oracle_schedule()
        assert np.array_equal(sch.save_t, g[name + "_save_t"]), name
        sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0), trajectories=xi.shape[2],
                      noise_increments=xi * sch.sqdt[:, None, None], **case.solve_kwargs())
        assert np.array_equal(sim.t, g[name + "_save_t"]), name
        assert cases.scaled_err(sim.out, g[name + "_out"]) < TOL, name


def test_large_pageable_host_output_goes_through_the_bounce_path_and_is_identical():
    """A pageable host output >= 32 MB leaves through page-locked bounce buffers and host copy threads (several waves, rows cut
    into chunks); it must equal the page-locked output bit for bit -- also for a ragged trajectory count, a caller-supplied
    pageable array with a wider pitch than the wave, forced small waves, and with the path switched off (host_threads = 1)."""
    case = cases.c1(T=10.0, saveat=0.1)
    n = 90_001                                                # 101 rows x 90 001 x 8 B = 73 MB, ragged
    ctx = S.get_context(0)
    prob = S.EnsembleProblem(case.problem())
    kw = dict(trajectories=n, seed=4, **case.solve_kwargs())
    ref = S.solve(prob, S.EM(), S.EnsembleB200(device=0), out=S.pinned_empty((101, 1, n)), **kw)
    a = S.solve(prob, S.EM(), S.EnsembleB200(device=0), **kw)                            # pageable, allocated by solve (untouched pages)
    pre = np.full((101, 1, n), np.nan)
    b = S.solve(prob, S.EM(), S.EnsembleB200(device=0, wave_traj=20_000), out=pre, **kw)   # several waves
    ctx.set_option("wave_traj", 0)
    ctx.set_option("host_threads", 1)
    try:
        c = S.solve(prob, S.EM(), S.EnsembleB200(device=0), **kw)
    finally:
        ctx.set_option("host_threads", 0)
    ctx.set_option("host_threads", 3)
    try:
        d = S.solve(prob, S.EM(), S.EnsembleB200(device=0), **kw)
    finally:
        ctx.set_option("host_threads", 0)
    for x in (a, b, c, d):
        assert np.array_equal(x.out, ref.out) and np.array_equal(x.status, ref.status)
    assert b.timing["waves"] >= 4 and b.out is pre
