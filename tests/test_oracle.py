"""CPU: the oracle against hand-derived known answers, the Random123 Philox known-answer
vectors, the committed golden fixtures, and the C++ restatement against the numpy restatement (bit for bit)."""
import math
import os

import numpy as np
import pytest

import c_oracle
import cases
import sdde_oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_v1.npz"))


def test_philox_known_answer_vectors():
    # Random123 kat_vectors for philox4x32_10
    def kat(c, k):
        r = O.philox4x32_10(*[np.array([x]) for x in c], k[0], k[1])
        return [int(x[0]) for x in r]
    assert kat([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_normals_golden_and_moments():
    z = O.philox_normals(GOLD["philox_ids"], 6, 3, 0xB2000001)
    assert np.array_equal(z, GOLD["philox_z"])
    z = O.philox_normals(np.arange(40000), 10, 1, 99)
    assert abs(z.mean()) < 4 / np.sqrt(z.size) and abs(z.var() - 1) < 6 * np.sqrt(2 / z.size)
    assert abs((z ** 4).mean() - 3) < 0.1


def _one_step(alg):
    m = O.Model(99, 1, 1, 1, True, 0, lambda u, ul, p, t: u, lambda u, ul, p, t: u)
    s = O.Schedule(t=np.array([0, 0.1]), dt=np.array([0.1]), sqdt=np.sqrt(np.array([0.1])),
                   lk_mode=np.zeros((1, 1), np.int32), lk_idx=-np.ones((1, 1), np.int64), lk_theta=np.zeros((1, 1)),
                   lk_time=np.zeros((1, 1)), save_t=np.array([0.1]), save_step=np.array([0]),
                   save_theta=np.array([1.0]), save_exact=np.array([1], np.int32))
    out, _ = O.solve_ensemble(m, [], [1.0], [0.0], [0.0], [1.0], s, alg, np.full((1, 1, 1), 0.2))
    return out[0, 0, 0]


def test_k1_one_step_known_answers():
    # f=u, g=u, u=1, dt=0.1, dW=0.2
    assert _one_step(O.ALG_EM_SPLIT) == pytest.approx(1.32, rel=1e-15)
    assert _one_step(O.ALG_EM_NOSPLIT) == pytest.approx(1.30, rel=1e-15)
    sq = np.sqrt(0.1)
    ggp = ((1.1 + sq) - 1.0) / sq
    assert _one_step(O.ALG_RKMIL_ITO) == pytest.approx(1.1 + 0.2 + ggp * (0.04 - 0.1) / 2, rel=1e-15)
    assert _one_step(O.ALG_RKMIL_ITO) == pytest.approx(1.260513167019495, rel=1e-14)
    assert _one_step(O.ALG_RKMIL_STRAT) == pytest.approx(1.32, rel=1e-15)


def test_k2_zero_noise_is_explicit_euler_for_the_dde():
    c = cases.c1(T=3.0, saveat=None)
    sch = c.oracle_schedule()
    out, _ = c.oracle(np.zeros((sch.n_steps, 1, 1)))
    a, b = -1.0, 0.5
    u = [1.0]
    for n in range(sch.n_steps):
        tl = sch.t[n] - 1.0
        if tl < 0:
            ul = 1.0
        else:
            i = int(sch.lk_idx[n, 0]); th = sch.lk_theta[n, 0]
            ul = u[i] if sch.lk_mode[n, 0] == O.LK_EXACT else (1 - th) * u[i - 1] + th * u[i]
        u.append(u[n] + sch.dt[n] * (a * u[n] + b * ul))
    assert np.allclose(out[:, 0, 0], np.array(u), rtol=1e-14)


def test_k3_exact_mean_recurrence_matches_ensemble_mean():
    c = cases.c1(T=4.0, saveat=None)
    sch = c.oracle_schedule()
    n = 20000
    dW = O.philox_normals(np.arange(n), sch.n_steps, 1, 5) * sch.sqdt[:, None, None]
    out, _ = c.oracle(dW)
    m = O.exact_mean_linear1(sch, -1.0, 0.5, 1.0, 1.0, 0.0, 1.0)
    se = out[:, 0, :].std(axis=1) / np.sqrt(n)
    assert np.all(np.abs(out[:, 0, :].mean(axis=1) - m)[1:] < 5 * se[1:] + 1e-12)


def test_k4_schedule_facts():
    s = O.build_schedule(0, 100, 0.05, [17.0], O.ALG_EM_SPLIT)
    assert s.n_steps == 2001 and s.tstops_hit == [17.0, 34.0, 100.0]
    assert s.dt[-1] == pytest.approx(3.75e-12, rel=0.01)
    short = [i for i in range(s.n_steps) if abs(s.dt[i] - 0.05) > 1e-15]
    assert short == [339, 679, 2000]
    assert int(np.sum(s.lk_mode[:, 0] == O.LK_EXACT)) == 3          # "3 exact of 1661"
    assert int(np.sum(s.lk_mode[:, 0] != O.LK_PRE)) == 1661
    s = O.build_schedule(0, 10, 0.01, [1.0], O.ALG_EM_SPLIT)
    assert s.n_steps == 1001 and s.dt[-1] == pytest.approx(1.7e-13, rel=0.01)
    s = O.build_schedule(0, 500, 0.05, [17.0], O.ALG_EM_SPLIT)
    assert s.n_steps == 10000
    s = O.build_schedule(0, 100, 0.05, [17.02], O.ALG_RKMIL_ITO)
    assert s.n_steps == 2002 and s.tstops_hit == [17.02, 34.04, 51.06, 100.0]
    s2 = O.build_schedule(0, 100, 0.05, [17.02], O.ALG_EM_SPLIT)
    assert s2.tstops_hit == [17.02, 34.04, 100.0]                   # EM tracks {tau}, {2 tau} only
    s = O.build_schedule(0, 100, 0.05, [17.02], O.ALG_RKMIL_ITO, order_discontinuity_t0=2)
    assert s.n_steps == 2001 and s.tstops_hit == [100.0] and s.dt[-1] == pytest.approx(3.5e-12, rel=0.02)
    assert s.lk_theta[400, 0] == pytest.approx(0.6, abs=1e-9)
    s = O.build_schedule(0, 20, 0.01, [1.0, 2.5], O.ALG_EM_SPLIT)
    assert s.n_steps == 2003 and s.tstops_hit == [1.0, 2.0, 2.5, 3.5, 5.0, 20.0]
    tiny = [s.dt[i] for i in range(s.n_steps) if s.dt[i] < 1e-10]
    assert len(tiny) == 3 and all(1e-15 < x < 1e-13 for x in tiny)


def test_k5_history_edge():
    # lookup at exactly t0 returns u0 even if h(p,t0) != u0; at t0 - eps returns h
    m = O.get_model(O.MODEL_LINEAR1)
    sch = O.build_schedule(0.0, 1.5, 0.5, [1.0], O.ALG_EM_SPLIT, save_everystep=True)
    assert sch.lk_mode[2, 0] == O.LK_EXACT and sch.lk_idx[2, 0] == 0          # t=1.0 - 1.0 == t0
    out, _ = O.solve_ensemble(m, [0.0, 1.0, 0.0], [2.0], [5.0], [0.0], [1.0], sch, O.ALG_EM_SPLIT,
                              np.zeros((sch.n_steps, 1, 1)))
    # f = u(t-1): steps 0,1 read h = 5, step 2 reads u0 = 2
    assert np.allclose(out[:, 0, 0], [2.0, 4.5, 7.0, 8.0])


def test_k7_saveat_alias_hazard():
    # H1: with saveat > tau the reference's delayed state is u0 once past the last saved point
    c = cases.c1(T=6.0, saveat=2.0)
    sch = c.oracle_schedule()
    dW = np.zeros((sch.n_steps, 1, 1))
    dense, _ = c.oracle(dW)
    alias, _ = c.oracle(dW, history_mode="reference_saveat_alias")
    assert not np.allclose(dense, alias, rtol=1e-6)
    # in alias mode with zero noise the delayed value is always u0=1 (prehistory 1, saved points only at 0,2,4,6 and
    # lookups beyond the last saved point return u0): u' = -u + 0.5 -> fixed point 0.5
    u = 1.0
    for n in range(sch.n_steps):
        u = u + sch.dt[n] * (-1.0 * u + 0.5 * 1.0)
        if abs(sch.t[n + 1] - 2.0) < 1e-12:
            assert alias[1, 0, 0] == pytest.approx(u, rel=1e-9)


def test_golden_fixtures_reproduce():
    for name, tau, alg, order0 in [("c2_em", 17.0, 0, 0), ("c3_rkmil_o0", 17.02, 2, 0), ("c3_rkmil_o2", 17.02, 2, 2)]:
        sv = O.expand_saveat(1.0, 0.0, 40.0)[0]
        sch = O.build_schedule(0.0, 40.0, 0.05, [tau], alg, order_discontinuity_t0=order0, saveat_times=sv)
        assert np.array_equal(sch.t, GOLD[name + "_t"]) and np.array_equal(sch.dt, GOLD[name + "_dt"])
        out, _ = O.solve_ensemble(O.get_model(O.MODEL_MACKEY_GLASS), [0.2, 0.1, 10.0, 0.1], [1.2], [1.2], [0.0], [tau],
                                  sch, alg, GOLD[name + "_xi"] * sch.sqdt[:, None, None])
        assert np.array_equal(out, GOLD[name + "_out"])


@pytest.mark.parametrize("mk,n", [
    (lambda: cases.c1(), 200), (lambda: cases.c1(saveat=None), 50), (lambda: cases.c2(T=60.0), 128),
    (lambda: cases.c2(T=60.0, alg=1), 64), (lambda: cases.c3(0, T=60.0), 128), (lambda: cases.c3(2, T=60.0, alg=3), 128),
    (lambda: cases.c4(T=8.0), 64), (lambda: cases.hayes(T=3.0, beta=-1.2, saveat=0.25), 64),
    (lambda: cases.sir(T=20.0, dt=0.02), 64), (lambda: cases.scalar_add(), 64),
    (lambda: cases.c2(T=60.0, alg=4), 64), (lambda: cases.c4(T=6.0).with_(alg=4), 32), (lambda: cases.sir(T=10.0, dt=0.02).with_(alg=4), 32),
    (lambda: cases.hayes(T=3.0, beta=-1.2, saveat=0.25, alg=4), 32)])
def test_cpp_restatement_equals_numpy_restatement_bitwise(mk, n):
    c = mk()
    sch = c.oracle_schedule()
    dW = cases.increments(sch, c.model.n_chan, n)
    ref, st = c.oracle(dW)
    sv, ss, se = O.expand_saveat(c.saveat, *c.tspan)
    out, st2, _, steps = c_oracle.solve(c.omodel.mid, c.alg, c.omodel.n_state, c.omodel.n_lag, c.omodel.n_noise, c.p,
                                        c.u0, c.h0, c.h1, c.lags, c.tspan[0], c.tspan[1], c.dt, sv, len(sch.save_t),
                                        save_start=ss, save_end=se, save_everystep=c.saveat is None, order0=c.order0,
                                        declared=c.declared, n_traj=n, dW=dW, n_threads=3)
    assert steps == sch.n_steps
    assert np.array_equal(out, ref) and np.array_equal(st, st2)


def test_cpp_philox_stream_matches_numpy():
    c = cases.c1(T=2.0)
    sch = c.oracle_schedule()
    n = 64
    z = O.philox_normals(np.arange(1000, 1000 + n), sch.n_steps, 1, 0xB2000001)
    ref, _ = c.oracle(z * sch.sqdt[:, None, None])
    sv, ss, se = O.expand_saveat(c.saveat, *c.tspan)
    out, _, _, _ = c_oracle.solve(c.omodel.mid, 0, 1, 1, 1, c.p, c.u0, c.h0, c.h1, c.lags, 0.0, 2.0, c.dt, sv,
                                  len(sch.save_t), n_traj=n, traj_offset=1000, seed=0xB2000001)
    assert cases.scaled_err(out, ref) < 1e-12      # libm vs numpy log/sin/cos differ by ulps


def test_moment_merge_is_associative_enough():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 2, 1000)) * 5 + 100
    full = O.moments(x)
    a, b, c = O.moments(x[..., :100]), O.moments(x[..., 100:550]), O.moments(x[..., 550:])
    m = O.merge_moments(O.merge_moments(a, b), c)
    assert np.allclose(m, full, rtol=1e-12)


def test_euler_heun_known_answer_and_order():
    # one step, f = u, g = u, u = 1, dt = 0.1, dW = 0.2 (no delay dependence): tmp = 1 + 0.1 + 0.2 = 1.3,
    # u+ = 1 + 0.05 (1 + 1.3) + 0.5 (1 + 1.3) 0.2 = 1.345
    assert _one_step(O.ALG_EULER_HEUN) == pytest.approx(1.345, rel=1e-15)
    s = O.build_schedule(0, 5, 0.01, [1.0], O.ALG_EULER_HEUN)
    assert s.lk_mode.shape == (501, 2) and s.tstops_hit == [1.0, 2.0, 5.0]          # alg_order 1/2 like EM
    # the second-stage lookup of step n is (numerically almost) the first-stage lookup of step n+1
    assert np.all(np.abs(s.lk_time[:-1, 1] - s.lk_time[1:, 0]) < 1e-12)


def test_oracle_quantile_definition_known_answers():
    """Statistics.quantile (alpha = beta = 1) hand-computed: v = 1..5 -> q(0.1) = 1.4, q(0.5) = 3, q(1) = 5;
    v = 1..4 -> median 2.5, q(0.25) = 1.75."""
    v5, v4 = np.array([5.0, 1.0, 4.0, 2.0, 3.0]), np.array([4.0, 1.0, 3.0, 2.0])
    assert O.julia_quantile(v5, 0.1) == 1.4 and O.julia_quantile(v5, 0.5) == 3.0 and O.julia_quantile(v5, 1.0) == 5.0
    assert O.julia_quantile(v5, 0.0) == 1.0 and O.julia_median(v5) == 3.0
    assert O.julia_median(v4) == 2.5 and O.julia_quantile(v4, 0.25) == 1.75
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1001)
    for p in (0.05, 0.3, 0.95):
        assert abs(O.julia_quantile(x, p) - np.quantile(x, p)) < 1e-14      # numpy 'linear' = the same type-7 definition


def test_oracle_noise_grid_interpolation_known_answers():
    """NoiseGrid lookups: exact nodes return the stored row, interior times the linear interpolant, the last node snaps."""
    ts = np.array([0.0, 0.5, 2.0])
    W = np.array([0.0, 1.0, 4.0]).reshape(3, 1, 1)
    got = O.noise_grid_on_nodes(ts, W, [0.0, 0.25, 0.5, 1.25, 2.0 - 2e-16])[:, 0, 0]
    assert np.array_equal(got, [0.0, 0.5, 1.0, 2.5, 4.0])
    assert np.array_equal(O.noise_grid_increments(W)[:, 0, 0], [1.0, 3.0])
    with pytest.raises(ValueError):
        O.noise_grid_on_nodes(ts, W, [2.5])


def test_oracle_sort_is_julias_total_order():
    v = np.array([0.0, -0.0, 1.0, -0.0, 0.0, -1.0])
    s = O._julia_sort(v)
    assert np.array_equal(np.signbit(s), [True, True, True, False, False, False]) and s[0] == -1.0 and s[-1] == 1.0
    # the median of an odd count is an order statistic itself: [-1, -0, -0, +0, 1] -> -0.0
    assert np.signbit(O.julia_median(np.array([0.0, -0.0, 1.0, -0.0, -1.0]))) and not np.signbit(O.julia_median(np.array([0.0, -0.0, 1.0, 0.0, -1.0])))


def test_export_for_julia_kit_is_consistent(tmp_path):
    """The hand-off kit for pinning the oracle against the real reference (tests/golden/export_for_julia.py ->
    julia/replay_fixtures.jl): the exported dense solutions restricted to the saveat grid are the committed golden vectors,
    the Brownian paths start at 0 and the manifest lists every array the Julia script reads."""
    import subprocess, sys, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([sys.executable, os.path.join(root, "tests", "golden", "export_for_julia.py"), str(tmp_path)], check=True,
                   capture_output=True)
    man = open(tmp_path / "manifest.toml").read()
    names = re.findall(r"^\[(\w+)\]", man, re.M)
    assert {"c2_em", "c3_rkmil_o0", "c3_rkmil_o2", "c4_em", "c2_euler_heun", "sra1_hayes", "sra3_hayes", "sosra_hayes"} <= set(names)
    assert os.path.getsize(tmp_path / "sosra_hayes_Z.bin") == os.path.getsize(tmp_path / "sosra_hayes_W.bin")
    g = np.load(os.path.join(root, "tests", "golden", "golden_v1.npz"))
    for name in names:
        nn = int(re.search(r"\[%s\].*?n_nodes = (\d+)" % name, man, re.S).group(1))
        ns = int(re.search(r"\[%s\].*?n_state = (\d+)" % name, man, re.S).group(1))
        nc = int(re.search(r"\[%s\].*?n_chan = (\d+)" % name, man, re.S).group(1))
        t = np.fromfile(tmp_path / f"{name}_t.bin")
        W = np.fromfile(tmp_path / f"{name}_W.bin").reshape(16, nc, nn)
        U = np.fromfile(tmp_path / f"{name}_dense.bin").reshape(16, ns, nn)
        assert t.size == nn and np.all(np.diff(t) > 0) and np.all(W[:, :, 0] == 0.0) and np.all(np.isfinite(U))
        if name + "_out" in g.files and name != "c4_em":
            # golden: saveat = 1.0 on [0, 40]; the save times are schedule nodes here, so the dense solution contains them
            idx = [int(np.argmin(np.abs(t - k))) for k in range(41)]
            if np.max(np.abs(t[idx] - np.arange(41))) > 1e-9:
                continue                     # re-phased grid (tau = 17.02, order 0): the integer times are interpolated saves
            # (not bit-equal: the kit feeds W = cumsum(dW) and the increments are re-derived from it, as the reference does)
            assert np.allclose(U[:, :, idx].transpose(2, 1, 0), g[name + "_out"], rtol=1e-10, atol=0)
    jl = open(os.path.join(root, "julia", "replay_fixtures.jl")).read()
    for key in ("n_nodes", "n_state", "n_chan", "has_z", "alg_option", "order_discontinuity_t0", "_dense.bin", "_W.bin", "_Z.bin", "_t.bin"):
        assert key in jl and (key.startswith("_") or key in man)


def _linsys_2x2(M, a=(0.0, 0.0)):
    """2 states, 1 lag (unused: B = 0), 2 noise channels: du_i = a_i u_i dt + sum_j M_ij u_i dW_j  (commutative columns)."""
    A = np.diag(a)
    p = np.concatenate([A.ravel(), np.zeros(4), np.zeros(2), np.asarray(M, float).ravel(), np.zeros(4)])
    return O.get_model(O.MODEL_LINSYS, 2, 1, 2), p


def test_k1_rkmil_commute_one_step_known_answer():
    """RKMilCommute on a 2 x 2 noise matrix, one step, hand-derived (see the arithmetic below): f = 0, u = (1, 2),
    G = [[0.1 u1, 0.2 u1], [0.3 u2, 0]], dt = 0.25, dW = (0.4, -0.2).
      G = [[.1,.2],[.6,0]];  K_1 = u + .5 G[:,1] = (1.05, 2.3) -> Dg_1 = [[.01,.02],[.18,0]];  K_2 = (1.1, 2) -> Dg_2 = [[.02,.04],[0,0]]
      Ito: J = dW dW'/2 - dt/2 I = [[-.045,-.04],[-.04,-.105]];  mil = Dg_1 J[:,1] + Dg_2 J[:,2] = (-.00625, -.0081)
      u+ = u + G dW + mil = (1 + 0 - .00625, 2 + .24 - .0081) = (0.99375, 2.2319)
      Stratonovich: J = dW dW'/2 -> mil = (0, .0144) -> u+ = (1.0, 2.2544)
    The same numbers follow from the analytic Milstein term sum_{j1,j2} M_ij1 M_ij2 u_i J[j2,j1] of this G."""
    model, p = _linsys_2x2([[0.1, 0.2], [0.3, 0.0]])
    dW = np.array([0.4, -0.2]).reshape(1, 2, 1)
    for alg, want in ((O.ALG_RKMILC_ITO, (0.99375, 2.2319)), (O.ALG_RKMILC_STRAT, (1.0, 2.2544))):
        sch = O.build_schedule(0.0, 0.25, 0.25, [1.0], alg, saveat_times=[], save_everystep=True)
        assert sch.n_steps == 1
        out, st = O.solve_ensemble(model, p, [1.0, 2.0], [1.0, 2.0], [0.0, 0.0], [1.0], sch, alg, dW)
        assert np.allclose(out[-1, :, 0], want, rtol=0, atol=2e-15), (alg, out[-1, :, 0])
        co, _, _, _ = __import__("c_oracle").solve(O.MODEL_LINSYS, alg, 2, 1, 2, p, [1.0, 2.0], [1.0, 2.0], [0.0, 0.0], [1.0], 0.0, 0.25, 0.25,
                                                   [], 2, save_everystep=True, dW=dW)
        assert np.array_equal(co, out)


def test_rkmil_commute_strong_order_one_against_exact_solution():
    """The commutative Milstein algebra itself, independent of any recollection of the reference's code: for
    du_i = a_i u_i dt + sum_j M_ij u_i dW_j the exact solution is u_i(0) exp((a_i - sum_j M_ij^2/2) t + sum_j M_ij W_j(t))
    (Ito) resp. exp(a_i t + sum_j M_ij W_j(t)) (Stratonovich); the scheme must converge to it with strong order 1.0
    (the reference's own acceptance band for its Milstein-type solvers: 1.0 +- 0.3, test/analyticless_convergence_tests.jl:62-91),
    where EM on the same paths shows 0.5."""
    M = np.array([[0.5, 0.3], [-0.4, 0.6]])
    a = (-0.3, 0.2)
    model, p = _linsys_2x2(M, a)
    n, T, fine = 400, 1.0, 2.0 ** -9
    rng = np.random.default_rng(11)
    dWf = rng.standard_normal((int(T / fine), 2, n)) * np.sqrt(fine)
    WT = dWf.sum(axis=0)
    u0 = np.array([1.0, 0.7])

    def err(alg, dt, exact):
        sch = O.build_schedule(0.0, T, dt, [1.0], alg, saveat_times=[T], save_start=False)
        k = int(round(dt / fine))
        dW = dWf.reshape(sch.n_steps, k, 2, n).sum(axis=1)
        out, _ = O.solve_ensemble(model, p, u0, u0, [0.0, 0.0], [1.0], sch, alg, dW)
        return np.mean(np.abs(out[-1] - exact))

    drift_ito = (np.array(a) - 0.5 * (M ** 2).sum(axis=1))[:, None] * T
    ex_ito = u0[:, None] * np.exp(drift_ito + M @ WT)
    ex_str = u0[:, None] * np.exp(np.array(a)[:, None] * T + M @ WT)
    dts = 2.0 ** -np.arange(7, 2, -1)
    for alg, exact, want in ((O.ALG_RKMILC_ITO, ex_ito, 1.0), (O.ALG_RKMILC_STRAT, ex_str, 1.0), (O.ALG_EM_SPLIT, ex_ito, 0.5)):
        slope = np.polyfit(np.log(dts), np.log([err(alg, dt, exact) for dt in dts]), 1)[0]
        assert abs(slope - want) < 0.3, (alg, slope)


def test_k1_wangli3smil_one_step_known_answers():
    """WangLi3SMil_A..F, one step of the Hayes test problem, against the composition of the three stage operators written
    out with plain floats (f = a u + b x_tau + c and g = alpha u + beta x_tau + gamma are affine in u, so the six orders give
    six different numbers; x_tau = h(p, -1) = 0 on the first step):
       D(x) = x + dt (a x + c);   C(x) = x - dt/2 * alpha (alpha x + gamma);   M(x) = x + (alpha x + gamma) dW + dW^2/2 * alpha (alpha x + gamma)
    where alpha (alpha x + gamma) is what the forward difference (g(x + g(x) sqrt(dt)) - g(x)) / sqrt(dt) gives for an affine g."""
    a, b, c, al, be, ga = -4.0, -2.0, 10.0, -1.3, -1.2, 1.1
    dt, dW = 0.1, 0.2
    D = lambda x: x + dt * (a * x + c)
    C = lambda x: x - (dt / 2) * (al * (al * x + ga))
    Mm = lambda x: x + (al * x + ga) * dW + (dW * dW / 2) * (al * (al * x + ga))
    ops = {"D": D, "C": C, "M": Mm}
    model = O.get_model(O.MODEL_HAYES)
    p = [1.0, a, b, c, al, be, ga]
    got = {}
    for alg, order in O.WANGLI_ORDER.items():
        x = 1.0
        for o in order:
            x = ops[o](x)
        sch = O.build_schedule(0.0, dt, dt, [1.0], alg, saveat_times=[], save_everystep=True)
        out, _ = O.solve_ensemble(model, p, [1.0], [1.0], [1.0], [1.0], sch, alg, np.full((1, 1, 1), dW))
        assert abs(out[-1, 0, 0] - x) < 5e-15, (order, out[-1, 0, 0], x)
        co, _, _, _ = __import__("c_oracle").solve(O.MODEL_HAYES, alg, 1, 1, 1, p, [1.0], [1.0], [1.0], [1.0], 0.0, dt, dt, [], 2,
                                                   save_everystep=True, dW=np.full((1, 1, 1), dW))
        assert np.array_equal(co, out)
        got[order] = out[-1, 0, 0]
    # C and M are affine maps with the same fixed point -gamma/alpha for an affine g, so they commute: 4 distinct values here
    assert len({round(v, 12) for v in got.values()}) == 4 and abs(got["DCM"] - got["DMC"]) < 1e-15 and abs(got["CMD"] - got["MCD"]) < 1e-15
    # all of them are Milstein to leading order: u + f dt + g dW + g g' (dW^2 - dt)/2 at u = 1
    mil = 1.0 + dt * (a + c) + (al + ga) * dW + al * (al + ga) * (dW * dW - dt) / 2
    assert all(abs(v - mil) < 0.25 for v in got.values())       # dt = 0.1 with a = -4 is coarse: O(dt^1.5..2) apart


def test_wangli3smil_strong_order_one_on_the_reference_convergence_problem():
    """test/analyticless_convergence_tests.jl:62-91 asserts strong order 1.0 +- 0.3 for WangLi3SMil_A..F on the Hayes problem
    (dts = 2^-8..2^-4 against a 2^-9 run, shared Brownian path); the same estimate with the oracle for two of the six orders
    (all six run on the device, tests/test_gpu_rng.py)."""
    n, T, fine = 100, 4.0, 2.0 ** -9
    rng = np.random.default_rng(2027)
    dWf = rng.standard_normal((int(round(T / fine)), 1, n)) * np.sqrt(fine)
    model = O.get_model(O.MODEL_HAYES)
    p = [1.0, -4.0, -2.0, 10.0, -1.3, 0.0, 1.1]

    def run(alg, dt):
        sch = O.build_schedule(0.0, T, dt, [1.0], alg, saveat_times=[T], save_start=False)
        k = int(round(dt / fine))
        out, _ = O.solve_ensemble(model, p, [1.0], [1.0], [1.0], [1.0], sch, alg, dWf.reshape(sch.n_steps, k, 1, n).sum(axis=1))
        return out[-1, 0]

    dts = 2.0 ** -np.arange(8, 3, -1)
    for alg in (O.ALG_WANGLI_A, O.ALG_WANGLI_F):
        ref = run(alg, fine)
        slope = np.polyfit(np.log(dts), np.log([np.mean(np.abs(run(alg, dt) - ref)) for dt in dts]), 1)[0]
        assert abs(slope - 1.0) < 0.3, (alg, slope)


# ----------------------------------------------------------------------------------------------
# SRA1 / SRA3 / SOSRA (Roessler's stochastic Runge-Kutta methods for additive noise, fixed step)
# ----------------------------------------------------------------------------------------------
def test_sra_tableaus_satisfy_the_order_conditions():
    """The eight conditions of strong order 1.5 for additive noise (Roessler 2010, SRA family) hold for the three restated
    tableaus to rounding: sum alpha = 1, sum beta1 = 1, sum beta2 = 0, beta1.c1 = 1, beta2.c1 = -1, alpha.(B0 e) = 1,
    alpha.(A0 e) = 1/2, alpha.(B0 e)^2 = 3/2, and the row sums A0 e = c0.  (They caught a transposed digit pair in the
    recollected SOSRA alpha_3.)"""
    for alg, t in O.SRA_TABLEAUS.items():
        al, b1, b2, c0, c1 = (np.array(x) for x in (t.alpha, t.beta1, t.beta2, t.c0, t.c1))
        A0, B0, e = np.array(t.A0), np.array(t.B0), np.ones(t.stages)
        res = [al.sum() - 1, b1.sum() - 1, b2.sum(), b1 @ c1 - 1, b2 @ c1 + 1, al @ (B0 @ e) - 1, al @ (A0 @ e) - 0.5,
               al @ (B0 @ e) ** 2 - 1.5, *(A0 @ e - c0)]
        assert max(abs(r) for r in res) < 4e-15, (alg, res)
        assert np.all(np.triu(A0) == 0) and np.all(np.triu(B0) == 0)           # explicit
        assert len(t.stage_sets()[0]) <= O.SRA_MAX_SETS


def test_k1_sra1_one_step_known_answer():
    """One SRA1 step of the reference's additive Hayes problem (test/test_prob_sol.jl:19,74-79, with beta != 0 so that g sees
    the delayed state) written out with plain floats the way the reference's hand-unrolled SRA1 perform_step! does it:
        gpdt = g(t + dt);  chi2 = (dW + dZ/sqrt 3)/2;  k1 = dt f(u, t);  k2 = dt f(u + 3 k1/4 + 3 chi2 gpdt/2, t + 3 dt/4)
        u+ = u + k1/3 + 2 k2/3 + chi2 (g(t) - gpdt) + dW gpdt
    Every f / g call looks its delayed state up at its own time: h(p, s) = 1 + s for s < 0."""
    a, b, c, be, ga = -4.0, -2.0, 10.0, 0.7, 0.1
    dt, dW, dZ, u = 0.1, 0.2, -0.15, 1.0
    x = lambda s: 1.0 + (s - 1.0)                                  # delayed state h(p, s - tau), tau = 1
    f = lambda v, s: a * v + b * x(s) + c
    g = lambda s: be * x(s) + ga
    gpdt, chi2 = g(dt), 0.5 * (dW + dZ / math.sqrt(3.0))
    k1 = dt * f(u, 0.0)
    k2 = dt * f(u + 3 * k1 / 4 + 3 * chi2 * gpdt / 2, 0.75 * dt)
    want = u + k1 / 3 + 2 * k2 / 3 + chi2 * (g(0.0) - gpdt) + dW * gpdt
    p = [1.0, a, b, c, 0.0, be, ga]
    noise = np.array([dW, dZ]).reshape(1, 2, 1)
    sch = O.build_schedule(0.0, dt, dt, [1.0], O.ALG_SRA1, saveat_times=[], save_everystep=True)
    assert sch.lk_mode.shape == (1, O.SRA_MAX_SETS) and np.all(sch.lk_mode == O.LK_PRE)
    assert np.allclose(sch.lk_theta[0], [-1.0, dt - 1.0, 0.75 * dt - 1.0, -1.0], rtol=0, atol=1e-16)   # sets: t, t+dt, t+3dt/4, pad
    out, _ = O.solve_ensemble(O.get_model(O.MODEL_HAYES), p, [u], [1.0], [1.0], [1.0], sch, O.ALG_SRA1, noise)
    assert abs(out[-1, 0, 0] - want) < 5e-15, (out[-1, 0, 0], want)
    co, _, _, _ = c_oracle.solve(O.MODEL_HAYES, O.ALG_SRA1, 1, 1, 1, p, [u], [1.0], [1.0], [1.0], 0.0, dt, dt, [], 2,
                                 save_everystep=True, dW=noise)
    assert np.array_equal(co, out)


def _coarsen_wz(dWf, dZf, hf, k):
    """Fine (dW, dZ) -> coarse (dW, dZ) over k fine steps of length hf, keeping the SAME Brownian path: the iterated
    integral I_(1,0) = int (W_s - W_t) ds = (h/2)(dW + dZ/sqrt 3) adds up as sum_k [I10_k + (W_k - W_0) hf]."""
    n = dWf.shape[0] // k
    w = dWf.reshape(n, k, *dWf.shape[1:])
    z = dZf.reshape(n, k, *dZf.shape[1:])
    i10f = 0.5 * hf * (w + z / math.sqrt(3.0))
    before = np.cumsum(w, axis=1) - w                               # W at the start of each fine step, relative to the coarse start
    i10 = (i10f + before * hf).sum(axis=1)
    dWc = w.sum(axis=1)
    h = hf * k
    return dWc, math.sqrt(3.0) * (2.0 * i10 / h - dWc)


@pytest.mark.parametrize("alg", [O.ALG_SRA1, O.ALG_SRA3, O.ALG_SOSRA])
def test_sra_strong_order_on_the_additive_hayes_problem(alg):
    """Strong order of the fixed-step SRA methods on the reference's additive Hayes problem (padd, test/test_prob_sol.jl:19)
    against a 2^-10 run of the same method on the same Brownian path (dW and the iterated integral coarsened exactly).
    * without the delayed term (b = 0: an additive-noise SDE) the methods show their nominal order 1.5; EM would show 1.0;
    * with it (b = -2) the stages evaluate f at t + c0 dt, whose delayed state comes from the LINEAR interpolant of the
      dense history ([EXT] the reference's sde_interpolant is linear too): a Brownian-bridge error of O(sqrt dt) per lookup,
      independent from step to step, which caps the global strong order at 1.0 -- a property of the reference's design
      that this path reproduces, not an implementation defect."""
    n, T, fine = 200, 2.0, 2.0 ** -10
    rng = np.random.default_rng(2028)
    nf = int(round(T / fine))
    dWf = rng.standard_normal((nf, 1, n)) * math.sqrt(fine)
    dZf = rng.standard_normal((nf, 1, n)) * math.sqrt(fine)
    model = O.get_model(O.MODEL_HAYES)

    def run(dt, b):
        k = int(round(dt / fine))
        dWc, dZc = _coarsen_wz(dWf, dZf, fine, k)
        sch = O.build_schedule(0.0, T, dt, [1.0], alg, saveat_times=[T], save_start=False)
        assert sch.n_steps == nf // k
        out, _ = O.solve_ensemble(model, [1.0, -4.0, b, 10.0, 0.0, 0.0, 0.1], [1.0], [1.0], [1.0], [1.0], sch, alg,
                                  np.concatenate([dWc, dZc], axis=1))
        return out[-1, 0]

    dts = 2.0 ** -np.arange(7, 2, -1)
    slopes = {}
    for b in (0.0, -2.0):
        ref = run(fine, b)
        errs = [np.mean(np.abs(run(dt, b) - ref)) for dt in dts]
        slopes[b] = np.polyfit(np.log(dts), np.log(errs), 1)[0]
    assert 1.3 < slopes[0.0] < 2.4, (alg, slopes)
    assert 0.8 < slopes[-2.0] < 1.8, (alg, slopes)


def test_sra_python_and_cpp_oracles_agree_bit_for_bit():
    """Both restatements, all three built-in tableaus + a caller-supplied one, off-grid lag (every stage interpolates the
    history at its own time) and a lag equal to dt (the t + dt stage reads the node being left)."""
    n = 6
    rng = np.random.default_rng(5)
    p = [1.0, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]
    custom = O.SraTableau(c0=(0.0, 0.5), c1=(0.25, 1.0), A0=((0.0, 0.0), (0.5, 0.0)), B0=((0.0, 0.0), (1.0, 0.0)),
                          alpha=(0.25, 0.75), beta1=(0.5, 0.5), beta2=(1.0, -1.0))
    for alg, tab, lag, dt in ((O.ALG_SRA1, None, 0.33, 0.05), (O.ALG_SRA3, None, 0.33, 0.05), (O.ALG_SOSRA, None, 0.05, 0.05),
                              (O.ALG_SRA_TABLEAU, custom, 0.33, 0.05)):
        sch = O.build_schedule(0.0, 1.0, dt, [lag], alg, saveat_times=[], save_everystep=True, tableau=tab)
        noise = rng.standard_normal((sch.n_steps, 2, n)) * sch.sqdt[:, None, None]
        pp = list(p); pp[0] = lag
        out, _ = O.solve_ensemble(O.get_model(O.MODEL_HAYES), pp, [1.0], [1.0], [1.0], [lag], sch, alg, noise, tableau=tab)
        co, _, _, steps = c_oracle.solve(O.MODEL_HAYES, alg, 1, 1, 1, pp, [1.0], [1.0], [1.0], [lag], 0.0, 1.0, dt, [], len(sch.save_t),
                                         save_everystep=True, n_traj=n, dW=noise, tableau=tab)
        assert steps in (sch.n_steps, sch.n_steps * n) and np.array_equal(co, out), alg


def _golden_v2():
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_golden_v2", os.path.join(here, "golden", "make_golden_v2.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, np.load(os.path.join(here, "golden", "golden_v2.npz"))


def test_golden_v2_fixtures_reproduce_in_both_oracles():
    """Round-2 algorithms (EM(false), RKMil Stratonovich, EulerHeun, RKMilCommute on a noise matrix, WangLi3SMil_A / _F, SRA1 /
    SRA3 / SOSRA): the committed vectors are reproduced bit for bit by the numpy oracle and by the C++ restatement."""
    mod, g = _golden_v2()
    for name, fx in mod.fixtures().items():
        sch, xi, out = mod.run(fx)
        assert np.array_equal(xi, g[name + "_xi"]) and np.array_equal(sch.t, g[name + "_t"]) and np.array_equal(sch.save_t, g[name + "_save_t"])
        assert np.array_equal(out, g[name + "_out"]), name
        model, p, u0, h0, h1, lags, tspan, dt, alg, saveat, rows, key = fx
        sv, ss, se = O.expand_saveat(saveat, *tspan)
        co, _, _, _ = c_oracle.solve(model.mid, alg, model.n_state, model.n_lag, model.n_noise, p, u0, h0, h1, lags, tspan[0], tspan[1], dt,
                                     sv, len(sch.save_t), save_start=ss, save_end=se, n_traj=mod.N, dW=xi * sch.sqdt[:, None, None])
        assert np.array_equal(co, out), name
