"""Order statistics across trajectories on the device (sdde_b200_quantiles) against the oracle's sort-based restatement of
Statistics.quantile / Statistics.median -- what [EXT] SciMLBase EnsembleSummary computes on the reference side.
Bit-exact: selection is integer work, the final interpolation is the same float64 expression."""
import numpy as np
import pytest

import cases
import sdde_b200 as S
import sdde_oracle as O

pytestmark = pytest.mark.gpu


def _device_quantiles(rows, probs, median, on_device):
    import torch
    n_rows, n = rows.shape
    sim = S.EnsembleSolution(np.arange(n_rows, dtype=float), None, None, None, False, 0.0, {})
    sim.out = torch.from_numpy(rows.reshape(n_rows, 1, n)).cuda() if on_device else rows.reshape(n_rows, 1, n)
    return S.timeseries_steps_quantile(sim, probs, median=median)[:, 0, :]


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 1000, 4097, 70001])
@pytest.mark.parametrize("on_device", [True, False])
def test_quantiles_equal_sorted_selection(n, on_device):
    rng = np.random.default_rng(1000 + n)
    rows = np.stack([
        rng.standard_normal(n),                                   # both signs
        1.2 + 1e-3 * rng.standard_normal(n),                      # one exponent, shared leading mantissa bits
        np.round(rng.standard_normal(n), 1),                      # many ties, +-0
        np.exp(40 * rng.standard_normal(n)),                      # wide dynamic range
        np.full(n, 0.75),                                         # all equal
        -np.abs(rng.standard_normal(n)) * 1e-310,                 # negative subnormals
    ])
    rows[2, ::7] = -0.0
    probs = [0.0, 0.05, 0.25, 0.5, 0.95, 1.0]
    got = _device_quantiles(rows, probs, True, on_device)
    ref = O.ensemble_quantiles(rows, probs, want_median=True)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (got, ref)


def test_quantiles_inf_and_nan_rows():
    rng = np.random.default_rng(5)
    n = 5000
    rows = rng.standard_normal((3, n))
    rows[0, :40] = np.inf
    rows[0, 40:60] = -np.inf
    rows[1, 123] = np.nan
    got = _device_quantiles(rows, [0.001, 0.5, 0.999], True, True)
    ref = O.ensemble_quantiles(rows, [0.001, 0.5, 0.999], want_median=True)
    assert np.all(np.isnan(got[1]))
    assert np.array_equal(got[[0, 2]], ref[[0, 2]], equal_nan=True)


def test_quantile_argument_errors():
    rows = np.zeros((1, 1000))
    with pytest.raises(S.SddeError):
        _device_quantiles(rows, [1.5], False, True)
    with pytest.raises(S.SddeError):
        _device_quantiles(rows, list(np.linspace(0.01, 0.99, 14)), False, True)     # > 12 distinct order statistics


def test_ensemble_summary_of_a_solved_ensemble():
    """EnsembleSummary on a C2-shaped solve kept in HBM: mean/var from the moment triples, median and 5 %/95 % bands by
    selection; equal to the oracle's statistics of the same paths."""
    case = cases.c2(T=20.0)
    sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(output="device"), trajectories=20000,
                  seed=11, moments=True, **case.solve_kwargs())
    summ = S.EnsembleSummary(sim, quantiles=(0.05, 0.95))
    paths = sim.out.cpu().numpy()
    ref = O.ensemble_quantiles(paths.reshape(-1, paths.shape[2]), [0.05, 0.95], want_median=True).reshape(paths.shape[0], paths.shape[1], 3)
    assert np.array_equal(summ.qlow, ref[..., 0]) and np.array_equal(summ.qhigh, ref[..., 1]) and np.array_equal(summ.med, ref[..., 2])
    assert np.allclose(summ.u, paths.mean(axis=2), rtol=1e-12) and np.allclose(summ.v, paths.var(axis=2, ddof=1), rtol=1e-9)
    assert summ.num_monte == 20000 and np.all(summ.qlow <= summ.med) and np.all(summ.med <= summ.qhigh)
    assert "quantile_kernel_ms" in sim.timing


def test_quantiles_full_size_row_properties():
    """BASELINE-size rows (10^6 trajectories): size-independent properties -- the selected value has exactly the right
    number of elements below and not above it."""
    import torch
    n = 1_000_000
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (1.0 + 0.2 * torch.randn(8, 1, n, dtype=torch.float64, device="cuda", generator=g))
    sim = S.EnsembleSolution(np.arange(8.0), x, None, None, False, 0.0, {})
    q = S.timeseries_steps_quantile(sim, [0.0, 0.05, 0.95, 1.0], median=True)[:, 0, :]
    xs = x[:, 0, :]
    assert np.array_equal(q[:, 0], xs.min(dim=1).values.cpu().numpy()) and np.array_equal(q[:, 3], xs.max(dim=1).values.cpu().numpy())
    for col, p in ((1, 0.05), (2, 0.95), (4, 0.5)):
        v = torch.from_numpy(q[:, col]).cuda()[:, None]
        below = (xs < v).sum(dim=1).cpu().numpy()
        not_above = (xs <= v).sum(dim=1).cpu().numpy()
        h = (n - 1) * p
        assert np.all(below <= np.floor(h) + 1) and np.all(not_above >= np.floor(h)), (p, below, not_above)


def test_summary_of_a_solution_solved_without_moments():
    """EnsembleSummary computes the missing mean/variance from the saved paths on the device (sdde_b200_moments), for
    device-resident and host-resident paths alike."""
    case = cases.c1(T=5.0)
    kw = dict(trajectories=30_001, seed=3, **case.solve_kwargs())
    for output in ("device", "host"):
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(output=output), **kw)
        assert sim.moments is None
        summ = S.EnsembleSummary(sim)
        paths = sim.out.cpu().numpy() if output == "device" else sim.out
        assert np.allclose(summ.u, paths.mean(axis=2), rtol=1e-12) and np.allclose(summ.v, paths.var(axis=2, ddof=1), rtol=1e-9)
        assert sim.moments is not None and np.all(sim.moments[..., 0] == 30_001)


def test_ensemble_analysis_per_step_and_per_point_forms():
    """[EXT] EnsembleAnalysis timestep_* / timepoint_* / timeseries_point_*: same names and argument order; order statistics are
    Statistics.quantile's (= numpy 'linear'), time points between saves use the linearly interpolated paths."""
    case = cases.c1(T=4.0, saveat=0.5)
    for output in ("host", "device"):
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output=output), trajectories=4001, seed=17,
                      moments=True, **case.solve_kwargs())
        out = sim.out.cpu().numpy() if hasattr(sim.out, "cpu") else sim.out
        i, t = 3, 1.7                                        # saved point 3 (t = 1.5); t = 1.7 lies between saves 3 and 4
        th = (t - 1.5) / 0.5
        xt = (1.0 - th) * out[3] + th * out[4]
        assert np.array_equal(np.asarray(S.get_timestep(sim, i).cpu() if hasattr(sim.out, "cpu") else S.get_timestep(sim, i)), out[i])
        m, v = S.timestep_meanvar(sim, i)
        assert np.allclose(m, out[i].mean(axis=1), rtol=1e-12) and np.allclose(v, out[i].var(axis=1, ddof=1), rtol=1e-9)
        assert np.allclose(S.timestep_quantile(sim, [0.1, 0.9], i), np.quantile(out[i], [0.1, 0.9], axis=1).T, rtol=1e-14)
        assert np.allclose(S.timestep_median(sim, i), np.median(out[i], axis=1), rtol=1e-14)
        m, v = S.timepoint_meanvar(sim, t)
        assert np.allclose(m, xt.mean(axis=1), rtol=1e-12) and np.allclose(v, xt.var(axis=1, ddof=1), rtol=1e-9)
        assert np.allclose(S.timepoint_quantile(sim, 0.25, t), np.quantile(xt, 0.25, axis=1), rtol=1e-13)
        assert np.allclose(S.timepoint_median(sim, t), np.median(xt, axis=1), rtol=1e-13)
        ms, vs = S.timeseries_point_meanvar(sim, [0.0, 1.7, 4.0])
        assert ms.shape == (3, 1) and np.allclose(ms[1], xt.mean(axis=1), rtol=1e-12) and np.allclose(ms[2], out[-1].mean(axis=1), rtol=1e-12)
        q = S.timeseries_point_quantile(sim, [0.05, 0.95], [0.25, 1.7])
        assert q.shape == (2, 1, 2) and np.allclose(q[1, 0], np.quantile(xt[0], [0.05, 0.95]), rtol=1e-13)
        assert np.allclose(S.timeseries_point_median(sim, [1.7])[0], np.median(xt, axis=1), rtol=1e-13)
        assert np.allclose(sim[5](t), xt[0, 5], rtol=1e-15)


import os as _os
_Q0, _QN = ([int(x) for x in _os.environ["SDDE_FUZZ_SEEDS"].split(":")] if _os.environ.get("SDDE_FUZZ_SEEDS") else [0, 8])


@pytest.mark.parametrize("seed", list(range(_Q0, _Q0 + _QN)))
def test_random_rows_quantiles_equal_sorted_selection(seed):
    """Randomised: row counts, lengths, value distributions (ties, few distinct values, one binade, wide range, signed zeros,
    infinities), probabilities and residency -- bit-exact against the sort-based oracle (SDDE_FUZZ_SEEDS widens the sweep)."""
    rng = np.random.default_rng(31000 + seed)
    n = int(rng.choice([1, 2, 5, 63, 64, 65, 257, 1000, 4096, 4097, 50_001, 200_003]))
    n_rows = int(rng.integers(1, 9))
    rows = np.empty((n_rows, n))
    for r in range(n_rows):
        kind = rng.integers(0, 8)
        if kind == 0: v = rng.standard_normal(n)
        elif kind == 1: v = 1.0 + rng.random(n) * 2.0 ** -int(rng.integers(1, 50))
        elif kind == 2: v = rng.integers(-3, 4, size=n).astype(float)
        elif kind == 3: v = np.exp(rng.uniform(-700, 700, size=n)) * rng.choice([-1.0, 1.0], size=n)
        elif kind == 4: v = np.full(n, float(rng.standard_normal()))
        elif kind == 5: v = np.round(rng.standard_normal(n), int(rng.integers(0, 3)))
        elif kind == 6: v = np.where(rng.random(n) < 0.5, 0.0, -0.0)
        else: v = rng.standard_normal(n) * 10.0 ** float(rng.integers(-300, 300))
        if rng.random() < 0.2 and n > 3:
            v[rng.integers(0, n, size=max(1, n // 50))] = rng.choice([np.inf, -np.inf])
        rows[r] = v
    k = int(rng.integers(1, 6))
    probs = sorted(set(float(x) for x in rng.choice([0.0, 0.01, 0.05, 0.1, 0.25, 1 / 3, 0.5, 0.75, 0.9, 0.95, 0.99, 1.0], size=k, replace=False)))
    median = bool(rng.integers(0, 2))
    got = _device_quantiles(rows, probs, median, bool(rng.integers(0, 2)))
    ref = O.ensemble_quantiles(rows, probs, want_median=median)
    assert got.shape == ref.shape
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), (seed, n, n_rows, probs, median)
