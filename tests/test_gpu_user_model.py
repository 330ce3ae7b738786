"""-m gpu: user-defined f/g compiled at run time with NVRTC -- the generality seam of
/root/reference/src/functionwrapper.jl:35-36 -- checked against the oracle with the same model written in numpy."""
import numpy as np
import pytest

import cases
import sdde_b200 as S
import sdde_oracle as O

pytestmark = pytest.mark.gpu

# a 2-state delayed oscillator with state-dependent diagonal noise, one lag, 3 parameters
SRC = r"""
struct UserModel {
    static constexpr int NS = 2, NL = 1, NW = 2, NP = 3;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double t) {
        du[0] = u[1];
        du[1] = -p[0] * u[0] - p[1] * ul[0][0] + sin(t);
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*ul)[NS], const P &p, double t) {
        gv[0] = p[2];
        gv[1] = p[2] * u[0] * ul[0][1];
    }
};
"""


def _f(u, ul, p, t):
    return np.stack([u[1], (-p[0] * u[0] - p[1] * ul[0][0]) + np.sin(t)])


def _g(u, ul, p, t):
    return np.stack([p[2] + 0.0 * u[0], (p[2] * u[0]) * ul[0][1]])


@pytest.mark.parametrize("alg", [0, 2])
def test_nvrtc_user_model_matches_oracle(alg):
    um = S.UserModel(SRC, 2, 1, 2, 3, True, name="delayed_oscillator")
    om = O.Model(1000, 2, 1, 2, True, 3, _f, _g)
    p = [1.0, 0.5, 0.2]
    prob = S.SDDEProblem(um.f, um.g, [1.0, 0.0], S.AffineHistory((1.0, 0.0), (0.1, 0.0)), (0.0, 6.0), p,
                         constant_lags=(0.7,))
    algo = cases.ALG_MAP[alg]
    kw = dict(dt=0.01, saveat=0.5)
    if alg >= 2:
        kw["adaptive"] = False
    n = 777
    sv = O.expand_saveat(0.5, 0.0, 6.0)[0]
    osch = O.build_schedule(0.0, 6.0, 0.01, [0.7], alg, saveat_times=sv)
    dW = cases.increments(osch, 2, n)
    sim = S.solve(S.EnsembleProblem(prob), algo, S.EnsembleB200(device=0), trajectories=n, noise_increments=dW, **kw)
    ref, _ = O.solve_ensemble(om, p, [1.0, 0.0], [1.0, 0.0], [0.1, 0.0], [0.7], osch, alg, dW)
    assert sim.out.shape == ref.shape
    assert cases.scaled_err(sim.out, ref) < 1e-12
    assert "user" in sim.timing["kernel_name"]
    # in-kernel Philox path of the JIT-compiled kernel runs and is finite
    sim2 = S.solve(S.EnsembleProblem(prob), algo, S.EnsembleB200(device=0), trajectories=5000, seed=11, moments=True, **kw)
    assert np.all(np.isfinite(sim2.out)) and sim2.moments[0, 0, 1] == 1.0


def test_nvrtc_compile_error_is_reported():
    bad = S.UserModel("struct UserModel { this is not C++ };", 1, 1, 1, 0, True)
    prob = S.SDDEProblem(bad.f, bad.g, 1.0, 1.0, (0.0, 1.0), (), constant_lags=(0.5,))
    with pytest.raises(S.SddeError) as ei:
        S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.1, trajectories=4)
    assert ei.value.code == -4 and "NVRTC" in str(ei.value)


def test_user_model_per_trajectory_params_noise_grid_and_rkmil():
    """The JIT path compiles one variant per (algorithm, noise source, parameter mode): exercise per-trajectory
    parameters + NoiseGrid + RKMil on the 2-state diagonal-noise user model."""
    um = S.UserModel(SRC, 2, 1, 2, 3, True, name="delayed_oscillator")
    om = O.Model(1000, 2, 1, 2, True, 3, _f, _g)
    n = 333
    rng = np.random.default_rng(9)
    pb = np.stack([1.0 + 0.2 * rng.random(n), 0.5 * np.ones(n), 0.1 + 0.1 * rng.random(n)])
    ub = np.stack([1.0 + 0.1 * rng.random(n), 0.05 * rng.random(n)])
    prob = S.SDDEProblem(um.f, um.g, [1.0, 0.0], S.AffineHistory((1.0, 0.0), (0.1, 0.0)), (0.0, 4.0), [1.0, 0.5, 0.2],
                         constant_lags=(0.7,))
    sv = O.expand_saveat(0.5, 0.0, 4.0)[0]
    osch = O.build_schedule(0.0, 4.0, 0.01, [0.7], 2, saveat_times=sv)
    dW = cases.increments(osch, 2, n)
    W = np.concatenate([np.zeros((1, 2, n)), np.cumsum(dW, axis=0)], axis=0)
    cur = np.zeros((2, n)); dWg = np.empty_like(dW)
    for k in range(osch.n_steps):
        dWg[k] = W[k + 1] - cur
        cur = cur + dWg[k]
    sim = S.solve(S.EnsembleProblem(prob, p_batch=pb, u0_batch=ub), S.RKMil(), S.EnsembleB200(device=0), trajectories=n,
                  noise_grid=W, dt=0.01, saveat=0.5, adaptive=False)
    ref, _ = O.solve_ensemble(om, pb, ub, [1.0, 0.0], [0.1, 0.0], [0.7], osch, 2, dWg)
    assert "ptraj1" in sim.timing["kernel_name"] and "noise2" in sim.timing["kernel_name"]
    assert cases.scaled_err(sim.out, ref) < 1e-12


SRC_H = r"""
struct UserModel {
    static constexpr int NS = 2, NL = 2, NW = 2, NP = 2;
    static constexpr bool DIAG = true;
    static constexpr bool HAS_H = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double t) {
        du[0] = -p[0] * ul[0][0] + ul[1][1];
        du[1] = -u[1] + ul[0][0] * ul[1][0];
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*ul)[NS], const P &p, double t) {
        gv[0] = p[1] * u[0];
        gv[1] = p[1] * ul[1][1];
    }
    template <class P> __device__ __forceinline__ static void h(double *out, const P &p, double t) {
        out[0] = cos(p[0] * t);
        out[1] = 0.5 + 0.1 * t * t;
    }
};
"""


@pytest.mark.parametrize("alg", [0, 3, 4])
@pytest.mark.parametrize("ptraj", [False, True])
def test_user_model_with_its_own_history_function(alg, ptraj):
    """A non-affine h(p, t) (cosine and parabola, depending on p) evaluated on the device for every lookup before t0;
    two lags, one of them off the step grid; EM / RKMil Stratonovich / EulerHeun; shared and per-trajectory p."""
    um = S.UserModel(SRC_H, 2, 2, 2, 2, True, name="own_history", history=True)

    def f(u, ul, p, t):
        return np.stack([-p[0] * ul[0][0] + ul[1][1], -u[1] + ul[0][0] * ul[1][0]])

    def g(u, ul, p, t):
        return np.stack([p[1] * u[0], p[1] * ul[1][1]])

    om = O.Model(1001, 2, 2, 2, True, 2, f, g)
    n = 257
    rng = np.random.default_rng(4)
    p = np.array([1.3, 0.2])
    pb = np.stack([1.0 + 0.5 * rng.random(n), 0.1 + 0.2 * rng.random(n)]) if ptraj else None
    lags = (0.5, 1.23)
    prob = S.SDDEProblem(um.f, um.g, [1.0, 0.5], um.h, (0.0, 3.0), p, constant_lags=lags)
    kw = dict(dt=0.01, saveat=0.25)
    if alg in (2, 3):
        kw["adaptive"] = False
    sv = O.expand_saveat(0.25, 0.0, 3.0)[0]
    osch = O.build_schedule(0.0, 3.0, 0.01, list(lags), alg, saveat_times=sv)
    dW = cases.increments(osch, 2, n)
    sim = S.solve(S.EnsembleProblem(prob, p_batch=pb), cases.ALG_MAP[alg], S.EnsembleB200(device=0), trajectories=n,
                  noise_increments=dW, **kw)

    def hist(pp, t):
        one = np.ones(n)
        return np.stack([np.cos(pp[0] * t) * one, (0.5 + 0.1 * t * t) * one])

    ref, _ = O.solve_ensemble(om, pb if ptraj else p, [1.0, 0.5], [0.0, 0.0], [0.0, 0.0], list(lags), osch, alg, dW, history_fn=hist)
    assert cases.scaled_err(sim.out, ref) < 1e-12
    # the history really matters: an affine stand-in gives different paths
    ref_aff, _ = O.solve_ensemble(om, pb if ptraj else p, [1.0, 0.5], [1.0, 0.5], [0.0, 0.0], list(lags), osch, alg, dW)
    assert cases.scaled_err(ref, ref_aff) > 1e-3


def test_user_model_history_argument_checks():
    um = S.UserModel(SRC_H, 2, 2, 2, 2, True, history=True)
    plain = S.UserModel(SRC, 2, 1, 2, 3, True)
    with pytest.raises(ValueError):
        S.SDDEProblem(um.f, um.g, [1.0, 0.5], 1.0, (0.0, 1.0), [1.0, 0.1], constant_lags=(0.5, 1.0))
    with pytest.raises(AttributeError):
        plain.h


def test_user_model_additive_noise_with_sosra():
    """A user model with additive (u-independent, time- and delay-dependent) noise under SOSRA: the NVRTC build of the SRA
    kernel (four lookup sets per step, two increments per channel) against the oracle."""
    src = r"""
struct UserModel {
    static constexpr int NS = 2, NL = 1, NW = 2, NP = 3;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double t) {
        du[0] = u[1];
        du[1] = -p[0] * u[0] - p[1] * ul[0][0] + sin(t);
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*ul)[NS], const P &p, double t) {
        gv[0] = p[2];
        gv[1] = p[2] * (1.0 + t) + 0.1 * ul[0][1];
    }
};
"""
    def g(u, ul, p, t):
        return np.stack([p[2] + 0.0 * u[0], p[2] * (1.0 + t) + 0.1 * ul[0][1]])
    um = S.UserModel(src, 2, 1, 2, 3, True, name="additive_oscillator")
    om = O.Model(1000, 2, 1, 2, True, 3, _f, g)
    p = [1.0, 0.5, 0.2]
    prob = S.SDDEProblem(um.f, um.g, [1.0, 0.0], S.AffineHistory((1.0, 0.0), (0.1, 0.0)), (0.0, 3.0), p, constant_lags=(0.7,))
    sv = O.expand_saveat(0.5, 0.0, 3.0)[0]
    osch = O.build_schedule(0.0, 3.0, 0.02, [0.7], O.ALG_SOSRA, saveat_times=sv)
    nz = cases.increments(osch, 4, 300)
    sim = S.solve(S.EnsembleProblem(prob), S.SOSRA(), S.EnsembleB200(device=0), trajectories=300, noise_increments=nz, dt=0.02,
                  saveat=0.5, adaptive=False)
    ref, _ = O.solve_ensemble(om, p, [1.0, 0.0], [1.0, 0.0], [0.1, 0.0], [0.7], osch, O.ALG_SOSRA, nz)
    assert cases.scaled_err(sim.out, ref) < 1e-12
