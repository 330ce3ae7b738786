"""``solve(EnsembleProblem, EM()/RKMil(), EnsembleB200(); dt, saveat, trajectories, ...)`` -> ``EnsembleSolution``.

Host-side mirror of the dispatch a Julia maintainer would add (julia/EnsembleB200.jl):
``SciMLBase.__solve(::AbstractEnsembleProblem{<:SDDEProblem}, ::Union{EM,RKMil}, ::EnsembleB200; trajectories, dt, saveat, ...)``
which replaces the per-trajectory loop over ``DiffEqBase.__solve`` (/root/reference/src/solve.jl:1-9).  Keyword names and
defaults follow ``__init`` (/root/reference/src/solve.jl:18-77) for the keywords that exist on this path; everything
else raises.  All numerical work happens in libsdde_b200.so; this module only marshals the descriptor.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import time
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from .problem import (EM, EulerHeun, RKMil, SRA, WangLi3SMil, EnsembleProblem, FinalState, MomentsReduction, SaveIdxs, SDDEProblem, UserModel,
                      _as_history)


@dataclass
class EnsembleB200:
    """The ensemble algorithm: where SciML has EnsembleThreads()/EnsembleDistributed(), this path has EnsembleB200().

    device       CUDA device of this process (default: LOCAL_RANK under torchrun, else 0)
    distributed  shard trajectories over the ranks of the initialised torch.distributed group (one process per GPU)
                 and merge ensemble moments over NCCL; paths stay on the rank that computed them
    output       'host' -> numpy arrays (D2H inside the call) | 'device' -> torch CUDA tensors (no D2H)
    """
    device: int | None = None
    distributed: bool = False
    output: str = "host"
    wave_traj: int | None = None
    devices: tuple | None = None   # several GPUs driven from THIS process (sdde_b200_solve_multi); host output only
    lanes_per_traj: int = 0      # 0/1: thread per trajectory; n_state (4): cooperative lanes (models with a row-wise form)


# ---------------------------------------------------------------------------------------------
# saveat expansion: /root/reference/src/solve.jl:32-34, 411-415, 659-677
# ---------------------------------------------------------------------------------------------
def _rat(x: float):
    """[EXT] Base.rat (julia/base/twiceprecision.jl): continued-fraction approximation num/den of a Float64 with both
    terms <= maxintfloat(Float32) = 2^24; returns the previous convergent when the next one would overflow that bound."""
    m = 16777216.0
    y, a, b, c, d = x, 1, 0, 0, 1
    while abs(y) <= m:
        f = math.trunc(y)
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > 16777216:
            return c, d
        if b != 0 and a / b == x:
            break
        y = math.inf if y == 0 else 1.0 / y
    return a, b


def _julia_range(a: float, s: float, b: float) -> np.ndarray:
    """Elements of the Julia Float64 range ``a:s:b`` ([EXT] Base.(:)(start::T, step::T, stop::T) where T<:IEEEFloat,
    julia/base/twiceprecision.jl).  When start, step and stop all have exact small-rational lifts the range is
    ``(start_n + k*step_n)/den`` evaluated in twice precision -- modelled here by exact integer arithmetic and ONE
    rounding; otherwise start and step are taken literally, the length is ``round((stop-start)/step) + 1`` minus an
    overshoot correction, and the elements are ``start + k*step`` in twice precision (exact rational sum, one rounding)."""
    from fractions import Fraction
    from math import gcd
    if s == 0:
        raise ValueError("range step cannot be zero")
    between = lambda lo, x, hi: lo <= x <= hi or hi <= x <= lo
    step_n, step_d = _rat(s)
    if step_d != 0 and step_n / step_d == s:
        start_n, start_d = _rat(a)
        stop_n, stop_d = _rat(b)
        if start_d != 0 and stop_d != 0 and start_n / start_d == a and stop_n / stop_d == b:
            den = start_d * (step_d // gcd(start_d, step_d))
            m = 9007199254740992.0
            if den != 0 and abs(a * den) <= m and abs(s * den) <= m and den % start_d == 0 and den % step_d == 0:
                start_n, step_n = round(a * den), round(s * den)
                dd = step_n * stop_d
                num = den * stop_n - stop_d * start_n + dd                          # "+ 1" folded into the dividend, as Julia does
                ln = max(0, abs(num) // abs(dd) * (1 if (num >= 0) == (dd > 0) else -1))   # Julia's div truncates toward zero
                if between(a, a + (ln - 1) * s, b + s / 2) and not between(a, a + ln * s, b):
                    return np.array([float(Fraction(start_n + k * step_n, den)) for k in range(ln)], dtype=np.float64)
    lf = (b - a) / s
    if lf < 0:
        ln = 0
    elif lf == 0:
        ln = 1
    else:
        ln = int(round(lf)) + 1                      # Python and Julia both round half to even
        last = a + (ln - 1) * s
        ln -= int(a < b < last) + int(a > b > last)
    fa, fs = Fraction(a), Fraction(s)
    return np.array([float(fa + k * fs) for k in range(ln)], dtype=np.float64)


_saveat_cache: dict = {}


def expand_saveat(saveat, t0: float, tf: float):
    """-> (saveat times pushed to the heap, save_start default, save_end default); /root/reference/src/solve.jl:661-677:
    a number s pushes ``(t0+s):s:tf`` when ``(t0:s:tf)[end] == tf``, else ``(t0+s):s:(tf-s)``."""
    if saveat is None:
        return np.zeros(0), True, True
    if np.isscalar(saveat):
        s = float(saveat)
        if not (s > 0) or not math.isfinite(s):
            raise ValueError("saveat must be positive and finite")
        if not (math.isfinite(t0) and math.isfinite(tf)) or (tf - t0) / s > 2.0e8:
            raise ValueError("saveat: more than 2e8 save points (or an infinite time span)")
        key = (s, float(t0), float(tf))
        if _saveat_cache.get("key") != key:          # repeated solves of one problem: the exact-rational range costs ~0.5 ms
            full = _julia_range(t0, s, tf)
            sv = _julia_range(t0 + s, s, tf) if full.size and full[-1] == tf else _julia_range(t0 + s, s, tf - s)
            _saveat_cache.update(key=key, sv=np.asarray(sv, dtype=np.float64))
        return _saveat_cache["sv"].copy(), True, True
    arr = np.asarray(list(saveat), dtype=np.float64)
    if arr.size == 0:
        return np.zeros(0), True, True
    if np.any(np.isnan(arr)):
        raise ValueError("saveat must not contain NaN")
    return np.sort(arr[(arr > t0) & (arr < tf)]), bool(np.any(arr == t0)), bool(np.any(arr == tf))


# ---------------------------------------------------------------------------------------------
# descriptor marshalling
# ---------------------------------------------------------------------------------------------
class _Marshalled:
    """Keeps every numpy buffer the descriptor points to alive (the Julia side uses GC.@preserve)."""

    def __init__(self):
        self.desc = capi.SddeDesc()
        self.keep = []


def _marshal(prob: SDDEProblem, model, alg, *, dt, saveat, save_everystep, save_start, save_end, tstops, adaptive,
             n_traj, traj_offset, seed, p_batch, u0_batch, noise, noise_mode, noise_mem, out_mem, want_paths,
             want_moments, lanes_per_traj=0, save_idxs=None, noise_t=None, d_discontinuities=()) -> _Marshalled:
    if not isinstance(alg, (EM, RKMil, EulerHeun, WangLi3SMil, SRA)):
        raise NotImplementedError(f"{type(alg).__name__} is not on this path (EM, RKMil, RKMilCommute, EulerHeun, WangLi3SMil_A-F, "
                                  "SRA1 / SRA3 / SOSRA / SRA(tableau), fixed step)")
    if adaptive is None:
        adaptive = isinstance(alg, (RKMil, SRA))      # [EXT] isadaptive(RKMil) = isadaptive(SRA1) = true, src/solve.jl:40
    if adaptive:
        raise NotImplementedError("adaptive stepping is outside this path: pass adaptive=False (RKMil is adaptive by "
                                  "default in the reference, src/solve.jl:40)")
    t0, tf = prob.tspan
    m = _Marshalled()
    d = m.desc
    d.struct_size = C.sizeof(capi.SddeDesc)
    d.model_id, d.alg = model.model_id, alg.alg_id
    if isinstance(alg, SRA):
        flat, stages = alg.flat_tableau()
        if flat is not None:
            m.keep_extra = flat
            d.sra_tableau, d.sra_stages = capi.np_ptr(flat).value, stages
        d.sra_order2 = int(round(2 * alg.order))
    d.n_state, d.n_noise, d.n_lag, d.n_param = model.n_state, model.n_noise, model.n_lag, model.n_param
    if len(prob.lags) != model.n_lag:
        raise ValueError(f"model {model.name} looks up {model.n_lag} lag(s), problem declares {len(prob.lags)}")
    if prob.u0.size != model.n_state:
        raise ValueError(f"u0 has {prob.u0.size} states, model {model.name} has {model.n_state}")
    if prob.p.size != model.n_param:
        raise ValueError(f"p has {prob.p.size} entries, model {model.name} takes {model.n_param}")
    d.order_discontinuity_t0 = prob.order_discontinuity_t0
    d.constant_lags_declared = 1 if len(prob.constant_lags) > 0 else 0
    sv, ss_def, se_def = expand_saveat(saveat, t0, tf)
    no_saveat = saveat is None or (not np.isscalar(saveat) and len(saveat) == 0)
    if save_everystep is None:
        save_everystep = no_saveat                                   # src/solve.jl:25
    if save_start is None:
        save_start = True if (save_everystep or no_saveat or np.isscalar(saveat)) else ss_def   # :32-34
    if save_end is None:
        save_end = True if (save_everystep or no_saveat or np.isscalar(saveat)) else se_def     # :411-415
    d.save_start, d.save_end, d.save_everystep = int(save_start), int(save_end), int(save_everystep)
    d.t0, d.tf, d.dt = t0, tf, float(dt)
    for i, lag in enumerate(prob.lags):
        d.lags[i] = lag
    hist = _as_history(prob.h, prob.p, t0, model.n_state)
    for j in range(model.n_state):
        d.hist_c0[j], d.hist_c1[j] = hist.c0[j], hist.c1[j]
    sv = np.ascontiguousarray(sv, dtype=np.float64)
    ts = np.ascontiguousarray(np.asarray(list(tstops), dtype=np.float64))
    m.keep += [sv, ts]
    d.saveat, d.n_saveat = capi.np_ptr(sv).value if sv.size else None, sv.size
    d.tstops, d.n_tstops = capi.np_ptr(ts).value if ts.size else None, ts.size
    # d_discontinuities keyword (src/solve.jl:20-23): Discontinuity(t, order) pairs, or bare times (order 0)
    dd = [(float(x[0]), x[1]) if isinstance(x, (tuple, list)) else (float(x), 0) for x in d_discontinuities]
    if any(round(2 * o) != 2 * o for _, o in dd):
        raise ValueError("discontinuity orders must be multiples of 1/2")
    dd_t = np.array([t for t, _ in dd], dtype=np.float64)
    dd_o = np.array([int(round(2 * o)) for _, o in dd], dtype=np.int32)
    m.keep += [dd_t, dd_o]
    if dd:
        d.d_disc_t, d.d_disc_order2, d.n_d_disc = capi.np_ptr(dd_t).value, capi.np_ptr(dd_o).value, len(dd)
    if p_batch is not None:
        if p_batch.shape != (model.n_param, n_traj):
            raise ValueError(f"p_batch must have shape ({model.n_param}, {n_traj})")
        m.keep.append(p_batch)
        d.params, d.param_stride = capi.np_ptr(p_batch).value, n_traj
    else:
        pp = np.ascontiguousarray(prob.p)
        m.keep.append(pp)
        d.params, d.param_stride = (capi.np_ptr(pp).value if pp.size else None), 0
    if u0_batch is not None:
        if u0_batch.shape != (model.n_state, n_traj):
            raise ValueError(f"u0_batch must have shape ({model.n_state}, {n_traj})")
        m.keep.append(u0_batch)
        d.u0, d.u0_stride = capi.np_ptr(u0_batch).value, n_traj
    else:
        uu = np.ascontiguousarray(prob.u0)
        m.keep.append(uu)
        d.u0, d.u0_stride = capi.np_ptr(uu).value, 0
    d.noise_mode, d.noise_mem = noise_mode, noise_mem
    d.noise = noise
    if noise_t is not None:
        m.keep.append(noise_t)
        d.noise_t, d.n_noise_t = capi.np_ptr(noise_t).value, noise_t.size
    d.out_mem = out_mem
    d.history_mode = 0
    d.lanes_per_traj = int(lanes_per_traj)
    if save_idxs is not None:
        idx = [int(i) for i in np.atleast_1d(save_idxs)]
        if not idx or any(i < 0 or i >= model.n_state for i in idx) or len(set(idx)) != len(idx):
            raise ValueError("save_idxs must be distinct 0-based state indices")
        d.n_save_idxs = len(idx)
        for k, i in enumerate(idx):
            d.save_idxs[k] = i
    d.want_paths, d.want_moments = int(want_paths), int(want_moments)
    d.n_traj, d.traj_offset = int(n_traj), int(traj_offset)
    d.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return m


@dataclass
class Schedule:
    t: np.ndarray
    dt: np.ndarray
    sqdt: np.ndarray
    lk_mode: np.ndarray
    lk_idx: np.ndarray
    lk_theta: np.ndarray
    save_t: np.ndarray
    save_step: np.ndarray
    save_theta: np.ndarray
    save_exact: np.ndarray
    tstops_hit: np.ndarray
    ring_slots: int
    min_gap: int

    @property
    def n_steps(self):
        return len(self.dt)


def _resolve_model(prob, ctx):
    model = prob.model
    if isinstance(model, UserModel):
        if ctx is None:
            raise RuntimeError("a UserModel needs a GPU context to be compiled")
        model = model.resolve(ctx)
    return model


def schedule(prob: SDDEProblem, alg, *, dt, saveat=None, save_everystep=None, save_start=None, save_end=None,
             tstops=(), adaptive=None, d_discontinuities=()) -> Schedule:
    """The trajectory-independent step/lookup/save schedule the library will use (host only, no GPU needed)."""
    model = prob.model
    if isinstance(model, UserModel):
        ns, nl, nw, npar, diag = model._shape
        from .problem import Model
        model = Model(model.name, capi.MODEL_USER_BASE, ns, nl, nw, npar, diag)
    m = _marshal(prob, model, alg, dt=dt, saveat=saveat, save_everystep=save_everystep, save_start=save_start,
                 save_end=save_end, tstops=tstops, adaptive=adaptive, n_traj=0, traj_offset=0, seed=0, p_batch=None,
                 u0_batch=None, noise=None, noise_mode=capi.NOISE_PHILOX, noise_mem=0, out_mem=0, want_paths=0,
                 want_moments=0, d_discontinuities=d_discontinuities)
    s = capi.SddeSchedule()
    capi.check(capi.lib().sdde_b200_schedule(C.byref(m.desc), C.byref(s)))
    try:
        N, K, L = s.n_steps, s.n_save, s.n_lag * s.n_stage

        def arr(ptr, n, dtype):
            return np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dtype, copy=True)

        return Schedule(
            t=arr(s.t, N + 1, np.float64), dt=arr(s.dt, N, np.float64), sqdt=arr(s.sqdt, N, np.float64),
            lk_mode=arr(s.lk_mode, N * L, np.int32).reshape(N, L), lk_idx=arr(s.lk_idx, N * L, np.int64).reshape(N, L),
            lk_theta=arr(s.lk_theta, N * L, np.float64).reshape(N, L),
            save_t=arr(s.save_t, K, np.float64), save_step=arr(s.save_step, K, np.int64),
            save_theta=arr(s.save_theta, K, np.float64), save_exact=arr(s.save_exact, K, np.int32),
            tstops_hit=arr(s.tstops_hit, s.n_tstops_hit, np.float64), ring_slots=s.ring_slots, min_gap=s.min_gap)
    finally:
        capi.lib().sdde_b200_schedule_free(C.byref(s))


# ---------------------------------------------------------------------------------------------
# solutions
# ---------------------------------------------------------------------------------------------
class RODESolution:
    """Lazy per-trajectory view: ``sol.t``, ``sol.u`` (list-of-states like Julia's Vector{uType}), ``sol.retcode``."""

    def __init__(self, ens, i):
        self._e, self._i = ens, i

    @property
    def t(self):
        return self._e.t

    @property
    def u(self):
        x = self._e.out[:, :, self._i]
        if hasattr(x, "cpu"):
            x = x.cpu().numpy()
        return x[:, 0] if self._e.scalar_state else x

    @property
    def retcode(self):
        st = self._e.status
        return "Success" if st is None or st[self._i] == 0 else "Unstable"

    def __getitem__(self, k):
        return self.u[k]

    def __call__(self, t):
        """``sol(t)``: the saved time series interpolated linearly ([EXT] the reference's ``sde_interpolant`` on a solution saved
        without dense output), no extrapolation.  ``t`` may be a scalar or an array of times."""
        v = self._e.at(t)[..., self._i]
        if hasattr(v, "cpu"):
            v = v.cpu().numpy()
        return v[..., 0] if self._e.scalar_state else v


def _bracket(ts, t):
    """-> (k, theta): t lies in [ts[k], ts[k+1]] with weight theta on the upper point; an exact hit gives theta = 0."""
    if not (ts[0] <= t <= ts[-1]):
        raise ValueError(f"t = {t} is outside the saved time span [{ts[0]}, {ts[-1]}] (solutions do not extrapolate)")
    k = int(np.searchsorted(ts, t, side="right")) - 1
    if ts[k] == t or k == len(ts) - 1:
        return min(k, len(ts) - 1), 0.0
    return k, float((t - ts[k]) / (ts[k + 1] - ts[k]))


class EnsembleSolution:
    """What the ensemble solve returns: dense ``out[n_save, n_state, trajectories]`` (time-major, trajectory
    contiguous) plus lazy per-trajectory ``RODESolution`` views (materialising 10^6 solution objects would cost more
    than the solve)."""

    def __init__(self, t, out, status, moments, scalar_state, elapsed, timing, traj_offset=0, n_total=None):
        self.t = t
        self.out = out
        self.status = status
        self.moments = moments          # [n_save, n_state, 3] (count, mean, M2) or None
        self.scalar_state = scalar_state
        self.elapsedTime = elapsed
        self.timing = timing
        self.traj_offset = traj_offset
        self.n_total = n_total
        self.converged = True

    def __len__(self):
        return 0 if self.out is None else self.out.shape[2]

    def __getitem__(self, i):
        if self.out is None:
            raise IndexError("paths were not requested (want_paths=False)")
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError(i)
        return RODESolution(self, i)

    @property
    def u(self):
        return _LazyList(self)

    def at(self, t):
        """All trajectories at time(s) ``t``: ``[n_state, trajectories]`` for a scalar t, ``[len(t), n_state, trajectories]`` for
        an array -- linear interpolation between the saved points (what ``sol(t)`` does per trajectory in the reference, here
        for the whole ensemble at once, on the device when the paths live there)."""
        if self.out is None:
            raise ValueError("paths were not requested (want_paths=False)")
        if np.ndim(t) > 0:
            rows = [self.at(float(x)) for x in np.asarray(t, dtype=np.float64).ravel()]
            if hasattr(rows[0], "data_ptr"):
                import torch
                return torch.stack(rows)
            return np.stack(rows)
        k, th = _bracket(np.asarray(self.t), float(t))
        if th == 0.0:
            return self.out[k]
        return (1.0 - th) * self.out[k] + th * self.out[k + 1]

    @property
    def retcodes(self):
        return None if self.status is None else np.where(self.status == 0, "Success", "Unstable")


class _LazyList:
    def __init__(self, e):
        self._e = e

    def __len__(self):
        return len(self._e)

    def __getitem__(self, i):
        return self._e[i]

    def __iter__(self):
        return (self._e[i] for i in range(len(self._e)))


# [EXT] SciMLBase.EnsembleAnalysis
def timeseries_steps_meanvar(sim: EnsembleSolution):
    """-> (mean[n_save, n_state], var[n_save, n_state]) over the WHOLE ensemble (all ranks when distributed)."""
    if sim.moments is None:
        raise ValueError("solve(..., moments=True) first")
    cnt, mean, m2 = sim.moments[..., 0], sim.moments[..., 1], sim.moments[..., 2]
    with np.errstate(invalid="ignore", divide="ignore"):
        var = m2 / (cnt - 1.0)
    return mean, var


def timeseries_steps_mean(sim):
    return timeseries_steps_meanvar(sim)[0]


def timepoint_meanvar(sim, t):
    """[EXT] EnsembleAnalysis.timepoint_meanvar(sim, t): mean and variance over the ensemble at time t -- from the moment
    triples when t is a saved point, else from the linearly interpolated paths (``sim.at(t)``, this rank's trajectories)."""
    k = int(np.argmin(np.abs(sim.t - t)))
    if sim.t[k] == t and sim.moments is not None:
        mean, var = timeseries_steps_meanvar(sim)
        return mean[k], var[k]
    x = sim.at(t)
    if hasattr(x, "cpu"):
        return x.mean(dim=-1).cpu().numpy(), x.var(dim=-1, unbiased=True).cpu().numpy()
    return x.mean(axis=-1), x.var(axis=-1, ddof=1)


def timeseries_steps_quantile(sim: EnsembleSolution, q, *, median=False, device=None, local=False):
    """[EXT] SciMLBase.EnsembleAnalysis.timeseries_steps_quantile (and _median): per saved time point and state, the
    quantile(s) ``q`` across this solution's trajectories -- Statistics.quantile's definition, selected exactly on the
    device (sdde_b200_quantiles; the paths stay in HBM when the solve ran with ``output='device'``).

    -> array [n_save, n_state, len(q)] (``[n_save, n_state]`` for a scalar q); with ``median=True`` the median is
    appended as the last entry of the last axis.  A solution that holds one rank's shard of a distributed solve yields the
    statistics of the WHOLE ensemble (collective: every rank must call; ``local=True`` restricts to this rank's paths)."""
    if sim.out is None:
        raise ValueError("solve(..., paths=True) first: order statistics need the saved paths")
    probs = np.atleast_1d(np.asarray(q, dtype=np.float64)) if q is not None else np.zeros(0)
    out = sim.out
    n_save, n_state, n_traj = out.shape
    on_dev = hasattr(out, "data_ptr")
    if on_dev:
        dev = out.device.index
        ptr = out.data_ptr()
    else:
        dev = device if device is not None else _default_device()
        out = np.ascontiguousarray(out, dtype=np.float64)
        ptr = capi.np_ptr(out).value
    ctx = capi.get_context(dev)
    ncol = probs.size + (1 if median else 0)
    res = np.empty((n_save * n_state, ncol), dtype=np.float64)
    if sim.n_total is not None and sim.n_total != n_traj and not local:
        # one rank's shard of a distributed ensemble: selection over ALL ranks (histograms all-reduced over NCCL)
        _quantiles_across_ranks(ctx, dev, ptr, on_dev, n_save * n_state, n_traj, sim.n_total, probs, median, res)
        res = res.reshape(n_save, n_state, ncol)
        return res[..., 0] if (np.isscalar(q) and not median) else res
    ms = C.c_double(0.0)
    capi.check(capi.lib().sdde_b200_quantiles(ctx.handle, ptr, capi.MEM_DEVICE if on_dev else capi.MEM_HOST, n_save * n_state,
                                              n_traj, n_traj, capi.np_ptr(probs).value if probs.size else None, probs.size,
                                              1 if median else 0, capi.np_ptr(res).value, C.byref(ms)))
    sim.timing["quantile_kernel_ms"] = ms.value
    res = res.reshape(n_save, n_state, ncol)
    if np.isscalar(q) and not median:
        return res[..., 0]
    return res


def _quantiles_across_ranks(ctx, dev, ptr, on_dev, n_rows, n_local, n_total, probs, median, res):
    """Exact order statistics of an ensemble sharded over the ranks of the torch.distributed group: per digit pass every
    rank histograms its own trajectories on its GPU, the histograms are summed with one all-reduce (NCCL over NVLink), and
    every rank narrows the same bins -- no path ever leaves its GPU."""
    import torch
    import torch.distributed as dist
    L = capi.lib()
    n_rank = C.c_int32(0)
    pp = capi.np_ptr(probs).value if probs.size else None
    capi.check(L.sdde_b200_qsel_ranks(n_total, pp, probs.size, 1 if median else 0, C.byref(n_rank)))
    tdev = torch.device(f"cuda:{dev}")
    hist = torch.empty(n_rows * n_rank.value * 2048, dtype=torch.int32, device=tdev)
    capi.check(L.sdde_b200_qsel_begin(ctx.handle, ptr if n_local else None, capi.MEM_DEVICE if on_dev else capi.MEM_HOST, n_rows,
                                      n_local, max(n_local, 1), n_total, pp, probs.size, 1 if median else 0, hist.data_ptr()))
    on_nccl = dist.get_backend() == "nccl"          # gloo (tests: two ranks sharing one GPU) reduces through the host

    def all_reduce_sum(t):
        if on_nccl:
            dist.all_reduce(t)
        else:
            h = t.cpu()
            dist.all_reduce(h)
            t.copy_(h)

    for p in range(6):
        capi.check(L.sdde_b200_qsel_local(ctx.handle, p))
        all_reduce_sum(hist)                        # counts per (row, rank, digit) over the whole ensemble
        torch.cuda.synchronize(tdev)
        capi.check(L.sdde_b200_qsel_apply(ctx.handle, p))
    nan = np.zeros(n_rows, dtype=np.int64)
    capi.check(L.sdde_b200_qsel_nan(ctx.handle, capi.np_ptr(nan).value, 0))
    t = torch.from_numpy(nan).to(tdev)
    all_reduce_sum(t)
    nan = t.cpu().numpy()
    capi.check(L.sdde_b200_qsel_nan(ctx.handle, capi.np_ptr(nan).value, 1))
    capi.check(L.sdde_b200_qsel_end(ctx.handle, capi.np_ptr(res).value))


# ---- [EXT] SciMLBase.EnsembleAnalysis: the per-time-step / per-time-point forms (same names and argument order) ----
def _rows(sim, rows):
    """A solution holding only ``rows`` ([k, n_state, trajectories]) of ``sim`` (same sharding information)."""
    return EnsembleSolution(np.arange(rows.shape[0], dtype=np.float64), rows, None, None, sim.scalar_state, 0.0, {},
                            traj_offset=sim.traj_offset, n_total=sim.n_total)


def get_timestep(sim, i):
    """All trajectories at saved point i: ``[n_state, trajectories]``."""
    return sim.out[i]


def get_timepoint(sim, t):
    """All trajectories at time t (linear interpolation between saved points): ``[n_state, trajectories]``."""
    return sim.at(t)


def timestep_meanvar(sim, i):
    if sim.moments is not None:
        mean, var = timeseries_steps_meanvar(sim)
        return mean[i], var[i]
    x = sim.out[i]
    if hasattr(x, "cpu"):
        return x.mean(dim=-1).cpu().numpy(), x.var(dim=-1, unbiased=True).cpu().numpy()
    return x.mean(axis=-1), x.var(axis=-1, ddof=1)


def timestep_mean(sim, i):
    return timestep_meanvar(sim, i)[0]


def timestep_quantile(sim, q, i, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.out[i:i + 1]), q, **kw)[0]


def timestep_median(sim, i, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.out[i:i + 1]), None, median=True, **kw)[0, :, -1]


def timepoint_mean(sim, t):
    return timepoint_meanvar(sim, t)[0]


def timepoint_quantile(sim, q, t, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.at(t)[None]), q, **kw)[0]


def timepoint_median(sim, t, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.at(t)[None]), None, median=True, **kw)[0, :, -1]


def timeseries_point_meanvar(sim, ts):
    mv = [timepoint_meanvar(sim, float(t)) for t in ts]
    return np.stack([m for m, _ in mv]), np.stack([v for _, v in mv])


def timeseries_point_mean(sim, ts):
    return timeseries_point_meanvar(sim, ts)[0]


def timeseries_point_quantile(sim, q, ts, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.at(np.asarray(ts, dtype=np.float64))), q, **kw)


def timeseries_point_median(sim, ts, **kw):
    return timeseries_steps_quantile(_rows(sim, sim.at(np.asarray(ts, dtype=np.float64))), None, median=True, **kw)[..., -1]


def timeseries_steps_median(sim: EnsembleSolution):
    return timeseries_steps_quantile(sim, None, median=True)[..., 0]


def _moments_of_paths(sim: EnsembleSolution) -> np.ndarray:
    """(count, mean, M2) per (save point, state) of a solution that was solved without ``moments=True`` (device kernels;
    a shard of a distributed solve is merged over the ranks like the solve itself does)."""
    if sim.out is None:
        raise ValueError("no saved paths")
    out = sim.out
    n_save, n_state, n_traj = out.shape
    on_dev = hasattr(out, "data_ptr")
    dev = out.device.index if on_dev else _default_device()
    if not on_dev:
        out = np.ascontiguousarray(out, dtype=np.float64)
    mom = np.zeros((n_save, n_state, 3), dtype=np.float64)
    capi.check(capi.lib().sdde_b200_moments(capi.get_context(dev).handle, out.data_ptr() if on_dev else capi.np_ptr(out).value,
                                            capi.MEM_DEVICE if on_dev else capi.MEM_HOST, n_save * n_state, n_traj, n_traj,
                                            capi.np_ptr(mom).value))
    if sim.n_total is not None and sim.n_total != n_traj:
        mom = _allgather_merge_moments(mom, dev)
    return mom


class EnsembleSummary:
    """[EXT] SciMLBase.EnsembleSummary(sim; quantiles=[0.05, 0.95]): mean ``u``, variance ``v``, median ``med`` and the
    two quantile bands ``qlow``/``qhigh`` per saved time point, all computed on the device."""

    def __init__(self, sim: EnsembleSolution, quantiles=(0.05, 0.95)):
        if len(quantiles) != 2:
            raise ValueError("quantiles must be (low, high)")
        if sim.moments is None:
            sim.moments = _moments_of_paths(sim)        # on the device, from the saved paths
        self.u, self.v = timeseries_steps_meanvar(sim)
        r = timeseries_steps_quantile(sim, list(quantiles), median=True)
        self.qlow, self.qhigh, self.med = r[..., 0], r[..., 1], r[..., 2]
        self.t = sim.t
        self.num_monte = sim.n_total or len(sim)
        self.elapsedTime = sim.elapsedTime
        self.converged = sim.converged


def merge_moments(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Chan merge of two (count, mean, M2) arrays through the library (host code of libsdde_b200)."""
    a = np.ascontiguousarray(a, dtype=np.float64).copy()
    b = np.ascontiguousarray(b, dtype=np.float64)
    capi.check(capi.lib().sdde_b200_merge_moments(capi.np_ptr(a), capi.np_ptr(b), a.size // 3))
    return a


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array backed by page-locked host memory (sdde_b200_alloc_pinned): pass it as ``solve(..., out=...)`` and reuse
    it across solves -- D2H then runs at PCIe speed (~45-50 GB/s instead of ~5 GB/s into pageable memory; allocating it
    costs ~0.4 ms/MB, so it only pays when reused).  Freed when the array (and every view of it) is garbage collected."""
    import weakref
    shape = tuple(int(x) for x in np.atleast_1d(shape))
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = capi.lib().sdde_b200_alloc_pinned(max(nbytes, 1))
    if not ptr:
        raise capi.SddeError(-100, capi.lib().sdde_b200_last_error().decode(errors="replace"))
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, capi.lib().sdde_b200_free_pinned, ptr)
    return arr


# ---------------------------------------------------------------------------------------------
# solve
# ---------------------------------------------------------------------------------------------
_sched_cache: dict = {}     # (n_save, n_steps, save_t) of the last schedule: repeated solves skip the host replay


def _default_device():
    import os
    return int(os.environ.get("LOCAL_RANK", "0"))


def _choose_seed(seed: int, prob_seed: int, across_ranks: bool) -> int:
    """seed keyword, else prob.seed, else a drawn one (src/solve.jl:302).  A drawn seed must be the SAME on every rank of a
    distributed solve -- the Philox streams are keyed by (seed, global trajectory id) -- so rank 0's draw is broadcast."""
    if seed != 0:
        return int(seed)
    if prob_seed != 0:
        return int(prob_seed)
    seed = int(np.random.SeedSequence().generate_state(2, np.uint32).view(np.uint64)[0]) or 1
    if across_ranks:
        import torch.distributed as dist
        box = [seed]
        dist.broadcast_object_list(box, src=0)
        seed = int(box[0])
    return seed


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous global-id range of rank r: [r*ceil(N/R), min(N, (r+1)*ceil(N/R)))  (SURVEY.md section 8(e))."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def solve(prob, alg, ensemblealg: EnsembleB200 | None = None, *, dt=0.0, trajectories=1, saveat=None,
          save_everystep=None, save_start=None, save_end=None, tstops=(), adaptive=None, seed=0, moments=False,
          paths=True, noise_increments=None, noise_grid=None, out=None, callback=None, save_idxs=None, traj_offset=0,
          d_discontinuities=(), **kwargs):
    """Ensemble solve on the B200 path.  ``prob`` is an ``EnsembleProblem`` (or an ``SDDEProblem`` -> one trajectory).

    noise_increments[n_steps, n_chan, trajectories]  supplied dW (validation; reference: prob.noise branch src/solve.jl:340-353)
    noise_grid[n_steps+1, n_chan, trajectories]      W on the schedule's nodes (NoiseGrid semantics), or a tuple
                   (t[n_grid], W[n_grid, n_chan, trajectories]) = NoiseGrid(t, W) on any grid covering tspan (W is
                   interpolated linearly at the step ends, like the reference's NoiseGrid)
    out                                               preallocated result buffer: numpy (ideally pinned) or torch CUDA tensor
    d_discontinuities                                 times, or (time, order) pairs, of discontinuities the user declares
                   (src/solve.jl:20-23): they become tstops and propagate by the constant lags like the one at t0
    traj_offset                                       global id of this call's first trajectory (default 0): the in-kernel Philox
                   streams are keyed by (seed, global trajectory id), so a slice [k, k+n) of a larger ensemble solved with
                   traj_offset=k reproduces those trajectories bit for bit (what the ranks of a sharded solve do internally)
    """
    if callback is not None:
        raise NotImplementedError("callbacks are outside this path (SURVEY.md section 8(f))")
    if kwargs:
        raise TypeError(f"unsupported keyword(s) on the EnsembleB200 path: {sorted(kwargs)}")
    single = isinstance(prob, SDDEProblem)
    eprob = EnsembleProblem(prob) if single else prob
    ea = ensemblealg or EnsembleB200()
    base = eprob.prob
    # EnsembleProblem.output_func / reduction -> what the device computes (julia/EnsembleB200.jl does the same mapping)
    if isinstance(eprob.output_func, FinalState):
        saveat, save_everystep, save_start, save_end = [base.tspan[1]], False, False, True
    elif isinstance(eprob.output_func, SaveIdxs):
        if save_idxs is not None and tuple(np.atleast_1d(save_idxs)) != eprob.output_func.idxs:
            raise ValueError("save_idxs keyword and SaveIdxs output_func disagree")
        save_idxs = eprob.output_func.idxs
    if isinstance(eprob.reduction, MomentsReduction):
        moments, paths = True, False
    n_total = int(trajectories)
    rank, world = 0, 1
    if ea.distributed:
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(n_total, rank, world)
    n_traj = hi - lo
    gid0 = int(traj_offset) + lo          # global id of this rank's first trajectory

    device = ea.device if ea.device is not None else (ea.devices[0] if ea.devices else _default_device())
    if isinstance(device, capi.Context):
        device = device.device
    ctx = capi.get_context(device)
    ctx.set_option("wave_traj", int(ea.wave_traj or 0))     # per solve: a forced wave size must not leak into later solves
    model = _resolve_model(base, ctx)
    if not paths and not moments:
        raise ValueError("nothing to return: ask for paths, moments or both")
    seed = _choose_seed(seed, base.seed, ea.distributed and world > 1)
    p_batch, u0_batch = eprob.batches(n_traj, gid0)      # per-trajectory data is indexed by global id

    t0w = time.perf_counter()
    noise_mode, noise_mem, noise_ptr, keep = capi.NOISE_PHILOX, capi.MEM_HOST, None, []
    grid_t = None
    if isinstance(noise_grid, tuple):                # NoiseGrid(t, W) on its own time grid
        grid_t = np.ascontiguousarray(noise_grid[0], dtype=np.float64)
        noise_grid = noise_grid[1]
        if grid_t.ndim != 1 or grid_t.size != noise_grid.shape[0]:
            raise ValueError("noise_grid=(t, W): t must be 1-d with one entry per row of W")
        keep.append(grid_t)
    nz = noise_increments if noise_increments is not None else noise_grid
    if nz is not None:
        noise_mode = capi.NOISE_SUPPLIED_DW if noise_increments is not None else capi.NOISE_SUPPLIED_GRID
        if hasattr(nz, "data_ptr"):                    # torch tensor
            if not nz.is_cuda or not nz.is_contiguous() or str(nz.dtype) != "torch.float64":
                raise ValueError("device noise must be a contiguous float64 CUDA tensor")
            noise_mem, noise_ptr = capi.MEM_DEVICE, nz.data_ptr()
        else:
            nz = np.ascontiguousarray(nz, dtype=np.float64)
            noise_ptr = capi.np_ptr(nz).value
        keep.append(nz)
        n_rows = model.n_chan * getattr(alg, "noise_mult", 1)       # SRA: dW rows, then dZ rows
        if nz.shape[1] != n_rows or nz.shape[2] != n_traj:
            raise ValueError(f"supplied noise must have shape [n, {n_rows}, {n_traj}]")

    m = _marshal(base, model, alg, dt=dt, saveat=saveat, save_everystep=save_everystep, save_start=save_start,
                 save_end=save_end, tstops=tstops, adaptive=adaptive, n_traj=n_traj, traj_offset=gid0, seed=seed,
                 p_batch=p_batch, u0_batch=u0_batch, noise=noise_ptr, noise_mode=noise_mode, noise_mem=noise_mem,
                 out_mem=capi.MEM_DEVICE if ea.output == "device" else capi.MEM_HOST, want_paths=paths,
                 want_moments=moments, lanes_per_traj=ea.lanes_per_traj, save_idxs=save_idxs, noise_t=grid_t,
                 d_discontinuities=d_discontinuities)
    d = m.desc
    skey = (d.alg, d.t0, d.tf, d.dt, tuple(d.lags[i] for i in range(d.n_lag)), d.order_discontinuity_t0,
            d.constant_lags_declared, d.save_start, d.save_end, d.save_everystep, m.keep[0].tobytes(), m.keep[1].tobytes(),
            m.keep[2].tobytes(), m.keep[3].tobytes(), d.sra_order2,
            m.keep_extra.tobytes() if getattr(m, "keep_extra", None) is not None else b"")
    if _sched_cache.get("key") != skey:
        sch = capi.SddeSchedule()
        capi.check(capi.lib().sdde_b200_schedule(C.byref(m.desc), C.byref(sch)))
        _sched_cache.update(key=skey, n_save=sch.n_save, n_steps=sch.n_steps,
                            save_t=np.ctypeslib.as_array(sch.save_t, shape=(max(sch.n_save, 1),))[:sch.n_save].copy())
        capi.lib().sdde_b200_schedule_free(C.byref(sch))
    n_save, n_steps, save_t = _sched_cache["n_save"], _sched_cache["n_steps"], _sched_cache["save_t"].copy()
    if nz is not None:
        need = n_steps + (1 if noise_mode == capi.NOISE_SUPPLIED_GRID else 0)
        if grid_t is None and nz.shape[0] < need:
            raise ValueError(f"supplied noise has {nz.shape[0]} rows, the schedule needs {need}")

    n_saved = model.n_state if save_idxs is None else len(np.atleast_1d(save_idxs))
    out_arr, out_ptr = None, None
    if paths:
        shape = (n_save, n_saved, n_traj)
        if out is not None:
            out_arr = out
            if tuple(out_arr.shape) != shape:
                raise ValueError(f"out must have shape {shape}")
        elif ea.output == "device":
            import torch
            out_arr = torch.empty(shape, dtype=torch.float64, device=f"cuda:{device}")
        else:
            out_arr = np.empty(shape, dtype=np.float64)
        if hasattr(out_arr, "data_ptr"):
            if out_arr.is_cuda != (ea.output == "device"):
                raise ValueError("out tensor residency does not match EnsembleB200(output=...)")
            if not out_arr.is_contiguous() or str(out_arr.dtype) != "torch.float64":
                raise ValueError("out must be a contiguous float64 tensor")   # the kernel writes 8 bytes per element
            if out_arr.is_cuda and out_arr.device.index != device:
                raise ValueError(f"out lives on cuda:{out_arr.device.index}, the solve runs on cuda:{device}")
            out_ptr = out_arr.data_ptr()
        else:
            if not out_arr.flags.c_contiguous or out_arr.dtype != np.float64:
                raise ValueError("out must be a C-contiguous float64 array")
            out_ptr = capi.np_ptr(out_arr).value
    m.desc.out_rows = n_save * n_saved          # capacity of out / moments: the library refuses to write more rows than this
    mom = np.zeros((n_save, n_saved, 3), dtype=np.float64) if moments else None
    # per-trajectory return codes only make sense with per-trajectory solutions; a moments-only run gets the count
    status = np.empty(n_traj, dtype=np.uint8) if paths else None
    status_ptr = capi.np_ptr(status).value if paths else None
    tm = capi.SddeTiming()
    if ea.devices and len(ea.devices) > 1:
        if ea.output != "host":
            raise ValueError("EnsembleB200(devices=...) returns host arrays")
        # an entry is a device index (this process's cached context of that GPU) or a capi.Context (two contexts may share a GPU)
        hs = (C.c_void_p * len(ea.devices))(*[dv.handle if isinstance(dv, capi.Context) else capi.get_context(dv).handle
                                               for dv in ea.devices])
        tms = (capi.SddeTiming * len(ea.devices))()
        capi.check(capi.lib().sdde_b200_solve_multi(hs, len(ea.devices), C.byref(m.desc), out_ptr,
                                                    capi.np_ptr(mom).value if moments else None,
                                                    status_ptr, C.cast(tms, C.c_void_p)))
        tm = tms[0]
        tm.n_unstable = sum(t.n_unstable for t in tms)
    else:
        capi.check(capi.lib().sdde_b200_solve(ctx.handle, C.byref(m.desc), out_ptr,
                                              capi.np_ptr(mom).value if moments else None, status_ptr,
                                              C.byref(tm)))
    if moments and ea.distributed and world > 1:
        mom = _allgather_merge_moments(mom, device)
    if hasattr(out_arr, "data_ptr") and not out_arr.is_cuda:
        out_arr = out_arr.numpy()
    sim = EnsembleSolution(save_t, out_arr, status, mom, base.scalar_state and n_saved == 1, time.perf_counter() - t0w, tm.as_dict(),
                           traj_offset=lo, n_total=n_total)
    sim.seed = seed
    sim.n_unstable = int(tm.n_unstable)          # trajectories that saw a non-finite state (ReturnCode.Unstable)
    if single:
        return sim[0] if paths else sim
    return sim


def _allgather_merge_moments(mom: np.ndarray, device: int) -> np.ndarray:
    """The path's single collective: all-gather the per-rank (count, mean, M2) triples over
    NCCL (gloo on CPU in tests) and Chan-merge them in rank order, so every rank ends with identical moments."""
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device(f"cuda:{device}") if backend == "nccl" else torch.device("cpu")
    local = torch.from_numpy(mom).to(dev)
    parts = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local)
    acc = parts[0].cpu().numpy().copy()
    for r in range(1, len(parts)):
        acc = merge_moments(acc.reshape(-1, 3), parts[r].cpu().numpy().reshape(-1, 3)).reshape(mom.shape)
    return acc
