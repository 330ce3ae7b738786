"""CPU oracle for the fixed-step constant-lag ensemble SDDE path  --  TEST INFRASTRUCTURE ONLY.

This file is a float64 restatement of what StochasticDelayDiffEq.jl v1.12.0 (``/root/reference``)
does for ``solve(prob::SDDEProblem, EM()/RKMil(); dt, adaptive=false, ...)``, one trajectory at a
time, vectorised here over trajectories with numpy (all arithmetic is element-wise, so every
trajectory sees exactly the scalar IEEE-754 operation sequence written below; numpy never fuses
multiply-add).

PARITY UNPINNED: the reference ships no golden vectors for this path and Julia
is not available in this environment, so this restatement is checked against hand-derived
known-answer values (K1..K7 in SURVEY.md section 8(c)), not against outputs of the reference itself.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product (``stochasticdelaydiffeq.jl_b200``) never does.

Reference lines followed (relative to /root/reference):
  * queues            src/solve.jl:613-713  (tstop_saveat_disc_handling)
  * first dt / sqdt   src/solve.jl:561-569
  * loop order        src/solve.jl:574-610  (loopheader! -> perform_step! -> loopfooter! -> handle_tstop!)
  * accept / next dt  src/integrators/interface.jl:1-30 + [EXT] StochasticDiffEq.apply_step!/modify_dt_for_tstops!
  * time update+snap  src/integrators/interface.jl:70-85
  * saving            src/integrators/interface.jl:357-430 (+ postamble! :100-117)
  * tstop / disc      src/integrators/utils.jl:1-22, 42-79, 91-118
  * history lookup    src/utils.jl:143-247 + src/solve.jl:407  ([EXT] DelayDiffEq.HistoryFunction,
                      StochasticDiffEq.LinearInterpolationData / sde_interpolation)
  * where h enters    src/functionwrapper.jl:8-9,35-36
  * EM / RKMil update [EXT] StochasticDiffEq perform_step! (call site src/solve.jl:586), SURVEY.md section 3.3
"""
from __future__ import annotations

import bisect
import heapq
import math
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np

# ----------------------------------------------------------------------------------------------
# algorithm ids (shared vocabulary with include/sdde_b200.h)
# ----------------------------------------------------------------------------------------------
ALG_EM_SPLIT = 0      # EM()  == EM(true)   [EXT] default split-step
ALG_EM_NOSPLIT = 1    # EM(false)
ALG_RKMIL_ITO = 2     # RKMil()  (interpretation = :Ito)
ALG_RKMIL_STRAT = 3   # RKMil(interpretation = :Stratonovich)
ALG_EULER_HEUN = 4    # EulerHeun()  (Stratonovich; second f/g evaluation at t + dt)
ALG_RKMILC_ITO = 5    # RKMilCommute()  -- commutative noise; on diagonal noise the RKMil update
ALG_RKMILC_STRAT = 6  # RKMilCommute(interpretation = :Stratonovich)
ALG_WANGLI_A, ALG_WANGLI_B, ALG_WANGLI_C, ALG_WANGLI_D, ALG_WANGLI_E, ALG_WANGLI_F = 7, 8, 9, 10, 11, 12   # WangLi3SMil_A .. _F
# stage operators of WangLi3SMil_x in execution order: 'D' drift, 'C' Ito correction, 'M' Stratonovich-Milstein diffusion
WANGLI_ORDER = {7: "DCM", 8: "DMC", 9: "CDM", 10: "CMD", 11: "MDC", 12: "MCD"}

# SRA1(), SRA3(), SOSRA(), SRA(tableau=...) with adaptive=false: Roessler's stochastic Runge-Kutta methods for ADDITIVE noise
# (the reference's smoke tests test/test_prob_sol.jl:74-87 run them adaptively; this path is fixed-step).  They draw a second
# increment dZ per step (src/solve.jl:308-312): the noise array carries 2 n_chan rows per step, dW first, then dZ.
ALG_SRA1, ALG_SRA3, ALG_SOSRA, ALG_SRA_TABLEAU = 13, 14, 15, 16
SRA_ALGS = (ALG_SRA1, ALG_SRA3, ALG_SOSRA, ALG_SRA_TABLEAU)
SRA_MAX_SETS = 4      # lookup sets per step in the schedule tables (distinct evaluation times, padded with t_n)


@dataclass
class SraTableau:
    """[EXT] StochasticDiffEq RosslerSRA: c0, c1, A0[i][j] / B0[i][j] (stage i from earlier stage j), alpha, beta1, beta2."""
    c0: tuple
    c1: tuple
    A0: tuple
    B0: tuple
    alpha: tuple
    beta1: tuple
    beta2: tuple

    @property
    def stages(self):
        return len(self.alpha)

    def flat(self):
        """Layout of sdde_desc.sra_tableau: [c0 | c1 | A0 row-major | B0 row-major | alpha | beta1 | beta2]."""
        out = list(self.c0) + list(self.c1)
        for M in (self.A0, self.B0):
            for row in M:
                out += list(row)
        return np.array(out + list(self.alpha) + list(self.beta1) + list(self.beta2), dtype=np.float64)

    def stage_sets(self):
        """Deduplicated evaluation times (set 0 = t_n itself) and the set each f / g evaluation uses."""
        cset = [0.0]

        def idx(c):
            if c not in cset:
                cset.append(c)
            return cset.index(c)
        fset, gset = [], []
        for i in range(self.stages):              # first-seen order over f_0, g_0, f_1, g_1, ...
            fset.append(idx(self.c0[i]))
            gset.append(idx(self.c1[i]))
        return cset, fset, gset


# [EXT] constructSRA1 / constructSRA3 (Roessler 2010, strong order 1.5 for additive noise) and constructSOSRA (Rackauckas &
# Nie 2018, stability-optimised; coefficients restated from memory -- the order conditions they must satisfy are checked
# in tests/test_oracle.py)
SRA_TABLEAUS = {
    ALG_SRA1: SraTableau(c0=(0.0, 0.75), c1=(1.0, 0.0), A0=((0.0, 0.0), (0.75, 0.0)), B0=((0.0, 0.0), (1.5, 0.0)),
                         alpha=(1.0 / 3.0, 2.0 / 3.0), beta1=(1.0, 0.0), beta2=(-1.0, 1.0)),
    ALG_SRA3: SraTableau(c0=(0.0, 1.0, 0.5), c1=(1.0, 0.0, 0.0),
                         A0=((0.0, 0.0, 0.0), (1.0, 0.0, 0.0), (0.25, 0.25, 0.0)),
                         B0=((0.0, 0.0, 0.0), (0.0, 0.0, 0.0), (1.0, 0.5, 0.0)),
                         alpha=(1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0), beta1=(1.0, 0.0, 0.0), beta2=(-1.0, 1.0, 0.0)),
    ALG_SOSRA: SraTableau(c0=(0.0, 0.6923962376159507, 1.0), c1=(0.0, 0.041248171110700504, 1.0),
                          A0=((0.0, 0.0, 0.0), (0.6923962376159507, 0.0, 0.0), (-3.1609142252828395, 4.1609142252828395, 0.0)),
                          B0=((0.0, 0.0, 0.0), (1.3371632704399763, 0.0, 0.0), (1.442371048468624, 1.8632741501139225, 0.0)),
                          alpha=(0.2889874966892885, 0.6859880440839937, 0.025024459226717772),
                          beta1=(-16.792534242221663, 17.514995785380226, 0.27753845684143835),
                          beta2=(0.4237535769069274, 0.6010381474428539, -1.0247917243497813)),
}


def sra_tableau(alg, tableau=None) -> SraTableau:
    if alg == ALG_SRA_TABLEAU:
        if tableau is None:
            raise ValueError("ALG_SRA_TABLEAU needs a tableau")
        return tableau
    return SRA_TABLEAUS[alg]


# [EXT] StochasticDiffEq.alg_order: EM 1//2, RKMil 1//1, SRA family 2//1  (call site src/solve.jl:248)
ALG_ORDER = {
    ALG_EM_SPLIT: Fraction(1, 2),
    ALG_EM_NOSPLIT: Fraction(1, 2),
    ALG_RKMIL_ITO: Fraction(1, 1),
    ALG_RKMIL_STRAT: Fraction(1, 1),
    ALG_EULER_HEUN: Fraction(1, 2),
    ALG_RKMILC_ITO: Fraction(1, 1),
    ALG_RKMILC_STRAT: Fraction(1, 1),
    **{a: Fraction(1, 1) for a in range(7, 13)},
    **{a: Fraction(2, 1) for a in SRA_ALGS},
}
N_STAGE = {ALG_EM_SPLIT: 1, ALG_EM_NOSPLIT: 1, ALG_RKMIL_ITO: 1, ALG_RKMIL_STRAT: 1, ALG_EULER_HEUN: 2, ALG_RKMILC_ITO: 1,
           ALG_RKMILC_STRAT: 1, **{a: 1 for a in range(7, 13)}, **{a: SRA_MAX_SETS for a in SRA_ALGS}}

# lookup modes in the schedule tables
LK_PRE = 0      # t < t0            -> user history h(p, t)
LK_EXACT = 1    # t == sol.t[i]     -> sol.u[i]
LK_INTERP = 2   # t in (t[i-1],t[i]) -> (1-theta) u[i-1] + theta u[i]
LK_U0 = 3       # t > sol.t[end]    -> the never-advanced history integrator's u = u0


def _eps(x: float) -> float:
    """Julia ``eps(x::Float64)`` (spacing of floats at |x|)."""
    return math.ulp(abs(x))


# ----------------------------------------------------------------------------------------------
# S -- step schedule (trajectory independent)
# ----------------------------------------------------------------------------------------------
@dataclass
class Schedule:
    t: np.ndarray            # [N+1] node times, t[0] = t0
    dt: np.ndarray           # [N]   dt_n used by step n (after modify_dt_for_tstops!)
    sqdt: np.ndarray         # [N]   sqrt(|dt_n|)
    lk_mode: np.ndarray      # [N, L] LK_*   ([N, n_stage*L] for two-stage algorithms, stage-major)
    lk_idx: np.ndarray       # [N, L] node index i (upper node for INTERP), -1 for PRE
    lk_theta: np.ndarray     # [N, L] theta for INTERP; lookup time for PRE
    lk_time: np.ndarray      # [N, L] t_n - lag (as computed in float64)
    save_t: np.ndarray       # [n_save] saved times in output order
    save_step: np.ndarray    # [n_save] step n whose end produced the save (-1 = initial condition)
    save_theta: np.ndarray   # [n_save] theta (nan-free): value=(1-th)*u_n + th*u_{n+1}
    save_exact: np.ndarray   # [n_save] 1 -> take u_{n+1} verbatim (or u0 for step -1)
    tstops_hit: list = field(default_factory=list)

    @property
    def n_steps(self) -> int:
        return len(self.dt)


def expand_saveat(saveat, t0: float, tf: float):
    """Return (interior+end save times, save_start, save_end) following src/solve.jl:32-34, 411-415, 659-677.

    ``saveat`` may be ``None``/empty (=> no saveat; caller decides save_everystep), a number or a sequence.
    For a number ``s`` the reference pushes ``(t0+s):s:tf`` when ``(t0:s:tf)[end] == tf`` else
    ``(t0+s):s:(tf-s)`` (src/solve.jl:662-671); the ranges are Julia Float64 ranges, restated in ``_float_range``.
    """
    if saveat is None:
        return np.zeros(0), True, True
    if np.isscalar(saveat):
        s = float(saveat)
        if s <= 0:
            raise ValueError("saveat must be positive")
        probe = _float_range(t0, s, tf)         # t0:s:tf
        if len(probe) > 0 and probe[-1] == tf:
            sv = _float_range(t0 + s, s, tf)
        else:
            sv = _float_range(t0 + s, s, tf - s)
        return np.asarray(sv, dtype=np.float64), True, True
    arr = np.asarray(list(saveat), dtype=np.float64)
    if arr.size == 0:
        return np.zeros(0), True, True
    save_start = bool(np.any(arr == t0))
    save_end = bool(np.any(arr == tf))
    interior = np.sort(arr[(arr > t0) & (arr < tf)])
    return interior, save_start, save_end


def _rational_lift(x: float):
    """[EXT] Base.rat(x) of julia/base/twiceprecision.jl (Julia 1.10): convergents of the continued fraction of x while
    numerator and denominator stay <= 2^24 (= maxintfloat(Float32)); (c, d) = previous convergent on overflow."""
    LIM = 1 << 24
    y = x
    a, b, c, d = 1, 0, 0, 1
    while abs(y) <= float(LIM):
        f = int(y)                      # trunc toward zero
        y = y - f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if abs(a) > LIM or abs(b) > LIM:
            return c, d
        if b != 0 and float(a) / float(b) == x:
            return a, b
        y = 1.0 / y if y != 0.0 else float("inf")
    return a, b


def _float_range(a: float, s: float, b: float) -> np.ndarray:
    """Elements of the Julia range ``a:s:b`` for Float64 ([EXT] Base.:(start, step, stop), twiceprecision.jl; MED-HIGH
    confidence, restated from knowledge of Julia 1.10 -- K-check: 0.1:0.1:0.5 == [0.1, 0.2, 0.3, 0.4, 0.5] exactly).

    Rational branch: start, step, stop all equal their small-rational lifts -> integer numerators over a common
    denominator, length by integer division, elements (n0 + k ns)/den rounded once (twice-precision evaluation).
    Literal branch: length = round((b-a)/s) + 1 (minus one on overshoot), elements a + k s rounded once."""
    from fractions import Fraction
    if s == 0.0:
        raise ValueError("range step cannot be zero")

    def inside(p, x, q):
        return (p <= x <= q) or (q <= x <= p)

    sn, sd = _rational_lift(s)
    if sd != 0 and sn / sd == s:
        an, ad = _rational_lift(a)
        bn, bd = _rational_lift(b)
        if ad != 0 and bd != 0 and an / ad == a and bn / bd == b:
            den = ad * (sd // math.gcd(ad, sd))
            big = float(1 << 53)
            if den != 0 and abs(a * den) <= big and abs(s * den) <= big and den % ad == 0 and den % sd == 0:
                n0, ns = int(round(a * den)), int(round(s * den))
                bot = ns * bd
                top = den * bn - bd * n0 + bot       # len = div(den*stop_n - stop_d*start_n + step_n*stop_d, step_n*stop_d)
                quo = abs(top) // abs(bot)
                if (top < 0) != (bot < 0):
                    quo = -quo                       # Julia's div truncates
                n = max(0, quo)
                if inside(a, a + (n - 1) * s, b + s / 2) and not inside(a, a + n * s, b):
                    return np.array([float(Fraction(n0 + k * ns, den)) for k in range(n)], dtype=np.float64)
    lf = (b - a) / s
    if lf < 0:
        n = 0
    elif lf == 0:
        n = 1
    else:
        n = int(round(lf)) + 1
        end = a + (n - 1) * s
        if (a < b < end) or (a > b > end):
            n -= 1
    A, S_ = Fraction(a), Fraction(s)
    return np.array([float(A + k * S_) for k in range(n)], dtype=np.float64)


def build_schedule(t0, tf, dt, lags, alg, *, order_discontinuity_t0=0, constant_lags_declared=True,
                   saveat_times=(), save_start=True, save_end=True, save_everystep=False,
                   tstops=(), neutral=False, maxsteps=50_000_000, d_discontinuities=(), tableau=None,
                   alg_order=None) -> Schedule:
    """Replay the reference's (trajectory-independent) time stepping in float64.

    Follows the queue construction of src/solve.jl:613-713 and the loop of src/solve.jl:574-610 with
    the non-adaptive branch of loopfooter! (src/integrators/interface.jl:70-85), handle_tstop! /
    handle_discontinuities! / add_next_discontinuities! (src/integrators/utils.jl).
    """
    t0 = float(t0); tf = float(tf); dt = float(dt)
    if not tf > t0:
        raise ValueError("only forward integration (tdir=+1) is supported on this path")
    lags = [float(x) for x in lags]
    alg_order = ALG_ORDER[alg] if alg_order is None else Fraction(alg_order)
    order0 = Fraction(order_discontinuity_t0)
    if dt == 0.0 and len(tstops) == 0:
        # src/solve.jl:121-123
        raise ValueError("Fixed timestep methods require a choice of dt or choosing the tstops")

    # ---- src/solve.jl:625-658 : tstops heap -------------------------------------------------
    add_prop = (order0 <= alg_order) and constant_lags_declared and len(lags) > 0
    tst: list[float] = []
    disc: list[tuple[float, Fraction]] = []
    if add_prop:
        maxlag = tf - t0
        next_order = order0 if neutral else order0 + Fraction(1, 2)
    # d_discontinuities keyword (src/solve.jl:20-23): (t, order) pairs, or bare times = order 0
    ddisc = [(float(d[0]), Fraction(d[1])) if isinstance(d, (tuple, list)) else (float(d), Fraction(0)) for d in d_discontinuities]
    if not ddisc and not add_prop and len(tstops) == 0:
        heapq.heappush(tst, tf)
    else:
        for x in tstops:
            x = float(x)
            if t0 < x <= tf:
                heapq.heappush(tst, x)
        for (x, _) in ddisc:                       # src/solve.jl:643-646
            if t0 < x <= tf:
                heapq.heappush(tst, x)
        if add_prop:
            for lag in lags:
                if lag < maxlag:
                    heapq.heappush(tst, t0 + lag)
        heapq.heappush(tst, tf)
    # ---- src/solve.jl:701-710 : discontinuity heap ------------------------------------------
    for (x, od) in ddisc:                          # src/solve.jl:690-699
        if t0 < x < tf and od <= alg_order + 1:
            heapq.heappush(disc, (x, od))
    if add_prop:
        for lag in lags:
            if lag < maxlag:
                heapq.heappush(disc, (t0 + lag, next_order))
    # ---- saveat heap (already expanded by the caller: src/solve.jl:659-677) -------------------
    sv = [float(x) for x in saveat_times]
    heapq.heapify(sv)

    ts = [t0]
    dts: list[float] = []
    save_t: list[float] = []
    save_step: list[int] = []
    save_theta: list[float] = []
    save_exact: list[int] = []
    hits: list[float] = []
    if save_start:
        save_t.append(t0); save_step.append(-1); save_theta.append(1.0); save_exact.append(1)

    t = t0
    n = 0
    while tst:
        while t < tst[0]:
            # modify_dt_for_tstops! [EXT] (called src/solve.jl:564 and from apply_step!):
            top = tst[0]
            dtn = min(dt, top - t) if dt != 0.0 else top - t
            # loopfooter! non-adaptive (src/integrators/interface.jl:70-80)
            ttmp = t + dtn
            tprev = t
            if abs(ttmp - top) < 10.0 * _eps(max(t, top)):
                t = top
            else:
                t = ttmp
            dts.append(dtn)
            ts.append(t)
            # savevalues! (src/integrators/interface.jl:357-430)
            while sv and sv[0] <= t:
                curt = heapq.heappop(sv)
                save_t.append(curt); save_step.append(n)
                if curt != t:
                    save_theta.append((curt - tprev) / dtn); save_exact.append(0)
                else:
                    save_theta.append(1.0); save_exact.append(1)
            if save_everystep:
                save_t.append(t); save_step.append(n); save_theta.append(1.0); save_exact.append(1)
            n += 1
            if n > maxsteps:
                raise RuntimeError("schedule exceeds maxsteps")
            if not tst:
                break
        # handle_tstop! (src/integrators/utils.jl:1-22)
        if tst:
            top = tst[0]
            if t == top:
                heapq.heappop(tst)
                hits.append(t)
            elif t > top:
                raise RuntimeError("stepped past a tstop (dtchangeable algorithm) -- reference errors here")
            if disc:
                # handle_discontinuities! (src/integrators/utils.jl:42-79)
                d = heapq.heappop(disc)
                order = d[1]
                while disc and disc[0][0] == t:
                    order = min(order, heapq.heappop(disc)[1])
                maxdt = 10.0 * _eps(t)
                while disc and abs(disc[0][0] - t) < maxdt:
                    order = min(order, heapq.heappop(disc)[1])
                while tst and abs(tst[0] - t) < maxdt:
                    heapq.heappop(tst)
                # add_next_discontinuities! (src/integrators/utils.jl:91-118)
                nxt = order if neutral else order + Fraction(1, 2)
                if nxt <= alg_order + Fraction(1, 2) and constant_lags_declared:
                    maxlag = tf - t
                    for lag in lags:
                        if lag < maxlag:
                            heapq.heappush(disc, (t + lag, nxt))
                            heapq.heappush(tst, t + lag)
    # postamble! -> solution_endpoint_match_cur_integrator! (src/integrators/interface.jl:100-117)
    if save_end and (len(save_t) == 0 or save_t[-1] != t):
        save_t.append(t); save_step.append(n - 1); save_theta.append(1.0); save_exact.append(1)

    N = len(dts)
    L = len(lags)
    NSTG = N_STAGE[alg]
    stage_c = [float(sg) for sg in range(NSTG)]          # EulerHeun: sets at t_n and t_n + dt_n
    if alg in SRA_ALGS:                                  # lookup set k at t_n + c_k dt_n, padded with t_n
        cset, _, _ = sra_tableau(alg, tableau).stage_sets()
        if len(cset) > SRA_MAX_SETS:
            raise ValueError("the SRA tableau has more than 4 distinct stage times")
        stage_c = cset + [0.0] * (SRA_MAX_SETS - len(cset))
    tarr = np.array(ts, dtype=np.float64)
    lk_mode = np.zeros((N, NSTG * L), dtype=np.int32)
    lk_idx = np.full((N, NSTG * L), -1, dtype=np.int64)
    lk_theta = np.zeros((N, NSTG * L), dtype=np.float64)
    lk_time = np.zeros((N, NSTG * L), dtype=np.float64)
    for k in range(N):
        tn = ts[k]
        for sg in range(NSTG):
            tq = tn + stage_c[sg] * dts[k]               # EulerHeun's second evaluation: f(tmp, p, t + dt)
            for l, lag in enumerate(lags):
                e = sg * L + l
                tl = tq - lag
                lk_time[k, e] = tl
                if tl < t0:
                    lk_mode[k, e] = LK_PRE; lk_theta[k, e] = tl
                    continue
                if tl > tn:                              # past sol.t[end] = t_n: constant u0 (HistoryFunction regime 3)
                    lk_mode[k, e] = LK_U0
                    continue
                # sde_interpolation on sol.t[1..k+1] (dense history: sol.t[end] == t_n)
                i = bisect.bisect_left(ts, tl, 0, k + 1)         # searchsortedfirst
                if ts[i] == tl:
                    lk_mode[k, e] = LK_EXACT; lk_idx[k, e] = i; lk_theta[k, e] = 1.0
                else:
                    h_ = ts[i] - ts[i - 1]
                    lk_mode[k, e] = LK_INTERP; lk_idx[k, e] = i
                    lk_theta[k, e] = (tl - ts[i - 1]) / h_
    dta = np.array(dts, dtype=np.float64)
    return Schedule(t=tarr, dt=dta, sqdt=np.sqrt(np.abs(dta)), lk_mode=lk_mode, lk_idx=lk_idx,
                    lk_theta=lk_theta, lk_time=lk_time,
                    save_t=np.array(save_t, dtype=np.float64), save_step=np.array(save_step, dtype=np.int64),
                    save_theta=np.array(save_theta, dtype=np.float64),
                    save_exact=np.array(save_exact, dtype=np.int32), tstops_hit=hits)


# ----------------------------------------------------------------------------------------------
# model zoo: f(u, ulag, p, t), g(u, ulag, p, t), shapes u [ns, nt], ulag [nl, ns, nt]
# p is a sequence of scalars or of [nt] arrays (per-trajectory parameters)
# ----------------------------------------------------------------------------------------------
MODEL_LINEAR1 = 0
MODEL_HAYES = 1
MODEL_MACKEY_GLASS = 2
MODEL_LINSYS = 3
MODEL_SIR = 4
MODEL_SCALAR_ADD = 5


def ipow(x, n: int):
    """x**n by right-to-left binary exponentiation; the SAME multiplication order is used by the
    C oracle and by the CUDA model (so all three agree bit-for-bit when FMA contraction is off)."""
    n = int(n)
    result = None
    base = x
    while n > 0:
        if n & 1:
            result = base if result is None else result * base
        n >>= 1
        if n:
            base = base * base
    return np.ones_like(x) if result is None else result


@dataclass
class Model:
    mid: int
    n_state: int
    n_lag: int
    n_noise: int
    diagonal: bool
    n_param: int
    f: callable
    g: callable          # diagonal: returns [ns, nt]; non-diagonal: returns [ns, m, nt]


def _linear1_f(u, ul, p, t):
    a, b, s = p
    return a * u + b * ul[0]


def _linear1_g(u, ul, p, t):
    a, b, s = p
    return s * u


def _hayes_f(u, ul, p, t):
    # test/test_prob_sol.jl:6-9  du = a*u + b*h(p,t-tau) + c
    tau, a, b, c, al, be, ga = p
    return (a * u + b * ul[0]) + c


def _hayes_g(u, ul, p, t):
    # test/test_prob_sol.jl:10-13 du = alpha*u + beta*h(p,t-tau) + gamma  (beta=0 gives the convergence-test g)
    tau, a, b, c, al, be, ga = p
    return (al * u + be * ul[0]) + ga


def _mg_f(u, ul, p, t):
    beta, gamma, nexp, sigma = p
    x = ul[0]
    xn = ipow(x, int(np.asarray(nexp).reshape(-1)[0]))
    return (beta * x) / (1.0 + xn) - gamma * u


def _mg_g(u, ul, p, t):
    beta, gamma, nexp, sigma = p
    return sigma * u


def make_linsys(ns: int, nl: int, m: int) -> Model:
    """du = (A u + sum_l B_l u(t-tau_l) + c) dt + G(u) dW,  G_ij = M_ij*u_i + N_ij.
    params (row-major): A[ns*ns], B_0..B_{nl-1}[ns*ns each], c[ns], M[ns*m], N[ns*m]."""
    npar = ns * ns * (1 + nl) + ns + 2 * ns * m

    def f(u, ul, p, t):
        out = []
        for i in range(ns):
            acc = p[i * ns + 0] * u[0]
            for j in range(1, ns):
                acc = acc + p[i * ns + j] * u[j]
            for l in range(nl):
                off = ns * ns * (1 + l)
                for j in range(ns):
                    acc = acc + p[off + i * ns + j] * ul[l][j]
            acc = acc + p[ns * ns * (1 + nl) + i]
            out.append(acc)
        return np.stack(out)

    def g(u, ul, p, t):
        offM = ns * ns * (1 + nl) + ns
        offN = offM + ns * m
        rows = []
        for i in range(ns):
            rows.append(np.stack([p[offM + i * m + j] * u[i] + p[offN + i * m + j] for j in range(m)]))
        return np.stack(rows)            # [ns, m, nt]

    return Model(MODEL_LINSYS, ns, nl, m, False, npar, f, g)


def _sir_f(u, ul, p, t):
    # test/nondiagonal_sparse_noise.jl:5-19
    beta, c, tau = p
    S, I, R = u
    N = (S + I) + R
    infection = ((beta * c) * I) / N * S
    Sd, Id, Rd = ul[0]
    Nd = (Sd + Id) + Rd
    recovery = ((beta * c) * Id) / Nd * Sd
    return np.stack([-infection, infection - recovery, recovery])


def _sir_g(u, ul, p, t):
    # test/nondiagonal_sparse_noise.jl:31-43, 3x2 pattern rows [1,0],[1,1],[0,1]
    beta, c, tau = p
    S, I, R = u
    N = (S + I) + R
    infection = ((beta * c) * I) / N * S
    Sd, Id, Rd = ul[0]
    Nd = (Sd + Id) + Rd
    recovery = ((beta * c) * Id) / Nd * Sd
    si = np.sqrt(np.maximum(0.0, infection))
    sr = np.sqrt(np.maximum(0.0, recovery))
    z = np.zeros_like(si)
    return np.stack([np.stack([-si, z]), np.stack([si, -sr]), np.stack([z, sr])])


def _sadd_f(u, ul, p, t):
    # test/events.jl:4  f = 0.3*h(p,t-p)
    a, b = p
    return a * ul[0]


def _sadd_g(u, ul, p, t):
    a, b = p
    return b + 0.0 * u


MODELS = {
    MODEL_LINEAR1: Model(MODEL_LINEAR1, 1, 1, 1, True, 3, _linear1_f, _linear1_g),
    MODEL_HAYES: Model(MODEL_HAYES, 1, 1, 1, True, 7, _hayes_f, _hayes_g),
    MODEL_MACKEY_GLASS: Model(MODEL_MACKEY_GLASS, 1, 1, 1, True, 4, _mg_f, _mg_g),
    MODEL_SIR: Model(MODEL_SIR, 3, 1, 2, False, 3, _sir_f, _sir_g),
    MODEL_SCALAR_ADD: Model(MODEL_SCALAR_ADD, 1, 1, 1, True, 2, _sadd_f, _sadd_g),
}


def get_model(mid: int, n_state=None, n_lag=None, n_noise=None) -> Model:
    if mid == MODEL_LINSYS:
        return make_linsys(int(n_state), int(n_lag), int(n_noise))
    return MODELS[mid]


# ----------------------------------------------------------------------------------------------
# N -- noise: Philox4x32-10 (Salmon et al. 2011, Random123) + Box-Muller in float64
# ----------------------------------------------------------------------------------------------
PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs are uint32-valued arrays (any broadcastable shapes)."""
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _box_muller_pair(wa, wb):
    """One Box-Muller pair from two Philox words (same bit allocation as csrc/sdde_kernels.cuh::normal_pair_from_words):
      radius:  m = (wb & 0xFF)<<32 | wa (40 bits), u = (m + 0.5) * 2^-40 in (0,1), r = sqrt(-2 ln u)
      angle :  q = wb >> 30 (quadrant), f = (wb >> 8) & 0x3FFFFF (22 bits), x = (f + 0.5) * 2^-22 - 1/2,
               theta = (pi/2) (q + x);   z0 = r cos(theta), z1 = r sin(theta)."""
    m = ((wb & np.uint64(0xFF)) << np.uint64(32)) | wa
    u = (m.astype(np.float64) + 0.5) * 2.0 ** -40
    r = np.sqrt(-2.0 * np.log(u))
    q = (wb >> np.uint64(30)).astype(np.int64)
    f = (wb >> np.uint64(8)) & np.uint64(0x3FFFFF)
    x = (f.astype(np.float64) + 0.5) * 2.0 ** -22 - 0.5
    s, c = np.sin((np.pi / 2) * x), np.cos((np.pi / 2) * x)
    # rotate by the quadrant exactly: q=0 (c,s), 1 (-s,c), 2 (-c,-s), 3 (s,-c)
    cq = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s)))
    sq = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c)))
    return r * cq, r * sq


def philox_normal_quad(traj_id, call, seed):
    """Four N(0,1) float64 variates for (global trajectory id, call index): one Philox4x32-10 block with
    counter = (traj_lo, traj_hi, call_lo, call_hi), key = (seed_lo, seed_hi); pair A from (w0, w1), pair B from (w2, w3).
    The kernel evaluates the same definition with range-restricted polynomials, here numpy's log/sin/cos are used, so
    the two agree to a few ulp, not bit for bit."""
    traj_id = np.asarray(traj_id, dtype=np.uint64)
    call = np.asarray(call, dtype=np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    w0, w1, w2, w3 = philox4x32_10(traj_id & _MASK32, traj_id >> np.uint64(32),
                                   call & _MASK32, call >> np.uint64(32),
                                   seed & 0xFFFFFFFF, seed >> 32)
    z0, z1 = _box_muller_pair(w0, w1)
    z2, z3 = _box_muller_pair(w2, w3)
    return z0, z1, z2, z3


def philox_normals(traj_ids, n_steps: int, n_noise: int, seed) -> np.ndarray:
    """xi[n, c, traj]: normal number q = n*n_noise + c comes from call q>>2, lane q&3."""
    traj_ids = np.asarray(traj_ids, dtype=np.uint64)
    nq = n_steps * n_noise
    ncall = (nq + 3) // 4
    calls = np.arange(ncall, dtype=np.uint64)[:, None]
    zs = philox_normal_quad(traj_ids[None, :], calls, seed)
    z = np.empty((ncall * 4, traj_ids.size), dtype=np.float64)
    for i in range(4):
        z[i::4] = zs[i]
    return z[:nq].reshape(n_steps, n_noise, traj_ids.size)


# ----------------------------------------------------------------------------------------------
# H + one-step algebra + V : the trajectory loop
# ----------------------------------------------------------------------------------------------
def _as_params(params, n_traj):
    """-> list of np arrays each shape [] or [nt]."""
    p = np.asarray(params, dtype=np.float64)
    if p.ndim == 1:
        return [np.float64(x) for x in p]
    assert p.shape[1] == n_traj
    return [p[i] for i in range(p.shape[0])]


def _history_affine(h0, h1, t, nt):
    """h(p,t)[j] = h0[j] + h1[j]*t ; returns [ns, nt]."""
    h0 = np.asarray(h0, dtype=np.float64)
    h1 = np.asarray(h1, dtype=np.float64)
    v = h0 + h1 * t
    return np.repeat(v[:, None], nt, axis=1)


def solve_ensemble(model: Model, params, u0, h0, h1, lags, sched: Schedule, alg, dW, *,
                   history_mode="dense", history_fn=None, return_dense=False, tableau=None):
    """Integrate ``n_traj`` trajectories with the supplied increments ``dW[n, c, traj]``.

    u0: [ns] or [ns, nt].  Returns (out[n_save, ns, nt], status[nt]) (+ dense [N+1, ns, nt]).
    history_mode:
      'dense'                  -- delayed state from the full step history (what save_everystep=true does;
                                   the semantics BASELINE.json's north_star describes)
      'reference_saveat_alias' -- delayed state from the SAVED points only, constant u0 past the last saved
                                   point (src/solve.jl:465-482 aliasing)
    """
    dW = np.asarray(dW, dtype=np.float64)
    N = sched.n_steps
    nt = dW.shape[2]
    ns, nl, m = model.n_state, model.n_lag, model.n_noise
    nc = ns if model.diagonal else m
    assert dW.shape[0] >= N and dW.shape[1] == (2 * nc if alg in SRA_ALGS else nc)
    p = _as_params(params, nt)
    u0a = np.asarray(u0, dtype=np.float64)
    u = np.repeat(u0a[:, None], nt, axis=1).copy() if u0a.ndim == 1 else u0a.copy()
    u0full = u.copy()
    t0 = float(sched.t[0])
    if history_fn is None:
        def history_fn(pp, t):
            return _history_affine(h0, h1, t, nt)

    hist_t = [t0]
    hist_u = [u.copy()]
    out = np.empty((len(sched.save_t), ns, nt), dtype=np.float64)
    # saved-solution arrays (used as history in alias mode)
    sol_t: list[float] = []
    sol_u: list[np.ndarray] = []
    ks = 0
    while ks < len(sched.save_t) and sched.save_step[ks] == -1:
        out[ks] = u; sol_t.append(t0); sol_u.append(u.copy()); ks += 1
    if history_mode == "reference_saveat_alias" and not sol_t:
        sol_t.append(t0); sol_u.append(u.copy())

    def lookup(tl):
        if history_mode == "dense":
            ht, hu = hist_t, hist_u
        else:
            ht, hu = sol_t, sol_u
        if tl < ht[0]:
            return history_fn(p, tl)                                  # prehistory
        if tl <= ht[-1]:
            i = bisect.bisect_left(ht, tl)                            # searchsortedfirst
            if ht[i] == tl:
                return hu[i]
            th = (tl - ht[i - 1]) / (ht[i] - ht[i - 1])
            return (1.0 - th) * hu[i - 1] + th * hu[i]                # [EXT] linear sde_interpolant
        return u0full                                                 # never-advanced stub integrator: u0

    for n in range(N):
        t = float(sched.t[n]); dt = float(sched.dt[n]); sq = float(sched.sqdt[n])
        ul = [lookup(t - float(lag)) for lag in lags]
        dw = dW[n]
        fu = model.f(u, ul, p, t)
        K = u + dt * fu
        if alg in SRA_ALGS:
            # [EXT] StochasticDiffEq SRAConstantCache (perform_step! of sri/sra methods, additive noise).  With
            # chi2 = (dW + dZ/sqrt(3))/2 = I_(1,0)/dt:
            #   H_i = uprev + dt sum_{j<i} A0[i][j] f(H_j, t + c0_j dt) + (sum_{j<i} B0[i][j] g(H_j, t + c1_j dt)) chi2
            #   u+  = uprev + dt sum_i alpha_i f_i + (sum_i beta1_i g_i) dW + (sum_i beta2_i g_i) chi2
            # Through the SDDE wrappers (src/functionwrapper.jl:8-9,35-36) every f / g call looks its delayed states up at
            # its own time argument, in the history that still ends at t_n.  (SRA1 has a hand-unrolled perform_step! in the
            # reference: same quantities, sums associated differently -- rounding-level differences only.)
            tab = sra_tableau(alg, tableau)
            dwv, dz = dw[:nc], dw[nc:2 * nc]
            chi2 = 0.5 * (dwv + dz / math.sqrt(3.0))

            def times(G, w):                                          # G w: diagonal or matrix noise
                if model.diagonal:
                    return G * w
                acc = G[:, 0, :] * w[0]
                for j in range(1, m):
                    acc = acc + G[:, j, :] * w[j]
                return acc
            fs, gs = [], []
            for i in range(tab.stages):
                H = u
                if i > 0:
                    A = tab.A0[i][0] * fs[0]
                    B = tab.B0[i][0] * gs[0]
                    for j in range(1, i):
                        A = A + tab.A0[i][j] * fs[j]
                        B = B + tab.B0[i][j] * gs[j]
                    H = (u + dt * A) + times(B, chi2)
                tf_ = t + tab.c0[i] * dt
                tg_ = t + tab.c1[i] * dt
                fs.append(model.f(H, [lookup(tf_ - float(lag)) for lag in lags], p, tf_))
                gs.append(model.g(H, [lookup(tg_ - float(lag)) for lag in lags], p, tg_))
            at = tab.alpha[0] * fs[0]
            bt = tab.beta1[0] * gs[0]
            e2 = tab.beta2[0] * gs[0]
            for i in range(1, tab.stages):
                at = at + tab.alpha[i] * fs[i]
                bt = bt + tab.beta1[i] * gs[i]
                e2 = e2 + tab.beta2[i] * gs[i]
            unew = ((u + dt * at) + times(bt, dwv)) + times(e2, chi2)
        elif alg == ALG_EULER_HEUN:
            # [EXT] EulerHeunConstantCache: the second f/g evaluation happens at t + dt, so its delayed states are
            # looked up at (t + dt) - tau in the history that still ends at t_n
            ul2 = [lookup((t + dt) - float(lag)) for lag in lags]
            g1 = model.g(u, ul, p, t)
            if model.diagonal:
                noise = g1 * dw
            else:
                noise = g1[:, 0, :] * dw[0]
                for j in range(1, m):
                    noise = noise + g1[:, j, :] * dw[j]
            tmp = (u + fu * dt) + noise
            f2 = model.f(tmp, ul2, p, t + dt)
            g2 = model.g(tmp, ul2, p, t + dt)
            if model.diagonal:
                noise2 = (0.5 * (g1 + g2)) * dw
            else:
                noise2 = (0.5 * (g1[:, 0, :] + g2[:, 0, :])) * dw[0]
                for j in range(1, m):
                    noise2 = noise2 + (0.5 * (g1[:, j, :] + g2[:, j, :])) * dw[j]
            unew = (u + (dt * 0.5) * (fu + f2)) + noise2
        elif alg in (ALG_EM_SPLIT, ALG_EM_NOSPLIT):
            uc = K if alg == ALG_EM_SPLIT else u
            gv = model.g(uc, ul, p, t)
            if model.diagonal:
                unew = K + gv * dw
            else:
                noise = gv[:, 0, :] * dw[0]
                for j in range(1, m):
                    noise = noise + gv[:, j, :] * dw[j]
                unew = K + noise
        elif alg in WANGLI_ORDER:
            # [EXT] StochasticDiffEq WangLi3SMil_A..F (Wang & Li's three-stage Milstein methods), diagonal noise.
            # LOW CONFIDENCE: restated from memory of src/perform_step/low_order.jl -- three "#stage" blocks per method, each
            # applying one operator to the running u, all f / g at the step's t (hence one history lookup per lag and step):
            #   D: u <- u + dt f(u)
            #   C: k = g(u); k1 = (g(u + k sqdt) - k)/sqdt; u <- u - (dt/2) k1
            #   M: k = g(u); k1 = (g(u + k sqdt) - k)/sqdt; u <- u + k dW + (dW^2/2) k1
            # D.C.M composes to u + f dt + g dW + g'g (dW^2 - dt)/2 + O(dt^1.5): Milstein.  The six letters are taken to be the
            # six orders of (D, C, M) (WANGLI_ORDER); the letter <-> order assignment is an ASSUMPTION.  What is pinned here
            # without Julia: hand-derived one-step answers per order and strong order 1.0 on the reference's own convergence
            # problem (tests/test_oracle.py); what needs Julia: julia/replay_fixtures.jl.
            if not model.diagonal:
                raise ValueError("WangLi3SMil is implemented for diagonal noise only")
            unew = u
            for op in WANGLI_ORDER[alg]:
                if op == "D":
                    unew = unew + dt * model.f(unew, ul, p, t)
                else:
                    k = model.g(unew, ul, p, t)
                    k1 = (model.g(unew + k * sq, ul, p, t) - k) / sq
                    if op == "C":
                        unew = unew - (dt * 0.5) * k1
                    else:
                        unew = (unew + k * dw) + ((dw * dw) * 0.5) * k1
        elif alg in (ALG_RKMILC_ITO, ALG_RKMILC_STRAT) and not model.diagonal:
            # [EXT] StochasticDiffEq RKMilCommuteConstantCache, noise matrix G [ns, m] with commutative columns (MED
            # confidence, restated from memory; the Milstein algebra itself is checked against the exact solution of a
            # commutative linear SDE in tests/test_oracle.py):
            #   J = dW dW'/2 (get_iterated_I, commutative);  Ito: J -= |dt|/2 I
            #   for j: K_j = (Ito ? K : uprev) + sqdt G[:, j];  Dg_j = (G(K_j) - G(u)) / sqdt;  mil += Dg_j J[:, j]
            #   u+ = uprev + dt f + G dW + mil
            # The stage point follows RKMil's diagonal branch (K for Ito, uprev for Stratonovich).  LOW confidence that the
            # reference's matrix branch makes this distinction -- but it must: under Stratonovich J has mean dt/2, so the
            # O(sqrt(dt)) drift contamination of Dg_j through K would cost the scheme its strong order 1 (measured 0.69).
            Lg = model.g(u, ul, p, t)                      # [ns, m, nt]
            mil = np.zeros_like(u)
            for j in range(m):
                Kj = (K if alg == ALG_RKMILC_ITO else u) + sq * Lg[:, j, :]
                Dg = (model.g(Kj, ul, p, t) - Lg) / sq
                for k in range(m):
                    Jkj = 0.5 * (dw[k] * dw[j])
                    if alg == ALG_RKMILC_ITO and k == j:
                        Jkj = Jkj - 0.5 * abs(dt)
                    mil = mil + Dg[:, k, :] * Jkj
            noise = Lg[:, 0, :] * dw[0]
            for j in range(1, m):
                noise = noise + Lg[:, j, :] * dw[j]
            unew = (K + noise) + mil
        else:
            if not model.diagonal:
                raise ValueError("RKMil requires diagonal noise")
            Lg = model.g(u, ul, p, t)
            if alg in (ALG_RKMIL_ITO, ALG_RKMILC_ITO):
                ut = K + Lg * sq
                ggp = (model.g(ut, ul, p, t) - Lg) / sq
                mil = ggp * (dw * dw - abs(dt)) / 2.0
            else:
                ut = u + Lg * sq
                ggp = (model.g(ut, ul, p, t) - Lg) / sq
                mil = ggp * (dw * dw) / 2.0
            unew = (K + Lg * dw) + mil
        # savevalues!
        while ks < len(sched.save_t) and sched.save_step[ks] == n:
            if sched.save_exact[ks]:
                val = unew
            else:
                th = float(sched.save_theta[ks])
                val = (1.0 - th) * u + th * unew
            out[ks] = val
            sol_t.append(float(sched.save_t[ks])); sol_u.append(np.array(val, copy=True))
            ks += 1
        u = unew
        hist_t.append(float(sched.t[n + 1])); hist_u.append(u)
    status = (~np.all(np.isfinite(out), axis=(0, 1)) | ~np.all(np.isfinite(u), axis=0)).astype(np.uint8)
    if return_dense:
        return out, status, np.stack(hist_u)
    return out, status


def moments(out: np.ndarray) -> np.ndarray:
    """(count, mean, M2) per (save, state) over trajectories -- [n_save, ns, 3]
    ([EXT] SciMLBase.EnsembleAnalysis.timeseries_steps_meanvar; var = M2/(count-1))."""
    cnt = np.full(out.shape[:2], out.shape[2], dtype=np.float64)
    mean = out.mean(axis=2)
    m2 = ((out - mean[:, :, None]) ** 2).sum(axis=2)
    return np.stack([cnt, mean, m2], axis=2)


def merge_moments(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Chan et al. pairwise merge of (count, mean, M2) triples."""
    na, ma, sa = a[..., 0], a[..., 1], a[..., 2]
    nb, mb, sb = b[..., 0], b[..., 1], b[..., 2]
    n = na + nb
    with np.errstate(invalid="ignore", divide="ignore"):
        d = mb - ma
        mean = np.where(n > 0, ma + d * (nb / np.where(n > 0, n, 1)), 0.0)
        m2 = sa + sb + d * d * (na * nb / np.where(n > 0, n, 1))
    return np.stack([n, mean, m2], axis=-1)


def exact_mean_linear1(sched: Schedule, a, b, u0, h0, h1, lag):
    """K3: exact ensemble-mean recurrence of the C1/C5 model under EM (split or not):
    E[u_{n+1}] = (1 + a dt_n) E[u_n] + b dt_n E[u~(t_n - tau)]   (SURVEY.md section 8(c) K3)."""
    N = sched.n_steps
    m = np.empty(N + 1)
    m[0] = u0
    for n in range(N):
        mode = sched.lk_mode[n, 0]
        if mode == LK_PRE:
            ml = h0 + h1 * sched.lk_theta[n, 0]
        elif mode == LK_EXACT:
            ml = m[sched.lk_idx[n, 0]]
        else:
            i = sched.lk_idx[n, 0]; th = sched.lk_theta[n, 0]
            ml = (1.0 - th) * m[i - 1] + th * m[i]
        m[n + 1] = m[n] + sched.dt[n] * (a * m[n] + b * ml)
    return m


# ---------------------------------------------------------------------------------------------------
# Order statistics across trajectories: what [EXT] SciMLBase EnsembleSummary computes on the reference side with
# EnsembleAnalysis.timeseries_steps_quantile / timeseries_steps_median -> [EXT] Statistics.quantile (alpha = beta = 1)
# and Statistics.median.  Checker for sdde_b200_quantiles; a full sort per row, formulas restated from Statistics.jl.
# ---------------------------------------------------------------------------------------------------
def _julia_sort(v):
    """sort(v) under Julia's isless for floats: a total order with -0.0 before +0.0 (numpy's sort leaves the order of the
    two zeros unspecified); NaNs are handled by the callers."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    b = v.view(np.uint64)
    key = np.where(b >> np.uint64(63) != 0, ~b, b | np.uint64(1 << 63))
    return v[np.argsort(key, kind="stable")]


def julia_quantile(v, p):
    """Statistics.quantile(v, p) for a float64 vector (sorted internally)."""
    v = _julia_sort(v)
    n = v.size
    if not 0.0 <= p <= 1.0:
        raise ValueError("input probability out of [0,1] range")
    if np.isnan(v).any():
        return np.nan                              # Julia raises; the device path reports NaN for such a row
    if n == 1:
        a = b = v[0]
        gamma = 0.0
    else:
        m = np.float64(1.0) + np.float64(p) * (1.0 - 1.0 - 1.0)
        aleph = np.float64(n) * np.float64(p) + m
        j = int(min(max(int(np.trunc(aleph)), 1), n - 1))
        gamma = float(min(max(aleph - j, 0.0), 1.0))
        a, b = v[j - 1], v[j]
    with np.errstate(invalid="ignore"):
        if np.isfinite(a) and np.isfinite(b):
            return a + np.float64(gamma) * (b - a)
        return (1.0 - np.float64(gamma)) * a + np.float64(gamma) * b


def julia_median(v):
    """Statistics.median(v): middle order statistic, or middle(a, b) = a/2 + b/2 of the two central ones."""
    v = _julia_sort(v)
    n = v.size
    if np.isnan(v).any():
        return np.nan
    if n % 2 == 1:
        return v[(n - 1) // 2]
    with np.errstate(invalid="ignore"):
        return v[n // 2 - 1] / 2 + v[n // 2] / 2


def ensemble_quantiles(paths, probs, want_median=False):
    """paths [n_rows, n_traj] -> [n_rows, len(probs) + want_median]."""
    paths = np.asarray(paths, dtype=np.float64)
    cols = len(probs) + (1 if want_median else 0)
    res = np.empty((paths.shape[0], cols))
    for i, row in enumerate(paths):
        for j, p in enumerate(probs):
            res[i, j] = julia_quantile(row, p)
        if want_median:
            res[i, -1] = julia_median(row)
    return res


# ---------------------------------------------------------------------------------------------------
# Supplied NoiseGrid(t, W) on its own time grid: [EXT] DiffEqNoiseProcess interpolate!(::NoiseGrid, u, p, t)
# (searchsortedfirst at t - 10 eps, exact-node shortcuts, linear interpolant) and calculate_step!'s snap of the last
# node (|t - ts[end]| < 100 eps), evaluated at the schedule's nodes.  Checker for the library's noise_regrid_kernel.
# ---------------------------------------------------------------------------------------------------
def noise_grid_on_nodes(ts, W, t_nodes):
    """W[n_grid, c, traj] on grid ts -> W at t_nodes, [len(t_nodes), c, traj]."""
    ts = np.asarray(ts, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    eps = np.finfo(np.float64).eps
    out = np.empty((len(t_nodes),) + W.shape[1:], dtype=np.float64)
    for k, t in enumerate(np.asarray(t_nodes, dtype=np.float64)):
        if abs(t - ts[-1]) < 100.0 * eps:
            t = ts[-1]
        if t > ts[-1] or t < ts[0]:
            raise ValueError("Solution interpolation cannot extrapolate past the final timepoint")
        i = int(np.searchsorted(ts, t - 10.0 * eps, side="left"))
        if i < len(ts) and ts[i] == t:
            out[k] = W[i]
        elif i > 0 and ts[i - 1] == t:
            out[k] = W[i - 1]
        else:
            th = (t - ts[i - 1]) / (ts[i] - ts[i - 1])
            out[k] = (1.0 - th) * W[i - 1] + th * W[i]
    return out


def noise_grid_increments(W_nodes):
    """calculate_step!/accept_step! bookkeeping: dW_n = W(t_{n+1}) - curW, curW += dW_n (curW starts at W(t0))."""
    cur = W_nodes[0].copy()
    dW = np.empty((W_nodes.shape[0] - 1,) + W_nodes.shape[1:], dtype=np.float64)
    for n in range(dW.shape[0]):
        dW[n] = W_nodes[n + 1] - cur
        cur = cur + dW[n]
    return dW
