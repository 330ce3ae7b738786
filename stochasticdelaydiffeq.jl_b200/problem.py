"""Host-side mirror of the SciML surface for this path: ``SDDEProblem``, ``EnsembleProblem``, ``EM``, ``RKMil``.

Names, argument order and meaning follow the reference's user API ([EXT] SciMLBase ``SDDEProblem(f, g, u0, h, tspan, p;
constant_lags, noise_rate_prototype, order_discontinuity_t0, seed)`` as consumed at /root/reference/src/solve.jl:128,
and ``EnsembleProblem(prob; prob_func)``), so that the parity tests read like the reference's own tests
(/root/reference/test/*.jl).  Julia closures cannot cross the C ABI; ``f`` and ``g`` are therefore handles into the
device model zoo (or a user model compiled from CUDA source, see ``UserModel``), and ``h`` is an affine history
``h(p, t) = c0 + c1*t`` (a Python callable ``h(p, t)`` is accepted and checked to be affine).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _capi as capi


# ---------------------------------------------------------------------------------------------
# algorithms ([EXT] StochasticDiffEq EM{split}, RKMil{interpretation})
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class EM:
    """Euler-Maruyama.  ``EM()`` is split-step like the reference's default ``EM(true)``."""
    split: bool = True

    @property
    def alg_id(self):
        return capi.ALG_EM_SPLIT if self.split else capi.ALG_EM_NOSPLIT


@dataclass(frozen=True)
class RKMil:
    """Derivative-free Milstein; fixed step only on this path (pass ``adaptive=False`` to ``solve``)."""
    interpretation: str = "Ito"

    def __post_init__(self):
        if self.interpretation not in ("Ito", "Stratonovich"):
            raise ValueError("interpretation must be 'Ito' or 'Stratonovich'")

    @property
    def alg_id(self):
        return capi.ALG_RKMIL_ITO if self.interpretation == "Ito" else capi.ALG_RKMIL_STRAT


@dataclass(frozen=True)
class RKMilCommute(RKMil):
    """Milstein for commutative noise ([EXT] StochasticDiffEq RKMilCommute; reference smoke test test/test_prob_sol.jl:44-49).
    On diagonal noise its update ``K + L dW + Dg (dW^2/2 [- dt/2])`` is the RKMil update written with the iterated
    integral J and the library runs the RKMil kernel; on a noise MATRIX with commutative columns (e.g. ``linsys`` with
    G_ij = M_ij u_i) it evaluates G once more per noise column: ``u+ = K + G dW + sum_j Dg_j J[:, j]``."""

    @property
    def alg_id(self):
        return capi.ALG_RKMILC_ITO if self.interpretation == "Ito" else capi.ALG_RKMILC_STRAT


@dataclass(frozen=True)
class WangLi3SMil:
    """``WangLi3SMil_A()`` .. ``WangLi3SMil_F()`` ([EXT] StochasticDiffEq; reference tests test/test_prob_sol.jl:50-61 and
    test/analyticless_convergence_tests.jl:62-91): Wang & Li's three-stage Milstein methods, diagonal noise, strong order 1.
    Restated as the six orders of (drift, Ito correction, Stratonovich-Milstein diffusion) -- LOW confidence in the
    letter <-> order assignment, see DESIGN.md section 6."""
    variant: str = "A"

    def __post_init__(self):
        if self.variant not in "ABCDEF" or len(self.variant) != 1:
            raise ValueError("variant must be one of 'A' .. 'F'")

    @property
    def alg_id(self):
        return capi.ALG_WANGLI3SMIL_A + "ABCDEF".index(self.variant)


def WangLi3SMil_A():
    return WangLi3SMil("A")


def WangLi3SMil_B():
    return WangLi3SMil("B")


def WangLi3SMil_C():
    return WangLi3SMil("C")


def WangLi3SMil_D():
    return WangLi3SMil("D")


def WangLi3SMil_E():
    return WangLi3SMil("E")


def WangLi3SMil_F():
    return WangLi3SMil("F")


@dataclass(frozen=True)
class EulerHeun:
    """Euler-Heun (Stratonovich), fixed step: second drift/diffusion evaluation at t + dt, i.e. a second set of
    delayed-state lookups per step ([EXT] StochasticDiffEq EulerHeun; reference smoke test test/test_prob_sol.jl:34)."""

    @property
    def alg_id(self):
        return capi.ALG_EULER_HEUN


@dataclass(frozen=True)
class SRA:
    """Roessler's stochastic Runge-Kutta methods for ADDITIVE noise, fixed step ([EXT] StochasticDiffEq ``SRA(tableau=...)``,
    ``SRA1()``, ``SRA3()``, ``SOSRA()``; the reference's smoke tests test/test_prob_sol.jl:74-87 run them adaptively -- pass
    ``adaptive=False`` here).  Strong order 1.5 when g does not depend on u.  Every step draws a second increment dZ
    (src/solve.jl:308-312), so supplied noise has ``2 * n_chan`` rows per step: dW first, then dZ (``NoiseGrid(t, W, Z)``).
    Every f / g evaluation of a stage looks its delayed states up at the stage's own time ``t + c dt``.

    ``tableau`` = dict with ``c0, c1, A0, B0, alpha, beta1, beta2`` (``A0[i][j]``: weight of stage j in stage i; at most
    3 stages and 4 distinct stage times); ``name`` selects a built-in one instead."""
    name: str = "SRA1"
    tableau: object = None
    order: float = 2.0          # [EXT] alg_order of the SRA family (depth of the discontinuity propagation)

    def __post_init__(self):
        if self.tableau is None and self.name not in ("SRA1", "SRA3", "SOSRA"):
            raise ValueError("built-in SRA tableaus: 'SRA1', 'SRA3', 'SOSRA' (pass tableau=... for another one)")

    @property
    def alg_id(self):
        if self.tableau is not None:
            return capi.ALG_SRA_TABLEAU
        return {"SRA1": capi.ALG_SRA1, "SRA3": capi.ALG_SRA3, "SOSRA": capi.ALG_SOSRA}[self.name]

    noise_mult = 2

    def flat_tableau(self):
        """-> (float64 array in the layout of sdde_desc.sra_tableau, stages) or (None, 0) for a built-in tableau."""
        if self.tableau is None:
            return None, 0
        tb = self.tableau
        s = len(tb["alpha"])
        A0, B0 = np.asarray(tb["A0"], dtype=np.float64), np.asarray(tb["B0"], dtype=np.float64)
        if A0.shape != (s, s) or B0.shape != (s, s) or any(len(tb[k]) != s for k in ("c0", "c1", "beta1", "beta2")):
            raise ValueError("SRA tableau: c0, c1, alpha, beta1, beta2 need s entries, A0 and B0 s x s")
        flat = np.concatenate([np.asarray(tb["c0"], dtype=np.float64), np.asarray(tb["c1"], dtype=np.float64), A0.ravel(), B0.ravel(),
                               np.asarray(tb["alpha"], dtype=np.float64), np.asarray(tb["beta1"], dtype=np.float64),
                               np.asarray(tb["beta2"], dtype=np.float64)])
        return np.ascontiguousarray(flat), s


def SRA1():
    return SRA("SRA1")


def SRA3():
    return SRA("SRA3")


def SOSRA():
    return SRA("SOSRA")


# ---------------------------------------------------------------------------------------------
# model zoo handles: model.f / model.g are what the user passes as f and g
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ModelFunction:
    model: "Model"
    role: str   # 'f' | 'g'


@dataclass(frozen=True)
class Model:
    name: str
    model_id: int
    n_state: int
    n_lag: int
    n_noise: int
    n_param: int
    diagonal: bool
    doc: str = ""

    @property
    def f(self):
        return ModelFunction(self, "f")

    @property
    def g(self):
        return ModelFunction(self, "g")

    @property
    def n_chan(self):
        return self.n_state if self.diagonal else self.n_noise


linear1 = Model("linear1", capi.MODEL_LINEAR1, 1, 1, 1, 3, True,
                "du = (a u + b u(t-tau)) dt + s u dW,  p = (a, b, s)")
hayes = Model("hayes", capi.MODEL_HAYES, 1, 1, 1, 7, True,
              "reference tests' Hayes equation, p = (tau, a, b, c, alpha, beta, gamma)")
mackey_glass = Model("mackey_glass", capi.MODEL_MACKEY_GLASS, 1, 1, 1, 4, True,
                     "du = (beta x/(1+x^n) - gamma u) dt + s u dW, x = u(t-tau), p = (beta, gamma, n, s)")
sir = Model("sir", capi.MODEL_SIR, 3, 1, 2, 3, False,
            "delayed SIR with 3x2 sparse noise (reference test nondiagonal_sparse_noise.jl), p = (beta, c, tau)")
scalar_add = Model("scalar_add", capi.MODEL_SCALAR_ADD, 1, 1, 1, 2, True,
                   "du = a u(t-tau) dt + b dW (reference test events.jl), p = (a, b)")


def linsys(n_state: int, n_lag: int, n_noise: int) -> Model:
    """du = (A u + sum_l B_l u(t-tau_l) + c) dt + G(u) dW, G_ij = M_ij u_i + N_ij;
    p = concat(A, B_1..B_L, c, M, N) row-major (use ``linsys_params``)."""
    npar = n_state * n_state * (1 + n_lag) + n_state + 2 * n_state * n_noise
    return Model(f"linsys_{n_state}_{n_lag}_{n_noise}", capi.MODEL_LINSYS, n_state, n_lag, n_noise, npar, False)


def linsys_params(A, Bs, c, M, N) -> np.ndarray:
    parts = [np.asarray(A, float).ravel()] + [np.asarray(B, float).ravel() for B in Bs]
    parts += [np.asarray(c, float).ravel(), np.asarray(M, float).ravel(), np.asarray(N, float).ravel()]
    return np.concatenate(parts)


class UserModel:
    """A model compiled at run time from CUDA C++ source (NVRTC, sm_100a) -- SURVEY.md section 8(f) rank 1.

    ``source`` must define ``struct UserModel`` with the interface documented in include/sdde_b200.h.
    """

    def __init__(self, source: str, n_state: int, n_lag: int, n_noise: int, n_param: int, diagonal: bool, name="user",
                 history: bool = False):
        self.source = source
        self.history = bool(history)   # the source defines HAS_H = true and h(out, p, t): pass ``model.h`` as the problem's h
        self._shape = (n_state, n_lag, n_noise, n_param, diagonal)
        self.name = name
        self._ids: dict[int, int] = {}

    def resolve(self, ctx: "capi.Context") -> Model:
        import ctypes as C
        key = id(ctx)
        if key not in self._ids:
            mid = C.c_int32(0)
            capi.check(capi.lib().sdde_b200_compile_model(ctx.handle, self.source.encode(), C.byref(mid)))
            self._ids[key] = mid.value
        ns, nl, nw, npar, diag = self._shape
        return Model(self.name, self._ids[key], ns, nl, nw, npar, diag)

    @property
    def f(self):
        return ModelFunction(self, "f")

    @property
    def g(self):
        return ModelFunction(self, "g")

    @property
    def h(self):
        """Handle of the model's own device history function (any h(p, t), not only affine)."""
        if not self.history:
            raise AttributeError("this UserModel was created without history=True")
        return ModelFunction(self, "h")


# ---------------------------------------------------------------------------------------------
# history functions
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class AffineHistory:
    """h(p, t)[j] = c0[j] + c1[j]*t for t < t0 (constant history: c1 = 0)."""
    c0: tuple
    c1: tuple

    def __call__(self, p, t):
        return np.asarray(self.c0) + np.asarray(self.c1) * t


def ConstantHistory(values) -> AffineHistory:
    v = tuple(float(x) for x in np.atleast_1d(values))
    return AffineHistory(v, tuple(0.0 for _ in v))


def _as_history(h, p, t0, n_state) -> AffineHistory:
    if isinstance(h, ModelFunction):                 # device history of a UserModel: the descriptor's affine slots are unused
        return AffineHistory(tuple(0.0 for _ in range(n_state)), tuple(0.0 for _ in range(n_state)))
    if isinstance(h, AffineHistory):
        if len(h.c0) != n_state:
            raise ValueError("history has the wrong number of states")
        return h
    if callable(h):
        ts = [t0 - 1.0, t0 - 2.0, t0 - 3.5]
        vals = [np.atleast_1d(np.asarray(h(p, t), dtype=float)) for t in ts]
        c1 = (vals[0] - vals[1]) / (ts[0] - ts[1])
        c0 = vals[0] - c1 * ts[0]
        if not np.allclose(c0 + c1 * ts[2], vals[2], rtol=1e-13, atol=1e-13):
            raise NotImplementedError("only affine history functions h(p,t) = c0 + c1*t cross the C ABI on this path")
        if c0.size != n_state:
            raise ValueError("history has the wrong number of states")
        return AffineHistory(tuple(c0), tuple(c1))
    return ConstantHistory(h)


# ---------------------------------------------------------------------------------------------
# SDDEProblem / EnsembleProblem
# ---------------------------------------------------------------------------------------------
class SDDEProblem:
    """``SDDEProblem(f, g, u0, h, tspan, p; constant_lags=(), noise_rate_prototype=None, order_discontinuity_t0=0,
    seed=0, neutral=False)`` -- fields as unpacked at /root/reference/src/solve.jl:128.

    ``lags`` supplies the delays the model looks up when they are not declared as ``constant_lags`` (the reference's
    nondiagonal_sparse_noise test does this: the lag lives in ``p`` and no discontinuity tracking happens)."""

    def __init__(self, f, g, u0, h, tspan, p=(), *, constant_lags=(), lags=None, dependent_lags=(),
                 noise_rate_prototype=None, noise=None, order_discontinuity_t0=0, neutral=False, seed=0):
        if not (isinstance(f, ModelFunction) and isinstance(g, ModelFunction) and f.role == "f" and g.role == "g"):
            raise TypeError("f and g must be <model>.f and <model>.g handles of the device model zoo (or of a UserModel)")
        if f.model is not g.model:
            raise ValueError("f and g must come from the same model")
        if isinstance(h, ModelFunction):
            if h.role != "h" or h.model is not f.model:
                raise ValueError("a device history must be <the same UserModel>.h")
        elif getattr(f.model, "history", False):
            raise ValueError("this UserModel defines its own history function: pass <model>.h as h")
        if dependent_lags:
            raise NotImplementedError("state-dependent lags are outside this path (SURVEY.md section 8(f))")
        if neutral:
            raise NotImplementedError("neutral problems are outside this path")
        self.model = f.model
        self.f, self.g = f, g
        self.u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64)).copy()
        self.scalar_state = np.isscalar(u0)
        self.tspan = (float(tspan[0]), float(tspan[1]))
        self.p = np.asarray(p, dtype=np.float64).ravel().copy()
        self.h = h
        self.constant_lags = tuple(float(x) for x in constant_lags)
        self.lags = tuple(float(x) for x in (lags if lags is not None else constant_lags))
        self.noise_rate_prototype = noise_rate_prototype
        self.noise = noise
        self.order_discontinuity_t0 = int(order_discontinuity_t0)
        self.neutral = False
        self.seed = int(seed)

    def remake(self, *, u0=None, p=None):
        q = SDDEProblem.__new__(SDDEProblem)
        q.__dict__.update(self.__dict__)
        if u0 is not None:
            q.u0 = np.atleast_1d(np.asarray(u0, dtype=np.float64)).copy()
        if p is not None:
            q.p = np.asarray(p, dtype=np.float64).ravel().copy()
        return q


# ---------------------------------------------------------------------------------------------
# EnsembleProblem.output_func / reduction ([EXT] SciMLBase): arbitrary host closures in SciML.  The forms below are the
# ones this path can express on the device (same names and meaning as in julia/EnsembleB200.jl); anything else raises.
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FinalState:
    """output_func ``(sol, i) -> (sol.u[end], false)``: only the state at tf is saved (saveat = [tf], save_start = False)."""


@dataclass(frozen=True)
class SaveIdxs:
    """output_func keeping some state components (0-based) == the ``save_idxs`` keyword (src/solve.jl:355-356)."""
    idxs: tuple

    def __post_init__(self):
        object.__setattr__(self, "idxs", tuple(int(i) for i in np.atleast_1d(self.idxs)))


@dataclass(frozen=True)
class MomentsReduction:
    """reduction accumulating (count, mean, M2) per saved time point and state -- what EnsembleAnalysis.timeseries_steps_meanvar
    computes from a finished ensemble.  On this path: moments on the device, no paths leave the GPU (24 bytes per saved value)."""


class EnsembleProblem:
    """``EnsembleProblem(prob; prob_func=None, output_func=None, reduction=None)`` ([EXT] SciMLBase).

    On this path a trajectory may differ from ``prob`` in ``p`` and ``u0`` only (what the kernels take per
    trajectory).  Either give ``prob_func(prob, i, repeat) -> SDDEProblem`` (called on the host for every i, as the
    reference does) or, for large ensembles, the vectorised ``p_batch[n_param, trajectories]`` /
    ``u0_batch[n_state, trajectories]``."""

    def __init__(self, prob: SDDEProblem, *, prob_func=None, output_func=None, reduction=None, p_batch=None,
                 u0_batch=None, safetycopy=True):
        if output_func is not None and not isinstance(output_func, (FinalState, SaveIdxs)):
            raise NotImplementedError("output_func is a host closure in SciML and cannot run on the device; expressible on this "
                                      "path: FinalState(), SaveIdxs(idxs), or the saveat / save_idxs keywords")
        if reduction is not None and not isinstance(reduction, MomentsReduction):
            raise NotImplementedError("reduction is a host closure in SciML and cannot run on the device; expressible on this "
                                      "path: MomentsReduction() (mean / variance per time point); order statistics: EnsembleSummary")
        self.output_func, self.reduction = output_func, reduction
        self.prob = prob
        self.prob_func = prob_func
        self.p_batch = None if p_batch is None else np.ascontiguousarray(p_batch, dtype=np.float64)
        self.u0_batch = None if u0_batch is None else np.ascontiguousarray(u0_batch, dtype=np.float64)

    def batches(self, trajectories: int, first: int = 0):
        """-> (p_batch or None, u0_batch or None) for trajectories [first, first+trajectories)."""
        pb, ub = self.p_batch, self.u0_batch
        if pb is not None:
            pb = np.ascontiguousarray(pb[:, first:first + trajectories])
        if ub is not None:
            ub = np.ascontiguousarray(ub[:, first:first + trajectories])
        if self.prob_func is not None:
            ps, us = [], []
            for i in range(first, first + trajectories):
                q = self.prob_func(self.prob, i + 1, 1)        # Julia's ensemble index is 1-based
                if q.model is not self.prob.model or q.tspan != self.prob.tspan or q.lags != self.prob.lags:
                    raise NotImplementedError("prob_func may change p and u0 only on this path")
                ps.append(q.p); us.append(q.u0)
            pb = np.ascontiguousarray(np.stack(ps, axis=1))
            ub = np.ascontiguousarray(np.stack(us, axis=1))
        return pb, ub
