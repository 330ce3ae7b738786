"""BASELINE.json configs C1..C5 and the reference-test models as PRODUCT-side problems only
(no oracle, no test infrastructure): what the measurement scripts of this directory solve."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdde_b200 as S  # noqa: E402

ALG_MAP = {0: S.EM(True), 1: S.EM(False), 2: S.RKMil("Ito"), 3: S.RKMil("Stratonovich"), 4: S.EulerHeun(),
           5: S.RKMilCommute("Ito"), 6: S.RKMilCommute("Stratonovich"), **{7 + i: S.WangLi3SMil("ABCDEF"[i]) for i in range(6)},
           13: S.SRA1(), 14: S.SRA3(), 15: S.SOSRA()}


class Case:
    def __init__(self, name, model, u0, h0, h1, tspan, p, lags, dt, saveat, alg=0, order0=0, declared=True):
        self.name, self.model = name, model
        self.u0, self.h0, self.h1 = list(u0), list(h0), list(h1)
        self.tspan, self.p, self.lags, self.dt, self.saveat = tspan, np.asarray(p, float), list(lags), dt, saveat
        self.alg, self.order0, self.declared = alg, order0, declared

    def problem(self, **kw):
        h = S.AffineHistory(tuple(self.h0), tuple(self.h1))
        u0 = self.u0[0] if len(self.u0) == 1 else self.u0
        return S.SDDEProblem(self.model.f, self.model.g, u0, h, self.tspan, self.p,
                             constant_lags=tuple(self.lags) if self.declared else (), lags=tuple(self.lags),
                             order_discontinuity_t0=self.order0, **kw)

    def solve_kwargs(self):
        kw = dict(dt=self.dt, saveat=self.saveat)
        if self.alg in (2, 3, 5, 6, 13, 14, 15):
            kw["adaptive"] = False
        return kw

    def algorithm(self):
        return ALG_MAP[self.alg]

    def with_(self, **kw):
        c = Case.__new__(Case)
        c.__dict__.update(self.__dict__)
        c.__dict__.update(kw)
        return c


def c1(T=10.0, saveat=0.1):
    return Case("C1", S.linear1, [1.0], [1.0], [0.0], (0.0, T), (-1.0, 0.5, 0.2), [1.0], 0.01, saveat)


def c2(T=100.0, tau=17.0, alg=0, order0=0):
    return Case("C2", S.mackey_glass, [1.2], [1.2], [0.0], (0.0, T), (0.2, 0.1, 10.0, 0.1), [tau], 0.05, 1.0, alg=alg, order0=order0)


def c3(order0=0, T=100.0, alg=2):
    c = c2(T=T, tau=17.02, alg=alg, order0=order0)
    c.name = f"C3_order{order0}"
    return c


def c4_params():
    A = -1.0 * np.eye(4) + 0.1 * (np.ones((4, 4)) - np.eye(4))
    B1 = 0.3 * np.eye(4)
    B2 = -0.2 * np.roll(np.eye(4), 1, axis=1)
    M = np.zeros((4, 2))
    for (i, j) in [(0, 0), (1, 0), (1, 1), (2, 1), (3, 1)]:
        M[i, j] = 0.1
    return S.linsys_params(A, [B1, B2], np.zeros(4), M, np.zeros((4, 2)))


def c4(T=20.0):
    u0 = [1.0, 0.5, 0.25, 0.125]
    return Case("C4", S.linsys(4, 2, 2), u0, u0, [0.0] * 4, (0.0, T), c4_params(), [1.0, 2.5], 0.01, 0.5)


def hayes(T=10.0, dt=2.0 ** -6, alg=0, beta=0.0, saveat=None):
    p = [1.0, -4.0, -2.0, 10.0, -1.3, beta, 1.1]
    return Case("hayes", S.hayes, [1.0], [1.0], [1.0], (0.0, T), p, [1.0], dt, saveat, alg=alg)


def sir(T=40.0, dt=0.1):
    return Case("sir", S.sir, [990.0, 10.0, 0.0], [1000.0, 0.0, 0.0], [0.0] * 3, (0.0, T), [0.05, 10.0, 4.0], [4.0], dt, 1.0,
                declared=False)


def scalar_add(T=10.0):
    return Case("scalar_add", S.scalar_add, [0.0], [0.0], [0.0], (0.0, T), [0.3, 0.1], [0.5], 0.01, 0.5)


def increments(case, n_traj, key=20261019):
    """Supplied Brownian increments on the product's own schedule: xi ~ N(0,1) (numpy Philox), dW_n = sqrt(dt_n) xi_n."""
    sch = S.schedule(case.problem(), case.algorithm(), **case.solve_kwargs())
    rng = np.random.Generator(np.random.Philox(key=key))
    return rng.standard_normal((sch.n_steps, case.model.n_chan, n_traj)) * sch.sqdt[:, None, None]
