"""Regenerates tests/golden/golden_v2.npz: ORACLE-generated regression vectors (see make_golden.py for what that does and
does not mean) for the round-2 algorithms -- EM(false), RKMil Stratonovich, EulerHeun, RKMilCommute on a noise matrix,
WangLi3SMil_A / _F, SRA1 / SRA3 / SOSRA -- each on 8 paths with shared normals, plus the schedule nodes.
Usage: python tests/golden/make_golden_v2.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import sdde_oracle as O  # noqa: E402

N = 8
PMUL = [1.0, -4.0, -2.0, 10.0, -1.3, -1.2, 1.1]      # test/test_prob_sol.jl:18
PADD = [0.33, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]       # additive, delayed term in g, lag off the grid


def fixtures():
    """name -> (oracle model, params, u0, h0, h1, lags, tspan, dt, alg, saveat, noise rows per step, key)"""
    mg = O.get_model(O.MODEL_MACKEY_GLASS)
    hay = O.get_model(O.MODEL_HAYES)
    pm = [0.2, 0.1, 10.0, 0.1]
    fx = {
        "mg_em_nosplit": (mg, pm, [1.2], [1.2], [0.0], [17.0], (0.0, 30.0), 0.05, O.ALG_EM_NOSPLIT, 1.0, 1, 11),
        "mg_rkmil_strat": (mg, pm, [1.2], [1.2], [0.0], [17.02], (0.0, 30.0), 0.05, O.ALG_RKMIL_STRAT, 1.0, 1, 12),
        "mg_euler_heun": (mg, pm, [1.2], [1.2], [0.0], [17.0], (0.0, 30.0), 0.05, O.ALG_EULER_HEUN, 1.0, 1, 13),
        "hayes_wangli_a": (hay, PMUL, [1.0], [1.0], [1.0], [1.0], (0.0, 3.0), 2.0 ** -6, O.ALG_WANGLI_A, 0.25, 1, 14),
        "hayes_wangli_f": (hay, PMUL, [1.0], [1.0], [1.0], [1.0], (0.0, 3.0), 2.0 ** -6, O.ALG_WANGLI_F, 0.25, 1, 15),
        "hayes_sra1": (hay, PADD, [1.0], [1.0], [1.0], [0.33], (0.0, 3.0), 0.05, O.ALG_SRA1, 0.25, 2, 16),
        "hayes_sra3": (hay, PADD, [1.0], [1.0], [1.0], [0.33], (0.0, 3.0), 0.05, O.ALG_SRA3, 0.25, 2, 17),
        "hayes_sosra": (hay, PADD, [1.0], [1.0], [1.0], [0.33], (0.0, 3.0), 0.05, O.ALG_SOSRA, 0.25, 2, 18),
    }
    import cases
    c = cases.c4(T=4.0)
    fx["c4_rkmil_commute"] = (c.omodel, list(np.asarray(c.p, float)), c.u0, c.h0, c.h1, c.lags, (0.0, 4.0), 0.01, O.ALG_RKMILC_ITO, 0.5, 2, 19)
    return fx


def run(fx):
    model, p, u0, h0, h1, lags, tspan, dt, alg, saveat, rows, key = fx
    sv, ss, se = O.expand_saveat(saveat, *tspan)
    sch = O.build_schedule(tspan[0], tspan[1], dt, lags, alg, saveat_times=sv, save_start=ss, save_end=se)
    xi = np.random.Generator(np.random.Philox(key=key)).standard_normal((sch.n_steps, rows, N))
    out, st = O.solve_ensemble(model, p, u0, h0, h1, lags, sch, alg, xi * sch.sqdt[:, None, None])
    assert np.all(st == 0)
    return sch, xi, out


def main():
    g = {}
    for name, fx in fixtures().items():
        sch, xi, out = run(fx)
        g[name + "_xi"], g[name + "_out"], g[name + "_t"], g[name + "_save_t"] = xi, out, sch.t, sch.save_t
    np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **g)
    print("wrote golden_v2.npz:", {k: v.shape for k, v in g.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
