"""Regenerates tests/golden/golden_v1.npz from the numpy oracle (oracle/sdde_oracle.py).

The reference (Julia) cannot run in this environment and holds no golden vectors for this path,
so these fixtures are ORACLE-generated regression vectors (they pin the restatement against accidental change and let
the GPU tests run without recomputing the oracle); they are NOT outputs of StochasticDelayDiffEq.jl.
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import sdde_oracle as O  # noqa: E402


def main():
    g = {}
    n = 16
    # C2 (Mackey-Glass, EM split) and C3 (tau=17.02, RKMil Ito, order0 = 0 and 2), T=40, 16 paths, shared increments
    for name, tau, alg, order0 in [("c2_em", 17.0, 0, 0), ("c3_rkmil_o0", 17.02, 2, 0), ("c3_rkmil_o2", 17.02, 2, 2)]:
        sv = O.expand_saveat(1.0, 0.0, 40.0)[0]
        sch = O.build_schedule(0.0, 40.0, 0.05, [tau], alg, order_discontinuity_t0=order0, saveat_times=sv)
        rng = np.random.Generator(np.random.Philox(key=20261019))
        xi = rng.standard_normal((sch.n_steps, 1, n))
        dW = xi * sch.sqdt[:, None, None]
        out, st = O.solve_ensemble(O.get_model(O.MODEL_MACKEY_GLASS), [0.2, 0.1, 10.0, 0.1], [1.2], [1.2], [0.0], [tau],
                                   sch, alg, dW)
        g[name + "_xi"] = xi
        g[name + "_out"] = out
        g[name + "_t"] = sch.t
        g[name + "_dt"] = sch.dt
    # C4 (4 states, 2 lags, 4x2 noise), T=6
    import cases
    c = cases.c4(T=6.0)
    sch = c.oracle_schedule()
    rng = np.random.Generator(np.random.Philox(key=4))
    xi = rng.standard_normal((sch.n_steps, 2, n))
    out, st = c.oracle(xi * sch.sqdt[:, None, None])
    g["c4_xi"], g["c4_out"], g["c4_t"] = xi, out, sch.t
    # Philox normals: seed 0xB2000001, trajectories 0..7 and 2^32+5, 6 steps x 3 channels
    ids = np.array(list(range(8)) + [2 ** 32 + 5], dtype=np.uint64)
    g["philox_ids"] = ids
    g["philox_z"] = O.philox_normals(ids, 6, 3, 0xB2000001)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **g)
    print("wrote golden_v1.npz:", {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()
