"""Throughput of every algorithm kernel on one problem shape: the reference's Hayes equation (test/test_prob_sol.jl), lag 1,
dt = 0.01, t in [0, 10] (1000+ steps), 10^6 paths, in-kernel Philox, saveat = 0.1, device-resident output."""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__)))]
import numpy as np
import sdde_b200 as S
PMUL = [1.0, -4.0, -2.0, 10.0, -1.3, -1.2, 1.1]
PADD = [1.0, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
algs = [("EM", S.EM(), PMUL, {}), ("EM(false)", S.EM(False), PMUL, {}), ("RKMil Ito", S.RKMil(), PMUL, {"adaptive": False}),
        ("RKMil Strat", S.RKMil("Stratonovich"), PMUL, {"adaptive": False}), ("EulerHeun", S.EulerHeun(), PMUL, {}),
        ("WangLi3SMil_A", S.WangLi3SMil_A(), PMUL, {}), ("EM additive", S.EM(), PADD, {}),
        ("SRA1", S.SRA1(), PADD, {"adaptive": False}), ("SRA3", S.SRA3(), PADD, {"adaptive": False}), ("SOSRA", S.SOSRA(), PADD, {"adaptive": False})]
for name, alg, p, kw in algs:
    prob = S.SDDEProblem(S.hayes.f, S.hayes.g, [1.0], lambda pp, t: np.ones(1) + t, (0.0, 10.0), p, constant_lags=(1.0,))
    best = 1e9
    for r in range(4):
        sim = S.solve(S.EnsembleProblem(prob), alg, S.EnsembleB200(device=0, output="device"), dt=0.01, saveat=0.1, trajectories=n, seed=3, **kw)
        if r: best = min(best, sim.timing["kernel_ms"])
    ns = sim.timing["n_steps"]
    print(f"{name:14s} steps {ns:5d} kernel {best:8.3f} ms  {n * ns / best / 1e-3:.3e} traj-steps/s  regs {sim.timing['regs_per_thread']:3d} {sim.timing['kernel_name']}", flush=True)
