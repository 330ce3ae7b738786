"""Latency of small ensembles (BASELINE.json configs[0]: 10^4 paths, linear SDDE, dt=0.01, T=10)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases
case = cases.c1(T=10.0, saveat=0.1)
for n in (1, 100, 10_000, 100_000, 1_000_000):
    for r in range(3):
        t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=3, **case.solve_kwargs())
        w = (time.perf_counter() - t0) * 1e3
    print(f"n={n:8d} kernel {sim.timing['kernel_ms']:.3f} ms  wall (host out) {w:.3f} ms  {sim.timing['n_steps']*n/sim.timing['kernel_ms']/1e-3:.3e} traj-steps/s")
