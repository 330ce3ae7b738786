"""C5: 1e8-trajectory scalar linear SDDE, moments only.  python tools/probe_c5.py [n_traj]"""
import sys, os, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import numpy as np
import sdde_b200 as S, _problems as cases
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
case = cases.c1(T=10.0, saveat=1.0)
for r in range(2):
    t0 = time.perf_counter()
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=0xB2000001,
                  moments=True, paths=False, **case.solve_kwargs())
    wall = time.perf_counter() - t0
    tm = sim.timing
    ts = tm["n_steps"] * n
    print(f"n={n} waves {tm['waves']} launches {tm['launches']} kernel {tm['kernel_ms']:.1f} ms wall {wall*1e3:.1f} ms  {ts/wall:.3e} traj-steps/s (wall)  "
          f"{ts/tm['kernel_ms']/1e-3:.3e} (kernel)  {tm['kernel_name']}")
mean, var = S.timeseries_steps_meanvar(sim)
print("count", sim.moments[-1, 0, 0], "mean(t=10)", mean[-1, 0], "var(t=10)", var[-1, 0],
      "(the exact-mean check of this configuration is tests/test_gpu_rng.py::test_k3_ensemble_mean_against_exact_recurrence)")
