"""C5: 1e8-trajectory scalar linear SDDE, moments only.  python tools/probe_c5.py [n_traj]"""
import sys, os, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")]
import numpy as np
import sdde_b200 as S, cases, sdde_oracle as O
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
sch = case.oracle_schedule()
m = O.exact_mean_linear1(sch, -1.0, 0.5, 1.0, 1.0, 0.0, 1.0)
mean, var = S.timeseries_steps_meanvar(sim)
idx = [int(np.argmin(np.abs(sch.t - t))) for t in sim.t]
z = (mean[:, 0] - m[idx])[1:] / np.sqrt(var[1:, 0] / n)
print("count", sim.moments[-1, 0, 0], "mean(t=10)", mean[-1, 0], "exact", m[idx][-1], "z-scores vs exact mean recurrence:", np.round(z, 2))
