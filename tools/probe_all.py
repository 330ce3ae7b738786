"""Throughput of every BASELINE.json config on one GPU (Philox noise, device-resident output)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases

def run(name, case, n, bytes_per_step, reps=3, **extra):
    for r in range(reps):
        sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0, output="device"),
                      trajectories=n, seed=3, **case.solve_kwargs(), **extra)
    tm = sim.timing
    ts = tm["n_steps"] * n
    rate = ts / tm["kernel_ms"] / 1e-3
    print(f"{name:34s} n={n:>9d} steps={tm['n_steps']:>6d} saves={tm['n_save']:>4d} kernel {tm['kernel_ms']:8.2f} ms  {rate:.3e} traj-steps/s  "
          f"{rate*bytes_per_step/1e9:6.0f} GB/s alg ({bytes_per_step:.1f} B/step)  regs {tm['regs_per_thread']}  {tm['kernel_name']}")

run("C1 linear1 EM T=10 saveat 0.1", cases.c1(), 1_000_000, 16 + 8 * 101 / 1001)
run("C2 MG EM tau=17 T=500", cases.c2(T=500.0), 1_000_000, 16.4)
run("C2 MG EM(false)", cases.c2(T=500.0, alg=1), 1_000_000, 16.4)
run("C3 MG RKMil tau=17.02 order0=0 T=100", cases.c3(0), 1_000_000, 16 + 8 * 101 / 2002)
run("C3 MG RKMil tau=17.02 order0=2 T=100", cases.c3(2), 1_000_000, 16 + 8 * 101 / 2001)
run("C3 MG RKMil Strat", cases.c3(0, alg=3), 1_000_000, 16 + 8 * 101 / 2002)
run("C4 linsys 4/2/2 EM T=20", cases.c4(), 1_000_000, 96 + 32 * 41 / 2003)   # two delayed nodes read + one written per step
run("C5 linear1 EM T=10 moments only", cases.c1(T=10.0, saveat=1.0), 4_000_000, 16 + 8 * 11 / 1001, moments=True, paths=False)
run("hayes EM T=10 dt=2^-6", cases.hayes(T=10.0, dt=2.0**-6, saveat=1.0), 1_000_000, 16.1)
run("sir EM T=20 dt=0.02", cases.sir(T=20.0, dt=0.02), 1_000_000, 48 + 24 * 21 / 1001)
