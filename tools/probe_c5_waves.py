"""C5 (10^8 scalar trajectories, moments only): wall time against the wave size.  python tools/probe_c5_waves.py"""
import sys, os, time
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases
n = 100_000_000
case = cases.c1(T=10.0, saveat=1.0)
for wave in (0, 1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24):
    for r in range(2):
        t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, wave_traj=wave or None), trajectories=n,
                      seed=0xB2000001, moments=True, paths=False, **case.solve_kwargs())
        wall = time.perf_counter() - t0
    S.get_context(0).set_option("wave_traj", 0)
    tm = sim.timing
    print(f"wave={wave or 'auto'} waves {tm['waves']} kernel {tm['kernel_ms']:.1f} ms wall {wall*1e3:.1f} ms {tm['n_steps']*n/wall:.3e} traj-steps/s mean {sim.moments[-1,0,1]:.12f}", flush=True)
