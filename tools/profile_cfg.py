"""One solve of a named config for ncu (device output, in-kernel Philox): python tools/profile_cfg.py <sir|c3strat|c3ito|c4|c1|hayes|eh|sra1|sosra> [reps]
Prints kernel ms per repetition; under ncu use reps = 1 and -k regex:sdde_solve -c 1."""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases
def _additive(alg):
    c = cases.hayes(T=10.0, dt=0.01, saveat=0.1, alg=alg)
    c.p = [1.0, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]
    return c


name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
case = {"sir": lambda: cases.sir(T=20.0, dt=0.02), "c3strat": lambda: cases.c3(0, alg=3), "c3ito": lambda: cases.c3(0, alg=2),
        "c4": lambda: cases.c4(), "c1": lambda: cases.c1(), "hayes": lambda: cases.hayes(T=10.0, dt=2.0 ** -6, saveat=1.0),
        "eh": lambda: cases.hayes(T=10.0, dt=0.01, saveat=0.1, alg=4), "sra1": lambda: _additive(13), "sosra": lambda: _additive(15)}[name]()
for r in range(reps):
    sim = S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0, output="device"),
                  trajectories=1_000_000, seed=3, **case.solve_kwargs())
    tm = sim.timing
    print(f"{name}: kernel {tm['kernel_ms']:.3f} ms  {1e6 * tm['n_steps'] / tm['kernel_ms'] / 1e-3:.3e} traj-steps/s  regs {tm['regs_per_thread']} "
          f"block {tm['block_threads']}  {tm['kernel_name']}", flush=True)
