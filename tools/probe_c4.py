"""C4 throughput: thread-per-trajectory vs cooperative lanes.  python tools/probe_c4.py [n_traj]"""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
case = cases.c4()
for lanes in (0, 4):
    for r in range(3):
        sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device", lanes_per_traj=lanes),
                      trajectories=n, seed=3, **case.solve_kwargs())
    tm = sim.timing
    ts = tm["n_steps"] * n
    print(f"lanes={lanes or 1} {tm['kernel_name']} regs {tm['regs_per_thread']} kernel {tm['kernel_ms']:.2f} ms  {ts/tm['kernel_ms']/1e-3:.3e} traj-steps/s  "
          f"{ts*(64+8*4*41/2003)/tm['kernel_ms']/1e-3/1e9:.0f} GB/s algorithmic (64.7 B/step)")
