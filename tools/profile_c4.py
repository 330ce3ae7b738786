import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import sdde_b200 as S, _problems as cases
case = cases.c4()
for r in range(2):
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), trajectories=1_000_000, seed=3, **case.solve_kwargs())
    print(sim.timing["kernel_ms"], sim.timing["n_steps"])
