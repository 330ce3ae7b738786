"""Pageable host output through the bounce path on other shapes: a single very long row (final states of 9 000 001 paths) and
a multi-state output (C4: 41 x 4 rows), each against the device-resident result."""
import sys, os
sys.path[:0] = [os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__))]
import numpy as np, sdde_b200 as S, _problems as cases
# one row longer than a bounce buffer: final state of 9 000 001 paths (72 MB row)
case = cases.c1(T=0.5, saveat=None)
prob = S.EnsembleProblem(case.problem(), output_func=S.FinalState())
n = 9_000_001
a = S.solve(prob, S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=9, dt=0.01)
ref = S.solve(prob, S.EM(), S.EnsembleB200(device=0, output="device"), trajectories=n, seed=9, dt=0.01)
print("final-state rows", a.out.shape, np.array_equal(a.out, ref.out.cpu().numpy()), a.timing["waves"])
# several states: C4, 300 001 paths, 41 x 4 rows
c4 = cases.c4(T=20.0)
n = 300_001
b = S.solve(S.EnsembleProblem(c4.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=9, **c4.solve_kwargs())
refb = S.solve(S.EnsembleProblem(c4.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), trajectories=n, seed=9, **c4.solve_kwargs())
print("c4 rows", b.out.shape, np.array_equal(b.out, refb.out.cpu().numpy()), b.timing["waves"])
