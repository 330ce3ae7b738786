"""A long integration: C2 over t in [0, 5000] (100 000 steps, ~290 trips round the history ring, 782 schedule chunks) against the
oracle under shared increments, and save_everystep over 70 000 steps (more save rows than a 16-bit counter holds)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, sdde_b200 as S, sdde_oracle as O, cases
case = cases.c2(T=5000.0).with_(saveat=10.0)
sch = case.oracle_schedule()
n = 48
dW = cases.increments(sch, 1, n)
t0 = time.time(); ref, st = case.oracle(dW); print("oracle", round(time.time() - t0, 1), "s; steps", sch.n_steps, "saves", len(sch.save_t))
sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, noise_increments=dW, **case.solve_kwargs())
print("C2 T=5000: scaled err", cases.scaled_err(sim.out, ref), "status equal", np.array_equal(sim.status, st))
# chaotic amplification over 5000 time units is large: also report the error over the first 1000 time units
k = int(np.searchsorted(sch.save_t, 1000.0))
print("   first 1000 time units:", cases.scaled_err(sim.out[:k], ref[:k]))
case = cases.c1(T=700.0).with_(saveat=None)
sch = case.oracle_schedule()
dW = cases.increments(sch, 1, 8)
ref, st = case.oracle(dW)
sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=8, noise_increments=dW, **case.solve_kwargs())
print("C1 save_everystep:", sim.out.shape, "scaled err", cases.scaled_err(sim.out, ref), "t equal", np.array_equal(sim.t, sch.save_t))
