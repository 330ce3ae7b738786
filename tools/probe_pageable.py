"""Host-output solve into PAGEABLE numpy memory (what a plain S.solve() returns) vs page-locked memory."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sdde_b200 as S
n = 1_000_000
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
pinned = S.pinned_empty((501, 1, n))
page = np.empty((501, 1, n))
page[:] = 0.0
for name, o in (("pinned", pinned), ("pageable (preallocated, touched)", page), ("pageable (allocated by solve)", None)):
    for r in range(3):
        t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=o)
        wall = (time.perf_counter() - t0) * 1e3
    print(f"{name}: wall {wall:.1f} ms  lib {sim.timing['total_ms']:.1f} ms  waves {sim.timing['waves']}", flush=True)
