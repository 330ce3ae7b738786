"""e2e (host output) wall time against the size of the first wave.  python tools/probe_ramp.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ctx = S.get_context(0)
for rs in (1 << 16, 1 << 15, 1 << 14, 1 << 13, 1 << 16, 1 << 14):
    ctx.set_option("ramp_start", rs)
    best = 1e9
    for r in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
        wall = (time.perf_counter() - t0) * 1e3
        if r > 0: best = min(best, wall)
    print(f"ramp_start={rs} waves={sim.timing['waves']} best wall {best:.2f} ms", flush=True)
