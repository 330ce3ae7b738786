"""auto host-output mode: per-solve arm, device times and wall time of consecutive C2 solves into one pinned buffer, with the
plain-copy ceiling of the same buffer before and after."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
dev = torch.empty((501, 1, n), dtype=torch.float64, device="cuda")
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
def ceiling():
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out.copy_(dev, non_blocking=True); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ctx = S.get_context(0)
print(f"plain copy {ceiling():.1f} ms")
for mode in (-1, 0, 1, -1):
    ctx.set_option("zero_copy", mode)
    for r in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
        w = (time.perf_counter() - t0) * 1e3
        tm = sim.timing
        print(f"mode {mode:2d} solve {r}: {'zero-copy' if 'mapped' in tm['kernel_name'] else 'staged   '} wall {w:6.1f} kernel {tm['kernel_ms']:5.1f} d2h {tm['d2h_ms']:5.1f}", flush=True)
    print(f"plain copy {ceiling():.1f} ms")
