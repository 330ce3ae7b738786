"""Host overhead per device-resident C2 solve: Python wall time vs the library's own wall time vs the kernel time."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ep = S.EnsembleProblem(prob)
ea = S.EnsembleB200(device=0, output="device")
out = torch.empty((501, 1, n), dtype=torch.float64, device="cuda")
def one(i):
    return S.solve(ep, S.EM(), ea, dt=0.05, saveat=1.0, trajectories=n, seed=1 + i, out=out)
for i in range(3): one(i)
torch.cuda.synchronize()
for i in range(6):
    t0 = time.perf_counter(); sim = one(i); w = (time.perf_counter() - t0) * 1e3
    print(f"wall {w:.3f} lib {sim.timing['total_ms']:.3f} kernel {sim.timing['kernel_ms']:.3f}  python {w - sim.timing['total_ms']:.3f}  lib-host {sim.timing['total_ms'] - sim.timing['kernel_ms']:.3f}")
pr = cProfile.Profile(); pr.enable()
for i in range(10): one(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
