import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
out = torch.empty((501, 1, 1_000_000), dtype=torch.float64, device="cuda:0")
for r in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0, trajectories=1_000_000, seed=r + 1, out=out)
    torch.cuda.synchronize(); w = (time.perf_counter() - t0) * 1e3
    print(f"call {r}: wall {w:.2f} ms  lib total {sim.timing['total_ms']:.2f} ms  kernel {sim.timing['kernel_ms']:.2f} ms")
