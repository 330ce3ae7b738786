import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
torch.cuda.set_device(0)
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
eprob = S.EnsembleProblem(prob)
out = torch.empty((501, 1, 1_000_000), dtype=torch.float64, device="cuda:0")
ea = S.EnsembleB200(device=0, distributed=False, output="device")
for mode in ("small_seed", "big_seed"):
    for r in range(8):
        seed = (r + 1) if mode == "small_seed" else 0xB2000001 + r
        t0 = time.perf_counter()
        sim = S.solve(eprob, S.EM(), ea, dt=0.05, saveat=1.0, trajectories=1_000_000, seed=seed, out=out, moments=False)
        w = (time.perf_counter() - t0) * 1e3
        print(f"{mode} call {r}: wall {w:.2f} ms  lib total {sim.timing['total_ms']:.2f} ms  kernel {sim.timing['kernel_ms']:.2f} ms h2d {sim.timing['h2d_bytes']}")
