"""staged host-output C2 solve: wall time vs the size of the first wave (option ramp_start)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ctx = S.get_context(0)
ctx.set_option("zero_copy", 0)
for rs in (65536, 4096, 8192, 16384, 32768, 65536, 131072):
    ctx.set_option("ramp_start", rs)
    ws = []
    for r in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
        ws.append((time.perf_counter() - t0) * 1e3)
    tm = sim.timing
    print(f"ramp_start={rs} waves={tm['waves']} wall min {min(ws):.2f} med {sorted(ws)[2]:.2f} kernel {tm['kernel_ms']:.1f} d2h {tm['d2h_ms']:.1f}", flush=True)
