"""e2e timeline probe: PCIe D2H rate (contiguous vs the strided wave copies) and the wall time of the host-output solve."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
dev = torch.empty((501, 1, n), dtype=torch.float64, device="cuda")
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best
t = timed(lambda: out.copy_(dev, non_blocking=True))
print(f"contiguous D2H 4.008 GB: {t*1e3:.1f} ms = {4.008/t:.1f} GB/s")
for w in (1 << 16, 1 << 19):
    t = timed(lambda: out[:, :, :w].copy_(dev[:, :, :w], non_blocking=True))
    print(f"strided D2H of {w} columns ({501*w*8/1e9:.2f} GB): {t*1e3:.1f} ms = {501*w*8/1e9/t:.1f} GB/s")
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
for r in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
    wall = (time.perf_counter() - t0) * 1e3
    tm = sim.timing
    print(f"waves={tm['waves']} wall {wall:.1f} ms  lib total {tm['total_ms']:.1f}  compute-stream {tm['kernel_ms']:.1f} ms  copy-stream {tm['d2h_ms']:.1f} ms -> {tm['d2h_bytes']/1e9/(tm['d2h_ms']*1e-3):.1f} GB/s", flush=True)
