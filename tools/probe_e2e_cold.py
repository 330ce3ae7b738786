"""e2e breakdown: wall time of the host-output C2 solve, zero-copy vs staged, warm vs cold library caches
(SDDE_DEBUG_TIMING=1 prints the library's own laps on stderr)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ctx = S.get_context(0)
for zc in (1, 0, 1, 0):
    ctx.set_option("zero_copy", zc)
    for cold in (0, 1):
        ws = []
        for r in range(4):
            if cold: ctx.set_option("drop_caches", 1)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
            ws.append((time.perf_counter() - t0) * 1e3)
        tm = sim.timing
        print(f"zero_copy={zc} cold={cold} waves={tm['waves']} wall {['%.1f' % w for w in ws]} lib total {tm['total_ms']:.1f} kernel {tm['kernel_ms']:.1f} d2h {tm['d2h_ms']:.1f} h2d_bytes {tm['h2d_bytes']}", flush=True)
