"""e2e probe: host pinned output, different wave sizes.  python tools/probe_e2e.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sdde_b200 as S
n = 1_000_000
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
out = torch.empty((501, 1, n), dtype=torch.float64, pin_memory=True)
for wave in (0, 1 << 17, 1 << 16, 1 << 15, 1 << 19):
    for r in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, wave_traj=wave or None), dt=0.05, saveat=1.0, trajectories=n, seed=5, out=out)
        wall = (time.perf_counter() - t0) * 1e3
    S.get_context(0).set_option("wave_traj", 0)
    tm = sim.timing
    print(f"wave={wave or 'auto'} waves={tm['waves']} wall {wall:.1f} ms  compute-stream {tm['kernel_ms']:.1f} ms  copy-stream {tm['d2h_ms']:.1f} ms  d2h {tm['d2h_bytes']/1e9:.2f} GB -> {tm['d2h_bytes']/1e9/(tm['d2h_ms']*1e-3):.1f} GB/s")
