"""Time sdde_b200_quantiles on the C2 bench output (501 rows x 10^6 trajectories, resident in HBM)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import sdde_b200 as S

n_rows, n = int(os.environ.get("ROWS", 501)), int(os.environ.get("NTRAJ", 1_000_000))
x = 1.0 + 0.2 * torch.randn(n_rows, 1, n, dtype=torch.float64, device="cuda")
sim = S.EnsembleSolution(np.arange(float(n_rows)), x, None, None, False, 0.0, {})
for rep in range(4):
    t0 = time.perf_counter()
    q = S.timeseries_steps_quantile(sim, [0.05, 0.95], median=True)
    wall = (time.perf_counter() - t0) * 1e3
    ms = sim.timing["quantile_kernel_ms"]
    print(f"rows={n_rows} n={n} kernel_ms={ms:.3f} wall_ms={wall:.3f} stream_GBs={6 * n_rows * n * 8 / ms / 1e6:.0f} "
          f"(6 passes x 8 B/value)  one_pass_equiv_GBs={n_rows * n * 8 / ms / 1e6:.0f}")
ref = torch.quantile(x[:4, 0, :], torch.tensor([0.05, 0.5, 0.95], dtype=torch.float64, device="cuda"), dim=1).T.cpu().numpy()
print("max |diff| vs torch.quantile on 4 rows:", np.abs(q[:4, 0, :][:, [0, 2, 1]] - ref).max())
