"""One pass of the order-statistics kernels and of the one-read moments kernel on the C2 output shape (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, sdde_b200 as S
n_rows, n = 501, 1_000_000
x = 1.0 + 0.2 * torch.randn(n_rows, 1, n, dtype=torch.float64, device="cuda")
sim = S.EnsembleSolution(np.arange(float(n_rows)), x, None, None, False, 0.0, {})
for r in range(2):
    summ = S.EnsembleSummary(sim)          # moments_kernel + merge, then qs_stream/qs_scan passes
    sim.moments = None
print("median[0]", summ.med[0, 0], "quantile kernels ms", sim.timing["quantile_kernel_ms"])
