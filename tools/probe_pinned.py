import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
import numpy as np, sdde_b200 as S, cases
L = S.lib(); S.get_context(0)
for mb in (8, 808, 4009):
    t0 = time.perf_counter(); p = L.sdde_b200_alloc_pinned(mb << 20); t1 = time.perf_counter(); L.sdde_b200_free_pinned(p); t2 = time.perf_counter()
    t3 = time.perf_counter(); a = np.empty(mb << 17); a[::512] = 0; t4 = time.perf_counter()
    print(f"{mb} MB: cudaHostAlloc {1e3*(t1-t0):.1f} ms, free {1e3*(t2-t1):.1f} ms; np.empty+first touch {1e3*(t4-t3):.1f} ms")
case = cases.c1(T=10.0, saveat=0.1)
n = 1_000_000
buf = L.sdde_b200_alloc_pinned(101 * n * 8)
out = np.ctypeslib.as_array((C.c_double * (101 * n)).from_address(buf)).reshape(101, 1, n)
for r in range(3):
    t0 = time.perf_counter()
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=n, seed=3, out=out, **case.solve_kwargs())
    print(f"pinned out: wall {1e3*(time.perf_counter()-t0):.2f} ms kernel-stream {sim.timing['kernel_ms']:.2f} ms")
