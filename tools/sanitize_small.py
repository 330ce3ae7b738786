"""Small end-to-end workload for compute-sanitizer (memcheck / racecheck): every kernel family once, tiny sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import sdde_b200 as S
import _problems as cases

def run(case, n=777, **kw):
    return S.solve(S.EnsembleProblem(case.problem()), cases.ALG_MAP[case.alg], S.EnsembleB200(device=0), trajectories=n, seed=3,
                   moments=True, **case.solve_kwargs(), **kw)

for case in (cases.c1(T=3.0), cases.c2(T=40.0), cases.c3(T=40.0), cases.c3(order0=2, T=40.0, alg=3), cases.c2(T=30.0, alg=4),
             cases.c4(T=4.0), cases.c4(T=2.0).with_(alg=5), cases.sir(T=4.0, dt=0.02), cases.scalar_add(T=3.0), cases.hayes(T=2.0),
             cases.hayes(T=2.0, alg=9), cases.c2(T=25.0, alg=12)):
    sim = run(case)
    print(case.name, "ok", float(np.asarray(sim.out)[-1].mean()), flush=True)
# SRA family: four lookup sets per step, two increments per channel (in-kernel Philox and supplied W, Z grids)
import sdde_b200 as _S
hp = cases.hayes(T=2.0, dt=0.05)
hp.p = np.array([0.33, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1]); hp.lags = [0.33]
for alg in (_S.SRA1(), _S.SRA3(), _S.SOSRA()):
    sim = S.solve(S.EnsembleProblem(hp.problem()), alg, S.EnsembleB200(device=0), trajectories=777, seed=3, moments=True, dt=0.05,
                  saveat=0.25, adaptive=False)
    print("sra", alg.name, "ok", float(np.asarray(sim.out)[-1].mean()), flush=True)
# supplied increments, per-trajectory parameters, several waves, device output, quantiles
case = cases.c1(T=2.0)
dW = cases.increments(case, 300)
sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, wave_traj=128), trajectories=300,
              noise_increments=dW, **case.solve_kwargs())
simd = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), trajectories=5000, seed=9,
               **case.solve_kwargs())
q = S.timeseries_steps_quantile(simd, [0.05, 0.95], median=True)
print("quantiles ok", q.shape, flush=True)
ea = S.EnsembleB200(device=0, lanes_per_traj=4)
c4 = cases.c4(T=2.0)
simc = S.solve(S.EnsembleProblem(c4.problem()), S.EM(), ea, trajectories=333, seed=3, **c4.solve_kwargs())
print("coop ok", flush=True)
# launch geometry: full CTAs + a tail round, and a wave that ends inside a warp (in-kernel Philox: the production path)
case = cases.c2(T=20.0)
for n in (148 * 1024 + 1500, 2 * 148 * 1024, 1025, 63):
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), trajectories=n, seed=5,
                  moments=True, **case.solve_kwargs())
    print("geometry ok", n, sim.timing["waves"], flush=True)
