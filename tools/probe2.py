"""probe2: python tools/probe2.py <n_traj> <T> <tau> [reps]  -- MG EM Philox, device output; prints kernel ms"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdde_b200 as S
n_traj = int(float(sys.argv[1])); T = float(sys.argv[2]); tau = float(sys.argv[3]); reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, T), (0.2, 0.1, 10.0, 0.1), constant_lags=(tau,))
for r in range(reps):
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0,
                  trajectories=n_traj, seed=1234)
    tm = sim.timing
print(f"tau={tau} n={n_traj} steps={tm['n_steps']} ring={tm['ring_slots']} kernel {tm['kernel_ms']:.2f} ms "
      f"{tm['n_steps']*n_traj/tm['kernel_ms']/1e-3:.3e} traj-steps/s")
