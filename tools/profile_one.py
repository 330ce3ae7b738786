"""One C2 solve for ncu: python tools/profile_one.py <n_traj> <T> [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdde_b200 as S
n_traj = int(float(sys.argv[1])); T = float(sys.argv[2]); reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, T), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
if os.environ.get("SDDE_CHUNK"): S.get_context(0).set_option("chunk_steps", int(os.environ["SDDE_CHUNK"]))
for r in range(reps):
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0,
                  trajectories=n_traj, seed=1234)
    print(sim.timing["kernel_ms"], sim.timing["n_steps"])
