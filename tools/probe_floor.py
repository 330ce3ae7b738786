"""Where is the memory floor of the step kernel?  The C2 bench launch (10^6 paths x 10 000 steps) with the lag shrunk so
that the history ring fits in L2 (tau = 0.2: 6 slots x 8 MB), against the real tau = 17 (342 slots, 2.7 GB, streamed from
HBM), and with sparse output.  Same instruction stream per step; only where the ring traffic lands changes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdde_b200 as S
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for tau, saveat in [(17.0, 1.0), (0.2, 1.0), (17.0, 100.0), (0.2, 100.0), (1.7, 1.0)]:
    prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(tau,))
    ms = []
    for r in range(reps):
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=saveat,
                      trajectories=1_000_000, seed=1234)
        ms.append(sim.timing["kernel_ms"])
    m = sorted(ms[1:])
    print(f"tau={tau:5.1f} saveat={saveat:5.1f} steps={sim.timing['n_steps']} ring_slots={sim.timing['ring_slots']} "
          f"ring {sim.timing['ring_bytes']/1e6:.0f} MB: min {m[0]:.3f} med {m[len(m)//2]:.3f} ms  {sim.timing['kernel_name']}", flush=True)
