"""Quick performance probe (not the bench): C2 at a given size, Philox noise, device-resident output."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sdde_b200 as S

n_traj = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
T = float(sys.argv[2]) if len(sys.argv) > 2 else 500.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = S.get_context(0)
import ctypes as C
v = C.c_double(0)
S._capi.check(S.lib().sdde_b200_dfma_peak(ctx.handle, C.byref(v))); print("dfma peak TFLOP/s", v.value)
S._capi.check(S.lib().sdde_b200_hbm_peak(ctx.handle, C.byref(v))); print("copy GB/s", v.value)
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, T), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
for opt in [dict(), dict(batch=1), dict(batch=2)]:
    for k, val in opt.items():
        ctx.set_option(k, val)
    for r in range(reps):
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0,
                      trajectories=n_traj, seed=1234)
        tm = sim.timing
        steps = tm["n_steps"] * n_traj
        print(opt, f"kernel {tm['kernel_ms']:.2f} ms  {steps / tm['kernel_ms'] / 1e-3:.3e} traj-steps/s  "
              f"{steps * 16.4 / tm['kernel_ms'] / 1e-3 / 1e9:.0f} GB/s algorithmic  regs {tm['regs_per_thread']} waves {tm['waves']}")
    ctx.set_option("batch", 0)
print(json.dumps(sim.timing))
