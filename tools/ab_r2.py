"""Same-box A/B of library builds on the C2 bench launch (10^6 paths x 10 000 steps, in-kernel Philox, device output).
usage: python tools/ab_r2.py [--rounds R] [--reps K] [--opts "k=v,k=v;k=v"] lib_a.so lib_b.so ...   ('default' = the in-tree library)
Each (library, option set) runs in its own subprocess; prints min / median kernel ms per combination and round."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %r)
import sdde_b200 as S
opts = sys.argv[1]; reps = int(sys.argv[2])
ctx = S.get_context(0)
for kv in [o for o in opts.split(",") if o]:
    k, v = kv.split("="); ctx.set_option(k, int(v))
prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ms = []
for r in range(reps):
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0,
                  trajectories=1_000_000, seed=1234)
    ms.append(sim.timing["kernel_ms"])
print(json.dumps({"ms": ms, "regs": sim.timing["regs_per_thread"], "block": sim.timing["block_threads"], "name": sim.timing["kernel_name"],
                  "mean_end": float(sim.out[-1].mean()) if hasattr(sim.out, "mean") else None}))
''' % ROOT
args = sys.argv[1:]
rounds, reps, optsets = 2, 5, ["lean=1", "lean=0"]
while args and args[0].startswith("--"):
    if args[0] == "--rounds": rounds = int(args[1]); args = args[2:]
    elif args[0] == "--reps": reps = int(args[1]); args = args[2:]
    elif args[0] == "--opts": optsets = args[1].split(";"); args = args[2:]
for rd in range(rounds):
    for lib in args:
        for lean in optsets:
            env = dict(os.environ)
            if lib != "default": env["SDDE_B200_LIB"] = os.path.join(ROOT, lib)
            p = subprocess.run([sys.executable, "-c", CHILD, str(lean), str(reps)], env=env, capture_output=True, text=True)
            if p.returncode != 0:
                print(f"{lib} {lean}: FAILED\n{p.stderr[-1500:]}"); continue
            d = json.loads(p.stdout.strip().splitlines()[-1])
            ms = sorted(d["ms"][1:]) if len(d["ms"]) > 1 else d["ms"]
            print(f"round {rd} {lib:28s} {lean:24s} min {ms[0]:7.3f} med {ms[len(ms)//2]:7.3f} ms  regs {d['regs']} block {d['block']} mean(t_end) {d['mean_end']}", flush=True)
