"""Sustained (power-capped) kernel time of the C2 launch as a function of where the ring traffic lands: tau = 17 (2.7 GB
ring, streamed from HBM) vs tau = 1.7 (0.29 GB) vs tau = 0.85 (0.15 GB, about L2-sized); 30 back-to-back launches each, mean
of the last 20, NVML SM clock and board power sampled during them."""
import sys, os, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdde_b200 as S, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
taus = [float(x) for x in sys.argv[1:]] or [17.0, 1.7, 0.85, 17.0]
for tau in taus:
    prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(tau,))
    samples, stop = [], threading.Event()
    def poll():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)); time.sleep(0.02)
    ms = []
    th = None
    for r in range(30):
        if r == 10:
            th = threading.Thread(target=poll, daemon=True); th.start()
        sim = S.solve(S.EnsembleProblem(prob), S.EM(), S.EnsembleB200(device=0, output="device"), dt=0.05, saveat=1.0, trajectories=1_000_000, seed=1234)
        ms.append(sim.timing["kernel_ms"])
    stop.set(); th.join()
    clk = sorted(c for c, _ in samples); pw = sorted(p for _, p in samples)
    print(f"tau={tau:5.2f} ring {sim.timing['ring_bytes']/1e6:6.0f} MB: first {ms[0]:.2f} sustained mean {sum(ms[10:])/20:.3f} ms  sm clock median {clk[len(clk)//2]} MHz  power median {pw[len(pw)//2]:.0f} W max {pw[-1]:.0f} W  {sim.timing['kernel_name']} regs {sim.timing['regs_per_thread']} mean(t_end) {float(sim.out[-1].mean()):.12f}", flush=True)
