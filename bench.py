#!/usr/bin/env python
"""bench.py -- trajectory-steps/sec of the ensemble SDDE path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's algorithm on the host CPU
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU

A "step" is one pass of the hot path over one batch: one whole ensemble solve of BASELINE.json configs[1]
(stochastic Mackey-Glass, tau=17, EM dt=0.05, t in [0,500] = 10 000 schedule steps, saveat 1.0 = 501 saved points,
10^6 trajectories PER GPU, in-kernel Philox increments).  Trajectories shard across ranks with no data-path collective
(weak scaling: per-GPU work fixed); the only collective is the NCCL all-gather of the (count, mean, M2) moment triples.

  value    = whole-job trajectory-steps/s with the result left in HBM (device-resident output, no D2H)
  e2e      = the same through the public API with HOST buffers, every step: the library's caches are dropped, so the
             schedule is replayed and its tables (~0.8 MB) are uploaded again (H2D), and the 4 GB of saved paths are
             written into pinned host memory (D2H) inside the timed region; `pcie_ceiling_gbs` is the rate plain
             cudaMemcpyAsync reaches for the same bytes with all ranks copying at once
  roofline = 16.4 algorithmic bytes per trajectory-step x trajectory-steps per launch
             / CUDA-event duration of the step kernel (events recorded by the library on its compute stream)
  configs  = the other BASELINE.json configs, short device-resident runs: C1, C3 (order_discontinuity_t0 = 0 and 2), C4
             (thread per trajectory and cooperative lanes), C5 at 10^8 paths IN TOTAL (NCCL moment merge inside the timed
             region), C2 strong (10^6 paths in total) -- each with its kernel, CUDA-event time and algorithmic GB/s
  checks   = agreement computed in the run: a slice of this rank's shard re-solved alone with its global trajectory
             offset (bit for bit), the library's merged moments against a Chan merge of per-rank two-pass moments and
             identical on every rank, the order statistics selected across ranks against a single-rank selection

The reference (Julia) cannot run here (no Julia in the image); the reference arm and cpu_baseline
time oracle/libsdde_oracle.so, the C++ restatement of the reference's per-trajectory algorithm, on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

METRIC = "trajectory-steps/sec (10^6-path Mackey-Glass SDDE, EM)"
UNIT = "trajectory-steps/s"
ALG_BYTES_PER_TRAJ_STEP = 16.4        # 8 B delayed-node read + 8 B new-node write + 8*501/10000 B saved output
T_END, DT, TAU = 500.0, 0.05, 17.0
N_STEPS, N_SAVE = 10000, 501          # schedule replay of this workload (asserted against the library's schedule below)
P_MG = (0.2, 0.1, 10.0, 0.1)
TRAJ_PER_GPU = 1_000_000


def workload_config(world, n_traj=TRAJ_PER_GPU):
    """The `config` block: a function of the world size only, identical in both arms (ours / reference)."""
    return dict(workload="C2 stochastic Mackey-Glass beta=0.2 gamma=0.1 n=10 sigma=0.1 tau=17 u0=h=1.2, EM(split) dt=0.05, "
                         "t in [0,500] (10000 steps), saveat=1.0 (501 points), independent N(0, dt) increments per step",
                trajectories_per_gpu=n_traj, trajectories_total=n_traj * world, n_steps=N_STEPS, n_save=N_SAVE,
                parallelism=f"trajectory-shard x{world}",
                l2="working set 2.7 GB ring + 4.0 GB output per launch >> 126 MB L2 (no explicit flush needed)")


def mg_problem(S):
    return S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, T_END), P_MG, constant_lags=(TAU,))


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is started
    before the warm-up (nvidia-smi takes a while to produce its first line); only samples whose timestamp falls inside
    [mark_start, mark_end] are reported."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines, self.t0, self.t1 = gpu_index, None, [], None, None

    def start(self):
        if os.environ.get("SDDE_BENCH_NO_SAMPLER"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw, n_all = [], [], set(), [], 0
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 10:
                continue
            n_all += 1
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts) + 0.06):
                continue
            try:
                sm.append(float(f[2])); mx.append(float(f[3])); pw.append(float(f[4]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[6:10]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "samples_total": n_all, "reasons": sorted(reasons)}


def cpu_reference_run(steps, warmup, sample_paths=None, rng="ziggurat"):
    """The reference's per-trajectory algorithm (C++ restatement) on all host cores; returns (value, info, ms per pass)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import sdde_oracle as O
    cores = os.cpu_count() or 1
    n = sample_paths or max(64, cores * 1500)
    sv = O.expand_saveat(1.0, 0.0, T_END)[0]
    times, n_steps = [], 0
    for i in range(warmup + steps):
        _, _, el, n_steps = c_oracle.solve(O.MODEL_MACKEY_GLASS, O.ALG_EM_SPLIT, 1, 1, 1, list(P_MG), [1.2], [1.2], [0.0], [TAU],
                                           0.0, T_END, DT, sv, 501, n_traj=n, seed=0xB2000001 + i, n_threads=cores, rng=rng)
        if i >= warmup:
            times.append(el)
    total = sum(times)
    value = n * n_steps * len(times) / total
    info = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "rng": ("xoroshiro128+ + 128-layer ziggurat (the cost class of the reference's Xoroshiro128Plus + randn; timing only, "
                    "oracle/sdde_oracle.cpp); the restatement omits Julia's per-trajectory __init, deepcopy and dynamic dispatch, so "
                    "it is expected to be faster than the Julia reference per core") if rng == "ziggurat" else
                   ("Philox4x32-10 + libm Box-Muller (log, sqrt, sin, cos per pair): the generator the parity tests share with the "
                    "device, several times dearer per normal than the reference's ziggurat"),
            "sample": f"{n} of the 10^6 trajectories x {n_steps} steps per pass, {len(times)} passes, C++ restatement of the "
                      f"reference's per-trajectory loop (oracle/sdde_oracle.cpp: growing history + binary search), std::thread "
                      f"static chunks over {cores} cores; not the Julia reference (Julia is not in the image)"}
    return value, info, total / len(times) * 1e3


# ------------------------------------------------------------------------------------------------------------------
# the other BASELINE.json configs (short, device-resident) and the agreement checks
# ------------------------------------------------------------------------------------------------------------------
def _c4_params(S):
    import numpy as np
    A = -1.0 * np.eye(4) + 0.1 * (np.ones((4, 4)) - np.eye(4))
    B1 = 0.3 * np.eye(4)
    B2 = -0.2 * np.roll(np.eye(4), 1, axis=1)
    M = np.zeros((4, 2))
    for (i, j) in [(0, 0), (1, 0), (1, 1), (2, 1), (3, 1)]:
        M[i, j] = 0.1
    return S.linsys_params(A, [B1, B2], np.zeros(4), M, np.zeros((4, 2)))


def config_block(S, torch, dist, local, world, barrier_sync, peak_gbs):
    """Every BASELINE.json config besides the headline one; per config: kernel name, CUDA-event kernel time (max over
    ranks), wall time of the solve call (max over ranks), whole-job trajectory-steps/s, algorithmic bytes per
    trajectory-step (16 n_state per lag read+write ... + saved output) and the fraction of the HBM figure."""
    dev = torch.device(f"cuda:{local}")
    u4 = [1.0, 0.5, 0.25, 0.125]
    mg = lambda tau, T, o0=0: S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, T), P_MG, constant_lags=(tau,),
                                            order_discontinuity_t0=o0)
    lin = S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0.0, 10.0), (-1.0, 0.5, 0.2), constant_lags=(1.0,))
    m4 = S.linsys(4, 2, 2)
    c4 = S.SDDEProblem(m4.f, m4.g, u4, S.AffineHistory(tuple(u4), (0.0,) * 4), (0.0, 20.0), _c4_params(S),
                       constant_lags=(1.0, 2.5))
    M = TRAJ_PER_GPU
    # name, problem, alg, solve kwargs, trajectories (total), sharded over ranks?, ensemble alg kwargs, bytes(n_state, n_lag reads)
    runs = [
        ("C1 linear scalar EM dt=0.01 T=10 saveat=0.1, 10^6 paths/GPU", lin, S.EM(), dict(dt=0.01, saveat=0.1), M * world, {}, 16.0, 1),
        ("C3 Mackey-Glass tau=17.02 RKMil(Ito) order_discontinuity_t0=0 T=100, 10^6 paths/GPU", mg(17.02, 100.0, 0), S.RKMil(),
         dict(dt=0.05, saveat=1.0, adaptive=False), M * world, {}, 16.0, 1),
        ("C3 Mackey-Glass tau=17.02 RKMil(Ito) order_discontinuity_t0=2 T=100, 10^6 paths/GPU", mg(17.02, 100.0, 2), S.RKMil(),
         dict(dt=0.05, saveat=1.0, adaptive=False), M * world, {}, 16.0, 1),
        ("C4 4-state 2-lag 4x2 noise EM dt=0.01 T=20 saveat=0.5, thread per trajectory, 10^6 paths/GPU", c4, S.EM(),
         dict(dt=0.01, saveat=0.5), M * world, {}, 96.0, 4),
        ("C4 same, cooperative lanes (4 lanes per trajectory), 2.5x10^5 paths/GPU", c4, S.EM(), dict(dt=0.01, saveat=0.5),
         (M // 4) * world, dict(lanes_per_traj=4), 96.0, 4),
        ("C5 linear scalar EM, 10^8 paths IN TOTAL, moments only at saveat=1.0, merge of the moment triples inside the timed region",
         lin, S.EM(), dict(dt=0.01, saveat=1.0, moments=True, paths=False), 100_000_000, {}, 16.0, 1),
        ("C2 strong: the headline workload with 10^6 paths IN TOTAL", mg(TAU, T_END), S.EM(), dict(dt=DT, saveat=1.0), M, {}, 16.0, 1),
    ]
    # the other fixed-step solvers of SURVEY.md section 8(f) rank 3 on the reference's own test problems (test/test_prob_sol.jl
    # Hayes equation, lag 1, dt = 0.01, t in [0, 10]; test/nondiagonal_sparse_noise.jl SIR), 10^6 paths per GPU
    hay = lambda p: S.SDDEProblem(S.hayes.f, S.hayes.g, 1.0, S.AffineHistory((1.0,), (1.0,)), (0.0, 10.0), p, constant_lags=(1.0,))
    pmul, padd = (1.0, -4.0, -2.0, 10.0, -1.3, -1.2, 1.1), (1.0, -4.0, -2.0, 10.0, 0.0, 0.6, 0.1)
    sir = S.SDDEProblem(S.sir.f, S.sir.g, [990.0, 10.0, 0.0], S.AffineHistory((1000.0, 0.0, 0.0), (0.0, 0.0, 0.0)), (0.0, 20.0),
                        (0.05, 10.0, 4.0), lags=(4.0,))
    runs += [
        ("A1 Hayes EulerHeun dt=0.01 T=10 saveat=0.1", hay(pmul), S.EulerHeun(), dict(dt=0.01, saveat=0.1), M * world, {}, 16.0, 1),
        ("A2 Hayes WangLi3SMil_A dt=0.01 T=10 saveat=0.1", hay(pmul), S.WangLi3SMil_A(), dict(dt=0.01, saveat=0.1), M * world, {}, 16.0, 1),
        ("A3 additive Hayes SRA1 (fixed step) dt=0.01 T=10 saveat=0.1", hay(padd), S.SRA1(), dict(dt=0.01, saveat=0.1, adaptive=False),
         M * world, {}, 16.0, 1),
        ("A4 additive Hayes SOSRA (fixed step) dt=0.01 T=10 saveat=0.1", hay(padd), S.SOSRA(), dict(dt=0.01, saveat=0.1, adaptive=False),
         M * world, {}, 16.0, 1),
        ("A5 delayed SIR, 3 states, 3x2 sparse noise, EM dt=0.02 T=20 saveat=1.0", sir, S.EM(), dict(dt=0.02, saveat=1.0), M * world, {}, 48.0, 3),
    ]
    out = []
    for name, prob, alg, kw, n_total, eakw, ring_bytes, n_state in runs:
        ea = S.EnsembleB200(device=local, distributed=world > 1, output="device", **eakw)
        ep = S.EnsembleProblem(prob)
        reps, best_wall, best_k, sim = 3, None, None, None
        for r in range(reps):
            barrier_sync()
            t0 = time.perf_counter()
            sim = S.solve(ep, alg, ea, trajectories=n_total, seed=0xC0F16 + r, **kw)
            barrier_sync()
            t = torch.tensor([time.perf_counter() - t0, sim.timing["kernel_ms"]], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if r > 0 and (best_wall is None or float(t[0]) < best_wall):        # first pass is the warm-up
                best_wall, best_k = float(t[0]), float(t[1])
        tm = sim.timing
        n_steps, n_save = tm["n_steps"], tm["n_save"]
        bps = ring_bytes + 8.0 * n_state * n_save / n_steps
        per_gpu = -(-n_total // world)
        gbs = bps * per_gpu * n_steps / (best_k * 1e-3) / 1e9
        out.append({"name": name, "kernel": tm["kernel_name"], "trajectories_total": n_total, "n_steps": n_steps, "n_save": n_save,
                    "waves_per_gpu": tm["waves"], "kernel_ms": best_k, "ms": best_wall * 1e3,
                    "value": n_total * n_steps / best_wall, "value_kernel_only": n_total * n_steps / (best_k * 1e-3),
                    "alg_bytes_per_step": bps, "achieved_gbs_per_gpu": gbs, "frac": gbs / peak_gbs,
                    "regs_per_thread": tm["regs_per_thread"], "block_threads": tm["block_threads"]})
        del sim
    return out


def checks_block(S, torch, dist, local, world, rank, eprob, out_dev, n_traj, seed):
    """Agreement checks computed in the run (all ranks take part; every entry is True on success)."""
    import numpy as np
    dev = torch.device(f"cuda:{local}")
    checks = {}

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t[0]))

    ea = S.EnsembleB200(device=local, distributed=world > 1, output="device")
    sim = S.solve(eprob, S.EM(), ea, dt=DT, saveat=1.0, trajectories=n_traj * world, seed=seed, out=out_dev, moments=True)
    lo = sim.traj_offset
    # (1) partition invariance: 4096 paths from the middle of this rank's shard, solved alone with their global offset
    k0 = (n_traj // 2) & ~1
    alone = S.solve(eprob, S.EM(), S.EnsembleB200(device=local, output="device"), dt=DT, saveat=1.0, trajectories=4096, seed=seed,
                    traj_offset=lo + k0)
    checks["slice_resolved_alone_bit_identical"] = all_true(torch.equal(alone.out, sim.out[:, :, k0:k0 + 4096]))
    # (2) merged moments: identical on every rank, and equal to the Chan merge of per-rank TWO-PASS moments
    x = sim.out[:, 0, :]
    mean = x.mean(dim=1)
    m2 = ((x - mean[:, None]) ** 2).sum(dim=1)
    mine = torch.stack([torch.full_like(mean, float(x.shape[1])), mean, m2], dim=1)
    parts = [torch.empty_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, mine)
    else:
        parts = [mine]
    acc = parts[0].cpu().numpy().copy()
    for r in range(1, world):
        acc = S.merge_moments(acc, parts[r].cpu().numpy())
    lib_m = np.asarray(sim.moments)[:, 0, :]
    checks["merged_moments_equal_two_pass_chan_merge"] = all_true(
        bool(np.array_equal(lib_m[:, 0], acc[:, 0])) and bool(np.allclose(lib_m[:, 1], acc[:, 1], rtol=1e-12, atol=1e-15)) and
        bool(np.allclose(lib_m[:, 2], acc[:, 2], rtol=1e-9, atol=1e-12)))
    mt = torch.from_numpy(np.ascontiguousarray(lib_m)).to(dev)
    mparts = [torch.empty_like(mt) for _ in range(world)]
    if world > 1:
        dist.all_gather(mparts, mt)
        checks["merged_moments_identical_on_all_ranks"] = all(torch.equal(mparts[0], p) for p in mparts)
    else:
        checks["merged_moments_identical_on_all_ranks"] = True
    del sim, alone, x
    # (3) order statistics selected over all ranks (histograms all-reduced over NCCL) == selection on one rank
    n_q = 200_001
    small = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 50.0), P_MG, constant_lags=(TAU,))
    kw = dict(dt=DT, saveat=1.0, trajectories=n_q, seed=seed + 1)
    d = S.solve(S.EnsembleProblem(small), S.EM(), ea, **kw)
    med_all = S.timeseries_steps_quantile(d, [0.05, 0.95], median=True)
    one = S.solve(S.EnsembleProblem(small), S.EM(), S.EnsembleB200(device=local, output="device"), **kw)
    med_one = S.timeseries_steps_quantile(one, [0.05, 0.95], median=True)
    checks["distributed_quantiles_equal_single_rank"] = all_true(bool(np.array_equal(med_all, med_one)))
    srt = torch.sort(one.out[:, 0, :], dim=1).values
    checks["median_equals_sorted_middle"] = all_true(bool(np.array_equal(med_one[:, 0, 2], srt[:, n_q // 2].cpu().numpy())))
    checks["all"] = all(bool(v) for v in checks.values())
    return checks


def pcie_ceiling(torch, dist, local, world, out_dev, out_host, barrier_sync):
    """What plain copy engines reach for the e2e arm's D2H bytes: every rank copies its device-resident paths into its pinned
    host buffer with ONE cudaMemcpyAsync at the same time; aggregate GB/s = all bytes / slowest rank."""
    dev = torch.device(f"cuda:{local}")
    best = None
    for r in range(3):
        barrier_sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out_host.copy_(out_dev, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if r > 0:
            best = float(t[0]) if best is None else min(best, float(t[0]))
    nbytes = out_dev.numel() * 8
    return {"pcie_ceiling_gbs": nbytes * world / (best * 1e-3) / 1e9, "pcie_ceiling_ms": best,
            "pcie_ceiling_per_gpu_gbs": nbytes / (best * 1e-3) / 1e9}


def isolated_launches(S, torch, local, launch, alg_bytes, peak_gbs, reps=5, idle_s=0.4):
    """Kernel time of single launches separated by idle gaps + the SM clock NVML reports while each one runs."""
    import threading
    samples, stop = [], threading.Event()
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)

        def poll():
            while not stop.is_set():
                samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                time.sleep(0.002)
        th = threading.Thread(target=poll, daemon=True)
        th.start()
    except Exception:
        th = None
    ms, mhz = [], []
    for i in range(reps):
        torch.cuda.synchronize()
        time.sleep(idle_s)
        t0 = time.perf_counter()
        sim = launch(i)
        t1 = time.perf_counter()
        ms.append(float(sim.timing["kernel_ms"]))
        inside = [c for (t, c) in samples if t0 + 0.25 * (t1 - t0) <= t <= t1]     # skip the ramp from the idle clock
        if inside:
            mhz.append(sorted(inside)[len(inside) // 2])
    stop.set()
    if th is not None:
        th.join(timeout=1.0)
    best = min(ms)
    med = sorted(ms)[len(ms) // 2]
    return {"kernel_ms": ms, "kernel_ms_median": med, "kernel_ms_min": best, "sm_mhz_during_launch": mhz,
            "achieved_median": alg_bytes / (med * 1e-3) / 1e9, "frac_median": alg_bytes / (med * 1e-3) / 1e9 / peak_gbs,
            "frac_best": alg_bytes / (best * 1e-3) / 1e9 / peak_gbs,
            "how": f"{reps} launches of the bench kernel, each after {idle_s} s of idle (board below its power limit); CUDA-event kernel time"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trajectories", type=int, default=TRAJ_PER_GPU)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the with-moments, summary, configs and checks arms (A/B timing runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        value, info, ms = cpu_reference_run(max(1, args.steps), max(0, args.warmup))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": workload_config(max(1, args.gpus), args.trajectories),
                          "cpu_baseline": info, "gpu_launches": 0,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    import sdde_b200 as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the SDDE path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier_sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_traj = args.trajectories
    n_total = n_traj * world
    prob = mg_problem(S)
    eprob = S.EnsembleProblem(prob)
    sch = S.schedule(prob, S.EM(), dt=DT, saveat=1.0)
    n_steps, n_save = sch.n_steps, len(sch.save_t)
    assert (n_steps, n_save) == (N_STEPS, N_SAVE)
    ctx = S.get_context(local)
    ea_dev = S.EnsembleB200(device=local, distributed=world > 1, output="device")
    out_dev = torch.empty((n_save, 1, n_traj), dtype=torch.float64, device=dev)

    def one_pass(i, ea, out, moments):
        return S.solve(eprob, S.EM(), ea, dt=DT, saveat=1.0, trajectories=n_total, seed=0xB2000001 + i, out=out,
                       moments=moments)

    # ---- device-resident arm: value + roofline ----
    W, K = max(3, args.warmup), max(1, args.steps)
    sampler = ClockSampler(local)
    sampler.start()
    for i in range(W):
        sim = one_pass(i, ea_dev, out_dev, False)
    barrier_sync()
    sampler.mark_start()
    t0 = time.perf_counter()
    kernel_ms, launches = [], 0
    for i in range(K):
        sim = one_pass(W + i, ea_dev, out_dev, False)
        kernel_ms.append(sim.timing["kernel_ms"])
        launches += sim.timing["launches"]
    barrier_sync()
    elapsed = time.perf_counter() - t0
    sampler.mark_end()
    clocks = sampler.stop()
    tmax = torch.tensor([elapsed, sum(kernel_ms) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    elapsed_max, kms_max = float(tmax[0]), float(tmax[1])
    value = n_total * n_steps * K / elapsed_max
    timing = sim.timing

    # one pass with ensemble moments: exercises the path's only collective (NCCL all-gather + Chan merge)
    simm = one_pass(0, ea_dev, out_dev, True)
    mean_end = float(simm.moments[-1, 0, 1])
    # the same passes WITH the ensemble mean/variance reduction inside the timed region (moment kernels over the 4 GB of
    # saved paths + the NCCL all-gather and Chan merge of the triples when N > 1): reported beside `value`
    barrier_sync()
    t0m = time.perf_counter()
    Km = 1 if args.no_extras else max(1, min(K, 5))
    for i in range(Km):
        simm = one_pass(W + i, ea_dev, out_dev, True)
    barrier_sync()
    elm = torch.tensor([time.perf_counter() - t0m], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elm, op=dist.ReduceOp.MAX)
    with_moments = {"value": n_total * n_steps * Km / float(elm[0]), "unit": UNIT, "ms_per_step": float(elm[0]) / Km * 1e3, "steps": Km,
                    "merge": "NCCL all_gather of (count, mean, M2) + Chan merge" if world > 1 else "single GPU: no collective"}

    peaks, peak_src = {}, "fallback (B200_PROFILING.md)"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peaks = {"hbm_gbs": 6650.0}

    # ---- summary arm (extra, not the headline): the ensemble never leaves HBM, the host receives mean, variance, median and
    #      5 % / 95 % bands per saved time point (EnsembleSummary: moment kernels + exact order statistics, selected over all
    #      ranks when the ensemble is sharded) -- the design answer to EnsembleProblem's `reduction` ----
    e2e_summary = None
    if not args.no_extras:
        for i in range(2):
            S.EnsembleSummary(one_pass(i, ea_dev, out_dev, True))
        barrier_sync()
        t0s = time.perf_counter()
        Ks = max(1, min(K, 5))
        for i in range(Ks):
            summ = S.EnsembleSummary(one_pass(W + i, ea_dev, out_dev, True))
        barrier_sync()
        els = torch.tensor([time.perf_counter() - t0s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(els, op=dist.ReduceOp.MAX)
        e2e_summary = {"value": n_total * n_steps * Ks / float(els[0]), "unit": UNIT, "ms_per_step": float(els[0]) / Ks * 1e3, "steps": Ks,
                       "d2h_bytes_per_step": int(summ.u.nbytes + summ.v.nbytes + summ.med.nbytes + summ.qlow.nbytes + summ.qhigh.nbytes),
                       "what": "solve + mean/variance + median + 5%/95% quantile bands of the WHOLE ensemble computed on the device(s), "
                               "only the summary crosses PCIe"}

    configs = checks = None
    if not args.no_extras:
        configs = config_block(S, torch, dist, local, world, barrier_sync, peaks["hbm_gbs"])
        checks = checks_block(S, torch, dist, local, world, rank, eprob, out_dev, n_traj, 0xB2000001)

    # ---- end-to-end arm: host buffers, pinned; per step the library's caches are dropped (schedule replay + table upload =
    #      the step's H2D) and the saved paths land in host memory (D2H) inside the timed region ----
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((n_save, 1, n_traj), dtype=torch.float64, pin_memory=True)
        ea_host = S.EnsembleB200(device=local, distributed=world > 1, output="host")
        detail = []                     # per solve: arm the library chose (auto: staged, zero-copy, then the faster), device times

        def note(tag, sim, wall):
            tm = sim.timing
            detail.append({"step": tag, "arm": "zero-copy" if "mapped host out" in tm["kernel_name"] else "staged",
                           "kernel_ms": round(tm["kernel_ms"], 2), "copy_ms": round(tm["d2h_ms"], 2), "wall_ms": round(wall * 1e3, 2)})
        for i in range(3):
            tw = time.perf_counter()
            simh = one_pass(i, ea_host, out_host, False)
            note("warmup", simh, time.perf_counter() - tw)
        barrier_sync()
        t0 = time.perf_counter()
        Ke = max(1, min(K, 3))
        h2d = d2h = 0
        for i in range(Ke):
            tw = time.perf_counter()
            ctx.set_option("drop_caches", 1)
            simh = one_pass(W + i, ea_host, out_host, False)
            checksum = float(out_host[-1, 0, :8].sum())          # the host reads the step's result
            note("timed", simh, time.perf_counter() - tw)
            h2d += int(simh.timing["h2d_bytes"]); d2h += int(simh.timing["d2h_bytes"])
        barrier_sync()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * n_steps * Ke / float(el[0]), "unit": UNIT,
               "h2d_bytes_per_step": h2d // Ke, "d2h_bytes_per_step": d2h // Ke,
               "ms_per_step": float(el[0]) / Ke * 1e3, "waves": simh.timing["waves"], "steps": Ke, "how": simh.timing["kernel_name"],
               "d2h_gbs_all_ranks": d2h / Ke * world / (float(el[0]) / Ke) / 1e9, "per_solve_rank0": detail}
        one_pass(0, ea_dev, out_dev, False)
        e2e.update(pcie_ceiling(torch, dist, local, world, out_dev, out_host, barrier_sync))
        del out_host
        if not args.no_extras:
            # the same solve when the caller passes NO page-locked buffer (a plain numpy result, fresh pageable pages each time):
            # bounce buffers + host copy threads inside the library; reported beside the page-locked figure above, not instead of it
            barrier_sync()
            tp = time.perf_counter()
            for i in range(2):
                simp = S.solve(eprob, S.EM(), ea_host, dt=DT, saveat=1.0, trajectories=n_total, seed=0xB2000001 + i)
                checksum = float(simp.out[-1, 0, :8].sum())
                del simp
            barrier_sync()
            elp = torch.tensor([time.perf_counter() - tp], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(elp, op=dist.ReduceOp.MAX)
            e2e["pageable_result_ms_per_step"] = float(elp[0]) / 2 * 1e3

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (per launch, this rank) ----
    alg_bytes = ALG_BYTES_PER_TRAJ_STEP * n_traj * n_steps
    k_ms = sum(kernel_ms) / K
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic, prof = None, {}
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        traffic = prof.get("bench_launch_dram_bytes")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "traffic_source": (f"ncu --set full capture of this kernel committed under profiles/ ({prof.get('capture')}, "
                                                       f"commit {prof.get('commit')}); not re-measured in this run") if traffic else None,
                "peak_source": peak_src, "kernel": timing["kernel_name"], "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "regs_per_thread": timing["regs_per_thread"],
                "block_threads": timing["block_threads"]}
    # The same launch timed ALONE: five launches, each after 0.4 s of idle, so the board is not at its power limit and the SM
    # clock is whatever an un-throttled launch gets (polled through NVML every 2 ms while the kernel runs).  Reported beside
    # the sustained figure above, never instead of it: it separates "the kernel stalls" from "the power cap sets the clock".
    if rank == 0 and not args.no_extras:
        try:
            roofline["isolated"] = isolated_launches(S, torch, local, lambda i: one_pass(100 + i, ea_dev, out_dev, False), alg_bytes,
                                                     peaks["hbm_gbs"])
        except Exception as ex:                                  # NVML missing: the timing is still valid, the clock is unknown
            roofline["isolated"] = {"error": repr(ex)}
    # the co-limiting resource: FP64 pipe (measured DFMA peak of this GPU, reported alongside; the micro-benchmark runs
    # AFTER every timed region, only the CPU leg below follows it)
    import ctypes as C
    v = C.c_double(0)
    S._capi.check(S.lib().sdde_b200_dfma_peak(ctx.handle, C.byref(v)))
    roofline["fp64"] = {"measured_dfma_peak_tflops": v.value, "fp64_pipe_pct_ncu": prof.get("fp64_pipe_pct"),
                        "note": "FP64-pipe utilisation comes from the committed ncu capture (profiles/), not from this run"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        # about 10-15 s of CPU work: 15 000 trajectories per core x 10 000 steps in one pass
        _, cpu_baseline, _ = cpu_reference_run(1, 0, sample_paths=(os.cpu_count() or 1) * 15000)
        # the same port with the generator the parity tests use (Philox + libm Box-Muller), on a third of the sample
        v_bm, _, _ = cpu_reference_run(1, 0, sample_paths=(os.cpu_count() or 1) * 5000, rng="philox")
        cpu_baseline["value_with_philox_box_muller"] = v_bm

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n_traj),
            "rng": "in-kernel Philox4x32-10 keyed by (seed, global trajectory id), 4 normals per call (40-bit radius, 24-bit angle per "
                   "Box-Muller pair), float64 Box-Muller from shared-memory tables",
            "ring_slots": timing["ring_slots"], "ensemble_mean_at_tf": mean_end,
            "clocks": clocks, "e2e": e2e, "with_moments": with_moments, "e2e_summary": e2e_summary, "gpu_launches": launches,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "kernel_ms_per_step_max_over_ranks": kms_max,
            "configs": configs, "checks": checks}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
