#!/usr/bin/env python
"""bench.py -- trajectory-steps/sec of the ensemble SDDE path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's algorithm on the host CPU
    torchrun --nproc-per-node N bench.py --gpus N ...        # one rank per GPU

A "step" is one pass of the hot path over one batch: one whole ensemble solve of BASELINE.json configs[1]
(stochastic Mackey-Glass, tau=17, EM dt=0.05, t in [0,500] = 10 000 schedule steps, saveat 1.0 = 501 saved points,
10^6 trajectories PER GPU, in-kernel Philox increments).  Trajectories shard across ranks with no data-path collective
(weak scaling: per-GPU work fixed); the only collective is the NCCL all-gather of the (count, mean, M2) moment triples.

  value  = whole-job trajectory-steps/s with the result left in HBM (device-resident output, no D2H)
  e2e    = the same through the public API with HOST buffers: descriptor H2D + D2H of the saved paths into pinned
           host memory inside the timed region
  roofline.achieved = 16.4 algorithmic bytes per trajectory-step x trajectory-steps per launch
           / CUDA-event duration of the step kernel (events recorded by the library on its compute stream)

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
WORKLOAD = dict(workload="C2 stochastic Mackey-Glass beta=0.2 gamma=0.1 n=10 sigma=0.1 tau=17 u0=h=1.2, EM(split) dt=0.05, "
                         "t in [0,500] (10000 steps), saveat=1.0 (501 points)", trajectories_per_gpu=1_000_000,
                noise="in-kernel Philox4x32-10 (4 normals per call: 40-bit radius, 24-bit angle per Box-Muller pair) + float64 Box-Muller",
                l2="working set 2.7 GB ring + 4.0 GB output per launch >> 126 MB L2 (no explicit flush needed)")
T_END, DT, TAU = 500.0, 0.05, 17.0
P_MG = (0.2, 0.1, 10.0, 0.1)


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
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


def cpu_reference_run(steps, warmup, sample_paths=None):
    """The reference's per-trajectory algorithm (C++ restatement) on all host cores; returns (value, info)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import c_oracle
    import sdde_oracle as O
    cores = os.cpu_count() or 1
    n = sample_paths or max(64, cores * 1500)
    sv = O.expand_saveat(1.0, 0.0, T_END)[0]
    times, n_steps = [], 0
    for i in range(warmup + steps):
        _, _, el, n_steps = c_oracle.solve(O.MODEL_MACKEY_GLASS, O.ALG_EM_SPLIT, 1, 1, 1, list(P_MG), [1.2], [1.2], [0.0], [TAU],
                                           0.0, T_END, DT, sv, 501, n_traj=n, seed=0xB2000001 + i, n_threads=cores)
        if i >= warmup:
            times.append(el)
    total = sum(times)
    value = n * n_steps * len(times) / total
    info = {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} of the 10^6 trajectories x {n_steps} steps per pass, {len(times)} passes, C++ restatement of the "
                      f"reference's per-trajectory loop (oracle/sdde_oracle.cpp), std::thread static chunks over {cores} cores; "
                      "not the Julia reference (Julia is not in the image)"}
    return value, info, total / len(times) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--trajectories", type=int, default=WORKLOAD["trajectories_per_gpu"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the with-moments and summary arms (A/B timing runs)")
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
                          "config": dict(WORKLOAD, trajectories_per_gpu=None, sample=info["sample"]),
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

    # ---- summary arm (extra, not the headline): the ensemble never leaves HBM, the host receives mean, variance, median and
    #      5 % / 95 % bands per saved time point (EnsembleSummary: moment kernels + exact order statistics on the device) ----
    e2e_summary = None
    if world == 1 and not args.no_extras:
        for i in range(2):
            S.EnsembleSummary(one_pass(i, ea_dev, out_dev, True))
        barrier_sync()
        t0s = time.perf_counter()
        Ks = max(1, min(K, 5))
        for i in range(Ks):
            summ = S.EnsembleSummary(one_pass(W + i, ea_dev, out_dev, True))
        barrier_sync()
        els = time.perf_counter() - t0s
        e2e_summary = {"value": n_total * n_steps * Ks / els, "unit": UNIT, "ms_per_step": els / Ks * 1e3, "steps": Ks,
                       "d2h_bytes_per_step": int(summ.u.nbytes + summ.v.nbytes + summ.med.nbytes + summ.qlow.nbytes + summ.qhigh.nbytes),
                       "what": "solve + mean/variance + median + 5%/95% quantile bands computed on the device, only the summary crosses PCIe"}

    # ---- end-to-end arm: host buffers, pinned, D2H inside the timed region ----
    e2e = None
    if not args.no_e2e:
        out_host = torch.empty((n_save, 1, n_traj), dtype=torch.float64, pin_memory=True)
        ea_host = S.EnsembleB200(device=local, distributed=world > 1, output="host")
        for i in range(2):
            simh = one_pass(i, ea_host, out_host, False)
        barrier_sync()
        t0 = time.perf_counter()
        Ke = max(1, min(K, 3))
        for i in range(Ke):
            simh = one_pass(W + i, ea_host, out_host, False)
            checksum = float(out_host[-1, 0, :8].sum())          # the host reads the step's result
        barrier_sync()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        e2e = {"value": n_total * n_steps * Ke / float(el[0]), "unit": UNIT,
               "h2d_bytes_per_step": int(simh.timing["h2d_bytes"]), "d2h_bytes_per_step": int(simh.timing["d2h_bytes"]),
               "ms_per_step": float(el[0]) / Ke * 1e3, "waves": simh.timing["waves"], "steps": Ke}
        del out_host

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (per launch, this rank) ----
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peaks = {"hbm_gbs": 6650.0}
    alg_bytes = ALG_BYTES_PER_TRAJ_STEP * n_traj * n_steps
    k_ms = sum(kernel_ms) / K
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))
        traffic = prof.get("bench_launch_dram_bytes")
    except Exception:
        prof = {}
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic, "peak_source": peak_src, "kernel": timing["kernel_name"], "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "regs_per_thread": timing["regs_per_thread"]}
    # the co-limiting resource: FP64 pipe (measured DFMA peak of this GPU, reported alongside)
    import ctypes as C
    v = C.c_double(0)
    ctx = S.get_context(local)
    S._capi.check(S.lib().sdde_b200_dfma_peak(ctx.handle, C.byref(v)))
    roofline["fp64"] = {"measured_dfma_peak_tflops": v.value, "fp64_pipe_pct_ncu": prof.get("fp64_pipe_pct"),
                        "note": "FP64-pipe utilisation comes from the committed ncu capture (profiles/), not from this run"}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        # about 10-15 s of CPU work: 15 000 trajectories per core x 10 000 steps in one pass (plus a short warm-up pass)
        _, cpu_baseline, _ = cpu_reference_run(1, 0, sample_paths=(os.cpu_count() or 1) * 15000)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": elapsed_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": dict(WORKLOAD, trajectories_per_gpu=n_traj, trajectories_total=n_total, n_steps=n_steps, n_save=n_save,
                           parallelism=f"trajectory-shard x{world}", ring_slots=timing["ring_slots"],
                           ensemble_mean_at_tf=mean_end),
            "clocks": clocks, "e2e": e2e, "with_moments": with_moments, "e2e_summary": e2e_summary, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernel_ms_per_step_max_over_ranks": kms_max}
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
