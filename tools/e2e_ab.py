"""Where is the end-to-end ceiling when N ranks send their saved paths to ONE host?  (e2e efficiency 0.24 at 8 GPUs.)
Run under torchrun (one rank per GPU):  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_ab.py [n_traj]
Per rank: the C2 bench output (501 x n_traj doubles = 4 GB at 10^6 paths).  Measures, all ranks at once (max over ranks):
  (a) plain cudaMemcpyAsync D2H of the device-resident paths into pinned host memory (copy engines)      = the ceiling
  (b) the same into pinned memory allocated write-combined / portable (SDDE_B200_PINNED_FLAGS)
  (c) the library's end-to-end solve, zero_copy = 1 (the step kernel's own stores cross PCIe)
  (d) the library's end-to-end solve, zero_copy = 0 (staged in HBM, cudaMemcpy2DAsync per wave overlapped with the next wave)
  (e) ranks ONE AT A TIME with plain copies (what a single GPU gets when the others are idle)
and prints the host topology the numbers depend on (numactl --hardware, nvidia-smi topo -m, CPU count)."""
import ctypes as C, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import sdde_b200 as S

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n_traj = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
torch.cuda.set_device(local); dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def maxr(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

prob = S.SDDEProblem(S.mackey_glass.f, S.mackey_glass.g, 1.2, 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1), constant_lags=(17.0,))
ep = S.EnsembleProblem(prob)
out_dev = torch.empty((501, 1, n_traj), dtype=torch.float64, device=dev)
nbytes = out_dev.numel() * 8
res = {"n_gpus": world, "bytes_per_rank": nbytes}
S.solve(ep, S.EM(), S.EnsembleB200(device=local, output="device"), dt=0.05, saveat=1.0, trajectories=n_traj, seed=1, out=out_dev)

def copy_ms(host, reps=4, only=None):
    best = 1e30
    for r in range(reps):
        sync()
        t0 = time.perf_counter()
        if only is None or only == rank:
            host.copy_(out_dev, non_blocking=True); torch.cuda.synchronize()
        el = maxr(time.perf_counter() - t0)
        if r: best = min(best, el)
    return best * 1e3

host = torch.empty((501, 1, n_traj), dtype=torch.float64, pin_memory=True)
ms = copy_ms(host); res["a_plain_copy_all_ranks"] = {"ms": ms, "gbs_total": nbytes * world / ms / 1e6, "gbs_per_gpu": nbytes / ms / 1e6}
if world > 1:
    per = [copy_ms(host, reps=3, only=r) for r in range(min(world, 2))]
    res["e_plain_copy_one_rank_at_a_time_gbs"] = [nbytes / m / 1e6 for m in per]
for flags, name in ((4, "b_write_combined"), (1, "b_portable")):
    os.environ["SDDE_B200_PINNED_FLAGS"] = str(flags)          # read by sdde_b200_alloc_pinned at every call
    ptr = S.lib().sdde_b200_alloc_pinned(nbytes)
    os.environ.pop("SDDE_B200_PINNED_FLAGS")
    if ptr:
        buf = (C.c_char * nbytes).from_address(ptr)
        arr = torch.frombuffer(buf, dtype=torch.float64).reshape(501, 1, n_traj)
        ms = copy_ms(arr); res[name] = {"ms": ms, "gbs_total": nbytes * world / ms / 1e6}
        del arr, buf
        S.lib().sdde_b200_free_pinned(ptr)

def e2e(zero_copy, reps=3, **opts):
    ctx = S.get_context(local); ctx.set_option("zero_copy", zero_copy)
    for k, v in opts.items(): ctx.set_option(k, v)
    ea = S.EnsembleB200(device=local, output="host")
    best = 1e30
    for r in range(reps + 1):
        sync(); t0 = time.perf_counter()
        sim = S.solve(ep, S.EM(), ea, dt=0.05, saveat=1.0, trajectories=n_traj, seed=2 + r, out=host)
        float(host[-1, 0, :8].sum())
        el = maxr(time.perf_counter() - t0)
        if r: best = min(best, el)
    ctx.set_option("zero_copy", -1)
    return {"ms": best * 1e3, "gbs_total": nbytes * world / best / 1e9, "traj_steps_per_s": n_traj * world * 1e4 / best, "how": sim.timing["kernel_name"][-24:], "waves": sim.timing["waves"]}

res["c_library_zero_copy"] = e2e(1)
res["d_library_staged"] = e2e(0)
res["d_library_staged_ramp_2e17"] = e2e(0, ramp_start=1 << 17)
S.get_context(local).set_option("ramp_start", 1 << 16)
if rank == 0:
    def sh(cmd):
        try: return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
        except Exception as e: return f"unavailable: {e}"
    res["host"] = {"cpus": os.cpu_count(), "numactl": sh("numactl --hardware 2>&1 | head -12"), "lscpu": sh("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)' "),
                   "topo": sh("nvidia-smi topo -m 2>&1 | head -14"), "mem": sh("free -g | head -2")}
    print(json.dumps(res, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
