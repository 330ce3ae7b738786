"""BASELINE.json configs[4]: 10^8-trajectory scalar linear SDDE, moments only, trajectories sharded over the ranks of a
torchrun job (strong scaling: the 10^8 paths are split), ensemble mean/variance merged with one NCCL all-gather.

    python tools/c5_sweep.py                         # 1 GPU
    torchrun --nproc-per-node N tools/c5_sweep.py    # N GPUs
"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import torch.distributed as dist
import sdde_b200 as S

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
N = int(float(os.environ.get("C5_TRAJ", "1e8")))
prob = S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0.0, 10.0), (-1.0, 0.5, 0.2), constant_lags=(1.0,))
ea = S.EnsembleB200(device=local, distributed=world > 1)
times = []
for r in range(3):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    sim = S.solve(S.EnsembleProblem(prob), S.EM(), ea, dt=0.01, saveat=1.0, trajectories=N, seed=0xB2000001, moments=True, paths=False)
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], device=f"cuda:{local}", dtype=torch.float64)
    if world > 1: dist.all_reduce(el, op=dist.ReduceOp.MAX)
    times.append(float(el[0]))
if rank == 0:
    mean, var = S.timeseries_steps_meanvar(sim)
    steps = sim.timing["n_steps"]
    print(json.dumps({"config": "C5", "n_gpus": world, "trajectories": N, "n_steps": steps, "seconds": min(times[1:]),
                      "trajectory_steps_per_s": N * steps / min(times[1:]), "count": float(sim.moments[-1, 0, 0]),
                      "mean_tf": float(mean[-1, 0]), "var_tf": float(var[-1, 0])}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
