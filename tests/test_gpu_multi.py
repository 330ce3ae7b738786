"""-m gpu, needs >= 2 GPUs (skipped otherwise): one process driving several GPUs through sdde_b200_solve_multi gives
exactly the single-GPU result (paths bit for bit; moments up to the merge order)."""
import numpy as np
import pytest

import cases
import sdde_b200 as S

pytestmark = pytest.mark.gpu


def _ngpu():
    return max(0, S.lib().sdde_b200_device_count())


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpus_one_process_equals_one_gpu():
    case = cases.c2(T=40.0)
    n = 5001
    kw = dict(trajectories=n, seed=77, moments=True, **case.solve_kwargs())
    a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
    b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(devices=(0, 1)), **kw)
    assert np.array_equal(a.out, b.out) and np.array_equal(a.status, b.status)
    assert np.allclose(a.moments, b.moments, rtol=1e-12)
    assert b.timing["moment_merge"] == 2          # NCCL all-gather over NVLink + device merge (the path's only collective)
    ctx = S.get_context(0)
    ctx.set_option("nccl", 0)
    try:
        c = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(devices=(0, 1)), **kw)
    finally:
        ctx.set_option("nccl", 1)
    assert c.timing["moment_merge"] == 1 and np.allclose(c.moments, b.moments, rtol=1e-13)
    # supplied noise + per-trajectory parameters are sliced per GPU
    c1 = cases.c1(T=3.0)
    sch = c1.oracle_schedule()
    dW = cases.increments(sch, 1, 1001)
    pb = np.stack([-1.0 + 0.1 * np.arange(1001) / 1001, 0.5 * np.ones(1001), 0.2 * np.ones(1001)])
    kw = dict(trajectories=1001, noise_increments=dW, **c1.solve_kwargs())
    a = S.solve(S.EnsembleProblem(c1.problem(), p_batch=pb), S.EM(), S.EnsembleB200(device=0), **kw)
    b = S.solve(S.EnsembleProblem(c1.problem(), p_batch=pb), S.EM(), S.EnsembleB200(devices=(0, 1)), **kw)
    assert np.array_equal(a.out, b.out)
    ref, _ = c1.oracle(dW, params=pb)
    assert cases.scaled_err(b.out, ref) < 1e-12


def _summary_worker(rank, world, port, q):
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "oracle"), os.path.join(root, "tests")]
    import numpy as np
    import torch
    import torch.distributed as dist
    import cases
    import sdde_b200 as S
    import sdde_oracle as O
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    # (1) synthetic rows, identical on both ranks; each rank holds its columns only (uneven shards, ties, a NaN on rank 1)
    n = 50_001
    rng = np.random.default_rng(123)
    x = np.stack([rng.standard_normal(n), 1.2 + 1e-3 * rng.standard_normal(n), np.round(rng.standard_normal(n), 1), np.full(n, 0.5),
                  rng.standard_normal(n)])
    x[4, n - 7] = np.nan
    lo, hi = S.shard_range(n, rank, world)
    shard = torch.from_numpy(np.ascontiguousarray(x[:, None, lo:hi])).cuda()
    sim = S.EnsembleSolution(np.arange(5.0), shard, None, None, False, 0.0, {}, traj_offset=lo, n_total=n)
    got = S.timeseries_steps_quantile(sim, [0.0, 0.05, 0.5, 0.95, 1.0], median=True)[:, 0, :]
    ref = O.ensemble_quantiles(x, [0.0, 0.05, 0.5, 0.95, 1.0], want_median=True)
    ok_synth = bool(np.array_equal(got, ref, equal_nan=True)) and bool(np.all(np.isnan(got[4])))
    loc = S.timeseries_steps_quantile(sim, [0.5], local=True)[:, 0, 0]
    ok_local = bool(np.array_equal(loc[:4], O.ensemble_quantiles(x[:4, lo:hi], [0.5])[:, 0]))
    # (2) a distributed solve: EnsembleSummary over both GPUs == statistics of the same ensemble solved on one GPU
    case = cases.c2(T=20.0)
    kw = dict(trajectories=20_001, seed=5, moments=True, **case.solve_kwargs())
    d = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=rank, distributed=True, output="device"), **kw)
    summ = S.EnsembleSummary(d)
    one = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=rank, output="device"), **kw)
    ref1 = S.EnsembleSummary(one)
    ok_solve = (bool(np.array_equal(summ.med, ref1.med)) and bool(np.array_equal(summ.qlow, ref1.qlow)) and
                bool(np.array_equal(summ.qhigh, ref1.qhigh)) and bool(np.allclose(summ.u, ref1.u, rtol=1e-12)) and summ.num_monte == 20_001)
    q.put((rank, ok_synth, ok_local, ok_solve))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_order_statistics_across_two_ranks_nccl():
    """One process per GPU (the torchrun layout): per digit pass each rank histograms its own trajectories, the histograms
    are all-reduced over NCCL, and both ranks end with the exact order statistics of the whole ensemble."""
    import os
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 200
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_summary_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True, True), (1, True, True, True)]
