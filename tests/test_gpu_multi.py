"""-m gpu: one process driving several contexts through sdde_b200_solve_multi gives exactly the single-context result
(paths bit for bit; moments up to the merge order), and the distributed order statistics equal the single-rank ones.
On a 1-GPU box the split / traj_offset / merge legs run with two contexts (or two ranks over gloo) sharing GPU 0; the
NCCL legs need >= 2 GPUs and are skipped otherwise."""
import numpy as np
import pytest

import cases
import sdde_b200 as S

pytestmark = pytest.mark.gpu


def _ngpu():
    return max(0, S.lib().sdde_b200_device_count())


def test_two_contexts_on_one_gpu_equal_one_context():
    """sdde_b200_solve_multi with two contexts that share GPU 0: contiguous global-id ranges, per-context traj_offset,
    repacked supplied noise, sliced per-trajectory parameters, host Chan merge of the moment triples (two ranks on one
    device cannot form an NCCL group, so the library takes its host merge)."""
    from sdde_b200 import _capi as capi
    ca, cb = capi.Context(0), capi.Context(0)
    try:
        case = cases.c2(T=40.0)
        n = 5001
        kw = dict(trajectories=n, seed=77, moments=True, **case.solve_kwargs())
        a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), **kw)
        b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(devices=(ca, cb)), **kw)
        assert np.array_equal(a.out, b.out) and np.array_equal(a.status, b.status)
        assert np.allclose(a.moments, b.moments, rtol=1e-12)
        assert b.timing["moment_merge"] == 1          # host merge
        # three contexts, uneven split (1667 + 1667 + 1667 = 5001), paths only
        cc = capi.Context(0)
        try:
            c3 = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(devices=(ca, cb, cc)), trajectories=n, seed=77,
                         **case.solve_kwargs())
        finally:
            cc.close()
        assert np.array_equal(a.out, c3.out)
        # supplied noise + per-trajectory parameters are sliced per context
        c1 = cases.c1(T=3.0)
        sch = c1.oracle_schedule()
        dW = cases.increments(sch, 1, 1001)
        pb = np.stack([-1.0 + 0.1 * np.arange(1001) / 1001, 0.5 * np.ones(1001), 0.2 * np.ones(1001)])
        kw = dict(trajectories=1001, noise_increments=dW, **c1.solve_kwargs())
        a = S.solve(S.EnsembleProblem(c1.problem(), p_batch=pb), S.EM(), S.EnsembleB200(device=0), **kw)
        b = S.solve(S.EnsembleProblem(c1.problem(), p_batch=pb), S.EM(), S.EnsembleB200(devices=(ca, cb)), **kw)
        assert np.array_equal(a.out, b.out)
        ref, _ = c1.oracle(dW, params=pb)
        assert cases.scaled_err(b.out, ref) < 1e-12
    finally:
        ca.close(); cb.close()


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


def _summary_worker(rank, world, port, q, backend="nccl"):
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
    gpu = rank if backend == "nccl" else 0            # gloo: both ranks share GPU 0 (1-GPU boxes)
    torch.cuda.set_device(gpu)
    if backend == "nccl":
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=torch.device(f"cuda:{gpu}"))
    else:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
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
    d = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=gpu, distributed=True, output="device"), **kw)
    summ = S.EnsembleSummary(d)
    one = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=gpu, output="device"), **kw)
    ref1 = S.EnsembleSummary(one)
    ok_solve = (bool(np.array_equal(summ.med, ref1.med)) and bool(np.array_equal(summ.qlow, ref1.qlow)) and
                bool(np.array_equal(summ.qhigh, ref1.qhigh)) and bool(np.allclose(summ.u, ref1.u, rtol=1e-12)) and summ.num_monte == 20_001)
    q.put((rank, ok_synth, ok_local, ok_solve))
    dist.barrier()
    dist.destroy_process_group()


def _run_two_ranks(backend):
    import os
    import torch.multiprocessing as mp
    port = 29700 + os.getpid() % 200 + (0 if backend == "nccl" else 250)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_summary_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True, True, True), (1, True, True, True)]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_order_statistics_across_two_ranks_nccl():
    """One process per GPU (the torchrun layout): per digit pass each rank histograms its own trajectories, the histograms
    are all-reduced over NCCL, and both ranks end with the exact order statistics of the whole ensemble."""
    _run_two_ranks("nccl")


def test_order_statistics_across_two_ranks_sharing_one_gpu():
    """The same two-rank selection with both ranks on GPU 0 and gloo as the group's backend (histograms reduced through the
    host): shard ranges, per-rank histograms, the agreed bin decisions, the distributed solve + merged moments -- every leg
    but the NCCL transport itself runs on a 1-GPU box."""
    _run_two_ranks("gloo")


def test_two_contexts_large_pageable_output_through_the_bounce_path():
    """sdde_b200_solve_multi into ONE large pageable host array: each context (own bounce buffers, own host copy threads, its own
    host thread inside the library) scatters its column range concurrently; equal to the single-context result."""
    from sdde_b200 import _capi as capi
    ca, cb = capi.Context(0), capi.Context(0)
    try:
        case = cases.c1(T=10.0, saveat=0.1)
        n = 130_001                                           # 101 rows x 65 000 x 8 B = 52 MB per context
        kw = dict(trajectories=n, seed=78, **case.solve_kwargs())
        a = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0, output="device"), **kw)
        b = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(devices=(ca, cb)), **kw)
        assert np.array_equal(a.out.cpu().numpy(), b.out) and np.array_equal(a.status, b.status)
    finally:
        ca.close(); cb.close()
