"""CPU, world_size 2 over gloo: the only collective of the path -- the (count, mean, M2) all-gather + Chan merge that
runs over NCCL on the GPUs -- and the rank -> trajectory-range partition."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    import torch.distributed as dist
    import sdde_b200 as S
    import sdde_oracle as O
    from sdde_b200.ensemble import _allgather_merge_moments
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(42)
    x = rng.standard_normal((5, 2, 1001)) * 3 + 10          # the "whole ensemble", identical on both ranks
    lo, hi = S.shard_range(1001, rank, world)
    local = O.moments(x[..., lo:hi])
    merged = _allgather_merge_moments(local, 0)
    from sdde_b200.ensemble import _choose_seed
    seeds = (_choose_seed(0, 0, True), _choose_seed(0, 0, True), _choose_seed(7, 0, True), _choose_seed(0, 9, True))
    q.put((rank, lo, hi, merged, O.moments(x), seeds))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_moment_merge_gloo():
    port = 29600 + os.getpid() % 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 501, 501, 1001)
    # a drawn seed is agreed on across ranks (and differs from draw to draw); explicit seeds pass through
    assert res[0][5] == res[1][5] and res[0][5][0] != res[0][5][1] and res[0][5][2:] == (7, 9)
    for _, _, _, merged, full, _ in res:
        assert np.allclose(merged[..., 0], 1001)
        assert np.allclose(merged, full, rtol=1e-12)
    assert np.array_equal(res[0][3], res[1][3])       # every rank ends with identical moments
