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
