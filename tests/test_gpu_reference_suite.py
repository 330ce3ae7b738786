"""-m gpu: the reference's own test files, restated for the algorithms on this path.

  test/test_prob_sol.jl                  -> every supported solver runs on the Hayes problem and sol.u[end] != 0
  test/analyticless_convergence_tests.jl -> strong orders (tests/test_gpu_rng.py::test_k6_*)
  test/nondiagonal_sparse_noise.jl       -> the delayed SIR model with a 3x2 sparse noise matrix runs (no assertion
                                            beyond "does not throw" in the reference; here additionally parity)
  test/events.jl                         -> callbacks are out of scope; its model (scalar additive noise) is covered by
                                            tests/test_gpu_parity.py::test_scalar_additive_noise
"""
import numpy as np
import pytest

import cases
import sdde_b200 as S

pytestmark = pytest.mark.gpu

PMUL = [1.0, -4.0, -2.0, 10.0, -1.3, -1.2, 1.1]
PADD = [1.0, -4.0, -2.0, 10.0, -0.0, -0.0, 0.1]


@pytest.mark.parametrize("alg,kw", [(S.EM(), {}), (S.EM(False), {}), (S.EulerHeun(), {}), (S.RKMil(), {"adaptive": False}),
                                    (S.RKMil("Stratonovich"), {"adaptive": False}), (S.RKMilCommute(), {"adaptive": False}),
                                    (S.RKMilCommute("Stratonovich"), {"adaptive": False})])
@pytest.mark.parametrize("p", [PMUL, PADD])
def test_prob_sol_smoke(alg, kw, p):
    # test/test_prob_sol.jl:15-43: h(p,t) = 1 + t, u0 = [1.0], tspan = (0, 0.1), dt = 0.01, constant_lags = (p[1],)
    prob = S.SDDEProblem(S.hayes.f, S.hayes.g, [1.0], lambda pp, t: np.ones(1) + t, (0.0, 0.1), p, constant_lags=(p[0],))
    sol = S.solve(prob, alg, dt=0.01, seed=1234, **kw)              # a single SDDEProblem -> one trajectory, like the reference
    assert sol.retcode == "Success"
    assert len(sol.t) == 11 and sol.t[-1] == 0.1
    assert np.all(np.isfinite(sol.u)) and np.all(sol.u[-1] != 0.0)   # @test sol.u[end] != zeros(1)


@pytest.mark.parametrize("alg", [S.WangLi3SMil(v) for v in "ABCDEF"])
def test_prob_sol_smoke_wangli(alg):
    # test/test_prob_sol.jl:50-61 (multiplicative problem)
    prob = S.SDDEProblem(S.hayes.f, S.hayes.g, [1.0], lambda pp, t: np.ones(1) + t, (0.0, 0.1), PMUL, constant_lags=(PMUL[0],))
    sol = S.solve(prob, alg, dt=0.01, seed=1234)
    assert sol.retcode == "Success" and len(sol.t) == 11 and np.all(np.isfinite(sol.u)) and np.all(sol.u[-1] != 0.0)


@pytest.mark.parametrize("alg", [S.SRA1(), S.SRA3(), S.SOSRA()])
def test_prob_sol_smoke_additive_sra(alg):
    # test/test_prob_sol.jl:72-87 "Additive problems": prob.p .= padd; solve(prob, SRA1(), dt = 0.01) ... (adaptively there,
    # fixed-step here)
    prob = S.SDDEProblem(S.hayes.f, S.hayes.g, [1.0], lambda pp, t: np.ones(1) + t, (0.0, 0.1), PADD, constant_lags=(PADD[0],))
    sol = S.solve(prob, alg, dt=0.01, seed=1234, adaptive=False)
    assert sol.retcode == "Success" and len(sol.t) == 11 and sol.t[-1] == 0.1
    assert np.all(np.isfinite(sol.u)) and np.all(sol.u[-1] != 0.0)
    with pytest.raises(NotImplementedError, match="adaptive"):
        S.solve(prob, alg, dt=0.01)


def test_nondiagonal_sparse_noise_runs():
    # test/nondiagonal_sparse_noise.jl:45-70 (fixed-step EM instead of adaptive LambaEM + callback)
    case = cases.sir(T=20.0, dt=0.02)
    sim = S.solve(S.EnsembleProblem(case.problem()), S.EM(), S.EnsembleB200(device=0), trajectories=256, seed=1234,
                  **case.solve_kwargs())
    assert sim.out.shape == (21, 3, 256) and np.all(np.isfinite(sim.out))
    # S + I + R is conserved by the drift and by the noise columns (each column sums to zero)
    tot = sim.out.sum(axis=1)
    assert np.allclose(tot, 1000.0, rtol=1e-12)


def test_plain_c_program_solves_on_the_gpu(tmp_path):
    """tests/c/abi_smoke.c (C99, only include/sdde_b200.h): schedule + an 8-path C2 solve through the bare C ABI."""
    import os
    import shutil
    import subprocess
    from sdde_b200 import _capi as capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([cc, "-std=c99", "-Wall", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"),
                    "-o", exe, "-L", libdir, "-lsdde_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "solve -> 0, u(0) = 1.2" in out, out
