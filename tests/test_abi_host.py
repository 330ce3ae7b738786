"""CPU: the C-ABI library loads and exports every symbol of include/sdde_b200.h; the host-side schedule builder of the
product equals the oracle's replay bit for bit; descriptor validation and the reference's error behaviour; host logic
of the SciML-surface mirror.  No compute entry point is exercised here (no GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
import sdde_b200 as S
import sdde_oracle as O
from sdde_b200 import _capi as capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "sdde_b200.h")).read()
    declared = sorted(set(re.findall(r"SDDE_API[^;(]*?\b(sdde_b200_\w+)\s*\(", hdr)))
    assert declared == sorted(capi.EXPORTS) and len(declared) == 27
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.sdde_b200_version() == 100


def test_rng_table_is_correctly_rounded():
    """The generator's table is computed on the host in long double (csrc/rng_table.cpp); every entry must be the float64
    nearest to the 200-bit mpmath value, up to the rare double-rounding case (<= 0.51 ulp).  Layout: sdde_rng_tables.cuh."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    tab = capi.rng_table()
    assert tab.shape == (512 + 512 + 128 + 128, 2)
    half_pi = mp.pi / 2

    def ulps(x, exact):
        if exact == 0:
            return 0.0 if x == 0.0 else np.inf
        return float(abs(mp.mpf(float(x)) - exact) / mp.mpf(float(np.spacing(abs(float(exact))))))

    worst = 0.0
    for k in range(512):
        c = mp.mpf(2) if k == 511 else 1 + (mp.mpf(k) + mp.mpf("0.5")) / 512
        inv = float(1 / c)
        assert tab[k, 0] == inv                                   # the reciprocal itself is correctly rounded
        m2ln = 2 * mp.log(mp.mpf(inv)) + (2 * mp.log(2) if k >= 256 else 0)
        worst = max(worst, ulps(tab[k, 1], m2ln))
    assert tab[511, 1] == 0.0
    for i in range(512):
        h, a = i >> 8, i & 255
        alpha = half_pi * (h + (mp.mpf(a) + mp.mpf("0.5")) / 2 ** 8 - mp.mpf("0.5"))
        worst = max(worst, ulps(tab[512 + i, 0], mp.cos(alpha)), ulps(tab[512 + i, 1], mp.sin(alpha)))
    for b in range(128):
        beta = half_pi * ((mp.mpf(b) + mp.mpf("0.5")) / 2 ** 15 - mp.mpf(1) / 2 ** 9)
        gamma = half_pi * ((mp.mpf(b) + mp.mpf("0.5")) / 2 ** 22 - mp.mpf(1) / 2 ** 16)
        worst = max(worst, ulps(tab[1024 + b, 0], mp.cos(beta)), ulps(tab[1024 + b, 1], mp.sin(beta)),
                    ulps(tab[1152 + b, 0], mp.cos(gamma)), ulps(tab[1152 + b, 1], mp.sin(gamma)))
    assert worst <= 0.51, worst
    # the angle decomposition covers the oracle's angles: theta = (pi/2)(q + (f + 1/2) 2^-22 - 1/2) = alpha + beta + gamma + s pi,
    # evaluated with the kernel's operation order (two angle additions)
    rng = np.random.default_rng(5)
    for wb in rng.integers(0, 2 ** 32, size=300, dtype=np.uint64):
        wb = int(wb)
        q, f = wb >> 30, (wb >> 8) & 0x3FFFFF
        theta = half_pi * (q + (mp.mpf(f) + mp.mpf("0.5")) / 2 ** 22 - mp.mpf("0.5"))
        ca, sa = tab[512 + ((wb >> 22) & 0x1FF)]
        cb, sb = tab[1024 + ((wb >> 15) & 0x7F)]
        cc, sc = tab[1152 + ((wb >> 8) & 0x7F)]
        cbc, sbc = cb * cc - sb * sc, sb * cc + cb * sc
        sgn = -1.0 if wb >> 31 else 1.0
        assert abs(sgn * (ca * cbc - sa * sbc) - float(mp.cos(theta))) < 5e-16
        assert abs(sgn * (sa * cbc + ca * sbc) - float(mp.sin(theta))) < 5e-16


def test_struct_layout_matches_header():
    src = '#include "sdde_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",' \
          'sizeof(sdde_desc),sizeof(sdde_schedule),sizeof(sdde_timing),offsetof(sdde_desc,params),offsetof(sdde_desc,seed));}'
    import subprocess, tempfile
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "a.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "a.c"), "-o", os.path.join(td, "a")], check=True)
        out = subprocess.run([os.path.join(td, "a")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [C.sizeof(capi.SddeDesc), C.sizeof(capi.SddeSchedule), C.sizeof(capi.SddeTiming),
                                     capi.SddeDesc.params.offset, capi.SddeDesc.seed.offset]


SCHED_CASES = [
    (lambda: cases.c1(), {}), (lambda: cases.c1(saveat=None), {}), (lambda: cases.c2(), {}), (lambda: cases.c2(T=500.0), {}),
    (lambda: cases.c3(0), {}), (lambda: cases.c3(2), {}), (lambda: cases.c3(0, alg=0), {}), (lambda: cases.c4(), {}),
    (lambda: cases.hayes(T=10.0, dt=2.0 ** -8), {}), (lambda: cases.sir(), {}), (lambda: cases.scalar_add(), {}),
    (lambda: cases.c1(T=3.3, saveat=[0.0, 0.7, 1.0, 2.2, 3.3]), {}), (lambda: cases.c1(T=3.3, saveat=[0.5, 1.5]), {}),
    (lambda: cases.c1(T=0.5, saveat=0.2), {}),                       # tf < tau: never reads its own past
    (lambda: cases.c1(T=2.0, saveat=0.003), {}),                     # several saves per step
    (lambda: cases.c1(T=5.0).with_(dt=1.7, lags=[1.0]), {}),         # dt > tau (allowed for EM)
    (lambda: cases.c2(T=60.0, alg=4), {}), (lambda: cases.c4(T=8.0).with_(alg=4), {}),      # EulerHeun: two lookup sets
    (lambda: cases.c1(T=5.0).with_(dt=0.25, lags=[0.25], alg=4), {}),                        # dt == tau: stage 2 lands on t_n
]


@pytest.mark.parametrize("mk,kw", SCHED_CASES)
def test_product_schedule_equals_oracle_bitwise(mk, kw):
    c = mk()
    s = S.schedule(c.problem(), cases.ALG_MAP[c.alg], **c.solve_kwargs())
    o = c.oracle_schedule()
    for k in ["t", "dt", "sqdt", "lk_mode", "lk_idx", "lk_theta", "save_t", "save_step", "save_theta", "save_exact"]:
        a, b = getattr(s, k), getattr(o, k)
        assert a.shape == b.shape and np.array_equal(a, b), k
    assert list(s.tstops_hit) == o.tstops_hit
    # ring capacity: no node that a later lookup needs is overwritten
    for n in range(s.n_steps):
        for l in range(s.lk_mode.shape[1]):
            if s.lk_mode[n, l] != 0:
                lo = s.lk_idx[n, l] - (1 if s.lk_mode[n, l] == 2 else 0)
                assert n - lo < s.ring_slots and n - s.lk_idx[n, l] >= s.min_gap


def test_user_tstops_and_quirk():
    c = cases.c1(T=3.0)
    s = S.schedule(c.problem(), S.EM(), dt=0.01, saveat=0.5, tstops=[0.333, 2.0, 7.0])
    o = O.build_schedule(0.0, 3.0, 0.01, [1.0], 0, saveat_times=O.expand_saveat(0.5, 0, 3.0)[0], tstops=[0.333, 2.0, 7.0])
    assert np.array_equal(s.t, o.t) and list(s.tstops_hit) == o.tstops_hit and 0.333 in s.t


def test_saveat_expansion_rules():
    sv, ss, se = S.expand_saveat(1.0, 0.0, 500.0)
    assert len(sv) == 500 and sv[0] == 1.0 and sv[-1] == 500.0 and ss and se
    sv, _, _ = S.expand_saveat(0.3, 0.0, 1.0)          # (t0:s:tf)[end] != tf -> stops at tf - s
    assert np.allclose(sv, [0.3, 0.6]) and sv[-1] <= 1.0 - 0.3
    sv, ss, se = S.expand_saveat([0.0, 0.5, 2.0, 9.0], 0.0, 2.0)
    assert list(sv) == [0.5] and ss and se
    sv, ss, se = S.expand_saveat([0.5], 0.0, 2.0)
    assert list(sv) == [0.5] and not ss and not se
    for case in [(0.1, 0.0, 10.0), (0.5, 0.0, 20.0), (1.0, 0.0, 100.0)]:
        assert np.array_equal(S.expand_saveat(*case)[0], O.expand_saveat(*case)[0])


def _desc_rc(prob, alg, **kw):
    try:
        S.schedule(prob, alg, **kw)
        return 0
    except S.SddeError as e:
        return e.code


def test_error_behaviour_mirrors_reference_guards():
    c = cases.c1()
    # src/solve.jl:121-123
    with pytest.raises(S.SddeError, match="Fixed timestep methods require a choice of dt"):
        S.schedule(c.problem(), S.EM(), dt=0.0)
    # RKMil is adaptive by default in the reference (src/solve.jl:40): this path needs adaptive=False
    with pytest.raises(NotImplementedError, match="adaptive"):
        S.schedule(c.problem(), S.RKMil(), dt=0.01)
    # src/solve.jl:581-585
    with pytest.raises(S.SddeError, match=r"dt > minimum\(constant_lags\)"):
        S.schedule(c.problem(), S.RKMil(), dt=1.5, adaptive=False)
    assert _desc_rc(c.problem(), S.EM(), dt=1.5) == 0       # EM is exempt
    with pytest.raises(S.SddeError):
        S.schedule(c.with_(tspan=(1.0, 0.0)).problem(), S.EM(), dt=0.1)
    with pytest.raises(NotImplementedError):
        S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0, 1), (1, 1, 1), constant_lags=(1.0,), neutral=True)
    with pytest.raises(TypeError):
        S.SDDEProblem(lambda u, h, p, t: u, S.linear1.g, 1.0, 1.0, (0, 1), (1, 1, 1), constant_lags=(1.0,))
    with pytest.raises(ValueError):
        S.SDDEProblem(S.linear1.f, S.hayes.g, 1.0, 1.0, (0, 1), (1, 1, 1), constant_lags=(1.0,))


def test_d_discontinuities_keyword_schedule_equals_oracle():
    """d_discontinuities (src/solve.jl:20-23): user-declared discontinuities become tstops (:643-646), enter the
    discontinuity heap when their order is <= alg_order + 1 (:690-699) and propagate by the constant lags like the one
    at t0 (src/integrators/utils.jl:42-118).  Product schedule == oracle replay, bit for bit; hand-checked facts below."""
    c = cases.c1(T=6.0, saveat=0.5)
    dd = [(0.37, 0), (2.2137, 1), 4.05, (4.4137, 3)]         # orders 0, 1, (bare time = 0), 3 (> alg_order + 1: a tstop only)
    for alg, algid in ((S.EM(), 0), (S.RKMil(), 2)):
        kw = c.with_(alg=algid).solve_kwargs()
        s = S.schedule(c.problem(), alg, d_discontinuities=dd, **kw)
        sv, ss, se = O.expand_saveat(c.saveat, *c.tspan)
        o = O.build_schedule(0.0, 6.0, c.dt, c.lags, algid, saveat_times=sv, save_start=ss, save_end=se, d_discontinuities=dd)
        for k in ["t", "dt", "sqdt", "lk_mode", "lk_idx", "lk_theta", "save_t", "save_step", "save_theta", "save_exact"]:
            assert np.array_equal(np.asarray(getattr(s, k)), np.asarray(getattr(o, k))), (algid, k)
        t = np.asarray(s.t)
        for x in (0.37, 2.2137, 4.05, 4.4137):                # every declared time is hit exactly
            assert np.any(t == x), x
        assert np.any(t == 0.37 + 1.0)                        # order 0 at 0.37 propagates by the lag (order 1/2 <= alg_order + 1/2)
        # order 1 at 2.2137: next order 3/2 -- tracked by RKMil (alg_order 1: 3/2 <= 3/2), not by EM (alg_order 1/2: 3/2 > 1)
        assert bool(np.any(t == 2.2137 + 1.0)) == (algid == 2)
        # (4.4137, order 3) is a tstop only: the steps after it restart from 4.4137, so 5.4137 is reached by plain stepping
        # either way -- what shows the difference is the tstop list
        assert bool(np.any(np.asarray(s.tstops_hit) == 2.2137 + 1.0)) == (algid == 2)
    base = S.schedule(c.problem(), S.EM(), **c.solve_kwargs())
    assert len(base.t) < len(s.t)
    with pytest.raises(ValueError):
        S.schedule(c.problem(), S.EM(), d_discontinuities=[(1.0, 0.25)], **c.solve_kwargs())


def test_julia_float_ranges_known_answers():
    """[EXT] Base.(:)(start, step, stop) for Float64 (twiceprecision.jl), as restated for saveat expansion
    (src/solve.jl:661-671).  Known answers of Julia itself: 0.1:0.1:0.5 and 0:0.1:1 hit the decimal values exactly (the
    rational lift), 1/3-steps land on correctly rounded thirds, a non-rational step falls back to start + k*step."""
    from sdde_b200.ensemble import _julia_range, _rat
    assert _rat(0.1) == (1, 10) and _rat(0.05) == (1, 20) and _rat(1 / 3) == (1, 3) and _rat(17.02) == (851, 50)
    assert _julia_range(0.1, 0.1, 0.5).tolist() == [0.1, 0.2, 0.3, 0.4, 0.5]
    assert _julia_range(0.0, 0.1, 1.0).tolist() == [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0]
    assert _julia_range(1 / 3, 1 / 3, 2.0).tolist() == [1 / 3, 2 / 3, 1.0, 4 / 3, 5 / 3, 2.0]
    assert _julia_range(0.0, 0.3, 1.0).tolist() == [0.0, 0.3, 0.6, 0.9]
    assert len(_julia_range(0.0, 1.0, 500.0)) == 501 and _julia_range(1.0, 1.0, 0.5).size == 0
    r = _julia_range(0.0, np.pi / 7, 3.0)                    # literal branch: len = round(3/(pi/7)) + 1 = 8 -> overshoots 3.0 -> 7
    assert r.size == 7 and abs(r[-1] - 6 * (np.pi / 7)) <= 5e-16 and r[-1] <= 3.0
    for (a, s_, b) in [(0.0, 0.1, 1.0), (0.1, 0.1, 0.5), (1 / 3, 1 / 3, 2.0), (0.0, np.pi / 7, 3.0), (0.25, 0.07, 9.3), (2.0, 1e-3, 2.5)]:
        assert np.array_equal(_julia_range(a, s_, b), O._float_range(a, s_, b))      # product == oracle restatement
    # saveat expansion built from them (number form): (t0+s):s:tf when (t0:s:tf)[end] == tf, else up to tf - s
    assert S.expand_saveat(0.3, 0.0, 1.0)[0].tolist() == [0.3, 0.6]
    assert S.expand_saveat(0.25, 0.0, 1.0)[0].tolist() == [0.25, 0.5, 0.75, 1.0]
    assert S.expand_saveat(1 / 3, 0.0, 1.0)[0].tolist() == [1 / 3, 2 / 3, 1.0]


def test_descriptor_mirrors_have_the_header_size():
    """The product's ctypes mirror, the oracle's own copy of it and the C header agree on sizeof(sdde_desc) and on the
    offsets of the fields added last (a stale mirror reads garbage past its end)."""
    import subprocess, tempfile
    import c_oracle
    src = ('#include "sdde_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n",sizeof(sdde_desc),'
           'offsetof(sdde_desc,noise_t),offsetof(sdde_desc,d_disc_t),offsetof(sdde_desc,n_d_disc),offsetof(sdde_desc,out_rows));}')
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "a.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "a.c"), "-o", os.path.join(td, "a")], check=True)
        want = [int(x) for x in subprocess.run([os.path.join(td, "a")], capture_output=True, text=True, check=True).stdout.split()]
    for D in (capi.SddeDesc, c_oracle.Desc):
        assert [C.sizeof(D), D.noise_t.offset, D.d_disc_t.offset, D.n_d_disc.offset, D.out_rows.offset] == want, D


def _c_struct_fields(hdr, name):
    """[(field, ctype, count)] of `typedef struct name {...} name;` in the header (pointers -> 'ptr')."""
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    consts = {k: int(v) for k, v in re.findall(r"#define (SDDE_MAX_\w+) (\d+)", hdr)}
    out = []
    for decl in filter(None, (d.strip() for d in body.split(";"))):
        m = re.match(r"(?:const\s+)?(\w+)\s+(.*)", decl)
        ctype, rest = m.group(1), m.group(2)
        for item in (x.strip() for x in rest.split(",")):
            ptr = item.startswith("*")
            item = item.lstrip("* ")
            am = re.match(r"(\w+)\[(\w+)\]", item)
            if am:
                n = am.group(2)
                out.append((am.group(1), ctype, consts[n] if n in consts else int(n)))
            else:
                out.append((item, "ptr" if ptr else ctype, 1))
    return out


def _julia_struct_fields(src, name):
    body = re.search(r"mutable struct %s\n(.*?)\nend" % name, src, re.S).group(1)
    out = []
    for ln in body.splitlines():
        m = re.match(r"\s*(\w+)::([^=]+?)\s*=", ln)
        if m:
            out.append((m.group(1), m.group(2).strip()))
    return out


def test_julia_struct_mirrors_match_header():
    """julia/EnsembleB200.jl cannot run here (no Julia in the image); its hand-written mirrors of sdde_desc, sdde_schedule
    and sdde_timing are therefore checked textually: same field names in the same order, Julia types of the same size and
    kind as the C types, same total size as gcc's sizeof -- and every ccall names an exported symbol with as many argument
    types as the header's prototype has parameters."""
    import subprocess, tempfile
    hdr = open(os.path.join(ROOT, "include", "sdde_b200.h")).read()
    jl = open(os.path.join(ROOT, "julia", "EnsembleB200.jl")).read()
    csize = {"int32_t": 4, "int64_t": 8, "uint64_t": 8, "double": 8, "ptr": 8, "char": 1}
    jkind = {"Int32": ("int32_t", 4), "Int64": ("int64_t", 8), "UInt64": ("uint64_t", 8), "Float64": ("double", 8), "UInt8": ("char", 1)}
    sizes = {}
    for cname, jname in (("sdde_desc", "SddeDesc"), ("sdde_schedule", "SddeSchedule"), ("sdde_timing", "SddeTiming")):
        cf, jf = _c_struct_fields(hdr, cname), _julia_struct_fields(jl, jname)
        assert [f[0] for f in cf] == [f[0] for f in jf], (cname, [f[0] for f in cf], [f[0] for f in jf])
        off = 0
        for (fname, ctype, count), (_, jt) in zip(cf, jf):
            if ctype == "ptr":
                assert jt.startswith("Ptr{"), (cname, fname, jt)
                sz, n = 8, 1
            elif count > 1:
                m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", jt)
                assert m and int(m.group(1)) == count and jkind[m.group(2)] == (ctype, csize[ctype]), (cname, fname, jt)
                sz, n = csize[ctype], count
            else:
                assert jkind[jt] == (ctype, csize[ctype]), (cname, fname, jt)
                sz, n = csize[ctype], 1
            off = (off + sz - 1) // sz * sz + sz * n              # natural alignment, as both compilers lay it out
        sizes[cname] = (off + 7) // 8 * 8
    src = '#include "sdde_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n",sizeof(sdde_desc),sizeof(sdde_schedule),sizeof(sdde_timing));}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "a.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "a.c"), "-o", os.path.join(td, "a")], check=True)
        want = [int(x) for x in subprocess.run([os.path.join(td, "a")], capture_output=True, text=True, check=True).stdout.split()]
    assert [sizes["sdde_desc"], sizes["sdde_schedule"], sizes["sdde_timing"]] == want
    # every ccall: exported symbol, argument-type count == parameter count of the header's prototype
    protos = {}
    for name, params in re.findall(r"SDDE_API[^;(]*?\b(sdde_b200_\w+)\s*\(([^;]*?)\)\s*;", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S), re.S):
        params = params.strip()
        protos[name] = 0 if params in ("", "void") else params.count(",") + 1
    calls = re.findall(r"ccall\(\(:(sdde_b200_\w+), libsdde\),\s*[\w{}.]+,\s*\(([^()]*(?:\{[^{}]*\}[^()]*)*)\)", jl)
    assert len(calls) >= 10
    for name, argt in calls:
        assert name in protos and name in capi.EXPORTS, name
        n = len([a for a in argt.split(",") if a.strip()])
        assert n == protos[name], (name, n, protos[name])


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = capi.lib()
    assert L.sdde_b200_device_count() < 0
    assert not L.sdde_b200_create(0)
    assert b"no CPU fallback" in L.sdde_b200_last_error()
    d = capi.SddeDesc()
    assert L.sdde_b200_solve(None, C.byref(d), None, None, None, None) <= capi.lib().sdde_b200_version() * 0 - 100
    with pytest.raises(S.SddeError):
        S.solve(cases.c1().problem(), S.EM(), dt=0.01)


def test_abi_rejects_bad_descriptors():
    L = capi.lib()
    sch = capi.SddeSchedule()
    d = capi.SddeDesc()
    assert L.sdde_b200_schedule(C.byref(d), C.byref(sch)) == -1 and b"struct_size" in L.sdde_b200_last_error()
    d.struct_size = C.sizeof(capi.SddeDesc)
    d.alg = 99
    assert L.sdde_b200_schedule(C.byref(d), C.byref(sch)) == -2
    assert L.sdde_b200_schedule(None, C.byref(sch)) == -1


def test_history_function_forms():
    from sdde_b200.problem import _as_history
    h = _as_history(lambda p, t: np.ones(1) + t, None, 0.0, 1)     # the reference tests' h(p,t) = 1 + t
    assert h.c0 == (1.0,) and h.c1 == (1.0,)
    assert _as_history(1.2, None, 0.0, 1).c0 == (1.2,)
    with pytest.raises(NotImplementedError):
        _as_history(lambda p, t: np.sin(t) * np.ones(1), None, 0.0, 1)


def test_shard_ranges_cover_exactly():
    for n, w in [(10 ** 6, 8), (10, 3), (7, 8), (0, 2), (10 ** 8, 4)]:
        rs = [S.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))


def test_merge_moments_matches_oracle():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 2, 900))
    a, b = O.moments(x[..., :301]), O.moments(x[..., 301:])
    got = S.merge_moments(a.reshape(-1, 3), b.reshape(-1, 3)).reshape(a.shape)
    assert np.allclose(got, O.merge_moments(a, b), rtol=1e-15)
    assert np.allclose(got, O.moments(x), rtol=1e-12)
    z = np.zeros_like(a)
    assert np.array_equal(S.merge_moments(z.reshape(-1, 3), a.reshape(-1, 3)).reshape(a.shape), a)


def test_ensemble_problem_batches_and_prob_func():
    c = cases.c1()
    pb = np.arange(30, dtype=float).reshape(3, 10)
    ep = S.EnsembleProblem(c.problem(), p_batch=pb)
    p, u = ep.batches(4, first=3)
    assert np.array_equal(p, pb[:, 3:7]) and u is None
    ep = S.EnsembleProblem(c.problem(), prob_func=lambda prob, i, rep: prob.remake(p=[i, 0, 0], u0=[2.0 * i]))
    p, u = ep.batches(3, first=2)
    assert np.array_equal(p[0], [3, 4, 5]) and np.array_equal(u[0], [6.0, 8.0, 10.0])     # 1-based like Julia
    with pytest.raises(NotImplementedError):
        S.EnsembleProblem(c.problem(), output_func=lambda sol, i: (sol, False))


def test_version_and_struct_guard_fields():
    L = capi.lib()
    assert L.sdde_b200_version() == 100
    assert L.sdde_b200_set_option(None, b"wave_traj", 1) == -1      # NULL context is rejected, not dereferenced
    d = capi.SddeDesc()
    assert hasattr(d, "out_pitch") and hasattr(d, "n_save_idxs") and len(d.save_idxs) == 8


def test_builtin_sra_tableaus_equal_the_oracles():
    """The coefficients the library integrates with (SRA1 / SRA3 / SOSRA) are bitwise the oracle's, whose tableaus are
    checked against the order conditions in tests/test_oracle.py."""
    from sdde_b200 import _capi as capi
    for alg, tab in O.SRA_TABLEAUS.items():
        out, st = np.zeros(33), C.c_int32(0)
        capi.check(capi.lib().sdde_b200_sra_tableau(alg, capi.np_ptr(out), C.byref(st)))
        flat = tab.flat()
        assert st.value == tab.stages and np.array_equal(out[:flat.size], flat), alg


def test_randomised_sra_schedules_equal_oracle_bitwise():
    """SRA1 / SRA3 / SOSRA / a caller's tableau: four lookup sets per step at t_n + c_k dt_n (the deduplicated stage times of
    the tableau, padded with t_n), discontinuities tracked to order 2 ([EXT] alg_order of the SRA family): the product's
    schedule equals the oracle's replay bit for bit on 60 random problems."""
    seed, trials = ([int(x) for x in os.environ["SDDE_FUZZ_SEEDS"].split(":")] if os.environ.get("SDDE_FUZZ_SEEDS") else [1315, 60])
    rng = np.random.default_rng(seed)
    custom = dict(c0=(0.0, 0.5), c1=(0.25, 1.0), A0=((0.0, 0.0), (0.5, 0.0)), B0=((0.0, 0.0), (1.0, 0.0)),
                  alpha=(0.25, 0.75), beta1=(0.5, 0.5), beta2=(1.0, -1.0))
    otab = O.SraTableau(**custom)
    for trial in range(trials):
        nl = int(rng.integers(1, 3))
        t0 = float(rng.choice([0.0, -1.5]))
        T = float(rng.choice([1.0, 3.7, 10.0]))
        dt = float(rng.choice([0.01, 0.05, 0.1, 2.0 ** -5, 0.03, 0.25]))
        lags = sorted(float(x) for x in rng.choice([0.3, 0.5, 1.0, 1.7, 17.02 / 4, 0.05 * 7, 0.25], size=nl, replace=False))
        dt = min(dt, min(lags) / float(rng.choice([1.0, 2.0])))          # src/solve.jl:581-585 guard; dt == tau allowed
        alg = int(rng.choice([13, 14, 15, 16]))
        order0 = int(rng.choice([0, 0, 1, 3]))
        saveat = rng.choice([None, 0.5, 0.37])
        saveat = None if saveat is None else float(saveat)
        model = S.hayes if nl == 1 else S.linsys(2, 2, 1)
        u0 = [1.0] * model.n_state
        prob = S.SDDEProblem(model.f, model.g, u0 if model.n_state > 1 else 1.0, S.ConstantHistory(u0), (t0, t0 + T),
                             np.zeros(model.n_param), constant_lags=tuple(lags), order_discontinuity_t0=order0)
        palg = S.SRA(tableau=custom) if alg == 16 else cases.ALG_MAP[alg]
        s = S.schedule(prob, palg, dt=dt, saveat=saveat, adaptive=False)
        sv, ss, se = O.expand_saveat(saveat, t0, t0 + T)
        o = O.build_schedule(t0, t0 + T, dt, lags, alg, order_discontinuity_t0=order0, saveat_times=sv, save_start=ss,
                             save_end=se, save_everystep=saveat is None, tableau=otab if alg == 16 else None)
        assert o.lk_mode.shape[1] == 4 * nl
        for k in ["t", "dt", "sqdt", "lk_mode", "lk_idx", "lk_theta", "save_t", "save_step", "save_theta", "save_exact"]:
            a, b = getattr(s, k), getattr(o, k)
            assert a.shape == b.shape and np.array_equal(a, b), (trial, k, t0, T, dt, lags, alg, order0, saveat)
        assert list(s.tstops_hit) == o.tstops_hit
    with pytest.raises(S.SddeError, match="explicit"):
        bad = dict(custom, A0=((0.0, 0.1), (0.5, 0.0)))
        S.schedule(prob, S.SRA(tableau=bad), dt=0.01, adaptive=False)
    with pytest.raises(NotImplementedError, match="adaptive"):
        S.schedule(prob, S.SRA1(), dt=0.01)                   # [EXT] isadaptive(SRA1): the reference would step adaptively


def test_randomised_schedules_equal_oracle_bitwise():
    """120 random (tspan, dt, lags, algorithm, order_discontinuity_t0, saveat) combinations: the product's C++ schedule
    builder and the oracle's Python replay agree on every bit of every table."""
    seed, trials = ([int(x) for x in os.environ["SDDE_FUZZ_SEEDS"].split(":")] if os.environ.get("SDDE_FUZZ_SEEDS") else [20261019, 120])
    rng = np.random.default_rng(seed)
    algs = [0, 1, 2, 3, 4]
    for trial in range(trials):
        nl = int(rng.integers(1, 3))
        t0 = float(rng.choice([0.0, -1.5, 2.0]))
        T = float(rng.choice([1.0, 3.7, 10.0, 25.0]))
        dt = float(rng.choice([0.01, 0.05, 0.1, 2.0 ** -5, 0.03, 0.25]))
        lags = sorted(float(x) for x in rng.choice([0.3, 0.5, 1.0, 1.7, 2.5, 17.02 / 4, 0.05 * 7, 0.25], size=nl, replace=False))
        alg = int(rng.choice(algs))
        if alg != 0 and alg != 1 and dt > min(lags):
            dt = min(lags) / 2                       # src/solve.jl:581-585 guard
        order0 = int(rng.choice([0, 0, 1, 2]))
        saveat = rng.choice([None, 0.5, 0.1, 1.0, 0.37])
        saveat = None if saveat is None else float(saveat)
        model = S.linear1 if nl == 1 else S.linsys(2, 2, 1)
        omodel_p = np.zeros(model.n_param)
        u0 = [1.0] * model.n_state
        prob = S.SDDEProblem(model.f, model.g, u0 if model.n_state > 1 else 1.0, S.ConstantHistory(u0), (t0, t0 + T),
                             omodel_p, constant_lags=tuple(lags), order_discontinuity_t0=order0)
        kw = dict(dt=dt, saveat=saveat)
        if alg in (2, 3):
            kw["adaptive"] = False
            if nl > 1:
                continue                             # RKMil needs diagonal noise; linsys has a noise matrix
        s = S.schedule(prob, cases.ALG_MAP[alg], **kw)
        sv, ss, se = O.expand_saveat(saveat, t0, t0 + T)
        o = O.build_schedule(t0, t0 + T, dt, lags, alg, order_discontinuity_t0=order0, saveat_times=sv, save_start=ss,
                             save_end=se, save_everystep=saveat is None)
        for k in ["t", "dt", "sqdt", "lk_mode", "lk_idx", "lk_theta", "save_t", "save_step", "save_theta", "save_exact"]:
            a, b = getattr(s, k), getattr(o, k)
            assert a.shape == b.shape and np.array_equal(a, b), (trial, k, t0, T, dt, lags, alg, order0, saveat)
        assert list(s.tstops_hit) == o.tstops_hit


def test_randomised_schedules_with_explicit_saveat_tstops_and_flags():
    """Explicit saveat vectors (with and without the end points), user tstops, save_start/save_end overrides and
    undeclared lags: product schedule == oracle replay, bit for bit, on 80 random problems."""
    rng = np.random.default_rng(77)
    for trial in range(80):
        t0 = float(rng.choice([0.0, 1.0]))
        T = float(rng.choice([2.0, 5.5, 12.0]))
        tf = t0 + T
        dt = float(rng.choice([0.01, 0.04, 0.1, 2.0 ** -4]))
        lag = float(rng.choice([0.5, 1.0, 1.7, 0.33]))
        alg = int(rng.choice([0, 1, 4]))
        declared = bool(rng.integers(0, 2))
        npts = int(rng.integers(1, 9))
        sv_user = sorted(set(float(np.round(x, 3)) for x in rng.uniform(t0, tf, size=npts)))
        if rng.integers(0, 2):
            sv_user = [t0] + sv_user
        if rng.integers(0, 2):
            sv_user = sv_user + [tf]
        tst = sorted(set(float(np.round(x, 2)) for x in rng.uniform(t0, tf, size=int(rng.integers(0, 4)))))
        prob = S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (t0, tf), (-1.0, 0.5, 0.2),
                             constant_lags=(lag,) if declared else (), lags=(lag,))
        s = S.schedule(prob, cases.ALG_MAP[alg], dt=dt, saveat=sv_user, tstops=tst)
        sv, ss, se = O.expand_saveat(sv_user, t0, tf)
        o = O.build_schedule(t0, tf, dt, [lag], alg, constant_lags_declared=declared, saveat_times=sv, save_start=ss, save_end=se,
                             save_everystep=False, tstops=tst)
        for k in ["t", "dt", "sqdt", "lk_mode", "lk_idx", "lk_theta", "save_t", "save_step", "save_theta", "save_exact"]:
            a, b = getattr(s, k), getattr(o, k)
            assert a.shape == b.shape and np.array_equal(a, b), (trial, k, t0, tf, dt, lag, alg, declared, sv_user, tst)
        assert list(s.tstops_hit) == o.tstops_hit


def test_order_statistic_planning_host_only():
    """sdde_b200_qsel_ranks (no GPU): how many distinct order statistics a request needs -- Statistics.quantile brackets
    each probability with two neighbours, exact hits and shared neighbours collapse, the median of an odd count is one."""
    import ctypes as C
    L = capi.lib()

    def ranks(n, probs, median):
        p = np.asarray(probs, dtype=np.float64)
        out = C.c_int32(-1)
        rc = L.sdde_b200_qsel_ranks(n, capi.np_ptr(p).value if p.size else None, p.size, 1 if median else 0, C.byref(out))
        return rc, out.value

    assert ranks(1, [0.3], True) == (0, 1)                     # a single trajectory: every statistic is that value
    assert ranks(5, [], True) == (0, 1)                        # odd count: the middle one
    assert ranks(4, [], True) == (0, 2)                        # even count: the two central ones
    assert ranks(5, [0.5], False) == (0, 2)                    # h = 3 exactly: v[3] and v[4] (gamma = 0)
    assert ranks(1001, [0.05, 0.95], True) == (0, 5)           # 2 + 2 + 1
    assert ranks(10, [0.0, 1.0], False) == (0, 4)              # clamped to [1, n-1]: (1,2) and (9,10)
    rc, _ = ranks(10, [1.5], False)
    assert rc == -1 and b"[0, 1]" in L.sdde_b200_last_error()
    rc, _ = ranks(1000, list(np.linspace(0.01, 0.99, 14)), False)
    assert rc == -2 and b"distinct order statistics" in L.sdde_b200_last_error()


def test_problem_surface_argument_checks_without_gpu():
    """Host-side mirror of the SciML surface: what is outside the path is refused with a clear error before any GPU work."""
    um_h = S.UserModel("struct UserModel {};", 2, 2, 2, 2, True, history=True)
    um = S.UserModel("struct UserModel {};", 2, 1, 2, 3, True)
    with pytest.raises(ValueError):                       # a model with its own h must be given that h
        S.SDDEProblem(um_h.f, um_h.g, [1.0, 0.5], 1.0, (0.0, 1.0), [1.0, 0.1], constant_lags=(0.5, 1.0))
    with pytest.raises(AttributeError):
        um.h
    with pytest.raises(ValueError):                       # h of another model
        S.SDDEProblem(um.f, um.g, [1.0, 0.5], um_h.h, (0.0, 1.0), [1.0, 0.1, 0.2], constant_lags=(0.5,))
    with pytest.raises(TypeError):                        # Python callables cannot cross the C ABI as f, g
        S.SDDEProblem(lambda *a: 0, lambda *a: 0, 1.0, 1.0, (0.0, 1.0))
    with pytest.raises(ValueError):
        S.SDDEProblem(S.linear1.f, S.hayes.g, 1.0, 1.0, (0.0, 1.0), (1.0, 1.0, 1.0), constant_lags=(1.0,))
    with pytest.raises(NotImplementedError):
        S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0.0, 1.0), (1.0, 1.0, 1.0), constant_lags=(1.0,), dependent_lags=(lambda u, p, t: 1.0,))
    with pytest.raises(NotImplementedError):
        S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0.0, 1.0), (1.0, 1.0, 1.0), constant_lags=(1.0,), neutral=True)
    prob = S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, (0.0, 1.0), (1.0, 1.0, 1.0), constant_lags=(1.0,))
    with pytest.raises(NotImplementedError):
        S.EnsembleProblem(prob, output_func=lambda sol, i: (sol, False))
    with pytest.raises(NotImplementedError):              # non-affine Python history for a built-in model
        S.schedule(S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, lambda p, t: np.cos(t), (0.0, 1.0), (1.0, 1.0, 1.0),
                                 constant_lags=(1.0,)), S.EM(), dt=0.1)
    with pytest.raises(NotImplementedError):              # RKMil is adaptive by default in the reference
        S.schedule(prob, S.RKMil(), dt=0.1)
    # RKMilCommute has its own ids (5 / 6); a diagonal model is served by the RKMil kernels inside the library
    assert (S.RKMilCommute().alg_id, S.RKMilCommute("Stratonovich").alg_id) == (5, 6) and isinstance(S.RKMilCommute(), S.RKMil)
    # output_func / reduction: the expressible forms are accepted, arbitrary closures are refused
    S.EnsembleProblem(prob, output_func=S.FinalState(), reduction=S.MomentsReduction())
    S.EnsembleProblem(prob, output_func=S.SaveIdxs([0]))
    with pytest.raises(NotImplementedError):
        S.EnsembleProblem(prob, reduction=lambda u, data, I: (u, False))
    with pytest.raises(ValueError):
        S.RKMil("Foo")


def test_plain_c_program_drives_the_abi(tmp_path):
    """tests/c/abi_smoke.c: a C99 program that includes only include/sdde_b200.h and links libsdde_b200.so -- the boundary
    carries no Python, torch or C++ types.  Host-only entry points here; with a GPU the same program also solves."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = str(tmp_path / "abi_smoke")
    subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "c", "abi_smoke.c"),
                    "-o", exe, "-L", libdir, "-lsdde_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "C1 steps 1001 saves 11" in out
    assert "C2 steps 2001 saves 101 ring 342" in out and "tstops 17 34 100" in out
    assert "bad struct_size -> -1 (refused)" in out


def test_schedule_structural_invariants_on_random_problems():
    """Independent re-derivation, from the lookup tables alone, of the two numbers the kernels trust: `ring_slots` (no node
    is overwritten while a later step may still read it) and `min_gap` (how many steps may be batched before a lookup could
    touch a node of the same batch); plus monotone time, steps that tile the span, saves inside their step."""
    rng = np.random.default_rng(4242)
    for trial in range(150):
        t0 = float(rng.choice([0.0, 0.5]))
        T = float(rng.choice([1.0, 3.3, 8.0]))
        dt = float(rng.choice([0.01, 0.03, 0.1, 2.0 ** -4]))
        nl = int(rng.integers(1, 3))
        lags = sorted(float(x) for x in rng.choice([0.05, 0.3, 0.31, 1.0, 1.7, 2.5], size=nl, replace=False))
        alg = int(rng.choice([0, 1, 2, 4]))
        if alg in (2, 4) and dt > min(lags):
            dt = min(lags) / 2
        model = S.linear1 if nl == 1 else S.linsys(2, 2, 1)
        if alg == 2 and nl > 1:
            alg = 0
        u0 = [1.0] * model.n_state
        prob = S.SDDEProblem(model.f, model.g, u0 if model.n_state > 1 else 1.0, S.ConstantHistory(u0), (t0, t0 + T),
                             np.zeros(model.n_param), constant_lags=tuple(lags), order_discontinuity_t0=int(rng.choice([0, 1, 2])))
        kw = dict(dt=dt, saveat=float(rng.choice([0.1, 0.37, 1.0])))
        if alg == 2:
            kw["adaptive"] = False
        s = S.schedule(prob, cases.ALG_MAP[alg], **kw)
        N = len(s.dt)
        assert np.all(np.diff(s.t) > 0) and s.t[0] == t0 and s.t[-1] == t0 + T
        assert np.all(s.dt > 0) and np.all(s.dt <= dt * (1 + 1e-12)) and abs(float(np.sum(s.dt)) - T) < 1e-9
        need, gap = 1, 2 ** 31 - 1
        for n in range(N):
            for e in range(s.lk_mode.shape[1]):
                mode, i = int(s.lk_mode[n, e]), int(s.lk_idx[n, e])
                if mode == 1:                       # exact node i
                    assert 0 <= i <= n and 0.0 <= s.lk_theta[n, e] <= 1.0
                    need = max(need, n - i); gap = min(gap, n - i)
                elif mode == 2:                     # between nodes i-1 and i
                    assert 1 <= i <= n and 0.0 < s.lk_theta[n, e] < 1.0
                    need = max(need, n - (i - 1)); gap = min(gap, n - i)
        assert s.ring_slots == max(2, need + 1), (trial, s.ring_slots, need)
        assert s.min_gap == gap, (trial, s.min_gap, gap)
        k = s.save_step >= 0
        assert np.all(s.save_t[k] <= s.t[s.save_step[k] + 1]) and np.all(s.save_t[k] >= s.t[s.save_step[k]])
        assert np.all((s.save_theta >= 0) & (s.save_theta <= 1))


def test_hopeless_or_non_finite_schedules_fail_fast():
    """A step count beyond the limit, an infinite span, NaN stops / save times are refused in milliseconds instead of being
    replayed (the reference would step until maxiters)."""
    import time
    def sched(**kw):
        prob = S.SDDEProblem(S.linear1.f, S.linear1.g, 1.0, 1.0, kw.pop("tspan", (0.0, 10.0)), (-1.0, 0.5, 0.2), constant_lags=(1.0,))
        return S.schedule(prob, S.EM(), **{"dt": 0.01, **kw})
    t0 = time.perf_counter()
    for kw, exc, pat in (({"dt": 1e-9}, S.SddeError, "maximum number of steps"), ({"dt": 1e-300}, S.SddeError, "maximum number of steps"),
                         ({"tspan": (0.0, float("inf"))}, S.SddeError, "finite"), ({"tstops": (float("nan"),)}, S.SddeError, "NaN"),
                         ({"saveat": 1e-9}, ValueError, "save points"), ({"saveat": float("nan")}, ValueError, "positive"),
                         ({"saveat": [float("nan"), 1.0]}, ValueError, "NaN"), ({"dt": float("nan")}, S.SddeError, "dt must be finite")):
        with pytest.raises(exc, match=pat):
            sched(**kw)
    assert time.perf_counter() - t0 < 2.0
    assert len(sched(tspan=(1e300, 2e300), dt=1e299).t) == 11 and len(sched(tstops=(5.0, 5.0, 5.0 + 1e-15)).t) == 1004


def test_solution_interpolation_and_timepoint_statistics():
    """``sol(t)`` / ``sim.at(t)``: linear interpolation between the saved points (the reference's sde_interpolant on a saved
    series), exact at saved points, no extrapolation; ``timepoint_meanvar`` at a time between saves uses it."""
    import torch
    t = np.array([0.0, 1.0, 2.0, 4.0])
    out = np.arange(4 * 2 * 3, dtype=float).reshape(4, 2, 3) ** 2
    for o in (out, torch.from_numpy(out)):
        e = S.EnsembleSolution(t, o, np.zeros(3, np.uint8), None, False, 0.0, {})
        assert np.array_equal(np.asarray(e.at(2.0)), out[2])
        assert np.allclose(np.asarray(e.at(3.0)), 0.5 * out[2] + 0.5 * out[3])
        assert np.allclose(e[1](0.25), 0.75 * out[0, :, 1] + 0.25 * out[1, :, 1])
        assert np.asarray(e.at([0.5, 4.0])).shape == (2, 2, 3) and e[2]([0.5, 4.0]).shape == (2, 2)
        m, v = S.timepoint_meanvar(e, 3.0)
        x = 0.5 * out[2] + 0.5 * out[3]
        assert np.allclose(m, x.mean(axis=1)) and np.allclose(v, x.var(axis=1, ddof=1))
        with pytest.raises(ValueError, match="extrapolate"):
            e.at(4.5)
    es = S.EnsembleSolution(t, out[:, :1, :], None, None, True, 0.0, {})
    assert np.isscalar(float(es[0](1.5))) and float(es[0](1.5)) == 0.5 * out[1, 0, 0] + 0.5 * out[2, 0, 0]
