"""ctypes loader of oracle/libsdde_oracle.so (the C++ restatement).  TEST INFRASTRUCTURE ONLY -- see sdde_oracle.cpp.

The descriptor is the product's own ``sdde_desc`` (include/sdde_b200.h); it is re-declared here so that the oracle does
not import the product package."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libsdde_oracle.so")


class Desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("model_id", C.c_int32), ("alg", C.c_int32),
        ("n_state", C.c_int32), ("n_noise", C.c_int32), ("n_lag", C.c_int32), ("n_param", C.c_int32),
        ("order_discontinuity_t0", C.c_int32), ("constant_lags_declared", C.c_int32),
        ("save_start", C.c_int32), ("save_end", C.c_int32), ("save_everystep", C.c_int32),
        ("noise_mode", C.c_int32), ("noise_mem", C.c_int32), ("out_mem", C.c_int32), ("history_mode", C.c_int32),
        ("want_paths", C.c_int32), ("want_moments", C.c_int32), ("lanes_per_traj", C.c_int32),
        ("n_saveat", C.c_int32), ("n_tstops", C.c_int32), ("reserved0", C.c_int32),
        ("t0", C.c_double), ("tf", C.c_double), ("dt", C.c_double),
        ("lags", C.c_double * 4), ("hist_c0", C.c_double * 8), ("hist_c1", C.c_double * 8),
        ("params", C.c_void_p), ("param_stride", C.c_int64), ("u0", C.c_void_p), ("u0_stride", C.c_int64),
        ("saveat", C.c_void_p), ("tstops", C.c_void_p), ("noise", C.c_void_p),
        ("n_traj", C.c_int64), ("traj_offset", C.c_int64), ("seed", C.c_uint64), ("out_pitch", C.c_int64),
        ("n_save_idxs", C.c_int32), ("save_idxs", C.c_int32 * 8), ("reserved1", C.c_int32),
        ("noise_t", C.c_void_p), ("n_noise_t", C.c_int64),
        ("d_disc_t", C.c_void_p), ("d_disc_order2", C.c_void_p), ("n_d_disc", C.c_int64), ("out_rows", C.c_int64),
        ("sra_tableau", C.c_void_p), ("sra_stages", C.c_int32), ("sra_order2", C.c_int32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            import subprocess
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = C.CDLL(LIB)
        _lib.sdde_oracle_solve.argtypes = [C.POINTER(Desc), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                           C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    return _lib


def solve(model_id, alg, n_state, n_lag, n_noise, params, u0, h0, h1, lags, t0, tf, dt, saveat_times, n_save, *,
          save_start=True, save_end=True, save_everystep=False, order0=0, declared=True, n_traj=1, traj_offset=0,
          seed=0, dW=None, noise_grid=None, n_threads=1, d_discontinuities=(), tableau=None, rng="philox"):
    """-> (out[n_save, n_state, n_traj], status[n_traj], elapsed seconds, steps)."""
    d = Desc()
    d.struct_size = C.sizeof(Desc)
    d.model_id, d.alg = model_id, alg
    d.n_state, d.n_noise, d.n_lag = n_state, n_noise, n_lag
    params = np.ascontiguousarray(params, dtype=np.float64)
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    d.n_param = params.shape[0]
    d.params = params.ctypes.data
    d.param_stride = n_traj if params.ndim == 2 else 0
    d.u0 = u0.ctypes.data
    d.u0_stride = n_traj if u0.ndim == 2 else 0
    d.order_discontinuity_t0, d.constant_lags_declared = order0, int(declared)
    d.save_start, d.save_end, d.save_everystep = int(save_start), int(save_end), int(save_everystep)
    d.t0, d.tf, d.dt = t0, tf, dt
    for i, x in enumerate(lags):
        d.lags[i] = x
    for j in range(n_state):
        d.hist_c0[j], d.hist_c1[j] = h0[j], h1[j]
    sv = np.ascontiguousarray(saveat_times, dtype=np.float64)
    d.saveat, d.n_saveat = (sv.ctypes.data if sv.size else None), sv.size
    d.n_traj, d.traj_offset, d.seed = n_traj, traj_offset, seed
    dd = [(float(x[0]), int(round(2 * x[1]))) if isinstance(x, (tuple, list)) else (float(x), 0) for x in d_discontinuities]
    dd_t = np.array([x[0] for x in dd], dtype=np.float64)
    dd_o = np.array([x[1] for x in dd], dtype=np.int32)
    if dd:
        d.d_disc_t, d.d_disc_order2, d.n_d_disc = dd_t.ctypes.data, dd_o.ctypes.data, len(dd)
    if 13 <= alg <= 16:      # SRA family: the C++ restatement takes the coefficients from the Python oracle's table
        import sdde_oracle as _O
        tab = _O.sra_tableau(alg, tableau)
        flat = tab.flat()
        d.sra_tableau, d.sra_stages = flat.ctypes.data, tab.stages
    nz, nrows = None, 0
    if dW is not None:
        nz = np.ascontiguousarray(dW, dtype=np.float64); d.noise_mode = 1; nrows = nz.shape[0]
    elif noise_grid is not None:
        nz = np.ascontiguousarray(noise_grid, dtype=np.float64); d.noise_mode = 2; nrows = nz.shape[0]
    d.noise = nz.ctypes.data if nz is not None else None
    if nz is None and rng == "ziggurat":        # timing only: xoroshiro128+ + ziggurat, the cost class of the reference's randn
        d.noise_mode = 100
    out = np.full((n_save, n_state, n_traj), np.nan)
    status = np.zeros(n_traj, dtype=np.uint8)
    el, steps = C.c_double(0), C.c_int64(0)
    rc = lib().sdde_oracle_solve(C.byref(d), n_save, nrows, out.ctypes.data, status.ctypes.data, int(n_threads),
                                 C.byref(el), C.byref(steps))
    if rc != 0:
        raise RuntimeError(f"sdde_oracle_solve failed: {rc}")
    return out, status, el.value, steps.value
