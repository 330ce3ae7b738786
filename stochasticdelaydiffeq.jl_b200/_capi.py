"""ctypes binding of libsdde_b200.so -- the in-container stand-in for Julia's ``ccall``.

Mirrors include/sdde_b200.h field for field.  There is no fallback: if the shared library is missing or no CUDA
device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SDDE_MAX_STATE, SDDE_MAX_LAG, SDDE_MAX_NOISE, SDDE_MAX_PARAM = 8, 4, 8, 128

ALG_EM_SPLIT, ALG_EM_NOSPLIT, ALG_RKMIL_ITO, ALG_RKMIL_STRAT, ALG_EULER_HEUN, ALG_RKMILC_ITO, ALG_RKMILC_STRAT = 0, 1, 2, 3, 4, 5, 6
ALG_WANGLI3SMIL_A = 7     # .. ALG_WANGLI3SMIL_F = 12
ALG_SRA1, ALG_SRA3, ALG_SOSRA, ALG_SRA_TABLEAU = 13, 14, 15, 16
MODEL_LINEAR1, MODEL_HAYES, MODEL_MACKEY_GLASS, MODEL_LINSYS, MODEL_SIR, MODEL_SCALAR_ADD = 0, 1, 2, 3, 4, 5
MODEL_USER_BASE = 1000
NOISE_PHILOX, NOISE_SUPPLIED_DW, NOISE_SUPPLIED_GRID = 0, 1, 2
MEM_HOST, MEM_DEVICE = 0, 1

_dp = C.POINTER(C.c_double)


class SddeDesc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("model_id", C.c_int32), ("alg", C.c_int32),
        ("n_state", C.c_int32), ("n_noise", C.c_int32), ("n_lag", C.c_int32), ("n_param", C.c_int32),
        ("order_discontinuity_t0", C.c_int32), ("constant_lags_declared", C.c_int32),
        ("save_start", C.c_int32), ("save_end", C.c_int32), ("save_everystep", C.c_int32),
        ("noise_mode", C.c_int32), ("noise_mem", C.c_int32), ("out_mem", C.c_int32), ("history_mode", C.c_int32),
        ("want_paths", C.c_int32), ("want_moments", C.c_int32), ("lanes_per_traj", C.c_int32),
        ("n_saveat", C.c_int32), ("n_tstops", C.c_int32), ("reserved0", C.c_int32),
        ("t0", C.c_double), ("tf", C.c_double), ("dt", C.c_double),
        ("lags", C.c_double * SDDE_MAX_LAG),
        ("hist_c0", C.c_double * SDDE_MAX_STATE), ("hist_c1", C.c_double * SDDE_MAX_STATE),
        ("params", C.c_void_p), ("param_stride", C.c_int64),
        ("u0", C.c_void_p), ("u0_stride", C.c_int64),
        ("saveat", C.c_void_p), ("tstops", C.c_void_p), ("noise", C.c_void_p),
        ("n_traj", C.c_int64), ("traj_offset", C.c_int64), ("seed", C.c_uint64), ("out_pitch", C.c_int64),
        ("n_save_idxs", C.c_int32), ("save_idxs", C.c_int32 * 8), ("reserved1", C.c_int32),
        ("noise_t", C.c_void_p), ("n_noise_t", C.c_int64),
        ("d_disc_t", C.c_void_p), ("d_disc_order2", C.c_void_p), ("n_d_disc", C.c_int64), ("out_rows", C.c_int64),
        ("sra_tableau", C.c_void_p), ("sra_stages", C.c_int32), ("sra_order2", C.c_int32),
    ]


class SddeSchedule(C.Structure):
    _fields_ = [
        ("n_steps", C.c_int64), ("n_save", C.c_int64), ("n_lag", C.c_int32), ("ring_slots", C.c_int32),
        ("min_gap", C.c_int32), ("n_tstops_hit", C.c_int32), ("n_stage", C.c_int32), ("reserved", C.c_int32),
        ("t", _dp), ("dt", _dp), ("sqdt", _dp),
        ("lk_mode", C.POINTER(C.c_int32)), ("lk_idx", C.POINTER(C.c_int64)), ("lk_theta", _dp),
        ("save_t", _dp), ("save_step", C.POINTER(C.c_int64)), ("save_theta", _dp),
        ("save_exact", C.POINTER(C.c_int32)), ("tstops_hit", _dp),
    ]


class SddeTiming(C.Structure):
    _fields_ = [
        ("kernel_ms", C.c_double), ("total_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("n_steps", C.c_int64), ("n_save", C.c_int64),
        ("launches", C.c_int32), ("waves", C.c_int32), ("ring_slots", C.c_int32), ("regs_per_thread", C.c_int32),
        ("block_threads", C.c_int32), ("traj_per_thread", C.c_int32), ("moment_merge", C.c_int32), ("n_unstable", C.c_int32),
        ("ring_bytes", C.c_int64),
        ("kernel_name", C.c_char * 128),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["kernel_name"] = self.kernel_name.decode()
        return d


# every symbol include/sdde_b200.h declares (tests check that the library exports them all)
EXPORTS = [
    "sdde_b200_version", "sdde_b200_device_count", "sdde_b200_last_error", "sdde_b200_create", "sdde_b200_destroy",
    "sdde_b200_set_option", "sdde_b200_schedule", "sdde_b200_schedule_free", "sdde_b200_solve", "sdde_b200_solve_multi",
    "sdde_b200_merge_moments", "sdde_b200_compile_model", "sdde_b200_philox_normals", "sdde_b200_sra_tableau", "sdde_b200_dfma_peak",
    "sdde_b200_hbm_peak", "sdde_b200_alloc_pinned", "sdde_b200_free_pinned", "sdde_b200_quantiles",
    "sdde_b200_qsel_ranks", "sdde_b200_qsel_begin", "sdde_b200_qsel_local", "sdde_b200_qsel_apply", "sdde_b200_qsel_nan",
    "sdde_b200_qsel_end", "sdde_b200_moments", "sdde_b200_rng_table",
]

LIB_PATH = os.environ.get("SDDE_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsdde_b200.so")
_lib = None


class SddeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsdde_b200 error {code}: {msg}")
        self.code = code
        self.message = msg


def lib():
    """Load libsdde_b200.so (built in-tree by ``__graft_entry__.build()`` / ``make -C .../csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if "SDDE_B200_NCCL_LIB" not in os.environ:
        # the library resolves libnccl at run time (sdde_b200_solve_multi); a process gets ONE libnccl (shared by SONAME),
        # so point it at the copy PyTorch ships -- a later `import torch` then finds the NCCL it was built against
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            for d in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(d, "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["SDDE_B200_NCCL_LIB"] = cand
                    break
        except Exception:
            pass
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for this path)")
    L = C.CDLL(LIB_PATH)
    L.sdde_b200_version.restype = C.c_int
    L.sdde_b200_device_count.restype = C.c_int
    L.sdde_b200_last_error.restype = C.c_char_p
    L.sdde_b200_create.restype = C.c_void_p
    L.sdde_b200_create.argtypes = [C.c_int]
    L.sdde_b200_destroy.argtypes = [C.c_void_p]
    L.sdde_b200_destroy.restype = None
    L.sdde_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.sdde_b200_schedule.argtypes = [C.POINTER(SddeDesc), C.POINTER(SddeSchedule)]
    L.sdde_b200_schedule_free.argtypes = [C.POINTER(SddeSchedule)]
    L.sdde_b200_schedule_free.restype = None
    L.sdde_b200_solve.argtypes = [C.c_void_p, C.POINTER(SddeDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SddeTiming)]
    L.sdde_b200_solve_multi.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.POINTER(SddeDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.sdde_b200_merge_moments.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.sdde_b200_compile_model.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int32)]
    L.sdde_b200_sra_tableau.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_int32)]
    L.sdde_b200_philox_normals.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]
    L.sdde_b200_quantiles.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int32,
                                      C.c_int32, C.c_void_p, _dp]
    L.sdde_b200_qsel_ranks.argtypes = [C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    L.sdde_b200_qsel_begin.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_void_p]
    L.sdde_b200_qsel_local.argtypes = [C.c_void_p, C.c_int32]
    L.sdde_b200_qsel_apply.argtypes = [C.c_void_p, C.c_int32]
    L.sdde_b200_qsel_nan.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.sdde_b200_qsel_end.argtypes = [C.c_void_p, C.c_void_p]
    L.sdde_b200_moments.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]
    if hasattr(L, "sdde_b200_rng_table"):      # absent only in older builds loaded through SDDE_B200_LIB for A/B timing
        L.sdde_b200_rng_table.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    L.sdde_b200_dfma_peak.argtypes = [C.c_void_p, _dp]
    L.sdde_b200_hbm_peak.argtypes = [C.c_void_p, _dp]
    L.sdde_b200_alloc_pinned.argtypes = [C.c_int64]
    L.sdde_b200_alloc_pinned.restype = C.c_void_p
    L.sdde_b200_free_pinned.argtypes = [C.c_void_p]
    L.sdde_b200_free_pinned.restype = None
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SddeError(rc, lib().sdde_b200_last_error().decode(errors="replace"))


def rng_table() -> np.ndarray:
    """Host copy of the normal generator's lookup table, shape (n_entries, 2) (layout: csrc/sdde_rng_tables.cuh)."""
    n = C.c_int64(0)
    check(lib().sdde_b200_rng_table(None, 0, C.byref(n)))
    out = np.empty((n.value, 2))
    check(lib().sdde_b200_rng_table(np_ptr(out), out.size, C.byref(n)))
    return out


def np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One GPU.  Wraps sdde_ctx*; destroyed with the object."""

    def __init__(self, device: int = 0):
        L = lib()
        self._h = L.sdde_b200_create(int(device))
        if not self._h:
            raise SddeError(-100, L.sdde_b200_last_error().decode(errors="replace"))
        self.device = int(device)
        for kv in filter(None, os.environ.get("SDDE_B200_OPTS", "").split(",")):     # A/B runs: "ring_blocked=0,lean=0"
            k, v = kv.split("=")
            self.set_option(k.strip(), int(v))

    def set_option(self, key: str, value: int):
        check(lib().sdde_b200_set_option(self._h, key.encode(), int(value)))

    def close(self):
        if getattr(self, "_h", None):
            lib().sdde_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h


_contexts: dict[int, Context] = {}


def get_context(device: int = 0) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
