"""B200-native ensemble integrator for constant-lag stochastic delay differential equations.

Drop-in for ONE path of SciML/StochasticDelayDiffEq.jl: ``solve(EnsembleProblem(SDDEProblem(...)), EM()/RKMil(),
EnsembleB200(); dt, saveat, trajectories)`` -- host code here only marshals a descriptor across the C ABI of
``libsdde_b200.so`` (include/sdde_b200.h); all arithmetic runs in hand-written sm_100a CUDA kernels
(csrc/sdde_kernels.cuh).  There is no CPU fallback.
"""
from . import _capi
from ._capi import SddeError, get_context, lib
from .ensemble import (EnsembleB200, EnsembleSolution, EnsembleSummary, Schedule, expand_saveat, merge_moments, pinned_empty, schedule, shard_range,
                       solve, timepoint_meanvar, timeseries_steps_mean, timeseries_steps_meanvar, timeseries_steps_median,
                       timeseries_steps_quantile, get_timestep, get_timepoint, timestep_meanvar, timestep_mean, timestep_quantile,
                       timestep_median, timepoint_mean, timepoint_quantile, timepoint_median, timeseries_point_meanvar,
                       timeseries_point_mean, timeseries_point_quantile, timeseries_point_median)
from .problem import (EM, EulerHeun, AffineHistory, ConstantHistory, EnsembleProblem, FinalState, Model, MomentsReduction, RKMil, RKMilCommute,
                      SaveIdxs, SDDEProblem, SOSRA, SRA, SRA1, SRA3, UserModel, WangLi3SMil, WangLi3SMil_A, WangLi3SMil_B, WangLi3SMil_C, WangLi3SMil_D,
                      WangLi3SMil_E, WangLi3SMil_F, hayes,
                      linear1, linsys, linsys_params, mackey_glass, scalar_add, sir)

__all__ = [
    "EM", "RKMil", "RKMilCommute", "EulerHeun", "SRA", "SRA1", "SRA3", "SOSRA", "WangLi3SMil", "WangLi3SMil_A", "WangLi3SMil_B", "WangLi3SMil_C", "WangLi3SMil_D",
    "WangLi3SMil_E", "WangLi3SMil_F", "FinalState", "SaveIdxs", "MomentsReduction", "SDDEProblem", "EnsembleProblem", "EnsembleB200", "EnsembleSolution", "solve", "schedule", "Schedule",
    "expand_saveat", "merge_moments", "pinned_empty", "shard_range", "timeseries_steps_meanvar", "timeseries_steps_mean",
    "timepoint_meanvar", "timeseries_steps_quantile", "timeseries_steps_median", "EnsembleSummary", "AffineHistory", "ConstantHistory", "Model", "UserModel", "linear1", "hayes", "mackey_glass",
    "sir", "scalar_add", "linsys", "linsys_params", "SddeError", "get_context", "lib",
    "get_timestep", "get_timepoint", "timestep_meanvar", "timestep_mean", "timestep_quantile", "timestep_median", "timepoint_mean", "timepoint_quantile", "timepoint_median", "timeseries_point_meanvar", "timeseries_point_mean", "timeseries_point_quantile", "timeseries_point_median",
]
__version__ = "0.1.0"
