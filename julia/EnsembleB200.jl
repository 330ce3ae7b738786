# EnsembleB200.jl -- the reference-side binding a StochasticDelayDiffEq.jl maintainer would add.
#
# NOT RUNNABLE IN THIS REPOSITORY'S CONTAINER (no Julia in the image); kept thin on purpose.  It adds ONE method to the
# SciML ensemble dispatch ([EXT] SciMLBase.__solve(::AbstractEnsembleProblem, alg, ensemblealg; trajectories, ...)):
# when the third positional argument of `solve` is `EnsembleB200()`, the per-trajectory loop over
# DiffEqBase.__solve(prob::AbstractSDDEProblem, alg) (src/solve.jl:1-9) is replaced by one ccall into libsdde_b200.so.
#
#   using StochasticDelayDiffEq; include("EnsembleB200.jl")
#   prob  = SDDEProblem(B200Models.mackey_glass..., 1.2, (p,t)->1.2, (0.0,500.0), (0.2,0.1,10.0,0.1); constant_lags=(17.0,))
#   sim   = solve(EnsembleProblem(prob), EM(), EnsembleB200(); dt=0.05, saveat=1.0, trajectories=10^6)
module SDDEB200

using SciMLBase, StochasticDiffEq, StochasticDelayDiffEq

const libsdde = get(ENV, "LIBSDDE_B200", "libsdde_b200.so")

struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    device::Cint
end
EnsembleB200() = EnsembleB200(0)

# Device-model handles stand in for f and g (Julia closures cannot cross the C ABI; see sdde_b200_compile_model for
# user CUDA source).  A handle is callable so that the same SDDEProblem still runs on the CPU path.
struct B200Model{F, G}
    id::Cint; n_state::Cint; n_lag::Cint; n_noise::Cint; n_param::Cint
    f::F; g::G
end

# mirrors `struct sdde_desc` of include/sdde_b200.h field for field
Base.@kwdef mutable struct SddeDesc
    struct_size::Int32 = 0; model_id::Int32 = 0; alg::Int32 = 0
    n_state::Int32 = 1; n_noise::Int32 = 1; n_lag::Int32 = 1; n_param::Int32 = 0
    order_discontinuity_t0::Int32 = 0; constant_lags_declared::Int32 = 1
    save_start::Int32 = 1; save_end::Int32 = 1; save_everystep::Int32 = 0
    noise_mode::Int32 = 0; noise_mem::Int32 = 0; out_mem::Int32 = 0; history_mode::Int32 = 0
    want_paths::Int32 = 1; want_moments::Int32 = 0; lanes_per_traj::Int32 = 0
    n_saveat::Int32 = 0; n_tstops::Int32 = 0; reserved0::Int32 = 0
    t0::Float64 = 0; tf::Float64 = 0; dt::Float64 = 0
    lags::NTuple{4, Float64} = (0.0, 0.0, 0.0, 0.0)
    hist_c0::NTuple{8, Float64} = ntuple(_ -> 0.0, 8); hist_c1::NTuple{8, Float64} = ntuple(_ -> 0.0, 8)
    params::Ptr{Float64} = C_NULL; param_stride::Int64 = 0
    u0::Ptr{Float64} = C_NULL; u0_stride::Int64 = 0
    saveat::Ptr{Float64} = C_NULL; tstops::Ptr{Float64} = C_NULL; noise::Ptr{Float64} = C_NULL
    n_traj::Int64 = 0; traj_offset::Int64 = 0; seed::UInt64 = 0
    out_pitch::Int64 = 0
    n_save_idxs::Int32 = 0; save_idxs::NTuple{8, Int32} = ntuple(_ -> Int32(0), 8); reserved1::Int32 = 0
    noise_t::Ptr{Float64} = C_NULL; n_noise_t::Int64 = 0
end

alg_id(::EM{true}) = Int32(0); alg_id(::EM{false}) = Int32(1)
alg_id(a::Union{RKMil, RKMilCommute}) = StochasticDiffEq.alg_interpretation(a) == :Ito ? Int32(2) : Int32(3)   # diagonal noise: same update
alg_id(::EulerHeun) = Int32(4)

last_error() = unsafe_string(ccall((:sdde_b200_last_error, libsdde), Cstring, ()))

const CTX = Dict{Cint, Ptr{Cvoid}}()
function context(dev)
    get!(CTX, dev) do
        h = ccall((:sdde_b200_create, libsdde), Ptr{Cvoid}, (Cint,), dev)
        h == C_NULL && error("libsdde_b200: " * last_error())
        h
    end
end

"""
    pinned_array(dims...) -> Array{Float64}

Page-locked host memory from the library (`sdde_b200_alloc_pinned`), wrapped as a Julia array and freed by a finalizer.
With a page-locked `out` the step kernel writes the saved paths straight into it over PCIe (no staging copy).
"""
function pinned_array(dims::Int...)
    ptr = ccall((:sdde_b200_alloc_pinned, libsdde), Ptr{Float64}, (Int64,), 8 * prod(dims))
    ptr == C_NULL && error("libsdde_b200: " * last_error())
    a = unsafe_wrap(Array, ptr, dims; own = false)
    finalizer(_ -> ccall((:sdde_b200_free_pinned, libsdde), Cvoid, (Ptr{Float64},), ptr), a)
    return a
end

function SciMLBase.__solve(ep::SciMLBase.AbstractEnsembleProblem, alg::Union{EM, RKMil, RKMilCommute, EulerHeun}, ea::EnsembleB200;
        trajectories, dt, saveat = Float64[], tstops = Float64[], adaptive = false, seed = UInt64(0), moments = false,
        save_idxs = nothing, noise_increments = nothing, pinned = true, kwargs...)
    prob = ep.prob
    prob isa SciMLBase.AbstractSDDEProblem || error("EnsembleB200 handles SDDEProblems only")
    adaptive && error("EnsembleB200 is a fixed-step path: pass adaptive=false")          # src/solve.jl:40
    isempty(kwargs) || error("unsupported keyword(s) on the EnsembleB200 path: $(keys(kwargs))")
    (ep.output_func === SciMLBase.DEFAULT_OUTPUT_FUNC && ep.reduction === SciMLBase.DEFAULT_REDUCTION) ||
        error("output_func / reduction run on the host in SciML; on this path use saveat + moments")
    m = prob.f.f::B200Model
    t0, tf = prob.tspan
    # saveat expansion exactly as tstop_saveat_disc_handling does it (src/solve.jl:659-677)
    sv = saveat isa Number ? collect(((t0:saveat:tf)[end] == tf) ? ((t0 + saveat):saveat:tf) : ((t0 + saveat):saveat:(tf - saveat))) :
         Float64[t for t in saveat if t0 < t < tf]
    save_start = saveat isa Number || isempty(saveat) || t0 in saveat
    save_end = saveat isa Number || isempty(saveat) || tf in saveat
    ts_user = collect(Float64, tstops)
    # affine history h(p,t) = c0 + c1 t (checked at a third point)
    h1 = prob.h(prob.p, t0 - 1.0); h2 = prob.h(prob.p, t0 - 2.0); h3 = prob.h(prob.p, t0 - 3.5)
    c1 = collect(Float64, h1 .- h2); c0 = collect(Float64, h1 .- c1 .* (t0 - 1.0))
    all(isapprox.(c0 .+ c1 .* (t0 - 3.5), h3; rtol = 1e-13, atol = 1e-13)) ||
        error("only affine history functions cross the C ABI for built-in models (user models may define h in CUDA source)")
    # EnsembleProblem.prob_func: a trajectory may differ in p and u0 (called on the host for every i, like the reference)
    per_traj = ep.prob_func !== SciMLBase.DEFAULT_PROB_FUNC
    if per_traj
        P = Array{Float64, 2}(undef, trajectories, m.n_param)           # == C layout [n_param][n_traj]
        U0 = Array{Float64, 2}(undef, trajectories, m.n_state)
        for i in 1:trajectories
            q = ep.prob_func(deepcopy(prob), i, 1)
            (q.tspan == prob.tspan && q.constant_lags == prob.constant_lags) || error("prob_func may change p and u0 only on this path")
            P[i, :] .= q.p; U0[i, :] .= q.u0
        end
        p, u0 = P, U0
    else
        p = collect(Float64, prob.p); u0 = collect(Float64, prob.u0 isa Number ? [prob.u0] : prob.u0)
    end
    idxs = save_idxs === nothing ? Int32[] : Int32[i - 1 for i in save_idxs]          # 0-based on the C side
    n_saved = isempty(idxs) ? Int(m.n_state) : length(idxs)
    nz = noise_increments === nothing ? Float64[] : noise_increments        # Array{Float64,3}(n_traj, n_chan, n_steps) == C [n][c][traj]
    d = SddeDesc(; struct_size = sizeof(SddeDesc), model_id = m.id, alg = alg_id(alg), n_state = m.n_state,
        n_noise = m.n_noise, n_lag = m.n_lag, n_param = m.n_param,
        order_discontinuity_t0 = prob.order_discontinuity_t0, constant_lags_declared = isempty(prob.constant_lags) ? 0 : 1,
        save_start, save_end, save_everystep = (isempty(sv) && !(saveat isa Number)) ? 1 : 0,
        want_moments = moments ? 1 : 0, n_saveat = length(sv), n_tstops = length(ts_user), t0, tf, dt,
        lags = ntuple(i -> i <= length(prob.constant_lags) ? Float64(prob.constant_lags[i]) : 0.0, 4),
        hist_c0 = ntuple(i -> i <= length(c0) ? c0[i] : 0.0, 8), hist_c1 = ntuple(i -> i <= length(c1) ? c1[i] : 0.0, 8),
        params = pointer(p), param_stride = per_traj ? trajectories : 0, u0 = pointer(u0), u0_stride = per_traj ? trajectories : 0,
        saveat = pointer(sv), tstops = pointer(ts_user), n_traj = trajectories,
        noise_mode = isempty(nz) ? 0 : 1, noise = isempty(nz) ? C_NULL : pointer(nz),
        n_save_idxs = length(idxs), save_idxs = ntuple(i -> i <= length(idxs) ? idxs[i] : Int32(0), 8),
        seed = seed == 0 ? (prob.seed == 0 ? rand(UInt64) : prob.seed) : seed)                # src/solve.jl:302
    n_save = length(sv) + save_start + (save_end && (isempty(sv) || sv[end] != tf))
    # == C layout [n_save][n_saved][n_traj]; page-locked so that the kernel writes it directly
    out = pinned ? pinned_array(trajectories, n_saved, n_save) : Array{Float64, 3}(undef, trajectories, n_saved, n_save)
    mom = zeros(Float64, 3, n_saved, n_save)
    status = zeros(UInt8, trajectories)
    rc = GC.@preserve p u0 sv ts_user nz out mom status ccall((:sdde_b200_solve, libsdde), Cint,
        (Ptr{Cvoid}, Ref{SddeDesc}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Cvoid}),
        context(ea.device), d, out, moments ? pointer(mom) : C_NULL, status, C_NULL)
    rc == 0 || error("libsdde_b200 ($rc): " * last_error())
    ts = vcat(save_start ? [t0] : Float64[], sv, (save_end && (isempty(sv) || sv[end] != tf)) ? [tf] : Float64[])
    # lazy per-trajectory solutions: views into `out`, built with the reference's own build_solution
    sols = [SciMLBase.build_solution(prob, alg, ts, [view(out, i, :, k) for k in 1:n_save];
                retcode = status[i] == 0 ? ReturnCode.Success : ReturnCode.Unstable) for i in 1:trajectories]
    sim = SciMLBase.EnsembleSolution(sols, 0.0, true)
    return moments ? (sim, mom) : sim          # mom[:, j, k] = (count, mean, M2) of state j at ts[k]
end

"""
    b200_quantiles(ea, out, probs; median = true) -> Array{Float64,3}  (n_col, n_state, n_save)

Median and quantile bands across trajectories of the `out` array of an EnsembleB200 solve (host memory here), selected
exactly on the device: what `EnsembleSummary(sim; quantiles)` computes with `timeseries_steps_quantile` /
`timeseries_steps_median` (Statistics.quantile / Statistics.median per time point), without materialising the
componentwise vectors.  Columns: one per probability, the median last.
"""
function b200_quantiles(ea::EnsembleB200, out::Array{Float64, 3}, probs::Vector{Float64}; median::Bool = true)
    n_traj, n_state, n_save = size(out)
    res = Array{Float64, 3}(undef, length(probs) + median, n_state, n_save)
    rc = GC.@preserve out probs res ccall((:sdde_b200_quantiles, libsdde), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Cint, Int64, Int64, Int64, Ptr{Float64}, Cint, Cint, Ptr{Float64}, Ptr{Float64}),
        context(ea.device), out, 0, n_state * n_save, n_traj, n_traj, probs, length(probs), median ? 1 : 0, res, C_NULL)
    rc == 0 || error("libsdde_b200 ($rc): " * last_error())
    return res
end

end # module
