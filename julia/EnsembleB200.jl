# EnsembleB200.jl -- the reference-side binding a StochasticDelayDiffEq.jl maintainer would add.
#
# NOT RUNNABLE IN THIS REPOSITORY'S CONTAINER (no Julia in the image).  What keeps it honest without Julia:
# tests/test_abi_host.py::test_julia_struct_mirrors_match_header parses the three struct mirrors below and compares field
# names, order, types and sizes with include/sdde_b200.h, and checks that every `ccall` names an exported symbol with the
# header's argument count.
#
# It adds ONE method to the SciML ensemble dispatch ([EXT] SciMLBase.__solve(::AbstractEnsembleProblem, alg, ensemblealg;
# trajectories, ...)): when the third positional argument of `solve` is `EnsembleB200()`, the per-trajectory loop over
# DiffEqBase.__solve(prob::AbstractSDDEProblem, alg) (src/solve.jl:1-9) is replaced by one ccall into libsdde_b200.so.
#
#   using StochasticDelayDiffEq; include("EnsembleB200.jl"); using .SDDEB200
#   f, g  = B200Models.mackey_glass(17.0)                      # device-model handles that are ALSO plain Julia functions
#   prob  = SDDEProblem(f, g, 1.2, (p, t) -> 1.2, (0.0, 500.0), (0.2, 0.1, 10.0, 0.1); constant_lags = (17.0,))
#   sim   = solve(EnsembleProblem(prob), EM(), EnsembleB200(); dt = 0.05, saveat = 1.0, trajectories = 10^6)
#   cpu   = solve(EnsembleProblem(prob), EM(), EnsembleThreads(); dt = 0.05, saveat = 1.0, trajectories = 100)   # same problem object
module SDDEB200

using SciMLBase, StochasticDiffEq, StochasticDelayDiffEq

export EnsembleB200, B200Models, FinalState, SaveIdxs, MomentsReduction, compile_model, pinned_array, b200_quantiles, b200_moments, check_sra_tableaus

const libsdde = get(ENV, "LIBSDDE_B200", "libsdde_b200.so")

"""
    EnsembleB200(device = 0)            one GPU
    EnsembleB200([0, 1, 2, 3])          several GPUs of this node from ONE Julia process (sdde_b200_solve_multi)
"""
struct EnsembleB200 <: SciMLBase.EnsembleAlgorithm
    devices::Vector{Cint}
end
EnsembleB200() = EnsembleB200(Cint[0])
EnsembleB200(device::Integer) = EnsembleB200(Cint[device])

# ------------------------------------------------------------------------------------------------------------------
# struct mirrors of include/sdde_b200.h (field for field; guarded by tests/test_abi_host.py)
# ------------------------------------------------------------------------------------------------------------------
Base.@kwdef mutable struct SddeDesc
    struct_size::Int32 = 0
    model_id::Int32 = 0
    alg::Int32 = 0
    n_state::Int32 = 1
    n_noise::Int32 = 1
    n_lag::Int32 = 1
    n_param::Int32 = 0
    order_discontinuity_t0::Int32 = 0
    constant_lags_declared::Int32 = 1
    save_start::Int32 = 1
    save_end::Int32 = 1
    save_everystep::Int32 = 0
    noise_mode::Int32 = 0
    noise_mem::Int32 = 0
    out_mem::Int32 = 0
    history_mode::Int32 = 0
    want_paths::Int32 = 1
    want_moments::Int32 = 0
    lanes_per_traj::Int32 = 0
    n_saveat::Int32 = 0
    n_tstops::Int32 = 0
    reserved0::Int32 = 0
    t0::Float64 = 0
    tf::Float64 = 0
    dt::Float64 = 0
    lags::NTuple{4, Float64} = (0.0, 0.0, 0.0, 0.0)
    hist_c0::NTuple{8, Float64} = ntuple(_ -> 0.0, 8)
    hist_c1::NTuple{8, Float64} = ntuple(_ -> 0.0, 8)
    params::Ptr{Float64} = C_NULL
    param_stride::Int64 = 0
    u0::Ptr{Float64} = C_NULL
    u0_stride::Int64 = 0
    saveat::Ptr{Float64} = C_NULL
    tstops::Ptr{Float64} = C_NULL
    noise::Ptr{Float64} = C_NULL
    n_traj::Int64 = 0
    traj_offset::Int64 = 0
    seed::UInt64 = 0
    out_pitch::Int64 = 0
    n_save_idxs::Int32 = 0
    save_idxs::NTuple{8, Int32} = ntuple(_ -> Int32(0), 8)
    reserved1::Int32 = 0
    noise_t::Ptr{Float64} = C_NULL
    n_noise_t::Int64 = 0
    d_disc_t::Ptr{Float64} = C_NULL
    d_disc_order2::Ptr{Int32} = C_NULL
    n_d_disc::Int64 = 0
    out_rows::Int64 = 0
    sra_tableau::Ptr{Float64} = C_NULL
    sra_stages::Int32 = 0
    sra_order2::Int32 = 0
end

Base.@kwdef mutable struct SddeSchedule
    n_steps::Int64 = 0
    n_save::Int64 = 0
    n_lag::Int32 = 0
    ring_slots::Int32 = 0
    min_gap::Int32 = 0
    n_tstops_hit::Int32 = 0
    n_stage::Int32 = 0
    reserved::Int32 = 0
    t::Ptr{Float64} = C_NULL
    dt::Ptr{Float64} = C_NULL
    sqdt::Ptr{Float64} = C_NULL
    lk_mode::Ptr{Int32} = C_NULL
    lk_idx::Ptr{Int64} = C_NULL
    lk_theta::Ptr{Float64} = C_NULL
    save_t::Ptr{Float64} = C_NULL
    save_step::Ptr{Int64} = C_NULL
    save_theta::Ptr{Float64} = C_NULL
    save_exact::Ptr{Int32} = C_NULL
    tstops_hit::Ptr{Float64} = C_NULL
end

Base.@kwdef mutable struct SddeTiming
    kernel_ms::Float64 = 0
    total_ms::Float64 = 0
    h2d_ms::Float64 = 0
    d2h_ms::Float64 = 0
    h2d_bytes::Int64 = 0
    d2h_bytes::Int64 = 0
    n_steps::Int64 = 0
    n_save::Int64 = 0
    launches::Int32 = 0
    waves::Int32 = 0
    ring_slots::Int32 = 0
    regs_per_thread::Int32 = 0
    block_threads::Int32 = 0
    traj_per_thread::Int32 = 0
    moment_merge::Int32 = 0
    n_unstable::Int32 = 0
    ring_bytes::Int64 = 0
    kernel_name::NTuple{128, UInt8} = ntuple(_ -> 0x00, 128)
end

# ------------------------------------------------------------------------------------------------------------------
# device-model handles: f and g of an SDDEProblem that are (a) ordinary Julia functions -- the SAME problem object runs
# on the reference's CPU path, which is also how parity against the real reference is pinned (julia/replay_fixtures.jl)
# -- and (b) carry the id of the device model that libsdde_b200 runs in their place (Julia closures cannot cross a C ABI;
# arbitrary models go through `compile_model` with CUDA source).
# ------------------------------------------------------------------------------------------------------------------
struct ModelInfo
    id::Int32
    n_state::Int32
    n_lag::Int32
    n_noise::Int32
    n_param::Int32
    diagonal::Bool
end

struct B200Fn{F} <: Function
    info::ModelInfo
    role::Symbol        # :f or :g
    cpu::F
end
(fn::B200Fn)(args...) = fn.cpu(args...)      # out-of-place (u, h, p, t) and in-place (du, u, h, p, t) forms alike

module B200Models
using ..SDDEB200: ModelInfo, B200Fn

# x^n by right-to-left binary exponentiation, the multiplication order of the device models (csrc/sdde_models.cuh `ipow`)
function ipow(x, n::Integer)
    n == 10 && return (x2 = x * x; x4 = x2 * x2; x8 = x4 * x4; x2 * x8)
    result = one(x); base = x; first = true
    while n > 0
        if isodd(n)
            result = first ? base : result * base
            first = false
        end
        n >>= 1
        n > 0 && (base = base * base)
    end
    return result
end

pair(info, f, g) = (B200Fn(info, :f, f), B200Fn(info, :g, g))

"du = (a u + b u(t-lag)) dt + s u dW, p = (a, b, s)   (BASELINE configs 1 and 5; SDDE_MODEL_LINEAR1)"
linear1(lag) = pair(ModelInfo(0, 1, 1, 1, 3, true),
    (u, h, p, t) -> p[1] * u + p[2] * h(p, t - lag),
    (u, h, p, t) -> p[3] * u)

"Hayes equation of test/test_prob_sol.jl:6-13, p = (tau, a, b, c, alpha, beta, gamma)   (SDDE_MODEL_HAYES)"
hayes() = pair(ModelInfo(1, 1, 1, 1, 7, true),
    (u, h, p, t) -> (p[2] * u + p[3] * h(p, t - p[1])) + p[4],
    (u, h, p, t) -> (p[5] * u + p[6] * h(p, t - p[1])) + p[7])

"stochastic Mackey-Glass, p = (beta, gamma, n, sigma)   (BASELINE configs 2 and 3; SDDE_MODEL_MACKEY_GLASS)"
mackey_glass(lag) = pair(ModelInfo(2, 1, 1, 1, 4, true),
    (u, h, p, t) -> (x = h(p, t - lag); (p[1] * x) / (1.0 + ipow(x, Int(p[3]))) - p[2] * u),
    (u, h, p, t) -> p[4] * u)

"""
n-state linear delay system with an n x m noise matrix, p = (A, B_1..B_L, c, M, N) row-major (csrc ModelLinSys):
du = (A u + sum_l B_l u(t - lag_l) + c) dt + G(u) dW, G_ij = M_ij u_i + N_ij   (BASELINE config 4; SDDE_MODEL_LINSYS).
In-place forms; pass `noise_rate_prototype = zeros(ns, nw)` to SDDEProblem.
"""
function linsys(ns::Int, lags::NTuple{L, Float64}, nw::Int) where {L}
    np = ns * ns * (1 + L) + ns + 2 * ns * nw
    f! = function (du, u, h, p, t)
        uls = ntuple(l -> h(p, t - lags[l]), L)
        for i in 1:ns
            acc = p[(i - 1) * ns + 1] * u[1]
            for j in 2:ns
                acc = acc + p[(i - 1) * ns + j] * u[j]
            end
            for l in 1:L, j in 1:ns
                acc = acc + p[ns * ns * l + (i - 1) * ns + j] * uls[l][j]
            end
            du[i] = acc + p[ns * ns * (1 + L) + i]
        end
        return nothing
    end
    g! = function (du, u, h, p, t)
        offM = ns * ns * (1 + L) + ns; offN = offM + ns * nw
        for i in 1:ns, j in 1:nw
            du[i, j] = p[offM + (i - 1) * nw + j] * u[i] + p[offN + (i - 1) * nw + j]
        end
        return nothing
    end
    return pair(ModelInfo(3, ns, L, nw, np, false), f!, g!)
end

"delayed SIR with 3 x 2 sparse noise of test/nondiagonal_sparse_noise.jl:5-43, p = (beta, c, tau)   (SDDE_MODEL_SIR)"
function sir()
    rates = function (u, h, p, t)
        hist = h(p, t - p[3])
        N = (u[1] + u[2]) + u[3]
        infection = ((p[1] * p[2]) * u[2]) / N * u[1]
        Nd = (hist[1] + hist[2]) + hist[3]
        recovery = ((p[1] * p[2]) * hist[2]) / Nd * hist[1]
        return infection, recovery
    end
    f! = function (du, u, h, p, t)
        inf, rec = rates(u, h, p, t)
        du[1] = -inf; du[2] = inf - rec; du[3] = rec
        return nothing
    end
    g! = function (du, u, h, p, t)
        inf, rec = rates(u, h, p, t)
        si = sqrt(max(0.0, inf)); sr = sqrt(max(0.0, rec))
        du[1, 1] = -si; du[1, 2] = 0.0
        du[2, 1] = si;  du[2, 2] = -sr
        du[3, 1] = 0.0; du[3, 2] = sr
        return nothing
    end
    return pair(ModelInfo(4, 3, 1, 2, 3, false), f!, g!)
end

"scalar additive-noise delay equation of test/events.jl:4-5, p = (a, b)   (SDDE_MODEL_SCALAR_ADD)"
scalar_add(lag) = pair(ModelInfo(5, 1, 1, 1, 2, true),
    (u, h, p, t) -> p[1] * h(p, t - lag),
    (u, h, p, t) -> p[2])

end # module B200Models

# the device model behind a problem: SDDEProblem(f, g, ...) stores f inside an SDDEFunction (prob.f.f); tolerate one more
# level of wrapping (FunctionWrappers / a user's own callable struct with an `f` field)
function model_info(prob)
    cand = Any[prob.f]
    for _ in 1:3
        c = cand[end]
        c isa B200Fn && return c.info
        hasproperty(c, :f) || break
        push!(cand, getproperty(c, :f))
    end
    cand[end] isa B200Fn && return cand[end].info
    error("EnsembleB200 needs device-model handles as f and g: build them with SDDEB200.B200Models.* or SDDEB200.compile_model")
end

# ------------------------------------------------------------------------------------------------------------------
# EnsembleProblem.output_func / reduction.  Both are arbitrary Julia closures in SciML and run on the host; the forms
# below are callable structs with the closure's CPU meaning (so the same EnsembleProblem runs under EnsembleThreads) that
# this path recognises and maps onto what the library computes on the device.  Anything else is refused.
# ------------------------------------------------------------------------------------------------------------------
"output_func: keep the final state only  ==  (sol, i) -> (sol.u[end], false)   -> saveat = [tf], save_start = false"
struct FinalState <: Function end
(::FinalState)(sol, i) = (sol.u[end], false)

"output_func: keep some state components  ==  save_idxs keyword (src/solve.jl:355-356)"
struct SaveIdxs <: Function
    idxs::Vector{Int}
end
(o::SaveIdxs)(sol, i) = (SciMLBase.build_solution(sol.prob, sol.alg, sol.t, [u[o.idxs] for u in sol.u]; retcode = sol.retcode), false)

"""
reduction: running (count, mean, M2) per saved time point and state, Chan-merged batch by batch -- what
EnsembleAnalysis.timeseries_steps_meanvar computes from a finished EnsembleSolution.  On this path: want_moments = 1,
want_paths = 0 (nothing but 24 bytes per saved value leaves the GPU); use `u_init = nothing`.
"""
struct MomentsReduction <: Function end
function (::MomentsReduction)(acc, batch, I)
    for sol in batch
        x = reduce(hcat, [vec(collect(u)) for u in sol.u])            # n_state x n_save
        if acc === nothing
            acc = (n = 1.0, mean = x, m2 = zero(x))
        else
            n = acc.n + 1; d = x .- acc.mean
            mean = acc.mean .+ d ./ n
            acc = (n = n, mean = mean, m2 = acc.m2 .+ d .* (x .- mean))
        end
    end
    return acc, false
end

# ------------------------------------------------------------------------------------------------------------------
# plumbing
# ------------------------------------------------------------------------------------------------------------------
alg_id(::EM{true}) = Int32(0)
alg_id(::EM{false}) = Int32(1)
alg_id(a::RKMil) = StochasticDiffEq.alg_interpretation(a) == :Ito ? Int32(2) : Int32(3)
alg_id(a::RKMilCommute) = StochasticDiffEq.alg_interpretation(a) == :Ito ? Int32(5) : Int32(6)   # diagonal models run the RKMil kernel
alg_id(::EulerHeun) = Int32(4)
alg_id(::WangLi3SMil_A) = Int32(7); alg_id(::WangLi3SMil_B) = Int32(8); alg_id(::WangLi3SMil_C) = Int32(9)
alg_id(::WangLi3SMil_D) = Int32(10); alg_id(::WangLi3SMil_E) = Int32(11); alg_id(::WangLi3SMil_F) = Int32(12)
# Roessler SRA methods for additive noise, adaptive = false; the library holds SRA1 / SRA3 / SOSRA, any other explicit tableau
# (SRA(), SRA2(), SOSRA2(), SRA(tableau = ...)) crosses the ABI as numbers: [c0 | c1 | A0 (row = stage) | B0 | alpha | beta1 | beta2]
alg_id(::SRA1) = Int32(13); alg_id(::SRA3) = Int32(14); alg_id(::SOSRA) = Int32(15)
alg_id(::Union{SRA, SRA2, SOSRA2}) = Int32(16)
sra_tableau_of(a::SRA) = a.tableau
sra_tableau_of(::SRA2) = StochasticDiffEq.constructSRA2()
sra_tableau_of(::SOSRA2) = StochasticDiffEq.constructSOSRA2()
function sra_flat(tab)                       # RosslerSRA stores A0', B0' (column = stage): Julia's column-major vec() of them IS the row-per-stage layout
    s = length(tab.α)
    s <= 3 || error("at most 3 SRA stages on this path")
    return Float64[tab.c₀; tab.c₁; vec(tab.A₀); vec(tab.B₀); tab.α; tab.β₁; tab.β₂], Int32(s)
end
"""
    check_sra_tableaus() -> Bool

Compares the Roessler tableaus the library integrates with (restated from memory on the CUDA side) with this session's
StochasticDiffEq `constructSRA1 / constructSRA3 / constructSOSRA`, entry by entry; run it once before trusting an SRA solve.
"""
function check_sra_tableaus(; atol = 0.0)
    ok = true
    for (id, name, tab) in ((Int32(13), "SRA1", StochasticDiffEq.constructSRA1()), (Int32(14), "SRA3", StochasticDiffEq.constructSRA3()),
                            (Int32(15), "SOSRA", StochasticDiffEq.constructSOSRA()))
        mine = zeros(Float64, 33); st = Ref{Int32}(0)
        check(ccall((:sdde_b200_sra_tableau, libsdde), Cint, (Int32, Ptr{Float64}, Ref{Int32}), id, mine, st))
        ref, s = sra_flat(tab)
        same = st[] == s && all(abs.(mine[1:length(ref)] .- ref) .<= atol)
        same || @warn "tableau of $name differs from StochasticDiffEq's" library = mine[1:length(ref)] reference = ref
        ok &= same
    end
    return ok
end

const B200Alg = Union{EM, RKMil, RKMilCommute, EulerHeun, WangLi3SMil_A, WangLi3SMil_B, WangLi3SMil_C, WangLi3SMil_D, WangLi3SMil_E, WangLi3SMil_F,
                      SRA, SRA1, SRA2, SRA3, SOSRA, SOSRA2}

last_error() = unsafe_string(ccall((:sdde_b200_last_error, libsdde), Cstring, ()))
check(rc) = rc == 0 || error("libsdde_b200 ($rc): " * last_error())

const CTX = Dict{Cint, Ptr{Cvoid}}()
function context(dev)
    get!(CTX, dev) do
        h = ccall((:sdde_b200_create, libsdde), Ptr{Cvoid}, (Cint,), dev)
        h == C_NULL && error("libsdde_b200: " * last_error())
        h
    end
end

"""
    compile_model(ea, cuda_source; n_state, n_lag, n_noise, n_param, diagonal, f, g) -> (f, g)

Register a user model from CUDA C++ source (`struct UserModel {...}`, see include/sdde_b200.h) on every GPU of `ea` and
return device-model handles; `f` / `g` are the Julia twins used when the problem runs on the CPU path.
"""
function compile_model(ea::EnsembleB200, source::AbstractString; n_state, n_lag, n_noise, n_param, diagonal, f, g)
    length(ea.devices) == 1 || error("user models are compiled per context: use one GPU per call")
    id = Ref{Int32}(0)
    check(ccall((:sdde_b200_compile_model, libsdde), Cint, (Ptr{Cvoid}, Cstring, Ref{Int32}), context(ea.devices[1]), source, id))
    info = ModelInfo(id[], n_state, n_lag, n_noise, n_param, diagonal)
    return B200Fn(info, :f, f), B200Fn(info, :g, g)
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

# number of saved points and their times, from the library's own schedule replay (never recomputed here: a binding that
# sizes `out` by its own count overflows it as soon as the two disagree, e.g. under save_everystep)
function saved_times(d::SddeDesc)
    s = SddeSchedule()
    check(ccall((:sdde_b200_schedule, libsdde), Cint, (Ref{SddeDesc}, Ref{SddeSchedule}), d, s))
    ts = copy(unsafe_wrap(Array, s.save_t, Int(s.n_save); own = false))
    nsteps = Int(s.n_steps)
    ccall((:sdde_b200_schedule_free, libsdde), Cvoid, (Ref{SddeSchedule},), s)
    return ts, nsteps
end

function SciMLBase.__solve(ep::SciMLBase.AbstractEnsembleProblem, alg::B200Alg, ea::EnsembleB200;
        trajectories, dt, saveat = Float64[], tstops = Float64[], d_discontinuities = (), adaptive = false, seed = UInt64(0),
        moments = false, save_idxs = nothing, save_everystep = isempty(saveat), save_start = nothing, save_end = nothing,
        noise_increments = nothing, pinned = true, kwargs...)
    prob = ep.prob
    prob isa SciMLBase.AbstractSDDEProblem || error("EnsembleB200 handles SDDEProblems only")
    adaptive && error("EnsembleB200 is a fixed-step path: pass adaptive=false")          # src/solve.jl:40
    isempty(kwargs) || error("unsupported keyword(s) on the EnsembleB200 path: $(keys(kwargs))")
    m = model_info(prob)
    t0, tf = prob.tspan
    # ---- output_func / reduction -> what the device computes ----
    want_paths = true
    of, red = ep.output_func, ep.reduction
    if of isa FinalState
        saveat = Float64[tf]; save_everystep = false; save_start = false; save_end = true
    elseif of isa SaveIdxs
        save_idxs = of.idxs
    elseif of !== SciMLBase.DEFAULT_OUTPUT_FUNC
        error("output_func is a Julia closure and cannot run on the device; expressible on this path: FinalState(), SaveIdxs(idxs), or the save_idxs / saveat keywords")
    end
    if red isa MomentsReduction
        moments = true; want_paths = false
    elseif red !== SciMLBase.DEFAULT_REDUCTION
        error("reduction is a Julia closure and cannot run on the device; expressible on this path: MomentsReduction() (mean / variance per time point), b200_quantiles for order statistics")
    end
    # ---- saveat expansion exactly as tstop_saveat_disc_handling does it (src/solve.jl:659-677); flags as __init (:25-34, 411-415) ----
    sv = saveat isa Number ? collect(((t0:saveat:tf)[end] == tf) ? ((t0 + saveat):saveat:tf) : ((t0 + saveat):saveat:(tf - saveat))) :
         Float64[t for t in saveat if t0 < t < tf]
    save_start === nothing && (save_start = save_everystep || isempty(saveat) || saveat isa Number || t0 in saveat)
    save_end === nothing && (save_end = save_everystep || isempty(saveat) || saveat isa Number || tf in saveat)
    ts_user = collect(Float64, tstops)
    dd_t = Float64[d isa Number ? d : d.t for d in d_discontinuities]
    dd_o = Int32[d isa Number ? 0 : Int32(2 * d.order) for d in d_discontinuities]
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
    nz = noise_increments === nothing ? Float64[] : noise_increments        # Array{Float64,3}(n_traj, n_chan, n_steps) == C [n][c][traj]; SRA family: 2 n_chan channels, dW rows first, then dZ
    d = SddeDesc(; struct_size = sizeof(SddeDesc), model_id = m.id, alg = alg_id(alg), n_state = m.n_state,
        n_noise = m.n_noise, n_lag = m.n_lag, n_param = m.n_param,
        order_discontinuity_t0 = prob.order_discontinuity_t0, constant_lags_declared = isempty(prob.constant_lags) ? 0 : 1,
        save_start = save_start ? 1 : 0, save_end = save_end ? 1 : 0, save_everystep = save_everystep ? 1 : 0,
        want_paths = want_paths ? 1 : 0, want_moments = moments ? 1 : 0,
        n_saveat = length(sv), n_tstops = length(ts_user), t0, tf, dt,
        lags = ntuple(i -> i <= length(prob.constant_lags) ? Float64(prob.constant_lags[i]) : 0.0, 4),
        hist_c0 = ntuple(i -> i <= length(c0) ? c0[i] : 0.0, 8), hist_c1 = ntuple(i -> i <= length(c1) ? c1[i] : 0.0, 8),
        params = pointer(p), param_stride = per_traj ? trajectories : 0, u0 = pointer(u0), u0_stride = per_traj ? trajectories : 0,
        saveat = pointer(sv), tstops = pointer(ts_user), n_traj = trajectories,
        noise_mode = isempty(nz) ? 0 : 1, noise = isempty(nz) ? C_NULL : pointer(nz),
        n_save_idxs = length(idxs), save_idxs = ntuple(i -> i <= length(idxs) ? idxs[i] : Int32(0), 8),
        d_disc_t = pointer(dd_t), d_disc_order2 = pointer(dd_o), n_d_disc = length(dd_t),
        seed = seed == 0 ? (prob.seed == 0 ? rand(UInt64) : prob.seed) : seed)                # src/solve.jl:302
    sra_tab = Float64[]
    if alg isa Union{SRA, SRA2, SOSRA2}
        sra_tab, d.sra_stages = sra_flat(sra_tableau_of(alg))
        d.sra_tableau = pointer(sra_tab)
        d.sra_order2 = Int32(2 * StochasticDiffEq.alg_order(alg))
    end
    GC.@preserve p u0 sv ts_user nz dd_t dd_o sra_tab begin
        ts, _ = saved_times(d)                                  # the library's count: never recomputed on this side
        n_save = length(ts)
        d.out_rows = n_save * n_saved                           # the library refuses to write more rows than the buffers hold
        # == C layout [n_save][n_saved][n_traj]; page-locked so that the kernel writes it directly
        out = !want_paths ? Array{Float64, 3}(undef, 0, 0, 0) :
              pinned ? pinned_array(trajectories, n_saved, n_save) : Array{Float64, 3}(undef, trajectories, n_saved, n_save)
        mom = zeros(Float64, 3, n_saved, n_save)
        status = zeros(UInt8, want_paths ? trajectories : 0)
        outp = want_paths ? pointer(out) : Ptr{Float64}(C_NULL)
        momp = moments ? pointer(mom) : Ptr{Float64}(C_NULL)
        stp = want_paths ? pointer(status) : Ptr{UInt8}(C_NULL)
        GC.@preserve out mom status begin
            if length(ea.devices) == 1
                check(ccall((:sdde_b200_solve, libsdde), Cint,
                    (Ptr{Cvoid}, Ref{SddeDesc}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Cvoid}),
                    context(ea.devices[1]), d, outp, momp, stp, C_NULL))
            else
                ctxs = Ptr{Cvoid}[context(dv) for dv in ea.devices]
                check(ccall((:sdde_b200_solve_multi, libsdde), Cint,
                    (Ptr{Ptr{Cvoid}}, Int32, Ref{SddeDesc}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ptr{Cvoid}),
                    ctxs, length(ctxs), d, outp, momp, stp, C_NULL))
            end
        end
    end
    if !want_paths            # MomentsReduction: the reduced value is the ensemble's `u`, as in SciML
        return SciMLBase.EnsembleSolution((t = ts, count = mom[1, :, :], mean = mom[2, :, :], m2 = mom[3, :, :]), 0.0, true)
    end
    # lazy per-trajectory solutions: views into `out`, built with the reference's own build_solution
    sols = [SciMLBase.build_solution(prob, alg, ts, [view(out, i, :, k) for k in 1:n_save];
                retcode = status[i] == 0 ? ReturnCode.Success : ReturnCode.Unstable) for i in 1:trajectories]
    sim = SciMLBase.EnsembleSolution(sols, 0.0, true)
    return moments ? (sim, mom) : sim          # mom[:, j, k] = (count, mean, M2) of state j at ts[k]
end

"""
    b200_moments(ea, out) -> Array{Float64,3}  (3, n_state, n_save)

(count, mean, M2) across trajectories of the `out` array of an earlier solve (EnsembleAnalysis.timeseries_steps_meanvar).
"""
function b200_moments(ea::EnsembleB200, out::Array{Float64, 3})
    n_traj, n_state, n_save = size(out)
    mom = Array{Float64, 3}(undef, 3, n_state, n_save)
    check(GC.@preserve out mom ccall((:sdde_b200_moments, libsdde), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int64, Int64, Ptr{Float64}),
        context(ea.devices[1]), out, 0, n_state * n_save, n_traj, n_traj, mom))
    return mom
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
    check(GC.@preserve out probs res ccall((:sdde_b200_quantiles, libsdde), Cint,
        (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Int64, Int64, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}),
        context(ea.devices[1]), out, 0, n_state * n_save, n_traj, n_traj, probs, length(probs), median ? 1 : 0, res, C_NULL))
    return res
end

end # module
