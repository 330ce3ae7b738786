# replay_fixtures.jl -- run the REFERENCE (StochasticDelayDiffEq.jl's own CPU path) on the parity fixtures of this
# repository and print how far the oracle's solution is from it.  This is the script that pins the oracle:
#
#     python tests/golden/export_for_julia.py /tmp/sdde_fixtures           # in this repository (needs no GPU)
#     julia --project=<env with StochasticDelayDiffEq, DiffEqNoiseProcess> julia/replay_fixtures.jl /tmp/sdde_fixtures
#
# NOT RUNNABLE IN THIS REPOSITORY'S CONTAINER (no Julia).  For every fixture and trajectory it builds the SAME SDDEProblem a
# user would (f, g from SDDEB200.B200Models: plain Julia functions), hands the reference the fixture's Brownian path as a
# NoiseGrid on the schedule's nodes (the prob.noise branch of src/solve.jl:340-353), solves with the fixture's algorithm,
# fixed dt and save_everystep = true (dense history == saved solution, SURVEY.md H1), and compares
#   * the reference's accepted time points sol.t with the oracle's schedule nodes (the [EXT] facts behind the host
#     schedule: tstop snapping, discontinuity propagation, alg_order),
#   * every state sol.u[n] with the oracle's dense solution (the [EXT] step formulas: EM split default, RKMil Ito /
#     Stratonovich, EulerHeun stage time, history interpolation regimes).
# Expected: schedule identical, max relative error <= 1e-12 (the bound the GPU path is held to against the oracle).
# A larger number names the [EXT] recollection that is wrong (SURVEY.md Appendix B) -- fix the oracle, regenerate the
# fixtures, re-run the GPU parity tests.

using StochasticDelayDiffEq, StochasticDiffEq, DiffEqNoiseProcess, TOML, Printf

include(joinpath(@__DIR__, "EnsembleB200.jl"))
using .SDDEB200: B200Models

dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "julia_export")
man = TOML.parsefile(joinpath(dir, "manifest.toml"))
n_traj = man["n_traj"]

readf64(path, dims...) = (a = Array{Float64}(undef, dims...); read!(path, a); a)

function algorithm(name, opt)
    name == "EM" && return EM(opt == "true")
    name == "RKMil" && return RKMil(interpretation = Symbol(opt))
    name == "EulerHeun" && return EulerHeun()
    startswith(name, "WangLi3SMil_") && return getfield(StochasticDiffEq, Symbol(name))()
    name in ("SRA1", "SRA3", "SOSRA") && return getfield(StochasticDiffEq, Symbol(name))()
    error("unknown algorithm $name")
end

function model(name, lags)
    name == "mackey_glass" && return B200Models.mackey_glass(lags[1]), nothing
    name == "linear1" && return B200Models.linear1(lags[1]), nothing
    name == "hayes" && return B200Models.hayes(), nothing
    name == "linsys_4_2_2" && return B200Models.linsys(4, (lags[1], lags[2]), 2), zeros(4, 2)
    error("unknown model $name")
end

worst = 0.0
for (name, fx) in sort(collect(filter(kv -> kv[2] isa Dict, man)); by = first)
    nn, ns, nc = fx["n_nodes"], fx["n_state"], fx["n_chan"]
    t = readf64(joinpath(dir, name * "_t.bin"), nn)
    W = readf64(joinpath(dir, name * "_W.bin"), nn, nc, n_traj)
    U = readf64(joinpath(dir, name * "_dense.bin"), nn, ns, n_traj)
    has_z = get(fx, "has_z", 0) == 1                    # SRA methods: second Brownian path, NoiseGrid(t, W, Z)
    Z = has_z ? readf64(joinpath(dir, name * "_Z.bin"), nn, nc, n_traj) : nothing
    (f, g), nrp = model(fx["model"], fx["lags"])
    alg = algorithm(fx["alg"], fx["alg_option"])
    scalar = ns == 1
    u0 = scalar ? fx["u0"][1] : collect(Float64, fx["u0"])
    h0 = scalar ? fx["h0"][1] : collect(Float64, fx["h0"])
    h = (p, t) -> h0
    err = 0.0; terr = 0.0; nmis = 0
    for i in 1:n_traj
        Wi = scalar && nc == 1 ? W[:, 1, i] : [W[n, :, i] for n in 1:nn]
        Zi = !has_z ? nothing : (scalar && nc == 1 ? Z[:, 1, i] : [Z[n, :, i] for n in 1:nn])
        noise = has_z ? NoiseGrid(t, Wi, Zi) : NoiseGrid(t, Wi)
        kw = nrp === nothing ? (;) : (; noise_rate_prototype = nrp)
        prob = SDDEProblem(f, g, u0, h, (fx["tspan"][1], fx["tspan"][2]), Tuple(fx["p"]); constant_lags = Tuple(fx["lags"]),
                           order_discontinuity_t0 = fx["order_discontinuity_t0"], noise = noise, kw...)
        sol = solve(prob, alg; dt = fx["dt"], adaptive = false, save_everystep = true)
        if length(sol.t) != nn
            nmis += 1
            continue
        end
        terr = max(terr, maximum(abs.(sol.t .- t)))
        for n in 1:nn
            un = scalar ? [sol.u[n]] : sol.u[n]
            err = max(err, maximum(abs.(un .- U[n, :, i]) ./ max.(abs.(U[n, :, i]), 1e-300)))
        end
    end
    global worst = max(worst, err)
    @printf("%-18s %-9s nodes %5d  schedule: %s (max |dt| %.1e)   max rel. error vs reference %.3e   %s\n", name, fx["alg"], nn,
            nmis == 0 ? "same length" : "LENGTH DIFFERS on $nmis paths", terr, err, (nmis == 0 && terr == 0.0 && err <= 1e-12) ? "PINNED" : "CHECK")
end
@printf("worst max rel. error over all fixtures: %.3e  (bound 1e-12)\n", worst)
