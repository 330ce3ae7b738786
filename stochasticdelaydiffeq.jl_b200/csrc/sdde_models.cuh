// sdde_models.cuh -- built-in model zoo (device code).
//
// Each model is the device-side form of what the reference builds at
// /root/reference/src/functionwrapper.jl:8-9,35-36: f(u,p,t) := f(u,h,p,t), g(u,p,t) := g(u,h,p,t), where the
// history function h has already been evaluated at t - lags[l] for every constant lag (ul[l][*]); that value is
// the same for every f/g call inside one step.
//
// Interface (also the contract for NVRTC user models, see sdde_b200_compile_model):
//   NS states, NL lags, NW noise channels, NP parameters, DIAG (diagonal noise: gv[NS]; else gv[NS*NW] row-major)
//   f(du, u, ul, p, t), g(gv, u, ul, p, t);  p[i] is either a kernel-argument constant or a per-trajectory register.
// This header is NVRTC-clean (no host includes).
#pragma once

namespace sdde {

// x^n, right-to-left binary exponentiation; the multiplication order is part of the model definition
// (the oracle uses the same order).
__device__ __forceinline__ double ipow(double x, int n) {
    if (n == 10) {  // same products, same order as the loop below (x^2 * x^8), without the loop overhead
        const double x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
        return x2 * x8;
    }
    // __dmul_rn: the products are never contracted into a following add (e.g. 1 + x^n), whatever shape the compiler gives
    // this loop in a particular kernel instantiation -- every batch body then rounds alike
    double result = 1.0, base = x;
    bool first = true;
    while (n > 0) {
        if (n & 1) { result = first ? base : __dmul_rn(result, base); first = false; }
        n >>= 1;
        if (n) base = __dmul_rn(base, base);
    }
    return result;
}

// n / d for a normal, positive, well-scaled denominator (here d = 1 + x^n >= 1): MUFU.RCP64H seed (~2^-20), one cubic
// (Halley) step y <- y (1 + e + e^2), e = 1 - d y  (relative error ~2^-60), then q = n y.  Error <= 0.5 ulp + 2^-7 ulp;
// no special-case branch, unlike the IEEE '/' operator's slow path.
__device__ __forceinline__ double div_pos(double n, double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    y = fma(y, t, y);
    return n * y;
}

// n / d for a denominator that is a positive NORMAL number well inside the exponent range (here: a conserved population
// N = S + I + R): reciprocal seed, cubic refinement (|1 - d y| ~ 2^-60), quotient, Markstein residual correction -- the
// IEEE quotient up to rare last-bit cases (<= 0.5 ulp + 2^-8 ulp), without the operator's special-case checks and
// branches (11 BRA + 5.7 BSSY/BSYNC per SIR step in profiles/r2h_sir_summary.txt).  d = 0 / inf / NaN give NaN: the
// trajectory is flagged Unstable exactly as with the IEEE operator's inf / NaN.
__device__ __forceinline__ double div_normal(double n, double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    y = fma(y, fma(e, e, e), y);
    const double q = n * y;
    const double r = fma(-d, q, n);
    return fma(r, y, q);
}

// sqrt(max(0, x)) as test/nondiagonal_sparse_noise.jl:33 writes it: zero for x <= 0 (and NaN, like fmax(0, NaN)), else
// MUFU.RSQ64H seed, one cubic step and a residual correction (the IEEE root up to rare last-bit cases); arguments below
// 1e-300 count as zero (error < 1e-150).  No special-case branches.
__device__ __forceinline__ double sqrt_clamped(double x) {
    const double xs = fmax(x, 1e-300);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(xs));
    const double r0 = xs * y;
    const double e = fma(-r0, y, 1.0);
    double r = fma(r0 * e, fma(e, 0.375, 0.5), r0);
    r = fma(fma(-r, r, xs), 0.5 * y, r);
    return x > 1e-300 ? r : 0.0;
}

// du = (a u + b u(t-tau)) dt + s u dW              BASELINE.json configs[0], configs[4]
struct ModelLinear1 {
    static constexpr int NS = 1, NL = 1, NW = 1, NP = 3;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double) {
        du[0] = fma(p[1], ul[0][0], p[0] * u[0]);      // explicit contraction: every batch body of the kernel rounds alike
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*)[NS], const P &p, double) {
        gv[0] = p[2] * u[0];
    }
};

// Hayes equation of the reference tests: test/test_prob_sol.jl:6-13, test/analyticless_convergence_tests.jl:6-13 (beta = 0)
struct ModelHayes {
    static constexpr int NS = 1, NL = 1, NW = 1, NP = 7;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double) {
        du[0] = fma(p[2], ul[0][0], p[1] * u[0]) + p[3];
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*ul)[NS], const P &p, double) {
        gv[0] = fma(p[5], ul[0][0], p[4] * u[0]) + p[6];
    }
};

// stochastic Mackey-Glass: BASELINE.json configs[1], configs[2]
// NEXP > 0: exponent fixed at compile time (the library picks NEXP = 10 when p[2] == 10 is shared by all trajectories)
template <int NEXP>
struct ModelMackeyGlassT {
    static constexpr int NS = 1, NL = 1, NW = 1, NP = 4;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double) {
        const double x = ul[0][0];
        const double xn = ipow(x, NEXP > 0 ? NEXP : (int)p[2]);
        // 1 + x^n >= 1 whenever n is even or x >= 0: branch-free division; otherwise the IEEE operator
        if (NEXP > 0 && (NEXP % 2) == 0) du[0] = fma(-p[1], u[0], div_pos(p[0] * x, 1.0 + xn));   // explicit contraction
        else du[0] = (p[0] * x) / (1.0 + xn) - p[1] * u[0];
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*)[NS], const P &p, double) {
        gv[0] = p[3] * u[0];
    }
};
typedef ModelMackeyGlassT<0> ModelMackeyGlass;
typedef ModelMackeyGlassT<10> ModelMackeyGlass10;

// n-state linear delay system with n x m noise: BASELINE.json configs[3]
//   du = (A u + sum_l B_l u(t-tau_l) + c) dt + G(u) dW,  G_ij = M_ij u_i + N_ij
//   p = A[NS*NS], B_0..B_{NL-1}[NS*NS], c[NS], M[NS*NW], N[NS*NW]   (row-major)
template <int NS_, int NL_, int NW_>
struct ModelLinSys {
    static constexpr int NS = NS_, NL = NL_, NW = NW_, NP = NS_ * NS_ * (1 + NL_) + NS_ + 2 * NS_ * NW_;
    static constexpr bool DIAG = false;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double acc = p[i * NS] * u[0];
#pragma unroll
            for (int j = 1; j < NS; ++j) acc = acc + p[i * NS + j] * u[j];
#pragma unroll
            for (int l = 0; l < NL; ++l)
#pragma unroll
                for (int j = 0; j < NS; ++j) acc = acc + p[NS * NS * (1 + l) + i * NS + j] * ul[l][j];
            du[i] = acc + p[NS * NS * (1 + NL) + i];
        }
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*)[NS], const P &p, double) {
        constexpr int offM = NS * NS * (1 + NL) + NS, offN = offM + NS * NW;
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int j = 0; j < NW; ++j) gv[i * NW + j] = p[offM + i * NW + j] * u[i] + p[offN + i * NW + j];
    }
    // row-wise interface for the cooperative-lanes kernel (sdde_coop.cuh): row i of f and of G, same operation order
    static constexpr bool HAS_ROWS = true;
    template <class P> __device__ __forceinline__ static double f_row(int i, const double *u, const double (*ul)[NS], const P &p, double) {
        double acc = p[i * NS] * u[0];
#pragma unroll
        for (int j = 1; j < NS; ++j) acc = acc + p[i * NS + j] * u[j];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int j = 0; j < NS; ++j) acc = acc + p[NS * NS * (1 + l) + i * NS + j] * ul[l][j];
        return acc + p[NS * NS * (1 + NL) + i];
    }
    template <class P> __device__ __forceinline__ static void g_row(int i, double *grow, const double *, double ui, const double (*)[NS], const P &p, double) {
        constexpr int offM = NS * NS * (1 + NL) + NS, offN = offM + NS * NW;
#pragma unroll
        for (int j = 0; j < NW; ++j) grow[j] = p[offM + i * NW + j] * ui + p[offN + i * NW + j];
    }
};

// delayed SIR with 3x2 sparse noise: test/nondiagonal_sparse_noise.jl:5-43
struct ModelSIR {
    static constexpr int NS = 3, NL = 1, NW = 2, NP = 3;
    static constexpr bool DIAG = false;
    template <class P> __device__ __forceinline__ static void rates(double &infection, double &recovery, const double *u,
                                                                    const double (*ul)[NS], const P &p) {
        const double N = (u[0] + u[1]) + u[2];
#ifdef SDDE_SIR_IEEE
        infection = ((p[0] * p[1]) * u[1]) / N * u[0];
#else
        infection = div_normal((p[0] * p[1]) * u[1], N) * u[0];       // N = S + I + R is conserved by drift and noise columns
#endif
        const double Nd = (ul[0][0] + ul[0][1]) + ul[0][2];
#ifdef SDDE_SIR_IEEE
        recovery = ((p[0] * p[1]) * ul[0][1]) / Nd * ul[0][0];
#else
        recovery = div_normal((p[0] * p[1]) * ul[0][1], Nd) * ul[0][0];
#endif
    }
    template <class P> __device__ __forceinline__ static void f(double *du, const double *u, const double (*ul)[NS], const P &p, double) {
        double inf, rec;
        rates(inf, rec, u, ul, p);
        du[0] = -inf; du[1] = inf - rec; du[2] = rec;
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *u, const double (*ul)[NS], const P &p, double) {
        double inf, rec;
        rates(inf, rec, u, ul, p);
#ifdef SDDE_SIR_IEEE
        const double si = sqrt(fmax(0.0, inf)), sr = sqrt(fmax(0.0, rec));
#else
        const double si = sqrt_clamped(inf), sr = sqrt_clamped(rec);
#endif
        gv[0] = -si; gv[1] = 0.0;
        gv[2] = si;  gv[3] = -sr;
        gv[4] = 0.0; gv[5] = sr;
    }
};

// scalar additive-noise delay equation: test/events.jl:4-5
struct ModelScalarAdd {
    static constexpr int NS = 1, NL = 1, NW = 1, NP = 2;
    static constexpr bool DIAG = true;
    template <class P> __device__ __forceinline__ static void f(double *du, const double *, const double (*ul)[NS], const P &p, double) {
        du[0] = p[0] * ul[0][0];
    }
    template <class P> __device__ __forceinline__ static void g(double *gv, const double *, const double (*)[NS], const P &p, double) {
        gv[0] = p[1];
    }
};

}  // namespace sdde
