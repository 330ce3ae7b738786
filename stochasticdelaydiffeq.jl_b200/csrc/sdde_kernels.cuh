// sdde_kernels.cuh -- the sm_100a step kernels of the ensemble SDDE path (device code, NVRTC-clean).
//
// One launch integrates one WAVE of trajectories over the WHOLE time span.  It replaces, for every trajectory of
// the wave at once, the reference's per-trajectory
//     __init            /root/reference/src/solve.jl:11-572      (register init from the descriptor)
//     solve! hot loop   /root/reference/src/solve.jl:574-610
//     loopheader!       /root/reference/src/integrators/interface.jl:1-30    (accept, next dW)
//     perform_step!     [EXT] StochasticDiffEq EM / RKMil, call site src/solve.jl:586
//     loopfooter!       /root/reference/src/integrators/interface.jl:70-85
//     savevalues!       /root/reference/src/integrators/interface.jl:357-430
//     HistoryFunction   [EXT] DelayDiffEq, wired at /root/reference/src/utils.jl:246
//     WienerProcess     [EXT] DiffEqNoiseProcess, set up at /root/reference/src/solve.jl:304-339
//
// Data layout in HBM (all float64):
//   ring   [ring_slots][NS][ring_pitch]   node j of the dense history lives in slot j % ring_slots; trajectory-contiguous
//   out    [n_save][NS][out_pitch]        time-major, trajectory-contiguous (the caller's array when it is on the device)
//   noise  [n][NC][noise_pitch]           supplied increments / grid values (validation modes only)
//   steps  packed StepRec<NL>[n_steps]    trajectory-independent schedule (schedule.cpp), staged through shared memory
//   saves  SaveRec[n_save]
// Thread mapping: one thread integrates TPT adjacent trajectories (TPT = 2 -> every ring/out access is a 128-bit
// double2 access, a warp touches 512 contiguous bytes, and the two independent recurrences give the FP64 pipe ILP).
#pragma once
#include "sdde_rng_tables.cuh"

namespace sdde {

typedef unsigned int u32;
typedef unsigned long long u64;

enum { K_LK_PRE = 0, K_LK_EXACT = 1, K_LK_INTERP = 2 };
enum { K_NOISE_PHILOX = 0, K_NOISE_DW = 1, K_NOISE_GRID = 2 };
enum { K_ALG_EM_SPLIT = 0, K_ALG_EM_NOSPLIT = 1, K_ALG_RKMIL_ITO = 2, K_ALG_RKMIL_STRAT = 3, K_ALG_EULER_HEUN = 4,
       K_ALG_RKMILC_ITO = 5, K_ALG_RKMILC_STRAT = 6,     // RKMilCommute on a noise MATRIX (diagonal models run it as RKMil)
       K_ALG_WANGLI = 7,                                 // WangLi3SMil_A..F: one kernel, the stage order is a run-time argument
       K_ALG_SRA = 8 };                                  // SRA1 / SRA3 / SOSRA / any Roessler SRA tableau (additive noise): one
                                                         // kernel, the tableau is a run-time argument (KArgs::sra)
// history lookup times per step: EulerHeun evaluates f and g a second time at t + dt; an SRA tableau evaluates f at
// t + c0_i dt and g at t + c1_i dt -- up to SDDE_SRA_MAX_SETS distinct times per step (the host deduplicates them)
#define SDDE_SRA_MAX_STAGES 3
#define SDDE_SRA_MAX_SETS 4
__host__ __device__ constexpr int alg_stages(int alg) { return alg == K_ALG_EULER_HEUN ? 2 : (alg == K_ALG_SRA ? SDDE_SRA_MAX_SETS : 1); }
// normals per noise channel and step: the SRA methods draw a second, independent increment dZ (W.dZ of
// /root/reference/src/solve.jl:308-312 `alg_needs_extra_process`); channels [0, NC) are dW, [NC, 2 NC) are dZ
__host__ __device__ constexpr int alg_noise_mult(int alg) { return alg == K_ALG_SRA ? 2 : 1; }
// batch depth a kernel is compiled with: multi-stage algorithms hold one node set per lookup time and step
#ifndef SDDE_SRA_BATCH
#define SDDE_SRA_BATCH 2   // two steps per batch: the Philox words of the next batch are produced one batch ahead and the lanes are compile-time (12.1 -> 10.3 ms on the additive Hayes problem, 10^6 x 1004 steps)
#endif
#ifndef SDDE_EH_BATCH_DIV
#define SDDE_EH_BATCH_DIV 1   // EulerHeun keeps the full batch depth: compile-time Philox lanes, words one batch ahead (4.82 -> 4.23 ms on Hayes, 10^6 x 1001 steps; 2 = halve it)
#endif
__host__ __device__ constexpr int alg_batch(int alg, int b) { return alg_stages(alg) >= 3 ? (b >= SDDE_SRA_BATCH ? SDDE_SRA_BATCH : 1) : ((alg_stages(alg) > 1 && b > 1) ? b / SDDE_EH_BATCH_DIV : b); }

#define SDDE_K_MAX_STATE 8
#define SDDE_K_MAX_PARAM 128
// CTA sizes.  Every CTA that draws in-kernel normals stages eight interleaved copies of the generator's table (160 KB,
// sdde_rng_tables.cuh) in shared memory, so the step kernels run ONE large CTA per SM: scalar models 512 threads,
// multi-state models 384 threads (168 registers per thread); the cooperative-lanes kernel 128 threads with one copy.
#ifndef SDDE_BLOCK_SCALAR
#define SDDE_BLOCK_SCALAR 512
#endif
#ifndef SDDE_BLOCK_MULTI
#define SDDE_BLOCK_MULTI 384
#endif
#define SDDE_BLOCK_COOP 128
#define SDDE_RNG_SMEM_BYTES(COPIES) (SDDE_RNG_TAB_ENTRIES * 16 * (COPIES))   // log + cos/sin tables of the normal generator
// first record of a batch whose steps the kernel may run through the lean body (pack_tables decides):
#define SDDE_MASK_FAST_EXACT (1 << 24)    // every lookup of every step hits a ring node exactly, on consecutive slots
#define SDDE_MASK_FAST_RING (1 << 25)     // every lookup of every step reads the ring (exact hits and interpolations mixed)
#define SDDE_MASK_FAST_INTERP (1 << 26)   // every lookup interpolates between CONSECUTIVE slots (the steady state of a lag on the grid)
// how many batches ahead the delayed-state nodes are pulled towards L1 (bytes in flight per SM = depth x 32 KB)
#ifndef SDDE_PF_DIST
#define SDDE_PF_DIST 1
#endif
#ifndef SDDE_PHILOX_ROUNDS
#define SDDE_PHILOX_ROUNDS 10
#endif

struct SaveRec {
    int step;       // step whose end produces this save (-1: initial condition)
    int exact;      // 1: take u_{n+1} verbatim
    double theta;   // else (1-theta) u_n + theta u_{n+1}
};

template <int NL>
struct StepRec {
    double t, dt, sqdt;
    int nsave;            // number of SaveRec produced by this step
    int ring_mask;        // lookup i = stage * NL + lag (i < 8).  bit i: reads the history ring; bit 8+i: its lower node is the
                          // upper node of the previous step's lookup i (kept in registers, not re-read); bit 16+i: constant
                          // u0 (past the end of the dense history); neither bit i nor 16+i: prehistory h(p, t - tau_l).
    long long off_store;  // BYTE offset (from the trajectory's ring column) of the slot that receives node n+1
    double inv_sqdt;      // 1 / sqrt|dt_n| (RKMil's difference quotient multiplies instead of dividing)
    struct Lk {
        double theta;         // interpolation weight; for prehistory: the lookup time t_n - tau_l
        long long off_lo;     // byte offsets of the two bracketing nodes; an exact hit is (lo = hi = node, theta = 1)
        long long off_hi;
        long long pad;
    } lk[NL];
};

// Roessler SRA tableau (additive noise) as the kernel sees it: stage i evaluates f with the delayed states of lookup set
// fset[i] at t + c0[i] dt and g with those of set gset[i] at t + c1[i] dt
struct SraTab {
    int s;                                       // stages (1 .. SDDE_SRA_MAX_STAGES)
    int fset[SDDE_SRA_MAX_STAGES], gset[SDDE_SRA_MAX_STAGES];
    double c0[SDDE_SRA_MAX_STAGES], c1[SDDE_SRA_MAX_STAGES];
    double A0[SDDE_SRA_MAX_STAGES][SDDE_SRA_MAX_STAGES], B0[SDDE_SRA_MAX_STAGES][SDDE_SRA_MAX_STAGES];   // [stage i][earlier stage j]
    double alpha[SDDE_SRA_MAX_STAGES], beta1[SDDE_SRA_MAX_STAGES], beta2[SDDE_SRA_MAX_STAGES];
};

struct KArgs {
    const unsigned char *steps;
    const SaveRec *saves;
    double *ring;
    double *out;
    const double *noise;
    const double *u0;        // per-trajectory initial state [NS][pitch] (wave local) or nullptr
    const double *ptraj;     // per-trajectory parameters   [NP][pitch] (wave local) or nullptr
    unsigned char *status;
    long long ring_pitch, out_pitch, out_off, noise_pitch, noise_off, u0_pitch, p_pitch;
    long long n_traj;        // valid trajectories of this wave
    long long traj_gid0;     // global id of the wave's first trajectory
    const double2 *rng_tab;  // generator table in global memory (sdde_rng_tables.cuh), staged per CTA
    long long slot_bytes;    // one ring slot = NS * ring_pitch * 8 bytes
    long long ring_bytes, out_bytes;   // extents of the ring and of `out` (bounds-checked builds, -DSDDE_CHECKED: every ring / out access traps when outside)
    int stage_order;         // K_ALG_WANGLI: the three stage operators in execution order, 2 bits each (0 drift, 1 Ito
                             // correction, 2 Stratonovich-Milstein diffusion), first stage in the low bits
    int n_full_ctas;         // CTAs [0, n_full_ctas) integrate BLOCK x TPT trajectories each, the remaining ("tail") CTAs
    int tail_traj;           // tail_traj each (a multiple of 32 x TPT): the last, partial round is spread over all SMs
    int n_steps, n_saves, ring_slots, chunk_steps, batch, noise_kind;
    int out_vec_ok;          // out rows are 16-byte aligned for double2 stores
    int n_saved;             // states saved per save point (save_idxs)
    int prefetch;            // step-by-step kernels: next step's nodes are older than this step (min_gap >= 1) -> L1 prefetch
    int save_row[SDDE_K_MAX_STATE];   // state j -> row within a save point, or -1 (not saved)
    u32 seed_lo, seed_hi;
    u32 rk[20];              // Philox round keys: rk[2r] = seed_lo + r*W0, rk[2r+1] = seed_hi + r*W1 (host precomputed)
    double hist_c0[SDDE_K_MAX_STATE], hist_c1[SDDE_K_MAX_STATE];
    double u0c[SDDE_K_MAX_STATE];
    double p[SDDE_K_MAX_PARAM];
    SraTab sra;              // K_ALG_SRA only
};

// tuning knobs per model shape
#ifndef SDDE_TUNE_B_SCALAR
#define SDDE_TUNE_B_SCALAR 4
#endif
#ifndef SDDE_TUNE_TPT_SCALAR
#define SDDE_TUNE_TPT_SCALAR 2
#endif
#ifndef SDDE_TUNE_FOOT_B2
#define SDDE_TUNE_FOOT_B2 4
#endif
#ifndef SDDE_TUNE_TPT_MULTI
#define SDDE_TUNE_TPT_MULTI 2
#endif
#ifndef SDDE_MIN_BLOCKS
#define SDDE_MIN_BLOCKS 4
#endif
#ifndef SDDE_TUNE_MULTI_MIN_NS
#define SDDE_TUNE_MULTI_MIN_NS 4
#endif
template <class M> struct Tune {
    static constexpr int TPT = (M::NS == 1) ? SDDE_TUNE_TPT_SCALAR : (M::NS >= SDDE_TUNE_MULTI_MIN_NS ? SDDE_TUNE_TPT_MULTI : 1);
    static constexpr int FOOT = M::NS * M::NL * TPT;              // doubles per prefetched step (one node set)
    static constexpr int B = (M::NS == 1 && M::NL == 1) ? SDDE_TUNE_B_SCALAR : (FOOT <= SDDE_TUNE_FOOT_B2 ? 2 : 1);
    // threads per CTA and resident CTAs per SM the register allocation aims for: 512 threads per SM with 128 registers
    // for the small models, 384 threads with 168 registers for multi-state models with two trajectories per thread
    // (they need the registers more than the occupancy: measured, profiles/README.md)
    static constexpr bool WIDE = (M::NS >= SDDE_TUNE_MULTI_MIN_NS && TPT > 1);
    static constexpr int BLOCK = WIDE ? SDDE_BLOCK_MULTI : SDDE_BLOCK_SCALAR;
#ifdef SDDE_MINB_SCALAR
    static constexpr int MINB = WIDE ? 1 : SDDE_MINB_SCALAR;       // A/B builds: few warps with many registers each
#else
    static constexpr int MINB = WIDE ? 1 : (512 / SDDE_BLOCK_SCALAR > 0 ? 512 / SDDE_BLOCK_SCALAR : 1);
#endif
};

// occupancy target per (model, algorithm): the model's, unless a measured exception is specialised next to the model
template <class M, int ALG> struct TuneK { static constexpr int MINB = Tune<M>::MINB; };

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw 2011; SDDE_PHILOX_ROUNDS = 7 is the paper's fastest Crush-resistant
// variant, build-time option) + Box-Muller in float64.
// Stream layout: counter = (traj_id_lo, traj_id_hi, call_lo, call_hi), key = (seed_lo, seed_hi);
// normal number q = step*NC + channel of a trajectory is lane (q & 3) of call (q >> 2): four normals per call.
// Replaces Xoroshiro128Plus + randn of /root/reference/src/solve.jl:313-336 (the bit stream is NOT reproduced,
// the distribution is: independent N(0,1), scaled by sqrt|dt_n| like [EXT] DiffEqNoiseProcess WHITE_NOISE_DIST).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(u32 c0, u32 c1, u32 c2, u32 c3, const u32 *__restrict__ rk, u32 (&o)[4]) {
#pragma unroll
    for (int r = 0; r < SDDE_PHILOX_ROUNDS; ++r) {
        const u64 p0 = (u64)0xD2511F53u * c0;
        const u64 p1 = (u64)0xCD9E8D57u * c2;
        const u32 n0 = (u32)(p1 >> 32) ^ c1 ^ rk[2 * r];
        const u32 n2 = (u32)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c1 = (u32)p1;
        c3 = (u32)p0;
        c0 = n0;
        c2 = n2;
    }
    o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
}

__device__ __forceinline__ void philox_round_keys(u32 k0, u32 k1, u32 *rk) {
    for (int r = 0; r < 10; ++r) { rk[2 * r] = k0 + (u32)r * 0x9E3779B9u; rk[2 * r + 1] = k1 + (u32)r * 0xBB67AE85u; }
}

__device__ __forceinline__ double rsqrt_approx(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
}

// FOUR N(0,1) variates from one Philox block: two Box-Muller pairs, pair A from (w0, w1), pair B from (w2, w3).
// Bit allocation per pair (wa, wb) -- identical in oracle/sdde_oracle.py and oracle/sdde_oracle.cpp:
//   radius: m = (wb & 0xFF)<<32 | wa  (40 bits), u = (m + 1/2) 2^-40 in (0,1), r = sqrt(-2 ln u)   [|z| <= 7.5 sigma]
//   angle : q = wb >> 30 (quadrant), f = (wb >> 8) & 0x3FFFFF (22 bits), x = (f + 1/2) 2^-22 - 1/2, theta = (pi/2)(q + x)
//   z0 = r cos(theta), z1 = r sin(theta)
// Why 40 + 24 bits and not 52 + 52: the 32x32->64-bit products of Philox share the FP64 datapath with DFMA on sm_100a
// (profiles/README.md), so bits are expensive.  The angle only has to make averages over the circle exact: for any
// smooth functional the mean over 2^24 equispaced directions is the trapezoidal rule of a smooth periodic function
// (error beyond any power of 2^-24), so each z stays N(0,1) to far below float64 rounding; the radius keeps 40 bits
// (tail truncated at 7.5 sigma, probability 5e-14 per draw; spacing of r at 6 sigma: 8e-5).
// Evaluation: every argument range is known, so no special cases and no libm calls; one shared-memory table
// (sdde_rng_tables.cuh, staged per CTA, COPIES interleaved copies, `tab` already points at this lane's copy):
//   -2 ln u  : u = 2^e' * M'; k = top 9 mantissa bits -> {1/c_k, -2 ln c'_k}; rho = fma(M, 1/c_k, -1) exact, |rho| <= 2^-9,
//              log1p(rho) = rho + rho^2 (L2 + L3 rho + L4 rho^2 + L5 rho^3)            (truncation < 2^-58)
//   sqrt     : MUFU.RSQ64H seed y, r0 = t y, e = 1 - r0 y, r = r0 (1 + e/2 + 3 e^2/8) (cubic step)   (<= 1 ulp)
//   sin, cos : the 2^23 angles of half a turn are alpha_i + beta_b + gamma_c with three tables of exact {cos, sin} (512 +
//              128 + 128 entries): TWO angle additions (8 FP64 instructions; <= 3 ulp absolute), the other half turn is a
//              sign flip of the radius.  (Round 1 evaluated sin / cos of a residual angle by polynomials: 11 FP64 and ~18
//              integer / select instructions, and its single-copy tables cost ~10 shared-memory wavefronts per lookup.)
// 25 FP64 instructions and four conflict-free 128-bit table loads per pair.
template <int COPIES>
__device__ __forceinline__ void normal_pair_from_words(u32 wa, u32 wb, const double2 *__restrict__ tab, double &z0, double &z1) {
    // ---- radius ----
    const double X = __hiloint2double((int)(0x3FF00000u | ((wb & 0xFFu) << 12) | (wa >> 20)), (int)(wa << 12));   // 1 + m 2^-40
    const double U = X - 0.99999999999954525264911353588104248046875;                         // (m + 1/2) 2^-40, exact
    const int hi = __double2hiint(U);
    const int k = (hi >> 11) & 511;
    const int n = 1023 - (hi >> 20) - ((hi >> 19) & 1);                                      // -e' in [0, 41]
    const double M = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(U));    // mantissa in [1,2)
    const double2 te = tab[k * COPIES];
    // ---- angle: table lookups issued early ----
    const double2 ca = tab[(SDDE_RNG_OFF_COARSE + ((wb >> 22) & 0x1FFu)) * COPIES];
    const double2 cb = tab[(SDDE_RNG_OFF_MID + ((wb >> 15) & 0x7Fu)) * COPIES];
    const double2 cc = tab[(SDDE_RNG_OFF_FINE + ((wb >> 8) & 0x7Fu)) * COPIES];
    // ---- -2 ln u ----
    const double rho = fma(M, te.x, -1.0);
    const double dn = __hiloint2double(0x43300000, n) - 4503599627370496.0;                  // (double)n
    double P = fma(rho, RNG_LOG1P_C[3], RNG_LOG1P_C[2]);
    P = fma(rho, P, RNG_LOG1P_C[1]);
    P = fma(rho, P, RNG_LOG1P_C[0]);
    const double l1p = fma(rho * rho, P, rho);
    double t = fma(dn, SDDE_RNG_TWO_LN2, te.y);
    t = fma(-2.0, l1p, t);                                                                   // -2 ln u  > 0
    // ---- sqrt(t) ----
    const double y = rsqrt_approx(t);
    const double r0 = t * y;
    const double e = fma(-r0, y, 1.0);
    const double g = fma(r0 * e, fma(e, 0.375, 0.5), r0);
    // ---- cos / sin (alpha + (beta + gamma)) ----
    const double cbc = fma(cb.x, cc.x, -(cb.y * cc.y));
    const double sbc = fma(cb.y, cc.x, cb.x * cc.y);
    const double cp = fma(ca.x, cbc, -(ca.y * sbc));
    const double sn = fma(ca.y, cbc, ca.x * sbc);
    const double gs = __hiloint2double(__double2hiint(g) ^ (int)(wb & 0x80000000u), __double2loint(g));   // second half turn: (-c, -s)
    z0 = gs * cp;
    z1 = gs * sn;
}

// normals 4*call .. 4*call+3 of trajectory gid
__device__ __forceinline__ void philox_words(u64 gid, u64 call, const u32 *__restrict__ rk, u32 (&w)[4]) {
    philox4x32_10((u32)gid, (u32)(gid >> 32), (u32)call, (u32)(call >> 32), rk, w);
}
template <int COPIES>
__device__ __forceinline__ void normal_quad_from_words(const u32 (&w)[4], const double2 *__restrict__ ltab, double (&z)[4]) {
    normal_pair_from_words<COPIES>(w[0], w[1], ltab, z[0], z[1]);
    normal_pair_from_words<COPIES>(w[2], w[3], ltab, z[2], z[3]);
}
template <int COPIES>
__device__ __forceinline__ void philox_normal_quad(u64 gid, u64 call, const u32 *__restrict__ rk,
                                                   const double2 *__restrict__ ltab, double (&z)[4]) {
    u32 w[4];
    philox_words(gid, call, rk, w);
    normal_quad_from_words<COPIES>(w, ltab, z);
}
// stage the table: COPIES interleaved copies (entry i of copy c at index COPIES * i + c); returns this thread's copy
template <int COPIES>
__device__ __forceinline__ const double2 *stage_rng_table(double2 *ltab, const double2 *__restrict__ gtab, int nthreads) {
    for (int i = threadIdx.x; i < SDDE_RNG_TAB_ENTRIES; i += nthreads) {
        const double2 v = gtab[i];
#pragma unroll
        for (int c = 0; c < COPIES; ++c) ltab[i * COPIES + c] = v;
    }
    return ltab + (threadIdx.x & (COPIES - 1));
}

// ------------------------------------------------------------------------------------------------
// Optional model-defined history function h(p, t) for t < t0 (user models: `static constexpr bool HAS_H = true;` and
// `template <class P> __device__ static void h(double *out, const P &p, double t)`); models without it use the
// descriptor's affine history c0 + c1*t.  The reference's h is an arbitrary callable (src/functionwrapper.jl:35-36).
// ------------------------------------------------------------------------------------------------
template <class...> struct VoidOf { typedef void type; };
template <class M, class = void> struct HasHist { static constexpr bool value = false; };
template <class M> struct HasHist<M, typename VoidOf<decltype(M::HAS_H)>::type> { static constexpr bool value = M::HAS_H; };

// ------------------------------------------------------------------------------------------------
// parameter accessors: kernel-argument constants (uniform) or per-trajectory registers (prob_func use case)
// ------------------------------------------------------------------------------------------------
struct ConstParams {
    const double *p;
    __device__ __forceinline__ double operator[](int i) const { return p[i]; }
};
template <int NP>
struct RegParams {
    double v[NP > 0 ? NP : 1];
    __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};

// ------------------------------------------------------------------------------------------------
// one step of one trajectory: [EXT] StochasticDiffEq perform_step! for EMConstantCache / RKMilConstantCache
// (SURVEY.md section 3.3).  ul = delayed states (same for every f/g evaluation of the step).
// ------------------------------------------------------------------------------------------------
template <class M, int ALG, class P>
__device__ __forceinline__ void sde_step(double (&u)[M::NS], const double (&uls)[alg_stages(ALG)][M::NL][M::NS], const P &p,
                                         double t, double dt, double sq, double isq, const double *dw, int stage_order = 0,
                                         const SraTab *sra = nullptr) {
    constexpr int NS = M::NS;
    const double (&ul)[M::NL][M::NS] = uls[0];
    if constexpr (ALG == K_ALG_SRA) {
        // [EXT] StochasticDiffEq SRAConstantCache (Roessler's SRA methods for additive noise; reference smoke tests
        // test/test_prob_sol.jl:74-87 run them adaptively -- here adaptive=false).  With chi2 = (dW + dZ/sqrt(3))/2 = I_(1,0)/dt:
        //   H_i = u + dt sum_{j<i} A0[i][j] f_j + (sum_{j<i} B0[i][j] g_j) chi2,   f_j = f(H_j, t + c0_j dt),  g_j = g(H_j, t + c1_j dt)
        //   u+  = u + dt sum_i alpha_i f_i + (sum_i beta1_i g_i) dW + (sum_i beta2_i g_i) chi2
        // Through the SDDE wrappers (src/functionwrapper.jl:8-9,35-36) every f / g evaluation looks the delayed states up at
        // ITS OWN time: the host deduplicates the stage times into lookup sets, `fset` / `gset` say which set an evaluation uses.
        constexpr int NL = M::NL, NW = M::DIAG ? 1 : M::NW, NC = M::DIAG ? NS : M::NW, NG = M::DIAG ? NS : NS * M::NW;
        const double *dz = dw + NC;
        double chi2[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) chi2[c] = 0.5 * (dw[c] + dz[c] * 0.57735026918962573);   // reference: dZ / sqrt(3) (1 ulp apart)
        double fs[SDDE_SRA_MAX_STAGES][NS], gs[SDDE_SRA_MAX_STAGES][NG];
#pragma unroll
        for (int i = 0; i < SDDE_SRA_MAX_STAGES; ++i) {
            if (i < sra->s) {                                  // warp-uniform
                double H[NS];
#pragma unroll
                for (int j = 0; j < NS; ++j) H[j] = u[j];
                if (i > 0) {
                    double A[NS], B[NG];
#pragma unroll
                    for (int j = 0; j < NS; ++j) A[j] = sra->A0[i][0] * fs[0][j];
#pragma unroll
                    for (int q = 0; q < NG; ++q) B[q] = sra->B0[i][0] * gs[0][q];
#pragma unroll
                    for (int jj = 1; jj < i; ++jj) {
#pragma unroll
                        for (int j = 0; j < NS; ++j) A[j] = A[j] + sra->A0[i][jj] * fs[jj][j];
#pragma unroll
                        for (int q = 0; q < NG; ++q) B[q] = B[q] + sra->B0[i][jj] * gs[jj][q];
                    }
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        double noise = B[j * NW] * chi2[M::DIAG ? j : 0];
#pragma unroll
                        for (int k = 1; k < NW; ++k) noise = noise + B[j * NW + k] * chi2[k];
                        H[j] = (u[j] + dt * A[j]) + noise;
                    }
                }
                double ulf[NL][NS], ulg[NL][NS];
                const int sf = sra->fset[i], sg = sra->gset[i];
#pragma unroll
                for (int l = 0; l < NL; ++l)
#pragma unroll
                    for (int j = 0; j < NS; ++j) {
                        double vf = uls[0][l][j], vg = uls[0][l][j];
#pragma unroll
                        for (int k = 1; k < SDDE_SRA_MAX_SETS; ++k) {
                            vf = (sf == k) ? uls[k][l][j] : vf;
                            vg = (sg == k) ? uls[k][l][j] : vg;
                        }
                        ulf[l][j] = vf; ulg[l][j] = vg;
                    }
                M::f(fs[i], H, ulf, p, t + sra->c0[i] * dt);
                M::g(gs[i], H, ulg, p, t + sra->c1[i] * dt);
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) fs[i][j] = 0.0;
#pragma unroll
                for (int q = 0; q < NG; ++q) gs[i][q] = 0.0;
            }
        }
        double at[NS], bt[NG], e2[NG];
#pragma unroll
        for (int j = 0; j < NS; ++j) at[j] = sra->alpha[0] * fs[0][j];
#pragma unroll
        for (int q = 0; q < NG; ++q) { bt[q] = sra->beta1[0] * gs[0][q]; e2[q] = sra->beta2[0] * gs[0][q]; }
#pragma unroll
        for (int i = 1; i < SDDE_SRA_MAX_STAGES; ++i) {
            if (i < sra->s) {
#pragma unroll
                for (int j = 0; j < NS; ++j) at[j] = at[j] + sra->alpha[i] * fs[i][j];
#pragma unroll
                for (int q = 0; q < NG; ++q) { bt[q] = bt[q] + sra->beta1[i] * gs[i][q]; e2[q] = e2[q] + sra->beta2[i] * gs[i][q]; }
            }
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            double bn = bt[j * NW] * dw[M::DIAG ? j : 0], en = e2[j * NW] * chi2[M::DIAG ? j : 0];
#pragma unroll
            for (int k = 1; k < NW; ++k) { bn = bn + bt[j * NW + k] * dw[k]; en = en + e2[j * NW + k] * chi2[k]; }
            u[j] = ((u[j] + dt * at[j]) + bn) + en;
        }
        return;
    }
    if constexpr (ALG == K_ALG_WANGLI && M::DIAG) {
        // [EXT] StochasticDiffEq WangLi3SMil_A..F (three-stage Milstein methods of Wang & Li, diagonal noise; reference smoke
        // and convergence tests test/test_prob_sol.jl:50-61, test/analyticless_convergence_tests.jl:62-91).  LOW confidence
        // restatement (see oracle/sdde_oracle.py): each step composes three stage operators, every f / g evaluated at the
        // step's t with the step's delayed states, the derivative g'g by a forward difference of width sqrt(dt):
        //   D  u <- u + dt f(u)
        //   C  k = g(u), k1 = (g(u + k sqrt(dt)) - k)/sqrt(dt);  u <- u - (dt/2) k1                  (Ito correction)
        //   M  k = g(u), k1 = (g(u + k sqrt(dt)) - k)/sqrt(dt);  u <- u + k dW + (dW^2/2) k1         (Stratonovich Milstein)
        // and the six letters are the six orders of (D, C, M); `stage_order` carries the order (warp-uniform branches).
#pragma unroll
        for (int sgi = 0; sgi < 3; ++sgi) {
            const int op = (stage_order >> (2 * sgi)) & 3;
            if (op == 0) {
                double du[NS];
                M::f(du, u, ul, p, t);
#pragma unroll
                for (int j = 0; j < NS; ++j) u[j] = u[j] + dt * du[j];
            } else {
                double k[NS], tmp[NS], k1[NS];
                M::g(k, u, ul, p, t);
#pragma unroll
                for (int j = 0; j < NS; ++j) tmp[j] = u[j] + k[j] * sq;
                M::g(k1, tmp, ul, p, t);
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const double d = (k1[j] - k[j]) * isq;
                    if (op == 1) u[j] = u[j] - (dt * 0.5) * d;
                    else u[j] = (u[j] + k[j] * dw[j]) + ((dw[j] * dw[j]) * 0.5) * d;
                }
            }
        }
        return;
    }
    if (ALG == K_ALG_EULER_HEUN) {
        // [EXT] EulerHeunConstantCache: tmp = u + f dt + g dW;  u+ = u + dt/2 (f + f(tmp, t+dt)) + 1/2 (g + g(tmp, t+dt)) dW
        constexpr int NW = M::DIAG ? 1 : M::NW, NG = M::DIAG ? NS : NS * M::NW;
        const double (&ul2)[M::NL][M::NS] = uls[alg_stages(ALG) - 1];
        double f1[NS], g1[NG], tmp[NS], f2[NS], g2[NG];
        M::f(f1, u, ul, p, t);
        M::g(g1, u, ul, p, t);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            // roundings pinned (the reference forms noise = g .* dW first, then muladd(f, dt, uprev) + noise): every batch body
            // of every instantiation rounds alike
            double noise = __dmul_rn(g1[i * NW], dw[M::DIAG ? i : 0]);
#pragma unroll
            for (int j = 1; j < NW; ++j) noise = fma(g1[i * NW + j], dw[j], noise);
            tmp[i] = __dadd_rn(fma(f1[i], dt, u[i]), noise);
        }
        M::f(f2, tmp, ul2, p, t + dt);
        M::g(g2, tmp, ul2, p, t + dt);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double noise2 = __dmul_rn(0.5 * (g1[i * NW] + g2[i * NW]), dw[M::DIAG ? i : 0]);
#pragma unroll
            for (int j = 1; j < NW; ++j) noise2 = fma(0.5 * (g1[i * NW + j] + g2[i * NW + j]), dw[j], noise2);
            u[i] = __dadd_rn(fma(dt * 0.5, f1[i] + f2[i], u[i]), noise2);
        }
        return;
    }
    double du[NS], K[NS];
    M::f(du, u, ul, p, t);
#pragma unroll
    for (int j = 0; j < NS; ++j) K[j] = u[j] + dt * du[j];
    if constexpr ((ALG == K_ALG_RKMILC_ITO || ALG == K_ALG_RKMILC_STRAT) && !M::DIAG) {
        // [EXT] RKMilCommuteConstantCache, noise matrix G (NS x NW) with commutative columns:
        //   J = dW dW'/2 (Ito: - |dt|/2 on the diagonal);  per column j: K_j = (Ito ? K : u) + sqrt(dt) G[:, j]  (the stage
        //   point of RKMil's diagonal branch; with K under Stratonovich the scheme drops to strong order ~0.7, tests/test_oracle.py),
        //   Dg_j = (G(K_j) - G(u)) / sqrt(dt),  correction += Dg_j J[:, j];   u+ = (K + G dW) + correction
        constexpr int NW = M::NW;
        double L[NS * NW], mil[NS];
        M::g(L, u, ul, p, t);
#pragma unroll
        for (int i = 0; i < NS; ++i) mil[i] = 0.0;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            double Kj[NS], gt[NS * NW];
#pragma unroll
            for (int i = 0; i < NS; ++i) Kj[i] = (ALG == K_ALG_RKMILC_ITO ? K[i] : u[i]) + sq * L[i * NW + j];
            M::g(gt, Kj, ul, p, t);
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                const double Jkj = (ALG == K_ALG_RKMILC_ITO && k == j) ? fma(0.5, dw[k] * dw[j], -0.5 * fabs(dt)) : 0.5 * (dw[k] * dw[j]);
#pragma unroll
                for (int i = 0; i < NS; ++i) mil[i] = mil[i] + ((gt[i * NW + k] - L[i * NW + k]) * isq) * Jkj;
            }
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double noise = L[i * NW] * dw[0];
#pragma unroll
            for (int j = 1; j < NW; ++j) noise = noise + L[i * NW + j] * dw[j];
            u[i] = (K[i] + noise) + mil[i];
        }
        return;
    }
    if (ALG == K_ALG_EM_SPLIT || ALG == K_ALG_EM_NOSPLIT) {
        if (M::DIAG) {
            double gv[NS];
            if (ALG == K_ALG_EM_SPLIT) M::g(gv, K, ul, p, t); else M::g(gv, u, ul, p, t);
#pragma unroll
            for (int j = 0; j < NS; ++j) u[j] = K[j] + gv[j] * dw[j];
        } else {
            constexpr int NW = M::NW;
            double gv[NS * NW];
            if (ALG == K_ALG_EM_SPLIT) M::g(gv, K, ul, p, t); else M::g(gv, u, ul, p, t);
#pragma unroll
            for (int i = 0; i < NS; ++i) {
                double noise = gv[i * NW] * dw[0];
#pragma unroll
                for (int j = 1; j < NW; ++j) noise = noise + gv[i * NW + j] * dw[j];
                u[i] = K[i] + noise;
            }
        }
    } else {
        // RKMil, diagonal noise only (the host rejects the non-diagonal combination)
        double L[NS], ut[NS], g2[NS];
        M::g(L, u, ul, p, t);
#pragma unroll
        for (int j = 0; j < NS; ++j) ut[j] = (ALG == K_ALG_RKMIL_ITO ? K[j] : u[j]) + L[j] * sq;
        M::g(g2, ut, ul, p, t);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            const double ggp = (g2[j] - L[j]) * isq;          // reference: ./ sqdt (1 ulp apart)
            const double w2 = (ALG == K_ALG_RKMIL_ITO) ? (dw[j] * dw[j] - fabs(dt)) : (dw[j] * dw[j]);
            const double mil = (ggp * w2) * 0.5;
            u[j] = (K[j] + L[j] * dw[j]) + mil;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// vector helpers: TPT adjacent trajectories per thread
// ------------------------------------------------------------------------------------------------
// -DSDDE_CHECKED: a verification build (compute-sanitizer is not available on the measurement pool): every ring and `out`
// access of the step kernel is checked against the buffer extents and traps when outside (tools/sanitize_small.py)
#ifdef SDDE_CHECKED
#define SDDE_CHECK_RING(a, ptr, nbytes) do { const long long _o = (const char *)(ptr) - (const char *)(a).ring; \
        if (_o < 0 || _o + (long long)(nbytes) > (a).ring_bytes) __trap(); } while (0)
#define SDDE_CHECK_OUT(a, ptr, nbytes) do { const long long _o = (const char *)(ptr) - (const char *)(a).out; \
        if (_o < 0 || _o + (long long)(nbytes) > (a).out_bytes) __trap(); } while (0)
#else
#define SDDE_CHECK_RING(a, ptr, nbytes) do { } while (0)
#define SDDE_CHECK_OUT(a, ptr, nbytes) do { } while (0)
#endif
template <int TPT> __device__ __forceinline__ void ld_traj(const double *p, double (&v)[TPT]);
template <> __device__ __forceinline__ void ld_traj<1>(const double *p, double (&v)[1]) { v[0] = *p; }
template <> __device__ __forceinline__ void ld_traj<2>(const double *p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2 *>(p);
    v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void ld_traj<4>(const double *p, double (&v)[4]) {
    const double2 t0 = *reinterpret_cast<const double2 *>(p), t1 = *reinterpret_cast<const double2 *>(p + 2);
    v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
}
template <int TPT> __device__ __forceinline__ void st_traj(double *p, const double (&v)[TPT]);
template <> __device__ __forceinline__ void st_traj<1>(double *p, const double (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void st_traj<2>(double *p, const double (&v)[2]) {
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
}

template <> __device__ __forceinline__ void st_traj<4>(double *p, const double (&v)[4]) {
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
}

__device__ __forceinline__ bool finite_d(double x) { return fabs(x) <= 1.7976931348623157e308; }

// ------------------------------------------------------------------------------------------------
// the step kernel: thread-per-trajectory (x TPT), whole time span in one launch
// ------------------------------------------------------------------------------------------------
template <class M, int TPT, bool PTRAJ>
struct TrajState {
    static constexpr int NS = M::NS, NC = M::DIAG ? M::NS : M::NW, NP = M::NP;
    double u[TPT][NS];
    RegParams<PTRAJ ? NP : 0> preg[TPT];
    double zc[TPT][4];        // cached normals of the current Philox call
    u64 zcall;                // its call index
    u32 wn[TPT][4];           // Philox words of call wcall, computed one batch ahead (batched path): the 10-round
    u64 wcall;                // dependent chain then has no consumer inside the batch and hides under its FP64 work
    double curW[TPT][2 * NC]; // NoiseGrid mode: current W (and Z for the algorithms that draw dZ)
    bool bad[TPT];
    int ks;                   // next SaveRec
    double carry[M::NL * SDDE_SRA_MAX_SETS][NS][TPT];   // upper nodes of the previous step's lookups (consecutive steps share a node)
};

// BW consecutive steps starting at global step index n (recs[0..BW)).  BW > 1: the delayed-state nodes of all BW
// steps are loaded up front (the host guarantees they are older than the batch) and n*NC is even, so which half of
// a Philox pair a step consumes is a compile-time fact.
// FAST selects a lean body for batches pack_tables marked on their first record (SDDE_MASK_FAST_*): every lookup of the
// batch reads the ring, the stored slots are consecutive ring slots without a wrap, dt is the same for all BW steps and
// no step saves more than 255 points.  The lean bodies take dt, sqrt|dt|, the store base and the packed save counts from
// recs[0] and have no prehistory / constant-u0 selects.  K_FAST_EXACT (every lookup hits a node exactly and the nodes are
// consecutive slots): one load per step from ONE base, no blend arithmetic -- five 128-bit shared-memory loads per batch
// instead of nine 64-bit loads per step.  K_FAST_RING (exact hits and interpolations mixed -- t_n - tau lands on either
// side of a node by rounding, so even "exact" lags interpolate with theta ~ 1e-16 on some steps): per-step theta and
// node offsets as in the general body.
// K_FAST_INTERP: every lookup interpolates and both bracketing nodes advance by one slot per step (lo(k) = hi(k-1)): loads
// from one base like the exact body, only theta is read per step.
enum { K_FAST_NONE = 0, K_FAST_EXACT = 1, K_FAST_RING = 2, K_FAST_INTERP = 3 };
template <class M, int ALG, int NOISE, bool PTRAJ, int TPT, int BW, int FAST>
__device__ __forceinline__ void run_batch(const KArgs &a, TrajState<M, TPT, PTRAJ> &st, const StepRec<M::NL * alg_stages(ALG)> *recs, int n,
                                          char *ringb, const double2 *ltab, long long tr0, long long last, bool out_vec,
                                          bool prefetch_next) {
    constexpr int NS = M::NS, NL = M::NL, NC = (M::DIAG ? M::NS : M::NW) * alg_noise_mult(ALG), NSTG = alg_stages(ALG), NLK = NL * NSTG;
    constexpr bool STATIC_RNG = (BW > 1) && ((BW * NC) % 4 == 0);   // batch start n*NC is then a multiple of 4 as well
    constexpr bool LEAN = FAST != K_FAST_NONE;
    typedef StepRec<NLK> Rec;
    const long long jstride = a.ring_pitch * 8;
    const long long slot = a.slot_bytes;
    // ---- all delayed-state loads of the batch, unconditionally (prehistory records point at slot 0) ----
    double xlo[BW][NLK][NS][TPT], xhi[BW][NLK][NS][TPT];
    if constexpr (FAST == K_FAST_EXACT || FAST == K_FAST_INTERP) {
#pragma unroll
        for (int l = 0; l < NLK; ++l) {
            const char *phi = ringb + recs[0].lk[l].off_hi;
            if (FAST == K_FAST_INTERP) {
                const bool reuse = (recs[0].ring_mask & (0x100 << l)) != 0;    // lo(n) is hi(n-1) of the previous batch: in registers
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    if (!reuse) { SDDE_CHECK_RING(a, phi - slot + j * jstride, 8 * TPT); ld_traj<TPT>(reinterpret_cast<const double *>(phi - slot + j * jstride), xlo[0][l][j]); }
                    else {
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) xlo[0][l][j][tt] = st.carry[l][j][tt];
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < BW; ++k)
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    SDDE_CHECK_RING(a, phi + k * slot + j * jstride, 8 * TPT);
                    ld_traj<TPT>(reinterpret_cast<const double *>(phi + k * slot + j * jstride), xhi[k][l][j]);
                    if (FAST == K_FAST_INTERP && k > 0) {
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) xlo[k][l][j][tt] = xhi[k - 1][l][j][tt];
                    }
                }
        }
    } else {
#pragma unroll
        for (int k = 0; k < BW; ++k) {
#pragma unroll
            for (int l = 0; l < NLK; ++l) {
                const char *plo = ringb + recs[k].lk[l].off_lo, *phi = ringb + recs[k].lk[l].off_hi;
#ifdef SDDE_NO_CARRY
                const bool reuse = false;
#else
                const bool reuse = (recs[k].ring_mask & (0x100 << l)) != 0;    // lo(n) is hi(n-1): already in registers
#endif
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    if (!reuse) { SDDE_CHECK_RING(a, plo + j * jstride, 8 * TPT); ld_traj<TPT>(reinterpret_cast<const double *>(plo + j * jstride), xlo[k][l][j]); }
                    SDDE_CHECK_RING(a, phi + j * jstride, 8 * TPT);
                    ld_traj<TPT>(reinterpret_cast<const double *>(phi + j * jstride), xhi[k][l][j]);
                    if (reuse) {
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) xlo[k][l][j][tt] = (k == 0) ? st.carry[l][j][tt] : xhi[k > 0 ? k - 1 : 0][l][j][tt];
                    }
                }
            }
        }
    }
#pragma unroll
    for (int l = 0; l < NLK; ++l)
#pragma unroll
        for (int j = 0; j < NS; ++j)
#pragma unroll
            for (int tt = 0; tt < TPT; ++tt) st.carry[l][j][tt] = xhi[BW - 1][l][j][tt];
    // batched path: pull the NEXT batch's nodes towards L1/L2 now; this batch's generator and recurrence work (about a
    // thousand cycles per warp) covers the HBM latency that the first use of a freshly loaded node otherwise exposes
    if (BW > 1 && prefetch_next) {
#pragma unroll
        for (int k = 0; k < BW; ++k)
#pragma unroll
            for (int l = 0; l < NLK; ++l) {
                const char *phi = ringb + recs[SDDE_PF_DIST * BW + k].lk[l].off_hi;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(phi + j * jstride));
                }
            }
    }
    // step-by-step path (multi-state models): pull the NEXT step's upper nodes towards L1 while this step computes
    // (they were written at least one step ago; recs[BW] is staged by the caller when prefetch_next is set)
    if (BW == 1 && prefetch_next) {
#pragma unroll
        for (int l = 0; l < NLK; ++l) {
            const char *phi = ringb + recs[BW].lk[l].off_hi;
#pragma unroll
            for (int j = 0; j < NS; ++j) asm volatile("prefetch.global.L1 [%0];" ::"l"(phi + j * jstride));
        }
    }
    // ---- increments of the whole batch first: accept_step! draws dW ~ N(0,|dt_n|) (src/integrators/interface.jl:449-454).
    //      This work is independent of the loads above, so it covers their latency. ----
    double dwb[BW][TPT][NC];
#pragma unroll
    for (int k = 0; k < BW; ++k) {
        const double sq = recs[LEAN ? 0 : k].sqdt;
        if (NOISE == K_NOISE_PHILOX) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const u64 q = (u64)(n + k) * NC + c;
                const u64 call = q >> 2;
                const bool fresh = STATIC_RNG ? (((k * NC + c) & 3) == 0) : (call != st.zcall);
                if (fresh) {
                    if (STATIC_RNG && (k * NC + c) == 0) {
                        // words of this call were produced during the previous batch (recomputed after a step-by-step
                        // stretch); produce the words of the NEXT batch's first call now, nothing in this batch waits for them
                        if (st.wcall != call) {
#pragma unroll
                            for (int tt = 0; tt < TPT; ++tt) philox_words((u64)(a.traj_gid0 + tr0 + tt), call, a.rk, st.wn[tt]);
                        }
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) {
                            normal_quad_from_words<SDDE_RNG_COPIES>(st.wn[tt], ltab, st.zc[tt]);
                            philox_words((u64)(a.traj_gid0 + tr0 + tt), call + (u64)(BW * NC / 4), a.rk, st.wn[tt]);
                        }
                        st.wcall = call + (u64)(BW * NC / 4);
                    } else {
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt)
                            philox_normal_quad<SDDE_RNG_COPIES>((u64)(a.traj_gid0 + tr0 + tt), call, a.rk, ltab, st.zc[tt]);
                    }
                    st.zcall = call;
                }
                const int lane = STATIC_RNG ? ((k * NC + c) & 3) : (int)(q & 3);
#pragma unroll
                for (int tt = 0; tt < TPT; ++tt) {
                    const double zlo = (lane & 1) ? st.zc[tt][1] : st.zc[tt][0];
                    const double zhi = (lane & 1) ? st.zc[tt][3] : st.zc[tt][2];
                    dwb[k][tt][c] = sq * ((lane & 2) ? zhi : zlo);
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c)
#pragma unroll
                for (int tt = 0; tt < TPT; ++tt) {
                    const long long tr = (tr0 + tt <= last) ? tr0 + tt : last;
                    const long long row = (NOISE == K_NOISE_GRID) ? ((long long)(n + k + 1) * NC + c) : ((long long)(n + k) * NC + c);
                    const double w = a.noise[row * a.noise_pitch + a.noise_off + tr];
                    if (NOISE == K_NOISE_GRID) { dwb[k][tt][c] = w - st.curW[tt][c]; st.curW[tt][c] += dwb[k][tt][c]; }
                    else dwb[k][tt][c] = w;
                }
        }
    }
    // ---- the steps: straight-line code (no branch inside the batch, so the scheduler can overlap the BW x TPT
    //      independent pieces: delayed-state drift terms and the u recurrences) ----
    double useq[BW + 1][TPT][NS];     // states at the batch's nodes (for the deferred savevalues!)
#pragma unroll
    for (int tt = 0; tt < TPT; ++tt)
#pragma unroll
        for (int j = 0; j < NS; ++j) useq[0][tt][j] = st.u[tt][j];
#pragma unroll
    for (int k = 0; k < BW; ++k) {
        const Rec &r = recs[LEAN ? 0 : k];
        const double tn = recs[k].t, dtn = r.dt, sq = r.sqdt, isq = r.inv_sqdt;
        // delayed states: [EXT] HistoryFunction regimes
        double ul[TPT][NSTG][NL][NS];
#pragma unroll
        for (int l = 0; l < NLK; ++l) {
            if constexpr (FAST == K_FAST_EXACT) {
#pragma unroll
                for (int j = 0; j < NS; ++j)
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt) ul[tt][l / NL][l % NL][j] = xhi[k][l][j][tt];
            } else if constexpr (FAST == K_FAST_RING || FAST == K_FAST_INTERP) {
                const double th = recs[k].lk[l].theta, omt = 1.0 - th;
#pragma unroll
                for (int j = 0; j < NS; ++j)
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt) ul[tt][l / NL][l % NL][j] = fma(th, xhi[k][l][j][tt], omt * xlo[k][l][j][tt]);   // explicit: every body rounds alike
            } else {
                const double th = r.lk[l].theta;
                const bool ring = (r.ring_mask & (1 << l)) != 0;
                const double omt = 1.0 - th;
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const double hv = a.hist_c0[j] + a.hist_c1[j] * th;
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt) {
                        const double iv = fma(th, xhi[k][l][j][tt], omt * xlo[k][l][j][tt]);
                        ul[tt][l / NL][l % NL][j] = ring ? iv : hv;
                    }
                }
                if constexpr (HasHist<M>::value) {
                    if (!ring) {                                    // prehistory record (warp-uniform: the schedule is shared)
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) {
                            double hv[NS];
                            if (PTRAJ) M::h(hv, st.preg[tt], th);
                            else M::h(hv, ConstParams{a.p}, th);
#pragma unroll
                            for (int j = 0; j < NS; ++j) ul[tt][l / NL][l % NL][j] = hv[j];
                        }
                    }
                }
                if (NSTG > 1 && (r.ring_mask & (0x10000 << l))) {   // second-stage lookup past the dense history: constant u0
#pragma unroll
                    for (int j = 0; j < NS; ++j)
#pragma unroll
                        for (int tt = 0; tt < TPT; ++tt) {
                            const long long tr = (tr0 + tt <= last) ? tr0 + tt : last;
                            ul[tt][l / NL][l % NL][j] = a.u0 ? a.u0[j * a.u0_pitch + tr] : a.u0c[j];
                        }
                }
            }
        }
        double (&dw)[TPT][NC] = dwb[k];
        // perform_step!
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            if (PTRAJ) sde_step<M, ALG>(st.u[tt], ul[tt], st.preg[tt], tn, dtn, sq, isq, dw[tt], a.stage_order, &a.sra);
            else sde_step<M, ALG>(st.u[tt], ul[tt], ConstParams{a.p}, tn, dtn, sq, isq, dw[tt], a.stage_order, &a.sra);
#pragma unroll
            for (int j = 0; j < NS; ++j) useq[k + 1][tt][j] = st.u[tt][j];
        }
        // dense history: node n+1 (the save_everystep push of src/integrators/interface.jl:406-417)
        char *pst = ringb + (LEAN ? recs[0].off_store + k * slot : recs[k].off_store);
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            double v[TPT];
#pragma unroll
            for (int tt = 0; tt < TPT; ++tt) v[tt] = st.u[tt][j];
            SDDE_CHECK_RING(a, pst + j * jstride, 8 * TPT);
            st_traj<TPT>(reinterpret_cast<double *>(pst + j * jstride), v);
        }
    }
    // [EXT] check_error! (src/solve.jl:578) looks at u after every step; here once per batch: every algorithm of this path
    // forms u_{n+1} = u_n + increments, so a non-finite u_n never turns finite again and nothing slips through between checks
#pragma unroll
    for (int tt = 0; tt < TPT; ++tt)
#pragma unroll
        for (int j = 0; j < NS; ++j) st.bad[tt] = st.bad[tt] || !finite_d(st.u[tt][j]);
    // ---- savevalues! (src/integrators/interface.jl:357-430), after the batch ----
    const unsigned long long packed = LEAN ? (unsigned long long)recs[0].lk[0].pad : ~0ull;   // lean: nsave of step k in byte k
    if (packed != 0) {
#pragma unroll
        for (int k = 0; k < BW; ++k) {
            const int ns = LEAN ? (int)((packed >> (8 * k)) & 255u) : recs[k].nsave;
            for (int s = 0; s < ns; ++s) {
                const SaveRec sr = a.saves[st.ks];
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    double v[TPT];
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt) {
                        v[tt] = sr.exact ? useq[k + 1][tt][j] : fma(sr.theta, useq[k + 1][tt][j], (1.0 - sr.theta) * useq[k][tt][j]);
                        st.bad[tt] = st.bad[tt] || !finite_d(v[tt]);
                    }
                    if (a.out && a.save_row[j] >= 0) {
                        double *op = a.out + ((long long)st.ks * a.n_saved + a.save_row[j]) * a.out_pitch + a.out_off + tr0;
                        if (out_vec) { if (tr0 <= last) { SDDE_CHECK_OUT(a, op, 8 * TPT); st_traj<TPT>(op, v); } }
                        else {
#pragma unroll
                            for (int tt = 0; tt < TPT; ++tt)
                                if (tr0 + tt <= last) { SDDE_CHECK_OUT(a, op + tt, 8); op[tt] = v[tt]; }
                        }
                    }
                }
                ++st.ks;
            }
        }
    }
}

template <class M, int ALG, int NOISE, bool PTRAJ, int TPT, int B>
__global__ void __launch_bounds__(Tune<M>::BLOCK, TuneK<M, ALG>::MINB) sdde_solve_kernel(const __grid_constant__ KArgs a) {
    constexpr int NS = M::NS, NL = M::NL, NC = (M::DIAG ? M::NS : M::NW) * alg_noise_mult(ALG), NP = M::NP, BLOCK = Tune<M>::BLOCK;
    constexpr int TAB_BYTES = (NOISE == K_NOISE_PHILOX) ? SDDE_RNG_SMEM_BYTES(SDDE_RNG_COPIES) : 0;   // supplied noise: no generator table
    typedef StepRec<NL * alg_stages(ALG)> Rec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *srec = reinterpret_cast<Rec *>(smem_raw + TAB_BYTES);
    // Trajectories of this CTA.  One CTA is resident per SM, so a wave of W trajectories would take ceil(W / (SMs x BLOCK x
    // TPT)) full rounds with the last one mostly idle (10^6 paths = 6.6 rounds on 148 SMs: 7 rounds of time).  The host
    // therefore launches whole rounds of full CTAs followed by ONE round of equally sized smaller "tail" CTAs; warps of a
    // tail CTA that have no trajectories leave before the first barrier.
    const bool full = (int)blockIdx.x < a.n_full_ctas;
    const long long cta0 = full ? (long long)blockIdx.x * (BLOCK * TPT)
                                : (long long)a.n_full_ctas * (BLOCK * TPT) + (long long)((int)blockIdx.x - a.n_full_ctas) * a.tail_traj;
    const int nact = full ? BLOCK : a.tail_traj / TPT;                      // active threads (a multiple of 32)
    if ((int)threadIdx.x >= nact) return;
    const double2 *ltab = nullptr;                                          // this lane's copy of the generator's table
    if (NOISE == K_NOISE_PHILOX) ltab = stage_rng_table<SDDE_RNG_COPIES>(reinterpret_cast<double2 *>(smem_raw), a.rng_tab, nact);

    const long long tr0 = cta0 + (long long)threadIdx.x * TPT;              // wave-local trajectory index
    const long long last = a.n_traj - 1;

    // ---- per-trajectory "__init": state, parameters, noise state (src/solve.jl:137, 302-339) ----
    TrajState<M, TPT, PTRAJ> st;
#pragma unroll
    for (int tt = 0; tt < TPT; ++tt) {
        const long long tr = (tr0 + tt <= last) ? tr0 + tt : last;
#pragma unroll
        for (int j = 0; j < NS; ++j) st.u[tt][j] = a.u0 ? a.u0[j * a.u0_pitch + tr] : a.u0c[j];
        if (PTRAJ) {
#pragma unroll
            for (int i = 0; i < NP; ++i) st.preg[tt].v[PTRAJ ? i : 0] = a.ptraj[i * a.p_pitch + tr];
        }
        st.zc[tt][0] = st.zc[tt][1] = st.zc[tt][2] = st.zc[tt][3] = 0.0;
        st.bad[tt] = false;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            st.curW[tt][c] = (NOISE == K_NOISE_GRID) ? a.noise[(long long)c * a.noise_pitch + a.noise_off + tr] : 0.0;
    }
    st.zcall = ~0ull;
    st.wcall = ~0ull;
    st.ks = 0;
#pragma unroll
    for (int l = 0; l < NL * SDDE_SRA_MAX_SETS; ++l)
#pragma unroll
        for (int j = 0; j < NS; ++j)
#pragma unroll
            for (int tt = 0; tt < TPT; ++tt) st.carry[l][j][tt] = 0.0;

    // ---- node 0 of the history = u0 (src/utils.jl:143-247: sol.t=[t0], sol.u=[u0]) ----
    char *ringb = reinterpret_cast<char *>(a.ring + tr0);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double v[TPT];
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) v[tt] = st.u[tt][j];
        SDDE_CHECK_RING(a, ringb + (long long)j * a.ring_pitch * 8, 8 * TPT);
        st_traj<TPT>(reinterpret_cast<double *>(ringb + (long long)j * a.ring_pitch * 8), v);
    }

    const bool out_vec = (TPT == 1) || (a.out_vec_ok && (tr0 + TPT - 1 <= last));
    // save_start (src/solve.jl:355-363)
    while (st.ks < a.n_saves && a.saves[st.ks].step < 0) {
        if (a.out) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                if (a.save_row[j] < 0) continue;
                double *op = a.out + ((long long)st.ks * a.n_saved + a.save_row[j]) * a.out_pitch + a.out_off + tr0;
                if (out_vec) {
                    double v[TPT];
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt) v[tt] = st.u[tt][j];
                    if (tr0 <= last) { SDDE_CHECK_OUT(a, op, 8 * TPT); st_traj<TPT>(op, v); }
                } else {
#pragma unroll
                    for (int tt = 0; tt < TPT; ++tt)
                        if (tr0 + tt <= last) { SDDE_CHECK_OUT(a, op + tt, 8); op[tt] = st.u[tt][j]; }
                }
            }
        }
        ++st.ks;
    }

    const bool batched = (B > 1) && (a.batch >= B);   // host: batch = min(B, min_gap + 1); chunk_steps % B == 0
    for (int n0 = 0; n0 < a.n_steps; n0 += a.chunk_steps) {
        const int cnt = (a.n_steps - n0 < a.chunk_steps) ? (a.n_steps - n0) : a.chunk_steps;
        constexpr int LOOK = (B > 1) ? SDDE_PF_DIST * B : 1;                                   // look-ahead records (prefetch)
        const int nrec = (a.n_steps - n0 - cnt >= LOOK) ? cnt + LOOK : a.n_steps - n0;
        __syncthreads();
        {   // stage this chunk of the schedule in shared memory
            const uint4 *src = reinterpret_cast<const uint4 *>(a.steps + (size_t)n0 * sizeof(Rec));
            uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + TAB_BYTES);
            const int nvec = nrec * (int)(sizeof(Rec) / 16);
            for (int i = threadIdx.x; i < nvec; i += nact) dst[i] = src[i];
        }
        __syncthreads();
        int nb = 0;
        if (batched) {
            for (; nb + B <= cnt; nb += B) {
                const int fm = srec[nb].ring_mask;
                const bool pfn = nb + (SDDE_PF_DIST + 1) * B <= nrec;   // the records of the batch to prefetch are staged
                if (fm & SDDE_MASK_FAST_EXACT)
                    run_batch<M, ALG, NOISE, PTRAJ, TPT, B, K_FAST_EXACT>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
                else if (fm & SDDE_MASK_FAST_INTERP)
                    run_batch<M, ALG, NOISE, PTRAJ, TPT, B, K_FAST_INTERP>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
                else if (fm & SDDE_MASK_FAST_RING)
                    run_batch<M, ALG, NOISE, PTRAJ, TPT, B, K_FAST_RING>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
                else
                    run_batch<M, ALG, NOISE, PTRAJ, TPT, B, K_FAST_NONE>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
            }
        }
        const bool pf = (B == 1) && a.prefetch;
        for (; nb < cnt; ++nb) {
            const int fm = (B == 1) ? srec[nb].ring_mask : 0;   // step-by-step kernels: the marks are per step
            const bool pfn = pf && (nb + 1 < nrec);
            if (fm & SDDE_MASK_FAST_EXACT)
                run_batch<M, ALG, NOISE, PTRAJ, TPT, 1, K_FAST_EXACT>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
            else if (fm & SDDE_MASK_FAST_RING)
                run_batch<M, ALG, NOISE, PTRAJ, TPT, 1, K_FAST_RING>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
            else
                run_batch<M, ALG, NOISE, PTRAJ, TPT, 1, K_FAST_NONE>(a, st, srec + nb, n0 + nb, ringb, ltab, tr0, last, out_vec, pfn);
        }
    }
    // [EXT] check_error! (src/solve.jl:578): NaN/Inf -> ReturnCode.Unstable, reported per trajectory
    if (a.status) {
#pragma unroll
        for (int tt = 0; tt < TPT; ++tt) {
            bool b = st.bad[tt];
#pragma unroll
            for (int j = 0; j < NS; ++j) b = b || !finite_d(st.u[tt][j]);
            if (tr0 + tt <= last) a.status[tr0 + tt] = b ? 1 : 0;
        }
    }
}

}  // namespace sdde
