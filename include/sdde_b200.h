/* sdde_b200.h -- C ABI of libsdde_b200.so: B200-native ensemble integrator for constant-lag SDDEs.
 *
 * This is the drop-in boundary for ONE path of SciML/StochasticDelayDiffEq.jl (v1.12.0):
 *
 *     solve(EnsembleProblem(SDDEProblem(f, g, u0, h, tspan, p; constant_lags)), EM()/RKMil(), ensemblealg;
 *           dt, saveat, trajectories, adaptive=false)
 *
 * i.e. what the reference executes, once per trajectory, through
 *     DiffEqBase.__solve        /root/reference/src/solve.jl:1-9
 *     DiffEqBase.__init         /root/reference/src/solve.jl:11-572
 *     DiffEqBase.solve!         /root/reference/src/solve.jl:574-610
 * driven by [EXT] SciMLBase.__solve(::AbstractEnsembleProblem, alg, ensemblealg; trajectories).
 * The Julia side binds these entry points with `ccall` from a new ensemble algorithm type
 * (`EnsembleB200`, see INTEGRATION.md and julia/EnsembleB200.jl); in this repository the same ABI is
 * driven by Python ctypes (stochasticdelaydiffeq.jl_b200/_capi.py).
 *
 * All entry points are extern "C", take plain pointers and sizes, never throw and never call back
 * into the host language.  Return value 0 = success; negative = error, message via
 * sdde_b200_last_error().  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with SDDE_ERR_CUDA.
 */
#ifndef SDDE_B200_H
#define SDDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDDE_B200_VERSION 100      /* 0.1.0 */

#if defined(__GNUC__)
#define SDDE_API __attribute__((visibility("default")))
#else
#define SDDE_API
#endif

#define SDDE_MAX_STATE 8
#define SDDE_MAX_LAG 4
#define SDDE_MAX_NOISE 8
#define SDDE_MAX_PARAM 128

/* ---- error codes (SURVEY.md section 8(b) "Errors") ---- */
#define SDDE_OK 0
#define SDDE_ERR_BAD_DESC (-1)     /* malformed descriptor (sizes, NULL pointers, tf <= t0, dt == 0 with no tstops ...)
                                      mirrors error() guards of src/solve.jl:121-123 */
#define SDDE_ERR_UNSUPPORTED (-2)  /* valid SciML request outside this path: RKMil with non-diagonal noise,
                                      dt > min(lag) with a non-EM algorithm (src/solve.jl:581-585), unknown model */
#define SDDE_ERR_NOMEM (-3)
#define SDDE_ERR_COMPILE (-4)      /* NVRTC compilation of a user model failed (log in last_error) */
#define SDDE_ERR_CUDA (-100)       /* -100 - cudaError_t */
#define SDDE_ERR_NCCL (-200)       /* -200 - ncclResult_t */

/* ---- algorithms: the reference's alg argument ([EXT] StochasticDiffEq EM{split}, RKMil{interpretation};
 *      dispatch site src/solve.jl:586) ---- */
enum { SDDE_ALG_EM_SPLIT = 0,      /* EM() == EM(true) */
       SDDE_ALG_EM_NOSPLIT = 1,    /* EM(false) */
       SDDE_ALG_RKMIL_ITO = 2,     /* RKMil(), adaptive=false */
       SDDE_ALG_RKMIL_STRAT = 3,   /* RKMil(interpretation=:Stratonovich), adaptive=false */
       SDDE_ALG_EULER_HEUN = 4,    /* EulerHeun() (Stratonovich): second f/g evaluation at t+dt (test/test_prob_sol.jl:34) */
       SDDE_ALG_RKMILC_ITO = 5,    /* RKMilCommute() (test/test_prob_sol.jl:44-49): commutative noise; a diagonal model runs the */
       SDDE_ALG_RKMILC_STRAT = 6,  /*   RKMil kernel, a noise-matrix model evaluates G once more per noise column */
       SDDE_ALG_WANGLI3SMIL_A = 7, /* WangLi3SMil_A .. _F (test/test_prob_sol.jl:50-61), diagonal noise: three-stage Milstein methods, */
       SDDE_ALG_WANGLI3SMIL_B = 8, /*   restated as the six orders of (drift, Ito correction, Stratonovich-Milstein diffusion) -- */
       SDDE_ALG_WANGLI3SMIL_C = 9, /*   LOW confidence in the letter <-> order map, see DESIGN.md section 6 */
       SDDE_ALG_WANGLI3SMIL_D = 10,
       SDDE_ALG_WANGLI3SMIL_E = 11,
       SDDE_ALG_WANGLI3SMIL_F = 12,
       SDDE_ALG_SRA1 = 13,         /* SRA1(), SRA3(), SOSRA() with adaptive=false (test/test_prob_sol.jl:74-87 runs them adaptively): */
       SDDE_ALG_SRA3 = 14,         /*   Roessler's stochastic Runge-Kutta methods for ADDITIVE noise; a second increment dZ per step  */
       SDDE_ALG_SOSRA = 15,        /*   (src/solve.jl:308-312), f at t + c0_i dt and g at t + c1_i dt with their own history lookups  */
       SDDE_ALG_SRA_TABLEAU = 16 };/* SRA(tableau=...): sdde_desc.sra_tableau (SRA2, SOSRA2, ... as the caller's StochasticDiffEq builds them) */

/* ---- built-in model zoo: the device-side replacement of the closures that
 *      src/functionwrapper.jl:8-9,35-36 builds around the user's f(u,h,p,t), g(u,h,p,t).
 *      Ids >= SDDE_MODEL_USER_BASE are returned by sdde_b200_compile_model (NVRTC). ---- */
enum { SDDE_MODEL_LINEAR1 = 0,       /* du=(a u + b u(t-tau))dt + s u dW            p=(a,b,s)                 */
       SDDE_MODEL_HAYES = 1,         /* test/test_prob_sol.jl:6-13                   p=(tau,a,b,c,alpha,beta,gamma) */
       SDDE_MODEL_MACKEY_GLASS = 2,  /* du=(beta x/(1+x^n) - gamma u)dt + s u dW     p=(beta,gamma,n,s), x=u(t-tau) */
       SDDE_MODEL_LINSYS = 3,        /* n-state linear delay system, n x m noise     p=(A,B_1..B_L,c,M,N) row-major */
       SDDE_MODEL_SIR = 4,           /* test/nondiagonal_sparse_noise.jl:5-43        p=(beta,c,tau)            */
       SDDE_MODEL_SCALAR_ADD = 5,    /* test/events.jl:4-5  du=a u(t-tau)dt + b dW   p=(a,b)                   */
       SDDE_MODEL_USER_BASE = 1000 };

/* ---- noise source: src/solve.jl:304-353 (WienerProcess vs prob.noise).  Channels per step: n_chan = n_state (diagonal
 *      noise) or n_noise (noise matrix); the SRA algorithms use 2 n_chan rows per step, dW in [0, n_chan), dZ in
 *      [n_chan, 2 n_chan) (NoiseGrid(t, W, Z)) ---- */
enum { SDDE_NOISE_PHILOX = 0,        /* in-kernel Philox4x32-10 + float64 Box-Muller, dW = sqrt|dt_n| * xi          */
       SDDE_NOISE_SUPPLIED_DW = 1,   /* noise[n][c][traj] = increment of step n (already scaled)                    */
       SDDE_NOISE_SUPPLIED_GRID = 2  /* noise[n][c][traj] = W(t_n) on the schedule's own nodes n=0..N (NoiseGrid):  */
};                                   /*   dW_n = W(t_{n+1}) - curW, curW += dW_n  ([EXT] DiffEqNoiseProcess)        */

/* ---- delayed-state source.  Only DENSE exists on the GPU: the reference's saveat aliasing
 *      (src/solve.jl:465-482) is reproduced by the oracle only, as a documented hazard. ---- */
enum { SDDE_HISTORY_DENSE = 0 };

/* ---- where a caller buffer lives ---- */
enum { SDDE_MEM_HOST = 0, SDDE_MEM_DEVICE = 1 };

/* Problem + solve-keyword descriptor.  POD; the caller owns every pointer for the duration of the call.
 * Field <-> reference map:
 *   model_id,params,u0,hist_c0,hist_c1,lags  SDDEProblem fields f,g,p,u0,h,constant_lags      src/solve.jl:128
 *   t0,tf                           prob.tspan                                       src/solve.jl:116-119
 *   dt, saveat, save_*, tstops      __init keyword arguments                         src/solve.jl:18-40
 *   order_discontinuity_t0          prob.order_discontinuity_t0                      src/solve.jl:100-105
 *   seed                            seed kw / prob.seed                              src/solve.jl:302
 *   noise_mode/noise                prob.noise                                       src/solve.jl:304-353
 *   n_traj                          `trajectories` of the ensemble solve ([EXT] SciMLBase)
 *   params+param_stride / u0+stride what EnsembleProblem.prob_func would vary per trajectory            */
typedef struct sdde_desc {
    int32_t struct_size;              /* = sizeof(sdde_desc); ABI guard */
    int32_t model_id;
    int32_t alg;
    int32_t n_state, n_noise, n_lag, n_param;
    int32_t order_discontinuity_t0;   /* default 0 */
    int32_t constant_lags_declared;   /* 1: lags were given as constant_lags (propagated-discontinuity tstops) */
    int32_t save_start, save_end, save_everystep;
    int32_t noise_mode;
    int32_t noise_mem;                /* SDDE_MEM_* of `noise` */
    int32_t out_mem;                  /* SDDE_MEM_* of `out` passed to sdde_b200_solve */
    int32_t history_mode;
    int32_t want_paths, want_moments;
    int32_t lanes_per_traj;           /* 0/1 = thread per trajectory (default); n_state (4) = cooperative lanes, one lane per
                                         state (sdde_coop.cuh; models with a row-wise form, shared parameters) */
    int32_t n_saveat, n_tstops;
    int32_t reserved0;
    double t0, tf, dt;
    double lags[SDDE_MAX_LAG];
    double hist_c0[SDDE_MAX_STATE];   /* h(p,t)[j] = hist_c0[j] + hist_c1[j]*t  for t < t0 */
    double hist_c1[SDDE_MAX_STATE];
    const double *params;             /* [n_param] if param_stride == 0 else params[k*param_stride + traj] */
    int64_t param_stride;
    const double *u0;                 /* [n_state] if u0_stride == 0 else u0[j*u0_stride + traj] */
    int64_t u0_stride;
    const double *saveat;             /* explicit save times (already expanded; src/solve.jl:659-677) */
    const double *tstops;             /* extra user tstops */
    const double *noise;              /* supplied noise, layout [n][channel][traj], row pitch = n_traj */
    int64_t n_traj;                   /* trajectories solved by THIS call */
    int64_t traj_offset;              /* global id of the first one (partition-invariant RNG streams) */
    uint64_t seed;
    int64_t out_pitch;                /* elements between consecutive rows of `out`; 0 = n_traj (dense) */
    int32_t n_save_idxs;              /* save_idxs keyword (src/solve.jl:355-356): 0 = save every state, else the number of */
    int32_t save_idxs[SDDE_MAX_STATE];/*   0-based state indices that are saved (out/moments then have n_save_idxs rows per save) */
    int32_t reserved1;
    const double *noise_t;            /* SDDE_NOISE_SUPPLIED_GRID only: NULL = `noise` holds W on the schedule's own nodes
                                         ([n_steps+1] rows per channel); else the n_noise_t ascending grid times of a
                                         NoiseGrid(t, W) on ANY grid covering [t0, tf] -- W(t) is then interpolated
                                         linearly at the step ends like [EXT] DiffEqNoiseProcess NoiseGrid does */
    int64_t n_noise_t;
    const double *d_disc_t;           /* d_discontinuities keyword (src/solve.jl:20-23): times of user-declared discontinuities; */
    const int32_t *d_disc_order2;     /*   2 x order of each (Discontinuity{tType, Rational{Int}}: orders are multiples of 1/2 here). */
    int64_t n_d_disc;                 /*   They become tstops (src/solve.jl:644-647) and seed the propagation heap (:688-699). */
    int64_t out_rows;                 /* capacity guard: rows (save points x saved states) the caller's out / moments buffers hold;
                                         0 = unchecked.  A solve that needs more rows fails with SDDE_ERR_BAD_DESC instead of
                                         writing past the buffers (e.g. save_everystep with a buffer sized for saveat). */
    const double *sra_tableau;        /* SDDE_ALG_SRA_TABLEAU: [c0(s) | c1(s) | A0(s x s, row i = stage i) | B0(s x s) | alpha(s) | beta1(s) |
                                         beta2(s)] of an explicit Roessler SRA tableau, s = sra_stages <= 3, at most 4 distinct stage
                                         times ([EXT] StochasticDiffEq RosslerSRA fields c0, c1, A0', B0', alpha, beta1, beta2) */
    int32_t sra_stages;
    int32_t sra_order2;               /* 2 x alg_order of the tableau (discontinuity propagation depth); 0 = 4 ([EXT] alg_order 2//1) */
} sdde_desc;

/* Host-side schedule (trajectory independent): the replay of src/solve.jl:574-610 time stepping with
 * tstops / discontinuity propagation (src/integrators/utils.jl) and the saveat logic of
 * src/integrators/interface.jl:357-430.  Arrays are owned by the library, freed by sdde_b200_schedule_free. */
typedef struct sdde_schedule {
    int64_t n_steps;
    int64_t n_save;
    int32_t n_lag;
    int32_t ring_slots;        /* history ring capacity the kernels use */
    int32_t min_gap;           /* min over steps of (n - newest node a lookup touches) */
    int32_t n_tstops_hit;
    int32_t n_stage;           /* history lookup times per step (2 for EulerHeun: t_n and t_n + dt_n) */
    int32_t reserved;
    double *t;                 /* [n_steps+1] */
    double *dt;                /* [n_steps]   */
    double *sqdt;              /* [n_steps]   */
    int32_t *lk_mode;          /* [n_steps][n_stage][n_lag] 0 prehistory, 1 exact node, 2 interpolate, 3 constant u0 */
    int64_t *lk_idx;           /* same shape: node index (upper node when interpolating), -1 otherwise */
    double *lk_theta;          /* same shape: theta, or the lookup time for prehistory */
    double *save_t;            /* [n_save] */
    int64_t *save_step;        /* [n_save] step whose end produced the save, -1 = initial condition */
    double *save_theta;        /* [n_save] */
    int32_t *save_exact;       /* [n_save] */
    double *tstops_hit;        /* [n_tstops_hit] */
} sdde_schedule;

typedef struct sdde_timing {
    double kernel_ms;          /* CUDA-event time of the step kernels (+ moments kernels) on the compute stream */
    double total_ms;           /* host wall time of the call */
    double h2d_ms, d2h_ms;     /* CUDA-event time on the copy streams */
    int64_t h2d_bytes, d2h_bytes;
    int64_t n_steps, n_save;
    int32_t launches;          /* kernels launched by this call */
    int32_t waves;             /* trajectory waves */
    int32_t ring_slots;
    int32_t regs_per_thread;
    int32_t block_threads;
    int32_t traj_per_thread;
    int32_t moment_merge;      /* sdde_b200_solve_multi: 0 = none, 1 = host Chan merge, 2 = NCCL all-gather + device merge */
    int32_t n_unstable;        /* trajectories of this call that saw a non-finite state (counted on the device, so a caller
                                  that passes status = NULL -- e.g. a moments-only run of 10^8 trajectories -- still learns
                                  it without a byte per trajectory crossing PCIe) */
    int64_t ring_bytes;
    char kernel_name[128];
} sdde_timing;

typedef struct sdde_ctx sdde_ctx;

SDDE_API int sdde_b200_version(void);
SDDE_API int sdde_b200_device_count(void);                       /* number of CUDA devices visible, <0 on error */
SDDE_API const char *sdde_b200_last_error(void);                 /* thread-local message of the last failing call */

/* One context drives one GPU (one process per GPU under torchrun; several contexts per process are fine). */
SDDE_API sdde_ctx *sdde_b200_create(int device);
SDDE_API void sdde_b200_destroy(sdde_ctx *ctx);
SDDE_API int sdde_b200_set_option(sdde_ctx *ctx, const char *key, int64_t value);  /* "wave_traj" (0 = auto), "batch" (0 = compiled-in
                                                                                    prefetch depth), "mem_limit" (bytes the device workspaces may use, 0 = 80 % of free memory),
                                                                                    "chunk_steps" (schedule records staged per
                                                                                    shared-memory refill, 1..1024), "ramp_start" (trajectories of the first wave when the
                                                                                    paths leave over PCIe; later waves grow x1.5; default 2^16, measured best on C2),
                                                                                    "host_threads" (host threads that copy a PAGEABLE host output out of the
                                                                                    page-locked bounce buffers; 0 = min(16, hardware threads), 1 = plain cudaMemcpy2DAsync), "bounce_min_mb" (smallest pageable
                                                                                    output that takes that path, default 32; 0 = always, used by the tests),
                                                                                    "nccl", "jit_linsys" (0/1), "zero_copy" (1: page-locked host `out` is written by the step kernel itself
                                                                                    over PCIe; 0: staged in HBM and copied wave by wave; -1 (default): small outputs zero-copy, outputs
                                                                                    >= 64 MB try both on the first two solves of their size class (power of two of the output bytes) and keep the faster) */

/* Host only: builds the step/lookup/save schedule for a descriptor (no GPU needed). */
SDDE_API int sdde_b200_schedule(const sdde_desc *d, sdde_schedule *out);
SDDE_API void sdde_b200_schedule_free(sdde_schedule *s);

/* Integrate d->n_traj trajectories.
 *   out      [n_save][n_state][n_traj] float64, time-major, trajectory-contiguous (may be NULL if !want_paths);
 *            host or device memory according to d->out_mem.
 *   moments  [n_save][n_state][3] = (count, mean, M2) over this call's trajectories, host memory (NULL if !want_moments)
 *   status   [n_traj] bytes, host memory, 0 = ok, 1 = non-finite state seen (-> ReturnCode.Unstable), may be NULL
 *   tm       optional timing/launch report                                                          */
SDDE_API int sdde_b200_solve(sdde_ctx *ctx, const sdde_desc *d, double *out, double *moments, uint8_t *status, sdde_timing *tm);

/* Same as sdde_b200_solve, over several GPUs from ONE host process (the Julia `ccall` case): the trajectories are split
 * into contiguous global-id ranges [r*ceil(N/R), ...), one per context, each range is solved concurrently on its GPU
 * (one host thread per context), paths land in the caller's `out` (host memory only), and the per-GPU moment triples
 * are all-gathered over NCCL (communicators from ncclCommInitAll, libnccl.so.2 resolved at run time) and Chan-merged in
 * rank order on the device; without NCCL, or with option "nccl" = 0, they are merged on the host.  Results are identical to a single-GPU solve (partition-invariant Philox
 * streams).  Under torchrun (one process per GPU) use sdde_b200_solve per rank + an NCCL all-gather of the triples. */
SDDE_API int sdde_b200_solve_multi(sdde_ctx *const *ctxs, int32_t n_ctx, const sdde_desc *d, double *out, double *moments,
                                   uint8_t *status, sdde_timing *tm /* [n_ctx] or NULL */);

/* Merge two (count, mean, M2) arrays of n triples: a <- merge(a, b)  (Chan et al.); host only. */
SDDE_API int sdde_b200_merge_moments(double *a, const double *b, int64_t n);

/* (count, mean, M2) across trajectories of saved rows that already exist (e.g. a device-resident `out` of an earlier solve that
 * did not ask for moments): the same one-read kernels sdde_b200_solve uses; [EXT] EnsembleAnalysis.timeseries_steps_meanvar.
 *   paths [n_rows][pitch] host or device (`mem`), first n_traj values per row; moments [n_rows][3] host memory. */
SDDE_API int sdde_b200_moments(sdde_ctx *ctx, const double *paths, int32_t mem, int64_t n_rows, int64_t n_traj, int64_t pitch, double *moments);

/* Order statistics across trajectories of saved rows, on the device: the reference side computes them with
 * [EXT] SciMLBase EnsembleSummary -> EnsembleAnalysis.timeseries_steps_quantile / timeseries_steps_median ->
 * Statistics.quantile(v, p) (alpha = beta = 1) / Statistics.median(v) on the per-time-point vectors of a solved ensemble.
 *   paths   [n_rows][pitch] float64 (the `out` of sdde_b200_solve viewed as n_save*n_state rows, pitch = n_traj), host or
 *           device memory according to `mem`; the first n_traj values of each row are used
 *   probs   n_prob probabilities in [0, 1]
 *   result  [n_rows][n_prob + (want_median != 0)] host memory; the median is the last column.  A row holding a NaN yields NaN.
 *   kernel_ms  optional: CUDA-event time of the selection kernels
 * Exact (bit-identical to selecting from the sorted row): radix selection on the order-preserving integer image. */
SDDE_API int sdde_b200_quantiles(sdde_ctx *ctx, const double *paths, int32_t mem, int64_t n_rows, int64_t n_traj, int64_t pitch,
                                 const double *probs, int32_t n_prob, int32_t want_median, double *result, double *kernel_ms);

/* The same selection in phases, for an ensemble whose trajectories are sharded over several processes / GPUs (one
 * context each, e.g. under torchrun): every rank histograms ITS columns of the rows, the caller sums the histograms over
 * the ranks (ncclAllReduce / torch.distributed.all_reduce on `d_hist`), every rank applies the same bin decisions, and
 * after QS passes = 6 rounds every rank holds the order statistics of the WHOLE ensemble (exact).
 *   sdde_b200_qsel_ranks   host only: how many distinct order statistics (n_rank) the request needs for n_total values
 *   sdde_b200_qsel_begin   paths = this rank's [n_rows][pitch] block (n_local columns used, may be 0); d_hist = a device
 *                          buffer of n_rows * n_rank * 2048 uint32 owned by the caller
 *   sdde_b200_qsel_local   pass p: zero d_hist, add this rank's histograms, synchronise          -> caller all-reduces d_hist
 *   sdde_b200_qsel_apply   pass p: narrow every rank's prefix by one digit from the summed histograms
 *   sdde_b200_qsel_nan     set = 0: counts[n_rows] <- NaNs seen by this rank; set = 1: install the summed counts
 *   sdde_b200_qsel_end     result [n_rows][n_prob + (want_median != 0)], host memory; closes the selection          */
SDDE_API int sdde_b200_qsel_ranks(int64_t n_total, const double *probs, int32_t n_prob, int32_t want_median, int32_t *n_rank);
SDDE_API int sdde_b200_qsel_begin(sdde_ctx *ctx, const double *paths, int32_t mem, int64_t n_rows, int64_t n_local, int64_t pitch,
                                  int64_t n_total, const double *probs, int32_t n_prob, int32_t want_median, uint32_t *d_hist);
SDDE_API int sdde_b200_qsel_local(sdde_ctx *ctx, int32_t pass);
SDDE_API int sdde_b200_qsel_apply(sdde_ctx *ctx, int32_t pass);
SDDE_API int sdde_b200_qsel_nan(sdde_ctx *ctx, int64_t *counts, int32_t set);
SDDE_API int sdde_b200_qsel_end(sdde_ctx *ctx, double *result);

/* Register a user model from CUDA C++ source (NVRTC, sm_100a).  The source must define
 *   struct UserModel { static constexpr int NS, NL, NW, NP; static constexpr bool DIAG;
 *                      template<class P> __device__ static void f(double*du,const double*u,const double(*ul)[NS],const P&p,double t);
 *                      template<class P> __device__ static void g(double*gv,const double*u,const double(*ul)[NS],const P&p,double t); };
 * Optional history function for t < t0 (any expression, not only affine): add
 *   static constexpr bool HAS_H = true;
 *   template<class P> __device__ static void h(double*out,const P&p,double t);
 * to the struct; it then replaces the descriptor's hist_c0/hist_c1 (the reference's h(p, t), src/functionwrapper.jl:35-36).
 * Returns a model id >= SDDE_MODEL_USER_BASE in *model_id. */
SDDE_API int sdde_b200_compile_model(sdde_ctx *ctx, const char *source, int32_t *model_id);

/* Test hooks / micro-benchmarks. */
SDDE_API int sdde_b200_philox_normals(sdde_ctx *ctx, uint64_t seed, int64_t traj_offset, int64_t n_traj,
                             int64_t n_steps, int32_t n_noise, double *out_host /* [n_steps][n_noise][n_traj] */);
/* Host copy of the normal generator's lookup table (layout: csrc/sdde_rng_tables.cuh); needs no device.
 * out may be NULL to query *n_entries (double2 entries; out needs 2 * n_entries doubles). */
/* The Roessler SRA tableau the library uses for alg = SDDE_ALG_SRA1 / SRA3 / SOSRA, in the layout of sdde_desc.sra_tableau
 * (out holds 5 s + 2 s^2 <= 33 doubles): lets the caller's side compare it with its own StochasticDiffEq constructSRA1 /
 * constructSRA3 / constructSOSRA before trusting a solve ([EXT] coefficients restated from memory, julia/replay_fixtures.jl). */
SDDE_API int sdde_b200_sra_tableau(int32_t alg, double *out, int32_t *stages);
SDDE_API int sdde_b200_rng_table(double *out, int64_t cap_doubles, int64_t *n_entries);
SDDE_API int sdde_b200_dfma_peak(sdde_ctx *ctx, double *tflops);       /* measured FP64 FMA throughput of this GPU */
SDDE_API int sdde_b200_hbm_peak(sdde_ctx *ctx, double *gbs);           /* measured device copy bandwidth (read+write) */

/* Pinned host memory helpers for the end-to-end path (Julia: unsafe_wrap the pointer). */
SDDE_API void *sdde_b200_alloc_pinned(int64_t bytes);
SDDE_API void sdde_b200_free_pinned(void *p);

#ifdef __cplusplus
}
#endif
#endif /* SDDE_B200_H */
