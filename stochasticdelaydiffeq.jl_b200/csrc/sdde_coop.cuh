// sdde_coop.cuh -- cooperative-lanes step kernel: NS lanes of a warp integrate ONE trajectory (lane j owns state j).
//
// BASELINE.json's north_star asks for a warp-cooperative mapping for systems with >= 4 states.  Here a group of
// LPT = NS (4 or 8) adjacent lanes shares a trajectory: each lane loads / interpolates / stores only its own state
// component (ring layout [slot][traj][NS], state fastest, so a warp still touches 256 contiguous bytes), the full
// state and delayed-state vectors are exchanged with warp shuffles, and each lane evaluates only ITS row of the
// drift and of the noise matrix (models expose f_row / g_row).  The per-trajectory register footprint is 1/NS of the
// thread-per-trajectory kernel, which is what matters for larger systems; for NS = 4 the thread-per-trajectory kernel
// is faster (measured in profiles/README.md) and stays the default.
// Same schedule tables, same Philox streams (one pair per trajectory and call, evaluated redundantly by the group's
// lanes -- in SIMT the idle lanes would occupy the issue slot anyway), same results as sdde_solve_kernel.
#pragma once
#include "sdde_kernels.cuh"

namespace sdde {

template <class M, int ALG, int NOISE>
__global__ void __launch_bounds__(SDDE_BLOCK_COOP) sdde_coop_kernel(const __grid_constant__ KArgs a) {
    constexpr int NS = M::NS, NL = M::NL, NW = M::NW, LPT = NS;
    static_assert(!M::DIAG && M::HAS_ROWS, "cooperative kernel needs a row-wise model with a noise matrix");
    static_assert(LPT == 4 || LPT == 8, "lanes per trajectory must be 4 or 8");
    typedef StepRec<NL> Rec;
    constexpr int SDDE_BLOCK = SDDE_BLOCK_COOP;
    constexpr int TAB_BYTES = (NOISE == K_NOISE_PHILOX) ? SDDE_RNG_SMEM_BYTES(1) : 0;   // one copy: small CTAs, many per SM
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Rec *srec = reinterpret_cast<Rec *>(smem_raw + TAB_BYTES);
    const double2 *ltab = nullptr;
    if (NOISE == K_NOISE_PHILOX) ltab = stage_rng_table<1>(reinterpret_cast<double2 *>(smem_raw), a.rng_tab, SDDE_BLOCK);
    const int j = threadIdx.x % LPT;                                                        // my state component
    const long long tr = ((long long)blockIdx.x * SDDE_BLOCK + threadIdx.x) / LPT;          // wave-local trajectory
    const long long last = a.n_traj - 1;
    const long long trc = tr <= last ? tr : last;
    const bool live = tr <= last;
    const ConstParams p{a.p};

    double u = a.u0 ? a.u0[j * a.u0_pitch + trc] : a.u0c[j];
    double zc[4] = {0.0, 0.0, 0.0, 0.0};
    u64 zcall = ~0ull;
    double curW[NW];
#pragma unroll
    for (int c = 0; c < NW; ++c) curW[c] = (NOISE == K_NOISE_GRID) ? a.noise[(long long)c * a.noise_pitch + a.noise_off + trc] : 0.0;
    bool bad = false;
    int ks = 0;

    char *ringb = reinterpret_cast<char *>(a.ring + tr * NS + j);     // ring layout [slot][traj][NS]
    *reinterpret_cast<double *>(ringb) = u;                            // node 0
    while (ks < a.n_saves && a.saves[ks].step < 0) {
        if (a.out && live && a.save_row[j] >= 0) a.out[((long long)ks * a.n_saved + a.save_row[j]) * a.out_pitch + a.out_off + tr] = u;
        ++ks;
    }
    const double hc0 = a.hist_c0[j], hc1 = a.hist_c1[j];

    for (int n0 = 0; n0 < a.n_steps; n0 += a.chunk_steps) {
        const int cnt = (a.n_steps - n0 < a.chunk_steps) ? (a.n_steps - n0) : a.chunk_steps;
        __syncthreads();
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.steps + (size_t)n0 * sizeof(Rec));
            uint4 *dst = reinterpret_cast<uint4 *>(smem_raw + TAB_BYTES);
            const int nvec = cnt * (int)(sizeof(Rec) / 16);
            for (int i = threadIdx.x; i < nvec; i += SDDE_BLOCK) dst[i] = src[i];
        }
        __syncthreads();
        for (int i = 0; i < cnt; ++i) {
            const Rec &r = srec[i];
            const int n = n0 + i;
            const double tn = r.t, dtn = r.dt, sq = r.sqdt;
            // my component of every delayed state, then the full vectors by shuffle
            double ul[NL][NS], uu[NS];
#pragma unroll
            for (int l = 0; l < NL; ++l) {
                const double th = r.lk[l].theta;
                const double lo = *reinterpret_cast<const double *>(ringb + r.lk[l].off_lo);
                const double hi = *reinterpret_cast<const double *>(ringb + r.lk[l].off_hi);
                const double mine = (r.ring_mask & (1 << l)) ? fma(th, hi, (1.0 - th) * lo) : (hc0 + hc1 * th);
#pragma unroll
                for (int s = 0; s < NS; ++s) ul[l][s] = __shfl_sync(0xffffffffu, mine, s, LPT);
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) uu[s] = __shfl_sync(0xffffffffu, u, s, LPT);
            // increments (identical in all lanes of the group: same trajectory id)
            double dw[NW];
#pragma unroll
            for (int c = 0; c < NW; ++c) {
                if (NOISE == K_NOISE_PHILOX) {
                    const u64 q = (u64)n * NW + c, call = q >> 2;
                    if (call != zcall) { philox_normal_quad<1>((u64)(a.traj_gid0 + tr), call, a.rk, ltab, zc); zcall = call; }
                    const double zlo = (q & 1) ? zc[1] : zc[0], zhi = (q & 1) ? zc[3] : zc[2];
                    dw[c] = sq * ((q & 2) ? zhi : zlo);
                } else {
                    const long long row = (NOISE == K_NOISE_GRID) ? ((long long)(n + 1) * NW + c) : ((long long)n * NW + c);
                    const double w = a.noise[row * a.noise_pitch + a.noise_off + trc];
                    if (NOISE == K_NOISE_GRID) { dw[c] = w - curW[c]; curW[c] += dw[c]; } else dw[c] = w;
                }
            }
            // perform_step! (EM, non-diagonal noise), row j only
            const double uold = u;
            const double fj = M::f_row(j, uu, ul, p, tn);
            const double K = u + dtn * fj;
            double grow[NW];
            if (ALG == K_ALG_EM_SPLIT) {
                double Kf[NS];
#pragma unroll
                for (int s = 0; s < NS; ++s) Kf[s] = __shfl_sync(0xffffffffu, K, s, LPT);
                M::g_row(j, grow, Kf, K, ul, p, tn);
            } else {
                M::g_row(j, grow, uu, u, ul, p, tn);
            }
            double noise = grow[0] * dw[0];
#pragma unroll
            for (int c = 1; c < NW; ++c) noise = noise + grow[c] * dw[c];
            u = K + noise;
            *reinterpret_cast<double *>(ringb + r.off_store) = u;
            for (int s = 0; s < r.nsave; ++s) {
                const SaveRec sr = a.saves[ks];
                const double v = sr.exact ? u : fma(sr.theta, u, (1.0 - sr.theta) * uold);
                bad = bad || !finite_d(v);
                if (a.out && live && a.save_row[j] >= 0) a.out[((long long)ks * a.n_saved + a.save_row[j]) * a.out_pitch + a.out_off + tr] = v;
                ++ks;
            }
        }
    }
    bad = bad || !finite_d(u);
    // any lane of the group bad -> trajectory bad
    unsigned b = bad ? 1u : 0u;
#pragma unroll
    for (int o = LPT / 2; o > 0; o >>= 1) b |= __shfl_xor_sync(0xffffffffu, b, o, LPT);
    if (a.status && live && j == 0) a.status[tr] = (unsigned char)b;
}

}  // namespace sdde
