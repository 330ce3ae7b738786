// sdde_summary.cuh -- exact order statistics across trajectories of the saved rows (sm_100a).
//
// Device half of sdde_b200_quantiles: for every saved row out[k][j][0..n_traj) select a handful of order statistics
// (the neighbours Statistics.quantile / Statistics.median need, see summary.hpp) WITHOUT sorting the row:
// most-significant-digit radix selection on the order-preserving 64-bit image of the doubles, 11 bits per pass
// (at most 6 passes).  Every pass streams the rows once, coalesced, builds per-CTA histograms in shared memory
// (the saved states of one time point share sign, exponent and leading mantissa bits, so
// per-thread runs of one bin are accumulated in a register first), and a small scan kernel narrows every rank to one bin.
// Ranks whose selected prefixes still coincide share one histogram ("leaders"), so median + two quantiles cost
// three histograms in the early passes, not six.  HBM-bound: 6 x 8 B per saved value.
#pragma once
#include <cstdint>

namespace sdde {

constexpr int QS_BITS = 11;
constexpr int QS_BINS = 1 << QS_BITS;
constexpr int QS_PASSES = 6;              // 5 x 11 bits + 9 bits
constexpr int QS_MAX_RANK = 12;
constexpr int QS_THREADS = 512;
constexpr int QS_UNROLL = 4;              // independent 256 B warp loads in flight per warp
constexpr int QS_CAP = 512;               // a rank whose bin holds <= QS_CAP values is finished from a gathered candidate list

enum { QS_MODE_HIST = 0, QS_MODE_GATHER = 1, QS_MODE_DONE = 2 };

struct QsRow {                             // selection state of one row
    unsigned long long prefix[QS_MAX_RANK];   // digits selected so far (right-aligned); the full key once the row is done
    long long k[QS_MAX_RANK];                 // rank among the values that share the prefix
    int leader[QS_MAX_RANK];                  // first rank with the same prefix (owns the histogram / candidate list)
    unsigned int n_cand[QS_MAX_RANK];         // gathered candidates per leader
    unsigned long long n_nan;
    int mode, pad;
};

struct QsRanks { long long k[QS_MAX_RANK]; };

__device__ __forceinline__ unsigned long long qs_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline int qs_shift(int pass) { return pass < QS_PASSES - 1 ? 64 - QS_BITS * (pass + 1) : 0; }
__host__ __device__ inline int qs_width(int pass) { return pass < QS_PASSES - 1 ? QS_BITS : 64 - QS_BITS * (QS_PASSES - 1); }

__global__ void qs_init_kernel(QsRow *rows, int n_rank, QsRanks ranks) {
    QsRow &q = rows[blockIdx.x];
    if (threadIdx.x < QS_MAX_RANK) {
        q.prefix[threadIdx.x] = 0ull;
        q.k[threadIdx.x] = threadIdx.x < n_rank ? ranks.k[threadIdx.x] : 0;
        q.leader[threadIdx.x] = 0;
        q.n_cand[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) { q.n_nan = 0ull; q.mode = QS_MODE_HIST; }
}

// One streaming pass over the rows.  grid (n_rows, n_split); dynamic shared memory n_rank * QS_BINS counters.
//   row in HIST mode   : histogram of digit PASS of the values that share a leader's prefix
//   row in GATHER mode : append those values' keys to the leader's candidate list (they are few)
//   row DONE           : nothing
// From the second pass on almost no value belongs to any leader, so the inner loop first looks the value's PREVIOUS
// digit up in a 2048-entry table of "leaders that chose this digit" and only then compares whole prefixes.  Histogram
// increments go through a thread-private run (bin, count): the saved states of one time point share sign, exponent and
// leading mantissa bits, so in the first pass a thread sees one bin for its whole slice and issues one atomic at the end.
template <int PASS>
__global__ void __launch_bounds__(QS_THREADS, 3) qs_stream_kernel(const double *__restrict__ paths, long long pitch, long long n,
                                                                  int n_rank, QsRow *__restrict__ rows,
                                                                  unsigned int *__restrict__ ghist,
                                                                  unsigned long long *__restrict__ cand) {
    extern __shared__ unsigned int sh[];
    __shared__ unsigned long long s_prefix[QS_MAX_RANK];
    __shared__ int s_slot[QS_MAX_RANK];
    __shared__ int s_nlead;
    __shared__ unsigned short s_tab[QS_BINS];
    constexpr int width = PASS < QS_PASSES - 1 ? QS_BITS : 64 - QS_BITS * (QS_PASSES - 1);
    constexpr int shift = PASS < QS_PASSES - 1 ? 64 - QS_BITS * (PASS + 1) : 0;
    const int row = blockIdx.x;                 // rows on grid.x (may exceed 65535), slices of a row on grid.y
    QsRow &q = rows[row];
    const int mode = q.mode;
    if (mode == QS_MODE_DONE) return;
    for (int i = threadIdx.x; i < QS_BINS; i += QS_THREADS) s_tab[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        int nl = 0;
        for (int r = 0; r < n_rank; ++r)
            if (q.leader[r] == r) {
                s_prefix[nl] = q.prefix[r]; s_slot[nl] = r;
                s_tab[q.prefix[r] & (QS_BINS - 1)] |= (unsigned short)(1u << nl);
                ++nl;
            }
        s_nlead = nl;
    }
    __syncthreads();
    const int nlead = s_nlead;
    const bool gather = mode == QS_MODE_GATHER;
    if (!gather) {
        for (int i = threadIdx.x; i < nlead * QS_BINS; i += QS_THREADS) sh[i] = 0u;
        __syncthreads();
    }

    const double *src = paths + (long long)row * pitch;
    const long long chunk = (long long)QS_THREADS * QS_UNROLL;
    const long long per = ((n + gridDim.y - 1) / gridDim.y + chunk - 1) / chunk * chunk;
    const long long lo = per * blockIdx.y, hi = lo + per < n ? lo + per : n;
    unsigned nan_seen = 0u;
    unsigned run_bin = 0u, run_cnt = 0u;          // thread-private run of identical (leader, digit) hits
    for (long long base = lo + threadIdx.x; base < hi; base += chunk) {
        double x[QS_UNROLL];
#pragma unroll
        for (int j = 0; j < QS_UNROLL; ++j) {
            const long long i = base + (long long)j * QS_THREADS;
            x[j] = i < hi ? __ldg(src + i) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < QS_UNROLL; ++j) {
            if (base + (long long)j * QS_THREADS >= hi) break;
            const unsigned long long key = qs_key(x[j]);
            const unsigned digit = (unsigned)(key >> shift) & (unsigned)((1 << width) - 1);
            unsigned mine = 1u;                        // leaders this value may belong to
            unsigned long long head = 0ull;
            if constexpr (PASS == 0) {
                nan_seen += x[j] != x[j];
            } else {
                head = key >> (shift + width);
                mine = (unsigned)s_tab[(unsigned)head & (QS_BINS - 1)];
            }
            while (mine) {
                const int l = __ffs(mine) - 1;
                mine &= mine - 1;
                if (head != s_prefix[l]) continue;
                if (gather) {
                    const unsigned pos = atomicAdd(&q.n_cand[s_slot[l]], 1u);
                    if (pos < (unsigned)QS_CAP) cand[((long long)row * n_rank + s_slot[l]) * QS_CAP + pos] = key;
                } else {
                    const unsigned bin = (unsigned)l * QS_BINS + digit;
                    if (bin != run_bin) {
                        if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
                        run_bin = bin; run_cnt = 0u;
                    }
                    ++run_cnt;
                }
            }
        }
    }
    if (gather) return;
    if (run_cnt) atomicAdd(&sh[run_bin], run_cnt);
    __syncthreads();
    for (int i = threadIdx.x; i < nlead * QS_BINS; i += QS_THREADS) {
        const unsigned v = sh[i];
        if (v) atomicAdd(&ghist[((long long)row * n_rank + s_slot[i >> QS_BITS]) * QS_BINS + (i & (QS_BINS - 1))], v);
    }
    if (PASS == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nan_seen += __shfl_xor_sync(0xffffffffu, nan_seen, o);
        if ((threadIdx.x & 31) == 0 && nan_seen) atomicAdd(&q.n_nan, (unsigned long long)nan_seen);
    }
}

typedef void (*QsStreamFn)(const double *, long long, long long, int, QsRow *, unsigned int *, unsigned long long *);
inline QsStreamFn qs_stream_fn(int pass) {
    switch (pass) {
        case 0: return qs_stream_kernel<0>;
        case 1: return qs_stream_kernel<1>;
        case 2: return qs_stream_kernel<2>;
        case 3: return qs_stream_kernel<3>;
        case 4: return qs_stream_kernel<4>;
        default: return qs_stream_kernel<5>;
    }
}

// One CTA per row, one warp per rank.
//   HIST mode  : find the bin holding rank k and descend into it; regroup equal prefixes; when every rank's bin is small,
//                switch the row to GATHER for the next streaming pass
//   GATHER mode: the leader's candidate list holds every value with the rank's prefix: pick the k-th by counting
__global__ void __launch_bounds__(32 * QS_MAX_RANK) qs_scan_kernel(QsRow *__restrict__ rows, const unsigned int *__restrict__ ghist,
                                                                   const unsigned long long *__restrict__ cand, int n_rank, int pass, int allow_gather) {
    __shared__ unsigned long long s_cand[QS_CAP];
    __shared__ int s_big;
    const int row = blockIdx.x, r = threadIdx.x >> 5, lane = threadIdx.x & 31;
    QsRow &q = rows[row];
    const int mode = q.mode;
    if (mode == QS_MODE_DONE) return;
    if (mode == QS_MODE_GATHER) {
        // ranks one after the other (the CTA shares one staging buffer); lists are short
        for (int rr = 0; rr < n_rank; ++rr) {
            const int L = q.leader[rr];
            const int c = (int)min(q.n_cand[L], (unsigned)QS_CAP);
            __syncthreads();
            for (int i = threadIdx.x; i < c; i += blockDim.x) s_cand[i] = cand[((long long)row * n_rank + L) * QS_CAP + i];
            __syncthreads();
            const long long k = q.k[rr];
            for (int i = threadIdx.x; i < c; i += blockDim.x) {
                const unsigned long long mine = s_cand[i];
                long long below = 0;
                for (int j = 0; j < c; ++j) {
                    const unsigned long long o = s_cand[j];
                    below += (o < mine) || (o == mine && j < i);
                }
                if (below == k) q.prefix[rr] = mine;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) q.mode = QS_MODE_DONE;
        return;
    }
    if (threadIdx.x == 0) s_big = 0;
    __syncthreads();
    const int width = qs_width(pass), per = (1 << width) / 32;
    if (r < n_rank) {
        const unsigned int *h = ghist + ((long long)row * n_rank + q.leader[r]) * QS_BINS + lane * per;
        const long long k = q.k[r];
        long long sum = 0;
        for (int i = 0; i < per; ++i) sum += h[i];
        long long incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        long long c = incl - sum;
        if (k >= c && k < incl) {
            int i = 0;
            for (; i < per - 1; ++i) {
                if (k < c + (long long)h[i]) break;
                c += h[i];
            }
            q.prefix[r] = (q.prefix[r] << width) | (unsigned long long)(lane * per + i);
            q.k[r] = k - c;
            if (h[i] > (unsigned)QS_CAP) s_big = 1;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int a = 0; a < n_rank; ++a) {
            int lead = a;
            for (int b = 0; b < a; ++b)
                if (q.prefix[b] == q.prefix[a]) { lead = b; break; }
            q.leader[a] = lead;
            q.n_cand[a] = 0u;
        }
        if (pass == QS_PASSES - 1) q.mode = QS_MODE_DONE;
        else if (allow_gather && !s_big && pass < QS_PASSES - 2) q.mode = QS_MODE_GATHER;
    }
}

}  // namespace sdde
