// unb_scan.cu -- candidate-vs-live-point pair scans (find_nearby, count_nearby,
// _subtract_nearby, compute_maxradiussq of ultranest/mlfriends.pyx) as sm_100a kernels.
//
// Arithmetic contract: the reference decides every pair with the
// k-sequential, non-fused fp64 distance  D = (((0 + (a0-b0)^2) + (a1-b1)^2) + ...)  and `<=`.
// Evaluating D costs 3 DP instructions per pair-dimension.  The kernels here get the SAME
// decisions with ~1 DFMA per pair-dimension:
//
//   filter : acc = h_i + sum_k a_ik * b_jk  (one DFMA per k, h_i rides in as the first addend)
//            with the identity  a.b + (r2 - |a|^2)/2 - |b|^2/2 = (r2 - D_true)/2.
//            A pair can only satisfy D <= r2 if acc >= thr_j, once h_i / thr_j are widened by
//            kappa * (|a|^2 + |b|^2 + r2), kappa = (8d+64) * 2^-53, which dominates every
//            rounding error of the filter AND of the reference's own sum (bound derived in
//            DESIGN.md "Filter bound").  The comparison uses only the high word of acc
//            (integer ISETP, off the DP pipe) with one more unit of slack.
//   decide : pairs that pass the filter (a few per candidate) are re-evaluated with the
//            reference's exact sequence (sq_step) and compared with `<=` -- so outputs are
//            bit-identical to the Cython, independent of the filter.
//
// The live block is streamed through shared memory tile by tile with TMA 1-D bulk copies
// (cp.async.bulk + mbarrier, double buffered); candidates live in registers (d <= 32) or in
// shared memory (d > 32).  No tensor cores: this is an fp64 CUDA-core scan.
#include "unb_internal.cuh"

#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int SCAN_THREADS = 128;
constexpr int TN = 4;               // live points per register-tile column group
constexpr int REG_TILE_N = 64;      // live points per shared-memory tile (register kernel)
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ double pos_inf() { return __longlong_as_double(0x7ff0000000000000LL); }

// ---------------------------------------------------------------------------------------
// live-block layout kernels
// ---------------------------------------------------------------------------------------

// one thread per tile slot: writes the coordinate rows and the squared norm
__global__ void k_live_build(const double *__restrict__ rows, int n, int d, int dr, int tile_n,
                             int ntiles, double *__restrict__ tiles, double *__restrict__ norms)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= ntiles * tile_n) return;
    int t = slot / tile_n, c = slot - t * tile_n;
    double *T = tiles + (size_t)t * (dr + 1) * tile_n + c;
    double na = 0.0;
    for (int k = 0; k < dr; k++) {
        double v = (slot < n && k < d) ? rows[(size_t)slot * d + k] : 0.0;
        T[(size_t)k * tile_n] = v;
        na = fma(v, v, na);
    }
    norms[slot] = (slot < n) ? na : 0.0;
}

__global__ void k_norm_max(const double *__restrict__ norms, int n, unsigned long long *out)
{
    double m = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        m = fmax(m, norms[i]);
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, f64_bits(m));
}

// h row of every tile.  THRESH: h_i = ((1+kappa) r2 - (1-kappa) |a_i|^2) / 2,  MIN: -|a_i|^2/2,
// padding slots get -1e300 so they can never pass a filter.
__global__ void k_live_set_h(const double *__restrict__ norms, int n, int dr, int tile_n,
                             int ntiles, int h_mode, double r2, double kappa,
                             double *__restrict__ tiles)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= ntiles * tile_n) return;
    int t = slot / tile_n, c = slot - t * tile_n;
    double h = -1e300;
    if (slot < n) {
        double na = norms[slot];
        if (h_mode == HMODE_THRESH) {
            double r2w = __dmul_rn(r2, __dadd_rn(1.0, kappa));
            double naw = __dmul_rn(na, __dsub_rn(1.0, kappa));
            h = __dmul_rn(0.5, __dsub_rn(r2w, naw));
        } else {
            h = __dmul_rn(-0.5, na);
        }
    }
    tiles[(size_t)t * (dr + 1) * tile_n + (size_t)dr * tile_n + c] = h;
}

// bootstrap: per-round compacted tiles (selected rows only, original order) with MIN-mode h
__global__ void k_gather_round_tiles(const double *__restrict__ rows, int d, int dr, int tile_n,
                                     const int *__restrict__ idxA, const int *__restrict__ offA,
                                     const int *__restrict__ nA, long long round_tile_stride,
                                     double *__restrict__ tiles)
{
    int round = blockIdx.y;
    int na = nA[round];
    int ntiles = (na + tile_n - 1) / tile_n;
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= ntiles * tile_n) return;
    int t = slot / tile_n, c = slot - t * tile_n;
    double *T = tiles + (size_t)round * round_tile_stride + (size_t)t * (dr + 1) * tile_n + c;
    int src = (slot < na) ? idxA[offA[round] + slot] : -1;
    double nrm = 0.0;
    for (int k = 0; k < dr; k++) {
        double v = (src >= 0 && k < d) ? rows[(size_t)src * d + k] : 0.0;
        T[(size_t)k * tile_n] = v;
        nrm = fma(v, v, nrm);
    }
    T[(size_t)dr * tile_n] = (src >= 0) ? __dmul_rn(-0.5, nrm) : -1e300;
}

// ---------------------------------------------------------------------------------------
// per-candidate scan state
// ---------------------------------------------------------------------------------------
template <int MODE>
struct CState {
    int valid;
    int first;      // FIND
    int thrkey;     // FIND / COUNT / SUBTRACT: high word of the filter threshold, minus 1
    int cnt;        // COUNT / SUBTRACT
    double best;    // MIN: exact minimum so far
    double amax;    // MIN: largest filter value seen
    double thr;     // MIN: amax - slack
    double slack;   // MIN
};

template <int MODE>
__device__ __forceinline__ void cs_init(CState<MODE> &s, bool valid, double nb,
                                        const ScanArgs &A)
{
    s.valid = valid;
    s.first = -1;
    s.cnt = 0;
    s.best = 1e300;
    s.amax = -1e300;
    if (MODE == SCAN_MIN) {
        double namax = __longlong_as_double((long long)*A.namax_bits);
        s.slack = __dmul_rn(A.kappa, __dadd_rn(namax, nb));
        s.thr = valid ? -pos_inf() : pos_inf();
        s.thrkey = 0;
    } else {
        double thr = __dmul_rn(0.5, __dmul_rn(nb, __dsub_rn(1.0, A.kappa)));
        s.thrkey = valid ? (__double2hiint(thr) - 1) : INT_MAX;
        s.thr = 0.0;
        s.slack = 0.0;
    }
}

template <int MODE>
__device__ __forceinline__ bool cs_flag(const CState<MODE> &s, double acc)
{
    if (MODE == SCAN_MIN) return acc >= s.thr;
    return __double2hiint(acc) >= s.thrkey;
}

template <int MODE>
__device__ __forceinline__ bool cs_done(const CState<MODE> &s)
{
    if (MODE == SCAN_FIND) return (!s.valid) || s.first >= 0;
    return !s.valid;
}

// ---------------------------------------------------------------------------------------
// register kernel: d <= 32 (DR = d rounded up to 4), TM candidates per thread
// ---------------------------------------------------------------------------------------
// The tile is read through a volatile pointer here on purpose: otherwise the compiler merges
// these loads with the filter loop's and keeps the whole tile column block alive (spilled)
// across the rarely taken branch.
template <int DR>
__device__ __forceinline__ double exact_dist_reg(const double (&a)[DR], const double *Tcol)
{
    const volatile double *Tv = Tcol;
    double D = 0.0;
#pragma unroll
    for (int k = 0; k < DR; k++) D = sq_step(D, Tv[k * REG_TILE_N], a[k]);
    return D;
}

// resident blocks per SM the register budget is planned for: candidates (2*TM*DR registers)
// + accumulators (2*TM*TN) + addressing/state
constexpr int reg_min_blocks(int DR, int TM)
{
    const int need = 2 * TM * DR + 2 * TM * TN + 48;
    return need <= 128 ? 4 : (need <= 168 ? 3 : 2);
}

template <int DR, int TM, int MODE>
__global__ void __launch_bounds__(SCAN_THREADS, reg_min_blocks(DR, TM)) k_scan_reg(const ScanArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    constexpr int TILE_DOUBLES = (DR + 1) * REG_TILE_N;
    constexpr uint32_t TILE_BYTES = TILE_DOUBLES * sizeof(double);
    double *tbuf = reinterpret_cast<double *>(smem_raw + 128);

    const int tid = threadIdx.x;
    const int round = blockIdx.y;
    const double *tiles = A.tiles + (long long)round * A.round_tile_stride;
    const int n_live = A.round_nlive ? A.round_nlive[round] : A.n_live;
    const int ntiles = (n_live + REG_TILE_N - 1) / REG_TILE_N;
    const long long n_items =
        A.round_nitems ? (long long)A.round_nitems[round]
                       : (A.n_items_dev ? (long long)*A.n_items_dev : A.n_items);
    const int *items = A.item_idx ? A.item_idx + (A.round_item_off ? A.round_item_off[round] : 0)
                                  : nullptr;
    const long long base = (long long)blockIdx.x * (SCAN_THREADS * TM);
    if (base >= n_items) return;
    const int d = A.d;

    // ---- candidates into registers
    double a[TM][DR];
    long long row[TM], orow[TM];
    CState<MODE> st[TM];
#pragma unroll
    for (int m = 0; m < TM; m++) {
        long long item = base + (long long)m * SCAN_THREADS + tid;
        bool valid = item < n_items;
        row[m] = valid ? (items ? (long long)items[item] : item) : -1;
        orow[m] = valid ? (A.out_row_idx ? (long long)A.out_row_idx[item] : row[m]) : -1;
        double nb = 0.0;
#pragma unroll
        for (int k = 0; k < DR; k++) {
            double v = (valid && k < d) ? A.cand[row[m] * d + k] : 0.0;
            a[m][k] = v;
            nb = fma(v, v, nb);
        }
        cs_init<MODE>(st[m], valid, nb, A);
        if (MODE == SCAN_SUBTRACT && valid)
            for (int k = 0; k < d; k++) A.out_rows[orow[m] * d + k] = 0.0;
    }

    // ---- TMA pipeline over live tiles
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
        tma_bulk_g2s(tbuf, tiles, TILE_BYTES, &bars[0]);
        if (ntiles > 1) {
            mbar_arrive_expect_tx(&bars[1], TILE_BYTES);
            tma_bulk_g2s(tbuf + TILE_DOUBLES, tiles + TILE_DOUBLES, TILE_BYTES, &bars[1]);
        }
    }

    bool warp_done = false;
    unsigned long long rechecks = 0;
    int inflight = -1;   // tile index still in flight when the loop is left early
    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        mbar_wait(&bars[buf], (t >> 1) & 1);
        const double *T = tbuf + buf * TILE_DOUBLES;
        if (!warp_done) {
#pragma unroll 1
            for (int g = 0; g < REG_TILE_N / TN; g++) {
                const double *Tg = T + g * TN;
                double acc[TM][TN];
                {
                    const double4 h = *reinterpret_cast<const double4 *>(Tg + DR * REG_TILE_N);
#pragma unroll
                    for (int m = 0; m < TM; m++) {
                        acc[m][0] = h.x; acc[m][1] = h.y; acc[m][2] = h.z; acc[m][3] = h.w;
                    }
                }
#pragma unroll
                for (int k = 0; k < DR; k++) {
                    const double4 b = *reinterpret_cast<const double4 *>(Tg + k * REG_TILE_N);
#pragma unroll
                    for (int m = 0; m < TM; m++) {
                        acc[m][0] = fma(a[m][k], b.x, acc[m][0]);
                        acc[m][1] = fma(a[m][k], b.y, acc[m][1]);
                        acc[m][2] = fma(a[m][k], b.z, acc[m][2]);
                        acc[m][3] = fma(a[m][k], b.w, acc[m][3]);
                    }
                }
                bool any = false;
#pragma unroll
                for (int m = 0; m < TM; m++)
#pragma unroll
                    for (int n = 0; n < TN; n++) any |= cs_flag<MODE>(st[m], acc[m][n]);
                if (any) {
                    // ---- decide: exact reference arithmetic for the pairs that passed the filter
                    const int gbase = t * REG_TILE_N + g * TN;
#pragma unroll
                    for (int m = 0; m < TM; m++) {
#pragma unroll
                        for (int n = 0; n < TN; n++) {
                            if (cs_flag<MODE>(st[m], acc[m][n])) {
                                rechecks++;
                                const double D = exact_dist_reg<DR>(a[m], Tg + n);
                                if (MODE == SCAN_MIN) {
                                    st[m].best = fmin(st[m].best, D);
                                    st[m].amax = fmax(st[m].amax, acc[m][n]);
                                    st[m].thr = __dsub_rn(st[m].amax, st[m].slack);
                                } else if (D <= A.r2) {
                                    if (MODE == SCAN_FIND) {
                                        st[m].first = gbase + n;
                                        st[m].thrkey = INT_MAX;   // later indices cannot win
                                    } else {
                                        st[m].cnt++;
                                        if (MODE == SCAN_SUBTRACT) {
                                            const volatile double *Tv = Tg + n;
                                            for (int k = 0; k < d; k++)
                                                A.out_rows[orow[m] * d + k] += Tv[k * REG_TILE_N];
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
            }
            bool done = true;
#pragma unroll
            for (int m = 0; m < TM; m++) done &= cs_done<MODE>(st[m]);
            warp_done = __all_sync(FULL, done);
        }
        const bool all_done = __syncthreads_and(warp_done);
        if (all_done) {
            if (t + 1 < ntiles) inflight = t + 1;
            break;
        }
        if (tid == 0 && t + 2 < ntiles) {
            mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
            tma_bulk_g2s(tbuf + buf * TILE_DOUBLES, tiles + (size_t)(t + 2) * TILE_DOUBLES,
                         TILE_BYTES, &bars[buf]);
        }
    }
    if (inflight >= 0) mbar_wait(&bars[inflight & 1], (inflight >> 1) & 1);

    // ---- results
    double blockmax = 0.0;
#pragma unroll
    for (int m = 0; m < TM; m++) {
        if (!st[m].valid) continue;
        if (MODE == SCAN_FIND) {
            if (A.out_idx) A.out_idx[orow[m]] = st[m].first;
            if (A.out_mask) A.out_mask[orow[m]] = st[m].first >= 0;
        } else if (MODE == SCAN_COUNT) {
            A.out_idx[orow[m]] = st[m].cnt;
        } else if (MODE == SCAN_SUBTRACT) {
            const double cnt = (double)st[m].cnt;
            for (int k = 0; k < d; k++) {
                double s = A.out_rows[orow[m] * d + k];
                A.out_rows[orow[m] * d + k] = __dsub_rn(A.cand[row[m] * d + k], __ddiv_rn(s, cnt));
            }
        } else {
            if (A.out_min) A.out_min[orow[m]] = st[m].best;
            blockmax = fmax(blockmax, st[m].best);
        }
    }
    if (MODE == SCAN_MIN) {
        for (int o = 16; o > 0; o >>= 1) blockmax = fmax(blockmax, __shfl_xor_sync(FULL, blockmax, o));
        if ((tid & 31) == 0 && A.out_round_max)
            atomicMax(A.out_round_max + round, f64_bits(blockmax));
    }
    if (A.stat_rechecks) {
        for (int o = 16; o > 0; o >>= 1) rechecks += __shfl_xor_sync(FULL, rechecks, o);
        if ((tid & 31) == 0 && rechecks) atomicAdd(A.stat_rechecks, rechecks);
    }
}

// ---------------------------------------------------------------------------------------
// membership ("any neighbour") kernel for MLFriends.inside: persistent blocks + slot refill
//
// inside() only needs to know WHETHER a live point lies within the radius, not which one comes
// first, so the scan order is free.  Each block streams the live tiles round-robin for as long as
// there is work; every thread owns TM candidate slots; a slot is retired at the first tile
// boundary after its candidate found a neighbour (or has seen all ntiles tiles) and is
// immediately refilled from a global work queue, so no lane idles while its warp-mates are
// still searching (the ordered kernel above pays the warp-maximum of the first-hit positions).
// Decisions are still made by the exact reference arithmetic, so the mask is bit-identical.
// ---------------------------------------------------------------------------------------
// filter pass of one shared-memory tile for the first TMA slots of every thread: records the
// pairs that pass the filter in pend[m] (bit g*4+n = live point of the tile)
template <int DR, int TM, int TMA>
__device__ __forceinline__ void tile_filter(const double (&a)[TM][DR], const int (&thrkey)[TM],
                                            unsigned long long (&pend)[TM], const double *T)
{
#pragma unroll 1
    for (int g = 0; g < REG_TILE_N / TN; g++) {
        const double *Tg = T + g * TN;
        double acc[TMA][TN];
        {
            const double4 h = *reinterpret_cast<const double4 *>(Tg + DR * REG_TILE_N);
#pragma unroll
            for (int m = 0; m < TMA; m++) {
                acc[m][0] = h.x; acc[m][1] = h.y; acc[m][2] = h.z; acc[m][3] = h.w;
            }
        }
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const double4 b = *reinterpret_cast<const double4 *>(Tg + k * REG_TILE_N);
#pragma unroll
            for (int m = 0; m < TMA; m++) {
                acc[m][0] = fma(a[m][k], b.x, acc[m][0]);
                acc[m][1] = fma(a[m][k], b.y, acc[m][1]);
                acc[m][2] = fma(a[m][k], b.z, acc[m][2]);
                acc[m][3] = fma(a[m][k], b.w, acc[m][3]);
            }
        }
#pragma unroll
        for (int m = 0; m < TMA; m++) {
            unsigned nib = 0;
#pragma unroll
            for (int n = 0; n < TN; n++)
                nib |= (__double2hiint(acc[m][n]) >= thrkey[m]) ? (1u << n) : 0u;
            pend[m] |= (unsigned long long)nib << (g * TN);
        }
    }
}

constexpr int ANY_STAGE_SLOTS = SCAN_THREADS;   // capacity of the compaction staging area
constexpr int COOP_G = 4;                       // survivors a warp advances together in the cooperative drain
constexpr int ANY_COOP_MAX = 24;                // survivors per block at which the drain turns cooperative

template <int DR, int TM>
__global__ void __launch_bounds__(SCAN_THREADS, reg_min_blocks(DR, TM))
k_inside_any(const ScanArgs A, int *__restrict__ queue_head)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    int *s_info = reinterpret_cast<int *>(smem_raw + 64);           // per warp: live slots | exhausted<<16
    constexpr int TILE_DOUBLES = (DR + 1) * REG_TILE_N;
    constexpr uint32_t TILE_BYTES = TILE_DOUBLES * sizeof(double);
    double *tbuf = reinterpret_cast<double *>(smem_raw + 128);
    double *stage = tbuf + 2 * TILE_DOUBLES;                        // [DR+2][ANY_STAGE_SLOTS]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const double *tiles = A.tiles;
    const int ntiles = (A.n_live + REG_TILE_N - 1) / REG_TILE_N;
    const int n_items = A.n_items_dev ? *A.n_items_dev : (int)A.n_items;
    const int d = A.d;
    const double thr_scale = __dmul_rn(0.5, __dsub_rn(1.0, A.kappa));

    double a[TM][DR];
    int row[TM], orow[TM], rem[TM], thrkey[TM], hit[TM];
    unsigned long long pend[TM];   // per slot: bit (g*4+n) = live point of the tile to decide
#pragma unroll
    for (int m = 0; m < TM; m++) {
        row[m] = -1; orow[m] = -1; rem[m] = 0; thrkey[m] = INT_MAX; hit[m] = 0; pend[m] = 0ull;
#pragma unroll
        for (int k = 0; k < DR; k++) a[m][k] = 0.0;
    }
    bool exhausted = false;

    // retire finished slots and refill every free slot from the queue (warp-synchronous)
    auto refill = [&]() {
#pragma unroll
        for (int m = 0; m < TM; m++) {
            if (row[m] >= 0 && (hit[m] || rem[m] <= 0)) {
                A.out_mask[orow[m]] = hit[m] ? 1 : 0;
                if (A.out_like && !hit[m]) A.out_like[orow[m]] = -pos_inf();
                row[m] = -1;
                thrkey[m] = INT_MAX;
            }
            const bool need = (row[m] < 0) && !exhausted;
            const unsigned ball = __ballot_sync(FULL, need);
            if (ball) {
                int base = 0;
                if (lane == 0) base = atomicAdd(queue_head, __popc(ball));
                base = __shfl_sync(FULL, base, 0);
                if (need) {
                    const int item = base + __popc(ball & ((1u << lane) - 1));
                    if (item < n_items) {
                        const int r = A.item_idx ? A.item_idx[item] : item;
                        row[m] = r;
                        orow[m] = A.out_row_idx ? A.out_row_idx[item] : r;
                        double nb = 0.0;
#pragma unroll
                        for (int k = 0; k < DR; k++) {
                            const double v = (k < d) ? A.cand[(size_t)r * d + k] : 0.0;
                            a[m][k] = v;
                            nb = fma(v, v, nb);
                        }
                        thrkey[m] = __double2hiint(__dmul_rn(nb, thr_scale)) - 1;
                        rem[m] = ntiles;
                        hit[m] = 0;
                    } else {
                        exhausted = true;
                    }
                }
            }
        }
        // one lane learning that the queue is empty is enough for the whole warp
        exhausted = __any_sync(FULL, exhausted);
    };

    refill();
    {
        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        if (__syncthreads_and(idle)) return;   // nothing claimed by this block
    }

    // The circular scan may start anywhere.  Start where the block's first proposal sits in the
    // batch, scaled to the live block: when the proposals ARE the live points in order
    // (`region.inside(active_u)`, integrator.py:1855) every proposal meets itself within a tile
    // or two instead of after up to ntiles; for unrelated proposals it is just some offset.
    if (tid == 0) s_info[8] = orow[0] >= 0 ? orow[0] : 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned start =
        (unsigned)(((long long)s_info[8] * ntiles) / (A.n_items > 0 ? A.n_items : 1)) % (unsigned)ntiles;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
        tma_bulk_g2s(tbuf, tiles + (size_t)(start % ntiles) * TILE_DOUBLES, TILE_BYTES, &bars[0]);
        mbar_arrive_expect_tx(&bars[1], TILE_BYTES);
        tma_bulk_g2s(tbuf + TILE_DOUBLES, tiles + (size_t)((start + 1) % ntiles) * TILE_DOUBLES,
                     TILE_BYTES, &bars[1]);
    }

    unsigned long long rechecks = 0;
    unsigned int tile_units = 0;
    bool warp_idle = false;   // a warp with no live slot skips the tile (drain phase)
    bool single = (TM == 1);  // block-uniform: every live slot sits in m = 0 (after a compaction)
    int cap = SCAN_THREADS * TM;   // slots the block is spread over
    for (unsigned tt = 0;; tt++) {
        const int buf = tt & 1;
        mbar_wait(&bars[buf], (tt >> 1) & 1);
        const double *T = tbuf + buf * TILE_DOUBLES;
        if (!warp_idle) {
            if (single) {
                tile_units += 1;
                tile_filter<DR, TM, 1>(a, thrkey, pend, T);
            } else {
                tile_units += TM;
                tile_filter<DR, TM, TM>(a, thrkey, pend, T);
            }
            // ---- decide: exact reference arithmetic for the recorded pairs, all lanes at once
            // (any order is fine for a membership test)
#pragma unroll
            for (int m = 0; m < TM; m++) {
                while (__any_sync(FULL, pend[m] != 0ull)) {
                    if (pend[m] != 0ull) {
                        const int col = __ffsll((long long)pend[m]) - 1;
                        pend[m] &= pend[m] - 1ull;
                        rechecks++;
                        const double D = exact_dist_reg<DR>(a[m], T + col);
                        unc_note(A, D);
                        if (D <= A.r2) {
                            hit[m] = 1;
                            thrkey[m] = INT_MAX;
                            pend[m] = 0ull;
                        }
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < TM; m++) rem[m]--;
            refill();
        }
        // ---- block bookkeeping: live slots per warp, queue state
        int mine = 0;
#pragma unroll
        for (int m = 0; m < TM; m++) mine += row[m] >= 0;
        int incl = mine;   // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int *info = s_info + (tt & 1) * (SCAN_THREADS / 32);   // double-buffered across iterations
        if (lane == 31) info[warp] = incl | ((exhausted ? 1 : 0) << 16);
        __syncthreads();   // also: every thread is done with tile `buf`
        int total = 0, warp_off = 0;
        bool all_exh = true;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; w++) {
            const int v = info[w];
            if (w < warp) warp_off += v & 0xffff;
            total += v & 0xffff;
            all_exh &= (v >> 16) != 0;
        }
        if (total == 0) {
            mbar_wait(&bars[(tt + 1) & 1], ((tt + 1) >> 1) & 1);   // tile tt+1 is always in flight
            break;
        }
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
            tma_bulk_g2s(tbuf + buf * TILE_DOUBLES,
                         tiles + (size_t)((start + tt + 2) % (unsigned)ntiles) * TILE_DOUBLES,
                         TILE_BYTES, &bars[buf]);
        }
        // ---- drain: once the queue is empty, pack the surviving slots into as few warps as
        // possible (slot m = 0 of threads 0..total-1), so the tail costs lanes, not warps
        if (all_exh && total <= cap / 2 && total <= ANY_STAGE_SLOTS) {
            int j = warp_off + incl - mine;
#pragma unroll
            for (int m = 0; m < TM; m++) {
                if (row[m] >= 0) {
#pragma unroll
                    for (int k = 0; k < DR; k++) stage[k * ANY_STAGE_SLOTS + j] = a[m][k];
                    stage[DR * ANY_STAGE_SLOTS + j] =
                        __hiloint2double(row[m], orow[m]);
                    stage[(DR + 1) * ANY_STAGE_SLOTS + j] =
                        __hiloint2double(rem[m], thrkey[m]);
                    j++;
                }
                row[m] = -1; thrkey[m] = INT_MAX; hit[m] = 0; pend[m] = 0ull;
            }
            __syncthreads();
            if (tid < total) {
#pragma unroll
                for (int k = 0; k < DR; k++) a[0][k] = stage[k * ANY_STAGE_SLOTS + tid];
                const double p0 = stage[DR * ANY_STAGE_SLOTS + tid];
                const double p1 = stage[(DR + 1) * ANY_STAGE_SLOTS + tid];
                row[0] = __double2hiint(p0); orow[0] = __double2loint(p0);
                rem[0] = __double2hiint(p1); thrkey[0] = __double2loint(p1);
            }
            __syncthreads();   // the staging area may be reused by the next compaction
            single = true;
            cap = (total + 31) / 32 * 32;
        }
        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        warp_idle = __all_sync(FULL, idle);
    }
    if (A.stat_rechecks) {
        for (int o = 16; o > 0; o >>= 1) rechecks += __shfl_xor_sync(FULL, rechecks, o);
        if (lane == 0 && rechecks) atomicAdd(A.stat_rechecks, rechecks);
    }
    if (A.stat_tiles && lane == 0) atomicAdd(A.stat_tiles, (unsigned long long)tile_units);
}

// ---------------------------------------------------------------------------------------
// membership kernel with an fp32 PRE-filter (k_inside_any32)
//
// Same persistent / refill / compaction structure as k_inside_any, but the filter runs in single
// precision: the FFMA pipe issues twice as fast as the DFMA pipe and the candidate registers
// halve.  Pairs are flagged when  acc32 = h32_i + sum_k fl32(a_ik) fl32(b_jk)  (FFMA chain)
// reaches thr32_j.  With u32 = 2^-24 every error source (conversion of both operands, the chain,
// rounding of h and thr) is bounded by (d+4) u32 (|a|^2+|b|^2+r2); h and thr are widened by
// kappa32 = (4d+32) u32 of the same quantity (h rounded up, thr rounded down), so every
// reference hit is flagged.  Flagged pairs are DECIDED in the reference's exact fp64 sequence
// from the fp64 rows in global memory (L2 resident), so the mask is still bit-identical.
// The host only selects this kernel when the slack is thin compared with r2 and all magnitudes
// are far from the fp32 range limits; a candidate whose own norm is out of range flags everything.
// ---------------------------------------------------------------------------------------
constexpr int ANY32_MAX_DR = 128;   // candidates of up to 128 dims fit the registers in fp32 (TM = 1)

constexpr int any32_min_blocks(int DR, int TM)
{
    const int need = TM * DR + 4 * TM + 48;
    // (+ ~48 transient registers of the batched fp64 decide loads)
    return need <= 80 ? 5 : (need <= 128 ? 4 : (need <= 168 ? 3 : 2));
}

// Per slot and group of 4 live points only the MAXIMUM of the four accumulators is compared in
// the common path (2-3 FMNMX + one FSETP instead of four compare/select/merge chains):
//   max <  thr_lo            nothing flagged (the usual case);
//   max >= thr_hi            a CERTAIN neighbour: with M = kappa32 (|a|^2_max + |b|^2 + r2) the
//                            error budget of the filter (widening of h and thr, conversions,
//                            chain: < kappa32 Q in total) gives  acc - thr >= M  =>  r2 - D > 0
//                            by more than the fp64 rounding of the reference's own distance, so
//                            the reference finds D <= r2 too and the slot retires without an
//                            exact evaluation (membership only needs existence);
//   otherwise                the flagged pairs of the group are recorded and decided exactly.
// thr_lo = -inf (candidate out of the fp32 range) flags everything, thr_hi = +inf then.
template <int DR, int TM, int TMA>
__device__ __forceinline__ void tile_filter32(const float (&a)[TM][DR], float (&thr_lo)[TM],
                                              const float (&thr_hi)[TM], const int (&row)[TM],
                                              int (&hit)[TM], unsigned long long (&pend)[TM],
                                              const float *T)
{
#pragma unroll 1
    for (int g = 0; g < REG_TILE_N / TN; g++) {
        const float *Tg = T + g * TN;
        // two packed FMAs (FFMA2) per slot and k: live points (0,1) and (2,3) of the group ride
        // in 64-bit register pairs, the proposal coordinate is the instruction's scalar operand.
        // Each half is one fmaf(): the filter's arithmetic (and its error budget) is unchanged.
        float2 acc01[TMA], acc23[TMA];
        {
            const float4 h = *reinterpret_cast<const float4 *>(Tg + DR * REG_TILE_N);
#pragma unroll
            for (int m = 0; m < TMA; m++) {
                acc01[m] = make_float2(h.x, h.y);
                acc23[m] = make_float2(h.z, h.w);
            }
        }
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const float4 b = *reinterpret_cast<const float4 *>(Tg + k * REG_TILE_N);
#pragma unroll
            for (int m = 0; m < TMA; m++) {
                const float2 am = make_float2(a[m][k], a[m][k]);
                acc01[m] = ffma2(am, make_float2(b.x, b.y), acc01[m]);
                acc23[m] = ffma2(am, make_float2(b.z, b.w), acc23[m]);
            }
        }
        float acc[TMA][TN];
#pragma unroll
        for (int m = 0; m < TMA; m++) {
            acc[m][0] = acc01[m].x; acc[m][1] = acc01[m].y;
            acc[m][2] = acc23[m].x; acc[m][3] = acc23[m].y;
        }
#pragma unroll
        for (int m = 0; m < TMA; m++) {
            const float mx = fmaxf(fmaxf(acc[m][0], acc[m][1]), fmaxf(acc[m][2], acc[m][3]));
            if (!(mx < thr_lo[m])) {            // also taken for NaN
                if (row[m] >= 0 && !hit[m]) {
                    if (mx >= thr_hi[m]) {
                        hit[m] = 1;
                        pend[m] = 0ull;
                        thr_lo[m] = __int_as_float(0x7f800000);
                    } else {
                        unsigned nib = 0;
#pragma unroll
                        for (int n = 0; n < TN; n++) nib |= !(acc[m][n] < thr_lo[m]) ? (1u << n) : 0u;
                        pend[m] |= (unsigned long long)nib << (g * TN);
                    }
                }
            }
        }
    }
}

template <int DR, int TM>
__global__ void __launch_bounds__(SCAN_THREADS, any32_min_blocks(DR, TM))
k_inside_any32(const ScanArgs A, int *__restrict__ queue_head)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    int *s_info = reinterpret_cast<int *>(smem_raw + 64);
    constexpr int TILE_FLOATS = (DR + 1) * REG_TILE_N;
    constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
    float *tbuf = reinterpret_cast<float *>(smem_raw + 128);
    uint32_t *stage = reinterpret_cast<uint32_t *>(tbuf + 2 * TILE_FLOATS);   // [DR+5][slots]

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const float *tiles = A.tiles32;
    const int ntiles = (A.n_live + REG_TILE_N - 1) / REG_TILE_N;
    const int n_items = A.n_items_dev ? *A.n_items_dev : (int)A.n_items;
    const int d = A.d;
    const double thr_scale = __dmul_rn(0.5, __dsub_rn(1.0, A.kappa32));
    // M = 1.001 kappa32 (|a|^2_max + r2 + |b|^2); namax32 = +inf switches the shortcut off
    const double sure_scale = __dmul_rn(1.001, A.kappa32), sure_base = __dadd_rn(A.namax32, A.r2);

    float a[TM][DR];
    int row[TM], orow[TM], rem[TM], hit[TM];
    float thr_lo[TM], thr_hi[TM];
    unsigned long long pend[TM];
#pragma unroll
    for (int m = 0; m < TM; m++) {
        row[m] = -1; orow[m] = -1; rem[m] = 0; hit[m] = 0; pend[m] = 0ull;
        thr_lo[m] = __int_as_float(0x7f800000); thr_hi[m] = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < DR; k++) a[m][k] = 0.f;
    }
    bool exhausted = false;

    // one work item into slot m of this thread
    auto load_item = [&](int m, int item) {
        const int r = A.item_idx ? A.item_idx[item] : item;
        row[m] = r;
        orow[m] = A.out_row_idx ? A.out_row_idx[item] : r;
        double nb = 0.0;
#pragma unroll
        for (int k = 0; k < DR; k++) {
            const double v = (k < d) ? A.cand[(size_t)r * d + k] : 0.0;
            a[m][k] = __double2float_rn(v);
            nb = fma(v, v, nb);
        }
        // threshold rounded DOWN; a zero / out-of-range norm flags everything.
        // thr_hi = thr + M rounded UP: the certain-neighbour level (tile_filter32)
        const float thr = __double2float_rd(__dmul_rn(nb, thr_scale));
        const bool in_range = nb < 1e30 && thr > 0.f;
        thr_lo[m] = in_range ? thr : -__int_as_float(0x7f800000);
        thr_hi[m] = in_range ? __double2float_ru(__dadd_rn(
                                   (double)thr, __dmul_rn(sure_scale, __dadd_rn(sure_base, nb))))
                             : __int_as_float(0x7f800000);
        rem[m] = ntiles;
        hit[m] = 0;
    };

    // binned launches (unb_cluster.cu): the items are sorted by the live tile whose centroid is
    // nearest, and a warp about to stream tile `cur_bin` takes its new items from that bin (the
    // following bins when it is empty) -- most proposals then meet a neighbour in their first tile
    const bool binned = A.bin_order != nullptr;
    int cur_bin = 0;

    auto refill = [&]() {
#pragma unroll
        for (int m = 0; m < TM; m++) {
            if (row[m] >= 0 && (hit[m] || rem[m] <= 0)) {
                A.out_mask[orow[m]] = hit[m] ? 1 : 0;
                if (A.out_like && !hit[m]) A.out_like[orow[m]] = -pos_inf();
                row[m] = -1;
                thr_lo[m] = __int_as_float(0x7f800000);
            }
            bool need = (row[m] < 0) && !exhausted;
            unsigned ball = __ballot_sync(FULL, need);
            if (!binned) {
                if (ball) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(queue_head, __popc(ball));
                    base = __shfl_sync(FULL, base, 0);
                    if (need) {
                        const int item = base + __popc(ball & ((1u << lane) - 1));
                        if (item < n_items) load_item(m, item);
                        else exhausted = true;
                    }
                }
            } else {
                while (ball) {   // warp-uniform
                    int base = -1, avail = 0, b = cur_bin;
                    if (lane == 0) {
                        const int want = __popc(ball);
                        for (int tries = 0; tries < ntiles; tries++) {
                            const int size_b = A.bin_start[b + 1] - A.bin_start[b];
                            if (*(volatile int *)(A.bin_head + b) < size_b) {
                                const int got = atomicAdd(A.bin_head + b, want);
                                if (got < size_b) {
                                    base = got;
                                    avail = min(want, size_b - got);
                                    break;
                                }
                            }
                            b = (b + 1 == ntiles) ? 0 : b + 1;
                        }
                    }
                    base = __shfl_sync(FULL, base, 0);
                    avail = __shfl_sync(FULL, avail, 0);
                    b = __shfl_sync(FULL, b, 0);
                    if (base < 0) {   // every bin is empty
                        exhausted = true;
                        break;
                    }
                    if (need && __popc(ball & ((1u << lane) - 1)) < avail) {
                        load_item(m, A.bin_order[A.bin_start[b] + base + __popc(ball & ((1u << lane) - 1))]);
                        need = false;
                    }
                    ball = __ballot_sync(FULL, need);
                }
            }
        }
        exhausted = __any_sync(FULL, exhausted);
    };

    if (binned) cur_bin = (int)(blockIdx.x % (unsigned)ntiles);   // blocks spread over the tiles
    refill();
    {
        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        if (__syncthreads_and(idle)) return;
    }

    // block-dependent start tile (see k_inside_any); binned launches start at their first bin
    if (tid == 0) s_info[8] = orow[0] >= 0 ? orow[0] : 0;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const unsigned start = binned
        ? blockIdx.x % (unsigned)ntiles
        : (unsigned)(((long long)s_info[8] * ntiles) / (A.n_items > 0 ? A.n_items : 1)) % (unsigned)ntiles;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
        tma_bulk_g2s(tbuf, tiles + (size_t)(start % ntiles) * TILE_FLOATS, TILE_BYTES, &bars[0]);
        mbar_arrive_expect_tx(&bars[1], TILE_BYTES);
        tma_bulk_g2s(tbuf + TILE_FLOATS, tiles + (size_t)((start + 1) % ntiles) * TILE_FLOATS,
                     TILE_BYTES, &bars[1]);
    }

    unsigned long long rechecks = 0;
    unsigned int tile_units = 0;
    bool warp_idle = false;
    bool single = (TM == 1);
    int cap = SCAN_THREADS * TM;
    for (unsigned tt = 0;; tt++) {
        const int buf = tt & 1;
        mbar_wait(&bars[buf], (tt >> 1) & 1);
        const float *T = tbuf + buf * TILE_FLOATS;
        if (!warp_idle) {
            if (single) {
                tile_units += 1;
                tile_filter32<DR, TM, 1>(a, thr_lo, thr_hi, row, hit, pend, T);
            } else {
                tile_units += TM;
                tile_filter32<DR, TM, TM>(a, thr_lo, thr_hi, row, hit, pend, T);
            }
            // ---- decide in exact fp64 from the fp64 rows (global memory, L2 resident)
            const int tile_first = (int)((start + tt) % (unsigned)ntiles) * REG_TILE_N;
#pragma unroll
            for (int m = 0; m < TM; m++) {
                while (__any_sync(FULL, pend[m] != 0ull)) {
                    if (pend[m] != 0ull) {
                        const int col = __ffsll((long long)pend[m]) - 1;
                        pend[m] &= pend[m] - 1ull;
                        rechecks++;
                        // padded slots of the last tile are only ever flagged by a proposal that
                        // flags everything (out-of-range norm): they are no live points
                        if (tile_first + col >= A.n_live) continue;
                        const int lrow = A.perm32 ? A.perm32[tile_first + col] : tile_first + col;
                        const double *lp = A.live_rows + (size_t)lrow * d;
                        const double *cp = A.cand + (size_t)row[m] * d;
                        // loads are issued in batches of 2 x 12 so that one L2 round trip covers
                        // a batch (the rows are L2 resident); padded terms add +0
                        double D = 0.0;
#pragma unroll
                        for (int k0 = 0; k0 < DR; k0 += 12) {
                            double lv[12], cv[12];
#pragma unroll
                            for (int j = 0; j < 12; j++) {
                                const bool in = (k0 + j) < d;
                                lv[j] = in ? __ldg(lp + k0 + j) : 0.0;
                                cv[j] = in ? __ldg(cp + k0 + j) : 0.0;
                            }
#pragma unroll
                            for (int j = 0; j < 12; j++) D = sq_step(D, lv[j], cv[j]);
                        }
                        unc_note(A, D);
                        if (D <= A.r2) {
                            hit[m] = 1;
                            thr_lo[m] = __int_as_float(0x7f800000);
                            pend[m] = 0ull;
                        }
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < TM; m++) rem[m]--;
            cur_bin = (int)((start + tt + 1) % (unsigned)ntiles);   // the tile this warp streams next
            refill();
        }
        int mine = 0;
#pragma unroll
        for (int m = 0; m < TM; m++) mine += row[m] >= 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        int *info = s_info + (tt & 1) * (SCAN_THREADS / 32);
        if (lane == 31) info[warp] = incl | ((exhausted ? 1 : 0) << 16);
        __syncthreads();
        int total = 0, warp_off = 0;
        bool all_exh = true;
#pragma unroll
        for (int w = 0; w < SCAN_THREADS / 32; w++) {
            const int v = info[w];
            if (w < warp) warp_off += v & 0xffff;
            total += v & 0xffff;
            all_exh &= (v >> 16) != 0;
        }
        if (total == 0) {
            mbar_wait(&bars[(tt + 1) & 1], ((tt + 1) >> 1) & 1);
            break;
        }
        if (tid == 0) {
            mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
            tma_bulk_g2s(tbuf + buf * TILE_FLOATS,
                         tiles + (size_t)((start + tt + 2) % (unsigned)ntiles) * TILE_FLOATS,
                         TILE_BYTES, &bars[buf]);
        }
        if (all_exh && total <= A.coop_max) {
            // ---- cooperative drain.  The last few proposals of a block are the unlucky long
            // scans; one proposal per lane would stream the remaining tiles at the price of a
            // full warp each.  Instead the survivors move to shared memory and every warp takes
            // one survivor at a time with its LANES spread over the tile's live points (lane l:
            // points 2l, 2l+1), so a tile costs ~4 DR instructions per survivor instead of
            // ~80 DR per warp, and the tail of the launch shrinks accordingly.
            int j = warp_off + incl - mine;
#pragma unroll
            for (int m = 0; m < TM; m++) {
                if (row[m] >= 0) {
#pragma unroll
                    for (int k = 0; k < DR; k++)
                        stage[k * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(a[m][k]);
                    stage[(DR + 0) * ANY_STAGE_SLOTS + j] = (uint32_t)row[m];
                    stage[(DR + 1) * ANY_STAGE_SLOTS + j] = (uint32_t)orow[m];
                    stage[(DR + 2) * ANY_STAGE_SLOTS + j] = (uint32_t)rem[m];
                    stage[(DR + 3) * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(thr_lo[m]);
                    stage[(DR + 4) * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(thr_hi[m]);
                    j++;
                }
                row[m] = -1;
            }
            volatile int *alive = s_info + 15;
            if (tid == 0) *alive = total;
            __syncthreads();
            unsigned int coop_units = 0;
            for (unsigned t2 = tt + 1;; t2++) {
                const int b2 = t2 & 1;
                mbar_wait(&bars[b2], (t2 >> 1) & 1);
                const float *T2 = tbuf + b2 * TILE_FLOATS;
                const int first2 = (int)((start + t2) % (unsigned)ntiles) * REG_TILE_N;
                // COOP_G survivors per warp pass: independent FFMA chains hide each other's latency
                for (int j0 = warp; j0 < total; j0 += (SCAN_THREADS / 32) * COOP_G) {
                    int jg[COOP_G], rg[COOP_G];
                    bool any_act = false;
#pragma unroll
                    for (int g = 0; g < COOP_G; g++) {
                        const int jj = j0 + g * (SCAN_THREADS / 32);
                        jg[g] = jj < total ? jj : total - 1;
                        rg[g] = jj < total ? (int)stage[(DR + 0) * ANY_STAGE_SLOTS + jj] : -1;
                        any_act |= rg[g] >= 0;   // finished earlier: warp-uniform
                    }
                    if (!any_act) continue;
                    const float2 h = *reinterpret_cast<const float2 *>(T2 + DR * REG_TILE_N + 2 * lane);
                    float2 acc2[COOP_G];
#pragma unroll
                    for (int g = 0; g < COOP_G; g++) acc2[g] = h;
#pragma unroll
                    for (int k = 0; k < DR; k++) {
                        const float2 bb = *reinterpret_cast<const float2 *>(T2 + k * REG_TILE_N + 2 * lane);
#pragma unroll
                        for (int g = 0; g < COOP_G; g++) {
                            const float c = __int_as_float((int)stage[k * ANY_STAGE_SLOTS + jg[g]]);
                            acc2[g] = ffma2(make_float2(c, c), bb, acc2[g]);
                        }
                    }
                    float acc0[COOP_G], acc1[COOP_G];
#pragma unroll
                    for (int g = 0; g < COOP_G; g++) { acc0[g] = acc2[g].x; acc1[g] = acc2[g].y; }
#pragma unroll
                    for (int g = 0; g < COOP_G; g++) {
                        if (rg[g] < 0) continue;
                        const int j2 = jg[g], r2_ = rg[g];
                        coop_units++;
                        const float tl = __int_as_float((int)stage[(DR + 3) * ANY_STAGE_SLOTS + j2]);
                        const float th = __int_as_float((int)stage[(DR + 4) * ANY_STAGE_SLOTS + j2]);
                        const bool f0 = !(acc0[g] < tl), f1 = !(acc1[g] < tl);   // also true for NaN
                        bool found = __any_sync(FULL, (acc0[g] >= th) || (acc1[g] >= th));
                        if (!found && __any_sync(FULL, f0 || f1)) {
                            // uncertain shell: every lane decides its own flagged pairs exactly
                            bool ok = false;
                            const double *cp = A.cand + (size_t)r2_ * d;
#pragma unroll 1
                            for (int which = 0; which < 2; which++) {
                                if ((which == 0 ? f0 : f1) && !ok && first2 + 2 * lane + which < A.n_live) {
                                    const int lslot = first2 + 2 * lane + which;
                                    const double *lp = A.live_rows + (size_t)(A.perm32 ? A.perm32[lslot] : lslot) * d;
                                    double D = 0.0;
                                    for (int kk = 0; kk < d; kk++) D = sq_step(D, __ldg(lp + kk), __ldg(cp + kk));
                                    unc_note(A, D);
                                    ok = D <= A.r2;
                                    rechecks++;
                                }
                            }
                            found = __any_sync(FULL, ok);
                        }
                        const int rem2 = (int)stage[(DR + 2) * ANY_STAGE_SLOTS + j2] - 1;
                        __syncwarp();
                        if (lane == 0) {
                            if (found || rem2 <= 0) {
                                const int o2 = (int)stage[(DR + 1) * ANY_STAGE_SLOTS + j2];
                                A.out_mask[o2] = found ? 1 : 0;
                                if (A.out_like && !found) A.out_like[o2] = -pos_inf();
                                stage[(DR + 0) * ANY_STAGE_SLOTS + j2] = 0xffffffffu;
                                atomicSub((int *)alive, 1);
                            } else {
                                stage[(DR + 2) * ANY_STAGE_SLOTS + j2] = (uint32_t)rem2;
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                const int left = *alive;
                __syncthreads();
                if (left == 0) {
                    mbar_wait(&bars[(t2 + 1) & 1], ((t2 + 1) >> 1) & 1);   // tile t2+1 is in flight
                    break;
                }
                if (tid == 0) {
                    mbar_arrive_expect_tx(&bars[b2], TILE_BYTES);
                    tma_bulk_g2s(tbuf + b2 * TILE_FLOATS,
                                 tiles + (size_t)((start + t2 + 2) % (unsigned)ntiles) * TILE_FLOATS,
                                 TILE_BYTES, &bars[b2]);
                }
            }
            tile_units += (coop_units + 31) / 32;   // one survivor-tile = 1/32 of a warp-tile of filter work
            break;
        }
        if (all_exh && total <= cap / 2 && total <= ANY_STAGE_SLOTS) {
            int j = warp_off + incl - mine;
#pragma unroll
            for (int m = 0; m < TM; m++) {
                if (row[m] >= 0) {
#pragma unroll
                    for (int k = 0; k < DR; k++)
                        stage[k * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(a[m][k]);
                    stage[(DR + 0) * ANY_STAGE_SLOTS + j] = (uint32_t)row[m];
                    stage[(DR + 1) * ANY_STAGE_SLOTS + j] = (uint32_t)orow[m];
                    stage[(DR + 2) * ANY_STAGE_SLOTS + j] = (uint32_t)rem[m];
                    stage[(DR + 3) * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(thr_lo[m]);
                    stage[(DR + 4) * ANY_STAGE_SLOTS + j] = (uint32_t)__float_as_int(thr_hi[m]);
                    j++;
                }
                row[m] = -1; thr_lo[m] = __int_as_float(0x7f800000); hit[m] = 0; pend[m] = 0ull;
            }
            __syncthreads();
            if (tid < total) {
#pragma unroll
                for (int k = 0; k < DR; k++)
                    a[0][k] = __int_as_float((int)stage[k * ANY_STAGE_SLOTS + tid]);
                row[0] = (int)stage[(DR + 0) * ANY_STAGE_SLOTS + tid];
                orow[0] = (int)stage[(DR + 1) * ANY_STAGE_SLOTS + tid];
                rem[0] = (int)stage[(DR + 2) * ANY_STAGE_SLOTS + tid];
                thr_lo[0] = __int_as_float((int)stage[(DR + 3) * ANY_STAGE_SLOTS + tid]);
                thr_hi[0] = __int_as_float((int)stage[(DR + 4) * ANY_STAGE_SLOTS + tid]);
            }
            __syncthreads();
            single = true;
            cap = (total + 31) / 32 * 32;
        }
        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        warp_idle = __all_sync(FULL, idle);
    }
    if (A.stat_rechecks) {
        for (int o = 16; o > 0; o >>= 1) rechecks += __shfl_xor_sync(FULL, rechecks, o);
        if (lane == 0 && rechecks) atomicAdd(A.stat_rechecks, rechecks);
    }
    if (A.stat_tiles && lane == 0) atomicAdd(A.stat_tiles, (unsigned long long)tile_units);
}

// ---------------------------------------------------------------------------------------
// k_inside_any32w: the same membership scan with WARP-INDEPENDENT streams.  Every warp owns two
// fp32 tile buffers and two mbarriers, issues its own TMA tile loads (an fp32 tile is 5 KB, so
// the 4x L2->shared traffic is cheap) and retires / refills / drains on its own: there is no
// block barrier and no block bookkeeping at all, so a warp that is deciding or refilling never
// holds up its neighbours.  Drain: once the queue is empty a warp moves the survivors of slot 1
// into free slot-0 lanes with shuffles and continues with the single-slot filter.
// ---------------------------------------------------------------------------------------
template <int DR, int TM>
__global__ void __launch_bounds__(SCAN_THREADS, any32_min_blocks(DR, TM))
k_inside_any32w(const ScanArgs A, int *__restrict__ queue_head)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int NW = SCAN_THREADS / 32;
    constexpr int TILE_FLOATS = (DR + 1) * REG_TILE_N;
    constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw) + warp * 2;          // 2 per warp
    float *tbuf = reinterpret_cast<float *>(smem_raw + 128) + (size_t)warp * 2 * TILE_FLOATS;
    static_assert(NW * 2 * sizeof(uint64_t) <= 128, "mbarrier header");

    const float *tiles = A.tiles32;
    const int ntiles = (A.n_live + REG_TILE_N - 1) / REG_TILE_N;
    const int n_items = A.n_items_dev ? *A.n_items_dev : (int)A.n_items;
    const int d = A.d;
    const double thr_scale = __dmul_rn(0.5, __dsub_rn(1.0, A.kappa32));
    // M = 1.001 kappa32 (|a|^2_max + r2 + |b|^2); namax32 = +inf switches the shortcut off
    const double sure_scale = __dmul_rn(1.001, A.kappa32), sure_base = __dadd_rn(A.namax32, A.r2);

    float a[TM][DR];
    int row[TM], orow[TM], rem[TM], hit[TM];
    float thr_lo[TM], thr_hi[TM];
    unsigned long long pend[TM];
#pragma unroll
    for (int m = 0; m < TM; m++) {
        row[m] = -1; orow[m] = -1; rem[m] = 0; hit[m] = 0; pend[m] = 0ull;
        thr_lo[m] = __int_as_float(0x7f800000); thr_hi[m] = __int_as_float(0x7f800000);
#pragma unroll
        for (int k = 0; k < DR; k++) a[m][k] = 0.f;
    }
    bool exhausted = false;

    auto refill = [&]() {
#pragma unroll
        for (int m = 0; m < TM; m++) {
            if (row[m] >= 0 && (hit[m] || rem[m] <= 0)) {
                A.out_mask[orow[m]] = hit[m] ? 1 : 0;
                if (A.out_like && !hit[m]) A.out_like[orow[m]] = -pos_inf();
                row[m] = -1;
                thr_lo[m] = __int_as_float(0x7f800000);
            }
            const bool need = (row[m] < 0) && !exhausted;
            const unsigned ball = __ballot_sync(FULL, need);
            if (ball) {
                int base = 0;
                if (lane == 0) base = atomicAdd(queue_head, __popc(ball));
                base = __shfl_sync(FULL, base, 0);
                if (need) {
                    const int item = base + __popc(ball & ((1u << lane) - 1));
                    if (item < n_items) {
                        const int r = A.item_idx ? A.item_idx[item] : item;
                        row[m] = r;
                        orow[m] = A.out_row_idx ? A.out_row_idx[item] : r;
                        double nb = 0.0;
#pragma unroll
                        for (int k = 0; k < DR; k++) {
                            const double v = (k < d) ? A.cand[(size_t)r * d + k] : 0.0;
                            a[m][k] = __double2float_rn(v);
                            nb = fma(v, v, nb);
                        }
                        // threshold rounded DOWN; a zero / out-of-range norm flags everything.
                        // thr_hi = thr + M rounded UP: the certain-neighbour level (tile_filter32)
                        const float thr = __double2float_rd(__dmul_rn(nb, thr_scale));
                        const bool in_range = nb < 1e30 && thr > 0.f;
                        thr_lo[m] = in_range ? thr : -__int_as_float(0x7f800000);
                        thr_hi[m] = in_range ? __double2float_ru(__dadd_rn(
                                                   (double)thr, __dmul_rn(sure_scale, __dadd_rn(sure_base, nb))))
                                             : __int_as_float(0x7f800000);
                        rem[m] = ntiles;
                        hit[m] = 0;
                    } else {
                        exhausted = true;
                    }
                }
            }
        }
        exhausted = __any_sync(FULL, exhausted);
    };

    refill();
    {
        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        if (__all_sync(FULL, idle)) return;   // nothing claimed by this warp
    }

    // block-dependent start tile (see k_inside_any): here per warp
    const unsigned start =
        (unsigned)(((long long)max(__shfl_sync(FULL, orow[0], 0), 0) * ntiles) /
                   (A.n_items > 0 ? A.n_items : 1)) % (unsigned)ntiles;
    if (lane == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
        mbar_arrive_expect_tx(&bars[0], TILE_BYTES);
        tma_bulk_g2s(tbuf, tiles + (size_t)(start % ntiles) * TILE_FLOATS, TILE_BYTES, &bars[0]);
        mbar_arrive_expect_tx(&bars[1], TILE_BYTES);
        tma_bulk_g2s(tbuf + TILE_FLOATS, tiles + (size_t)((start + 1) % ntiles) * TILE_FLOATS,
                     TILE_BYTES, &bars[1]);
    }
    __syncwarp();

    unsigned long long rechecks = 0;
    unsigned int tile_units = 0;
    bool single = (TM == 1);   // warp-uniform: every live slot of the warp sits in m = 0
    for (unsigned tt = 0;; tt++) {
        const int buf = tt & 1;
        mbar_wait(&bars[buf], (tt >> 1) & 1);
        const float *T = tbuf + buf * TILE_FLOATS;
        if (single) {
            tile_units += 1;
            tile_filter32<DR, TM, 1>(a, thr_lo, thr_hi, row, hit, pend, T);
        } else {
            tile_units += TM;
            tile_filter32<DR, TM, TM>(a, thr_lo, thr_hi, row, hit, pend, T);
        }
        const int tile_first = (int)((start + tt) % (unsigned)ntiles) * REG_TILE_N;
#pragma unroll
        for (int m = 0; m < TM; m++) {
            while (__any_sync(FULL, pend[m] != 0ull)) {
                if (pend[m] != 0ull) {
                    const int col = __ffsll((long long)pend[m]) - 1;
                    pend[m] &= pend[m] - 1ull;
                    rechecks++;
                    if (tile_first + col >= A.n_live) continue;   // padded slot, no live point
                    const int lrow = A.perm32 ? A.perm32[tile_first + col] : tile_first + col;
                    const double *lp = A.live_rows + (size_t)lrow * d;
                    const double *cp = A.cand + (size_t)row[m] * d;
                    double D = 0.0;
#pragma unroll
                    for (int k0 = 0; k0 < DR; k0 += 12) {
                        double lv[12], cv[12];
#pragma unroll
                        for (int j = 0; j < 12; j++) {
                            const bool in = (k0 + j) < d;
                            lv[j] = in ? __ldg(lp + k0 + j) : 0.0;
                            cv[j] = in ? __ldg(cp + k0 + j) : 0.0;
                        }
#pragma unroll
                        for (int j = 0; j < 12; j++) D = sq_step(D, lv[j], cv[j]);
                    }
                    unc_note(A, D);
                    if (D <= A.r2) {
                        hit[m] = 1;
                        thr_lo[m] = __int_as_float(0x7f800000);
                        pend[m] = 0ull;
                    }
                }
            }
        }
#pragma unroll
        for (int m = 0; m < TM; m++) rem[m]--;
        refill();

        bool idle = true;
#pragma unroll
        for (int m = 0; m < TM; m++) idle &= row[m] < 0;
        if (__all_sync(FULL, idle)) {
            mbar_wait(&bars[(tt + 1) & 1], ((tt + 1) >> 1) & 1);   // tile tt+1 is always in flight
            break;
        }
        // ---- drain: move the survivors of the upper slots into free slot-0 lanes (shuffles)
        if (TM > 1 && exhausted && !single) {
#pragma unroll
            for (int m = 1; m < TM; m++) {
                const unsigned free0 = __ballot_sync(FULL, row[0] < 0);
                const unsigned livem = __ballot_sync(FULL, row[m] >= 0);
                if (livem != 0u && __popc(livem) <= __popc(free0)) {
                    // the r-th free lane adopts the slot of the r-th live lane
                    const bool taker = (free0 >> lane) & 1u;
                    const int rank = __popc(free0 & ((1u << lane) - 1));
                    const bool takes = taker && rank < __popc(livem);
                    const int src = takes ? (int)__fns(livem, 0, rank + 1) : lane;
#pragma unroll
                    for (int k = 0; k < DR; k++) {
                        const float v = __shfl_sync(FULL, a[m][k], src);
                        if (takes) a[0][k] = v;
                    }
                    const int r_ = __shfl_sync(FULL, row[m], src);
                    const int o_ = __shfl_sync(FULL, orow[m], src);
                    const int e_ = __shfl_sync(FULL, rem[m], src);
                    const float t_ = __shfl_sync(FULL, thr_lo[m], src);
                    const float th_ = __shfl_sync(FULL, thr_hi[m], src);
                    // which live lanes were adopted: the first popc(livem) free lanes took them all
                    if (takes) { row[0] = r_; orow[0] = o_; rem[0] = e_; thr_lo[0] = t_; thr_hi[0] = th_; hit[0] = 0; pend[0] = 0ull; }
                    if ((livem >> lane) & 1u) { row[m] = -1; thr_lo[m] = __int_as_float(0x7f800000); hit[m] = 0; pend[m] = 0ull; }
                }
            }
            bool upper = false;
#pragma unroll
            for (int m = 1; m < TM; m++) upper |= row[m] >= 0;
            single = !__any_sync(FULL, upper);
        }
        __syncwarp();   // every lane is done with tile `buf`
        if (lane == 0) {
            mbar_arrive_expect_tx(&bars[buf], TILE_BYTES);
            tma_bulk_g2s(tbuf + buf * TILE_FLOATS,
                         tiles + (size_t)((start + tt + 2) % (unsigned)ntiles) * TILE_FLOATS,
                         TILE_BYTES, &bars[buf]);
        }
    }
    if (A.stat_rechecks) {
        for (int o = 16; o > 0; o >>= 1) rechecks += __shfl_xor_sync(FULL, rechecks, o);
        if (lane == 0 && rechecks) atomicAdd(A.stat_rechecks, rechecks);
    }
    if (A.stat_tiles && lane == 0) atomicAdd(A.stat_tiles, (unsigned long long)tile_units);
}

// fp32 image of the live block, always in 64-point tiles (independent of the fp64 tile size):
// coordinates rounded to nearest, h row for radius r2 rounded UP
__global__ void k_live_build32(const double *__restrict__ rows, const double *__restrict__ norms,
                               const int *__restrict__ perm, int n, int d, int dr, int ntiles,
                               double r2, double kappa32, float *__restrict__ tiles32)
{
    int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= ntiles * REG_TILE_N) return;
    int t = slot / REG_TILE_N, c = slot - t * REG_TILE_N;
    float *F = tiles32 + (size_t)t * (dr + 1) * REG_TILE_N + c;
    const int src = (slot < n) ? (perm ? perm[slot] : slot) : -1;   // clustered tiles: slot -> live row
    for (int k = 0; k < dr; k++)
        F[(size_t)k * REG_TILE_N] =
            (src >= 0 && k < d) ? __double2float_rn(rows[(size_t)src * d + k]) : 0.f;
    float h = -1e30f;
    if (slot < n) {
        const double r2w = __dmul_rn(r2, __dadd_rn(1.0, kappa32));
        const double naw = __dmul_rn(norms[src], __dsub_rn(1.0, kappa32));
        h = __double2float_ru(__dmul_rn(0.5, __dsub_rn(r2w, naw)));
    }
    F[(size_t)dr * REG_TILE_N] = h;
}

// ---------------------------------------------------------------------------------------
// generic kernel: any d that fits shared memory; candidates k-major in shared memory,
// one candidate per thread, 8 live points per register group
// ---------------------------------------------------------------------------------------
constexpr int GTN = 8;

template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_gen(const ScanArgs A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);
    const int dr = A.dr, d = A.d, tile_n = A.tile_n;
    const int tile_doubles = (dr + 1) * tile_n;
    const uint32_t tile_bytes = (uint32_t)tile_doubles * sizeof(double);
    double *cs = reinterpret_cast<double *>(smem_raw + 128);       // [dr][SCAN_THREADS]
    double *tbuf = cs + (size_t)dr * SCAN_THREADS;

    const int tid = threadIdx.x;
    const int round = blockIdx.y;
    const double *tiles = A.tiles + (long long)round * A.round_tile_stride;
    const int n_live = A.round_nlive ? A.round_nlive[round] : A.n_live;
    const int ntiles = (n_live + tile_n - 1) / tile_n;
    const long long n_items =
        A.round_nitems ? (long long)A.round_nitems[round]
                       : (A.n_items_dev ? (long long)*A.n_items_dev : A.n_items);
    const int *items = A.item_idx ? A.item_idx + (A.round_item_off ? A.round_item_off[round] : 0)
                                  : nullptr;
    const long long base = (long long)blockIdx.x * SCAN_THREADS;
    if (base >= n_items) return;

    const long long item = base + tid;
    const bool valid = item < n_items;
    const long long row = valid ? (items ? (long long)items[item] : item) : -1;
    const long long orow = valid ? (A.out_row_idx ? (long long)A.out_row_idx[item] : row) : -1;
    double nb = 0.0;
    for (int k = 0; k < dr; k++) {
        double v = (valid && k < d) ? A.cand[row * d + k] : 0.0;
        cs[(size_t)k * SCAN_THREADS + tid] = v;
        nb = fma(v, v, nb);
    }
    CState<MODE> st;
    cs_init<MODE>(st, valid, nb, A);
    if (MODE == SCAN_SUBTRACT && valid)
        for (int k = 0; k < d; k++) A.out_rows[orow * d + k] = 0.0;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&bars[0], tile_bytes);
        tma_bulk_g2s(tbuf, tiles, tile_bytes, &bars[0]);
        if (ntiles > 1) {
            mbar_arrive_expect_tx(&bars[1], tile_bytes);
            tma_bulk_g2s(tbuf + tile_doubles, tiles + tile_doubles, tile_bytes, &bars[1]);
        }
    }

    bool warp_done = false;
    unsigned long long rechecks = 0;
    int inflight = -1;
    const double *mycs = cs + tid;
    for (int t = 0; t < ntiles; t++) {
        const int buf = t & 1;
        mbar_wait(&bars[buf], (t >> 1) & 1);
        const double *T = tbuf + (size_t)buf * tile_doubles;
        if (!warp_done) {
            for (int g = 0; g < tile_n / GTN; g++) {
                const double *Tg = T + g * GTN;
                double acc[GTN];
                {
                    const double4 h0 = *reinterpret_cast<const double4 *>(Tg + (size_t)dr * tile_n);
                    const double4 h1 = *reinterpret_cast<const double4 *>(Tg + (size_t)dr * tile_n + 4);
                    acc[0] = h0.x; acc[1] = h0.y; acc[2] = h0.z; acc[3] = h0.w;
                    acc[4] = h1.x; acc[5] = h1.y; acc[6] = h1.z; acc[7] = h1.w;
                }
#pragma unroll 4
                for (int k = 0; k < dr; k++) {
                    const double av = mycs[(size_t)k * SCAN_THREADS];
                    const double4 b0 = *reinterpret_cast<const double4 *>(Tg + (size_t)k * tile_n);
                    const double4 b1 = *reinterpret_cast<const double4 *>(Tg + (size_t)k * tile_n + 4);
                    acc[0] = fma(av, b0.x, acc[0]); acc[1] = fma(av, b0.y, acc[1]);
                    acc[2] = fma(av, b0.z, acc[2]); acc[3] = fma(av, b0.w, acc[3]);
                    acc[4] = fma(av, b1.x, acc[4]); acc[5] = fma(av, b1.y, acc[5]);
                    acc[6] = fma(av, b1.z, acc[6]); acc[7] = fma(av, b1.w, acc[7]);
                }
                bool any = false;
#pragma unroll
                for (int n = 0; n < GTN; n++) any |= cs_flag<MODE>(st, acc[n]);
                if (any) {
                    const int gbase = t * tile_n + g * GTN;
#pragma unroll
                    for (int n = 0; n < GTN; n++) {
                        if (cs_flag<MODE>(st, acc[n])) {
                            rechecks++;
                            double D = 0.0;
                            const volatile double *Tv = Tg + n;
                            for (int k = 0; k < d; k++)
                                D = sq_step(D, Tv[(size_t)k * tile_n], mycs[(size_t)k * SCAN_THREADS]);
                            if (MODE == SCAN_MIN) {
                                st.best = fmin(st.best, D);
                                st.amax = fmax(st.amax, acc[n]);
                                st.thr = __dsub_rn(st.amax, st.slack);
                            } else if (D <= A.r2) {
                                if (MODE == SCAN_FIND) {
                                    st.first = gbase + n;
                                    st.thrkey = INT_MAX;
                                } else {
                                    st.cnt++;
                                    if (MODE == SCAN_SUBTRACT)
                                        for (int k = 0; k < d; k++)
                                            A.out_rows[orow * d + k] += Tv[(size_t)k * tile_n];
                                }
                            }
                        }
                    }
                }
            }
            warp_done = __all_sync(FULL, cs_done<MODE>(st));
        }
        const bool all_done = __syncthreads_and(warp_done);
        if (all_done) {
            if (t + 1 < ntiles) inflight = t + 1;
            break;
        }
        if (tid == 0 && t + 2 < ntiles) {
            mbar_arrive_expect_tx(&bars[buf], tile_bytes);
            tma_bulk_g2s(tbuf + (size_t)buf * tile_doubles, tiles + (size_t)(t + 2) * tile_doubles,
                         tile_bytes, &bars[buf]);
        }
    }
    if (inflight >= 0) mbar_wait(&bars[inflight & 1], (inflight >> 1) & 1);

    double blockmax = 0.0;
    if (valid) {
        if (MODE == SCAN_FIND) {
            if (A.out_idx) A.out_idx[orow] = st.first;
            if (A.out_mask) A.out_mask[orow] = st.first >= 0;
        } else if (MODE == SCAN_COUNT) {
            A.out_idx[orow] = st.cnt;
        } else if (MODE == SCAN_SUBTRACT) {
            const double cnt = (double)st.cnt;
            for (int k = 0; k < d; k++) {
                double s = A.out_rows[orow * d + k];
                A.out_rows[orow * d + k] = __dsub_rn(A.cand[row * d + k], __ddiv_rn(s, cnt));
            }
        } else {
            if (A.out_min) A.out_min[orow] = st.best;
            blockmax = st.best;
        }
    }
    if (MODE == SCAN_MIN) {
        for (int o = 16; o > 0; o >>= 1) blockmax = fmax(blockmax, __shfl_xor_sync(FULL, blockmax, o));
        if ((tid & 31) == 0 && A.out_round_max)
            atomicMax(A.out_round_max + round, f64_bits(blockmax));
    }
    if (A.stat_rechecks) {
        for (int o = 16; o > 0; o >>= 1) rechecks += __shfl_xor_sync(FULL, rechecks, o);
        if ((tid & 31) == 0 && rechecks) atomicAdd(A.stat_rechecks, rechecks);
    }
}

// ---------------------------------------------------------------------------------------
// plain exact kernel: the reference's triple loop, one thread per candidate, no filter, no
// shared memory.  Slow by design; used for UNB_OPT_EXACT_ONLY (validation of the filtered
// kernels on the GPU itself) and for d too large for the tiled kernels.
// ---------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_exact(const ScanArgs A)
{
    const int round = blockIdx.y;
    const int d = A.d;
    const int n_live = A.round_nlive ? A.round_nlive[round] : A.n_live;
    const long long n_items =
        A.round_nitems ? (long long)A.round_nitems[round]
                       : (A.n_items_dev ? (long long)*A.n_items_dev : A.n_items);
    const int *items = A.item_idx ? A.item_idx + (A.round_item_off ? A.round_item_off[round] : 0)
                                  : nullptr;
    const int *lidx = A.live_idx ? A.live_idx + (A.round_live_off ? A.round_live_off[round] : 0)
                                 : nullptr;
    const long long item = (long long)blockIdx.x * SCAN_THREADS + threadIdx.x;
    const bool valid = item < n_items;
    double best = 1e300;
    if (valid) {
        const long long row = items ? (long long)items[item] : item;
        const long long orow = A.out_row_idx ? (long long)A.out_row_idx[item] : row;
        const double *b = A.cand + row * d;
        int first = -1, cnt = 0;
        if (MODE == SCAN_SUBTRACT)
            for (int k = 0; k < d; k++) A.out_rows[orow * d + k] = 0.0;
        for (int i = 0; i < n_live; i++) {
            const double *a = A.live_rows + (size_t)(lidx ? lidx[i] : i) * d;
            double D = 0.0;
            for (int k = 0; k < d; k++) D = sq_step(D, a[k], b[k]);
            if (MODE == SCAN_MIN) {
                best = fmin(best, D);
            } else if (D <= A.r2) {
                if (MODE == SCAN_FIND) { first = i; break; }
                cnt++;
                if (MODE == SCAN_SUBTRACT)
                    for (int k = 0; k < d; k++) A.out_rows[orow * d + k] += a[k];
            }
        }
        if (MODE == SCAN_FIND) {
            if (A.out_idx) A.out_idx[orow] = first;
            if (A.out_mask) A.out_mask[orow] = first >= 0;
        } else if (MODE == SCAN_COUNT) {
            A.out_idx[orow] = cnt;
        } else if (MODE == SCAN_SUBTRACT) {
            for (int k = 0; k < d; k++) {
                double s = A.out_rows[orow * d + k];
                A.out_rows[orow * d + k] = __dsub_rn(b[k], __ddiv_rn(s, (double)cnt));
            }
        } else if (A.out_min) {
            A.out_min[orow] = best;
        }
    }
    if (MODE == SCAN_MIN) {
        double m = valid ? best : 0.0;
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, o));
        if ((threadIdx.x & 31) == 0 && A.out_round_max)
            atomicMax(A.out_round_max + round, f64_bits(m));
    }
}

// ---------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------
template <typename K>
int set_smem(unb_ctx *ctx, K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)bytes));
    return UNB_OK;
}

template <int DR, int TM, int MODE>
int launch_reg(unb_ctx *ctx, const ScanArgs &a, int rounds, long long max_items, cudaStream_t s)
{
    const size_t smem = 128 + 2 * (size_t)(DR + 1) * REG_TILE_N * sizeof(double);
    UNB_TRY(set_smem(ctx, k_scan_reg<DR, TM, MODE>, smem));
    long long bx = (max_items + SCAN_THREADS * TM - 1) / (SCAN_THREADS * TM);
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)rounds);
    k_scan_reg<DR, TM, MODE><<<grid, SCAN_THREADS, smem, s>>>(a);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

template <int DR, int MODE>
int launch_reg_tm(unb_ctx *ctx, const ScanArgs &a, int rounds, long long max_items, cudaStream_t s)
{
    // candidates per thread: the LDS.128 : DFMA ratio is 1 : 2*TM, so TM >= 2 keeps the shared
    // memory pipe below the DP pipe; tiny d affords 4.  SUBTRACT keeps TM=1 (its hit path
    // read-modify-writes global rows), and small batches favour more blocks.
    constexpr int TM_BIG = (DR <= 8) ? 4 : 2;
    if constexpr (MODE == SCAN_SUBTRACT) {
        return launch_reg<DR, 1, MODE>(ctx, a, rounds, max_items, s);
    } else {
        if (max_items * rounds < (long long)ctx->sm_count * SCAN_THREADS * 2 * TM_BIG)
            return launch_reg<DR, 1, MODE>(ctx, a, rounds, max_items, s);
        return launch_reg<DR, TM_BIG, MODE>(ctx, a, rounds, max_items, s);
    }
}

template <int MODE>
int launch_mode(unb_ctx *ctx, const ScanArgs &a, int rounds, long long max_items, cudaStream_t s)
{
    if (ctx->exact_only || a.tiles == nullptr) {
        long long bx = (max_items + SCAN_THREADS - 1) / SCAN_THREADS;
        if (bx < 1) bx = 1;
        dim3 grid((unsigned)bx, (unsigned)rounds);
        k_scan_exact<MODE><<<grid, SCAN_THREADS, 0, s>>>(a);
        ctx->launches++;
        UNB_CUDA(ctx, cudaGetLastError());
        return UNB_OK;
    }
    if (a.dr <= 32 && a.tile_n == REG_TILE_N) {
        switch (a.dr) {
        case 4: return launch_reg_tm<4, MODE>(ctx, a, rounds, max_items, s);
        case 8: return launch_reg_tm<8, MODE>(ctx, a, rounds, max_items, s);
        case 12: return launch_reg_tm<12, MODE>(ctx, a, rounds, max_items, s);
        case 16: return launch_reg_tm<16, MODE>(ctx, a, rounds, max_items, s);
        case 20: return launch_reg_tm<20, MODE>(ctx, a, rounds, max_items, s);
        case 24: return launch_reg_tm<24, MODE>(ctx, a, rounds, max_items, s);
        case 28: return launch_reg_tm<28, MODE>(ctx, a, rounds, max_items, s);
        case 32: return launch_reg_tm<32, MODE>(ctx, a, rounds, max_items, s);
        default: break;
        }
    }
    const size_t smem = 128 + ((size_t)a.dr * SCAN_THREADS + 2 * (size_t)(a.dr + 1) * a.tile_n) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_scan_gen<MODE>, smem));
    long long bx = (max_items + SCAN_THREADS - 1) / SCAN_THREADS;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, (unsigned)rounds);
    k_scan_gen<MODE><<<grid, SCAN_THREADS, smem, s>>>(a);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

template <int DR, int TM>
int launch_any(unb_ctx *ctx, const ScanArgs &a, int *queue_head, cudaStream_t s)
{
    const size_t smem = 128 + (2 * (size_t)(DR + 1) * REG_TILE_N +
                               (size_t)(DR + 2) * ANY_STAGE_SLOTS) * sizeof(double);
    static int per_sm = 0;   // per instantiation; the engine drives one device per process
    if (per_sm == 0) {
        UNB_TRY(set_smem(ctx, k_inside_any<DR, TM>, smem));
        UNB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inside_any<DR, TM>,
                                                                   SCAN_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    long long bx = (a.n_items + SCAN_THREADS * TM - 1) / (SCAN_THREADS * TM);
    const long long resident = (long long)per_sm * ctx->sm_count;
    if (bx > resident) bx = resident;
    if (bx < 1) bx = 1;
    k_inside_any<DR, TM><<<(unsigned)bx, SCAN_THREADS, smem, s>>>(a, queue_head);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

template <int DR, int TM>
int launch_any32(unb_ctx *ctx, const ScanArgs &a, int *queue_head, cudaStream_t s)
{
    const bool blocksync = ctx->block_kernel != 0;   // UNB_OPT_BLOCK_KERNEL
    long long bx = (a.n_items + SCAN_THREADS * TM - 1) / (SCAN_THREADS * TM);
    // Warp-independent streams win on small launches (latency: no barrier, warps retire alone;
    // measured 0.127 vs 0.150 ms at 4096 proposals, 333 vs 428 us per integrator iteration); the
    // block-synchronous kernel wins on bulk launches (block-wide drain compaction, 4x less tile
    // traffic; 0.66 vs 0.71 ms at 2^20).  Per-warp tile buffers also need d <= 32.
    if (!blocksync && DR <= 32 && a.n_items <= 16384 && !a.bin_order) {
        const size_t smem = 128 + (size_t)(SCAN_THREADS / 32) * 2 * (DR + 1) * REG_TILE_N * sizeof(float);
        static int per_sm_w = 0;
        if (per_sm_w == 0) {
            UNB_TRY(set_smem(ctx, k_inside_any32w<DR, TM>, smem));
            UNB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                              &per_sm_w, k_inside_any32w<DR, TM>, SCAN_THREADS, smem));
            if (per_sm_w < 1) per_sm_w = 1;
        }
        const long long resident = (long long)per_sm_w * ctx->sm_count;
        if (bx > resident) bx = resident;
        if (bx < 1) bx = 1;
        k_inside_any32w<DR, TM><<<(unsigned)bx, SCAN_THREADS, smem, s>>>(a, queue_head);
        ctx->launches++;
        UNB_CUDA(ctx, cudaGetLastError());
        return UNB_OK;
    }
    const size_t smem = 128 + (2 * (size_t)(DR + 1) * REG_TILE_N + (size_t)(DR + 5) * ANY_STAGE_SLOTS) *
                                  sizeof(float);
    static int per_sm = 0;   // per instantiation; the engine drives one device per process
    if (per_sm == 0) {
        UNB_TRY(set_smem(ctx, k_inside_any32<DR, TM>, smem));
        UNB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_inside_any32<DR, TM>,
                                                                   SCAN_THREADS, smem));
        if (per_sm < 1) per_sm = 1;
    }
    const long long resident = (long long)per_sm * ctx->sm_count;
    if (bx > resident) bx = resident;
    if (bx < 1) bx = 1;
    int coop_max = ctx->coop_max < 0 ? ANY_COOP_MAX : ctx->coop_max;   // UNB_OPT_COOP_MAX
    if (coop_max > ANY_STAGE_SLOTS) coop_max = ANY_STAGE_SLOTS;
    ScanArgs ac = a;
    ac.coop_max = coop_max;
    k_inside_any32<DR, TM><<<(unsigned)bx, SCAN_THREADS, smem, s>>>(ac, queue_head);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

template <int DR>
int launch_any_tm(unb_ctx *ctx, const ScanArgs &a, int *queue_head, cudaStream_t s)
{
    if (a.tiles32) {   // fp32 pre-filter selected by the caller
        constexpr int TM32 = (DR <= 8) ? 4 : 2;
        if (a.n_items < (long long)ctx->sm_count * SCAN_THREADS * 2 * TM32)
            return launch_any32<DR, 1>(ctx, a, queue_head, s);
        return launch_any32<DR, TM32>(ctx, a, queue_head, s);
    }
    constexpr int TM_BIG = (DR <= 8) ? 4 : 2;
    if (a.n_items < (long long)ctx->sm_count * SCAN_THREADS * 2 * TM_BIG)
        return launch_any<DR, 1>(ctx, a, queue_head, s);
    return launch_any<DR, TM_BIG>(ctx, a, queue_head, s);
}

}  // namespace

// membership-only scan (mask, no index): persistent refill kernel when the shape allows,
// else the ordered FIND kernel.  queue_head must be a zeroed device int.
int unb_launch_inside_any(unb_ctx *ctx, const ScanArgs &a, int *queue_head, cudaStream_t s)
{
    if (a.n_items <= 0) return UNB_OK;
    const bool mask_only = !ctx->exact_only && a.out_mask && !a.out_idx && a.n_live > 0;
    if (mask_only && a.tiles32 && a.dr > 32 && a.dr <= ANY32_MAX_DR) {
        // 32 < d <= 128: fp32 pre-filter kernel only, one candidate per thread
        switch (a.dr) {
#define UNB_CASE32(DR_) case DR_: return launch_any32<DR_, 1>(ctx, a, queue_head, s);
            UNB_CASE32(36) UNB_CASE32(40) UNB_CASE32(44) UNB_CASE32(48) UNB_CASE32(52)
            UNB_CASE32(56) UNB_CASE32(60) UNB_CASE32(64) UNB_CASE32(68) UNB_CASE32(72)
            UNB_CASE32(76) UNB_CASE32(80) UNB_CASE32(84) UNB_CASE32(88) UNB_CASE32(92)
            UNB_CASE32(96) UNB_CASE32(100) UNB_CASE32(104) UNB_CASE32(108) UNB_CASE32(112)
            UNB_CASE32(116) UNB_CASE32(120) UNB_CASE32(124) UNB_CASE32(128)
#undef UNB_CASE32
        default: break;
        }
    }
    if (mask_only && a.dr <= 32 && (a.tiles32 || (a.tiles && a.tile_n == REG_TILE_N))) {
        switch (a.dr) {
        case 4: return launch_any_tm<4>(ctx, a, queue_head, s);
        case 8: return launch_any_tm<8>(ctx, a, queue_head, s);
        case 12: return launch_any_tm<12>(ctx, a, queue_head, s);
        case 16: return launch_any_tm<16>(ctx, a, queue_head, s);
        case 20: return launch_any_tm<20>(ctx, a, queue_head, s);
        case 24: return launch_any_tm<24>(ctx, a, queue_head, s);
        case 28: return launch_any_tm<28>(ctx, a, queue_head, s);
        case 32: return launch_any_tm<32>(ctx, a, queue_head, s);
        default: break;
        }
    }
    return unb_launch_scan(ctx, SCAN_FIND, a, 1, s);
}

// ---------------------------------------------------------------------------------------
// exported (library-internal) entry points
// ---------------------------------------------------------------------------------------
double unb_kappa(size_t d) { return (8.0 * (double)d + 64.0) * 1.1102230246251565e-16; }

// largest tile that lets the generic kernel keep 128 candidates + 2 tiles in 227 KB
size_t unb_pick_tile_n(size_t d)
{
    size_t dr = (d + 3) / 4 * 4;
    if (dr <= 32) return REG_TILE_N;
    const size_t budget = 227 * 1024 - 128;
    for (size_t tn = 64; tn >= 8; tn /= 2) {
        size_t need = (dr * SCAN_THREADS + 2 * (dr + 1) * tn) * sizeof(double);
        if (need <= budget) return tn;
    }
    return 0;   // does not fit: caller falls back to the plain exact kernel
}

int unb_live_build(unb_ctx *ctx, LiveTiles &L, const double *rows_dev, size_t n, size_t d,
                   cudaStream_t s)
{
    L.valid = false;
    L.t32_valid = false;
    L.cluster_valid = false;
    L.t32_clustered = false;
    L.n = n;
    L.d = d;
    L.dr = (d + 3) / 4 * 4;
    L.tile_n = unb_pick_tile_n(d);
    L.h_mode = HMODE_NONE;
    UNB_TRY(unb_reserve(ctx, L.rows, (n ? n : 1) * d * sizeof(double)));
    if (rows_dev != L.rows.p && n)
        UNB_CUDA(ctx, cudaMemcpyAsync(L.rows.p, rows_dev, n * d * sizeof(double),
                                      cudaMemcpyDeviceToDevice, s));
    UNB_TRY(unb_reserve(ctx, L.namax, sizeof(unsigned long long)));
    UNB_CUDA(ctx, cudaMemsetAsync(L.namax.p, 0, sizeof(unsigned long long), s));
    if (L.tile_n == 0 || n == 0) {   // plain exact kernels only
        L.ntiles = 0;
        L.valid = true;
        return UNB_OK;
    }
    L.ntiles = (n + L.tile_n - 1) / L.tile_n;
    const size_t slots = L.ntiles * L.tile_n;
    UNB_TRY(unb_reserve(ctx, L.tiles, L.ntiles * (L.dr + 1) * L.tile_n * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, L.norms, slots * sizeof(double)));
    k_live_build<<<(unsigned)((slots + 127) / 128), 128, 0, s>>>(
        (const double *)L.rows.p, (int)n, (int)d, (int)L.dr, (int)L.tile_n, (int)L.ntiles,
        (double *)L.tiles.p, (double *)L.norms.p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    k_norm_max<<<8, 256, 0, s>>>((const double *)L.norms.p, (int)n,
                                 (unsigned long long *)L.namax.p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    L.valid = true;
    return UNB_OK;
}

int unb_live_update_rows(unb_ctx *ctx, LiveTiles &L, const int *rows_dev_idx, size_t nrows,
                         cudaStream_t s)
{
    if (!nrows || L.ntiles == 0) return UNB_OK;
    // one launch refreshes every image of the touched rows (unb_live.cu); the fp32 image follows in
    // place unless its tiles are in cluster order (then it is rebuilt through the permutation)
    const bool keep32 = L.t32_valid && !L.t32_clustered;
    LiveUpdateArgs a;
    memset(&a, 0, sizeof(a));
    a.rows = (const double *)L.rows.p;
    a.idx = rows_dev_idx;
    a.nrows = (int)nrows;
    a.n = (int)L.n;
    a.d = (int)L.d;
    a.dr = (int)L.dr;
    a.tile_n = (int)L.tile_n;
    a.tiles = (double *)L.tiles.p;
    a.norms = (double *)L.norms.p;
    a.namax = (unsigned long long *)L.namax.p;
    a.h_mode = L.h_mode;
    a.h_r2 = L.h_r2;
    a.kappa = unb_kappa(L.d);
    a.tiles32 = keep32 ? (float *)L.tiles32.p : nullptr;
    a.t32_r2 = L.t32_r2;
    a.kappa32 = unb_kappa32(L.d);
    UNB_TRY(unb_launch_live_update(ctx, a, s));
    if (!keep32) L.t32_valid = false;
    return UNB_OK;
}

// the host's copy of the norm bound follows row updates without a device round trip: same
// k-sequential FMA chain as the kernels, so the two bounds agree bit for bit
void unb_live_note_host_row(LiveTiles &L, const double *row)
{
    double na = 0.0;
    for (size_t k = 0; k < L.d; k++) na = std::fma(row[k], row[k], na);
    if (na > L.namax_host) L.namax_host = na;
}

int unb_live_set_h(unb_ctx *ctx, LiveTiles &L, int h_mode, double r2, cudaStream_t s)
{
    if (L.ntiles == 0) return UNB_OK;
    if (L.h_mode == h_mode && (h_mode == HMODE_MIN || L.h_r2 == r2)) return UNB_OK;
    const size_t slots = L.ntiles * L.tile_n;
    k_live_set_h<<<(unsigned)((slots + 127) / 128), 128, 0, s>>>(
        (const double *)L.norms.p, (int)L.n, (int)L.dr, (int)L.tile_n, (int)L.ntiles, h_mode, r2,
        unb_kappa(L.d), (double *)L.tiles.p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    L.h_mode = h_mode;
    L.h_r2 = r2;
    return UNB_OK;
}

double unb_kappa32(size_t d) { return (4.0 * (double)d + 32.0) * 5.9604644775390625e-08; }

int unb_live_prepare32(unb_ctx *ctx, LiveTiles &L, double r2, bool *usable, cudaStream_t s)
{
    *usable = false;
    // needs the squared norms (computed with the fp64 tiles) and candidates that fit registers
    if (!ctx->filter_fp32 || ctx->exact_only || L.ntiles == 0 || L.dr > ANY32_MAX_DR)
        return UNB_OK;
    if (!L.t32_valid) {   // (re)built block: fetch the largest squared norm once
        unsigned long long bits = 0;
        UNB_CUDA(ctx, cudaMemcpyAsync(&bits, L.namax.p, sizeof(bits), cudaMemcpyDeviceToHost, s));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));
        memcpy(&L.namax_host, &bits, sizeof(double));
    }
    const double k32 = unb_kappa32(L.d);
    // ranges far from fp32 overflow/underflow, and a shell of false alarms that is thin
    // compared with the radius (otherwise the fp64 filter is the better tool)
    if (!(r2 >= 1e-30 && r2 <= 1e30 && L.namax_host <= 1e30)) return UNB_OK;
    if (!(k32 * (2.0 * L.namax_host + r2) <= r2 / 16.0)) return UNB_OK;
    const size_t ntiles32 = (L.n + REG_TILE_N - 1) / REG_TILE_N;
    // clustered tile order (unb_cluster.cu): computed once per (re)built block when a large launch
    // asks for it, kept across in-place row updates (a permutation stays a permutation)
    const bool can_cluster = L.dr <= 32 && ntiles32 >= 4 && ntiles32 <= unb_cluster_max_tiles();
    if (L.want_cluster && !L.cluster_valid && can_cluster) {
        const size_t K = ntiles32;
        UNB_TRY(unb_reserve(ctx, L.perm32, ntiles32 * REG_TILE_N * sizeof(int)));
        UNB_TRY(unb_reserve(ctx, L.cl_scratch_i, (3 * L.n + 4 * K + 8) * sizeof(int)));
        UNB_TRY(unb_reserve(ctx, L.cl_scratch_f, 2 * K * L.d * sizeof(float)));
        UNB_TRY(unb_reserve(ctx, L.ctiles32, ((K + REG_TILE_N - 1) / REG_TILE_N) * (L.dr + 1) * REG_TILE_N * sizeof(float)));
        UNB_TRY(unb_launch_cluster_live(ctx, (const double *)L.rows.p, (int)L.n, (int)L.d, (int)K,
                                        (int *)L.perm32.p, (int *)L.cl_scratch_i.p,
                                        (float *)L.cl_scratch_f.p, s));
        L.cluster_valid = true;
    }
    const bool clustered = L.cluster_valid && can_cluster;
    if (!L.t32_valid || L.t32_r2 != r2 || L.t32_clustered != clustered) {
        const size_t slots = ntiles32 * REG_TILE_N;
        UNB_TRY(unb_reserve(ctx, L.tiles32, ntiles32 * (L.dr + 1) * REG_TILE_N * sizeof(float)));
        k_live_build32<<<(unsigned)((slots + 127) / 128), 128, 0, s>>>(
            (const double *)L.rows.p, (const double *)L.norms.p,
            clustered ? (const int *)L.perm32.p : nullptr, (int)L.n, (int)L.d, (int)L.dr,
            (int)ntiles32, r2, k32, (float *)L.tiles32.p);
        ctx->launches++;
        UNB_CUDA(ctx, cudaGetLastError());
        if (clustered)   // tile centroids follow the rows (in-place row updates move them a little)
            UNB_TRY(unb_launch_tile_centroids(ctx, (const double *)L.rows.p, (const int *)L.perm32.p,
                                              (int)L.n, (int)L.d, (int)L.dr, (int)ntiles32,
                                              (float *)L.ctiles32.p, s));
        L.t32_valid = true;
        L.t32_r2 = r2;
        L.t32_clustered = clustered;
    }
    *usable = true;
    return UNB_OK;
}

int unb_launch_gather_round_tiles(unb_ctx *ctx, const double *rows, int n, int d, int dr,
                                  int tile_n, const int *idxA, const int *offA, const int *nA,
                                  int rounds, long long round_tile_stride, double *tiles,
                                  cudaStream_t s)
{
    const int max_tiles = (n + tile_n - 1) / tile_n;
    dim3 grid((unsigned)((max_tiles * tile_n + 127) / 128), (unsigned)rounds);
    k_gather_round_tiles<<<grid, 128, 0, s>>>(rows, d, dr, tile_n, idxA, offA, nA,
                                              round_tile_stride, tiles);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_scan(unb_ctx *ctx, int mode, const ScanArgs &a, int rounds, cudaStream_t s)
{
    if (a.n_items <= 0 || rounds <= 0) return UNB_OK;
    switch (mode) {
    case SCAN_FIND: return launch_mode<SCAN_FIND>(ctx, a, rounds, a.n_items, s);
    case SCAN_COUNT: return launch_mode<SCAN_COUNT>(ctx, a, rounds, a.n_items, s);
    case SCAN_SUBTRACT: return launch_mode<SCAN_SUBTRACT>(ctx, a, rounds, a.n_items, s);
    case SCAN_MIN: return launch_mode<SCAN_MIN>(ctx, a, rounds, a.n_items, s);
    default: return unb_fail(ctx, UNB_ERR_ARG, "unknown scan mode %d", mode);
    }
}
