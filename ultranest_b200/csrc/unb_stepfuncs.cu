// unb_stepfuncs.cu -- population step-sampler helpers: the device side of
// ultranest/stepfuncs.pyx
//   * within_unit_cube                     stepfuncs.pyx:22-52
//   * evolve_prepare / evolve_update       stepfuncs.pyx:57-94, 99-183
//   * evolve, fused with a built-in prior transform and likelihood   stepfuncs.pyx:189-282
//   * step_back                            stepfuncs.pyx:285-334
//   * update_vectorised_slice_sampler      stepfuncs.pyx:537-630
// and of the inner loop of PopulationSimpleSliceSampler.__next__ (popstepsampler.py:940-965)
// with the population state resident on the device between iterations (unb_popslice_*).
// One thread owns one walker / worker / point; walkers are independent, so every kernel is a
// plain row kernel.  Random draws stay on the host (np.random, reference order) and arrive as
// plain arrays.  Built with -fmad=false: multiply-adds are non-fused like the NumPy expressions.
#include "unb_internal.cuh"
#include "unb_loglike.cuh"

#include <climits>
#include <cmath>
#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int SF_THREADS = 128;
constexpr size_t SF_SMEM_BUDGET = 200 * 1024;
constexpr int STEP_BACK_MAXCOLS = 2048;   // flag bits per walker held in local memory

__host__ __device__ inline int sf_odd(int d) { return d | 1; }

// ---------------------------------------------------------------------------------------
// small host helpers
// ---------------------------------------------------------------------------------------
inline cudaStream_t S0(unb_ctx *ctx) { return ctx->lane[0].stream; }

int up(unb_ctx *ctx, DevBuf &b, const void *src, size_t bytes, cudaStream_t s)
{
    UNB_TRY(unb_reserve(ctx, b, bytes ? bytes : 8));
    if (!bytes) return UNB_OK;
    UNB_CUDA(ctx, cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, s));
    ctx->h2d_bytes += (long long)bytes;
    return UNB_OK;
}

int down(unb_ctx *ctx, void *dst, const DevBuf &b, size_t bytes, cudaStream_t s)
{
    if (!bytes) return UNB_OK;
    UNB_CUDA(ctx, cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += (long long)bytes;
    return UNB_OK;
}

inline unsigned blocks_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

int check(unb_ctx *ctx)
{
    if (!ctx) return UNB_ERR_ARG;
    UNB_CUDA(ctx, cudaSetDevice(ctx->device));
    return UNB_OK;
}

int row_threads_for(int doubles_per_thread)
{
    size_t t = SF_SMEM_BUDGET / ((size_t)doubles_per_thread * sizeof(double));
    if (t >= (size_t)SF_THREADS) return SF_THREADS;
    return (int)(t / 32 * 32);
}

template <typename K>
int allow_smem(unb_ctx *ctx, K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return UNB_OK;
}

// ---------------------------------------------------------------------------------------
// within_unit_cube (stepfuncs.pyx:30-34): every coordinate strictly inside (0, 1)
// ---------------------------------------------------------------------------------------
__global__ void k_within_unit_cube(const double *__restrict__ u, long long n, int d,
                                   unsigned char *__restrict__ acceptable)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *row = u + i * d;
    bool ok = true;
    for (int j = 0; j < d; j++) {
        const double x = row[j];
        if (!(0.0 < x && x < 1.0)) { ok = false; break; }
    }
    acceptable[i] = ok ? 1 : 0;
}

// evolve_prepare (stepfuncs.pyx:67-69)
__global__ void k_evolve_prepare(const unsigned char *__restrict__ searching_left,
                                 const unsigned char *__restrict__ searching_right, long long n,
                                 unsigned char *__restrict__ search_right,
                                 unsigned char *__restrict__ bisecting)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool sl = searching_left[i] != 0, sr = searching_right[i] != 0;
    search_right[i] = (!sl && sr) ? 1 : 0;
    bisecting[i] = !(sl || sr) ? 1 : 0;
}

// exclusive prefix count of non-zero flags (one block; the rank of an acceptable walker is its
// row in the compacted likelihood vector, stepfuncs.pyx:152-156)
__global__ void __launch_bounds__(1024) k_flag_ranks(const unsigned char *__restrict__ flags, int n,
                                                     int *__restrict__ rank, int *__restrict__ total)
{
    __shared__ int warp_sum[32];
    __shared__ int carry, chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const bool f = i < n && flags[i] != 0;
        const unsigned ball = __ballot_sync(FULL, f);
        const int within = __popc(ball & ((1u << lane) - 1));
        if (lane == 0) warp_sum[warp] = __popc(ball);
        __syncthreads();
        if (warp == 0) {
            const int v = warp_sum[lane];
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int other = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += other;
            }
            warp_sum[lane] = incl - v;            // exclusive warp offsets
            if (lane == 31) chunk_total = incl;
        }
        __syncthreads();
        if (i < n) rank[i] = carry + warp_sum[warp] + within;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__device__ __forceinline__ void slice_state_update(bool succ, bool search_right, bool bisecting,
                                                   double &t, double &left, double &right, bool &sl,
                                                   bool &sr, bool &succ_out)
{
    // stepfuncs.pyx:158-183
    if (succ) {
        if (sl) left = __dmul_rn(left, 2.0);
        else if (search_right) right = __dmul_rn(right, 2.0);
    } else {
        if (sl) sl = false;
        else if (search_right) sr = false;
    }
    if (bisecting) {
        if (t < 0) left = t;
        else right = t;
        if (succ) t = __longlong_as_double(0x7ff8000000000000LL);
        succ_out = succ;
    } else {
        succ_out = false;
    }
}

__global__ void k_evolve_update(const unsigned char *__restrict__ acceptable,
                                const double *__restrict__ Lnew, const int *__restrict__ rank,
                                long long n_lnew, double Lmin,
                                const unsigned char *__restrict__ search_right,
                                const unsigned char *__restrict__ bisecting,
                                double *__restrict__ currentt, double *__restrict__ current_left,
                                double *__restrict__ current_right,
                                unsigned char *__restrict__ searching_left,
                                unsigned char *__restrict__ searching_right,
                                unsigned char *__restrict__ success, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool succ = success[i] != 0;   // :152-156 only ever sets the flag
    if (acceptable[i]) {
        const int j = rank[i];
        if (j < n_lnew && Lnew[j] > Lmin) succ = true;
    }
    double t = currentt[i], left = current_left[i], right = current_right[i];
    bool sl = searching_left[i] != 0, sr = searching_right[i] != 0, out;
    slice_state_update(succ, search_right[i] != 0, bisecting[i] != 0, t, left, right, sl, sr, out);
    currentt[i] = t;
    current_left[i] = left;
    current_right[i] = right;
    searching_left[i] = sl ? 1 : 0;
    searching_right[i] = sr ? 1 : 0;
    success[i] = out ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// evolve with a device transform + likelihood (stepfuncs.pyx:250-274 as ONE kernel): proposal on
// the slice, unit-cube test, v = transform(u), L = loglike(v), slice-state update.  currentt of
// the bisecting walkers already holds the host's uniform draws (:255).
// ---------------------------------------------------------------------------------------
struct EvolveArgs {
    double *u;                 // in: slice origins; out: proposals (the reference's alias, :252)
    const double *v;
    double *t, *left, *right;
    unsigned char *sl, *sr;
    unsigned char *acceptable, *success;
    double *like;              // out: likelihood of the acceptable walkers, -inf elsewhere
    long long n;
    int d;
    int xform_kind;
    const double *xform_scale, *xform_lo;
    int loglike_kind;
    const double *lparams;
    double Lmin;
};

__device__ __forceinline__ void apply_xform(int kind, const double *scale, const double *lo,
                                            const double *u, double *v, int d)
{
    if (kind == UNB_XFORM_SCALE_SHIFT)
        for (int k = 0; k < d; k++) v[k] = __dadd_rn(__dmul_rn(u[k], __ldg(scale + k)), __ldg(lo + k));
    else
        for (int k = 0; k < d; k++) v[k] = u[k];
}

__global__ void k_evolve_fused(const EvolveArgs A)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int d = A.d, ds = sf_odd(d);
    double *un = rowbuf + (size_t)threadIdx.x * 2 * ds;   // proposal, then v = transform(u)
    double *tt = un + ds;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    bool sl = A.sl[i] != 0, sr = A.sr[i] != 0;
    const bool search_right = !sl && sr, bisecting = !(sl || sr);
    double t = A.t[i], left = A.left[i], right = A.right[i];
    const double coef = sl ? left : (search_right ? right : t);   // :253-256
    double *urow = A.u + i * d;
    const double *vrow = A.v + i * d;
    bool acceptable = true;
    for (int k = 0; k < d; k++) {
        const double x = __dadd_rn(urow[k], __dmul_rn(vrow[k], coef));
        un[k] = x;
        urow[k] = x;
        acceptable = acceptable && (0.0 < x && x < 1.0);
    }
    double like = -__longlong_as_double(0x7ff0000000000000LL);
    if (acceptable) {
        apply_xform(A.xform_kind, A.xform_scale, A.xform_lo, un, un, d);
        like = loglike_row(A.loglike_kind, un, d, tt, A.lparams);
    }
    bool out;
    slice_state_update(acceptable && like > A.Lmin, search_right, bisecting, t, left, right, sl, sr, out);
    A.t[i] = t;
    A.left[i] = left;
    A.right[i] = right;
    A.sl[i] = sl ? 1 : 0;
    A.sr[i] = sr ? 1 : 0;
    A.acceptable[i] = acceptable ? 1 : 0;
    A.success[i] = out ? 1 : 0;
    A.like[i] = like;
}

// ---------------------------------------------------------------------------------------
// step_back (stepfuncs.pyx:306-334)
// ---------------------------------------------------------------------------------------
__global__ void k_i64_max(const long long *__restrict__ x, long long n, long long *__restrict__ out)
{
    long long best = LLONG_MIN;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        best = x[i] > best ? x[i] : best;
    for (int o = 16; o > 0; o >>= 1) {
        const long long other = __shfl_xor_sync(FULL, best, o);
        best = other > best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out, best);
}

__global__ void k_step_back(double Lmin, double *__restrict__ allL, long long nwalkers, long long ncols,
                            long long *__restrict__ generation, double *__restrict__ currentt,
                            const long long *__restrict__ gen_max)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwalkers) return;
    long long max_width = *gen_max + 1;                 // :308
    if (max_width > ncols) max_width = ncols;
    if (max_width <= 0) return;
    unsigned long long bits[STEP_BACK_MAXCOLS / 64];
    double *row = allL + i * ncols;
    int remaining = 0;
    for (long long w = 0; w * 64 < max_width; w++) {
        unsigned long long b = 0;
        for (int k = 0; k < 64 && w * 64 + k < max_width; k++)
            if (row[w * 64 + k] < Lmin) b |= 1ull << k;   // :309
        bits[w] = b;
        remaining += __popcll(b);
    }
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    long long g = generation[i];
    bool touched = false;
    while (remaining > 0) {                              // :319-334, this walker's share
        const long long ga = g < 0 ? g + ncols : g;      // NumPy index wrap of allL[i, g]
        const long long gb = g < 0 ? g + max_width : g;  // ... and of below_threshold_parent[., g]
        // out-of-range generations raise IndexError in the reference; the host mirror validates
        // them before the launch (stepfuncs.step_back), the kernel never writes out of bounds
        if (ga < 0 || gb < 0 || ga >= ncols || gb >= max_width) break;
        row[ga] = nan;
        const unsigned long long bit = 1ull << (gb & 63);
        if (bits[gb >> 6] & bit) {
            bits[gb >> 6] &= ~bit;
            remaining--;
        }
        g -= 1;
        touched = true;
    }
    if (touched) {
        generation[i] = g;
        currentt[i] = nan;
    }
}

// ---------------------------------------------------------------------------------------
// update_vectorised_slice_sampler (stepfuncs.pyx:608-628).  The reference walks the workers in
// order; only workers of the same point interact, so the thread of point w replays its workers in
// order and leaves the others alone.
// ---------------------------------------------------------------------------------------
struct SliceArgs {
    const double *t;
    double *tleft, *tright;
    const double *pL, *pu, *pp;
    const long long *worker_running;
    long long *status;
    double thr, shrink;
    double *allu, *allL, *allp;
    long long popsize;
    int d, nparams;
    unsigned long long *discarded;
    // round-robin worker layout (what :623-628 produces, and arange): worker l serves point
    // running[l % count].  NULL: arbitrary worker_running, every point scans all workers.
    const long long *running;
    long long count;
};

__global__ void k_slice_update(const SliceArgs S)
{
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long npoints = S.running ? S.count : S.popsize;
    long long disc = 0;
    if (r < npoints) {
        const long long w = S.running ? S.running[r] : r;
        const long long first = S.running ? r : 0, step = S.running ? S.count : 1;
        double tl = S.tleft[w], tr = S.tright[w];
        long long st = S.status[w], src = -1;
        for (long long l = first; l < S.popsize; l += step) {
            if (!S.running && S.worker_running[l] != w) continue;
            const double t = S.t[l];
            const double L = S.pL[l];
            if (t > tr || t < tl) {                       // :609-612
                if (L > S.thr) disc++;
                continue;
            }
            if (0 < t && t < tr) tr = __ddiv_rn(t, S.shrink);   // :613-614
            if (0 > t && t > tl) tl = __ddiv_rn(t, S.shrink);   // :615-616
            if (L > S.thr && st == 0) {                   // :617-621
                st = 1;
                src = l;
            }
        }
        S.tleft[w] = tl;
        S.tright[w] = tr;
        S.status[w] = st;
        if (src >= 0) {
            for (int k = 0; k < S.d; k++) S.allu[w * S.d + k] = S.pu[src * S.d + k];
            S.allL[w] = S.pL[src];
            for (int k = 0; k < S.nparams; k++) S.allp[w * S.nparams + k] = S.pp[src * S.nparams + k];
        }
    }
    for (int o = 16; o > 0; o >>= 1) disc += __shfl_xor_sync(FULL, disc, o);
    if ((threadIdx.x & 31) == 0 && disc) atomicAdd(S.discarded, (unsigned long long)disc);
}

// :623-628 -- the workers go round robin to the points still running, in point order.
// One block: ordered compaction of the running points, then worker j takes running[j % count].
__global__ void __launch_bounds__(1024) k_slice_reassign(const long long *__restrict__ status,
                                                         long long popsize,
                                                         long long *__restrict__ running,
                                                         long long *__restrict__ worker_running,
                                                         long long *__restrict__ n_running)
{
    __shared__ int warp_sum[32];
    __shared__ int carry, chunk_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (long long base = 0; base < popsize; base += 1024) {
        const long long k = base + threadIdx.x;
        const bool f = k < popsize && status[k] == 0;
        const unsigned ball = __ballot_sync(FULL, f);
        const int within = __popc(ball & ((1u << lane) - 1));
        if (lane == 0) warp_sum[warp] = __popc(ball);
        __syncthreads();
        if (warp == 0) {
            const int v = warp_sum[lane];
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int other = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += other;
            }
            warp_sum[lane] = incl - v;
            if (lane == 31) chunk_total = incl;
        }
        __syncthreads();
        if (f) running[carry + warp_sum[warp] + within] = k;
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    const int count = carry;
    if (threadIdx.x == 0) *n_running = count;
    if (count == 0) return;   // `while ... and (status == 0).any()` never runs: workers unchanged
    __threadfence_block();
    for (long long j = threadIdx.x; j < popsize; j += 1024) worker_running[j] = running[j % count];
}

// proposals of one pass (popstepsampler.py:941-951) with device transform + likelihood
struct ProposeArgs {
    const double *pos;          // uniform draws of this pass
    const double *tleft, *tright, *allu, *v;
    const long long *worker_running;
    double *t, *pu, *pp, *pL;
    long long popsize;
    int d;
    int xform_kind;
    const double *xform_scale, *xform_lo;
    int loglike_kind;
    const double *lparams;
};

__global__ void k_popslice_propose(const ProposeArgs P)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int d = P.d, ds = sf_odd(d);
    double *pv = rowbuf + (size_t)threadIdx.x * 2 * ds;
    double *tt = pv + ds;
    const long long l = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= P.popsize) return;
    const long long w = P.worker_running[l];
    const double tl = P.tleft[w], tr = P.tright[w];
    // t = tleft_worker + (tright_worker - tleft_worker) * slice_position   (:943)
    const double t = __dadd_rn(tl, __dmul_rn(__dsub_rn(tr, tl), P.pos[l]));
    P.t[l] = t;
    const double *urow = P.allu + w * d, *vrow = P.v + w * d;
    double *pu = P.pu + l * d, *pp = P.pp + l * d;
    for (int k = 0; k < d; k++) {
        const double x = __dadd_rn(urow[k], __dmul_rn(t, vrow[k]));   // points + t * v_worker (:947)
        pu[k] = x;
        pv[k] = x;
    }
    apply_xform(P.xform_kind, P.xform_scale, P.xform_lo, pv, pv, d);
    for (int k = 0; k < d; k++) pp[k] = pv[k];
    P.pL[l] = loglike_row(P.loglike_kind, pv, d, tt, P.lparams);
}

int upload_desc(unb_ctx *ctx, int xform_kind, const double *scale, const double *lo, int loglike_kind,
                const double *lparams, size_t d, cudaStream_t s, const double **scale_dev,
                const double **lo_dev, const double **lp_dev)
{
    if (xform_kind != UNB_XFORM_IDENTITY && xform_kind != UNB_XFORM_SCALE_SHIFT)
        return unb_fail(ctx, UNB_ERR_ARG, "unknown transform kind %d", xform_kind);
    if (loglike_kind != UNB_LOGLIKE_GAUSS && loglike_kind != UNB_LOGLIKE_ROSENBROCK &&
        loglike_kind != UNB_LOGLIKE_EGGBOX)
        return unb_fail(ctx, UNB_ERR_ARG, "the fused step needs a device likelihood");
    if (xform_kind == UNB_XFORM_SCALE_SHIFT && (!scale || !lo))
        return unb_fail(ctx, UNB_ERR_ARG, "null transform parameters");
    if (loglike_kind == UNB_LOGLIKE_GAUSS && !lparams)
        return unb_fail(ctx, UNB_ERR_ARG, "gaussian likelihood needs parameters");
    // [scale d][lo d][lparams d+2]
    std::vector<double> buf(3 * d + 2, 0.0);
    if (xform_kind == UNB_XFORM_SCALE_SHIFT) {
        memcpy(&buf[0], scale, d * sizeof(double));
        memcpy(&buf[d], lo, d * sizeof(double));
    }
    if (loglike_kind == UNB_LOGLIKE_GAUSS) memcpy(&buf[2 * d], lparams, (d + 2) * sizeof(double));
    UNB_TRY(up(ctx, ctx->sf_params, buf.data(), buf.size() * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));   // buf is a temporary
    const double *base = (const double *)ctx->sf_params.p;
    *scale_dev = base;
    *lo_dev = base + d;
    *lp_dev = base + 2 * d;
    return UNB_OK;
}

}  // namespace

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" int unb_within_unit_cube(unb_ctx *ctx, const double *u, size_t n, size_t ndim,
                                    uint8_t *acceptable)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (n == 0) return UNB_OK;
    if (!u || !acceptable || ndim == 0) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = S0(ctx);
    UNB_TRY(up(ctx, ctx->sf[0], u, n * ndim * sizeof(double), s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[1], n));
    k_within_unit_cube<<<blocks_for(n, SF_THREADS), SF_THREADS, 0, s>>>(
        (const double *)ctx->sf[0].p, (long long)n, (int)ndim, (unsigned char *)ctx->sf[1].p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    UNB_TRY(down(ctx, acceptable, ctx->sf[1], n, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

extern "C" int unb_evolve_prepare(unb_ctx *ctx, const uint8_t *searching_left,
                                  const uint8_t *searching_right, size_t n, uint8_t *search_right,
                                  uint8_t *bisecting)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (n == 0) return UNB_OK;
    if (!searching_left || !searching_right || !search_right || !bisecting)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = S0(ctx);
    UNB_TRY(up(ctx, ctx->sf[0], searching_left, n, s));
    UNB_TRY(up(ctx, ctx->sf[1], searching_right, n, s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[2], n));
    UNB_TRY(unb_reserve(ctx, ctx->sf[3], n));
    k_evolve_prepare<<<blocks_for(n, SF_THREADS), SF_THREADS, 0, s>>>(
        (const unsigned char *)ctx->sf[0].p, (const unsigned char *)ctx->sf[1].p, (long long)n,
        (unsigned char *)ctx->sf[2].p, (unsigned char *)ctx->sf[3].p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    UNB_TRY(down(ctx, search_right, ctx->sf[2], n, s));
    UNB_TRY(down(ctx, bisecting, ctx->sf[3], n, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

extern "C" int unb_evolve_update(unb_ctx *ctx, const uint8_t *acceptable, const double *Lnew,
                                 size_t n_lnew, double Lmin, const uint8_t *search_right,
                                 const uint8_t *bisecting, double *currentt, double *current_left,
                                 double *current_right, uint8_t *searching_left,
                                 uint8_t *searching_right, uint8_t *success, size_t n)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (n == 0) return UNB_OK;
    if (!acceptable || (!Lnew && n_lnew) || !search_right || !bisecting || !currentt || !current_left ||
        !current_right || !searching_left || !searching_right || !success)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (n >= (size_t)1 << 30) return unb_fail(ctx, UNB_ERR_ARG, "population too large");
    cudaStream_t s = S0(ctx);
    const size_t nb = n * sizeof(double);
    UNB_TRY(up(ctx, ctx->sf[0], acceptable, n, s));
    UNB_TRY(up(ctx, ctx->sf[1], Lnew, n_lnew * sizeof(double), s));
    UNB_TRY(up(ctx, ctx->sf[2], search_right, n, s));
    UNB_TRY(up(ctx, ctx->sf[3], bisecting, n, s));
    UNB_TRY(up(ctx, ctx->sf[4], currentt, nb, s));
    UNB_TRY(up(ctx, ctx->sf[5], current_left, nb, s));
    UNB_TRY(up(ctx, ctx->sf[6], current_right, nb, s));
    UNB_TRY(up(ctx, ctx->sf[7], searching_left, n, s));
    UNB_TRY(up(ctx, ctx->sf[8], searching_right, n, s));
    UNB_TRY(up(ctx, ctx->sf[9], success, n, s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[10], (n + 2) * sizeof(int)));
    int *rank = (int *)ctx->sf[10].p;
    k_flag_ranks<<<1, 1024, 0, s>>>((const unsigned char *)ctx->sf[0].p, (int)n, rank, rank + n);
    k_evolve_update<<<blocks_for(n, SF_THREADS), SF_THREADS, 0, s>>>(
        (const unsigned char *)ctx->sf[0].p, (const double *)ctx->sf[1].p, rank, (long long)n_lnew, Lmin,
        (const unsigned char *)ctx->sf[2].p, (const unsigned char *)ctx->sf[3].p, (double *)ctx->sf[4].p,
        (double *)ctx->sf[5].p, (double *)ctx->sf[6].p, (unsigned char *)ctx->sf[7].p,
        (unsigned char *)ctx->sf[8].p, (unsigned char *)ctx->sf[9].p, (long long)n);
    ctx->launches += 2;
    UNB_CUDA(ctx, cudaGetLastError());
    UNB_TRY(down(ctx, currentt, ctx->sf[4], nb, s));
    UNB_TRY(down(ctx, current_left, ctx->sf[5], nb, s));
    UNB_TRY(down(ctx, current_right, ctx->sf[6], nb, s));
    UNB_TRY(down(ctx, searching_left, ctx->sf[7], n, s));
    UNB_TRY(down(ctx, searching_right, ctx->sf[8], n, s));
    UNB_TRY(down(ctx, success, ctx->sf[9], n, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

namespace {

// one chunk of walkers through the fused evolve kernel.  One device blob, laid out
//   [v | u t left right sl sr | like acceptable success]
// so that the upload is the first two groups and the download the last two, each ONE copy through
// pinned staging (a step moves nine small arrays; nine pageable copies cost more than the kernel).
int evolve_chunk(unb_ctx *ctx, EvolveArgs A, int threads, double *currentu, const double *currentv,
                 double *currentt, double *current_left, double *current_right,
                 uint8_t *searching_left, uint8_t *searching_right, size_t n, size_t ndim,
                 uint8_t *acceptable, uint8_t *success, double *like, cudaStream_t s)
{
    const size_t nb = n * sizeof(double), rb = n * ndim * sizeof(double);
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t o_v = 0, o_u = o_v + al(rb), o_t = o_u + al(rb), o_l = o_t + al(nb), o_r = o_l + al(nb),
                 o_sl = o_r + al(nb), o_sr = o_sl + al(n), o_like = o_sr + al(n), o_acc = o_like + al(nb),
                 o_succ = o_acc + al(n), total = o_succ + al(n);
    UNB_TRY(unb_reserve(ctx, ctx->sf[0], total));
    UNB_TRY(unb_wait_small(ctx));
    UNB_TRY(unb_reserve_pinned(ctx, ctx->pin_small, total));
    char *dev = (char *)ctx->sf[0].p, *pin = (char *)ctx->pin_small.p;
    memcpy(pin + o_v, currentv, rb);
    memcpy(pin + o_u, currentu, rb);
    memcpy(pin + o_t, currentt, nb);
    memcpy(pin + o_l, current_left, nb);
    memcpy(pin + o_r, current_right, nb);
    memcpy(pin + o_sl, searching_left, n);
    memcpy(pin + o_sr, searching_right, n);
    UNB_CUDA(ctx, cudaMemcpyAsync(dev, pin, o_like, cudaMemcpyHostToDevice, s));
    ctx->h2d_bytes += (long long)o_like;
    A.v = (const double *)(dev + o_v);
    A.u = (double *)(dev + o_u);
    A.t = (double *)(dev + o_t);
    A.left = (double *)(dev + o_l);
    A.right = (double *)(dev + o_r);
    A.sl = (unsigned char *)(dev + o_sl);
    A.sr = (unsigned char *)(dev + o_sr);
    A.like = (double *)(dev + o_like);
    A.acceptable = (unsigned char *)(dev + o_acc);
    A.success = (unsigned char *)(dev + o_succ);
    A.n = (long long)n;
    const size_t smem = (size_t)threads * 2 * sf_odd((int)ndim) * sizeof(double);
    UNB_TRY(allow_smem(ctx, k_evolve_fused, smem));
    k_evolve_fused<<<blocks_for(n, threads), threads, smem, s>>>(A);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    UNB_CUDA(ctx, cudaMemcpyAsync(pin + o_u, dev + o_u, total - o_u, cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += (long long)(total - o_u);
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    memcpy(currentu, pin + o_u, rb);
    memcpy(currentt, pin + o_t, nb);
    memcpy(current_left, pin + o_l, nb);
    memcpy(current_right, pin + o_r, nb);
    memcpy(searching_left, pin + o_sl, n);
    memcpy(searching_right, pin + o_sr, n);
    memcpy(like, pin + o_like, nb);
    memcpy(acceptable, pin + o_acc, n);
    memcpy(success, pin + o_succ, n);
    return UNB_OK;
}

}  // namespace

extern "C" int unb_evolve(unb_ctx *ctx, const unb_step_desc *desc, double Lmin, double *currentu,
                          const double *currentv, double *currentt, double *current_left,
                          double *current_right, uint8_t *searching_left, uint8_t *searching_right,
                          size_t n, size_t ndim, uint8_t *acceptable, uint8_t *success, double *like)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (n == 0) return UNB_OK;
    if (!desc || !currentu || !currentv || !currentt || !current_left || !current_right ||
        !searching_left || !searching_right || !acceptable || !success || !like || ndim == 0)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    const int threads = row_threads_for(2 * sf_odd((int)ndim));
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%zu too large for the row kernels", ndim);
    cudaStream_t s = S0(ctx);
    EvolveArgs A;
    memset(&A, 0, sizeof(A));
    UNB_TRY(upload_desc(ctx, desc->xform_kind, desc->xform_scale, desc->xform_lo, desc->loglike_kind,
                        desc->lparams, ndim, s, &A.xform_scale, &A.xform_lo, &A.lparams));
    A.d = (int)ndim;
    A.xform_kind = desc->xform_kind;
    A.loglike_kind = desc->loglike_kind;
    A.Lmin = Lmin;
    // walkers are independent: large populations go through in chunks that keep the pinned staging
    // blob around 32 MB (UNB_OPT_CHUNK_ROWS overrides the chunk length, for tests)
    size_t chunk = (32u << 20) / ((2 * ndim + 4) * sizeof(double) + 4);
    if (ctx->chunk_rows > 0) chunk = (size_t)ctx->chunk_rows;
    if (chunk < 1) chunk = 1;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t cn = n - off < chunk ? n - off : chunk;
        UNB_TRY(evolve_chunk(ctx, A, threads, currentu + off * ndim, currentv + off * ndim, currentt + off,
                             current_left + off, current_right + off, searching_left + off,
                             searching_right + off, cn, ndim, acceptable + off, success + off,
                             like + off, s));
    }
    return UNB_OK;
}

extern "C" int unb_step_back(unb_ctx *ctx, double Lmin, double *allL, size_t nwalkers, size_t ncols,
                             int64_t *generation, double *currentt)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (nwalkers == 0 || ncols == 0) return UNB_OK;
    if (!allL || !generation || !currentt) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (ncols > (size_t)STEP_BACK_MAXCOLS)
        return unb_fail(ctx, UNB_ERR_ARG, "step_back supports chains of up to %d generations",
                        STEP_BACK_MAXCOLS);
    cudaStream_t s = S0(ctx);
    UNB_TRY(up(ctx, ctx->sf[0], allL, nwalkers * ncols * sizeof(double), s));
    UNB_TRY(up(ctx, ctx->sf[1], generation, nwalkers * sizeof(int64_t), s));
    UNB_TRY(up(ctx, ctx->sf[2], currentt, nwalkers * sizeof(double), s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[3], sizeof(long long)));
    const long long lowest = LLONG_MIN;
    UNB_CUDA(ctx, cudaMemcpyAsync(ctx->sf[3].p, &lowest, sizeof(lowest), cudaMemcpyHostToDevice, s));
    unsigned nblk = blocks_for(nwalkers, 256);
    if (nblk > 1024) nblk = 1024;
    k_i64_max<<<nblk, 256, 0, s>>>((const long long *)ctx->sf[1].p, (long long)nwalkers,
                                   (long long *)ctx->sf[3].p);
    k_step_back<<<blocks_for(nwalkers, 64), 64, 0, s>>>(
        Lmin, (double *)ctx->sf[0].p, (long long)nwalkers, (long long)ncols, (long long *)ctx->sf[1].p,
        (double *)ctx->sf[2].p, (const long long *)ctx->sf[3].p);
    ctx->launches += 2;
    UNB_CUDA(ctx, cudaGetLastError());
    UNB_TRY(down(ctx, allL, ctx->sf[0], nwalkers * ncols * sizeof(double), s));
    UNB_TRY(down(ctx, generation, ctx->sf[1], nwalkers * sizeof(int64_t), s));
    UNB_TRY(down(ctx, currentt, ctx->sf[2], nwalkers * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

namespace {

int check_workers(unb_ctx *ctx, const int64_t *worker_running, size_t popsize)
{
    for (size_t l = 0; l < popsize; l++)
        if (worker_running[l] < 0 || (size_t)worker_running[l] >= popsize)
            return unb_fail(ctx, UNB_ERR_ARG, "worker_running[%zu]=%lld outside the population", l,
                            (long long)worker_running[l]);
    return UNB_OK;
}

int launch_slice_update(unb_ctx *ctx, const SliceArgs &S, long long *running, long long *worker_running,
                        long long *n_running, cudaStream_t s)
{
    UNB_CUDA(ctx, cudaMemsetAsync(S.discarded, 0, sizeof(unsigned long long), s));
    k_slice_update<<<blocks_for((size_t)(S.running ? S.count : S.popsize), SF_THREADS), SF_THREADS, 0, s>>>(S);
    k_slice_reassign<<<1, 1024, 0, s>>>(S.status, S.popsize, running, worker_running, n_running);
    ctx->launches += 2;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

}  // namespace

extern "C" int unb_update_vectorised_slice_sampler(
    unb_ctx *ctx, const double *t, double *tleft, double *tright, const double *proposed_L,
    const double *proposed_u, const double *proposed_p, int64_t *worker_running, int64_t *status,
    double likelihood_threshold, double shrink_factor, double *allu, double *allL, double *allp,
    size_t popsize, size_t ndim, size_t nparams, int64_t *discarded)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;   // the scratch buffers are shared with the slice-loop session
    if (discarded) *discarded = 0;
    if (popsize == 0) return UNB_OK;
    if (!t || !tleft || !tright || !proposed_L || !proposed_u || !proposed_p || !worker_running ||
        !status || !allu || !allL || !allp || !discarded)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    UNB_TRY(check_workers(ctx, worker_running, popsize));
    cudaStream_t s = S0(ctx);
    const size_t nb = popsize * sizeof(double), ib = popsize * sizeof(int64_t);
    UNB_TRY(up(ctx, ctx->sf[0], t, nb, s));
    UNB_TRY(up(ctx, ctx->sf[1], tleft, nb, s));
    UNB_TRY(up(ctx, ctx->sf[2], tright, nb, s));
    UNB_TRY(up(ctx, ctx->sf[3], proposed_L, nb, s));
    UNB_TRY(up(ctx, ctx->sf[4], proposed_u, nb * ndim, s));
    UNB_TRY(up(ctx, ctx->sf[5], proposed_p, nb * nparams, s));
    UNB_TRY(up(ctx, ctx->sf[6], worker_running, ib, s));
    UNB_TRY(up(ctx, ctx->sf[7], status, ib, s));
    UNB_TRY(up(ctx, ctx->sf[8], allu, nb * ndim, s));
    UNB_TRY(up(ctx, ctx->sf[9], allL, nb, s));
    UNB_TRY(up(ctx, ctx->sf[10], allp, nb * nparams, s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[11], ib + 2 * sizeof(long long)));
    SliceArgs S;
    S.t = (const double *)ctx->sf[0].p;
    S.tleft = (double *)ctx->sf[1].p;
    S.tright = (double *)ctx->sf[2].p;
    S.pL = (const double *)ctx->sf[3].p;
    S.pu = (const double *)ctx->sf[4].p;
    S.pp = (const double *)ctx->sf[5].p;
    S.worker_running = (const long long *)ctx->sf[6].p;
    S.status = (long long *)ctx->sf[7].p;
    S.thr = likelihood_threshold;
    S.shrink = shrink_factor;
    S.allu = (double *)ctx->sf[8].p;
    S.allL = (double *)ctx->sf[9].p;
    S.allp = (double *)ctx->sf[10].p;
    S.popsize = (long long)popsize;
    S.d = (int)ndim;
    S.nparams = (int)nparams;
    long long *running = (long long *)ctx->sf[11].p;
    long long *counters = running + popsize;   // [0] n_running, [1] discarded
    S.discarded = (unsigned long long *)(counters + 1);
    // The reference only ever produces round-robin layouts (arange, then :623-628): the first
    // `count` workers serve distinct points in ascending order and worker l repeats worker
    // l % count.  Then point running[r] is served by the workers r, r + count, ... -- O(popsize)
    // instead of every point scanning every worker.
    size_t count = 1;
    while (count < popsize && worker_running[count] > worker_running[count - 1]) count++;
    bool round_robin = true;
    for (size_t l = count; l < popsize && round_robin; l++)
        round_robin = worker_running[l] == worker_running[l - count];
    S.running = nullptr;
    S.count = 0;
    if (round_robin) {
        UNB_CUDA(ctx, cudaMemcpyAsync(running, worker_running, count * sizeof(int64_t),
                                      cudaMemcpyHostToDevice, s));
        S.running = running;
        S.count = (long long)count;
    }
    UNB_TRY(launch_slice_update(ctx, S, running, (long long *)ctx->sf[6].p, counters, s));
    UNB_TRY(down(ctx, tleft, ctx->sf[1], nb, s));
    UNB_TRY(down(ctx, tright, ctx->sf[2], nb, s));
    UNB_TRY(down(ctx, worker_running, ctx->sf[6], ib, s));
    UNB_TRY(down(ctx, status, ctx->sf[7], ib, s));
    UNB_TRY(down(ctx, allu, ctx->sf[8], nb * ndim, s));
    UNB_TRY(down(ctx, allL, ctx->sf[9], nb, s));
    UNB_TRY(down(ctx, allp, ctx->sf[10], nb * nparams, s));
    long long host_counters[2] = {0, 0};
    UNB_CUDA(ctx, cudaMemcpyAsync(host_counters, counters, sizeof(host_counters), cudaMemcpyDeviceToHost, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    *discarded = host_counters[1];
    return UNB_OK;
}

// ---- device-resident slice loop (popstepsampler.py:916-965) ----------------------------------
// buffers: sf[0] allu, [1] allL, [2] allp, [3] v, [4] tleft, [5] tright, [6] worker_running,
//          [7] status, [8] pos, [9] t, [10] pu, [11] pp, [12] pL, [13] running + counters
extern "C" int unb_popslice_begin(unb_ctx *ctx, const unb_step_desc *desc, const double *allu,
                                  const double *allL, const double *v, const double *tleft,
                                  const double *tright, size_t popsize, size_t ndim,
                                  double likelihood_threshold, double shrink_factor)
{
    UNB_TRY(check(ctx));
    ctx->ps_active = false;
    if (!desc || !allu || !allL || !v || !tleft || !tright || popsize == 0 || ndim == 0)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer or empty population");
    if (row_threads_for(2 * sf_odd((int)ndim)) < 32)
        return unb_fail(ctx, UNB_ERR_ARG, "ndim=%zu too large for the row kernels", ndim);
    cudaStream_t s = S0(ctx);
    UNB_TRY(upload_desc(ctx, desc->xform_kind, desc->xform_scale, desc->xform_lo, desc->loglike_kind,
                        desc->lparams, ndim, s, &ctx->ps_scale, &ctx->ps_lo, &ctx->ps_lparams));
    const size_t nb = popsize * sizeof(double), rb = nb * ndim, ib = popsize * sizeof(int64_t);
    UNB_TRY(up(ctx, ctx->sf[0], allu, rb, s));
    UNB_TRY(up(ctx, ctx->sf[1], allL, nb, s));
    UNB_TRY(up(ctx, ctx->sf[3], v, rb, s));
    UNB_TRY(up(ctx, ctx->sf[4], tleft, nb, s));
    UNB_TRY(up(ctx, ctx->sf[5], tright, nb, s));
    // allp = NaN, worker_running = arange, status = 0   (popstepsampler.py:908, 932-934)
    std::vector<double> nanrows(popsize * ndim, std::nan(""));
    std::vector<int64_t> ar(popsize);
    for (size_t i = 0; i < popsize; i++) ar[i] = (int64_t)i;
    UNB_TRY(up(ctx, ctx->sf[2], nanrows.data(), rb, s));
    UNB_TRY(up(ctx, ctx->sf[6], ar.data(), ib, s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[7], ib));
    UNB_CUDA(ctx, cudaMemsetAsync(ctx->sf[7].p, 0, ib, s));
    UNB_TRY(unb_reserve(ctx, ctx->sf[8], nb));
    UNB_TRY(unb_reserve(ctx, ctx->sf[9], nb));
    UNB_TRY(unb_reserve(ctx, ctx->sf[10], rb));
    UNB_TRY(unb_reserve(ctx, ctx->sf[11], rb));
    UNB_TRY(unb_reserve(ctx, ctx->sf[12], nb));
    UNB_TRY(unb_reserve(ctx, ctx->sf[13], ib + 2 * sizeof(long long)));
    UNB_CUDA(ctx, cudaMemcpyAsync(ctx->sf[13].p, ar.data(), ib, cudaMemcpyHostToDevice, s));
    ctx->ps_count = popsize;   // workers start on their own point: running = arange
    UNB_CUDA(ctx, cudaStreamSynchronize(s));   // the staging vectors are temporaries
    ctx->ps_popsize = popsize;
    ctx->ps_ndim = ndim;
    ctx->ps_xform_kind = desc->xform_kind;
    ctx->ps_loglike_kind = desc->loglike_kind;
    ctx->ps_thr = likelihood_threshold;
    ctx->ps_shrink = shrink_factor;
    ctx->ps_active = true;
    return UNB_OK;
}

extern "C" int unb_popslice_iterate(unb_ctx *ctx, const double *slice_position, int64_t *n_running,
                                    int64_t *discarded)
{
    UNB_TRY(check(ctx));
    if (!ctx->ps_active) return unb_fail(ctx, UNB_ERR_STATE, "unb_popslice_begin has not been called");
    if (!slice_position || !n_running || !discarded) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = S0(ctx);
    const size_t popsize = ctx->ps_popsize, ndim = ctx->ps_ndim;
    // the draws go through pinned staging so the copy is one asynchronous DMA
    UNB_TRY(unb_wait_small(ctx));
    UNB_TRY(unb_reserve_pinned(ctx, ctx->pin_small, popsize * sizeof(double)));
    memcpy(ctx->pin_small.p, slice_position, popsize * sizeof(double));
    UNB_TRY(up(ctx, ctx->sf[8], ctx->pin_small.p, popsize * sizeof(double), s));
    ProposeArgs P;
    P.pos = (const double *)ctx->sf[8].p;
    P.tleft = (const double *)ctx->sf[4].p;
    P.tright = (const double *)ctx->sf[5].p;
    P.allu = (const double *)ctx->sf[0].p;
    P.v = (const double *)ctx->sf[3].p;
    P.worker_running = (const long long *)ctx->sf[6].p;
    P.t = (double *)ctx->sf[9].p;
    P.pu = (double *)ctx->sf[10].p;
    P.pp = (double *)ctx->sf[11].p;
    P.pL = (double *)ctx->sf[12].p;
    P.popsize = (long long)popsize;
    P.d = (int)ndim;
    P.xform_kind = ctx->ps_xform_kind;
    P.xform_scale = ctx->ps_scale;
    P.xform_lo = ctx->ps_lo;
    P.loglike_kind = ctx->ps_loglike_kind;
    P.lparams = ctx->ps_lparams;
    const int threads = row_threads_for(2 * sf_odd((int)ndim));
    const size_t smem = (size_t)threads * 2 * sf_odd((int)ndim) * sizeof(double);
    UNB_TRY(allow_smem(ctx, k_popslice_propose, smem));
    k_popslice_propose<<<blocks_for(popsize, threads), threads, smem, s>>>(P);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    SliceArgs S;
    S.t = P.t;
    S.tleft = (double *)ctx->sf[4].p;
    S.tright = (double *)ctx->sf[5].p;
    S.pL = P.pL;
    S.pu = P.pu;
    S.pp = P.pp;
    S.worker_running = P.worker_running;
    S.status = (long long *)ctx->sf[7].p;
    S.thr = ctx->ps_thr;
    S.shrink = ctx->ps_shrink;
    S.allu = (double *)ctx->sf[0].p;
    S.allL = (double *)ctx->sf[1].p;
    S.allp = (double *)ctx->sf[2].p;
    S.popsize = (long long)popsize;
    S.d = (int)ndim;
    S.nparams = (int)ndim;
    long long *running = (long long *)ctx->sf[13].p;
    long long *counters = running + popsize;
    S.discarded = (unsigned long long *)(counters + 1);
    S.running = running;   // layout left by the previous pass (k_slice_reassign) or by begin
    S.count = (long long)ctx->ps_count;
    UNB_TRY(launch_slice_update(ctx, S, running, (long long *)ctx->sf[6].p, counters, s));
    long long host_counters[2] = {0, 0};
    UNB_CUDA(ctx, cudaMemcpyAsync(host_counters, counters, sizeof(host_counters), cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += sizeof(host_counters);
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    *n_running = host_counters[0];
    *discarded = host_counters[1];
    if (host_counters[0] > 0) ctx->ps_count = (size_t)host_counters[0];   // else: layout unchanged
    return UNB_OK;
}

extern "C" int unb_popslice_end(unb_ctx *ctx, double *allu, double *allp, double *allL, double *tleft,
                                double *tright, int64_t *status)
{
    UNB_TRY(check(ctx));
    if (!ctx->ps_active) return unb_fail(ctx, UNB_ERR_STATE, "unb_popslice_begin has not been called");
    cudaStream_t s = S0(ctx);
    const size_t nb = ctx->ps_popsize * sizeof(double), rb = nb * ctx->ps_ndim;
    if (allu) UNB_TRY(down(ctx, allu, ctx->sf[0], rb, s));
    if (allL) UNB_TRY(down(ctx, allL, ctx->sf[1], nb, s));
    if (allp) UNB_TRY(down(ctx, allp, ctx->sf[2], rb, s));
    if (tleft) UNB_TRY(down(ctx, tleft, ctx->sf[4], nb, s));
    if (tright) UNB_TRY(down(ctx, tright, ctx->sf[5], nb, s));
    if (status) UNB_TRY(down(ctx, status, ctx->sf[7], ctx->ps_popsize * sizeof(int64_t), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}
