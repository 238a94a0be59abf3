// unb_region.cu -- row-wise kernels of the MLFriends region path:
//   * ellipsoid membership  (_inside_ellipsoid, mlfriends.pyx:882-912; einsum order of SURVEY
//     fact 5) fused with the candidate transform and the compaction of the survivors, i.e. the
//     first two stages of MLFriends.inside (mlfriends.pyx:1202-1206),
//   * Scaling/Affine layer transform / untransform (mlfriends.pyx:605-620, 737-752),
//   * vectorised likelihoods (docs/gauss.py:25-27, examples/testeggbox.py:9-11,
//     examples/testrosenbrock.py:10-13) with NumPy's pairwise-summation order,
//   * bootstrap enlargement factor (mlfriends.pyx:1060-1062),
//   * mean pair distance (mlfriends.pyx:229-270).
// One thread owns one row; the row's working vector lives in shared memory with an odd
// double-stride so the 64-bit accesses of a warp are bank-conflict free.  The whole library is
// built with -fmad=false: every multiply-add below is non-fused unless it says fma().
#include "unb_internal.cuh"
#include "unb_loglike.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr size_t ROW_SMEM_BUDGET = 200 * 1024;

__host__ __device__ inline int odd_stride(int d) { return d | 1; }

// threads per block so that threads * odd_stride(d) doubles fit the budget (multiple of 32)
inline int row_threads(int d)
{
    size_t per = (size_t)odd_stride(d) * sizeof(double);
    size_t t = ROW_SMEM_BUDGET / per;
    if (t >= 128) return 128;
    return (int)(t / 32 * 32);
}

template <typename K>
int set_smem(unb_ctx *ctx, K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)bytes));
    return UNB_OK;
}

// region parameters in __constant__ memory, row stride dr = d rounded up to 4, zero padded
constexpr int PREP_MAXD = 32;     // register kernel
constexpr int CONST_MAXD = 32;    // parameters in __constant__ memory (3 x 32^2 doubles = 24 KB)
__constant__ double c_ell_center[CONST_MAXD];
__constant__ double c_ell_invcov[CONST_MAXD * CONST_MAXD];
// inverse covariance FOLDED onto the upper triangle (S_jj = A_jj, S_jk = A_jk + A_kj for j < k,
// 0 below): d^T A d = sum_{j<=k} d_j S_jk d_k -- half the multiply-adds of the ellipsoid filter
__constant__ double c_ell_fold[CONST_MAXD * CONST_MAXD];
__constant__ double c_xf_shift[CONST_MAXD];
__constant__ double c_xf_mat[CONST_MAXD * CONST_MAXD];

// ---------------------------------------------------------------------------------------
// ellipsoid (+ transform + compaction)
// ---------------------------------------------------------------------------------------
__global__ void k_prep(const PrepArgs P)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int d = P.d, ds = odd_stride(d);
    double *my = rowbuf + (size_t)threadIdx.x * ds;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < P.m;
    bool inside = valid;
    if (P.center) {
        double acc = 0.0;
        if (valid) {
            const double *p = P.pts + j * d;
            for (int k = 0; k < d; k++)
                my[k] = __dsub_rn(p[k], __ldg(P.center + k));
            // np.einsum('ij,jk,ik->i'): acc += (d_j * A_jk) * d_k, j outer, k inner (buffered chunks)
            acc = einsum_quadform(my, P.invcov, d);
        }
        inside = valid && (acc <= P.r2);
        if (valid && P.mask) P.mask[j] = inside ? 1 : 0;
    }
    if (P.layer_kind < 0) return;

    // compaction of the survivors (warp-aggregated; order is irrelevant because every result
    // is scattered back through items[])
    const unsigned ball = __ballot_sync(FULL, inside);
    int base = 0;
    if ((threadIdx.x & 31) == 0 && ball) base = atomicAdd(P.n_items, __popc(ball));
    base = __shfl_sync(FULL, base, 0);
    if (!inside) return;
    const int pos = base + __popc(ball & ((1u << (threadIdx.x & 31)) - 1));
    P.items[pos] = (int)j;
    const double *p = P.pts + j * d;
    double *out = P.tcand + (size_t)pos * d;
    if (P.layer_kind == UNB_LAYER_AFFINE) {
        // DEFINED order: x = w - ctr, t_j = fma-chain over k ascending
        for (int k = 0; k < d; k++) my[k] = __dsub_rn(p[k], __ldg(P.shift + k));
        for (int jj = 0; jj < d; jj++) {
            double t = 0.0;
            for (int k = 0; k < d; k++) t = fma(my[k], __ldg(P.mat + (size_t)k * d + jj), t);
            out[jj] = t;
        }
    } else if (P.layer_kind == UNB_LAYER_SCALING) {
        for (int k = 0; k < d; k++)
            out[k] = __ddiv_rn(__dsub_rn(p[k], __ldg(P.shift + k)), __ldg(P.mat + k));
    } else {
        for (int k = 0; k < d; k++) out[k] = p[k];
    }
}

// ---------------------------------------------------------------------------------------
// register variant for d <= 32: the candidate lives in registers and the region's matrices in
// __constant__ memory, so every multiply takes its matrix operand straight from the constant
// bank (no load instruction at all).  Matrices are stored with row stride DR (d rounded up to
// 4) and zero padding; padded terms add +0 and leave the reference's sequence unchanged.
// Optionally fuses the vectorised likelihood of the rows that pass the ellipsoid, so the
// proposals are read from HBM exactly once.
// ---------------------------------------------------------------------------------------

// resident blocks the register kernel is compiled for (register cap 128 / 168 / 255)
constexpr int prep_min_blocks(int DR) { return DR <= 20 ? 5 : (DR <= 24 ? 4 : (DR <= 28 ? 3 : 2)); }

// NumPy's pairwise sum (pw_block, n <= 128) of a register vector: every index is a compile-time
// constant after unrolling, so the terms never leave the register file.
template <int DR>
__device__ __forceinline__ double pw_sum_regs(const double (&t)[DR], int n)
{
    if (DR < 8 || n < 8) {
        double res = 0.0;
#pragma unroll
        for (int i = 0; i < DR; i++)
            if (i < n) res = __dadd_rn(res, t[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int q = 0; q < 8; q++) r[q] = t[q < DR ? q : 0];
    const int n8 = n - (n % 8);
#pragma unroll
    for (int i = 8; i + 8 <= DR; i += 8) {
        if (i < n8) {
#pragma unroll
            for (int q = 0; q < 8; q++) r[q] = __dadd_rn(r[q], t[i + q]);
        }
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = 8; i < DR; i++)
        if (i >= n8 && i < n) res = __dadd_rn(res, t[i]);
    return res;
}

// the built-in likelihoods of loglike_row for a row held in registers (same arithmetic)
template <int DR>
__device__ __forceinline__ double loglike_regs(int kind, const double (&p)[DR], const double *row, int d,
                                               const double *lp)
{
    double t[DR];
    if (kind == UNB_LOGLIKE_GAUSS) {
        const double sigma = __ldg(lp + d), norm_const = __ldg(lp + d + 1);
#pragma unroll
        for (int i = 0; i < DR; i++) {
            const double z = (i < d) ? __ddiv_rn(__dsub_rn(p[i], __ldg(lp + i)), sigma) : 0.0;
            t[i] = __dmul_rn(z, z);
        }
        return __dsub_rn(__dmul_rn(-0.5, pw_sum_regs<DR>(t, d)), norm_const);
    } else if (kind == UNB_LOGLIKE_ROSENBROCK) {
#pragma unroll
        for (int i = 0; i < DR; i++) {
            const double a = p[i], b = p[i + 1 < DR ? i + 1 : i];
            const double u = __dsub_rn(b, __dmul_rn(a, a));
            const double v = __dsub_rn(1.0, a);
            t[i] = __dadd_rn(__dmul_rn(100.0, __dmul_rn(u, u)), __dmul_rn(v, v));
        }
        return __dmul_rn(-2.0, pw_sum_regs<DR>(t, d - 1));
    }
    // eggbox: cos() is a long routine -- a rolled loop over the row in memory (L1 resident)
    // instead of DR inlined copies
    double chi = 1.0;
#pragma unroll 1
    for (int i = 0; i < d; i++) chi = __dmul_rn(chi, cos(__ddiv_rn(row[i], 2.0)));
    return pow(__dadd_rn(2.0, chi), 5.0);
}

template <int DR>
__device__ __forceinline__ void prep_load_row(const PrepArgs &P, long long j, bool vec, double (&p)[DR])
{
    const int d = P.d;
    const bool valid = j < P.m;
    if (vec) {
        // even d: rows are 16-byte aligned, read them as 128-bit loads
        const double2 *row2 = reinterpret_cast<const double2 *>(P.pts + j * d);
#pragma unroll
        for (int k = 0; k < DR; k += 2) {
            double2 v = make_double2(0.0, 0.0);
            if (valid && k < d) v = row2[k >> 1];
            p[k] = v.x;
            p[k + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int k = 0; k < DR; k++) p[k] = (valid && k < d) ? P.pts[j * d + k] : 0.0;
    }
}

template <int DR>
__global__ void __launch_bounds__(128, prep_min_blocks(DR)) k_prep_reg(const PrepArgs P)
{
    const int d = P.d;
    const bool vec = (d & 1) == 0 && (reinterpret_cast<uintptr_t>(P.pts) & 15) == 0;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double p[DR];
    prep_load_row<DR>(P, j, vec, p);
    {
        const bool valid = j < P.m;
        bool inside = valid;
        if (P.center) {
            double dl[DR];
#pragma unroll
            for (int k = 0; k < DR; k++) dl[k] = __dsub_rn(p[k], P.center_arg[k]);
            // filter: r_fast = d^T (A d) with fused multiply-adds on the folded matrix
            // (d(d+1)/2 + 2d DFMA instead of the einsum's 3 d^2 non-fused operations).  Both r_fast and the reference's sequential
            // einsum value lie within (d^2+2d+4) u * sum|d_j A_jk d_k| <= tol of d^T A d, with
            // sum|...| <= |d|^2 ||A||_F, so outside the band [r2 - tol, r2 + tol] the comparison
            // is already decided; inside the band the exact einsum order decides.
            double nd = 0.0, rfast = 0.0;
#pragma unroll
            for (int jj = 0; jj < DR; jj++) {
                double y = 0.0;
#pragma unroll
                for (int k = jj; k < DR; k++) y = fma(c_ell_fold[jj * DR + k], dl[k], y);   // folded: k >= jj
                rfast = fma(dl[jj], y, rfast);
                nd = fma(dl[jj], dl[jj], nd);
            }
            const double tol = __dmul_rn(P.ell_tol_scale, nd);
            bool in = rfast <= P.r2;
            const bool band = !(fabs(__dsub_rn(rfast, P.r2)) > tol);   // also true for NaN
            if (__any_sync(FULL, band && valid)) {
                if (band) {
                    double acc = 0.0;
#pragma unroll
                    for (int jj = 0; jj < DR; jj++)
#pragma unroll
                        for (int k = 0; k < DR; k++)
                            acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dl[jj], c_ell_invcov[jj * DR + k]), dl[k]));
                    in = acc <= P.r2;
                }
            }
            inside = valid && in;
            if (valid && P.mask) P.mask[j] = inside ? 1 : 0;
        }
        if (P.like && valid) {
            double like = -__longlong_as_double(0x7ff0000000000000LL);
            if (inside) like = loglike_regs<DR>(P.loglike_kind, p, P.pts + j * d, d, P.lparams);
            P.like[j] = like;
        }
        if (P.layer_kind >= 0) {
            const unsigned ball = __ballot_sync(FULL, inside);
            int base = 0;
            if ((threadIdx.x & 31) == 0 && ball) base = atomicAdd(P.n_items, __popc(ball));
            base = __shfl_sync(FULL, base, 0);
            if (inside) {
                const int pos = base + __popc(ball & ((1u << (threadIdx.x & 31)) - 1));
                P.items[pos] = (int)j;
                double *out = P.tcand + (size_t)pos * d;
                double o[DR];
                if (P.layer_kind == UNB_LAYER_AFFINE) {
                    double x[DR];
#pragma unroll
                    for (int k = 0; k < DR; k++) x[k] = __dsub_rn(p[k], c_xf_shift[k]);
#pragma unroll
                    for (int jj = 0; jj < DR; jj++) {
                        double t = 0.0;
#pragma unroll
                        for (int k = 0; k < DR; k++) t = fma(x[k], c_xf_mat[k * DR + jj], t);
                        o[jj] = t;
                    }
                } else if (P.layer_kind == UNB_LAYER_SCALING) {
#pragma unroll
                    for (int k = 0; k < DR; k++) o[k] = __ddiv_rn(__dsub_rn(p[k], c_xf_shift[k]), c_xf_mat[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < DR; k++) o[k] = p[k];
                }
                if ((d & 1) == 0 && (reinterpret_cast<uintptr_t>(P.tcand) & 15) == 0) {
                    double2 *out2 = reinterpret_cast<double2 *>(out);   // pos * d is even
#pragma unroll
                    for (int k = 0; k < DR; k += 2)
                        if (k < d) out2[k >> 1] = make_double2(o[k], o[k + 1]);
                } else {
#pragma unroll
                    for (int k = 0; k < DR; k++)
                        if (k < d) out[k] = o[k];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// tile variant for 32 < d (rows no longer fit registers): the block's TILE_PTS proposals sit
// row-major in shared memory (odd stride: a warp's 64-bit reads of one column are conflict
// free) and the two d x d products -- ellipsoid filter  Y = delta A  and layer transform
// t = x T -- run as register-blocked products: lane = 4 proposals (lane + 32 i), warp = chunks
// of 8 matrix columns.  The warp first copies its d x 8 matrix panel (zero-padded copy, 128-bit
// coalesced loads, all in flight together) into its own shared-memory slot, so one reduction
// step is 4 LDS.64 + 4 broadcast LDS.128 for 32 independent DFMA and never waits on L2.
// In the layer transform the reduction index runs ascending in ONE accumulator per output,
// which is exactly its defined order; the ellipsoid product only feeds the
// filter (band decided by the reference's einsum order, like in k_prep_reg), so it is free to
// use the folded upper-triangular matrix -- half the multiply-adds.
// ---------------------------------------------------------------------------------------
constexpr int TILE_PTS = 128;
constexpr int TILE_WARPS = 8;
constexpr int TILE_THREADS = TILE_WARPS * 32;
constexpr size_t TILE_SMEM_BUDGET = 216 * 1024;

__host__ __device__ inline size_t tile_smem_doubles(int d)
{
    return (size_t)TILE_PTS * odd_stride(d) + (size_t)TILE_WARPS * d * 8;
}

__device__ __forceinline__ void tile_product(const double *__restrict__ sm, int ds, int lane,
                                             double *__restrict__ panel,
                                             const double *__restrict__ Mpad, int ja, int jb,
                                             int dp, int c0, double (&acc)[4][8])
{
    // panel[j - ja][0..8) = Mpad[j][c0 .. c0+8),  ja <= j < jb
    {
        const double2 *src = reinterpret_cast<const double2 *>(Mpad + (size_t)ja * dp + c0);
        double2 *dst = reinterpret_cast<double2 *>(panel);
        const int units = (jb - ja) * 4, hstep = dp >> 1;
        __syncwarp();   // the previous chunk's reads of the slot are done
        for (int u = lane; u < units; u += 32) dst[u] = __ldg(src + (size_t)(u >> 2) * hstep + (u & 3));
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[i][c] = 0.0;
    const double *r0 = sm + (size_t)lane * ds;
    const double *r1 = r0 + 32 * ds, *r2 = r0 + 64 * ds, *r3 = r0 + 96 * ds;
    const double2 *prow = reinterpret_cast<const double2 *>(panel);
#pragma unroll 2
    for (int j = ja; j < jb; j++, prow += 4) {
        const double2 a01 = prow[0], a23 = prow[1], a45 = prow[2], a67 = prow[3];
        const double a[8] = {a01.x, a01.y, a23.x, a23.y, a45.x, a45.y, a67.x, a67.y};
        const double x[4] = {r0[j], r1[j], r2[j], r3[j]};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int c = 0; c < 8; c++) acc[i][c] = fma(x[i], a[c], acc[i][c]);
    }
}

// stage  pts[row0 .. row0+TILE_PTS) - shift  into the shared-memory tile (rows >= nvalid: zeros).
// The raw rows travel global -> shared memory as 8-byte cp.async copies (LDGSTS: no register
// staging, every copy of the thread in flight at once); the thread then subtracts the shift from
// exactly the elements it copied, so only its own cp.async group has to be waited for.
__device__ __forceinline__ void tile_stage(double *sm, int ds, const double *__restrict__ pts,
                                           long long row0, int nvalid, int d,
                                           const double *__restrict__ shift)
{
    const double *src = pts + row0 * d;
    const int q = TILE_THREADS / d, r = TILE_THREADS % d;   // a thread strides TILE_THREADS elements
    const int pt0 = threadIdx.x / d, k0 = threadIdx.x % d;
    const int total = TILE_PTS * d, nv = nvalid * d;
    int pt = pt0, k = k0;
    for (int e = threadIdx.x; e < total; e += TILE_THREADS) {
        double *dst = sm + pt * ds + k;
        if (e < nv)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src + e)
                         : "memory");
        else
            *dst = 0.0;
        pt += q;
        k += r;
        if (k >= d) { k -= d; pt++; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    if (shift) {
        pt = pt0;
        k = k0;
        for (int e = threadIdx.x; e < nv; e += TILE_THREADS) {
            double *dst = sm + pt * ds + k;
            *dst = __dsub_rn(*dst, __ldg(shift + k));
            pt += q;
            k += r;
            if (k >= d) { k -= d; pt++; }
        }
    }
}

__global__ void __launch_bounds__(TILE_THREADS) k_prep_tile(const PrepArgs P)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_pos[TILE_PTS];
    const int d = P.d, ds = odd_stride(d), dp = P.pad_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // per-warp matrix panel slots behind the tile (16-byte aligned: even offset)
    double *panels = sm + (((size_t)TILE_PTS * ds + 1) & ~(size_t)1);
    double *panel = panels + (size_t)warp * d * 8;
    const long long row0 = (long long)blockIdx.x * TILE_PTS;
    const long long left = P.m - row0;
    const int nvalid = left < TILE_PTS ? (int)left : TILE_PTS;
    const long long j = row0 + tid;
    const bool valid = tid < nvalid;   // threads >= TILE_PTS own no proposal
    const int nchunks = dp >> 3;
    bool inside = valid;

    if (P.center) {
        tile_stage(sm, ds, P.pts, row0, nvalid, d, P.center);
        __syncthreads();
        // delta^T A delta = sum over j <= c of delta_j S_jc delta_c with S the folded matrix
        // (S_jj = A_jj, S_jc = A_jc + A_cj for j < c, 0 below the diagonal): chunk ch only
        // needs the rows j < 8 ch + 8.  The rows of all chunks, laid end to end, are cut into
        // TILE_WARPS equal segments -- any partition of the (j, c) products sums to the same
        // filter value -- so the warps carry equal work whatever d is.
        double rp[4] = {0.0, 0.0, 0.0, 0.0};
        {
            int total_rows = 0;
            for (int ch = 0; ch < nchunks; ch++) total_rows += min(8 * ch + 8, d);
            const int lo = (int)((long long)total_rows * warp / TILE_WARPS);
            const int hi = (int)((long long)total_rows * (warp + 1) / TILE_WARPS);
            int prefix = 0;
            for (int ch = 0; ch < nchunks && prefix < hi; ch++) {
                const int rows = min(8 * ch + 8, d);
                const int ja = max(lo - prefix, 0), jb = min(hi - prefix, rows);
                prefix += rows;
                if (ja >= jb) continue;
                const int c0 = ch << 3;
                double acc[4][8];
                tile_product(sm, ds, lane, panel, P.invcov_pad, ja, jb, dp, c0, acc);
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (c0 + c < d) {
#pragma unroll
                        for (int i = 0; i < 4; i++)
                            rp[i] = fma(acc[i][c], sm[(size_t)(lane + 32 * i) * ds + c0 + c], rp[i]);
                    }
                }
            }
        }
        // the warp's partial sums go through its own panel slot (d*8 >= TILE_PTS doubles)
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; i++) panel[lane + 32 * i] = rp[i];
        __syncthreads();
        if (tid < TILE_PTS) {
            const double *my = sm + (size_t)tid * ds;
            double rfast = 0.0, nd = 0.0;
#pragma unroll
            for (int w = 0; w < TILE_WARPS; w++) rfast = __dadd_rn(rfast, panels[(size_t)w * d * 8 + tid]);
            for (int k = 0; k < d; k++) nd = fma(my[k], my[k], nd);
            // same band argument as k_prep_reg: any summation order of the d^2 products (fused
            // or not) stays within (d^2+2d+4) u |delta|^2 ||A||_F of delta^T A delta; tol is 2x that
            const double tol = __dmul_rn(P.ell_tol_scale, nd);
            bool in = rfast <= P.r2;
            const bool band = !(fabs(__dsub_rn(rfast, P.r2)) > tol);   // also true for NaN
            if (band && valid) {
                in = einsum_quadform(my, P.invcov, d) <= P.r2;
            }
            inside = valid && in;
            if (valid && P.mask) P.mask[j] = inside ? 1 : 0;
        }
    }

    if (P.layer_kind >= 0) {
        const unsigned ball = __ballot_sync(FULL, inside);
        int base = 0;
        if (lane == 0 && ball) base = atomicAdd(P.n_items, __popc(ball));
        base = __shfl_sync(FULL, base, 0);
        const int pos = inside ? base + __popc(ball & ((1u << lane) - 1)) : -1;
        if (tid < TILE_PTS) s_pos[tid] = pos;
        if (inside) P.items[pos] = (int)j;
        // barrier: s_pos visible, every warp done with the delta tile and the panel slots
        const int nsurv = __syncthreads_count(inside);
        if (nsurv) {
            if (P.layer_kind == UNB_LAYER_AFFINE) {
                tile_stage(sm, ds, P.pts, row0, nvalid, d, P.shift);
                __syncthreads();
                for (int ch = warp; ch < nchunks; ch += TILE_WARPS) {
                    const int c0 = ch << 3;
                    double acc[4][8];
                    tile_product(sm, ds, lane, panel, P.mat_pad, 0, d, dp, c0, acc);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const int pos_i = s_pos[lane + 32 * i];
                        if (pos_i < 0) continue;
                        double *out = P.tcand + (size_t)pos_i * d + c0;
#pragma unroll
                        for (int c = 0; c < 8; c++)
                            if (c0 + c < d) out[c] = acc[i][c];
                    }
                }
            } else {
                const double *src = P.pts + row0 * d;
                const int nv = nvalid * d;
                for (int e = tid; e < nv; e += TILE_THREADS) {
                    const int pt = e / d, k = e - pt * d;
                    const int pos_e = s_pos[pt];
                    if (pos_e < 0) continue;
                    double v = src[e];
                    if (P.layer_kind == UNB_LAYER_SCALING)
                        v = __ddiv_rn(__dsub_rn(v, __ldg(P.shift + k)), __ldg(P.mat + k));
                    P.tcand[(size_t)pos_e * d + k] = v;
                }
            }
        }
    }
    if (P.like) {
        __syncthreads();   // the tile becomes per-thread scratch of the likelihood
        if (valid) {
            double like = -__longlong_as_double(0x7ff0000000000000LL);
            if (inside) like = loglike_row(P.loglike_kind, P.pts + j * d, d, sm + (size_t)tid * ds, P.lparams);
            P.like[j] = like;
        }
    }
}

// ---------------------------------------------------------------------------------------
// layer transforms
// ---------------------------------------------------------------------------------------
__global__ void k_transform(int kind, int inverse, const double *__restrict__ in, long long m,
                            int d, const double *__restrict__ shift,
                            const double *__restrict__ mat, double *__restrict__ out)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int ds = odd_stride(d);
    double *my = rowbuf + (size_t)threadIdx.x * ds;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double *p = in + j * d;
    double *o = out + j * d;
    if (kind == UNB_LAYER_AFFINE) {
        if (!inverse) {
            for (int k = 0; k < d; k++) my[k] = __dsub_rn(p[k], __ldg(shift + k));
            for (int jj = 0; jj < d; jj++) {
                double t = 0.0;
                for (int k = 0; k < d; k++) t = fma(my[k], __ldg(mat + (size_t)k * d + jj), t);
                o[jj] = t;
            }
        } else {
            for (int k = 0; k < d; k++) my[k] = p[k];
            for (int jj = 0; jj < d; jj++) {
                double t = 0.0;
                for (int k = 0; k < d; k++) t = fma(my[k], __ldg(mat + (size_t)k * d + jj), t);
                o[jj] = __dadd_rn(t, __ldg(shift + jj));
            }
        }
    } else if (kind == UNB_LAYER_SCALING) {
        if (!inverse)
            for (int k = 0; k < d; k++)
                o[k] = __ddiv_rn(__dsub_rn(p[k], __ldg(shift + k)), __ldg(mat + k));
        else
            for (int k = 0; k < d; k++)
                o[k] = __dadd_rn(__dmul_rn(p[k], __ldg(mat + k)), __ldg(shift + k));
    } else {
        for (int k = 0; k < d; k++) o[k] = p[k];
    }
}

// ---------------------------------------------------------------------------------------
// vectorised likelihoods.  lp = device parameter block:
//   GAUSS: centers[d], sigma, norm_const
// mask (nullable): rows with mask == 0 get -inf (integrator.py:1797-1802 only evaluates accepted
// rows and leaves the rest at -inf).
// ---------------------------------------------------------------------------------------
__global__ void k_loglike(int kind, const double *__restrict__ params, int d, long long n,
                          double *__restrict__ like, const unsigned char *__restrict__ mask,
                          const double *__restrict__ lp)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int ds = odd_stride(d);
    double *t = rowbuf + (size_t)threadIdx.x * ds;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (mask && !mask[j]) {
        like[j] = -__longlong_as_double(0x7ff0000000000000LL);
        return;
    }
    like[j] = loglike_row(kind, params + j * d, d, t, lp);
}

// ---------------------------------------------------------------------------------------
// tail of the fused refill: for every region member  v = transform(u)  ->  tregion.inside(v)
// (einsum order of _inside_ellipsoid)  ->  loglike(v)  ->  logl > Lmin.  Per thread two
// shared-memory rows: v and the likelihood's term vector.
// ---------------------------------------------------------------------------------------
__global__ void k_refill_tail(const TailArgs T)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int d = T.d, ds = odd_stride(d);
    double *v = rowbuf + (size_t)threadIdx.x * 2 * ds;
    double *t = v + ds;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = j < T.m;
    bool member = false, tpass = false, accepted = false;
    if (valid) {
        const double *p = T.pts + j * d;
        member = T.have_mask ? (T.flags[j] != 0) : true;
        if (member && T.check_cube)
            for (int k = 0; k < d; k++) member = member && (p[k] > 0.0) && (p[k] < 1.0);
        double like = -__longlong_as_double(0x7ff0000000000000LL);
        if (member) {
            if (T.xform_kind == UNB_XFORM_SCALE_SHIFT)
                for (int k = 0; k < d; k++)
                    v[k] = __dadd_rn(__dmul_rn(p[k], __ldg(T.xform_scale + k)), __ldg(T.xform_lo + k));
            else
                for (int k = 0; k < d; k++) v[k] = p[k];
            tpass = true;
            if (T.treg_center) {
                for (int k = 0; k < d; k++) t[k] = __dsub_rn(v[k], __ldg(T.treg_center + k));
                tpass = einsum_quadform(t, T.treg_invcov, d) <= T.treg_r2;
            }
            if (tpass) {
                like = loglike_row(T.loglike_kind, v, d, t, T.lparams);
                accepted = like > T.Lmin;
            }
        }
        T.like[j] = like;
        T.flags[j] = (unsigned char)((member ? UNB_REFILL_MEMBER : 0) | (tpass ? UNB_REFILL_TREGION : 0) |
                                     (accepted ? UNB_REFILL_ACCEPTED : 0));
    }
    const unsigned bm = __ballot_sync(FULL, member), bt = __ballot_sync(FULL, tpass),
                   ba = __ballot_sync(FULL, accepted);
    if ((threadIdx.x & 31) == 0) {
        if (bm) atomicAdd(T.counts + 0, __popc(bm));
        if (bt) atomicAdd(T.counts + 1, __popc(bt));
        if (ba) atomicAdd(T.counts + 2, __popc(ba));
    }
}

// ---------------------------------------------------------------------------------------
// bootstrap enlargement: per round, max over the left-out rows of the einsum form
// ---------------------------------------------------------------------------------------
__global__ void k_enlargement_f(const double *__restrict__ u, int d,
                                const int *__restrict__ item_idx,
                                const int *__restrict__ round_item_off,
                                const int *__restrict__ round_nitems,
                                const double *__restrict__ ctrs,
                                const double *__restrict__ invcovs,
                                unsigned long long *__restrict__ out_round_key)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int ds = odd_stride(d);
    double *my = rowbuf + (size_t)threadIdx.x * ds;
    const int round = blockIdx.y;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = item < round_nitems[round];
    unsigned long long key = 0;   // below every finite double's key
    if (valid) {
        const int row = item_idx[round_item_off[round] + item];
        const double *p = u + (size_t)row * d;
        const double *ctr = ctrs + (size_t)round * d;
        const double *A = invcovs + (size_t)round * d * d;
        for (int k = 0; k < d; k++) my[k] = __dsub_rn(p[k], __ldg(ctr + k));
        key = f64_key(einsum_quadform(my, A, d));
    }
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(FULL, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(out_round_key + round, key);
}

// ---------------------------------------------------------------------------------------
// mean pair distance: thread j accumulates sqrt(D_ij) over i < j in the same cluster
// ---------------------------------------------------------------------------------------
__global__ void k_pairdist(const double *__restrict__ pts, const long long *__restrict__ ids,
                           int n, int d, double *__restrict__ partial_sum,
                           long long *__restrict__ partial_cnt)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double total = 0.0;
    long long cnt = 0;
    const long long cj = ids[j];
    if (cj != 0) {
        const double *b = pts + (size_t)j * d;
        for (int i = 0; i < j; i++) {
            if (ids[i] != cj) continue;
            const double *a = pts + (size_t)i * d;
            double D = 0.0;
            for (int k = 0; k < d; k++) D = sq_step(D, a[k], b[k]);
            total = __dadd_rn(total, sqrt(D));
            cnt++;
        }
    }
    partial_sum[j] = total;
    partial_cnt[j] = cnt;
}

// ---------------------------------------------------------------------------------------
// fp64 pipe peak probe: 8 independent DFMA chains per thread, no memory traffic.  Gives the
// measured denominator for the compute roofline of the scan kernels on THIS device and clock.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;   // keeps the chains alive
}

__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float seed)
{
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    float a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float m = 1.0000001f, c = 1e-7f;
    for (int i = 0; i < iters; i++) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
    }
    const float r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678f) out[0] = r;
}

// packed form: 8 independent FFMA2 chains (16 lane-FMAs per iteration and thread)
template <bool BCAST>
__global__ void __launch_bounds__(256) k_fp32x2_peak(float *out, int iters, float seed)
{
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed + threadIdx.x + i, seed - i);
    const float ms = 1.0000001f + 1e-7f * (float)(threadIdx.x & 1);
    const float2 m = BCAST ? make_float2(ms, ms) : make_float2(ms, 1.0000002f);
    const float2 c = make_float2(1e-7f, 2e-7f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = BCAST ? ffma2(m, a[i], c) : ffma2(a[i], m, c);
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) r += a[i].x + a[i].y;
    if (r == 12345.678f) out[0] = r;
}

}  // namespace

int unb_launch_fp32_peak(unb_ctx *ctx, float *scratch, int blocks, int iters, cudaStream_t s)
{
    if (iters < 0) {   // packed forms: -iters iterations; blocks < 0 selects the broadcast form
        if (blocks < 0) k_fp32x2_peak<true><<<-blocks, 256, 0, s>>>(scratch, -iters, 1.0f);
        else k_fp32x2_peak<false><<<blocks, 256, 0, s>>>(scratch, -iters, 1.0f);
        ctx->launches++;
        UNB_CUDA(ctx, cudaGetLastError());
        return UNB_OK;
    }
    k_fp32_peak<<<blocks, 256, 0, s>>>(scratch, iters, 1.0f);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_fp64_peak(unb_ctx *ctx, double *scratch, int blocks, int iters, cudaStream_t s)
{
    k_fp64_peak<<<blocks, 256, 0, s>>>(scratch, iters, 1.0);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

size_t unb_const_maxd() { return CONST_MAXD; }


// the tile prep kernel serves 32 < d as long as its 128-row tile fits shared memory
bool unb_tile_prep_fits(int d)
{
    return d > PREP_MAXD && tile_smem_doubles(d) * sizeof(double) <= TILE_SMEM_BUDGET;
}

size_t unb_max_rowwise_d() { return ROW_SMEM_BUDGET / sizeof(double) / 32 - 1; }

// -- __constant__ parameter block of the register prep kernel -------------------------------
static const unb_ctx *g_const_owner = nullptr;
static long long g_const_version = -1;

// a destroyed context must not stay the recorded owner of the constant block
void unb_prep_forget_ctx(const unb_ctx *ctx)
{
    if (g_const_owner == ctx) {
        g_const_owner = nullptr;
        g_const_version = -1;
    }
}

// what the constant block currently holds (zero-padded images), so that a call re-sends only the
// arrays that changed: the integrator moves the ellipsoid centre every iteration
// (integrator.py:2756) but the matrices only at a region rebuild
static std::vector<double> g_img_center, g_img_invcov, g_img_shift, g_img_mat;

static std::vector<double> padded_image(const std::vector<double> &src, size_t rows, size_t cols, size_t stride)
{
    std::vector<double> buf(rows * stride, 0.0);
    for (size_t r = 0; r < rows; r++)
        for (size_t c = 0; c < cols; c++) buf[r * stride + c] = src[r * cols + c];
    return buf;
}

static bool same_image(const std::vector<double> &a, const std::vector<double> &b)
{
    return a.size() == b.size() && (a.empty() || memcmp(a.data(), b.data(), a.size() * sizeof(double)) == 0);
}

static int send_image(unb_ctx *ctx, const void *symbol, const std::vector<double> &img, cudaStream_t s)
{
    UNB_CUDA(ctx, cudaMemcpyToSymbolAsync(symbol, img.data(), img.size() * sizeof(double), 0,
                                          cudaMemcpyHostToDevice, s));
    return UNB_OK;
}

// (re)load the region's ellipsoid / layer parameters into constant memory if they changed
int unb_prep_sync_constants(unb_ctx *ctx, cudaStream_t s)
{
    RegionState &R = ctx->region;
    if (g_const_owner == ctx && g_const_version == R.param_version) return UNB_OK;
    const size_t d = R.have_ellipsoid ? R.ell_d : R.live.d;
    if (d > (size_t)CONST_MAXD) return UNB_OK;
    const size_t dr = (d + 3) / 4 * 4;
    if (g_const_owner != ctx) {
        g_img_center.clear(); g_img_invcov.clear(); g_img_shift.clear(); g_img_mat.clear();
    }
    // (the ellipsoid centre is a kernel argument of k_prep_reg, PrepArgs::center_arg)
    std::vector<double> center, invcov, shift, mat;
    if (R.have_ellipsoid && R.ell_d == d) invcov = padded_image(R.ell_invcov_h, d, d, dr);
    if (R.layer_kind == UNB_LAYER_AFFINE && R.layer_d == d) {
        shift = padded_image(R.layer_shift_h, 1, d, dr);
        mat = padded_image(R.layer_mat_h, d, d, dr);
    } else if (R.layer_kind == UNB_LAYER_SCALING && R.layer_d == d) {
        shift = padded_image(R.layer_shift_h, 1, d, dr);
        mat.assign(dr, 1.0);
        for (size_t k = 0; k < d; k++) mat[k] = R.layer_mat_h[k];
    }
    const bool new_center = !center.empty() && !same_image(center, g_img_center);
    const bool new_invcov = !invcov.empty() && !same_image(invcov, g_img_invcov);
    const bool new_shift = !shift.empty() && !same_image(shift, g_img_shift);
    const bool new_mat = !mat.empty() && !same_image(mat, g_img_mat);
    if (new_center || new_invcov || new_shift || new_mat) {
        // the constant block is process-global: kernels of ANOTHER context may still read the old
        // values (the previous owner's streams are not ours to name, so wait for the device)
        if (g_const_owner != nullptr && g_const_owner != ctx) UNB_CUDA(ctx, cudaDeviceSynchronize());
        // kernels of either lane may still read the old values
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        std::vector<double> fold;
        if (new_center) UNB_TRY(send_image(ctx, c_ell_center, center, s));
        if (new_invcov) {
            UNB_TRY(send_image(ctx, c_ell_invcov, invcov, s));
            fold.assign(d * dr, 0.0);
            for (size_t r = 0; r < d; r++) {
                fold[r * dr + r] = R.ell_invcov_h[r * d + r];
                for (size_t c = r + 1; c < d; c++)
                    fold[r * dr + c] = R.ell_invcov_h[r * d + c] + R.ell_invcov_h[c * d + r];
            }
            UNB_TRY(send_image(ctx, c_ell_fold, fold, s));
        }
        if (new_shift) UNB_TRY(send_image(ctx, c_xf_shift, shift, s));
        if (new_mat) UNB_TRY(send_image(ctx, c_xf_mat, mat, s));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));   // the images are read from pageable memory
        if (new_center) g_img_center.swap(center);
        if (new_invcov) g_img_invcov.swap(invcov);
        if (new_shift) g_img_shift.swap(shift);
        if (new_mat) g_img_mat.swap(mat);
    }
    g_const_owner = ctx;
    g_const_version = R.param_version;
    return UNB_OK;
}

template <int DR>
static int launch_prep_reg(unb_ctx *ctx, const PrepArgs &p, cudaStream_t s)
{
    k_prep_reg<DR><<<(unsigned)((p.m + 127) / 128), 128, 0, s>>>(p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_prep(unb_ctx *ctx, const PrepArgs &p, cudaStream_t s)
{
    if (p.m <= 0) return UNB_OK;
    if (p.use_constants && p.d > CONST_MAXD)
        return unb_fail(ctx, UNB_ERR_ARG, "constant-memory prep needs ndim <= %d", CONST_MAXD);
    if (p.use_constants && p.d <= PREP_MAXD) {
        switch ((p.d + 3) / 4 * 4) {
        case 4: return launch_prep_reg<4>(ctx, p, s);
        case 8: return launch_prep_reg<8>(ctx, p, s);
        case 12: return launch_prep_reg<12>(ctx, p, s);
        case 16: return launch_prep_reg<16>(ctx, p, s);
        case 20: return launch_prep_reg<20>(ctx, p, s);
        case 24: return launch_prep_reg<24>(ctx, p, s);
        case 28: return launch_prep_reg<28>(ctx, p, s);
        case 32: return launch_prep_reg<32>(ctx, p, s);
        default: break;
        }
    }
    if (p.pad_stride > 0 && unb_tile_prep_fits(p.d)) {
        if ((p.center && !p.invcov_pad) || (p.layer_kind == UNB_LAYER_AFFINE && !p.mat_pad) ||
            p.pad_stride % 8 != 0 || p.pad_stride < p.d)
            return unb_fail(ctx, UNB_ERR_ARG, "tile prep kernel needs the padded matrices");
        const size_t smem = (tile_smem_doubles(p.d) + 2) * sizeof(double);
        UNB_CUDA(ctx, cudaFuncSetAttribute(k_prep_tile, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(TILE_SMEM_BUDGET + 1024)));
        k_prep_tile<<<(unsigned)((p.m + TILE_PTS - 1) / TILE_PTS), TILE_THREADS, smem, s>>>(p);
        ctx->launches++;
        UNB_CUDA(ctx, cudaGetLastError());
        return UNB_OK;
    }
    const int threads = row_threads(p.d);
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the row kernels", p.d);
    const size_t smem = (size_t)threads * odd_stride(p.d) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_prep, smem));
    k_prep<<<(unsigned)((p.m + threads - 1) / threads), threads, smem, s>>>(p);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_transform(unb_ctx *ctx, int kind, bool inverse, const double *in, long long m,
                         int d, const double *shift, const double *mat, double *out,
                         cudaStream_t s)
{
    if (m <= 0) return UNB_OK;
    const int threads = row_threads(d);
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the row kernels", d);
    const size_t smem = (size_t)threads * odd_stride(d) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_transform, smem));
    k_transform<<<(unsigned)((m + threads - 1) / threads), threads, smem, s>>>(
        kind, inverse ? 1 : 0, in, m, d, shift, mat, out);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_loglike(unb_ctx *ctx, int kind, const double *params, int d, long long n,
                       double *like, const unsigned char *mask, const double *lparams_dev,
                       cudaStream_t s)
{
    if (n <= 0) return UNB_OK;
    const int threads = row_threads(d);
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the row kernels", d);
    const size_t smem = (size_t)threads * odd_stride(d) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_loglike, smem));
    k_loglike<<<(unsigned)((n + threads - 1) / threads), threads, smem, s>>>(
        kind, params, d, n, like, mask, lparams_dev);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_refill_tail(unb_ctx *ctx, const TailArgs &t, cudaStream_t s)
{
    if (t.m <= 0) return UNB_OK;
    const int threads = row_threads(2 * odd_stride(t.d) + 1);
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the row kernels", t.d);
    const size_t smem = (size_t)threads * 2 * odd_stride(t.d) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_refill_tail, smem));
    k_refill_tail<<<(unsigned)((t.m + threads - 1) / threads), threads, smem, s>>>(t);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_enlargement_f(unb_ctx *ctx, const double *u, int d, const int *item_idx,
                             const int *round_item_off, const int *round_nitems,
                             int max_items, const double *ctrs, const double *invcovs,
                             int rounds, unsigned long long *out_round_key, cudaStream_t s)
{
    if (rounds <= 0 || max_items <= 0) return UNB_OK;
    int threads = row_threads(d);
    if (threads < 32) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the row kernels", d);
    if (threads > 64) threads = 64;   // more blocks: the rounds are small
    const size_t smem = (size_t)threads * odd_stride(d) * sizeof(double);
    UNB_TRY(set_smem(ctx, k_enlargement_f, smem));
    dim3 grid((unsigned)((max_items + threads - 1) / threads), (unsigned)rounds);
    k_enlargement_f<<<grid, threads, smem, s>>>(u, d, item_idx, round_item_off, round_nitems,
                                               ctrs, invcovs, out_round_key);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_pairdist(unb_ctx *ctx, const double *pts, const long long *ids, int n, int d,
                        double *partial_sum, long long *partial_cnt, cudaStream_t s)
{
    if (n <= 0) return UNB_OK;
    k_pairdist<<<(unsigned)((n + 63) / 64), 64, 0, s>>>(pts, ids, n, d, partial_sum, partial_cnt);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}
