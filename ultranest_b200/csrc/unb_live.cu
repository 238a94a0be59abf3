// unb_live.cu -- in-place row updates of the mirrored live block in ONE launch.
//
// The integrator replaces one live point per iteration (integrator.py:2753-2756:
// `region.u[worst] = u; region.unormed[worst] = transform(u)`), then calls `region.inside(active_u)`
// (integrator.py:1855).  Every image the scans read must follow: the row-major fp64 rows (copied by
// the caller), the fp64 tile column, the squared norm, the filter offset h of the current scan
// mode, the fp32 tile column and its offset.  Doing that in one small kernel -- instead of
// invalidating the derived images and rebuilding them with four launches and a device->host read of
// the largest norm -- is what keeps the per-iteration call short.
//
// The largest squared norm is kept as a running UPPER BOUND (atomicMax here, exact again at the
// next full rebuild): it only enters error margins (fp32 filter usability, certain-neighbour level,
// candidate set of the max-min scan), which stay valid, merely a hair wider, under an overestimate.
#include "unb_internal.cuh"

namespace {

__global__ void k_live_update(const LiveUpdateArgs A)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.nrows) return;
    const int slot = A.idx[r];
    if (slot < 0 || slot >= A.n) return;
    const double *row = A.rows + (size_t)slot * A.d;
    // fp64 tile column + squared norm (same k-sequential FMA chain as the full build)
    double na = 0.0;
    double *T = nullptr;
    if (A.tiles) {
        const int t = slot / A.tile_n, c = slot - t * A.tile_n;
        T = A.tiles + (size_t)t * (A.dr + 1) * A.tile_n + c;
    }
    for (int k = 0; k < A.dr; k++) {
        const double v = (k < A.d) ? row[k] : 0.0;
        if (T) T[(size_t)k * A.tile_n] = v;
        na = fma(v, v, na);
    }
    A.norms[slot] = na;
    atomicMax(A.namax, f64_bits(na));
    if (T && A.h_mode != HMODE_NONE) {
        double h;
        if (A.h_mode == HMODE_THRESH) {
            const double r2w = __dmul_rn(A.h_r2, __dadd_rn(1.0, A.kappa));
            const double naw = __dmul_rn(na, __dsub_rn(1.0, A.kappa));
            h = __dmul_rn(0.5, __dsub_rn(r2w, naw));
        } else {
            h = __dmul_rn(-0.5, na);
        }
        T[(size_t)A.dr * A.tile_n] = h;
    }
    if (A.tiles32) {   // fp32 image, always 64-point tiles, row order
        const int t = slot / 64, c = slot - t * 64;
        float *F = A.tiles32 + (size_t)t * (A.dr + 1) * 64 + c;
        for (int k = 0; k < A.dr; k++)
            F[(size_t)k * 64] = (k < A.d) ? __double2float_rn(row[k]) : 0.f;
        const double r2w = __dmul_rn(A.t32_r2, __dadd_rn(1.0, A.kappa32));
        const double naw = __dmul_rn(na, __dsub_rn(1.0, A.kappa32));
        F[(size_t)A.dr * 64] = __double2float_ru(__dmul_rn(0.5, __dsub_rn(r2w, naw)));
    }
}

// ---------------------------------------------------------------------------------------
// per-round first and second moments of the SELECTED rows of a bootstrap round, about a fixed
// reference point c0 (so that the later  sum(yy^T) - n ybar ybar^T  does not cancel):
//   sums[r][p] = sum_i y_ip,  sxx[r][p][q] = sum_i y_ip y_iq,  y = u[idx] - c0.
// One block per (slice of rows, round): rows staged in shared memory, thread t accumulates the
// (p, q) pairs t, t + blockDim, ... over the slice and adds them to the round's totals.
// These are APPROXIMATE inputs of the enlargement screen (mlfriends.py): any summation order will
// do, the rounds that can decide the result are recomputed with the reference's own NumPy algebra.
// ---------------------------------------------------------------------------------------
constexpr int MOM_ROWS = 64;
constexpr int MOM_THREADS = 256;

__global__ void __launch_bounds__(MOM_THREADS) k_round_moments(const double *__restrict__ u, int d,
                                                               const int *__restrict__ idxA,
                                                               const int *__restrict__ offA,
                                                               const int *__restrict__ nA,
                                                               const double *__restrict__ c0,
                                                               double *__restrict__ sums,
                                                               double *__restrict__ sxx)
{
    extern __shared__ __align__(16) double y[];   // MOM_ROWS x d
    const int round = blockIdx.y;
    const int n = nA[round];
    const int first = blockIdx.x * MOM_ROWS;
    if (first >= n) return;
    const int rows = (n - first) < MOM_ROWS ? (n - first) : MOM_ROWS;
    const int *idx = idxA + offA[round] + first;
    for (int e = threadIdx.x; e < rows * d; e += MOM_THREADS) {
        const int i = e / d, k = e - i * d;
        y[e] = u[(size_t)idx[i] * d + k] - c0[k];
    }
    __syncthreads();
    for (int pq = threadIdx.x; pq < d * d + d; pq += MOM_THREADS) {
        double acc = 0.0;
        if (pq < d * d) {
            const int p = pq / d, q = pq - p * d;
            if (q < p) continue;   // symmetric: upper triangle only
            for (int i = 0; i < rows; i++) acc = fma(y[i * d + p], y[i * d + q], acc);
            atomicAdd(sxx + (size_t)round * d * d + pq, acc);
        } else {
            const int p = pq - d * d;
            for (int i = 0; i < rows; i++) acc += y[i * d + p];
            atomicAdd(sums + (size_t)round * d + p, acc);
        }
    }
}

}  // namespace

int unb_launch_round_moments(unb_ctx *ctx, const double *u, int d, const int *idxA, const int *offA,
                             const int *nA, int max_rows, int rounds, const double *c0, double *sums,
                             double *sxx, cudaStream_t s)
{
    if (rounds <= 0 || max_rows <= 0) return UNB_OK;
    const size_t smem = (size_t)MOM_ROWS * d * sizeof(double);
    if (smem > 200 * 1024) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the moments kernel", d);
    if (smem > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(k_round_moments, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UNB_CUDA(ctx, cudaMemsetAsync(sums, 0, (size_t)rounds * d * sizeof(double), s));
    UNB_CUDA(ctx, cudaMemsetAsync(sxx, 0, (size_t)rounds * d * d * sizeof(double), s));
    dim3 grid((unsigned)((max_rows + MOM_ROWS - 1) / MOM_ROWS), (unsigned)rounds);
    k_round_moments<<<grid, MOM_THREADS, smem, s>>>(u, d, idxA, offA, nA, c0, sums, sxx);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_live_update(unb_ctx *ctx, const LiveUpdateArgs &a, cudaStream_t s)
{
    if (a.nrows <= 0) return UNB_OK;
    k_live_update<<<(unsigned)((a.nrows + 63) / 64), 64, 0, s>>>(a);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}
