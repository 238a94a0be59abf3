// unb_cluster.cu -- locality for the membership scan: clustered live tiles + proposals binned by
// their nearest tile centroid.
//
// MLFriends.inside() only asks WHETHER a live point lies within the radius, so the order in which
// a proposal meets the live points is free.  In the headline geometry (N_live = 4000, d = 20) a
// random pair is within the radius with probability 1/280, i.e. a proposal streams ~6 of the 63
// live tiles before its first hit.  The hits are not spread evenly, though: the neighbours of a
// proposal sit in a cap around its own direction.  So
//
//   1. the fp32 live tiles are built from the live points SORTED BY CLUSTER (a few Lloyd
//      iterations of k-means with one centroid per tile; clusters laid end to end and cut into
//      tiles of 64), and every tile gets a centroid;
//   2. every proposal is binned by the tile whose centroid is nearest (one more "tile" of filter
//      work: the centroids are a 64-point tile themselves);
//   3. a block of the membership kernel that is about to stream tile T refills its free slots from
//      bin T (falling back to the following bins when it is empty).
//
// A proposal then finds its first neighbour in the first tile it sees 84 % of the time and streams
// 2.2 tiles on average instead of 5.6 (measured on the bench geometry).  None of this touches a
// decision: the permutation only reorders the live points inside the any-neighbour scan, every
// flagged pair is still decided by the reference's exact fp64 sequence on the original rows.
#include "unb_internal.cuh"

#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int CL_TILE_N = 64;          // live points per fp32 tile (REG_TILE_N of unb_scan.cu)
constexpr int KM_THREADS = 128;
constexpr int KM_ITERS = 6;

// ---- k-means over the live rows (fp32 arithmetic: this only shapes a heuristic) -------------

__global__ void k_km_init(const double *__restrict__ rows, int n, int d, int K, float *__restrict__ cent)
{
    const int c = blockIdx.x;
    const long long src = (long long)c * n / K;
    for (int k = threadIdx.x; k < d; k += blockDim.x) cent[(size_t)c * d + k] = (float)rows[(size_t)src * d + k];
}

// nearest centroid of every live point; accumulates the new centroid sums
__global__ void __launch_bounds__(KM_THREADS) k_km_assign(const double *__restrict__ rows, int n, int d, int K,
                                                          const float *__restrict__ cent, int *__restrict__ label,
                                                          float *__restrict__ sums, int *__restrict__ counts)
{
    extern __shared__ float s_cent[];   // K x d
    for (int i = threadIdx.x; i < K * d; i += blockDim.x) s_cent[i] = cent[i];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *p = rows + (size_t)i * d;
    int best = 0;
    float bestd = 3.4e38f;
    for (int c = 0; c < K; c++) {
        const float *cc = s_cent + (size_t)c * d;
        float dist = 0.f;
        for (int k = 0; k < d; k++) {
            const float diff = (float)p[k] - cc[k];
            dist = fmaf(diff, diff, dist);
        }
        if (dist < bestd) { bestd = dist; best = c; }
    }
    label[i] = best;
    if (sums) {
        for (int k = 0; k < d; k++) atomicAdd(sums + (size_t)best * d + k, (float)p[k]);
        atomicAdd(counts + best, 1);
    }
}

__global__ void k_km_update(int d, int K, const float *__restrict__ sums, const int *__restrict__ counts,
                            float *__restrict__ cent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= K * d) return;
    const int c = i / d;
    if (counts[c] > 0) cent[i] = sums[i] / (float)counts[c];
}

// ---- generic binning: histogram -> exclusive scan -> scatter ------------------------------------

__global__ void k_hist(const int *__restrict__ key, const int *__restrict__ n_dev, int n_host, int nbins,
                       int *__restrict__ counts)
{
    const int n = n_dev ? *n_dev : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(counts + key[i], 1);
    (void)nbins;
}

// start[0 .. nbins] = exclusive prefix of counts; zeroes fill[] and head[]
__global__ void k_bin_scan(const int *__restrict__ counts, int nbins, int *__restrict__ start,
                           int *__restrict__ fill, int *__restrict__ head)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int run = 0;
    for (int b = 0; b < nbins; b++) {
        start[b] = run;
        run += counts[b];
        fill[b] = 0;
        if (head) head[b] = 0;
    }
    start[nbins] = run;
}

// order[start[key] + position within the bin] = i (warp-aggregated claims)
__global__ void k_bin_scatter(const int *__restrict__ key, const int *__restrict__ n_dev, int n_host,
                              const int *__restrict__ start, int *__restrict__ fill, int *__restrict__ order)
{
    const int n = n_dev ? *n_dev : n_host;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    const int b = valid ? key[i] : -1;
    const unsigned peers = __match_any_sync(FULL, b);
    if (!valid) return;
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(fill + b, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[start[b] + base + __popc(peers & ((1u << lane) - 1))] = i;
}

// ---- centroid tiles ----------------------------------------------------------------------------
// one block per live tile: mean of its members (fp64 rows through perm) -> the centroid tiles, same
// k-major layout as the live tiles: centroid c of centroid-tile c/64 sits in column c%64; row dr holds
// -|c|^2/2 so that  score = c.b - |c|^2/2  is largest for the nearest centroid; unused columns -1e30
__global__ void __launch_bounds__(64) k_tile_centroids(const double *__restrict__ rows, const int *__restrict__ perm,
                                                       int n, int d, int dr, int ntiles, float *__restrict__ ctiles)
{
    __shared__ double s_part[64];
    const int t = blockIdx.x;                   // live tile index == centroid index
    const int nct = (ntiles + CL_TILE_N - 1) / CL_TILE_N;
    float *C = ctiles + (size_t)(t / CL_TILE_N) * (dr + 1) * CL_TILE_N + (t % CL_TILE_N);
    const int first = t * CL_TILE_N;
    const int members = (n - first) < CL_TILE_N ? (n - first) : CL_TILE_N;
    double nrm = 0.0;
    for (int k = 0; k < dr; k++) {
        double v = 0.0;
        if (k < d && threadIdx.x < members) v = rows[(size_t)perm[first + threadIdx.x] * d + k];
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int o = 32; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_part[threadIdx.x] += s_part[threadIdx.x + o];
            __syncthreads();
        }
        const double mean = members > 0 ? s_part[0] / members : 0.0;
        if (threadIdx.x == 0) C[(size_t)k * CL_TILE_N] = (float)mean;
        nrm += mean * mean;
        __syncthreads();
    }
    if (threadIdx.x == 0) C[(size_t)dr * CL_TILE_N] = (float)(-0.5 * nrm);
    (void)nct;
}

__global__ void k_fill_f32(float *__restrict__ p, long long n, float v)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- nearest tile centroid of every work item ---------------------------------------------------
// one thread per item, row in fp32 registers, the centroid tiles staged through shared memory.
template <int DR>
__global__ void __launch_bounds__(128) k_nearest_centroid(const double *__restrict__ cand, int d,
                                                          const int *__restrict__ item_idx,
                                                          const int *__restrict__ n_dev, int n_host,
                                                          const float *__restrict__ ctiles, int ntiles,
                                                          int *__restrict__ bin_of, int *__restrict__ counts)
{
    __shared__ __align__(16) float s_tile[(DR + 1) * CL_TILE_N];
    __shared__ int s_hist[256];
    const int n = n_dev ? *n_dev : n_host;
    const int item = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = item < n;
    for (int b = threadIdx.x; b < 256; b += blockDim.x) s_hist[b] = 0;
    float a[DR];
    if (valid) {
        const int r = item_idx ? item_idx[item] : item;
#pragma unroll
        for (int k = 0; k < DR; k++) a[k] = (k < d) ? (float)cand[(size_t)r * d + k] : 0.f;
    } else {
#pragma unroll
        for (int k = 0; k < DR; k++) a[k] = 0.f;
    }
    float best = -3.4e38f;
    int best_c = 0;
    const int nct = (ntiles + CL_TILE_N - 1) / CL_TILE_N;
    for (int ct = 0; ct < nct; ct++) {
        __syncthreads();
        const float *src = ctiles + (size_t)ct * (DR + 1) * CL_TILE_N;
        for (int i = threadIdx.x; i < (DR + 1) * CL_TILE_N; i += blockDim.x) s_tile[i] = src[i];
        __syncthreads();
#pragma unroll 1
        for (int g = 0; g < CL_TILE_N / 4; g++) {
            const float *Tg = s_tile + g * 4;
            const float4 h = *reinterpret_cast<const float4 *>(Tg + DR * CL_TILE_N);
            float2 acc01 = make_float2(h.x, h.y), acc23 = make_float2(h.z, h.w);
#pragma unroll
            for (int k = 0; k < DR; k++) {
                const float4 b = *reinterpret_cast<const float4 *>(Tg + k * CL_TILE_N);
                const float2 ak = make_float2(a[k], a[k]);
                acc01 = ffma2(ak, make_float2(b.x, b.y), acc01);
                acc23 = ffma2(ak, make_float2(b.z, b.w), acc23);
            }
            const int c0 = ct * CL_TILE_N + g * 4;
            if (acc01.x > best && c0 + 0 < ntiles) { best = acc01.x; best_c = c0 + 0; }
            if (acc01.y > best && c0 + 1 < ntiles) { best = acc01.y; best_c = c0 + 1; }
            if (acc23.x > best && c0 + 2 < ntiles) { best = acc23.x; best_c = c0 + 2; }
            if (acc23.y > best && c0 + 3 < ntiles) { best = acc23.y; best_c = c0 + 3; }
        }
    }
    if (valid) {
        bin_of[item] = best_c;
        if (ntiles <= 256) atomicAdd(&s_hist[best_c], 1);
        else atomicAdd(counts + best_c, 1);
    }
    __syncthreads();
    if (ntiles <= 256)
        for (int b = threadIdx.x; b < ntiles; b += blockDim.x)
            if (s_hist[b]) atomicAdd(counts + b, s_hist[b]);
}

template <int DR>
int launch_nearest(unb_ctx *ctx, const double *cand, int d, const int *item_idx, const int *n_dev,
                   long long n_host, const float *ctiles, int ntiles, int *bin_of, int *counts,
                   cudaStream_t s)
{
    k_nearest_centroid<DR><<<(unsigned)((n_host + 127) / 128), 128, 0, s>>>(
        cand, d, item_idx, n_dev, (int)n_host, ctiles, ntiles, bin_of, counts);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

}  // namespace

size_t unb_cluster_max_tiles() { return 256; }

// k-means over the live rows -> perm (tile slot -> live row, clusters laid end to end).
// scratch_i: at least 3 n + 4 K + 8 ints; scratch_f: at least 2 K d floats.
int unb_launch_cluster_live(unb_ctx *ctx, const double *rows, int n, int d, int K, int *perm,
                            int *scratch_i, float *scratch_f, cudaStream_t s)
{
    if (n <= 0 || K <= 0) return UNB_OK;
    int *label = scratch_i;                 // [n]
    int *counts = scratch_i + n;            // [K]
    int *start = counts + K;                // [K + 1]
    int *fill = start + K + 1;              // [K]
    float *cent = scratch_f;                // [K d]
    float *sums = scratch_f + (size_t)K * d;
    const size_t smem = (size_t)K * d * sizeof(float);
    if (smem > 200 * 1024) return unb_fail(ctx, UNB_ERR_ARG, "live block too large for tile clustering");
    if (smem > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(k_km_assign, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_km_init<<<K, 32, 0, s>>>(rows, n, d, K, cent);
    const unsigned nb = (unsigned)((n + KM_THREADS - 1) / KM_THREADS);
    for (int it = 0; it < KM_ITERS; it++) {
        UNB_CUDA(ctx, cudaMemsetAsync(sums, 0, (size_t)K * d * sizeof(float), s));
        UNB_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int), s));
        k_km_assign<<<nb, KM_THREADS, smem, s>>>(rows, n, d, K, cent, label, sums, counts);
        k_km_update<<<(unsigned)((K * d + 127) / 128), 128, 0, s>>>(d, K, sums, counts, cent);
        ctx->launches += 2;
    }
    // final labels for the final centroids, then a counting sort by label
    UNB_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)K * sizeof(int), s));
    k_km_assign<<<nb, KM_THREADS, smem, s>>>(rows, n, d, K, cent, label, nullptr, nullptr);
    k_hist<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(label, nullptr, n, K, counts);
    k_bin_scan<<<1, 32, 0, s>>>(counts, K, start, fill, nullptr);
    k_bin_scatter<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(label, nullptr, n, start, fill, perm);
    ctx->launches += 5;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

// centroid of every fp32 live tile (members through perm) in tile layout; ctiles holds
// ceil(ntiles / 64) * (dr + 1) * 64 floats
int unb_launch_tile_centroids(unb_ctx *ctx, const double *rows, const int *perm, int n, int d, int dr,
                              int ntiles, float *ctiles, cudaStream_t s)
{
    const int nct = (ntiles + CL_TILE_N - 1) / CL_TILE_N;
    const long long total = (long long)nct * (dr + 1) * CL_TILE_N;
    k_fill_f32<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(ctiles, total, 0.f);
    // unused centroid columns must never win: their offset row is very negative
    for (int ct = 0; ct < nct; ct++) {
        float *hrow = ctiles + (size_t)ct * (dr + 1) * CL_TILE_N + (size_t)dr * CL_TILE_N;
        k_fill_f32<<<1, 64, 0, s>>>(hrow, CL_TILE_N, -1e30f);
    }
    k_tile_centroids<<<ntiles, 64, 0, s>>>(rows, perm, n, d, dr, ntiles, ctiles);
    ctx->launches += 2 + nct;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

// bins the work items of a membership launch by nearest tile centroid:
//   bin_of[item], order[] (items sorted by bin), meta = counts[ntiles] | start[ntiles+1] | fill[ntiles] | head[ntiles]
int unb_launch_bin_items(unb_ctx *ctx, const double *cand, int d, int dr, const int *item_idx,
                         const int *n_items_dev, long long n_items, const float *ctiles, int ntiles,
                         int *bin_of, int *order, int *meta, cudaStream_t s)
{
    if (n_items <= 0) return UNB_OK;
    int *counts = meta, *start = meta + ntiles, *fill = start + ntiles + 1, *head = fill + ntiles;
    UNB_CUDA(ctx, cudaMemsetAsync(counts, 0, (size_t)ntiles * sizeof(int), s));
    switch (dr) {
#define UNB_NEAREST(DR_) case DR_: UNB_TRY(launch_nearest<DR_>(ctx, cand, d, item_idx, n_items_dev, n_items, ctiles, ntiles, bin_of, counts, s)); break;
        UNB_NEAREST(4) UNB_NEAREST(8) UNB_NEAREST(12) UNB_NEAREST(16)
        UNB_NEAREST(20) UNB_NEAREST(24) UNB_NEAREST(28) UNB_NEAREST(32)
#undef UNB_NEAREST
    default: return unb_fail(ctx, UNB_ERR_ARG, "binning needs ndim <= 32");
    }
    k_bin_scan<<<1, 32, 0, s>>>(counts, ntiles, start, fill, head);
    k_bin_scatter<<<(unsigned)((n_items + 255) / 256), 256, 0, s>>>(bin_of, n_items_dev, (int)n_items, start, fill, order);
    ctx->launches += 2;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}
