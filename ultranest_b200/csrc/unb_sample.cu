// unb_sample.cu -- device-side proposal generation ("throughput mode") for the MLFriends region.
//
// The reference draws proposals on the host from np.random (legacy MT19937) and filters them
// (mlfriends.pyx:1096-1112 sample_from_boundingbox, :1135-1160 sample_from_wrapping_ellipsoid).
// The parity path of this library keeps that stream on the host.  Behind an explicit flag the
// draws can instead be made ON THE DEVICE with a counter-based generator (Philox4x32-10, keyed by
// the caller's seed; the counter is the proposal's global index), so nothing crosses PCIe on the
// way in and only accepted rows come back.  Statistically the same proposals, NOT the same
// random stream: a seeded run in this mode is reproducible but differs from the reference's.
//
//   k_draw          proposals + unit-cube mask (one thread per proposal, row in shared memory)
//   k_finish_mask   member &= cube [&& like > Lmin]
//   k_block_counts / k_scan_counts / k_scatter_rows
//                   ORDERED compaction of the accepted rows (candidate order is kept, like the
//                   reference's boolean-mask indexing, so results do not depend on scheduling)
#include "unb_internal.cuh"

#include <cstring>

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int DRAW_THREADS = 128;
constexpr int COMPACT_ROWS = 1024;   // rows per block of the compaction kernels

__host__ __device__ inline int sample_stride(int d) { return d | 1; }

// Philox4x32-10 (Salmon et al. 2011, "Parallel random numbers: as easy as 1, 2, 3")
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// 52-bit uniform in the OPEN interval (0, 1) from two 32-bit words: (x + 1/2) / 2^52 is exact in a
// double for every 52-bit x, so neither 0 nor 1 can come out (log(u) and u**(1/d) stay finite)
__device__ __forceinline__ double u52(unsigned a, unsigned b)
{
    return ((double)(a >> 6) * 67108864.0 + (double)(b >> 6) + 0.5) * (1.0 / 4503599627370496.0);
}

struct DrawArgs {
    int method;               // UNB_SAMPLE_*
    long long m;
    int d;
    unsigned long long seed, offset;   // offset: global index of this call's first proposal
    const double *center;     // [d]
    const double *axes_T;     // [d x d] row-major (ellipsoid_axes_T)
    double scale;             // sqrt(enlarge)
    double *out;              // [m x d]
    unsigned char *cube;      // [m]: every coordinate strictly inside (0, 1)
};

__global__ void __launch_bounds__(DRAW_THREADS) k_draw(const DrawArgs A)
{
    extern __shared__ __align__(16) double rowbuf[];
    const int d = A.d, ds = sample_stride(d);
    double *z = rowbuf + (size_t)threadIdx.x * ds;
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.m) return;
    const unsigned long long g = A.offset + (unsigned long long)j;
    const uint2 key = make_uint2((unsigned)A.seed, (unsigned)(A.seed >> 32));
    double *w = A.out + (size_t)j * d;
    bool inside = true;
    if (A.method == UNB_SAMPLE_UNIT_CUBE) {
        // np.random.uniform(size=(nsamples, ndim)) (mlfriends.pyx:1105): two coordinates per block
        for (int k = 0; k < d; k += 2) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)(k >> 1), 1u), key);
            const double a = u52(r.x, r.y), b = u52(r.z, r.w);
            w[k] = a;
            if (k + 1 < d) w[k + 1] = b;
        }
    } else {
        // z ~ N(0, I) (Box-Muller, two normals per Philox block), z /= |z|,
        // u = z * enlarge**0.5 * U**(1/d), w = center + u . axes_T   (mlfriends.pyx:1145-1151)
        double nrm = 0.0;
        for (int k = 0; k < d; k += 2) {
            const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), (unsigned)(k >> 1), 0u), key);
            const double rad = sqrt(-2.0 * log(u52(r.x, r.y)));
            double sn, cs;
            sincospi(2.0 * u52(r.z, r.w), &sn, &cs);
            z[k] = rad * cs;
            nrm = fma(z[k], z[k], nrm);
            if (k + 1 < d) {
                z[k + 1] = rad * sn;
                nrm = fma(z[k + 1], z[k + 1], nrm);
            }
        }
        const uint4 r = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 0x7fffffffu, 0u), key);
        const double f = A.scale * pow(u52(r.x, r.y), 1.0 / (double)d) / sqrt(nrm);
        for (int k = 0; k < d; k++) z[k] *= f;
        for (int c = 0; c < d; c++) {
            double acc = __ldg(A.center + c);
            for (int k = 0; k < d; k++) acc = fma(z[k], __ldg(A.axes_T + (size_t)k * d + c), acc);
            w[c] = acc;
            inside &= (acc > 0.0) && (acc < 1.0);
        }
    }
    A.cube[j] = inside ? 1 : 0;
}

__global__ void k_finish_mask(unsigned char *__restrict__ mask, const unsigned char *__restrict__ cube,
                              const double *__restrict__ like, double Lmin, int use_lmin, long long m)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    bool ok = mask[j] != 0 && cube[j] != 0;
    if (ok && use_lmin) ok = like[j] > Lmin;
    mask[j] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256) k_block_counts(const unsigned char *__restrict__ mask, long long m,
                                                      int *__restrict__ counts)
{
    __shared__ int s_cnt[8];
    const long long base = (long long)blockIdx.x * COMPACT_ROWS;
    int c = 0;
    for (int i = threadIdx.x; i < COMPACT_ROWS; i += 256) {
        const long long j = base + i;
        c += (j < m && mask[j]) ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += s_cnt[w];
        counts[blockIdx.x] = t;
    }
}

// exclusive scan of the block counts by one block (serial over chunks of 1024; the list is short:
// 1024 blocks for 2^20 rows); counts[nblocks] receives the total
__global__ void __launch_bounds__(1024) k_scan_counts(int *__restrict__ counts, int nblocks, int *__restrict__ total)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if ((threadIdx.x & 31) >= o) incl += t;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int wv = s_warp[threadIdx.x], winc = wv;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, winc, o);
                if (threadIdx.x >= o) winc += t;
            }
            s_warp[threadIdx.x] = winc - wv;   // exclusive warp offsets
        }
        __syncthreads();
        const int carry = s_carry;
        if (i < nblocks) counts[i] = carry + s_warp[threadIdx.x >> 5] + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__global__ void __launch_bounds__(256) k_scatter_rows(const unsigned char *__restrict__ mask, long long m, int d,
                                                      const int *__restrict__ block_off,
                                                      const double *__restrict__ rows,
                                                      const double *__restrict__ like,
                                                      double *__restrict__ out_rows,
                                                      double *__restrict__ out_like,
                                                      int *__restrict__ out_index)
{
    __shared__ int s_dst[COMPACT_ROWS];
    __shared__ int s_warp[8];
    __shared__ int s_run;
    const long long base = (long long)blockIdx.x * COMPACT_ROWS;
    if (threadIdx.x == 0) s_run = block_off[blockIdx.x];
    __syncthreads();
    // destination of every accepted row of the block, in row order
    for (int i0 = 0; i0 < COMPACT_ROWS; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const long long j = base + i;
        const bool ok = j < m && mask[j];
        const unsigned ball = __ballot_sync(FULL, ok);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) s_warp[warp] = __popc(ball);
        __syncthreads();
        int off = s_run;
        for (int w = 0; w < warp; w++) off += s_warp[w];
        s_dst[i] = ok ? off + __popc(ball & ((1u << lane) - 1)) : -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += s_warp[w];
            s_run += t;
        }
        __syncthreads();
    }
    // cooperative copy: consecutive threads move consecutive doubles of a row
    const long long nrows = (m - base) < COMPACT_ROWS ? (m - base) : COMPACT_ROWS;
    for (long long e = threadIdx.x; e < nrows * d; e += 256) {
        const int i = (int)(e / d), k = (int)(e - (long long)i * d);
        const int dst = s_dst[i];
        if (dst >= 0) out_rows[(size_t)dst * d + k] = rows[(size_t)(base + i) * d + k];
    }
    for (int i = threadIdx.x; i < nrows; i += 256) {
        const int dst = s_dst[i];
        if (dst >= 0) {
            if (out_like) out_like[dst] = like[base + i];
            if (out_index) out_index[dst] = (int)(base + i);
        }
    }
}

}  // namespace

int unb_launch_draw(unb_ctx *ctx, int method, long long m, int d, unsigned long long seed,
                    unsigned long long offset, const double *center_dev, const double *axes_T_dev,
                    double scale, double *out_dev, unsigned char *cube_dev, cudaStream_t s)
{
    if (m <= 0) return UNB_OK;
    const size_t smem = (size_t)DRAW_THREADS * sample_stride(d) * sizeof(double);
    if (smem > 200 * 1024) return unb_fail(ctx, UNB_ERR_ARG, "ndim=%d too large for the proposal generator", d);
    if (smem > 48 * 1024)
        UNB_CUDA(ctx, cudaFuncSetAttribute(k_draw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DrawArgs A;
    memset(&A, 0, sizeof(A));
    A.method = method;
    A.m = m;
    A.d = d;
    A.seed = seed;
    A.offset = offset;
    A.center = center_dev;
    A.axes_T = axes_T_dev;
    A.scale = scale;
    A.out = out_dev;
    A.cube = cube_dev;
    k_draw<<<(unsigned)((m + DRAW_THREADS - 1) / DRAW_THREADS), DRAW_THREADS, smem, s>>>(A);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

int unb_launch_finish_mask(unb_ctx *ctx, unsigned char *mask_dev, const unsigned char *cube_dev,
                           const double *like_dev, double Lmin, bool use_lmin, long long m,
                           cudaStream_t s)
{
    if (m <= 0) return UNB_OK;
    k_finish_mask<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(mask_dev, cube_dev, like_dev, Lmin,
                                                              use_lmin ? 1 : 0, m);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

// ordered compaction; scratch_counts must hold (nblocks + 1) ints; *total_dev receives the count
int unb_launch_compact_rows(unb_ctx *ctx, const unsigned char *mask_dev, long long m, int d,
                            const double *rows_dev, const double *like_dev, int *scratch_counts,
                            int *total_dev, double *out_rows_dev, double *out_like_dev,
                            int *out_index_dev, cudaStream_t s)
{
    if (m <= 0) {
        UNB_CUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(int), s));
        return UNB_OK;
    }
    const int nblocks = (int)((m + COMPACT_ROWS - 1) / COMPACT_ROWS);
    k_block_counts<<<nblocks, 256, 0, s>>>(mask_dev, m, scratch_counts);
    k_scan_counts<<<1, 1024, 0, s>>>(scratch_counts, nblocks, total_dev);
    k_scatter_rows<<<nblocks, 256, 0, s>>>(mask_dev, m, d, scratch_counts, rows_dev, like_dev,
                                           out_rows_dev, out_like_dev, out_index_dev);
    ctx->launches += 3;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

size_t unb_compact_scratch_ints(long long m) { return (size_t)((m + COMPACT_ROWS - 1) / COMPACT_ROWS) + 2; }
