// unb_internal.cuh -- shared declarations of the sm_100a MLFriends engine (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ultranest_b200.h"

// NVTX range around an ABI call (header-only NVTX v3: a no-op unless a tool is attached), so that
// a timeline of a run shows the region calls by name next to their kernels
struct UnbRange {
    explicit UnbRange(const char *name) { nvtxRangePushA(name); }
    ~UnbRange() { nvtxRangePop(); }
};
#define UNB_RANGE(name) UnbRange unb_range_guard__(name)

// ---------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------
struct unb_ctx;
int unb_fail(unb_ctx *ctx, int code, const char *fmt, ...);

#define UNB_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t err__ = (call);                                                          \
        if (err__ != cudaSuccess)                                                            \
            return unb_fail((ctx), UNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call,             \
                            cudaGetErrorString(err__), __FILE__, __LINE__);                  \
    } while (0)

#define UNB_TRY(call)                                                                        \
    do {                                                                                     \
        int rc__ = (call);                                                                   \
        if (rc__ != UNB_OK) return rc__;                                                     \
    } while (0)

// ---------------------------------------------------------------------------------------
// growable device / pinned-host buffers
// ---------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
};
int unb_reserve(unb_ctx *ctx, DevBuf &b, size_t bytes);
int unb_reserve_pinned(unb_ctx *ctx, PinBuf &b, size_t bytes);
int unb_wait_small(unb_ctx *ctx);   // before (re)using ctx->pin_small

// ---------------------------------------------------------------------------------------
// tiled live block
//
// The (N x d) row-major t-space live block (MLFriends.unormed) is re-laid out in HBM as
// ceil(N / tile_n) tiles.  One tile is (dr + 1) rows of tile_n doubles, k-major:
//     row k < dr : coordinate k of the tile's tile_n live points (0 for k >= d or i >= N)
//     row dr     : the filter offset h_i (depends on scan mode / radius, see unb_scan.cu)
// so a tile is ONE contiguous chunk that a single cp.async.bulk (TMA 1-D) moves into shared
// memory, and inside shared memory the coordinates of 4 consecutive live points for a fixed k
// are two broadcast LDS.128.  dr = d rounded up to a multiple of 4.
// ---------------------------------------------------------------------------------------
struct LiveTiles {
    DevBuf tiles;     // ntiles * (dr+1) * tile_n doubles
    DevBuf rows;      // n * d doubles, row-major copy (exact-path kernels, gathers)
    DevBuf norms;     // n doubles: computed squared norms na_i (k-sequential FMA)
    DevBuf namax;     // 1 double: max_i na_i (as ordered uint64)
    size_t n = 0, d = 0, dr = 0, tile_n = 0, ntiles = 0;
    int h_mode = -1;      // which h row is currently stored (HMODE_*)
    double h_r2 = 0.0;
    bool valid = false;
    // fp32 copy of the tiles for the single-precision pre-filter of the membership kernel
    DevBuf tiles32;       // ntiles * (dr+1) * tile_n floats, h row built for t32_r2
    bool t32_valid = false;
    double t32_r2 = 0.0;
    double namax_host = 0.0;   // max squared norm, fetched when the fp32 tiles are built
    // locality (unb_cluster.cu): the fp32 tiles may hold the live points in CLUSTER order
    // (perm32: tile slot -> live row) with one centroid per tile (ctiles32, tile layout)
    bool want_cluster = false;     // requested by a large membership launch on the region's block
    bool cluster_valid = false;    // perm32 / ctiles32 describe the current rows
    bool t32_clustered = false;    // tiles32 was built through perm32
    DevBuf perm32, ctiles32, cl_scratch_i, cl_scratch_f;
};

enum { HMODE_NONE = -1, HMODE_THRESH = 0, HMODE_MIN = 1 };

enum ScanMode {
    SCAN_FIND = 0,      // first index within r2 (ordered, early exit)   find_nearby
    SCAN_COUNT = 1,     // number within r2                               count_nearby
    SCAN_SUBTRACT = 2,  // count + ordered coordinate sum of hits         _subtract_nearby
    SCAN_MIN = 3        // exact min distance                             compute_maxradiussq
};

enum XformKind { XF_NONE = 0, XF_SCALING = 1, XF_AFFINE = 2 };

// Arguments shared by all scan kernels
struct ScanArgs {
    // live side
    const double *tiles;      // tiled live block (of this launch / of round 0); NULL -> exact kernel
    const float *tiles32;     // fp32 tiles (membership kernel with the fp32 pre-filter), nullable
    const int *perm32;        // nullable: fp32 tile slot -> live row (clustered tiles)
    // proposals binned by nearest tile centroid (nullable: one global work queue instead)
    const int *bin_order;     // work items sorted by bin
    const int *bin_start;     // [ntiles + 1]
    int *bin_head;            // [ntiles] claimed so far (device counters, zeroed per launch)
    double kappa32;           // slack factor of the fp32 filter
    int coop_max;             // block membership kernel: survivors at which the drain turns cooperative
    double namax32;           // max squared norm of the live block (certain-neighbour level); +inf: off
    const double *live_rows;  // row-major (n x d) live block (plain exact kernel)
    const int *live_idx;      // nullable: exact kernel scans rows live_idx[i] (bootstrap rounds)
    const int *round_live_off;   // nullable: offset of the round's list in live_idx
    int n_live;
    int tile_n;
    int d;
    int dr;
    // per-round addressing (bootstrap): tiles + blockIdx.y * round_tile_stride
    long long round_tile_stride;   // in doubles; 0 for single-round launches
    const int *round_nlive;        // nullable
    const int *round_nitems;       // nullable
    const int *round_item_off;     // nullable: offset of the round's item list in item_idx
    // candidate side
    const double *cand;       // row-major (M x d), already in the live block's space
    const int *item_idx;      // nullable: work item -> candidate row
    const int *out_row_idx;   // nullable: work item -> output row (default: candidate row)
    const int *n_items_dev;   // nullable: device-resident item count (after compaction)
    long long n_items;        // host-known item count (upper bound if n_items_dev != NULL)
    // scan parameters
    double r2;
    double kappa;             // filter slack factor (see unb_scan.cu)
    const unsigned long long *namax_bits;   // device: bits of max live squared norm (SCAN_MIN)
    // outputs (indexed by output row)
    long long *out_idx;       // FIND: first index / COUNT: count            (nullable for FIND)
    unsigned char *out_mask;  // FIND: idx >= 0                               (nullable)
    double *out_rows;         // SUBTRACT: (M x d) result rows
    double *out_min;          // MIN: per-candidate exact min distance        (nullable)
    double *out_like;         // any-kernel: rows without a neighbour get -inf (nullable)
    unsigned long long *out_round_max;   // MIN: per-round max over candidates (ordered bits)
    // Transform tolerance (membership kernels): the device whitens proposals in a DEFINED order,
    // the reference with OpenBLAS dgemm; the two t-rows differ by a few ulp, so a pair distance
    // within unc_tau of the radius could be decided differently by the reference.  Such exact
    // decisions are counted here, and the host shim re-decides the call with the reference's own
    // np.dot transform (never observed in 4e7 proposals; ~1e-12 relative wide).  0 / NULL: off.
    double unc_tau;
    unsigned int *unc_count;
    unsigned long long *stat_rechecks;   // diagnostic counter (nullable)
    unsigned long long *stat_tiles;      // diagnostic: warp-tiles filtered (nullable)
};

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
struct RegionState {
    LiveTiles live;
    std::vector<double> snapshot;   // host copy of the mirrored unormed (row diffing)
    int layer_kind = UNB_LAYER_IDENTITY;
    size_t layer_d = 0;
    DevBuf layer_shift, layer_mat;
    std::vector<double> layer_shift_h, layer_mat_h, ell_center_h, ell_invcov_h;   // host snapshots
    bool have_ellipsoid = false;
    size_t ell_d = 0;
    double enlarge = 0.0;
    double ell_fro = 0.0;     // Frobenius norm of the inverse covariance (ellipsoid filter band)
    DevBuf ell_center, ell_invcov;
    DevBuf ell_invcov_pad, layer_mat_pad;   // zero-padded copies (row stride pad_stride) for k_prep_tile
    size_t pad_stride = 0;
    bool have_radius = false;
    double r2 = 0.0;
    double unc_tau = 0.0;          // transform tolerance on pair distances (0: not reported)
    long long param_version = 0;   // bumped whenever layer / ellipsoid parameters change
};

// one in-flight chunk of a host-buffer call: its own stream, device scratch and pinned staging,
// so chunk c+1's H2D overlaps chunk c's kernels and D2H
struct Lane {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_in = nullptr;    // H2D from pin_in finished
    cudaEvent_t ev_done = nullptr;  // everything of the chunk finished
    DevBuf cand, tcand, items, counter, mask, idx, like;
    DevBuf smp_cube, smp_counts, smp_rows, smp_like, smp_n;   // device-side proposal generation
    DevBuf bin_of, bin_order, bin_meta;                       // proposals binned by nearest tile centroid
    PinBuf pin_in, pin_mask, pin_like, pin_idx, pin_n;
    // deferred copy-out of a staged result
    unsigned char *pend_mask = nullptr;
    double *pend_like = nullptr;
    long long *pend_idx = nullptr;
    size_t pend_rows = 0;
};

struct unb_ctx {
    int device = 0;
    int sm_count = 0;
    std::string err;
    long long launches = 0;
    long long h2d_bytes = 0, d2h_bytes = 0;
    long long last_rechecks = 0;
    long long last_tile_visits = 0;
    int exact_only = 0;
    int filter_fp32 = 1;
    int sure_level = 1;      // UNB_OPT_SURE_LEVEL
    int coop_max = -1;       // UNB_OPT_COOP_MAX (-1: the kernel's default)
    int block_kernel = 0;    // UNB_OPT_BLOCK_KERNEL
    long long bin_min_rows = 1 << 17;   // UNB_OPT_BIN_MIN_ROWS: launches from this size on use clustered tiles + bins
    long long chunk_rows = 0;

    Lane lane[2];
    RegionState region;
    LiveTiles scratch_live;          // stateless calls
    DevBuf aux0, aux1, aux2, aux3, stat;
    DevBuf unc;                      // one uint: exact decisions inside the transform tolerance
    long long last_uncertain = 0;    // ... of the last host-buffer membership call
    unsigned int unc_seen = 0;       // value of the (monotone) device counter at the last read
    DevBuf lparams;                  // likelihood parameters
    DevBuf refill_params;            // fused refill: xform scale/lo, tregion center/invcov, counters
    // bootstrap scratch
    DevBuf boot_rows, boot_u, boot_tiles, boot_idx, boot_meta, boot_out, boot_ell;
    PinBuf pin_small;
    cudaEvent_t ev_small = nullptr;   // last asynchronous read of pin_small (row patches of the mirror)
    bool ev_small_pending = false;
    // device-side proposal generation (unb_sample.cu)
    DevBuf smp_axes, smp_center;
    std::vector<double> smp_axes_h;
    // population step-sampler helpers (unb_stepfuncs.cu): scratch buffers + slice-loop session
    DevBuf sf[14];
    DevBuf sf_params;
    bool ps_active = false;
    size_t ps_popsize = 0, ps_ndim = 0, ps_count = 0;
    int ps_xform_kind = 0, ps_loglike_kind = 0;
    const double *ps_scale = nullptr, *ps_lo = nullptr, *ps_lparams = nullptr;
    double ps_thr = 0.0, ps_shrink = 1.0;
};

// ---------------------------------------------------------------------------------------
// launchers implemented in the .cu files
// ---------------------------------------------------------------------------------------
int unb_live_build(unb_ctx *ctx, LiveTiles &L, const double *rows_dev, size_t n, size_t d,
                   cudaStream_t s);
int unb_live_update_rows(unb_ctx *ctx, LiveTiles &L, const int *rows_dev_idx, size_t nrows,
                         cudaStream_t s);
void unb_live_note_host_row(LiveTiles &L, const double *row);
int unb_live_set_h(unb_ctx *ctx, LiveTiles &L, int h_mode, double r2, cudaStream_t s);
// builds / refreshes the fp32 tiles for radius r2; *usable says whether the fp32 pre-filter is
// safe and worthwhile for this block and radius (ranges, slack thin compared with r2)
int unb_live_prepare32(unb_ctx *ctx, LiveTiles &L, double r2, bool *usable, cudaStream_t s);
double unb_kappa32(size_t d);
double unb_kappa(size_t d);
size_t unb_pick_tile_n(size_t d);
int unb_launch_scan(unb_ctx *ctx, int mode, const ScanArgs &a, int rounds, cudaStream_t s);
int unb_launch_inside_any(unb_ctx *ctx, const ScanArgs &a, int *queue_head, cudaStream_t s);
int unb_launch_gather_round_tiles(unb_ctx *ctx, const double *rows, int n, int d, int dr,
                                  int tile_n, const int *idxA, const int *offA,
                                  const int *nA, int rounds, long long round_tile_stride,
                                  double *tiles, cudaStream_t s);

// ellipsoid membership (+ optional candidate transform and compaction of the survivors)
struct PrepArgs {
    const double *pts;        // (m x d) candidates, u-space
    long long m;
    int d;
    const double *center;     // ellipsoid (NULL: every row passes)
    double center_arg[32];    // register kernel (d <= 32): the centre travels as a kernel argument --
                              // the integrator moves it every iteration (integrator.py:2756), and a
                              // kernel argument needs neither an upload nor a stream synchronisation
    const double *invcov;
    double r2;
    unsigned char *mask;      // out: ellipsoid mask (1 byte per row); NULL to skip
    int layer_kind;           // -1: no transform/compaction stage
    const double *shift;
    const double *mat;
    double *tcand;            // out: compacted transformed rows
    int *items;               // out: original row of each compacted row
    int *n_items;             // in/out: device counter (must be zeroed)
    // register kernel (d <= 32) only:
    int use_constants;        // ellipsoid / layer parameters come from __constant__ memory
    double ell_tol_scale;     // 2 (d^2+2d+8) u ||invcov||_F: band half-width per unit |delta|^2 (2x the derived bound)
    double *like;             // out: fused likelihood of the rows inside the ellipsoid (nullable)
    int loglike_kind;
    const double *lparams;    // device likelihood parameter block
    // tile kernel (32 < d): zero-padded row-major copies, row stride pad_stride (multiple of 8)
    int pad_stride;           // 0: tile kernel not requested
    const double *invcov_pad;   // FOLDED inverse covariance: A_jj on, A_jc + A_cj above, 0 below the diagonal
    const double *mat_pad;
};
bool unb_tile_prep_fits(int d);
int unb_prep_sync_constants(unb_ctx *ctx, cudaStream_t s);
size_t unb_const_maxd();
void unb_prep_forget_ctx(const unb_ctx *ctx);
int unb_launch_prep(unb_ctx *ctx, const PrepArgs &p, cudaStream_t s);
// tail of the fused refill (integrator.py:1790-1805): user transform, tregion, likelihood, Lmin
struct TailArgs {
    const double *pts;           // (m x d) proposals, u-space
    long long m;
    int d;
    unsigned char *flags;        // in: region membership byte (ignored when !have_mask); out: flag bits
    int have_mask;
    int check_cube;
    int xform_kind;
    const double *xform_scale, *xform_lo;
    const double *treg_center, *treg_invcov;   // nullable
    double treg_r2;
    int loglike_kind;
    const double *lparams;
    double Lmin;
    double *like;                // out
    int *counts;                 // [3] member / tregion / accepted (atomically incremented)
};
int unb_launch_refill_tail(unb_ctx *ctx, const TailArgs &t, cudaStream_t s);
int unb_launch_transform(unb_ctx *ctx, int kind, bool inverse, const double *in, long long m,
                         int d, const double *shift, const double *mat, double *out,
                         cudaStream_t s);
int unb_launch_loglike(unb_ctx *ctx, int kind, const double *params, int d, long long n,
                       double *like, const unsigned char *mask, const double *lparams_dev,
                       cudaStream_t s);
int unb_launch_enlargement_f(unb_ctx *ctx, const double *u, int d, const int *item_idx,
                             const int *round_item_off, const int *round_nitems,
                             int max_items, const double *ctrs, const double *invcovs,
                             int rounds, unsigned long long *out_round_key, cudaStream_t s);
int unb_launch_pairdist(unb_ctx *ctx, const double *pts, const long long *ids, int n, int d,
                        double *partial_sum, long long *partial_cnt, cudaStream_t s);
size_t unb_max_rowwise_d();
// one-launch in-place row update of every image of the live block (unb_live.cu)
struct LiveUpdateArgs {
    const double *rows;       // row-major fp64 rows (already patched)
    const int *idx;           // rows to refresh
    int nrows, n, d, dr, tile_n;
    double *tiles;            // fp64 tiles (nullable)
    double *norms;
    unsigned long long *namax;   // running UPPER BOUND of the largest squared norm (atomicMax)
    int h_mode;               // HMODE_* currently stored in the fp64 tiles' h row (HMODE_NONE: skip)
    double h_r2, kappa;
    float *tiles32;           // fp32 tiles in row order (nullable)
    double t32_r2, kappa32;
};
int unb_launch_live_update(unb_ctx *ctx, const LiveUpdateArgs &a, cudaStream_t s);
int unb_launch_round_moments(unb_ctx *ctx, const double *u, int d, const int *idxA, const int *offA,
                             const int *nA, int max_rows, int rounds, const double *c0, double *sums,
                             double *sxx, cudaStream_t s);
// clustered live tiles + binned proposals (unb_cluster.cu)
size_t unb_cluster_max_tiles();
int unb_launch_cluster_live(unb_ctx *ctx, const double *rows, int n, int d, int K, int *perm,
                            int *scratch_i, float *scratch_f, cudaStream_t s);
int unb_launch_tile_centroids(unb_ctx *ctx, const double *rows, const int *perm, int n, int d, int dr,
                              int ntiles, float *ctiles, cudaStream_t s);
int unb_launch_bin_items(unb_ctx *ctx, const double *cand, int d, int dr, const int *item_idx,
                         const int *n_items_dev, long long n_items, const float *ctiles, int ntiles,
                         int *bin_of, int *order, int *meta, cudaStream_t s);
// device-side proposal generation (unb_sample.cu)
int unb_launch_draw(unb_ctx *ctx, int method, long long m, int d, unsigned long long seed,
                    unsigned long long offset, const double *center_dev, const double *axes_T_dev,
                    double scale, double *out_dev, unsigned char *cube_dev, cudaStream_t s);
int unb_launch_finish_mask(unb_ctx *ctx, unsigned char *mask_dev, const unsigned char *cube_dev,
                           const double *like_dev, double Lmin, bool use_lmin, long long m,
                           cudaStream_t s);
int unb_launch_compact_rows(unb_ctx *ctx, const unsigned char *mask_dev, long long m, int d,
                            const double *rows_dev, const double *like_dev, int *scratch_counts,
                            int *total_dev, double *out_rows_dev, double *out_like_dev,
                            int *out_index_dev, cudaStream_t s);
size_t unb_compact_scratch_ints(long long m);
int unb_launch_fp64_peak(unb_ctx *ctx, double *scratch, int blocks, int iters, cudaStream_t s);
int unb_launch_fp32_peak(unb_ctx *ctx, float *scratch, int blocks, int iters, cudaStream_t s);

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}

// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem,
                                             uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// order-preserving map non-negative double <-> uint64 (for atomicMax on doubles >= 0)
__device__ __forceinline__ unsigned long long f64_bits(double x)
{
    return static_cast<unsigned long long>(__double_as_longlong(x));
}

// total-order key for any finite double (atomicMax on possibly negative values)
__device__ __forceinline__ unsigned long long f64_key(double x)
{
    unsigned long long b = static_cast<unsigned long long>(__double_as_longlong(x));
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// the reference's k-sequential, non-fused squared distance (mlfriends.pyx:178-180).
// The whole library is compiled with -fmad=false; the intrinsics make the intent explicit.
__device__ __forceinline__ double sq_step(double acc, double a, double b)
{
    double diff = __dsub_rn(a, b);
    return __dadd_rn(acc, __dmul_rn(diff, diff));
}

// np.einsum('ij,jk,ik->i', delta, A, delta) of the reference (mlfriends.pyx:910, 1062, 1432):
// acc += (delta_j * A_jk) * delta_k, j outer, k inner, no FMA -- and, because NumPy reduces through
// its BUFFERED iterator (np.getbufsize() = 8192 elements), the partial sum restarts every
// floor(8192 / d) rows j and the partials are added to the result in order.  For d <= 90 that is
// the plain sequential sum; beyond, the chunking changes the last bits (probed against the compiled
// reference at d = 91 ... 150, tests/test_oracle_vs_reference.py).  `A` may be global or constant.
__device__ __forceinline__ double einsum_quadform(const double *delta, const double *__restrict__ A,
                                                  int d)
{
    int rows = 8192 / d;
    if (rows < 1) rows = 1;
    double total = 0.0;
    for (int j0 = 0; j0 < d; j0 += rows) {
        const int j1 = (j0 + rows < d) ? j0 + rows : d;
        double acc = 0.0;
        for (int jj = j0; jj < j1; jj++) {
            const double dj = delta[jj];
            const double *Arow = A + (size_t)jj * d;
            for (int k = 0; k < d; k++)
                acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(dj, __ldg(Arow + k)), delta[k]));
        }
        total = __dadd_rn(total, acc);
    }
    return total;
}

// Packed single-precision FMA (Blackwell FFMA2): d.x = a.x*b.x + c.x, d.y = a.y*b.y + c.y, each
// lane rounded to nearest once, i.e. exactly two fmaf().  One issue slot for two FMAs; a scalar
// broadcast make_float2(s, s) folds into the instruction's .F32 operand form (no extra MOV).
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c)
{
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long *>(&a);
    rb = *reinterpret_cast<unsigned long long *>(&b);
    rc = *reinterpret_cast<unsigned long long *>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// an exact membership decision that the reference could take differently (see ScanArgs::unc_tau)
__device__ __forceinline__ void unc_note(const ScanArgs &A, double D)
{
    if (A.unc_count != nullptr && fabs(__dsub_rn(D, A.r2)) <= A.unc_tau) atomicAdd(A.unc_count, 1u);
}

#endif  // __CUDACC__
