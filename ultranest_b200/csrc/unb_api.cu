// unb_api.cu -- extern "C" entry points declared in include/ultranest_b200.h: context
// management, host<->device staging (pinned, double-buffered lanes) and the orchestration
// of the kernels in unb_scan.cu / unb_region.cu.  No compute happens on the host here.
#include "unb_internal.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <new>
#include <thread>

// ---------------------------------------------------------------------------------------
// infrastructure
// ---------------------------------------------------------------------------------------
int unb_fail(unb_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

int unb_reserve(unb_ctx *ctx, DevBuf &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return UNB_OK;
    if (b.p) {
        // buffers may still be in use by enqueued work of either lane
        UNB_CUDA(ctx, cudaDeviceSynchronize());
        UNB_CUDA(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = bytes + bytes / 4;
    cap = (cap + 255) / 256 * 256;
    cudaError_t e = cudaMalloc(&b.p, cap);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return unb_fail(ctx, UNB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
    }
    b.cap = cap;
    return UNB_OK;
}

int unb_reserve_pinned(unb_ctx *ctx, PinBuf &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return UNB_OK;
    if (b.p) {
        UNB_CUDA(ctx, cudaDeviceSynchronize());
        UNB_CUDA(ctx, cudaFreeHost(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    size_t cap = bytes + bytes / 4;
    cudaError_t e = cudaMallocHost(&b.p, cap);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return unb_fail(ctx, UNB_ERR_NOMEM, "cudaMallocHost(%zu) failed: %s", cap, cudaGetErrorString(e));
    }
    b.cap = cap;
    return UNB_OK;
}

// pin_small may still be read by an asynchronous copy of an earlier call (row patches of the mirror)
int unb_wait_small(unb_ctx *ctx)
{
    if (ctx->ev_small_pending) {
        UNB_CUDA(ctx, cudaEventSynchronize(ctx->ev_small));
        ctx->ev_small_pending = false;
    }
    return UNB_OK;
}

namespace {

void free_dev(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}
void free_pin(PinBuf &b)
{
    if (b.p) cudaFreeHost(b.p);
    b.p = nullptr;
    b.cap = 0;
}
void free_live(LiveTiles &L)
{
    free_dev(L.tiles); free_dev(L.rows); free_dev(L.norms); free_dev(L.namax); free_dev(L.tiles32);
    free_dev(L.perm32); free_dev(L.ctiles32); free_dev(L.cl_scratch_i); free_dev(L.cl_scratch_f);
    L.valid = false;
}

inline cudaStream_t S0(unb_ctx *ctx) { return ctx->lane[0].stream; }

int h2d(unb_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    if (!bytes) return UNB_OK;
    UNB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    ctx->h2d_bytes += (long long)bytes;
    return UNB_OK;
}
int d2h(unb_ctx *ctx, void *dst, const void *src, size_t bytes, cudaStream_t s)
{
    if (!bytes) return UNB_OK;
    UNB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    ctx->d2h_bytes += (long long)bytes;
    return UNB_OK;
}

// pageable -> pinned staging copy on several host threads (one thread moves ~10 GB/s, the PCIe
// link takes 46 GB/s, so a single memcpy would be the bottleneck of the host pipeline)
void par_memcpy(void *dst, const void *src, size_t bytes)
{
    const size_t min_slice = 4u << 20;
    unsigned hw = std::thread::hardware_concurrency();
    size_t nthreads = std::min<size_t>(hw ? std::min(hw, 8u) : 4u, bytes / min_slice);
    if (nthreads <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> pool;
    const size_t slice = (bytes / nthreads + 63) / 64 * 64;
    for (size_t t = 1; t < nthreads; t++) {
        const size_t off = t * slice;
        if (off >= bytes) break;
        const size_t len = std::min(slice, bytes - off);
        pool.emplace_back([=]() { memcpy((char *)dst + off, (const char *)src + off, len); });
    }
    memcpy(dst, src, std::min(slice, bytes));
    for (auto &th : pool) th.join();
}

bool host_is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int check_ctx(unb_ctx *ctx)
{
    if (!ctx) return UNB_ERR_ARG;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return unb_fail(ctx, UNB_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return UNB_OK;
}

int stat_reset(unb_ctx *ctx, cudaStream_t s)
{
    UNB_TRY(unb_reserve(ctx, ctx->stat, 2 * sizeof(unsigned long long)));   // rechecks, tile visits
    UNB_CUDA(ctx, cudaMemsetAsync(ctx->stat.p, 0, 2 * sizeof(unsigned long long), s));
    return UNB_OK;
}
int stat_fetch(unb_ctx *ctx, cudaStream_t s)
{
    unsigned long long v = 0;
    UNB_CUDA(ctx, cudaMemcpyAsync(&v, ctx->stat.p, sizeof(v), cudaMemcpyDeviceToHost, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    ctx->last_rechecks = (long long)v;
    return UNB_OK;
}

// The tolerance only means something next to the radius it was derived for: it is dropped when the
// radius changes (unb_region_set_radius) and ignored unless it is tiny against the radius (the
// certain-neighbour margin of the fp32 filter has 1e-6 r2 of slack to absorb it).
double eff_tau(const unb_ctx *ctx)
{
    const RegionState &R = ctx->region;
    return (R.have_radius && R.unc_tau > 0.0 && R.unc_tau <= 1e-7 * R.r2) ? R.unc_tau : 0.0;
}

// counter of uncertain exact decisions (transform tolerance): zeroed before a membership call ...
int unc_reset(unb_ctx *ctx, cudaStream_t s)
{
    // the device counter is monotone (zeroed once); a call's count is the increase it observes,
    // which saves a memset per call on the latency-critical per-iteration path
    if (!ctx->unc.p) {
        UNB_TRY(unb_reserve(ctx, ctx->unc, 2 * sizeof(unsigned int)));
        UNB_CUDA(ctx, cudaMemsetAsync(ctx->unc.p, 0, 2 * sizeof(unsigned int), s));
        ctx->unc_seen = 0;
    }
    ctx->last_uncertain = 0;
    return UNB_OK;
}
// ... and read back without a synchronisation of its own: every lane copies the (monotone) counter
// into its pinned slot after its last kernel; the later of the two copies has seen everything, so
// after the call's final stream synchronisations the larger value is the total
bool unc_active(const unb_ctx *ctx)
{
    return ctx->unc.p && eff_tau(ctx) > 0.0 && ctx->region.layer_kind == UNB_LAYER_AFFINE;
}
int unc_copy(unb_ctx *ctx, int lane)
{
    if (!unc_active(ctx)) return UNB_OK;
    Lane &ln = ctx->lane[lane];
    UNB_TRY(unb_reserve_pinned(ctx, ln.pin_n, 4 * sizeof(int)));
    return d2h(ctx, (unsigned int *)ln.pin_n.p + 2, ctx->unc.p, sizeof(unsigned int), ln.stream);
}
void unc_collect(unb_ctx *ctx, bool lane0, bool lane1)
{
    if (!unc_active(ctx)) return;
    unsigned int v = 0;
    if (lane0 && ctx->lane[0].pin_n.p) v = std::max(v, ((const unsigned int *)ctx->lane[0].pin_n.p)[2]);
    if (lane1 && ctx->lane[1].pin_n.p) v = std::max(v, ((const unsigned int *)ctx->lane[1].pin_n.p)[2]);
    ctx->last_uncertain = (long long)(unsigned int)(v - ctx->unc_seen);
    ctx->unc_seen = v;
}

ScanArgs scan_args_for(const LiveTiles &L)
{
    ScanArgs a;
    memset(&a, 0, sizeof(a));
    a.tiles = L.ntiles ? (const double *)L.tiles.p : nullptr;
    a.live_rows = (const double *)L.rows.p;
    a.n_live = (int)L.n;
    a.tile_n = (int)L.tile_n;
    a.d = (int)L.d;
    a.dr = (int)L.dr;
    a.kappa = unb_kappa(L.d);
    a.namax_bits = (const unsigned long long *)L.namax.p;
    return a;
}

// largest squared norm of the live block for the certain-neighbour level of the fp32 membership
// filter (unb_scan.cu, tile_filter32); UNB_OPT_SURE_LEVEL = 0 switches the shortcut off (tests)
double sure_namax(const unb_ctx *ctx, const LiveTiles &L)
{
    return ctx->sure_level ? L.namax_host : (double)INFINITY;
}

// The radius the FILTERS of a live block are built for.  For the region's block it is widened by
// the transform tolerance (ScanArgs::unc_tau): a pair whose device distance is within that
// tolerance ABOVE the radius must still reach the exact decision, where it is reported as
// uncertain.  Decisions themselves always compare with the true radius.
double filter_r2(const unb_ctx *ctx, const LiveTiles &L, double r2)
{
    return (&L == &ctx->region.live && r2 == ctx->region.r2) ? r2 + eff_tau(ctx) : r2;
}

// threshold-mode h row + (when safe and enabled) the fp32 image used by the membership kernel
int prepare_threshold(unb_ctx *ctx, LiveTiles &L, double r2_true, cudaStream_t s, ScanArgs *a)
{
    const double r2 = filter_r2(ctx, L, r2_true);
    UNB_TRY(unb_live_set_h(ctx, L, HMODE_THRESH, r2, s));
    bool ok32 = false;
    UNB_TRY(unb_live_prepare32(ctx, L, r2, &ok32, s));
    if (!ok32) L.t32_valid = false;   // never leave a stale image that looks usable
    if (a && ok32) {
        a->tiles32 = (const float *)L.tiles32.p;
        a->perm32 = L.t32_clustered ? (const int *)L.perm32.p : nullptr;
        a->kappa32 = unb_kappa32(L.d);
        a->namax32 = sure_namax(ctx, L);
    }
    return UNB_OK;
}

// large launches against clustered fp32 tiles: bin the work items by nearest tile centroid
// (unb_cluster.cu) so that the membership kernel starts every proposal where hits are likely
int attach_bins(unb_ctx *ctx, Lane &ln, const LiveTiles &L, ScanArgs &a, cudaStream_t s)
{
    if (!a.tiles32 || !a.perm32 || !L.t32_clustered || ctx->exact_only || ctx->bin_min_rows <= 0 ||
        a.n_items < ctx->bin_min_rows || L.dr > 32 || a.out_idx)
        return UNB_OK;
    const int ntiles = (int)((L.n + 63) / 64);
    UNB_TRY(unb_reserve(ctx, ln.bin_of, (size_t)a.n_items * sizeof(int)));
    UNB_TRY(unb_reserve(ctx, ln.bin_order, (size_t)a.n_items * sizeof(int)));
    UNB_TRY(unb_reserve(ctx, ln.bin_meta, (size_t)(4 * ntiles + 2) * sizeof(int)));
    int *meta = (int *)ln.bin_meta.p;
    UNB_TRY(unb_launch_bin_items(ctx, a.cand, a.d, a.dr, a.item_idx, a.n_items_dev, a.n_items,
                                 (const float *)L.ctiles32.p, ntiles, (int *)ln.bin_of.p,
                                 (int *)ln.bin_order.p, meta, s));
    a.bin_order = (const int *)ln.bin_order.p;
    a.bin_start = meta + ntiles;
    a.bin_head = meta + ntiles + (ntiles + 1) + ntiles;
    return UNB_OK;
}

// First-neighbour indices in TWO PHASES when the fp32 membership kernel is available
// (find_nearby, mlfriends.pyx:143-183): (1) the any-neighbour kernel marks the candidates that have
// a neighbour at all -- every other candidate is finished with -1 and never enters the ordered
// scan; (2) the ordered first-index kernel runs over the members only.  In the full-scan regime
// (no neighbours) FIND then costs what the membership test costs (3.8 instead of 6.5 ms per 2^20
// candidates at d = 20, 3.6 instead of 10.8 ms per 2^17 at d = 100); with neighbours everywhere
// it pays the membership pass on top (+12 %).  Indices are those of the ordered exact kernel.
int find_two_phase(unb_ctx *ctx, Lane &ln, ScanArgs a, cudaStream_t s, bool *done)
{
    *done = false;
    if (!a.tiles32 || ctx->exact_only || !a.out_idx || a.item_idx || a.out_row_idx || a.n_items_dev ||
        a.n_items < 4096 || a.dr > 128)
        return UNB_OK;
    const size_t n = (size_t)a.n_items;
    UNB_TRY(unb_reserve(ctx, ln.mask, n));
    UNB_TRY(unb_reserve(ctx, ln.items, n * sizeof(int)));
    UNB_TRY(unb_reserve(ctx, ln.counter, 4 * sizeof(int)));
    UNB_TRY(unb_reserve(ctx, ln.smp_counts, unb_compact_scratch_ints((long long)n) * sizeof(int)));
    UNB_CUDA(ctx, cudaMemsetAsync(ln.counter.p, 0, 4 * sizeof(int), s));
    UNB_CUDA(ctx, cudaMemsetAsync(a.out_idx, 0xff, n * sizeof(long long), s));   // -1 everywhere
    ScanArgs any = a;
    any.out_idx = nullptr;
    unsigned char *user_mask = a.out_mask;
    any.out_mask = (unsigned char *)ln.mask.p;
    any.stat_rechecks = nullptr;
    UNB_TRY(unb_launch_inside_any(ctx, any, (int *)ln.counter.p + 1, s));
    int *n_members = (int *)ln.counter.p + 2;
    UNB_TRY(unb_launch_compact_rows(ctx, (const unsigned char *)ln.mask.p, (long long)n, 0, nullptr,
                                    nullptr, (int *)ln.smp_counts.p, n_members, nullptr, nullptr,
                                    (int *)ln.items.p, s));
    if (user_mask)
        UNB_CUDA(ctx, cudaMemcpyAsync(user_mask, ln.mask.p, n, cudaMemcpyDeviceToDevice, s));
    ScanArgs ord = a;
    ord.tiles32 = nullptr;
    ord.out_mask = nullptr;
    ord.item_idx = (const int *)ln.items.p;
    ord.out_row_idx = (const int *)ln.items.p;
    ord.n_items_dev = n_members;
    UNB_TRY(unb_launch_scan(ctx, SCAN_FIND, ord, 1, s));
    *done = true;
    return UNB_OK;
}

// a launch of `m` proposals against the region's live block: ask for clustered tiles when large
void request_cluster(unb_ctx *ctx, LiveTiles &L, size_t m)
{
    if (ctx->bin_min_rows > 0 && (long long)m >= ctx->bin_min_rows) L.want_cluster = true;
}

// upload a live block into L and build its tiles
int live_from_host(unb_ctx *ctx, LiveTiles &L, const double *rows, size_t n, size_t d,
                   cudaStream_t s)
{
    UNB_TRY(unb_reserve(ctx, L.rows, (n ? n : 1) * d * sizeof(double)));
    UNB_TRY(h2d(ctx, L.rows.p, rows, n * d * sizeof(double), s));
    return unb_live_build(ctx, L, (const double *)L.rows.p, n, d, s);
}

int check_dims(unb_ctx *ctx, size_t n, size_t d)
{
    if (d == 0) return unb_fail(ctx, UNB_ERR_ARG, "ndim must be >= 1");
    if (d > (1u << 20)) return unb_fail(ctx, UNB_ERR_ARG, "ndim too large");
    if (n > 0x7fffffffULL) return unb_fail(ctx, UNB_ERR_ARG, "too many rows (max 2^31-1 per call)");
    return UNB_OK;
}

// generic host-buffer pair scan against an already built live block
int scan_host(unb_ctx *ctx, LiveTiles &L, int mode, const double *bpts, size_t nb, double r2,
              long long *out_idx_host, double *out_rows_host, double *out_max_host)
{
    Lane &ln = ctx->lane[0];
    cudaStream_t s = ln.stream;
    const size_t d = L.d;
    UNB_TRY(unb_live_set_h(ctx, L, mode == SCAN_MIN ? HMODE_MIN : HMODE_THRESH, filter_r2(ctx, L, r2), s));
    UNB_TRY(unb_reserve(ctx, ln.cand, nb * d * sizeof(double)));
    UNB_TRY(h2d(ctx, ln.cand.p, bpts, nb * d * sizeof(double), s));
    UNB_TRY(stat_reset(ctx, s));
    ScanArgs a = scan_args_for(L);
    a.cand = (const double *)ln.cand.p;
    a.n_items = (long long)nb;
    a.r2 = r2;
    a.stat_rechecks = (unsigned long long *)ctx->stat.p;
    if (mode == SCAN_FIND || mode == SCAN_COUNT) {
        UNB_TRY(unb_reserve(ctx, ln.idx, nb * sizeof(long long)));
        a.out_idx = (long long *)ln.idx.p;
    } else if (mode == SCAN_SUBTRACT) {
        UNB_TRY(unb_reserve(ctx, ln.tcand, nb * d * sizeof(double)));
        a.out_rows = (double *)ln.tcand.p;
    } else {
        UNB_TRY(unb_reserve(ctx, ctx->aux0, sizeof(unsigned long long)));
        UNB_CUDA(ctx, cudaMemsetAsync(ctx->aux0.p, 0, sizeof(unsigned long long), s));
        a.out_round_max = (unsigned long long *)ctx->aux0.p;
    }
    bool two_phase = false;
    if (mode == SCAN_FIND) {
        ScanArgs pre;
        memset(&pre, 0, sizeof(pre));
        UNB_TRY(prepare_threshold(ctx, L, r2, s, &pre));   // fp32 image when safe for this radius
        ScanArgs a2 = a;
        a2.tiles32 = pre.tiles32;
        a2.perm32 = pre.perm32;
        a2.kappa32 = pre.kappa32;
        a2.namax32 = pre.namax32;
        UNB_TRY(find_two_phase(ctx, ln, a2, s, &two_phase));
    }
    if (!two_phase) UNB_TRY(unb_launch_scan(ctx, mode, a, 1, s));
    if (mode == SCAN_FIND || mode == SCAN_COUNT)
        UNB_TRY(d2h(ctx, out_idx_host, ln.idx.p, nb * sizeof(long long), s));
    else if (mode == SCAN_SUBTRACT)
        UNB_TRY(d2h(ctx, out_rows_host, ln.tcand.p, nb * d * sizeof(double), s));
    else
        UNB_TRY(d2h(ctx, out_max_host, ctx->aux0.p, sizeof(double), s));
    return stat_fetch(ctx, s);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------
extern "C" int unb_abi_version(void) { return UNB_ABI_VERSION; }

extern "C" int unb_ctx_create(int device, unb_ctx **out)
{
    if (!out) return UNB_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) return UNB_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return UNB_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return UNB_ERR_CUDA;
    if (prop.major != 10) {
        fprintf(stderr, "ultranest_b200: device %d is sm_%d%d; this library is built for sm_100a only\n",
                device, prop.major, prop.minor);
        return UNB_ERR_CUDA;
    }
    unb_ctx *ctx = new (std::nothrow) unb_ctx();
    if (!ctx) return UNB_ERR_NOMEM;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    // clustered tiles + binned proposals are OPT-IN (UNB_OPT_BIN_MIN_ROWS): measured on B200 they
    // halve the tiles a proposal streams (192k -> 90k warp-tiles per 2^20 proposals) but the binning
    // passes (0.12 + 0.17 ms) cost more than the membership kernel gains (0.49 -> 0.40 ms); see
    // DESIGN.md 4.1 "Tried, measured, not enabled"
    ctx->bin_min_rows = 0;
    for (int i = 0; i < 2; i++) {
        Lane &ln = ctx->lane[i];
        if (cudaStreamCreateWithFlags(&ln.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ln.ev_in, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ln.ev_done, cudaEventDisableTiming) != cudaSuccess) {
            delete ctx;
            return UNB_ERR_CUDA;
        }
    }
    *out = ctx;
    return UNB_OK;
}

extern "C" int unb_ctx_destroy(unb_ctx *ctx)
{
    if (!ctx) return UNB_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    unb_prep_forget_ctx(ctx);
    for (int i = 0; i < 2; i++) {
        Lane &ln = ctx->lane[i];
        free_dev(ln.cand); free_dev(ln.tcand); free_dev(ln.items); free_dev(ln.counter);
        free_dev(ln.mask); free_dev(ln.idx); free_dev(ln.like);
        free_dev(ln.smp_cube); free_dev(ln.smp_counts); free_dev(ln.smp_rows); free_dev(ln.smp_like);
        free_dev(ln.smp_n);
        free_dev(ln.bin_of); free_dev(ln.bin_order); free_dev(ln.bin_meta);
        free_pin(ln.pin_in); free_pin(ln.pin_mask); free_pin(ln.pin_like); free_pin(ln.pin_idx);
        free_pin(ln.pin_n);
        if (ln.ev_in) cudaEventDestroy(ln.ev_in);
        if (ln.ev_done) cudaEventDestroy(ln.ev_done);
        if (ln.stream) cudaStreamDestroy(ln.stream);
    }
    free_live(ctx->region.live);
    free_live(ctx->scratch_live);
    free_dev(ctx->region.layer_shift); free_dev(ctx->region.layer_mat);
    free_dev(ctx->region.ell_center); free_dev(ctx->region.ell_invcov);
    free_dev(ctx->region.ell_invcov_pad); free_dev(ctx->region.layer_mat_pad);
    free_dev(ctx->aux0); free_dev(ctx->aux1); free_dev(ctx->aux2); free_dev(ctx->aux3);
    free_dev(ctx->stat); free_dev(ctx->unc); free_dev(ctx->lparams); free_dev(ctx->refill_params);
    free_dev(ctx->boot_rows); free_dev(ctx->boot_u); free_dev(ctx->boot_tiles);
    free_dev(ctx->boot_idx); free_dev(ctx->boot_meta); free_dev(ctx->boot_out);
    free_dev(ctx->boot_ell);
    free_pin(ctx->pin_small);
    if (ctx->ev_small) cudaEventDestroy(ctx->ev_small);
    free_dev(ctx->smp_axes); free_dev(ctx->smp_center);
    for (DevBuf &b : ctx->sf) free_dev(b);
    free_dev(ctx->sf_params);
    delete ctx;
    return UNB_OK;
}

extern "C" const char *unb_last_error(const unb_ctx *ctx)
{
    return ctx ? ctx->err.c_str() : "null context";
}

extern "C" int unb_ctx_set_option(unb_ctx *ctx, int option, int64_t value)
{
    if (!ctx) return UNB_ERR_ARG;
    switch (option) {
    case UNB_OPT_EXACT_ONLY: ctx->exact_only = value ? 1 : 0; return UNB_OK;
    case UNB_OPT_CHUNK_ROWS: ctx->chunk_rows = value > 0 ? value : 0; return UNB_OK;
    case UNB_OPT_FILTER_FP32: ctx->filter_fp32 = value ? 1 : 0; return UNB_OK;
    case UNB_OPT_SURE_LEVEL: ctx->sure_level = value ? 1 : 0; return UNB_OK;
    case UNB_OPT_COOP_MAX: ctx->coop_max = value < 0 ? 0 : (int)(value > 1 << 20 ? 1 << 20 : value); return UNB_OK;
    case UNB_OPT_BLOCK_KERNEL: ctx->block_kernel = value ? 1 : 0; return UNB_OK;
    case UNB_OPT_BIN_MIN_ROWS: ctx->bin_min_rows = value > 0 ? value : 0; return UNB_OK;
    default: return unb_fail(ctx, UNB_ERR_ARG, "unknown option %d", option);
    }
}

extern "C" int unb_ctx_get_stat(unb_ctx *ctx, int stat, int64_t *value)
{
    if (!ctx || !value) return UNB_ERR_ARG;
    switch (stat) {
    case UNB_STAT_KERNEL_LAUNCHES: *value = ctx->launches; return UNB_OK;
    case UNB_STAT_RECHECKS: *value = ctx->last_rechecks; return UNB_OK;
    case UNB_STAT_H2D_BYTES: *value = ctx->h2d_bytes; return UNB_OK;
    case UNB_STAT_D2H_BYTES: *value = ctx->d2h_bytes; return UNB_OK;
    case UNB_STAT_UNCERTAIN: *value = ctx->last_uncertain; return UNB_OK;
    case UNB_STAT_TILE_VISITS: {
        // refresh from the device counter (written by the any-neighbour kernel)
        unsigned long long v[2] = {0, 0};
        if (ctx->stat.p) {
            UNB_CUDA(ctx, cudaDeviceSynchronize());
            UNB_CUDA(ctx, cudaMemcpy(v, ctx->stat.p, sizeof(v), cudaMemcpyDeviceToHost));
        }
        *value = (long long)v[1];
        return UNB_OK;
    }
    default: return unb_fail(ctx, UNB_ERR_ARG, "unknown stat %d", stat);
    }
}

// measured DFMA issue rate of this device (warp-wide fused multiply-adds x 32 lanes per second)
extern "C" int unb_fp64_peak(unb_ctx *ctx, double *dfma_per_s)
{
    UNB_TRY(check_ctx(ctx));
    if (!dfma_per_s) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = S0(ctx);
    UNB_TRY(unb_reserve(ctx, ctx->aux0, 64));
    const int blocks = ctx->sm_count * 8, iters = 20000;
    cudaEvent_t e0, e1;
    UNB_CUDA(ctx, cudaEventCreate(&e0));
    UNB_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        UNB_CUDA(ctx, cudaEventRecord(e0, s));
        UNB_TRY(unb_launch_fp64_peak(ctx, (double *)ctx->aux0.p, blocks, iters, s));
        UNB_CUDA(ctx, cudaEventRecord(e1, s));
        UNB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        UNB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)blocks * 256.0 * 8.0 * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *dfma_per_s = best;
    return UNB_OK;
}

extern "C" int unb_fp32_peak_form(unb_ctx *ctx, int form, double *lane_fma_per_s)
{
    UNB_TRY(check_ctx(ctx));
    if (!lane_fma_per_s) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (form < 0 || form > 2) return unb_fail(ctx, UNB_ERR_ARG, "unknown instruction form %d", form);
    cudaStream_t s = S0(ctx);
    UNB_TRY(unb_reserve(ctx, ctx->aux0, 64));
    const int blocks = ctx->sm_count * 8, iters = form == 0 ? 40000 : 20000;
    const double per_iter = form == 0 ? 8.0 : 16.0;   // lane-FMAs per thread and iteration
    cudaEvent_t e0, e1;
    UNB_CUDA(ctx, cudaEventCreate(&e0));
    UNB_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        UNB_CUDA(ctx, cudaEventRecord(e0, s));
        UNB_TRY(unb_launch_fp32_peak(ctx, (float *)ctx->aux0.p, form == 2 ? -blocks : blocks,
                                     form == 0 ? iters : -iters, s));
        UNB_CUDA(ctx, cudaEventRecord(e1, s));
        UNB_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.f;
        UNB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double rate = (double)blocks * 256.0 * per_iter * iters / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *lane_fma_per_s = best;
    return UNB_OK;
}

// the fp32 roofline denominator: the best rate over the instruction forms (the membership filter
// issues FFMA2 with a scalar multiplicand; a device where the packed form is no faster reports
// the scalar rate)
extern "C" int unb_fp32_peak(unb_ctx *ctx, double *ffma_per_s)
{
    UNB_TRY(check_ctx(ctx));
    if (!ffma_per_s) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    double best = 0.0;
    for (int form = 0; form < 3; form++) {
        double r = 0.0;
        UNB_TRY(unb_fp32_peak_form(ctx, form, &r));
        if (r > best) best = r;
    }
    *ffma_per_s = best;
    return UNB_OK;
}

extern "C" int unb_ctx_synchronize(unb_ctx *ctx)
{
    UNB_TRY(check_ctx(ctx));
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
    return UNB_OK;
}

// ---------------------------------------------------------------------------------------
// stateless scans
// ---------------------------------------------------------------------------------------
extern "C" int unb_find_nearby(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                               size_t nb, size_t ndim, double radiussq, int64_t *nnearby)
{
    UNB_RANGE("unb_find_nearby");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, na > nb ? na : nb, ndim));
    if (nb == 0) return UNB_OK;
    if (!bpts || !nnearby || (na && !apts)) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (na == 0) {   // no members: every entry is -1 (mlfriends.pyx:176)
        for (size_t j = 0; j < nb; j++) nnearby[j] = -1;
        return UNB_OK;
    }
    UNB_TRY(live_from_host(ctx, ctx->scratch_live, apts, na, ndim, S0(ctx)));
    return scan_host(ctx, ctx->scratch_live, SCAN_FIND, bpts, nb, radiussq,
                     (long long *)nnearby, nullptr, nullptr);
}

extern "C" int unb_count_nearby(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                                size_t nb, size_t ndim, double radiussq, int64_t *nnearby)
{
    UNB_RANGE("unb_count_nearby");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, na > nb ? na : nb, ndim));
    if (nb == 0) return UNB_OK;
    if (!bpts || !nnearby || (na && !apts)) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (na == 0) {
        for (size_t j = 0; j < nb; j++) nnearby[j] = 0;
        return UNB_OK;
    }
    UNB_TRY(live_from_host(ctx, ctx->scratch_live, apts, na, ndim, S0(ctx)));
    return scan_host(ctx, ctx->scratch_live, SCAN_COUNT, bpts, nb, radiussq,
                     (long long *)nnearby, nullptr, nullptr);
}

extern "C" int unb_subtract_nearby(unb_ctx *ctx, const double *apts, size_t n, size_t ndim,
                                   double radiussq, double *bpts_out)
{
    UNB_RANGE("unb_subtract_nearby");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    if (n == 0) return UNB_OK;
    if (!apts || !bpts_out) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    UNB_TRY(live_from_host(ctx, ctx->scratch_live, apts, n, ndim, S0(ctx)));
    return scan_host(ctx, ctx->scratch_live, SCAN_SUBTRACT, apts, n, radiussq, nullptr, bpts_out,
                     nullptr);
}

extern "C" int unb_compute_maxradiussq(unb_ctx *ctx, const double *apts, size_t na,
                                       const double *bpts, size_t nb, size_t ndim,
                                       double *maxd_out)
{
    UNB_RANGE("unb_compute_maxradiussq");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, na > nb ? na : nb, ndim));
    if (!maxd_out) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (nb == 0) {   // maxd stays 0 (mlfriends.pyx:209)
        *maxd_out = 0.0;
        return UNB_OK;
    }
    if (na == 0) {   // mind stays 1e300 for every b; (float)1e300 = inf
        *maxd_out = (double)(float)1e300;
        return UNB_OK;
    }
    UNB_TRY(live_from_host(ctx, ctx->scratch_live, apts, na, ndim, S0(ctx)));
    double maxd = 0.0;
    UNB_TRY(scan_host(ctx, ctx->scratch_live, SCAN_MIN, bpts, nb, 0.0, nullptr, nullptr, &maxd));
    *maxd_out = (double)(float)maxd;   // C `float` return of the reference (mlfriends.pyx:188)
    return UNB_OK;
}

extern "C" int unb_mean_pair_distance(unb_ctx *ctx, const double *pts, const int64_t *clusterids,
                                      size_t n, size_t ndim, double *out)
{
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    if (!out || (n && (!pts || !clusterids))) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = S0(ctx);
    UNB_TRY(unb_reserve(ctx, ctx->aux0, n * ndim * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ctx->aux1, n * sizeof(long long)));
    UNB_TRY(unb_reserve(ctx, ctx->aux2, n * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ctx->aux3, n * sizeof(long long)));
    UNB_TRY(h2d(ctx, ctx->aux0.p, pts, n * ndim * sizeof(double), s));
    UNB_TRY(h2d(ctx, ctx->aux1.p, clusterids, n * sizeof(long long), s));
    UNB_TRY(unb_launch_pairdist(ctx, (const double *)ctx->aux0.p, (const long long *)ctx->aux1.p,
                                (int)n, (int)ndim, (double *)ctx->aux2.p, (long long *)ctx->aux3.p, s));
    std::vector<double> ps(n);
    std::vector<long long> pc(n);
    UNB_TRY(d2h(ctx, ps.data(), ctx->aux2.p, n * sizeof(double), s));
    UNB_TRY(d2h(ctx, pc.data(), ctx->aux3.p, n * sizeof(long long), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    double total = 0.0;
    long long npairs = 0;
    for (size_t j = 0; j < n; j++) {   // final fold of the per-row device partials
        total += ps[j];
        npairs += pc[j];
    }
    *out = total / (double)npairs;
    return UNB_OK;
}

// zero-padded row-major copy (row stride = ndim rounded up to 8) for the tile prep kernel
static size_t pad8(size_t d) { return (d + 7) / 8 * 8; }
static int upload_padded8(unb_ctx *ctx, DevBuf &dst, const double *mat, size_t ndim, cudaStream_t s,
                          bool fold = false)
{
    const size_t dp = pad8(ndim);
    std::vector<double> buf(ndim * dp, 0.0);
    for (size_t r = 0; r < ndim; r++) {
        if (!fold) {
            memcpy(&buf[r * dp], mat + r * ndim, ndim * sizeof(double));
            continue;
        }
        // quadratic form folded onto the upper triangle (the ellipsoid filter of k_prep_tile)
        buf[r * dp + r] = mat[r * ndim + r];
        for (size_t c = r + 1; c < ndim; c++) buf[r * dp + c] = mat[r * ndim + c] + mat[c * ndim + r];
    }
    UNB_TRY(unb_reserve(ctx, dst, buf.size() * sizeof(double)));
    UNB_TRY(h2d(ctx, dst.p, buf.data(), buf.size() * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));   // buf is a temporary
    return UNB_OK;
}

extern "C" int unb_inside_ellipsoid(unb_ctx *ctx, const double *points, size_t m, size_t ndim,
                                    const double *center, const double *invcov,
                                    double square_radius, uint8_t *mask)
{
    UNB_RANGE("unb_inside_ellipsoid");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, m, ndim));
    if (m == 0) return UNB_OK;
    if (!points || !center || !invcov || !mask) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    Lane &ln = ctx->lane[0];
    cudaStream_t s = ln.stream;
    UNB_TRY(unb_reserve(ctx, ln.cand, m * ndim * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.mask, m));
    UNB_TRY(unb_reserve(ctx, ctx->aux0, ndim * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ctx->aux1, ndim * ndim * sizeof(double)));
    UNB_TRY(h2d(ctx, ln.cand.p, points, m * ndim * sizeof(double), s));
    UNB_TRY(h2d(ctx, ctx->aux0.p, center, ndim * sizeof(double), s));
    UNB_TRY(h2d(ctx, ctx->aux1.p, invcov, ndim * ndim * sizeof(double), s));
    PrepArgs p;
    memset(&p, 0, sizeof(p));
    p.pts = (const double *)ln.cand.p;
    p.m = (long long)m;
    p.d = (int)ndim;
    p.center = (const double *)ctx->aux0.p;
    p.invcov = (const double *)ctx->aux1.p;
    p.r2 = square_radius;
    p.mask = (unsigned char *)ln.mask.p;
    p.layer_kind = -1;
    if (!ctx->exact_only && unb_tile_prep_fits((int)ndim)) {
        UNB_TRY(upload_padded8(ctx, ctx->aux2, invcov, ndim, s, true));
        double fro = 0.0;
        for (size_t i = 0; i < ndim * ndim; i++) fro += invcov[i] * invcov[i];
        p.pad_stride = (int)pad8(ndim);
        p.invcov_pad = (const double *)ctx->aux2.p;
        p.ell_tol_scale = 2.0 * ((double)(ndim * ndim + 2 * ndim + 8)) * 1.1102230246251565e-16 *
                          std::sqrt(fro) * (1.0 + 1e-9);
    }
    UNB_TRY(unb_launch_prep(ctx, p, s));
    UNB_TRY(d2h(ctx, mask, ln.mask.p, m, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

namespace {
int transform_host(unb_ctx *ctx, int kind, bool inverse, const double *in, size_t m, size_t d,
                   const double *shift, size_t shift_n, const double *mat, size_t mat_n,
                   double *out)
{
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, m, d));
    if (m == 0) return UNB_OK;
    if (!in || !out || !shift || !mat) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    Lane &ln = ctx->lane[0];
    cudaStream_t s = ln.stream;
    UNB_TRY(unb_reserve(ctx, ln.cand, m * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.tcand, m * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ctx->aux0, shift_n * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ctx->aux1, mat_n * sizeof(double)));
    UNB_TRY(h2d(ctx, ln.cand.p, in, m * d * sizeof(double), s));
    UNB_TRY(h2d(ctx, ctx->aux0.p, shift, shift_n * sizeof(double), s));
    UNB_TRY(h2d(ctx, ctx->aux1.p, mat, mat_n * sizeof(double), s));
    UNB_TRY(unb_launch_transform(ctx, kind, inverse, (const double *)ln.cand.p, (long long)m, (int)d,
                                 (const double *)ctx->aux0.p, (const double *)ctx->aux1.p,
                                 (double *)ln.tcand.p, s));
    UNB_TRY(d2h(ctx, out, ln.tcand.p, m * d * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}
}  // namespace

extern "C" int unb_transform_scaling(unb_ctx *ctx, const double *w, size_t m, size_t ndim,
                                     const double *mean, const double *std, double *out)
{
    return transform_host(ctx, UNB_LAYER_SCALING, false, w, m, ndim, mean, ndim, std, ndim, out);
}
extern "C" int unb_untransform_scaling(unb_ctx *ctx, const double *ww, size_t m, size_t ndim,
                                       const double *mean, const double *std, double *out)
{
    return transform_host(ctx, UNB_LAYER_SCALING, true, ww, m, ndim, mean, ndim, std, ndim, out);
}
extern "C" int unb_transform_affine(unb_ctx *ctx, const double *w, size_t m, size_t ndim,
                                    const double *ctr, const double *T, double *out)
{
    return transform_host(ctx, UNB_LAYER_AFFINE, false, w, m, ndim, ctr, ndim, T, ndim * ndim, out);
}
extern "C" int unb_untransform_affine(unb_ctx *ctx, const double *ww, size_t m, size_t ndim,
                                      const double *ctr, const double *invT, double *out)
{
    return transform_host(ctx, UNB_LAYER_AFFINE, true, ww, m, ndim, ctr, ndim, invT, ndim * ndim, out);
}

// ---------------------------------------------------------------------------------------
// stateful region
// ---------------------------------------------------------------------------------------
extern "C" int unb_region_sync_live(unb_ctx *ctx, const double *unormed, size_t n, size_t ndim,
                                    int64_t *rows_changed)
{
    UNB_RANGE("unb_region_sync_live");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    if (!unormed || n == 0) return unb_fail(ctx, UNB_ERR_ARG, "empty live block");
    RegionState &R = ctx->region;
    cudaStream_t s = S0(ctx);
    const size_t rowb = ndim * sizeof(double);
    bool full = !R.live.valid || R.live.n != n || R.live.d != ndim || R.snapshot.size() != n * ndim;
    std::vector<int> changed;
    if (!full) {
        const char *cur = (const char *)unormed;
        const char *old = (const char *)R.snapshot.data();
        if (memcmp(cur, old, n * rowb) != 0) {
            for (size_t i = 0; i < n; i++)
                if (memcmp(cur + i * rowb, old + i * rowb, rowb) != 0) changed.push_back((int)i);
            if (changed.size() > n / 4) full = true;
        }
    }
    if (full) {
        // both lanes may still read the old mirror
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        R.snapshot.assign(unormed, unormed + n * ndim);
        UNB_TRY(live_from_host(ctx, R.live, unormed, n, ndim, s));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));
        if (rows_changed) *rows_changed = (int64_t)n;
        return UNB_OK;
    }
    if (!changed.empty()) {
        // Everything below is stream-ordered on lane 0; nothing waits for the device: the pinned
        // staging block is guarded by an event (waited for by its next writer), and the host's
        // bound of the largest squared norm follows the patched rows without a read-back.
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        UNB_TRY(unb_wait_small(ctx));
        const size_t k = changed.size();
        UNB_TRY(unb_reserve_pinned(ctx, ctx->pin_small, k * (rowb + sizeof(int))));
        char *stage = (char *)ctx->pin_small.p;
        int *idx_stage = (int *)(stage + k * rowb);
        for (size_t r = 0; r < k; r++) {
            const size_t i = (size_t)changed[r];
            memcpy(stage + r * rowb, unormed + i * ndim, rowb);
            memcpy(R.snapshot.data() + i * ndim, unormed + i * ndim, rowb);
            idx_stage[r] = changed[r];
            unb_live_note_host_row(R.live, unormed + i * ndim);
            UNB_TRY(h2d(ctx, (char *)R.live.rows.p + i * rowb, stage + r * rowb, rowb, s));
        }
        UNB_TRY(unb_reserve(ctx, ctx->aux3, k * sizeof(int)));
        UNB_TRY(h2d(ctx, ctx->aux3.p, idx_stage, k * sizeof(int), s));
        UNB_TRY(unb_live_update_rows(ctx, R.live, (const int *)ctx->aux3.p, k, s));
        if (!ctx->ev_small) UNB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_small, cudaEventDisableTiming));
        UNB_CUDA(ctx, cudaEventRecord(ctx->ev_small, s));
        ctx->ev_small_pending = true;
    }
    if (rows_changed) *rows_changed = (int64_t)changed.size();
    return UNB_OK;
}

extern "C" int unb_region_set_layer(unb_ctx *ctx, int kind, const double *shift, const double *mat,
                                    size_t ndim)
{
    UNB_TRY(check_ctx(ctx));
    RegionState &R = ctx->region;
    cudaStream_t s = S0(ctx);
    if (kind != UNB_LAYER_IDENTITY && kind != UNB_LAYER_SCALING && kind != UNB_LAYER_AFFINE)
        return unb_fail(ctx, UNB_ERR_ARG, "unknown layer kind %d", kind);
    if (kind != UNB_LAYER_IDENTITY) {
        if (!shift || !mat || ndim == 0) return unb_fail(ctx, UNB_ERR_ARG, "null layer parameters");
        const size_t mat_n = kind == UNB_LAYER_AFFINE ? ndim * ndim : ndim;
        if (R.layer_kind == kind && R.layer_d == ndim && R.layer_shift_h.size() == ndim &&
            R.layer_mat_h.size() == mat_n &&
            memcmp(R.layer_shift_h.data(), shift, ndim * sizeof(double)) == 0 &&
            memcmp(R.layer_mat_h.data(), mat, mat_n * sizeof(double)) == 0)
            return UNB_OK;   // unchanged since the last call
        R.layer_shift_h.assign(shift, shift + ndim);
        R.layer_mat_h.assign(mat, mat + mat_n);
        R.param_version++;
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        UNB_TRY(unb_reserve(ctx, R.layer_shift, ndim * sizeof(double)));
        UNB_TRY(unb_reserve(ctx, R.layer_mat, mat_n * sizeof(double)));
        UNB_TRY(h2d(ctx, R.layer_shift.p, shift, ndim * sizeof(double), s));
        UNB_TRY(h2d(ctx, R.layer_mat.p, mat, mat_n * sizeof(double), s));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));
        if (kind == UNB_LAYER_AFFINE && unb_tile_prep_fits((int)ndim))
            UNB_TRY(upload_padded8(ctx, R.layer_mat_pad, mat, ndim, s));
    }
    if (R.layer_kind != kind) R.param_version++;
    R.layer_kind = kind;
    R.layer_d = ndim;
    return UNB_OK;
}

extern "C" int unb_region_set_ellipsoid(unb_ctx *ctx, const double *center, const double *invcov,
                                        double enlarge, size_t ndim)
{
    UNB_TRY(check_ctx(ctx));
    if (!center || !invcov || ndim == 0) return unb_fail(ctx, UNB_ERR_ARG, "null ellipsoid");
    RegionState &R = ctx->region;
    cudaStream_t s = S0(ctx);
    R.enlarge = enlarge;
    const bool same_shape = R.have_ellipsoid && R.ell_d == ndim && R.ell_center_h.size() == ndim &&
                            R.ell_invcov_h.size() == ndim * ndim;
    const bool same_invcov = same_shape && memcmp(R.ell_invcov_h.data(), invcov, ndim * ndim * sizeof(double)) == 0;
    const bool same_center = same_shape && memcmp(R.ell_center_h.data(), center, ndim * sizeof(double)) == 0;
    if (same_invcov && same_center) return UNB_OK;   // unchanged since the last call
    if (same_invcov) {
        // only the centre moved (every iteration of a run, integrator.py:2756): refresh the device
        // copy in stream order -- no matrix upload, no constant-bank update, no synchronisation of
        // this stream (the register prep kernel takes the centre as a kernel argument)
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        R.ell_center_h.assign(center, center + ndim);
        UNB_TRY(h2d(ctx, R.ell_center.p, R.ell_center_h.data(), ndim * sizeof(double), s));
        // kernels that read the centre from device memory (d > 32) may run on a caller's stream
        if (ndim > 32) UNB_CUDA(ctx, cudaStreamSynchronize(s));
        return UNB_OK;
    }
    R.ell_center_h.assign(center, center + ndim);
    R.ell_invcov_h.assign(invcov, invcov + ndim * ndim);
    {
        double fro = 0.0;
        for (size_t i = 0; i < ndim * ndim; i++) fro += invcov[i] * invcov[i];
        R.ell_fro = std::sqrt(fro) * (1.0 + 1e-9);
    }
    R.param_version++;
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
    UNB_TRY(unb_reserve(ctx, R.ell_center, ndim * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, R.ell_invcov, ndim * ndim * sizeof(double)));
    UNB_TRY(h2d(ctx, R.ell_center.p, center, ndim * sizeof(double), s));
    UNB_TRY(h2d(ctx, R.ell_invcov.p, invcov, ndim * ndim * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    if (unb_tile_prep_fits((int)ndim)) UNB_TRY(upload_padded8(ctx, R.ell_invcov_pad, invcov, ndim, s, true));
    R.ell_d = ndim;
    R.enlarge = enlarge;
    R.have_ellipsoid = true;
    return UNB_OK;
}

extern "C" int unb_region_set_transform_tolerance(unb_ctx *ctx, double tau)
{
    UNB_TRY(check_ctx(ctx));
    if (!(tau >= 0.0) || !std::isfinite(tau)) return unb_fail(ctx, UNB_ERR_ARG, "tolerance must be finite and >= 0");
    ctx->region.unc_tau = tau;
    return UNB_OK;
}

extern "C" int unb_region_set_radius(unb_ctx *ctx, double maxradiussq)
{
    if (!ctx) return UNB_ERR_ARG;
    if (!ctx->region.have_radius || ctx->region.r2 != maxradiussq)
        ctx->region.unc_tau = 0.0;   // a transform tolerance belongs to the radius it was derived for
    ctx->region.r2 = maxradiussq;
    ctx->region.have_radius = true;
    return UNB_OK;
}

namespace {

int region_ready(unb_ctx *ctx, bool need_ellipsoid)
{
    RegionState &R = ctx->region;
    if (!R.live.valid) return unb_fail(ctx, UNB_ERR_STATE, "region live block not set");
    if (!R.have_radius) return unb_fail(ctx, UNB_ERR_STATE, "region radius not set");
    if (need_ellipsoid) {
        if (!R.have_ellipsoid) return unb_fail(ctx, UNB_ERR_STATE, "region ellipsoid not set");
        if (R.ell_d != R.live.d) return unb_fail(ctx, UNB_ERR_STATE, "ellipsoid/live ndim mismatch");
        if (R.layer_kind != UNB_LAYER_IDENTITY && R.layer_d != R.live.d)
            return unb_fail(ctx, UNB_ERR_STATE, "layer/live ndim mismatch");
    }
    return UNB_OK;
}

// enqueue MLFriends.inside (+ optional likelihood) for device-resident rows on lane `ln`
// ellipsoid membership only (RobustEllipsoidRegion / SimpleRegion / WrappingEllipsoid .inside)
int enqueue_ellipsoid(unb_ctx *ctx, cudaStream_t s, const double *pts_dev, size_t m,
                      unsigned char *mask_dev)
{
    RegionState &R = ctx->region;
    const size_t d = R.ell_d;
    const bool use_const = d <= unb_const_maxd();
    if (use_const) UNB_TRY(unb_prep_sync_constants(ctx, s));
    PrepArgs p;
    memset(&p, 0, sizeof(p));
    p.pts = pts_dev;
    p.m = (long long)m;
    p.d = (int)d;
    p.center = (const double *)R.ell_center.p;
    if (d <= 32) memcpy(p.center_arg, R.ell_center_h.data(), d * sizeof(double));
    p.invcov = (const double *)R.ell_invcov.p;
    p.r2 = R.enlarge;
    p.mask = mask_dev;
    p.layer_kind = -1;
    p.use_constants = use_const ? 1 : 0;
    p.ell_tol_scale = 2.0 * ((double)(d * d + 2 * d + 8)) * 1.1102230246251565e-16 * R.ell_fro;
    if (!ctx->exact_only && unb_tile_prep_fits((int)d)) {
        p.pad_stride = (int)pad8(d);
        p.invcov_pad = (const double *)R.ell_invcov_pad.p;
    }
    return unb_launch_prep(ctx, p, s);
}

int enqueue_inside(unb_ctx *ctx, Lane &ln, cudaStream_t s, const double *pts_dev, size_t m,
                   unsigned char *mask_dev, long long *idx_dev, double *like_dev,
                   int loglike_kind, bool use_ellipsoid = true, bool ellipsoid_only = false)
{
    if (ellipsoid_only) return enqueue_ellipsoid(ctx, s, pts_dev, m, mask_dev);
    RegionState &R = ctx->region;
    const size_t d = R.live.d;
    UNB_TRY(unb_reserve(ctx, ln.tcand, m * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.items, m * sizeof(int)));
    UNB_TRY(unb_reserve(ctx, ln.counter, 2 * sizeof(int)));   // [0] survivors, [1] work-queue head
    UNB_CUDA(ctx, cudaMemsetAsync(ln.counter.p, 0, 2 * sizeof(int), s));
    if (idx_dev) UNB_CUDA(ctx, cudaMemsetAsync(idx_dev, 0xff, m * sizeof(long long), s));
    // mask-only requests on a tiled live block go through the register prep kernel (constant
    // memory parameters, fused likelihood) and the persistent any-neighbour kernel
    // (d <= 32: register prep kernel; 32 < d <= 128: generic prep kernel + fp32 membership kernel)
    const bool have32 = R.live.t32_valid && R.live.t32_r2 == filter_r2(ctx, R.live, R.r2) && ctx->filter_fp32;
    const bool reg_prep = !ctx->exact_only && !idx_dev && R.live.ntiles > 0 && R.live.dr <= 32;
    const bool use_any = reg_prep || (!ctx->exact_only && !idx_dev && have32);
    // 32 < d: tile prep kernel (register-blocked products out of shared memory)
    const bool tile_prep = !ctx->exact_only && unb_tile_prep_fits((int)d);
    const bool fuse_like = (reg_prep || (tile_prep && use_any)) && like_dev &&
                           loglike_kind != UNB_LOGLIKE_NONE;
    // ellipsoid parameters through the constant bank whenever they fit (d <= 56)
    const bool const_prep = reg_prep || (!ctx->exact_only && R.live.d <= unb_const_maxd());
    if (const_prep) UNB_TRY(unb_prep_sync_constants(ctx, s));
    PrepArgs p;
    memset(&p, 0, sizeof(p));
    p.pts = pts_dev;
    p.m = (long long)m;
    p.d = (int)d;
    p.center = use_ellipsoid ? (const double *)R.ell_center.p : nullptr;
    if (d <= 32 && R.ell_center_h.size() == d) memcpy(p.center_arg, R.ell_center_h.data(), d * sizeof(double));
    p.invcov = (const double *)R.ell_invcov.p;
    p.r2 = R.enlarge;
    p.mask = mask_dev;
    p.layer_kind = R.layer_kind;
    p.shift = (const double *)R.layer_shift.p;
    p.mat = (const double *)R.layer_mat.p;
    p.tcand = (double *)ln.tcand.p;
    p.items = (int *)ln.items.p;
    p.n_items = (int *)ln.counter.p;
    p.use_constants = const_prep ? 1 : 0;
    p.ell_tol_scale = 2.0 * ((double)(d * d + 2 * d + 8)) * 1.1102230246251565e-16 * R.ell_fro;
    if (tile_prep) {
        p.pad_stride = (int)pad8(d);
        p.invcov_pad = (const double *)R.ell_invcov_pad.p;
        p.mat_pad = (const double *)R.layer_mat_pad.p;
    }
    if (fuse_like) {
        p.like = like_dev;
        p.loglike_kind = loglike_kind;
        p.lparams = (const double *)ctx->lparams.p;
    }
    UNB_TRY(unb_launch_prep(ctx, p, s));
    ScanArgs a = scan_args_for(R.live);
    a.cand = (const double *)ln.tcand.p;
    a.out_row_idx = (const int *)ln.items.p;
    a.n_items_dev = (const int *)ln.counter.p;
    a.n_items = (long long)m;
    a.r2 = R.r2;
    a.out_mask = mask_dev;
    a.out_idx = idx_dev;
    if (use_any && R.layer_kind == UNB_LAYER_AFFINE && eff_tau(ctx) > 0.0 && ctx->unc.p) {
        // proposals were whitened on the device in the defined order: report exact decisions that
        // the reference's np.dot transform could turn around (the shim re-decides such calls)
        a.unc_tau = eff_tau(ctx);
        a.unc_count = (unsigned int *)ctx->unc.p;
    }
    if (use_any) {
        a.out_like = fuse_like ? like_dev : nullptr;
        if (have32) {
            a.tiles32 = (const float *)R.live.tiles32.p;   // prepared by the caller (set_h stage)
            a.perm32 = R.live.t32_clustered ? (const int *)R.live.perm32.p : nullptr;
            a.kappa32 = unb_kappa32(R.live.d);
            a.namax32 = sure_namax(ctx, R.live);
            UNB_TRY(attach_bins(ctx, ln, R.live, a, s));
        }
        UNB_TRY(unb_launch_inside_any(ctx, a, (int *)ln.counter.p + 1, s));
    } else {
        UNB_TRY(unb_launch_scan(ctx, SCAN_FIND, a, 1, s));
    }
    if (like_dev && loglike_kind != UNB_LOGLIKE_NONE && !fuse_like)
        UNB_TRY(unb_launch_loglike(ctx, loglike_kind, pts_dev, (int)d, (long long)m, like_dev,
                                   mask_dev, (const double *)ctx->lparams.p, s));
    return UNB_OK;
}

void lane_flush(Lane &ln)
{
    if (!ln.pend_rows) return;
    cudaEventSynchronize(ln.ev_done);
    if (ln.pend_mask) memcpy(ln.pend_mask, ln.pin_mask.p, ln.pend_rows);
    if (ln.pend_like) memcpy(ln.pend_like, ln.pin_like.p, ln.pend_rows * sizeof(double));
    if (ln.pend_idx) memcpy(ln.pend_idx, ln.pin_idx.p, ln.pend_rows * sizeof(long long));
    ln.pend_rows = 0;
    ln.pend_mask = nullptr;
    ln.pend_like = nullptr;
    ln.pend_idx = nullptr;
}

int upload_lparams(unb_ctx *ctx, int kind, const double *lparams, size_t d, cudaStream_t s)
{
    if (kind == UNB_LOGLIKE_GAUSS) {
        if (!lparams) return unb_fail(ctx, UNB_ERR_ARG, "gaussian likelihood needs parameters");
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
        UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
        UNB_TRY(unb_reserve(ctx, ctx->lparams, (d + 2) * sizeof(double)));
        UNB_TRY(h2d(ctx, ctx->lparams.p, lparams, (d + 2) * sizeof(double), s));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));
    } else if (kind != UNB_LOGLIKE_NONE && kind != UNB_LOGLIKE_EGGBOX && kind != UNB_LOGLIKE_ROSENBROCK) {
        return unb_fail(ctx, UNB_ERR_ARG, "unknown likelihood kind %d", kind);
    }
    return UNB_OK;
}

// Row counts of the pipeline chunks of a host-buffer call.  The call is bound by the H2D copies
// (PCIe); what it pays on top is the kernels + D2H of the LAST chunk, which nothing overlaps.  So
// the chunks are full-sized while there is plenty left and halve towards the end (down to 2^15
// rows): 3.40 -> 3.25 ms per 2^20 x 20 batch on B200 (tools/chunk_sweep.py).  An explicit
// UNB_OPT_CHUNK_ROWS gives uniform chunks.
std::vector<size_t> chunk_plan(const unb_ctx *ctx, size_t m)
{
    std::vector<size_t> plan;
    if (ctx->chunk_rows > 0) {
        const size_t c = (size_t)ctx->chunk_rows;
        for (size_t off = 0; off < m; off += c) plan.push_back(std::min(c, m - off));
        return plan;
    }
    const size_t base = (size_t)1 << 18, floor_rows = (size_t)1 << 15;
    size_t left = m;
    while (left > base) {
        plan.push_back(base);
        left -= base;
    }
    while (left > 0) {   // taper
        size_t c = left > 2 * floor_rows ? (left + 1) / 2 : left;
        plan.push_back(c);
        left -= c;
    }
    return plan;
}

// chunked, double-buffered host pipeline: H2D(c+1) overlaps kernels(c) and D2H(c)
int inside_host(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask, int64_t *idx_out,
                double *like, int loglike_kind, bool use_ellipsoid = true,
                bool ellipsoid_only = false)
{
    RegionState &R = ctx->region;
    const size_t d = ellipsoid_only ? R.ell_d : R.live.d;
    const size_t rowb = d * sizeof(double);
    const std::vector<size_t> plan = chunk_plan(ctx, m);
    const size_t chunk = *std::max_element(plan.begin(), plan.end());   // buffer size per lane
    if (!ellipsoid_only) {
        request_cluster(ctx, R.live, chunk);
        UNB_TRY(prepare_threshold(ctx, R.live, R.r2, S0(ctx), nullptr));
        UNB_TRY(unc_reset(ctx, S0(ctx)));
    }
    UNB_CUDA(ctx, cudaStreamSynchronize(S0(ctx)));
    const bool src_pinned = host_is_pinned(pts);
    const bool mask_pinned = host_is_pinned(mask);
    const bool like_pinned = like ? host_is_pinned(like) : true;
    const bool idx_pinned = idx_out ? host_is_pinned(idx_out) : true;
    int rc = UNB_OK;
    size_t c = 0;
    for (size_t off = 0; c < plan.size() && rc == UNB_OK; off += plan[c], c++) {
        const size_t rows = plan[c];
        Lane &ln = ctx->lane[c & 1];
        cudaStream_t s = ln.stream;
        lane_flush(ln);
        auto body = [&]() -> int {
            UNB_TRY(unb_reserve(ctx, ln.cand, chunk * rowb));
            UNB_TRY(unb_reserve(ctx, ln.mask, chunk));
            if (idx_out) UNB_TRY(unb_reserve(ctx, ln.idx, chunk * sizeof(long long)));
            if (like) UNB_TRY(unb_reserve(ctx, ln.like, chunk * sizeof(double)));
            if (src_pinned) {
                UNB_TRY(h2d(ctx, ln.cand.p, pts + off * d, rows * rowb, s));
            } else {
                UNB_TRY(unb_reserve_pinned(ctx, ln.pin_in, chunk * rowb));
                UNB_CUDA(ctx, cudaEventSynchronize(ln.ev_in));
                par_memcpy(ln.pin_in.p, pts + off * d, rows * rowb);
                UNB_TRY(h2d(ctx, ln.cand.p, ln.pin_in.p, rows * rowb, s));
                UNB_CUDA(ctx, cudaEventRecord(ln.ev_in, s));
            }
            UNB_TRY(enqueue_inside(ctx, ln, s, (const double *)ln.cand.p, rows,
                                   (unsigned char *)ln.mask.p,
                                   idx_out ? (long long *)ln.idx.p : nullptr,
                                   like ? (double *)ln.like.p : nullptr, loglike_kind,
                                   use_ellipsoid, ellipsoid_only));
            ln.pend_rows = 0;
            if (mask_pinned) {
                UNB_TRY(d2h(ctx, mask + off, ln.mask.p, rows, s));
            } else {
                UNB_TRY(unb_reserve_pinned(ctx, ln.pin_mask, chunk));
                UNB_TRY(d2h(ctx, ln.pin_mask.p, ln.mask.p, rows, s));
                ln.pend_mask = mask + off;
                ln.pend_rows = rows;
            }
            if (like) {
                if (like_pinned) {
                    UNB_TRY(d2h(ctx, like + off, ln.like.p, rows * sizeof(double), s));
                } else {
                    UNB_TRY(unb_reserve_pinned(ctx, ln.pin_like, chunk * sizeof(double)));
                    UNB_TRY(d2h(ctx, ln.pin_like.p, ln.like.p, rows * sizeof(double), s));
                    ln.pend_like = like + off;
                    ln.pend_rows = rows;
                }
            }
            if (idx_out) {
                if (idx_pinned) {
                    UNB_TRY(d2h(ctx, idx_out + off, ln.idx.p, rows * sizeof(long long), s));
                } else {
                    UNB_TRY(unb_reserve_pinned(ctx, ln.pin_idx, chunk * sizeof(long long)));
                    UNB_TRY(d2h(ctx, ln.pin_idx.p, ln.idx.p, rows * sizeof(long long), s));
                    ln.pend_idx = (long long *)idx_out + off;
                    ln.pend_rows = rows;
                }
            }
            UNB_CUDA(ctx, cudaEventRecord(ln.ev_done, s));
            return UNB_OK;
        };
        rc = body();
    }
    if (rc == UNB_OK && !ellipsoid_only) {
        rc = unc_copy(ctx, 0);
        if (rc == UNB_OK && c > 1) rc = unc_copy(ctx, 1);
    }
    for (int i = 0; i < 2; i++) {
        cudaStreamSynchronize(ctx->lane[i].stream);
        lane_flush(ctx->lane[i]);
    }
    if (rc == UNB_OK) {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return unb_fail(ctx, UNB_ERR_CUDA, "inside pipeline: %s", cudaGetErrorString(e));
        if (!ellipsoid_only) unc_collect(ctx, true, c > 1);
    }
    return rc;
}

}  // namespace

extern "C" int unb_region_inside(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                                 int64_t *idx_out)
{
    UNB_RANGE("unb_region_inside");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, true));
    if (m == 0) return UNB_OK;
    if (!pts || !mask) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return inside_host(ctx, pts, m, mask, idx_out, nullptr, UNB_LOGLIKE_NONE);
}

extern "C" int unb_region_inside_ellipsoid(unb_ctx *ctx, const double *pts, size_t m,
                                           uint8_t *mask)
{
    UNB_RANGE("unb_region_inside_ellipsoid");
    UNB_TRY(check_ctx(ctx));
    if (!ctx->region.have_ellipsoid) return unb_fail(ctx, UNB_ERR_STATE, "region ellipsoid not set");
    if (m == 0) return UNB_OK;
    if (!pts || !mask) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return inside_host(ctx, pts, m, mask, nullptr, nullptr, UNB_LOGLIKE_NONE, true, true);
}

extern "C" int unb_region_inside_ellipsoid_dev(unb_ctx *ctx, const double *pts_dev, size_t m,
                                               uint8_t *mask_dev, void *stream)
{
    UNB_TRY(check_ctx(ctx));
    if (!ctx->region.have_ellipsoid) return unb_fail(ctx, UNB_ERR_STATE, "region ellipsoid not set");
    if (m == 0) return UNB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    return enqueue_ellipsoid(ctx, s, pts_dev, m, mask_dev);
}

extern "C" int unb_region_friends(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                                  int64_t *idx_out)
{
    UNB_RANGE("unb_region_friends");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, true));
    if (m == 0) return UNB_OK;
    if (!pts || !mask) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return inside_host(ctx, pts, m, mask, idx_out, nullptr, UNB_LOGLIKE_NONE, false);
}

extern "C" int unb_region_inside_loglike(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                                         double *like, int loglike_kind, const double *lparams)
{
    UNB_RANGE("unb_region_inside_loglike");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, true));
    if (m == 0) return UNB_OK;
    if (!pts || !mask || !like) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    UNB_TRY(upload_lparams(ctx, loglike_kind, lparams, ctx->region.live.d, S0(ctx)));
    return inside_host(ctx, pts, m, mask, nullptr, like, loglike_kind);
}

// ---- fused refill (integrator.py:1773-1837) -------------------------------------------------
extern "C" int unb_region_refill(unb_ctx *ctx, const double *u, size_t m, size_t ndim,
                                 const unb_refill_desc *desc, uint8_t *flags, double *like,
                                 int64_t *counts)
{
    UNB_RANGE("unb_region_refill");
    UNB_TRY(check_ctx(ctx));
    if (!desc || !counts) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    counts[0] = counts[1] = counts[2] = 0;
    if (desc->region_mode < 0 || desc->region_mode > 2)
        return unb_fail(ctx, UNB_ERR_ARG, "region_mode must be 0, 1 or 2");
    if (desc->xform_kind != UNB_XFORM_IDENTITY && desc->xform_kind != UNB_XFORM_SCALE_SHIFT)
        return unb_fail(ctx, UNB_ERR_ARG, "unknown transform kind %d", desc->xform_kind);
    if (desc->xform_kind == UNB_XFORM_SCALE_SHIFT && (!desc->xform_scale || !desc->xform_lo))
        return unb_fail(ctx, UNB_ERR_ARG, "scale/shift transform needs parameters");
    if (desc->treg_center && !desc->treg_invcov)
        return unb_fail(ctx, UNB_ERR_ARG, "tregion needs its inverse covariance");
    if (desc->loglike_kind == UNB_LOGLIKE_NONE)
        return unb_fail(ctx, UNB_ERR_ARG, "refill needs a device likelihood");
    if (ndim == 0 || ndim > unb_max_rowwise_d() / 2)
        return unb_fail(ctx, UNB_ERR_ARG, "ndim=%zu outside the refill kernel's range", ndim);
    RegionState &R = ctx->region;
    if (desc->region_mode != 0) {
        UNB_TRY(region_ready(ctx, true));
        if (R.live.d != ndim) return unb_fail(ctx, UNB_ERR_ARG, "ndim does not match the region");
    }
    if (m == 0) return UNB_OK;
    if (!u || !flags || !like) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    const size_t d = ndim;
    cudaStream_t s0 = S0(ctx);
    UNB_TRY(upload_lparams(ctx, desc->loglike_kind, desc->lparams, d, s0));
    // parameter block: scale[d] lo[d] center[d] invcov[d*d] counters(3 ints, padded)
    const size_t nparam = 3 * d + d * d;
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
    UNB_TRY(unb_reserve(ctx, ctx->refill_params, (nparam + 2) * sizeof(double)));
    double *P = (double *)ctx->refill_params.p;
    if (desc->xform_kind == UNB_XFORM_SCALE_SHIFT) {
        UNB_TRY(h2d(ctx, P, desc->xform_scale, d * sizeof(double), s0));
        UNB_TRY(h2d(ctx, P + d, desc->xform_lo, d * sizeof(double), s0));
    }
    if (desc->treg_center) {
        UNB_TRY(h2d(ctx, P + 2 * d, desc->treg_center, d * sizeof(double), s0));
        UNB_TRY(h2d(ctx, P + 3 * d, desc->treg_invcov, d * d * sizeof(double), s0));
    }
    int *cnt_dev = (int *)(P + nparam);
    UNB_CUDA(ctx, cudaMemsetAsync(cnt_dev, 0, 4 * sizeof(int), s0));
    const size_t rowb = d * sizeof(double);
    const std::vector<size_t> plan = chunk_plan(ctx, m);
    const size_t chunk = *std::max_element(plan.begin(), plan.end());
    if (desc->region_mode != 0) {
        request_cluster(ctx, R.live, chunk);
        UNB_TRY(prepare_threshold(ctx, R.live, R.r2, s0, nullptr));
        UNB_TRY(unc_reset(ctx, s0));
    }
    UNB_CUDA(ctx, cudaStreamSynchronize(s0));

    const bool src_pinned = host_is_pinned(u);
    const bool flags_pinned = host_is_pinned(flags);
    const bool like_pinned = host_is_pinned(like);
    int rc = UNB_OK;
    size_t c = 0;
    for (size_t off = 0; c < plan.size() && rc == UNB_OK; off += plan[c], c++) {
        const size_t rows = plan[c];
        Lane &ln = ctx->lane[c & 1];
        cudaStream_t s = ln.stream;
        lane_flush(ln);
        auto body = [&]() -> int {
            UNB_TRY(unb_reserve(ctx, ln.cand, chunk * rowb));
            UNB_TRY(unb_reserve(ctx, ln.mask, chunk));
            UNB_TRY(unb_reserve(ctx, ln.like, chunk * sizeof(double)));
            if (src_pinned) {
                UNB_TRY(h2d(ctx, ln.cand.p, u + off * d, rows * rowb, s));
            } else {
                UNB_TRY(unb_reserve_pinned(ctx, ln.pin_in, chunk * rowb));
                UNB_CUDA(ctx, cudaEventSynchronize(ln.ev_in));
                par_memcpy(ln.pin_in.p, u + off * d, rows * rowb);
                UNB_TRY(h2d(ctx, ln.cand.p, ln.pin_in.p, rows * rowb, s));
                UNB_CUDA(ctx, cudaEventRecord(ln.ev_in, s));
            }
            if (desc->region_mode != 0)
                UNB_TRY(enqueue_inside(ctx, ln, s, (const double *)ln.cand.p, rows,
                                       (unsigned char *)ln.mask.p, nullptr, nullptr,
                                       UNB_LOGLIKE_NONE, desc->region_mode == 2, false));
            TailArgs t;
            memset(&t, 0, sizeof(t));
            t.pts = (const double *)ln.cand.p;
            t.m = (long long)rows;
            t.d = (int)d;
            t.flags = (unsigned char *)ln.mask.p;
            t.have_mask = desc->region_mode != 0;
            t.check_cube = desc->check_cube;
            t.xform_kind = desc->xform_kind;
            t.xform_scale = P;
            t.xform_lo = P + d;
            t.treg_center = desc->treg_center ? P + 2 * d : nullptr;
            t.treg_invcov = P + 3 * d;
            t.treg_r2 = desc->treg_enlarge;
            t.loglike_kind = desc->loglike_kind;
            t.lparams = (const double *)ctx->lparams.p;
            t.Lmin = desc->Lmin;
            t.like = (double *)ln.like.p;
            t.counts = cnt_dev;
            UNB_TRY(unb_launch_refill_tail(ctx, t, s));
            ln.pend_rows = 0;
            if (flags_pinned) {
                UNB_TRY(d2h(ctx, flags + off, ln.mask.p, rows, s));
            } else {
                UNB_TRY(unb_reserve_pinned(ctx, ln.pin_mask, chunk));
                UNB_TRY(d2h(ctx, ln.pin_mask.p, ln.mask.p, rows, s));
                ln.pend_mask = flags + off;
                ln.pend_rows = rows;
            }
            if (like_pinned) {
                UNB_TRY(d2h(ctx, like + off, ln.like.p, rows * sizeof(double), s));
            } else {
                UNB_TRY(unb_reserve_pinned(ctx, ln.pin_like, chunk * sizeof(double)));
                UNB_TRY(d2h(ctx, ln.pin_like.p, ln.like.p, rows * sizeof(double), s));
                ln.pend_like = like + off;
                ln.pend_rows = rows;
            }
            UNB_CUDA(ctx, cudaEventRecord(ln.ev_done, s));
            return UNB_OK;
        };
        rc = body();
    }
    if (rc == UNB_OK && desc->region_mode != 0) {
        rc = unc_copy(ctx, 0);
        if (rc == UNB_OK && c > 1) rc = unc_copy(ctx, 1);
    }
    for (int i = 0; i < 2; i++) {
        cudaStreamSynchronize(ctx->lane[i].stream);
        lane_flush(ctx->lane[i]);
    }
    if (rc != UNB_OK) return rc;
    if (desc->region_mode != 0) unc_collect(ctx, true, c > 1);
    int cnt_host[4] = {0, 0, 0, 0};
    UNB_TRY(d2h(ctx, cnt_host, cnt_dev, 3 * sizeof(int), s0));
    UNB_CUDA(ctx, cudaStreamSynchronize(s0));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return unb_fail(ctx, UNB_ERR_CUDA, "refill pipeline: %s", cudaGetErrorString(e));
    for (int i = 0; i < 3; i++) counts[i] = cnt_host[i];
    return UNB_OK;
}

// ---- device-side proposal generation (mlfriends.pyx:1096-1112, 1135-1160 without the host RNG) ----
namespace {

// draw `rows` proposals starting at global index `offset`, filter them and compact the accepted
// rows (draw order) into out_rows/out_like (device); *n_out_dev receives the count.  All on `s`,
// with the scratch buffers of lane `ln`.
int sample_enqueue(unb_ctx *ctx, Lane &ln, const unb_sample_desc *desc, unsigned long long offset,
                   size_t rows, double *out_rows_dev, double *out_like_dev, int *n_out_dev,
                   cudaStream_t s)
{
    RegionState &R = ctx->region;
    const size_t d = R.live.d;
    const bool want_like = desc->loglike_kind != UNB_LOGLIKE_NONE;
    UNB_TRY(unb_reserve(ctx, ln.cand, rows * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.mask, rows));
    UNB_TRY(unb_reserve(ctx, ln.smp_cube, rows));
    if (want_like) UNB_TRY(unb_reserve(ctx, ln.like, rows * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.smp_counts, unb_compact_scratch_ints((long long)rows) * sizeof(int)));
    UNB_TRY(unb_launch_draw(ctx, desc->method, (long long)rows, (int)d, desc->seed, offset,
                            (const double *)R.ell_center.p, (const double *)ctx->smp_axes.p,
                            sqrt(R.enlarge), (double *)ln.cand.p, (unsigned char *)ln.smp_cube.p, s));
    // wrapping-ellipsoid draws lie inside the ellipsoid by construction (the reference does not
    // test it either, mlfriends.pyx:1152-1158); unit-cube draws get the full inside()
    UNB_TRY(enqueue_inside(ctx, ln, s, (const double *)ln.cand.p, rows, (unsigned char *)ln.mask.p,
                           nullptr, want_like ? (double *)ln.like.p : nullptr, desc->loglike_kind,
                           desc->method == UNB_SAMPLE_UNIT_CUBE, false));
    UNB_TRY(unb_launch_finish_mask(ctx, (unsigned char *)ln.mask.p, (const unsigned char *)ln.smp_cube.p,
                                   want_like ? (const double *)ln.like.p : nullptr, desc->Lmin,
                                   want_like && desc->use_lmin, (long long)rows, s));
    return unb_launch_compact_rows(ctx, (const unsigned char *)ln.mask.p, (long long)rows, (int)d,
                                   (const double *)ln.cand.p,
                                   want_like ? (const double *)ln.like.p : nullptr,
                                   (int *)ln.smp_counts.p, n_out_dev, out_rows_dev,
                                   want_like ? out_like_dev : nullptr, nullptr, s);
}

int sample_prepare(unb_ctx *ctx, const unb_sample_desc *desc, size_t nsamples_hint, cudaStream_t s)
{
    if (!desc) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (desc->method != UNB_SAMPLE_WRAPPING_ELLIPSOID && desc->method != UNB_SAMPLE_UNIT_CUBE)
        return unb_fail(ctx, UNB_ERR_ARG, "unknown sampling method %d", desc->method);
    if (desc->use_lmin && desc->loglike_kind == UNB_LOGLIKE_NONE)
        return unb_fail(ctx, UNB_ERR_ARG, "the Lmin cut needs a device likelihood");
    UNB_TRY(region_ready(ctx, true));
    RegionState &R = ctx->region;
    const size_t d = R.live.d;
    if (desc->method == UNB_SAMPLE_WRAPPING_ELLIPSOID) {
        if (!desc->axes_T) return unb_fail(ctx, UNB_ERR_ARG, "wrapping-ellipsoid draws need axes_T");
        if (!(R.enlarge > 0.0)) return unb_fail(ctx, UNB_ERR_NUMERIC, "enlarge must be positive");
        if (ctx->smp_axes_h.size() != d * d ||
            memcmp(ctx->smp_axes_h.data(), desc->axes_T, d * d * sizeof(double)) != 0) {
            UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
            UNB_CUDA(ctx, cudaStreamSynchronize(s));
            ctx->smp_axes_h.assign(desc->axes_T, desc->axes_T + d * d);
            UNB_TRY(unb_reserve(ctx, ctx->smp_axes, d * d * sizeof(double)));
            UNB_TRY(h2d(ctx, ctx->smp_axes.p, ctx->smp_axes_h.data(), d * d * sizeof(double), s));
            UNB_CUDA(ctx, cudaStreamSynchronize(s));
        }
    }
    if (desc->loglike_kind != UNB_LOGLIKE_NONE && desc->lparams)
        UNB_TRY(upload_lparams(ctx, desc->loglike_kind, desc->lparams, d, s));
    request_cluster(ctx, R.live, nsamples_hint);
    return prepare_threshold(ctx, R.live, R.r2, s, nullptr);
}

}  // namespace

extern "C" int unb_region_sample_dev(unb_ctx *ctx, const unb_sample_desc *desc, size_t nsamples,
                                     double *rows_out_dev, double *like_out_dev, int32_t *n_out_dev,
                                     void *stream)
{
    UNB_RANGE("unb_region_sample_dev");
    UNB_TRY(check_ctx(ctx));
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    UNB_TRY(sample_prepare(ctx, desc, nsamples, s));
    if (!rows_out_dev || !n_out_dev) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (nsamples > 0x7fffffffULL) return unb_fail(ctx, UNB_ERR_ARG, "too many rows (max 2^31-1 per call)");
    if (nsamples == 0) {
        UNB_CUDA(ctx, cudaMemsetAsync(n_out_dev, 0, sizeof(int), s));
        return UNB_OK;
    }
    return sample_enqueue(ctx, ctx->lane[0], desc, desc->offset, nsamples, rows_out_dev, like_out_dev,
                          (int *)n_out_dev, s);
}

extern "C" int unb_region_sample(unb_ctx *ctx, const unb_sample_desc *desc, size_t nsamples,
                                 double *rows_out, double *like_out, int64_t *n_out, int64_t *counts)
{
    UNB_RANGE("unb_region_sample");
    UNB_TRY(check_ctx(ctx));
    if (!n_out) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    *n_out = 0;
    if (counts) counts[0] = counts[1] = counts[2] = 0;
    cudaStream_t s = S0(ctx);
    UNB_TRY(sample_prepare(ctx, desc,
                           std::min<size_t>(nsamples, ctx->chunk_rows > 0 ? (size_t)ctx->chunk_rows : (size_t)(1 << 18)), s));
    if (nsamples == 0) return UNB_OK;
    if (!rows_out) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    const size_t d = ctx->region.live.d;
    const bool want_like = desc->loglike_kind != UNB_LOGLIKE_NONE && like_out;
    // chunks alternate between the two lanes: the kernels of chunk c+1 run while the accepted rows
    // of chunk c travel back.  The row copy of a chunk needs its count, hence the per-chunk sync of
    // that lane only.
    const size_t chunk = std::min<size_t>(nsamples, ctx->chunk_rows > 0 ? (size_t)ctx->chunk_rows : (size_t)(1 << 18));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));   // sample_prepare ran on lane 0's stream
    for (int i = 0; i < 2; i++) {
        Lane &ln = ctx->lane[i];
        UNB_TRY(unb_reserve(ctx, ln.smp_rows, chunk * d * sizeof(double)));
        UNB_TRY(unb_reserve(ctx, ln.smp_like, chunk * sizeof(double)));
        UNB_TRY(unb_reserve(ctx, ln.smp_n, 4 * sizeof(int)));
        UNB_TRY(unb_reserve_pinned(ctx, ln.pin_n, 4 * sizeof(int)));
    }
    size_t filled = 0;
    const size_t nchunks = (nsamples + chunk - 1) / chunk;
    auto launch = [&](size_t c) -> int {
        Lane &ln = ctx->lane[c & 1];
        const size_t off = c * chunk, rows = std::min(chunk, nsamples - off);
        UNB_TRY(sample_enqueue(ctx, ln, desc, desc->offset + off, rows, (double *)ln.smp_rows.p,
                               (double *)ln.smp_like.p, (int *)ln.smp_n.p, ln.stream));
        return d2h(ctx, ln.pin_n.p, ln.smp_n.p, sizeof(int), ln.stream);
    };
    auto collect = [&](size_t c) -> int {
        Lane &ln = ctx->lane[c & 1];
        UNB_CUDA(ctx, cudaStreamSynchronize(ln.stream));
        const int n_acc = *(const int *)ln.pin_n.p;
        if (n_acc > 0) {
            UNB_TRY(d2h(ctx, rows_out + filled * d, ln.smp_rows.p, (size_t)n_acc * d * sizeof(double), ln.stream));
            if (want_like)
                UNB_TRY(d2h(ctx, like_out + filled, ln.smp_like.p, (size_t)n_acc * sizeof(double), ln.stream));
            filled += (size_t)n_acc;
        }
        return UNB_OK;
    };
    UNB_TRY(launch(0));
    for (size_t c = 1; c < nchunks; c++) {
        UNB_TRY(launch(c));
        UNB_TRY(collect(c - 1));
    }
    UNB_TRY(collect(nchunks - 1));
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[0].stream));
    UNB_CUDA(ctx, cudaStreamSynchronize(ctx->lane[1].stream));
    *n_out = (int64_t)filled;
    if (counts) counts[2] = (int64_t)filled;
    return UNB_OK;
}

extern "C" int unb_sample_draw(unb_ctx *ctx, int method, size_t nsamples, size_t ndim, uint64_t seed,
                               uint64_t offset, const double *center, const double *axes_T,
                               double enlarge, double *rows_out, uint8_t *cube_out)
{
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, nsamples, ndim));
    if (nsamples == 0) return UNB_OK;
    if (!rows_out || !cube_out) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (method != UNB_SAMPLE_WRAPPING_ELLIPSOID && method != UNB_SAMPLE_UNIT_CUBE)
        return unb_fail(ctx, UNB_ERR_ARG, "unknown sampling method %d", method);
    const size_t d = ndim;
    cudaStream_t s = S0(ctx);
    Lane &ln = ctx->lane[0];
    UNB_TRY(unb_reserve(ctx, ln.cand, nsamples * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.smp_cube, nsamples));
    const double *c_dev = nullptr, *a_dev = nullptr;
    if (method == UNB_SAMPLE_WRAPPING_ELLIPSOID) {
        if (!center || !axes_T || !(enlarge > 0.0))
            return unb_fail(ctx, UNB_ERR_ARG, "wrapping-ellipsoid draws need center, axes_T, enlarge > 0");
        UNB_TRY(unb_reserve(ctx, ctx->smp_center, d * sizeof(double)));
        UNB_TRY(unb_reserve(ctx, ctx->aux2, d * d * sizeof(double)));
        UNB_TRY(h2d(ctx, ctx->smp_center.p, center, d * sizeof(double), s));
        UNB_TRY(h2d(ctx, ctx->aux2.p, axes_T, d * d * sizeof(double), s));
        c_dev = (const double *)ctx->smp_center.p;
        a_dev = (const double *)ctx->aux2.p;
    }
    UNB_TRY(unb_launch_draw(ctx, method, (long long)nsamples, (int)d, seed, offset, c_dev, a_dev,
                            method == UNB_SAMPLE_WRAPPING_ELLIPSOID ? sqrt(enlarge) : 1.0,
                            (double *)ln.cand.p, (unsigned char *)ln.smp_cube.p, s));
    UNB_TRY(d2h(ctx, rows_out, ln.cand.p, nsamples * d * sizeof(double), s));
    UNB_TRY(d2h(ctx, cube_out, ln.smp_cube.p, nsamples, s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

extern "C" int unb_region_inside_dev(unb_ctx *ctx, const double *pts_dev, size_t m,
                                     uint8_t *mask_dev, void *stream)
{
    UNB_RANGE("unb_region_inside_dev");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, true));
    if (m == 0) return UNB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    request_cluster(ctx, ctx->region.live, m);
    UNB_TRY(prepare_threshold(ctx, ctx->region.live, ctx->region.r2, s, nullptr));
    return enqueue_inside(ctx, ctx->lane[0], s, pts_dev, m, mask_dev, nullptr, nullptr,
                          UNB_LOGLIKE_NONE);
}

extern "C" int unb_region_inside_loglike_dev(unb_ctx *ctx, const double *pts_dev, size_t m,
                                             uint8_t *mask_dev, double *like_dev, int loglike_kind,
                                             const double *lparams, void *stream)
{
    UNB_RANGE("unb_region_inside_loglike_dev");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, true));
    if (m == 0) return UNB_OK;
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    if (lparams) UNB_TRY(upload_lparams(ctx, loglike_kind, lparams, ctx->region.live.d, s));
    request_cluster(ctx, ctx->region.live, m);
    UNB_TRY(prepare_threshold(ctx, ctx->region.live, ctx->region.r2, s, nullptr));
    return enqueue_inside(ctx, ctx->lane[0], s, pts_dev, m, mask_dev, nullptr, like_dev,
                          loglike_kind);
}

extern "C" int unb_region_find_nearby(unb_ctx *ctx, const double *tpts, size_t m, int64_t *nnearby)
{
    UNB_RANGE("unb_region_find_nearby");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, false));
    if (m == 0) return UNB_OK;
    if (!tpts || !nnearby) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return scan_host(ctx, ctx->region.live, SCAN_FIND, tpts, m, ctx->region.r2,
                     (long long *)nnearby, nullptr, nullptr);
}

namespace {
// mask-only scan of host candidates against a built live block (any-neighbour kernel)
int has_neighbour_host(unb_ctx *ctx, LiveTiles &L, const double *tpts, size_t m, double r2,
                       uint8_t *mask)
{
    Lane &ln = ctx->lane[0];
    cudaStream_t s = ln.stream;
    const size_t d = L.d;
    ScanArgs pre;
    memset(&pre, 0, sizeof(pre));
    UNB_TRY(prepare_threshold(ctx, L, r2, s, &pre));
    UNB_TRY(unb_reserve(ctx, ln.cand, m * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.mask, m));
    UNB_TRY(unb_reserve(ctx, ln.counter, 2 * sizeof(int)));
    UNB_TRY(h2d(ctx, ln.cand.p, tpts, m * d * sizeof(double), s));
    UNB_CUDA(ctx, cudaMemsetAsync(ln.counter.p, 0, 2 * sizeof(int), s));
    UNB_TRY(stat_reset(ctx, s));
    ScanArgs a = scan_args_for(L);
    a.tiles32 = pre.tiles32;
    a.perm32 = pre.perm32;     // clustered fp32 tiles: slot -> live row (must travel with tiles32)
    a.kappa32 = pre.kappa32;
    a.namax32 = pre.namax32;
    a.cand = (const double *)ln.cand.p;
    a.n_items = (long long)m;
    a.r2 = r2;
    a.out_mask = (unsigned char *)ln.mask.p;
    a.stat_rechecks = (unsigned long long *)ctx->stat.p;
    UNB_TRY(unb_launch_inside_any(ctx, a, (int *)ln.counter.p + 1, s));
    UNB_TRY(d2h(ctx, mask, ln.mask.p, m, s));
    return stat_fetch(ctx, s);
}
}  // namespace

// `find_nearby(apts, bpts, r2, out); out >= 0` without the index (the growth test of
// _update_clusters, mlfriends.pyx:304-307)
extern "C" int unb_has_neighbour(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                                 size_t nb, size_t ndim, double radiussq, uint8_t *mask)
{
    UNB_RANGE("unb_has_neighbour");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, na > nb ? na : nb, ndim));
    if (nb == 0) return UNB_OK;
    if (!bpts || !mask || (na && !apts)) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (na == 0) {
        memset(mask, 0, nb);
        return UNB_OK;
    }
    UNB_TRY(live_from_host(ctx, ctx->scratch_live, apts, na, ndim, S0(ctx)));
    return has_neighbour_host(ctx, ctx->scratch_live, bpts, nb, radiussq, mask);
}

// mask-only variant of unb_region_find_nearby (what `find_nearby(...) >= 0` callers need): runs
// the any-neighbour kernel instead of the ordered first-index scan
extern "C" int unb_region_has_neighbour(unb_ctx *ctx, const double *tpts, size_t m, uint8_t *mask)
{
    UNB_RANGE("unb_region_has_neighbour");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, false));
    if (m == 0) return UNB_OK;
    if (!tpts || !mask) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return has_neighbour_host(ctx, ctx->region.live, tpts, m, ctx->region.r2, mask);
}

extern "C" int unb_region_find_nearby_dev(unb_ctx *ctx, const double *tpts_dev, size_t m,
                                          int64_t *nnearby_dev, uint8_t *mask_dev, void *stream)
{
    UNB_RANGE("unb_region_find_nearby_dev");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, false));
    if (m == 0) return UNB_OK;
    if (!tpts_dev || (!nnearby_dev && !mask_dev)) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    ScanArgs pre;
    memset(&pre, 0, sizeof(pre));
    if (!nnearby_dev) request_cluster(ctx, ctx->region.live, m);
    UNB_TRY(prepare_threshold(ctx, ctx->region.live, ctx->region.r2, s, &pre));
    ScanArgs a = scan_args_for(ctx->region.live);
    if (!nnearby_dev) {
        a.tiles32 = pre.tiles32;
        a.perm32 = pre.perm32;
        a.kappa32 = pre.kappa32;
        a.namax32 = pre.namax32;
    }
    a.cand = tpts_dev;
    a.n_items = (long long)m;
    a.r2 = ctx->region.r2;
    a.out_idx = (long long *)nnearby_dev;
    a.out_mask = mask_dev;
    Lane &ln = ctx->lane[0];
    if (nnearby_dev) {
        ScanArgs a2 = a;
        a2.tiles32 = pre.tiles32;
        a2.perm32 = pre.perm32;
        a2.kappa32 = pre.kappa32;
        a2.namax32 = pre.namax32;
        bool two_phase = false;
        UNB_TRY(find_two_phase(ctx, ln, a2, s, &two_phase));
        if (two_phase) return UNB_OK;
        return unb_launch_scan(ctx, SCAN_FIND, a, 1, s);
    }
    // mask only: the persistent any-neighbour kernel MLFriends.inside uses
    UNB_TRY(unb_reserve(ctx, ln.counter, 2 * sizeof(int)));
    UNB_CUDA(ctx, cudaMemsetAsync(ln.counter.p, 0, 2 * sizeof(int), s));
    UNB_TRY(stat_reset(ctx, s));
    a.stat_rechecks = (unsigned long long *)ctx->stat.p;
    a.stat_tiles = (unsigned long long *)ctx->stat.p + 1;
    UNB_TRY(attach_bins(ctx, ln, ctx->region.live, a, s));
    return unb_launch_inside_any(ctx, a, (int *)ln.counter.p + 1, s);
}

extern "C" int unb_region_count_nearby(unb_ctx *ctx, const double *tpts, size_t m, int64_t *nnearby)
{
    UNB_RANGE("unb_region_count_nearby");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(region_ready(ctx, false));
    if (m == 0) return UNB_OK;
    if (!tpts || !nnearby) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    return scan_host(ctx, ctx->region.live, SCAN_COUNT, tpts, m, ctx->region.r2,
                     (long long *)nnearby, nullptr, nullptr);
}

// ---------------------------------------------------------------------------------------
// bootstrap
// ---------------------------------------------------------------------------------------
namespace {

// what bootstrap_enqueue leaves behind: per-active-round results on the device
struct BootRounds {
    std::vector<int> round_of;         // active slot -> round
    unsigned long long *dMax = nullptr;   // [R] bits of the round's max-min squared distance
    unsigned long long *dF = nullptr;     // [R] total-order keys of the round's einsum maximum
    bool want_d = false, want_f = false;
};

// Enqueues the device half of rounds [round_lo, round_hi) on stream s: index lists of the active
// rounds (host), gather of the per-round tiles, the max-min scan of all rounds in one launch and
// the per-round einsum maxima.  Nothing is read back here.
int bootstrap_enqueue(unb_ctx *ctx, const double *unormed, const double *u, size_t n, size_t ndim,
                      const uint8_t *selected, size_t round_lo, size_t round_hi, const double *ctrs,
                      const double *invcovs, bool want_d, bool want_f, cudaStream_t s, BootRounds *B)
{
    const size_t d = ndim;
    B->want_d = want_d;
    B->want_f = want_f;
    // host: index lists of the active rounds (selection masks come from the host RNG in the
    // reference's order; rounds with all/none selected are skipped like mlfriends.pyx:1048-1049)
    std::vector<int> &round_of = B->round_of;
    std::vector<int> idxA, idxB, offA, offB, nA, nB;
    for (size_t r = round_lo; r < round_hi; r++) {
        const uint8_t *sel = selected + r * n;
        size_t ca = 0;
        for (size_t i = 0; i < n; i++) ca += sel[i] ? 1 : 0;
        if (ca == 0 || ca == n) continue;
        round_of.push_back((int)r);
        offA.push_back((int)idxA.size());
        offB.push_back((int)idxB.size());
        nA.push_back((int)ca);
        nB.push_back((int)(n - ca));
        for (size_t i = 0; i < n; i++) {
            if (sel[i]) idxA.push_back((int)i);
            else idxB.push_back((int)i);
        }
    }
    const int R = (int)round_of.size();
    if (R == 0) return UNB_OK;

    // device: rows, index lists, meta
    if (want_d) {
        UNB_TRY(unb_reserve(ctx, ctx->boot_rows, n * d * sizeof(double)));
        UNB_TRY(h2d(ctx, ctx->boot_rows.p, unormed, n * d * sizeof(double), s));
    }
    const size_t nidx = idxA.size() + idxB.size();
    UNB_TRY(unb_reserve(ctx, ctx->boot_idx, nidx * sizeof(int)));
    int *dA = (int *)ctx->boot_idx.p;
    int *dB = dA + idxA.size();
    UNB_TRY(h2d(ctx, dA, idxA.data(), idxA.size() * sizeof(int), s));
    UNB_TRY(h2d(ctx, dB, idxB.data(), idxB.size() * sizeof(int), s));
    UNB_TRY(unb_reserve(ctx, ctx->boot_meta, 4 * (size_t)R * sizeof(int)));
    int *dOffA = (int *)ctx->boot_meta.p, *dOffB = dOffA + R, *dNA = dOffB + R, *dNB = dNA + R;
    UNB_TRY(h2d(ctx, dOffA, offA.data(), R * sizeof(int), s));
    UNB_TRY(h2d(ctx, dOffB, offB.data(), R * sizeof(int), s));
    UNB_TRY(h2d(ctx, dNA, nA.data(), R * sizeof(int), s));
    UNB_TRY(h2d(ctx, dNB, nB.data(), R * sizeof(int), s));
    UNB_TRY(unb_reserve(ctx, ctx->boot_out, 2 * (size_t)R * sizeof(unsigned long long)));
    UNB_CUDA(ctx, cudaMemsetAsync(ctx->boot_out.p, 0, 2 * (size_t)R * sizeof(unsigned long long), s));
    unsigned long long *dMax = (unsigned long long *)ctx->boot_out.p;
    unsigned long long *dF = dMax + R;
    B->dMax = dMax;
    B->dF = dF;

    int maxB = 0;
    for (int r = 0; r < R; r++) maxB = std::max(maxB, nB[r]);
    UNB_TRY(stat_reset(ctx, s));
    if (want_d) {
        // per-round compacted tiles of the selected rows
        const size_t dr = (d + 3) / 4 * 4;
        const size_t tile_n = unb_pick_tile_n(d);
        ScanArgs a;
        memset(&a, 0, sizeof(a));
        a.live_rows = (const double *)ctx->boot_rows.p;
        a.live_idx = dA;
        a.round_live_off = dOffA;
        a.d = (int)d;
        a.dr = (int)dr;
        a.tile_n = (int)tile_n;
        a.kappa = unb_kappa(d);
        a.round_nlive = dNA;
        a.round_nitems = dNB;
        a.round_item_off = dOffB;
        a.cand = (const double *)ctx->boot_rows.p;
        a.item_idx = dB;
        a.out_round_max = dMax;
        a.n_items = maxB;
        a.stat_rechecks = (unsigned long long *)ctx->stat.p;
        if (tile_n > 0 && !ctx->exact_only) {
            const size_t max_tiles = (n + tile_n - 1) / tile_n;
            const long long stride = (long long)(max_tiles * (dr + 1) * tile_n);
            UNB_TRY(unb_reserve(ctx, ctx->boot_tiles, (size_t)R * (size_t)stride * sizeof(double)));
            UNB_TRY(unb_launch_gather_round_tiles(ctx, (const double *)ctx->boot_rows.p, (int)n,
                                                  (int)d, (int)dr, (int)tile_n, dA, dOffA, dNA, R,
                                                  stride, (double *)ctx->boot_tiles.p, s));
            a.tiles = (const double *)ctx->boot_tiles.p;
            a.round_tile_stride = stride;
            // max squared norm over ALL rows bounds every round's subset
            UNB_TRY(unb_live_build(ctx, ctx->scratch_live, (const double *)ctx->boot_rows.p, n, d, s));
            a.namax_bits = (const unsigned long long *)ctx->scratch_live.namax.p;
        }
        UNB_TRY(unb_launch_scan(ctx, SCAN_MIN, a, R, s));
    }

    if (want_f) {
        UNB_TRY(unb_reserve(ctx, ctx->boot_u, n * d * sizeof(double)));
        UNB_TRY(h2d(ctx, ctx->boot_u.p, u, n * d * sizeof(double), s));
        // the pinned block may still be read by an earlier call's copy on another stream
        UNB_TRY(unb_wait_small(ctx));
        UNB_CUDA(ctx, cudaStreamSynchronize(s));
        UNB_TRY(unb_reserve_pinned(ctx, ctx->pin_small, (size_t)R * (d + d * d) * sizeof(double)));
        double *hc = (double *)ctx->pin_small.p;
        double *ha = hc + (size_t)R * d;
        for (int r = 0; r < R; r++) {
            memcpy(hc + (size_t)r * d, ctrs + (size_t)round_of[r] * d, d * sizeof(double));
            memcpy(ha + (size_t)r * d * d, invcovs + (size_t)round_of[r] * d * d, d * d * sizeof(double));
        }
        UNB_TRY(unb_reserve(ctx, ctx->boot_ell, (size_t)R * (d + d * d) * sizeof(double)));
        UNB_TRY(h2d(ctx, ctx->boot_ell.p, hc, (size_t)R * (d + d * d) * sizeof(double), s));
        const double *dC = (const double *)ctx->boot_ell.p;
        const double *dAinv = dC + (size_t)R * d;
        UNB_TRY(unb_launch_enlargement_f(ctx, (const double *)ctx->boot_u.p, (int)d, dB, dOffB, dNB,
                                         maxB, dC, dAinv, R, dF, s));
    }
    return UNB_OK;
}

// Folds the per-round device results into the 5 doubles one allreduce(MAX) ships
// (integrator.py:395-404): [max radius^2 (each round rounded to float32 like the reference's C
// `float` return), max enlargement, failure flag (a round with f <= 0 or non-finite,
// mlfriends.pyx:1063-1065), tag, -tag].  The tag pair lets the ranks verify after the MAX that
// they all worked on the same selection masks (max(tag) == -max(-tag) iff all tags agree).
__global__ void k_boot_fold(const unsigned long long *__restrict__ dMax,
                            const unsigned long long *__restrict__ dF, int R, int want_d, int want_f,
                            double failed_in, double tag, double *__restrict__ out)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double maxd = 0.0, maxf = 0.0, failed = failed_in;
    for (int r = 0; r < R; r++) {
        if (want_d) {
            const double v = (double)(float)__longlong_as_double((long long)dMax[r]);
            maxd = v > maxd ? v : maxd;
        }
        if (want_f) {
            const unsigned long long key = dF[r];
            const unsigned long long bits = (key >> 63) ? (key & 0x7fffffffffffffffULL) : ~key;
            const double f = __longlong_as_double((long long)bits);
            if (!(f > 0.0) || isinf(f) || isnan(f)) failed = 1.0;
            else maxf = f > maxf ? f : maxf;
        }
    }
    out[0] = maxd;
    out[1] = maxf;
    out[2] = failed;
    out[3] = tag;
    out[4] = -tag;
}

}  // namespace

extern "C" int unb_region_bootstrap(unb_ctx *ctx, const double *unormed, const double *u, size_t n,
                                    size_t ndim, const uint8_t *selected, size_t nrounds,
                                    size_t round_lo, size_t round_hi, const double *ctrs,
                                    const double *invcovs, double *maxd_out, double *f_out)
{
    UNB_RANGE("unb_region_bootstrap");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    const bool want_d = unormed && maxd_out;
    const bool want_f = u && ctrs && invcovs && f_out;
    if (!selected || n == 0 || (!want_d && !want_f))
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer / empty live block");
    if (round_hi > nrounds) round_hi = nrounds;
    if (round_lo >= round_hi) return UNB_OK;
    cudaStream_t s = S0(ctx);
    for (size_t r = round_lo; r < round_hi; r++) {   // inactive rounds report 0
        if (want_d) maxd_out[r] = 0.0;
        if (want_f) f_out[r] = 0.0;
    }
    BootRounds B;
    UNB_TRY(bootstrap_enqueue(ctx, unormed, u, n, ndim, selected, round_lo, round_hi, ctrs, invcovs,
                              want_d, want_f, s, &B));
    const int R = (int)B.round_of.size();
    if (R == 0) return UNB_OK;
    std::vector<unsigned long long> out(2 * (size_t)R);
    UNB_TRY(d2h(ctx, out.data(), ctx->boot_out.p, 2 * (size_t)R * sizeof(unsigned long long), s));
    UNB_TRY(stat_fetch(ctx, s));
    for (int r = 0; r < R; r++) {
        if (want_d) {
            double maxd;
            memcpy(&maxd, &out[r], sizeof(double));
            maxd_out[B.round_of[r]] = (double)(float)maxd;   // C `float` return (mlfriends.pyx:188)
        }
        if (want_f) {
            unsigned long long key = out[R + r];
            unsigned long long bits = (key >> 63) ? (key & 0x7fffffffffffffffULL) : ~key;
            double f;
            memcpy(&f, &bits, sizeof(double));
            f_out[B.round_of[r]] = f;
        }
    }
    return UNB_OK;
}

extern "C" int unb_region_bootstrap_moments(unb_ctx *ctx, const double *u, size_t n, size_t ndim,
                                            const uint8_t *selected, size_t nrounds, size_t round_lo,
                                            size_t round_hi, const double *c0, double *sums, double *sxx,
                                            int64_t *counts)
{
    UNB_RANGE("unb_region_bootstrap_moments");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    if (!u || !selected || !c0 || !sums || !sxx || !counts || n == 0)
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer / empty live block");
    if (round_hi > nrounds) round_hi = nrounds;
    if (round_lo >= round_hi) return UNB_OK;
    cudaStream_t s = S0(ctx);
    const size_t d = ndim;
    const int R = (int)(round_hi - round_lo);
    std::vector<int> idxA, offA, nA;
    int maxA = 0;
    for (size_t r = round_lo; r < round_hi; r++) {
        const uint8_t *sel = selected + r * n;
        offA.push_back((int)idxA.size());
        for (size_t i = 0; i < n; i++)
            if (sel[i]) idxA.push_back((int)i);
        nA.push_back((int)idxA.size() - offA.back());
        counts[r] = nA.back();
        maxA = std::max(maxA, nA.back());
    }
    UNB_TRY(unb_reserve(ctx, ctx->boot_u, n * d * sizeof(double)));
    UNB_TRY(h2d(ctx, ctx->boot_u.p, u, n * d * sizeof(double), s));
    UNB_TRY(unb_reserve(ctx, ctx->boot_idx, (idxA.size() + 1) * sizeof(int)));
    UNB_TRY(h2d(ctx, ctx->boot_idx.p, idxA.data(), idxA.size() * sizeof(int), s));
    UNB_TRY(unb_reserve(ctx, ctx->boot_meta, 4 * (size_t)R * sizeof(int)));
    int *dOff = (int *)ctx->boot_meta.p, *dN = dOff + R;
    UNB_TRY(h2d(ctx, dOff, offA.data(), R * sizeof(int), s));
    UNB_TRY(h2d(ctx, dN, nA.data(), R * sizeof(int), s));
    UNB_TRY(unb_reserve(ctx, ctx->boot_ell, ((size_t)R * (d + d * d) + d) * sizeof(double)));
    double *dSums = (double *)ctx->boot_ell.p, *dSxx = dSums + (size_t)R * d, *dC0 = dSxx + (size_t)R * d * d;
    UNB_TRY(h2d(ctx, dC0, c0, d * sizeof(double), s));
    UNB_TRY(unb_launch_round_moments(ctx, (const double *)ctx->boot_u.p, (int)d, (const int *)ctx->boot_idx.p,
                                     dOff, dN, maxA, R, dC0, dSums, dSxx, s));
    UNB_TRY(d2h(ctx, sums + round_lo * d, dSums, (size_t)R * d * sizeof(double), s));
    UNB_TRY(d2h(ctx, sxx + round_lo * d * d, dSxx, (size_t)R * d * d * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}

extern "C" int unb_region_bootstrap_fold_dev(unb_ctx *ctx, const double *unormed, const double *u,
                                             size_t n, size_t ndim, const uint8_t *selected,
                                             size_t nrounds, size_t round_lo, size_t round_hi,
                                             const double *ctrs, const double *invcovs,
                                             int host_failed, double tag, double *out5_dev,
                                             void *stream)
{
    UNB_RANGE("unb_region_bootstrap_fold_dev");
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, ndim));
    const bool want_d = unormed != nullptr;
    const bool want_f = u && ctrs && invcovs;
    if (!selected || !out5_dev || n == 0 || (!want_d && !want_f))
        return unb_fail(ctx, UNB_ERR_ARG, "null pointer / empty live block");
    if (round_hi > nrounds) round_hi = nrounds;
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    BootRounds B;
    if (round_lo < round_hi)
        UNB_TRY(bootstrap_enqueue(ctx, unormed, u, n, ndim, selected, round_lo, round_hi, ctrs,
                                  invcovs, want_d, want_f, s, &B));
    k_boot_fold<<<1, 32, 0, s>>>(B.dMax, B.dF, (int)B.round_of.size(), want_d ? 1 : 0,
                                 want_f ? 1 : 0, host_failed ? 1.0 : 0.0, tag, out5_dev);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}

// ---------------------------------------------------------------------------------------
// likelihoods
// ---------------------------------------------------------------------------------------
namespace {
int loglike_host(unb_ctx *ctx, int kind, const double *params, size_t d, size_t n, double *like,
                 const double *lparams)
{
    UNB_TRY(check_ctx(ctx));
    UNB_TRY(check_dims(ctx, n, d));
    if (n == 0) return UNB_OK;
    if (!params || !like) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    if (kind == UNB_LOGLIKE_ROSENBROCK && d < 2) return unb_fail(ctx, UNB_ERR_ARG, "rosenbrock needs ndim >= 2");
    Lane &ln = ctx->lane[0];
    cudaStream_t s = ln.stream;
    UNB_TRY(upload_lparams(ctx, kind, lparams, d, s));
    UNB_TRY(unb_reserve(ctx, ln.cand, n * d * sizeof(double)));
    UNB_TRY(unb_reserve(ctx, ln.like, n * sizeof(double)));
    UNB_TRY(h2d(ctx, ln.cand.p, params, n * d * sizeof(double), s));
    UNB_TRY(unb_launch_loglike(ctx, kind, (const double *)ln.cand.p, (int)d, (long long)n,
                               (double *)ln.like.p, nullptr, (const double *)ctx->lparams.p, s));
    UNB_TRY(d2h(ctx, like, ln.like.p, n * sizeof(double), s));
    UNB_CUDA(ctx, cudaStreamSynchronize(s));
    return UNB_OK;
}
}  // namespace

extern "C" int unb_loglike_gauss(unb_ctx *ctx, const double *params, size_t d, size_t n,
                                 double *like, const double *centers, double sigma,
                                 double norm_const)
{
    if (!ctx || !centers) return UNB_ERR_ARG;
    std::vector<double> lp(d + 2);
    memcpy(lp.data(), centers, d * sizeof(double));
    lp[d] = sigma;
    lp[d + 1] = norm_const;
    return loglike_host(ctx, UNB_LOGLIKE_GAUSS, params, d, n, like, lp.data());
}

extern "C" int unb_loglike_rosenbrock(unb_ctx *ctx, const double *params, size_t d, size_t n,
                                      double *like)
{
    return loglike_host(ctx, UNB_LOGLIKE_ROSENBROCK, params, d, n, like, nullptr);
}

extern "C" int unb_loglike_eggbox(unb_ctx *ctx, const double *params, size_t d, size_t n,
                                  double *like)
{
    return loglike_host(ctx, UNB_LOGLIKE_EGGBOX, params, d, n, like, nullptr);
}

extern "C" int unb_loglike_gauss_dev(unb_ctx *ctx, const double *params_dev, size_t d, size_t n,
                                     double *like_dev, const uint8_t *mask_dev,
                                     const double *centers, double sigma, double norm_const,
                                     void *stream)
{
    UNB_TRY(check_ctx(ctx));
    if (n == 0) return UNB_OK;
    if (!params_dev || !like_dev || !centers) return unb_fail(ctx, UNB_ERR_ARG, "null pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : S0(ctx);
    std::vector<double> lp(d + 2);
    memcpy(lp.data(), centers, d * sizeof(double));
    lp[d] = sigma;
    lp[d + 1] = norm_const;
    UNB_TRY(upload_lparams(ctx, UNB_LOGLIKE_GAUSS, lp.data(), d, s));
    return unb_launch_loglike(ctx, UNB_LOGLIKE_GAUSS, params_dev, (int)d, (long long)n, like_dev,
                              mask_dev, (const double *)ctx->lparams.p, s);
}
