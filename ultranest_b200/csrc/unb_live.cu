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

}  // namespace

int unb_launch_live_update(unb_ctx *ctx, const LiveUpdateArgs &a, cudaStream_t s)
{
    if (a.nrows <= 0) return UNB_OK;
    k_live_update<<<(unsigned)((a.nrows + 63) / 64), 64, 0, s>>>(a);
    ctx->launches++;
    UNB_CUDA(ctx, cudaGetLastError());
    return UNB_OK;
}
