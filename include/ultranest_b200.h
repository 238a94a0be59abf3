/*
 * ultranest_b200.h -- C ABI of the B200-native MLFriends region engine.
 *
 * This is the drop-in boundary for ONE hot path of JohannesBuchner/UltraNest: the MLFriends
 * region subsystem of ultranest/mlfriends.pyx plus the vectorised-likelihood batch call.
 * Every entry point names the reference interface it replaces (file:line relative to the
 * reference checkout).  Conventions, all taken from the reference's own foreign-function
 * precedent (languages/c/mylib.c:24-41, languages/c/runc.py:8-28):
 *
 *   - plain pointers and sizes only; arrays are C-contiguous, row-major, float64 unless
 *     stated; indices are int64 (mlfriends.pyx:25-26); masks are 1 byte per row (NumPy bool);
 *   - the CALLER owns every host buffer, outputs are caller-allocated (mlfriends.pyx:147:
 *     `nnearby` is an out-parameter); the library owns device memory behind `unb_ctx`;
 *     no host pointer is retained after a call returns;
 *   - every function returns UNB_OK (0) or a negative UNB_ERR_* code; the message is at
 *     unb_last_error(ctx).  The Python shim maps codes to the reference's exception types;
 *   - one host thread per ctx, results are valid on return (synchronous semantics), except
 *     the `_dev` variants, which take DEVICE pointers (e.g. torch.Tensor.data_ptr()) plus a
 *     cudaStream_t (as void*; NULL = the ctx's own stream) and only enqueue work.
 *
 * There is no CPU fallback anywhere behind this ABI: a missing/unsupported GPU is an error.
 */
#ifndef ULTRANEST_B200_H
#define ULTRANEST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UNB_ABI_VERSION 1

#define UNB_OK 0
#define UNB_ERR_CUDA (-1)        /* CUDA runtime/driver failure (incl. no device)        */
#define UNB_ERR_ARG (-2)         /* bad argument (shape mismatch, NULL, d too large ...) */
#define UNB_ERR_STATE (-3)       /* region state incomplete for the requested call       */
#define UNB_ERR_NOMEM (-4)       /* host or device allocation failed                     */
#define UNB_ERR_NUMERIC (-5)     /* non-positive distances etc. (-> LinAlgError)         */

/* unb_ctx_set_option keys */
#define UNB_OPT_EXACT_ONLY 1     /* 1: bypass the filtered scans, run the plain exact-order
                                    kernels (slow; for validation)                        */
#define UNB_OPT_CHUNK_ROWS 2     /* rows per H2D/compute/D2H pipeline chunk (host API)    */
#define UNB_OPT_FILTER_FP32 3    /* 1 (default): the membership kernel pre-filters pairs in
                                    fp32 (decisions stay exact fp64); 0: fp64 filter only   */

#define UNB_OPT_SURE_LEVEL 4     /* 1 (default): the fp32 membership filter retires a proposal on a
                                    CERTAIN neighbour (filter value above the whole error budget)
                                    without an exact evaluation; 0: every flagged pair is decided
                                    in exact fp64 (A/B switch of the parity tests)            */
#define UNB_OPT_COOP_MAX 5       /* survivors per block at which the membership kernel's drain turns
                                    cooperative (0 disables; default 24)                       */
#define UNB_OPT_BLOCK_KERNEL 6   /* 1: always the block-synchronous membership kernel, also for
                                    small launches (default 0: warp-independent kernel there)  */

#define UNB_OPT_BIN_MIN_ROWS 7    /* membership launches of at least this many proposals against the
                                    region's live block use clustered live tiles and proposals binned
                                    by nearest tile centroid (default 0 = never: on B200 the binning
                                    passes cost more than the scan gains).  Scheduling only: masks
                                    do not depend on it                                          */

/* unb_ctx_get_stat keys */
#define UNB_STAT_KERNEL_LAUNCHES 1   /* kernels launched by this ctx since creation       */
#define UNB_STAT_RECHECKS 2          /* filtered-scan pairs that went to the exact path
                                        in the last scan call (diagnostic)                */
#define UNB_STAT_H2D_BYTES 3
#define UNB_STAT_D2H_BYTES 4
#define UNB_STAT_UNCERTAIN 6         /* exact membership decisions of the last host-buffer
                                        unb_region_inside / _friends / _inside_loglike / _refill call
                                        that lie within the transform tolerance of the radius
                                        (unb_region_set_transform_tolerance); 0 = every decision is
                                        the reference's                                          */
#define UNB_STAT_TILE_VISITS 5       /* warp x tile filter passes of the last
                                        unb_region_find_nearby_dev(mask-only) call (diagnostic) */

/* transform-layer kinds for unb_region_set_layer */
#define UNB_LAYER_IDENTITY 0
#define UNB_LAYER_SCALING 1      /* ScalingLayer.transform, mlfriends.pyx:605-611          */
#define UNB_LAYER_AFFINE 2       /* AffineLayer.transform,  mlfriends.pyx:737-743          */

/* likelihood kinds for the fused inside+loglike call */
#define UNB_LOGLIKE_NONE 0
#define UNB_LOGLIKE_GAUSS 1
#define UNB_LOGLIKE_EGGBOX 2
#define UNB_LOGLIKE_ROSENBROCK 3

typedef struct unb_ctx unb_ctx;

/* ------------------------------------------------------------------ context */
int unb_abi_version(void);
int unb_ctx_create(int device, unb_ctx **out);
int unb_ctx_destroy(unb_ctx *ctx);
const char *unb_last_error(const unb_ctx *ctx);
int unb_ctx_set_option(unb_ctx *ctx, int option, int64_t value);
int unb_ctx_get_stat(unb_ctx *ctx, int stat, int64_t *value);
int unb_ctx_synchronize(unb_ctx *ctx);
/* measured fp64 FMA rate of the device (lane-FMAs per second): the compute-roofline denominator
 * bench.py reports next to the HBM one */
int unb_fp64_peak(unb_ctx *ctx, double *dfma_per_s);
int unb_fp32_peak(unb_ctx *ctx, double *ffma_per_s);
/* the same probe by instruction form: 0 scalar FFMA, 1 packed FFMA2 (fma.rn.f32x2, register
 * pairs), 2 FFMA2 with a scalar (broadcast) multiplicand -- the form of the membership filter */
int unb_fp32_peak_form(unb_ctx *ctx, int form, double *lane_fma_per_s);

/* --------------------------------------------- stateless scans (host buffers) */

/* replaces find_nearby(apts, bpts, radiussq, nnearby), mlfriends.pyx:143-183:
 * nnearby[j] = FIRST i with ||apts[i]-bpts[j]||^2 <= radiussq (fp64, k-sequential,
 * non-fused), else -1.  Bit-exact. */
int unb_find_nearby(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                    size_t nb, size_t ndim, double radiussq, int64_t *nnearby);

/* `find_nearby(...) >= 0` as a mask, without the index: the absorption test of _update_clusters
 * (mlfriends.pyx:304-307).  Uses the any-neighbour kernel. */
int unb_has_neighbour(unb_ctx *ctx, const double *apts, size_t na, const double *bpts, size_t nb,
                      size_t ndim, double radiussq, uint8_t *mask);

/* replaces count_nearby (cdef), mlfriends.pyx:31-68: number of i within radiussq. */
int unb_count_nearby(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                     size_t nb, size_t ndim, double radiussq, int64_t *nnearby);

/* replaces _subtract_nearby(apts, bpts, radiussq), mlfriends.pyx:73-113. */
int unb_subtract_nearby(unb_ctx *ctx, const double *apts, size_t n, size_t ndim,
                        double radiussq, double *bpts_out);

/* replaces compute_maxradiussq(apts, bpts) (cdef float), mlfriends.pyx:188-224:
 * max_j min_i ||a_i-b_j||^2, returned rounded to float32 like the C `float` return. */
int unb_compute_maxradiussq(unb_ctx *ctx, const double *apts, size_t na, const double *bpts,
                            size_t nb, size_t ndim, double *maxd_out);

/* replaces compute_mean_pair_distance(pts, clusterids), mlfriends.pyx:229-270.
 * (Deterministic but tree-ordered sum: agrees with the sequential reference to ~1e-13
 * relative, not bit-exact; documented in DESIGN.md.) */
int unb_mean_pair_distance(unb_ctx *ctx, const double *pts, const int64_t *clusterids,
                           size_t n, size_t ndim, double *out);

/* replaces _inside_ellipsoid(points, center, invcov, square_radius), mlfriends.pyx:882-912
 * (einsum 'ij,jk,ik->i' order: acc += (d_j*A_jk)*d_k, j outer, k inner; `<=`). Bit-exact. */
int unb_inside_ellipsoid(unb_ctx *ctx, const double *points, size_t m, size_t ndim,
                         const double *center, const double *invcov, double square_radius,
                         uint8_t *mask);

/* replace ScalingLayer/AffineLayer.transform / .untransform, mlfriends.pyx:605-620, 737-752.
 * Scaling is bit-exact; affine uses the DEFINED order (k-sequential FMA) because the
 * reference's np.dot is BLAS-kernel specific (see DESIGN.md "Affine transform"). */
int unb_transform_scaling(unb_ctx *ctx, const double *w, size_t m, size_t ndim,
                          const double *mean, const double *std, double *out);
int unb_untransform_scaling(unb_ctx *ctx, const double *ww, size_t m, size_t ndim,
                            const double *mean, const double *std, double *out);
int unb_transform_affine(unb_ctx *ctx, const double *w, size_t m, size_t ndim,
                         const double *ctr, const double *T, double *out);
int unb_untransform_affine(unb_ctx *ctx, const double *ww, size_t m, size_t ndim,
                           const double *ctr, const double *invT, double *out);

/* ------------------------------------------------- stateful region (device mirror) */

/* Mirror of MLFriends.unormed (the t-space live block, mlfriends.pyx:982).  The library
 * keeps a host snapshot; calling this again with a mutated array (integrator.py:2753-2754
 * patches rows in place) re-uploads only the rows that changed.  *rows_changed may be NULL. */
int unb_region_sync_live(unb_ctx *ctx, const double *unormed, size_t n, size_t ndim,
                         int64_t *rows_changed);

/* transformLayer parameters used for candidates: kind SCALING: shift=mean[d], mat=std[d];
 * AFFINE: shift=ctr[d], mat=T[d*d]; IDENTITY: both NULL. */
int unb_region_set_layer(unb_ctx *ctx, int kind, const double *shift, const double *mat,
                         size_t ndim);

/* MLFriends.ellipsoid_center / ellipsoid_invcov / enlarge (mlfriends.pyx:1226-1227, 1254). */
int unb_region_set_ellipsoid(unb_ctx *ctx, const double *center, const double *invcov,
                             double enlarge, size_t ndim);

/* MLFriends.maxradiussq */
int unb_region_set_radius(unb_ctx *ctx, double maxradiussq);
/* AffineLayer.transform is np.dot(w - ctr, T) (mlfriends.pyx:743), i.e. OpenBLAS dgemm, whose
 * summation order no other implementation can reproduce; the fused calls whiten proposals on the
 * device in a defined order instead, so the two t-rows differ by a few ulp.  `tau` is the caller's
 * bound on what that can do to a pair distance near the radius.  The filters are built for
 * radius + tau, decisions compare with the radius, and every exact decision with
 * |D - radius| <= tau is counted (UNB_STAT_UNCERTAIN) so that the caller can re-decide the call with
 * the reference's own transform.  0 (default) switches the reporting off. */
int unb_region_set_transform_tolerance(unb_ctx *ctx, double tau);

/* replaces MLFriends.inside(pts), mlfriends.pyx:1186-1211 (ellipsoid -> transform ->
 * find_nearby >= 0) in one device pipeline.  idx_out (optional, may be NULL) receives the
 * first-neighbour index or -1. */
int unb_region_inside(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                      int64_t *idx_out);
int unb_region_inside_dev(unb_ctx *ctx, const double *pts_dev, size_t m, uint8_t *mask_dev,
                          void *stream);
/* Ellipsoid stage alone against the mirrored ellipsoid, through the same chunked pipeline:
 * RobustEllipsoidRegion / SimpleRegion / WrappingEllipsoid `.inside` (mlfriends.pyx:1374-1390,
 * 1628-1649).  Needs only unb_region_set_ellipsoid. */
int unb_region_inside_ellipsoid(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask);
int unb_region_inside_ellipsoid_dev(unb_ctx *ctx, const double *pts_dev, size_t m,
                                    uint8_t *mask_dev, void *stream);
/* Same pipeline without the ellipsoid stage: transform + neighbour scan of u-space proposals,
 * i.e. `find_nearby(self.unormed, transformLayer.transform(w), ...) >= 0` of
 * sample_from_wrapping_ellipsoid (mlfriends.pyx:1155-1158). */
int unb_region_friends(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                       int64_t *idx_out);

/* find_nearby / count_nearby of t-space candidates against the mirrored live block
 * (mlfriends.pyx:1110, 1125, 1157, 1088). */
int unb_region_find_nearby(unb_ctx *ctx, const double *tpts, size_t m, int64_t *nnearby);
int unb_region_count_nearby(unb_ctx *ctx, const double *tpts, size_t m, int64_t *nnearby);
/* `find_nearby(self.unormed, tpts, ...) >= 0` without the index (any-neighbour kernel) */
int unb_region_has_neighbour(unb_ctx *ctx, const double *tpts, size_t m, uint8_t *mask);
/* device-pointer variant of the scan alone (nnearby_dev or mask_dev may be NULL) */
int unb_region_find_nearby_dev(unb_ctx *ctx, const double *tpts_dev, size_t m,
                               int64_t *nnearby_dev, uint8_t *mask_dev, void *stream);

/* Bootstrapped radius + ellipsoid enlargement, the device half of
 * MLFriends.compute_enlargement (mlfriends.pyx:1044-1066) for `nrounds` rounds at once:
 *   selected : nrounds x n bytes, the rounds' selection masks (host RNG, reference order);
 *   unormed  : n x d, t-space live block (NULL with maxd_out NULL: skip the radius scan);
 *   u        : n x d, u-space live block (NULL: skip the enlargement);
 *   ctrs     : nrounds x d, invcovs: nrounds x d x d, per-round bounding_ellipsoid centre and
 *              inv(cov) from the host (LAPACK stays on the host, as in the reference);
 *   maxd_out : nrounds, compute_maxradiussq(unormed[sel], unormed[~sel]) rounded to float32;
 *   f_out    : nrounds, einsum('ij,jk,ik->i', u[~sel]-ctr, a, u[~sel]-ctr).max()
 *              (skipped when u/ctrs/invcovs/f_out is NULL).
 * round_lo/round_hi select the slice of rounds this process computes (multi-GPU sharding,
 * integrator.py:388-404); other entries of the outputs are left untouched. */
int unb_region_bootstrap(unb_ctx *ctx, const double *unormed, const double *u, size_t n,
                         size_t ndim, const uint8_t *selected, size_t nrounds,
                         size_t round_lo, size_t round_hi, const double *ctrs,
                         const double *invcovs, double *maxd_out, double *f_out);

/* The same rounds, folded ON THE DEVICE into the buffer one allreduce(MAX) ships -- the collective
 * site of _update_region_bootstrap (integrator.py:388-404, comm.gather + np.max + comm.bcast there):
 *   out5_dev : DEVICE pointer to 5 doubles (e.g. a torch tensor handed to torch.distributed):
 *              [max_r maxd_r (each float32-rounded), max_r f_r, failed, tag, -tag];
 *   failed   : 1 if host_failed != 0 (the caller's d x d algebra failed on this rank) or a round has
 *              f <= 0 / non-finite (mlfriends.pyx:1063-1065); NaN under MAX is undefined in NCCL, so
 *              failure travels as its own flag (the reference ships NaN, integrator.py:391-393);
 *   tag      : any value that is equal on all ranks iff they hold the same selection masks (a
 *              checksum); after the MAX, out[3] == -out[4] proves it.
 * Only enqueues work on `stream` (NULL: the ctx's own stream); nothing is copied back. */
int unb_region_bootstrap_fold_dev(unb_ctx *ctx, const double *unormed, const double *u, size_t n,
                                  size_t ndim, const uint8_t *selected, size_t nrounds,
                                  size_t round_lo, size_t round_hi, const double *ctrs,
                                  const double *invcovs, int host_failed, double tag,
                                  double *out5_dev, void *stream);

/* First and second moments of the SELECTED rows of every bootstrap round about the reference point
 * c0[ndim], accumulated on the device:  sums[r][p] = sum_i y_ip,  sxx[r][p][q] = sum_i y_ip y_iq
 * (upper triangle q >= p filled), y = u[selected[r]] - c0;  counts[r] = selected rows.  They give the
 * per-round bounding_ellipsoid (mlfriends.pyx:426-476: mean, np.cov) up to summation order -- good
 * enough to SCREEN the rounds: MLFriends.compute_enlargement needs only max_r f_r, so only the
 * rounds whose screened f is within a margin of the maximum are recomputed with the reference's own
 * NumPy algebra (the result stays bit-identical, the host does 1-2 rounds of np.cov + inv instead of
 * 30).  Rounds outside [round_lo, round_hi) are left untouched. */
int unb_region_bootstrap_moments(unb_ctx *ctx, const double *u, size_t n, size_t ndim,
                                 const uint8_t *selected, size_t nrounds, size_t round_lo,
                                 size_t round_hi, const double *c0, double *sums, double *sxx,
                                 int64_t *counts);

/* ---------------------------------------------- vectorised likelihood batch call */

/* Same (params, d, n, like) shape as languages/c/mylib.c:33 my_c_likelihood_vectorized. */
/* docs/gauss.py:25-27; norm_const = 0.5*log(2*pi*sigma^2)*d evaluated by the host. Bit-exact
 * w.r.t. NumPy (pairwise-sum order reproduced). */
int unb_loglike_gauss(unb_ctx *ctx, const double *params, size_t d, size_t n, double *like,
                      const double *centers, double sigma, double norm_const);
/* examples/testrosenbrock.py:10-13. Bit-exact w.r.t. NumPy. */
int unb_loglike_rosenbrock(unb_ctx *ctx, const double *params, size_t d, size_t n,
                           double *like);
/* examples/testeggbox.py:9-11 (cos/pow: few-ulp parity, not bit-exact). */
int unb_loglike_eggbox(unb_ctx *ctx, const double *params, size_t d, size_t n, double *like);

int unb_loglike_gauss_dev(unb_ctx *ctx, const double *params_dev, size_t d, size_t n,
                          double *like_dev, const uint8_t *mask_dev, const double *centers,
                          double sigma, double norm_const, void *stream);

/* Fused proposal evaluation (integrator.py:1776-1804 without the host
 * compaction): mask = region.inside(pts); like[j] = loglike(pts[j]) where mask[j], else -inf.
 * One H2D of pts, D2H of mask+like.  lparams: GAUSS -> centers[d], sigma, norm_const packed as
 * d+2 doubles; others NULL. */
int unb_region_inside_loglike(unb_ctx *ctx, const double *pts, size_t m, uint8_t *mask,
                              double *like, int loglike_kind, const double *lparams);
int unb_region_inside_loglike_dev(unb_ctx *ctx, const double *pts_dev, size_t m,
                                  uint8_t *mask_dev, double *like_dev, int loglike_kind,
                                  const double *lparams, void *stream);

/* ---- fused refill: ReactiveNestedSampler._refill_samples,
 * integrator.py:1773-1837, as ONE device pipeline per batch of proposals:
 *   region membership  (MLFriends.inside, mlfriends.pyx:1186-1211; or the friends test alone for
 *                       draws that already lie in the wrapping ellipsoid, mlfriends.pyx:1155-1160)
 *   -> v = transform(u)                       (integrator.py:1790; identity or u*scale+lo)
 *   -> tregion.inside(v)                      (integrator.py:1793-1796; WrappingEllipsoid.inside,
 *                                              mlfriends.pyx:1628-1649, all dimensions variable)
 *   -> logl = loglike(v) for the survivors    (integrator.py:1802-1803), -inf elsewhere
 *   -> accepted = logl > Lmin                 (integrator.py:1805)
 * One H2D of the proposals; D2H of one flag byte and one double per row.
 * flags[j]: bit0 = region member (what region.sample() would have returned),
 *           bit1 = also inside tregion (a likelihood call was spent: `nc`),
 *           bit2 = logl > Lmin.
 * counts[3] = number of rows with bit0 / bit1 / bit2. */
#define UNB_XFORM_IDENTITY 0
#define UNB_XFORM_SCALE_SHIFT 1   /* v = u * scale + lo, two roundings like NumPy's `u * scale + lo` */

#define UNB_REFILL_MEMBER 1
#define UNB_REFILL_TREGION 2
#define UNB_REFILL_ACCEPTED 4

typedef struct unb_refill_desc {
    int32_t region_mode;        /* 0: rows are members already (no region test); 1: friends test
                                   only; 2: wrapping ellipsoid + friends (MLFriends.inside) */
    int32_t check_cube;         /* rows with a coordinate outside (0,1) are no members
                                   (mlfriends.pyx:1154) */
    int32_t xform_kind;         /* UNB_XFORM_* */
    int32_t loglike_kind;       /* UNB_LOGLIKE_* (not NONE) */
    const double *xform_scale;  /* [ndim] (SCALE_SHIFT) */
    const double *xform_lo;     /* [ndim] */
    const double *treg_center;  /* [ndim]; NULL: no tregion (integrator.py:1797-1800) */
    const double *treg_invcov;  /* [ndim x ndim] */
    double treg_enlarge;
    const double *lparams;      /* as for unb_region_inside_loglike */
    double Lmin;
} unb_refill_desc;

int unb_region_refill(unb_ctx *ctx, const double *u, size_t m, size_t ndim,
                      const unb_refill_desc *desc, uint8_t *flags, double *like,
                      int64_t *counts);

/* ------------------------------------------------ device-side proposal generation
 * "Throughput mode" of MLFriends.sample (mlfriends.pyx:1162-1184): the draws of
 * sample_from_wrapping_ellipsoid (mlfriends.pyx:1135-1160) or sample_from_boundingbox
 * (mlfriends.pyx:1096-1112) are made ON THE DEVICE by a counter-based generator (Philox4x32-10;
 * key = seed, counter = offset + row), filtered by the region exactly like the host-RNG path
 * (unit cube, wrapping ellipsoid, neighbour test; optionally the likelihood and `like > Lmin`), and
 * only the accepted rows are returned, in draw order.  No proposal crosses PCIe on the way in.
 * NOT the reference's random stream (np.random MT19937): statistically equivalent proposals, a
 * different seeded sequence; the parity path (host RNG) is the default everywhere. */
#define UNB_SAMPLE_WRAPPING_ELLIPSOID 0
#define UNB_SAMPLE_UNIT_CUBE 1

typedef struct unb_sample_desc {
    int32_t method;             /* UNB_SAMPLE_* */
    int32_t loglike_kind;       /* UNB_LOGLIKE_*; NONE: no likelihood */
    int32_t use_lmin;           /* 1: drop rows with like <= Lmin (needs a likelihood) */
    int32_t reserved;
    uint64_t seed;              /* generator key */
    uint64_t offset;            /* global index of the first draw of this call: calls with
                                   disjoint [offset, offset + nsamples) ranges never overlap */
    const double *axes_T;       /* [ndim x ndim] MLFriends.ellipsoid_axes_T (WRAPPING_ELLIPSOID) */
    const double *lparams;      /* as for unb_region_inside_loglike */
    double Lmin;
} unb_sample_desc;

/* rows_out: caller-allocated nsamples x ndim; like_out: nsamples or NULL; *n_out = accepted rows;
 * counts (nullable) = [draws inside the unit cube, region members, accepted]. */
int unb_region_sample(unb_ctx *ctx, const unb_sample_desc *desc, size_t nsamples, double *rows_out,
                      double *like_out, int64_t *n_out, int64_t *counts);
/* device-resident variant: rows_out_dev (nsamples x ndim), like_out_dev (nullable), n_out_dev (one
 * int32) are DEVICE pointers; only enqueues work on `stream`. */
int unb_region_sample_dev(unb_ctx *ctx, const unb_sample_desc *desc, size_t nsamples,
                          double *rows_out_dev, double *like_out_dev, int32_t *n_out_dev,
                          void *stream);
/* the raw draws (before any region filter) and their unit-cube mask: lets a test restate the
 * generator on the host.  center / axes_T / enlarge as MLFriends.ellipsoid_center, .ellipsoid_axes_T,
 * .enlarge; ignored for UNB_SAMPLE_UNIT_CUBE. */
int unb_sample_draw(unb_ctx *ctx, int method, size_t nsamples, size_t ndim, uint64_t seed,
                    uint64_t offset, const double *center, const double *axes_T, double enlarge,
                    double *rows_out, uint8_t *cube_out);

/* ------------------------------------------------ population step-sampler helpers
 * (SURVEY 8-f rank 2): ultranest/stepfuncs.pyx, the compiled helpers of the vectorised slice
 * samplers in ultranest/popstepsampler.py.  Boolean arrays are NumPy bool (1 byte), integer
 * arrays int64 (stepfuncs.pyx:16-17).  Arrays the reference updates in place are in/out here.
 * Random draws are made by the host (np.random, reference order) and passed in as arrays. */

/* within_unit_cube(u), stepfuncs.pyx:22-52: acceptable[i] = all(0 < u[i,:] < 1). */
int unb_within_unit_cube(unb_ctx *ctx, const double *u, size_t n, size_t ndim, uint8_t *acceptable);

/* evolve_prepare(searching_left, searching_right), stepfuncs.pyx:57-94. */
int unb_evolve_prepare(unb_ctx *ctx, const uint8_t *searching_left, const uint8_t *searching_right,
                       size_t n, uint8_t *search_right, uint8_t *bisecting);

/* evolve_update(...), stepfuncs.pyx:99-183.  Lnew[n_lnew]: one value per acceptable walker in
 * walker order; currentt, current_left, current_right, searching_left, searching_right and
 * success are updated in place. */
int unb_evolve_update(unb_ctx *ctx, const uint8_t *acceptable, const double *Lnew, size_t n_lnew,
                      double Lmin, const uint8_t *search_right, const uint8_t *bisecting,
                      double *currentt, double *current_left, double *current_right,
                      uint8_t *searching_left, uint8_t *searching_right, uint8_t *success, size_t n);

/* Built-in prior transform + likelihood evaluated inside the fused step kernels
 * (UNB_XFORM_*, UNB_LOGLIKE_*; parameters as for unb_refill_desc). */
typedef struct unb_step_desc {
    int32_t xform_kind;
    int32_t loglike_kind;
    const double *xform_scale;  /* [ndim] (SCALE_SHIFT) */
    const double *xform_lo;     /* [ndim] */
    const double *lparams;      /* GAUSS: centers[ndim], sigma, norm_const */
} unb_step_desc;

/* evolve(transform, loglike, Lmin, ...), stepfuncs.pyx:189-282, as ONE kernel for a device
 * transform + likelihood: proposal on the slice (currentu is overwritten by the proposals, the
 * reference's `unew = currentu` alias, :252), unit-cube test, v = transform(u), L = loglike(v),
 * slice-state update.  currentt of the bisecting walkers must already hold the uniform draws
 * of :255.  Out: acceptable[n] (nc = its sum), success[n], like[n] (-inf where not acceptable). */
int unb_evolve(unb_ctx *ctx, const unb_step_desc *desc, double Lmin, double *currentu,
               const double *currentv, double *currentt, double *current_left,
               double *current_right, uint8_t *searching_left, uint8_t *searching_right, size_t n,
               size_t ndim, uint8_t *acceptable, uint8_t *success, double *like);

/* step_back(Lmin, allL, generation, currentt), stepfuncs.pyx:285-334; allL is
 * (nwalkers x ncols) row-major, ncols <= 2048.  In place. */
int unb_step_back(unb_ctx *ctx, double Lmin, double *allL, size_t nwalkers, size_t ncols,
                  int64_t *generation, double *currentt);

/* update_vectorised_slice_sampler(...), stepfuncs.pyx:537-630.  In place on tleft, tright,
 * worker_running, status, allu, allL, allp; *discarded = the returned count. */
int unb_update_vectorised_slice_sampler(
    unb_ctx *ctx, const double *t, double *tleft, double *tright, const double *proposed_L,
    const double *proposed_u, const double *proposed_p, int64_t *worker_running, int64_t *status,
    double likelihood_threshold, double shrink_factor, double *allu, double *allL, double *allp,
    size_t popsize, size_t ndim, size_t nparams, int64_t *discarded);

/* Inner loop of PopulationSimpleSliceSampler.__next__ (popstepsampler.py:916-965) with the
 * population resident on the device:
 *   begin   : upload allu, allL, v and the slice limits; allp = NaN, worker_running = arange,
 *             status = 0 (popstepsampler.py:908, 932-934);
 *   iterate : one pass -- t from the uniform draws (:942-943), proposals (:945-947), device
 *             transform + likelihood (:949-950), update_vectorised_slice_sampler (:954-957);
 *             returns the number of points still running and the discarded evaluations;
 *   end     : download the state (any pointer may be NULL).
 * Any other step-helper call on the same context ends the session. */
int unb_popslice_begin(unb_ctx *ctx, const unb_step_desc *desc, const double *allu,
                       const double *allL, const double *v, const double *tleft,
                       const double *tright, size_t popsize, size_t ndim,
                       double likelihood_threshold, double shrink_factor);
int unb_popslice_iterate(unb_ctx *ctx, const double *slice_position, int64_t *n_running,
                         int64_t *discarded);
int unb_popslice_end(unb_ctx *ctx, double *allu, double *allp, double *allL, double *tleft,
                     double *tright, int64_t *status);

#ifdef __cplusplus
}
#endif
#endif /* ULTRANEST_B200_H */
