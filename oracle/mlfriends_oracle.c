/*
 * oracle/mlfriends_oracle.c -- CPU restatement of the UltraNest MLFriends hot loops.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the *checker* for the CUDA product in
 * ultranest_b200/csrc; it is never linked into, imported by, or called from the
 * product path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.
 *
 * Every function restates, in plain C, the arithmetic of one reference loop and cites
 * the reference file:line it follows (paths relative to /root/reference).  The
 * arithmetic contract that makes results bit-identical to the compiled Cython:
 *   - fp64, k-sequential accumulation  d = d + (a-b)*(a-b)   (no FMA contraction;
 *     built with -ffp-contract=off and without -march, like the reference's -O3 build;
 *     Cython lowers `(x)**2` on doubles to pow(x, 2.0), which gcc folds to x*x),
 *   - comparisons are `<=`,
 *   - compute_maxradiussq returns a C float (mlfriends.pyx:188), i.e. (double)(float)maxd.
 * Parity pinning: tests/test_oracle_vs_reference.py checks every function here against
 * the compiled reference in oracle/_ref (and against committed golden vectors in
 * tests/golden/ generated from that reference).
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared -o liboracle.so mlfriends_oracle.c -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

/* squared distance, restating the k-loop shared by mlfriends.pyx:64-66, 102-104,
 * 178-180, 217-219, 263-265 */
static inline double sqdist(const double *a, const double *b, size_t ndim)
{
    double d = 0.0;
    for (size_t k = 0; k < ndim; k++) {
        double diff = a[k] - b[k];
        d = d + diff * diff;
    }
    return d;
}

/* count_nearby: mlfriends.pyx:31-68 */
void orc_count_nearby(const double *apts, size_t na, const double *bpts, size_t nb,
                      size_t ndim, double radiussq, int64_t *nnearby)
{
    for (size_t j = 0; j < nb; j++) {
        nnearby[j] = 0;
        for (size_t i = 0; i < na; i++) {
            double d = sqdist(apts + i * ndim, bpts + j * ndim, ndim);
            if (d <= radiussq)
                nnearby[j] += 1;
        }
    }
}

/* find_nearby: mlfriends.pyx:143-183 -- FIRST index i with d <= radiussq, else -1 */
void orc_find_nearby(const double *apts, size_t na, const double *bpts, size_t nb,
                     size_t ndim, double radiussq, int64_t *nnearby)
{
    for (size_t j = 0; j < nb; j++) {
        nnearby[j] = -1;
        for (size_t i = 0; i < na; i++) {
            double d = sqdist(apts + i * ndim, bpts + j * ndim, ndim);
            if (d <= radiussq) {
                nnearby[j] = (int64_t)i;
                break;
            }
        }
    }
}

/* _subtract_nearby: mlfriends.pyx:73-113.  Sum of all points within radiussq
 * (including the point itself), accumulated in ascending i, then
 * b[j,k] = a[j,k] - b[j,k] / float(nnearby). */
void orc_subtract_nearby(const double *apts, size_t n, size_t ndim, double radiussq,
                         double *bpts)
{
    for (size_t j = 0; j < n; j++) {
        size_t nnearby = 0;
        for (size_t k = 0; k < ndim; k++)
            bpts[j * ndim + k] = 0.0;
        for (size_t i = 0; i < n; i++) {
            double d = sqdist(apts + i * ndim, apts + j * ndim, ndim);
            if (d <= radiussq) {
                nnearby += 1;
                for (size_t k = 0; k < ndim; k++)
                    bpts[j * ndim + k] += apts[i * ndim + k];
            }
        }
        for (size_t k = 0; k < ndim; k++)
            bpts[j * ndim + k] = apts[j * ndim + k] - bpts[j * ndim + k] / (double)nnearby;
    }
}

/* compute_maxradiussq: mlfriends.pyx:188-224.  max_j min_i d(a_i,b_j); mind starts at
 * 1e300, maxd at 0; the C return type is `float`, so the double is rounded to
 * float32 on return. */
double orc_maxradiussq(const double *apts, size_t na, const double *bpts, size_t nb,
                       size_t ndim)
{
    double maxd = 0.0;
    for (size_t j = 0; j < nb; j++) {
        double mind = 1e300;
        for (size_t i = 0; i < na; i++) {
            double d = sqdist(apts + i * ndim, bpts + j * ndim, ndim);
            mind = (mind < d) ? mind : d;       /* min(mind, d) */
        }
        maxd = (maxd > mind) ? maxd : mind;     /* max(maxd, mind) */
    }
    return (double)(float)maxd;
}

/* same scan, but the un-rounded double (for tolerance studies only) */
double orc_maxradiussq_double(const double *apts, size_t na, const double *bpts, size_t nb,
                              size_t ndim)
{
    double maxd = 0.0;
    for (size_t j = 0; j < nb; j++) {
        double mind = 1e300;
        for (size_t i = 0; i < na; i++) {
            double d = sqdist(apts + i * ndim, bpts + j * ndim, ndim);
            mind = (mind < d) ? mind : d;
        }
        maxd = (maxd > mind) ? maxd : mind;
    }
    return maxd;
}

/* bootstrapped variant used by MLFriends.compute_enlargement / compute_maxradiussq
 * (mlfriends.pyx:1004-1012, 1044-1054): A = rows with selected[i] != 0 in original
 * order, B = the rest; no gather copies are needed to restate the arithmetic. */
double orc_maxradiussq_selected(const double *pts, size_t n, size_t ndim,
                                const uint8_t *selected)
{
    double maxd = 0.0;
    for (size_t j = 0; j < n; j++) {
        if (selected[j]) continue;
        double mind = 1e300;
        for (size_t i = 0; i < n; i++) {
            if (!selected[i]) continue;
            double d = sqdist(pts + i * ndim, pts + j * ndim, ndim);
            mind = (mind < d) ? mind : d;
        }
        maxd = (maxd > mind) ? maxd : mind;
    }
    return (double)(float)maxd;
}

/* compute_mean_pair_distance: mlfriends.pyx:229-270.  Sequential sum of sqrt over
 * same-cluster pairs (ids != 0), i < j; `Npairs` is a C int. Returns total/Npairs
 * (NaN/inf if Npairs == 0, like the reference's float division would raise). */
double orc_mean_pair_distance(const double *pts, const int64_t *clusterids, size_t na,
                              size_t ndim)
{
    double total_dist = 0.0;
    int npairs = 0;
    for (size_t j = 0; j < na; j++) {
        if (clusterids[j] == 0) continue;
        for (size_t i = 0; i < j; i++) {
            if (clusterids[j] == clusterids[i]) {
                double pair_dist = sqdist(pts + i * ndim, pts + j * ndim, ndim);
                total_dist += sqrt(pair_dist);
                npairs += 1;
            }
        }
    }
    return total_dist / npairs;
}

/* The quadratic form of np.einsum('ij,jk,ik->i', delta, A, delta) for one row, in NumPy's order:
 * acc += (delta_j * A_jk) * delta_k, j outer, k inner, sequential, no FMA -- and
 * NumPy reduces through its buffered iterator (np.getbufsize() = 8192 elements): the partial sum
 * restarts every floor(8192 / ndim) rows j and the partials are added to the result in order.
 * Up to ndim = 90 that is the plain sequential sum; from 91 on the chunking shows in the last bits
 * (probed bitwise against the compiled reference for ndim in {2 ... 150}). */
static double orc_einsum_quadform(const double *delta, const double *a, size_t ndim)
{
    size_t rows = 8192 / ndim;
    if (rows < 1) rows = 1;
    double total = 0.0;
    for (size_t j0 = 0; j0 < ndim; j0 += rows) {
        size_t j1 = j0 + rows < ndim ? j0 + rows : ndim;
        double acc = 0.0;
        for (size_t j = j0; j < j1; j++)
            for (size_t k = 0; k < ndim; k++)
                acc = acc + (delta[j] * a[j * ndim + k]) * delta[k];
        total = total + acc;
    }
    return total;
}

/* _inside_ellipsoid: mlfriends.pyx:882-912.  d = points - center (rounded), then
 * np.einsum('ij,jk,ik->i', d, invcov, d), whose accumulation order on NumPy >= 1.2x
 * is  acc += (d_j * A_jk) * d_k,  j outer, k inner, sequential, no FMA, in buffered chunks
 * (orc_einsum_quadform above); mask = r <= square_radius.
 * r_out may be NULL. */
void orc_inside_ellipsoid(const double *points, size_t m, size_t ndim,
                          const double *center, const double *invcov,
                          double square_radius, uint8_t *mask, double *r_out)
{
    double delta[1024];
    for (size_t p = 0; p < m; p++) {
        for (size_t k = 0; k < ndim; k++)
            delta[k] = points[p * ndim + k] - center[k];
        double acc = orc_einsum_quadform(delta, invcov, ndim);
        if (r_out) r_out[p] = acc;
        mask[p] = (acc <= square_radius) ? 1 : 0;
    }
}

/* ScalingLayer.transform: mlfriends.pyx:605-611, (w - mean) / std elementwise. */
void orc_transform_scaling(const double *w, size_t m, size_t ndim, const double *mean,
                           const double *std, double *out)
{
    for (size_t p = 0; p < m; p++)
        for (size_t k = 0; k < ndim; k++)
            out[p * ndim + k] = (w[p * ndim + k] - mean[k]) / std[k];
}

/* AffineLayer.transform: mlfriends.pyx:737-743 is np.dot(w - ctr, T), i.e. OpenBLAS
 * dgemm, whose summation order is kernel/CPU specific and NOT reproducible
 * (SURVEY fact 6).  The product therefore DEFINES the order below -- x = w - ctr
 * (rounded), out_j = fma(x_{d-1}, T[d-1][j], ... fma(x_0, T[0][j], 0)) -- and this
 * oracle restates that definition so the CUDA kernel can be checked bit-exactly.
 * Parity of the definition against np.dot is pinned only to 1e-13 relative
 * (tests), and mask parity through inside() is pinned against the reference on
 * the committed golden candidates. */
void orc_transform_affine(const double *w, size_t m, size_t ndim, const double *ctr,
                          const double *T, double *out)
{
    double x[1024];
    for (size_t p = 0; p < m; p++) {
        for (size_t k = 0; k < ndim; k++)
            x[k] = w[p * ndim + k] - ctr[k];
        for (size_t j = 0; j < ndim; j++) {
            double acc = 0.0;
            for (size_t k = 0; k < ndim; k++)
                acc = fma(x[k], T[k * ndim + j], acc);
            out[p * ndim + j] = acc;
        }
    }
}

/* AffineLayer.untransform: mlfriends.pyx:745-752, np.dot(ww, invT) + ctr; same
 * defined order as above, then + ctr (rounded). */
void orc_untransform_affine(const double *ww, size_t m, size_t ndim, const double *ctr,
                            const double *invT, double *out)
{
    for (size_t p = 0; p < m; p++) {
        for (size_t j = 0; j < ndim; j++) {
            double acc = 0.0;
            for (size_t k = 0; k < ndim; k++)
                acc = fma(ww[p * ndim + k], invT[k * ndim + j], acc);
            out[p * ndim + j] = acc + ctr[j];
        }
    }
}

/* NumPy's pairwise summation of a contiguous double row, as used by
 * ndarray.sum(axis=1) on a C-contiguous (n, d) array
 * (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum; PW_BLOCKSIZE=128):
 *   n < 8: sequential from 0.;  n <= 128: 8 accumulators, combined
 *   ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail added sequentially;
 *   n > 128: split at n2 = n/2, n2 -= n2 % 8, recurse.                           */
static double np_pairwise_sum(const double *a, size_t n)
{
    if (n < 8) {
        double res = 0.;
        for (size_t i = 0; i < n; i++)
            res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        size_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++)
            res += a[i];
        return res;
    } else {
        size_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

double orc_np_pairwise_sum(const double *a, size_t n) { return np_pairwise_sum(a, n); }

/* Vectorised Gaussian log-likelihood, docs/gauss.py:25-27 / examples/testfeatures.py:54-56:
 *   like = -0.5 * (((theta - centers)/sigma)**2).sum(axis=1) - 0.5*log(2*pi*sigma**2)*ndim
 * `norm_const` is the host-evaluated  0.5 * np.log(2*np.pi*sigma**2) * ndim  (Python
 * evaluates it left to right; passing it in keeps libm's log out of the comparison).
 * ABI shape follows languages/c/mylib.c:33 (params, d, n, like). */
void orc_loglike_gauss(const double *params, size_t d, size_t n, double *like,
                       const double *centers, double sigma, double norm_const)
{
    double t[4096];
    for (size_t j = 0; j < n; j++) {
        for (size_t i = 0; i < d; i++) {
            double z = (params[j * d + i] - centers[i]) / sigma;
            t[i] = z * z;
        }
        like[j] = -0.5 * np_pairwise_sum(t, d) - norm_const;
    }
}

/* Rosenbrock, examples/testrosenbrock.py:10-13:
 *   a = theta[:,:-1]; b = theta[:,1:]
 *   -2 * (100 * (b - a**2)**2 + (1 - a)**2).sum(axis=1)
 * The (n, d-1) temporary is C-contiguous, so sum(axis=1) is the pairwise sum above. */
void orc_loglike_rosenbrock(const double *params, size_t d, size_t n, double *like)
{
    double t[4096];
    for (size_t j = 0; j < n; j++) {
        for (size_t i = 0; i + 1 < d; i++) {
            double a = params[j * d + i], b = params[j * d + i + 1];
            double u = b - a * a;
            double v = 1 - a;
            t[i] = 100 * (u * u) + v * v;
        }
        like[j] = -2 * np_pairwise_sum(t, d - 1);
    }
}

/* Eggbox, examples/testeggbox.py:9-11:  chi = cos(z/2).prod(axis=1); (2 + chi)**5.
 * prod is sequential from 1.0; cos and pow come from libm here (NumPy uses its own
 * SIMD cos on some CPUs), so parity for this likelihood is tolerance-based
 * (few ulp), never bit-exact -- see DESIGN.md. */
void orc_loglike_eggbox(const double *params, size_t d, size_t n, double *like)
{
    for (size_t j = 0; j < n; j++) {
        double chi = 1.0;
        for (size_t i = 0; i < d; i++)
            chi *= cos(params[j * d + i] / 2.);
        like[j] = pow(2. + chi, 5);
    }
}

/* Bootstrap ellipsoid enlargement factor for one round, mlfriends.pyx:1060-1062:
 *   delta = u[~selected] - ctr;  f = einsum('ij,jk,ik->i', delta, a, delta).max()
 * with the same einsum order as orc_inside_ellipsoid.  Returns -inf for empty B. */
double orc_enlargement_f(const double *u, size_t n, size_t ndim, const uint8_t *selected,
                         const double *ctr, const double *a)
{
    double delta[1024];
    double f = -INFINITY;
    for (size_t p = 0; p < n; p++) {
        if (selected[p]) continue;
        for (size_t k = 0; k < ndim; k++)
            delta[k] = u[p * ndim + k] - ctr[k];
        double acc = orc_einsum_quadform(delta, a, ndim);
        if (acc > f) f = acc;
    }
    return f;
}
