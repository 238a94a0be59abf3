/*
 * oracle/stepfuncs_oracle.c -- CPU restatement of the population step-sampler helpers of
 * ultranest/stepfuncs.pyx.
 *
 * TEST INFRASTRUCTURE ONLY: the checker for ultranest_b200/csrc/unb_stepfuncs.cu; never linked
 * into, imported by, or called from the product path.
 *
 * Every function restates one reference loop in plain C and cites the reference file:line it
 * follows (paths relative to /root/reference).  Boolean arrays are NumPy bool = one byte;
 * integer arrays are int64 (stepfuncs.pyx:16-17).  Pinned bit-for-bit against the compiled
 * reference (oracle/_ref/ultranest/stepfuncs*.so) by tests/test_stepfuncs_oracle.py.
 *
 * Build:  gcc -O2 -ffp-contract=off -fPIC -shared -o libstepfuncs_oracle.so stepfuncs_oracle.c -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* within_unit_cube: stepfuncs.pyx:22-52 -- every coordinate strictly inside (0, 1) */
void sfo_within_unit_cube(const double *u, size_t popsize, size_t ndim, uint8_t *acceptable)
{
    for (size_t i = 0; i < popsize; i++) {
        acceptable[i] = 1;
        for (size_t j = 0; j < ndim; j++) {
            double x = u[i * ndim + j];
            if (!(0.0 < x && x < 1.0)) {
                acceptable[i] = 0;
                break;
            }
        }
    }
}

/* evolve_prepare: stepfuncs.pyx:57-94 */
void sfo_evolve_prepare(const uint8_t *searching_left, const uint8_t *searching_right, size_t n,
                        uint8_t *search_right, uint8_t *bisecting)
{
    for (size_t i = 0; i < n; i++) {
        search_right[i] = !searching_left[i] && searching_right[i];
        bisecting[i] = !(searching_left[i] || searching_right[i]);
    }
}

/* evolve_update: stepfuncs.pyx:99-183.  Lnew holds one value per acceptable walker, in
 * walker order (:152-156); everything else is per walker and written in place. */
void sfo_evolve_update(const uint8_t *acceptable, const double *Lnew, double Lmin,
                       const uint8_t *search_right, const uint8_t *bisecting, double *currentt,
                       double *current_left, double *current_right, uint8_t *searching_left,
                       uint8_t *searching_right, uint8_t *success, size_t popsize)
{
    size_t j = 0;
    for (size_t k = 0; k < popsize; k++) {
        if (acceptable[k]) {
            if (Lnew[j] > Lmin) success[k] = 1;
            j++;
        }
    }
    for (size_t i = 0; i < popsize; i++) {
        if (success[i] != 0) {            /* :161-165 step out further while still accepting */
            if (searching_left[i]) current_left[i] *= 2;
            else if (search_right[i]) current_right[i] *= 2;
        } else {                          /* :167-171 done stepping out when rejected */
            if (searching_left[i]) searching_left[i] = 0;
            else if (search_right[i]) searching_right[i] = 0;
        }
        if (bisecting[i]) {               /* :173-181 */
            if (currentt[i] < 0) current_left[i] = currentt[i];
            else current_right[i] = currentt[i];
            if (success[i] != 0) currentt[i] = (double)NAN;
        } else {
            success[i] = 0;               /* :183 */
        }
    }
}

/* step_back: stepfuncs.pyx:285-334.  allL is (nwalkers x ncols) row-major.  The reference
 * builds below_threshold = allL[:, :max_width] < Lmin once (:308-309) and then peels entries
 * off the back of each problematic walker's chain until no flagged entry is left; walkers are
 * independent, so the `while True` sweep (:319-334) is restated walker by walker.  NumPy's
 * negative-index wrap of allL[i, g] / below[.., g] is kept for fidelity. */
void sfo_step_back(double Lmin, double *allL, size_t nwalkers, size_t ncols, int64_t *generation,
                   double *currentt)
{
    if (nwalkers == 0) return;
    int64_t max_width = generation[0];
    for (size_t i = 1; i < nwalkers; i++)
        if (generation[i] > max_width) max_width = generation[i];
    max_width += 1;
    if (max_width > (int64_t)ncols) max_width = (int64_t)ncols;   /* slice clamps */
    if (max_width <= 0) return;
    uint8_t below[max_width];
    for (size_t i = 0; i < nwalkers; i++) {
        int remaining = 0;
        for (int64_t c = 0; c < max_width; c++) {
            below[c] = allL[i * ncols + c] < Lmin;
            remaining += below[c];
        }
        while (remaining > 0) {
            int64_t g = generation[i];
            generation[i] -= 1;
            currentt[i] = (double)NAN;
            int64_t ga = g < 0 ? g + (int64_t)ncols : g;      /* allL[i, g] */
            allL[i * ncols + ga] = (double)NAN;
            int64_t gb = g < 0 ? g + max_width : g;           /* below_threshold_parent[.., g] */
            if (below[gb]) {
                below[gb] = 0;
                remaining--;
            }
        }
    }
}

/* update_vectorised_slice_sampler: stepfuncs.pyx:537-630.  Sequential over the workers l: the
 * slice of point worker_running[l] shrinks, and the first worker whose proposal beats the
 * threshold moves the point.  Returns `discarded`. */
int64_t sfo_update_vectorised_slice_sampler(
    const double *t, double *tleft, double *tright, const double *proposed_L,
    const double *proposed_u, const double *proposed_p, int64_t *worker_running, int64_t *status,
    double Likelihood_threshold, double shrink_factor, double *allu, double *allL, double *allp,
    int64_t popsize, size_t ndim, size_t nparams)
{
    int64_t discarded = 0;
    for (int64_t l = 0; l < popsize; l++) {
        int64_t w = worker_running[l];
        if (t[l] > tright[w] || t[l] < tleft[w]) {            /* :609-612 */
            if (proposed_L[l] > Likelihood_threshold) discarded += 1;
            continue;
        }
        if (0 < t[l] && t[l] < tright[w]) tright[w] = t[l] / shrink_factor;   /* :613-614 */
        if (0 > t[l] && t[l] > tleft[w]) tleft[w] = t[l] / shrink_factor;     /* :615-616 */
        if (proposed_L[l] > Likelihood_threshold && status[w] == 0) {         /* :617-621 */
            status[w] = 1;
            memcpy(allu + w * ndim, proposed_u + l * ndim, ndim * sizeof(double));
            allL[w] = proposed_L[l];
            memcpy(allp + w * nparams, proposed_p + l * nparams, nparams * sizeof(double));
        }
    }
    /* :623-628 hand the workers to the points still running, round robin in point order */
    int any_running = 0;
    for (int64_t k = 0; k < popsize; k++) any_running |= status[k] == 0;
    int64_t j = 0;
    while (j < popsize && any_running) {
        for (int64_t k = 0; k < popsize; k++) {
            if (status[k] == 0 && j < popsize) {
                worker_running[j] = k;
                j += 1;
            }
        }
    }
    return discarded;
}
