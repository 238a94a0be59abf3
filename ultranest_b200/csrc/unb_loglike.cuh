// unb_loglike.cuh -- device-side row functions shared by the row kernels (unb_region.cu) and the
// population step-sampler kernels (unb_stepfuncs.cu): NumPy's pairwise summation order and the
// built-in vectorised likelihoods (docs/gauss.py:25-27, examples/testrosenbrock.py:10-13,
// examples/testeggbox.py:9-11).  Everything is non-fused unless it says fma().
#pragma once
#include "unb_internal.cuh"

// ---------------------------------------------------------------------------------------
// NumPy pairwise summation (numpy/_core/src/umath/loops_utils.h.src, PW_BLOCKSIZE = 128)
// ---------------------------------------------------------------------------------------
static __device__ double pw_block(const double *a, int n)
{
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r0 = a[0], r1 = a[1], r2 = a[2], r3 = a[3], r4 = a[4], r5 = a[5], r6 = a[6], r7 = a[7];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 = __dadd_rn(r0, a[i + 0]); r1 = __dadd_rn(r1, a[i + 1]);
        r2 = __dadd_rn(r2, a[i + 2]); r3 = __dadd_rn(r3, a[i + 3]);
        r4 = __dadd_rn(r4, a[i + 4]); r5 = __dadd_rn(r5, a[i + 5]);
        r6 = __dadd_rn(r6, a[i + 6]); r7 = __dadd_rn(r7, a[i + 7]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)),
                           __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
    for (; i < n; i++) res = __dadd_rn(res, a[i]);
    return res;
}

static __device__ double np_pairwise_sum(const double *a, int n)
{
    if (n <= 128) return pw_block(a, n);
    // explicit post-order walk of the halving tree (n2 = n/2 rounded down to a multiple of 8)
    int off[20], len[20], phase[20];
    double left[20];
    int sp = 0;
    off[0] = 0; len[0] = n; phase[0] = 0; left[0] = 0.0;
    double ret = 0.0;
    while (sp >= 0) {
        if (len[sp] <= 128) {
            ret = pw_block(a + off[sp], len[sp]);
            sp--;
            continue;
        }
        int n2 = len[sp] / 2;
        n2 -= n2 % 8;
        if (phase[sp] == 0) {
            phase[sp] = 1;
            off[sp + 1] = off[sp]; len[sp + 1] = n2; phase[sp + 1] = 0;
            sp++;
        } else if (phase[sp] == 1) {
            left[sp] = ret;
            phase[sp] = 2;
            off[sp + 1] = off[sp] + n2; len[sp + 1] = len[sp] - n2; phase[sp + 1] = 0;
            sp++;
        } else {
            ret = __dadd_rn(left[sp], ret);
            sp--;
        }
    }
    return ret;
}

__device__ __forceinline__ double loglike_row(int kind, const double *p, int d, double *t,
                                              const double *lp)
{
    if (kind == UNB_LOGLIKE_GAUSS) {
        const double sigma = __ldg(lp + d), norm_const = __ldg(lp + d + 1);
        for (int i = 0; i < d; i++) {
            double z = __ddiv_rn(__dsub_rn(p[i], __ldg(lp + i)), sigma);
            t[i] = __dmul_rn(z, z);
        }
        return __dsub_rn(__dmul_rn(-0.5, np_pairwise_sum(t, d)), norm_const);
    } else if (kind == UNB_LOGLIKE_ROSENBROCK) {
        for (int i = 0; i + 1 < d; i++) {
            const double a = p[i], b = p[i + 1];
            const double u = __dsub_rn(b, __dmul_rn(a, a));
            const double v = __dsub_rn(1.0, a);
            t[i] = __dadd_rn(__dmul_rn(100.0, __dmul_rn(u, u)), __dmul_rn(v, v));
        }
        return __dmul_rn(-2.0, np_pairwise_sum(t, d - 1));
    }
    double chi = 1.0;   // eggbox
    for (int i = 0; i < d; i++) chi = __dmul_rn(chi, cos(__ddiv_rn(p[i], 2.0)));
    return pow(__dadd_rn(2.0, chi), 5.0);
}

