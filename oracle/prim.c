/* prim.c -- plain C restatement of the third-party arithmetic primitives the
 * reference's hot path delegates to.  TEST INFRASTRUCTURE ONLY (see
 * oracle/__init__.py): it pins the numpy/scipy routines the oracle calls, so
 * parity does not rest on those libraries alone.
 *
 * The reference (pure Python) calls, per chunk:
 *   scipy.signal.sosfilt   core/numerical.py:334,399,402,410  (scipy 1.18.1, _sosfilt.pyx: DF2T)
 *   scipy.signal.lfilter   core/numerical.py:445,508,511,519  (_sigtools linear_filter: DF2T)
 *   scipy.signal.upfirdn   via resample_poly, numerical.py:610,631 (_upfirdn_apply.pyx)
 *   numpy.fft.rfft         numerical.py:214,235,699 (pocketfft); here a direct O(n^2) DFT
 *   numpy.convolve         what oaconvolve's result is defined to equal
 * Each function below is the published algorithm of that routine, written
 * from its definition.  Build: make -C oracle  ->  oracle/_build/liboracle_prim.so
 */
#include <math.h>
#include <stddef.h>

/* y[i] = sum_k h[k] x[i-k], full convolution, length n + m - 1 (numpy.convolve 'full') */
void prim_convolve_full(const double *x, long n, const double *h, long m, double *y) {
    for (long i = 0; i < n + m - 1; ++i) {
        long double acc = 0.0L;
        long k0 = i - (n - 1) > 0 ? i - (n - 1) : 0;
        long k1 = i < m - 1 ? i : m - 1;
        for (long k = k0; k <= k1; ++k) acc += (long double)h[k] * (long double)x[i - k];
        y[i] = (double)acc;
    }
}

/* Cascaded biquads, direct form II transposed, state zi[nsec][2] updated in
 * place (scipy.signal.sosfilt semantics for one 1-D signal). */
void prim_sosfilt(const double *sos, long nsec, const double *x, long n, double *zi, double *y) {
    for (long i = 0; i < n; ++i) {
        double v = x[i];
        for (long s = 0; s < nsec; ++s) {
            const double *c = sos + 6 * s;
            double b0 = c[0] / c[3], b1 = c[1] / c[3], b2 = c[2] / c[3];
            double a1 = c[4] / c[3], a2 = c[5] / c[3];
            double out = b0 * v + zi[2 * s];
            zi[2 * s] = b1 * v - a1 * out + zi[2 * s + 1];
            zi[2 * s + 1] = b2 * v - a2 * out;
            v = out;
        }
        y[i] = v;
    }
}

/* Transfer-function filter, direct form II transposed, order = max(nb, na) - 1,
 * state z[order] updated in place (scipy.signal.lfilter semantics). */
void prim_lfilter(const double *b, long nb, const double *a, long na, const double *x, long n,
                  double *z, double *y) {
    long order = (nb > na ? nb : na) - 1;
    for (long i = 0; i < n; ++i) {
        double b0 = nb > 0 ? b[0] / a[0] : 0.0;
        double out = order > 0 ? z[0] + b0 * x[i] : b0 * x[i];
        for (long k = 1; k <= order; ++k) {
            double bk = k < nb ? b[k] / a[0] : 0.0;
            double ak = k < na ? a[k] / a[0] : 0.0;
            double next = k < order ? z[k] : 0.0;
            z[k - 1] = next + bk * x[i] - ak * out;
        }
        y[i] = out;
    }
}

/* upfirdn: upsample by `up` (zero insertion), FIR filter h, keep every `down`-th
 * sample.  y has ceil(((n - 1) * up + nh) / down) samples. */
long prim_upfirdn_len(long n, long nh, long up, long down) {
    long full = (n - 1) * up + nh;
    return (full + down - 1) / down;
}
void prim_upfirdn(const double *h, long nh, const double *x, long n, long up, long down,
                  double *y) {
    long ny = prim_upfirdn_len(n, nh, up, down);
    for (long j = 0; j < ny; ++j) {
        long t = j * down;                 /* index in the upsampled, filtered stream */
        long double acc = 0.0L;
        /* contributing inputs k: 0 <= t - k*up < nh */
        long kmax = t / up;
        if (kmax > n - 1) kmax = n - 1;
        for (long k = kmax; k >= 0; --k) {
            long hi = t - k * up;
            if (hi >= nh) break;
            acc += (long double)h[hi] * (long double)x[k];
        }
        y[j] = (double)acc;
    }
}

/* Real-input DFT by definition: X[k] = sum_j x[j] exp(-2 pi i j k / n), k <= n/2. */
void prim_rdft(const double *x, long n, double *re, double *im) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (long k = 0; k <= n / 2; ++k) {
        long double sr = 0.0L, si = 0.0L;
        for (long j = 0; j < n; ++j) {
            long m = (long)(((long long)j * k) % n);
            long double ang = two_pi * (long double)m / (long double)n;
            sr += (long double)x[j] * cosl(ang);
            si -= (long double)x[j] * sinl(ang);
        }
        re[k] = (double)sr;
        im[k] = (double)si;
    }
}
