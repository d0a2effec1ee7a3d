/*
 * TEST INFRASTRUCTURE ONLY -- see dfft.h.
 *
 * Stockham autosort mixed-radix DFT in IEEE double (radix 4/2/3/5 butterflies,
 * generic O(r^2) butterfly for other small primes, Bluestein chirp-z for
 * lengths with a prime factor > 31).  Twiddles come from long-double sincos so
 * each table entry is correctly rounded to within one ulp.
 */
#include "dfft.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define DFFT_MAX_GENERIC_RADIX 31

struct dfft_plan_s {
    int n;
    int nfac;
    int fac[64];
    double* tw; /* n complex: exp(-2 pi i k / n) */
    /* Bluestein members (bluestein != 0) */
    int bluestein;
    int m;
    dfft_plan* sub;
    double* chirp; /* n complex: exp(-pi i k^2 / n) */
    double* bk;    /* m complex: DFT of the wrapped conjugate chirp */
};

static void fill_twiddles(double* tw, int n) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    for (int k = 0; k < n; k++) {
        long double a = two_pi * (long double)k / (long double)n;
        tw[2 * k] = (double)cosl(a);
        tw[2 * k + 1] = (double)(-sinl(a));
    }
}

static int largest_prime_factor(int n) {
    int best = 1;
    for (int p = 2; (long)p * p <= n; p++) {
        while (n % p == 0) {
            best = p;
            n /= p;
        }
    }
    if (n > 1) best = n;
    return best;
}

dfft_plan* dfft_plan_create(int n) {
    dfft_plan* p = (dfft_plan*)calloc(1, sizeof(dfft_plan));
    p->n = n;
    if (n > 1 && largest_prime_factor(n) > DFFT_MAX_GENERIC_RADIX) {
        p->bluestein = 1;
        int m = 1;
        while (m < 2 * n - 1) m <<= 1;
        p->m = m;
        p->sub = dfft_plan_create(m);
        p->chirp = (double*)malloc(sizeof(double) * 2 * n);
        const long double pi = 3.141592653589793238462643383279502884L;
        for (int k = 0; k < n; k++) {
            long long k2 = ((long long)k * k) % (2LL * n);
            long double a = pi * (long double)k2 / (long double)n;
            p->chirp[2 * k] = (double)cosl(a);
            p->chirp[2 * k + 1] = (double)(-sinl(a));
        }
        p->bk = (double*)calloc(2 * (size_t)m, sizeof(double));
        for (int k = 0; k < n; k++) {
            p->bk[2 * k] = p->chirp[2 * k];
            p->bk[2 * k + 1] = -p->chirp[2 * k + 1];
            if (k > 0) {
                p->bk[2 * (m - k)] = p->chirp[2 * k];
                p->bk[2 * (m - k) + 1] = -p->chirp[2 * k + 1];
            }
        }
        double* scratch = (double*)malloc(sizeof(double) * dfft_scratch_len(p->sub));
        dfft_execute(p->sub, p->bk, scratch);
        free(scratch);
        return p;
    }
    p->tw = (double*)malloc(sizeof(double) * 2 * (size_t)(n > 0 ? n : 1));
    fill_twiddles(p->tw, n);
    int rem = n;
    while (rem % 4 == 0) { p->fac[p->nfac++] = 4; rem /= 4; }
    while (rem % 2 == 0) { p->fac[p->nfac++] = 2; rem /= 2; }
    for (int q = 3; rem > 1; q += 2) {
        while (rem % q == 0) { p->fac[p->nfac++] = q; rem /= q; }
    }
    return p;
}

void dfft_plan_destroy(dfft_plan* p) {
    if (!p) return;
    if (p->sub) dfft_plan_destroy(p->sub);
    free(p->tw);
    free(p->chirp);
    free(p->bk);
    free(p);
}

long dfft_scratch_len(const dfft_plan* p) {
    if (p->bluestein) return 2L * p->m + dfft_scratch_len(p->sub);
    return 2L * p->n;
}

static inline void cmul(double ar, double ai, double br, double bi, double* cr, double* ci) {
    *cr = ar * br - ai * bi;
    *ci = ar * bi + ai * br;
}

static void stage_generic(const dfft_plan* pl, int r, int m, int s, const double* a, double* b) {
    const int n = pl->n;
    const double* tw = pl->tw;
    const int step_r = n / r;
    double xr[DFFT_MAX_GENERIC_RADIX], xi[DFFT_MAX_GENERIC_RADIX];
    for (int pp = 0; pp < m; pp++) {
        for (int q = 0; q < s; q++) {
            for (int k = 0; k < r; k++) {
                xr[k] = a[2 * (q + s * (pp + k * m))];
                xi[k] = a[2 * (q + s * (pp + k * m)) + 1];
            }
            for (int j = 0; j < r; j++) {
                double sr = xr[0], si = xi[0];
                for (int k = 1; k < r; k++) {
                    int idx = ((j * k) % r) * step_r;
                    double tr, ti;
                    cmul(xr[k], xi[k], tw[2 * idx], tw[2 * idx + 1], &tr, &ti);
                    sr += tr;
                    si += ti;
                }
                int widx = pp * j * s;
                double orr, oi;
                cmul(sr, si, tw[2 * widx], tw[2 * widx + 1], &orr, &oi);
                b[2 * (q + s * (r * pp + j))] = orr;
                b[2 * (q + s * (r * pp + j)) + 1] = oi;
            }
        }
    }
}

static void stage2(const dfft_plan* pl, int m, int s, const double* a, double* b) {
    const double* tw = pl->tw;
    for (int pp = 0; pp < m; pp++) {
        double wr = tw[2 * (pp * s)], wi = tw[2 * (pp * s) + 1];
        const double* a0 = a + 2 * (s * pp);
        const double* a1 = a + 2 * (s * (pp + m));
        double* b0 = b + 2 * (s * (2 * pp));
        double* b1 = b + 2 * (s * (2 * pp + 1));
        for (int q = 0; q < s; q++) {
            double ur = a0[2 * q], ui = a0[2 * q + 1], vr = a1[2 * q], vi = a1[2 * q + 1];
            b0[2 * q] = ur + vr;
            b0[2 * q + 1] = ui + vi;
            cmul(ur - vr, ui - vi, wr, wi, &b1[2 * q], &b1[2 * q + 1]);
        }
    }
}

static void stage4(const dfft_plan* pl, int m, int s, const double* a, double* b) {
    const double* tw = pl->tw;
    for (int pp = 0; pp < m; pp++) {
        int i1 = pp * s, i2 = 2 * pp * s, i3 = 3 * pp * s;
        double w1r = tw[2 * i1], w1i = tw[2 * i1 + 1];
        double w2r = tw[2 * i2], w2i = tw[2 * i2 + 1];
        double w3r = tw[2 * i3], w3i = tw[2 * i3 + 1];
        const double* a0 = a + 2 * (s * pp);
        const double* a1 = a + 2 * (s * (pp + m));
        const double* a2 = a + 2 * (s * (pp + 2 * m));
        const double* a3 = a + 2 * (s * (pp + 3 * m));
        double* b0 = b + 2 * (s * (4 * pp));
        double* b1 = b0 + 2 * s;
        double* b2 = b1 + 2 * s;
        double* b3 = b2 + 2 * s;
        for (int q = 0; q < s; q++) {
            double x0r = a0[2 * q], x0i = a0[2 * q + 1];
            double x1r = a1[2 * q], x1i = a1[2 * q + 1];
            double x2r = a2[2 * q], x2i = a2[2 * q + 1];
            double x3r = a3[2 * q], x3i = a3[2 * q + 1];
            double t0r = x0r + x2r, t0i = x0i + x2i;
            double t1r = x0r - x2r, t1i = x0i - x2i;
            double t2r = x1r + x3r, t2i = x1i + x3i;
            /* (-i) * (x1 - x3) */
            double t3r = x1i - x3i, t3i = -(x1r - x3r);
            b0[2 * q] = t0r + t2r;
            b0[2 * q + 1] = t0i + t2i;
            cmul(t1r + t3r, t1i + t3i, w1r, w1i, &b1[2 * q], &b1[2 * q + 1]);
            cmul(t0r - t2r, t0i - t2i, w2r, w2i, &b2[2 * q], &b2[2 * q + 1]);
            cmul(t1r - t3r, t1i - t3i, w3r, w3i, &b3[2 * q], &b3[2 * q + 1]);
        }
    }
}

static void stage3(const dfft_plan* pl, int m, int s, const double* a, double* b) {
    const double* tw = pl->tw;
    const double k3 = 0.86602540378443864676372317075294; /* sin(2 pi / 3) */
    for (int pp = 0; pp < m; pp++) {
        int i1 = pp * s, i2 = 2 * pp * s;
        double w1r = tw[2 * i1], w1i = tw[2 * i1 + 1];
        double w2r = tw[2 * i2], w2i = tw[2 * i2 + 1];
        const double* a0 = a + 2 * (s * pp);
        const double* a1 = a + 2 * (s * (pp + m));
        const double* a2 = a + 2 * (s * (pp + 2 * m));
        double* b0 = b + 2 * (s * (3 * pp));
        double* b1 = b0 + 2 * s;
        double* b2 = b1 + 2 * s;
        for (int q = 0; q < s; q++) {
            double x0r = a0[2 * q], x0i = a0[2 * q + 1];
            double x1r = a1[2 * q], x1i = a1[2 * q + 1];
            double x2r = a2[2 * q], x2i = a2[2 * q + 1];
            double tr = x1r + x2r, ti = x1i + x2i;
            double ur = x0r - 0.5 * tr, ui = x0i - 0.5 * ti;
            /* (-i k3) * (x1 - x2) */
            double vr = k3 * (x1i - x2i), vi = -k3 * (x1r - x2r);
            b0[2 * q] = x0r + tr;
            b0[2 * q + 1] = x0i + ti;
            cmul(ur + vr, ui + vi, w1r, w1i, &b1[2 * q], &b1[2 * q + 1]);
            cmul(ur - vr, ui - vi, w2r, w2i, &b2[2 * q], &b2[2 * q + 1]);
        }
    }
}

static void stage5(const dfft_plan* pl, int m, int s, const double* a, double* b) {
    const double* tw = pl->tw;
    const double c1 = 0.30901699437494742410229341718282;  /* cos(2 pi/5) */
    const double c2 = -0.80901699437494742410229341718282; /* cos(4 pi/5) */
    const double s1 = 0.95105651629515357211643933337938;  /* sin(2 pi/5) */
    const double s2 = 0.58778525229247312916870595463907;  /* sin(4 pi/5) */
    for (int pp = 0; pp < m; pp++) {
        double wr[5], wi[5];
        for (int j = 1; j < 5; j++) {
            wr[j] = tw[2 * (pp * j * s)];
            wi[j] = tw[2 * (pp * j * s) + 1];
        }
        const double* a0 = a + 2 * (s * pp);
        const double* a1 = a + 2 * (s * (pp + m));
        const double* a2 = a + 2 * (s * (pp + 2 * m));
        const double* a3 = a + 2 * (s * (pp + 3 * m));
        const double* a4 = a + 2 * (s * (pp + 4 * m));
        double* b0 = b + 2 * (s * (5 * pp));
        double* b1 = b0 + 2 * s;
        double* b2 = b1 + 2 * s;
        double* b3 = b2 + 2 * s;
        double* b4 = b3 + 2 * s;
        for (int q = 0; q < s; q++) {
            double x0r = a0[2 * q], x0i = a0[2 * q + 1];
            double t1r = a1[2 * q] + a4[2 * q], t1i = a1[2 * q + 1] + a4[2 * q + 1];
            double t2r = a2[2 * q] + a3[2 * q], t2i = a2[2 * q + 1] + a3[2 * q + 1];
            double t3r = a1[2 * q] - a4[2 * q], t3i = a1[2 * q + 1] - a4[2 * q + 1];
            double t4r = a2[2 * q] - a3[2 * q], t4i = a2[2 * q + 1] - a3[2 * q + 1];
            double m1r = x0r + c1 * t1r + c2 * t2r, m1i = x0i + c1 * t1i + c2 * t2i;
            double m2r = x0r + c2 * t1r + c1 * t2r, m2i = x0i + c2 * t1i + c1 * t2i;
            double n1r = s1 * t3r + s2 * t4r, n1i = s1 * t3i + s2 * t4i;
            double n2r = s2 * t3r - s1 * t4r, n2i = s2 * t3i - s1 * t4i;
            /* (-i) * n = (n_i, -n_r) */
            b0[2 * q] = x0r + t1r + t2r;
            b0[2 * q + 1] = x0i + t1i + t2i;
            cmul(m1r + n1i, m1i - n1r, wr[1], wi[1], &b1[2 * q], &b1[2 * q + 1]);
            cmul(m2r + n2i, m2i - n2r, wr[2], wi[2], &b2[2 * q], &b2[2 * q + 1]);
            cmul(m2r - n2i, m2i + n2r, wr[3], wi[3], &b3[2 * q], &b3[2 * q + 1]);
            cmul(m1r - n1i, m1i + n1r, wr[4], wi[4], &b4[2 * q], &b4[2 * q + 1]);
        }
    }
}

static void stockham(const dfft_plan* pl, double* x, double* y) {
    int len = pl->n, s = 1;
    double *a = x, *b = y;
    for (int f = 0; f < pl->nfac; f++) {
        int r = pl->fac[f];
        int m = len / r;
        switch (r) {
            case 2: stage2(pl, m, s, a, b); break;
            case 3: stage3(pl, m, s, a, b); break;
            case 4: stage4(pl, m, s, a, b); break;
            case 5: stage5(pl, m, s, a, b); break;
            default: stage_generic(pl, r, m, s, a, b); break;
        }
        len = m;
        s *= r;
        double* t = a;
        a = b;
        b = t;
    }
    if (a != x) memcpy(x, a, sizeof(double) * 2 * (size_t)pl->n);
}

void dfft_execute(const dfft_plan* p, double* data, double* scratch) {
    if (p->n <= 1) return;
    if (!p->bluestein) {
        stockham(p, data, scratch);
        return;
    }
    const int n = p->n, m = p->m;
    double* a = scratch;
    double* sub_scratch = scratch + 2 * (size_t)m;
    for (int k = 0; k < n; k++) cmul(data[2 * k], data[2 * k + 1], p->chirp[2 * k], p->chirp[2 * k + 1], &a[2 * k], &a[2 * k + 1]);
    memset(a + 2 * (size_t)n, 0, sizeof(double) * 2 * (size_t)(m - n));
    dfft_execute(p->sub, a, sub_scratch);
    for (int k = 0; k < m; k++) {
        double r, i;
        cmul(a[2 * k], a[2 * k + 1], p->bk[2 * k], p->bk[2 * k + 1], &r, &i);
        a[2 * k] = r;
        a[2 * k + 1] = -i; /* conjugate: inverse transform through the forward plan */
    }
    dfft_execute(p->sub, a, sub_scratch);
    const double inv = 1.0 / (double)m;
    for (int k = 0; k < n; k++) {
        double r = a[2 * k] * inv, i = -a[2 * k + 1] * inv;
        cmul(r, i, p->chirp[2 * k], p->chirp[2 * k + 1], &data[2 * k], &data[2 * k + 1]);
    }
}

void dfft_r2c_2d(int n0, int n1, const double* in, double* out, int nthreads) {
    const int fw = n1 / 2 + 1;
    dfft_plan* prow = dfft_plan_create(n1);
    dfft_plan* pcol = dfft_plan_create(n0);
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
    if (nthreads > omp_get_max_threads()) nthreads = omp_get_max_threads(); /* honour OMP_NUM_THREADS */
#endif
    const int npairs = (n0 + 1) / 2;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
#endif
    {
        double* buf = (double*)malloc(sizeof(double) * 2 * (size_t)(n1 > n0 ? n1 : n0));
        long sl_r = dfft_scratch_len(prow), sl_c = dfft_scratch_len(pcol);
        double* scratch = (double*)malloc(sizeof(double) * (size_t)(sl_r > sl_c ? sl_r : sl_c));
        /* rows, two real rows per complex transform */
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int pr = 0; pr < npairs; pr++) {
            const int ra = 2 * pr, rb = 2 * pr + 1;
            const double* xa = in + (size_t)ra * n1;
            const double* xb = rb < n0 ? in + (size_t)rb * n1 : NULL;
            for (int i = 0; i < n1; i++) {
                buf[2 * i] = xa[i];
                buf[2 * i + 1] = xb ? xb[i] : 0.0;
            }
            dfft_execute(prow, buf, scratch);
            double* oa = out + 2 * (size_t)ra * fw;
            double* ob = xb ? out + 2 * (size_t)rb * fw : NULL;
            for (int k = 0; k < fw; k++) {
                int kc = (n1 - k) % n1;
                double zr = buf[2 * k], zi = buf[2 * k + 1];
                double cr = buf[2 * kc], ci = -buf[2 * kc + 1];
                oa[2 * k] = 0.5 * (zr + cr);
                oa[2 * k + 1] = 0.5 * (zi + ci);
                if (ob) {
                    /* (z - conj(zc)) / (2i) */
                    ob[2 * k] = 0.5 * (zi - ci);
                    ob[2 * k + 1] = -0.5 * (zr - cr);
                }
            }
        }
        /* columns */
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (int c = 0; c < fw; c++) {
            for (int r = 0; r < n0; r++) {
                buf[2 * r] = out[2 * ((size_t)r * fw + c)];
                buf[2 * r + 1] = out[2 * ((size_t)r * fw + c) + 1];
            }
            dfft_execute(pcol, buf, scratch);
            for (int r = 0; r < n0; r++) {
                out[2 * ((size_t)r * fw + c)] = buf[2 * r];
                out[2 * ((size_t)r * fw + c) + 1] = buf[2 * r + 1];
            }
        }
        free(buf);
        free(scratch);
    }
    dfft_plan_destroy(prow);
    dfft_plan_destroy(pcol);
}
