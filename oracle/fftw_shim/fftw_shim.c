/*
 * TEST INFRASTRUCTURE ONLY -- the ten FFTW entry points the reference calls
 * (src/fft_processing.c:20-59), implemented on oracle/dfft.c so that the
 * unmodified reference sources can be linked in an image without FFTW.
 */
#include "fftw3.h"

#include <stdlib.h>

#include "../dfft.h"

struct phd_shim_plan_s {
    int n0, n1;
    double* in;
    fftw_complex* out;
    int nthreads;
};

static int g_nthreads = 1;

int fftw_init_threads(void) { return 1; }
void fftw_plan_with_nthreads(int nthreads) { g_nthreads = nthreads > 0 ? nthreads : 1; }
fftw_complex* fftw_alloc_complex(unsigned long n) { return (fftw_complex*)malloc(sizeof(fftw_complex) * n); }
double* fftw_alloc_real(unsigned long n) { return (double*)malloc(sizeof(double) * n); }

fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double* in, fftw_complex* out, unsigned flags) {
    (void)flags;
    fftw_plan p = (fftw_plan)malloc(sizeof(struct phd_shim_plan_s));
    if (!p) return NULL;
    p->n0 = n0;
    p->n1 = n1;
    p->in = in;
    p->out = out;
    p->nthreads = g_nthreads;
    return p;
}

void fftw_execute(const fftw_plan p) { dfft_r2c_2d(p->n0, p->n1, p->in, (double*)p->out, p->nthreads); }
void fftw_destroy_plan(fftw_plan p) { free(p); }
void fftw_free(void* p) { free(p); }
void fftw_cleanup_threads(void) {}
void fftw_cleanup(void) {}
