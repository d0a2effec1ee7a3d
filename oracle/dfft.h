/*
 * TEST INFRASTRUCTURE ONLY -- double-precision DFT used by the CPU oracle.
 *
 * Restates the published definition of the transform FFTW's
 * fftw_plan_dft_r2c_2d computes (FFTW 3.3.x manual, "What FFTW Really
 * Computes": unnormalised forward DFT, sign -1, output n0 x (n1/2+1),
 * row major).  FFTW itself is an un-vendored dependency of the reference
 * (src/CMakeLists.txt:15-17,29; version hint fftw-3.3.10 in .gitignore:32).
 * Used by oracle/fftw_shim (so the unmodified reference links) and by
 * oracle/photohive_oracle.c.  Never linked into the product library.
 */
#ifndef PHD_ORACLE_DFFT_H
#define PHD_ORACLE_DFFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dfft_plan_s dfft_plan;

/* 1-D complex plan for any n >= 1 (mixed radix for smooth n, Bluestein otherwise). */
dfft_plan* dfft_plan_create(int n);
void dfft_plan_destroy(dfft_plan* p);
/* In-place forward transform of n interleaved (re,im) doubles; scratch holds dfft_scratch_len(p) doubles. */
void dfft_execute(const dfft_plan* p, double* data, double* scratch);
long dfft_scratch_len(const dfft_plan* p);

/* Real 2-D forward transform: in[n0][n1] -> out[n0][n1/2+1][2], nthreads OpenMP threads (>=1). */
void dfft_r2c_2d(int n0, int n1, const double* in, double* out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
