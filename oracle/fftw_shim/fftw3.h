/*
 * TEST INFRASTRUCTURE ONLY -- stand-in for <fftw3.h>.
 *
 * The reference links FFTW 3.3.x (src/CMakeLists.txt:15-17,29), which is not
 * vendored in /root/reference and not installed in this image.  This header
 * declares exactly the ten symbols the reference calls
 * (src/fft_processing.c:20,21,24,25,34,47,53,54,56,58,59) so that the
 * UNMODIFIED reference sources compile; fftw_shim.c implements them with an
 * accurate double-precision mixed-radix / Bluestein DFT.  The transform is the
 * textbook unnormalised forward DFT, half spectrum (n1/2+1 columns), row major.
 *
 * Nothing in the product path includes this file.
 */
#ifndef PHD_ORACLE_FFTW3_SHIM_H
#define PHD_ORACLE_FFTW3_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct phd_shim_plan_s* fftw_plan;

#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

int fftw_init_threads(void);
void fftw_plan_with_nthreads(int nthreads);
fftw_complex* fftw_alloc_complex(unsigned long n);
double* fftw_alloc_real(unsigned long n);
fftw_plan fftw_plan_dft_r2c_2d(int n0, int n1, double* in, fftw_complex* out, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_free(void* p);
void fftw_cleanup_threads(void);
void fftw_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
