// Internal declarations shared by the CUDA translation units of libreport_data.so.
// Nothing here is part of the C ABI (include/photohive_dsp.h).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "photohive_dsp.h"

#define PHD_CELL_Q 5          // u64 quantities per palette cell in global memory: count, n255, sum max, sum s, sum hue fraction
#define PHD_MAX_GROUPS 2048   // T = h*s*v + v + 1 upper bound (shared-memory tables)
#define PHD_MAX_BINS 8192     // na*nr upper bound
#define PHD_MAX_FACTORS 24
#define PHD_NCS_SMALL 832     // compile-time chunk-array stride of the 256-thread front end (cells + twins + 32 scratch <= 832)

// Fixed-point scales of the integer accumulators (all sums are exact integers => order independent).
#define PHD_S_SHIFT 20        // saturation and hue fraction of the palette cells: value * 2^20
#define PHD_T_SHIFT 22        // wrapped hue sums per parent: degrees * 2^22
#define PHD_LN_SHIFT 20       // ln(power): value * 2^20 (value < 2^6)

// Opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the function IN THE CURRENT
// DEVICE'S context, so it is set once per device (a process may hold contexts on several GPUs).
#define PHD_ALLOW_SMEM(func, bytes)                                                                           \
    do {                                                                                                      \
        static bool done_[64];                                                                                \
        int d_ = 0;                                                                                           \
        cudaGetDevice(&d_);                                                                                   \
        if ((unsigned)d_ < 64u && !done_[d_]) {                                                               \
            cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));            \
            done_[d_] = true;                                                                                 \
        }                                                                                                     \
    } while (0)

typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;

// Everything a kernel needs to know about the job; passed by value.
struct DevParams {
    int W, H, fw;        // full image, fw = W/2+1 spectrum columns
    int Hp;              // column pitch of the transposed spectrum / bin map: H rounded up to a multiple of 4
    int dw, dh, ds;      // HSV (possibly downsampled) image and the rate
    long long npx, hpx;  // W*H and dw*dh
    int fe_threads;      // front-end CTA size (256 or 512); a chunk is fe_threads * 16 HSV pixels
    int chunk;
    int nchunks;         // ceil(hpx / chunk)
    int cpp, nspans;     // front-end walk: chunks per walk (span) and spans per image, set per launch (phd_fe_plan)
    int hp, sp, vp, T;
    int ncls, NC;        // palette classes sp*vp+2 and cells ncls*hp*4 (pixel_cells.cuh)
    double Lh, Ls, Lv, bt, gt;
    double coverage;
    int L;               // linked_list_size
    float qw, svw;
    int nr, na, nbins;
    double streak, magthr;
    int denom;
    int max_boxes;
    size_t image_stride;
    int aligned16;       // image base and stride are 16-byte aligned
};

// Per-group merge plan written by the palette-select kernel, read by palette-accumulate.
struct GroupPlan {
    short slot;   // parent slot this group's pixels are summed into (-1: none)
    short mode;   // 0 none, 1 every pixel, 2 tie (partial, see below)
    int cstar;    // tie: chunks < cstar fully accepted, chunk == cstar accepts the first `need` pixels
    int need;
    int clast;    // tie: chunk holding the group's last pixel when that pixel alone survives, else -1
};

// Per-parent integer accumulators (palette-accumulate -> finalize).
struct SlotAcc {
    u64 cnt, summax, n255, s_sum, t_sum;  // s_sum * 2^PHD_S_SHIFT, t_sum * 2^PHD_T_SHIFT
};

// Per-image scalar accumulators of the front end.
struct ImageAcc {
    u64 sum[3];    // sum of k per channel
    u64 sumsq[3];  // sum of k^2 per channel
    u64 s_sum;     // sum of s * 2^PHD_S_SHIFT over the HSV image
    u64 dc_valid;  // general-input route (f64path.cu): `dc` holds X[0,0] of the reference's transform
    double dc;
    u64 pad;
};

// Accumulators of the general-input (planes of arbitrary doubles) route, one image per call (f64path.cu).
struct F64Work {
    double* acc;     // [10]  sum r,g,b; sum of squares r,g,b; sum gray; sum of saturations; (Br+Bg+Bb)/3; unused
    double* slots;   // [T][4]  per parent slot: sum v, sum s, sum wrapped hue
    double* sharp;   // [max_boxes][2]  sum f, sum f^2
    float* gray32;   // [npx]  (gray - average) * 255000, the row transform's input
};

struct SharpAcc {
    long long s1;  // sum f        (f = integer Laplacian of the gray numerator)
    u64 s2lo, s2hi; // sum f^2 split as 32-bit halves
    u64 pad;
};

struct FftPlan {
    int n;
    int m;      // 0: the radix plan below is that of length n.  > 0: Bluestein -- length n has a large prime factor, the
                // transform is a circular convolution through length m >= 2n-1 (a product of small radices), and the
                // radix plan and pass tables below are those of length m
    int nfac;
    int fac[PHD_MAX_FACTORS];
    int twp_off[PHD_MAX_FACTORS];  // start of pass f's table inside twp
    const float2* chirp;  // Bluestein: n entries exp(-i pi k^2 / n)
    const float2* bhat;   // Bluestein: m entries, the length-m transform of the conjugate chirp, divided by m
    const float2* tw;   // device, n entries exp(-2 pi i k / n)
    const float2* twp;  // device, per-pass tables: pass f (radix r, m = n/r butterflies) reads
                        // twp[twp_off[f] + (j-1)*m + b], j = 1..r-1 -- consecutive lanes, consecutive entries
};

// A transform too long for shared memory (image sides beyond 12,800 / ~11,000 pixels; the reference admits 24,494):
// four-step through HBM, length L = n1 * n2, or Bluestein through such an L when n has no usable factorisation (fft.cu).
struct PhdLongFft {
    int n;        // logical length
    int L;        // transformed length: n, or the Bluestein length >= 2n-1
    int n1, n2;   // L = n1 * n2
    int blue;     // 1: Bluestein
    FftPlan p1, p2;        // runtime-radix plans of n1 and n2
    float2* mem;           // one allocation holding every table below
    const float2* twL;     // L entries exp(-2 pi i k / L)
    const float2* chirp;   // n entries (Bluestein)
    const float2* bhat;    // L entries (Bluestein)
};
int phd_long_fft_create(int n, PhdLongFft* lf, cudaStream_t st);  // 0 ok
void phd_long_fft_destroy(PhdLongFft* lf);
// rows of ONE image (packed bytes, or the float gray plane of the general-input route) -> transposed half spectrum
int phd_launch_long_rows(const uint8_t* rgb, const float* gray32, const DevParams& P, const PhdLongFft& lf, float2* buf0,
                         float2* buf1, float2* specT, cudaStream_t st, int* launches);
// columns of ONE image + blur-profile epilogue (ws accumulators of that image), or the power spectrum (test hook)
struct Workspace;
int phd_launch_long_cols(const DevParams& P, const PhdLongFft& lf, const float2* specT, float2* buf0, float2* buf1,
                         const u16* binmapT, Workspace& ws, float* power_out, cudaStream_t st, int* launches);

// Device workspace for one sub-batch.
struct Workspace {
    int capacity;  // images
    u16* counts_chunk;   // [cap][nchunks][T]
    u64* cells;          // [cap][PHD_CELL_Q][NC]   zeroed per sub-batch, filled by the front end
    u64* cells_tie;      // [cap][PHD_CELL_Q][NC]   cells of the partly accepted tie groups (zeroed by palette_select)
    u32* span32;         // [spans of the launch][3][NC]  per front-end walk: count, n255, sum max
    u64* span64;         // [spans of the launch][2][NC]  per front-end walk: sum s, sum hue fraction (both * 2^20)
    u32* work;           // [cap * nchunks]         (image, chunk) items of the tie path
    u32* work_n;         // [1]
    u32* queue;          // [1] task counter of the fused front-end + row kernel
    u32* hist;           // [cap][T]
    ImageAcc* iacc;      // [cap]
    GroupPlan* plan;     // [cap][T]
    int* pal_n;          // [cap]
    int* parent_ids;     // [cap][T]
    int* tie_list;       // [cap][T]
    int* tie_n;          // [cap]
    int* tie_groups;     // [cap]
    long long* dropped;  // [cap]
    SlotAcc* sacc;       // [cap][T]
    float2* spec;        // [fft cap][fw][Hp]  row-transformed half spectrum, transposed
    u64* binsum;         // [cap][nbins]
    u32* maxpow;         // [cap] float bits
    SharpAcc* sharp;     // [cap][max_boxes]
    int* boxes;          // [cap][max_boxes][4]
};

// Front-end walk plan: a CTA (or a front-end task of the fused kernel) walks `cpp` consecutive chunks of one image --
// long walks amortise the table load and the final flush, but there must be enough walks to fill 148 SMs.
static inline void phd_fe_plan(DevParams& P, int nimg) {
    const long long total = (long long)P.nchunks * nimg;
    int cpp = (int)(total / (148 * 12));
    cpp = cpp < 1 ? 1 : (cpp > 32 ? 32 : cpp);
    P.cpp = cpp;
    P.nspans = (P.nchunks + cpp - 1) / cpp;
}
// Largest number of walks any launch of up to `cap` images can have (sizes span32 / span64).
static inline size_t phd_fe_max_spans(const DevParams& P0, int cap) {
    DevParams P = P0;
    size_t m = 0;
    for (int n = 1; n <= cap; n++) {
        phd_fe_plan(P, n);
        const size_t t = (size_t)P.nspans * n;
        if (t > m) m = t;
    }
    return m;
}

// ---- launchers (each in its own .cu) --------------------------------------------------------
void phd_launch_pixels(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                       const unsigned char* exc, Workspace& ws, cudaStream_t st, int* launches);
// from_hist: the group totals are already in ws.hist (general-input route); the cells are not consulted
void phd_launch_palette_select(const DevParams& P, int nimg, const double* centres, const float* sv_f, Workspace& ws,
                               cudaStream_t st, int* launches, bool from_hist = false);
void phd_launch_palette_ties(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                             const unsigned char* exc, Workspace& ws, cudaStream_t st, int* launches);
void phd_launch_group_sweep(const DevParams& P, const unsigned char* tabs, const unsigned char* exc, bool fast,
                            u16* out_dev, cudaStream_t st);
size_t phd_exc_table_bytes(int hp);  // exceptional-colour table of h_partitions = hp (pixel_cells.cuh)
void phd_launch_build_cell_tables(const DevParams& P, unsigned char* tables_dev, unsigned char* exc_dev, int* ok_dev,
                                  cudaStream_t st);
size_t phd_cell_tables_size();
size_t phd_pixels_smem(const DevParams& P);

int phd_fft_plan_factors(int n, int* fac, int* nfac);  // 0 ok, nonzero unsupported
int phd_fft_make_plan(int n, FftPlan* pl);             // radix plan of n, or a Bluestein plan (pl->m > 0); 0 ok
void phd_fft_fill_bluestein(float2* chirp_dev, float2* bhat_dev, int n, int m, cudaStream_t st);
void phd_fill_twiddles(float2* dev_tw, int n, cudaStream_t st);
size_t phd_fft_pass_table_entries(const FftPlan& pl);
void phd_fft_fill_pass_tables(float2* dev, FftPlan& pl, cudaStream_t st);  // also sets pl.twp_off / pl.twp
bool phd_launch_front_rows(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                           const unsigned char* exc, const FftPlan& row, Workspace& ws, cudaStream_t st, int* launches);
int phd_launch_fft_rows(const uint8_t* rgb, const DevParams& P, int nimg, const FftPlan& row, float2* spec,
                        cudaStream_t st, int* launches);
int phd_launch_fft_cols_blur(const DevParams& P, int nimg, const FftPlan& col, float2* spec, const u16* binmap,
                             Workspace& ws, float* power_out, cudaStream_t st, int* launches);
void phd_launch_bin_map(int W, int H, int Hp, int nr, int na, u16* map_dev, int* counts_dev, cudaStream_t st);

void phd_launch_sharpness(const uint8_t* rgb, const DevParams& P, int nimg, int max_w, int max_h, Workspace& ws,
                          cudaStream_t st, int* launches);
// f64: accumulators of the general-input route (one image) instead of the integer ones, or nullptr
void phd_launch_finalize(const DevParams& P, int nimg, const double* centres, const int* bincount, Workspace& ws,
                         const phd_flat_layout& lay, unsigned char* records_dev, cudaStream_t st, int* launches,
                         const F64Work* f64 = nullptr);

size_t phd_fft_cols_smem(const DevParams& P, const FftPlan* col, int* tile_cols);

// general-input route (f64path.cu)
void phd_launch_f64_front(const double* planes, const DevParams& P, F64Work& fw, Workspace& ws, cudaStream_t st, int* launches);
void phd_launch_f64_accumulate(const double* planes, const DevParams& P, const double* centres, F64Work& fw, Workspace& ws,
                               cudaStream_t st, int* launches);
size_t phd_f64_accumulate_smem(const DevParams& P);
int phd_launch_fft_rows_gray(const float* gray32, const DevParams& P, const FftPlan& row, float2* specT, cudaStream_t st,
                             int* launches);
