/*
 * photohive_dsp.h -- C ABI of libreport_data.so (B200-native build).
 *
 * Part 1 is the DROP-IN boundary: the three entry points and every struct the reference's
 * Python ctypes layer (lib.py:20-37, structures.py) and C callers (src/test/test.c:76-84) bind.
 * Names, argument order, struct layout (x86-64 SysV) and ownership rules are those of the
 * reference headers cited on each declaration (paths relative to the reference root).
 *
 * Part 2 is ADDITIVE: a batch interface on packed 8-bit RGB (host or device resident) that
 * returns flat, fixed-stride report records.  It is what the throughput metric is measured on;
 * a record converts to a drop-in Full_Report_Data with phd_flat_to_full_report().
 *
 * Plain pointers and sizes only; no CUDA or torch types cross this boundary.
 */
#ifndef PHOTOHIVE_DSP_H
#define PHOTOHIVE_DSP_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define PHD_API __attribute__((visibility("default")))
#else
#define PHD_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ===================================================================================== */
/* Part 1 -- drop-in boundary                                                             */
/* ===================================================================================== */

typedef double Pixel; /* src/types.h:5 */
typedef double Bin;   /* src/blur_profile.h:8 */

/* src/image_processing.h:12-17 (32 bytes: int, pad, 3 doubles) */
typedef struct Pixel_HSV {
    int parent_id;
    double h;
    double s;
    double v;
} Pixel_HSV;

/* src/image_processing.h:31-36 -- planar, row major, values in [0,1]; height comes first */
typedef struct Image_RGB {
    int height, width;
    Pixel* r;
    Pixel* g;
    Pixel* b;
} Image_RGB;

/* src/image_processing.h:63-66 */
typedef struct Image_PGM {
    int height, width;
    Pixel* data;
} Image_PGM;

/* src/image_processing.h:73-80 */
typedef struct RGB_Statistics {
    Pixel Br, Bg, Bb, Cr, Cg, Cb;
} RGB_Statistics;

/* src/image_processing.h:92-98 -- boxes are [top,bottom) x [left,right) */
typedef struct Crop_Boundaries {
    int N;
    int* top;
    int* bottom;
    int* left;
    int* right;
} Crop_Boundaries;

/* src/color_quantization.h:11-15 */
typedef struct Color_Palette {
    int N;
    Pixel_HSV* averages;
    Pixel* percentages;
} Color_Palette;

/* src/blur_profile.h:20-24 -- bins[angle][radius], one malloc per angle row */
typedef struct Blur_Profile {
    int num_angle_bins, num_radius_bins;
    int angle_bin_size, radius_bin_size;
    Bin** bins;
} Blur_Profile;

/* src/blur_profile.h:52-61 */
typedef struct Blur_Vector {
    int angle;
    float magnitude;
} Blur_Vector;
typedef struct Blur_Vector_Group {
    int len_vectors;
    Blur_Vector* blur_vectors;
} Blur_Vector_Group;

/* src/utilities.h:25-28 */
typedef struct Sharpnesses {
    int N;
    Pixel* sharpness;
} Sharpnesses;

/* src/utilities.h:30-37 */
typedef struct Full_Report_Data {
    RGB_Statistics* rgb_stats;
    Color_Palette* color_palette;
    Blur_Profile* blur_profile;
    Blur_Vector_Group* blur_vectors;
    Pixel average_saturation;
    Sharpnesses* sharpness; /* NULL when no Crop_Boundaries were passed (src/filtering.c:152-154) */
} Full_Report_Data;

/*
 * Replaces src/interface.c:20-94 (declared src/interface.h:16-23).
 * Returns a malloc-owned report, or NULL (+ a message on stderr) when the reference would
 * (src/utilities.c:64-87: NULL image or planes, a side < 350, > 120 M pixels, aspect outside
 * [1/5, 5]) or when the GPU path cannot serve the request (no CUDA device, an image side beyond
 * the shared-memory FFT -- see DESIGN.md "Limits").
 * Planes holding k/255.0 (8-bit images, what utils.py:30-37 produces) take the integer pipeline; any other doubles
 * (16-bit images, create_test_rgb of src/debug.c:53 ...) take a two-pass FP64 route (csrc/f64path.cu).
 * The input image and boxes are not modified and stay owned by the caller.
 */
PHD_API Full_Report_Data* get_full_report_data(Image_RGB* image, Crop_Boundaries* salient_characters,
                                       int h_partitions, int s_partitions, int v_partitions,
                                       double black_thresh, double gray_thresh,
                                       double coverage_thresh, int linked_list_size,
                                       int downsample_rate, int radius_partitions, int angle_partitions,
                                       float quantity_weight, float saturation_value_weight,
                                       double fft_streak_thresh, double magnitude_thresh,
                                       int blur_cutoff_ratio_denom);

/* Replaces src/interface.c:97-111 (declared src/interface.h:26): frees every sub-object, then *report = NULL. */
PHD_API void free_full_report(Full_Report_Data** report);

/* Replaces src/blur_profile.c:140-180 (declared src/blur_profile.h:78); host side, called by
 * core.py:223 with the IMAGE height and width.  Returns a malloc-owned Image_PGM. */
PHD_API Image_PGM* get_blur_profile_visual(Blur_Profile* blur_profile, int height, int width);

/* ===================================================================================== */
/* Part 2 -- additive batch interface                                                     */
/* ===================================================================================== */

/* The 15 scalar arguments of get_full_report_data, in the same order (src/interface.h:16-23). */
typedef struct phd_params {
    int h_partitions, s_partitions, v_partitions;
    double black_thresh, gray_thresh, coverage_thresh;
    int linked_list_size, downsample_rate, radius_partitions, angle_partitions;
    float quantity_weight, saturation_value_weight;
    double fft_streak_thresh, magnitude_thresh;
    int blur_cutoff_ratio_denom;
} phd_params;

/* Defaults of core.py:442-448. */
PHD_API void phd_default_params(phd_params* p);

/* Fixed-size head of one flat report record.  It is followed, at the offsets returned by
 * phd_flat_layout(), by: double palette_hsv[3*T]; double palette_pct[T]; int parent_ids[T];
 * double blur_bins[na*nr]; double sharpness[max_boxes]   (T = h*s*v + v + 1). */
typedef struct phd_flat_head {
    double rgb_stats[6]; /* Br Bg Bb Cr Cg Cb */
    double average_saturation;
    double max_power;      /* max of the raw power spectrum (diagnostic) */
    long long dropped_pixels; /* pixels lost by the reference's tie path (diagnostic) */
    int palette_n;
    int tie_groups;        /* diagnostic */
    int n_sharpness;       /* -1: no boxes were given (Full_Report_Data.sharpness == NULL) */
    int angle_bin_size, radius_bin_size;
    int num_angle_bins, num_radius_bins;
    int status;            /* 0 ok; nonzero: record invalid (PHD_E_*) */
    int blur_vec_angle[10];
    float blur_vec_mag[10];
} phd_flat_head;

typedef struct phd_flat_layout {
    size_t record_bytes; /* stride between records, multiple of 16 */
    size_t off_palette_hsv, off_palette_pct, off_parent_ids, off_blur_bins, off_sharpness;
    int T, na, nr, max_boxes;
} phd_flat_layout;

enum {
    PHD_OK = 0,
    PHD_E_REJECTED = 1,    /* the reference's pre-checks would return NULL */
    PHD_E_BAD_PARAMS = 2,
    PHD_E_NO_DEVICE = 3,
    PHD_E_CUDA = 4,
    PHD_E_UNSUPPORTED = 5, /* e.g. a transform length with a prime factor the FFT does not cover */
    PHD_E_NOT_8BIT = 6     /* no longer produced: non-8-bit planes are served by the FP64 route of get_full_report_data */
};

typedef struct phd_context phd_context; /* one per CUDA device; owns streams, plans, workspaces */

PHD_API int phd_context_create(int device, phd_context** out);
PHD_API void phd_context_destroy(phd_context* ctx);
PHD_API const char* phd_last_error(const phd_context* ctx);

PHD_API int phd_flat_get_layout(const phd_params* p, int max_boxes, phd_flat_layout* out);

/*
 * The hot path on a batch: n_images packed 8-bit RGB images (r,g,b interleaved, row major,
 * `image_stride` bytes apart, all width x height), resident on the host or on ctx's device
 * (detected from the pointer).  `boxes` is NULL or n_images*max_boxes quadruples
 * {top,bottom,left,right} (host memory); every image uses max_boxes boxes.
 * `records` receives n_images records of layout phd_flat_get_layout(p, max_boxes); it may be a
 * host or a device pointer.  Synchronous: returns after the records are written.
 * Boxes must satisfy 0 <= top <= bottom <= height and 0 <= left <= right <= width (crop_pgm,
 * src/image_processing.c:215-219, refuses anything else), at most 65535 per image; an empty box yields NaN as in
 * the reference.  Otherwise PHD_E_BAD_PARAMS.
 * ORDERING CONTRACT for device-resident input / output: the library works on its own non-blocking streams and does
 * not know the caller's.  Whatever produced `rgb` (a kernel or a copy on another stream) must have COMPLETED before
 * the call (synchronise that stream or its event first), and `rgb` / `records` must live on ctx's device.  On return
 * all work of the call has completed, so the caller may reuse or free every buffer at once.
 */
PHD_API int phd_get_reports_u8(phd_context* ctx, const uint8_t* rgb, int n_images, int width, int height,
                       size_t image_stride, const int* boxes, int max_boxes, const phd_params* p,
                       void* records);

/*
 * The same call on the GPUs of one box (images are independent, there is no exchange step
 * and no collective).  ctxs[0..n_ctx-1] are contexts on DISTINCT devices; context g takes the contiguous range
 * [g*n_images/n_ctx, (g+1)*n_images/n_ctx) of the batch on a host thread of its own and its device writes the
 * records of that range straight into `records` -- the gather is the join of those threads.  `rgb` and `records` are
 * HOST buffers (pinned buffers copy at the full PCIe rate; pageable ones go through the threaded staging uploader).
 * Returns the first non-zero status of any range.  Replaces the per-image loop a caller of the reference's
 * get_full_report_data (src/interface.c:20-94) would write over a photo collection.
 */
PHD_API int phd_get_reports_u8_multi(phd_context* const* ctxs, int n_ctx, const uint8_t* rgb, int n_images, int width,
                             int height, size_t image_stride, const int* boxes, int max_boxes, const phd_params* p,
                             void* records);

/* Builds the drop-in, malloc-owned report from one flat record (free with free_full_report). */
PHD_API Full_Report_Data* phd_flat_to_full_report(const void* record, const phd_flat_layout* layout);

/* Device timing of the last phd_get_reports_u8 call on ctx, in milliseconds (CUDA events on the
 * pipeline's own stream): [0] whole pipeline, [1] front end, [2] palette select, [3] palette tie
 * path, [4] row FFT, [5] column FFT + blur binning, [6] sharpness, [7] finalize.
 * Returns the number of kernel launches of that call. */
PHD_API int phd_last_timing(const phd_context* ctx, float ms[8]);
/* How many times each of those stages was launched in that call (sub-batches), same indexing; [0] = 1. */
PHD_API int phd_last_stage_launches(const phd_context* ctx, int n[8]);
/* 1 when that call ran the front end and the row FFT as ONE launch (role-switching persistent CTAs; its time is then
 * reported as stage 1 and stage 4 is zero), 0 when they were two launches. */
PHD_API int phd_last_fused(const phd_context* ctx);

/* Test hooks (parity tests call these through the C ABI; they are not needed by applications). */
PHD_API int phd_debug_group_sweep(phd_context* ctx, const phd_params* p, uint16_t* out_2pow24 /* host */);       /* product path (integer fast path + FP64 edge path) */
PHD_API int phd_debug_group_sweep_exact(phd_context* ctx, const phd_params* p, uint16_t* out_2pow24 /* host */); /* FP64 transcription of the reference arithmetic */
PHD_API int phd_debug_bin_map(phd_context* ctx, int width, int height, int nr, int na,
                      uint16_t* map /* host, height*(width/2+1) */, int* counts /* host, na*nr */);
PHD_API int phd_debug_power_spectrum(phd_context* ctx, const uint8_t* rgb /* host */, int width, int height,
                             float* power /* host, height*(width/2+1), |FFT(gray - 0.5)|^2 */);
PHD_API int phd_debug_group_counts(phd_context* ctx, const uint8_t* rgb /* host */, int width, int height,
                           const phd_params* p, int* counts /* host, T */);

#ifdef __cplusplus
}
#endif
#endif /* PHOTOHIVE_DSP_H */
