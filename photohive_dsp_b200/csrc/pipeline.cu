// Host side of libreport_data.so: the C ABI of include/photohive_dsp.h.
//
// Part 1 mirrors src/interface.c:20-111 (get_full_report_data / free_full_report) and
// src/blur_profile.c:140-180 (get_blur_profile_visual); the stage ORDER of interface.c:28-93 is kept as a
// sequence of kernels on one stream.  Part 2 is the batch interface.  There is no CPU implementation of
// any stage in this file: without a CUDA device every entry point fails loudly.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "phd_internal.h"

namespace {

struct ShapePlan {
    int W, H, nr, na;
    FftPlan row, col;
    float2* tw_row = nullptr;
    float2* tw_col = nullptr;
    u16* binmap = nullptr;
    int* bincount = nullptr;
    // sides too long for the shared-memory kernels (fft.cu: four-step transforms through HBM)
    bool long_row = false, long_col = false;
    PhdLongFft lrow{}, lcol{};
};

struct ParamTables {
    phd_params p;
    double* centres = nullptr;  // device [3*T]: group centre h, s, v
    float* sv_f = nullptr;      // device [T]: (float)(s*v) of the centre
    unsigned char* tabs = nullptr;  // device: class / reciprocal tables of pixel_cells.cuh
    unsigned char* exc = nullptr;   // device: exceptional-colour table (pixel_cells.cuh, phd_exc_bytes); shared by every
                                    // parameter set with the same h_partitions, owned by phd_context::exc_tables
};

}  // namespace

struct phd_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::mutex mu;
    std::vector<ShapePlan> shapes;
    std::vector<ParamTables> tables;
    std::vector<std::pair<int, unsigned char*>> exc_tables;  // (h_partitions, exceptional-colour table: 2.4 MB at 18 bins)
    Workspace ws{};
    unsigned char* ws_zero = nullptr;  // one allocation holding every accumulator that must start at zero
    size_t ws_zero_bytes = 0;
    size_t ws_key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned char* d_rgb = nullptr;
    size_t d_rgb_bytes = 0;
    // host input: two staging buffers filled by a copy stream while the previous sub-batch computes
    cudaStream_t copy_stream = nullptr;
    unsigned char* d_stage[2] = {nullptr, nullptr};
    size_t d_stage_bytes[2] = {0, 0};
    cudaEvent_t ev_copied[2] = {nullptr, nullptr};
    cudaEvent_t ev_consumed[2] = {nullptr, nullptr};
    unsigned char* d_records = nullptr;
    size_t d_records_bytes = 0;
    // drop-in call (get_full_report_data): the three planes of doubles, their packed 8-bit copy and the exactness flag,
    // kept between calls, and one stream per plane for the uploads
    double* d_planes = nullptr;
    size_t d_planes_bytes = 0;
    unsigned char* d_u8 = nullptr;
    size_t d_u8_bytes = 0;
    int* d_flag = nullptr;
    // general-input route (f64path.cu): accumulators (one zeroed allocation) and the float gray plane
    F64Work f64{};
    unsigned char* f64_zero = nullptr;
    size_t f64_zero_bytes = 0;
    unsigned char* f64_gray = nullptr;
    size_t f64_gray_bytes = 0;
    unsigned char* long_buf[2] = {nullptr, nullptr};  // work buffers of the long (HBM) transforms, one image
    size_t long_buf_bytes[2] = {0, 0};
    // uploader threads of the drop-in call: each owns a stream, two pinned slices and their "slice free again" events
    static constexpr int kUpThreads = 8;
    static constexpr size_t kUpSlice = 2u << 20;
    unsigned char* h_ring = nullptr;  // [kUpThreads][2][kUpSlice], pinned
    cudaStream_t up_stream[kUpThreads] = {};
    cudaEvent_t up_ev[kUpThreads][2] = {};
    std::mutex dropin_mu;  // serialises whole drop-in calls (they share d_u8 across two locked sections of mu)
    std::vector<cudaEvent_t> events;
    std::vector<int> spans;  // (stage, start event index, end event index) triples of the last call
    size_t events_used = 0;
    float last_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int last_stage_launches[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int last_launches = 0;
    int last_fused = 0;  // the last call ran front end + row FFT as one launch (its time is reported as stage 1)
    char err[512] = {0};
};

namespace {

#define CUDA_TRY(ctx, expr)                                                                               \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            snprintf((ctx)->err, sizeof((ctx)->err), "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), \
                     __FILE__, __LINE__, #expr);                                                          \
            fprintf(stderr, "photohive_dsp: %s\n", (ctx)->err);                                           \
            return PHD_E_CUDA;                                                                            \
        }                                                                                                 \
    } while (0)

int fail(phd_context* ctx, int code, const char* msg) {
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    fprintf(stderr, "photohive_dsp: %s\n", msg);
    return code;
}

// src/utilities.c:64-87 (limits :11-13)
bool reference_rejects(int width, int height, char* why, size_t n) {
    if (height < 350 || width < 350) {
        snprintf(why, n, "Error: Image height and width must be greater than 350. Height: %d\tWidth%d", height, width);
        return true;
    }
    if ((long long)height * width > 12000LL * 10000LL) {
        snprintf(why, n, "Error: Image must have less than %d pixels.", 12000 * 10000);
        return true;
    }
    const float aspect = (float)height / (float)width;
    if (aspect < (1.0 / 5.0) || aspect > (5.0 / 1.0)) {
        snprintf(why, n, "Error: Invalid aspect ratio: %f", aspect);
        return true;
    }
    return false;
}

bool same_params_for_tables(const phd_params& a, const phd_params& b) {
    return a.h_partitions == b.h_partitions && a.s_partitions == b.s_partitions &&
           a.v_partitions == b.v_partitions && a.black_thresh == b.black_thresh && a.gray_thresh == b.gray_thresh;
}

int check_params(phd_context* ctx, const phd_params* p, int max_boxes) {
    if (!p) return fail(ctx, PHD_E_BAD_PARAMS, "params pointer is NULL");
    if (p->h_partitions <= 0 || p->s_partitions <= 0 || p->v_partitions <= 0)
        return fail(ctx, PHD_E_BAD_PARAMS, "ERROR: h_partitions, s_partitions and v_partitions must be nonzero");
    const long long T = (long long)p->h_partitions * p->s_partitions * p->v_partitions + p->v_partitions + 1;
    if (T > PHD_MAX_GROUPS) return fail(ctx, PHD_E_BAD_PARAMS, "palette grid larger than PHD_MAX_GROUPS groups");
    if (p->h_partitions > 360) return fail(ctx, PHD_E_BAD_PARAMS, "h_partitions > 360 gives a zero-width hue bin");
    // src/color_quantization.c:41 divides 360 / h_partitions in integers; when it does not divide, the hue index
    // reaches h_partitions and the reference writes past its group array
    if (360 % p->h_partitions != 0)
        return fail(ctx, PHD_E_UNSUPPORTED, "h_partitions must divide 360 (the reference overflows its group array otherwise)");
    if ((long long)p->s_partitions * p->v_partitions + 2 > 255)
        return fail(ctx, PHD_E_UNSUPPORTED, "s_partitions * v_partitions + 2 must fit one byte in this build");
    {
        const long long NC = ((long long)p->s_partitions * p->v_partitions + 2) * p->h_partitions * 4;
        if (phd_cell_tables_size() / 2 + (size_t)(NC + 32) * 16 + (size_t)NC * 28 > 200 * 1024)  // one of the two tables
            return fail(ctx, PHD_E_UNSUPPORTED, "palette grid too fine for the shared-memory cells of this build");
    }
    if (!(p->black_thresh >= 0.0 && p->black_thresh < 1.0 && p->gray_thresh >= 0.0 && p->gray_thresh < 1.0))
        return fail(ctx, PHD_E_BAD_PARAMS, "black_thresh and gray_thresh must lie in [0, 1)");
    if (p->radius_partitions <= 0 || p->angle_partitions <= 1 ||
        (long long)p->radius_partitions * p->angle_partitions > PHD_MAX_BINS || p->angle_partitions > 180)
        return fail(ctx, PHD_E_BAD_PARAMS, "radius/angle partitions out of range");
    if (p->linked_list_size <= 0) return fail(ctx, PHD_E_BAD_PARAMS, "linked_list_size must be positive");
    if (p->blur_cutoff_ratio_denom <= 0) return fail(ctx, PHD_E_BAD_PARAMS, "blur_cutoff_ratio_denom must be positive");
    if (max_boxes < 0) return fail(ctx, PHD_E_BAD_PARAMS, "negative box count");
    return PHD_OK;
}

void fill_dev_params(DevParams& P, const phd_params& p, int W, int H, int max_boxes, size_t stride, int aligned16) {
    memset(&P, 0, sizeof(P));
    P.W = W; P.H = H; P.fw = W / 2 + 1;
    P.Hp = (H + 3) / 4 * 4;
    P.ds = p.downsample_rate > 1 ? p.downsample_rate : 1;
    P.dw = P.ds > 1 ? W / P.ds : W;
    P.dh = P.ds > 1 ? H / P.ds : H;
    P.npx = (long long)W * H;
    P.hpx = (long long)P.dw * P.dh;
    P.hp = p.h_partitions; P.sp = p.s_partitions; P.vp = p.v_partitions;
    P.T = P.hp * P.sp * P.vp + P.vp + 1;
    P.ncls = P.sp * P.vp + 2;
    P.NC = P.ncls * P.hp * 4;
    // three 256-thread CTAs per SM with the three-word chunk layout while the cells and their max == 255 twins fit
    // its fixed stride (PHD_NCS_SMALL in frontend.cu) and the colour (0,0,0) is black but 255 is not; one 512-thread
    // CTA with four words otherwise
    const int twins = (P.sp + 1) * P.hp * 4;
    // (its drain gives every (class, hue bin) pair one thread: pairs with a twin first, the rest from the next warp on)
    const int drain_threads = (((P.sp + 1) * P.hp + 31) & ~31) + (P.sp * (P.vp - 1) + 1) * P.hp;
    P.fe_threads = (P.NC + twins + 32 <= 832 && drain_threads <= 256 && p.black_thresh > 0.0 && p.black_thresh <= 0.999999)
                       ? 256 : 512;
    P.chunk = P.fe_threads * 16;
    P.nchunks = (int)((P.hpx + P.chunk - 1) / P.chunk);
    // src/color_quantization.c:41-45
    P.Lh = (double)(360 / P.hp);
    P.Ls = (1 - p.gray_thresh) / P.sp;
    P.Lv = (1 - p.black_thresh) / P.vp;
    P.bt = p.black_thresh; P.gt = p.gray_thresh;
    P.coverage = p.coverage_thresh;
    P.L = p.linked_list_size;
    P.qw = p.quantity_weight; P.svw = p.saturation_value_weight;
    P.nr = p.radius_partitions; P.na = p.angle_partitions; P.nbins = P.nr * P.na;
    P.streak = p.fft_streak_thresh; P.magthr = p.magnitude_thresh; P.denom = p.blur_cutoff_ratio_denom;
    P.max_boxes = max_boxes;
    P.image_stride = stride;
    P.aligned16 = aligned16;
}

// Group centres exactly as initialize_octree computes them (src/color_quantization.c:58-98); host doubles, no FMA.
// The caches of per-parameter tables and per-shape plans are bounded (a photo or crop collection with many sizes must
// not grow device memory without limit): beyond the bound the least recently used entry is freed -- after the context's
// stream has drained, since a queued kernel may still read it.
constexpr size_t kMaxTables = 16, kMaxExcTables = 4, kMaxShapes = 32;

int get_tables(phd_context* ctx, const phd_params& p, ParamTables** out) {
    for (size_t i = 0; i < ctx->tables.size(); i++)
        if (same_params_for_tables(ctx->tables[i].p, p)) {
            if (i + 1 != ctx->tables.size()) std::rotate(ctx->tables.begin() + i, ctx->tables.begin() + i + 1, ctx->tables.end());
            ctx->tables.back().p = p;  // the non-table parameters of this call
            *out = &ctx->tables.back();
            return PHD_OK;
        }
    if (ctx->tables.size() >= kMaxTables) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ParamTables& old = ctx->tables.front();
        cudaFree(old.centres); cudaFree(old.sv_f); cudaFree(old.tabs);
        ctx->tables.erase(ctx->tables.begin());
    }
    const int hp = p.h_partitions, sp = p.s_partitions, vp = p.v_partitions;
    const int T = hp * sp * vp + vp + 1;
    std::vector<double> c(3 * (size_t)T, 0.0);
    std::vector<float> sv(T, 0.f);
    double* gh = c.data();
    double* gs = gh + T;
    double* gv = gs + T;
    volatile double Lh = (double)(360 / hp);
    volatile double Ls = (1 - p.gray_thresh) / sp;
    volatile double Lv = (1 - p.black_thresh) / vp;
    volatile double half_h = Lh / 2, s_offs = Ls / 2 + p.gray_thresh, v_offs = Lv / 2 + p.black_thresh;
    int i = 0;
    for (int h = 0; h < hp; h++)
        for (int s = 0; s < sp; s++)
            for (int v = 0; v < vp; v++) {
                i = (h * sp + s) * vp + v;
                volatile double a = h * Lh, b = s * Ls, d = v * Lv;  // products rounded before the add
                gh[i] = a + half_h;
                gs[i] = b + s_offs;
                gv[i] = d + v_offs;
            }
    volatile double L_gray = (1.0f - p.black_thresh) / (double)vp;
    for (int j = 0; j < vp; j++) {
        i++;
        volatile double a = L_gray * j;
        gh[i] = 0; gs[i] = 0; gv[i] = a + v_offs;
    }
    i++;
    gh[i] = gs[i] = gv[i] = 0;
    for (int g = 0; g < T; g++) {
        volatile double prod = gs[g] * gv[g];
        sv[g] = (float)prod;
    }
    ParamTables t;
    t.p = p;
    CUDA_TRY(ctx, cudaMalloc(&t.centres, sizeof(double) * 3 * T));
    CUDA_TRY(ctx, cudaMalloc(&t.sv_f, sizeof(float) * T));
    CUDA_TRY(ctx, cudaMemcpyAsync(t.centres, c.data(), sizeof(double) * 3 * T, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(t.sv_f, sv.data(), sizeof(float) * T, cudaMemcpyHostToDevice, ctx->stream));
    // per-parameter tables, built on the device with the reference's double arithmetic
    DevParams P;
    fill_dev_params(P, p, 1024, 1024, 0, 0, 0);
    int* ok_dev = nullptr;
    int ok = 1;
    CUDA_TRY(ctx, cudaMalloc(&t.tabs, phd_cell_tables_size()));
    unsigned char* new_exc = nullptr;
    for (auto& e : ctx->exc_tables)
        if (e.first == hp) t.exc = e.second;
    if (!t.exc) {
        if (ctx->exc_tables.size() >= kMaxExcTables) {
            // drop the oldest code table no cached parameter set refers to any more
            for (size_t i = 0; i < ctx->exc_tables.size(); i++) {
                bool used = false;
                for (auto& o : ctx->tables) used = used || o.exc == ctx->exc_tables[i].second;
                if (!used) {
                    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
                    cudaFree(ctx->exc_tables[i].second);
                    ctx->exc_tables.erase(ctx->exc_tables.begin() + i);
                    break;
                }
            }
        }
        CUDA_TRY(ctx, cudaMalloc(&new_exc, phd_exc_table_bytes(hp)));
        t.exc = new_exc;
    }
    CUDA_TRY(ctx, cudaMalloc(&ok_dev, sizeof(int)));
    CUDA_TRY(ctx, cudaMemcpyAsync(ok_dev, &ok, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    phd_launch_build_cell_tables(P, t.tabs, new_exc, ok_dev, ctx->stream);
    CUDA_TRY(ctx, cudaMemcpyAsync(&ok, ok_dev, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ok_dev);
    if (!ok) {
        cudaFree(t.centres); cudaFree(t.sv_f); cudaFree(t.tabs); cudaFree(new_exc);
        return fail(ctx, PHD_E_UNSUPPORTED, "thresholds put a colour outside the palette grid (the reference would index past it)");
    }
    if (new_exc) ctx->exc_tables.push_back({hp, new_exc});
    ctx->tables.push_back(t);
    *out = &ctx->tables.back();
    return PHD_OK;
}

int get_shape(phd_context* ctx, int W, int H, int nr, int na, ShapePlan** out) {
    for (size_t i = 0; i < ctx->shapes.size(); i++) {
        const ShapePlan& c = ctx->shapes[i];
        if (c.W == W && c.H == H && c.nr == nr && c.na == na) {
            if (i + 1 != ctx->shapes.size()) std::rotate(ctx->shapes.begin() + i, ctx->shapes.begin() + i + 1, ctx->shapes.end());
            *out = &ctx->shapes.back();
            return PHD_OK;
        }
    }
    if (ctx->shapes.size() >= kMaxShapes) {
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ShapePlan& old = ctx->shapes.front();
        cudaFree(old.tw_row); cudaFree(old.tw_col); cudaFree(old.binmap); cudaFree(old.bincount);
        phd_long_fft_destroy(&old.lrow); phd_long_fft_destroy(&old.lcol);
        ctx->shapes.erase(ctx->shapes.begin());
    }
    ShapePlan s;
    s.W = W; s.H = H; s.nr = nr; s.na = na;
    // sides beyond what two shared-memory buffers hold (rows: 12,800 pixels; columns: ~11,000 next to the blur bins) are
    // transformed through HBM; the reference admits sides up to 24,494 (src/utilities.c:11-13)
    s.long_row = (size_t)W * 2 * sizeof(float2) > 200 * 1024;
    s.long_col = (size_t)((H + 3) / 4 * 4) * 2 * sizeof(float2) + (size_t)2 * nr * na * sizeof(u32) > 200 * 1024;
    memset(&s.row, 0, sizeof(s.row));
    memset(&s.col, 0, sizeof(s.col));
    s.row.n = W; s.col.n = H;
    if ((!s.long_row && phd_fft_make_plan(W, &s.row)) || (!s.long_col && phd_fft_make_plan(H, &s.col)))
        return fail(ctx, PHD_E_UNSUPPORTED, "image side has too many or too large prime factors: FFT length not supported by this build");
    if ((s.long_row && phd_long_fft_create(W, &s.lrow, ctx->stream)) || (s.long_col && phd_long_fft_create(H, &s.lcol, ctx->stream)))
        return fail(ctx, PHD_E_UNSUPPORTED, "cannot plan the long (HBM) transform of this image side");
    const int Hp = (H + 3) / 4 * 4;
    const size_t nspec = (size_t)(W / 2 + 1) * Hp;
    const size_t pe_row = s.long_row ? 0 : phd_fft_pass_table_entries(s.row), pe_col = s.long_col ? 0 : phd_fft_pass_table_entries(s.col);
    // one allocation per direction: n twiddles, the pass tables, and for a Bluestein plan the chirp (n) and its transform (m)
    CUDA_TRY(ctx, cudaMalloc(&s.tw_row, sizeof(float2) * (W + pe_row + (s.row.m > 0 ? W + s.row.m : 0))));
    CUDA_TRY(ctx, cudaMalloc(&s.tw_col, sizeof(float2) * (H + pe_col + (s.col.m > 0 ? H + s.col.m : 0))));
    CUDA_TRY(ctx, cudaMalloc(&s.binmap, sizeof(u16) * nspec + 16));  // + 16: the column kernel rounds its last slice copy up
    CUDA_TRY(ctx, cudaMalloc(&s.bincount, sizeof(int) * nr * na));
    if (!s.long_row) {
        phd_fill_twiddles(s.tw_row, W, ctx->stream);
        phd_fft_fill_pass_tables(s.tw_row + W, s.row, ctx->stream);
    }
    if (!s.long_col) {
        phd_fill_twiddles(s.tw_col, H, ctx->stream);
        phd_fft_fill_pass_tables(s.tw_col + H, s.col, ctx->stream);
    }
    if (s.row.m > 0) {
        float2* chirp = s.tw_row + W + pe_row;
        phd_fft_fill_bluestein(chirp, chirp + W, W, s.row.m, ctx->stream);
        s.row.chirp = chirp; s.row.bhat = chirp + W;
    }
    if (s.col.m > 0) {
        float2* chirp = s.tw_col + H + pe_col;
        phd_fft_fill_bluestein(chirp, chirp + H, H, s.col.m, ctx->stream);
        s.col.chirp = chirp; s.col.bhat = chirp + H;
    }
    phd_launch_bin_map(W, H, Hp, nr, na, s.binmap, s.bincount, ctx->stream);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    s.row.tw = s.tw_row;
    s.col.tw = s.tw_col;

    ctx->shapes.push_back(s);
    *out = &ctx->shapes.back();
    return PHD_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int ensure_workspace(phd_context* ctx, const DevParams& P, int cap, int cap_fft) {
    const size_t key[8] = {(size_t)cap, (size_t)P.T * 65536 + (size_t)P.NC, (size_t)P.nchunks, (size_t)P.H, (size_t)P.fw,
                           (size_t)P.nbins, (size_t)P.max_boxes, (size_t)cap_fft};
    if (memcmp(key, ctx->ws_key, sizeof(key)) == 0 && ctx->ws.capacity == cap) return PHD_OK;
    Workspace& w = ctx->ws;
    cudaFree(w.counts_chunk); cudaFree(w.plan); cudaFree(w.pal_n); cudaFree(w.parent_ids); cudaFree(w.tie_list);
    cudaFree(w.tie_n); cudaFree(w.tie_groups); cudaFree(w.dropped); cudaFree(w.spec); cudaFree(w.boxes);
    cudaFree(w.cells_tie); cudaFree(w.work); cudaFree(w.span32); cudaFree(w.span64);
    cudaFree(ctx->ws_zero);
    memset(&w, 0, sizeof(w));
    ctx->ws_zero = nullptr;
    const size_t T = P.T, c = cap;
    CUDA_TRY(ctx, cudaMalloc(&w.counts_chunk, sizeof(u16) * c * P.nchunks * T));
    CUDA_TRY(ctx, cudaMalloc(&w.plan, sizeof(GroupPlan) * c * T));
    CUDA_TRY(ctx, cudaMalloc(&w.cells_tie, sizeof(u64) * c * PHD_CELL_Q * P.NC));
    CUDA_TRY(ctx, cudaMalloc(&w.work, sizeof(u32) * c * P.nchunks));
    {
        const size_t spans = phd_fe_max_spans(P, cap);
        CUDA_TRY(ctx, cudaMalloc(&w.span32, sizeof(u32) * spans * 3 * P.NC));
        CUDA_TRY(ctx, cudaMalloc(&w.span64, sizeof(u64) * spans * 2 * P.NC));
    }
    CUDA_TRY(ctx, cudaMalloc(&w.pal_n, sizeof(int) * c));
    CUDA_TRY(ctx, cudaMalloc(&w.parent_ids, sizeof(int) * c * T));
    CUDA_TRY(ctx, cudaMalloc(&w.tie_list, sizeof(int) * c * T));
    CUDA_TRY(ctx, cudaMalloc(&w.tie_n, sizeof(int) * c));
    CUDA_TRY(ctx, cudaMalloc(&w.tie_groups, sizeof(int) * c));
    CUDA_TRY(ctx, cudaMalloc(&w.dropped, sizeof(long long) * c));
    CUDA_TRY(ctx, cudaMalloc(&w.spec, sizeof(float2) * (size_t)cap_fft * P.Hp * P.fw));
    CUDA_TRY(ctx, cudaMalloc(&w.boxes, sizeof(int) * 4 * c * (P.max_boxes > 0 ? P.max_boxes : 1)));
    // zero-initialised accumulators, contiguous so one memset per sub-batch clears them
    size_t off = 0;
    const size_t o_hist = off; off = align_up(off + sizeof(u32) * c * T, 256);
    const size_t o_iacc = off; off = align_up(off + sizeof(ImageAcc) * c, 256);
    const size_t o_sacc = off; off = align_up(off + sizeof(SlotAcc) * c * T, 256);
    const size_t o_bins = off; off = align_up(off + sizeof(u64) * c * P.nbins, 256);
    const size_t o_maxp = off; off = align_up(off + sizeof(u32) * c, 256);
    const size_t o_sharp = off; off = align_up(off + sizeof(SharpAcc) * c * (P.max_boxes > 0 ? P.max_boxes : 1), 256);
    const size_t o_cells = off; off = align_up(off + sizeof(u64) * c * PHD_CELL_Q * P.NC, 256);
    const size_t o_workn = off; off = align_up(off + sizeof(u32), 256);
    const size_t o_queue = off; off = align_up(off + sizeof(u32), 256);
    CUDA_TRY(ctx, cudaMalloc(&ctx->ws_zero, off));
    ctx->ws_zero_bytes = off;
    w.hist = reinterpret_cast<u32*>(ctx->ws_zero + o_hist);
    w.iacc = reinterpret_cast<ImageAcc*>(ctx->ws_zero + o_iacc);
    w.sacc = reinterpret_cast<SlotAcc*>(ctx->ws_zero + o_sacc);
    w.binsum = reinterpret_cast<u64*>(ctx->ws_zero + o_bins);
    w.maxpow = reinterpret_cast<u32*>(ctx->ws_zero + o_maxp);
    w.sharp = reinterpret_cast<SharpAcc*>(ctx->ws_zero + o_sharp);
    w.cells = reinterpret_cast<u64*>(ctx->ws_zero + o_cells);
    w.work_n = reinterpret_cast<u32*>(ctx->ws_zero + o_workn);
    w.queue = reinterpret_cast<u32*>(ctx->ws_zero + o_queue);
    w.capacity = cap;
    memcpy(ctx->ws_key, key, sizeof(key));
    return PHD_OK;
}

int ensure_bytes(phd_context* ctx, unsigned char** buf, size_t* have, size_t need) {
    if (*have >= need) return PHD_OK;
    cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    CUDA_TRY(ctx, cudaMalloc(buf, need));
    *have = need;
    return PHD_OK;
}

// Upload of PAGEABLE host memory.  A pageable cudaMemcpy moves about 11 GB/s on this platform; here a few threads
// copy 2 MB slices into their own pinned slices and queue the DMA on their own streams (double buffered), which is
// bound by the host's memcpy bandwidth instead (about 4x faster).  Blocks until every byte is on the device.
struct UploadSeg {
    const unsigned char* src;
    unsigned char* dst;
    size_t bytes;
};

cudaError_t threaded_upload(phd_context* ctx, const UploadSeg* segs, int nseg) {
    constexpr int NT = phd_context::kUpThreads;
    constexpr size_t SL = phd_context::kUpSlice;
    cudaError_t e = cudaSuccess;
    if (!ctx->h_ring) e = cudaHostAlloc(&ctx->h_ring, (size_t)NT * 2 * SL, cudaHostAllocDefault);
    for (int t = 0; t < NT && e == cudaSuccess; t++) {
        if (!ctx->up_stream[t]) e = cudaStreamCreateWithFlags(&ctx->up_stream[t], cudaStreamNonBlocking);
        for (int b = 0; b < 2 && e == cudaSuccess; b++)
            if (!ctx->up_ev[t][b]) e = cudaEventCreateWithFlags(&ctx->up_ev[t][b], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return e;
    std::vector<size_t> first(nseg + 1, 0);  // first slice of every segment
    for (int i = 0; i < nseg; i++) first[i + 1] = first[i] + (segs[i].bytes + SL - 1) / SL;
    const size_t nslices = first[nseg];
    if (nslices == 0) return cudaSuccess;
    int nthreads = (int)std::min<size_t>(NT, nslices);
    const unsigned hc = std::thread::hardware_concurrency();
    if (hc > 0 && (unsigned)nthreads > hc) nthreads = (int)hc;
    cudaError_t te[NT];
    auto upload = [&](int t) {
        cudaSetDevice(ctx->device);
        cudaError_t err = cudaSuccess;
        size_t mine = 0;
        int seg = 0;
        for (size_t sl = t; sl < nslices && err == cudaSuccess; sl += nthreads, mine++) {
            while (sl >= first[seg + 1]) seg++;
            const size_t off = (sl - first[seg]) * SL;
            const size_t bytes = std::min(SL, segs[seg].bytes - off);
            const int b = (int)(mine & 1);
            unsigned char* stage = ctx->h_ring + ((size_t)t * 2 + b) * SL;
            if (mine >= 2) err = cudaEventSynchronize(ctx->up_ev[t][b]);  // the slice's previous DMA is done
            if (err != cudaSuccess) break;
            memcpy(stage, segs[seg].src + off, bytes);
            err = cudaMemcpyAsync(segs[seg].dst + off, stage, bytes, cudaMemcpyHostToDevice, ctx->up_stream[t]);
            if (err == cudaSuccess) err = cudaEventRecord(ctx->up_ev[t][b], ctx->up_stream[t]);
        }
        if (err == cudaSuccess) err = cudaStreamSynchronize(ctx->up_stream[t]);
        te[t] = err;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; t++) pool.emplace_back(upload, t);
    upload(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < nthreads; t++)
        if (te[t] != cudaSuccess) e = te[t];
    return e;
}

bool is_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool is_pageable_host_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

// Row and column stage of `nimg` images: the shared-memory kernels, or image by image through HBM for long sides.
int long_buffers(phd_context* ctx, const ShapePlan& shape, const DevParams& P) {
    size_t need = 0;
    if (shape.long_row) need = std::max(need, (size_t)(P.Hp / 2) * shape.lrow.L * sizeof(float2));
    if (shape.long_col) need = std::max(need, (size_t)P.fw * shape.lcol.L * sizeof(float2));
    for (int b = 0; b < 2 && need; b++) {
        const int rc = ensure_bytes(ctx, &ctx->long_buf[b], &ctx->long_buf_bytes[b], need);
        if (rc != PHD_OK) return rc;
    }
    return PHD_OK;
}

int launch_rows_any(phd_context* ctx, const ShapePlan& shape, const uint8_t* rgb, size_t image_stride, const float* gray32,
                    const DevParams& P, int nimg, float2* spec, cudaStream_t st, int* launches) {
    if (!shape.long_row) {
        if (gray32) return phd_launch_fft_rows_gray(gray32, P, shape.row, spec, st, launches);
        return phd_launch_fft_rows(rgb, P, nimg, shape.row, spec, st, launches);
    }
    int rc = long_buffers(ctx, shape, P);
    if (rc != PHD_OK) return rc;
    for (int i = 0; i < nimg; i++)
        if (phd_launch_long_rows(rgb ? rgb + (size_t)i * image_stride : nullptr, gray32, P, shape.lrow,
                                 reinterpret_cast<float2*>(ctx->long_buf[0]), reinterpret_cast<float2*>(ctx->long_buf[1]),
                                 spec + (size_t)i * P.fw * P.Hp, st, launches))
            return 1;
    return 0;
}

int launch_cols_any(phd_context* ctx, const ShapePlan& shape, const DevParams& P, int nimg, float2* spec, Workspace& ws,
                    float* power_out, cudaStream_t st, int* launches) {
    if (!shape.long_col) return phd_launch_fft_cols_blur(P, nimg, shape.col, spec, shape.binmap, ws, power_out, st, launches);
    int rc = long_buffers(ctx, shape, P);
    if (rc != PHD_OK) return rc;
    for (int i = 0; i < nimg; i++) {
        Workspace one = ws;  // accumulators of image i
        one.iacc += i;
        one.binsum += (size_t)i * P.nbins;
        one.maxpow += i;
        if (phd_launch_long_cols(P, shape.lcol, spec + (size_t)i * P.fw * P.Hp, reinterpret_cast<float2*>(ctx->long_buf[0]),
                                 reinterpret_cast<float2*>(ctx->long_buf[1]), shape.binmap, one, power_out, st, launches))
            return 1;
    }
    return 0;
}

// Images per FFT sub-batch (the row-transformed spectra of one sub-batch live in `spec`).  Measured on B200
// (profiles/): bigger launches win over keeping the spectra L2 resident (48.4 k images/s at 48 images per launch,
// 50.5 k at 512 vs 45.9 k at 11), so the sub-batch is the whole palette batch, capped at 8 GB of spectra.
int pick_fft_batch(const DevParams& P, int n_images) {
    const char* env = getenv("PHD_SUB_BATCH");
    long long sub = env ? atoll(env) : 0;
    if (sub <= 0) {
        const double per = (double)P.Hp * P.fw * sizeof(float2);
        sub = (long long)(8.0e9 / per);
        if (sub < 1) sub = 1;
    }
    if (sub > n_images) sub = n_images;
    return (int)sub;
}

// Images per palette group (front end, parent selection, accumulation, sharpness, finalize): large, so the
// one-CTA-per-image kernels fill the machine; bounded by the per-chunk histogram storage.
int pick_palette_batch(const DevParams& P, int n_images, bool host_input) {
    const char* env = getenv("PHD_PALETTE_BATCH");
    long long pb = env ? atoll(env) : 0;
    if (pb <= 0) {
        pb = host_input ? 32 : 512;  // host input: small sub-batches keep the copy engine and the SMs both busy
        const double per = (double)P.nchunks * P.T * sizeof(u16);
        const long long cap = (long long)(1024.0 * 1024 * 1024 / per);
        if (pb > cap) pb = cap;
        if (pb < 1) pb = 1;
    }
    if (pb > n_images) pb = n_images;
    return (int)pb;
}

enum { ST_FRONT = 1, ST_SELECT, ST_ACCUM, ST_ROWS, ST_COLS, ST_SHARP, ST_FINAL };

// The pipeline on device-resident input.  records_dev: device buffer for n_images records.
int run_pipeline(phd_context* ctx, const uint8_t* rgb_host_or_dev, bool input_on_device, int n_images, int W, int H,
                 size_t image_stride, const int* boxes_host, int max_boxes, const phd_params& p,
                 unsigned char* records_dev, const phd_flat_layout& lay) {
    DevParams P;
    const size_t tight = (size_t)W * H * 3;
    const size_t dev_stride = input_on_device ? image_stride : align_up(tight, 16);
    int aligned16 = 1;
    if (input_on_device) aligned16 = (((uintptr_t)rgb_host_or_dev % 16) == 0 && image_stride % 16 == 0) ? 1 : 0;
    fill_dev_params(P, p, W, H, max_boxes, dev_stride, aligned16);
    if (P.dw < 1 || P.dh < 1) return fail(ctx, PHD_E_BAD_PARAMS, "downsample_rate leaves no pixels");

    ShapePlan* shape;
    ParamTables* tab;
    int rc;
    if ((rc = get_shape(ctx, W, H, P.nr, P.na, &shape)) != PHD_OK) return rc;
    if ((rc = get_tables(ctx, p, &tab)) != PHD_OK) return rc;
    const int pb = pick_palette_batch(P, n_images, !input_on_device);
    const int fb = pick_fft_batch(P, pb);
    if ((rc = ensure_workspace(ctx, P, pb, fb)) != PHD_OK) return rc;
    if (!input_on_device)
        for (int b = 0; b < 2; b++)
            if ((rc = ensure_bytes(ctx, &ctx->d_stage[b], &ctx->d_stage_bytes[b], dev_stride * pb)) != PHD_OK) return rc;
    cudaStream_t st = ctx->stream;
    ctx->spans.clear();
    ctx->events_used = 0;
    auto mark = [&](int* idx) -> int {
        if (ctx->events_used == ctx->events.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return 1;
            ctx->events.push_back(e);
        }
        *idx = (int)ctx->events_used++;
        return cudaEventRecord(ctx->events[*idx], st) != cudaSuccess;
    };
    for (int i = 0; i < 8; i++) ctx->last_stage_launches[i] = 0;
    auto span = [&](int stage, int a, int b) {
        ctx->spans.push_back(stage); ctx->spans.push_back(a); ctx->spans.push_back(b);
        ctx->last_stage_launches[stage]++;
    };
    int launches = 0;
    int e_begin, e_end, e0, e1;
    if (mark(&e_begin)) return fail(ctx, PHD_E_CUDA, "cudaEventRecord failed");
    // host input: sub-batch i+1 is copied (copy stream) while sub-batch i computes
    auto stage_copy = [&](int first, int slot, bool wait_consumed) -> cudaError_t {
        const int n = (n_images - first < pb) ? (n_images - first) : pb;
        cudaError_t e = cudaSuccess;
        if (wait_consumed) e = cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_consumed[slot], 0);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(ctx->d_stage[slot], dev_stride, rgb_host_or_dev + (size_t)first * image_stride,
                                  image_stride, tight, n, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copied[slot], ctx->copy_stream);
        return e;
    };
    // pageable host input: the same double buffering, but the copy is threaded_upload, which blocks the host -- so
    // sub-batch i+1 is uploaded AFTER the kernels of sub-batch i are queued, while they run
    const bool pageable = !input_on_device && is_pageable_host_pointer(rgb_host_or_dev);
    auto stage_upload = [&](int first, int slot) -> cudaError_t {
        const int n = (n_images - first < pb) ? (n_images - first) : pb;
        std::vector<UploadSeg> segs(n);
        for (int i = 0; i < n; i++)
            segs[i] = {rgb_host_or_dev + (size_t)(first + i) * image_stride, ctx->d_stage[slot] + (size_t)i * dev_stride, tight};
        return threaded_upload(ctx, segs.data(), n);
    };
    if (pageable) {
        CUDA_TRY(ctx, cudaStreamSynchronize(st));  // the staging buffers are reused
        CUDA_TRY(ctx, stage_upload(0, 0));
    } else if (!input_on_device) {
        // order the copy stream after whatever this context did before (the staging buffers are reused)
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_consumed[0], st));
        CUDA_TRY(ctx, stage_copy(0, 0, true));
    }
    int batch_index = 0;
    for (int first = 0; first < n_images; first += pb, batch_index++) {
        const int n = (n_images - first < pb) ? (n_images - first) : pb;
        const uint8_t* d_in;
        if (input_on_device) d_in = rgb_host_or_dev + (size_t)first * image_stride;
        else if (pageable) d_in = ctx->d_stage[batch_index & 1];
        else {
            const int slot = batch_index & 1;
            if (first + pb < n_images) CUDA_TRY(ctx, stage_copy(first + pb, slot ^ 1, batch_index >= 1));
            CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_copied[slot], 0));
            d_in = ctx->d_stage[slot];
        }
        int max_w = 0, max_h = 0;
        if (max_boxes > 0) {
            const int* b = boxes_host + (size_t)first * max_boxes * 4;
            for (int i = 0; i < n * max_boxes; i++) {
                const int w = b[4 * i + 3] - b[4 * i + 2], h = b[4 * i + 1] - b[4 * i];
                if (w > max_w) max_w = w;
                if (h > max_h) max_h = h;
            }
            if (max_w > W) max_w = W;
            if (max_h > H) max_h = H;
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ws.boxes, b, sizeof(int) * 4 * (size_t)n * max_boxes,
                                          cudaMemcpyHostToDevice, st));
        }
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->ws_zero, 0, ctx->ws_zero_bytes, st));
        phd_fe_plan(P, n);
        mark(&e0);
        // PHD_FUSED=1: front end and row FFT as one launch of role-switching persistent CTAs (fft.cu: k_front_rows) where
        // a fused kernel exists for the shape and the spectra of the whole group fit the spectrum buffer.  Records are
        // byte-identical either way (tested); measured on B200 the fused launch is 5 % SLOWER than the two launches
        // (45.1 vs 25.5 + 17.6 ms per 4096 images: both roles are bound by the same shared-memory pipe, and the register
        // prefetch of the row stage had to go to fit 80 registers), so it is opt-in.  profiles/README.md has the numbers.
        const bool fuse_enabled = getenv("PHD_FUSED") && atoi(getenv("PHD_FUSED")) != 0;
        const bool fused = fuse_enabled && fb >= n &&
                           phd_launch_front_rows(d_in, P, n, tab->tabs, tab->exc, shape->row, ctx->ws, st, &launches);
        ctx->last_fused = fused ? 1 : 0;
        if (!fused) phd_launch_pixels(d_in, P, n, tab->tabs, tab->exc, ctx->ws, st, &launches);
        mark(&e1); span(ST_FRONT, e0, e1); e0 = e1;
        phd_launch_palette_select(P, n, tab->centres, tab->sv_f, ctx->ws, st, &launches);
        mark(&e1); span(ST_SELECT, e0, e1); e0 = e1;
        phd_launch_palette_ties(d_in, P, n, tab->tabs, tab->exc, ctx->ws, st, &launches);
        mark(&e1); span(ST_ACCUM, e0, e1); e0 = e1;
        for (int f0 = 0; f0 < n; f0 += fb) {
            const int nf = (n - f0 < fb) ? (n - f0) : fb;
            Workspace sub = ctx->ws;  // accumulators of images f0.. of this group
            sub.iacc += f0;
            sub.binsum += (size_t)f0 * P.nbins;
            sub.maxpow += f0;
            if (!fused) {
                if (launch_rows_any(ctx, *shape, d_in + (size_t)f0 * dev_stride, dev_stride, nullptr, P, nf, ctx->ws.spec, st, &launches))
                    return fail(ctx, PHD_E_UNSUPPORTED, "row FFT of this image width cannot be served");
                mark(&e1); span(ST_ROWS, e0, e1); e0 = e1;
            }
            if (launch_cols_any(ctx, *shape, P, nf, ctx->ws.spec, sub, nullptr, st, &launches))
                return fail(ctx, PHD_E_UNSUPPORTED, "column FFT of this image height cannot be served");
            mark(&e1); span(ST_COLS, e0, e1); e0 = e1;
        }
        phd_launch_sharpness(d_in, P, n, max_w, max_h, ctx->ws, st, &launches);
        mark(&e1); span(ST_SHARP, e0, e1); e0 = e1;
        phd_launch_finalize(P, n, tab->centres, shape->bincount, ctx->ws, lay,
                            records_dev + (size_t)first * lay.record_bytes, st, &launches);
        mark(&e1); span(ST_FINAL, e0, e1);
        if (!input_on_device) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_consumed[batch_index & 1], st));
        CUDA_TRY(ctx, cudaGetLastError());
        if (pageable && first + pb < n_images) {
            const int next_slot = (batch_index & 1) ^ 1;
            if (batch_index >= 1) CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_consumed[next_slot]));  // its reader is done
            CUDA_TRY(ctx, stage_upload(first + pb, next_slot));
        }
    }
    if (mark(&e_end)) return fail(ctx, PHD_E_CUDA, "cudaEventRecord failed");
    span(0, e_begin, e_end);
    ctx->last_launches = launches;
    return PHD_OK;
}

// The general-input route: ONE image given as three device planes of arbitrary doubles (see f64path.cu).  Same stage
// order as run_pipeline; the palette is done in two FP64 passes around the unchanged parent selection, the FFT reads a
// float gray plane.
int run_pipeline_f64(phd_context* ctx, const double* planes_dev, int W, int H, const int* boxes_host, int max_boxes,
                     const phd_params& p, unsigned char* records_dev, const phd_flat_layout& lay) {
    DevParams P;
    fill_dev_params(P, p, W, H, max_boxes, 0, 0);
    if (P.dw < 1 || P.dh < 1) return fail(ctx, PHD_E_BAD_PARAMS, "downsample_rate leaves no pixels");
    ShapePlan* shape;
    ParamTables* tab;
    int rc;
    if ((rc = get_shape(ctx, W, H, P.nr, P.na, &shape)) != PHD_OK) return rc;
    if ((rc = get_tables(ctx, p, &tab)) != PHD_OK) return rc;
    if ((rc = ensure_workspace(ctx, P, 1, 1)) != PHD_OK) return rc;
    if (phd_f64_accumulate_smem(P) > 200 * 1024)
        return fail(ctx, PHD_E_UNSUPPORTED, "palette grid too fine for the general-input route of this build");
    P.cpp = P.nchunks;  // one span: the tie path of this route ranks pixels itself, nothing is folded from span sums
    P.nspans = 1;
    const size_t nb1 = max_boxes > 0 ? max_boxes : 1;
    const size_t zbytes = sizeof(double) * (10 + 4 * (size_t)P.T + 2 * nb1);
    if ((rc = ensure_bytes(ctx, &ctx->f64_zero, &ctx->f64_zero_bytes, zbytes)) != PHD_OK) return rc;
    if ((rc = ensure_bytes(ctx, &ctx->f64_gray, &ctx->f64_gray_bytes, sizeof(float) * (size_t)P.npx)) != PHD_OK) return rc;
    F64Work& fw = ctx->f64;
    fw.acc = reinterpret_cast<double*>(ctx->f64_zero);
    fw.slots = fw.acc + 10;
    fw.sharp = fw.slots + 4 * (size_t)P.T;
    fw.gray32 = reinterpret_cast<float*>(ctx->f64_gray);

    cudaStream_t st = ctx->stream;
    ctx->spans.clear();
    ctx->events_used = 0;
    for (int i = 0; i < 8; i++) ctx->last_stage_launches[i] = 0;
    int launches = 0;
    if (max_boxes > 0)
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ws.boxes, boxes_host, sizeof(int) * 4 * (size_t)max_boxes, cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->ws_zero, 0, ctx->ws_zero_bytes, st));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->f64_zero, 0, zbytes, st));
    phd_launch_f64_front(planes_dev, P, fw, ctx->ws, st, &launches);
    phd_launch_palette_select(P, 1, tab->centres, tab->sv_f, ctx->ws, st, &launches, true);
    phd_launch_f64_accumulate(planes_dev, P, tab->centres, fw, ctx->ws, st, &launches);
    if (launch_rows_any(ctx, *shape, nullptr, 0, fw.gray32, P, 1, ctx->ws.spec, st, &launches))
        return fail(ctx, PHD_E_UNSUPPORTED, "row FFT of this image width cannot be served");
    if (launch_cols_any(ctx, *shape, P, 1, ctx->ws.spec, ctx->ws, nullptr, st, &launches))
        return fail(ctx, PHD_E_UNSUPPORTED, "column FFT of this image height cannot be served");
    phd_launch_finalize(P, 1, tab->centres, shape->bincount, ctx->ws, lay, records_dev, st, &launches, &fw);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->last_launches = launches;
    ctx->last_fused = 0;
    return PHD_OK;
}

int collect_timing(phd_context* ctx) {
    for (int i = 0; i < 8; i++) ctx->last_ms[i] = 0.f;
    for (size_t i = 0; i + 2 < ctx->spans.size(); i += 3) {
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->events[ctx->spans[i + 1]], ctx->events[ctx->spans[i + 2]]));
        ctx->last_ms[ctx->spans[i]] += ms;
    }
    return PHD_OK;
}

// Image_RGB planes (doubles k/255.0) -> packed 8-bit, with an exactness flag.
__global__ void k_ingest_f64(const double* __restrict__ r, const double* __restrict__ g, const double* __restrict__ b,
                             long long npx, uint8_t* __restrict__ out, int* __restrict__ not_8bit) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npx) return;
    const double v[3] = {r[i], g[i], b[i]};
    bool bad = false;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int k = (int)(v[c] * 255.0 + 0.5);
        k = min(max(k, 0), 255);
        if (__ddiv_rn((double)k, 255.0) != v[c]) bad = true;
        out[3 * i + c] = (uint8_t)k;
    }
    if (bad) atomicOr(not_8bit, 1);
}

std::mutex g_default_mu;
phd_context* g_default_ctx = nullptr;

phd_context* default_context() {
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_ctx) {
        int dev = 0;
        const char* env = getenv("PHD_DEVICE");
        if (env) dev = atoi(env);
        if (phd_context_create(dev, &g_default_ctx) != PHD_OK) g_default_ctx = nullptr;
    }
    return g_default_ctx;
}

}  // namespace

// =============================================================================================
// Part 2: batch interface
// =============================================================================================
extern "C" {

void phd_default_params(phd_params* p) {
    p->h_partitions = 18; p->s_partitions = 2; p->v_partitions = 3;
    p->black_thresh = 0.1; p->gray_thresh = 0.1; p->coverage_thresh = 0.95;
    p->linked_list_size = 1000; p->downsample_rate = 1;
    p->radius_partitions = 40; p->angle_partitions = 72;
    p->quantity_weight = 0.1f; p->saturation_value_weight = 0.9f;
    p->fft_streak_thresh = 1.20; p->magnitude_thresh = 0.3; p->blur_cutoff_ratio_denom = 2;
}

int phd_context_create(int device, phd_context** out) {
    if (!out) return PHD_E_BAD_PARAMS;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        fprintf(stderr, "photohive_dsp: no CUDA device available; this library has no CPU path\n");
        return PHD_E_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) {
        fprintf(stderr, "photohive_dsp: device %d out of range (%d devices)\n", device, ndev);
        return PHD_E_NO_DEVICE;
    }
    phd_context* ctx = new phd_context();
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int b = 0; ok && b < 2; b++)
        ok = cudaEventCreateWithFlags(&ctx->ev_copied[b], cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&ctx->ev_consumed[b], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        fprintf(stderr, "photohive_dsp: cannot initialise CUDA device %d: %s\n", device,
                cudaGetErrorString(cudaGetLastError()));
        delete ctx;
        return PHD_E_CUDA;
    }
    *out = ctx;
    return PHD_OK;
}

void phd_context_destroy(phd_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->shapes) {
        cudaFree(s.tw_row); cudaFree(s.tw_col); cudaFree(s.binmap); cudaFree(s.bincount);
        phd_long_fft_destroy(&s.lrow); phd_long_fft_destroy(&s.lcol);
    }
    for (auto& t : ctx->tables) { cudaFree(t.centres); cudaFree(t.sv_f); cudaFree(t.tabs); }
    for (auto& e : ctx->exc_tables) cudaFree(e.second);
    Workspace& w = ctx->ws;
    cudaFree(w.counts_chunk); cudaFree(w.plan); cudaFree(w.pal_n); cudaFree(w.parent_ids); cudaFree(w.tie_list);
    cudaFree(w.tie_n); cudaFree(w.tie_groups); cudaFree(w.dropped); cudaFree(w.spec); cudaFree(w.boxes);
    cudaFree(w.cells_tie); cudaFree(w.work); cudaFree(w.span32); cudaFree(w.span64);
    cudaFree(ctx->ws_zero); cudaFree(ctx->d_rgb); cudaFree(ctx->d_records);
    cudaFree(ctx->d_stage[0]); cudaFree(ctx->d_stage[1]);
    cudaFree(ctx->d_planes); cudaFree(ctx->d_u8); cudaFree(ctx->d_flag);
    cudaFree(ctx->f64_zero); cudaFree(ctx->f64_gray);
    cudaFree(ctx->long_buf[0]); cudaFree(ctx->long_buf[1]);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    for (int t = 0; t < phd_context::kUpThreads; t++) {
        if (ctx->up_stream[t]) cudaStreamDestroy(ctx->up_stream[t]);
        for (int b = 0; b < 2; b++)
            if (ctx->up_ev[t][b]) cudaEventDestroy(ctx->up_ev[t][b]);
    }
    for (auto e : ctx->events) cudaEventDestroy(e);
    for (int b = 0; b < 2; b++) {
        if (ctx->ev_copied[b]) cudaEventDestroy(ctx->ev_copied[b]);
        if (ctx->ev_consumed[b]) cudaEventDestroy(ctx->ev_consumed[b]);
    }
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* phd_last_error(const phd_context* ctx) { return ctx ? ctx->err : "no context"; }

int phd_flat_get_layout(const phd_params* p, int max_boxes, phd_flat_layout* out) {
    if (!p || !out || max_boxes < 0) return PHD_E_BAD_PARAMS;
    if (p->h_partitions <= 0 || p->s_partitions <= 0 || p->v_partitions <= 0 || p->radius_partitions <= 0 ||
        p->angle_partitions <= 0)
        return PHD_E_BAD_PARAMS;
    const size_t T = (size_t)p->h_partitions * p->s_partitions * p->v_partitions + p->v_partitions + 1;
    const size_t nb = (size_t)p->radius_partitions * p->angle_partitions;
    size_t off = align_up(sizeof(phd_flat_head), 16);
    out->off_palette_hsv = off; off += sizeof(double) * 3 * T;
    out->off_palette_pct = off; off += sizeof(double) * T;
    out->off_parent_ids = off; off = align_up(off + sizeof(int) * T, 8);
    out->off_blur_bins = off; off += sizeof(double) * nb;
    out->off_sharpness = off; off += sizeof(double) * (size_t)max_boxes;
    out->record_bytes = align_up(off, 16);
    out->T = (int)T; out->na = p->angle_partitions; out->nr = p->radius_partitions; out->max_boxes = max_boxes;
    return PHD_OK;
}

int phd_get_reports_u8(phd_context* ctx, const uint8_t* rgb, int n_images, int width, int height,
                       size_t image_stride, const int* boxes, int max_boxes, const phd_params* p, void* records) {
    if (!ctx) return PHD_E_BAD_PARAMS;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->err[0] = 0;
    if (!rgb || !records || n_images <= 0) return fail(ctx, PHD_E_BAD_PARAMS, "Error: Image pointer is NULL.");
    int rc = check_params(ctx, p, max_boxes);
    if (rc != PHD_OK) return rc;
    if (max_boxes > 0 && !boxes) return fail(ctx, PHD_E_BAD_PARAMS, "max_boxes > 0 but boxes is NULL");
    if (max_boxes > 65535) return fail(ctx, PHD_E_BAD_PARAMS, "more than 65535 boxes per image");
    // crop_pgm (src/image_processing.c:215-219) refuses boxes outside the image (the reference then dereferences NULL);
    // inverted boxes make it allocate a negative size.  Empty boxes (top == bottom) are served like the reference: NaN.
    for (long long i = 0; i < (long long)n_images * max_boxes; i++) {
        const int t = boxes[4 * i], b = boxes[4 * i + 1], l = boxes[4 * i + 2], r = boxes[4 * i + 3];
        if (t < 0 || l < 0 || b > height || r > width || t > b || l > r)
            return fail(ctx, PHD_E_BAD_PARAMS, "Error: crop boundaries outside of image boundaries.");
    }
    char why[256];
    if (reference_rejects(width, height, why, sizeof(why))) return fail(ctx, PHD_E_REJECTED, why);
    if (image_stride < (size_t)width * height * 3) return fail(ctx, PHD_E_BAD_PARAMS, "image_stride smaller than one image");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    phd_flat_layout lay;
    phd_flat_get_layout(p, max_boxes, &lay);
    const bool in_dev = is_device_pointer(rgb);
    const bool out_dev = is_device_pointer(records);
    unsigned char* rec_dev = (unsigned char*)records;
    if (!out_dev) {
        if ((rc = ensure_bytes(ctx, &ctx->d_records, &ctx->d_records_bytes, lay.record_bytes * (size_t)n_images)) != PHD_OK)
            return rc;
        rec_dev = ctx->d_records;
    }
    rc = run_pipeline(ctx, rgb, in_dev, n_images, width, height, image_stride, boxes, max_boxes, *p, rec_dev, lay);
    if (rc != PHD_OK) {
        // nothing of this context may still read the caller's buffers once the error is returned
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        for (int t = 0; t < phd_context::kUpThreads; t++)
            if (ctx->up_stream[t]) cudaStreamSynchronize(ctx->up_stream[t]);
        cudaGetLastError();
        return rc;
    }
    if (!out_dev)
        CUDA_TRY(ctx, cudaMemcpyAsync(records, rec_dev, lay.record_bytes * (size_t)n_images, cudaMemcpyDeviceToHost,
                                      ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    return collect_timing(ctx);
}

// images are independent, so a batch shards over the GPUs of one box with no exchange step.
// One host thread per context (device); thread g takes the contiguous range [g*B/G, (g+1)*B/G) of the batch and its
// device writes the records straight into that range of the caller's array -- the gather is the join.
int phd_get_reports_u8_multi(phd_context* const* ctxs, int n_ctx, const uint8_t* rgb, int n_images, int width, int height,
                             size_t image_stride, const int* boxes, int max_boxes, const phd_params* p, void* records) {
    if (!ctxs || n_ctx <= 0 || !rgb || !records || n_images <= 0 || !p) return PHD_E_BAD_PARAMS;
    for (int g = 0; g < n_ctx; g++) {
        if (!ctxs[g]) return PHD_E_BAD_PARAMS;
        for (int o = 0; o < g; o++)
            if (ctxs[o] == ctxs[g]) return PHD_E_BAD_PARAMS;  // one range per context
    }
    if (is_device_pointer(rgb) || is_device_pointer(records)) {
        fprintf(stderr, "photohive_dsp: phd_get_reports_u8_multi takes HOST buffers (a device buffer belongs to one GPU)\n");
        return PHD_E_BAD_PARAMS;
    }
    phd_flat_layout lay;
    if (phd_flat_get_layout(p, max_boxes, &lay) != PHD_OK) return PHD_E_BAD_PARAMS;
    std::vector<int> rcs(n_ctx, PHD_OK);
    auto work = [&](int g) {
        const long long lo = (long long)n_images * g / n_ctx, hi = (long long)n_images * (g + 1) / n_ctx;
        if (hi <= lo) return;
        rcs[g] = phd_get_reports_u8(ctxs[g], rgb + (size_t)lo * image_stride, (int)(hi - lo), width, height, image_stride,
                                    boxes ? boxes + (size_t)lo * max_boxes * 4 : nullptr, max_boxes, p,
                                    (unsigned char*)records + (size_t)lo * lay.record_bytes);
    };
    std::vector<std::thread> pool;
    for (int g = 1; g < n_ctx; g++) pool.emplace_back(work, g);
    work(0);
    for (auto& t : pool) t.join();
    for (int g = 0; g < n_ctx; g++)
        if (rcs[g] != PHD_OK) return rcs[g];
    return PHD_OK;
}

int phd_last_timing(const phd_context* ctx, float ms[8]) {
    if (!ctx) return 0;
    for (int i = 0; i < 8; i++) ms[i] = ctx->last_ms[i];
    return ctx->last_launches;
}

int phd_last_fused(const phd_context* ctx) { return ctx ? ctx->last_fused : 0; }

int phd_last_stage_launches(const phd_context* ctx, int n[8]) {
    if (!ctx) return PHD_E_BAD_PARAMS;
    for (int i = 0; i < 8; i++) n[i] = ctx->last_stage_launches[i];
    return PHD_OK;
}

Full_Report_Data* phd_flat_to_full_report(const void* record, const phd_flat_layout* lay) {
    if (!record || !lay) return NULL;
    const unsigned char* rec = (const unsigned char*)record;
    const phd_flat_head* h = (const phd_flat_head*)rec;
    if (h->status != 0) return NULL;
    const double* hsv = (const double*)(rec + lay->off_palette_hsv);
    const double* pct = (const double*)(rec + lay->off_palette_pct);
    const int* pid = (const int*)(rec + lay->off_parent_ids);
    const double* bins = (const double*)(rec + lay->off_blur_bins);
    const double* sharp = (const double*)(rec + lay->off_sharpness);

    // every block is checked; on exhaustion whatever was built is released through the same path the caller would use
    Full_Report_Data* r = (Full_Report_Data*)calloc(1, sizeof(Full_Report_Data));
    if (!r) return NULL;
    bool ok = true;
    r->rgb_stats = (RGB_Statistics*)calloc(1, sizeof(RGB_Statistics));
    if (r->rgb_stats) {
        r->rgb_stats->Br = h->rgb_stats[0]; r->rgb_stats->Bg = h->rgb_stats[1]; r->rgb_stats->Bb = h->rgb_stats[2];
        r->rgb_stats->Cr = h->rgb_stats[3]; r->rgb_stats->Cg = h->rgb_stats[4]; r->rgb_stats->Cb = h->rgb_stats[5];
    } else ok = false;
    r->average_saturation = h->average_saturation;

    Color_Palette* cp = (Color_Palette*)calloc(1, sizeof(Color_Palette));
    r->color_palette = cp;
    if (cp) {
        cp->N = h->palette_n;
        cp->averages = (Pixel_HSV*)calloc(cp->N > 0 ? cp->N : 1, sizeof(Pixel_HSV));
        cp->percentages = (Pixel*)calloc(cp->N > 0 ? cp->N : 1, sizeof(Pixel));
        if (cp->averages && cp->percentages) {
            for (int i = 0; i < cp->N; i++) {
                cp->averages[i].parent_id = pid[i];
                cp->averages[i].h = hsv[3 * i]; cp->averages[i].s = hsv[3 * i + 1]; cp->averages[i].v = hsv[3 * i + 2];
                cp->percentages[i] = pct[i];
            }
        } else ok = false;
    } else ok = false;

    Blur_Profile* bp = (Blur_Profile*)calloc(1, sizeof(Blur_Profile));
    r->blur_profile = bp;
    if (bp) {
        bp->num_radius_bins = h->num_radius_bins;
        bp->angle_bin_size = h->angle_bin_size; bp->radius_bin_size = h->radius_bin_size;
        bp->bins = (Bin**)calloc(h->num_angle_bins > 0 ? h->num_angle_bins : 1, sizeof(Bin*));
        if (bp->bins) {
            bp->num_angle_bins = h->num_angle_bins;  // free_full_report walks this many rows (NULL rows are fine for free)
            for (int a = 0; a < bp->num_angle_bins; a++) {
                bp->bins[a] = (Bin*)malloc(sizeof(Bin) * (bp->num_radius_bins > 0 ? bp->num_radius_bins : 1));
                if (!bp->bins[a]) { ok = false; continue; }
                memcpy(bp->bins[a], bins + (size_t)a * bp->num_radius_bins, sizeof(Bin) * bp->num_radius_bins);
            }
        } else ok = false;
    } else ok = false;

    Blur_Vector_Group* bv = (Blur_Vector_Group*)calloc(1, sizeof(Blur_Vector_Group));
    r->blur_vectors = bv;
    if (bv) {
        bv->blur_vectors = (Blur_Vector*)calloc(10, sizeof(Blur_Vector));
        if (bv->blur_vectors) {
            bv->len_vectors = 10;
            for (int k = 0; k < 10; k++) { bv->blur_vectors[k].angle = h->blur_vec_angle[k]; bv->blur_vectors[k].magnitude = h->blur_vec_mag[k]; }
        } else ok = false;
    } else ok = false;

    r->sharpness = NULL;
    if (h->n_sharpness >= 0) {
        Sharpnesses* s = (Sharpnesses*)calloc(1, sizeof(Sharpnesses));
        r->sharpness = s;
        if (s) {
            s->sharpness = (Pixel*)calloc(h->n_sharpness > 0 ? h->n_sharpness : 1, sizeof(Pixel));
            if (s->sharpness) {
                s->N = h->n_sharpness;
                for (int i = 0; i < s->N; i++) s->sharpness[i] = sharp[i];
            } else ok = false;
        } else ok = false;
    }
    if (!ok) {
        fprintf(stderr, "photohive_dsp: out of host memory while building the report\n");
        free_full_report(&r);
        return NULL;
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------------------------
static int group_sweep(phd_context* ctx, const phd_params* p, uint16_t* out, bool fast) {
    if (!ctx || !out) return PHD_E_BAD_PARAMS;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_params(ctx, p, 0);
    if (rc != PHD_OK) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevParams P;
    fill_dev_params(P, *p, 1024, 1024, 0, 0, 0);
    ParamTables* tab;
    if ((rc = get_tables(ctx, *p, &tab)) != PHD_OK) return rc;
    u16* d;
    CUDA_TRY(ctx, cudaMalloc(&d, sizeof(u16) << 24));
    phd_launch_group_sweep(P, tab->tabs, tab->exc, fast, d, ctx->stream);
    cudaError_t e = cudaMemcpyAsync(out, d, sizeof(u16) << 24, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    CUDA_TRY(ctx, e);
    return PHD_OK;
}

int phd_debug_group_sweep(phd_context* ctx, const phd_params* p, uint16_t* out) { return group_sweep(ctx, p, out, true); }
int phd_debug_group_sweep_exact(phd_context* ctx, const phd_params* p, uint16_t* out) { return group_sweep(ctx, p, out, false); }

int phd_debug_bin_map(phd_context* ctx, int width, int height, int nr, int na, uint16_t* map, int* counts) {
    if (!ctx) return PHD_E_BAD_PARAMS;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ShapePlan* s;
    int rc = get_shape(ctx, width, height, nr, na, &s);
    if (rc != PHD_OK) return rc;
    if (map) {
        // the device map is stored transposed ([x][k], pitch Hp); hand it back row major
        const int fw = width / 2 + 1, Hp = (height + 3) / 4 * 4;
        std::vector<u16> t((size_t)fw * Hp);
        CUDA_TRY(ctx, cudaMemcpy(t.data(), s->binmap, sizeof(u16) * t.size(), cudaMemcpyDeviceToHost));
        for (int k = 0; k < height; k++)
            for (int x = 0; x < fw; x++) map[(size_t)k * fw + x] = t[(size_t)x * Hp + k];
    }
    if (counts) CUDA_TRY(ctx, cudaMemcpy(counts, s->bincount, sizeof(int) * nr * na, cudaMemcpyDeviceToHost));
    return PHD_OK;
}

int phd_debug_power_spectrum(phd_context* ctx, const uint8_t* rgb, int width, int height, float* power) {
    if (!ctx || !rgb || !power) return PHD_E_BAD_PARAMS;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    phd_params p;
    phd_default_params(&p);
    DevParams P;
    const size_t tight = (size_t)width * height * 3, stride = align_up(tight, 16);
    fill_dev_params(P, p, width, height, 0, stride, 1);
    ShapePlan* s;
    int rc = get_shape(ctx, width, height, P.nr, P.na, &s);
    if (rc != PHD_OK) return rc;
    if ((rc = ensure_workspace(ctx, P, 1, 1)) != PHD_OK) return rc;
    if ((rc = ensure_bytes(ctx, &ctx->d_rgb, &ctx->d_rgb_bytes, stride)) != PHD_OK) return rc;
    const size_t nspec = (size_t)P.fw * height;
    float* d_pow;
    CUDA_TRY(ctx, cudaMalloc(&d_pow, sizeof(float) * nspec));
    int launches = 0;
    cudaError_t e = cudaMemcpyAsync(ctx->d_rgb, rgb, tight, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        if (launch_rows_any(ctx, *s, ctx->d_rgb, stride, nullptr, P, 1, ctx->ws.spec, ctx->stream, &launches) ||
            launch_cols_any(ctx, *s, P, 1, ctx->ws.spec, ctx->ws, d_pow, ctx->stream, &launches)) {
            cudaFree(d_pow);
            return fail(ctx, PHD_E_UNSUPPORTED, "FFT does not fit shared memory");
        }
        e = cudaMemcpyAsync(power, d_pow, sizeof(float) * nspec, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d_pow);
    CUDA_TRY(ctx, e);
    return PHD_OK;
}

int phd_debug_group_counts(phd_context* ctx, const uint8_t* rgb, int width, int height, const phd_params* p, int* counts) {
    if (!ctx || !rgb || !counts) return PHD_E_BAD_PARAMS;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_params(ctx, p, 0);
    if (rc != PHD_OK) return rc;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    DevParams P;
    const size_t tight = (size_t)width * height * 3, stride = align_up(tight, 16);
    fill_dev_params(P, *p, width, height, 0, stride, 1);
    ParamTables* tab;
    if ((rc = get_tables(ctx, *p, &tab)) != PHD_OK) return rc;
    if ((rc = ensure_workspace(ctx, P, 1, 1)) != PHD_OK) return rc;
    if ((rc = ensure_bytes(ctx, &ctx->d_rgb, &ctx->d_rgb_bytes, stride)) != PHD_OK) return rc;
    int launches = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_rgb, rgb, tight, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->ws_zero, 0, ctx->ws_zero_bytes, ctx->stream));
    phd_fe_plan(P, 1);
    phd_launch_pixels(ctx->d_rgb, P, 1, tab->tabs, tab->exc, ctx->ws, ctx->stream, &launches);
    phd_launch_palette_select(P, 1, tab->centres, tab->sv_f, ctx->ws, ctx->stream, &launches);
    CUDA_TRY(ctx, cudaMemcpyAsync(counts, ctx->ws.hist, sizeof(int) * P.T, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PHD_OK;
}

// =============================================================================================
// Part 1: drop-in entry points
// =============================================================================================
Full_Report_Data* get_full_report_data(Image_RGB* image, Crop_Boundaries* crop, int h_partitions, int s_partitions,
                                       int v_partitions, double black_thresh, double gray_thresh,
                                       double coverage_thresh, int linked_list_size, int downsample_rate,
                                       int radius_partitions, int angle_partitions, float quantity_weight,
                                       float saturation_value_weight, double fft_streak_thresh,
                                       double magnitude_thresh, int blur_cutoff_ratio_denom) {
    // pre_compute_error_checks, src/utilities.c:64-87, same order and messages
    if (image == NULL) {
        fprintf(stderr, "Error: Image pointer is NULL.\n");
        return NULL;
    }
    char why[256];
    if (reference_rejects(image->width, image->height, why, sizeof(why))) {
        fprintf(stderr, "%s\n", why);
        return NULL;
    }
    if (image->r == NULL || image->g == NULL || image->b == NULL) {
        fprintf(stderr, "Error: At least one color channel was a NULL pointer.\n");
        return NULL;
    }
    phd_params p;
    p.h_partitions = h_partitions; p.s_partitions = s_partitions; p.v_partitions = v_partitions;
    p.black_thresh = black_thresh; p.gray_thresh = gray_thresh; p.coverage_thresh = coverage_thresh;
    p.linked_list_size = linked_list_size; p.downsample_rate = downsample_rate;
    p.radius_partitions = radius_partitions; p.angle_partitions = angle_partitions;
    p.quantity_weight = quantity_weight; p.saturation_value_weight = saturation_value_weight;
    p.fft_streak_thresh = fft_streak_thresh; p.magnitude_thresh = magnitude_thresh;
    p.blur_cutoff_ratio_denom = blur_cutoff_ratio_denom;

    phd_context* ctx = default_context();
    if (!ctx) return NULL;
    const int W = image->width, H = image->height;
    const long long npx = (long long)W * H;
    const int nb = crop ? crop->N : 0;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);  // check_params writes the context's error buffer
        if (check_params(ctx, &p, nb) != PHD_OK) return NULL;
    }
    std::vector<int> boxes((size_t)(nb > 0 ? nb : 1) * 4);
    for (int i = 0; i < nb; i++) {
        boxes[4 * i] = crop->top[i]; boxes[4 * i + 1] = crop->bottom[i];
        boxes[4 * i + 2] = crop->left[i]; boxes[4 * i + 3] = crop->right[i];
        // crop_pgm (src/image_processing.c:215-219) refuses these and the reference then dereferences NULL
        if (crop->right[i] > W || crop->left[i] > W || crop->bottom[i] > H || crop->top[i] > H || crop->left[i] < 0 ||
            crop->right[i] < 0 || crop->top[i] < 0 || crop->bottom[i] < 0) {
            fprintf(stderr, "Error: crop boundaries outside of image boundaries.\n");
            return NULL;
        }
    }

    // planes -> packed 8-bit on the device.  The planes are pageable host memory (24 B per pixel): threaded_upload;
    // all buffers are kept in the context between calls.
    int flag = 0;
    Full_Report_Data* result = NULL;
    const size_t stride = align_up((size_t)npx * 3, 16);
    std::lock_guard<std::mutex> dropin_lk(ctx->dropin_mu);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        cudaSetDevice(ctx->device);
        cudaError_t e = cudaSuccess;
        const size_t plane_bytes = sizeof(double) * (size_t)npx;
        if (ctx->d_planes_bytes < 3 * plane_bytes) {
            cudaFree(ctx->d_planes);
            ctx->d_planes = nullptr; ctx->d_planes_bytes = 0;
            e = cudaMalloc(&ctx->d_planes, 3 * plane_bytes);
            if (e == cudaSuccess) ctx->d_planes_bytes = 3 * plane_bytes;
        }
        if (e == cudaSuccess && ctx->d_u8_bytes < stride) {
            cudaFree(ctx->d_u8);
            ctx->d_u8 = nullptr; ctx->d_u8_bytes = 0;
            e = cudaMalloc(&ctx->d_u8, stride);
            if (e == cudaSuccess) ctx->d_u8_bytes = stride;
        }
        if (e == cudaSuccess && !ctx->d_flag) e = cudaMalloc(&ctx->d_flag, sizeof(int));
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream);
        if (e == cudaSuccess) {
            const double* src[3] = {image->r, image->g, image->b};
            UploadSeg segs[3];
            for (int c = 0; c < 3; c++)
                segs[c] = {reinterpret_cast<const unsigned char*>(src[c]),
                           reinterpret_cast<unsigned char*>(ctx->d_planes + (size_t)c * npx), plane_bytes};
            e = threaded_upload(ctx, segs, 3);
        }
        if (e == cudaSuccess) {
            k_ingest_f64<<<(unsigned)((npx + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_planes, ctx->d_planes + npx,
                                                                                  ctx->d_planes + 2 * npx, npx, ctx->d_u8,
                                                                                  ctx->d_flag);
            e = cudaMemcpyAsync(&flag, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            fprintf(stderr, "photohive_dsp: CUDA error while uploading the image: %s\n", cudaGetErrorString(e));
            return NULL;
        }
    }
    phd_flat_layout lay;
    phd_flat_get_layout(&p, nb, &lay);
    std::vector<unsigned char> rec(lay.record_bytes);
    int rc;
    if (flag) {
        // values that are not k/255: the general-input route on the planes already on the device (f64path.cu)
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->err[0] = 0;
        rc = ensure_bytes(ctx, &ctx->d_records, &ctx->d_records_bytes, lay.record_bytes);
        if (rc == PHD_OK) rc = run_pipeline_f64(ctx, ctx->d_planes, W, H, nb > 0 ? boxes.data() : NULL, nb, p, ctx->d_records, lay);
        if (rc == PHD_OK && cudaMemcpyAsync(rec.data(), ctx->d_records, lay.record_bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
            rc = PHD_E_CUDA;
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess && rc == PHD_OK) rc = PHD_E_CUDA;
    } else {
        rc = phd_get_reports_u8(ctx, ctx->d_u8, 1, W, H, stride, nb > 0 ? boxes.data() : NULL, nb, &p, rec.data());
    }
    if (rc != PHD_OK) return NULL;
    result = phd_flat_to_full_report(rec.data(), &lay);
    if (result && !crop && result->sharpness) {  // unreachable (n_sharpness = -1 when nb == 0), kept for clarity
        free(result->sharpness->sharpness);
        free(result->sharpness);
        result->sharpness = NULL;
    }
    // A non-NULL Crop_Boundaries with N == 0 still yields a Sharpnesses object in the reference (filtering.c:158-160).
    if (result && crop && !result->sharpness) {
        result->sharpness = (Sharpnesses*)malloc(sizeof(Sharpnesses));
        result->sharpness->N = 0;
        result->sharpness->sharpness = (Pixel*)calloc(1, sizeof(Pixel));
    }
    return result;
}

void free_full_report(Full_Report_Data** report) {
    if (!report || !*report) return;
    Full_Report_Data* r = *report;
    if (r->color_palette) {
        free(r->color_palette->averages);
        free(r->color_palette->percentages);
        free(r->color_palette);
        r->color_palette = NULL;
    }
    if (r->blur_profile) {
        for (int a = 0; a < r->blur_profile->num_angle_bins; a++) free(r->blur_profile->bins[a]);
        free(r->blur_profile->bins);
        free(r->blur_profile);
        r->blur_profile = NULL;
    }
    if (r->blur_vectors) {
        free(r->blur_vectors->blur_vectors);
        free(r->blur_vectors);
        r->blur_vectors = NULL;
    }
    if (r->sharpness) {
        free(r->sharpness->sharpness);
        free(r->sharpness);
        r->sharpness = NULL;
    }
    free(r->rgb_stats);
    r->rgb_stats = NULL;
    free(r);
    *report = NULL;
}

// Host-side visualiser, src/blur_profile.c:140-180.  Not on the hot path.
Image_PGM* get_blur_profile_visual(Blur_Profile* bp, int height, int width) {
    if (!bp || height <= 0 || width <= 0) return NULL;
    Image_PGM* out = (Image_PGM*)malloc(sizeof(Image_PGM));
    if (!out) {
        fprintf(stderr, "Error creating output image.\n");
        return NULL;
    }
    out->height = height;
    out->width = width;
    out->data = (Pixel*)calloc((size_t)height * width, sizeof(Pixel));
    const double REF_PI = 3.14159265;
    for (int y = 0; y < height; y++) {
        const double dy = (y < height / 2) ? -(double)y : (double)(height - y);
        for (int x = 0; x < width; x++) {
            const double dx = x;
            const double r = sqrt(dx * dx + dy * dy);
            const double phi = atan2(dy, dx);
            int r_bin = (int)(r / bp->radius_bin_size);
            if (r_bin >= bp->num_radius_bins) r_bin = bp->num_radius_bins - 1;
            int phi_bin = (int)((phi + REF_PI * 0.5f) / REF_PI * (double)(bp->num_angle_bins - 1));
            if (phi_bin >= bp->num_angle_bins) phi_bin = bp->num_angle_bins - 1;
            if (phi_bin < 0) phi_bin = 0;
            out->data[(size_t)y * width + x] = bp->bins[phi_bin][r_bin];
        }
    }
    return out;
}

}  // extern "C"
