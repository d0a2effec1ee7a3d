// Per-image epilogue: turns the integer accumulators into one flat report record.
//
// Replaces the closing arithmetic of get_rgb_statistics / get_hsv_average
// (src/image_processing.c:533-553, src/filtering.c:125-147), calculate_avg_hsv
// (src/color_quantization.c:557-573), the averaging loop of calculate_blur_profile
// (src/blur_profile.c:56-58,106-116) with pgm_normalize_fft's G_s (src/fft_processing.c:192),
// vectorize_blur_profile (src/blur_profile.c:324-416, convolve_1d src/filtering.c:12-24) and the
// variance/mean of get_variance_sharpness (src/filtering.c:170-174).  One CTA per image.
#include <math.h>

#include "cell_reduce.cuh"

namespace {

__device__ __forceinline__ double u128_to_double(unsigned __int128 v) {
    return (double)(u64)(v >> 64) * 18446744073709551616.0 + (double)(u64)v;
}
__device__ __forceinline__ double i128_to_double(__int128 v) {
    return v < 0 ? -u128_to_double((unsigned __int128)(-v)) : u128_to_double((unsigned __int128)v);
}

__global__ void __launch_bounds__(256) k_finalize(DevParams P, const double* __restrict__ centres,
                                                  const int* __restrict__ bincount,
                                                  const ImageAcc* __restrict__ iacc, const int* __restrict__ pal_n,
                                                  const int* __restrict__ parent_ids,
                                                  SlotAcc* sacc, const GroupPlan* __restrict__ plan_g,
                                                  const int* __restrict__ tie_list, const int* __restrict__ tie_n,
                                                  const u64* __restrict__ cells_tie_g,
                                                  const u64* __restrict__ binsum,
                                                  const u32* __restrict__ maxpow, const SharpAcc* __restrict__ sharp,
                                                  const int* __restrict__ boxes, const int* __restrict__ tie_groups,
                                                  const long long* __restrict__ dropped, phd_flat_layout lay,
                                                  unsigned char* __restrict__ records,
                                                  const double* __restrict__ f64_acc, const double* __restrict__ f64_slots,
                                                  const double* __restrict__ f64_sharp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* bins = reinterpret_cast<double*>(smem_raw);  // [nbins]
    const int img = blockIdx.x, tid = threadIdx.x, T = P.T;
    unsigned char* rec = records + (size_t)img * lay.record_bytes;
    phd_flat_head* head = reinterpret_cast<phd_flat_head*>(rec);
    double* out_hsv = reinterpret_cast<double*>(rec + lay.off_palette_hsv);
    double* out_pct = reinterpret_cast<double*>(rec + lay.off_palette_pct);
    int* out_pid = reinterpret_cast<int*>(rec + lay.off_parent_ids);
    double* out_bins = reinterpret_cast<double*>(rec + lay.off_blur_bins);
    double* out_sharp = reinterpret_cast<double*>(rec + lay.off_sharpness);

    const ImageAcc a = iacc[img];
    const double np = (double)P.npx;
    const int N = pal_n[img];
    // records carry padding bytes nothing below writes: clear the record here so that equal images give equal bytes
    // (this used to be a separate memset per call)
    for (size_t i = tid; i < lay.record_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(rec)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // --- partly accepted tie groups: fold the pixels k_palette_ties accepted into their parent's sums ---
    {
        const int nt = f64_slots ? 0 : tie_n[img];  // the general-input route has already added its tie pixels
        const u64* ct = cells_tie_g + (size_t)img * PHD_CELL_Q * P.NC;
        for (int k = tid; k < nt; k += blockDim.x) {
            const int g = tie_list[(size_t)img * T + k];
            const GroupPlan gp = plan_g[(size_t)img * T + g];
            const GroupSums S = phd_reduce_group(ct, P, g, centres[parent_ids[(size_t)img * T + gp.slot]]);
            SlotAcc* A = sacc + (size_t)img * T + gp.slot;
            if (S.summax) atomicAdd(&A->summax, S.summax);
            if (S.n255) atomicAdd(&A->n255, S.n255);
            if (S.s_sum) atomicAdd(&A->s_sum, S.s_sum);
            if (S.t_sum) atomicAdd(&A->t_sum, (u64)S.t_sum);
        }
        __threadfence();
        __syncthreads();
    }

    // --- palette averages ---
    const double inv_total = 1.0 / (double)P.hpx;
    for (int j = tid; j < T; j += blockDim.x) {
        double h = 0, s = 0, v = 0, pct = 0;
        int pid = -1;
        if (j < N) {
            pid = parent_ids[(size_t)img * T + j];
            SlotAcc A;  // read past L1: other threads of this CTA have just added to it with atomics
            {
                const u64* q = reinterpret_cast<const u64*>(sacc + (size_t)img * T + j);
                A.cnt = __ldcg(q); A.summax = __ldcg(q + 1); A.n255 = __ldcg(q + 2); A.s_sum = __ldcg(q + 3); A.t_sum = __ldcg(q + 4);
            }
            const double off = 180.0 - centres[pid];
            const double inv = 1.0 / (double)A.cnt;
            if (f64_slots) {  // general-input route: FP64 sums of v, s and the wrapped hue
                h = f64_slots[4 * j + 2] * inv - off;
                s = f64_slots[4 * j + 1] * inv;
                v = f64_slots[4 * j] * inv;
            } else {
                h = (double)A.t_sum * (1.0 / (double)(1 << PHD_T_SHIFT)) * inv - off;
                s = (double)A.s_sum * (1.0 / (double)(1 << PHD_S_SHIFT)) * inv;
                v = ((double)(A.summax - 255ull * A.n255) / 255.0 + (double)A.n255 * 0.999999) * inv;
            }
            if (h < 0) h += 360;
            else if (h > 360) h -= 360;
            pct = (double)A.cnt * inv_total;
        }
        out_hsv[3 * j] = h; out_hsv[3 * j + 1] = s; out_hsv[3 * j + 2] = v;
        out_pct[j] = pct;
        out_pid[j] = pid;
    }

    // --- blur bins: mean of ln(p) per bin, scaled by G_s ---
    double maxp = (double)__uint_as_float(maxpow[img]);
    const double Gs = 1.0 / (2.0 * log(sqrt(maxp) + 1.0));
    for (int b = tid; b < P.nbins; b += blockDim.x) {
        const int c = bincount[b];
        double val = 0.0;
        const u64 q = binsum[(size_t)img * P.nbins + b];
        if (c != 0 && q != 0) val = ((double)q * (1.0 / (double)(1 << PHD_LN_SHIFT)) / (double)c) * Gs;
        bins[b] = val;
        out_bins[b] = val;
    }

    // --- sharpness ---
    for (int k = tid; k < P.max_boxes; k += blockDim.x) {
        const int* bx = boxes + ((size_t)img * P.max_boxes + k) * 4;
        const int w = bx[3] - bx[2], h = bx[1] - bx[0];
        double r = nan("");
        if (f64_sharp && w > 0 && h > 0) {
            const double dn = (double)w * (double)h;
            const double avg = f64_sharp[2 * k] / dn;
            r = (f64_sharp[2 * k + 1] / dn - avg * avg) / avg;
        } else if (w > 0 && h > 0 && bx[2] >= 0 && bx[0] >= 0 && bx[3] <= P.W && bx[1] <= P.H) {
            const SharpAcc S = sharp[(size_t)img * P.max_boxes + k];
            const long long n = (long long)w * h;
            // n*S2 - S1^2 exactly, then var = that / (n^2 * 255000^2), avg = S1 / (n * 255000)
            const unsigned __int128 s2 = ((unsigned __int128)S.s2hi << 32) + S.s2lo;
            const __int128 num = (__int128)(s2 * (unsigned __int128)n) - (__int128)S.s1 * (__int128)S.s1;
            const double dn = (double)n;
            const double var = i128_to_double(num) / (dn * dn * 255000.0 * 255000.0);
            const double avg = (double)S.s1 / (dn * 255000.0);
            r = var / avg;
        }
        out_sharp[k] = r;
    }
    __syncthreads();
    // per-angle sums of the first nr / denom radius bins (vectorize_blur_profile, src/blur_profile.c:340-350)
    for (int i = tid; i < P.na; i += blockDim.x) {
        double t = 0;
        for (int j = 0; j < P.nr / P.denom; j++) t += bins[i * P.nr + j];
        bins[P.nbins + i] = t;
    }
    __syncthreads();

    if (tid == 0) {
        for (int c = 0; c < 3 && f64_acc; c++) {
            const double mean = f64_acc[c] / np;
            head->rgb_stats[c] = mean;
            head->rgb_stats[3 + c] = sqrt(fmax(f64_acc[3 + c] / np - mean * mean, 0.0));
        }
        for (int c = 0; c < 3 && !f64_acc; c++) {
            const double mean = (double)a.sum[c] / 255.0 / np;
            const unsigned __int128 num = (unsigned __int128)a.sumsq[c] * (unsigned __int128)P.npx -
                                          (unsigned __int128)a.sum[c] * (unsigned __int128)a.sum[c];
            head->rgb_stats[c] = mean;
            head->rgb_stats[3 + c] = sqrt(u128_to_double(num) / (np * np * 65025.0));
        }
        head->average_saturation = f64_acc ? f64_acc[7] / (double)P.hpx
                                           : (double)a.s_sum * (1.0 / (double)(1 << PHD_S_SHIFT)) / (double)P.hpx;
        head->max_power = maxp;
        head->dropped_pixels = dropped[img];
        head->palette_n = N;
        head->tie_groups = tie_groups[img];
        head->n_sharpness = P.max_boxes > 0 ? P.max_boxes : -1;
        head->angle_bin_size = 180 / P.na;
        head->radius_bin_size = (int)(sqrt((double)(P.fw * P.fw + P.H * P.H / 4)) / (double)P.nr);
        head->num_angle_bins = P.na;
        head->num_radius_bins = P.nr;
        head->status = 0;

        // --- blur vectors (A.6) ---
        const int na = P.na, nr = P.nr;
        for (int k = 0; k < 10; k++) { head->blur_vec_angle[k] = 0; head->blur_vec_mag[k] = 0.f; }
        const int rc = nr / P.denom;
        // tot[] and smooth[] live after the bins in shared memory
        double* tot = bins + P.nbins;
        double* sm = tot + na;
        double avg = 0;
        for (int i = 0; i < na; i++) avg += tot[i];  // tot[] was summed by na threads above, each in the reference's order
        avg /= na;
        for (int i = 0; i < na; i++) {
            double r = 0;
            for (int j = 0; j < 5; j++) r += tot[(i - j + na) % na] * 1.0;
            sm[i] = r / 5;
        }
        int maxima[10], nmax = 0;
        const double thr = avg * P.streak;
        if (sm[0] > sm[na - 1] && sm[0] > sm[1])
            if (sm[0] > thr && nmax < 10) maxima[nmax++] = 0;
        for (int i = 1; i < na - 1; i++)
            if (sm[i] > sm[i - 1] && sm[i] > sm[i + 1])
                if (sm[i] > thr && nmax < 10) maxima[nmax++] = i;
        if (sm[na - 1] > sm[na - 2] && sm[na - 1] > sm[0])
            if (sm[na - 1] > thr && nmax < 10) maxima[nmax++] = na - 1;
        for (int k = 0; k < nmax; k++) {
            const int idx = (maxima[k] + na / 2) % na;
            const double* sig = bins + (size_t)idx * nr;
            double bavg = 0;
            for (int j = 0; j < rc; j++) bavg += sig[j];
            if (bavg > avg) continue;
            int R = nr;
            for (int j = 0; j < nr; j++)
                if (sig[j] < P.magthr) { R = j; break; }
            head->blur_vec_mag[k] = __fdiv_rn((float)R, (float)nr);
            head->blur_vec_angle[k] = (int)__fsub_rn(__fmul_rn(180.0f, __fdiv_rn((float)idx, (float)na)), 90.0f);
        }
    }
}

}  // namespace

void phd_launch_finalize(const DevParams& P, int nimg, const double* centres, const int* bincount, Workspace& ws,
                         const phd_flat_layout& lay, unsigned char* records_dev, cudaStream_t st, int* launches,
                         const F64Work* f64) {
    const size_t smem = ((size_t)P.nbins + 2 * (size_t)P.na) * sizeof(double);
    PHD_ALLOW_SMEM((k_finalize), 200 * 1024);
    k_finalize<<<nimg, 256, smem, st>>>(P, centres, bincount, ws.iacc, ws.pal_n, ws.parent_ids, ws.sacc, ws.plan,
                                        ws.tie_list, ws.tie_n, ws.cells_tie, ws.binsum,
                                        ws.maxpow, ws.sharp, ws.boxes, ws.tie_groups, ws.dropped, lay, records_dev,
                                        f64 ? f64->acc : nullptr, f64 ? f64->slots : nullptr, f64 ? f64->sharp : nullptr);
    *launches += 1;
}
