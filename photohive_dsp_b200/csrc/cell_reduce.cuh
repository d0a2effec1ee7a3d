// Palette cells -> the sums calculate_avg_hsv (src/color_quantization.c:510-576) takes over the pixels of ONE group
// once that group is known to end up under a parent with centre hue `parent_h`.
//
// For a pixel with hue h the reference adds t = h + off (off = 180 - parent_h), minus 360 if t > 360, plus 360 if
// t < 0.  Within a half-bin cell the decision is the same for every pixel (seams are half-bin boundaries), and the
// pixels that sit exactly ON a boundary were sorted by k_build_exc onto the side the reference's doubles put them:
//     sub 1 / 3  interior of the lower / upper half bin: decided at the cell's mid-point
//     sub 0      on the bin's lower edge b, low side :  t = b + off, a result <= 0 wraps up   (range (0, 360])
//     sub 2      on the bin's upper edge b, high side:  t = b + off, a result >= 360 wraps down (range [0, 360))
#pragma once

#include "phd_internal.h"

struct GroupSums {
    u64 cnt, n255, summax, s_sum;
    long long t_sum;  // * 2^PHD_T_SHIFT
};

__device__ __forceinline__ void phd_reduce_pair(const u64* __restrict__ cells, int NC, int pair, int j, double Lh,
                                                double off, GroupSums& S) {
    const double half = Lh * 0.5;
#pragma unroll
    for (int sub = 0; sub < 4; sub++) {
        const int cell = pair * 4 + sub;
        const u64 cnt = cells[cell];
        if (!cnt) continue;
        const double n = (double)cnt;
        double tsum;
        if (sub == 1 || sub == 3) {
            const double lo = (double)j * Lh + (sub == 3 ? half : 0.0);
            const double hsum = n * lo + (double)cells[4 * (size_t)NC + cell] * half * (1.0 / (double)(1 << PHD_S_SHIFT));
            const double tm = lo + half * 0.5 + off;
            const double w = tm > 360.0 ? -360.0 : (tm < 0.0 ? 360.0 : 0.0);
            tsum = hsum + n * (off + w);
        } else if (sub == 0) {
            double t = (double)j * Lh + off;
            if (t <= 0.0) t += 360.0;
            else if (t > 360.0) t -= 360.0;
            tsum = n * t;
        } else {
            double t = (double)(j + 1) * Lh + off;
            if (t < 0.0) t += 360.0;
            else if (t >= 360.0) t -= 360.0;
            tsum = n * t;
        }
        S.cnt += cnt;
        S.n255 += cells[(size_t)NC + cell];
        S.summax += cells[2 * (size_t)NC + cell];
        S.s_sum += cells[3 * (size_t)NC + cell];
        S.t_sum += __double2ll_rn(tsum * (double)(1 << PHD_T_SHIFT));
    }
}

// All cells of reference group g (cells: the image's [PHD_CELL_Q][NC] block).
__device__ __forceinline__ GroupSums phd_reduce_group(const u64* __restrict__ cells, const DevParams& P, int g,
                                                      double parent_h) {
    GroupSums S{0, 0, 0, 0, 0};
    const int spvp = P.sp * P.vp;
    const double off = 180.0 - parent_h;
    if (g < P.hp * spvp) {
        const int j = g / spvp, cls = g - j * spvp;
        phd_reduce_pair(cells, P.NC, cls * P.hp + j, j, P.Lh, off, S);
    } else if (g == P.T - (P.vp + 1) || g == P.T - 1) {
        const int cls = (g == P.T - 1) ? spvp + 1 : spvp;
        for (int j = 0; j < P.hp; j++) phd_reduce_pair(cells, P.NC, cls * P.hp + j, j, P.Lh, off, S);
    }
    return S;
}

// First cell and cell count of group g (for zeroing / iterating).
__device__ __forceinline__ void phd_group_cell_range(const DevParams& P, int g, int* first, int* count) {
    const int spvp = P.sp * P.vp;
    if (g < P.hp * spvp) {
        const int j = g / spvp, cls = g - j * spvp;
        *first = (cls * P.hp + j) * 4;
        *count = 4;
    } else if (g == P.T - (P.vp + 1) || g == P.T - 1) {
        const int cls = (g == P.T - 1) ? spvp + 1 : spvp;
        *first = cls * P.hp * 4;
        *count = P.hp * 4;
    } else {
        *first = 0;
        *count = 0;
    }
}
