// Integer fast path for the palette group of an 8-bit pixel, with an exact FP64 slow path.
//
// The group id is a pure function of (R,G,B) and the palette parameters.  Evaluating the reference's
// double arithmetic per pixel (hsv_exact.cuh) costs ~5 FP64 divisions; here
//   * the value bin / black test is a 256-entry table over max(R,G,B);
//   * the saturation bin / gray test is a per-max threshold list over min(R,G,B) (s = (max-min)/max is
//     monotone in min for fixed max; the table builder checks that monotonicity on all 65,536 pairs);
//   * the hue bin is exact rational arithmetic: h*q = 60*(off*q + p) with q = max-min, p the signed
//     channel difference of the sector; floor(h / Lh) is an integer division (reciprocal table + remainder
//     fix-up).  Rationals that are not exactly on a bin edge are at least 1/(255*Lh) away from it, 10 orders
//     of magnitude more than the reference's rounding error, so the integer result equals the reference's.
//     Pixels exactly on an edge (remainder 0) whose FP value is not trivially exact take the FP64 path of
//     hsv_exact.cuh (out of line: it is rare and large).
// All tables are built ON THE DEVICE with the exact arithmetic (k_build_pal_tables), and the combination is
// verified against the CPU oracle for all 2^24 colours (tests/test_gpu_parity.py).
//
// The fixed-point values used for the SUMS (saturation 2^-30, hue 2^-22) come from 64-bit reciprocal
// tables; they carry ~1e-9 relative error and feed no discrete decision.
#pragma once

#include "hsv_exact.cuh"

#define PHD_HQ_SHIFT PHD_T_SHIFT
#define PHD_HQ_360 (360ll << PHD_HQ_SHIFT)

// Thresholds per max value are stored in rows of SPW = round_up(sp, 4) u16: [0] gray limit (min >= it -> gray),
// [j] = smallest min with Si < j (so Si = #{j >= 1 : min < row[j]}), unused entries 0.
__host__ __device__ inline int phd_spw(int sp) { return (sp + 3) / 4 * 4; }

// Device buffer layout of the per-parameter tables (every section 16-byte aligned).
__host__ __device__ inline size_t phd_pal_tables_bytes(int sp) {
    return 256 /*vtab*/ + 2048 /*rs*/ + 2048 /*rh*/ + 1024 /*rd*/ + 512 * (size_t)phd_spw(sp) /*sthr*/;
}

struct PalTablesView {
    const unsigned char* vtab;  // [256]  value bin, 0xFF = black
    const u64* rs;              // [256]  round(2^46 / max)
    const u64* rh;              // [256]  round(60 * 2^38 / q)
    const u32* rd;              // [256]  ceil(2^32 / (Lh * q)) (q >= 1): reciprocal of the hue-bin width in h*q units
    const u16* sthr;            // [256][SPW]
};

__device__ __forceinline__ PalTablesView phd_pal_tables_view(const unsigned char* base) {
    PalTablesView v;
    v.vtab = base;
    v.rs = reinterpret_cast<const u64*>(base + 256);
    v.rh = reinterpret_cast<const u64*>(base + 256 + 2048);
    v.rd = reinterpret_cast<const u32*>(base + 256 + 4096);
    v.sthr = reinterpret_cast<const u16*>(base + 256 + 4096 + 1024);
    return v;
}

// Copies the tables into shared memory (dst 16-byte aligned).
__device__ __forceinline__ void phd_pal_tables_to_smem(unsigned char* dst, const unsigned char* __restrict__ src, int sp) {
    const int n16 = (int)(phd_pal_tables_bytes(sp) / 16);
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
}

// Hue bin of a pixel through the reference's double arithmetic (rare path, kept out of line).
__device__ __noinline__ int phd_hue_bin_exact(int R, int G, int B, const double* __restrict__ k255, double Lh) {
    const HsvD e = phd_hsv_exact(R, G, B, k255);
    return (int)__ddiv_rn(e.h, Lh);
}

// Wrap decision of calculate_avg_hsv (color_quantization.c:538-548) through the double arithmetic:
// +1 subtract 360, -1 add 360, 0 leave.
__device__ __noinline__ int phd_wrap_exact(int R, int G, int B, const double* __restrict__ k255, double off) {
    const HsvD e = phd_hsv_exact(R, G, B, k255);
    const double t = __dadd_rn(e.h, off);
    return t > 360.0 ? 1 : (t < 0.0 ? -1 : 0);
}

struct FastPx {
    int gid;
    int mx, mn, q;
    int off;  // sector base in units of 60 degrees: 0, 2, 4
    int p;    // signed numerator of the hue fraction p/q
};

struct FastCfg {  // palette constants hoisted out of the pixel loop
    int sp, vp, spw, Lhi, gray_gid, black_gid, T;
    double Lh;
};

__device__ __forceinline__ FastCfg phd_fast_cfg(const DevParams& P) {
    FastCfg c;
    c.sp = P.sp; c.vp = P.vp; c.spw = phd_spw(P.sp); c.Lhi = (int)P.Lh; c.T = P.T;
    c.gray_gid = P.T - (P.vp + 1); c.black_gid = P.T - 1; c.Lh = P.Lh;
    return c;
}

__device__ __forceinline__ FastPx phd_group_fast(int R, int G, int B, const PalTablesView& tb, const FastCfg& C,
                                                 const double* __restrict__ k255) {
    FastPx o;
    const int mx = max(R, max(G, B)), mn = min(R, min(G, B));
    const int q = mx - mn;
    const bool isR = (R == mx), isG = (!isR) && (G == mx);
    const int a = isR ? G : (isG ? B : R);
    const int b = isR ? B : (isG ? R : G);
    const int p = a - b;
    const int off = isR ? 0 : (isG ? 2 : 4);
    o.mx = mx; o.mn = mn; o.q = q; o.p = p; o.off = off;
    const int vi = tb.vtab[mx];
    // saturation class from the threshold row of this max
    const u16* th = tb.sthr + mx * C.spw;
    const uint2 t03 = *reinterpret_cast<const uint2*>(th);
    const int t0 = t03.x & 0xffff, t1 = t03.x >> 16, t2 = t03.y & 0xffff, t3 = t03.y >> 16;
    int si = (mn < t1) + (mn < t2) + (mn < t3);
    for (int j = 4; j < C.sp; j++) si += (mn < (int)th[j]);
    // hue bin: floor(num / den) with num = h*q, den = Lh*q
    int hi = 0;
    if (q != 0) {
        int num = 60 * (off * q + p);
        if (num < 0) num += 360 * q;
        const int den = C.Lhi * q;
        hi = (int)__umulhi((u32)num, tb.rd[q]);
        int rem = num - hi * den;
        if (rem < 0) { hi--; rem += den; }
        else if (rem >= den) { hi++; rem -= den; }
        if (rem == 0 && p != 0 && p != q && p != -q && vi != 0xFF && mn < t0)
            hi = phd_hue_bin_exact(R, G, B, k255, C.Lh);  // on a bin edge with an inexact FP value
    }
    int g = (hi * C.sp + si) * C.vp + vi;
    g = (mn >= t0) ? C.gray_gid : g;
    g = (vi == 0xFF) ? C.black_gid : g;
    o.gid = min(max(g, 0), C.T - 1);
    return o;
}

// saturation * 2^30 (rounded); matches rgb2hsv's special cases (max==0 -> 0, delta==max -> 0.999999)
__device__ __forceinline__ u32 phd_sat_q30(const FastPx& f, const PalTablesView& tb) {
    const u32 v = (u32)(((u64)f.q * tb.rs[f.mx] + 32768ull) >> 16);
    return f.mn == 0 ? (f.mx == 0 ? 0u : 1073740750u /* round(0.999999 * 2^30) */) : v;
}

// hue * 2^22 in [0, 360*2^22)
__device__ __forceinline__ long long phd_hue_q22(const FastPx& f, const PalTablesView& tb) {
    const int ap = f.p < 0 ? -f.p : f.p;
    const long long frac = (long long)(((u64)ap * tb.rh[f.q] + 32768ull) >> 16);  // 60 * |p|/q * 2^22 (rh[0] = 0)
    long long h = ((long long)(60 * f.off) << PHD_HQ_SHIFT) + (f.p < 0 ? -frac : frac);
    if (h < 0) h += PHD_HQ_360;
    return f.q == 0 ? 0ll : h;
}
