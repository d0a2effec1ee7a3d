// Exact per-pixel HSV conversion and palette-group binning for 8-bit colours.
//
// The reference works on doubles k/255.0 (utils.py:30-37) and its bin edges follow IEEE double
// rounding, so every operation that feeds a discrete decision is done here with the explicitly
// rounded, never-contracted intrinsics (__dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn) in the order of
// src/image_processing.c:384-414 (rgb2hsv) and src/color_quantization.c:127-146 (arm_octree).
// tests/test_gpu_parity.py sweeps all 2^24 colours against the oracle.
#pragma once

#include "phd_internal.h"

struct HsvD {
    double h, s, v;
    int mx;  // integer max channel (v is mx/255.0, or 0.999999 when mx == 255)
};

// k255: shared-memory table of (double)k / 255.0, k = 0..255
__device__ __forceinline__ void phd_fill_k255(double* k255) {
    for (int k = threadIdx.x; k < 256; k += blockDim.x) k255[k] = __ddiv_rn((double)k, 255.0);
}

__device__ __forceinline__ HsvD phd_hsv_exact(int R, int G, int B, const double* __restrict__ k255) {
    const int mxi = max(R, max(G, B));
    const int mni = min(R, min(G, B));
    const double r = k255[R], g = k255[G], b = k255[B];
    const double mx = k255[mxi], mn = k255[mni];
    const double d = __dsub_rn(mx, mn);
    double h;
    if (mxi == mni) h = 0.0;
    else if (R == mxi) h = __dmul_rn(60.0, __ddiv_rn(__dsub_rn(g, b), d));
    else if (G == mxi) h = __dmul_rn(60.0, __dadd_rn(2.0, __ddiv_rn(__dsub_rn(b, r), d)));
    else h = __dmul_rn(60.0, __dadd_rn(4.0, __ddiv_rn(__dsub_rn(r, g), d)));
    while (h < 0.0) h = __dadd_rn(h, 360.0);
    HsvD o;
    o.h = h;
    o.mx = mxi;
    o.v = (mxi == 255) ? 0.999999 : mx;
    o.s = (mxi == 0) ? 0.0 : ((mni == 0) ? 0.999999 : __ddiv_rn(d, mx));
    return o;
}

__device__ __forceinline__ int phd_group_exact(const HsvD& p, const DevParams& P) {
    int g;
    if (p.v < P.bt) g = P.T - 1;
    else if (p.s < P.gt) g = P.T - (P.vp + 1);  // (int)(v - black) binds first: always the first gray group
    else {
        const int Vi = (int)__ddiv_rn(__dsub_rn(p.v, P.bt), P.Lv);
        const int Si = (int)__ddiv_rn(__dsub_rn(p.s, P.gt), P.Ls);
        const int Hi = (int)__ddiv_rn(p.h, P.Lh);
        g = (Hi * P.sp + Si) * P.vp + Vi;
    }
    // The reference indexes out of bounds here when h_partitions does not divide 360; keep memory safe.
    return min(max(g, 0), P.T - 1);
}

// 48 bytes = 16 packed RGB pixels into 12 words.
__device__ __forceinline__ void phd_load48(const uint8_t* __restrict__ p, u32 (&w)[12], bool aligned, long long valid_bytes) {
    if (aligned && valid_bytes >= 48) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) {
            u32 v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                long long o = 4 * i + j;
                u32 byte = (o < valid_bytes) ? (u32)__ldg(p + o) : 0u;
                v |= byte << (8 * j);
            }
            w[i] = v;
        }
    }
}

__device__ __forceinline__ int phd_byte_of(const u32 (&w)[12], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 255; }

// Source pixel index (in the full image) of HSV pixel i: identity, or the reference's downsample
// walk src/image_processing.c:344-366 (row stride N-1, :351).
__device__ __forceinline__ long long phd_src_index(long long i, const DevParams& P) {
    if (P.ds <= 1) return i;
    long long y = i / P.dw, x = i - y * P.dw;
    return y * (long long)(P.ds - 1) * P.W + x * P.ds;
}
