// Laplacian-variance sharpness of the salient crop boxes.
//
// Replaces get_variance_sharpness (src/filtering.c:151-183): crop_pgm (src/image_processing.c:213-232),
// the zero-padded 3x3 Laplacian (src/filtering.c:40-50,81-107) and get_average/get_variance
// (:125-147).  The gray value of a pixel is G/255000 with the exact integer numerator
// G = 299R+587G+114B, so the filter response is an exact integer and the two moments are
// accumulated as integers; finalize turns them into variance/mean.
//
// f(x,y) = 8 g(x,y) - (sum of the 8 neighbours) = 9 g(x,y) - (r(y-1) + r(y) + r(y+1)),  r(y) = g(x-1,y)+g(x,y)+g(x+1,y),
// with g = 0 outside the CROP (the reference pads the cropped image with zeros).  A warp owns a strip of 126 output
// columns: every lane holds FOUR adjacent pixels (the 128 pixels of the warp include one halo column on each side),
// read as four aligned 32-bit words per row (12 useful bytes, funnel-shifted into place) and turned into gray
// numerators by two-way dot products on the packed bytes; the horizontal neighbours across lanes come from two
// shuffles per row, and the last two row sums of every column stay in registers while the warp walks down.
#include "phd_internal.h"

namespace {

constexpr int kStripCols = 126;  // output columns per warp
constexpr int kStripRows = 96;   // output rows per warp

// 299 R + 587 G + 114 B of pixel i (0..3) of 12 packed bytes (IDP.2A on 16-bit weight pairs, any byte phase)
__device__ __forceinline__ int gray_of(const u32 (&w)[3], int i) {
    constexpr u32 RG = 299u | (587u << 16), B_ = 114u, _R = 299u << 16, GB = 587u | (114u << 16);
    if (i == 0) return (int)__dp2a_hi(B_, w[0], __dp2a_lo(RG, w[0], 0u));       // R G B .
    if (i == 1) return (int)__dp2a_lo(GB, w[1], __dp2a_hi(_R, w[0], 0u));       // . . . R | G B
    if (i == 2) return (int)__dp2a_lo(B_, w[2], __dp2a_hi(RG, w[1], 0u));       // . . R G | B
    return (int)__dp2a_hi(GB, w[2], __dp2a_lo(_R, w[2], 0u));                   // . R G B
}

__global__ void __launch_bounds__(256) k_sharpness(const uint8_t* __restrict__ rgb, DevParams P,
                                                   const int* __restrict__ boxes, SharpAcc* __restrict__ acc,
                                                   int strips_x) {
    const int img = blockIdx.z, box = blockIdx.y;
    const int* bx = boxes + ((size_t)img * P.max_boxes + box) * 4;
    const int top = bx[0], bottom = bx[1], left = bx[2], right = bx[3];
    const int w = right - left, h = bottom - top;
    if (w <= 0 || h <= 0 || left < 0 || top < 0 || right > P.W || bottom > P.H) return;  // finalize reports NaN
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int strip = blockIdx.x * 8 + wid;  // warp-uniform
    const int sy = strip / strips_x, sx = strip - sy * strips_x;
    const int x0 = sx * kStripCols, y0 = sy * kStripRows;
    long long s1 = 0;
    u64 s2 = 0;
    if (x0 < w && y0 < h) {
        const int xl = x0 - 1 + 4 * lane;  // crop column of this lane's first pixel
        bool in_x[4], out_x[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int x = xl + j;
            in_x[j] = x >= 0 && x < w;
            out_x[j] = in_x[j] && x >= x0 && x < x0 + kStripCols;
        }
        const bool all_in = in_x[0] && in_x[3];
        const bool any_in = xl + 3 >= 0 && xl < w;
        const uint8_t* img_base = rgb + (size_t)img * P.image_stride;
        const uint8_t* img_end = img_base + (size_t)P.W * P.H * 3;
        const uint8_t* col = img_base + ((long long)top * P.W + left + xl) * 3;  // may point before the row for xl = -1
        int gp[4] = {0, 0, 0, 0}, rpp[4] = {0, 0, 0, 0}, rp[4] = {0, 0, 0, 0};
        auto row = [&](int y, int (&g)[4], int (&r)[4]) {  // gray of the lane's four pixels and their 3-wide row sums
            g[0] = g[1] = g[2] = g[3] = 0;
            if (any_in && y >= 0 && y < h) {
                const uint8_t* q = col + (size_t)y * P.W * 3;
                const uint8_t* qa = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(q) & ~(uintptr_t)3);
                if (all_in && qa + 16 <= img_end) {
                    const u32* p4 = reinterpret_cast<const u32*>(qa);
                    const u32 l0 = __ldg(p4), l1 = __ldg(p4 + 1), l2 = __ldg(p4 + 2), l3 = __ldg(p4 + 3);
                    const u32 sh = (u32)(reinterpret_cast<uintptr_t>(q) & 3) * 8;
                    const u32 wv[3] = {__funnelshift_r(l0, l1, sh), __funnelshift_r(l1, l2, sh), __funnelshift_r(l2, l3, sh)};
#pragma unroll
                    for (int j = 0; j < 4; j++) g[j] = gray_of(wv, j);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (in_x[j])
                            g[j] = 299 * (int)__ldg(q + 3 * j) + 587 * (int)__ldg(q + 3 * j + 1) + 114 * (int)__ldg(q + 3 * j + 2);
                }
            }
            const int l = __shfl_up_sync(0xffffffffu, g[3], 1), rr = __shfl_down_sync(0xffffffffu, g[0], 1);
            r[0] = l + g[0] + g[1];
            r[1] = g[0] + g[1] + g[2];
            r[2] = g[1] + g[2] + g[3];
            r[3] = g[2] + g[3] + rr;
        };
        int g[4], r[4];
        row(y0 - 1, g, rpp);
        row(y0, gp, rp);
        const int y_end = min(y0 + kStripRows, h);
        int t1 = 0;  // |f| <= 9 * 255000 and at most 4 * kStripRows terms per lane: fits 32 bits
        for (int y = y0 + 1; y <= y_end; y++) {  // row y completes the window of output row y-1
            row(y, g, r);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int f = out_x[j] ? 9 * gp[j] - (rpp[j] + rp[j] + r[j]) : 0;
                t1 += f;
                s2 += (u64)((long long)f * (long long)f);
                rpp[j] = rp[j]; rp[j] = r[j]; gp[j] = g[j];
            }
        }
        s1 = t1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ long long r1[8];
    __shared__ u64 r2[8];
    if (lane == 0) { r1[wid] = s1; r2[wid] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0;
        u64 b = 0;
        for (int i = 0; i < 8; i++) { a += r1[i]; b += r2[i]; }
        SharpAcc* A = acc + (size_t)img * P.max_boxes + box;
        if (a) atomicAdd(reinterpret_cast<u64*>(&A->s1), (u64)a);
        if (b) {
            atomicAdd(&A->s2lo, b & 0xffffffffull);
            atomicAdd(&A->s2hi, b >> 32);
        }
    }
}

}  // namespace

void phd_launch_sharpness(const uint8_t* rgb, const DevParams& P, int nimg, int max_w, int max_h, Workspace& ws,
                          cudaStream_t st, int* launches) {
    if (P.max_boxes <= 0 || max_w <= 0 || max_h <= 0) return;
    // grid covers the largest box of the sub-batch; warps outside their own box retire immediately
    const int strips_x = (max_w + kStripCols - 1) / kStripCols, strips_y = (max_h + kStripRows - 1) / kStripRows;
    dim3 grid((strips_x * strips_y + 7) / 8, P.max_boxes, nimg);
    k_sharpness<<<grid, 256, 0, st>>>(rgb, P, ws.boxes, ws.sharp, strips_x);
    *launches += 1;
}
