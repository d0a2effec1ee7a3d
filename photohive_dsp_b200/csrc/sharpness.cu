// Laplacian-variance sharpness of the salient crop boxes.
//
// Replaces get_variance_sharpness (src/filtering.c:151-183): crop_pgm (src/image_processing.c:213-232),
// the zero-padded 3x3 Laplacian (src/filtering.c:40-50,81-107) and get_average/get_variance
// (:125-147).  The gray value of a pixel is G/255000 with the exact integer numerator
// G = 299R+587G+114B, so the filter response is an exact integer and the two moments are
// accumulated as integers; finalize turns them into variance/mean.
//
// f(x,y) = 8 g(x,y) - (sum of the 8 neighbours) = 9 g(x,y) - (r(y-1) + r(y) + r(y+1)),  r(y) = g(x-1,y)+g(x,y)+g(x+1,y),
// with g = 0 outside the CROP (the reference pads the cropped image with zeros).  A warp owns a strip of 30 output
// columns (its 32 lanes also carry the two halo columns), walks down the rows keeping the last two row sums in
// registers, and gets the horizontal neighbours by shuffle: three byte loads per pixel instead of 27.
#include "phd_internal.h"

namespace {

constexpr int kStripCols = 30;  // output columns per warp
constexpr int kStripRows = 64;  // output rows per warp

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
        const int x = x0 - 1 + lane;              // crop coordinates of this lane's column (halo lanes 0 and 31)
        const bool in_x = x >= 0 && x < w;
        const bool out_lane = lane >= 1 && lane <= kStripCols && x < w;
        const uint8_t* col = rgb + (size_t)img * P.image_stride + ((size_t)top * P.W + left + (in_x ? x : 0)) * 3;
        auto row = [&](int y, int* g, int* r) {  // gray of (x, y) and the 3-wide row sum, zero outside the crop
            int v = 0;
            if (in_x && y >= 0 && y < h) {
                const uint8_t* q = col + (size_t)y * P.W * 3;
                v = 299 * (int)__ldg(q) + 587 * (int)__ldg(q + 1) + 114 * (int)__ldg(q + 2);
            }
            const int l = __shfl_up_sync(0xffffffffu, v, 1), rr = __shfl_down_sync(0xffffffffu, v, 1);
            *g = v;
            *r = l + v + rr;
        };
        int g_prev, r_prev2, r_prev;  // centre of row y-1, row sums of rows y-2 and y-1
        int g, r;
        row(y0 - 1, &g, &r_prev2);
        row(y0, &g_prev, &r_prev);
        const int y_end = min(y0 + kStripRows, h);
        for (int y = y0 + 1; y <= y_end; y++) {  // row y completes the window of output row y-1
            row(y, &g, &r);
            if (out_lane) {
                const int f = 9 * g_prev - (r_prev2 + r_prev + r);
                s1 += f;
                s2 += (u64)((long long)f * (long long)f);
            }
            r_prev2 = r_prev; r_prev = r; g_prev = g;
        }
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
