// Laplacian-variance sharpness of the salient crop boxes.
//
// Replaces get_variance_sharpness (src/filtering.c:151-183): crop_pgm (src/image_processing.c:213-232),
// the zero-padded 3x3 Laplacian (src/filtering.c:40-50,81-107) and get_average/get_variance
// (:125-147).  The gray value of a pixel is G/255000 with the exact integer numerator
// G = 299R+587G+114B, so the filter response is an exact integer and the two moments are
// accumulated as integers; finalize turns them into variance/mean.
#include "phd_internal.h"

namespace {

__device__ __forceinline__ int gray_num(const uint8_t* __restrict__ base, int W, int y, int x) {
    const uint8_t* q = base + ((size_t)y * W + x) * 3;
    return 299 * (int)__ldg(q) + 587 * (int)__ldg(q + 1) + 114 * (int)__ldg(q + 2);
}

__global__ void __launch_bounds__(256) k_sharpness(const uint8_t* __restrict__ rgb, DevParams P,
                                                   const int* __restrict__ boxes, SharpAcc* __restrict__ acc,
                                                   int tiles_x) {
    const int img = blockIdx.z, box = blockIdx.y;
    const int* bx = boxes + ((size_t)img * P.max_boxes + box) * 4;
    const int top = bx[0], bottom = bx[1], left = bx[2], right = bx[3];
    const int w = right - left, h = bottom - top;
    if (w <= 0 || h <= 0 || left < 0 || top < 0 || right > P.W || bottom > P.H) return;  // finalize reports NaN
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    if (tx * 32 >= w || ty * 32 >= h) return;  // CTA-uniform
    const int x = tx * 32 + (threadIdx.x & 31);
    const int y0 = ty * 32 + (threadIdx.x >> 5) * 4;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    long long s1 = 0;
    u64 s2 = 0;
    if (x < w) {
        for (int dy = 0; dy < 4; dy++) {
            const int y = y0 + dy;
            if (y >= h) break;
            int f = 0;
#pragma unroll
            for (int fy = -1; fy <= 1; fy++)
#pragma unroll
                for (int fx = -1; fx <= 1; fx++) {
                    const int iy = y + fy, ix = x + fx;
                    if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
                        const int g = gray_num(base, P.W, iy + top, ix + left);
                        f += (fy == 0 && fx == 0) ? 8 * g : -g;
                    }
                }
            s1 += f;
            s2 += (u64)((long long)f * (long long)f);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    __shared__ long long r1[8];
    __shared__ u64 r2[8];
    if ((threadIdx.x & 31) == 0) { r1[threadIdx.x >> 5] = s1; r2[threadIdx.x >> 5] = s2; }
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
    // grid covers the largest box of the sub-batch; CTAs outside their own box retire immediately
    const int tiles_x = (max_w + 31) / 32, tiles_y = (max_h + 31) / 32;
    dim3 grid(tiles_x * tiles_y, P.max_boxes, nimg);
    k_sharpness<<<grid, 256, 0, st>>>(rgb, P, ws.boxes, ws.sharp, tiles_x);
    *launches += 1;
}
