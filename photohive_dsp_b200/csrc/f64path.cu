// Planes of doubles that are NOT of the form k/255 (e.g. the reference's own create_test_rgb, src/debug.c:53, or 16-bit
// images divided by 255): the general-input route of get_full_report_data.
//
// The 8-bit pipeline (frontend.cu) rests on exact integer arithmetic over 2^24 colours; arbitrary doubles have no such
// structure, so this route follows the reference literally, in FP64, with two passes over the pixels:
//   k_f64_stats       get_rgb_statistics sums over the full image (src/image_processing.c:545-556, src/filtering.c:125-147)
//   k_f64_gray        rgb2pgm (:505-512) minus the channel-mean average (remove_dc_bias), scaled by 255000 like the 8-bit
//                     numerators: the FP32 input of the row transform
//   k_f64_classify    downsample_rgb walk + rgb2hsv + arm_octree bin of every HSV pixel (:344-417,
//                     src/color_quantization.c:108-161): per-chunk group counts (for the raster-order tie rule), group
//                     totals, saturation sum
//   (k_palette_select decides parents and merges from the group totals, as for 8-bit images)
//   k_f64_accumulate  calculate_avg_hsv (src/color_quantization.c:510-576): every pixel again, added to its parent with the
//                     reference's wrap t = h + 180 - h_parent; tie groups keep the first `take` pixels in raster order and
//                     the last one (:411-451)
//   k_f64_sharpness   get_variance_sharpness on the gray doubles (src/filtering.c:40-107,151-183)
// Sums are FP64 atomics: results agree with the reference to ~1e-12, but not bit for bit from run to run (the 8-bit route
// is; this one trades that for generality).  Throughput is not a goal here -- the route exists so that every input the
// reference accepts is served.
#include "hsv_exact.cuh"

namespace {

struct HsvP { double h, s, v; };

// rgb2hsv on doubles, operation by operation (src/image_processing.c:384-414; fmax/fmin, no FMA contraction)
__device__ __forceinline__ HsvP hsv_of_doubles(double r, double g, double b) {
    const double mx = fmax(fmax(r, g), b), mn = fmin(fmin(r, g), b);
    const double d = __dsub_rn(mx, mn);
    double h;
    if (d == 0) h = 0;
    else if (mx == r) h = __dmul_rn(60.0, __ddiv_rn(__dsub_rn(g, b), d));
    else if (mx == g) h = __dmul_rn(60.0, __dadd_rn(2.0, __ddiv_rn(__dsub_rn(b, r), d)));
    else h = __dmul_rn(60.0, __dadd_rn(4.0, __ddiv_rn(__dsub_rn(r, g), d)));
    if (0 < h && h < 360) {
    } else if (h < 0) {
        while (h < 0) h = __dadd_rn(h, 360.0);
    } else if (h > 360) {
        while (h > 360) h = __dsub_rn(h, 360.0);
    }
    HsvP o;
    o.h = h;
    o.v = (mx == 1) ? 0.999999 : mx;
    o.s = (mx == 0) ? 0.0 : ((d == mx) ? 0.999999 : __ddiv_rn(d, mx));
    return o;
}

__device__ __forceinline__ int group_of(const HsvP& p, const DevParams& P) {
    HsvD q;
    q.h = p.h; q.s = p.s; q.v = p.v; q.mx = 0;
    return phd_group_exact(q, P);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    return t;
}

// planes: [3][npx] doubles (r, g, b).  acc: [10] = sum r,g,b, sum of squares r,g,b, sum gray, (saturation sum),
// (Br+Bg+Bb)/3, unused.
__global__ void __launch_bounds__(256) k_f64_stats(const double* __restrict__ planes, DevParams P, double* __restrict__ acc) {
    __shared__ double red[8];
    const double* r = planes;
    const double* g = planes + P.npx;
    const double* b = planes + 2 * P.npx;
    double s[7] = {0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.npx; i += (long long)gridDim.x * blockDim.x) {
        const double x = r[i], y = g[i], z = b[i];
        s[0] += x; s[1] += y; s[2] += z;
        s[3] += x * x; s[4] += y * y; s[5] += z * z;
        s[6] += __dadd_rn(__dadd_rn(__dmul_rn(0.299, x), __dmul_rn(0.587, y)), __dmul_rn(0.114, z));
    }
    for (int k = 0; k < 7; k++) {
        const double t = block_sum(s[k], red);
        if (threadIdx.x == 0) atomicAdd(&acc[k], t);
    }
}

// The reference transforms gray - avg with avg = (Br+Bg+Bb)/3 (src/interface.c:78, src/blur_profile.c:233-238).
// X[0,0] = sum(gray) - P*avg is handed to the column kernel the way cols_fix_dc expects it.
__global__ void k_f64_dc(DevParams P, double* __restrict__ acc, ImageAcc* __restrict__ iacc) {
    const double np = (double)P.npx;
    const double avg = (acc[0] / np + acc[1] / np + acc[2] / np) / 3.0;
    acc[8] = avg;
    iacc->dc = acc[6] - np * avg;
    iacc->dc_valid = 1;
}

// rgb2pgm (src/image_processing.c:505-512) minus that average, scaled by 255000 like the 8-bit numerators: the FP32
// input of the row transform.  (Subtracting the true mean, not 0.5 as the 8-bit route does, keeps the float mantissa
// for the image's variation: create_test_rgb is almost constant.)
__global__ void __launch_bounds__(256) k_f64_gray(const double* __restrict__ planes, DevParams P,
                                                  const double* __restrict__ acc, float* __restrict__ gray32) {
    const double* r = planes;
    const double* g = planes + P.npx;
    const double* b = planes + 2 * P.npx;
    const double avg = acc[8];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.npx; i += (long long)gridDim.x * blockDim.x) {
        const double gr = __dadd_rn(__dadd_rn(__dmul_rn(0.299, r[i]), __dmul_rn(0.587, g[i])), __dmul_rn(0.114, b[i]));
        gray32[i] = (float)((gr - avg) * 255000.0);
    }
}

__global__ void __launch_bounds__(256) k_f64_classify(const double* __restrict__ planes, DevParams P,
                                                      u16* __restrict__ counts_chunk, u32* __restrict__ hist,
                                                      double* __restrict__ acc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32* cnt = reinterpret_cast<u32*>(smem_raw);  // [T]
    __shared__ double red[8];
    const int chunk = blockIdx.x, tid = threadIdx.x, T = P.T;
    const double* r = planes;
    const double* g = planes + P.npx;
    const double* b = planes + 2 * P.npx;
    for (int i = tid; i < T; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    const int ppt = P.chunk / 256;
    double ssum = 0;
    for (int i = 0; i < ppt; i++) {
        const long long p = (long long)chunk * P.chunk + (long long)tid * ppt + i;
        if (p >= P.hpx) break;
        const long long src = phd_src_index(p, P);
        const HsvP px = hsv_of_doubles(r[src], g[src], b[src]);
        ssum += px.s;
        atomicAdd(&cnt[group_of(px, P)], 1u);
    }
    const double t = block_sum(ssum, red);  // (its barriers also complete the shared counts)
    if (tid == 0) atomicAdd(&acc[7], t);
    u16* cc = counts_chunk + (size_t)chunk * T;
    for (int i = tid; i < T; i += blockDim.x) {
        cc[i] = (u16)cnt[i];
        if (cnt[i]) atomicAdd(&hist[i], cnt[i]);
    }
}

// slots: [T][4] doubles per parent slot: sum v, sum s, sum t (wrapped hue), unused.
__global__ void __launch_bounds__(256) k_f64_accumulate(const double* __restrict__ planes, DevParams P,
                                                        const double* __restrict__ centres,
                                                        const GroupPlan* __restrict__ plan_g,
                                                        const int* __restrict__ parent_ids,
                                                        double* __restrict__ slots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = P.T;
    double* acc = reinterpret_cast<double*>(smem_raw);         // [T][3]
    u16* gid = reinterpret_cast<u16*>(acc + 3 * (size_t)T);   // [chunk] group of every pixel of the chunk
    __shared__ int scan[256];
    __shared__ int sh_last, sh_any;
    const int chunk = blockIdx.x, tid = threadIdx.x;
    const double* r = planes;
    const double* g = planes + P.npx;
    const double* b = planes + 2 * P.npx;
    for (int i = tid; i < 3 * T; i += blockDim.x) acc[i] = 0;
    if (tid == 0) sh_any = 0;
    __syncthreads();
    const int ppt = P.chunk / 256;
    const long long c0 = (long long)chunk * P.chunk;
    auto add = [&](const HsvP& px, int slot) {
        const double off = __dsub_rn(180.0, centres[parent_ids[slot]]);
        double t = __dadd_rn(px.h, off);
        if (t > 360) t = __dsub_rn(t, 360.0);
        else if (t < 0) t = __dadd_rn(t, 360.0);
        atomicAdd(&acc[3 * slot], px.v);
        atomicAdd(&acc[3 * slot + 1], px.s);
        atomicAdd(&acc[3 * slot + 2], t);
    };
    auto pixel = [&](int li) -> HsvP {
        const long long src = phd_src_index(c0 + li, P);
        return hsv_of_doubles(r[src], g[src], b[src]);
    };
    for (int i = 0; i < ppt; i++) {
        const int li = tid * ppt + i;
        u16 gi = 0xffff;
        if (c0 + li < P.hpx) {
            const HsvP px = pixel(li);
            gi = (u16)group_of(px, P);
            const GroupPlan gp = plan_g[gi];
            if (gp.mode == 1) add(px, gp.slot);
            else if (gp.mode == 2) {
                if (chunk < gp.cstar) add(px, gp.slot);  // inside the accepted prefix
                else if ((chunk == gp.cstar && gp.need > 0) || chunk == gp.clast) sh_any = 1;  // ranked below
            }
        }
        gid[li] = gi;
    }
    __syncthreads();
    if (sh_any) {
        // tie groups whose cut-off chunk or last pixel is here: raster rank inside the chunk (src/color_quantization.c:435-440)
        for (int gsel = 0; gsel < T; gsel++) {
            const GroupPlan gp = plan_g[gsel];
            if (gp.mode != 2) continue;
            const bool partial = (gp.cstar == chunk && gp.need > 0), last = (gp.clast == chunk);
            if (!partial && !last) continue;  // uniform across the block
            int mine = 0, my_last = -1;
            for (int i = 0; i < ppt; i++)
                if (gid[tid * ppt + i] == gsel) { mine++; my_last = tid * ppt + i; }
            scan[tid] = mine;
            if (tid == 0) sh_last = -1;
            __syncthreads();
            if (tid == 0) {
                int run = 0;
                for (int t = 0; t < 256; t++) { const int c = scan[t]; scan[t] = run; run += c; }
            }
            if (my_last >= 0) atomicMax(&sh_last, my_last);
            __syncthreads();
            int rank = scan[tid];
            const int last_idx = sh_last;
            for (int i = 0; i < ppt; i++) {
                const int li = tid * ppt + i;
                if (gid[li] != gsel) continue;
                // the last pixel lies beyond the accepted prefix (take < n), so the two conditions never meet in one pixel
                const bool take = (partial && rank < gp.need) || (last && li == last_idx);
                rank++;
                if (take) add(pixel(li), gp.slot);
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < 3 * T; i += blockDim.x)
        if (acc[i] != 0) atomicAdd(&slots[4 * (i / 3) + (i % 3)], acc[i]);
}

// out: [boxes][2] doubles: sum f, sum f^2 with f the 3x3 Laplacian of the crop (zero outside the CROP,
// src/filtering.c:81-107) of the gray doubles.
__global__ void __launch_bounds__(256) k_f64_sharpness(const double* __restrict__ planes, DevParams P,
                                                       const int* __restrict__ boxes, double* __restrict__ out) {
    __shared__ double red[8];
    const int box = blockIdx.y;
    const int top = boxes[4 * box], bottom = boxes[4 * box + 1], left = boxes[4 * box + 2], right = boxes[4 * box + 3];
    const int w = right - left, h = bottom - top;
    if (w <= 0 || h <= 0) return;
    const double* r = planes;
    const double* g = planes + P.npx;
    const double* b = planes + 2 * P.npx;
    auto gray = [&](int y, int x) -> double {
        if (y < top || y >= bottom || x < left || x >= right) return 0.0;
        const long long i = (long long)y * P.W + x;
        return __dadd_rn(__dadd_rn(__dmul_rn(0.299, r[i]), __dmul_rn(0.587, g[i])), __dmul_rn(0.114, b[i]));
    };
    double s1 = 0, s2 = 0;
    const long long n = (long long)w * h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int y = top + (int)(i / w), x = left + (int)(i % w);
        double nb = 0;
        for (int dy = -1; dy <= 1; dy++)
            for (int dx = -1; dx <= 1; dx++)
                if (dy || dx) nb += gray(y + dy, x + dx);
        const double f = 8.0 * gray(y, x) - nb;
        s1 += f;
        s2 += f * f;
    }
    const double a = block_sum(s1, red), c = block_sum(s2, red);
    if (threadIdx.x == 0) {
        atomicAdd(&out[2 * box], a);
        atomicAdd(&out[2 * box + 1], c);
    }
}

}  // namespace

size_t phd_f64_accumulate_smem(const DevParams& P) { return (size_t)3 * P.T * sizeof(double) + (size_t)P.chunk * sizeof(u16); }

void phd_launch_f64_front(const double* planes, const DevParams& P, F64Work& fw, Workspace& ws, cudaStream_t st, int* launches) {
    int blocks = (int)((P.npx + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    k_f64_stats<<<blocks, 256, 0, st>>>(planes, P, fw.acc);
    k_f64_dc<<<1, 1, 0, st>>>(P, fw.acc, ws.iacc);
    k_f64_gray<<<blocks, 256, 0, st>>>(planes, P, fw.acc, fw.gray32);
    k_f64_classify<<<P.nchunks, 256, (size_t)P.T * sizeof(u32), st>>>(planes, P, ws.counts_chunk, ws.hist, fw.acc);
    *launches += 4;
}

void phd_launch_f64_accumulate(const double* planes, const DevParams& P, const double* centres, F64Work& fw, Workspace& ws,
                               cudaStream_t st, int* launches) {
    const size_t smem = phd_f64_accumulate_smem(P);
    PHD_ALLOW_SMEM((k_f64_accumulate), 200 * 1024);
    k_f64_accumulate<<<P.nchunks, 256, smem, st>>>(planes, P, centres, ws.plan, ws.parent_ids, fw.slots);
    *launches += 1;
    if (P.max_boxes > 0) {
        k_f64_sharpness<<<dim3(64, P.max_boxes), 256, 0, st>>>(planes, P, ws.boxes, fw.sharp);
        *launches += 1;
    }
}
