// Pixel front end and palette accumulation.
//
//   k_frontend            replaces downsample_rgb / rgb2hsv / get_rgb_statistics / get_hsv_average and the
//                         counting half of arm_octree (src/image_processing.c:344-417,533-553,
//                         src/color_quantization.c:108-161): one read of the packed RGB bytes gives the
//                         channel sums, the saturation sum and the per-chunk palette histogram.
//   k_palette_accumulate  replaces the pixel moving of group_irregular_pixels and calculate_avg_hsv
//                         (src/color_quantization.c:342-479,510-576): second pass (RGB is L2 resident)
//                         that sums wrapped hue / s / v per parent, honouring the tie-path rule
//                         "first `room` pixels in raster order + the very last pixel".
//   k_group_sweep         test hook: group id of all 2^24 colours.
//
// All accumulators are integers (fixed point where needed) so results do not depend on scheduling.
#include "hsv_exact.cuh"

namespace {

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
template <bool FUSED_STATS>
__global__ void __launch_bounds__(PHD_FE_THREADS) k_frontend(const uint8_t* __restrict__ rgb, DevParams P,
                                                             u16* __restrict__ counts_chunk, u32* __restrict__ hist,
                                                             ImageAcc* __restrict__ iacc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* k255 = reinterpret_cast<double*>(smem_raw);
    u32* sh_hist = reinterpret_cast<u32*>(smem_raw + 256 * sizeof(double));
    __shared__ u64 red[7][PHD_FE_THREADS / 32];

    const int img = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    phd_fill_k255(k255);
    for (int g = tid; g < P.T; g += blockDim.x) sh_hist[g] = 0;
    __syncthreads();

    const long long p0 = (long long)chunk * PHD_CHUNK + (long long)tid * PHD_PX_PER_THREAD;
    u64 s_acc = 0;
    u32 sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    if (p0 < P.hpx) {
        if (P.ds <= 1) {
            u32 w[12];
            const long long valid = (P.hpx - p0) * 3;
            phd_load48(base + p0 * 3, w, P.aligned16 != 0, valid);
#pragma unroll
            for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
                if (p0 + i < P.hpx) {
                    const int R = phd_byte_of(w, 3 * i), G = phd_byte_of(w, 3 * i + 1), B = phd_byte_of(w, 3 * i + 2);
                    const HsvD px = phd_hsv_exact(R, G, B, k255);
                    atomicAdd(&sh_hist[phd_group_exact(px, P)], 1u);
                    s_acc += (u64)__double2ll_rn(px.s * (double)(1 << PHD_S_SHIFT));
                    if (FUSED_STATS) {
                        sum[0] += R; sum[1] += G; sum[2] += B;
                        sq[0] += R * R; sq[1] += G * G; sq[2] += B * B;
                    }
                }
            }
        } else {
            for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
                if (p0 + i < P.hpx) {
                    const uint8_t* q = base + phd_src_index(p0 + i, P) * 3;
                    const int R = __ldg(q), G = __ldg(q + 1), B = __ldg(q + 2);
                    const HsvD px = phd_hsv_exact(R, G, B, k255);
                    atomicAdd(&sh_hist[phd_group_exact(px, P)], 1u);
                    s_acc += (u64)__double2ll_rn(px.s * (double)(1 << PHD_S_SHIFT));
                }
            }
        }
    }
    // block reduction of the scalar sums -> one 64-bit global atomic each
    u64 vals[7] = {s_acc, sum[0], sum[1], sum[2], sq[0], sq[1], sq[2]};
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll
    for (int k = 0; k < 7; k++) {
        if (!FUSED_STATS && k > 0) break;
        u64 v = warp_sum_u64(vals[k]);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (tid < 7 && (FUSED_STATS || tid == 0)) {
        u64 v = 0;
        for (int w = 0; w < PHD_FE_THREADS / 32; w++) v += red[tid][w];
        ImageAcc* a = iacc + img;
        u64* dst = tid == 0 ? &a->s_sum : (tid <= 3 ? &a->sum[tid - 1] : &a->sumsq[tid - 4]);
        if (v) atomicAdd(dst, v);
    }
    // per-chunk histogram (needed for raster ranks in the tie path) + image histogram
    u16* cc = counts_chunk + ((size_t)img * P.nchunks + chunk) * P.T;
    for (int g = tid; g < P.T; g += blockDim.x) {
        const u32 c = sh_hist[g];
        cc[g] = (u16)c;
        if (c) atomicAdd(&hist[(size_t)img * P.T + g], c);
    }
}

// Channel sums over the FULL image when the HSV image is downsampled (src/interface.c:50-55).
__global__ void __launch_bounds__(256) k_rgb_stats(const uint8_t* __restrict__ rgb, DevParams P,
                                                   ImageAcc* __restrict__ iacc) {
    __shared__ u64 red[6][8];
    const int img = blockIdx.y, tid = threadIdx.x;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    u64 sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    for (long long p = (long long)blockIdx.x * blockDim.x + tid; p < P.npx; p += (long long)gridDim.x * blockDim.x) {
        const uint8_t* q = base + p * 3;
        const u32 R = __ldg(q), G = __ldg(q + 1), B = __ldg(q + 2);
        sum[0] += R; sum[1] += G; sum[2] += B;
        sq[0] += R * R; sq[1] += G * G; sq[2] += B * B;
    }
    const int lane = tid & 31, wid = tid >> 5;
    for (int k = 0; k < 3; k++) {
        u64 a = warp_sum_u64(sum[k]), b = warp_sum_u64(sq[k]);
        if (lane == 0) { red[k][wid] = a; red[3 + k][wid] = b; }
    }
    __syncthreads();
    if (tid < 6) {
        u64 v = 0;
        for (int w = 0; w < 8; w++) v += red[tid][w];
        ImageAcc* a = iacc + img;
        if (v) atomicAdd(tid < 3 ? &a->sum[tid] : &a->sumsq[tid - 3], v);
    }
}

// ------------------------------------------------------------------------------------------
struct SlotSm {
    u32* cnt; u32* summax; u32* n255; u32* s_lo; u32* s_hi; u32* t_lo; u32* t_hi;
};

__device__ __forceinline__ void slot_add(const SlotSm& S, int slot, double off, const HsvD& px) {
    double t = __dadd_rn(px.h, off);
    if (t > 360.0) t = __dsub_rn(t, 360.0);
    else if (t < 0.0) t = __dadd_rn(t, 360.0);
    const u32 tq = (u32)__double2ll_rn(t * (double)(1 << PHD_T_SHIFT));
    const u32 sq = (u32)__double2ll_rn(px.s * (double)(1 << PHD_S_SHIFT));
    atomicAdd(&S.cnt[slot], 1u);
    atomicAdd(&S.summax[slot], (u32)px.mx);
    if (px.mx == 255) atomicAdd(&S.n255[slot], 1u);
    atomicAdd(&S.s_lo[slot], sq & 0x7fffu);
    atomicAdd(&S.s_hi[slot], sq >> 15);
    atomicAdd(&S.t_lo[slot], tq & 0x7fffu);
    atomicAdd(&S.t_hi[slot], tq >> 15);
}

__global__ void __launch_bounds__(PHD_FE_THREADS) k_palette_accumulate(
    const uint8_t* __restrict__ rgb, DevParams P, const double* __restrict__ centres,
    const GroupPlan* __restrict__ plan_g, const int* __restrict__ pal_n, const int* __restrict__ parent_ids,
    const int* __restrict__ tie_list, const int* __restrict__ tie_n, SlotAcc* __restrict__ sacc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = P.T;
    double* k255 = reinterpret_cast<double*>(smem_raw);
    double* off = k255 + 256;                                   // [T]
    GroupPlan* plan = reinterpret_cast<GroupPlan*>(off + T);    // [T]
    u32* acc = reinterpret_cast<u32*>(plan + T);                // [7][T]
    u16* gid_cache = reinterpret_cast<u16*>(acc + 7 * T);       // [PHD_CHUNK]
    __shared__ int scan[PHD_FE_THREADS];
    __shared__ int sh_last;

    const int img = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    const int N = pal_n[img];
    phd_fill_k255(k255);
    for (int g = tid; g < T; g += blockDim.x) plan[g] = plan_g[(size_t)img * T + g];
    for (int j = tid; j < N; j += blockDim.x) off[j] = __dsub_rn(180.0, centres[parent_ids[(size_t)img * T + j]]);
    for (int i = tid; i < 7 * T; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    SlotSm S{acc, acc + T, acc + 2 * T, acc + 3 * T, acc + 4 * T, acc + 5 * T, acc + 6 * T};

    const long long p0 = (long long)chunk * PHD_CHUNK + (long long)tid * PHD_PX_PER_THREAD;
    u32 w[12];
    if (p0 < P.hpx && P.ds <= 1) phd_load48(base + p0 * 3, w, P.aligned16 != 0, (P.hpx - p0) * 3);
#pragma unroll
    for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
        int gid = 0xffff;
        if (p0 + i < P.hpx) {
            int R, G, B;
            if (P.ds <= 1) {
                R = phd_byte_of(w, 3 * i); G = phd_byte_of(w, 3 * i + 1); B = phd_byte_of(w, 3 * i + 2);
            } else {
                const uint8_t* q = base + phd_src_index(p0 + i, P) * 3;
                R = __ldg(q); G = __ldg(q + 1); B = __ldg(q + 2);
            }
            const HsvD px = phd_hsv_exact(R, G, B, k255);
            gid = phd_group_exact(px, P);
            const GroupPlan gp = plan[gid];
            if (gp.mode == 1 || (gp.mode == 2 && chunk < gp.cstar)) slot_add(S, gp.slot, off[gp.slot], px);
        }
        gid_cache[tid * PHD_PX_PER_THREAD + i] = (u16)gid;
    }
    __syncthreads();

    // Tie groups whose partial chunk, or whose last pixel, falls in this chunk: ordered pass.
    const int nt = tie_n[img];
    for (int k = 0; k < nt; k++) {
        const int g = tie_list[(size_t)img * T + k];
        const GroupPlan gp = plan[g];
        const bool partial = (gp.cstar == chunk && gp.need > 0);
        const bool last = (gp.clast == chunk);
        if (!partial && !last) continue;  // uniform across the block
        int mine = 0, my_last = -1;
#pragma unroll
        for (int i = 0; i < PHD_PX_PER_THREAD; i++)
            if (gid_cache[tid * PHD_PX_PER_THREAD + i] == g) { mine++; my_last = tid * PHD_PX_PER_THREAD + i; }
        scan[tid] = mine;
        if (tid == 0) sh_last = -1;
        __syncthreads();
        // exclusive prefix over 256 threads (small, done naively by warp 0 lanes in sequence chunks)
        if (tid == 0) {
            int run = 0;
            for (int t = 0; t < PHD_FE_THREADS; t++) { int c = scan[t]; scan[t] = run; run += c; }
        }
        if (my_last >= 0) atomicMax(&sh_last, my_last);
        __syncthreads();
        int rank = scan[tid];
        const int last_idx = sh_last;
        for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
            const int li = tid * PHD_PX_PER_THREAD + i;
            if (gid_cache[li] != g) continue;
            const bool take = (partial && rank < gp.need) || (last && li == last_idx && !(partial && rank < gp.need));
            rank++;
            if (!take) continue;
            const uint8_t* q = base + phd_src_index((long long)chunk * PHD_CHUNK + li, P) * 3;
            const HsvD px = phd_hsv_exact(__ldg(q), __ldg(q + 1), __ldg(q + 2), k255);
            slot_add(S, gp.slot, off[gp.slot], px);
        }
        __syncthreads();
    }
    __syncthreads();

    for (int j = tid; j < N; j += blockDim.x) {
        const u32 c = S.cnt[j];
        if (!c) continue;
        SlotAcc* a = sacc + (size_t)img * T + j;
        atomicAdd(&a->cnt, (u64)c);
        atomicAdd(&a->summax, (u64)S.summax[j]);
        if (S.n255[j]) atomicAdd(&a->n255, (u64)S.n255[j]);
        atomicAdd(&a->s_sum, ((u64)S.s_hi[j] << 15) + S.s_lo[j]);
        atomicAdd(&a->t_sum, ((u64)S.t_hi[j] << 15) + S.t_lo[j]);
    }
}

__global__ void __launch_bounds__(256) k_group_sweep(DevParams P, u16* __restrict__ out) {
    __shared__ double k255[256];
    phd_fill_k255(k255);
    __syncthreads();
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;  // r<<16 | g<<8 | b
    const HsvD px = phd_hsv_exact((c >> 16) & 255, (c >> 8) & 255, c & 255, k255);
    out[c] = (u16)phd_group_exact(px, P);
}

}  // namespace

void phd_launch_frontend(const uint8_t* rgb, const DevParams& P, int nimg, const double* centres, Workspace& ws,
                         cudaStream_t st, int* launches) {
    (void)centres;
    dim3 grid(P.nchunks, nimg);
    const size_t smem = 256 * sizeof(double) + (size_t)P.T * sizeof(u32);
    if (P.ds <= 1) {
        k_frontend<true><<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, ws.counts_chunk, ws.hist, ws.iacc);
        *launches += 1;
    } else {
        k_frontend<false><<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, ws.counts_chunk, ws.hist, ws.iacc);
        int blocks = (int)((P.npx + 256LL * 16 - 1) / (256LL * 16));
        if (blocks < 1) blocks = 1;
        k_rgb_stats<<<dim3(blocks, nimg), 256, 0, st>>>(rgb, P, ws.iacc);
        *launches += 2;
    }
}

void phd_launch_palette_accumulate(const uint8_t* rgb, const DevParams& P, int nimg, const double* centres,
                                   Workspace& ws, cudaStream_t st, int* launches) {
    dim3 grid(P.nchunks, nimg);
    const size_t smem = 256 * sizeof(double) + (size_t)P.T * (sizeof(double) + sizeof(GroupPlan) + 7 * sizeof(u32)) +
                        PHD_CHUNK * sizeof(u16);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_palette_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    k_palette_accumulate<<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, centres, ws.plan, ws.pal_n, ws.parent_ids,
                                                             ws.tie_list, ws.tie_n, ws.sacc);
    *launches += 1;
}

void phd_launch_group_sweep(const DevParams& P, u16* out_dev, cudaStream_t st) {
    k_group_sweep<<<(1 << 24) / 256, 256, 0, st>>>(P, out_dev);
}
