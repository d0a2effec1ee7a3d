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
#include "hsv_fast.cuh"

namespace {

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// One CTA per chunk of PHD_CHUNK HSV pixels; each thread owns 16 consecutive pixels (48 bytes, three 16-byte
// loads).  Histogram increments are run-length merged per thread before they hit shared memory.
template <bool FUSED_STATS>
__global__ void __launch_bounds__(PHD_FE_THREADS) k_frontend(const uint8_t* __restrict__ rgb, DevParams P,
                                                             const unsigned char* __restrict__ pal_tables,
                                                             u16* __restrict__ counts_chunk, u32* __restrict__ hist,
                                                             ImageAcc* __restrict__ iacc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* k255 = reinterpret_cast<double*>(smem_raw);
    unsigned char* tb_raw = smem_raw + 256 * sizeof(double);
    u32* sh_hist = reinterpret_cast<u32*>(tb_raw + phd_pal_tables_bytes(P.sp));
    __shared__ u64 red[7][PHD_FE_THREADS / 32];

    const int img = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    phd_fill_k255(k255);
    phd_pal_tables_to_smem(tb_raw, pal_tables, P.sp);
    for (int g = tid; g < P.T; g += blockDim.x) sh_hist[g] = 0;
    __syncthreads();
    const PalTablesView tb = phd_pal_tables_view(tb_raw);
    const FastCfg C = phd_fast_cfg(P);

    const long long p0 = (long long)chunk * PHD_CHUNK + (long long)tid * PHD_PX_PER_THREAD;
    u64 s_acc = 0;
    u32 sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    int run_gid = -1;
    u32 run_n = 0;
    if (p0 < P.hpx) {
        u32 w[12];
        if (P.ds <= 1) phd_load48(base + p0 * 3, w, P.aligned16 != 0, (P.hpx - p0) * 3);
#pragma unroll
        for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
            if (p0 + i < P.hpx) {
                int R, G, B;
                if (P.ds <= 1) {
                    R = phd_byte_of(w, 3 * i); G = phd_byte_of(w, 3 * i + 1); B = phd_byte_of(w, 3 * i + 2);
                } else {
                    const uint8_t* q = base + phd_src_index(p0 + i, P) * 3;
                    R = __ldg(q); G = __ldg(q + 1); B = __ldg(q + 2);
                }
                const FastPx px = phd_group_fast(R, G, B, tb, C, k255);
                if (px.gid != run_gid) {
                    if (run_n) atomicAdd(&sh_hist[run_gid], run_n);
                    run_gid = px.gid;
                    run_n = 0;
                }
                run_n++;
                s_acc += phd_sat_q30(px, tb);
                if (FUSED_STATS) {
                    sum[0] += R; sum[1] += G; sum[2] += B;
                    sq[0] += R * R; sq[1] += G * G; sq[2] += B * B;
                }
            }
        }
        if (run_n) atomicAdd(&sh_hist[run_gid], run_n);
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
// Per-thread run of consecutive pixels that go to the same parent slot; flushed with native 32-bit
// shared-memory atomics (ATOMS.ADD) when the slot changes.
struct SlotRun {
    int slot;
    u32 summax, n255;
    u64 s, t;
};

struct SlotSm {
    u32* summax; u32* n255; u32* s_lo; u32* s_hi; u32* t_lo; u32* t_hi;
};

__device__ __forceinline__ void run_flush(const SlotSm& S, SlotRun& r) {
    if (r.slot >= 0) {
        atomicAdd(&S.summax[r.slot], r.summax);
        if (r.n255) atomicAdd(&S.n255[r.slot], r.n255);
        atomicAdd(&S.s_lo[r.slot], (u32)(r.s & 0xffffull));
        atomicAdd(&S.s_hi[r.slot], (u32)(r.s >> 16));
        atomicAdd(&S.t_lo[r.slot], (u32)(r.t & 0xffffull));
        atomicAdd(&S.t_hi[r.slot], (u32)(r.t >> 16));
    }
    r.slot = -1; r.summax = 0; r.n255 = 0; r.s = 0; r.t = 0;
}

// Adds one pixel to its parent slot: t = wrap(h + off) as calculate_avg_hsv does (color_quantization.c:538-548).
// The wrap decision is discrete; within 0.01 degree of the 0/360 seam it is taken from the exact FP64 replay.
__device__ __forceinline__ void run_add(const SlotSm& S, SlotRun& r, int slot, long long off_q, double off_d,
                                        const FastPx& px, const PalTablesView& tb, int R, int G, int B,
                                        const double* __restrict__ k255) {
    if (slot != r.slot) {
        run_flush(S, r);
        r.slot = slot;
    }
    long long t = phd_hue_q22(px, tb) + off_q;
    const long long eps = (1ll << PHD_HQ_SHIFT) / 100;
    const long long d360 = t - PHD_HQ_360;
    if ((d360 > -eps && d360 < eps) || (t > -eps && t < eps)) {
        const HsvD e = phd_hsv_exact(R, G, B, k255);
        const double te = __dadd_rn(e.h, off_d);
        if (te > 360.0) t -= PHD_HQ_360;
        else if (te < 0.0) t += PHD_HQ_360;
        if (t < 0) t = 0;  // the fixed-point value may sit a hair on the other side of the seam
    } else if (t > PHD_HQ_360) t -= PHD_HQ_360;
    else if (t < 0) t += PHD_HQ_360;
    r.t += (u64)t;
    r.s += phd_sat_q30(px, tb);
    r.summax += (u32)px.mx;
    r.n255 += (px.mx == 255);
}

__global__ void __launch_bounds__(PHD_FE_THREADS) k_palette_accumulate(
    const uint8_t* __restrict__ rgb, DevParams P, const double* __restrict__ centres,
    const unsigned char* __restrict__ pal_tables, const GroupPlan* __restrict__ plan_g,
    const int* __restrict__ pal_n, const int* __restrict__ parent_ids, const int* __restrict__ tie_list,
    const int* __restrict__ tie_n, SlotAcc* __restrict__ sacc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = P.T;
    double* k255 = reinterpret_cast<double*>(smem_raw);
    unsigned char* tb_raw = smem_raw + 256 * sizeof(double);    // 16-byte aligned, size multiple of 16
    double* off = reinterpret_cast<double*>(tb_raw + phd_pal_tables_bytes(P.sp));  // [T]
    long long* off_q = reinterpret_cast<long long*>(off + T);   // [T]
    GroupPlan* plan = reinterpret_cast<GroupPlan*>(off_q + T);  // [T]
    u32* acc = reinterpret_cast<u32*>(plan + T);                // [6][T]
    u16* gid_cache = reinterpret_cast<u16*>(acc + 6 * T);       // [PHD_CHUNK]
    __shared__ int scan[PHD_FE_THREADS];
    __shared__ int sh_last;

    const int img = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    const int N = pal_n[img];
    phd_fill_k255(k255);
    phd_pal_tables_to_smem(tb_raw, pal_tables, P.sp);
    for (int g = tid; g < T; g += blockDim.x) plan[g] = plan_g[(size_t)img * T + g];
    for (int j = tid; j < N; j += blockDim.x) {
        const double o = __dsub_rn(180.0, centres[parent_ids[(size_t)img * T + j]]);
        off[j] = o;
        off_q[j] = __double2ll_rn(o * (double)(1 << PHD_HQ_SHIFT));
    }
    for (int i = tid; i < 6 * T; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    const PalTablesView tb = phd_pal_tables_view(tb_raw);
    const FastCfg C = phd_fast_cfg(P);
    SlotSm S{acc, acc + T, acc + 2 * T, acc + 3 * T, acc + 4 * T, acc + 5 * T};
    SlotRun run{-1, 0, 0, 0, 0};

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
            const FastPx px = phd_group_fast(R, G, B, tb, C, k255);
            gid = px.gid;
            const GroupPlan gp = plan[gid];
            if (gp.mode == 1 || (gp.mode == 2 && chunk < gp.cstar))
                run_add(S, run, gp.slot, off_q[gp.slot], off[gp.slot], px, tb, R, G, B, k255);
        }
        gid_cache[tid * PHD_PX_PER_THREAD + i] = (u16)gid;
    }
    run_flush(S, run);
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
        if (tid == 0) {
            int acc_run = 0;
            for (int t = 0; t < PHD_FE_THREADS; t++) { int c = scan[t]; scan[t] = acc_run; acc_run += c; }
        }
        if (my_last >= 0) atomicMax(&sh_last, my_last);
        __syncthreads();
        int rank = scan[tid];
        const int last_idx = sh_last;
        for (int i = 0; i < PHD_PX_PER_THREAD; i++) {
            const int li = tid * PHD_PX_PER_THREAD + i;
            if (gid_cache[li] != g) continue;
            const bool take = (partial && rank < gp.need) || (last && li == last_idx);
            rank++;
            if (!take) continue;
            const uint8_t* q = base + phd_src_index((long long)chunk * PHD_CHUNK + li, P) * 3;
            const int R = __ldg(q), G = __ldg(q + 1), B = __ldg(q + 2);
            const FastPx px = phd_group_fast(R, G, B, tb, C, k255);
            run_add(S, run, gp.slot, off_q[gp.slot], off[gp.slot], px, tb, R, G, B, k255);
        }
        run_flush(S, run);
        __syncthreads();
    }
    __syncthreads();

    for (int j = tid; j < N; j += blockDim.x) {
        const u64 sm = S.summax[j];
        const u64 sv = ((u64)S.s_hi[j] << 16) + S.s_lo[j];
        const u64 tv = ((u64)S.t_hi[j] << 16) + S.t_lo[j];
        if (!(sm | sv | tv)) continue;
        SlotAcc* a = sacc + (size_t)img * T + j;
        atomicAdd(&a->summax, sm);
        if (S.n255[j]) atomicAdd(&a->n255, (u64)S.n255[j]);
        if (sv) atomicAdd(&a->s_sum, sv);
        if (tv) atomicAdd(&a->t_sum, tv);
    }
}

template <bool FAST>
__global__ void __launch_bounds__(256) k_group_sweep(DevParams P, const unsigned char* __restrict__ pal_tables,
                                                     u16* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* k255 = reinterpret_cast<double*>(smem_raw);
    unsigned char* tb_raw = smem_raw + 256 * sizeof(double);
    phd_fill_k255(k255);
    if (FAST) phd_pal_tables_to_smem(tb_raw, pal_tables, P.sp);
    __syncthreads();
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;  // r<<16 | g<<8 | b
    const int R = (c >> 16) & 255, G = (c >> 8) & 255, B = c & 255;
    if (FAST) {
        const PalTablesView tb = phd_pal_tables_view(tb_raw);
        const FastCfg C = phd_fast_cfg(P);
        out[c] = (u16)phd_group_fast(R, G, B, tb, C, k255).gid;
    } else {
        const HsvD px = phd_hsv_exact(R, G, B, k255);
        out[c] = (u16)phd_group_exact(px, P);
    }
}

// Builds the per-parameter tables with the exact arithmetic; one thread per max value.
__global__ void __launch_bounds__(256) k_build_pal_tables(DevParams P, unsigned char* __restrict__ out,
                                                          int* __restrict__ ok) {
    __shared__ double k255[256];
    phd_fill_k255(k255);
    __syncthreads();
    const int m = threadIdx.x;
    unsigned char* vtab = out;
    u64* rs = reinterpret_cast<u64*>(out + 256);
    u64* rh = reinterpret_cast<u64*>(out + 256 + 2048);
    u32* rd = reinterpret_cast<u32*>(out + 256 + 4096);
    u16* sthr = reinterpret_cast<u16*>(out + 256 + 4096 + 1024);
    const int spw = phd_spw(P.sp);
    // value bin / black (rgb2hsv :408, arm_octree :129,141)
    const double v = (m == 255) ? 0.999999 : k255[m];
    int vi = 0xFF;
    if (!(v < P.bt)) vi = min(max((int)__ddiv_rn(__dsub_rn(v, P.bt), P.Lv), 0), 254);
    vtab[m] = (unsigned char)vi;
    rs[m] = m ? ((1ull << 46) + (u64)(m / 2)) / (u64)m : 0ull;
    rh[m] = m ? ((60ull << 38) + (u64)(m / 2)) / (u64)m : 0ull;
    {
        const u64 den = (u64)(int)P.Lh * (u64)m;
        rd[m] = den ? (u32)(((1ull << 32) + den - 1) / den) : 0u;
    }
    // saturation classes along min = 0..m: gray (-1) or Si, must be non-increasing
    u16* th = sthr + m * spw;
    for (int j = 0; j < spw; j++) th[j] = j < P.sp ? 0xFFFF : 0;
    int prev = 1 << 30;
    bool mono = true;
    for (int mn = 0; mn <= m; mn++) {
        double s;
        if (m == 0) s = 0.0;
        else if (mn == 0) s = 0.999999;
        else s = __ddiv_rn(__dsub_rn(k255[m], k255[mn]), k255[m]);
        int c = -1;
        if (!(s < P.gt)) c = (int)__ddiv_rn(__dsub_rn(s, P.gt), P.Ls);
        if (c > prev || c >= P.sp) mono = false;
        // first min at which the class drops below j
        for (int j = 0; j < P.sp; j++)
            if (c < j && th[j] == 0xFFFF) th[j] = (u16)mn;
        prev = c;
    }
    if (!mono) atomicAnd(ok, 0);
}

}  // namespace

void phd_launch_frontend(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* pal_tables,
                         Workspace& ws, cudaStream_t st, int* launches) {
    dim3 grid(P.nchunks, nimg);
    const size_t smem = 256 * sizeof(double) + phd_pal_tables_bytes(P.sp) + (size_t)P.T * sizeof(u32);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_frontend<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_frontend<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    if (P.ds <= 1) {
        k_frontend<true><<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, pal_tables, ws.counts_chunk, ws.hist, ws.iacc);
        *launches += 1;
    } else {
        k_frontend<false><<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, pal_tables, ws.counts_chunk, ws.hist, ws.iacc);
        int blocks = (int)((P.npx + 256LL * 16 - 1) / (256LL * 16));
        if (blocks < 1) blocks = 1;
        k_rgb_stats<<<dim3(blocks, nimg), 256, 0, st>>>(rgb, P, ws.iacc);
        *launches += 2;
    }
}

void phd_launch_palette_accumulate(const uint8_t* rgb, const DevParams& P, int nimg, const double* centres,
                                   const unsigned char* pal_tables, Workspace& ws, cudaStream_t st, int* launches) {
    dim3 grid(P.nchunks, nimg);
    const size_t smem = 256 * sizeof(double) +
                        (size_t)P.T * (sizeof(double) + sizeof(long long) + sizeof(GroupPlan) + 6 * sizeof(u32)) +
                        PHD_CHUNK * sizeof(u16) + phd_pal_tables_bytes(P.sp);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_palette_accumulate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    k_palette_accumulate<<<grid, PHD_FE_THREADS, smem, st>>>(rgb, P, centres, pal_tables, ws.plan, ws.pal_n,
                                                             ws.parent_ids, ws.tie_list, ws.tie_n, ws.sacc);
    *launches += 1;
}

void phd_launch_group_sweep(const DevParams& P, const unsigned char* pal_tables, bool fast, u16* out_dev,
                            cudaStream_t st) {
    const size_t smem = 256 * sizeof(double) + phd_pal_tables_bytes(P.sp);
    if (fast) k_group_sweep<true><<<(1 << 24) / 256, 256, smem, st>>>(P, pal_tables, out_dev);
    else k_group_sweep<false><<<(1 << 24) / 256, 256, smem, st>>>(P, pal_tables, out_dev);
}

void phd_launch_build_pal_tables(const DevParams& P, unsigned char* tables_dev, int* ok_dev, cudaStream_t st) {
    k_build_pal_tables<<<1, 256, 0, st>>>(P, tables_dev, ok_dev);
}

size_t phd_pal_tables_size(int sp) { return phd_pal_tables_bytes(sp); }
