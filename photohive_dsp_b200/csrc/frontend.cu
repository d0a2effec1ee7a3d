// Pixel front end: ONE pass over the packed RGB bytes produces everything the palette, the saturation mean and
// the channel statistics need.
//
//   k_pixels          replaces downsample_rgb / rgb2hsv / get_rgb_statistics / get_hsv_average, arm_octree and the
//                     summing of calculate_avg_hsv (src/image_processing.c:344-417,533-553,
//                     src/color_quantization.c:108-161,510-576).  Every pixel is classified into a CELL
//                     (pixel_cells.cuh) and its count / max / saturation / hue contributions are added to that
//                     cell with shared-memory integer atomics; per-parent sums for whatever parents are selected
//                     later follow from the cells (palette_select.cu), so no second pass over the image is needed.
//   k_palette_ties    the one part of group_irregular_pixels (src/color_quantization.c:411-451) that depends on
//                     raster order: tied groups keep "the first `room` pixels + the last pixel".  Only the few
//                     chunks that hold those pixels are revisited (work list from palette_select).
//   k_rgb_stats       channel sums over the FULL image when the HSV image is downsampled (src/interface.c:50-55).
//   k_build_cell_tables / k_build_exc   per-parameter tables, computed with the reference's double arithmetic.
//   k_group_sweep     test hook: group id of all 2^24 colours.
//
// All accumulators are integers (fixed point where needed) so results do not depend on scheduling.
#include "frontend_walk.cuh"

namespace {

template <int THREADS, bool DS, int NCS, bool DB>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1) k_pixels(const uint8_t* __restrict__ rgb, DevParams P,
                                                    const unsigned char* __restrict__ tabs_g,
                                                    const unsigned char* __restrict__ exc,
                                                    u16* __restrict__ counts_chunk, u64* __restrict__ cells_g,
                                                    u32* __restrict__ span32, u64* __restrict__ span64,
                                                    ImageAcc* __restrict__ iacc) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int c_begin = blockIdx.x * P.cpp, c_end = min(c_begin + P.cpp, P.nchunks);
    fe_walk<THREADS, DS, NCS, DB>(smem_raw, rgb, P, tabs_g, exc, blockIdx.y, blockIdx.x, c_begin, c_end, true,
                                  counts_chunk, cells_g, span32, span64, iacc);
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
// Tie path.  Work item = (image, chunk) holding pixels of a tied group that are only PARTLY accepted.  The chunk
// is classified again; accepted pixels (chunks before the group's c* entirely, the first `need` pixels of chunk
// c* in raster order, and the group's very last pixel) are added to the image's tie cells, which finalize folds
// into the parents with the same cell arithmetic as everything else.
__global__ void __launch_bounds__(256) k_palette_ties(const uint8_t* __restrict__ rgb, DevParams P,
                                                      const unsigned char* __restrict__ tabs_g,
                                                      const unsigned char* __restrict__ exc,
                                                      const GroupPlan* __restrict__ plan_g,
                                                      const int* __restrict__ tie_list, const int* __restrict__ tie_n,
                                                      const u32* __restrict__ work, const u32* __restrict__ work_n,
                                                      u64* __restrict__ cells_tie) {
    constexpr int MAXT = 64;  // tie groups of one image whose plans are cached in shared memory
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* tb_raw = smem_raw;
    u16* tie_cache = reinterpret_cast<u16*>(smem_raw + phd_cell_tables_bytes());  // [chunk] tie index of each pixel
    u16* tie_of_pair = tie_cache + P.chunk;                                       // [ncls*hp] (cls, hue bin) -> tie index
    unsigned char* cls_has_tie = reinterpret_cast<unsigned char*>(tie_of_pair + P.ncls * P.hp);  // [ncls]
    constexpr u16 NONE = 0xffff;
    __shared__ GroupPlan tplan[MAXT];
    __shared__ int scan[256];
    __shared__ int sh_last;
    const int tid = threadIdx.x, T = P.T, NC = P.NC;
    const int ppt = P.chunk / 256;  // consecutive pixels per thread (16 or 32)
    const int npairs = P.ncls * P.hp, spvp = P.sp * P.vp;
    const u32 n_items = *work_n;
    if (blockIdx.x >= n_items) return;
    exc = phd_exc_biased(exc);
    phd_cell_tabs_to_smem(tb_raw, tabs_g);
    const unsigned char* svtab = tb_raw;
    const int qs = 20;  // tie cells are kept in Q20 like the global cells
    const CellCfg K = phd_cell_cfg(P, qs);
    const bool fast_ok = P.ds <= 1 && P.aligned16 != 0;

    auto accept = [&](u64* ct, const PixOut& o) {
        atomicAdd(ct + o.cell, 1ull);
        if (o.w0 >> 16) atomicAdd(ct + NC + o.cell, 1ull);
        atomicAdd(ct + 2 * NC + o.cell, (u64)o.mx);
        atomicAdd(ct + 3 * NC + o.cell, (u64)(o.sbits - PHD_MAGIC_RN_BITS) << (20 - qs));
        atomicAdd(ct + 4 * NC + o.cell, (u64)(o.hbits - PHD_MAGIC_RN_BITS) << (20 - qs));
    };

    int cur_img = -1, nt = 0;
    for (u32 item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int img = (int)(work[item] / (u32)P.nchunks), chunk = (int)(work[item] % (u32)P.nchunks);
        const uint8_t* base = rgb + (size_t)img * P.image_stride;
        u64* ct = cells_tie + (size_t)img * PHD_CELL_Q * NC;
        const long long c0 = (long long)chunk * P.chunk;
        __syncthreads();  // the previous item's caches are no longer read
        if (img != cur_img) {
            // (cls, hue bin) -> index of its partly accepted tie group in this image's tie list
            cur_img = img;
            nt = tie_n[img];
            for (int i = tid; i < npairs; i += 256) tie_of_pair[i] = NONE;
            for (int i = tid; i < P.ncls; i += 256) cls_has_tie[i] = 0;
            for (int k = tid; k < min(nt, MAXT); k += 256) tplan[k] = plan_g[(size_t)img * T + tie_list[(size_t)img * T + k]];
            __syncthreads();
            for (int k = 0; k < nt; k++) {
                const int g = tie_list[(size_t)img * T + k];
                if (g < P.hp * spvp) {
                    if (tid == 0) {
                        const int j = g / spvp, cls = g - j * spvp;
                        tie_of_pair[cls * P.hp + j] = (u16)k;
                        cls_has_tie[cls] = 1;
                    }
                } else {
                    const int cls = (g == T - 1) ? spvp + 1 : spvp;
                    for (int j = tid; j < P.hp; j += 256) tie_of_pair[cls * P.hp + j] = (u16)k;
                    if (tid == 0) cls_has_tie[cls] = 1;
                }
            }
            __syncthreads();
        }
        auto get_plan = [&](int k) -> GroupPlan {
            return k < MAXT ? tplan[k] : plan_g[(size_t)img * T + tie_list[(size_t)img * T + k]];
        };
        // classify the chunk; pixels of tie groups whose accepted prefix covers the whole chunk are taken at once
        const long long p0 = c0 + (long long)tid * ppt;
        for (int i0 = 0; i0 < ppt; i0 += 16) {
            u32 w[12];
            const bool vec = fast_ok && p0 + i0 + 16 <= P.hpx;
            if (vec) load48_aligned(base + (p0 + i0) * 3, w);
#pragma unroll 4
            for (int i = 0; i < 16; i++) {
                const int li = tid * ppt + i0 + i;
                u16 t = NONE;
                if (c0 + li < P.hpx) {
                    int R, G, B;
                    if (vec) { R = packed_byte(w, 3 * i); G = packed_byte(w, 3 * i + 1); B = packed_byte(w, 3 * i + 2); }
                    else {
                        const uint8_t* q = base + phd_src_index(c0 + li, P) * 3;
                        R = __ldg(q); G = __ldg(q + 1); B = __ldg(q + 2);
                    }
                    // most pixels are ruled out by their saturation/value class alone: (max, min) -> class
                    const int mx = max(R, max(G, B)), mn = min(R, min(G, B));
                    if (cls_has_tie[svtab[((mx * mx + mx) >> 1) + mn]]) {
                        const PixOut o = phd_pixel(R, G, B, svtab, K, exc);
                        t = tie_of_pair[o.cell >> 2];
                        if (t != NONE) {
                            // whole chunk accepted -- unless it belongs to a span before c*'s, whose sums
                            // palette_select has already folded in (this chunk is here for another tie group)
                            const int cs = get_plan(t).cstar;
                            if (chunk < cs && chunk >= (cs / P.cpp) * P.cpp) accept(ct, o);
                        }
                    }
                }
                tie_cache[li] = t;
            }
        }
        __syncthreads();
        for (int k = 0; k < nt; k++) {
            const GroupPlan gp = get_plan(k);
            const bool partial = (gp.cstar == chunk && gp.need > 0);
            const bool last = (gp.clast == chunk);
            if (!partial && !last) continue;  // uniform across the block
            int mine = 0, my_last = -1;
            for (int i = 0; i < ppt; i++)
                if (tie_cache[tid * ppt + i] == k) { mine++; my_last = tid * ppt + i; }
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
                if (tie_cache[li] != k) continue;
                const bool take = (partial && rank < gp.need) || (last && li == last_idx);
                rank++;
                if (take) {
                    const uint8_t* q = base + phd_src_index(c0 + li, P) * 3;
                    accept(ct, phd_pixel(__ldg(q), __ldg(q + 1), __ldg(q + 2), svtab, K, exc));
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
template <bool FAST>
__global__ void __launch_bounds__(256) k_group_sweep(DevParams P, const unsigned char* __restrict__ tabs_g,
                                                     const unsigned char* __restrict__ exc, u16* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* k255 = reinterpret_cast<double*>(smem_raw);
    unsigned char* tb_raw = smem_raw + 256 * sizeof(double);
    phd_fill_k255(k255);
    if (FAST) phd_cell_tabs_to_smem(tb_raw, tabs_g);
    __syncthreads();
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;  // r<<16 | g<<8 | b
    const int R = (c >> 16) & 255, G = (c >> 8) & 255, B = c & 255;
    if (FAST) {
        const CellCfg K = phd_cell_cfg(P, 20);
        const PixOut o = phd_pixel(R, G, B, tb_raw, K, phd_exc_biased(exc));
        out[c] = (u16)phd_cell_group(o.cell, P);
    } else {
        const HsvD px = phd_hsv_exact(R, G, B, k255);
        out[c] = (u16)phd_group_exact(px, P);
    }
}

// Per-parameter tables with the reference's arithmetic (pixel_cells.cuh); one thread per max value.
__global__ void __launch_bounds__(256) k_build_cell_tables(DevParams P, unsigned char* __restrict__ out,
                                                           int* __restrict__ ok) {
    __shared__ double k255[256];
    phd_fill_k255(k255);
    __syncthreads();
    const int m = threadIdx.x;
    unsigned char* svtab = out;
    const int spvp = P.sp * P.vp;
    // value bin / black (rgb2hsv :408, arm_octree :129,141)
    const double v = (m == 255) ? 0.999999 : k255[m];
    const bool black = v < P.bt;
    const int vi_raw = black ? 0 : (int)__ddiv_rn(__dsub_rn(v, P.bt), P.Lv);
    if (vi_raw < 0 || vi_raw >= P.vp) atomicAnd(ok, 0);  // the reference would index past its grid
    const int vi = min(max(vi_raw, 0), P.vp - 1);
    for (int mn = 0; mn <= m; mn++) {
        double s;
        if (m == 0) s = 0.0;
        else if (mn == 0) s = 0.999999;
        else s = __ddiv_rn(__dsub_rn(k255[m], k255[mn]), k255[m]);
        int cls;
        if (black) cls = spvp + 1;
        else if (s < P.gt) cls = spvp;
        else {
            const int si = (int)__ddiv_rn(__dsub_rn(s, P.gt), P.Ls);
            if (si < 0 || si >= P.sp) atomicAnd(ok, 0);
            cls = min(max(si, 0), P.sp - 1) * P.vp + vi;
        }
        svtab[((m * m + m) >> 1) + mn] = (unsigned char)cls;
        // second table (three-word front end): pixels with max == 255 get the TWIN class of their class -- ncls + si
        // for the top value bin of saturation bin si, ncls + sp for gray; black cannot occur there (bt <= 0.999999
        // is a precondition of that kernel variant) and keeps its class
        int cls3 = cls;
        if (m == 255 && cls != spvp + 1) cls3 = P.ncls + (cls == spvp ? P.sp : cls / P.vp);
        svtab[PHD_TRI_SIZE + ((m * m + m) >> 1) + mn] = (unsigned char)cls3;
    }
}

// Exceptional-colour table (pixel_cells.cuh): for every colour whose hue is exactly k * Lh/2, what the reference's
// doubles make of it, stored at entry (tri(max) + min, k) as (cell delta relative to the ordinary cell cls*4hp + 2k + 1,
// full).  Every exceptional colour has its own entry ((max, min, k) determine the colour); the others stay zero.
__global__ void __launch_bounds__(256) k_build_exc(DevParams P, unsigned char* __restrict__ out, int* __restrict__ ok) {
    __shared__ double k255[256];
    phd_fill_k255(k255);
    __syncthreads();
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;  // R | G << 8 | B << 16
    const int R = c & 255, G = (c >> 8) & 255, B = (c >> 16) & 255;
    const int mx = max(R, max(G, B)), mn = min(R, min(G, B)), q = mx - mn;
    if (q != 0) {
        const bool isR = (R == mx), isG = (G == mx);
        const int p = isR ? (G - B) : (isG ? (B - R) : (R - G));
        int num2 = 120 * ((isR ? 0 : (isG ? 2 : 4)) * q + p);
        if (num2 < 0) num2 += 720 * q;
        const int Lhi = (int)P.Lh, den = Lhi * q;
        if (num2 % den == 0) {
            const int k = num2 / den;                   // the pixel sits exactly on b = k * Lh/2
            const double b = (double)k * (P.Lh * 0.5);  // exact (Lh is a small integer)
            const HsvD e = phd_hsv_exact(R, G, B, k255);
            const int hi_ref = (int)__ddiv_rn(e.h, P.Lh);
            // side of a wrap seam at b (calculate_avg_hsv :538-548): t = h + off with b + off == 0 or 360
            bool below;
            if (b < 180.0) below = __dadd_rn(e.h, -b) < 0.0;               // t < 0 -> t + 360 (ends near 360)
            else below = !(__dadd_rn(e.h, __dsub_rn(360.0, b)) > 360.0);   // not wrapped (stays near 360)
            // chunk-index delta (pixel_cells.cuh, CHUNK INDEX) relative to the ordinary ci = cls*4hp + k
            const int to_rare = 2 * P.hp;  // same hue bin, from sub 1 (k even) to sub 0; one less: sub 2 of the bin below
            int delta, full = 0;
            if ((k & 1) == 0) {
                const int dj = hi_ref - (k >> 1);
                if (dj == 0) delta = below ? to_rare : 0;  // (j, on the lower edge, low side) | (j, lower half)
                else if (dj == -1 && k > 0) {              // (j-1, end of the upper half) | (j-1, on the upper edge)
                    delta = below ? -1 : to_rare - 1;
                    full = below ? 1 : 0;
                } else {
                    delta = 0;
                    atomicAnd(ok, 0);
                }
            } else {
                if (hi_ref != (k >> 1)) atomicAnd(ok, 0);
                delta = below ? -1 : 0;  // (j, end of the lower half) | (j, upper half)
                full = below ? 1 : 0;
            }
            unsigned char* ent = out + PHD_EXC_ENTRY * ((size_t)(((mx * mx + mx) >> 1) + mn) * (2 * P.hp) + k);
            *reinterpret_cast<short*>(ent) = (short)delta;
            if (full != (delta == -1 ? 1 : 0)) atomicAnd(ok, 0);  // the table relies on "end of the half bin <=> delta == -1"
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------

size_t phd_pixels_smem(const DevParams& P) {
    const size_t ncs = (P.fe_threads == 256) ? PHD_NCS_SMALL : (size_t)P.NC + 32;
    return phd_cell_tables_bytes() + (P.fe_threads == 256 ? 6 : 4) * ncs * sizeof(u32) + (size_t)P.NC * (2 * sizeof(u64) + 3 * sizeof(u32));
}

void phd_launch_pixels(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                       const unsigned char* exc, Workspace& ws, cudaStream_t st, int* launches) {
    const size_t smem = phd_pixels_smem(P);
    PHD_ALLOW_SMEM((k_pixels<256, false, PHD_NCS_SMALL, true>), 200 * 1024);
    PHD_ALLOW_SMEM((k_pixels<256, true, PHD_NCS_SMALL, true>), 200 * 1024);
    PHD_ALLOW_SMEM((k_pixels<512, false, 0, false>), 200 * 1024);
    PHD_ALLOW_SMEM((k_pixels<512, true, 0, false>), 200 * 1024);
    dim3 grid(P.nspans, nimg);  // one walk of P.cpp chunks per CTA (phd_fe_plan)
    const bool ds = P.ds > 1;
    if (P.fe_threads == 256) {
        if (ds) k_pixels<256, true, PHD_NCS_SMALL, true><<<grid, 256, smem, st>>>(rgb, P, tabs, exc, ws.counts_chunk, ws.cells, ws.span32, ws.span64, ws.iacc);
        else k_pixels<256, false, PHD_NCS_SMALL, true><<<grid, 256, smem, st>>>(rgb, P, tabs, exc, ws.counts_chunk, ws.cells, ws.span32, ws.span64, ws.iacc);
    } else {
        if (ds) k_pixels<512, true, 0, false><<<grid, 512, smem, st>>>(rgb, P, tabs, exc, ws.counts_chunk, ws.cells, ws.span32, ws.span64, ws.iacc);
        else k_pixels<512, false, 0, false><<<grid, 512, smem, st>>>(rgb, P, tabs, exc, ws.counts_chunk, ws.cells, ws.span32, ws.span64, ws.iacc);
    }
    *launches += 1;
    if (ds) {
        int blocks = (int)((P.npx + 256LL * 16 - 1) / (256LL * 16));
        if (blocks < 1) blocks = 1;
        k_rgb_stats<<<dim3(blocks, nimg), 256, 0, st>>>(rgb, P, ws.iacc);
        *launches += 1;
    }
}

void phd_launch_palette_ties(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                             const unsigned char* exc, Workspace& ws, cudaStream_t st, int* launches) {
    const size_t smem = phd_cell_tables_bytes() + ((size_t)P.chunk + (size_t)P.ncls * P.hp) * sizeof(u16) + (size_t)P.ncls + 16;
    PHD_ALLOW_SMEM((k_palette_ties), 200 * 1024);
    // CTAs stride over the (image, chunk) work list, whose length only the device knows; CTAs beyond it retire at once.
    // Small batches get up to 128 CTAs per image so that a single image's chunks are not walked by four CTAs.
    long long want = (long long)nimg * (P.nchunks < 128 ? P.nchunks : 128);
    int grid = (int)(want > 148 * 4 ? 148 * 4 : want);
    k_palette_ties<<<grid, 256, smem, st>>>(rgb, P, tabs, exc, ws.plan, ws.tie_list, ws.tie_n, ws.work, ws.work_n,
                                            ws.cells_tie);
    *launches += 1;
}

void phd_launch_group_sweep(const DevParams& P, const unsigned char* tabs, const unsigned char* exc, bool fast,
                            u16* out_dev, cudaStream_t st) {
    const size_t smem = 256 * sizeof(double) + phd_cell_tables_bytes();
    PHD_ALLOW_SMEM((k_group_sweep<true>), 100 * 1024);
    if (fast) k_group_sweep<true><<<(1 << 24) / 256, 256, smem, st>>>(P, tabs, exc, out_dev);
    else k_group_sweep<false><<<(1 << 24) / 256, 256, 256 * sizeof(double), st>>>(P, tabs, exc, out_dev);
}

// exc_dev == nullptr: the exceptional-colour codes of this h_partitions already exist (they depend on nothing else)
void phd_launch_build_cell_tables(const DevParams& P, unsigned char* tables_dev, unsigned char* exc_dev, int* ok_dev,
                                  cudaStream_t st) {
    k_build_cell_tables<<<1, 256, 0, st>>>(P, tables_dev, ok_dev);
    if (exc_dev) {
        cudaMemsetAsync(exc_dev, 0, phd_exc_bytes(P.hp), st);
        k_build_exc<<<(1 << 24) / 256, 256, 0, st>>>(P, exc_dev, ok_dev);
    }
}

size_t phd_exc_table_bytes(int hp) { return phd_exc_bytes(hp); }

size_t phd_cell_tables_size() { return 2 * phd_cell_tables_bytes(); }  // ordinary classes, then the table with twin classes
