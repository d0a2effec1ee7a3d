// Palette parent selection and merge plan, one CTA per image, entirely on the device so a batch
// needs no host round trip.
//
// Replaces find_valid_octree_parents (src/color_quantization.c:174-203) with saliency :588-595,
// compare_quantities :601-611 and custom_sort (src/utilities.c:132-153), and the decision part of
// group_irregular_pixels (src/color_quantization.c:342-479) with get_node_distance_heuristic
// :253-288.  The comparator is not a total order (float difference truncated to int, x86
// cvttss2si overflow -> INT_MIN), so the reference's insertion sort is reproduced step by step
// rather than replaced by "a sort".  The tie path follows the behaviour of the shipped -O0
// binary: the first tied parent always wins (get_distance_pixel_to_parent has no return value,
// :303-311) and pixels beyond the tail node's free room are dropped except the last (:435-440).
#include <limits.h>

#include "cell_reduce.cuh"

namespace {

__device__ __forceinline__ int trunc_f2i_x86(float d) {
    if (!(d > -2147483904.0f && d < 2147483648.0f)) return INT_MIN;
    return (int)d;
}

__device__ double centre_dist(const double* gh, const double* gs, const double* gv, int T, int vp, int a, int p) {
    const int gray_start = T - (vp + 1), black = T - 1;
    if (a < gray_start && p < gray_start) {
        double hd = fabs(__dsub_rn(gh[a], gh[p]));
        if (hd > 180.0) hd = __dsub_rn(360.0, hd);
        hd = __dmul_rn(hd, (1.0) / (360.0));
        const double sd = __dsub_rn(gs[a], gs[p]), vd = __dsub_rn(gv[a], gv[p]);
        return __dadd_rn(__dadd_rn(__dmul_rn(hd, hd), __dmul_rn(sd, sd)), __dmul_rn(vd, vd));
    }
    if ((gray_start <= a && a < black && p < gray_start) || (gray_start <= p && p < black && a < gray_start)) {
        const double sd = __dsub_rn(gs[a], gs[p]), vd = __dsub_rn(gv[a], gv[p]);
        return __dadd_rn(__dmul_rn(sd, sd), __dmul_rn(vd, vd));
    }
    const double vd = __dsub_rn(gv[a], gv[p]);
    return __dmul_rn(vd, vd);
}

__global__ void __launch_bounds__(256) k_palette_select(DevParams P, const double* __restrict__ centres,
                                                        const float* __restrict__ sv_f,
                                                        const u64* __restrict__ cells_g,
                                                        const u16* __restrict__ counts_chunk,
                                                        GroupPlan* __restrict__ plan_g, int* __restrict__ pal_n,
                                                        int* __restrict__ parent_ids, int* __restrict__ tie_list,
                                                        int* __restrict__ tie_n, int* __restrict__ tie_groups,
                                                        long long* __restrict__ dropped, SlotAcc* __restrict__ sacc,
                                                        u32* __restrict__ hist, ImageAcc* __restrict__ iacc,
                                                        u64* __restrict__ cells_tie_g,
                                                        const u32* __restrict__ span32,
                                                        const u64* __restrict__ span64, u32* __restrict__ work,
                                                        u32* __restrict__ work_n, const int from_hist) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int T = P.T, tid = threadIdx.x, img = blockIdx.x;
    int* n = reinterpret_cast<int*>(smem_raw);       // [T] pixel counts
    float* sal = reinterpret_cast<float*>(n + T);    // [T]
    int* ids = reinterpret_cast<int*>(sal + T);      // [T] sorted group ids (parents first)
    int* slot = ids + T;                             // [T] group -> parent slot or -1
    int* first = slot + T;                           // [T] nearest parent slot (first of the tied ones)
    int* nmin = first + T;                           // [T] number of equally near parents
    int* fill = nmin + T;                            // [T] per parent slot: occupancy of its tail node
    int* pend = fill + T;                            // [T] per parent slot: group whose last pixel is pending
    int* take = pend + T;                            // [T] tie groups: pixels accepted from the front
    int* keep = take + T;                            // [T] tie groups: last pixel survives
    int* scnt = keep + T;                            // [T] per parent slot: pixels that end up in it
    u32* cbits = reinterpret_cast<u32*>(scnt + T);   // [(nchunks+31)/32] chunks the tie path must revisit
    int* scan_list = nmin;                           // [T] partly accepted tie groups (nmin[] is dead by then)
    __shared__ int sh_N, sh_nt, sh_nscan;
    __shared__ int sh_wsum[32];
    __shared__ u64 sh_ssum;

    const double* gh = centres;
    const double* gs = centres + T;
    const double* gv = centres + 2 * T;
    const u64* cells = cells_g + (size_t)img * PHD_CELL_Q * P.NC;
    const int nbw = (P.nchunks + 31) / 32;

    // pixel count of every group, and the image's saturation sum, from the cells
    for (int g = tid; g < T; g += blockDim.x) n[g] = 0;
    for (int w = tid; w < nbw; w += blockDim.x) cbits[w] = 0;
    if (tid == 0) { sh_ssum = 0; sh_nscan = 0; }
    __syncthreads();
    if (from_hist) {  // general-input route: the totals come from k_f64_classify; no cells, no saturation sum here
        for (int g = tid; g < T; g += blockDim.x) n[g] = (int)hist[(size_t)img * T + g];
    } else {
        u64 ssum = 0;
        for (int pair = tid; pair < P.ncls * P.hp; pair += blockDim.x) {
            u64 c = 0;
#pragma unroll
            for (int sub = 0; sub < 4; sub++) {
                c += cells[pair * 4 + sub];
                ssum += cells[3 * (size_t)P.NC + pair * 4 + sub];
            }
            if (c) {
                const int cls = pair / P.hp, j = pair - cls * P.hp, spvp = P.sp * P.vp;
                const int g = cls < spvp ? j * spvp + cls : (cls == spvp ? T - (P.vp + 1) : T - 1);
                atomicAdd(&n[g], (int)c);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
        if ((tid & 31) == 0 && ssum) atomicAdd(&sh_ssum, ssum);
    }
    __syncthreads();
    if (tid == 0 && !from_hist) iacc[img].s_sum = sh_ssum;

    for (int g = tid; g < T; g += blockDim.x) {
        const int c = n[g];
        hist[(size_t)img * T + g] = (u32)c;
        // saliency, float arithmetic in source order (:588-595)
        const float w = __fadd_rn(P.qw, __fmul_rn(P.svw, sv_f[g]));
        sal[g] = __fmul_rn(__fmul_rn((float)c, w), 1000.0f);
        ids[g] = g;
        slot[g] = -1;
        take[g] = 0;
        keep[g] = 0;
    }
    __syncthreads();

    // Fast path: when every pair of saliencies is either equal or at least 1 apart and nothing can overflow
    // the float->int truncation, the reference's comparator is a strict weak order and its insertion sort
    // is the stable descending sort -- computed here as a parallel rank sort.
    // Differences in (0,1) but no overflow (every |value| < 2^30): the comparator sees two elements as "equal" exactly
    // when their float difference is below 1 in magnitude.  Cut the value-sorted sequence into CHAINS at the gaps >= 1:
    // an element passes every element of a lower chain and stops at every element of a higher one (float subtraction is
    // monotone), so the insertion sort's result is the chains in value order, and inside a chain whatever the insertion
    // sort makes of the chain's members alone, inserted in index order.  Only chains with a gap in (0,1) differ from
    // the rank sort, they are short, and one thread each replays them (fine palettes made most images take the serial
    // replay: 15 of 84 ms per 4096 images of BASELINE config 5).
    // Anything else (|values| >= 2^30, NaN, a chain longer than kMaxChain): one warp replays the whole insertion sort.
    int* rank = first;  // scratch, reused below
    for (int g = tid; g < T; g += blockDim.x) {
        const float sg = sal[g];
        int r = 0;
        for (int o = 0; o < T; o++) {
            const float so = sal[o];
            r += (so > sg) || (so == sg && o < g);
        }
        rank[g] = r;
    }
    __syncthreads();
    for (int g = tid; g < T; g += blockDim.x) ids[rank[g]] = g;
    __syncthreads();
    constexpr int kMaxChain = 48;
    int ok = 1, no_ovf = 1;
    for (int i = tid; i < T; i += blockDim.x) {
        const float a = sal[ids[i]];
        if (!(fabsf(a) < 1073741824.0f)) { ok = 0; no_ovf = 0; }
        if (i + 1 < T) {
            const float d = __fsub_rn(a, sal[ids[i + 1]]);
            if (!(d == 0.0f || d >= 1.0f)) ok = 0;
        }
    }
    int safe = __syncthreads_and(ok);
    if (!safe && __syncthreads_and(no_ovf)) {
        // chain repair: the thread at a chain's first position walks it, and replays it if it has a gap in (0,1)
        int* order = nmin;  // scratch [T] (dead until the nearest-parent step): the repaired ids of the chains
        int fail = 0;
        for (int i = tid; i < T; i += blockDim.x) order[i] = ids[i];
        __syncthreads();  // the chain threads below overwrite their members' entries
        for (int i = tid; i < T; i += blockDim.x) {
            if (i > 0 && !(__fsub_rn(sal[ids[i - 1]], sal[ids[i]]) >= 1.0f)) continue;  // not a chain start
            int e = i + 1;
            bool inner = false;
            while (e < T) {
                const float d = __fsub_rn(sal[ids[e - 1]], sal[ids[e]]);
                if (d >= 1.0f) break;
                inner = inner || d != 0.0f;
                e++;
            }
            if (!inner) continue;  // all equal: the rank sort's index order is what the insertion sort gives
            const int m = e - i;
            if (m > kMaxChain) { fail = 1; continue; }
            int mem[kMaxChain], out[kMaxChain];
            for (int k = 0; k < m; k++) {  // members by ascending group index = insertion order
                const int x = ids[i + k];
                int pos = k;
                while (pos > 0 && mem[pos - 1] > x) { mem[pos] = mem[pos - 1]; pos--; }
                mem[pos] = x;
            }
            for (int k = 0; k < m; k++) {
                const int x = mem[k];
                const float v = sal[x];
                int pos = k;
                while (pos > 0 && trunc_f2i_x86(__fsub_rn(sal[out[pos - 1]], v)) < 0) { out[pos] = out[pos - 1]; pos--; }
                out[pos] = x;
            }
            for (int k = 0; k < m; k++) order[i + k] = out[k];
        }
        const int any_fail = __syncthreads_or(fail);
        if (!any_fail) {
            for (int i = tid; i < T; i += blockDim.x) ids[i] = order[i];
            safe = 1;
        }
        __syncthreads();
    }
    if (!safe) {
        for (int g = tid; g < T; g += blockDim.x) ids[g] = g;
        __syncthreads();
    }

    if (!safe && tid < 32) {
        // insertion sort exactly as custom_sort walks it (utilities.c:132-153), one element at a time, but the walk of
        // an element is done 32 positions per step by a warp: the element passes predecessor p while
        // (int)(sal[p] - sal[x]) < 0 and stops at the FIRST predecessor (from the right) that does not let it pass.
        const int lane = tid;
        for (int i = 1; i < T; i++) {
            const int x = ids[i];
            const float v = sal[x];
            int pos = 0;
            for (int hi = i; hi > 0; hi -= 32) {
                const int j = hi - 1 - lane;  // lane 0 looks at the nearest predecessor
                bool blocks = false;
                if (j >= 0) blocks = !(trunc_f2i_x86(__fsub_rn(sal[ids[j]], v)) < 0);
                const unsigned m = __ballot_sync(0xffffffffu, blocks);
                if (m) { pos = hi - (__ffs(m) - 1); break; }
            }
            if (pos < i) {  // shift ids[pos .. i-1] one to the right, from the right end, then drop x into the gap
                for (int hi = i; hi > pos; hi -= 32) {
                    const int j = hi - 1 - lane;
                    int val = 0;
                    if (j >= pos) val = ids[j];
                    __syncwarp();
                    if (j >= pos) ids[j + 1] = val;
                    __syncwarp();
                }
                if (lane == 0) ids[pos] = x;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // Parents: the first N groups in sorted order whose pixel counts reach the coverage goal (:342-360: goal -= count
    // until goal <= 0).  A block-wide prefix sum over contiguous pieces of the sorted order instead of one thread's walk
    // (871 dependent shared-memory loads with the fine palette, the other seven warps waiting at the barrier).
    // coverage_thresh > 1 never reaches the goal in the reference (it then reads an unset array); every group becomes a
    // parent here instead (sh_N starts at T).
    {
        const int goal = (int)((double)P.hpx * P.coverage);
        const int per = (T + (int)blockDim.x - 1) / (int)blockDim.x;
        const int i0 = min(tid * per, T), i1 = min(i0 + per, T);
        int local = 0;
        for (int i = i0; i < i1; i++) local += n[ids[i]];
        const int lane = tid & 31, wid = tid >> 5;
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (tid == 0) sh_N = T;
        if (lane == 31) sh_wsum[wid] = incl;
        __syncthreads();
        int before = incl - local;
        for (int w = 0; w < wid; w++) before += sh_wsum[w];
        int cum = before;
        for (int i = i0; i < i1; i++) {
            cum += n[ids[i]];
            if (cum >= goal) { atomicMin(&sh_N, i + 1); break; }
        }
        __syncthreads();
    }
    const int N = sh_N;
    for (int j = tid; j < N; j += blockDim.x) {
        slot[ids[j]] = j;
        const int pn = n[ids[j]];
        fill[j] = pn > 0 ? ((pn - 1) % P.L) + 1 : 0;
        pend[j] = -1;
        scnt[j] = pn;
    }
    __syncthreads();

    // nearest parent(s) of every non-empty non-parent group (:370-392): the smallest distance, how many parents are at
    // exactly that distance, and the first of them in parent order.  One WARP per group, the parents spread over its lanes
    // (a thread per group left most threads idle -- few groups need the search, each needs all N parents in FP64).
    {
        const int lane = tid & 31, wid = tid >> 5, nwarps = (int)blockDim.x >> 5;
        for (int g = wid; g < T; g += nwarps) {
            if (n[g] == 0 || slot[g] >= 0) {
                if (lane == 0) { first[g] = -1; nmin[g] = 0; }
                continue;
            }
            double m = (double)T * (double)T;
            int cnt = 0, fi = 0x7fffffff;
            for (int j = lane; j < N; j += 32) {
                const double d = centre_dist(gh, gs, gv, T, P.vp, g, ids[j]);
                if (d < m) { m = d; cnt = 1; fi = j; }
                else if (d == m) cnt++;
            }
            // the warp's minimum; lanes that hold it contribute their counts and their first index
            double wm = m;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wm = fmin(wm, __shfl_xor_sync(0xffffffffu, wm, o));
            int c = (cnt > 0 && m == wm) ? cnt : 0, f = (cnt > 0 && m == wm) ? fi : 0x7fffffff;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c += __shfl_xor_sync(0xffffffffu, c, o);
                f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
            }
            if (lane == 0) { first[g] = c > 0 ? f : -1; nmin[g] = c; }
        }
    }
    __syncthreads();

    if (tid == 0) {
        int ties = 0, nt = 0;
        long long drop = 0;
        for (int g = 0; g < T; g++) {
            if (n[g] == 0 || slot[g] >= 0 || first[g] < 0) continue;
            const int f = first[g];
            if (nmin[g] > 1) {
                ties++;
                const int room = P.L - fill[f];
                const int tk = room < n[g] ? room : n[g];
                take[g] = tk;
                fill[f] += tk;
                scnt[f] += tk;
                if (tk < n[g]) {
                    if (pend[f] >= 0) { keep[pend[f]] = 0; drop += 1; scnt[f] -= 1; }
                    pend[f] = g;
                    keep[g] = 1;
                    scnt[f] += 1;
                    drop += n[g] - tk - 1;
                }
                if (tk < n[g]) tie_list[(size_t)img * T + nt++] = g;  // partly accepted: raster order matters
            } else {
                take[g] = -1;
                if (pend[f] >= 0) { keep[pend[f]] = 0; drop += 1; pend[f] = -1; scnt[f] -= 1; }
                fill[f] = ((n[g] - 1) % P.L) + 1;
                scnt[f] += n[g];
            }
        }
        sh_nt = nt;
        pal_n[img] = N;
        tie_n[img] = nt;
        tie_groups[img] = ties;
        dropped[img] = drop;
    }
    __syncthreads();

    // write the plan; partly accepted tie groups locate their partial / last chunk from the per-chunk counts and
    // flag the chunks the tie kernel has to revisit
    for (int g = tid; g < T; g += blockDim.x) {
        GroupPlan gp;
        gp.slot = -1; gp.mode = 0; gp.cstar = -1; gp.need = 0; gp.clast = -1;
        if (slot[g] >= 0) { gp.slot = (short)slot[g]; gp.mode = 1; }
        else if (n[g] > 0 && first[g] >= 0) {
            gp.slot = (short)first[g];
            if (take[g] < 0 || take[g] >= n[g]) gp.mode = 1;  // sole nearest parent, or a tie whose pixels all fit
            else {
                gp.mode = 2;  // c*, need and clast are filled in by the warp scan below
                scan_list[atomicAdd(&sh_nscan, 1)] = g;
            }
        }
        plan_g[(size_t)img * T + g] = gp;
        pend[g] = gp.mode;  // (pend[] and fill[] are dead: reused as the group's mode and parent slot for the fold below)
        fill[g] = gp.slot;
    }
    __syncthreads();
    // Everything that joins a parent as a whole: fold its cells into the parent's sums.  One thread per (class, hue bin)
    // pair, not per group: the gray and the black group span every hue bin, and a single thread walking their 4 * hp cells
    // with dependent global loads was most of this kernel's latency for a single image.
    if (!from_hist) {
        const int spvp = P.sp * P.vp;
        for (int pair = tid; pair < P.ncls * P.hp; pair += blockDim.x) {
            const int cls = pair / P.hp, j = pair - cls * P.hp;
            const int g = cls < spvp ? j * spvp + cls : (cls == spvp ? T - (P.vp + 1) : T - 1);
            if (pend[g] != 1) continue;
            GroupSums S{0, 0, 0, 0, 0};
            phd_reduce_pair(cells, P.NC, pair, j, P.Lh, 180.0 - gh[ids[fill[g]]], S);
            SlotAcc* a = sacc + (size_t)img * T + fill[g];
            if (S.summax) atomicAdd(&a->summax, S.summax);
            if (S.n255) atomicAdd(&a->n255, S.n255);
            if (S.s_sum) atomicAdd(&a->s_sum, S.s_sum);
            if (S.t_sum) atomicAdd(&a->t_sum, (u64)S.t_sum);
        }
    }
    for (int j = tid; j < T; j += blockDim.x) {
        parent_ids[(size_t)img * T + j] = j < N ? ids[j] : -1;
        if (j < N) atomicAdd(&sacc[(size_t)img * T + j].cnt, (u64)scnt[j]);
    }
    __syncthreads();
    // one WARP per partly accepted tie group walks the per-chunk counts (32 chunks per step): the chunk c* where the
    // accepted prefix ends, how many of its pixels are still accepted, the chunk of the group's last pixel, and the
    // chunks the tie kernel has to revisit
    {
        const int lane = tid & 31, wid = tid >> 5, nwarps = blockDim.x >> 5;
        for (int e = wid; e < sh_nscan; e += nwarps) {
            const int g = scan_list[e];
            const u16* cc = counts_chunk + (size_t)img * P.nchunks * T + g;
            const int tk = take[g];
            int cum = 0, cstar = -1, need = 0, clast = -1;
            bool found = (tk == 0);
            // pass 1: the chunk c* where the accepted prefix of `tk` pixels ends, how many pixels of c* are still
            // accepted, and the chunk of the group's last pixel
            for (int base = 0; base < P.nchunks; base += 32) {
                const int c = base + lane;
                const int k = c < P.nchunks ? (int)cc[(size_t)c * T] : 0;
                int incl = k;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                const unsigned nz = __ballot_sync(0xffffffffu, k > 0);
                if (nz) clast = base + 31 - __clz(nz);
                if (!found) {
                    const int before = cum + incl - k;
                    const unsigned hit = __ballot_sync(0xffffffffu, k > 0 && cum + incl >= tk);
                    if (hit) {
                        const int L = __ffs(hit) - 1;
                        cstar = base + L;
                        need = tk - __shfl_sync(0xffffffffu, before, L);
                        found = true;
                    }
                    cum += __shfl_sync(0xffffffffu, incl, 31);
                }
            }
            if (!keep[g]) clast = -1;
            // Every front-end walk (span of P.cpp chunks) that ends before c* is accepted as a whole: its cells come
            // from the walk's stored sums.  Only the chunks of c*'s own span up to c* (and the chunk of the last
            // pixel) are looked at again by the tie kernel.
            const int sstar = cstar >= 0 ? cstar / P.cpp : 0;
            const int c_first = sstar * P.cpp;
            if (cstar >= 0)
                for (int c = c_first + lane; c <= cstar; c += 32)
                    if (cc[(size_t)c * T] > 0 && (c < cstar || need > 0)) atomicOr(&cbits[c >> 5], 1u << (c & 31));
            if (lane == 0) {
                GroupPlan* gp = plan_g + (size_t)img * T + g;
                gp->cstar = cstar;
                gp->need = need;
                gp->clast = clast;
                if (clast >= 0) atomicOr(&cbits[clast >> 5], 1u << (clast & 31));
            }
            if (lane == 0) take[g] = sstar;  // (take[] is dead from here on: reused for the fold below)
        }
    }
    __syncthreads();
    // Tie cells of every partly accepted group: the sums of its wholly accepted spans (zero when there is none).  All
    // threads share the (quantity, cell) items; a thread walks the spans of its item with eight loads in flight.
    if (!from_hist) {
        u64* ct = cells_tie_g + (size_t)img * PHD_CELL_Q * P.NC;
        for (int e = 0; e < sh_nscan; e++) {
            const int g = scan_list[e], sstar = take[g];
            int c0, nc;
            phd_group_cell_range(P, g, &c0, &nc);
            for (int i = tid; i < PHD_CELL_Q * nc; i += blockDim.x) {
                const int q = i / nc, cell = c0 + (i % nc);
                const size_t s0 = (size_t)img * P.nspans;
                u64 v = 0;
                if (q < 3) {
                    const u32* b = span32 + (s0 * 3 + q) * P.NC + cell;
                    const size_t st = (size_t)3 * P.NC;
                    int sp = 0;
                    for (; sp + 8 <= sstar; sp += 8) {
                        u32 t[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) t[u] = b[(size_t)(sp + u) * st];
#pragma unroll
                        for (int u = 0; u < 8; u++) v += t[u];
                    }
                    for (; sp < sstar; sp++) v += b[(size_t)sp * st];
                } else {
                    const u64* b = span64 + (s0 * 2 + (q - 3)) * P.NC + cell;
                    const size_t st = (size_t)2 * P.NC;
                    int sp = 0;
                    for (; sp + 8 <= sstar; sp += 8) {
                        u64 t[8];
#pragma unroll
                        for (int u = 0; u < 8; u++) t[u] = b[(size_t)(sp + u) * st];
#pragma unroll
                        for (int u = 0; u < 8; u++) v += t[u];
                    }
                    for (; sp < sstar; sp++) v += b[(size_t)sp * st];
                }
                ct[(size_t)q * P.NC + cell] = v;
            }
        }
    }
    __syncthreads();
    for (int w = tid; w < nbw; w += blockDim.x) {
        u32 bits = cbits[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            work[atomicAdd(work_n, 1u)] = (u32)img * (u32)P.nchunks + (u32)(w * 32 + b);
        }
    }
}

}  // namespace

void phd_launch_palette_select(const DevParams& P, int nimg, const double* centres, const float* sv_f, Workspace& ws,
                               cudaStream_t st, int* launches, bool from_hist) {
    const size_t smem = (size_t)P.T * 11 * sizeof(int) + (size_t)((P.nchunks + 31) / 32) * sizeof(u32);
    PHD_ALLOW_SMEM((k_palette_select), 200 * 1024);
    k_palette_select<<<nimg, 256, smem, st>>>(P, centres, sv_f, ws.cells, ws.counts_chunk, ws.plan, ws.pal_n,
                                              ws.parent_ids, ws.tie_list, ws.tie_n, ws.tie_groups, ws.dropped, ws.sacc,
                                              ws.hist, ws.iacc, ws.cells_tie, ws.span32, ws.span64, ws.work, ws.work_n,
                                              from_hist ? 1 : 0);
    *launches += 1;
}
