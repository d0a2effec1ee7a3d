// The front end's walk over consecutive chunks of one image (see frontend.cu for what the front end replaces).
// Included by frontend.cu (k_pixels) and fused.cu (k_front_rows).
#pragma once

#include "pixel_cells.cuh"

namespace {

constexpr int kPixUnroll = 15;  // the 15 pixels after the first: fully unrolled (partial unrolling measured slower)

__device__ __forceinline__ u64 warp_sum_u64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_sum_u32(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Byte b (0..47) of 48 packed bytes, zero extended: one byte permute.
__device__ __forceinline__ int packed_byte(const u32 (&w)[12], int b) {
    return (int)__byte_perm(w[b >> 2], 0u, 0x4440u + (u32)(b & 3));
}

// Channel sums of 16 pixels: bytes are regrouped into channel-pure words (two permutes each) and reduced with
// dp4a (sum: dot with 1,1,1,1; sum of squares: dot with itself).
__device__ __forceinline__ void channel_sums(const u32 (&w)[12], u32 (&sum)[3], u32 (&sq)[3]) {
#pragma unroll
    for (int g = 0; g < 4; g++) {  // pixels 4g .. 4g+3 live in words 3g .. 3g+2
        const u32 a = w[3 * g], b = w[3 * g + 1], c = w[3 * g + 2];
        // bytes of (a,b,c): R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
        const u32 r = __byte_perm(__byte_perm(a, b, 0x0630), c, 0x5210);   // R0 R1 R2 R3
        const u32 gg = __byte_perm(__byte_perm(a, b, 0x0741), c, 0x6210);  // G0 G1 G2 G3
        const u32 bb = __byte_perm(__byte_perm(a, b, 0x0052), c, 0x7410);  // B0 B1 B2 B3
        sum[0] = __dp4a(r, 0x01010101u, sum[0]);
        sum[1] = __dp4a(gg, 0x01010101u, sum[1]);
        sum[2] = __dp4a(bb, 0x01010101u, sum[2]);
        sq[0] = __dp4a(r, r, sq[0]);
        sq[1] = __dp4a(gg, gg, sq[1]);
        sq[2] = __dp4a(bb, bb, sq[2]);
    }
}

// Per-thread run of consecutive pixels that fall into the same cell.  EVERY pixel issues four native 32-bit
// shared-memory reductions (ATOMS.ADD, no branch): when the cell changed they carry the finished run to its cell,
// otherwise they add to a per-lane scratch cell (index NC + lane: conflict free, never read).  A branch-free loop
// lets the scheduler overlap the 16 pixels of a thread -- with a branch per pixel the table / exceptional-colour
// load latencies were exposed (profiles/).  sbits/hbits carry the float->int magic bias; the drain removes
// count * bias (mod 2^32).
struct CellRun {
    u32 addr;  // shared-space byte address of the cell's word 0
    u32 w0, mx, s, h;
};

// STRIDE_B > 0: compile-time byte stride between the word arrays (immediate offsets); 0: runtime stride.
// NW = 4: words w0, mx, s, h.  NW = 3: word 0 is (count << 20 | sum max) and travels in r.mx; r.w0 is unused.
template <int STRIDE_B, int NW>
__device__ __forceinline__ void run_emit(u32 addr, const CellRun& r, u32 stride_b) {
    if (NW == 3) {
        static_assert(NW == 4 || STRIDE_B > 0, "the three-word layout has a compile-time stride");
        asm volatile(
            "red.shared.add.u32 [%0], %1;\n\t"
            "red.shared.add.u32 [%0+%4], %2;\n\t"
            "red.shared.add.u32 [%0+%5], %3;"
            ::"r"(addr), "r"(r.mx), "r"(r.s), "r"(r.h), "n"(STRIDE_B), "n"(2 * STRIDE_B));
    } else if (STRIDE_B > 0) {
        asm volatile(
            "red.shared.add.u32 [%0], %1;\n\t"
            "red.shared.add.u32 [%0+%5], %2;\n\t"
            "red.shared.add.u32 [%0+%6], %3;\n\t"
            "red.shared.add.u32 [%0+%7], %4;"
            ::"r"(addr), "r"(r.w0), "r"(r.mx), "r"(r.s), "r"(r.h), "n"(STRIDE_B), "n"(2 * STRIDE_B), "n"(3 * STRIDE_B));
    } else {
        asm volatile(
            "red.shared.add.u32 [%0], %1;\n\t"
            "red.shared.add.u32 [%5], %2;\n\t"
            "red.shared.add.u32 [%6], %3;\n\t"
            "red.shared.add.u32 [%7], %4;"
            ::"r"(addr), "r"(r.w0), "r"(r.mx), "r"(r.s), "r"(r.h), "r"(addr + stride_b), "r"(addr + 2 * stride_b),
              "r"(addr + 3 * stride_b));
    }
}

// word 0 of the three-word layout: one pixel and its max
__device__ __forceinline__ u32 phd_w3_word0(const PixOut& o) { return o.mx | (1u << 20); }

template <int STRIDE_B, int NW>
__device__ __forceinline__ void run_start(CellRun& r, u32 base, const PixOut& o) {
    r.addr = base + 4u * (u32)o.cell;
    r.w0 = o.w0;
    r.mx = NW == 3 ? phd_w3_word0(o) : o.mx;
    r.s = o.sbits;
    r.h = o.hbits;
}

template <int STRIDE_B, int NW>
__device__ __forceinline__ void run_step(CellRun& r, u32 base, u32 scratch, u32 stride_b, const PixOut& o) {
    const u32 addr = base + 4u * (u32)o.cell;
    const bool change = (addr != r.addr);
    run_emit<STRIDE_B, NW>(change ? r.addr : scratch, r, stride_b);
    // accumulator = accumulator * keep + new: one multiply-add each instead of a select and an add (measured: neutral
    // while the kernel was bound by four atomics per pixel, -2 % with three)
    const u32 keep = change ? 0u : 1u;
    const u32 m0 = NW == 3 ? phd_w3_word0(o) : o.mx;
    if (NW == 4) asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r.w0) : "r"(keep), "r"(o.w0));
    asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r.mx) : "r"(keep), "r"(m0));
    asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r.s) : "r"(keep), "r"(o.sbits));
    asm("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r.h) : "r"(keep), "r"(o.hbits));
    r.addr = addr;
}

__device__ __forceinline__ void load48_aligned(const uint8_t* __restrict__ p, u32 (&w)[12]) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    w[8] = c.x; w[9] = c.y; w[10] = c.z; w[11] = c.w;
}

// ------------------------------------------------------------------------------------------
// One CTA walks `cpp` consecutive chunks of one image.  A chunk is THREADS*16 HSV pixels; each thread owns 16
// consecutive pixels (48 bytes, three 16-byte loads, prefetched one chunk ahead).  Shared memory: the class table,
// TWO sets of chunk cell words (4 arrays each, fed by reductions; while chunk c+1 fills one set the other is
// drained, so there is one barrier per chunk), and the CTA's running cell sums (plain adds in the drain, one owner
// thread per cell), flushed to the image's global cells once per CTA.
//   chunk word 0: count | n255 << 16     1: sum max     2: sum s * 2^QS     3: sum hue fraction * 2^QS
// QS = 32 - log2(chunk pixels), so a whole chunk cannot overflow 32 bits.
// NCS: compile-time stride (words) between the chunk word arrays, 0 = P.NC + 32 at run time.
// DB: two chunk sets (one barrier per chunk); false = one set and a second barrier after the drain (large palettes).
// fe_walk is the body of that walk as a device function so that two kernels can run it: k_pixels (one walk per CTA,
// frontend.cu) and the front-end role of k_front_rows (fused.cu: persistent CTAs that take front-end walks and row-FFT
// tasks from one queue).  `span` is the walk's index inside the image: besides adding its cell sums to the image's
// cells, the walk stores them as the image's span `span` (span32 / span64), from which the tie path takes every span
// that lies wholly before a tie group's cut-off chunk without looking at a pixel again (palette_select.cu).
// load_table: the class table is not in shared memory yet (first walk of a CTA, or another role used the memory).
// The caller guarantees a CTA-wide barrier between the previous user of the shared memory and this call.
template <int THREADS, bool DS, int NCS, bool DB, bool PREFETCH = true>
__device__ __forceinline__ void fe_walk(unsigned char* smem_raw, const uint8_t* __restrict__ rgb, const DevParams& P,
                                        const unsigned char* __restrict__ tabs_g,
                                        const unsigned char* __restrict__ exc, const int img, const int span,
                                        const int c_begin, const int c_end, const bool load_table,
                                        u16* __restrict__ counts_chunk, u64* __restrict__ cells_g,
                                        u32* __restrict__ span32, u64* __restrict__ span64,
                                        ImageAcc* __restrict__ iacc) {
    constexpr int CHUNK = THREADS * 16;
    constexpr int QS = (THREADS == 256) ? 20 : 19;
    static_assert(THREADS == 256 || THREADS == 512, "chunk size / QS pairs");
    // Three chunk words per cell instead of four (the kernel is bound by the shared-memory atomics): count and sum of
    // max share word 0 (count << 20 | sum max: 4096 pixels * 255 < 2^20), and the pixels with max == 255 -- whose count
    // the v clamp needs -- go to TWIN cells behind the ordinary ones (the twin class ids come straight out of a
    // second class table, so the pixel loop pays nothing) and are folded into their base cells by the drain.
    constexpr bool W3 = (THREADS == 256);
    constexpr int W3_THREADS = W3 ? THREADS : 1;
    constexpr int NW = W3 ? 3 : 4;
    static_assert(!W3 || (NCS > 0 && DB), "the three-word layout is the double-buffered fixed-stride variant");
    const int NC = P.NC;
    const int NCW = W3 ? NC + (P.sp + 1) * P.hp * 4 : NC;  // cells of the chunk arrays: ordinary + twins
    const int ncs = NCS > 0 ? NCS : NC + 32;  // word stride of the chunk arrays (cells + 32 scratch cells)
    unsigned char* tb_raw = smem_raw;
    u32* chunkW = reinterpret_cast<u32*>(smem_raw + phd_cell_tables_bytes());      // [DB ? 2 : 1][NW][ncs]
    u64* acc_s = reinterpret_cast<u64*>(chunkW + (DB ? 2 : 1) * NW * ncs);         // [NC]
    u64* acc_h = acc_s + NC;                                                       // [NC]
    u32* acc_cnt = reinterpret_cast<u32*>(acc_h + NC);                             // [NC]
    u32* acc_n255 = acc_cnt + NC;
    u32* acc_mx = acc_n255 + NC;
    __shared__ u64 red[6][THREADS / 32];
    __shared__ u32 gb[2][2];  // per chunk set: pixels of the gray and of the black group
    // W3: the (class, hue bin) pair each thread drains after every chunk: x = pair | twin pair << 16 (0xffff: none),
    // y = slot of the pair's group in the per-chunk group counts (0xffff: gray / black, summed through gb[]) |
    // (0 gray, 1 black) << 16 | (repairs wrapped all-black chunks) << 17.  Pairs that have a twin (top value bin of every
    // saturation bin, gray) sit together in the first warps and the others start at the next warp boundary, so that no
    // warp runs the twin drain for a few lanes only.
    __shared__ uint4 dmeta[W3_THREADS];  // z, w: chunk-index word offset (cls*4hp + 2j) of the pair and of its twin

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    if (load_table) phd_cell_tabs_to_smem(tb_raw, tabs_g + (W3 ? phd_cell_tables_bytes() : 0));  // W3: the table with the twin classes
    for (int i = tid; i < (DB ? 2 : 1) * NW * ncs; i += THREADS) chunkW[i] = 0;
    for (int i = tid; i < NC; i += THREADS) {
        acc_s[i] = 0; acc_h[i] = 0; acc_cnt[i] = 0; acc_n255[i] = 0; acc_mx[i] = 0;
    }
    if (tid < 4) gb[tid >> 1][tid & 1] = 0;
    __syncthreads();
    const unsigned char* svtab = tb_raw;
    const CellCfg K = phd_cell_cfg(P, QS);
    const int spvp = P.sp * P.vp, hp = P.hp, npairs_colour = spvp * hp;
    const int twin_pair0 = NC / 4;                  // first twin pair of the chunk arrays
    const int black_cell = (spvp + 1) * hp * 4 + 1; // cell of the colour (0,0,0)
    int saw = 0;                                    // this thread's last drain met a non-empty cell
    if (W3) {
        const int n_top = P.sp * hp, n_tw = n_top + hp, t0 = (n_tw + 31) & ~31;
        const int per_s = (P.vp - 1) * hp, n_rest = P.sp * per_s, r = tid - t0;
        u32 d_pair = 0xffff, d_twin = 0xffff, d_cc = 0xffff, flags = 0;
        if (tid < n_top) {
            const int si = tid / hp, j = tid - si * hp, cls = si * P.vp + P.vp - 1;
            d_pair = cls * hp + j; d_twin = twin_pair0 + si * hp + j; d_cc = j * spvp + cls;
        } else if (tid < n_tw) {
            const int j = tid - n_top;
            d_pair = spvp * hp + j; d_twin = twin_pair0 + P.sp * hp + j;
        } else if (r >= 0 && r < n_rest) {
            const int si = r / per_s, rem = r - si * per_s, vi = rem / hp, j = rem - vi * hp, cls = si * P.vp + vi;
            d_pair = cls * hp + j; d_cc = j * spvp + cls;
        } else if (r >= n_rest && r < n_rest + hp) {
            const int j = r - n_rest;
            d_pair = (spvp + 1) * hp + j; flags = 1u | (j == 0 ? 2u : 0u);
        }
        auto src_of = [&](u32 pair) -> u32 { const u32 c = pair / (u32)hp; return c * 4u * (u32)hp + 2u * (pair - c * (u32)hp); };
        dmeta[tid] = make_uint4(d_pair | (d_twin << 16), d_cc | (flags << 16), d_pair != 0xffff ? src_of(d_pair) : 0u,
                                d_twin != 0xffff ? src_of(d_twin) : 0u);
    }
    {   // keep the table pointer in registers: the compiler would reload it from the constant bank per pixel
        unsigned long long e;
        asm volatile("mov.u64 %0, %1;" : "=l"(e) : "l"(phd_exc_biased(exc)));
        exc = reinterpret_cast<const unsigned char*>(e);
    }
    const u32 cw_base0 = (u32)__cvta_generic_to_shared(chunkW);
    const u32 stride_b = 4u * (u32)ncs;
    const u32 set_b = (u32)NW * stride_b;  // bytes between the two chunk sets

    u32 sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    const bool fast_ok = !DS && P.aligned16 != 0;
    u32 w[12];
    {
        const long long p0 = (long long)c_begin * CHUNK + (long long)tid * 16;
        if (PREFETCH && fast_ok && p0 + 16 <= P.hpx) load48_aligned(base + p0 * 3, w);
    }
    for (int chunk = c_begin; chunk < c_end; chunk++) {
        const int set = DB ? ((chunk - c_begin) & 1) : 0;
        const u32 cw_base = cw_base0 + (u32)set * set_b;
        const u32 scratch = cw_base + 4u * (u32)(NCW + lane);
        // phd_pixel<true> leaves the float bias of the half bin in its chunk index: the base absorbs it (mod 2^32)
        const u32 cw_ci = cw_base - 4u * PHD_MAGIC_FLOOR_BITS;
        const long long p0 = (long long)chunk * CHUNK + (long long)tid * 16;
        if (fast_ok && p0 + 16 <= P.hpx) {
            // PREFETCH: the next chunk's bytes are in flight (12 registers) while this one is processed; without it the
            // chunk's own bytes are loaded here (the fused kernel, where other roles' CTAs cover the latency and the
            // registers are needed)
            u32 wn[12];
            const long long pn = p0 + CHUNK;
            const bool more = PREFETCH && (chunk + 1 < c_end) && (pn + 16 <= P.hpx);
            if (!PREFETCH) load48_aligned(base + p0 * 3, w);
            if (more) load48_aligned(base + pn * 3, wn);
            channel_sums(w, sum, sq);
            CellRun run;
            {
                const PixOut o = phd_pixel<true>(packed_byte(w, 0), packed_byte(w, 1), packed_byte(w, 2), svtab, K, exc);
                run_start<4 * NCS, NW>(run, cw_ci, o);
            }
#pragma unroll kPixUnroll
            for (int i = 1; i < 16; i++)
                run_step<4 * NCS, NW>(run, cw_ci, scratch, stride_b,
                                  phd_pixel<true>(packed_byte(w, 3 * i), packed_byte(w, 3 * i + 1), packed_byte(w, 3 * i + 2),
                                            svtab, K, exc));
            run_emit<4 * NCS, NW>(run.addr, run, stride_b);
            if (more) {
#pragma unroll
                for (int i = 0; i < 12; i++) w[i] = wn[i];
            }
        } else if (p0 < P.hpx) {
            // image tail, unaligned input or downsampled HSV grid: one pixel at a time
            CellRun run{0xffffffffu, 0, 0, 0, 0};
            bool any = false;
#pragma unroll 1
            for (int i = 0; i < 16; i++) {
                if (p0 + i >= P.hpx) break;
                const uint8_t* q = base + phd_src_index(p0 + i, P) * 3;
                const int R = __ldg(q), G = __ldg(q + 1), B = __ldg(q + 2);
                if (!DS) {
                    sum[0] += R; sum[1] += G; sum[2] += B;
                    sq[0] += R * R; sq[1] += G * G; sq[2] += B * B;
                }
                const PixOut o = phd_pixel<true>(R, G, B, svtab, K, exc);
                if (!any) {
                    run_start<4 * NCS, NW>(run, cw_ci, o);
                    any = true;
                } else {
                    run_step<4 * NCS, NW>(run, cw_ci, scratch, stride_b, o);
                }
            }
            if (any) run_emit<4 * NCS, NW>(run.addr, run, stride_b);
        }
        // the only barrier of the chunk: set `set` is complete, the other set is free again.  W3: it also tells whether
        // the previous chunk's drain met any non-empty cell -- a chunk of 4096 pure black pixels wraps word 0 of its
        // one cell to zero (count 4096 << 20, sum max 0, s = h = 0) and looks empty; every other full cell is caught by
        // its non-zero sum max (see drain3)
        const int any_prev = W3 ? __syncthreads_or(saw) : (__syncthreads(), 1);
        saw = 0;
        // drain: chunk words -> running sums, per-chunk group counts (needed for raster ranks in the tie path)
        u32* cw = chunkW + (size_t)set * NW * ncs;
        u16* cc = counts_chunk + ((size_t)img * P.nchunks + chunk) * P.T;
        if (DB && tid == 0 && chunk > c_begin) {  // gray / black totals of the PREVIOUS chunk are complete now
            u16* ccp = cc - P.T;
            const u32 wrapped = any_prev ? 0u : (u32)CHUNK;
            ccp[P.T - (P.vp + 1)] = (u16)gb[set ^ 1][0];
            ccp[P.T - 1] = (u16)(gb[set ^ 1][1] + wrapped);
            gb[set ^ 1][0] = 0; gb[set ^ 1][1] = 0;
        }
        // W3: (source pair of the chunk arrays, pair of the running sums it is added to, twin = its pixels have max 255)
        // chunk words of a (class, hue bin) pair in the chunk-index layout (pixel_cells.cuh): the two ordinary half-bin
        // cells (sub 1, 3) are the word pair at `src` (= cls*4hp + 2j), the two edge cells (sub 0, 2) the pair 2hp further
        const int rare_off = 2 * hp;
        auto ld_pair = [&](u32* arr, int src) -> uint4 {
            const uint2 o = *reinterpret_cast<const uint2*>(arr + src), r = *reinterpret_cast<const uint2*>(arr + src + rare_off);
            return make_uint4(r.x, o.x, r.y, o.y);
        };
        auto zero_pair = [&](u32* arr, int src) {
            *reinterpret_cast<uint2*>(arr + src) = make_uint2(0, 0);
            *reinterpret_cast<uint2*>(arr + src + rare_off) = make_uint2(0, 0);
        };
        auto drain3 = [&](int src, int dst, bool twin) -> u32 {
            const uint4 a = ld_pair(cw, src);
            if ((a.x | a.y | a.z | a.w) == 0) return 0;
            saw = 1;
            const uint4 sv = ld_pair(cw + ncs, src), hv = ld_pair(cw + 2 * ncs, src);
            zero_pair(cw, src); zero_pair(cw + ncs, src); zero_pair(cw + 2 * ncs, src);
            // count field; a cell that took the whole chunk wrapped it to 0 but kept its sum of max
            auto cnt = [](u32 w0) -> u32 { const u32 c = w0 >> 20; return (c == 0 && w0 != 0) ? (u32)CHUNK : c; };
            const uint4 n = make_uint4(cnt(a.x), cnt(a.y), cnt(a.z), cnt(a.w));
            uint4* ac = reinterpret_cast<uint4*>(acc_cnt) + dst;
            uint4* am = reinterpret_cast<uint4*>(acc_mx) + dst;
            uint4 t = *ac; t.x += n.x; t.y += n.y; t.z += n.z; t.w += n.w; *ac = t;
            if (twin) {
                uint4* an = reinterpret_cast<uint4*>(acc_n255) + dst;
                t = *an; t.x += n.x; t.y += n.y; t.z += n.z; t.w += n.w; *an = t;
            }
            t = *am; t.x += a.x & 0xfffffu; t.y += a.y & 0xfffffu; t.z += a.z & 0xfffffu; t.w += a.w & 0xfffffu; *am = t;
            ulonglong2* as = reinterpret_cast<ulonglong2*>(acc_s) + 2 * dst;
            ulonglong2* ah = reinterpret_cast<ulonglong2*>(acc_h) + 2 * dst;
            ulonglong2 u = as[0]; u.x += sv.x - n.x * PHD_MAGIC_RN_BITS; u.y += sv.y - n.y * PHD_MAGIC_RN_BITS; as[0] = u;
            u = as[1]; u.x += sv.z - n.z * PHD_MAGIC_RN_BITS; u.y += sv.w - n.w * PHD_MAGIC_RN_BITS; as[1] = u;
            u = ah[0]; u.x += hv.x - n.x * PHD_MAGIC_RN_BITS; u.y += hv.y - n.y * PHD_MAGIC_RN_BITS; ah[0] = u;
            u = ah[1]; u.x += hv.z - n.z * PHD_MAGIC_RN_BITS; u.y += hv.w - n.w * PHD_MAGIC_RN_BITS; ah[1] = u;
            return n.x + n.y + n.z + n.w;
        };
        // the four sub-cells of a (class, hue bin) pair are adjacent in every array: 16-byte accesses
        auto drain_pair = [&](int pair) -> u32 {
            const int pc = pair / hp, src = pc * 4 * hp + 2 * (pair - pc * hp);
            const uint4 c = ld_pair(cw, src);
            if ((c.x | c.y | c.z | c.w) == 0) return 0;
            const uint4 m = ld_pair(cw + ncs, src), sv = ld_pair(cw + 2 * ncs, src), hv = ld_pair(cw + 3 * ncs, src);
            zero_pair(cw, src); zero_pair(cw + ncs, src); zero_pair(cw + 2 * ncs, src); zero_pair(cw + 3 * ncs, src);
            const uint4 n = make_uint4(c.x & 0xffffu, c.y & 0xffffu, c.z & 0xffffu, c.w & 0xffffu);
            uint4* ac = reinterpret_cast<uint4*>(acc_cnt) + pair;
            uint4* an = reinterpret_cast<uint4*>(acc_n255) + pair;
            uint4* am = reinterpret_cast<uint4*>(acc_mx) + pair;
            uint4 t = *ac; t.x += n.x; t.y += n.y; t.z += n.z; t.w += n.w; *ac = t;
            t = *an; t.x += c.x >> 16; t.y += c.y >> 16; t.z += c.z >> 16; t.w += c.w >> 16; *an = t;
            t = *am; t.x += m.x; t.y += m.y; t.z += m.z; t.w += m.w; *am = t;
            // sums carry count * bias (mod 2^32): the true chunk sums fit 32 bits
            ulonglong2* as = reinterpret_cast<ulonglong2*>(acc_s) + 2 * pair;
            ulonglong2* ah = reinterpret_cast<ulonglong2*>(acc_h) + 2 * pair;
            ulonglong2 u = as[0]; u.x += sv.x - n.x * PHD_MAGIC_RN_BITS; u.y += sv.y - n.y * PHD_MAGIC_RN_BITS; as[0] = u;
            u = as[1]; u.x += sv.z - n.z * PHD_MAGIC_RN_BITS; u.y += sv.w - n.w * PHD_MAGIC_RN_BITS; as[1] = u;
            u = ah[0]; u.x += hv.x - n.x * PHD_MAGIC_RN_BITS; u.y += hv.y - n.y * PHD_MAGIC_RN_BITS; ah[0] = u;
            u = ah[1]; u.x += hv.z - n.z * PHD_MAGIC_RN_BITS; u.y += hv.w - n.w * PHD_MAGIC_RN_BITS; ah[1] = u;
            return n.x + n.y + n.z + n.w;
        };
        if (W3) {
            // this thread's pair (see dmeta): read here so that nothing about it stays live through the pixel loop
            const uint4 dm = dmeta[tid];
            const int d_pair = (int)(dm.x & 0xffffu), d_twin = (int)(dm.x >> 16), d_cc = (int)(dm.y & 0xffffu);
            if (d_pair != 0xffff) {
                u32 n = drain3((int)dm.z, d_pair, false);
                if (d_twin != 0xffff) n += drain3((int)dm.w, d_pair, true);
                if (d_cc != 0xffff) cc[d_cc] = (u16)n;
                else if (n) atomicAdd(&gb[set][(dm.y >> 16) & 1u], n);
                // the owner of the black pair of hue bin 0 repairs a wrapped all-black PREVIOUS chunk (see the barrier)
                if ((dm.y >> 17) && chunk > c_begin && !any_prev) acc_cnt[black_cell] += (u32)CHUNK;
            }
        } else {
            for (int pair = tid; pair < npairs_colour; pair += THREADS) {
                const int cls = pair / hp, j = pair - cls * hp;
                cc[j * spvp + cls] = (u16)drain_pair(pair);
            }
            // gray and black: all hue bins collapse into one group each; handled by the LAST threads so that the
            // colour pairs and these spread over different warps
            for (int k = THREADS - 1 - tid; k < 2 * hp; k += THREADS) {
                const int which = k / hp, j = k - which * hp;
                const u32 cnt = drain_pair((spvp + which) * hp + j);
                if (cnt) atomicAdd(&gb[set][which], cnt);
            }
        }
        if (tid >= 64 && tid < 64 + P.vp - 1) cc[P.T - P.vp + (tid - 64)] = 0;  // gray groups 2.. are never used
        if (!DB) {
            __syncthreads();
            if (tid == 0) {
                cc[P.T - (P.vp + 1)] = (u16)gb[0][0];
                cc[P.T - 1] = (u16)gb[0][1];
                gb[0][0] = 0; gb[0][1] = 0;
            }
        }
    }
    const int any_last = W3 ? __syncthreads_or(saw) : (__syncthreads(), 1);
    if (DB && tid == 0 && c_end > c_begin) {
        const int set = (c_end - 1 - c_begin) & 1;
        u16* cc = counts_chunk + ((size_t)img * P.nchunks + (c_end - 1)) * P.T;
        const u32 wrapped = any_last ? 0u : (u32)CHUNK;
        cc[P.T - (P.vp + 1)] = (u16)gb[set][0];
        cc[P.T - 1] = (u16)(gb[set][1] + wrapped);
    }
    if (W3 && c_end > c_begin && !any_last) {  // the last chunk was a wrapped all-black one
        if (tid == 0) acc_cnt[black_cell] += (u32)CHUNK;
        __syncthreads();
    }

    // flush the CTA's cell sums (Q20 in global memory whatever QS is) and the channel sums
    u64* cg = cells_g + (size_t)img * PHD_CELL_Q * NC;
    u32* s32 = span32 + ((size_t)img * P.nspans + span) * 3 * NC;
    u64* s64 = span64 + ((size_t)img * P.nspans + span) * 2 * NC;
    for (int i = tid; i < NC; i += THREADS) {
        const u32 n = acc_cnt[i];
        // the walk's own sums (plain coalesced stores, zeros included: the arrays are never cleared)
        s32[i] = n; s32[NC + i] = acc_n255[i]; s32[2 * NC + i] = acc_mx[i];
        s64[i] = acc_s[i] << (20 - QS); s64[NC + i] = acc_h[i] << (20 - QS);
        if (n) {
            atomicAdd(cg + i, (u64)n);
            if (acc_n255[i]) atomicAdd(cg + NC + i, (u64)acc_n255[i]);
            atomicAdd(cg + 2 * NC + i, (u64)acc_mx[i]);
            atomicAdd(cg + 3 * NC + i, acc_s[i] << (20 - QS));
            atomicAdd(cg + 4 * NC + i, acc_h[i] << (20 - QS));
        }
    }
    if (!DS) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const u64 a = warp_sum_u64(sum[k]), b = warp_sum_u64(sq[k]);
            if (lane == 0) { red[k][wid] = a; red[3 + k][wid] = b; }
        }
        __syncthreads();
        if (tid < 6) {
            u64 v = 0;
            for (int w2 = 0; w2 < THREADS / 32; w2++) v += red[tid][w2];
            ImageAcc* a = iacc + img;
            if (v) atomicAdd(tid < 3 ? &a->sum[tid] : &a->sumsq[tid - 3], v);
        }
    }
}

}  // namespace
