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
template <int THREADS, bool DS, int NCS, bool DB>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 3 : 1) k_pixels(const uint8_t* __restrict__ rgb, DevParams P,
                                                    const unsigned char* __restrict__ tabs_g,
                                                    const unsigned char* __restrict__ exc, int cpp,
                                                    u16* __restrict__ counts_chunk, u64* __restrict__ cells_g,
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
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    __shared__ uint2 dmeta[W3_THREADS];

    const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    phd_cell_tabs_to_smem(tb_raw, tabs_g + (W3 ? phd_cell_tables_bytes() : 0));  // W3: the table with the twin classes
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
        dmeta[tid] = make_uint2(d_pair | (d_twin << 16), d_cc | (flags << 16));
    }
    {   // keep the table pointer in registers: the compiler would reload it from the constant bank per pixel
        unsigned long long e;
        asm volatile("mov.u64 %0, %1;" : "=l"(e) : "l"(exc));
        exc = reinterpret_cast<const unsigned char*>(e);
    }
    const u32 cw_base0 = (u32)__cvta_generic_to_shared(chunkW);
    const u32 stride_b = 4u * (u32)ncs;
    const u32 set_b = (u32)NW * stride_b;  // bytes between the two chunk sets

    u32 sum[3] = {0, 0, 0}, sq[3] = {0, 0, 0};
    const int c_begin = blockIdx.x * cpp, c_end = min(c_begin + cpp, P.nchunks);
    const bool fast_ok = !DS && P.aligned16 != 0;
    u32 w[12];
    {
        const long long p0 = (long long)c_begin * CHUNK + (long long)tid * 16;
        if (fast_ok && p0 + 16 <= P.hpx) load48_aligned(base + p0 * 3, w);
    }
    for (int chunk = c_begin; chunk < c_end; chunk++) {
        const int set = DB ? ((chunk - c_begin) & 1) : 0;
        const u32 cw_base = cw_base0 + (u32)set * set_b;
        const u32 scratch = cw_base + 4u * (u32)(NCW + lane);
        const long long p0 = (long long)chunk * CHUNK + (long long)tid * 16;
        if (fast_ok && p0 + 16 <= P.hpx) {
            u32 wn[12];
            const long long pn = p0 + CHUNK;
            const bool more = (chunk + 1 < c_end) && (pn + 16 <= P.hpx);
            if (more) load48_aligned(base + pn * 3, wn);  // next chunk's bytes are in flight while this one is processed
            channel_sums(w, sum, sq);
            CellRun run;
            {
                const PixOut o = phd_pixel(packed_byte(w, 0), packed_byte(w, 1), packed_byte(w, 2), svtab, K, exc);
                run_start<4 * NCS, NW>(run, cw_base, o);
            }
#pragma unroll kPixUnroll
            for (int i = 1; i < 16; i++)
                run_step<4 * NCS, NW>(run, cw_base, scratch, stride_b,
                                  phd_pixel(packed_byte(w, 3 * i), packed_byte(w, 3 * i + 1), packed_byte(w, 3 * i + 2),
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
                const PixOut o = phd_pixel(R, G, B, svtab, K, exc);
                if (!any) {
                    run_start<4 * NCS, NW>(run, cw_base, o);
                    any = true;
                } else {
                    run_step<4 * NCS, NW>(run, cw_base, scratch, stride_b, o);
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
        auto drain3 = [&](int src, int dst, bool twin) -> u32 {
            uint4* ap = reinterpret_cast<uint4*>(cw) + src;
            const uint4 a = *ap;
            if ((a.x | a.y | a.z | a.w) == 0) return 0;
            saw = 1;
            uint4* sp4 = reinterpret_cast<uint4*>(cw + ncs) + src;
            uint4* hp4 = reinterpret_cast<uint4*>(cw + 2 * ncs) + src;
            const uint4 sv = *sp4, hv = *hp4;
            const uint4 z = make_uint4(0, 0, 0, 0);
            *ap = z; *sp4 = z; *hp4 = z;
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
            uint4* w0p = reinterpret_cast<uint4*>(cw) + pair;
            const uint4 c = *w0p;
            if ((c.x | c.y | c.z | c.w) == 0) return 0;
            uint4* w1p = reinterpret_cast<uint4*>(cw + ncs) + pair;
            uint4* w2p = reinterpret_cast<uint4*>(cw + 2 * ncs) + pair;
            uint4* w3p = reinterpret_cast<uint4*>(cw + 3 * ncs) + pair;
            const uint4 m = *w1p, sv = *w2p, hv = *w3p;
            const uint4 z = make_uint4(0, 0, 0, 0);
            *w0p = z; *w1p = z; *w2p = z; *w3p = z;
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
            const uint2 dm = dmeta[tid];
            const int d_pair = (int)(dm.x & 0xffffu), d_twin = (int)(dm.x >> 16), d_cc = (int)(dm.y & 0xffffu);
            if (d_pair != 0xffff) {
                u32 n = drain3(d_pair, d_pair, false);
                if (d_twin != 0xffff) n += drain3(d_twin, d_pair, true);
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
    for (int i = tid; i < NC; i += THREADS) {
        const u32 n = acc_cnt[i];
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
                        if (t != NONE && chunk < get_plan(t).cstar) accept(ct, o);  // whole chunk accepted
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
        const PixOut o = phd_pixel(R, G, B, tb_raw, K, exc);
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

// Exceptional-colour codes (pixel_cells.cuh): for every colour whose hue is exactly k * Lh/2, what the reference's
// doubles make of it.  code = full << 7 | (cell delta + 4), relative to the ordinary cell cls*4hp + 2k + 1.
__global__ void __launch_bounds__(256) k_build_exc(DevParams P, unsigned char* __restrict__ out, int* __restrict__ ok) {
    __shared__ double k255[256];
    phd_fill_k255(k255);
    __syncthreads();
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;  // R | G << 8 | B << 16
    const int R = c & 255, G = (c >> 8) & 255, B = (c >> 16) & 255;
    const int mx = max(R, max(G, B)), mn = min(R, min(G, B)), q = mx - mn;
    int code = 0;
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
            int delta, full = 0;
            if ((k & 1) == 0) {
                const int dj = hi_ref - (k >> 1);
                if (dj == 0) delta = below ? -1 : 0;  // (j, on the lower edge, low side) | (j, lower half)
                else if (dj == -1 && k > 0) {         // (j-1, end of the upper half) | (j-1, on the upper edge)
                    delta = below ? -2 : -3;
                    full = below ? 1 : 0;
                } else {
                    delta = 0;
                    atomicAnd(ok, 0);
                }
            } else {
                if (hi_ref != (k >> 1)) atomicAnd(ok, 0);
                delta = below ? -2 : 0;  // (j, end of the lower half) | (j, upper half)
                full = below ? 1 : 0;
            }
            code = (full << 7) | (delta + 4);
        }
    }
    out[c] = (unsigned char)code;
}

}  // namespace

// ------------------------------------------------------------------------------------------
#define PHD_NCS_SMALL 832  // compile-time chunk-array stride of the 256-thread variant (cells + twins + 32 scratch <= 832)

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
    // chunks per CTA: long walks amortise the table load and the final flush; enough CTAs to fill 148 SMs
    long long total = (long long)P.nchunks * nimg;
    int cpp = (int)(total / (148 * 12));
    cpp = cpp < 1 ? 1 : (cpp > 32 ? 32 : cpp);
    dim3 grid((P.nchunks + cpp - 1) / cpp, nimg);
    const bool ds = P.ds > 1;
    if (P.fe_threads == 256) {
        if (ds) k_pixels<256, true, PHD_NCS_SMALL, true><<<grid, 256, smem, st>>>(rgb, P, tabs, exc, cpp, ws.counts_chunk, ws.cells, ws.iacc);
        else k_pixels<256, false, PHD_NCS_SMALL, true><<<grid, 256, smem, st>>>(rgb, P, tabs, exc, cpp, ws.counts_chunk, ws.cells, ws.iacc);
    } else {
        if (ds) k_pixels<512, true, 0, false><<<grid, 512, smem, st>>>(rgb, P, tabs, exc, cpp, ws.counts_chunk, ws.cells, ws.iacc);
        else k_pixels<512, false, 0, false><<<grid, 512, smem, st>>>(rgb, P, tabs, exc, cpp, ws.counts_chunk, ws.cells, ws.iacc);
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
    if (exc_dev) k_build_exc<<<(1 << 24) / 256, 256, 0, st>>>(P, exc_dev, ok_dev);
}

size_t phd_cell_tables_size() { return 2 * phd_cell_tables_bytes(); }  // ordinary classes, then the table with twin classes
