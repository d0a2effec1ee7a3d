// Per-pixel palette arithmetic of the single-pass front end: colour -> (cell, fixed-point h/s/v contributions).
//
// WHY CELLS.  The reference bins every pixel into a palette group (arm_octree, src/color_quantization.c:108-161),
// picks parent groups, moves the other groups' pixels to their nearest parent and only then averages
// wrap(h + 180 - h_parent), s and v per parent (calculate_avg_hsv, :510-576).  The wrap makes the hue sum depend
// on the parent, which is unknown while the pixels stream by.  But the wrap seams sit at h_parent +- 180, and
// group centres are bin mid-points, so every seam is a hue-bin edge or a hue-bin mid-point.  Splitting each
// (saturation/value class, hue bin) into HALF bins therefore gives statistics from which the sums for ANY
// parent assignment follow exactly -- one pass over the pixels, no second pass for the averages:
//
//     cell = ((cls * hp) + j) * 4 + sub        cls: si*vp+vi | sp*vp (gray) | sp*vp+1 (black);  j: hue bin
//         sub 1 / 3 : lower / upper half of hue bin j (interior pixels, plus edge pixels that behave like them)
//         sub 0     : pixels exactly on the bin's LOWER edge that belong to bin j but sit on the low side of a seam
//         sub 2     : pixels exactly on the bin's UPPER edge that the reference's rounding put into bin j
//                     (so that cell = cls*4hp + 2*halfbin + 1 for every ordinary pixel)
//     per cell: count, n(max==255), sum max, sum s*2^QS, sum (h - half-bin start)/(Lh/2)*2^QS
//
// CHUNK INDEX.  The pixel loop of the front end does not address its shared-memory chunk words by `cell` but by
//     ci = cls * 4hp + hb                       ordinary pixels: hb = 2j + (upper half), i.e. sub 1 / 3
//        = cls * 4hp + 2hp + 2j + (sub >> 1)    the rare edge cells, sub 0 / 2
// so that the cells that take 95 % of the pixels are CONSECUTIVE words: with cell = 4 pair + sub they all sat in odd
// words, i.e. the reductions of a warp shared 16 of the 32 banks (ncu: 23 M of 49 M shared-memory wavefronts of the
// kernel were bank conflicts of the reductions, and the LSU data pipe was its top unit at 82 %).  phd_pixel<true>
// returns ci, phd_pixel<false> the cell; phd_cell_from_ci converts.
//
// EXACTNESS.  hue = 60*(off + p/q) is a rational; the half-bin index is floor(num2/den) with num2 = 120*(off*q+p)
// (+720q if negative), den = Lh*q, evaluated in FP32 on exactly representable integers (< 2^24) with a guard eps
// that is far below the smallest possible distance 1/den of a non-integer quotient from an integer, and the
// remainder rem = num2 - hb*den is exact.  rem != 0: the reference's double rounding (~1e-14) cannot move the pixel
// across a half-bin boundary.  rem == 0 ("exceptional", ~2-15 % of 8-bit pixels): the pixel is exactly on a
// boundary and the reference's IEEE-double result decides (a) its hue bin, (b) on which side of a wrap seam it
// falls.  Both are pure functions of the 24-bit colour and h_partitions; k_build_exc evaluates the reference's
// double arithmetic (hsv_exact.cuh) once for all 2^24 colours into a code table that the pixel loop reads for
// exceptional pixels only.  An exceptional colour is identified by (max, min, half bin) -- the hue k * Lh/2 fixes sector
// and p, hence the third channel -- so the table is exc[tri(max, min)][2 hp] of TWO bytes: (cell delta, signed; 1 if the
// hue fraction counts as the END of its half bin): 2.4 MB at 18 hue bins instead of a byte per 24-bit colour (16 MB),
// (a signed 16-bit chunk-index delta: 2.4 MB),
// indexed by two numbers the loop has anyway (the class-table index and the half bin), and applied with two
// multiply-adds -- the pixel loop is bound by the ALU pipe (PRMT / LOP3 / IADD3 / SEL), and the former byte-permuted
// 24-bit index + bit-field decode put 7 such instructions on every pixel for the ~5 % that are exceptional.
// tests/test_gpu_parity.py sweeps all 2^24 colours against the CPU oracle.
#pragma once

#include "hsv_exact.cuh"

#define PHD_TRI_SIZE 32896  // 256*257/2: (max, min) pairs with min <= max

// ---- parameter table copied to shared memory by every CTA ----------------------------------------------------
//   svtab[tri(mx)+mn]  u8   class of the pixel from (max, min): si*vp+vi, sp*vp (gray), sp*vp+1 (black)
// (reciprocals of max and of Lh*(max-min) come from MUFU.RCP: 1 ulp is ample, see phd_pixel)
__host__ __device__ inline size_t phd_cell_tables_bytes() { return PHD_TRI_SIZE; }

__device__ __forceinline__ void phd_cell_tabs_to_smem(unsigned char* dst, const unsigned char* __restrict__ src) {
    const int n16 = (int)(phd_cell_tables_bytes() / 16);
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d[i] = __ldg(s + i);
}

// Constants of the pixel loop, hoisted into registers.
struct CellCfg {
    int hp4;        // hp * 4: cells per class
    float Lh;       // hue-bin width in degrees (an integer)
    float eps;      // guard of the half-bin floor
    float qscale;   // 2^QS
    u32 sat1_bits;  // bits of MAGIC + round(0.999999 * 2^QS)
    u32 full_val;   // 2^QS - 1: hue fraction of an edge pixel that counts as the END of its half bin
    u32 hb_n;       // 2 * hp: half bins = entries per (max, min) pair of the exceptional-colour table
};

#define PHD_MAGIC_RN 12582912.0f   // 1.5 * 2^23: x + MAGIC has round(x) in its low mantissa bits (|x| < 2^22)
#define PHD_MAGIC_RN_BITS 0x4B400000u
#define PHD_MAGIC_FLOOR 8388608.0f // 2^23, added with round-down: floor(x) in the low mantissa bits (0 <= x < 2^23)
#define PHD_MAGIC_FLOOR_BITS 0x4B000000u

// A non-integer num2/den is at least 1/den >= hp/91800 away from an integer; the FP32 evaluation errs by less than
// 2hp * 3 * 2^-24 (reciprocal 1 ulp, product, sum); eps sits a quarter of the way.
__host__ __device__ inline float phd_cell_eps(int hp) { return 0.25f * (float)hp / 91800.0f; }

__device__ __forceinline__ CellCfg phd_cell_cfg(const DevParams& P, int qs) {
    CellCfg c;
    c.hp4 = P.hp * 4;
    c.Lh = (float)P.Lh;
    c.eps = phd_cell_eps(P.hp);
    c.qscale = (float)(1u << qs);
    c.sat1_bits = PHD_MAGIC_RN_BITS + (u32)__double2uint_rn(0.999999 * (double)(1u << qs));
    c.full_val = (1u << qs) - 1u;
    c.hb_n = 2u * (u32)P.hp;
    return c;
}

struct PixOut {
    int cell;
    u32 w0;      // 1 + ((max == 255) << 16)
    u32 mx;
    u32 sbits;   // MAGIC_BITS + s * 2^QS      (the run accumulators subtract count * MAGIC_BITS at flush time)
    u32 hbits;   // MAGIC_BITS + fraction of the half bin * 2^QS
};

__device__ __forceinline__ float phd_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Exceptional-colour table: entry (t, k) -- t = tri(max) + min, k = half bin -- is the s16 chunk-index delta at
// 2 * (t * 2hp + k): 0, 2hp, 2hp - 1, or -1.  -1 (the pixel belongs to the half bin below) is also the one case in which
// its hue fraction counts as the END of that half bin, so the table needs no second field -- and the pixel loop ONE load:
// with two byte loads per exceptional pixel the loads alone (a few scattered lanes each, L2 latency) were ~12 % of the
// kernel's LSU wavefronts.
#define PHD_EXC_ENTRY 2
__host__ __device__ inline size_t phd_exc_bytes(int hp) { return (size_t)PHD_TRI_SIZE * 2 * (size_t)hp * PHD_EXC_ENTRY; }
// The pixel loop indexes the table with the half bin still carrying its float bias (the bits of 2^23 + k): kernels
// pass phd_pixel the table pointer moved back by that bias once, instead of subtracting it per pixel.
__device__ __forceinline__ const unsigned char* phd_exc_biased(const unsigned char* exc) {
    return exc - (unsigned long long)PHD_EXC_ENTRY * (unsigned long long)PHD_MAGIC_FLOOR_BITS;
}

// chunk index -> cell (see CHUNK INDEX above); cls is the class of ci (an exceptional pixel never leaves its class block)
__device__ __forceinline__ int phd_cell_from_ci(int ci, int cls, int hp) {
    const int r = ci - cls * 4 * hp;
    const int rare = r >= 2 * hp ? 1 : 0, rr = r - rare * 2 * hp;  // rr = 2j + (sub >> 1)
    return (cls * hp + (rr >> 1)) * 4 + 2 * (rr & 1) + 1 - rare;
}

// exc: the exceptional-colour table, biased (phd_exc_biased).  CI: o.cell is the chunk index ci instead of the cell.
template <bool CI = false>
__device__ __forceinline__ PixOut phd_pixel(int R, int G, int B, const unsigned char* __restrict__ svtab,
                                            const CellCfg& K, const unsigned char* __restrict__ exc) {
    PixOut o;
    const int mx = max(R, max(G, B)), mn = min(R, min(G, B));
    const int q = mx - mn;
    // sector with the reference's tie priority r, g, b (src/image_processing.c:394-397)
    const bool isR = (R == mx), isG = (G == mx);
    const int p = isR ? (G - B) : (isG ? (B - R) : (R - G));
    const float offk = isR ? 0.0f : (isG ? 240.0f : 480.0f);
    const float pf = (float)p, qf = (float)q;
    float num2 = fmaf(120.0f, pf, offk * qf);          // 120 * (off*q + p), exact
    if (num2 < 0.0f) num2 = fmaf(720.0f, qf, num2);    // h < 0 -> h + 360 (:398-404)
    const int tri = ((mx * mx + mx) >> 1) + mn;
    const int cls = svtab[tri];
    // half-bin index and exact remainder
    // Lh*q exactly for q >= 1 (the tiny addend rounds away); q == 0: num2 == 0 and any positive den gives half bin 0
    const float den = fmaf(K.Lh, qf, 1e-30f);
    const float rden = phd_rcp(den);
    const float y = fmaf(num2, rden, K.eps);
    const float hbm = __fadd_rd(y, PHD_MAGIC_FLOOR);   // bits: 0x4B000000 + floor(y)
    const float hbf = hbm - PHD_MAGIC_FLOOR;
    const float rem = fmaf(-hbf, den, num2);           // exact: integers below 2^24
    const float frac = rem * rden;                     // in [0, 1)
    u32 hbits = __float_as_uint(fmaf(frac, K.qscale, PHD_MAGIC_RN));
    // chunk index ci = cls*4hp + halfbin (the half bin still carries its float bias: removed below, or by the caller's
    // base address when CI)
    int cell = cls * K.hp4 + (int)__float_as_uint(hbm);
    if (rem == 0.0f && q != 0) {
        // exactly on a half-bin boundary: the reference's double rounding decides (k_build_exc).  frac == 0 here, so
        // hbits == MAGIC_RN_BITS and the END of the half bin is one multiply-add away.
        const unsigned char* e = exc + (unsigned long long)PHD_EXC_ENTRY * (u32)((u32)tri * K.hb_n + __float_as_uint(hbm));
        const int delta = (int)__ldg(reinterpret_cast<const short*>(e));
        cell += delta;
        if (delta < 0) hbits += K.full_val;
    }
    // saturation (src/image_processing.c:412-414): 0 | 0.999999 | delta/max
    // (delta == max gives 2^QS, the clamp turns it into 0.999999; max == 0 gives 0 through the fmax)
    const u32 sraw = __float_as_uint(fmaf(qf * phd_rcp(fmaxf((float)mx, 1.0f)), K.qscale, PHD_MAGIC_RN));
    const u32 sbits = min(sraw, K.sat1_bits);
    // CI: ci + PHD_MAGIC_FLOOR_BITS (the caller's base address absorbs the bias); otherwise the cell
    o.cell = CI ? cell : phd_cell_from_ci(cell - (int)PHD_MAGIC_FLOOR_BITS, cls, (int)(K.hb_n >> 1));
    o.w0 = (mx == 255) ? 0x10001u : 1u;
    o.mx = (u32)mx;
    o.sbits = sbits;
    o.hbits = hbits;
    return o;
}

// ---- cell -> reference group id -------------------------------------------------------------------------------
__device__ __forceinline__ int phd_cell_group(int cell, const DevParams& P) {
    const int pair = cell >> 2;
    const int cls = pair / P.hp, j = pair - cls * P.hp;
    const int spvp = P.sp * P.vp;
    if (cls < spvp) return j * spvp + cls;
    return cls == spvp ? P.T - (P.vp + 1) : P.T - 1;
}
