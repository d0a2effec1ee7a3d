// Hand-written 2-D real FFT of the grayscale image with the blur-profile binning fused into the
// column pass.  No cuFFT anywhere in the library.
//
// Data flow (per image):  packed RGB --k_rows--> specT[x][k] (row-transformed half spectrum, TRANSPOSED so a
// spectrum column is contiguous) --k_cols--> integer blur bins + max power.  The final spectrum is never
// written: power, ln, polar binning happen in the column kernel's epilogue.
//
//   k_rows_t / k_rows_generic   first half of pgm_fft (src/fft_processing.c:18-63) with rgb2pgm fused in front
//       (src/image_processing.c:505-512).  Rows are read as packed RGB, turned into the exact integer gray
//       numerators 299R+587G+114B, two rows are packed as one complex sequence, transformed in shared memory
//       (Stockham autosort) and split into the two half spectra.  A constant 127500 (= 0.5 gray) is removed
//       before the transform for FP32 headroom; DC is repaired exactly in the column epilogue.
//   k_cols_t / k_cols_generic   second half of pgm_fft, remove_dc_bias (src/blur_profile.c:233-238),
//       pgm_normalize_fft (src/fft_processing.c:173-200) and the accumulation loop of calculate_blur_profile
//       (src/blur_profile.c:87-100).  Columns (contiguous in specT) and the matching slice of the bin-id map
//       are staged into shared memory with cp.async.bulk (the TMA copy engine) signalled through an mbarrier,
//       transformed, then power -> (p<1 ? 0 : ln p) -> polar bin -> run-length merged per thread ->
//       shared-memory integer bins -> 64-bit global integer bins; max power by atomicMax.
//       G_s is a scalar and is applied after averaging, in finalize.
//   k_bin_map   cartesian_to_polar_conversion + the bin index arithmetic (src/blur_profile.c:427-458, :94-97;
//       newton_int_sqrt src/utilities.c:43-52); image independent, built once per (W,H,nr,na) and cached.
//
// The *_t kernels are compile-time specialised (length and radix plan as template arguments: radix-15/16/25...
// register butterflies with folded constants) for the lengths of PHD_FFT_PLANS (BASELINE.json's shapes, video and
// camera sizes); every other length runs the *_generic kernels (runtime radix list with register butterflies for
// 2..13, 15..19, 21, 25, an O(p^2) pass for any other prime p up to 40 and Bluestein's chirp-z for lengths with a larger
// prime factor: arbitrary crops are served rather than refused).
#include <cuda.h>  // CUtensorMap (types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "fft_tables.cuh"
#include "frontend_walk.cuh"
#include "phd_internal.h"

namespace {

#define PHD_GRAY_BIAS 127500
#define PHD_MAX_PRIME 12799  // = any side the shared-memory kernels serve; large primes are slow (O(p^2)), not refused
constexpr int kRowThreads = 256;
constexpr int kColThreads = 512;
constexpr bool kColsPacked = true;  // column butterflies on the FP32x2 pipe (see Cx)
constexpr int kColsMinBlocks1080 = 1; // register hint of the 1080-point column kernel (see k_cols_t)

// Experiment knobs (-D...): pass twiddles generated from ONE loaded base twiddle per butterfly (w, w^2 = w*w, w^3 = w^2*w,
// w^4 = (w^2)^2, ...: a product tree of depth <= 5) instead of R-1 table loads.
#ifndef PHD_ROWS_TWPOW
#define PHD_ROWS_TWPOW 1
#endif
#ifndef PHD_COLS_TWPOW
#define PHD_COLS_TWPOW 1
#endif
#ifndef PHD_ROWS_TMA
#define PHD_ROWS_TMA 1  // row kernel: the transposed spectrum leaves through a TMA tensor store (see rows_walk)
#endif
#ifndef PHD_RT_TWPOW
#define PHD_RT_TWPOW 0   // runtime-radix passes: twiddles by the product tree (measured faster, not accurate enough in Bluestein)
#endif
#ifndef PHD_ROWS_PK
#define PHD_ROWS_PK 0   // row butterflies on the packed FP32x2 pipe
#endif

// Complex arithmetic, scalar (PK = false) or on the packed FP32x2 pipe of sm_100 (PK = true: PTX
// add/sub/mul/fma.rn.f32x2 -> SASS FADD2 / FMUL2 / FFMA2, one instruction for both components, scalars broadcast
// as an operand modifier).  Packed halves the instruction count of the complex adds and scalings but not their
// time on the FP32 pipe: it pays in the column kernel, which is bound by instruction issue (-7 %), and costs a
// little in the row kernel, which is bound by latency (+1 % at 1920, +6 % at 6000) -- so rows stay scalar.
template <bool PK> struct Cx;
template <> struct Cx<false> {
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
    static __device__ __forceinline__ float2 fmas(float2 a, float c, float2 b) {  // a * c + b, c real
        return make_float2(fmaf(a.x, c, b.x), fmaf(a.y, c, b.y));
    }
    static __device__ __forceinline__ float2 scale(float2 a, float c) { return make_float2(a.x * c, a.y * c); }
};
#define PHD_PK2(op, r, a, b)                                                                                     \
    asm("{ .reg .b64 pa, pb, pc; mov.b64 pa, {%2,%3}; mov.b64 pb, {%4,%5}; " op ".rn.f32x2 pc, pa, pb;"           \
        " mov.b64 {%0,%1}, pc; }" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y))
template <> struct Cx<true> {
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { float2 r; PHD_PK2("add", r, a, b); return r; }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) { float2 r; PHD_PK2("sub", r, a, b); return r; }
    static __device__ __forceinline__ float2 fmas(float2 a, float c, float2 b) {
        float2 r;
        asm("{ .reg .b64 pa, pb, pc, pd; mov.b64 pa, {%2,%3}; mov.b64 pb, {%4,%4}; mov.b64 pc, {%5,%6};"
            " fma.rn.f32x2 pd, pa, pb, pc; mov.b64 {%0,%1}, pd; }"
            : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(c), "f"(b.x), "f"(b.y));
        return r;
    }
    static __device__ __forceinline__ float2 scale(float2 a, float c) {
        float2 r;
        asm("{ .reg .b64 pa, pb, pc; mov.b64 pa, {%2,%3}; mov.b64 pb, {%4,%4}; mul.rn.f32x2 pc, pa, pb;"
            " mov.b64 {%0,%1}, pc; }" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(c));
        return r;
    }
};
// a * (c + i s): the cross terms by two scalar multiplies, the rest one (packed) FMA
template <bool PK>
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
    return Cx<PK>::fmas(a, c, make_float2(-a.y * s, a.x * s));
}
template <bool PK>
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return cmul_cs<PK>(a, b.x, b.y); }
// i * (a - b) = (b.y - a.y, a.x - b.x): the rotated difference the odd outputs of a butterfly add and subtract
__device__ __forceinline__ float2 isub(float2 a, float2 b) { return make_float2(b.y - a.y, a.x - b.x); }

// ------------------------------------------------------------------------------------------
// Register butterflies.  Radix<R, PK>::run is an in-place R-point DFT, natural order in and out.
// ------------------------------------------------------------------------------------------
template <int R, bool PK> struct Radix;

template <bool PK> struct Radix<2, PK> {
    using C = Cx<PK>;
    static __device__ __forceinline__ void run(float2 (&x)[2]) {
        const float2 a = x[0], b = x[1];
        x[0] = C::add(a, b);
        x[1] = C::sub(a, b);
    }
};
template <bool PK> struct Radix<4, PK> {
    using C = Cx<PK>;
    static __device__ __forceinline__ void run(float2 (&x)[4]) {
        const float2 t0 = C::add(x[0], x[2]), t1 = C::sub(x[0], x[2]);
        const float2 t2 = C::add(x[1], x[3]), t3 = isub(x[3], x[1]);  // -i (x1 - x3)
        x[0] = C::add(t0, t2);
        x[1] = C::add(t1, t3);
        x[2] = C::sub(t0, t2);
        x[3] = C::sub(t1, t3);
    }
};
template <bool PK> struct Radix<3, PK> {
    using C = Cx<PK>;
    static __device__ __forceinline__ void run(float2 (&x)[3]) {
        const float k3 = 0.86602540378443864676f;
        const float2 t = C::add(x[1], x[2]);
        const float2 u = C::fmas(t, -0.5f, x[0]);
        const float2 v = C::scale(isub(x[2], x[1]), k3);  // -i k3 (x1 - x2)
        x[0] = C::add(x[0], t);
        x[1] = C::add(u, v);
        x[2] = C::sub(u, v);
    }
};
template <bool PK> struct Radix<5, PK> {
    using C = Cx<PK>;
    static __device__ __forceinline__ void run(float2 (&x)[5]) {
        const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
        const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
        const float2 t1 = C::add(x[1], x[4]), t2 = C::add(x[2], x[3]);
        const float2 r3 = isub(x[4], x[1]), r4 = isub(x[3], x[2]);  // -i (x1 - x4), -i (x2 - x3)
        const float2 m1 = C::fmas(t2, c2, C::fmas(t1, c1, x[0]));
        const float2 m2 = C::fmas(t2, c1, C::fmas(t1, c2, x[0]));
        const float2 n1 = C::fmas(r4, s2, C::scale(r3, s1));
        const float2 n2 = C::fmas(r4, -s1, C::scale(r3, s2));
        x[0] = C::add(C::add(x[0], t1), t2);
        x[1] = C::add(m1, n1);
        x[4] = C::sub(m1, n1);
        x[2] = C::add(m2, n2);
        x[3] = C::sub(m2, n2);
    }
};

// RA*RB-point DFT from RA- and RB-point ones (Cooley-Tukey in registers, constants folded at compile time):
// n = n1*RB + n2, k = k1 + RA*k2;  X[k] = sum_n2 W_N^(n2 k1) [sum_n1 x[n] W_RA^(n1 k1)] W_RB^(n2 k2).
template <int RA, int RB, bool PK>
struct Composite {
    static __device__ __forceinline__ void run(float2 (&x)[RA * RB]) {
        constexpr int N = RA * RB;
        float2 t[N];
#pragma unroll
        for (int n2 = 0; n2 < RB; n2++) {
            float2 y[RA];
#pragma unroll
            for (int n1 = 0; n1 < RA; n1++) y[n1] = x[n1 * RB + n2];
            Radix<RA, PK>::run(y);
#pragma unroll
            for (int k1 = 0; k1 < RA; k1++) {
                const int e = (n2 * k1) % N;
                if (e == 0) t[n2 * RA + k1] = y[k1];
                else t[n2 * RA + k1] = cmul_cs<PK>(y[k1], PhdTw<N>::c(e), PhdTw<N>::s(e));
            }
        }
#pragma unroll
        for (int k1 = 0; k1 < RA; k1++) {
            float2 z[RB];
#pragma unroll
            for (int n2 = 0; n2 < RB; n2++) z[n2] = t[n2 * RA + k1];
            Radix<RB, PK>::run(z);
#pragma unroll
            for (int k2 = 0; k2 < RB; k2++) x[k1 + RA * k2] = z[k2];
        }
    }
};
// Odd prime P by the definition, folded on the conjugate symmetry W^(P-k) = conj(W^k): with s_k = x[k] + x[P-k],
// d_k = x[k] - x[P-k],  X[j] = A_j + i B_j and X[P-j] = A_j - i B_j,  A_j = x0 + sum_k cos(jk) s_k,
// B_j = sum_k TWS(jk) d_k  ((P-1)^2 FMAs in all, constants folded at compile time).
template <int P, bool PK>
struct PrimeRadix {
    using C = Cx<PK>;
    static __device__ __forceinline__ void run(float2 (&x)[P]) {
        constexpr int h = (P - 1) / 2;
        float2 s[h], d[h];
#pragma unroll
        for (int k = 1; k <= h; k++) {
            s[k - 1] = C::add(x[k], x[P - k]);
            d[k - 1] = isub(x[k], x[P - k]);  // i d_k, so that i B_j accumulates directly
        }
        const float2 x0 = x[0];
        float2 sum = x0;
#pragma unroll
        for (int k = 0; k < h; k++) sum = C::add(sum, s[k]);
        x[0] = sum;
#pragma unroll
        for (int j = 1; j <= h; j++) {
            float2 A = x0, B = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 1; k <= h; k++) {
                const int e = (j * k) % P;
                A = C::fmas(s[k - 1], PhdTw<P>::c(e), A);
                B = C::fmas(d[k - 1], PhdTw<P>::s(e), B);
            }
            x[j] = C::add(A, B);
            x[P - j] = C::sub(A, B);
        }
    }
};
template <bool PK> struct Radix<7, PK> : PrimeRadix<7, PK> {};
template <bool PK> struct Radix<11, PK> : PrimeRadix<11, PK> {};
template <bool PK> struct Radix<13, PK> : PrimeRadix<13, PK> {};
template <bool PK> struct Radix<17, PK> : PrimeRadix<17, PK> {};
template <bool PK> struct Radix<19, PK> : PrimeRadix<19, PK> {};
template <bool PK> struct Radix<6, PK> : Composite<2, 3, PK> {};
template <bool PK> struct Radix<8, PK> : Composite<2, 4, PK> {};
template <bool PK> struct Radix<9, PK> : Composite<3, 3, PK> {};
template <bool PK> struct Radix<10, PK> : Composite<2, 5, PK> {};
template <bool PK> struct Radix<12, PK> : Composite<3, 4, PK> {};
template <bool PK> struct Radix<15, PK> : Composite<3, 5, PK> {};
template <bool PK> struct Radix<16, PK> : Composite<4, 4, PK> {};
template <bool PK> struct Radix<18, PK> : Composite<2, 9, PK> {};
template <bool PK> struct Radix<21, PK> : Composite<3, 7, PK> {};
template <bool PK> struct Radix<25, PK> : Composite<5, 5, PK> {};

// Per-pass twiddle tables of the compile-time plans: pass i (radix R, stride S, M = N/R butterflies) reads
// twp[off_i + (j-1)*M + b] = exp(-2 pi i * (b - b % S) * j / N), j = 1..R-1 -- consecutive lanes read consecutive
// entries (the plain table indexed by pps*j scatters a warp over up to 2*j cache lines).  Only M/S of a pass's bases
// differ, but a compact table indexed by b/S measured SLOWER (rows +19 %, columns +30 % at 1080p): one entry per
// butterfly it stays.
template <int N, int R0, int R1, int R2, int R3>
struct PlanT {
    static constexpr int off0 = 0;
    static constexpr int off1 = off0 + (R0 - 1) * (N / R0);
    static constexpr int off2 = off1 + (R1 - 1) * (N / R1);
    static constexpr int off3 = off2 + (R2 - 1) * (N / R2);
};

// Padded layout of the first pass's input (one extra float2 every 16): the producers write 16 consecutive
// elements per lane, and 17-element lane strides keep their 8-byte stores conflict free.
__device__ __forceinline__ int pad16(int e) { return e + (e >> 4); }

// One Stockham pass, compile-time length N, radix R, stride S (product of the radices already applied), over
// nfft sequences laid out fstride apart in shared memory.  Reads are unit stride across lanes; the first pass
// (S == 1) writes with stride R, which is why the plans start with an odd radix (conflict-free 8-byte stores).
// GT > 0: the CTA is split into groups of GT threads and group f owns sequence f for the whole transform, so the
// passes of one sequence synchronise on a NAMED barrier of GT threads (bar.sync 1+f) instead of the whole CTA:
// warps only ever wait for the siblings that work on the same sequence.  GT == 0: all threads share all sequences.
template <int GT>
__device__ __forceinline__ void seq_sync() {
    if (GT == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)threadIdx.x / GT), "n"(GT) : "memory");
}

// w^2 for a unit complex w
__device__ __forceinline__ float2 csqr(float2 w) { return make_float2(fmaf(w.x, w.x, -w.y * w.y), 2.0f * w.x * w.y); }
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// OPAD / IPAD: the output of this pass is skewed by OPAD elements per block of R*S outputs, the input of this pass was
// written with such a skew of IPAD per sub-sequence (see fft_skew).
template <int N, int R, int S, bool PADIN, int GT, bool PK, bool TWPOW = false, int OPAD = 0, int IPAD = 0>
__device__ __forceinline__ void pass_t(const float2* __restrict__ in, float2* __restrict__ out,
                                       const float2* __restrict__ twp, int nfft, int fstride_in, int fstride_out) {
    constexpr int M = N / R;
    // padded input: with M % 16 == 0 the pad of b + k * M splits into pad16(b) + k * (M + M / 16); otherwise per element
    const int total = GT > 0 ? M : nfft * M;
    const int first = GT > 0 ? (int)threadIdx.x % GT : (int)threadIdx.x;
    const int step = GT > 0 ? GT : (int)blockDim.x;
    const int f_fixed = GT > 0 ? (int)threadIdx.x / GT : 0;
    if (GT > 0 && f_fixed >= nfft) return;
    for (int idx = first; idx < total; idx += step) {
        const int f = GT > 0 ? f_fixed : idx / M;
        const int b = idx - (GT > 0 ? 0 : f * M);
        const int q = b % S;
        const int pps = b - q;
        const float2* a = in + f * fstride_in;
        float2* y = out + f * fstride_out;
        float2 x[R];
        if (PADIN) {
            if constexpr (M % 16 == 0) {
                const int b0 = pad16(b);
#pragma unroll
                for (int k = 0; k < R; k++) x[k] = a[b0 + k * (M + M / 16)];
            } else {
#pragma unroll
                for (int k = 0; k < R; k++) x[k] = a[pad16(b + k * M)];
            }
        } else {
#pragma unroll
            for (int k = 0; k < R; k++) x[k] = a[b + k * (M + IPAD)];
        }
        Radix<R, PK>::run(x);
        if (OPAD > 0) y += OPAD * (b / S);
        if (S * R == N) {  // last pass: every twiddle is 1
#pragma unroll
            for (int j = 0; j < R; j++) y[q + j * S] = x[j];
        } else if (TWPOW) {
            // one table entry per butterfly (the j = 1 row), its powers by squaring / multiplying
            float2 w[R];
            w[1] = __ldg(&twp[b]);
            y[R * pps + q] = x[0];
#pragma unroll
            for (int j = 1; j < R; j++) {
                if (j > 1) w[j] = (j % 2 == 0) ? csqr(w[j / 2]) : cmul2(w[j - 1], w[1]);
                y[R * pps + q + j * S] = cmulf<PK>(x[j], w[j]);
            }
        } else {
            y[R * pps + q] = x[0];
#pragma unroll
            for (int j = 1; j < R; j++) y[R * pps + q + j * S] = cmulf<PK>(x[j], __ldg(&twp[(j - 1) * M + b]));
        }
    }
}

// Up to four passes; R3 == 1 means a three-pass plan.  a: input (padded layout when PADIN), b: scratch; sequence f
// lives at a + f * fstride_a and b + f * fstride in EVERY pass -- with thread groups (GT > 0) the groups drift apart by
// whole passes, so a sequence must never be written into another sequence's region of either buffer.
// Returns the buffer holding the result (unpadded; stride fstride for b, fstride_a for a).  The barrier after the
// LAST pass is left to the caller (it usually needs a CTA-wide one there anyway).
// Skew of the second pass's output.  Consecutive butterflies b of that pass write consecutive elements q = b % R0 of one
// block of R0*R1 outputs and then jump to the next block: in a half warp (8-byte accesses) the lane after the jump lands
// on the bank pair of an earlier lane unless the block stride is congruent to R0 modulo 16 -- two-way conflicts on every
// store of the pass (ncu: 11 % of the row kernel's shared-memory wavefronts).  The third pass reads sub-sequence k at
// k * (N / R2) = k * R0 * R1, i.e. exactly those blocks, so skewing every block by `fft_skew` elements costs nothing on
// the reading side.  Only where the buffer has the room: the padded staging layout (N / 16 spare elements).
__host__ __device__ constexpr int fft_skew(int n, int r0, int r1, int r2, bool padin) {
    const int p = (((r0 * (1 - r1)) % 16) + 16) % 16;
    return (padin && p * (r2 - 1) <= n / 16) ? p : 0;
}

template <int N, int R0, int R1, int R2, int R3, bool PADIN, int GT, bool PK, bool TWPOW = false>
__device__ __forceinline__ float2* fft_run_t(float2* a, float2* b, const float2* __restrict__ twp, int nfft,
                                             int fstride_a, int fstride) {
    static_assert(R0 * R1 * R2 * R3 == N, "radix plan does not multiply to N");
    using PL = PlanT<N, R0, R1, R2, R3>;
    pass_t<N, R0, 1, PADIN, GT, PK, TWPOW>(a, b, twp + PL::off0, nfft, fstride_a, fstride);
    seq_sync<GT>();
    constexpr int SKEW = R3 == 1 ? fft_skew(N, R0, R1, R2, PADIN) : 0;
    pass_t<N, R1, R0, false, GT, PK, TWPOW, SKEW>(b, a, twp + PL::off1, nfft, fstride, fstride_a);
    seq_sync<GT>();
    pass_t<N, R2, R0 * R1, false, GT, PK, false, 0, SKEW>(a, b, twp + PL::off2, nfft, fstride_a, fstride);
    if (R3 == 1) return b;
    seq_sync<GT>();
    pass_t<N, (R3 == 1 ? 2 : R3), (R3 == 1 ? N / 2 : R0 * R1 * R2), false, GT, PK>(b, a, twp + PL::off3, nfft, fstride, fstride_a);
    return a;
}

// ---- runtime-radix fallback (any length whose prime factors are <= PHD_MAX_PRIME) ------------
template <int R>
__device__ __forceinline__ void pass_rt(const float2* __restrict__ in, float2* __restrict__ out, int n, int s,
                                        const float2* __restrict__ twp, int nbatch, int bstride) {
    const int m = n / R;
    const int total = nbatch * m;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int col = idx / m;
        const int b = idx - col * m;
        const float2* a = in + col * bstride;
        float2* y = out + col * bstride;
        const int bs = b / s, q = b - bs * s;
        const int pps = b - q;
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = a[b + k * m];
        Radix<R, false>::run(x);
        y[R * pps + q] = x[0];
        // Table twiddles here.  The product tree of pass_t was 6..12 % faster on these kernels too, but in the Bluestein
        // route (two transforms and three chirp products per sequence) its few extra ulps per twiddle pushed one low-valued
        // blur bin of a 358x600 image to 1.5e-4 relative (tools/soak.py, case 86 of seed 21): accuracy first on this path.
        if (PHD_RT_TWPOW != 0) {
            float2 w[R];
            w[1] = __ldg(&twp[b]);
#pragma unroll
            for (int j = 1; j < R; j++) {
                if (j > 1) w[j] = (j % 2 == 0) ? csqr(w[j / 2]) : cmul2(w[j - 1], w[1]);
                y[R * pps + q + j * s] = cmulf<false>(x[j], w[j]);
            }
        } else {
#pragma unroll
            for (int j = 1; j < R; j++) y[R * pps + q + j * s] = cmulf<false>(x[j], __ldg(&twp[(j - 1) * m + b]));
        }
    }
}

// other primes (and any radix without a register butterfly): O(r^2) pass straight from the definition, one thread per
// OUTPUT (sequence, butterfly, j) so that even a prime transform length (a single butterfly per sequence) keeps the
// whole CTA busy; lanes of a warp share the inputs (shared-memory broadcast) and gather the twiddles from L1.
__device__ void pass_rt_prime(int r, const float2* __restrict__ in, float2* __restrict__ out, int n, int s,
                              const float2* __restrict__ tw, const float2* __restrict__ twp, int nbatch, int bstride) {
    const int m = n / r;
    const int total = nbatch * m * r;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int rest = idx / r, j = idx - rest * r;
        const int col = rest / m, b = rest - col * m;
        const float2* a = in + col * bstride;
        float2* y = out + col * bstride;
        const int bs = b / s, q = b - bs * s;
        const int pps = b - q;
        float2 acc = a[b];
        int e = 0;  // (j * k) % r, stepped
        for (int k = 1; k < r; k++) {
            e += j;
            if (e >= r) e -= r;
            acc = Cx<false>::add(acc, cmulf<false>(a[b + k * m], __ldg(&tw[e * m])));
        }
        y[r * pps + q + j * s] = j ? cmulf<false>(acc, __ldg(&twp[(j - 1) * m + b])) : acc;
    }
}

__device__ float2* fft_run_rt(const FftPlan& pl, float2* bufA, float2* bufB, int nbatch, int bstride) {
    float2* a = bufA;
    float2* b = bufB;
    int s = 1;
    const int len = pl.m > 0 ? pl.m : pl.n;  // a Bluestein plan carries the radix plan of its padded length
    for (int f = 0; f < pl.nfac; f++) {
        const int r = pl.fac[f];
        const float2* twp = pl.twp + pl.twp_off[f];
        switch (r) {
            case 2: pass_rt<2>(a, b, len, s, twp, nbatch, bstride); break;
            case 3: pass_rt<3>(a, b, len, s, twp, nbatch, bstride); break;
            case 4: pass_rt<4>(a, b, len, s, twp, nbatch, bstride); break;
            case 5: pass_rt<5>(a, b, len, s, twp, nbatch, bstride); break;
            case 6: pass_rt<6>(a, b, len, s, twp, nbatch, bstride); break;
            case 7: pass_rt<7>(a, b, len, s, twp, nbatch, bstride); break;
            case 8: pass_rt<8>(a, b, len, s, twp, nbatch, bstride); break;
            case 9: pass_rt<9>(a, b, len, s, twp, nbatch, bstride); break;
            case 10: pass_rt<10>(a, b, len, s, twp, nbatch, bstride); break;
            case 11: pass_rt<11>(a, b, len, s, twp, nbatch, bstride); break;
            case 12: pass_rt<12>(a, b, len, s, twp, nbatch, bstride); break;
            case 13: pass_rt<13>(a, b, len, s, twp, nbatch, bstride); break;
            case 15: pass_rt<15>(a, b, len, s, twp, nbatch, bstride); break;
            case 16: pass_rt<16>(a, b, len, s, twp, nbatch, bstride); break;
            case 17: pass_rt<17>(a, b, len, s, twp, nbatch, bstride); break;
            case 18: pass_rt<18>(a, b, len, s, twp, nbatch, bstride); break;
            case 19: pass_rt<19>(a, b, len, s, twp, nbatch, bstride); break;
            case 21: pass_rt<21>(a, b, len, s, twp, nbatch, bstride); break;
            case 25: pass_rt<25>(a, b, len, s, twp, nbatch, bstride); break;
            default: pass_rt_prime(r, a, b, len, s, pl.tw, twp, nbatch, bstride); break;
        }
        __syncthreads();
        s *= r;
        float2* t = a; a = b; b = t;
    }
    return a;
}

// Bluestein (chirp-z) for lengths with a large prime factor: X[k] = w[k] * sum_j (x[j] w[j]) conj(w)[k-j] with
// w[k] = exp(-i pi k^2 / n), the sum being a circular convolution of length m >= 2n-1 done with two transforms of the
// small-radix length m (the inverse one as conj(FFT(conj(.)))).  O(m log m) per sequence instead of the O(n p) of the
// direct pass over a prime factor p (src/fft_processing.c:34 leaves this choice to FFTW, which does the same).
// In: nbatch sequences of n values at stride bstride (>= m) in bufA.  Returns the buffer holding the n results.
__device__ float2* fft_run_blue(const FftPlan& pl, float2* bufA, float2* bufB, int nbatch, int bstride) {
    const int n = pl.n, m = pl.m;
    for (int idx = threadIdx.x; idx < nbatch * m; idx += blockDim.x) {
        const int col = idx / m, k = idx - col * m;
        float2* p = bufA + col * bstride + k;
        *p = k < n ? cmulf<false>(*p, __ldg(&pl.chirp[k])) : make_float2(0.f, 0.f);
    }
    __syncthreads();
    float2* r = fft_run_rt(pl, bufA, bufB, nbatch, bstride);
    float2* other = (r == bufA) ? bufB : bufA;
    for (int idx = threadIdx.x; idx < nbatch * m; idx += blockDim.x) {
        const int col = idx / m, k = idx - col * m;
        float2* p = r + col * bstride + k;
        const float2 v = cmulf<false>(*p, __ldg(&pl.bhat[k]));
        *p = make_float2(v.x, -v.y);
    }
    __syncthreads();
    float2* r2 = fft_run_rt(pl, r, other, nbatch, bstride);
    for (int idx = threadIdx.x; idx < nbatch * n; idx += blockDim.x) {
        const int col = idx / n, k = idx - col * n;
        float2* p = r2 + col * bstride + k;
        const float2 v = *p;
        *p = cmulf<false>(make_float2(v.x, -v.y), __ldg(&pl.chirp[k]));
    }
    __syncthreads();
    return r2;
}

// chirp[k] = exp(-i pi k^2 / n); bhat = (1/m) * DFT_m of the conjugate chirp laid out circularly (index -j at m - j).
// Evaluated straight from the definition in double precision, once per shape.
__global__ void k_bluestein_tables(float2* __restrict__ chirp, float2* __restrict__ bhat, int n, int m) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        double sn, cs;
        sincospi((double)(((long long)k * k) % (2LL * n)) / (double)n, &sn, &cs);
        chirp[k] = make_float2((float)cs, (float)(-sn));
    }
    if (k < m) {
        double re = 0, im = 0;
        for (int j = -(n - 1); j < n; j++) {
            // exp(+i pi j^2 / n) * exp(-2 pi i j k / m): one angle, both parts reduced exactly in integers
            const double a1 = (double)(((long long)j * j) % (2LL * n)) / (double)n;
            long long jk = ((long long)j * k) % m;
            if (jk < 0) jk += m;
            const double a2 = 2.0 * (double)jk / (double)m;
            double sn, cs;
            sincospi(a1 - a2, &sn, &cs);
            re += cs;
            im += sn;
        }
        bhat[k] = make_float2((float)(re / m), (float)(im / m));
    }
}

__global__ void k_twiddles(float2* tw, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)n, &s, &c);
    tw[k] = make_float2((float)c, (float)(-s));
}

// Pass tables of a compile-time plan (see PlanT): one launch per pass.
__global__ void k_pass_twiddles(float2* out, int n, int r, int s) {
    const int m = n / r;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (r - 1) * m) return;
    const int j = i / m + 1, b = i - (j - 1) * m;
    const int pps = b - b % s;
    double sn, cs;
    sincospi(2.0 * (double)(((long long)pps * j) % n) / (double)n, &sn, &cs);
    out[i] = make_float2((float)cs, (float)(-sn));
}

// Gray numerator 299 R + 587 G + 114 B - 127500 of pixel i (0..15) of 48 packed bytes: two 2-way dot products
// (16-bit weights x bytes, IDP.2A) straight on the packed words, whatever the pixel's byte alignment.
template <int NWORDS>
__device__ __forceinline__ int gray16(const u32 (&w)[NWORDS], int i) {
    const int b = 3 * i, k = b >> 2, s = b & 3;
    const u32 bias = (u32)(-PHD_GRAY_BIAS);
    constexpr u32 RG = 299u | (587u << 16), B_ = 114u, _R = 299u << 16, GB = 587u | (114u << 16);
    if (s == 0) return (int)__dp2a_hi(B_, w[k], __dp2a_lo(RG, w[k], bias));       // R G B .
    if (s == 1) return (int)__dp2a_hi(GB, w[k], __dp2a_lo(_R, w[k], bias));       // . R G B
    if (s == 2) return (int)__dp2a_lo(B_, w[k + 1], __dp2a_hi(RG, w[k], bias));   // . . R G | B
    return (int)__dp2a_lo(GB, w[k + 1], __dp2a_hi(_R, w[k], bias));               // . . . R | G B
}

// ------------------------------------------------------------------------------------------
// Rows, specialised: one CTA transforms PAIRS row pairs at a time (each pair = one packed complex sequence) and
// writes, for every spectrum column x, the 2*PAIRS consecutive entries specT[x][2*PAIRS*j ..] (32-byte sectors
// with PAIRS = 2.  PAIRS = 1 writes 16-byte half sectors: measured 1.9x SLOWER on 3840- and 6000-pixel rows
// despite three times the resident CTAs -- partial-sector writes are that expensive -- so every shape uses 2).
// Requires W % 16 == 0, H % (2*PAIRS) == 0, 16-byte aligned rows.
// ------------------------------------------------------------------------------------------
// CTAs per SM the register allocation of the row kernel must allow: what shared memory admits, but never so many
// that a thread is left with fewer than 64 registers (the radix-15..25 butterflies spill below that).
constexpr int rows_min_blocks(int smem_bytes, int threads) {
    const int by_smem = smem_bytes <= 72 * 1024 ? 3 : (smem_bytes <= 110 * 1024 ? 2 : 1);
    const int by_regs = 65536 / (threads * 64) < 1 ? 1 : 65536 / (threads * 64);
    return by_smem < by_regs ? by_smem : by_regs;
}

// The row kernel stages SEG pixels of a row pair per thread (rows_walk): 16 where width and first-pass sub-length are
// multiples of 16, else 8 for widths that are a multiple of 8 (1080, 3000, 600, 1800: portrait video and photos).
template <int N, int R0>
constexpr int rows_seg() { return (N % 16 == 0 && (N / R0) % 16 == 0) ? 16 : 8; }
template <int N, int R0>
constexpr bool rows_t_ok() { return N % 8 == 0; }

// shared memory of the row kernel with the TMA store: two buffers of two padded row pairs, each rounded up to 1 KB
constexpr int rows_tma_smem(int n) { return 2 * ((2 * (n + n / 16) * 8 + 1023) / 1024 * 1024); }

// rows_walk: the body as a device function -- steps q_begin, q_begin + q_step, ... < q_end of image `img` -- shared by
// k_rows_t (one walk per CTA) and the row role of k_front_rows (persistent CTAs, tasks from a queue).
// TMA tensor store of one box (inner 8 floats = the four rows of a step, up to 256 spectrum columns, one image) from a
// shared-memory tile laid out [column][32 bytes] with the 32-byte swizzle of the tensor map (SASS: UTMASTG).
__device__ __forceinline__ void tma_store_box(const CUtensorMap* tmap, const void* smem_tile, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
                 "r"((u32)__cvta_generic_to_shared(smem_tile)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// TMA_OUT: the step's output (for every spectrum column k the 32 bytes specT[k][4q .. 4q+3]) of the first `ktma` columns
// is assembled as a tile in shared memory (the staging buffer is free after the last pass) and leaves through the TMA
// engine; the other columns are stored directly.  Written directly, every lane of
// a store instruction goes to a different 128-byte line: 961 lane-sectors per step through the LSU, whose data pipe is
// the top unit of this kernel (ncu: 78 %); measured, the direct stores cost 2.3 of the kernel's 8.1 ms per 2048 1080p
// images even when they are redirected to an L2-resident region, i.e. the cost is the requests, not the DRAM writes.
// SEG: pixels per staging task, 16 (three 16-byte loads per row) or 8 (three 8-byte loads: widths that are a multiple of 8
// but not of 16, e.g. the 1080-pixel rows of portrait video -- rows then start on 8-byte boundaries only).
template <int N, int R0, int R1, int R2, int R3, int THREADS, int PAIRS, bool PREFETCH = true, bool TMA_OUT = false, int SEG = 16>
__device__ __forceinline__ void rows_walk(unsigned char* smem_raw, const uint8_t* __restrict__ rgb, const DevParams& P,
                                          const float2* __restrict__ twp, float2* __restrict__ specT, const int img,
                                          const int q_begin, const int q_end, const int q_step,
                                          const CUtensorMap* tmap = nullptr, const int ktma = 0,
                                          const CUtensorMap* tmap_tail = nullptr) {
    constexpr int NP = N + N / 16;  // padded length
    // TMA_OUT: both buffers hold the padded layout and start on a 1 KB boundary (either can be the staging buffer, and
    // either can hold the swizzled output tile): the two swap roles every step, see below
    constexpr int BUF = TMA_OUT ? (PAIRS * NP * 8 + 1023) / 1024 * 1024 / 8 : PAIRS * NP;  // float2 per buffer
    float2* bufA = reinterpret_cast<float2*>(smem_raw);  // [PAIRS][NP]  (also pass scratch: [PAIRS][N] fits)
    float2* bufB = bufA + BUF;                           // [PAIRS][N]
    const uint8_t* img_base = rgb + (size_t)img * P.image_stride;
    // Per step: PAIRS row pairs x N/16 segments of 16 pixels; one task (thread) holds the segment of BOTH rows of
    // its pair (2 x three 16-byte loads).  The loads of the next step are issued before the passes of this one and
    // stay in registers meanwhile.
    constexpr int SEGS = N / SEG;
    static_assert(SEG == 16 || SEG == 8, "16 or 8 pixels per staging task");
    static_assert(N % SEG == 0, "whole segments");
    static_assert(R3 == 1, "the output loop below reads the result with stride N (buffer b)");
    constexpr int GT = THREADS / PAIRS;  // threads of one pair: they stage, transform and synchronise among themselves
    static_assert(SEGS <= GT && GT % 32 == 0, "one staging task per thread of the pair's group");
    const int pair = threadIdx.x / GT, seg = threadIdx.x % GT;
    const bool has_task = seg < SEGS;
    // (SEG == 16 keeps its six 16-byte registers and builds the word arrays at the point of use: ptxas schedules that
    // form 2 % better than word arrays filled by the loads)
    uint4 a0, b0, c0, a1, b1, c1;
    uint2 d0[3], d1[3];
    auto load_step = [&](int q) {
        const uint8_t* base = img_base + (size_t)(2 * PAIRS * q) * N * 3;
        const uint8_t* p0 = base + (size_t)(2 * pair) * N * 3 + (size_t)seg * (3 * SEG);
        const uint8_t* p1 = base + (size_t)(2 * pair + 1) * N * 3 + (size_t)seg * (3 * SEG);
        // plain cached loads: the three pieces of neighbouring threads share 128-byte lines, and loads that bypass L1
        // (ld.global.nc.L1::no_allocate) measured 20 % slower for the whole kernel
        if constexpr (SEG == 16) {
            const uint4* s0 = reinterpret_cast<const uint4*>(p0);
            const uint4* s1 = reinterpret_cast<const uint4*>(p1);
            a0 = __ldg(s0); b0 = __ldg(s0 + 1); c0 = __ldg(s0 + 2);
            a1 = __ldg(s1); b1 = __ldg(s1 + 1); c1 = __ldg(s1 + 2);
        } else {
            const uint2* s0 = reinterpret_cast<const uint2*>(p0);
            const uint2* s1 = reinterpret_cast<const uint2*>(p1);
#pragma unroll
            for (int j = 0; j < 3; j++) { d0[j] = __ldg(s0 + j); d1[j] = __ldg(s1 + j); }
        }
    };
    if (PREFETCH && has_task && q_begin < q_end) load_step(q_begin);
    const int fw = N / 2 + 1;
    static_assert(!TMA_OUT || PAIRS == 2, "the tile holds 32 bytes per column");
    for (int q = q_begin; q < q_end; q += q_step) {
        if (has_task) {
            if (!PREFETCH) load_step(q);  // no registers held across the passes (fused kernel)
            if constexpr (SEG == 16) {
                const u32 w0[12] = {a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, b0.z, b0.w, c0.x, c0.y, c0.z, c0.w};
                const u32 w1[12] = {a1.x, a1.y, a1.z, a1.w, b1.x, b1.y, b1.z, b1.w, c1.x, c1.y, c1.z, c1.w};
                float2* dst = bufA + pair * NP + seg * 17;
#pragma unroll
                for (int i = 0; i < 16; i++) dst[i] = make_float2((float)gray16(w0, i), (float)gray16(w1, i));
            } else {
                const u32 w0[6] = {d0[0].x, d0[0].y, d0[1].x, d0[1].y, d0[2].x, d0[2].y};
                const u32 w1[6] = {d1[0].x, d1[0].y, d1[1].x, d1[1].y, d1[2].x, d1[2].y};
                float2* dst = bufA + pair * NP + pad16(seg * 8);  // an 8-run never straddles a pad
#pragma unroll
                for (int i = 0; i < 8; i++) dst[i] = make_float2((float)gray16(w0, i), (float)gray16(w1, i));
            }
            if (PREFETCH && q + q_step < q_end) load_step(q + q_step);
        }
        // TMA_OUT: the first pass writes bufB, which holds the previous step's output tile: the engine must have read it.
        // The wait sits behind the staging (which wrote the OTHER buffer), so it is over long before thread 0 gets here.
        if (TMA_OUT && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();  // CTA wide: the previous step's output loop (all threads read every pair's result) is over
        const float2* z = fft_run_t<N, R0, R1, R2, R3, true, GT, PHD_ROWS_PK != 0, PHD_ROWS_TWPOW != 0>(bufA, bufB, twp, PAIRS, NP, N);
        __syncthreads();  // every pair's spectrum is complete
        if (TMA_OUT) {
            // tile in the staging buffer (free since the last pass read it): [fw][32 bytes], the 16-byte halves swapped in
            // rows 4..7 of every 8 (32B swizzle).  The buffers then swap roles: the next step stages into the buffer
            // that holds z now (free after this loop and its barrier) while the engine reads the tile.
            unsigned char* tile = reinterpret_cast<unsigned char*>(bufA);
            for (int k = threadIdx.x; k < fw; k += blockDim.x) {
                const int kc = k == 0 ? 0 : N - k;
                float4 v[2];
#pragma unroll
                for (int pr = 0; pr < 2; pr++) {
                    const float2 zk = z[pr * N + k], zc = z[pr * N + kc];
                    v[pr] = make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
                }
                if (k >= ktma) {  // columns beyond the engine's share: straight to global memory
                    float* o = reinterpret_cast<float*>(specT + (size_t)img * fw * P.Hp + 4 * q + (size_t)k * P.Hp);
                    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(v[0].x), "f"(v[0].y),
                                 "f"(v[0].z), "f"(v[0].w), "f"(v[1].x), "f"(v[1].y), "f"(v[1].z), "f"(v[1].w)
                                 : "memory");
                    continue;
                }
                float4* row = reinterpret_cast<float4*>(tile + 32 * k);
                const int sw = (k >> 2) & 1;
                row[sw] = v[0];
                row[sw ^ 1] = v[1];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
            __syncthreads();
            if (threadIdx.x == 0) {
                // whole boxes of 256 columns, then the ktma % 256 remaining ones through a map whose box is that wide
                int k0 = 0;
                for (; k0 + 256 <= ktma; k0 += 256) tma_store_box(tmap, tile + 32 * k0, 8 * q, k0, img);
                if (k0 < ktma) tma_store_box(tmap_tail, tile + 32 * k0, 8 * q, k0, img);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            float2* t = bufA; bufA = bufB; bufB = t;
            continue;
        }
        float2* out = specT + (size_t)img * fw * P.Hp + 2 * PAIRS * q;
        for (int k = threadIdx.x; k < fw; k += blockDim.x) {
            const int kc = k == 0 ? 0 : N - k;
            float4 v[PAIRS];
#pragma unroll
            for (int pr = 0; pr < PAIRS; pr++) {
                const float2 zk = z[pr * N + k], zc = z[pr * N + kc];
                v[pr] = make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
            }
            float* o = reinterpret_cast<float*>(out + (size_t)k * P.Hp);
            if (PAIRS == 2) {
                // one 32-byte store = one full sector per thread (sm_100 256-bit vector store)
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(v[0].x), "f"(v[0].y),
                             "f"(v[0].z), "f"(v[0].w), "f"(v[PAIRS - 1].x), "f"(v[PAIRS - 1].y), "f"(v[PAIRS - 1].z),
                             "f"(v[PAIRS - 1].w)
                             : "memory");
            } else {
#pragma unroll
                for (int pr = 0; pr < PAIRS; pr++) reinterpret_cast<float4*>(o)[pr] = v[pr];
            }
        }
        // the next step's staging writes bufA (last read by pass 3, barrier passed); its first pass writes bufB
        // only after the barrier that follows the staging, i.e. after every thread finished this output loop
    }
    // the CTA's shared memory must stay until the engine has read the last tile; the writes complete with the grid
    if (TMA_OUT && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int N, int R0, int R1, int R2, int R3, int THREADS, int PAIRS, int SEG = 16>
__global__ void __launch_bounds__(THREADS, rows_min_blocks(PAIRS * (2 * N + N / 16) * 8, THREADS)) k_rows_t(const uint8_t* __restrict__ rgb, DevParams P,
                                                    const float2* __restrict__ twp, float2* __restrict__ specT) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // steps blockIdx.x, blockIdx.x + gridDim.x, ...: CTAs that run together work on neighbouring rows, so the
    // pieces they write into the same 128-byte lines of the transposed spectrum meet in L2
    rows_walk<N, R0, R1, R2, R3, THREADS, PAIRS, true, false, SEG>(smem_raw, rgb, P, twp, specT, blockIdx.y, blockIdx.x,
                                                                   P.H / (2 * PAIRS), gridDim.x);
}
// the same walk with the TMA tensor store (tmap: rank 3 over specT, built by launch_rows_t)
template <int N, int R0, int R1, int R2, int R3, int THREADS, int SEG = 16>
__global__ void __launch_bounds__(THREADS, rows_min_blocks(rows_tma_smem(N), THREADS)) k_rows_tma(const uint8_t* __restrict__ rgb, DevParams P,
                                                    const float2* __restrict__ twp, float2* __restrict__ specT,
                                                    const __grid_constant__ CUtensorMap tmap,
                                                    const __grid_constant__ CUtensorMap tmap_tail, int ktma) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    rows_walk<N, R0, R1, R2, R3, THREADS, 2, true, true, SEG>(smem_raw, rgb, P, twp, specT, blockIdx.y, blockIdx.x, P.H / 4,
                                                               gridDim.x, &tmap, ktma, &tmap_tail);
}

// ------------------------------------------------------------------------------------------
// Front end + row FFT as ONE launch of persistent, role-switching CTAs.  Both stages stream the same packed RGB; run
// one after the other each leaves issue slots idle for a different reason (the front end stalls on its shared-memory
// atomics and its per-chunk barrier, the row FFT on the latency between its short passes), and the second reader
// finds nothing of the image in L2 any more.  Here every CTA (256 threads, the front end's shared memory, three per
// SM) pops tasks from one queue that lists, image by image, the image's front-end walks and its row-FFT tasks
// interleaved; at any moment the CTAs of an SM are a mix of both roles, working on the same few images, so the two
// instruction streams fill each other's stalls and the second read of an image hits L2.  Tasks never wait for one
// another (the column pass that needs both runs after this launch), so there is nothing to deadlock on.
// A row task p of `row_parts` takes the steps p, p + row_parts, ...: tasks popped together write neighbouring rows.
// Neither role prefetches into registers here (PREFETCH = false): with the 80 registers three CTAs per SM allow, the
// front end's 12 and the row stage's 24 prefetch registers were spilled (31 M local loads per 256 images, served by L2
// because the shared-memory carve-out leaves almost no L1); the other roles' CTAs cover the exposed load latency.
template <int N, int R0, int R1, int R2>
__global__ void __launch_bounds__(256, 3) k_front_rows(const uint8_t* __restrict__ rgb, DevParams P,
                                                       const unsigned char* __restrict__ tabs_g,
                                                       const unsigned char* __restrict__ exc,
                                                       u16* __restrict__ counts_chunk, u64* __restrict__ cells_g,
                                                       u32* __restrict__ span32, u64* __restrict__ span64,
                                                       ImageAcc* __restrict__ iacc, const float2* __restrict__ twp,
                                                       float2* __restrict__ specT, u32* __restrict__ queue, int nimg,
                                                       int row_parts) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ u32 s_task;
    const u32 tpi = (u32)(P.nspans + row_parts);  // tasks per image
    const u32 total = (u32)nimg * tpi;
    bool table_in_smem = false;
    for (;;) {
        __syncthreads();  // the previous task is done with the shared memory (and s_task has been read)
        if (threadIdx.x == 0) s_task = atomicAdd(queue, 1u);
        __syncthreads();
        const u32 t = s_task;
        if (t >= total) break;
        const int img = (int)(t / tpi), k = (int)(t - (u32)img * tpi);
        // tasks 0 .. tpi-1 of an image: row tasks spread evenly among the walks (r = row tasks before task k)
        const int r = (int)(((long long)k * row_parts) / tpi), r1 = (int)(((long long)(k + 1) * row_parts) / tpi);
        if (r1 > r) {
            rows_walk<N, R0, R1, R2, 1, 256, 2, false>(smem_raw, rgb, P, twp, specT, img, r, P.H / 4, row_parts);
            table_in_smem = false;
        } else {
            const int span = k - r, c_begin = span * P.cpp, c_end = min(c_begin + P.cpp, P.nchunks);
            fe_walk<256, false, PHD_NCS_SMALL, true, true>(smem_raw, rgb, P, tabs_g, exc, img, span, c_begin, c_end,
                                                            !table_in_smem, counts_chunk, cells_g, span32, span64, iacc);
            table_in_smem = true;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Rows, generic (any width, runtime radix plan): a CTA walks strided groups of 2*NP rows (NP = 2: quads).  The rows'
// bytes are staged into shared memory (16-byte vector loads when the rows are 16-byte aligned), each thread then
// converts "its" pixels of the NP row pairs to gray numerators, NP packed complex sequences are transformed, and every
// spectrum entry of the group leaves as one 32-byte sector (NP = 2) or half a sector (NP = 1: rows of 6401..12800
// pixels, whose two pairs no longer fit shared memory).  Rows past the image bottom count as gray 0.5 (a zero sequence
// after the bias) and land in the Hp padding of the transposed spectrum.
// N > 0: the width is the compile-time length N with plan R0 R1 R2 [R3] -- planned widths that k_rows_t / k_rows_tma do
// not take: those that are not a multiple of 8 (900, 1050), and the four-pass widths 7680 / 5120 (PHD_FFT_PLANS4, NP = 1):
// same staging, compile-time passes, THREADS threads.
// gray32 != nullptr: the input is a plane of floats (general-input route: (gray - 0.5) * 255000 of one image) instead
// of packed bytes.
template <int NP, int N = 0, int R0 = 1, int R1 = 1, int R2 = 1, int THREADS = 512, int R3 = 1>
__global__ void __launch_bounds__(THREADS) k_rows_generic(const uint8_t* __restrict__ rgb, DevParams P, FftPlan pl,
                                                              float2* __restrict__ specT,
                                                              const float* __restrict__ gray32 = nullptr) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int W = N > 0 ? N : P.W;
    const int L = (N == 0 && pl.m > 0) ? pl.m : W;       // sequence stride: the padded length of a Bluestein plan
    float2* bufA = reinterpret_cast<float2*>(smem_raw);  // [NP][L]
    float2* bufB = bufA + NP * L;                        // [NP][L]; first holds the raw bytes of the rows (6 W NP <= 8 L NP)
    unsigned char* raw = reinterpret_cast<unsigned char*>(bufB);
    const int img = blockIdx.y;
    const int ngroups = P.Hp / (2 * NP);
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    const int row_bytes = 3 * W;
    // widest load every row start allows: 16 bytes, else 8 (e.g. 1080-pixel rows), else 4, else single bytes
    const int gran = (int)((reinterpret_cast<uintptr_t>(base) | (uintptr_t)row_bytes) & 15);
    const int vw = gran == 0 ? 16 : ((gran & 7) == 0 ? 8 : ((gran & 3) == 0 ? 4 : 1));
    for (int q = blockIdx.x; q < ngroups; q += gridDim.x) {
        __syncthreads();  // the previous group's output loop has finished reading bufB
        for (int r = 0; r < 2 * NP && !gray32; r++) {
            const int row = 2 * NP * q + r;
            unsigned char* dst = raw + (size_t)r * row_bytes;
            if (row >= P.H) continue;
            const uint8_t* src = base + (size_t)row * row_bytes;
            if (vw == 16) {
                const uint4* s4 = reinterpret_cast<const uint4*>(src);
                uint4* d4 = reinterpret_cast<uint4*>(dst);
                for (int i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) d4[i] = __ldg(s4 + i);
            } else if (vw == 8) {
                const uint2* s2 = reinterpret_cast<const uint2*>(src);
                uint2* d2 = reinterpret_cast<uint2*>(dst);
                for (int i = threadIdx.x; i < row_bytes / 8; i += blockDim.x) d2[i] = __ldg(s2 + i);
            } else if (vw == 4) {
                const u32* s1 = reinterpret_cast<const u32*>(src);
                u32* d1 = reinterpret_cast<u32*>(dst);
                for (int i = threadIdx.x; i < row_bytes / 4; i += blockDim.x) d1[i] = __ldg(s1 + i);
            } else {
                for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) dst[i] = __ldg(src + i);
            }
        }
        __syncthreads();
        // Four pixels per task where the row allows it (width and row bytes multiples of 4, bytes not a float plane):
        // three aligned words per row instead of twelve byte loads -- the staged compile-time widths (1080, 3000, ...:
        // portrait video and photos) spent more LSU wavefronts on this conversion than on a pass of the transform.
        const bool quad = !gray32 && (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(raw) | (uintptr_t)row_bytes) % 4 == 0);
        for (int idx = threadIdx.x; quad && idx < NP * (W / 4); idx += blockDim.x) {
            const int pair = idx / (W / 4), x4 = idx - pair * (W / 4);
            int g[2][4];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = 2 * NP * q + 2 * pair + h;
                const u32* pw = reinterpret_cast<const u32*>(raw + (size_t)(2 * pair + h) * row_bytes) + 3 * x4;
                const u32 a = pw[0], b = pw[1], c = pw[2];  // R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
                const bool in = row < P.H;
                g[h][0] = in ? 299 * (int)(a & 255u) + 587 * (int)((a >> 8) & 255u) + 114 * (int)((a >> 16) & 255u) - PHD_GRAY_BIAS : 0;
                g[h][1] = in ? 299 * (int)(a >> 24) + 587 * (int)(b & 255u) + 114 * (int)((b >> 8) & 255u) - PHD_GRAY_BIAS : 0;
                g[h][2] = in ? 299 * (int)((b >> 16) & 255u) + 587 * (int)(b >> 24) + 114 * (int)(c & 255u) - PHD_GRAY_BIAS : 0;
                g[h][3] = in ? 299 * (int)((c >> 8) & 255u) + 587 * (int)((c >> 16) & 255u) + 114 * (int)(c >> 24) - PHD_GRAY_BIAS : 0;
            }
            float2* dst = bufA + pair * L + 4 * x4;
#pragma unroll
            for (int i = 0; i < 4; i++) dst[i] = make_float2((float)g[0][i], (float)g[1][i]);
        }
        for (int idx = threadIdx.x; !quad && idx < NP * W; idx += blockDim.x) {
            const int pair = idx / W, x = idx - pair * W;
            int g[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int row = 2 * NP * q + 2 * pair + h;
                const unsigned char* px = raw + (size_t)(2 * pair + h) * row_bytes + 3 * x;
                g[h] = (row < P.H && !gray32) ? 299 * (int)px[0] + 587 * (int)px[1] + 114 * (int)px[2] - PHD_GRAY_BIAS : 0;
            }
            float2 v = make_float2((float)g[0], (float)g[1]);
            if (gray32) {
                const int row = 2 * NP * q + 2 * pair;
                v.x = row < P.H ? __ldg(gray32 + (size_t)row * W + x) : 0.f;
                v.y = row + 1 < P.H ? __ldg(gray32 + (size_t)(row + 1) * W + x) : 0.f;
            }
            bufA[pair * L + x] = v;
        }
        __syncthreads();
        const float2* z;
        if constexpr (N > 0) {
            // table twiddles here: the product tree measured slower in this variant (3000-pixel rows 55 -> 84 us per image)
            z = fft_run_t<N, R0, R1, R2, R3, false, THREADS / NP, false>(bufA, bufB, pl.twp, NP, N, N);
            __syncthreads();
        } else {
            z = pl.m > 0 ? fft_run_blue(pl, bufA, bufB, NP, L) : fft_run_rt(pl, bufA, bufB, NP, W);
        }
        float2* out = specT + (size_t)img * P.fw * P.Hp + 2 * NP * q;
        for (int k = threadIdx.x; k < P.fw; k += blockDim.x) {
            const int kc = k == 0 ? 0 : W - k;
            float4 v[NP];
#pragma unroll
            for (int pr = 0; pr < NP; pr++) {
                const float2 zk = z[pr * L + k], zc = z[pr * L + kc];
                v[pr] = make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
            }
            float* o = reinterpret_cast<float*>(out + (size_t)k * P.Hp);
            if (NP == 2) {
                asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(v[0].x), "f"(v[0].y),
                             "f"(v[0].z), "f"(v[0].w), "f"(v[NP - 1].x), "f"(v[NP - 1].y), "f"(v[NP - 1].z), "f"(v[NP - 1].w)
                             : "memory");
            } else {
                *reinterpret_cast<float4*>(o) = v[0];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// mbarrier / bulk-copy helpers (PTX; SASS shows UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// ---- epilogue pieces shared by both column kernels ------------------------------------------------------------
// res: transformed columns in shared memory (column c at res + c*cs), values scaled by 255000 (gray numerators);
// map: bin ids of the same columns (shared or global, column c at map + c*ms).
#define PHD_POWER_SCALE (255000.0 * 255000.0)

// DC coefficient: the transform ran on gray - 0.5; the reference transforms gray - (Br+Bg+Bb)/3
// (interface.c:78, blur_profile.c:233-238).  By linearity only X[0,0] differs; it is rebuilt from the exact channel
// sums and stored so that the common power arithmetic below reproduces it.
__device__ __forceinline__ void cols_fix_dc(const DevParams& P, const ImageAcc* __restrict__ iacc, int img, float2* res00) {
    const ImageAcc a = iacc[img];
    if (a.dc_valid) {  // general-input route: computed from the doubles (f64path.cu)
        *res00 = make_float2((float)(a.dc * 255000.0), 0.f);
        return;
    }
    const double np = (double)P.npx;
    const double avg = ((double)a.sum[0] / 255.0 / np + (double)a.sum[1] / 255.0 / np + (double)a.sum[2] / 255.0 / np) / 3.0;
    const double gsum = (299.0 * (double)a.sum[0] + 587.0 * (double)a.sum[1] + 114.0 * (double)a.sum[2]) / 255000.0;
    const double dc = gsum - np * avg;
    *res00 = make_float2((float)(dc * 255000.0), 0.f);
}

// power -> (p < 1 ? 0 : ln p) as fixed point 2^-20 -> polar bin, run-length merged per thread (bins change slowly
// along a column) into the shared integer bins.  Returns the running max of the RAW power (re^2 + im^2).
// GT > 0: column c belongs to the thread group c (threads c*GT .. c*GT+GT-1, the group that transformed it).
// FULL: H is a multiple of SEG (no bounds check per element).
template <int SEG, int GT, bool FULL = false>
__device__ __forceinline__ float cols_accumulate(int H, int ncol, const float2* res, int cs, const u16* map, int ms,
                                                 u32* bin_lo, u32* bin_hi, float mymax) {
    const float thr = (float)PHD_POWER_SCALE;                            // p >= 1  <=>  raw >= 255000^2
    const float kq = (float)(0.69314718055994530942 * (1 << PHD_LN_SHIFT));  // ln 2 * 2^20
    const float lgq = (float)(-2.0 * 17.960137721520944 * 0.69314718055994530942 * (1 << PHD_LN_SHIFT));  // ln(1/255000^2) * 2^20
    const int segs = (H + SEG - 1) / SEG;
    const int first = GT > 0 ? (int)threadIdx.x % GT : (int)threadIdx.x;
    const int total = GT > 0 ? ((int)threadIdx.x / GT < ncol ? segs : 0) : ncol * segs;
    for (int task = first; task < total; task += (GT > 0 ? GT : (int)blockDim.x)) {
        const int c = GT > 0 ? (int)threadIdx.x / GT : task / segs, k0 = (task - (GT > 0 ? 0 : c * segs)) * SEG;
        const float2* rp = res + c * cs + k0;
        const u16* mp = map + c * ms + k0;
        int run_bin = -1;
        u32 run_sum = 0;
#pragma unroll
        for (int i = 0; i < SEG; i++) {
            if (FULL || k0 + i < H) {
                const float2 v = rp[i];
                const float raw = fmaf(v.x, v.x, v.y * v.y);
                mymax = fmaxf(mymax, raw);
                if (raw >= thr) {
                    // ln p * 2^20 in one multiply-add; clamped at 0: it can round a hair below 0 when p == 1
                    const u32 q = (u32)__float2int_rn(fmaxf(fmaf(__log2f(raw), kq, lgq), 0.f));
                    const int bin = mp[i];
                    if (bin != run_bin) {
                        if (run_sum) {
                            atomicAdd(&bin_lo[run_bin], run_sum & 0x1fffu);
                            atomicAdd(&bin_hi[run_bin], run_sum >> 13);
                        }
                        run_bin = bin;
                        run_sum = 0;
                    }
                    run_sum += q;
                }
            }
        }
        if (run_sum) {
            atomicAdd(&bin_lo[run_bin], run_sum & 0x1fffu);
            atomicAdd(&bin_hi[run_bin], run_sum >> 13);
        }
    }
    return mymax;
}

// One flush per CTA: max power (scaled to p) and the shared bins into the image's 64-bit global bins.
__device__ __forceinline__ void cols_flush(const DevParams& P, int img, float mymax, const u32* bin_lo, const u32* bin_hi,
                                           float* sh_max, u64* __restrict__ binsum, u32* __restrict__ maxpow) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((threadIdx.x & 31) == 0) sh_max[threadIdx.x >> 5] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, sh_max[w]);
        atomicMax(&maxpow[img], __float_as_uint(m * (float)(1.0 / PHD_POWER_SCALE)));
    }
    u64* dst = binsum + (size_t)img * P.nbins;
    for (int b = threadIdx.x; b < P.nbins; b += blockDim.x) {
        const u64 v = ((u64)bin_hi[b] << 13) + bin_lo[b];
        if (v) atomicAdd(&dst[b], v);
    }
}

// test hook: the power spectrum itself, row major
__device__ __forceinline__ void cols_write_power(const DevParams& P, int img, int x0, int ncol, const float2* res, int cs,
                                                 float* __restrict__ power_out) {
    const float inv_scale2 = (float)(1.0 / PHD_POWER_SCALE);
    for (int idx = threadIdx.x; idx < ncol * P.H; idx += blockDim.x) {
        const int c = idx / P.H, k = idx - c * P.H;
        const float2 v = res[c * cs + k];
        power_out[((size_t)img * P.H + k) * P.fw + x0 + c] = (v.x * v.x + v.y * v.y) * inv_scale2;
    }
}

// Columns, specialised.  A CTA walks `gpc` consecutive groups of NB contiguous columns of ONE image: the shared
// bins are zeroed and flushed once per CTA, and while a group's epilogue runs the bulk-copy engine already
// fetches the next group's columns (into the buffer the last FFT pass no longer reads) and bin-id slices.
// Requires Hp == N (H % 4 == 0), cols_t_ok<N, NB>() (16-byte multiples for the bulk copies) and a 3-pass plan.
// Two CTAs fit an SM either way.  MINB is only the register-allocation hint of __launch_bounds__ (0 = none): measured,
// the 1080-point kernel is fastest with 1 (58 registers; 0 -> 48 and 2 -> 60 are 1-2 % slower), the longer ones with 2.
template <int N, int R0, int R1, int R2, int R3, int NB, int MINB, bool WRITE_POWER>
__global__ void __launch_bounds__(kColThreads, MINB) k_cols_t(DevParams P, const float2* __restrict__ tw,
                                                        const float2* __restrict__ specT,
                                                        const u16* __restrict__ binmapT,
                                                        const ImageAcc* __restrict__ iacc, u64* __restrict__ binsum,
                                                        u32* __restrict__ maxpow, float* __restrict__ power_out, int gpc) {
    static_assert(R3 == 1, "the prefetch below assumes the result lands in bufB");
    constexpr int GT = kColThreads / NB;  // thread group of one column (see seq_sync)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + NB * N;
    u16* smap = reinterpret_cast<u16*>(bufB + NB * N);  // [2][NB*N]
    u32* bin_lo = reinterpret_cast<u32*>(smap + 2 * NB * N);
    u32* bin_hi = bin_lo + P.nbins;
    __shared__ __align__(8) u64 bar;
    __shared__ float sh_max[kColThreads / 32];

    const int img = blockIdx.y;
    const int ngroups = (P.fw + NB - 1) / NB;
    const int g_begin = blockIdx.x * gpc, g_end = min(g_begin + gpc, ngroups);
    auto fetch = [&](int g, int slot) {  // thread 0 only
        const int x0 = g * NB, ncol = min(NB, P.fw - x0);
        const u32 cbytes = (u32)(ncol * N * sizeof(float2)), mbytes = ((u32)(ncol * N * sizeof(u16)) + 15u) & ~15u;
        mbar_expect_tx(&bar, cbytes + mbytes);
        bulk_g2s(bufA, specT + ((size_t)img * P.fw + x0) * N, cbytes, &bar);
        bulk_g2s(smap + slot * NB * N, binmapT + (size_t)x0 * N, mbytes, &bar);
    };
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fetch(g_begin, 0);
    }
    if (!WRITE_POWER)
        for (int b = threadIdx.x; b < 2 * P.nbins; b += blockDim.x) bin_lo[b] = 0;
    __syncthreads();  // barrier init visible to every waiter
    float mymax = 0.f;
    for (int g = g_begin; g < g_end; g++) {
        const int it = g - g_begin;
        const int x0 = g * NB, ncol = min(NB, P.fw - x0);
        mbar_wait(&bar, it & 1);
        float2* res = fft_run_t<N, R0, R1, R2, R3, false, GT, kColsPacked, PHD_COLS_TWPOW != 0>(bufA, bufB, tw, ncol, N, N);
        __syncthreads();
        if (threadIdx.x == 0) {
            if (g + 1 < g_end) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // bufA was written by pass 2
                fetch(g + 1, (it + 1) & 1);
            }
            if (!WRITE_POWER && x0 == 0) cols_fix_dc(P, iacc, img, res);  // thread 0 also owns point (0,0) below
        }
        if (WRITE_POWER) cols_write_power(P, img, x0, ncol, res, N, power_out);
        else {
            constexpr int SEG = (NB * N + kColThreads - 1) / kColThreads;
            mymax = cols_accumulate<SEG, GT, N % SEG == 0>(N, ncol, res, N, smap + (it & 1) * NB * N, N, bin_lo, bin_hi, mymax);
        }
        // bufB is rewritten by the next group's first pass: a column is binned and rewritten by the thread group
        // that transformed it, so the hand-over is that group's barrier (the test hook spreads its writes over
        // all threads).  The only CTA-wide barrier of the loop is the one above that frees bufA for the prefetch;
        // giving each group its own mbarrier and prefetch removes that one too but measured 0.5% slower.
        if (WRITE_POWER) __syncthreads();
        else seq_sync<GT>();
    }
    if (!WRITE_POWER) cols_flush(P, img, mymax, bin_lo, bin_hi, sh_max, binsum, maxpow);
}

// Columns, generic: runtime radix plan, plain loads, bin ids read from global memory.
template <bool WRITE_POWER>
__global__ void __launch_bounds__(kColThreads, 2) k_cols_generic(DevParams P, FftPlan pl, int TC,
                                                              const float2* __restrict__ specT,
                                                              const u16* __restrict__ binmapT,
                                                              const ImageAcc* __restrict__ iacc,
                                                              u64* __restrict__ binsum, u32* __restrict__ maxpow,
                                                              float* __restrict__ power_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int Hp = P.Hp;
    const int L = pl.m > 0 ? pl.m : Hp;  // column stride in shared memory: the padded length of a Bluestein plan
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)TC * L;
    u32* bin_lo = reinterpret_cast<u32*>(bufB + (size_t)TC * L);
    u32* bin_hi = bin_lo + P.nbins;
    __shared__ float sh_max[kColThreads / 32];

    const int img = blockIdx.y;
    const int x0 = blockIdx.x * TC;
    const int ncol = min(TC, P.fw - x0);
    const float2* src = specT + ((size_t)img * P.fw + x0) * Hp;
    if (!WRITE_POWER)
        for (int b = threadIdx.x; b < 2 * P.nbins; b += blockDim.x) bin_lo[b] = 0;
    for (int idx = threadIdx.x; idx < ncol * Hp; idx += blockDim.x) {
        const int c = idx / Hp, k = idx - c * Hp;
        bufA[c * L + k] = src[idx];
    }
    __syncthreads();
    float2* res = pl.m > 0 ? fft_run_blue(pl, bufA, bufB, ncol, L) : fft_run_rt(pl, bufA, bufB, ncol, L);
    if (WRITE_POWER) {
        cols_write_power(P, img, x0, ncol, res, L, power_out);
        return;
    }
    if (x0 == 0 && threadIdx.x == 0) cols_fix_dc(P, iacc, img, res);  // thread 0 also owns point (0,0) below
    const float mymax = cols_accumulate<8, 0>(P.H, ncol, res, L, binmapT + (size_t)x0 * Hp, Hp, bin_lo, bin_hi, 0.f);
    __syncthreads();
    cols_flush(P, img, mymax, bin_lo, bin_hi, sh_max, binsum, maxpow);
}

// ------------------------------------------------------------------------------------------
__device__ int newton_isqrt(double val) {
    if (val == 0.0) return 0;
    double x = val;
    for (;;) {
        const double s = __dmul_rn(0.5, __dadd_rn(x, __ddiv_rn(val, x)));
        if (fabs(__dsub_rn(s, x)) < 1.0) return (int)s;
        x = s;
    }
}

// Bin id of spectrum element (row k, column x), stored transposed: map[x*Hp + k].
__global__ void k_bin_map(int W, int H, int Hp, int nr, int na, u16* __restrict__ map, int* __restrict__ counts) {
    const int fw = W / 2 + 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)fw * H) return;
    const int x = (int)(i / H), k = (int)(i - (long long)x * H);
    const int hb = H / 2 + (H % 2 == 1 ? 1 : 0);
    // rows k >= H - hb are written last by the reference loop (bottom half overwrites the middle row)
    int y;
    double phi;
    if (k >= H - hb) { y = H - 1 - k; phi = atan2((double)y, (double)x); }
    else { y = k; phi = -atan2((double)y, (double)x); }
    const int r2 = x * x + y * y;
    const double REF_PI = 3.14159265;
    const double half_pi = __dmul_rn(REF_PI, (double)0.5f);
    const int pb = (int)__dmul_rn(__ddiv_rn(__dadd_rn(phi, half_pi), REF_PI), (double)(na - 1));
    const double rbs = (double)((fw * fw + H * H / 4) / (nr * nr));
    int rb = newton_isqrt(__ddiv_rn((double)r2, rbs));
    if (rb == nr) rb--;
    int bin = pb * nr + rb;
    bin = min(max(bin, 0), na * nr - 1);
    map[(size_t)x * Hp + k] = (u16)bin;
    atomicAdd(&counts[bin], 1);
}

// ------------------------------------------------------------------------------------------
// Transforms that do not fit shared memory (image sides beyond 12,800 / ~11,000 pixels; src/utilities.c:11-13 admits
// 120 MP at aspect <= 5, i.e. sides up to 24,494): four-step through HBM.  L = n1 * n2,
//   X[k1 + n1 k2] = sum_j2 W_L^(j2 k1) [ sum_j1 x[j1 n2 + j2] W_n1^(j1 k1) ] W_n2^(j2 k2):
// stage A transforms the n2 strided sub-sequences of length n1 (a tile of them per CTA, coalesced along j2), applies the
// twiddle and stores [k1][j2]; stage B transforms the n1 contiguous sub-sequences of length n2 and stores X in natural
// order.  Both run the runtime-radix passes on their tiles in shared memory.  A length without a usable factorisation
// goes through Bluestein around the same two stages.  A rare path (panoramas): correctness first, several HBM passes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_long_stage_a(const float2* __restrict__ src, float2* __restrict__ dst, size_t stride,
                                                      FftPlan p1, int n1, int n2, int L, const float2* __restrict__ twL,
                                                      int TB) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)TB * n1;
    const size_t seq = blockIdx.y;
    const int j20 = blockIdx.x * TB, nt = min(TB, n2 - j20);
    const float2* s = src + seq * stride;
    for (int idx = threadIdx.x; idx < nt * n1; idx += blockDim.x) {
        const int j1 = idx / nt, t = idx - j1 * nt;
        bufA[t * n1 + j1] = s[(size_t)j1 * n2 + j20 + t];
    }
    __syncthreads();
    const float2* r = fft_run_rt(p1, bufA, bufB, nt, n1);
    float2* d = dst + seq * stride;
    for (int idx = threadIdx.x; idx < nt * n1; idx += blockDim.x) {
        const int k1 = idx / nt, t = idx - k1 * nt, j2 = j20 + t;
        const float2 w = __ldg(&twL[(int)(((long long)j2 * k1) % L)]);
        d[(size_t)k1 * n2 + j2] = cmulf<false>(r[t * n1 + k1], w);
    }
}

__global__ void __launch_bounds__(256) k_long_stage_b(const float2* __restrict__ src, float2* __restrict__ dst, size_t stride,
                                                      FftPlan p2, int n1, int n2, int TB) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)TB * n2;
    const size_t seq = blockIdx.y;
    const int k10 = blockIdx.x * TB, nt = min(TB, n1 - k10);
    const float2* s = src + seq * stride + (size_t)k10 * n2;
    for (int idx = threadIdx.x; idx < nt * n2; idx += blockDim.x) bufA[idx] = s[idx];
    __syncthreads();
    const float2* r = fft_run_rt(p2, bufA, bufB, nt, n2);
    float2* d = dst + seq * stride;
    for (int idx = threadIdx.x; idx < nt * n2; idx += blockDim.x) {
        const int k2 = idx / nt, t = idx - k2 * nt;
        d[(size_t)(k10 + t) + (size_t)n1 * k2] = r[t * n2 + k2];
    }
}

// Bluestein around the long transform (see fft_run_blue): x w, zero padded | conj(. * bhat) | conj(.) * w
__global__ void k_long_chirp_in(float2* __restrict__ data, size_t stride, int n, int L, const float2* __restrict__ chirp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    float2* p = data + (size_t)blockIdx.y * stride + k;
    *p = k < n ? cmulf<false>(*p, __ldg(&chirp[k])) : make_float2(0.f, 0.f);
}
__global__ void k_long_mul_bhat(float2* __restrict__ data, size_t stride, int L, const float2* __restrict__ bhat) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    float2* p = data + (size_t)blockIdx.y * stride + k;
    const float2 v = cmulf<false>(*p, __ldg(&bhat[k]));
    *p = make_float2(v.x, -v.y);
}
__global__ void k_long_chirp_out(float2* __restrict__ data, size_t stride, int n, const float2* __restrict__ chirp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float2* p = data + (size_t)blockIdx.y * stride + k;
    const float2 v = *p;
    *p = cmulf<false>(make_float2(v.x, -v.y), __ldg(&chirp[k]));
}
// tables of a long plan: W_L^k, and for Bluestein the chirp and the circular conjugate chirp (to be transformed)
__global__ void k_long_tables(float2* __restrict__ twL, float2* __restrict__ chirp, float2* __restrict__ bpad, int n, int L,
                              int blue) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    double sn, cs;
    sincospi(2.0 * (double)k / (double)L, &sn, &cs);
    twL[k] = make_float2((float)cs, (float)(-sn));
    if (!blue) return;
    // bpad[j] = exp(+i pi j^2 / n) for |j| < n laid out circularly, 0 elsewhere
    const int j = k < n ? k : (k > L - n ? L - k : -1);
    if (j >= 0) {
        sincospi((double)(((long long)j * j) % (2LL * n)) / (double)n, &sn, &cs);
        bpad[k] = make_float2((float)cs, (float)sn);
        if (k < n) chirp[k] = make_float2((float)cs, (float)(-sn));
    } else {
        bpad[k] = make_float2(0.f, 0.f);
    }
}
__global__ void k_long_scale(float2* __restrict__ data, int L, float f) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < L) data[k] = make_float2(data[k].x * f, data[k].y * f);
}

// rows of one image as packed pairs: data[pair][x] = (gray(row 2 pair, x), gray(row 2 pair + 1, x)) numerators
__global__ void __launch_bounds__(256) k_long_gray(const uint8_t* __restrict__ rgb, const float* __restrict__ gray32, DevParams P,
                                                   float2* __restrict__ data, size_t stride) {
    const int pair = blockIdx.y;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < P.W; x += gridDim.x * blockDim.x) {
        float g[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int row = 2 * pair + h;
            if (row >= P.H) g[h] = 0.f;
            else if (gray32) g[h] = __ldg(gray32 + (size_t)row * P.W + x);
            else {
                const uint8_t* px = rgb + ((size_t)row * P.W + x) * 3;
                g[h] = (float)(299 * (int)__ldg(px) + 587 * (int)__ldg(px + 1) + 114 * (int)__ldg(px + 2) - PHD_GRAY_BIAS);
            }
        }
        data[(size_t)pair * stride + x] = make_float2(g[0], g[1]);
    }
}
// split the pair spectra into the two rows' half spectra and write them transposed (see k_rows_generic)
__global__ void __launch_bounds__(256) k_long_rows_out(const float2* __restrict__ data, size_t stride, DevParams P,
                                                       float2* __restrict__ specT) {
    const int pair = blockIdx.y;
    const float2* z = data + (size_t)pair * stride;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.fw; k += gridDim.x * blockDim.x) {
        const int kc = k == 0 ? 0 : P.W - k;
        const float2 zk = z[k], zc = z[kc];
        *reinterpret_cast<float4*>(specT + (size_t)k * P.Hp + 2 * pair) =
            make_float4(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y), 0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
    }
}
__global__ void __launch_bounds__(256) k_long_cols_in(const float2* __restrict__ specT, DevParams P, float2* __restrict__ data,
                                                      size_t stride) {
    const int x = blockIdx.y;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.H; k += gridDim.x * blockDim.x)
        data[(size_t)x * stride + k] = specT[(size_t)x * P.Hp + k];
}
// one CTA per transformed column (in HBM): the same epilogue as the shared-memory column kernels
template <bool WRITE_POWER>
__global__ void __launch_bounds__(kColThreads) k_long_cols_epi(DevParams P, float2* __restrict__ data, size_t stride,
                                                               const u16* __restrict__ binmapT,
                                                               const ImageAcc* __restrict__ iacc, u64* __restrict__ binsum,
                                                               u32* __restrict__ maxpow, float* __restrict__ power_out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u32* bin_lo = reinterpret_cast<u32*>(smem_raw);
    u32* bin_hi = bin_lo + P.nbins;
    __shared__ float sh_max[kColThreads / 32];
    const int x = blockIdx.x;
    float2* res = data + (size_t)x * stride;
    if (WRITE_POWER) {
        cols_write_power(P, 0, x, 1, res, 0, power_out);
        return;
    }
    for (int b = threadIdx.x; b < 2 * P.nbins; b += blockDim.x) bin_lo[b] = 0;
    if (x == 0 && threadIdx.x == 0) cols_fix_dc(P, iacc, 0, res);
    __syncthreads();
    const float mymax = cols_accumulate<8, 0>(P.H, 1, res, 0, binmapT + (size_t)x * P.Hp, 0, bin_lo, bin_hi, 0.f);
    __syncthreads();
    cols_flush(P, 0, mymax, bin_lo, bin_hi, sh_max, binsum, maxpow);
}

// Tensor map of the transposed spectra of `nimg` images for the row kernel's TMA store: floats, rank 3 =
// (2 Hp floats of a spectrum column | fw columns | images), box = (8 floats = 4 rows, up to 256 columns, 1 image), 32-byte
// swizzle (the kernel writes its tile with the same pattern, conflict free).  false: no driver entry point / refused --
// the caller falls back to the kernel with direct stores.
typedef CUresult (*PhdEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
bool make_spec_tmap(CUtensorMap* tm, float2* specT, const DevParams& P, int nimg, int box_cols) {
    static PhdEncodeTiled encode = []() -> PhdEncodeTiled {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (getenv("PHD_NO_TMA_STORE")) return nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<PhdEncodeTiled>(fn);
    }();
    if (!encode) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)2 * P.Hp, (cuuint64_t)P.fw, (cuuint64_t)nimg};
    const cuuint64_t strides[2] = {(cuuint64_t)P.Hp * sizeof(float2), (cuuint64_t)P.fw * P.Hp * sizeof(float2)};
    const cuuint32_t box[3] = {8, (cuuint32_t)box_cols, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, specT, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (getenv("PHD_DEBUG")) fprintf(stderr, "[phd] spectrum tensor map (%d images, fw %d, Hp %d): %s\n", nimg, P.fw, P.Hp, rc == CUDA_SUCCESS ? "ok" : "REFUSED");
    return rc == CUDA_SUCCESS;
}

// ---- dispatch tables of the specialised shapes -------------------------------------------------
template <int N, int R0, int R1, int R2, int R3, int PAIRS>
void launch_rows_t(const uint8_t* rgb, const DevParams& P, int nimg, const float2* tw, float2* specT, cudaStream_t st) {
    constexpr int SEG = rows_seg<N, R0>();
    constexpr int THREADS = ((PAIRS * (N / SEG) + 127) / 128) * 128;  // one staging task (SEG pixels of a row pair) per thread
    const size_t smem = (size_t)PAIRS * (2 * N + N / 16) * sizeof(float2);
    PHD_ALLOW_SMEM((k_rows_t<N, R0, R1, R2, R3, THREADS, PAIRS, SEG>), (int)smem);
    // CTAs walk strided steps.  gx CTAs per image: the launch runs in ceil(gx * nimg / resident) waves of
    // ceil(nsteps / gx) steps (+ about one step's worth of start-up per CTA: the first loads are not prefetched); pick
    // the gx that minimises that product.  One CTA per image -- what "about four waves in all" gave for large batches --
    // left the last of 4.6 waves 40 % empty: 8 % of the kernel at 2048 1080p images.
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 2048 / THREADS) per_sm = 2048 / THREADS;
    if (per_sm < 1) per_sm = 1;
    const int nsteps = P.H / (2 * PAIRS);
    const long long resident = (long long)per_sm * 148;
    int gx = 1;
    long long best = -1;
    for (int c = 1; c <= nsteps; c++) {
        const int per_cta = (nsteps + c - 1) / c;
        if (per_cta < 4 && c > 1) break;  // shorter walks only add start-up cost
        const long long waves = ((long long)c * nimg + resident - 1) / resident;
        const long long cost = waves * (per_cta + 1);
        if (best < 0 || cost < best) { best = cost; gx = c; }
    }
    if constexpr (PAIRS == 2 && PHD_ROWS_TMA != 0) {
        // How many spectrum columns leave through the engine; the rest is stored directly.  Measured at 1080p (961 columns,
        // three CTAs per SM; ms per 2048 images): 0: 8.08, 256: 7.85, 512: 7.54, 768: 7.33, all: 7.3 .. 9.4 and erratic --
        // with three CTAs feeding it 32-byte box rows the engine itself becomes the limit, so LSU and engine share the
        // work (80 %, in whole boxes).  Rows of 3840 / 6000 pixels (one CTA per SM) are fastest with everything through
        // the engine (4K: 15.9 -> 15.3 ms per 256 images; 24 MP: 20.7 with 2304 of 3001 columns, 20.0 with all).
        static const int ktma_env = getenv("PHD_ROWS_KTMA") ? atoi(getenv("PHD_ROWS_KTMA")) : -1;
        int ktma = rows_tma_smem(N) <= 110 * 1024 ? (P.fw * 4 / 5) / 256 * 256 : P.fw;
        if (ktma_env >= 0) ktma = ktma_env < P.fw ? ktma_env : P.fw;
        CUtensorMap tm, tm_tail;
        if (ktma > 0 && make_spec_tmap(&tm, specT, P, nimg, 256) &&
            make_spec_tmap(&tm_tail, specT, P, nimg, ktma % 256 ? ktma % 256 : 256)) {
            PHD_ALLOW_SMEM((k_rows_tma<N, R0, R1, R2, R3, THREADS, SEG>), rows_tma_smem(N));
            k_rows_tma<N, R0, R1, R2, R3, THREADS, SEG><<<dim3(gx, nimg), THREADS, rows_tma_smem(N), st>>>(rgb, P, tw, specT, tm, tm_tail, ktma);
            return;
        }
    }
    k_rows_t<N, R0, R1, R2, R3, THREADS, PAIRS, SEG><<<dim3(gx, nimg), THREADS, smem, st>>>(rgb, P, tw, specT);
}

template <int N, int R0, int R1, int R2, int R3, int NB, int MINB = 2>
void launch_cols_t(const DevParams& P, int nimg, const float2* tw, const float2* specT, const u16* binmapT,
                   Workspace& ws, float* power_out, cudaStream_t st) {
    const size_t smem = (size_t)2 * NB * N * sizeof(float2) + (size_t)2 * NB * N * sizeof(u16) + (size_t)2 * P.nbins * sizeof(u32);
    PHD_ALLOW_SMEM((k_cols_t<N, R0, R1, R2, R3, NB, MINB, false>), 200 * 1024);
    PHD_ALLOW_SMEM((k_cols_t<N, R0, R1, R2, R3, NB, MINB, true>), 200 * 1024);
    // groups per CTA: long walks amortise the bin zero/flush; pick the walk length whose CTA count fills whole
    // waves of the machine (2 CTAs of this size per SM, 148 SMs)
    const int ngroups = (P.fw + NB - 1) / NB;
    const int slots = 2 * 148;
    int best = 1;
    double best_cost = 1e30;
    for (int gpc = 1; gpc <= 16 && gpc <= ngroups; gpc++) {
        const long long ctas = (long long)((ngroups + gpc - 1) / gpc) * nimg;
        const long long waves = (ctas + slots - 1) / slots;
        const double cost = (double)waves * (gpc + 0.8);  // 0.8 group-equivalents of per-CTA overhead
        if (cost < best_cost - 1e-9) { best_cost = cost; best = gpc; }
    }
    dim3 grid((ngroups + best - 1) / best, nimg);
    if (power_out)
        k_cols_t<N, R0, R1, R2, R3, NB, MINB, true><<<grid, kColThreads, smem, st>>>(P, tw, specT, binmapT, ws.iacc, ws.binsum, ws.maxpow, power_out, best);
    else
        k_cols_t<N, R0, R1, R2, R3, NB, MINB, false><<<grid, kColThreads, smem, st>>>(P, tw, specT, binmapT, ws.iacc, ws.binsum, ws.maxpow, nullptr, best);
}

}  // namespace

// Compile-time specialised transform lengths: X(N, R0, R1, R2, NB).  (Second and third block: 8K / 5K heights, QHD+, WSXGA+, DCI 4K, 2.7K
// action-camera frames, 5 / 6 / 18 MP sensors and a few more display sizes; 7680 has no three-pass plan with radices <= 25 and stays on the runtime-radix kernels.)  One three-pass radix plan per length serves the
// row kernel (image width N: needs N % 16 == 0 and (N / R0) % 16 == 0) and the column kernel (image height N: needs
// N % 8 == 0; NB = columns per CTA group).  An odd first radix keeps the stride-R0 stores of the first pass conflict
// free.  BASELINE shapes first, then the usual video and camera sizes in both orientations (1080p, 4K, 12 / 16 / 20 /
// 24 MP sensors at 4:3 and 3:2); any other length runs the runtime-radix kernels.
#define PHD_FFT_PLANS(X)                                                                                           \
    X(1920, 15, 8, 16, 2) X(3840, 15, 16, 16, 1) X(6000, 15, 25, 16, 1) X(1080, 9, 10, 12, 4) X(2160, 15, 9, 16, 2) \
    X(4000, 25, 10, 16, 1) X(1280, 5, 16, 16, 4) X(2560, 10, 16, 16, 2) X(1024, 4, 16, 16, 4) X(2048, 8, 16, 16, 2) \
    X(800, 5, 10, 16, 4) X(640, 5, 8, 16, 4) X(720, 9, 10, 8, 4) X(1440, 9, 10, 16, 2) X(768, 3, 16, 16, 4)        \
    X(1536, 6, 16, 16, 2) X(600, 15, 8, 5, 4) X(480, 15, 8, 4, 4) X(960, 15, 8, 8, 4) X(1200, 15, 8, 10, 4)         \
    X(1600, 25, 8, 8, 2) X(2448, 17, 9, 16, 2) X(3000, 15, 8, 25, 1) X(3024, 21, 12, 12, 1) X(3264, 17, 12, 16, 1)  \
    X(3456, 18, 12, 16, 1) X(3648, 19, 12, 16, 1) X(4032, 21, 12, 16, 1) X(4608, 18, 16, 16, 1) X(5472, 19, 18, 16, 1)  \
    X(4320, 15, 16, 18, 1) X(2880, 15, 12, 16, 1) X(1800, 15, 10, 12, 2) X(3200, 25, 8, 16, 1) X(900, 9, 10, 10, 4)       \
    X(1680, 15, 16, 7, 2) X(1050, 15, 10, 7, 4) X(1152, 9, 8, 16, 2) X(2400, 15, 10, 16, 1)                               \
    X(4096, 16, 16, 16, 1) X(5184, 18, 18, 16, 1) X(2000, 25, 5, 16, 2) X(2704, 13, 13, 16, 2) X(1520, 19, 5, 16, 4)     \
    X(2592, 9, 18, 16, 2) X(1944, 9, 12, 18, 2)

// Row widths with a compile-time FOUR-pass plan (no three radices <= 25 multiply to them): 8K and 5K frames.  One row pair
// per CTA in the staged row kernel (two pairs do not fit shared memory), X(N, R0, R1, R2, R3).  As column lengths they run
// the runtime-radix kernel with the same factors.
#define PHD_FFT_PLANS4(X) X(7680, 15, 8, 8, 8) X(5120, 5, 16, 8, 8)

static bool special_radices(int n, int r[4]) {
#define PHD_X(N, R0, R1, R2, NB) if (n == N) { r[0] = R0; r[1] = R1; r[2] = R2; r[3] = 1; return true; }
    PHD_FFT_PLANS(PHD_X)
#undef PHD_X
#define PHD_X(N, R0, R1, R2, R3) if (n == N) { r[0] = R0; r[1] = R1; r[2] = R2; r[3] = R3; return true; }
    PHD_FFT_PLANS4(PHD_X)
#undef PHD_X
    return false;
}

// Radix plan of length n: the compile-time plan when n has one (the same pass tables then serve both kernel
// families), otherwise few, large radices (16 15 12 10 9 8 6 5 4 3 2 have register butterflies; any other prime
// up to PHD_MAX_PRIME gets the O(p^2) pass).  An odd radix goes first: its stride-R stores are conflict free.
int phd_fft_plan_factors(int n, int* fac, int* nfac) {
    int r4[4];
    if (special_radices(n, r4)) {
        int k = 0;
        for (int i = 0; i < 4; i++) if (r4[i] > 1) fac[k++] = r4[i];
        *nfac = k;
        return 0;
    }
    int a2 = 0, a3 = 0, a5 = 0, k = 0, rem = n;
    while (rem % 2 == 0) { a2++; rem /= 2; }
    while (rem % 3 == 0) { a3++; rem /= 3; }
    while (rem % 5 == 0) { a5++; rem /= 5; }
    int odd[PHD_MAX_FACTORS], no = 0, even[PHD_MAX_FACTORS], ne = 0;
    for (int q = 7; rem > 1 && q <= rem; q += 2)
        while (rem % q == 0) {
            if (q > PHD_MAX_PRIME || no >= PHD_MAX_FACTORS) return 1;  // the O(p^2) pass is not meant for huge primes
            odd[no++] = q;
            rem /= q;
        }
    if (rem != 1) return 1;
    auto push = [&](int* arr, int& cnt, int v) { if (cnt < PHD_MAX_FACTORS) arr[cnt++] = v; };
    while (a3 > 0 && a5 > 0) { push(odd, no, 15); a3--; a5--; }
    while (a3 >= 2) { push(odd, no, 9); a3 -= 2; }
    while (a5 >= 2) { push(odd, no, 25); a5 -= 2; }
    if (a5 == 1) { if (a2 >= 1) { push(even, ne, 10); a2--; } else push(odd, no, 5); a5 = 0; }
    if (a3 == 1) {
        if (a2 >= 2) { push(even, ne, 12); a2 -= 2; }
        else if (a2 == 1) { push(even, ne, 6); a2--; }
        else push(odd, no, 3);
        a3 = 0;
    }
    while (a2 >= 4) { push(even, ne, 16); a2 -= 4; }
    if (a2 == 3) push(even, ne, 8);
    else if (a2 == 2) push(even, ne, 4);
    else if (a2 == 1) push(even, ne, 2);
    if (no + ne > PHD_MAX_FACTORS) return 1;
    // largest odd radix first, then everything else, big radices early
    auto sort_desc = [](int* arr, int cnt) {
        for (int i = 1; i < cnt; i++) for (int j = i; j > 0 && arr[j] > arr[j - 1]; j--) { int t = arr[j]; arr[j] = arr[j - 1]; arr[j - 1] = t; }
    };
    sort_desc(odd, no);
    sort_desc(even, ne);
    // a plain prime (O(p^2) pass) should not lead; prefer a register butterfly (15, 25, 9, 5, 3) in front
    int lead = -1;
    for (int i = 0; i < no; i++) if (odd[i] == 15 || odd[i] == 25 || odd[i] == 9 || odd[i] == 5 || odd[i] == 3) { lead = i; break; }
    if (lead >= 0) fac[k++] = odd[lead];
    for (int i = 0; i < ne; i++) fac[k++] = even[i];
    for (int i = 0; i < no; i++) if (i != lead) fac[k++] = odd[i];
    *nfac = k;
    return 0;
}

// Plan of length n.  A length whose largest prime factor exceeds kBlueMinPrime is transformed by Bluestein through the
// smallest 2^a 3^b 5^c >= 2n-1, provided that padded length still fits the shared-memory kernels; everything else (and
// longer sides) keeps the direct radix plan with its O(p^2) pass for odd primes.
int phd_fft_make_plan(int n, FftPlan* pl) {
    constexpr int kBlueMinPrime = 40, kBlueMaxLen = 11000;
    memset(pl, 0, sizeof(*pl));
    pl->n = n;
    int r4[4];
    int largest = 1, rem = n;
    for (int q = 2; q * q <= rem; q++)
        while (rem % q == 0) { largest = q > largest ? q : largest; rem /= q; }
    if (rem > largest) largest = rem;
    if (!special_radices(n, r4) && largest > kBlueMinPrime) {
        int best = 0;
        for (long long a = 1; a <= kBlueMaxLen; a *= 2)
            for (long long b = a; b <= kBlueMaxLen; b *= 3)
                for (long long c = b; c <= kBlueMaxLen; c *= 5)
                    if (c >= 2LL * n - 1 && (best == 0 || c < best)) best = (int)c;
        if (best > 0) {
            pl->m = best;
            return phd_fft_plan_factors(best, pl->fac, &pl->nfac);
        }
    }
    return phd_fft_plan_factors(n, pl->fac, &pl->nfac);
}

void phd_fft_fill_bluestein(float2* chirp_dev, float2* bhat_dev, int n, int m, cudaStream_t st) {
    k_bluestein_tables<<<(m + 127) / 128, 128, 0, st>>>(chirp_dev, bhat_dev, n, m);
}

size_t phd_fft_pass_table_entries(const FftPlan& pl) {
    size_t e = 0;
    const int len = pl.m > 0 ? pl.m : pl.n;
    for (int f = 0; f < pl.nfac; f++) e += (size_t)(pl.fac[f] - 1) * (len / pl.fac[f]);
    return e;
}

void phd_fft_fill_pass_tables(float2* dev, FftPlan& pl, cudaStream_t st) {
    int s = 1, off = 0;
    for (int f = 0; f < pl.nfac; f++) {
        const int r = pl.fac[f];
        const int len = pl.m > 0 ? pl.m : pl.n;
        const int cnt = (r - 1) * (len / r);
        pl.twp_off[f] = off;
        if (cnt > 0) k_pass_twiddles<<<(cnt + 255) / 256, 256, 0, st>>>(dev + off, len, r, s);
        off += cnt;
        s *= r;
    }
    pl.twp = dev;
}

void phd_fill_twiddles(float2* dev_tw, int n, cudaStream_t st) {
    k_twiddles<<<(n + 255) / 256, 256, 0, st>>>(dev_tw, n);
}

// (rows_t_ok / rows_seg: which lengths the row kernel takes, and with which staging granularity -- defined above)
template <int N, int R0, int R1, int R2>
static void launch_rows_t_if(const uint8_t* rgb, const DevParams& P, int nimg, const float2* tw, float2* specT, cudaStream_t st) {
    if constexpr (rows_t_ok<N, R0>()) launch_rows_t<N, R0, R1, R2, 1, 2>(rgb, P, nimg, tw, specT, st);
}
// Lengths the specialised column kernel takes: bulk copies need 16-byte multiples.  A column is N * 8 bytes (always one),
// a column's bin-id slice N * 2 bytes: whole groups of NB slices are a multiple of 16 when NB * N % 8 == 0, and the one
// partial group at the end of an image has its slice copy rounded up to 16 bytes (the map has the slack, pipeline.cu).
template <int N, int NB>
constexpr bool cols_t_ok() { return N % 4 == 0 && (NB * N) % 8 == 0; }
template <int N, int R0, int R1, int R2, int NB>
static void launch_cols_t_if(const DevParams& P, int nimg, const float2* tw, const float2* specT, const u16* binmapT,
                             Workspace& ws, float* power_out, cudaStream_t st) {
    if constexpr (cols_t_ok<N, NB>()) launch_cols_t<N, R0, R1, R2, 1, NB, (N == 1080 ? kColsMinBlocks1080 : 2)>(P, nimg, tw, specT, binmapT, ws, power_out, st);
}

static int rows_generic_grid(size_t smem, int ngroups, int nimg) {
    int per_sm = (int)((220 * 1024) / smem);
    if (per_sm > 8) per_sm = 8;
    if (per_sm < 1) per_sm = 1;
    long long want = (long long)per_sm * 148 * 4;
    int gx = (int)((want + nimg - 1) / nimg);
    if (gx > ngroups) gx = ngroups;
    return gx < 1 ? 1 : gx;
}
// Lengths with a plan that the row kernel cannot take (not a multiple of 8: 900, 1050): staged bytes + compile-time passes.
template <int N, int R0, int R1, int R2>
static void launch_rows_staged_if(const uint8_t* rgb, const DevParams& P, int nimg, const FftPlan& row, float2* specT,
                                  cudaStream_t st) {
    if constexpr (!rows_t_ok<N, R0>()) {
        constexpr int TH = N <= 2048 ? 256 : 512;
        const size_t smem = (size_t)N * 4 * sizeof(float2);
        PHD_ALLOW_SMEM((k_rows_generic<2, N, R0, R1, R2, TH>), 200 * 1024);
        k_rows_generic<2, N, R0, R1, R2, TH><<<dim3(rows_generic_grid(smem, P.Hp / 4, nimg), nimg), TH, smem, st>>>(rgb, P, row, specT);
    }
}

// Lengths the fused front-end + row kernel is built for: the row role must fit the front end's CTA (256 threads, i.e.
// at most 128 sixteen-pixel segments per row, two row pairs) and its shared memory (three CTAs per SM).
template <int N, int R0>
constexpr bool fused_ok() { return rows_seg<N, R0>() == 16 && N / 16 <= 128 && 2 * (2 * N + N / 16) * 8 <= 70 * 1024; }
template <int N, int R0, int R1, int R2>
static void launch_front_rows_if(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                                 const unsigned char* exc, const float2* twp, Workspace& ws, cudaStream_t st) {
    if constexpr (fused_ok<N, R0>()) {
        size_t smem = (size_t)2 * (2 * N + N / 16) * sizeof(float2);
        if (phd_pixels_smem(P) > smem) smem = phd_pixels_smem(P);
        PHD_ALLOW_SMEM((k_front_rows<N, R0, R1, R2>), 100 * 1024);
        // row tasks about as long as a front-end walk (measured alone: a walk of 32 chunks ~ 1/16 of 6.2 us x 444 CTAs,
        // the rows of an image ~ 4.3 us x 444 CTAs): H/4 steps split into parts of ~24 steps
        int parts = (P.H / 4 + 23) / 24;
        if (getenv("PHD_ROW_PARTS")) parts = atoi(getenv("PHD_ROW_PARTS"));
        if (parts < 1) parts = 1;
        const long long tasks = (long long)nimg * (P.nspans + parts);
        const int grid = (int)(tasks < 148 * 3 ? tasks : 148 * 3);
        k_front_rows<N, R0, R1, R2><<<grid, 256, smem, st>>>(rgb, P, tabs, exc, ws.counts_chunk, ws.cells, ws.span32,
                                                             ws.span64, ws.iacc, twp, ws.spec, ws.queue, nimg, parts);
    }
}

static bool rows_fast_ok(const DevParams& P) {
    int r[4];
    return P.H % 4 == 0 && P.Hp == P.H && P.aligned16 && special_radices(P.W, r) && r[3] == 1;
}

// Front end and row FFT of `nimg` images as one launch (see k_front_rows); returns false when this shape / parameter
// set has no fused kernel (the caller then launches the two stages separately).
bool phd_launch_front_rows(const uint8_t* rgb, const DevParams& P, int nimg, const unsigned char* tabs,
                           const unsigned char* exc, const FftPlan& row, Workspace& ws, cudaStream_t st, int* launches) {
    if (P.fe_threads != 256 || P.ds > 1 || !rows_fast_ok(P)) return false;
    switch (P.W) {
#define PHD_X(N, R0, R1, R2, NB) \
    case N: if (fused_ok<N, R0>()) { launch_front_rows_if<N, R0, R1, R2>(rgb, P, nimg, tabs, exc, row.twp, ws, st); *launches += 1; return true; } break;
        PHD_FFT_PLANS(PHD_X)
#undef PHD_X
    }
    return false;
}

int phd_launch_fft_rows(const uint8_t* rgb, const DevParams& P, int nimg, const FftPlan& row, float2* specT,
                        cudaStream_t st, int* launches) {
    *launches += 1;
    if (rows_fast_ok(P)) {
        switch (P.W) {
#define PHD_X(N, R0, R1, R2, NB) \
    case N: if (rows_t_ok<N, R0>()) { launch_rows_t_if<N, R0, R1, R2>(rgb, P, nimg, row.twp, specT, st); return 0; } break;
            PHD_FFT_PLANS(PHD_X)
#undef PHD_X
        }
    }
    switch (P.W) {
#define PHD_X(N, R0, R1, R2, NB) \
    case N: if (!rows_t_ok<N, R0>()) { launch_rows_staged_if<N, R0, R1, R2>(rgb, P, nimg, row, specT, st); return 0; } break;
        PHD_FFT_PLANS(PHD_X)
#undef PHD_X
    }
    switch (P.W) {
#define PHD_X(N, R0, R1, R2, R3)                                                                                        \
    case N: {                                                                                                           \
        const size_t sm = (size_t)N * 2 * sizeof(float2);                                                               \
        PHD_ALLOW_SMEM((k_rows_generic<1, N, R0, R1, R2, 512, R3>), 200 * 1024);                                        \
        k_rows_generic<1, N, R0, R1, R2, 512, R3><<<dim3(rows_generic_grid(sm, P.Hp / 2, nimg), nimg), 512, sm, st>>>(  \
            rgb, P, row, specT);                                                                                        \
        return 0;                                                                                                       \
    }
        PHD_FFT_PLANS4(PHD_X)
#undef PHD_X
    }
    size_t smem = (size_t)(row.m > 0 ? row.m : P.W) * 4 * sizeof(float2);  // two row pairs, two buffers (Bluestein: padded)
    const bool one_pair = smem > 200 * 1024;          // rows longer than 6400 pixels: one pair per CTA
    if (one_pair) smem /= 2;
    if (smem > 200 * 1024) return 1;
    PHD_ALLOW_SMEM((k_rows_generic<1>), 200 * 1024);
    PHD_ALLOW_SMEM((k_rows_generic<2>), 200 * 1024);
    {
        const int gx = rows_generic_grid(smem, P.Hp / (one_pair ? 2 : 4), nimg);
        // one CTA per SM when four rows need more than a third of the shared memory: give it 16 warps
        if (one_pair) k_rows_generic<1><<<dim3(gx, nimg), 512, smem, st>>>(rgb, P, row, specT);
        else k_rows_generic<2><<<dim3(gx, nimg), smem > 72 * 1024 ? 512 : kRowThreads, smem, st>>>(rgb, P, row, specT);
    }
    return 0;
}

// Row transform of ONE image given as a plane of floats (general-input route): the runtime-radix kernel serves every width.
int phd_launch_fft_rows_gray(const float* gray32, const DevParams& P, const FftPlan& row, float2* specT, cudaStream_t st,
                             int* launches) {
    *launches += 1;
    size_t smem = (size_t)(row.m > 0 ? row.m : P.W) * 4 * sizeof(float2);
    const bool one_pair = smem > 200 * 1024;
    if (one_pair) smem /= 2;
    if (smem > 200 * 1024) return 1;
    PHD_ALLOW_SMEM((k_rows_generic<1>), 200 * 1024);
    PHD_ALLOW_SMEM((k_rows_generic<2>), 200 * 1024);
    const int gx = rows_generic_grid(smem, P.Hp / (one_pair ? 2 : 4), 1);
    if (one_pair) k_rows_generic<1><<<dim3(gx, 1), 512, smem, st>>>(nullptr, P, row, specT, gray32);
    else k_rows_generic<2><<<dim3(gx, 1), smem > 72 * 1024 ? 512 : kRowThreads, smem, st>>>(nullptr, P, row, specT, gray32);
    return 0;
}

size_t phd_fft_cols_smem(const DevParams& P, const FftPlan* col, int* tile_cols) {
    const size_t bins = (size_t)2 * P.nbins * sizeof(u32);
    const size_t budget = 110 * 1024;  // two CTAs per SM
    const size_t len = (col && col->m > 0) ? (size_t)col->m : (size_t)P.Hp;  // Bluestein: the padded length
    int tc = 8;
    while (tc > 1 && (size_t)tc * len * 2 * sizeof(float2) + bins > budget) tc >>= 1;
    *tile_cols = tc;
    return (size_t)tc * len * 2 * sizeof(float2) + bins;
}

int phd_launch_fft_cols_blur(const DevParams& P, int nimg, const FftPlan& col, float2* specT, const u16* binmapT,
                             Workspace& ws, float* power_out, cudaStream_t st, int* launches) {
    *launches += 1;
    if (P.Hp == P.H) {
        switch (P.H) {
#define PHD_X(N, R0, R1, R2, NB) \
    case N: if (cols_t_ok<N, NB>()) { launch_cols_t_if<N, R0, R1, R2, NB>(P, nimg, col.twp, specT, binmapT, ws, power_out, st); return 0; } break;
            PHD_FFT_PLANS(PHD_X)
#undef PHD_X
        }
    }
    int tc;
    const size_t smem = phd_fft_cols_smem(P, &col, &tc);
    if (smem > 200 * 1024) return 1;
    PHD_ALLOW_SMEM((k_cols_generic<false>), 200 * 1024);
    PHD_ALLOW_SMEM((k_cols_generic<true>), 200 * 1024);
    dim3 grid((P.fw + tc - 1) / tc, nimg);
    if (power_out)
        k_cols_generic<true><<<grid, kColThreads, smem, st>>>(P, col, tc, specT, binmapT, ws.iacc, ws.binsum, ws.maxpow, power_out);
    else
        k_cols_generic<false><<<grid, kColThreads, smem, st>>>(P, col, tc, specT, binmapT, ws.iacc, ws.binsum, ws.maxpow, nullptr);
    return 0;
}

void phd_launch_bin_map(int W, int H, int Hp, int nr, int na, u16* map_dev, int* counts_dev, cudaStream_t st) {
    const long long n = (long long)(W / 2 + 1) * H;
    cudaMemsetAsync(counts_dev, 0, sizeof(int) * na * nr, st);
    cudaMemsetAsync(map_dev, 0, sizeof(u16) * (size_t)(W / 2 + 1) * Hp, st);
    k_bin_map<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, H, Hp, nr, na, map_dev, counts_dev);
}

// ------------------------------------------------------------------------------------------
// Long transforms through HBM (kernels above): plan, run, and the row / column stages built on them.
// ------------------------------------------------------------------------------------------
namespace {
bool long_smooth(int x) {  // every prime factor has a register butterfly or a cheap O(p^2) pass
    for (int q = 2; q <= 40 && x > 1; q++)
        while (x % q == 0) x /= q;
    return x == 1;
}
bool long_split(int L, int* n1, int* n2) {
    int r = (int)sqrt((double)L);
    for (int d = r; d >= 2; d--) {
        if (L % d) continue;
        const int e = L / d;
        if (e > 4096) break;
        if (long_smooth(d) && long_smooth(e)) { *n1 = d; *n2 = e; return true; }
    }
    return false;
}
int long_tile(int len) {  // sub-sequences per CTA: two buffers of tile * len complex values within 96 KB
    int tb = 6144 / len;
    return tb < 1 ? 1 : (tb > 16 ? 16 : tb);
}
cudaError_t long_fft_raw(const PhdLongFft& lf, float2* data, float2* tmp, int nseq, size_t stride, cudaStream_t st,
                         int* launches) {
    const int ta = long_tile(lf.n1), tb = long_tile(lf.n2);
    PHD_ALLOW_SMEM((k_long_stage_a), 100 * 1024);
    PHD_ALLOW_SMEM((k_long_stage_b), 100 * 1024);
    k_long_stage_a<<<dim3((lf.n2 + ta - 1) / ta, nseq), 256, (size_t)2 * ta * lf.n1 * sizeof(float2), st>>>(
        data, tmp, stride, lf.p1, lf.n1, lf.n2, lf.L, lf.twL, ta);
    k_long_stage_b<<<dim3((lf.n1 + tb - 1) / tb, nseq), 256, (size_t)2 * tb * lf.n2 * sizeof(float2), st>>>(
        tmp, data, stride, lf.p2, lf.n1, lf.n2, tb);
    *launches += 2;
    return cudaGetLastError();
}
// transform of logical length lf.n of nseq sequences at `stride` (>= lf.L) complex values; result in data[.][0 .. n)
cudaError_t long_fft_run(const PhdLongFft& lf, float2* data, float2* tmp, int nseq, size_t stride, cudaStream_t st,
                         int* launches) {
    if (!lf.blue) return long_fft_raw(lf, data, tmp, nseq, stride, st, launches);
    const dim3 gl((lf.L + 255) / 256, nseq), gn((lf.n + 255) / 256, nseq);
    k_long_chirp_in<<<gl, 256, 0, st>>>(data, stride, lf.n, lf.L, lf.chirp);
    cudaError_t e = long_fft_raw(lf, data, tmp, nseq, stride, st, launches);
    if (e != cudaSuccess) return e;
    k_long_mul_bhat<<<gl, 256, 0, st>>>(data, stride, lf.L, lf.bhat);
    e = long_fft_raw(lf, data, tmp, nseq, stride, st, launches);
    if (e != cudaSuccess) return e;
    k_long_chirp_out<<<gn, 256, 0, st>>>(data, stride, lf.n, lf.chirp);
    *launches += 3;
    return cudaGetLastError();
}
size_t sub_plan_entries(const FftPlan& pl) { return (size_t)pl.n + phd_fft_pass_table_entries(pl); }
}  // namespace

int phd_long_fft_create(int n, PhdLongFft* lf, cudaStream_t st) {
    memset(lf, 0, sizeof(*lf));
    lf->n = n;
    lf->L = n;
    if (!long_split(n, &lf->n1, &lf->n2)) {
        lf->blue = 1;
        long long best = 0;
        for (long long a = 1; a < (1LL << 24); a *= 2)
            for (long long b = a; b < (1LL << 24); b *= 3)
                for (long long c = b; c < (1LL << 24); c *= 5)
                    if (c >= 2LL * n - 1 && (best == 0 || c < best)) best = c;
        lf->L = (int)best;
        if (!long_split(lf->L, &lf->n1, &lf->n2)) return 1;
    }
    lf->p1.n = lf->n1;
    lf->p2.n = lf->n2;
    if (phd_fft_plan_factors(lf->n1, lf->p1.fac, &lf->p1.nfac) || phd_fft_plan_factors(lf->n2, lf->p2.fac, &lf->p2.nfac)) return 1;
    const size_t e1 = sub_plan_entries(lf->p1), e2 = sub_plan_entries(lf->p2);
    const size_t total = e1 + e2 + (size_t)lf->L + (lf->blue ? (size_t)n + lf->L : 0);
    if (cudaMalloc(&lf->mem, sizeof(float2) * total) != cudaSuccess) return 2;
    float2* t1 = lf->mem;
    float2* t2 = t1 + e1;
    float2* twL = t2 + e2;
    float2* chirp = twL + lf->L;
    float2* bhat = chirp + n;
    phd_fill_twiddles(t1, lf->n1, st);
    phd_fft_fill_pass_tables(t1 + lf->n1, lf->p1, st);
    lf->p1.tw = t1;
    phd_fill_twiddles(t2, lf->n2, st);
    phd_fft_fill_pass_tables(t2 + lf->n2, lf->p2, st);
    lf->p2.tw = t2;
    lf->twL = twL;
    float2* tmp = nullptr;
    if (lf->blue && cudaMalloc(&tmp, sizeof(float2) * lf->L) != cudaSuccess) return 2;
    k_long_tables<<<(lf->L + 255) / 256, 256, 0, st>>>(twL, lf->blue ? chirp : nullptr, lf->blue ? bhat : nullptr, n, lf->L, lf->blue);
    if (lf->blue) {
        // bhat = (1/L) * DFT_L of the circular conjugate chirp, by the long transform itself
        lf->chirp = chirp;
        lf->bhat = bhat;
        int launches = 0;
        const cudaError_t e = long_fft_raw(*lf, bhat, tmp, 1, (size_t)lf->L, st, &launches);
        k_long_scale<<<(lf->L + 255) / 256, 256, 0, st>>>(bhat, lf->L, 1.0f / (float)lf->L);
        cudaStreamSynchronize(st);
        cudaFree(tmp);
        if (e != cudaSuccess) return 2;
    }
    return cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess ? 0 : 2;
}

void phd_long_fft_destroy(PhdLongFft* lf) {
    if (lf && lf->mem) cudaFree(lf->mem);
    if (lf) memset(lf, 0, sizeof(*lf));
}

int phd_launch_long_rows(const uint8_t* rgb, const float* gray32, const DevParams& P, const PhdLongFft& lf, float2* buf0,
                         float2* buf1, float2* specT, cudaStream_t st, int* launches) {
    const int npairs = P.Hp / 2;
    const size_t stride = (size_t)lf.L;
    k_long_gray<<<dim3(32, npairs), 256, 0, st>>>(rgb, gray32, P, buf0, stride);
    if (long_fft_run(lf, buf0, buf1, npairs, stride, st, launches) != cudaSuccess) return 1;
    k_long_rows_out<<<dim3(16, npairs), 256, 0, st>>>(buf0, stride, P, specT);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int phd_launch_long_cols(const DevParams& P, const PhdLongFft& lf, const float2* specT, float2* buf0, float2* buf1,
                         const u16* binmapT, Workspace& ws, float* power_out, cudaStream_t st, int* launches) {
    const size_t stride = (size_t)lf.L;
    k_long_cols_in<<<dim3(16, P.fw), 256, 0, st>>>(specT, P, buf0, stride);
    if (long_fft_run(lf, buf0, buf1, P.fw, stride, st, launches) != cudaSuccess) return 1;
    const size_t smem = (size_t)2 * P.nbins * sizeof(u32);
    PHD_ALLOW_SMEM((k_long_cols_epi<false>), 100 * 1024);
    if (power_out) k_long_cols_epi<true><<<P.fw, kColThreads, 0, st>>>(P, buf0, stride, binmapT, ws.iacc, ws.binsum, ws.maxpow, power_out);
    else k_long_cols_epi<false><<<P.fw, kColThreads, smem, st>>>(P, buf0, stride, binmapT, ws.iacc, ws.binsum, ws.maxpow, nullptr);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
