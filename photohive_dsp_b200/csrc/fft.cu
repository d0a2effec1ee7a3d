// Hand-written 2-D real FFT of the grayscale image with the blur-profile binning fused into the
// column pass.  No cuFFT anywhere in the library.
//
//   k_fft_rows       first half of pgm_fft (src/fft_processing.c:18-63) with rgb2pgm fused in front
//                    (src/image_processing.c:505-512): two image rows are read as packed RGB, turned
//                    into the exact integer gray numerators 299R+587G+114B, packed as one complex
//                    sequence, transformed in shared memory (Stockham autosort, mixed radix) and
//                    split into the two half spectra.  A constant 127500 (= 0.5 gray) is removed
//                    before the transform for FP32 headroom; DC is repaired exactly in the epilogue.
//   k_fft_cols_blur  second half of pgm_fft, remove_dc_bias (src/blur_profile.c:233-238),
//                    pgm_normalize_fft (src/fft_processing.c:173-200) and the accumulation loop of
//                    calculate_blur_profile (src/blur_profile.c:87-100): column tiles are transformed
//                    in shared memory, then power -> (p<1 ? 0 : ln p) -> polar bin (cached id map) ->
//                    shared-memory integer bins -> 64-bit global integer bins; max power by atomicMax.
//                    G_s is a scalar and is applied after averaging, in finalize.
//   k_bin_map        cartesian_to_polar_conversion + the bin index arithmetic
//                    (src/blur_profile.c:427-458, :94-97; newton_int_sqrt src/utilities.c:43-52),
//                    image independent, built once per (W,H,nr,na) and cached.
#include <math.h>

#include "phd_internal.h"

namespace {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void butterfly(float2 (&x)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&x)[2]) {
    float2 a = x[0], b = x[1];
    x[0] = caddf(a, b);
    x[1] = csubf(a, b);
}
template <>
__device__ __forceinline__ void butterfly<4>(float2 (&x)[4]) {
    float2 t0 = caddf(x[0], x[2]), t1 = csubf(x[0], x[2]);
    float2 t2 = caddf(x[1], x[3]), t3 = mul_mi(csubf(x[1], x[3]));
    x[0] = caddf(t0, t2);
    x[1] = caddf(t1, t3);
    x[2] = csubf(t0, t2);
    x[3] = csubf(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<3>(float2 (&x)[3]) {
    const float k3 = 0.86602540378443864676f;
    float2 t = caddf(x[1], x[2]);
    float2 u = make_float2(x[0].x - 0.5f * t.x, x[0].y - 0.5f * t.y);
    float2 d = csubf(x[1], x[2]);
    float2 v = make_float2(k3 * d.y, -k3 * d.x);
    x[0] = caddf(x[0], t);
    x[1] = caddf(u, v);
    x[2] = csubf(u, v);
}
template <>
__device__ __forceinline__ void butterfly<5>(float2 (&x)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    float2 t1 = caddf(x[1], x[4]), t2 = caddf(x[2], x[3]);
    float2 t3 = csubf(x[1], x[4]), t4 = csubf(x[2], x[3]);
    float2 m1 = make_float2(x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y);
    float2 m2 = make_float2(x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y);
    float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
    float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
    x[0] = make_float2(x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y);
    x[1] = make_float2(m1.x + n1.y, m1.y - n1.x);
    x[4] = make_float2(m1.x - n1.y, m1.y + n1.x);
    x[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
    x[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
}

// One Stockham pass of radix R over `nbatch` sequences of length n laid out `bstride` apart.
// in/out are shared-memory buffers; s is the product of the radices already applied.
template <int R>
__device__ __forceinline__ void fft_pass(const float2* __restrict__ in, float2* __restrict__ out, int n, int s,
                                         const float2* __restrict__ tw, int nbatch, int bstride) {
    const int m = n / R;
    const int total = nbatch * m;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int col = idx / m;
        const int b = idx - col * m;
        const float2* a = in + col * bstride;
        float2* y = out + col * bstride;
        const int q = b % s;
        const int pps = b - q;
        float2 x[R];
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = a[b + k * m];
        butterfly<R>(x);
        y[R * pps + q] = x[0];
#pragma unroll
        for (int j = 1; j < R; j++) y[R * pps + q + j * s] = cmulf(x[j], __ldg(&tw[pps * j]));
    }
}

// Generic odd prime radix (7..31): O(r^2) butterfly with table twiddles.
__device__ void fft_pass_generic(int r, const float2* __restrict__ in, float2* __restrict__ out, int n, int s,
                                 const float2* __restrict__ tw, int nbatch, int bstride) {
    const int m = n / r;
    const int total = nbatch * m;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int col = idx / m;
        const int b = idx - col * m;
        const float2* a = in + col * bstride;
        float2* y = out + col * bstride;
        const int q = b % s;
        const int pps = b - q;
        for (int j = 0; j < r; j++) {
            float2 acc = a[b];
            for (int k = 1; k < r; k++) acc = caddf(acc, cmulf(a[b + k * m], __ldg(&tw[((j * k) % r) * m])));
            y[r * pps + q + j * s] = j ? cmulf(acc, __ldg(&tw[pps * j])) : acc;
        }
    }
}

// Runs every pass; returns the buffer that holds the result.
__device__ float2* fft_run(const FftPlan& pl, float2* bufA, float2* bufB, int nbatch, int bstride) {
    float2* a = bufA;
    float2* b = bufB;
    int s = 1;
    for (int f = 0; f < pl.nfac; f++) {
        const int r = pl.fac[f];
        switch (r) {
            case 2: fft_pass<2>(a, b, pl.n, s, pl.tw, nbatch, bstride); break;
            case 3: fft_pass<3>(a, b, pl.n, s, pl.tw, nbatch, bstride); break;
            case 4: fft_pass<4>(a, b, pl.n, s, pl.tw, nbatch, bstride); break;
            case 5: fft_pass<5>(a, b, pl.n, s, pl.tw, nbatch, bstride); break;
            default: fft_pass_generic(r, a, b, pl.n, s, pl.tw, nbatch, bstride); break;
        }
        __syncthreads();
        s *= r;
        float2* t = a; a = b; b = t;
    }
    return a;
}

__global__ void k_twiddles(float2* tw, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)n, &s, &c);
    tw[k] = make_float2((float)c, (float)(-s));
}

// ------------------------------------------------------------------------------------------
// Row pass: one CTA per row pair.
__global__ void __launch_bounds__(256) k_fft_rows(const uint8_t* __restrict__ rgb, DevParams P, FftPlan pl,
                                                  float2* __restrict__ spec) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + P.W;
    const int img = blockIdx.y, pr = blockIdx.x;
    const int ra = 2 * pr, rb = 2 * pr + 1;
    const bool has_b = rb < P.H;
    const uint8_t* base = rgb + (size_t)img * P.image_stride;
    const uint8_t* pa = base + (size_t)ra * P.W * 3;
    const uint8_t* pb = base + (size_t)rb * P.W * 3;
    for (int x = threadIdx.x; x < P.W; x += blockDim.x) {
        const int ga = 299 * (int)__ldg(pa + 3 * x) + 587 * (int)__ldg(pa + 3 * x + 1) + 114 * (int)__ldg(pa + 3 * x + 2);
        int gb = 127500;
        if (has_b) gb = 299 * (int)__ldg(pb + 3 * x) + 587 * (int)__ldg(pb + 3 * x + 1) + 114 * (int)__ldg(pb + 3 * x + 2);
        bufA[x] = make_float2((float)(ga - 127500), (float)(gb - 127500));
    }
    __syncthreads();
    const float2* z = fft_run(pl, bufA, bufB, 1, 0);
    float2* oa = spec + ((size_t)img * P.H + ra) * P.fw;
    float2* ob = spec + ((size_t)img * P.H + rb) * P.fw;
    for (int k = threadIdx.x; k < P.fw; k += blockDim.x) {
        const int kc = k == 0 ? 0 : P.W - k;
        const float2 zk = z[k];
        const float2 zc = make_float2(z[kc].x, -z[kc].y);
        oa[k] = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
        if (has_b) ob[k] = make_float2(0.5f * (zk.y - zc.y), -0.5f * (zk.x - zc.x));
    }
}

// ------------------------------------------------------------------------------------------
// Column pass + blur binning: one CTA per tile of TC adjacent spectrum columns.
template <bool WRITE_POWER>
__global__ void __launch_bounds__(256) k_fft_cols_blur(DevParams P, FftPlan pl, int TC,
                                                       const float2* __restrict__ spec,
                                                       const u16* __restrict__ binmap,
                                                       const ImageAcc* __restrict__ iacc, u64* __restrict__ binsum,
                                                       u32* __restrict__ maxpow, float* __restrict__ power_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int H = P.H;
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)TC * H;
    u32* bin_lo = reinterpret_cast<u32*>(bufB + (size_t)TC * H);
    u32* bin_hi = bin_lo + P.nbins;
    __shared__ float sh_max[8];

    const int img = blockIdx.y;
    const int x0 = blockIdx.x * TC;
    const int ncol = min(TC, P.fw - x0);
    const float2* src = spec + (size_t)img * H * P.fw;
    if (!WRITE_POWER)
        for (int b = threadIdx.x; b < 2 * P.nbins; b += blockDim.x) bin_lo[b] = 0;
    for (int idx = threadIdx.x; idx < H * TC; idx += blockDim.x) {
        const int k = idx / TC, c = idx - k * TC;
        bufA[(size_t)c * H + k] = c < ncol ? src[(size_t)k * P.fw + x0 + c] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    const float2* res = fft_run(pl, bufA, bufB, ncol, H);

    const float inv_scale2 = (float)(1.0 / (255000.0 * 255000.0));
    float mymax = 0.f;
    for (int idx = threadIdx.x; idx < H * ncol; idx += blockDim.x) {
        const int k = idx / ncol, c = idx - k * ncol;
        const int x = x0 + c;
        const float2 v = res[(size_t)c * H + k];
        float p = (v.x * v.x + v.y * v.y) * inv_scale2;
        if (WRITE_POWER) {
            power_out[((size_t)img * H + k) * P.fw + x] = p;
        } else {
            if (x == 0 && k == 0) {
                // DC: sum(gray - avg) from the exact channel sums (interface.c:78, blur_profile.c:233-238)
                const ImageAcc a = iacc[img];
                const double np = (double)P.npx;
                const double avg = ((double)a.sum[0] / 255.0 / np + (double)a.sum[1] / 255.0 / np +
                                    (double)a.sum[2] / 255.0 / np) / 3.0;
                const double gsum = (299.0 * (double)a.sum[0] + 587.0 * (double)a.sum[1] + 114.0 * (double)a.sum[2]) / 255000.0;
                const double dc = gsum - np * avg;
                p = (float)(dc * dc);
            }
            mymax = fmaxf(mymax, p);
            if (p >= 1.0f) {
                const u32 q = (u32)__float2int_rn(logf(p) * (float)(1 << PHD_LN_SHIFT));
                const int bin = binmap[(size_t)k * P.fw + x];
                atomicAdd(&bin_lo[bin], q & 0x1fffu);
                atomicAdd(&bin_hi[bin], q >> 13);
            }
        }
    }
    if (WRITE_POWER) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
    if ((threadIdx.x & 31) == 0) sh_max[threadIdx.x >> 5] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, sh_max[w]);
        atomicMax(&maxpow[img], __float_as_uint(m));
    }
    u64* dst = binsum + (size_t)img * P.nbins;
    for (int b = threadIdx.x; b < P.nbins; b += blockDim.x) {
        const u64 v = ((u64)bin_hi[b] << 13) + bin_lo[b];
        if (v) atomicAdd(&dst[b], v);
    }
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

__global__ void k_bin_map(int W, int H, int nr, int na, u16* __restrict__ map, int* __restrict__ counts) {
    const int fw = W / 2 + 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)fw * H) return;
    const int k = (int)(i / fw), x = (int)(i - (long long)k * fw);
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
    map[i] = (u16)bin;
    atomicAdd(&counts[bin], 1);
}

}  // namespace

int phd_fft_plan_factors(int n, int* fac, int* nfac) {
    int rem = n, k = 0;
    while (rem % 4 == 0) { fac[k++] = 4; rem /= 4; }
    while (rem % 2 == 0) { fac[k++] = 2; rem /= 2; }
    for (int q = 3; rem > 1; q += 2) {
        while (rem % q == 0) {
            if (q > 31 || k >= PHD_MAX_FACTORS) return 1;  // large prime factor: not covered yet
            fac[k++] = q;
            rem /= q;
        }
    }
    *nfac = k;
    return 0;
}

void phd_fill_twiddles(float2* dev_tw, int n, cudaStream_t st) {
    k_twiddles<<<(n + 255) / 256, 256, 0, st>>>(dev_tw, n);
}

int phd_launch_fft_rows(const uint8_t* rgb, const DevParams& P, int nimg, const FftPlan& row, float2* spec,
                        cudaStream_t st, int* launches) {
    const size_t smem = (size_t)P.W * 2 * sizeof(float2);
    if (smem > 200 * 1024) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_fft_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid((P.H + 1) / 2, nimg);
    k_fft_rows<<<grid, 256, smem, st>>>(rgb, P, row, spec);
    *launches += 1;
    return 0;
}

size_t phd_fft_cols_smem(const DevParams& P, int* tile_cols) {
    const size_t bins = (size_t)2 * P.nbins * sizeof(u32);
    const size_t budget = 200 * 1024;
    int tc = 8;
    while (tc > 1 && (size_t)tc * P.H * 2 * sizeof(float2) + bins > budget) tc >>= 1;
    *tile_cols = tc;
    return (size_t)tc * P.H * 2 * sizeof(float2) + bins;
}

int phd_launch_fft_cols_blur(const DevParams& P, int nimg, const FftPlan& col, float2* spec, const u16* binmap,
                             Workspace& ws, float* power_out, cudaStream_t st, int* launches) {
    int tc;
    const size_t smem = phd_fft_cols_smem(P, &tc);
    if (smem > 200 * 1024) return 1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_fft_cols_blur<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(k_fft_cols_blur<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    dim3 grid((P.fw + tc - 1) / tc, nimg);
    if (power_out)
        k_fft_cols_blur<true><<<grid, 256, smem, st>>>(P, col, tc, spec, binmap, ws.iacc, ws.binsum, ws.maxpow, power_out);
    else
        k_fft_cols_blur<false><<<grid, 256, smem, st>>>(P, col, tc, spec, binmap, ws.iacc, ws.binsum, ws.maxpow, nullptr);
    *launches += 1;
    return 0;
}

void phd_launch_bin_map(int W, int H, int nr, int na, u16* map_dev, int* counts_dev, cudaStream_t st) {
    const long long n = (long long)(W / 2 + 1) * H;
    cudaMemsetAsync(counts_dev, 0, sizeof(int) * na * nr, st);
    k_bin_map<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, H, nr, na, map_dev, counts_dev);
}
