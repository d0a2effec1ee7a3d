// Micro-benchmark: issue rate of the packed FP32x2 instructions (FADD2 / FFMA2, PTX add/fma.rn.f32x2) on sm_100a
// against their scalar forms, alone and interleaved with shared-memory loads -- decides whether the FFT butterflies
// (issue bound, about half of their instructions complex adds and multiplies) should use them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/f32x2 tools/ubench/f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8  // independent chains per thread

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm volatile("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7};"
                 "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
                 : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm volatile("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb;"
                 "mov.b64 {%0,%1}, rc; }" : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}

// MODE 0: scalar FFMA, 2*CH per iteration   1: FFMA2, CH per iteration (same flops)
// MODE 2: scalar FADD                      3: FADD2
// MODE 4: FFMA2 + one LDS.64 per FFMA2      5: scalar FFMA (2 per) + one LDS.64
template <int MODE>
__global__ void __launch_bounds__(256) k(float2* out, const float2* in, long long* clk) {
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = in[i];
    __syncthreads();
    float2 v[CH];
#pragma unroll
    for (int c = 0; c < CH; c++) v[c] = in[threadIdx.x + c * 32];
    const float2 m = in[5], a = in[6];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CH; c++) {
            if (MODE == 0 || MODE == 5) { v[c].x = fmaf(v[c].x, m.x, a.x); v[c].y = fmaf(v[c].y, m.y, a.y); }
            if (MODE == 1 || MODE == 4) v[c] = fma2(v[c], m, a);
            if (MODE == 2) { v[c].x += a.x; v[c].y += a.y; }
            if (MODE == 3) v[c] = add2(v[c], a);
            if (MODE == 4 || MODE == 5) {
                const float2 l = sm[(threadIdx.x + it + c * 32) & 1023];
                v[c].x += l.x;
            }
        }
    }
    const long long t1 = clock64();
    float2 s = v[0];
#pragma unroll
    for (int c = 1; c < CH; c++) { s.x += v[c].x; s.y += v[c].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_chain) {
    const int ctas = 148 * 4;
    float2 *out, *in;
    long long* clk;
    cudaMalloc(&out, ctas * 256 * sizeof(float2));
    cudaMalloc(&in, 1024 * sizeof(float2));
    cudaMemset(in, 0, 1024 * sizeof(float2));
    cudaMalloc(&clk, ctas * sizeof(long long));
    k<MODE><<<ctas, 256>>>(out, in, clk);
    k<MODE><<<ctas, 256>>>(out, in, clk);
    cudaDeviceSynchronize();
    long long h[148 * 4];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; i++) avg += (double)h[i] / ctas;
    // 4 CTAs x 8 warps per SM; warp instructions of the listed kinds per clock64 tick per SM.  clock64 ticks slower
    // than the boosted SM clock on this part (scalar FFMA shows 5.26 per tick = the 4 per SM clock peak): compare rows.
    const double winstr = 4.0 * 8 * ITERS * CH * instr_per_chain;
    printf("%-44s %8.0f ticks  %5.2f warp-instr/tick/SM\n", name, avg, winstr / avg);
    cudaFree(out); cudaFree(in); cudaFree(clk);
}

int main() {
    run<0>("scalar FFMA (2 per complex)", 2);
    run<1>("FFMA2 (1 per complex)", 1);
    run<2>("scalar FADD (2 per complex)", 2);
    run<3>("FADD2 (1 per complex)", 1);
    run<4>("FFMA2 + LDS.64 + FADD (3 instr)", 3);
    run<5>("2 FFMA + LDS.64 + FADD (4 instr)", 4);
    return 0;
}
