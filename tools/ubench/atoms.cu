// Micro-benchmarks that decide the histogram design (shared-memory atomics vs alternatives) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/atoms tools/ubench/atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;

#define ITERS 4096
#define TBL 1024

__device__ __forceinline__ u32 lcg(u32& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

template <int MODE>
__global__ void __launch_bounds__(256) k(u32* out, u64* gtab, int gsize) {
    __shared__ u32 tab[TBL * 2];
    __shared__ u32 priv[32 * 64 * 8 / 8];  // lane-private [entry][lane], 64 entries per warp... (8 warps)
    for (int i = threadIdx.x; i < TBL * 2; i += blockDim.x) tab[i] = 0;
    __syncthreads();
    u32 s = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    u32 acc = 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
        const u32 r = lcg(s);
        if (MODE == 0) atomicAdd(&tab[r & (TBL - 1)], 1u);                  // spread, no return
        if (MODE == 1) acc += atomicAdd(&tab[r & (TBL - 1)], 1u);           // spread, with return
        if (MODE == 2) atomicAdd(&tab[(r & 31) * 32], 1u);                  // all lanes same bank, 32 addresses
        if (MODE == 3) atomicAdd(&tab[it & (TBL - 1)], r);                  // warp-uniform address
        if (MODE == 4) atomicAdd(reinterpret_cast<u64*>(tab) + (r & (TBL - 1)), (u64)r);  // 64-bit shared
        if (MODE == 5) { u32* p = &priv[((r & 7) * 8 + warp) * 32 + lane]; *p += r; }     // lane-private RMW
        if (MODE == 6) atomicAdd(&gtab[(r * 2654435761u >> 8) % (u32)gsize], (u64)r);      // global red spread
        if (MODE == 7) acc += __match_any_sync(0xffffffffu, r & 1023);
        if (MODE == 8) acc += __match_any_sync(0xffffffffu, r & 3);
        if (MODE == 9) { atomicAdd(&tab[r & (TBL - 1)], 1u); atomicAdd(&tab[TBL + (r & (TBL - 1))], r & 255); }  // two tables same index
        if (MODE == 10) { acc += r; }  // baseline loop
        if (MODE == 11) atomicAdd(&tab[(r & 15) * 2], 1u);                  // 16 hot addresses (heavy same-address conflicts)
        if (MODE == 13) atomicAdd(&tab[(gtab[it & 15] + it) & (TBL - 1)], r);                   // warp-uniform address the compiler cannot prove uniform
        if (MODE == 14) atomicAdd(&tab[((gtab[it & 15] + it) & (TBL - 1)) ^ (lane & 1)], r);     // two addresses per warp
        if (MODE == 15) { const u32 c = ((gtab[it & 15] + it) & 255) + (lane >> 3); atomicAdd(&tab[c], 1u); atomicAdd(&tab[c + 576], r & 255); atomicAdd(&tab[c + 1152], r); }  // 4 distinct cells per warp, 3 words
        if (MODE == 16) { const u32 c = r & 511; atomicAdd(&tab[c], 1u); atomicAdd(&tab[c + 512], r & 255); atomicAdd(&tab[c + 1024], r); atomicAdd(&tab[c + 1536], r >> 3); }  // random cells, 4 words
        if (MODE == 12) { const u32 o = atomicAdd(&tab[r & (TBL - 1)], r << 12); if (o + (r << 12) < o) atomicAdd(&tab[TBL + (r & (TBL - 1))], 1u); }  // 64-bit by carry
    }
    __syncthreads();
    u32 t = acc;
    for (int i = threadIdx.x; i < TBL * 2; i += blockDim.x) t += tab[i];
    if (MODE == 5) for (int i = threadIdx.x; i < 2048; i += blockDim.x) t += priv[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
void run(const char* name, u32* out, u64* gtab, int gsize) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int ctas = sms * 4;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<ctas, 256>>>(out, gtab, gsize);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<ctas, 256>>>(out, gtab, gsize);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double laneops = (double)ctas * 256 * ITERS;
    const double cyc = ms * 1e-3 * clk * 1e3;
    printf("%-44s %8.3f ms  %7.3f lane-ops/clk/SM (at %d MHz nominal)  %s\n", name, ms, laneops / cyc / sms, clk / 1000,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    u32* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int gsize = 1 << 20;
    u64* gtab; cudaMalloc(&gtab, gsize * 8); cudaMemset(gtab, 0, gsize * 8);
    run<10>("baseline loop (lcg only)", out, gtab, gsize);
    run<0>("ATOMS.ADD u32 spread 1024, no return", out, gtab, gsize);
    run<1>("ATOMS.ADD u32 spread 1024, with return", out, gtab, gsize);
    run<2>("ATOMS.ADD u32 same bank, 32 addresses", out, gtab, gsize);
    run<3>("ATOMS.ADD u32 warp-uniform address", out, gtab, gsize);
    run<11>("ATOMS.ADD u32 16 hot addresses", out, gtab, gsize);
    run<4>("atomicAdd u64 shared spread", out, gtab, gsize);
    run<12>("u32 add + carry (64-bit emulation)", out, gtab, gsize);
    run<9>("2x ATOMS.ADD spread", out, gtab, gsize);
    cudaMemset(gtab, 0, gsize * 8);
    run<13>("ATOMS.ADD runtime warp-uniform address", out, gtab, gsize);
    run<14>("ATOMS.ADD two addresses per warp", out, gtab, gsize);
    run<15>("3x ATOMS, 4 cells per warp (smooth image)", out, gtab, gsize);
    run<16>("4x ATOMS, random cells of 512 (noise image)", out, gtab, gsize);
    run<5>("lane-private LDS/add/STS", out, gtab, gsize);
    run<6>("RED.ADD u64 global spread 1M", out, gtab, gsize);
    run<7>("match.any 1024 keys", out, gtab, gsize);
    run<8>("match.any 4 keys", out, gtab, gsize);
    return 0;
}
