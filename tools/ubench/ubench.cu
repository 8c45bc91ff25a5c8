// Throw-away instruction-rate probes used to choose the direction-bit extraction scheme.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CHAINS 8
#define UNROLL 8

template <int KIND> __global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t seed) {
    __shared__ uint32_t tab[24 * 24 * 32 / 4 + 32];  // small; conflict-free indexing by lane
    uint32_t a[CHAINS], b[CHAINS], acc = 0;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 24 * 24 * 8 + 32; i += blockDim.x) tab[i] = i * seed;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = seed * (c + 1) + t; b[c] = seed ^ (t * 2654435761u + c); }
    const uint32_t c1 = 0xfffffff7u + (seed & 1);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
                if (KIND == 0) {  // VIMNMX with predicates + 2 predicated adds
                    bool p, q;
                    a[c] = __vibmax_s16x2(a[c], b[c], &p, &q);
                    b[c] = __vadd2(b[c], c1);
                    if (p) acc += (1u << (2 * c));
                    if (q) acc += (2u << (2 * c));
                } else if (KIND == 1) {  // VIMNMX with predicates + 2 ballots
                    bool p, q;
                    a[c] = __vibmax_s16x2(a[c], b[c], &p, &q);
                    b[c] = __vadd2(b[c], c1);
                    acc ^= __ballot_sync(0xffffffffu, p);
                    acc += __ballot_sync(0xffffffffu, q);
                } else if (KIND == 2) {  // conflict-free LDS + address add + pack
                    uint32_t i0 = (a[c] & 0x1f80u) + lane * 4, i1 = (b[c] & 0x1f80u) + lane * 4;
                    uint32_t lo = *(const uint32_t *)((const char *)tab + i0);
                    uint32_t hi = *(const uint32_t *)((const char *)tab + i1);
                    a[c] = hi * 65536u + lo + a[c];
                    b[c] = __viaddmax_s16x2(b[c], c1, a[c]);
                } else if (KIND == 3) {  // PRMT + VIADDMNMX (both ALU?)
                    a[c] = __byte_perm(a[c], b[c], 0x5410);
                    b[c] = __viaddmax_s16x2(b[c], c1, a[c]);
                } else if (KIND == 4) {  // LOP3 + VIADDMNMX
                    a[c] = (a[c] & 0xfffcfffcu) | (b[c] & 0x00030003u);
                    b[c] = __viaddmax_s16x2(b[c], c1, a[c]);
                } else if (KIND == 5) {  // ballot only
                    acc += __ballot_sync(0xffffffffu, (a[c] + u) & 1);
                    a[c] = __viaddmax_s16x2(a[c], c1, b[c]);
                } else if (KIND == 6) {  // tag approach cell: 8 ALU + imads
                    uint32_t h = a[c] & 0xfffcfffcu;
                    uint32_t m = __vadd2(h, c1);
                    uint32_t f1 = __viaddmax_s16x2(b[c], c1, m);
                    uint32_t f1n = (f1 & 0xfffcfffcu) | 0x00010001u;
                    uint32_t hh = __viaddmax_s16x2(__vmaxs2(f1n, b[c]), c1, m);
                    acc = acc * 16u + (a[c] - h) + 4u * (f1 - f1n);
                    a[c] = hh; b[c] = f1n;
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= a[c] + b[c];
    out[t] = acc;
}

template <int KIND> void run(const char *name, double instr_per_inner) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 2048;
    uint32_t *out; cudaMalloc(&out, threads * blocks * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<KIND><<<blocks, threads>>>(out, iters, 12345u); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    double inner = (double)threads * blocks * iters * UNROLL * CHAINS;
    double clk = 1.965e9 * prop.multiProcessorCount * 4;  // SMSP-cycles per second
    printf("%-34s %8.3f ms  inner/s=%.3e  SMSP-cycles per warp-inner=%.2f  (%s)\n", name, best,
           inner / (best * 1e-3), clk * (best * 1e-3) / (inner / 32), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    run<0>("vibmax+vadd+2 pred-add", 4);
    run<1>("vibmax+vadd+2 ballot", 4);
    run<2>("2 LDS(cf)+2 addr+pack+viaddmax", 6);
    run<3>("PRMT+VIADDMNMX", 2);
    run<4>("LOP3+VIADDMNMX", 2);
    run<5>("ballot+VIADDMNMX", 2);
    run<6>("tag cell", 10);
    return 0;
}
