// Instruction issue-rate probes (cycles per warp-instruction per SM sub-partition) for the
// instructions of the packed fill kernels.  Build: nvcc -arch=sm_100a -O3 -o rates rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define CHAINS 8
#define UNROLL 8

template <int KIND> __global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t seed) {
    __shared__ uint32_t tab[4096];
    uint32_t a[CHAINS], b[CHAINS], acc = 0, acc2 = 0;
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = i * seed;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = seed * (c + 1) + t; b[c] = seed ^ (t * 2654435761u + c); }
    const uint32_t c1 = 0xfffffff7u + (seed & 1);
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
#pragma unroll
            for (int c = 0; c < CHAINS; c++) {
#define MP(out, x, y, k) asm volatile("{\n\t.reg .pred pl, ph;\n\t.reg .s16 x0, x1, y0, y1;\n\t" \
                        "max.s16x2 %0, %3, %4;\n\tmov.b32 {x0, x1}, %0;\n\tmov.b32 {y0, y1}, %3;\n\t" \
                        "setp.eq.s16 pl, x0, y0;\n\tsetp.eq.s16 ph, x1, y1;\n\t" \
                        "@pl add.u32 %1, %1, " #k ";\n\t@ph add.u32 %2, %2, " #k ";\n\t}" \
                        : "=&r"(out), "+r"(acc), "+r"(acc2) : "r"(x), "r"(y))
#define MU(out, x, y, k) asm volatile("{\n\t" \
                        "max.s16x2 %0, %3, %4;\n\t" \
                        "add.u32 %1, %1, %0;\n\tadd.u32 %2, %2, %3;\n\t}" \
                        : "=&r"(out), "+r"(acc), "+r"(acc2) : "r"(x), "r"(y))
#define MN(out, x, y, k) asm volatile("max.s16x2 %0, %1, %2;" : "=r"(out) : "r"(x), "r"(y))
#define MI(out, x, y, k) asm volatile("{\n\t.reg .pred pl, ph;\n\t.reg .s16 x0, x1, y0, y1;\n\t" \
                        "max.s16x2 %0, %3, %4;\n\tmov.b32 {x0, x1}, %0;\n\tmov.b32 {y0, y1}, %3;\n\t" \
                        "setp.eq.s16 pl, x0, y0;\n\tsetp.eq.s16 ph, x1, y1;\n\t" \
                        "@pl mad.lo.u32 %1, %5, " #k ", %1;\n\t@ph mad.lo.u32 %2, %5, " #k ", %2;\n\t}" \
                        : "=&r"(out), "+r"(acc), "+r"(acc2) : "r"(x), "r"(y), "r"(one))
#define MPU(out, x, y, k) asm volatile("{\n\t.reg .pred pl, ph;\n\t.reg .u16 x0, x1, y0, y1;\n\t" \
                        "max.u16x2 %0, %3, %4;\n\tmov.b32 {x0, x1}, %0;\n\tmov.b32 {y0, y1}, %3;\n\t" \
                        "setp.eq.u16 pl, x0, y0;\n\tsetp.eq.u16 ph, x1, y1;\n\t" \
                        "@pl add.u32 %1, %1, " #k ";\n\t@ph add.u32 %2, %2, " #k ";\n\t}" \
                        : "=&r"(out), "+r"(acc), "+r"(acc2) : "r"(x), "r"(y))
#define MNU(out, x, y, k) asm volatile("max.u16x2 %0, %1, %2;" : "=r"(out) : "r"(x), "r"(y))
#define CELL32(M_) { uint32_t fe, t, to; \
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(fe) : "r"(b[c]), "r"(c1)); \
                    M_(b[c], a[c], fe, 0x4); \
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(fe) : "r"(a[c]), "r"(c1)); \
                    M_(t, b[c], fe, 0x8); \
                    M_(to, t, b[c], 0x2); \
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(to) : "r"(to), "r"(c1)); \
                    asm volatile("add.u32 %0, %1, %2;" : "=r"(fe) : "r"(a[c]), "r"(t)); \
                    M_(a[c], fe, to, 0x1); }
#define CELL(M_) { uint32_t fe, t, to; \
                    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(fe) : "r"(b[c]), "r"(c1)); \
                    M_(b[c], a[c], fe, 0x4); \
                    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(fe) : "r"(a[c]), "r"(c1)); \
                    M_(t, b[c], fe, 0x8); \
                    M_(to, t, b[c], 0x2); \
                    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(to) : "r"(to), "r"(c1)); \
                    asm volatile("add.s16x2 %0, %1, %2;" : "=r"(fe) : "r"(a[c]), "r"(t)); \
                    M_(a[c], fe, to, 0x1); }
                const uint32_t one = seed & 1;
                if (KIND == 0) CELL(MP)
                else if (KIND == 1) CELL(MU)
                else if (KIND == 2) CELL(MN)
                else if (KIND == 3) CELL(MI)
                else if (KIND == 4) CELL32(MPU)
                else if (KIND == 5) CELL32(MNU)
                else if (KIND == 6) CELL32(MP)
            }
        }
    }
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= a[c] + b[c];
    out[t] = acc + acc2;
}

template <int KIND> void run(const char *name, int warps_per_smsp) {
    int dev; cudaGetDevice(&dev); cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    const int threads = 32 * warps_per_smsp * 4 > 256 ? 256 : 32 * warps_per_smsp * 4;
    const int blocks = prop.multiProcessorCount * (32 * warps_per_smsp * 4 / threads), iters = 2048;
    uint32_t *out; cudaMalloc(&out, threads * blocks * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0); k<KIND><<<blocks, threads>>>(out, iters, 12345u); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r) best = ms < best ? ms : best;
    }
    double inner = (double)threads * blocks * iters * UNROLL * CHAINS;
    double clk = 1.965e9 * prop.multiProcessorCount * 4;  // SMSP-cycles per second
    printf("%-52s w/smsp=%2d %8.3f ms  SMSP-cycles per warp-inner = %.2f  (%s)\n", name, warps_per_smsp, best,
           clk * (best * 1e-3) / (inner / 32), cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int w : {16, 3}) {
        run<0>("cell: pred adds", w);
        run<1>("cell: unpredicated adds", w);
        run<2>("cell: no adds (4 add16 + 4 max)", w);
        run<3>("cell: pred IMAD", w);
        run<4>("cell32: u16x2 max + pred adds + 32-bit adds", w);
        run<5>("cell32: no pred adds (4 add32 + 4 maxu)", w);
        run<6>("cell32: s16x2 max + pred adds + 32-bit adds", w);
    }
    return 0;
}
