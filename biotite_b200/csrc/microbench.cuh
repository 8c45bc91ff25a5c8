// Register-only instruction-rate probes: the denominators of the integer-ALU
// roofline the fill kernels are measured against.
#pragma once
#include <stdint.h>

namespace bt {

constexpr int kMicroChains = 8;   // independent dependency chains per thread
constexpr int kMicroUnroll = 8;   // rounds per loop iteration

__host__ __device__ constexpr int microbench_instr_per_iter(int kind) {
    // per round and chain: kind 0/1 -> 2 DPX instructions; kind 2 -> 1 DPX + 1 IMAD
    return kMicroChains * kMicroUnroll * 2 + 0 * kind;
}

template <int KIND>
__global__ void __launch_bounds__(256) microbench_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t a[kMicroChains], b[kMicroChains];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kMicroChains; k++) {
        a[k] = seed * (k + 1) + t;
        b[k] = seed ^ (t * 2654435761u + k);
    }
    const uint32_t c1 = 0xfffffff7u + (seed & 1), c2 = 0x00010001u * (seed & 3);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < kMicroUnroll; u++) {
#pragma unroll
            for (int k = 0; k < kMicroChains; k++) {
                if (KIND == 0) {
                    a[k] = __viaddmax_s16x2(a[k], c1, b[k]);
                    b[k] = __vimax3_s16x2(b[k], a[k], c2);
                } else if (KIND == 1) {
                    a[k] = (uint32_t)__viaddmax_s32((int)a[k], (int)c1, (int)b[k]);
                    b[k] = (uint32_t)__vimax3_s32((int)b[k], (int)a[k], (int)c2);
                } else {
                    a[k] = __viaddmax_s16x2(a[k], c1, b[k]);
                    b[k] = b[k] * 5u + a[k];  // IMAD: the FMA pipe issues beside the ALU pipe
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kMicroChains; k++) acc ^= a[k] + b[k];
    out[t] = acc;
}

}  // namespace bt
