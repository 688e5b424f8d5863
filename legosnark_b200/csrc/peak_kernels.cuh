// peak_kernels.cuh — roofline microbenchmarks for the integer multiply-add pipe
// (b200_imad_peak): the measured denominators of the IMAD roofline in bench.py.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace b200 {

// kind 0: IMAD.WIDE.U32 (64-bit accumulate), 8 independent chains per thread
static __global__ void __launch_bounds__(256) k_peak_imad_wide(uint64_t *out, uint32_t iters, uint32_t seed)
{
    uint64_t acc[8];
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k + threadIdx.x;
#pragma unroll 8
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a), "r"(b));
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567u) out[0] = s;
}

// kind 1: IMAD (32-bit), 8 independent chains
static __global__ void __launch_bounds__(256) k_peak_imad(uint64_t *out, uint32_t iters, uint32_t seed)
{
    uint32_t acc[8];
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = k + threadIdx.x;
#pragma unroll 8
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= acc[k];
    if (s == 0x1234567u) out[0] = s;
}

// kind 2: dependent Fq::mul chains, 2 per thread
static __global__ void __launch_bounds__(256) k_peak_modmul(uint64_t *out, uint32_t iters, uint32_t seed)
{
    Fq x = Fq::one(), y = Fq::r2();
    x.l[0] ^= (seed + threadIdx.x) & 0xffffu;
    y.l[0] ^= (blockIdx.x) & 0xffffu;
    for (uint32_t it = 0; it < iters; it++) {
        x = Fq::mul(x, y);
        y = Fq::mul(y, x);
    }
    if (x.l[0] == 0x1234567u && y.l[3] == 77u) out[0] = x.l[1];
}

}  // namespace b200
