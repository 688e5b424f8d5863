// peak_kernels.cuh — roofline microbenchmarks for the integer multiply-add pipe
// (b200_imad_peak): the measured denominators of the IMAD roofline in bench.py.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace b200 {

// ptxas optimises straight through inline PTX: a `mad.wide` whose multiplicands are loop
// invariant is hoisted, and one with a 64-bit addend and a linear dependence is split into
// IMAD.WIDE(RZ) + IADD3 pairs.  Both kernels below therefore use loop-carried, non-linear
// operands, and their SASS was checked (cuobjdump) to contain the intended instruction mix.
//
// kind 0: the multiply-add stream of the Montgomery product itself — detail::mont_step rows
// (IMAD.WIDE.U32 with carry-out + IMAD.WIDE.U32.X chains sharing one multiplier, plus the
// one IMAD that forms the quotient digit), two independent accumulator sets per thread.
// Counted: 16 wide + 1 narrow multiply-adds per step.
static __global__ void __launch_bounds__(256) k_peak_imad_wide(uint64_t *out, uint32_t iters, uint32_t seed)
{
    uint32_t e0[8], o0[8], e1[8], o1[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        e0[k] = seed * (k + 1) + threadIdx.x;
        o0[k] = seed * (k + 9) ^ blockIdx.x;
        e1[k] = seed * (k + 17) + threadIdx.x * 3;
        o1[k] = seed * (k + 25) ^ (blockIdx.x * 5);
    }
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
        detail::mont_step<FqParams>(e0, o0, o0, o0[3] ^ e0[5], false);
        detail::mont_step<FqParams>(e1, o1, o1, o1[3] ^ e1[5], false);
        detail::mont_step<FqParams>(o0, e0, e0, e0[3] ^ o0[5], false);
        detail::mont_step<FqParams>(o1, e1, e1, e1[3] ^ o1[5], false);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= e0[k] ^ o0[k] ^ e1[k] ^ o1[k];
    if (s == 0x1234567u) out[0] = s;
}

// kind 1: IMAD (32 x 32 + 32, low half), 8 independent chains, acc = acc * neighbour + acc
static __global__ void __launch_bounds__(256) k_peak_imad(uint64_t *out, uint32_t iters, uint32_t seed)
{
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = (seed + k) * 2654435761u + threadIdx.x + blockIdx.x;
#pragma unroll 4
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(acc[k]) : "r"(acc[(k + 3) & 7]));
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
