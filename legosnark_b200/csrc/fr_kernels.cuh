// fr_kernels.cuh — scalar-field (Fr) vector kernels on either side of the MSMs
// (SURVEY.md §8(f) rows 2 and 3):
//
//   k_fr_fold        CPPoly::prove's witness folding (LS/gadgets/poly.h:52-67) and, with no
//                    witness output, MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234):
//                    level i pairs (2p, 2p+1):  w[p] = t[2p+1] - t[2p],
//                    t'[p] = -t[2p] (r_i - 1) + t[2p+1] r_i = t[2p] + r_i w[p]
//   k_fr_bind_hi     DPMle::pushRandomness (LS/prototools/mle.h:199-210): pairs (p, p + half)
//   k_fr_fft_pass    libfqfft's basic radix-2 domain (FQFFT/evaluation_domain/domains/
//                    basic_radix2_domain_aux.tcc:42-75: bit reversal + log n butterfly stages;
//                    basic_radix2_domain.tcc: FFT / iFFT / cosetFFT / icosetFFT), several
//                    stages per pass in shared memory
//
// All values are Montgomery-form Fr elements as the reference holds them (fp.hpp:42); field
// arithmetic is exact, so any evaluation order gives the reference's limbs.
#pragma once
#include <cuda_runtime.h>

#include "field.cuh"

namespace b200 {

// ------------------------------------------------------------------------------
// multilinear folding
// ------------------------------------------------------------------------------
constexpr int FOLD_THREADS = 256;
constexpr int FOLD_LEVELS = 9;  // a block folds 2 * FOLD_THREADS = 2^9 values down to one

// in: 2^(d - i0) values.  The block folds its 512-value tile through levels i0 .. i0 + nl - 1
// (nl <= 9; nl < 9 only when fewer levels remain) and writes 512 >> nl values to `next`.
// w != nullptr: level i's witness coefficients go to w[start_i + p] with
// start_i = 2^d - 2^(d - i) (the reference's w_coeffs layout, poly.h:57-66).
static __global__ void __launch_bounds__(FOLD_THREADS) k_fr_fold(const Fr *__restrict__ in, const Fr *__restrict__ r, uint32_t d,
                                                                 uint32_t i0, uint32_t nl, Fr *__restrict__ next, Fr *__restrict__ w)
{
    __shared__ Fr sm[FOLD_THREADS];
    const size_t len = (size_t)1 << (d - i0);          // values at level i0
    const size_t pairs0 = len >> 1;
    const size_t p0 = (size_t)blockIdx.x * FOLD_THREADS + threadIdx.x;
    const size_t full = (size_t)1 << d;
    // level i0: straight from global memory (64 contiguous bytes per thread)
    Fr t = Fr::zero();
    if (p0 < pairs0) {
        const Fr a = in[2 * p0], b = in[2 * p0 + 1];
        const Fr diff = Fr::sub(b, a);
        if (w) w[(full - len) + p0] = diff;
        t = Fr::add(a, Fr::mul(r[i0], diff));
    }
    if (nl == 1) {
        if (p0 < pairs0) next[p0] = t;
        return;
    }
    sm[threadIdx.x] = t;
    __syncthreads();
    uint32_t active = FOLD_THREADS >> 1;
    for (uint32_t q = 1; q < nl; q++, active >>= 1) {
        const size_t lenq = len >> q;                  // values at level i0 + q
        const size_t pq = (size_t)blockIdx.x * active + threadIdx.x;
        Fr a, b;
        const bool live = threadIdx.x < active && pq < (lenq >> 1);
        if (live) {
            a = sm[2 * threadIdx.x];
            b = sm[2 * threadIdx.x + 1];
        }
        __syncthreads();
        if (live) {
            const Fr diff = Fr::sub(b, a);
            if (w) w[(full - lenq) + pq] = diff;
            t = Fr::add(a, Fr::mul(r[i0 + q], diff));
            sm[threadIdx.x] = t;
            if (q == nl - 1) next[pq] = t;
        }
        __syncthreads();
    }
}

// out[p] = table[p] (1 - r) + table[p + half] r   (eqbit(false, r) = 1 - r, eqbit(true, r) = r)
static __global__ void __launch_bounds__(256) k_fr_bind_hi(const Fr *__restrict__ table, size_t half, const Fr *__restrict__ r,
                                                           Fr *__restrict__ out)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= half) return;
    const Fr a = table[p], b = table[p + half];
    out[p] = Fr::add(a, Fr::mul(*r, Fr::sub(b, a)));
}

// ------------------------------------------------------------------------------
// radix-2 domain
// ------------------------------------------------------------------------------
// Fr::root_of_unity (order 2^28) and Fr::multiplicative_generator = 5, Montgomery form
// (alt_bn128_init.cpp:57-61; bn128_init.cpp has the same values)
__device__ __forceinline__ Fr fr_root_of_unity_2_28()
{
    Fr x;
    const uint32_t v[8] = {0x80d13d9cu, 0x636e7355u, 0x2445ffd6u, 0xa22bf374u, 0x1eb203d8u, 0x56452ac0u, 0x2963f9e7u, 0x1860ef94u};
#pragma unroll
    for (int i = 0; i < 8; i++) x.l[i] = v[i];
    return x;
}
constexpr uint32_t FR_TWO_ADICITY = 28;  // alt_bn128_Fr::s

__device__ inline Fr fr_pow(const Fr &base, uint64_t e)
{
    Fr r = Fr::one();
    if (e == 0) return r;
    int top = 63 - __clzll((long long)e);
    r = base;
    for (int i = top - 1; i >= 0; i--) {
        r = Fr::sqr(r);
        if ((e >> i) & 1ull) r = Fr::mul(r, base);
    }
    return r;
}

// consts[0] = omega (get_root_of_unity(2^logn), field_utils.tcc:38-51), [1] = omega^-1,
// [2] = (2^logn)^-1 (basic_radix2_domain.tcc iFFT: FieldT(a.size()).inverse()),
// [3] = g (coset shift, or one), [4] = g^-1.  One thread.
static __global__ void k_fr_domain_consts(uint32_t logn, const Fr *__restrict__ g, Fr *__restrict__ consts)
{
    if (threadIdx.x || blockIdx.x) return;
    Fr omega = fr_root_of_unity_2_28();
    for (uint32_t i = FR_TWO_ADICITY; i > logn; --i) omega = Fr::sqr(omega);
    consts[0] = omega;
    consts[1] = Fr::inv(omega);
    Fr n = Fr::zero();
    n.l[logn >> 5] = 1u << (logn & 31);
    consts[2] = Fr::inv(Fr::to_mont(n));
    const Fr gg = g ? *g : Fr::one();
    consts[3] = gg;
    consts[4] = Fr::inv(gg);
}

// out[i] = scale * base^i, i < count (scale == nullptr: 1).  Thread = run of 16 powers.
static __global__ void __launch_bounds__(128) k_fr_pow_table(const Fr *__restrict__ base, const Fr *__restrict__ scale, size_t count,
                                                             Fr *__restrict__ out)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= count) return;
    const Fr b = *base;
    Fr cur = fr_pow(b, i0);
    if (scale) cur = Fr::mul(cur, *scale);
    for (size_t t = 0; t < 16 && i0 + t < count; t++) {
        out[i0 + t] = cur;
        cur = Fr::mul(cur, b);
    }
}

constexpr int FFT_THREADS = 256;
constexpr int FFT_TILE_LOG = 10;  // a block holds 2^10 elements = 32 KB of shared memory

struct FftPass {
    uint32_t logn;  // transform size
    uint32_t s0;    // butterfly stages already done (this pass does stages s0 .. s0 + k - 1)
    uint32_t k;     // stages of this pass
    uint32_t t;     // log2 of the run of low-index neighbours a block takes along (s0 > 0): coalescing
};

// One pass = k butterfly stages on 2^(k + t) elements per block.  Stage s pairs indices that differ
// in bit s (m = 2^s, CLRS / basic_radix2_domain_aux.tcc:58-72) with twiddle omega^(j n / 2m),
// j = idx mod m, read from tw[] = omega^0 .. omega^(n/2 - 1).
// s0 == 0: the block owns 2^k consecutive outputs and loads src[bitreverse(idx)] (the swap loop at
// :48-53), times pre[bitreverse(idx)] when a coset shift is applied before the transform
// (_multiply_by_coset, :163-171).  The last pass multiplies by post[idx] (icosetFFT: n^-1 g^-idx)
// or by the scalar *post_scalar (iFFT: n^-1).
static __global__ void __launch_bounds__(FFT_THREADS) k_fr_fft_pass(const Fr *__restrict__ src, Fr *__restrict__ dst, FftPass P,
                                                                    const Fr *__restrict__ tw, const Fr *__restrict__ pre,
                                                                    const Fr *__restrict__ post, const Fr *__restrict__ post_scalar)
{
    // 8 limb planes of E words (a one-word-per-32 padding of the planes was measured: no change, the passes are
    // bound by the multiply-add pipe, not by shared-memory wavefronts)
    extern __shared__ uint32_t sm[];
    const uint32_t E = 1u << (P.k + P.t);
#define FFT_AT(i, e) sm[(i) * E + (e)]
    const uint32_t tmask = (1u << P.t) - 1u;
    // block -> (hi, lo0): idx = hi 2^(s0+k) + mid 2^s0 + lo0 + l
    size_t hi, lo0;
    if (P.s0 == 0) {
        hi = blockIdx.x;
        lo0 = 0;
    } else {
        const uint32_t runs = 1u << (P.s0 - P.t);
        hi = blockIdx.x / runs;
        lo0 = (size_t)(blockIdx.x % runs) << P.t;
    }
    const size_t base = (hi << (P.s0 + P.k)) + lo0;
    for (uint32_t e = threadIdx.x; e < E; e += FFT_THREADS) {
        const uint32_t mid = e >> P.t, l = e & tmask;
        const size_t idx = base + ((size_t)mid << P.s0) + l;
        Fr v;
        if (P.s0 == 0) {
            const size_t ridx = (size_t)(__brevll((unsigned long long)idx) >> (64 - P.logn));
            v = src[ridx];
            if (pre) v = Fr::mul(v, pre[ridx]);
        } else {
            v = src[idx];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) FFT_AT(i, e) = v.l[i];
    }
    __syncthreads();
    // two stages per trip through shared memory (radix 4 in registers): the thread owns the four
    // elements whose `mid` differs in bits q and q + 1.  Stage q uses one twiddle for both of its
    // butterflies (bit q + 1 lies above the stage), stage q + 1 uses two.
    uint32_t q = 0;
    for (; q + 1 < P.k; q += 2) {
        const uint32_t sh_a = P.logn - (P.s0 + q) - 1;  // stage q exponent shift; stage q + 1: sh_a - 1
        for (uint32_t g4 = threadIdx.x; g4 < (E >> 2); g4 += FFT_THREADS) {
            const uint32_t l = g4 & tmask, rest = g4 >> P.t;
            const uint32_t mid_lo = rest & ((1u << q) - 1u), mid_hi = rest >> q;
            const uint32_t mid0 = (mid_hi << (q + 2)) | mid_lo;
            const uint32_t e00 = (mid0 << P.t) | l, step = 1u << (q + P.t);
            const uint32_t e01 = e00 + step, e10 = e00 + 2 * step, e11 = e00 + 3 * step;
            const size_t ja = ((size_t)mid_lo << P.s0) + lo0 + l;        // idx mod 2^(s0+q)
            const size_t jb1 = ja + ((size_t)1 << (P.s0 + q));             // idx mod 2^(s0+q+1) when bit q is set
            Fr a00, a01, a10, a11;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a00.l[i] = FFT_AT(i, e00);
                a01.l[i] = FFT_AT(i, e01);
                a10.l[i] = FFT_AT(i, e10);
                a11.l[i] = FFT_AT(i, e11);
            }
            // stage q
            if (ja) {
                const Fr wa = tw[ja << sh_a];
                a01 = Fr::mul(wa, a01);
                a11 = Fr::mul(wa, a11);
            }
            Fr t = a01;
            a01 = Fr::sub(a00, t);
            a00 = Fr::add(a00, t);
            t = a11;
            a11 = Fr::sub(a10, t);
            a10 = Fr::add(a10, t);
            // stage q + 1: (a00, a10) with omega^(ja ...), (a01, a11) with omega^(jb1 ...)
            if (ja) a10 = Fr::mul(tw[ja << (sh_a - 1)], a10);
            a11 = Fr::mul(tw[jb1 << (sh_a - 1)], a11);
            t = a10;
            a10 = Fr::sub(a00, t);
            a00 = Fr::add(a00, t);
            t = a11;
            a11 = Fr::sub(a01, t);
            a01 = Fr::add(a01, t);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                FFT_AT(i, e00) = a00.l[i];
                FFT_AT(i, e01) = a01.l[i];
                FFT_AT(i, e10) = a10.l[i];
                FFT_AT(i, e11) = a11.l[i];
            }
        }
        __syncthreads();
    }
    for (; q < P.k; q++) {  // odd stage count: one radix-2 stage
        const uint32_t s = P.s0 + q;
        const uint32_t tw_shift = P.logn - s - 1;  // exponent = j * n / 2^(s+1)
        for (uint32_t bf = threadIdx.x; bf < (E >> 1); bf += FFT_THREADS) {
            const uint32_t l = bf & tmask, rest = bf >> P.t;
            const uint32_t mid_lo = rest & ((1u << q) - 1u), mid_hi = rest >> q;
            const uint32_t mid0 = (mid_hi << (q + 1)) | mid_lo;
            const uint32_t e0 = (mid0 << P.t) | l, e1 = e0 + (1u << (q + P.t));
            const size_t j = ((size_t)mid_lo << P.s0) + lo0 + l;
            Fr a, b;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                a.l[i] = FFT_AT(i, e0);
                b.l[i] = FFT_AT(i, e1);
            }
            const Fr tt = j ? Fr::mul(tw[j << tw_shift], b) : b;
            const Fr lo = Fr::add(a, tt), hi2 = Fr::sub(a, tt);
#pragma unroll
            for (int i = 0; i < 8; i++) {
                FFT_AT(i, e0) = lo.l[i];
                FFT_AT(i, e1) = hi2.l[i];
            }
        }
        __syncthreads();
    }
    const bool last = P.s0 + P.k == P.logn;
    for (uint32_t e = threadIdx.x; e < E; e += FFT_THREADS) {
        const uint32_t mid = e >> P.t, l = e & tmask;
        const size_t idx = base + ((size_t)mid << P.s0) + l;
        Fr v;
#pragma unroll
        for (int i = 0; i < 8; i++) v.l[i] = FFT_AT(i, e);
        if (last && post) v = Fr::mul(v, post[idx]);
        else if (last && post_scalar) v = Fr::mul(v, *post_scalar);
        dst[idx] = v;
    }
#undef FFT_AT
}

}  // namespace b200
