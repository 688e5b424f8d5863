// fr_kernels.cuh — scalar-field (Fr) vector kernels on either side of the MSMs
// (SURVEY.md §8(f) rows 2 and 3):
//
//   k_fr_fold        CPPoly::prove's witness folding (LS/gadgets/poly.h:52-67) and, with no
//                    witness output, MultiVPolyT::evalMLE (LS/prototools/polytools.h:207-234):
//                    level i pairs (2p, 2p+1):  w[p] = t[2p+1] - t[2p],
//                    t'[p] = -t[2p] (r_i - 1) + t[2p+1] r_i = t[2p] + r_i w[p]
//   k_fr_bind_hi     DPMle::pushRandomness (LS/prototools/mle.h:199-210): pairs (p, p + half)
//   k_fr_eq_step     DPBeta::compute_eq_tbl (mle.h:93-105), one level per launch
//   k_fr_matrix_mle  DPMatrixMle's constructor (mle.h:241-259): matrix rows folded with the eq table
//   k_fr_sumcheck_round  the sum over p inside CPSumcheck::make_new_h_poly (LS/gadgets/sumcheck.h:85-106)
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
// sum-check dynamic-programming tables (LS/prototools/mle.h, LS/gadgets/sumcheck.h)
// ------------------------------------------------------------------------------
// One level of DPBeta::compute_eq_tbl (mle.h:93-105), restated as written there:
//     tmp[p] = eqbit(p >= 2^j, r[j]) * dst[p >> 1],   p < 2^(j+1)
// (level 0 is the same line with dst = {1}).  The reference indexes the previous level by p >> 1, not by the
// low bits of p; the kernel keeps that indexing, so the table is the reference's table, limb for limb.
static __global__ void __launch_bounds__(256) k_fr_eq_step(const Fr *__restrict__ src, const Fr *__restrict__ r, uint32_t j,
                                                           Fr *__restrict__ dst)
{
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ((size_t)2 << j)) return;
    const Fr rj = r[j];
    const Fr e = p >= ((size_t)1 << j) ? rj : Fr::sub(Fr::one(), rj);  // eqbit(bool, r), mle.cc:13-16
    const Fr prev = j == 0 ? Fr::one() : src[p >> 1];
    dst[p] = Fr::mul(e, prev);
}

// DPMatrixMle's constructor (mle.h:241-259): v[r] = sum_l A[(l << d) + r] * eq[l], n = 2^d rows l and columns r.
// Block = 32 columns x 8 row groups: a warp reads 32 consecutive elements of one matrix row (1 KB), the eight
// partial sums of a column meet in shared memory.  gridDim.y splits the rows further; the y-slices are added by
// k_fr_vec_add_slices.
constexpr int MMLE_THREADS = 256;
static __global__ void __launch_bounds__(MMLE_THREADS) k_fr_matrix_mle(const Fr *__restrict__ A, const Fr *__restrict__ eq, uint32_t d,
                                                                       Fr *__restrict__ slices)
{
    __shared__ Fr sm[MMLE_THREADS];
    const size_t n = (size_t)1 << d;
    const uint32_t cl = threadIdx.x & 31u, lg = threadIdx.x >> 5;
    const size_t r = (size_t)blockIdx.x * 32 + cl;
    const size_t rows_per = (n + gridDim.y - 1) / gridDim.y;
    const size_t l0 = (size_t)blockIdx.y * rows_per, l1 = min(n, l0 + rows_per);
    Fr acc = Fr::zero();
    if (r < n)
        for (size_t l = l0 + lg; l < l1; l += 8) acc = Fr::add(acc, Fr::mul(A[(l << d) + r], eq[l]));
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (lg == 0 && r < n) {
#pragma unroll 1
        for (int g = 1; g < 8; g++) acc = Fr::add(acc, sm[g * 32 + cl]);
        slices[(size_t)blockIdx.y * n + r] = acc;
    }
}

// out[i] = sum_y slices[y * n + i]
static __global__ void __launch_bounds__(256) k_fr_vec_add_slices(const Fr *__restrict__ slices, size_t n, uint32_t ny, Fr *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr acc = slices[i];
    for (uint32_t y = 1; y < ny; y++) acc = Fr::add(acc, slices[(size_t)y * n + i]);
    out[i] = acc;
}

// The sum inside CPSumcheck::make_new_h_poly (sumcheck.h:85-106) for two committed polynomials:
//     S(x) = sum_{p < half} w[p] * (a[p] (1 - x) + a[p + half] x) * (b[p] (1 - x) + b[p + half] x)
// (getMLEPoly, mle.h:218-227: eqbit_poly(0) * v0 + eqbit_poly(1) * v1), w = the beta suffix table of the round or
// nullptr (DPBetaDummy: the matrix sum-check).  The three coefficients are exact field sums, so any
// association gives the reference's limbs; the linear factor eqbit_poly(rho[j]) * beta_pre of a real beta is
// applied by the caller to the three sums.  Block partials -> part[block][3]; k_fr_sum3 adds them.
constexpr int SC_THREADS = 256;
static __global__ void __launch_bounds__(SC_THREADS) k_fr_sumcheck_round(const Fr *__restrict__ a, const Fr *__restrict__ b,
                                                                         const Fr *__restrict__ w, size_t half, Fr *__restrict__ part)
{
    __shared__ Fr sm[3][SC_THREADS];
    Fr c0 = Fr::zero(), c1 = Fr::zero(), c2 = Fr::zero();
    for (size_t p = (size_t)blockIdx.x * SC_THREADS + threadIdx.x; p < half; p += (size_t)gridDim.x * SC_THREADS) {
        Fr a0 = a[p], a1 = Fr::sub(a[p + half], a0);
        const Fr b0 = b[p], b1 = Fr::sub(b[p + half], b0);
        if (w) {
            const Fr wp = w[p];
            a0 = Fr::mul(a0, wp);
            a1 = Fr::mul(a1, wp);
        }
        c0 = Fr::add(c0, Fr::mul(a0, b0));
        c1 = Fr::add(c1, Fr::mul_add(a0, b1, a1, b0));
        c2 = Fr::add(c2, Fr::mul(a1, b1));
    }
    sm[0][threadIdx.x] = c0;
    sm[1][threadIdx.x] = c1;
    sm[2][threadIdx.x] = c2;
    __syncthreads();
    for (int o = SC_THREADS / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int k = 0; k < 3; k++) sm[k][threadIdx.x] = Fr::add(sm[k][threadIdx.x], sm[k][threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x < 3) part[(size_t)blockIdx.x * 3 + threadIdx.x] = sm[threadIdx.x][0];
}

// out[k] = sum_blocks part[block][k], k < 3 (one warp per coefficient)
static __global__ void __launch_bounds__(96) k_fr_sum3(const Fr *__restrict__ part, uint32_t nblocks, Fr *__restrict__ out)
{
    __shared__ Fr sm[96];
    const uint32_t k = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    Fr acc = Fr::zero();
    for (uint32_t i = lane; i < nblocks; i += 32) acc = Fr::add(acc, part[(size_t)i * 3 + k]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (lane == 0) {
        for (int i = 1; i < 32; i++) acc = Fr::add(acc, sm[k * 32 + i]);
        out[k] = acc;
    }
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

// P[i] *= (c1 * ratio^i - c0)^-1, i < n: the per-point divisions of step_radix2_domain::divide_by_Z_on_coset
// (FQFFT/evaluation_domain/domains/step_radix2_domain.tcc:219-232: one Fp inversion per element on the host there).
// Thread = run of INVG_RUN consecutive elements: the denominators are formed from ratio^i0 by one product each,
// inverted together (Montgomery's trick: 3 products per element + one Fermat inversion per run; field_utils.tcc:171-194
// is the reference's own batch_invert), and applied.  The inverse of a field element is unique: same limbs.
constexpr int INVG_RUN = 32;
static __global__ void __launch_bounds__(128) k_fr_scale_inv_geometric(Fr *__restrict__ P, size_t n, const Fr *__restrict__ consts /* c1, ratio, c0 */)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * INVG_RUN;
    if (i0 >= n) return;
    const Fr c1 = consts[0], ratio = consts[1], c0 = consts[2];
    const uint32_t cnt = (uint32_t)min((size_t)INVG_RUN, n - i0);
    Fr den[INVG_RUN], pre[INVG_RUN];
    Fr t = Fr::mul(c1, fr_pow(ratio, i0));  // c1 * ratio^i
    Fr acc = Fr::one();
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; k++) {
        den[k] = Fr::sub(t, c0);
        pre[k] = acc;
        acc = Fr::mul(acc, den[k]);
        t = Fr::mul(t, ratio);
    }
    Fr inv = Fr::inv(acc);
#pragma unroll 1
    for (int k = (int)cnt - 1; k >= 0; k--) {
        const Fr di = Fr::mul(inv, pre[k]);
        inv = Fr::mul(inv, den[k]);
        P[i0 + k] = Fr::mul(P[i0 + k], di);
    }
}

// out[i] = in[i] * a0 * a_ratio^i / prod_f (c1_f * ratio_f^i - c0_f), i < n, f < nf (1 or 2); in == nullptr: in[i] = 1.
// The per-point divisions of libfqfft's Lagrange evaluations, _basic_radix2_evaluate_all_lagrange_polynomials
// (FQFFT/evaluation_domain/domains/basic_radix2_domain_aux.tcc:225-233: u[i] = l * (t - r).inverse(), l *= omega, r *= omega)
// and step_radix2_domain::evaluate_all_lagrange_polynomials (step_radix2_domain.tcc:171-176), one Fp inversion per point on the
// host there.  Same run structure as k_fr_scale_inv_geometric; the numerator rides in the prefix products.
// consts: a0, a_ratio, then (c1, ratio, c0) per factor.
static __global__ void __launch_bounds__(128) k_fr_geometric_quotients(const Fr *in, Fr *out /* may alias in */, size_t n,
                                                                       const Fr *__restrict__ consts, int nf)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * INVG_RUN;
    if (i0 >= n) return;
    const Fr a_ratio = consts[1], r1 = consts[3], z1 = consts[4];
    const Fr r2 = nf > 1 ? consts[6] : Fr::one(), z2 = nf > 1 ? consts[7] : Fr::zero();
    const uint32_t cnt = (uint32_t)min((size_t)INVG_RUN, n - i0);
    Fr den[INVG_RUN], pre[INVG_RUN];
    Fr num = Fr::mul(consts[0], fr_pow(a_ratio, i0));
    Fr t1 = Fr::mul(consts[2], fr_pow(r1, i0));
    Fr t2 = nf > 1 ? Fr::mul(consts[5], fr_pow(r2, i0)) : Fr::one();
    Fr acc = Fr::one();
#pragma unroll 1
    for (uint32_t k = 0; k < cnt; k++) {
        Fr d = Fr::sub(t1, z1);
        if (nf > 1) d = Fr::mul(d, Fr::sub(t2, z2));
        den[k] = d;
        pre[k] = Fr::mul(acc, num);
        acc = Fr::mul(acc, d);
        num = Fr::mul(num, a_ratio);
        t1 = Fr::mul(t1, r1);
        if (nf > 1) t2 = Fr::mul(t2, r2);
    }
    Fr inv = Fr::inv(acc);
#pragma unroll 1
    for (int k = (int)cnt - 1; k >= 0; k--) {
        Fr q = Fr::mul(inv, pre[k]);
        inv = Fr::mul(inv, den[k]);
        if (in) q = Fr::mul(q, in[i0 + k]);
        out[i0 + k] = q;
    }
}

// x[i] = x[i] * y[i] - z[i], i < n: H on the coset before the division, r1cs_to_qap.tcc:270-300
static __global__ void __launch_bounds__(256) k_fr_mul_sub(Fr *__restrict__ x, const Fr *__restrict__ y, const Fr *__restrict__ z, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = Fr::sub(Fr::mul(x[i], y[i]), z[i]);
}

// P[i] *= s, i < n
static __global__ void __launch_bounds__(256) k_fr_scale(Fr *__restrict__ P, size_t n, const Fr *__restrict__ s)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) P[i] = Fr::mul(P[i], *s);
}

// ------------------------------------------------------------------------------
// step_radix2_domain (FQFFT/evaluation_domain/domains/step_radix2_domain.tcc): domains of 2^k + 2^r points, the one
// libfqfft picks for the 128 x 128 matrix product of BASELINE.json configs[3] (2^21 + 1).  The two radix-2 transforms
// inside run through k_fr_fft_pass; these kernels are the O(m) loops the reference wraps around them (:38-71, :73-139),
// which are serial host code there (omega_i *= omega chains).  ow[] = omega^i, i < big (the twiddle table of the
// 2 * big domain); gp[] = g^i, i < big + small, for the coset variants (_multiply_by_coset, aux.tcc:172-180).
// ------------------------------------------------------------------------------
// FFT, first loop (:43-50): c[i] = a[i] + a[i+big], d[i] = omega^i (a[i] - a[i+big]) for i < small; c[i] = a[i], d[i] = omega^i a[i] above
static __global__ void __launch_bounds__(256) k_step_fwd_pre(const Fr *__restrict__ a, const Fr *__restrict__ gp, const Fr *__restrict__ ow,
                                                             size_t big, size_t small, Fr *__restrict__ c, Fr *__restrict__ d)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= big) return;
    Fr lo = a[i];
    if (gp) lo = Fr::mul(lo, gp[i]);
    if (i < small) {
        Fr hi = a[i + big];
        if (gp) hi = Fr::mul(hi, gp[i + big]);
        c[i] = Fr::add(lo, hi);
        d[i] = Fr::mul(ow[i], Fr::sub(lo, hi));
    } else {
        c[i] = lo;
        d[i] = Fr::mul(ow[i], lo);
    }
}

// part[jc * small + i] = sum over j = j0 + jc, j0 + jc + J, ... < compr of (scale ? scale[i + j small] : 1) * d[i + j small]
// (the e[i] sums of FFT :52-60 with j0 = 0; the U1[i] -= sum tmp[i + j small] of iFFT :113-120 with j0 = 1 and scale = omega^i)
static __global__ void __launch_bounds__(256) k_step_strided_partial(const Fr *__restrict__ d, const Fr *__restrict__ scale, size_t small,
                                                                     size_t compr, uint32_t j0, uint32_t J, Fr *__restrict__ part)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= small * J) return;
    const size_t i = t % small, jc = t / small;
    Fr acc = Fr::zero();
    for (size_t j = j0 + jc; j < compr; j += J) {
        Fr v = d[i + j * small];
        if (scale) v = Fr::mul(v, scale[i + j * small]);
        acc = Fr::add(acc, v);
    }
    part[jc * small + i] = acc;
}
// out[i] = sum_jc part[jc small + i]
static __global__ void __launch_bounds__(256) k_step_strided_final(const Fr *__restrict__ part, size_t small, uint32_t J, Fr *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= small) return;
    Fr acc = part[i];
    for (uint32_t jc = 1; jc < J; jc++) acc = Fr::add(acc, part[(size_t)jc * small + i]);
    out[i] = acc;
}

// iFFT, last loops (:106-138) given U0 = iFFT_big(a[0..big)), U1 = iFFT_small(a[big..)), S[i] = sum_{j >= 1} omega^(i + j small) U0[i + j small]:
//   U1'[i] = (U1[i] - S[i]) omega^-i;  a[i] = (U0[i] + U1'[i]) / 2, a[big + i] = (U0[i] - U1'[i]) / 2 for i < small;  a[i] = U0[i] above;
// then the optional coset scaling a[i] *= g^-i (icosetFFT :148-152).  owi[] = omega^-i; half = 1/2.
static __global__ void __launch_bounds__(256) k_step_inv_post(const Fr *__restrict__ U0, const Fr *__restrict__ U1, const Fr *__restrict__ S,
                                                              const Fr *__restrict__ owi, const Fr *__restrict__ gip, const Fr *__restrict__ half,
                                                              size_t big, size_t small, Fr *__restrict__ a)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= big) return;
    const Fr u0 = U0[i];
    if (i < small) {
        const Fr u1 = Fr::mul(Fr::sub(U1[i], S[i]), owi[i]);
        Fr lo = Fr::mul(Fr::add(u0, u1), *half), hi = Fr::mul(Fr::sub(u0, u1), *half);
        if (gip) {
            lo = Fr::mul(lo, gip[i]);
            hi = Fr::mul(hi, gip[i + big]);
        }
        a[i] = lo;
        a[i + big] = hi;
    } else {
        a[i] = gip ? Fr::mul(u0, gip[i]) : u0;
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
