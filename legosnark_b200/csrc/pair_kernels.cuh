// pair_kernels.cuh — batch-affine bucket accumulation (north star: "batch-affine bucket reduction"; the reference's
// own structure is batch_to_special + mixed_add, multiexp.tcc:240-242, with batch_invert, field_utils.tcc:171-194).
//
// The XYZZ mixed addition of k_accumulate costs 10 field products (1232 multiply-adds).  An AFFINE addition
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// costs 1 product + 1 squaring + 1 product once 1 / (x2 - x1) is known, and Montgomery's trick shares one inversion
// among all the denominators of a batch for 3 more products each: 5 products + 1 squaring = 788 multiply-adds.
// A bucket is a sum of points that are all known up front, so its additions can be arranged as a balanced tree whose
// levels are batches of independent additions:
//
//   level 0   slots (2q, 2q+1) of the bucket-ordered `entries` (bases, signs applied)      -> PA1[q]
//   level 1   (PA1[2r], PA1[2r+1])                                                          -> PA2[r]
//   ...       the remaining ceil(cnt / 2^levels) points of a bucket are summed by k_accumulate_pa (XYZZ)
//
// Bucket ranges are aligned to 2^levels slots by the sort (SortGeom::align_log), so pairs never straddle buckets; a slot
// holding ENTRY_PAD is absent (a pair with one absent side copies the other through).
//
// One kernel per level.  Thread g of T owns pairs g, g + T, g + 2T, ... (K = ceil(P / T) of them):
//   forward   prefix[k T + g] = product of the thread's denominators before pair k (coalesced scratch);
//   block     the threads' products are multiplied up in shared memory (two Hillis-Steele scans: prefix and suffix
//             products), ONE Fermat inversion per block, every thread gets the inverse of its own product;
//   backward  pairs in reverse: 1 / den_k = inv * prefix_k, inv *= den_k; lambda, x3, y3 -> dst.
// Exceptional cases (the reference's adders handle them too, alt_bn128_g1.cpp:139-195): equal points (tangent:
// lambda = 3 x^2 / 2 y, denominator 2 y), opposite points (result: the affine zero (0, 0), denominator 1), an absent or
// zero operand (copy, denominator 1).
#pragma once
#include "msm_kernels.cuh"

namespace b200 {

constexpr int PAIR_THREADS = 256;

template <class F>
struct PairOperands {
    F x1, y1, x2, y2;
    int kind;  // 0: nothing to write; 1: copy (x1, y1); 2: chord; 3: tangent at (x1, y1); 4: result is zero
};

// level 0: operands from the bases through the entries; LEVEL > 0: from the previous level's array
template <class F, int LEVEL>
__device__ __forceinline__ PairOperands<F> load_pair(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                      const Affine<F> *__restrict__ prev, size_t q)
{
    PairOperands<F> o;
    o.kind = 0;
    bool have1, have2;
    Affine<F> p1, p2;
    if (LEVEL == 0) {
        const uint2 e = *reinterpret_cast<const uint2 *>(entries + 2 * q);
        have1 = e.x != ENTRY_PAD;
        have2 = e.y != ENTRY_PAD;
        if (have1) {
            p1 = bases[e.x & 0x7fffffffu];
            if (e.x >> 31) p1.y = F::neg(p1.y);
        }
        if (have2) {
            p2 = bases[e.y & 0x7fffffffu];
            if (e.y >> 31) p2.y = F::neg(p2.y);
        }
    } else {
        // element j of level LEVEL covers slots [j 2^LEVEL, (j + 1) 2^LEVEL): present iff its first slot holds an entry
        have1 = entries[(2 * q) << LEVEL] != ENTRY_PAD;
        have2 = have1 && entries[(2 * q + 1) << LEVEL] != ENTRY_PAD;
        if (have1) p1 = prev[2 * q];
        if (have2) p2 = prev[2 * q + 1];
    }
    if (!have1) return o;
    o.x1 = p1.x;
    o.y1 = p1.y;
    o.kind = 1;
    if (!have2 || p2.is_inf()) return o;
    if (p1.is_inf()) {
        o.x1 = p2.x;
        o.y1 = p2.y;
        return o;
    }
    o.x2 = p2.x;
    o.y2 = p2.y;
    if (p1.x == p2.x) o.kind = (p1.y == p2.y) ? 3 : 4;
    else o.kind = 2;
    return o;
}

template <class F>
__device__ __forceinline__ F pair_denominator(const PairOperands<F> &o)
{
    if (o.kind == 2) return F::sub(o.x2, o.x1);
    if (o.kind == 3) return F::dbl(o.y1);
    return F::one();
}

template <class F, int LEVEL>
__global__ void __launch_bounds__(PAIR_THREADS) k_pair_add(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                           const Affine<F> *__restrict__ prev, const uint32_t *__restrict__ totals,
                                                           F *__restrict__ prefix, Affine<F> *__restrict__ dst)
{
    __shared__ F sm[PAIR_THREADS];
    __shared__ F sm_inv;
    const size_t P = ((size_t)totals[5] + ((2u << LEVEL) - 1)) >> (LEVEL + 1);  // pairs at this level
    const size_t T = (size_t)gridDim.x * PAIR_THREADS, g = (size_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
    const uint32_t K = (uint32_t)((P + T - 1) / T);
    // ---- forward: running product of the denominators
    F acc = F::one();
    for (uint32_t k = 0; k < K; k++) {
        const size_t q = (size_t)k * T + g;
        if (q >= P) break;
        const PairOperands<F> o = load_pair<F, LEVEL>(bases, entries, prev, q);
        prefix[q] = acc;
        if (o.kind == 2 || o.kind == 3) acc = F::mul(acc, pair_denominator(o));
    }
    // ---- block: inverse of every thread's product from one inversion
    const uint32_t t = threadIdx.x;
    sm[t] = acc;
    __syncthreads();
    F pre = acc;  // inclusive prefix product over the threads
    for (uint32_t d = 1; d < PAIR_THREADS; d <<= 1) {
        F other;
        if (t >= d) other = sm[t - d];
        __syncthreads();
        if (t >= d) {
            pre = F::mul(pre, other);
            sm[t] = pre;
        }
        __syncthreads();
    }
    if (t == PAIR_THREADS - 1) sm_inv = F::inv(pre);
    __syncthreads();
    const F inv_total = sm_inv;
    const F excl = t ? sm[t - 1] : F::one();
    __syncthreads();
    sm[t] = acc;
    __syncthreads();
    F suf = acc;  // inclusive suffix product
    for (uint32_t d = 1; d < PAIR_THREADS; d <<= 1) {
        F other;
        if (t + d < PAIR_THREADS) other = sm[t + d];
        __syncthreads();
        if (t + d < PAIR_THREADS) {
            suf = F::mul(suf, other);
            sm[t] = suf;
        }
        __syncthreads();
    }
    const F after = t + 1 < PAIR_THREADS ? sm[t + 1] : F::one();
    F inv = F::mul(F::mul(inv_total, excl), after);  // 1 / acc
    // ---- backward
    for (int k = (int)K - 1; k >= 0; k--) {
        const size_t q = (size_t)k * T + g;
        if (q >= P) continue;
        const PairOperands<F> o = load_pair<F, LEVEL>(bases, entries, prev, q);
        if (o.kind == 0) continue;
        Affine<F> r;
        if (o.kind == 1) {
            r.x = o.x1;
            r.y = o.y1;
        } else if (o.kind == 4) {
            r = Affine<F>::inf();
        } else {
            const F den = pair_denominator(o);
            const F dinv = F::mul(inv, prefix[q]);
            inv = F::mul(inv, den);
            F lambda, xs;
            if (o.kind == 2) {
                lambda = F::mul(F::sub(o.y2, o.y1), dinv);
                xs = F::add(o.x1, o.x2);
            } else {
                const F xx = F::sqr(o.x1);
                lambda = F::mul(F::add(F::dbl(xx), xx), dinv);
                xs = F::dbl(o.x1);
            }
            r.x = F::sub(F::sqr(lambda), xs);
            r.y = F::sub(F::mul(lambda, F::sub(o.x1, r.x)), o.y1);
        }
        dst[q] = r;
    }
}

// the tail of every task after `levels` tree levels: its ceil(len / 2^levels) remaining points, XYZZ mixed additions
template <class F>
__global__ void __launch_bounds__(128) k_accumulate_pa(const Affine<F> *__restrict__ pa, uint32_t levels, const uint2 *__restrict__ meta,
                                                        const uint32_t *__restrict__ order, const uint32_t *__restrict__ totals,
                                                        XYZZ<F> *__restrict__ partial, const uint32_t *__restrict__ task_bucket,
                                                        const XYZZ<F> *__restrict__ seed)
{
    const uint32_t ntasks = totals[1];
    const uint32_t gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= ntasks) return;
    const uint32_t t = order[gidx];
    const uint2 m = meta[t];  // m.x: first slot (a multiple of 2^levels), m.y: entries
    const Affine<F> *src = pa + (m.x >> levels);
    const uint32_t cnt = (m.y + (1u << levels) - 1) >> levels;
    XYZZ<F> acc = XYZZ<F>::inf();
    if (seed) {
        const uint32_t tb = task_bucket[t];
        if (tb >> 31) acc = seed[tb & 0x3fffffffu];
    }
    if (sizeof(F) <= 32) {  // G1: next point in registers during the addition; G2: no room (see k_accumulate)
        Affine<F> p = src[0];
        for (uint32_t k = 0; k < cnt; k++) {
            const Affine<F> q = p;
            if (k + 1 < cnt) p = src[k + 1];
            if (!q.is_inf()) xyzz_madd(acc, q.x, q.y, false);
        }
    } else {
        for (uint32_t k = 0; k < cnt; k++) {
            const Affine<F> q = src[k];
            if (!q.is_inf()) xyzz_madd(acc, q.x, q.y, false);
        }
    }
    partial[t] = acc;
}

}  // namespace b200
