// msm_kernels.cuh — the device side of the MSM / batch_exp engine.
//
// Pipeline for out = sum_i s_i P_i (replaces multi_exp_inner<BDLO12>,
// LFF/algebra/scalar_multiplication/multiexp.tcc:165-282):
//
//   k_ingest            Jacobian bases -> affine (x,y) + zero flags; one Fermat
//                       inversion per thread shared by its K points (Montgomery
//                       trick; replaces batch_to_special, multiexp.tcc:683-715)
//   k_digit_count       s_i -> standard form (as_bigint, fp.tcc:227-238), signed
//                       c-bit digits, histogram over (window, |digit|) buckets
//   k_scan_*            exclusive scan: bucket offsets and accumulation-task offsets
//   k_digit_scatter     counting-sort scatter: bucket-ordered list of (index, sign)
//   k_task_meta/order   split buckets into tasks of <= L entries, order tasks by
//                       length (longest first) so the lanes of a warp do equal work
//   k_accumulate        one thread per task: gather affine points, XYZZ mixed adds
//   k_bucket_combine    buckets that were split: sum their task partials (warp each)
//   k_window_reduce1/2  per window: sum_j j * B_j by segment running sums, a small
//                       scalar multiplication per segment and tree reductions
//   host                Horner over the W window sums (host_arith.hpp)
//
// Reference semantics kept: zero scalars and zero bases contribute nothing,
// repeated / equal / opposite bases go through the doubling / infinity branches of
// the adders, any Jacobian representative is accepted.
#pragma once
#include <cuda_runtime.h>

#include "curve.cuh"

namespace b200 {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct MsmGeom {
    uint32_t c;   // window bits
    uint32_t W;   // windows
    uint32_t B;   // buckets per window = 2^(c-1)
    uint32_t L;   // max entries per accumulation task
    uint32_t NB;  // W * B
};

// ------------------------------------------------------------------------------
// Jacobian -> affine with a per-thread shared inversion.  Thread t owns points
// t, t+T, t+2T, ... (coalesced across the warp).  OUT_JAC: write (x, y, 1) /
// (0,1,0) Jacobian images in place of affine pairs (batch_to_special semantics).
// ------------------------------------------------------------------------------
template <class F, bool OUT_JAC>
__global__ void __launch_bounds__(128) k_ingest(const Jacobian<F> *__restrict__ in, void *__restrict__ out_,
                                               uint8_t *__restrict__ flags, F *__restrict__ prefix, size_t n)
{
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const F one = F::one();
    F acc = one;
    bool any = false;
    for (size_t i = t; i < n; i += T) {
        const F z = in[i].z;
        if (z.is_zero() || z == one) continue;
        prefix[i] = acc;
        acc = F::mul(acc, z);
        any = true;
    }
    F inv = any ? F::inv(acc) : one;
    if (t >= n) return;
    // walk this thread's points in reverse
    const size_t last = t + ((n - 1 - t) / T) * T;
    for (size_t i = last;; i -= T) {
        const Jacobian<F> p = in[i];
        Affine<F> a;
        bool isinf = false;
        if (p.z.is_zero()) {
            a = Affine<F>::inf();
            isinf = true;
        } else if (p.z == one) {
            a.x = p.x;
            a.y = p.y;
        } else {
            const F zi = F::mul(inv, prefix[i]);
            inv = F::mul(inv, p.z);
            const F z2 = F::sqr(zi);
            a.x = F::mul(p.x, z2);
            a.y = F::mul(p.y, F::mul(z2, zi));
        }
        if (OUT_JAC) {
            Jacobian<F> *out = reinterpret_cast<Jacobian<F> *>(out_);
            out[i] = isinf ? Jacobian<F>::inf() : Jacobian<F>{a.x, a.y, one};
        } else {
            Affine<F> *out = reinterpret_cast<Affine<F> *>(out_);
            out[i] = a;
        }
        if (flags) flags[i] = isinf ? 1 : 0;
        if (i == t) break;
    }
}

// flags for bases that are already affine on the device ((0,0) = zero)
template <class F>
__global__ void k_affine_flags(const Affine<F> *__restrict__ pts, uint8_t *__restrict__ flags, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = pts[i].is_inf() ? 1 : 0;
}

// ------------------------------------------------------------------------------
// signed-digit recoding
// ------------------------------------------------------------------------------
template <class Fn>
__device__ __forceinline__ void for_each_digit(const Fr &s, uint32_t c, uint32_t W, Fn fn)
{
    uint32_t limbs[9];
#pragma unroll
    for (int i = 0; i < 8; i++) limbs[i] = s.l[i];
    limbs[8] = 0;
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t k = 0; k < W; k++) {
        const uint32_t bit = k * c;
        const uint32_t limb = bit >> 5, sh = bit & 31u;
        uint32_t raw = 0;
        if (limb < 8) {
            const uint64_t two = (uint64_t)limbs[limb] | ((uint64_t)limbs[limb + 1] << 32);
            raw = (uint32_t)(two >> sh) & mask;
        }
        const uint32_t d = raw + carry;
        uint32_t mag, neg;
        if (d > half) {
            mag = (1u << c) - d;
            neg = 1;
            carry = 1;
        } else {
            mag = d;
            neg = 0;
            carry = 0;
        }
        if (mag) fn(k, mag, neg);
    }
}

static __global__ void __launch_bounds__(256) k_digit_count(const Fr *__restrict__ scalars_mont, const uint8_t *__restrict__ flags,
                                                      size_t n, MsmGeom g, uint32_t *__restrict__ cnt)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) return;
    const Fr s = Fr::from_mont(scalars_mont[i]);
    for_each_digit(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t) { atomicAdd(&cnt[k * g.B + (mag - 1)], 1u); });
}

static __global__ void __launch_bounds__(256) k_digit_scatter(const Fr *__restrict__ scalars_mont, const uint8_t *__restrict__ flags,
                                                        size_t n, MsmGeom g, uint32_t *__restrict__ cursor,
                                                        uint32_t *__restrict__ entries)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) return;
    const Fr s = Fr::from_mont(scalars_mont[i]);
    for_each_digit(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t neg) {
        const uint32_t pos = atomicAdd(&cursor[k * g.B + (mag - 1)], 1u);
        entries[pos] = (uint32_t)i | (neg << 31);
    });
}

// ------------------------------------------------------------------------------
// exclusive scan over the bucket counts; element = (entries, tasks)
// ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tasks_of(uint32_t cnt, uint32_t L) { return (cnt + L - 1) / L; }

// phase 1: per-tile totals
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile_sums(const uint32_t *__restrict__ cnt, uint32_t NB, uint32_t L,
                                                                 uint2 *__restrict__ tile_sums)
{
    __shared__ uint2 sm[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint2 v = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        if (idx < NB) {
            const uint32_t cv = cnt[idx];
            v.x += cv;
            v.y += tasks_of(cv, L);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v.x += __shfl_down_sync(0xffffffffu, v.x, o);
        v.y += __shfl_down_sync(0xffffffffu, v.y, o);
    }
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint2 s = make_uint2(0, 0);
        for (int w = 0; w < SCAN_THREADS / 32; w++) {
            s.x += sm[w].x;
            s.y += sm[w].y;
        }
        tile_sums[blockIdx.x] = s;
    }
}

// phase 2: one block scans the tile totals in place (exclusive) and writes the grand totals
static __global__ void __launch_bounds__(1024) k_scan_tiles(uint2 *__restrict__ tile_sums, uint32_t ntiles, uint32_t *__restrict__ totals)
{
    __shared__ uint2 warp_tot[32];
    __shared__ uint2 carry_sm;
    if (threadIdx.x == 0) carry_sm = make_uint2(0, 0);
    __syncthreads();
    for (uint32_t base = 0; base < ntiles; base += 1024) {
        const uint32_t idx = base + threadIdx.x;
        uint2 v = idx < ntiles ? tile_sums[idx] : make_uint2(0, 0);
        uint2 incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ax = __shfl_up_sync(0xffffffffu, incl.x, o);
            const uint32_t ay = __shfl_up_sync(0xffffffffu, incl.y, o);
            if ((threadIdx.x & 31) >= o) {
                incl.x += ax;
                incl.y += ay;
            }
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint2 w = warp_tot[threadIdx.x];
            uint2 wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t ax = __shfl_up_sync(0xffffffffu, wi.x, o);
                const uint32_t ay = __shfl_up_sync(0xffffffffu, wi.y, o);
                if (threadIdx.x >= o) {
                    wi.x += ax;
                    wi.y += ay;
                }
            }
            warp_tot[threadIdx.x] = make_uint2(wi.x - w.x, wi.y - w.y);  // exclusive warp offsets
        }
        __syncthreads();
        const uint2 wo = warp_tot[threadIdx.x >> 5];
        const uint2 c0 = carry_sm;
        const uint2 excl = make_uint2(c0.x + wo.x + incl.x - v.x, c0.y + wo.y + incl.y - v.y);
        if (idx < ntiles) tile_sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_sm = make_uint2(excl.x + v.x, excl.y + v.y);
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        totals[0] = carry_sm.x;  // total entries
        totals[1] = carry_sm.y;  // total tasks
    }
}

// phase 3: per-tile exclusive scan + tile offset; writes off/cursor (entries) and toff (tasks)
static __global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t *__restrict__ cnt, uint32_t NB, uint32_t L,
                                                             const uint2 *__restrict__ tile_sums, uint32_t *__restrict__ off,
                                                             uint32_t *__restrict__ cursor, uint32_t *__restrict__ toff)
{
    __shared__ uint2 warp_tot[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t cv[SCAN_ITEMS];
    uint2 tsum = make_uint2(0, 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        cv[k] = idx < NB ? cnt[idx] : 0u;
        tsum.x += cv[k];
        tsum.y += tasks_of(cv[k], L);
    }
    uint2 incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t ax = __shfl_up_sync(0xffffffffu, incl.x, o);
        const uint32_t ay = __shfl_up_sync(0xffffffffu, incl.y, o);
        if ((threadIdx.x & 31) >= o) {
            incl.x += ax;
            incl.y += ay;
        }
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint2 wo = make_uint2(0, 0);
    for (int w = 0; w < (int)(threadIdx.x >> 5); w++) {
        wo.x += warp_tot[w].x;
        wo.y += warp_tot[w].y;
    }
    const uint2 t0 = tile_sums[blockIdx.x];
    uint32_t e = t0.x + wo.x + incl.x - tsum.x;
    uint32_t t = t0.y + wo.y + incl.y - tsum.y;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        const uint32_t idx = base + k;
        if (idx < NB) {
            off[idx] = e;
            cursor[idx] = e;
            toff[idx] = t;
        }
        e += cv[k];
        t += tasks_of(cv[k], L);
    }
}

// ------------------------------------------------------------------------------
// accumulation tasks: task t of bucket b covers entries [off[b] + j L, +len)
// ------------------------------------------------------------------------------
// largest b with toff[b] <= t and a non-empty task range (toff is non-decreasing)
__device__ __forceinline__ uint32_t bucket_of_task(const uint32_t *__restrict__ toff, uint32_t NB, uint32_t t)
{
    uint32_t lo = 0, hi = NB;  // find first index with toff[idx] > t
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (toff[mid] <= t) lo = mid + 1;
        else hi = mid;
    }
    return lo - 1;
}

static __global__ void __launch_bounds__(256) k_task_meta(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                                                    const uint32_t *__restrict__ toff, const uint32_t *__restrict__ totals,
                                                    MsmGeom g, uint2 *__restrict__ meta, uint32_t *__restrict__ len_hist)
{
    extern __shared__ uint32_t sh_hist[];  // L + 1 bins
    for (uint32_t k = threadIdx.x; k <= g.L; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
    const uint32_t ntasks = totals[1];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ntasks) {
        const uint32_t b = bucket_of_task(toff, g.NB, t);
        const uint32_t j = t - toff[b];
        const uint32_t rem = cnt[b] - j * g.L;
        const uint32_t len = rem < g.L ? rem : g.L;
        meta[t] = make_uint2(off[b] + j * g.L, len);
        atomicAdd(&sh_hist[len], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= g.L; k += blockDim.x)
        if (sh_hist[k]) atomicAdd(&len_hist[k], sh_hist[k]);
}

// descending-length exclusive scan of the length histogram (L + 1 <= 1024 bins): afterwards
// len_start[len] = number of tasks strictly longer than len; len_cursor is a working copy.
static __global__ void __launch_bounds__(1024) k_len_scan(uint32_t *__restrict__ len_hist, uint32_t *__restrict__ len_cursor, uint32_t L)
{
    __shared__ uint32_t warp_tot[32];
    const uint32_t tid = threadIdx.x;
    const uint32_t v = tid <= L ? len_hist[L - tid] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= (uint32_t)o) incl += a;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        const uint32_t w = warp_tot[tid];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= (uint32_t)o) wi += a;
        }
        warp_tot[tid] = wi - w;
    }
    __syncthreads();
    const uint32_t excl = warp_tot[tid >> 5] + incl - v;
    if (tid <= L) {
        len_hist[L - tid] = excl;
        len_cursor[L - tid] = excl;
    }
}

static __global__ void __launch_bounds__(256) k_task_order(const uint2 *__restrict__ meta, const uint32_t *__restrict__ totals,
                                                     MsmGeom g, uint32_t *__restrict__ len_cursor, uint32_t *__restrict__ order)
{
    extern __shared__ uint32_t sh[];  // [0..L] block histogram, [L+1..2L+1] block base
    uint32_t *sh_hist = sh, *sh_base = sh + (g.L + 1);
    for (uint32_t k = threadIdx.x; k <= g.L; k += blockDim.x) sh_hist[k] = 0;
    __syncthreads();
    const uint32_t ntasks = totals[1];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0, local = 0;
    const bool live = t < ntasks;
    if (live) {
        len = meta[t].y;
        local = atomicAdd(&sh_hist[len], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k <= g.L; k += blockDim.x)
        if (sh_hist[k]) sh_base[k] = atomicAdd(&len_cursor[k], sh_hist[k]);
    __syncthreads();
    if (live) order[sh_base[len] + local] = t;
}

// one thread per task, longest tasks first
template <class F>
__global__ void __launch_bounds__(128) k_accumulate(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                     const uint2 *__restrict__ meta, const uint32_t *__restrict__ order,
                                                     const uint32_t *__restrict__ totals, XYZZ<F> *__restrict__ partial)
{
    const uint32_t ntasks = totals[1];
    const uint32_t gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= ntasks) return;
    const uint32_t t = order[gidx];
    const uint2 m = meta[t];
    const uint32_t *e = entries + m.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    uint32_t cur = e[0];
    Affine<F> p = bases[cur & 0x7fffffffu];
    for (uint32_t k = 0; k < m.y; k++) {
        const uint32_t neg = cur >> 31;
        const Affine<F> q = p;
        if (k + 1 < m.y) {  // prefetch the next point while this one is added
            cur = e[k + 1];
            p = bases[cur & 0x7fffffffu];
        }
        xyzz_madd(acc, q.x, q.y, neg != 0);
    }
    partial[t] = acc;
}

// ------------------------------------------------------------------------------
// warp-level helpers on XYZZ values
// ------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ XYZZ<F> shfl_down_point(const XYZZ<F> &p, int delta)
{
    XYZZ<F> r;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&p);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) dst[i] = __shfl_down_sync(0xffffffffu, src[i], delta);
    return r;
}

template <class F>
__device__ __forceinline__ XYZZ<F> warp_sum_point(XYZZ<F> v)
{
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        const XYZZ<F> other = shfl_down_point(v, o);
        xyzz_add_cold(&v, &other);
    }
    return v;  // lane 0 holds the sum
}

// buckets that were split into several tasks: one warp sums the partials into the first slot
template <class F>
__global__ void __launch_bounds__(128) k_bucket_combine(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                         MsmGeom g, XYZZ<F> *__restrict__ partial)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t b0 = warp * 32; b0 < g.NB; b0 += nwarps * 32) {
        // each lane inspects one bucket; the warp then serves the split ones in turn
        const uint32_t b = b0 + lane;
        const uint32_t nt = b < g.NB ? tasks_of(cnt[b], g.L) : 0;
        uint32_t pending = __ballot_sync(0xffffffffu, nt >= 2);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint32_t bb = b0 + src;
            const uint32_t ntb = __shfl_sync(0xffffffffu, nt, src);
            const uint32_t t0 = toff[bb];
            XYZZ<F> acc = XYZZ<F>::inf();
            for (uint32_t k = lane; k < ntb; k += 32) {
                const XYZZ<F> q = partial[t0 + k];
                xyzz_add_cold(&acc, &q);
            }
            acc = warp_sum_point(acc);
            if (lane == 0) partial[t0] = acc;
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------------------
// window reduction: W_k = sum_{j=1..B} j * bucket[k][j]
// ------------------------------------------------------------------------------
constexpr int RED_THREADS = 128;

// block (x = block within window, y = window); thread = segment of S consecutive buckets
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_window_reduce1(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                                 const XYZZ<F> *__restrict__ partial, MsmGeom g, uint32_t S,
                                                                 XYZZ<F> *__restrict__ block_out)
{
    __shared__ XYZZ<F> sm[RED_THREADS / 32];
    const uint32_t k = blockIdx.y;
    const uint32_t seg = blockIdx.x * RED_THREADS + threadIdx.x;
    const uint32_t j0 = seg * S;  // 0-based bucket index of the segment start; bucket j has weight j + 1
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    if (j0 < g.B) {
        for (int jj = (int)S - 1; jj >= 0; jj--) {
            const uint32_t j = j0 + (uint32_t)jj;
            if (j < g.B) {
                const uint32_t b = k * g.B + j;
                if (cnt[b]) {
                    const XYZZ<F> q = partial[toff[b]];
                    xyzz_add_cold(&run, &q);
                }
            }
            xyzz_add_cold(&acc, &run);
        }
        // acc = sum (jj + 1) bucket[j0 + jj], run = sum bucket[j0 + jj]
        if (j0) {
            const XYZZ<F> wrun = xyzz_mul_small(run, j0);
            xyzz_add_cold(&acc, &wrun);
        }
    }
    acc = warp_sum_point(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        XYZZ<F> v = threadIdx.x < RED_THREADS / 32 ? sm[threadIdx.x] : XYZZ<F>::inf();
        v = warp_sum_point(v);
        if (threadIdx.x == 0) block_out[k * gridDim.x + blockIdx.x] = v;
    }
}

// one warp per window sums the block results
template <class F>
__global__ void __launch_bounds__(32) k_window_reduce2(const XYZZ<F> *__restrict__ block_out, uint32_t nblk,
                                                       XYZZ<F> *__restrict__ window_sums)
{
    const uint32_t k = blockIdx.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = threadIdx.x; i < nblk; i += 32) {
        const XYZZ<F> q = block_out[k * nblk + i];
        xyzz_add_cold(&acc, &q);
    }
    acc = warp_sum_point(acc);
    if (threadIdx.x == 0) window_sums[k] = acc;
}

// ------------------------------------------------------------------------------
// fixed-base tables (get_window_table / windowed_exp / batch_exp, multiexp.tcc:547-646)
// table[o][d] = d * 2^(o w) * g, affine, (0,0) for d = 0; rows = ceil(254 / w)
// ------------------------------------------------------------------------------
// thread = run of M consecutive multiples of row base g_o, written as Jacobian (normalised by k_ingest after)
template <class F>
__global__ void __launch_bounds__(128) k_table_rows(const Affine<F> *__restrict__ row_bases, uint32_t w, uint32_t M,
                                                     Jacobian<F> *__restrict__ table_jac)
{
    const uint32_t o = blockIdx.y;
    const uint32_t run = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t d0 = run * M;
    const uint32_t rowlen = 1u << w;
    if (d0 >= rowlen) return;
    const Affine<F> g = row_bases[o];
    XYZZ<F> acc = xyzz_mul_small(XYZZ<F>::from_affine(g), d0);
    Jacobian<F> *row = table_jac + (size_t)o * rowlen;
    for (uint32_t t = 0; t < M && d0 + t < rowlen; t++) {
        row[d0 + t] = acc.to_jacobian();
        if (!g.is_inf()) xyzz_madd_cold(&acc, &g, false);
    }
}

// out[i] = (coeff * s_i) * g via the affine table; Jacobian out (normalised by k_ingest after)
template <class F>
__global__ void __launch_bounds__(128) k_batch_exp(const Affine<F> *__restrict__ table, uint32_t w, uint32_t rows,
                                                    const Fr *__restrict__ scalars_mont, const Fr *__restrict__ coeff_mont,
                                                    size_t n, Jacobian<F> *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = scalars_mont[i];
    if (coeff_mont) s = Fr::mul(*coeff_mont, s);  // coeff * v[i] (multiexp.tcc:666)
    s = Fr::from_mont(s);
    uint32_t limbs[9];
#pragma unroll
    for (int k = 0; k < 8; k++) limbs[k] = s.l[k];
    limbs[8] = 0;
    const uint32_t mask = (1u << w) - 1u;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t o = 0; o < rows; o++) {
        const uint32_t bit = o * w;
        const uint32_t limb = bit >> 5, sh = bit & 31u;
        const uint64_t two = (uint64_t)limbs[limb] | ((uint64_t)limbs[limb + 1] << 32);
        const uint32_t d = (uint32_t)(two >> sh) & mask;
        if (d) {
            const Affine<F> p = table[((size_t)o << w) + d];
            if (!p.is_inf()) xyzz_madd(acc, p.x, p.y, false);
        }
    }
    out[i] = acc.to_jacobian();
}

// ------------------------------------------------------------------------------
// parity hooks (element-wise)
// ------------------------------------------------------------------------------
template <class T, class Op>
__global__ void k_elementwise(const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ out, size_t n, Op op)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T x = a[i];
    const T y = b ? b[i] : x;
    out[i] = op(x, y);
}

}  // namespace b200
