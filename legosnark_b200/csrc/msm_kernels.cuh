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
    uint32_t NB;  // Wb * B
    // Precomputed keys (k_key_level): the resident table also holds 2^(c k) P_i for every
    // window k (level k at [k * pre_stride, +count)), so the digit of window k selects a point of
    // level k and ALL windows share one set of B buckets (Wb = 1): no per-window reduction, no
    // Horner doublings, and c can grow (fewer windows = fewer additions) without paying
    // W * 2^(c-1) bucket sums.  Without a precomputed table Wb = W and pre_stride = 0.
    uint32_t Wb;          // windows of buckets: W, or 1 for a precomputed key
    uint32_t pre_stride;  // points per level of the precomputed table (0: plain key)
    uint32_t pre_off;     // first point of this MSM inside its level
    uint32_t ones;        // 1: scalars equal to one were set aside by k_digit_count and summed by k_sum_ones
    // window reduction of a precomputed key (one window): the device returns the per-job sums S_0, T_0 .. T_{jobs-2}
    // and the host forms S_0 + 2^logS sum_b 2^b T_b (a few dozen cheap host operations instead of a serial chain
    // of doublings and a second combine on a lone warp).  red_jobs == 0: the device returns the window sums.
    uint32_t red_jobs, red_logS;
    // marginal-sum form of the same (k_reduce_marginals): red_hb + red_lb = log2(segments); the host gets
    // 1 + red_hb + red_lb points.  red_hb == 0 && red_lb == 0: the bit decomposition over all segments (red_jobs).
    uint32_t red_hb, red_lb;
};
// geometry of the radix-partition sort (sort_kernels.cuh / engine_sort.cu)
struct SortGeom {
    uint32_t low_bits;        // buckets per partition = 1 << low_bits
    uint32_t NP;              // partitions = ceil(NB / 2^low_bits)
    uint32_t tile;            // scalars per block of k_part_scatter (W * tile <= PART_STAGE_ITEMS)
    uint32_t hist_per_block;  // scalars per block of k_part_hist (multiple of 256)
    // batch-affine accumulation (pair_kernels.cuh): every bucket's range of `entries` starts on a multiple of 2^align_log
    // slots and is padded with ENTRY_PAD up to one, so that slots (2q, 2q+1), (4r .. 4r+3) never straddle buckets.
    // 0: dense ranges as in round 1.
    uint32_t align_log;
};
constexpr uint32_t ENTRY_PAD = 0xffffffffu;  // no point index reaches 2^31 - 1 (W n < 2^31 is checked for keys; sign bit set + all ones)
__host__ __device__ __forceinline__ uint32_t result_points(const MsmGeom &g) { return g.red_jobs ? g.red_jobs : g.Wb; }
__host__ __device__ __forceinline__ uint32_t bucket_base(const MsmGeom &g, uint32_t k) { return g.pre_stride ? 0u : k * g.B; }

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

// Precomputed keys: next level of the table, out[i] = 2^c * in[i] (Jacobian; k_ingest turns it
// into the affine level).  One-off work per key: (W - 1) * c doublings per base.
template <class F>
__global__ void __launch_bounds__(128) k_key_level(const Affine<F> *__restrict__ in, uint32_t c, size_t n,
                                                    Jacobian<F> *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> p = XYZZ<F>::from_affine(in[i]);
    if (!p.is_inf())
        for (uint32_t t = 0; t < c; t++) xyzz_dbl_cold(&p);
    out[i] = p.to_jacobian();
}

// ------------------------------------------------------------------------------
// signed-digit recoding
// ------------------------------------------------------------------------------
// The scalar is consumed from the bottom: the window is the low c bits, then the whole 256-bit value moves
// right by c with eight funnel shifts (static register indices; the round-1 form indexed a limb array with a
// runtime index, which put it in local memory and cost ~45 instructions per window -- the sort kernels were
// instruction-bound on it: profiles/r2d ncu capture, sm throughput 78 %).  2 <= c <= 24.
template <class Fn>
__device__ __forceinline__ void for_each_window(const Fr &s, uint32_t c, uint32_t W, Fn fn)
{
    uint32_t l0 = s.l[0], l1 = s.l[1], l2 = s.l[2], l3 = s.l[3], l4 = s.l[4], l5 = s.l[5], l6 = s.l[6], l7 = s.l[7];
    const uint32_t mask = (1u << c) - 1u;
    const uint32_t half = 1u << (c - 1);
    uint32_t carry = 0;
    for (uint32_t k = 0; k < W; k++) {
        const uint32_t d = (l0 & mask) + carry;
        l0 = __funnelshift_r(l0, l1, c);
        l1 = __funnelshift_r(l1, l2, c);
        l2 = __funnelshift_r(l2, l3, c);
        l3 = __funnelshift_r(l3, l4, c);
        l4 = __funnelshift_r(l4, l5, c);
        l5 = __funnelshift_r(l5, l6, c);
        l6 = __funnelshift_r(l6, l7, c);
        l7 >>= c;
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
        fn(k, mag, neg);  // every window, mag == 0: nothing to add
    }
}

template <class Fn>
__device__ __forceinline__ void for_each_digit(const Fr &s, uint32_t c, uint32_t W, Fn fn)
{
    for_each_window(s, c, W, [&](uint32_t k, uint32_t mag, uint32_t neg) {
        if (mag) fn(k, mag, neg);
    });
}

// Histogram / cursor atomics of one warp on (mostly) distinct buckets go out one per lane.  When many
// lanes of the warp hit the SAME bucket -- small scalars put half of all points into the carry bucket
// (k, 1) of the first empty window, equal scalars put everything into one bucket per window -- the
// same-address atomics serialise in L2.  All 32 lanes call this with their bucket (or NO_BUCKET):
// a cheap neighbour test (one shuffle, one ballot; never fires on uniform digits) switches the warp
// to match_any aggregation: one atomic per distinct bucket, ranks by popc.  Returns the lane's slot.
constexpr uint32_t NO_BUCKET = 0xffffffffu;
__device__ __forceinline__ uint32_t warp_bucket_add(uint32_t *__restrict__ counters, uint32_t bucket)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nb = __shfl_down_sync(0xffffffffu, bucket, 1);
    const uint32_t dup = __ballot_sync(0xffffffffu, bucket != NO_BUCKET && lane < 31u && bucket == nb);
    if (__popc(dup) < 2) return bucket != NO_BUCKET ? atomicAdd(&counters[bucket], 1u) : 0u;
    const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
    const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
    uint32_t base = 0;
    if (bucket != NO_BUCKET && lane == leader) base = atomicAdd(&counters[bucket], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

// Pass 1: one thread per point.  The scalar leaves Montgomery form once, its W signed digits are
// counted into the bucket histogram and stored window-major, digits[k * stride + i] = |d| | sign << 31
// (0 = nothing to add), so that pass 2 streams them back coalesced.
// Scalars equal to ONE never enter the sort (ones_idx != nullptr): like the reference's
// multi_exp_with_mixed_addition, which adds those bases directly (multiexp.tcc:455-487; 0/1-heavy
// witness vectors are the common LegoSNARK / Groth16 input), their indices are appended to a list --
// one ballot and one atomic per warp -- and summed by k_sum_ones; in the sort they would all hit
// bucket 1 of window 0 (n same-address atomics, one giant bucket).
static __global__ void __launch_bounds__(256) k_digit_count(const Fr *__restrict__ scalars_mont, const uint8_t *__restrict__ flags,
                                                      size_t n, size_t stride, MsmGeom g, uint32_t *__restrict__ cnt,
                                                      uint32_t *__restrict__ digits, uint32_t *__restrict__ ones_idx,
                                                      uint32_t *__restrict__ ones_cnt)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inrange = i < n;
    const bool live = inrange && !flags[i];
    Fr s = Fr::zero();
    if (live) s = Fr::from_mont(scalars_mont[i]);
    bool one = false;
    if (ones_idx) {  // every lane of the warp takes part in the ballot
        one = live && s.l[0] == 1u && (s.l[1] | s.l[2] | s.l[3] | s.l[4] | s.l[5] | s.l[6] | s.l[7]) == 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, one);
        if (m) {
            const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs(m) - 1u;
            uint32_t base = 0;
            if (lane == leader) base = atomicAdd(ones_cnt, (uint32_t)__popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (one) ones_idx[base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
        }
    }
    // every lane walks all W windows (warp_bucket_add is warp-collective); lanes without a point, with a
    // zero base or with a scalar set aside above carry s = 0: all digits zero
    if (!live || one) s = Fr::zero();
    for_each_window(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t neg) {
        if (inrange) digits[(size_t)k * stride + i] = mag ? (mag | (neg << 31)) : 0u;
        warp_bucket_add(cnt, mag ? bucket_base(g, k) + (mag - 1) : NO_BUCKET);
    });
}

// Pass 2: grid (points / 4, windows), four consecutive digits per thread (one 128-bit load;
// every window's row of `digits` starts 16-byte aligned: row stride = n rounded up to 4).
// Blocks are dispatched window-major, so the scattered 4-byte writes of one window (n * 4 bytes
// of `entries`) and its cursors stay L2-resident instead of thrashing DRAM sectors across all
// W * n entries.
static __global__ void __launch_bounds__(256) k_digit_scatter(const uint32_t *__restrict__ digits, size_t n, size_t stride, MsmGeom g,
                                                        uint32_t *__restrict__ cursor, uint32_t *__restrict__ entries)
{
    const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const uint32_t k = blockIdx.y;
    uint4 v = make_uint4(0, 0, 0, 0);  // lanes past the end stay for the warp-collective cursor update
    if (i0 < n) v = *reinterpret_cast<const uint4 *>(digits + (size_t)k * stride + i0);
    const uint32_t d[4] = {v.x, v.y, v.z, v.w};
    const uint32_t bb = bucket_base(g, k);
    const uint32_t pbase = g.pre_stride ? k * g.pre_stride + g.pre_off : 0u;  // level k of a precomputed key
#pragma unroll
    for (int t = 0; t < 4; t++) {
        const bool live = d[t] && i0 + t < n;
        const uint32_t pos = warp_bucket_add(cursor, live ? bb + ((d[t] & 0x7fffffffu) - 1) : NO_BUCKET);
        if (live) entries[pos] = (pbase + (uint32_t)(i0 + t)) | (d[t] & 0x80000000u);
    }
}

// ------------------------------------------------------------------------------
// exclusive scan over the bucket counts; element = (entries, tasks)
// ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tasks_of(uint32_t cnt, uint32_t L) { return (cnt + L - 1) / L; }
constexpr uint32_t BIG_TASKS = 32;     // buckets with more tasks than this are combined by k_big_combine
constexpr uint32_t BIG_CHUNK = 1024;   // partials one warp sums per pass

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
        totals[2] = 0;           // number of split buckets (filled by k_task_meta)
        totals[4] = 0;           // number of big buckets (more than BIG_TASKS tasks; filled by k_task_meta)
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
                                                    const uint32_t *__restrict__ toff, uint32_t *__restrict__ totals,
                                                    MsmGeom g, uint2 *__restrict__ meta, uint32_t *__restrict__ len_hist,
                                                    uint32_t *__restrict__ split, uint32_t *__restrict__ big,
                                                    uint32_t *__restrict__ task_bucket)
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
        task_bucket[t] = b | (j == 0 ? 0x80000000u : 0u) | (cnt[b] <= g.L ? 0x40000000u : 0u);  // bit 31: first task; bit 30: only task
        atomicAdd(&sh_hist[len], 1u);
        if (j == 0 && cnt[b] > g.L) {  // bucket spans several tasks
            if (tasks_of(cnt[b], g.L) > BIG_TASKS) big[atomicAdd(&totals[4], 1u)] = b;
            else split[atomicAdd(&totals[2], 1u)] = b;
        }
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
// MINB: resident blocks per SM asked of ptxas (G2 only: 3 caps the kernel at 168 registers; knob `g2_blocks`)
template <class F, int MINB = 0, bool PAIRED = false>
__global__ void __launch_bounds__(128, MINB) k_accumulate(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                     const uint2 *__restrict__ meta, const uint32_t *__restrict__ order,
                                                     const uint32_t *__restrict__ totals, XYZZ<F> *__restrict__ partial,
                                                     const uint32_t *__restrict__ task_bucket, const XYZZ<F> *seed,
                                                     XYZZ<F> *dense_out)
{
    const uint32_t ntasks = totals[1];
    const uint32_t gidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (gidx >= ntasks) return;
    const uint32_t t = order[gidx];
    const uint2 m = meta[t];
    const uint32_t *e = entries + m.x;
    XYZZ<F> acc = XYZZ<F>::inf();
    const uint32_t tb = (seed || dense_out) ? task_bucket[t] : 0u;
    if (seed) {  // later chunks of a pipelined MSM: the first task of a bucket continues from the bucket's sum so far
        if (tb >> 31) acc = seed[tb & 0x3fffffffu];
    }
    if (sizeof(F) <= 32) {
        // G1: the next point is loaded into registers while this one is added
        uint32_t cur = e[0];
        Affine<F> p = bases[cur & 0x7fffffffu];
        for (uint32_t k = 0; k < m.y; k++) {
            const uint32_t neg = cur >> 31;
            const Affine<F> q = p;
            if (k + 1 < m.y) {
                cur = e[k + 1];
                p = bases[cur & 0x7fffffffu];
            }
            if constexpr (PAIRED) xyzz_madd_paired(acc, q.x, q.y, neg != 0);
            else xyzz_madd(acc, q.x, q.y, neg != 0);
        }
    } else {
        // G2: accumulator (64 registers) + one point (32) + the product temporaries already fill the
        // register file; a second point in flight spills.  The next point (one 128-byte line) is
        // pulled into L2 instead and loaded when its turn comes.
        uint32_t cur = e[0];
        for (uint32_t k = 0; k < m.y; k++) {
            const uint32_t neg = cur >> 31;
            const Affine<F> q = bases[cur & 0x7fffffffu];
            if (k + 1 < m.y) {
                cur = e[k + 1];
                asm volatile("prefetch.global.L2 [%0];" ::"l"(bases + (cur & 0x7fffffffu)));
            }
            xyzz_madd(acc, q.x, q.y, neg != 0);
        }
    }
    // pipelined MSMs: a bucket that is ONE task writes its new total straight into the dense per-bucket array that lives across
    // the chunks (seed and dense_out are the same array; only this thread touches the bucket), so the fold pass only has to
    // copy the buckets that were split into several tasks
    if (dense_out && (tb & 0x40000000u)) dense_out[tb & 0x3fffffffu] = acc;
    else partial[t] = acc;
}

// ------------------------------------------------------------------------------
// G2: two lanes per task.  One thread per task needs the whole XYZZ<Fq2> accumulator (64 registers), a point (32) and the
// product temporaries: 190 registers, two blocks per SM, and the multiply-add pipe idles on dependency stalls
// (k_accumulate<Fq2>: 80 % of the IMAD bound).  Here lane 2i holds the c0 components and lane 2i+1 the c1 components of
// task i's accumulator and operands ("half" elements): additions are component-wise, and a product
//     (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + (a0 b1 + a1 b0) u
// is ONE fused two-product Montgomery pass per lane (Fq::mul_add, 200 multiply-adds) after the partners swap their halves
// with 16 shuffles: the same 400 multiply-adds as Fq2::mul, split over two lanes, with half the registers per lane.
// All branches are uniform within a pair; shuffles name only the pair in their mask.
// ------------------------------------------------------------------------------
__device__ __forceinline__ Fq pair_swap(const Fq &a, uint32_t pmask)
{
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(pmask, a.l[i], 1);
    return r;
}
// own half of a * b
__device__ __forceinline__ Fq h2_mul(const Fq &a, const Fq &b, uint32_t par, uint32_t pmask)
{
    const Fq pa = pair_swap(a, pmask), pb = pair_swap(b, pmask);
    const Fq x1 = par ? pa : a;            // lane 0: a0 b0 + (-a1) b1;   lane 1: a0 b1 + a1 b0
    const Fq x2 = par ? a : Fq::neg(pa);
    return Fq::mul_add(x1, b, x2, pb);
}
// own half of a^2 (complex squaring, fp2.tcc:111-120): (a0 + a1)(a0 - a1) | 2 a0 a1
__device__ __forceinline__ Fq h2_sqr(const Fq &a, uint32_t par, uint32_t pmask)
{
    const Fq pa = pair_swap(a, pmask);
    const Fq u = par ? Fq::dbl(pa) : Fq::add(a, pa);
    const Fq v = par ? a : Fq::sub(a, pa);
    return Fq::mul(u, v);
}
__device__ __forceinline__ bool h2_is_zero(const Fq &a, uint32_t pmask)
{
    const int z = a.is_zero() ? 1 : 0;
    const int o = __shfl_xor_sync(pmask, z, 1);  // both lanes of the pair always execute the shuffle (no short-circuit)
    return (z & o) != 0;
}
__device__ __forceinline__ Fq2 h2_full(const Fq &a, uint32_t par, uint32_t pmask)
{
    const Fq pa = pair_swap(a, pmask);
    return par ? Fq2{pa, a} : Fq2{a, pa};
}
__device__ __forceinline__ Fq h2_own(const Fq2 &a, uint32_t par) { return par ? a.c1 : a.c0; }

struct HalfXYZZ {
    Fq x, y, zz, zzz;
};

// acc += (x2, +-y2), madd-2008-s on half elements; same case analysis as xyzz_madd
__device__ __forceinline__ void h2_madd(HalfXYZZ &acc, const Fq &x2, const Fq &y2_in, bool negate, uint32_t par, uint32_t pmask)
{
    const Fq y2 = Fq::cneg(y2_in, negate);
    if (h2_is_zero(acc.zz, pmask)) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = acc.zzz = par ? Fq::zero() : Fq::one();
        return;
    }
    const Fq U2 = h2_mul(x2, acc.zz, par, pmask);
    const Fq S2 = h2_mul(y2, acc.zzz, par, pmask);
    const Fq P = Fq::sub(U2, acc.x);
    const Fq R = Fq::sub(S2, acc.y);
    if (h2_is_zero(P, pmask)) {
        if (h2_is_zero(R, pmask)) {  // same point: both lanes double the full point and keep their half (rare)
            const XYZZ<Fq2> d = xyzz_dbl_affine<Fq2>(h2_full(x2, par, pmask), h2_full(y2, par, pmask));
            acc.x = h2_own(d.x, par);
            acc.y = h2_own(d.y, par);
            acc.zz = h2_own(d.zz, par);
            acc.zzz = h2_own(d.zzz, par);
        } else {
            acc.x = acc.y = acc.zz = acc.zzz = Fq::zero();
        }
        return;
    }
    const Fq PP = h2_sqr(P, par, pmask);
    const Fq PPP = h2_mul(P, PP, par, pmask);
    const Fq Q = h2_mul(acc.x, PP, par, pmask);
    const Fq X3 = Fq::sub(Fq::sub(h2_sqr(R, par, pmask), PPP), Fq::dbl(Q));
    acc.y = Fq::sub(h2_mul(R, Fq::sub(Q, X3), par, pmask), h2_mul(acc.y, PPP, par, pmask));
    acc.x = X3;
    acc.zz = h2_mul(acc.zz, PP, par, pmask);
    acc.zzz = h2_mul(acc.zzz, PPP, par, pmask);
}

// k_accumulate for G2 with two lanes per task (64 tasks per 128-thread block)
static __global__ void __launch_bounds__(128, 4) k_accumulate_g2pair(const Affine<Fq2> *__restrict__ bases, const uint32_t *__restrict__ entries,
                                                                   const uint2 *__restrict__ meta, const uint32_t *__restrict__ order,
                                                                   const uint32_t *__restrict__ totals, XYZZ<Fq2> *__restrict__ partial,
                                                                   const uint32_t *__restrict__ task_bucket, const XYZZ<Fq2> *__restrict__ seed)
{
    const uint32_t ntasks = totals[1];
    const uint32_t gidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 1;
    if (gidx >= ntasks) return;  // both lanes of a pair leave together
    const uint32_t lane = threadIdx.x & 31u, par = lane & 1u, pmask = 3u << (lane & ~1u);
    const uint32_t t = order[gidx];
    const uint2 m = meta[t];
    const uint32_t *e = entries + m.x;
    HalfXYZZ acc;
    acc.x = acc.y = acc.zz = acc.zzz = Fq::zero();
    if (seed) {
        const uint32_t tb = task_bucket[t];
        if (tb >> 31) {
            const Fq *sp = reinterpret_cast<const Fq *>(seed + (tb & 0x7fffffffu));  // x.c0 x.c1 y.c0 y.c1 zz.c0 zz.c1 zzz.c0 zzz.c1
            acc.x = sp[par];
            acc.y = sp[2 + par];
            acc.zz = sp[4 + par];
            acc.zzz = sp[6 + par];
        }
    }
    uint32_t cur = e[0];
    const Fq *bp = reinterpret_cast<const Fq *>(bases + (cur & 0x7fffffffu));  // x.c0 x.c1 y.c0 y.c1
    Fq px = bp[par], py = bp[2 + par];
    for (uint32_t k = 0; k < m.y; k++) {
        const uint32_t neg = cur >> 31;
        const Fq qx = px, qy = py;
        if (k + 1 < m.y) {
            cur = e[k + 1];
            bp = reinterpret_cast<const Fq *>(bases + (cur & 0x7fffffffu));
            px = bp[par];
            py = bp[2 + par];
        }
        h2_madd(acc, qx, qy, neg != 0, par, pmask);
    }
    Fq *op = reinterpret_cast<Fq *>(partial + t);
    op[par] = acc.x;
    op[2 + par] = acc.y;
    op[4 + par] = acc.zz;
    op[6 + par] = acc.zzz;
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
__device__ __forceinline__ XYZZ<F> warp_sum_point(XYZZ<F> v, int width = 32)
{
#pragma unroll 1
    for (int o = width >> 1; o > 0; o >>= 1) {
        const XYZZ<F> other = shfl_down_point(v, o);
        // lanes at or above o get their own value back from the shuffle: adding it would send them through the doubling
        // path, which the whole warp then waits for at every level
        if ((int)(threadIdx.x & 31u) < o) xyzz_add_cold(&v, &other);
    }
    return v;  // lane 0 holds the sum
}

// ------------------------------------------------------------------------------
// Quad-cooperative addition.  The window reduction and every tree sum are chains of DEPENDENT point additions on
// warps that have their scheduler to themselves: a lone warp needs 5-7 us for the 14 products of one XYZZ addition,
// one after the other.  Here the four lanes of a quad (lane & ~3 .. lane | 3) hold the SAME accumulator and the SAME
// operand (replicated) and split the products of add-2008-s over four levels:
//     level 1   U1 = X1 ZZ2      U2 = X2 ZZ1      S1 = Y1 ZZZ2      S2 = Y2 ZZZ1        (all-gather; P, R on every lane)
//     level 2   PP = P^2         RR = R^2         ZZ12 = ZZ1 ZZ2    ZZZ12 = ZZZ1 ZZZ2   (PP to all, RR to lane 1)
//     level 3   PPP = P PP       Q = U1 PP        ZZ3 = ZZ12 PP     --                  (PPP to all)
//     level 4   V = S1 PPP       T = R (Q - X3)   --                ZZZ3 = ZZZ12 PPP    (Y3 = T - V on lane 1; all-gather)
// i.e. the depth of 4 products instead of 14, for ~90 shuffles.  Same products as xyzz_add, so the same coordinates
// (bit for bit) on every lane.  All branches are uniform within a quad because the state is replicated.
// ------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ F quad_bcast(const F &v, int src, uint32_t qmask)
{
    F r;
    const uint32_t *a = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(F) / 4); i++) d[i] = __shfl_sync(qmask, a[i], src, 4);
    return r;
}

template <class F>
__device__ __noinline__ void xyzz_add_quad(XYZZ<F> *acc_, const XYZZ<F> *q_)
{
    XYZZ<F> &acc = *acc_;
    const XYZZ<F> &q = *q_;
    if (q.is_inf()) return;
    if (acc.is_inf()) {
        acc = q;
        return;
    }
    const uint32_t ql = threadIdx.x & 3u;
    const uint32_t qmask = 0xfu << ((threadIdx.x & 31u) & ~3u);
    // level 1
    F m1;
    {
        const F a = ql == 0 ? acc.x : ql == 1 ? q.x : ql == 2 ? acc.y : q.y;
        const F b = ql == 0 ? q.zz : ql == 1 ? acc.zz : ql == 2 ? q.zzz : acc.zzz;
        m1 = F::mul(a, b);
    }
    const F U1 = quad_bcast(m1, 0, qmask), U2 = quad_bcast(m1, 1, qmask), S1 = quad_bcast(m1, 2, qmask), S2 = quad_bcast(m1, 3, qmask);
    const F P = F::sub(U2, U1);
    const F R = F::sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) xyzz_dbl_cold(&acc);  // the same point: every lane doubles its copy
        else acc = XYZZ<F>::inf();
        return;
    }
    // level 2: P^2, R^2 (as plain products: one convergent routine for the quad; the value is the same) and ZZ12, ZZZ12
    F m2;
    {
        const F a = ql == 0 ? P : ql == 1 ? R : ql == 2 ? acc.zz : acc.zzz;
        const F b = ql == 0 ? P : ql == 1 ? R : ql == 2 ? q.zz : q.zzz;
        m2 = F::mul(a, b);
    }
    const F PP = quad_bcast(m2, 0, qmask);
    // level 3 (lane 3 has nothing to do: it repeats lane 2's product to stay convergent)
    F m3;
    {
        const F a = ql == 0 ? P : ql == 1 ? U1 : m2;  // lanes 2, 3: ZZ12 resp. ZZZ12 (lane 3's result is not used)
        m3 = F::mul(a, PP);
    }
    const F PPP = quad_bcast(m3, 0, qmask);
    // level 4
    F X3 = F::zero(), m4;
    {
        F a, b;
        if (ql == 1) {  // m2 = RR, m3 = Q
            X3 = F::sub(F::sub(m2, PPP), F::dbl(m3));
            a = R;
            b = F::sub(m3, X3);
        } else {
            a = ql == 0 ? S1 : m2;  // lane 3: ZZZ12; lane 2 repeats a product it does not need
            b = PPP;
        }
        m4 = F::mul(a, b);
    }
    const F V = quad_bcast(m4, 0, qmask);
    const F Y3 = F::sub(m4, V);  // meaningful on lane 1
    acc.x = quad_bcast(X3, 1, qmask);
    acc.y = quad_bcast(Y3, 1, qmask);
    acc.zz = quad_bcast(m3, 2, qmask);
    acc.zzz = quad_bcast(m4, 3, qmask);
}

// sum over the eight quads of a warp (every quad holds its value replicated); quad 0 ends up with the warp's sum
template <class F>
__device__ __forceinline__ XYZZ<F> warp_sum_quads(XYZZ<F> v, int quads = 8)
{
#pragma unroll 1
    for (int o = quads >> 1; o > 0; o >>= 1) {
        const XYZZ<F> other = shfl_down_point(v, o * 4);
        if ((int)((threadIdx.x & 31u) >> 2) < o) xyzz_add_quad(&v, &other);  // the other quads would add a value to itself
    }
    return v;
}
// sum over all quads of a block of 128 threads; the first quad of the block holds the result
template <class F>
__device__ __forceinline__ XYZZ<F> block_sum_quads(XYZZ<F> v, XYZZ<F> *sm)
{
    v = warp_sum_quads(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        const uint32_t qd = threadIdx.x >> 2;
        v = qd < (blockDim.x >> 5) ? sm[qd] : XYZZ<F>::inf();
        v = warp_sum_quads(v, (int)(blockDim.x >> 5));
    }
    return v;
}

// buckets that were split into 2..BIG_TASKS tasks (listed by k_task_meta): their task partials are
// summed into the bucket's first slot.  Four lanes serve one bucket (eight buckets per warp and
// step: the common case is 2-3 partials, e.g. the short top window).  With nothing split the
// kernel returns at once.
template <class F>
__global__ void __launch_bounds__(128) k_bucket_combine(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                         const uint32_t *__restrict__ split, const uint32_t *__restrict__ totals,
                                                         MsmGeom g, XYZZ<F> *__restrict__ partial)
{
    const uint32_t nsplit = totals[2];
    const uint32_t lane = threadIdx.x & 31, sub = lane >> 2, sl = lane & 3;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 8; base < nsplit; base += nwarps * 8) {
        const uint32_t i = base + sub;
        const bool live = i < nsplit;
        const uint32_t b = live ? split[i] : 0u;
        const uint32_t nt = live ? tasks_of(cnt[b], g.L) : 0u;
        const uint32_t t0 = live ? toff[b] : 0u;
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t k = sl; k < nt; k += 4) {
            const XYZZ<F> q = partial[t0 + k];
            xyzz_add_cold(&acc, &q);
        }
        XYZZ<F> other = shfl_down_point(acc, 2);
        if (sl < 2) xyzz_add_cold(&acc, &other);
        other = shfl_down_point(acc, 1);
        if (sl == 0) xyzz_add_cold(&acc, &other);
        if (live && sl == 0) partial[t0] = acc;
    }
}

// Hot buckets (more than BIG_TASKS tasks: a bucket that received a large share of ALL points -- equal
// scalars, the carry bucket of small scalars, an almost empty top window on a precomputed key).  Pass p
// works on the partials at stride BIG_CHUNK^p: every warp of the grid takes chunks of BIG_CHUNK of them
// (any bucket), sums them (lanes stride, then a shuffle tree) and leaves the chunk sum in the chunk's
// first slot; after ceil(log_BIG_CHUNK(tasks)) passes the bucket sum sits in its first slot, where the
// window reduction expects it.  Passes are separate launches (grid-wide dependency).
template <class F>
__global__ void __launch_bounds__(128) k_big_combine(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                      const uint32_t *__restrict__ big, const uint32_t *__restrict__ totals,
                                                      MsmGeom g, uint32_t pass, XYZZ<F> *__restrict__ partial)
{
    const uint32_t nbig = totals[4];
    if (nbig == 0) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    size_t stride = 1;
    for (uint32_t p = 0; p < pass; p++) stride *= BIG_CHUNK;
    uint32_t first = 0;  // global chunk index of this bucket's first chunk: chunks are dealt round-robin over all buckets
    for (uint32_t bi = 0; bi < nbig; bi++) {
        const uint32_t b = big[bi];
        const size_t nt = tasks_of(cnt[b], g.L);
        const size_t count = (nt + stride - 1) / stride;  // partials still standing at this pass
        if (count <= 1) continue;
        const uint32_t chunks = (uint32_t)((count + BIG_CHUNK - 1) / BIG_CHUNK);
        const size_t t0 = toff[b];
        for (uint32_t j = (warp + nwarps - first % nwarps) % nwarps; j < chunks; j += nwarps) {
            const size_t m0 = (size_t)j * BIG_CHUNK, m1 = min(m0 + BIG_CHUNK, count);
            XYZZ<F> acc = XYZZ<F>::inf();
            for (size_t m = m0 + lane; m < m1; m += 32) {
                const XYZZ<F> q = partial[t0 + m * stride];
                xyzz_add_cold(&acc, &q);
            }
            acc = warp_sum_point(acc);
            if (lane == 0) partial[t0 + m0 * stride] = acc;
            __syncwarp();
        }
        first += chunks;
    }
}

// pipelined MSMs (host bases arriving in chunks): after chunk j has been sorted and accumulated,
// its bucket sums (first task slot of every non-empty bucket) go into the dense array that lives
// across chunks; the first chunk initialises it.  Later chunks seed the first task of every bucket
// with dense[b] (k_accumulate's `seed`), so this pass is a copy of the new totals, not a point
// addition per bucket (round 1 added: 3 x 2^19 XYZZ additions per 2^20-point MSM, 0.3 ms).
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fold(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                      const XYZZ<F> *__restrict__ partial, uint32_t NB, bool first, bool seeded,
                                                      XYZZ<F> *__restrict__ dense)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= NB) return;
    if (!cnt[b]) {
        if (first) dense[b] = XYZZ<F>::inf();
        return;
    }
    const XYZZ<F> q = partial[toff[b]];
    if (first || seeded) {  // seeded: the chunk's accumulation started from dense[b] (k_accumulate), so q is the new total
        dense[b] = q;
    } else {
        XYZZ<F> acc = dense[b];
        xyzz_add_cold(&acc, &q);
        dense[b] = acc;
    }
}

// the same for the buckets that were split into several tasks only (lists of k_task_meta / k_task_emit: `split`, totals[2]
// entries; `big`, totals[4]); every other non-empty bucket was written into `dense` by k_accumulate itself
template <class F>
__global__ void __launch_bounds__(128) k_bucket_fold_lists(const uint32_t *__restrict__ toff, const XYZZ<F> *__restrict__ partial,
                                                            const uint32_t *__restrict__ split, const uint32_t *__restrict__ big,
                                                            const uint32_t *__restrict__ totals, XYZZ<F> *__restrict__ dense)
{
    const uint32_t ns = totals[2], nb = totals[4];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < ns + nb; i += gridDim.x * blockDim.x) {
        const uint32_t b = i < ns ? split[i] : big[i - ns];
        dense[b] = partial[toff[b]];
    }
}

// ------------------------------------------------------------------------------
// window reduction: R_k = sum_{j=0..B-1} (j + 1) * bucket[k][j]
// (the running-sum loop of multiexp.tcc:244-278, restructured for parallelism)
//
// level 1  k_reduce_segments: one thread per segment of S = 2^logS consecutive buckets runs
//          the classic running sum inside its segment:
//              run_s = sum_t bucket[S s + t],   acc_s = sum_t (t + 1) bucket[S s + t]
//          so that  R_k = sum_s acc_s + S * sum_s s * run_s.
// level 2  k_reduce_bits: the second term by binary decomposition of s, which needs only
//          plain sums (tree reductions, no serial chain over segments):
//              sum_s s run_s = sum_b 2^b T_b,   T_b = sum_{s : bit b of s set} run_s.
//          Block (job, part, k): job 0 sums the acc_s, job b + 1 sums T_b; each job is shared
//          by RED2_SPLIT blocks, which double their partial sum b + logS times themselves (the
//          doubling is linear); the last block of a window to finish adds all results into R_k.
// ------------------------------------------------------------------------------
constexpr int RED_THREADS = 128;
constexpr int RED2_THREADS = 128;

// `dense` != nullptr: bucket sums come from the dense per-bucket array that k_bucket_fold keeps
// across the chunks of a pipelined MSM (empty buckets hold infinity); otherwise from the task
// partials of the single sort.
template <class F>
__device__ __forceinline__ XYZZ<F> load_bucket_sum(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                   const XYZZ<F> *__restrict__ partial, const XYZZ<F> *__restrict__ dense, uint32_t b)
{
    if (dense) return dense[b];
    if (cnt[b]) return partial[toff[b]];
    return XYZZ<F>::inf();
}

// The two additions of every step are inlined (the accumulators stay in registers; the out-of-line
// adder of round 1 kept them in local memory) and the next bucket's sum is loaded while they run.
template <class F>
__global__ void __launch_bounds__(RED_THREADS) k_reduce_segments(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ toff,
                                                                  const XYZZ<F> *__restrict__ partial, const XYZZ<F> *__restrict__ dense,
                                                                  MsmGeom g, uint32_t logS, XYZZ<F> *__restrict__ seg_run,
                                                                  XYZZ<F> *__restrict__ seg_acc)
{
    const uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;  // global segment index over all windows (blocks of 32..RED_THREADS)
    const uint32_t nseg = g.NB >> logS;
    if (seg >= nseg) return;
    const uint32_t b0 = seg << logS;  // segments never straddle windows: S divides B
    XYZZ<F> run = XYZZ<F>::inf(), acc = XYZZ<F>::inf();
    int t = (1 << logS) - 1;
    XYZZ<F> q = load_bucket_sum<F>(cnt, toff, partial, dense, b0 + (uint32_t)t);
#pragma unroll 1
    for (; t >= 0; t--) {
        const XYZZ<F> cur = q;
        if (t > 0) q = load_bucket_sum<F>(cnt, toff, partial, dense, b0 + (uint32_t)(t - 1));
        if (sizeof(F) <= 32) {
            xyzz_add(run, cur);
            xyzz_add(acc, run);
        } else {
            xyzz_add_cold(&run, &cur);
            xyzz_add_cold(&acc, &run);
        }
    }
    seg_run[seg] = run;
    seg_acc[seg] = acc;
}

// ------------------------------------------------------------------------------
// level 2 for ONE window of buckets (precomputed keys), by marginal sums.  With M = 2^(hb + lb) segments,
// s = hi * 2^lb + lo:
//     sum_s s run_s = 2^lb sum_hi hi R_hi + sum_lo lo C_lo,   R_hi = sum_lo run_(hi,lo),  C_lo = sum_hi run_(hi,lo)
// k_reduce_marginals: one block per output -- the 2^hb row sums R, the 2^lb column sums C and the 2^hb partial
//     sums of acc_s (A_hi = sum_lo acc_(hi,lo)): 3 additions per segment instead of ~9 for the bit
//     decomposition over all M segments;
// k_reduce_marginal_bits: one block per output -- T^R_b = sum over the hi with bit b set of R_hi, T^C_b likewise,
//     and A = sum_hi A_hi: 1 + hb + lb points go to the host, which applies the weights (a few dozen host
//     operations): R = A + 2^logS (2^lb Horner(T^R) + Horner(T^C)).
// ------------------------------------------------------------------------------
template <class F>
__device__ __forceinline__ XYZZ<F> block_sum_point(XYZZ<F> v, XYZZ<F> *sm)
{
    v = warp_sum_point(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : XYZZ<F>::inf();
        v = warp_sum_point(v, (int)(blockDim.x >> 5));  // blockDim.x / 32 is a power of two
    }
    return v;  // thread 0 holds the block's sum
}

// one block per output: these sums are chains of dependent point additions (a lone warp needs ~7 us per
// addition), so every output gets 128 lanes: at most two values per lane, then the 5 + 2 level block tree
constexpr int MARG_THREADS = 128;
template <class F>
__global__ void __launch_bounds__(MARG_THREADS) k_reduce_marginals(const XYZZ<F> *__restrict__ seg_run, const XYZZ<F> *__restrict__ seg_acc,
                                                                   uint32_t hb, uint32_t lb, XYZZ<F> *__restrict__ marg)
{
    __shared__ XYZZ<F> sm[MARG_THREADS / 32];
    const uint32_t H = 1u << hb, Lo = 1u << lb;
    const uint32_t w = blockIdx.x;  // 0..H-1: R_hi; H..H+Lo-1: C_lo; then A_hi
    const XYZZ<F> *src;
    uint32_t count, stride;
    if (w < H) {
        src = seg_run + (size_t)w * Lo;
        count = Lo;
        stride = 1;
    } else if (w < H + Lo) {
        src = seg_run + (w - H);
        count = H;
        stride = Lo;
    } else {
        src = seg_acc + (size_t)(w - H - Lo) * Lo;
        count = Lo;
        stride = 1;
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = threadIdx.x; i < count; i += MARG_THREADS) {
        const XYZZ<F> q = src[(size_t)i * stride];
        xyzz_add_cold(&acc, &q);
    }
    acc = block_sum_point(acc, sm);
    if (threadIdx.x == 0) marg[w] = acc;
}

template <class F>
__global__ void __launch_bounds__(MARG_THREADS) k_reduce_marginal_bits(const XYZZ<F> *__restrict__ marg, uint32_t hb, uint32_t lb,
                                                                       XYZZ<F> *__restrict__ out)
{
    __shared__ XYZZ<F> sm[MARG_THREADS / 32];
    const uint32_t H = 1u << hb, Lo = 1u << lb;
    const uint32_t job = blockIdx.x;  // 0: A; 1..hb: T^R_b; hb+1..hb+lb: T^C_b
    const XYZZ<F> *src;
    uint32_t count, bit = 0;
    bool all = false;
    if (job == 0) {
        src = marg + H + Lo;
        count = H;
        all = true;
    } else if (job <= hb) {
        src = marg;
        count = H >> 1;
        bit = job - 1;
    } else {
        src = marg + H;
        count = Lo >> 1;
        bit = job - 1 - hb;
    }
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = threadIdx.x; i < count; i += MARG_THREADS) {
        const uint32_t s0 = all ? i : (((i >> bit) << (bit + 1)) | (1u << bit) | (i & ((1u << bit) - 1u)));
        const XYZZ<F> q = src[s0];
        xyzz_add_cold(&acc, &q);
    }
    acc = block_sum_point(acc, sm);
    if (threadIdx.x == 0) out[job] = acc;
}

// Sum of the bases whose scalar is one (list built by k_digit_count).  Every thread adds its strided
// share with mixed additions, blocks tree-sum, the last block to finish sums the block results
// into `out` (first == false: adds to what an earlier chunk of a pipelined MSM left there).
constexpr int ONES_THREADS = 128;
template <class F>
__global__ void __launch_bounds__(ONES_THREADS) k_sum_ones(const Affine<F> *__restrict__ bases, const uint32_t *__restrict__ ones_idx,
                                                            const uint32_t *__restrict__ ones_cnt, XYZZ<F> *__restrict__ part,
                                                            uint32_t *__restrict__ done, bool first, XYZZ<F> *__restrict__ out)
{
    __shared__ XYZZ<F> sm[ONES_THREADS / 32];
    __shared__ uint32_t ticket;
    const uint32_t count = *ones_cnt;
    if (count == 0) {
        if (first && blockIdx.x == 0 && threadIdx.x == 0) *out = XYZZ<F>::inf();
        return;
    }
    // no more blocks than the list can feed (the others leave at once and take no ticket)
    const uint32_t nblocks = min(gridDim.x, (count + ONES_THREADS - 1) / ONES_THREADS);
    if (blockIdx.x >= nblocks) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t e = blockIdx.x * ONES_THREADS + threadIdx.x; e < count; e += nblocks * ONES_THREADS) {
        const Affine<F> p = bases[ones_idx[e]];
        xyzz_madd_cold(&acc, &p, false);
    }
    acc = block_sum_point(acc, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = acc;
        __threadfence();
        ticket = atomicAdd(done, 1u);
    }
    __syncthreads();
    if (ticket != nblocks - 1) return;
    __threadfence();
    if (threadIdx.x < 32) {
        XYZZ<F> v = XYZZ<F>::inf();
        for (uint32_t o = threadIdx.x; o < nblocks; o += 32) {
            const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(&part[o]);
            XYZZ<F> q;
            uint32_t *d = reinterpret_cast<uint32_t *>(&q);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = p[i];
            xyzz_add_cold(&v, &q);
        }
        v = warp_sum_point(v);
        if (threadIdx.x == 0) {
            if (!first) {
                const XYZZ<F> prev = *out;
                xyzz_add_cold(&v, &prev);
            }
            *out = v;
            *done = 0;  // ready for the next call
        }
    }
}

// grid ((njobs + 1) * RED2_SPLIT, W): job 0 sums twice as many points as the others and gets
// twice the blocks; all blocks of a launch fit the machine in one wave (serial depth matters
// here, not throughput: a point addition issued by a lone warp takes several microseconds)
constexpr int RED2_SPLIT = 2;  // blocks per job with W windows of buckets; a precomputed key (one window) uses more

template <class F>
__global__ void __launch_bounds__(RED2_THREADS) k_reduce_bits(const XYZZ<F> *__restrict__ seg_run, const XYZZ<F> *__restrict__ seg_acc,
                                                               uint32_t M, uint32_t logS, uint32_t split, uint32_t per_job,
                                                               XYZZ<F> *__restrict__ job_out, uint32_t *__restrict__ done,
                                                               XYZZ<F> *__restrict__ window_sums)
{
    __shared__ XYZZ<F> sm[RED2_THREADS / 32];
    __shared__ uint32_t ticket;
    const uint32_t nout = gridDim.x, k = blockIdx.y;
    const uint32_t job = blockIdx.x < 2 * split ? 0 : blockIdx.x / split - 1;
    const uint32_t nparts = job == 0 ? 2 * split : split;
    const uint32_t part = job == 0 ? blockIdx.x : blockIdx.x % split;
    const XYZZ<F> *src = (job == 0 ? seg_acc : seg_run) + (size_t)k * M;
    // the segments this job sums: all of them (job 0) or those with bit (job - 1) set, enumerated
    // densely so that every thread gets the same share
    const uint32_t count = job == 0 ? M : M >> 1;
    const uint32_t bit = job == 0 ? 0 : job - 1;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = part * RED2_THREADS + threadIdx.x; i < count; i += nparts * RED2_THREADS) {
        const uint32_t s0 = job == 0 ? i : (((i >> bit) << (bit + 1)) | (1u << bit) | (i & ((1u << bit) - 1u)));
        const XYZZ<F> q = src[s0];
        xyzz_add_cold(&acc, &q);
    }
    acc = block_sum_point(acc, sm);
    if (threadIdx.x == 0) {
        if (job > 0 && !per_job)
            for (uint32_t i = 0; i < bit + logS; i++) xyzz_dbl_cold(&acc);
        job_out[(size_t)k * nout + blockIdx.x] = acc;
        __threadfence();
        ticket = atomicAdd(&done[per_job ? job : k], 1u);
    }
    __syncthreads();
    if (per_job) {
        // one window (precomputed key): the last block of each job sums that job's partials; weights on the host
        if (ticket != nparts - 1) return;
        __threadfence();
        const uint32_t first = job == 0 ? 0 : (job + 1) * split;
        XYZZ<F> v = XYZZ<F>::inf();
        if (threadIdx.x < nparts) {
            const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(&job_out[first + threadIdx.x]);
            uint32_t *d = reinterpret_cast<uint32_t *>(&v);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = p[i];
        }
        v = block_sum_point(v, sm);
        if (threadIdx.x == 0) {
            window_sums[job] = v;
            done[job] = 0;
        }
        return;
    }
    if (ticket != nout - 1) return;
    // last block of window k: every block's result is visible; the whole block sums them
    __threadfence();
    XYZZ<F> v = XYZZ<F>::inf();
    for (uint32_t o = threadIdx.x; o < nout; o += RED2_THREADS) {
        const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(&job_out[(size_t)k * nout + o]);
        XYZZ<F> q;
        uint32_t *d = reinterpret_cast<uint32_t *>(&q);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = p[i];
        xyzz_add_cold(&v, &q);
    }
    v = block_sum_point(v, sm);  // sm is free again: all threads passed the barrier above
    if (threadIdx.x == 0) {
        window_sums[k] = v;
        done[k] = 0;  // ready for the next call
    }
}

// k_reduce_bits with one QUAD per partial sum (xyzz_add_quad): same jobs, same block / ticket structure, a third of the
// latency per addition.  A block of 128 lanes is 32 quads; quad j of part p takes the elements p * 32 + j, + nparts * 32, ...
template <class F>
__global__ void __launch_bounds__(RED2_THREADS) k_reduce_bits_quad(const XYZZ<F> *__restrict__ seg_run, const XYZZ<F> *__restrict__ seg_acc,
                                                                    uint32_t M, uint32_t logS, uint32_t split, uint32_t per_job,
                                                                    XYZZ<F> *__restrict__ job_out, uint32_t *__restrict__ done,
                                                                    XYZZ<F> *__restrict__ window_sums)
{
    constexpr uint32_t NQ = RED2_THREADS / 4;
    __shared__ XYZZ<F> sm[RED2_THREADS / 32];
    __shared__ uint32_t ticket;
    const uint32_t nout = gridDim.x, k = blockIdx.y, quad = threadIdx.x >> 2;
    const uint32_t job = blockIdx.x < 2 * split ? 0 : blockIdx.x / split - 1;
    const uint32_t nparts = job == 0 ? 2 * split : split;
    const uint32_t part = job == 0 ? blockIdx.x : blockIdx.x % split;
    const XYZZ<F> *src = (job == 0 ? seg_acc : seg_run) + (size_t)k * M;
    const uint32_t count = job == 0 ? M : M >> 1;
    const uint32_t bit = job == 0 ? 0 : job - 1;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint32_t i = part * NQ + quad; i < count; i += nparts * NQ) {
        const uint32_t s0 = job == 0 ? i : (((i >> bit) << (bit + 1)) | (1u << bit) | (i & ((1u << bit) - 1u)));
        const XYZZ<F> q = src[s0];
        xyzz_add_quad(&acc, &q);
    }
    acc = block_sum_quads(acc, sm);
    if (threadIdx.x == 0) {
        if (job > 0 && !per_job)
            for (uint32_t i = 0; i < bit + logS; i++) xyzz_dbl_cold(&acc);
        job_out[(size_t)k * nout + blockIdx.x] = acc;
        __threadfence();
        ticket = atomicAdd(&done[per_job ? job : k], 1u);
    }
    __syncthreads();
    const uint32_t last = per_job ? nparts - 1 : nout - 1;
    if (ticket != last) return;
    // last block of the job (one window) or of window k: sum the partials, a quad per stride
    __threadfence();
    const uint32_t first = per_job ? (job == 0 ? 0 : (job + 1) * split) : 0;
    const uint32_t total = per_job ? nparts : nout;
    const XYZZ<F> *part_out = job_out + (per_job ? 0 : (size_t)k * nout) + first;
    XYZZ<F> v = XYZZ<F>::inf();
    for (uint32_t o = quad; o < total; o += NQ) {
        const volatile uint32_t *p = reinterpret_cast<const volatile uint32_t *>(&part_out[o]);
        XYZZ<F> q;
        uint32_t *d = reinterpret_cast<uint32_t *>(&q);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = p[i];
        xyzz_add_quad(&v, &q);
    }
    v = block_sum_quads(v, sm);  // sm is free again: all threads passed the barrier above
    if (threadIdx.x == 0) {
        window_sums[per_job ? job : k] = v;
        done[per_job ? job : k] = 0;  // ready for the next call
    }
}

// ------------------------------------------------------------------------------
// small MSMs (n <= SMALL_MAX_N: CPlink's prove is one G1 MSM of 1026 points,
// LS/gadgets/subspace.cc:78-85; sparse-matrix keygen issues thousands of 1-2 term ones,
// LS/utils/sparsemexp.h:62-90).  At these sizes the multi-kernel pipeline is all launch and
// dependency latency, so ONE kernel does the whole job on the bases as they came from the host
// (Jacobian, any Z): block k owns window k; it recodes every scalar, counting-sorts the window's
// digits in shared memory, one warp per bucket sums its points (lanes stride, then a shuffle
// tree), and warp 0 forms sum_j (j + 1) B_j with a suffix scan.  Serial depth ~ n/512 + 13
// point operations; the W window sums go to the host Horner like the big path's.
// ------------------------------------------------------------------------------
constexpr uint32_t SMALL_C = 5;                       // window bits
constexpr uint32_t SMALL_NBK = 1u << (SMALL_C - 1);   // buckets = warps per block
constexpr uint32_t SMALL_THREADS = 32 * SMALL_NBK;
constexpr uint32_t SMALL_W = (255 + SMALL_C - 1) / SMALL_C;
constexpr uint32_t SMALL_MAX_N = 4096;

// the bases come as uploaded from the host (Jacobian, any Z) or from a resident key (affine)
template <class F>
__device__ __forceinline__ bool small_base_is_zero(const Jacobian<F> &p) { return p.z.is_zero(); }
template <class F>
__device__ __forceinline__ bool small_base_is_zero(const Affine<F> &p) { return p.is_inf(); }
template <class F>
__device__ __forceinline__ void small_base_add(XYZZ<F> &acc, Jacobian<F> p, bool neg)
{
    if (neg) p.y = F::neg(p.y);
    if (p.z == F::one()) {
        const Affine<F> a{p.x, p.y};
        xyzz_madd_cold(&acc, &a, false);
    } else {
        const XYZZ<F> q = XYZZ<F>::from_jacobian(p);
        xyzz_add_cold(&acc, &q);
    }
}
template <class F>
__device__ __forceinline__ void small_base_add(XYZZ<F> &acc, const Affine<F> &p, bool neg)
{
    xyzz_madd_cold(&acc, &p, neg);
}

template <class F, class BaseT>
__global__ void __launch_bounds__(SMALL_THREADS) k_msm_small(const BaseT *__restrict__ bases, const Fr *__restrict__ scalars_mont,
                                                              uint32_t n, XYZZ<F> *__restrict__ window_sums)
{
    __shared__ uint16_t sh_idx[SMALL_MAX_N];  // bucket-ordered point indices, bit 15 = negate
    __shared__ int8_t sh_dig[SMALL_MAX_N];    // signed digit of point i in this window
    __shared__ uint32_t sh_cnt[SMALL_NBK], sh_off[SMALL_NBK + 1], sh_cur[SMALL_NBK];
    __shared__ XYZZ<F> sh_bsum[SMALL_NBK];
    const uint32_t k = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < SMALL_NBK) sh_cnt[tid] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SMALL_THREADS) {
        int d = 0;
        if (!small_base_is_zero(bases[i])) {
            const Fr s = Fr::from_mont(scalars_mont[i]);
            for_each_digit(s, SMALL_C, SMALL_W, [&](uint32_t kk, uint32_t mag, uint32_t neg) {
                if (kk == k) d = neg ? -(int)mag : (int)mag;
            });
        }
        sh_dig[i] = (int8_t)d;
        if (d) atomicAdd(&sh_cnt[(d < 0 ? -d : d) - 1], 1u);
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t o = 0;
        for (uint32_t j = 0; j < SMALL_NBK; j++) {
            sh_off[j] = o;
            sh_cur[j] = o;
            o += sh_cnt[j];
        }
        sh_off[SMALL_NBK] = o;
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SMALL_THREADS) {
        const int d = sh_dig[i];
        if (d) {
            const uint32_t pos = atomicAdd(&sh_cur[(d < 0 ? -d : d) - 1], 1u);
            sh_idx[pos] = (uint16_t)(i | (d < 0 ? 0x8000u : 0u));
        }
    }
    __syncthreads();
    // warp = bucket
    {
        XYZZ<F> acc = XYZZ<F>::inf();
        for (uint32_t e = sh_off[warp] + lane; e < sh_off[warp + 1]; e += 32) {
            const uint32_t ix = sh_idx[e];
            small_base_add(acc, bases[ix & 0x7fffu], (ix & 0x8000u) != 0);
        }
        // shuffle tree only as deep as the bucket is full (tiny MSMs: 0-2 entries per bucket)
        const uint32_t m = sh_off[warp + 1] - sh_off[warp];
        int width = 1;
        while (width < 32 && (uint32_t)width < m) width <<= 1;
        acc = warp_sum_point(acc, width);
        if (lane == 0) sh_bsum[warp] = acc;
    }
    __syncthreads();
    if (warp == 0) {
        // suffix sums P_l = sum_{m >= l} B_m, then sum_l P_l = sum_m (m + 1) B_m
        XYZZ<F> v = lane < SMALL_NBK ? sh_bsum[lane] : XYZZ<F>::inf();
#pragma unroll 1
        for (uint32_t o = 1; o < SMALL_NBK; o <<= 1) {
            const XYZZ<F> other = shfl_down_point(v, (int)o);
            if (lane + o < SMALL_NBK) xyzz_add_cold(&v, &other);
        }
        v = warp_sum_point(v, (int)SMALL_NBK);
        if (lane == 0) window_sums[k] = v;
    }
}

// ------------------------------------------------------------------------------
// batched tiny MSMs (b200_msm_batch_*: the per-column multi_exp calls of mtxmultiexp, LS/gadgets/subspace.cc:18-25).
// k_batch_terms: one thread per term, s_i * P_i by a fixed 4-bit window (15 table additions, then 4 doublings + at
// most one addition per window; the table lives in local memory, the chain is latency-bound and all terms run
// side by side).  k_batch_sums: one thread per MSM adds its terms.
// ------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(64) k_batch_terms(const Jacobian<F> *__restrict__ bases, const Fr *__restrict__ scalars_mont, size_t n,
                                                     XYZZ<F> *__restrict__ term_out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Jacobian<F> p = bases[i];
    XYZZ<F> r = XYZZ<F>::inf();
    if (!p.z.is_zero()) {
        const Fr s = Fr::from_mont(scalars_mont[i]);
        XYZZ<F> tab[15];  // tab[d - 1] = d * P
        tab[0] = XYZZ<F>::from_jacobian(p);
        for (int d = 1; d < 15; d++) {
            tab[d] = tab[d - 1];
            xyzz_add_cold(&tab[d], &tab[0]);
        }
#pragma unroll 1
        for (int w = 63; w >= 0; w--) {
            if (!r.is_inf())
                for (int t = 0; t < 4; t++) xyzz_dbl_cold(&r);
            const uint32_t d = (s.l[w >> 3] >> ((w & 7) * 4)) & 15u;
            if (d) xyzz_add_cold(&r, &tab[d - 1]);
        }
    }
    term_out[i] = r;
}

template <class F>
__global__ void __launch_bounds__(64) k_batch_sums(const XYZZ<F> *__restrict__ term, const uint64_t *__restrict__ offsets, size_t count,
                                                    Jacobian<F> *__restrict__ out)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    XYZZ<F> acc = XYZZ<F>::inf();
    for (uint64_t i = offsets[j]; i < offsets[j + 1]; i++) {
        const XYZZ<F> q = term[i];
        xyzz_add_cold(&acc, &q);
    }
    out[j] = acc.to_jacobian();
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
