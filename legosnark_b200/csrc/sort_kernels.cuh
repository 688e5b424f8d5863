// sort_kernels.cuh — Pippenger bucket assignment as a two-level radix partition on the bucket id
// (window, |digit|), staged through shared memory.
//
// Replaces the loop nest of multi_exp_inner<BDLO12> that walks every (window, point) pair and adds
// the point to buckets[id] (LFF/algebra/scalar_multiplication/multiexp.tcc:209-238): the engine first
// groups the W * n (bucket, point) pairs by bucket so that one thread can sum a bucket from a
// contiguous list.  The round-1 sort did this with one global atomic per pair in a histogram pass and
// one more in the scatter pass (k_digit_count / k_digit_scatter in msm_kernels.cuh, kept as the
// fallback for geometries this file does not cover); here the global atomics are per (block, partition):
//
//   k_part_hist     scalars leave Montgomery form once (stored for the next pass); every block counts
//                   its digits per PARTITION (the high bits of the bucket id) in shared memory and
//                   adds its NP counters to the global ones; scalars equal to one are set aside
//   k_part_scan     exclusive scan of the NP partition counts -> partition regions in `items`
//   k_part_scatter  MSD pass: a block recomputes the digits of its tile of scalars, ranks them per
//                   partition in shared memory, stages the (entry, bucket) pairs partition-major in
//                   shared memory, reserves one range per partition with one global atomic and copies
//                   the staged pairs out in runs of consecutive addresses
//   k_part_sort     LSD pass: one block per partition (2^low_bits buckets).  Counting sort in shared
//                   memory: histogram of the low bits, scan, ranks; the bucket-ordered entries are
//                   staged in shared memory and written back linearly (partitions above the staging
//                   capacity scatter straight to their L2-resident region).  Emits cnt[] / off[] of
//                   its buckets and the partition's number of accumulation tasks.
//   k_part_scan2    exclusive scan of the per-partition task counts
//   k_task_emit     one block per partition: task offsets of its buckets, task descriptors, and the
//                   longest-first order over the whole sort (counting sort on the task length: the
//                   32 lanes of a k_accumulate warp get tasks of equal length, the last wave the shortest)
//
// Traffic at n = 2^20, c = 20, W = 13 (13.6 M pairs): 32 MB scalars read + 32 MB written and read back,
// 109 MB of pairs written and read (twice, the second time from L2), 54 MB of entries written: ~300 MB.
#pragma once
#include "msm_kernels.cuh"

namespace b200 {

constexpr uint32_t PART_MAX_NP = 4096;         // partitions (shared histogram of the MSD pass)
constexpr uint32_t PART_MAX_LOW = 12;          // at most 4096 buckets per partition (shared counters of the LSD pass)
constexpr uint32_t PART_STAGE_ITEMS = 11264;   // (entry, bucket) pairs staged per block of the MSD pass: 88 KB (two blocks per SM with 2048 partitions)
constexpr uint32_t FINE_STAGE = 10240;         // entries staged per block of the LSD pass: 40 KB (five blocks per SM)
constexpr uint32_t PART_THREADS = 256;           // k_part_hist
constexpr uint32_t SCAT_THREADS = 1024;          // k_part_scatter: one scalar per thread, two blocks per SM
constexpr uint32_t FINE_THREADS = 384;          // 5 x 384 threads per SM: 2048 partitions run in 2.8 waves (ncu: r2e had 2.3 waves at 3 x 512)

__host__ __device__ __forceinline__ uint32_t part_npad(uint32_t NP) { return (NP + 1u) & ~1u; }
inline size_t part_scatter_smem(const SortGeom &sg) { return (size_t)part_npad(sg.NP) * 3 * 4 + (size_t)PART_STAGE_ITEMS * 8; }
inline size_t part_sort_smem(const SortGeom &sg, uint32_t L) { return ((size_t)2 << sg.low_bits) * 4 + (size_t)FINE_STAGE * 4 + (size_t)(L + 1) * 4; }
inline size_t task_emit_smem(const SortGeom &sg, uint32_t L) { return ((size_t)3 << sg.low_bits) * 4 + (size_t)(2 * (L + 1)) * 4; }

// in-place exclusive scan of sh[0..n) by the whole block; every thread gets the total.  tmp: 33 words of
// shared memory.  Ends with a barrier (sh and tmp[32] are readable by everyone afterwards).
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t *sh, uint32_t n, uint32_t *tmp)
{
    const uint32_t T = blockDim.x, tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t per = (n + T - 1) / T;
    const uint32_t lo = min(tid * per, n), hi = min(lo + per, n);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; i++) sum += sh[i];
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t a = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += a;
    }
    if (lane == 31) tmp[warp] = incl;
    __syncthreads();
    if (tid < 32) {
        const uint32_t w = tid < (T >> 5) ? tmp[tid] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, wi, o);
            if (tid >= (uint32_t)o) wi += a;
        }
        tmp[tid] = wi - w;
        if (tid == 31) tmp[32] = wi;
    }
    __syncthreads();
    uint32_t run = tmp[warp] + incl - sum;
    for (uint32_t i = lo; i < hi; i++) {
        const uint32_t v = sh[i];
        sh[i] = run;
        run += v;
    }
    __syncthreads();
    return tmp[32];
}

// ------------------------------------------------------------------------------
// pass 1: standard-form scalars, partition histogram, ones list
// ------------------------------------------------------------------------------
static __global__ void __launch_bounds__(PART_THREADS) k_part_hist(const Fr *__restrict__ scalars_mont, const uint8_t *__restrict__ flags,
                                                                    size_t n, MsmGeom g, SortGeom sg, Fr *__restrict__ std_out,
                                                                    uint32_t *__restrict__ pcount, uint32_t *__restrict__ ones_idx,
                                                                    uint32_t *__restrict__ ones_cnt)
{
    extern __shared__ uint32_t sh_hist[];
    for (uint32_t p = threadIdx.x; p < sg.NP; p += PART_THREADS) sh_hist[p] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * sg.hist_per_block;
    for (uint32_t j = 0; j < sg.hist_per_block; j += PART_THREADS) {  // uniform trip count: the ballot below is warp-collective
        const size_t i = base + j + threadIdx.x;
        const bool inrange = i < n;
        const bool live = inrange && !flags[i];
        Fr s = Fr::zero();
        if (live) s = Fr::from_mont(scalars_mont[i]);
        bool one = false;
        if (ones_idx) {
            one = live && s.l[0] == 1u && (s.l[1] | s.l[2] | s.l[3] | s.l[4] | s.l[5] | s.l[6] | s.l[7]) == 0u;
            const uint32_t m = __ballot_sync(0xffffffffu, one);
            if (m) {
                const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs(m) - 1u;
                uint32_t b = 0;
                if (lane == leader) b = atomicAdd(ones_cnt, (uint32_t)__popc(m));
                b = __shfl_sync(0xffffffffu, b, leader);
                if (one) ones_idx[b + __popc(m & ((1u << lane) - 1u))] = (uint32_t)i;
            }
        }
        if (!live || one) s = Fr::zero();  // no digits: zero base, zero scalar, or set aside above
        if (inrange) std_out[i] = s;
        for_each_digit(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t) {
            atomicAdd(&sh_hist[(bucket_base(g, k) + (mag - 1)) >> sg.low_bits], 1u);
        });
    }
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < sg.NP; p += PART_THREADS)
        if (sh_hist[p]) atomicAdd(&pcount[p], sh_hist[p]);
}

// exclusive scan of up to PART_MAX_NP counters by one block: start[] and a working copy; total -> *total_out.
// With len_hist: also the descending exclusive scan of the task-length histogram (L + 1 bins), so that
// len_cursor[len] = number of tasks of the whole sort that are longer than len (the longest-first order).
static __global__ void __launch_bounds__(1024) k_part_scan(const uint32_t *__restrict__ count, uint32_t NP, uint32_t *__restrict__ start,
                                                            uint32_t *__restrict__ cursor, uint32_t *__restrict__ total_out,
                                                            uint32_t *__restrict__ zero_a, uint32_t *__restrict__ zero_b,
                                                            const uint32_t *__restrict__ len_hist, uint32_t *__restrict__ len_cursor, uint32_t L,
                                                            uint32_t pad, uint32_t align_mask, uint32_t *__restrict__ raw_total_out)
{
    __shared__ uint32_t sh[PART_MAX_NP];
    __shared__ uint32_t tmp[33];
    __shared__ uint32_t raw;
    if (threadIdx.x == 0) raw = 0;
    __syncthreads();
    // pad / align_mask != 0: room for the per-bucket alignment padding of the partition (SortGeom::align_log)
    uint32_t mine = 0;
    for (uint32_t p = threadIdx.x; p < NP; p += 1024) {
        const uint32_t c = count[p];
        mine += c;
        sh[p] = (c + pad + align_mask) & ~align_mask;
    }
    if (mine) atomicAdd(&raw, mine);
    __syncthreads();
    if (threadIdx.x == 0 && raw_total_out) *raw_total_out = raw;
    const uint32_t total = block_excl_scan(sh, NP, tmp);
    for (uint32_t p = threadIdx.x; p < NP; p += 1024) {
        start[p] = sh[p];
        if (cursor) cursor[p] = sh[p];
    }
    if (threadIdx.x == 0) {
        *total_out = total;
        if (zero_a) *zero_a = 0;
        if (zero_b) *zero_b = 0;
    }
    if (len_hist) {
        __syncthreads();
        for (uint32_t k = threadIdx.x; k <= L; k += 1024) sh[k] = len_hist[L - k];
        __syncthreads();
        block_excl_scan(sh, L + 1, tmp);
        for (uint32_t k = threadIdx.x; k <= L; k += 1024) len_cursor[L - k] = sh[k];
    }
}

// ------------------------------------------------------------------------------
// pass 2 (MSD): partition-major (entry, bucket) pairs
// ------------------------------------------------------------------------------
static __global__ void __launch_bounds__(SCAT_THREADS) k_part_scatter(const Fr *__restrict__ std_scalars, size_t n, MsmGeom g, SortGeom sg,
                                                                       uint32_t *__restrict__ pcursor, uint2 *__restrict__ items)
{
    extern __shared__ uint32_t sh[];
    __shared__ uint32_t tmp[33];
    const uint32_t NPp = part_npad(sg.NP);
    uint32_t *hist = sh, *loc = sh + NPp, *gdelta = sh + 2 * NPp;
    uint2 *stage = reinterpret_cast<uint2 *>(sh + 3 * NPp);
    for (uint32_t p = threadIdx.x; p < sg.NP; p += SCAT_THREADS) hist[p] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * sg.tile;
    const uint32_t cnt_t = (uint32_t)min((size_t)sg.tile, n - base);  // tile <= SCAT_THREADS: one scalar per thread, kept in registers
    const bool live = threadIdx.x < cnt_t;
    Fr s = Fr::zero();
    if (live) s = std_scalars[base + threadIdx.x];
    for_each_digit(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t) {
        atomicAdd(&hist[(bucket_base(g, k) + (mag - 1)) >> sg.low_bits], 1u);
    });
    __syncthreads();
    for (uint32_t p = threadIdx.x; p < sg.NP; p += SCAT_THREADS) loc[p] = hist[p];
    __syncthreads();
    const uint32_t total = block_excl_scan(loc, sg.NP, tmp);
    for (uint32_t p = threadIdx.x; p < sg.NP; p += SCAT_THREADS) {
        const uint32_t h = hist[p];
        if (h) gdelta[p] = atomicAdd(&pcursor[p], h) - loc[p];
        hist[p] = 0;  // becomes the rank cursor of the second pass
    }
    __syncthreads();
    {
        const uint32_t i = (uint32_t)(base + threadIdx.x);
        for_each_digit(s, g.c, g.W, [&](uint32_t k, uint32_t mag, uint32_t neg) {
            const uint32_t bucket = bucket_base(g, k) + (mag - 1);
            const uint32_t p = bucket >> sg.low_bits;
            const uint32_t pbase = g.pre_stride ? k * g.pre_stride + g.pre_off : 0u;  // level k of a precomputed key
            const uint32_t r = atomicAdd(&hist[p], 1u);
            stage[loc[p] + r] = make_uint2((pbase + i) | (neg << 31), bucket);
        });
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += SCAT_THREADS) {
        const uint2 it = stage[j];
        items[gdelta[it.y >> sg.low_bits] + j] = it;
    }
}

// counter update of a warp on shared counters; same-bucket pile-ups (equal scalars, the carry bucket of
// small scalars) are aggregated with match_any like warp_bucket_add does for global counters
__device__ __forceinline__ uint32_t warp_shared_add(uint32_t *counters, uint32_t bucket)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t nb = __shfl_down_sync(0xffffffffu, bucket, 1);
    const uint32_t dup = __ballot_sync(0xffffffffu, bucket != NO_BUCKET && lane < 31u && bucket == nb);
    if (__popc(dup) < 4) return bucket != NO_BUCKET ? atomicAdd(&counters[bucket], 1u) : 0u;
    const uint32_t peers = __match_any_sync(0xffffffffu, bucket);
    const uint32_t leader = (uint32_t)__ffs(peers) - 1u;
    uint32_t base = 0;
    if (bucket != NO_BUCKET && lane == leader) base = atomicAdd(&counters[bucket], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

// ------------------------------------------------------------------------------
// pass 3 (LSD): counting sort of one partition in shared memory
// ------------------------------------------------------------------------------
constexpr int FINE_UNROLL = 4;  // loads in flight per thread in the two sweeps (the sweeps are latency-bound)

static __global__ void __launch_bounds__(FINE_THREADS) k_part_sort(const uint2 *__restrict__ items, const uint32_t *__restrict__ pstart,
                                                                    const uint32_t *__restrict__ pcount, MsmGeom g, SortGeom sg,
                                                                    uint32_t *__restrict__ cnt, uint32_t *__restrict__ off,
                                                                    uint32_t *__restrict__ entries, uint32_t *__restrict__ ptasks,
                                                                    uint32_t *__restrict__ len_hist)
{
    extern __shared__ uint32_t sh[];
    __shared__ uint32_t tmp[33];
    __shared__ uint32_t sh_tasks;
    const uint32_t PB = 1u << sg.low_bits, L = g.L;
    uint32_t *cnt_sh = sh, *cur_sh = sh + PB, *stage = sh + 2 * PB, *lh = stage + FINE_STAGE;
    const uint32_t p = blockIdx.x, b0 = p << sg.low_bits;
    const uint32_t nb = min(PB, g.NB - b0);
    const uint32_t ps = pstart[p], pc = pcount[p];
    const uint32_t amask = (1u << sg.align_log) - 1u;
    for (uint32_t b = threadIdx.x; b < PB; b += FINE_THREADS) cnt_sh[b] = 0;
    for (uint32_t k = threadIdx.x; k <= L; k += FINE_THREADS) lh[k] = 0;
    if (threadIdx.x == 0) sh_tasks = 0;
    __syncthreads();
    for (uint32_t basej = 0; basej < pc; basej += FINE_UNROLL * FINE_THREADS) {  // uniform trip count: warp-collective update
        uint32_t bk[FINE_UNROLL];
#pragma unroll
        for (int u = 0; u < FINE_UNROLL; u++) {
            const uint32_t j = basej + u * FINE_THREADS + threadIdx.x;
            bk[u] = j < pc ? items[ps + j].y - b0 : NO_BUCKET;
        }
#pragma unroll
        for (int u = 0; u < FINE_UNROLL; u++) warp_shared_add(cnt_sh, bk[u]);
    }
    __syncthreads();
    uint32_t t = 0;
    for (uint32_t b = threadIdx.x; b < PB; b += FINE_THREADS) {
        const uint32_t c = cnt_sh[b];
        cur_sh[b] = (c + amask) & ~amask;  // slots of the bucket: its entries, padded to the alignment
        if (b < nb) {
            cnt[b0 + b] = c;
            t += tasks_of(c, L);
            if (c) {  // task lengths of this bucket: c / L full tasks and the remainder
                const uint32_t full = c / L, rem = c - full * L;
                if (full) atomicAdd(&lh[L], full);
                if (rem) atomicAdd(&lh[rem], 1u);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31u) == 0 && t) atomicAdd(&sh_tasks, t);
    __syncthreads();
    const uint32_t slots = block_excl_scan(cur_sh, PB, tmp);  // == pc without alignment
    for (uint32_t b = threadIdx.x; b < nb; b += FINE_THREADS) off[b0 + b] = ps + cur_sh[b];
    for (uint32_t k = threadIdx.x; k <= L; k += FINE_THREADS)
        if (lh[k]) atomicAdd(&len_hist[k], lh[k]);
    if (threadIdx.x == 0) ptasks[p] = sh_tasks;
    __syncthreads();
    const bool staged = slots <= FINE_STAGE;
    if (amask) {  // padding slots read as ENTRY_PAD; the real entries overwrite their slots after the barrier
        for (uint32_t j = threadIdx.x; j < slots; j += FINE_THREADS) {
            if (staged) stage[j] = ENTRY_PAD;
            else entries[ps + j] = ENTRY_PAD;
        }
        // the unused end of the partition's region (it was sized for the worst-case padding, k_part_scan)
        const uint32_t alloc = (pc + (amask << sg.low_bits) + amask) & ~amask;
        for (uint32_t j = slots + threadIdx.x; j < alloc; j += FINE_THREADS) entries[ps + j] = ENTRY_PAD;
        __syncthreads();
    }
    for (uint32_t basej = 0; basej < pc; basej += FINE_UNROLL * FINE_THREADS) {
        uint2 it[FINE_UNROLL];
#pragma unroll
        for (int u = 0; u < FINE_UNROLL; u++) {
            const uint32_t j = basej + u * FINE_THREADS + threadIdx.x;
            it[u] = j < pc ? items[ps + j] : make_uint2(0u, NO_BUCKET);
        }
#pragma unroll
        for (int u = 0; u < FINE_UNROLL; u++) {
            const bool live = it[u].y != NO_BUCKET;
            const uint32_t pos = warp_shared_add(cur_sh, live ? it[u].y - b0 : NO_BUCKET);
            if (live) {
                if (staged) stage[pos] = it[u].x;
                else entries[ps + pos] = it[u].x;
            }
        }
    }
    if (staged) {
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < slots; j += FINE_THREADS) entries[ps + j] = stage[j];
    }
}

// ------------------------------------------------------------------------------
// accumulation tasks of one partition: toff[], meta[], longest-first order[], split / big bucket lists
// ------------------------------------------------------------------------------
static __global__ void __launch_bounds__(FINE_THREADS) k_task_emit(const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ off,
                                                                    const uint32_t *__restrict__ ptstart, MsmGeom g, SortGeom sg,
                                                                    uint32_t *__restrict__ toff, uint2 *__restrict__ meta,
                                                                    uint32_t *__restrict__ order, uint32_t *__restrict__ totals,
                                                                    uint32_t *__restrict__ split, uint32_t *__restrict__ big,
                                                                    uint32_t *__restrict__ len_cursor, uint32_t *__restrict__ task_bucket)
{
    extern __shared__ uint32_t sh[];
    __shared__ uint32_t tmp[33];
    const uint32_t PB = 1u << sg.low_bits, L = g.L;
    uint32_t *tk = sh, *cs = sh + PB, *os = sh + 2 * PB, *lh = sh + 3 * PB, *lstart = lh + (L + 1);
    const uint32_t p = blockIdx.x, b0 = p << sg.low_bits;
    const uint32_t nb = min(PB, g.NB - b0);
    const uint32_t t0 = ptstart[p];
    for (uint32_t b = threadIdx.x; b < PB; b += FINE_THREADS) {
        const uint32_t c = b < nb ? cnt[b0 + b] : 0u;
        cs[b] = c;
        os[b] = b < nb ? off[b0 + b] : 0u;
        tk[b] = tasks_of(c, L);
    }
    for (uint32_t k = threadIdx.x; k <= L; k += FINE_THREADS) lh[k] = 0;
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += FINE_THREADS) {
        const uint32_t c = cs[b];
        if (!c) continue;
        const uint32_t full = c / L, rem = c - full * L;
        if (full) atomicAdd(&lh[L], full);
        if (rem) atomicAdd(&lh[rem], 1u);
        if (c > L) {  // bucket spans several tasks: k_bucket_combine / k_big_combine sum its partials
            if (tk[b] > BIG_TASKS) big[atomicAdd(&totals[4], 1u)] = b0 + b;
            else split[atomicAdd(&totals[2], 1u)] = b0 + b;
        }
    }
    __syncthreads();
    const uint32_t ntask = block_excl_scan(tk, PB, tmp);
    for (uint32_t b = threadIdx.x; b < nb; b += FINE_THREADS) toff[b0 + b] = t0 + tk[b];
    // longest first over the WHOLE sort (k_accumulate's last wave then holds the shortest tasks): this block
    // reserves, per task length, a range behind the tasks of that length other blocks already placed
    // (len_cursor starts at the number of longer tasks, k_part_scan); lh[len] becomes the block's rank cursor
    for (uint32_t k = threadIdx.x; k <= L; k += FINE_THREADS) {
        const uint32_t h = lh[k];
        lstart[k] = h ? atomicAdd(&len_cursor[k], h) : 0u;
        lh[k] = 0;
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < ntask; t += FINE_THREADS) {
        uint32_t lo = 0, hi = PB;  // first bucket with tk[b] > t; the one before it owns task t
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (tk[mid] <= t) lo = mid + 1;
            else hi = mid;
        }
        const uint32_t b = lo - 1, j = t - tk[b];
        const uint32_t rem = cs[b] - j * L;
        const uint32_t len = rem < L ? rem : L;
        meta[t0 + t] = make_uint2(os[b] + j * L, len);
        task_bucket[t0 + t] = (b0 + b) | (j == 0 ? 0x80000000u : 0u) | (cs[b] <= L ? 0x40000000u : 0u);  // bit 31: first task; bit 30: only task
        order[lstart[len] + atomicAdd(&lh[len], 1u)] = t0 + t;
    }
}

}  // namespace b200
