// engine_impl.cuh — per-group (F = Fq for G1, Fq2 for G2) orchestration templates.
// Included by engine_g1.cu and engine_g2.cu only.
#pragma once
#include "engine_common.hpp"
#include "host_arith.hpp"
#include "host_copy.hpp"
#include "msm_kernels.cuh"
#include "pair_kernels.cuh"
#include "test_ops.cuh"

namespace b200 {
namespace eng {

// radix-partition bucket sort (engine_sort.cu)
bool partition_geometry(const MsmGeom &g, size_t n, int sms, SortGeom *out, uint32_t align_log);
void enqueue_partition_sort(Device &D, cudaStream_t st, const MsmGeom &g, const SortGeom &sg, const uint8_t *d_flags,
                            const Fr *d_scalars, size_t n);

// ------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------
// cost model (ms) shared by the plain and the precomputed-key geometries: bucket sort +
// accumulation are linear in the W * n entries; the window reduction costs 2 XYZZ additions per
// bucket but never less than its serial depth; the last stage is pure latency.
inline double geometry_cost(size_t n, uint32_t c, bool pre)
{
    const double W = std::ceil(255.0 / c), B = (double)(1u << (c - 1));
    const double entries = W * (double)n;
    const double acc = std::max(entries * 1.7e-7, 0.15 + entries * 0.5e-7);
    const double red = std::max((pre ? 1.0 : W) * B * 5.4e-7, 0.16) + 0.25;
    return acc + red;
}

// pre_c != 0: geometry of a precomputed key (window bits fixed when the key was extended)
inline MsmGeom choose_geometry(size_t n, uint32_t pre_c = 0, uint32_t pre_stride = 0, uint32_t pre_off = 0, bool wide_field = false)
{
    MsmGeom g;
    uint32_t best_c = 4;
    if (pre_c) {
        best_c = pre_c;
    } else if (g_tune_c > 0) {
        best_c = (uint32_t)std::min(std::max(g_tune_c, 2), 24);
    } else {
        // fitted to B200 measurements (profiles/, DESIGN.md "window choice")
        double best = 1e300;
        for (uint32_t c = 4; c <= 22; c++) {
            const double cost = geometry_cost(n, c, false);
            if (cost < best) {
                best = cost;
                best_c = c;
            }
        }
    }
    g.c = best_c;
    g.W = (255 + g.c - 1) / g.c;
    g.B = 1u << (g.c - 1);
    g.Wb = pre_c ? 1 : g.W;
    g.pre_stride = pre_c ? pre_stride : 0;
    g.pre_off = pre_c ? pre_off : 0;
    g.ones = g_tune_ones ? 1 : 0;
    g.red_jobs = g.red_logS = g.red_hb = g.red_lb = 0;
    g.NB = g.Wb * g.B;
    if (g_tune_L > 0) {
        g.L = (uint32_t)std::min(std::max(g_tune_L, 1), 1023);
    } else {
        const double avg = (double)n * g.W / (double)g.NB;
        // a precomputed key's top window holds only 254 - (W - 1) c bits: its n digits pile onto the lowest
        // buckets of the shared set (c = 20: 85 extra entries on 12 388 buckets at 2^20).  A task length that
        // covers them keeps those buckets whole (no split + combine pass); beyond 512 they are hot buckets anyway.
        double pile = 0;
        if (pre_c) {
            const int top_bits = 254 - (int)((g.W - 1) * g.c);
            if (top_bits > 0 && top_bits < 31) pile = 1.35 * (double)n / (double)(1u << top_bits);
            if (avg + pile > 400) pile = 0;
        }
        uint32_t L = 32;
        while (L < std::max(2.0 * avg, avg + pile + 4.0 * std::sqrt(avg + pile)) && L < 512) L <<= 1;
        // k_accumulate runs one thread per task: keep several waves of tasks in flight even when
        // few buckets hold many entries each (a precomputed key with a small window)
        // ... unless the buckets themselves are already that many tasks (one task per bucket needs no combine)
        const double want_tasks = 148.0 * 512.0 * 6.0;
        const double cap = (double)n * g.W / want_tasks;
        // G2 keeps 256 threads per SM resident and pays three times as much for every combine: once the buckets alone are a
        // wave and a half of tasks they stay whole (2^18, c = 17: 2.97 -> 2.70 ms, profiles/r4r_stage_task_len.jsonl; G1 at
        // the same size has 0.86 waves of buckets and prefers the split, 1.01 vs 1.10 ms)
        const bool whole_buckets = wide_field && (double)g.NB >= 1.5 * 148.0 * 256.0;
        while (pre_c && !whole_buckets && (double)g.NB < want_tasks && L > 32 && L > cap) L >>= 1;
        g.L = L;
    }
    return g;
}

// window bits for extending a key of n bases (b200_key_precompute_*), from the sweeps in
// profiles/r09d_precompute_sweep_2p*.jsonl and profiles/r2j_precompute_window_sweep.jsonl: c = 17 (255 = 15 x 17: a full top
// window) up to 2^18, 20 beyond.  Round 1 switched to c = 22 from 2^25 on; measured in round 2 that geometry loses at every
// size (2^24: 38.2 vs 35.7 ms; 2^26: 457 vs ~145 ms, profiles/r2p_identity_check.log, r2q_stage.jsonl): its top window holds
// 12 bits, so n / 3097 entries pile onto each of the lowest 3097 buckets.  c = 20 needs 13 n < 2^31 table entries (n <= 2^27).
inline uint32_t choose_precompute_window(size_t n)
{
    if (n < ((size_t)1 << 19)) return 17;
    return (size_t)13 * n < ((size_t)1 << 31) ? 20 : 22;
}
// a (sub-)range of a precomputed key takes the precomputed path when it fills the one big bucket set to
// at least ~4 entries per bucket; shorter ranges run on the plain level-0 bases
inline bool precomputed_pays(size_t n, uint32_t c)
{
    const double W = std::ceil(255.0 / c), B = (double)(1u << (c - 1));
    return (double)n * W >= 4.0 * B;
}

template <class F>
struct HostOf;
template <>
struct HostOf<Fq> {
    typedef host::HFq type;
    static constexpr int group = 0;
    static constexpr size_t jac_limbs = 12;
};
template <>
struct HostOf<Fq2> {
    typedef host::HFq2 type;
    static constexpr int group = 1;
    static constexpr size_t jac_limbs = 24;
};

// ------------------------------------------------------------------------------
// bases: Jacobian (host image, already on the device) -> affine + flags
// ------------------------------------------------------------------------------
template <class F, bool OUT_JAC>
void run_ingest(Device &D, cudaStream_t st, const Jacobian<F> *d_in, void *d_out, uint8_t *d_flags, size_t n)
{
    if (n == 0) return;
    D.prefix.ensure(n * sizeof(F));
    const uint32_t blocks = std::max<uint32_t>(1, std::min<uint32_t>(cdiv(n, 128 * 16), (uint32_t)D.sms * 8));
    LAUNCH(D, (k_ingest<F, OUT_JAC>), blocks, 128, 0, st, d_in, d_out, d_flags, D.prefix.as<F>(), n);
}

inline void fill_stats(const Device &D, size_t n, const MsmGeom &g, const uint32_t *totals, double finalize_us, double h2d,
                       double d2h)
{
    g_stats = b200_stats_t{};
    g_stats.n = n;
    g_stats.window_bits = g.c;
    g_stats.num_windows = g.W;
    g_stats.chunk_len = g.L;
    g_stats.kernel_launches = D.launches;
    g_stats.num_entries = totals[0];
    g_stats.num_tasks = totals[1];
    g_stats.host_finalize_us = finalize_us;
    g_stats.h2d_bytes = h2d;
    g_stats.d2h_bytes = d2h;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, D.ev[2], D.ev[3]) == cudaSuccess) g_stats.accumulate_ms = ms;
    if (cudaEventElapsedTime(&ms, D.ev[0], D.ev[1]) == cudaSuccess) g_stats.device_ms = ms;
    if (cudaEventElapsedTime(&ms, D.ev[0], D.ev[2]) == cudaSuccess) g_stats.sort_ms = ms;
}

// ------------------------------------------------------------------------------
// one device, one shard.  The MSM is enqueued in three stages so that a host-buffer call can
// pipeline its upload: plan (geometry + scratch), sort + accumulate (once, or once per chunk
// of points followed by a fold into the dense bucket array), window reduction.
// ------------------------------------------------------------------------------
struct MsmPlan {
    MsmGeom g;
    size_t max_entries, max_tasks, nseg;
    uint32_t ntiles, logS, M, njobs, split;
};

// n: points of the whole shard (decides the geometry); chunk_max: most points one sort handles
template <class F>
MsmPlan plan_msm(Device &D, cudaStream_t st, size_t n, size_t chunk_max, bool dense, const MsmGeom *forced = nullptr)
{
    MsmPlan P;
    const MsmGeom g = P.g = forced ? *forced : choose_geometry(n, 0, 0, 0, sizeof(F) == 64);
    P.max_entries = (size_t)g.W * chunk_max;
    if (P.max_entries >= (1ull << 32)) throw CudaError{"MSM shard too large: W * n must stay below 2^32 entries"};
    P.max_tasks = P.max_entries / g.L + g.NB;
    P.ntiles = cdiv(g.NB, SCAN_TILE);
    // window reduction geometry: segments of S = 2^logS buckets, M segments per window
    // Segment length by a two-term model (ms; fitted to profiles/r02_launches_*): stage 1 runs
    // 2 * 2^logS serial additions per thread in whole waves of resident warps (an XYZZ addition
    // holds the multiply-add pipe of its scheduler for t_add); stage 2 costs ~6 additions per
    // segment on a few hundred blocks.
    {
        const double t_add = (sizeof(F) == 32 ? 3.9e-3 : 11.5e-3);            // ms per warp-wide addition
        const double wps = sizeof(F) == 32 ? 4.0 : 2.0;                       // resident warps per scheduler
        const double cap = D.sms * 4.0 * wps;
        double best = 1e300;
        P.logS = 0;
        for (uint32_t ls = 0; ls <= 8 && (1u << ls) <= g.B; ls++) {
            const double nseg = (double)(g.NB >> ls), nwarps = std::ceil(nseg / 32.0);
            const double waves = std::ceil(nwarps / cap);
            // a lone warp cannot keep the pipe busy (dependent carry chains): ~1.6x slower per addition
            const double share = std::max(1.6, std::min(wps, std::ceil(nwarps / (D.sms * 4.0))));
            const double t1 = waves * (double)(2u << ls) * t_add * share;
            // one window of buckets (precomputed key): stage 2 spreads every job over `split` blocks,
            // each thread sums nseg / (2 * split * 128) values before the block tree
            const double t2 = g.Wb == 1 ? 0.12 + nseg / (2.0 * 16 * RED2_THREADS) * t_add * 1.6
                                        : 0.3 + 2.7e-6 * nseg * (sizeof(F) == 32 ? 1.0 : 3.0);
            if (t1 + t2 < best) {
                best = t1 + t2;
                P.logS = ls;
            }
        }
    }
    if (g_tune_logS >= 0 && (1u << g_tune_logS) <= g.B) P.logS = (uint32_t)g_tune_logS;
    P.M = g.B >> P.logS;
    P.njobs = 1;
    while ((1u << (P.njobs - 1)) < P.M) P.njobs++;  // job 0 + one job per bit of the segment index
    P.nseg = (size_t)g.Wb * P.M;
    // blocks per job of stage 2: W windows already fill the machine with 2; a lone window
    // (precomputed key) spreads each job over 8-16 blocks
    P.split = RED2_SPLIT;
    if (g.Wb == 1) P.split = P.M <= (1u << 14) ? 8 : 16;  // swept: profiles/r15b_reduce_sweep_*.jsonl
    if (g_tune_quads) {
        // quad-cooperative stage 2 (k_reduce_bits_quad): a block is 32 partial sums instead of 128; as many blocks per job as one
        // wave of the machine holds (four blocks of 128 lanes per SM), but at least one element per quad
        const uint32_t wave = (uint32_t)D.sms * 4u, per_window = std::max(1u, wave / std::max(1u, g.Wb));
        P.split = std::max(1u, std::min(per_window / (P.njobs + 1), std::max(1u, (P.M >> 1) / 32u)));
        P.split = std::min(P.split, 64u);
    }
    if (g_tune_split > 0) P.split = (uint32_t)std::min(g_tune_split, 64);
    if (g.Wb == 1 && g_tune_host_horner) {  // one window: per-job sums to the host (MsmGeom::red_jobs); also from the dense array of a chunked call
        P.g.red_jobs = P.njobs;
        P.g.red_logS = P.logS;
        uint32_t lgM = 0;
        while ((1u << lgM) < P.M) lgM++;
        if (g_tune_marginals && lgM >= 2) {  // marginal sums (k_reduce_marginals): 1 + hb + lb points to the host
            P.g.red_hb = (lgM + 1) / 2;
            P.g.red_lb = lgM - P.g.red_hb;
            P.g.red_jobs = 1 + lgM;
        }
    }
    const uint32_t nres = result_points(P.g);

    D.cnt.ensure((size_t)g.NB * 4);
    D.off.ensure((size_t)g.NB * 4);
    D.cursor.ensure((size_t)g.NB * 4);
    D.toff.ensure((size_t)g.NB * 4);
    D.tile_sums.ensure((size_t)P.ntiles * sizeof(uint2));
    D.totals.ensure(32);
    D.entries.ensure((P.max_entries + (size_t)3 * g.NB + 4 * 4096 + 16) * 4);  // + alignment padding of the batch-affine path
    D.digits.ensure((size_t)g.W * ((chunk_max + 3) & ~(size_t)3) * 4);
    D.meta.ensure(P.max_tasks * sizeof(uint2));
    D.order.ensure(P.max_tasks * 4);
    D.task_bucket.ensure(P.max_tasks * 4);
    D.len_hist.ensure((size_t)(g.L + 1) * 4);
    D.len_cursor.ensure((size_t)(g.L + 1) * 4);
    D.partial.ensure(P.max_tasks * sizeof(XYZZ<F>));
    D.seg_run.ensure(P.nseg * sizeof(XYZZ<F>));
    D.seg_acc.ensure(P.nseg * sizeof(XYZZ<F>));
    D.job_out.ensure(std::max((size_t)g.Wb * (P.njobs + 1) * P.split, (size_t)3 << ((P.g.red_hb ? P.g.red_hb : 1))) * sizeof(XYZZ<F>));
    D.split.ensure(std::min<size_t>(g.NB, P.max_tasks) * 4 + 4);
    D.big.ensure(std::min<size_t>(g.NB, P.max_tasks / BIG_TASKS + 1) * 4 + 4);
    if (dense) D.bucket_sum.ensure((size_t)g.NB * sizeof(XYZZ<F>));
    if (D.done.cap < (size_t)std::max<uint32_t>(g.Wb, 64) * 4) {
        D.done.ensure(1024 * 4);
        CK(cudaMemsetAsync(D.done.p, 0, D.done.cap, st));  // k_reduce_bits leaves the counters at zero
    }
    D.window_sums.ensure((size_t)nres * sizeof(XYZZ<F>));
    D.ensure_pinned((size_t)(nres + 2) * sizeof(XYZZ<F>) + 64);
    if (dense && g_tune_overlap_sort) {  // the second set of sort outputs (Device::alt), same sizes
        D.swap_sort_set();
        D.cnt.ensure((size_t)g.NB * 4);
        D.off.ensure((size_t)g.NB * 4);
        D.toff.ensure((size_t)g.NB * 4);
        D.totals.ensure(32);
        D.entries.ensure((P.max_entries + (size_t)3 * g.NB + 4 * 4096 + 16) * 4);
        D.meta.ensure(P.max_tasks * sizeof(uint2));
        D.order.ensure(P.max_tasks * 4);
        D.task_bucket.ensure(P.max_tasks * 4);
        D.split.ensure(std::min<size_t>(g.NB, P.max_tasks) * 4 + 4);
        D.big.ensure(std::min<size_t>(g.NB, P.max_tasks / BIG_TASKS + 1) * 4 + 4);
        if (g.ones) D.ones_idx.ensure(chunk_max * 4);
        D.swap_sort_set();
    }
    if (g.ones) {
        D.ones_idx.ensure(chunk_max * 4);
        D.ones_part.ensure((size_t)D.sms * 2 * sizeof(XYZZ<F>));
        D.ones_sum.ensure(sizeof(XYZZ<F>));
        if (!D.ones_done.p) {
            D.ones_done.ensure(4);
            CK(cudaMemsetAsync(D.ones_done.p, 0, 4, st));  // k_sum_ones leaves the ticket at zero
        }
    }
    return P;
}

// two lanes per task for G2 (k_accumulate_g2pair); the Fq overload is never called (sizeof(F) == 64 guards the call site)
inline void launch_accumulate_g2pair(Device &D, cudaStream_t st, const Affine<Fq2> *d_aff, const uint32_t *entries, const uint2 *meta,
                                     const uint32_t *order, const uint32_t *totals, XYZZ<Fq2> *partial, const uint32_t *tbk,
                                     const XYZZ<Fq2> *seed, size_t max_tasks)
{
    LAUNCH(D, k_accumulate_g2pair, cdiv(max_tasks, 64), 128, 0, st, d_aff, entries, meta, order, totals, partial, tbk, seed);
}
inline void launch_accumulate_g2pair(Device &, cudaStream_t, const Affine<Fq> *, const uint32_t *, const uint2 *, const uint32_t *,
                                     const uint32_t *, XYZZ<Fq> *, const uint32_t *, const XYZZ<Fq> *, size_t)
{
}

// bucket sort of the digits of `n` scalars and accumulation of the bucket (task) sums into D.partial
template <class F>
// dense_direct (pipelined MSMs): k_accumulate writes single-task buckets straight into D.bucket_sum; returns whether it did
// (the optional accumulation kernels do not, the caller then folds every bucket)
// sort_st / sorted (pipelined MSMs, knob overlap_sort): the sort runs on its own stream and `st` waits for the event before
// the accumulation, so that the next chunk's sort can run under this chunk's accumulation
bool enqueue_sort_accumulate(Device &D, cudaStream_t st, const MsmPlan &P, const Affine<F> *d_aff, const uint8_t *d_flags,
                             const Fr *d_scalars, size_t n, bool first_chunk = true, bool seeded = false, bool dense_direct = false,
                             cudaStream_t sort_st = nullptr, cudaEvent_t sorted = nullptr)
{
    const MsmGeom &g = P.g;
    SortGeom sg;
    const uint32_t ba = g_tune_sort ? (uint32_t)g_tune_ba : 0u;  // batch-affine tree levels (needs the aligned bucket ranges of the partition sort)
    const bool part_sort = g_tune_sort && partition_geometry(g, n, D.sms, &sg, ba);
    cudaStream_t main_st = st;
    if (sort_st) st = sort_st;  // everything up to the accumulation is enqueued on the sort stream
    if (!part_sort) {
        CK(cudaMemsetAsync(D.cnt.p, 0, (size_t)g.NB * 4, st));
        CK(cudaMemsetAsync((char *)D.totals.p + 12, 0, 4, st));  // totals[3]: scalars equal to one
        CK(cudaMemsetAsync(D.len_hist.p, 0, (size_t)(g.L + 1) * 4, st));
    }

    uint32_t *cnt = D.cnt.as<uint32_t>(), *off = D.off.as<uint32_t>(), *cursor = D.cursor.as<uint32_t>();
    uint32_t *toff = D.toff.as<uint32_t>(), *totals = D.totals.as<uint32_t>(), *entries = D.entries.as<uint32_t>();
    uint2 *tile_sums = D.tile_sums.as<uint2>(), *meta = D.meta.as<uint2>();
    uint32_t *order = D.order.as<uint32_t>(), *len_hist = D.len_hist.as<uint32_t>(), *len_cursor = D.len_cursor.as<uint32_t>();
    XYZZ<F> *partial = D.partial.as<XYZZ<F>>();

    const size_t max_tasks = (size_t)g.W * n / g.L + g.NB;
    uint32_t *split = D.split.as<uint32_t>(), *big = D.big.as<uint32_t>();
    if (part_sort) {
        enqueue_partition_sort(D, st, g, sg, d_flags, d_scalars, n);
    } else {
    const uint32_t pblocks = cdiv(n, 256);
    const size_t dstride = (n + 3) & ~(size_t)3;
    LAUNCH(D, k_digit_count, pblocks, 256, 0, st, d_scalars, d_flags, n, dstride, g, cnt, D.digits.as<uint32_t>(),
           g.ones ? D.ones_idx.as<uint32_t>() : (uint32_t *)nullptr, totals + 3);
    LAUNCH(D, k_scan_tile_sums, P.ntiles, SCAN_THREADS, 0, st, cnt, g.NB, g.L, tile_sums);
    LAUNCH(D, k_scan_tiles, 1, 1024, 0, st, tile_sums, P.ntiles, totals);
    LAUNCH(D, k_scan_apply, P.ntiles, SCAN_THREADS, 0, st, cnt, g.NB, g.L, tile_sums, off, cursor, toff);
    LAUNCH(D, k_digit_scatter, dim3(cdiv(n, 1024), g.W), 256, 0, st, D.digits.as<uint32_t>(), n, dstride, g, cursor, entries);
    const uint32_t tblocks = cdiv(max_tasks, 256);
    LAUNCH(D, k_task_meta, tblocks, 256, (g.L + 1) * 4, st, cnt, off, toff, totals, g, meta, len_hist, split, big, D.task_bucket.as<uint32_t>());
    LAUNCH(D, k_len_scan, 1, 1024, 0, st, len_hist, len_cursor, g.L);
    LAUNCH(D, k_task_order, tblocks, 256, 2 * (g.L + 1) * 4, st, meta, totals, g, len_cursor, order);
    }
    if (sort_st) {
        CK(cudaEventRecord(sorted, sort_st));
        CK(cudaStreamWaitEvent(main_st, sorted, 0));
    }
    st = main_st;
    const uint32_t *tbk = D.task_bucket.as<uint32_t>();
    const XYZZ<F> *seed = seeded ? D.bucket_sum.as<XYZZ<F>>() : (const XYZZ<F> *)nullptr;  // chunk > 0 of a pipelined MSM
    XYZZ<F> *dout = dense_direct ? D.bucket_sum.as<XYZZ<F>>() : (XYZZ<F> *)nullptr;
    bool wrote_dense = dense_direct;
    CK(cudaEventRecord(D.ev[2], st));
    if (part_sort && ba) {
        // levels of independent affine pair additions with shared inversions, then the XYZZ tail (pair_kernels.cuh)
        const size_t slots = (size_t)g.W * n + (size_t)3 * g.NB + 4 * 4096 + 16;
        D.prefix.ensure((slots / 2 + 1) * sizeof(F));
        D.pa1.ensure((slots / 2 + 1) * sizeof(Affine<F>));
        const uint32_t pblocks = (uint32_t)D.sms * 2;
        LAUNCH(D, (k_pair_add<F, 0>), pblocks, PAIR_THREADS, 0, st, d_aff, (const uint32_t *)entries, (const Affine<F> *)nullptr,
               (const uint32_t *)totals, D.prefix.as<F>(), D.pa1.as<Affine<F>>());
        const Affine<F> *last = D.pa1.as<Affine<F>>();
        if (ba >= 2) {
            D.pa2.ensure((slots / 4 + 1) * sizeof(Affine<F>));
            LAUNCH(D, (k_pair_add<F, 1>), pblocks, PAIR_THREADS, 0, st, d_aff, (const uint32_t *)entries, last, (const uint32_t *)totals,
                   D.prefix.as<F>(), D.pa2.as<Affine<F>>());
            last = D.pa2.as<Affine<F>>();
        }
        LAUNCH(D, (k_accumulate_pa<F>), cdiv(max_tasks, 128), 128, 0, st, last, ba, (const uint2 *)meta, (const uint32_t *)order,
               (const uint32_t *)totals, partial, tbk, seed);
        wrote_dense = false;
    } else if (sizeof(F) == 64 && g_tune_g2pair) {
        launch_accumulate_g2pair(D, st, d_aff, entries, meta, order, totals, partial, tbk, seed, max_tasks);
        wrote_dense = false;
    } else if (sizeof(F) == 32 && g_tune_g1paired) {  // independent products of the mixed addition issued in pairs
        if (g_tune_g1paired == 2)
            LAUNCH(D, (k_accumulate<F, (sizeof(F) == 32 ? 4 : 0), sizeof(F) == 32>), cdiv(max_tasks, 128), 128, 0, st, d_aff, entries, meta, order,
                   totals, partial, tbk, seed, dout);
        else
            LAUNCH(D, (k_accumulate<F, 0, sizeof(F) == 32>), cdiv(max_tasks, 128), 128, 0, st, d_aff, entries, meta, order, totals, partial, tbk,
                   seed, dout);
    } else if (sizeof(F) == 64 && g_tune_g2blocks == 1) {  // ptxas free to use 255 registers (252 used), two blocks per SM
        LAUNCH(D, (k_accumulate<F, (sizeof(F) == 64 ? 1 : 0)>), cdiv(max_tasks, 128), 128, 0, st, d_aff, entries, meta, order, totals, partial,
               tbk, seed, dout);
    } else if (sizeof(F) == 64 && g_tune_g2blocks == 3) {
        LAUNCH(D, (k_accumulate<F, (sizeof(F) == 64 ? 3 : 0)>), cdiv(max_tasks, 128), 128, 0, st, d_aff, entries, meta, order, totals, partial,
               tbk, seed, dout);
    } else {
        LAUNCH(D, (k_accumulate<F>), cdiv(max_tasks, 128), 128, 0, st, d_aff, entries, meta, order, totals, partial, tbk, seed, dout);
    }
    CK(cudaEventRecord(D.ev[3], st));
    LAUNCH(D, (k_bucket_combine<F>), (uint32_t)D.sms * 8, 128, 0, st, cnt, toff, split, totals, g, partial);
    {   // hot buckets: as many passes as the largest possible bucket (all tasks in one) needs; idle ones return at once
        size_t reach = BIG_CHUNK;
        for (uint32_t pass = 0;; pass++) {
            LAUNCH(D, (k_big_combine<F>), (uint32_t)D.sms * 4, 128, 0, st, cnt, toff, big, totals, g, pass, partial);
            if (reach >= max_tasks) break;
            reach *= BIG_CHUNK;
        }
    }
    if (g.ones)  // level 0 of a precomputed key starts at the table's base: this MSM's bases begin pre_off points in
        LAUNCH(D, (k_sum_ones<F>), (uint32_t)D.sms * 2, ONES_THREADS, 0, st, d_aff + (g.pre_stride ? g.pre_off : 0),
               (const uint32_t *)D.ones_idx.as<uint32_t>(), (const uint32_t *)(totals + 3), D.ones_part.as<XYZZ<F>>(),
               D.ones_done.as<uint32_t>(), first_chunk, D.ones_sum.as<XYZZ<F>>());
    return wrote_dense;
}

// direct: k_accumulate already wrote every single-task bucket into the dense array (which was cleared before the first
// chunk); only the buckets that were split into several tasks are copied
template <class F>
void enqueue_fold(Device &D, cudaStream_t st, const MsmPlan &P, bool first, bool direct)
{
    if (direct)
        LAUNCH(D, (k_bucket_fold_lists<F>), (uint32_t)D.sms * 4, 128, 0, st, (const uint32_t *)D.toff.as<uint32_t>(),
               (const XYZZ<F> *)D.partial.as<XYZZ<F>>(), (const uint32_t *)D.split.as<uint32_t>(), (const uint32_t *)D.big.as<uint32_t>(),
               (const uint32_t *)D.totals.as<uint32_t>(), D.bucket_sum.as<XYZZ<F>>());
    else
        LAUNCH(D, (k_bucket_fold<F>), cdiv(P.g.NB, 128), 128, 0, st, D.cnt.as<uint32_t>(), D.toff.as<uint32_t>(), D.partial.as<XYZZ<F>>(),
               P.g.NB, first, !first, D.bucket_sum.as<XYZZ<F>>());
}

// window reduction + D2H of the W window sums (and the entry / task totals of the last sort)
template <class F>
void enqueue_reduce(Device &D, cudaStream_t st, const MsmPlan &P, bool dense)
{
    const MsmGeom &g = P.g;
    XYZZ<F> *seg_run = D.seg_run.as<XYZZ<F>>(), *seg_acc = D.seg_acc.as<XYZZ<F>>(), *wsums = D.window_sums.as<XYZZ<F>>();
    // block size of stage 1: small blocks spread a grid of a few hundred warps evenly over the SMs (1024 blocks of 32 threads on
    // 148 SMs: 7 or 6 per SM; 256 blocks of 128: 2 or 1)
    const uint32_t rb = (uint32_t)std::min(std::max(g_tune_red_block, 32), RED_THREADS) & ~31u;
    LAUNCH(D, (k_reduce_segments<F>), cdiv(P.nseg, rb), rb, 0, st, D.cnt.as<uint32_t>(), D.toff.as<uint32_t>(),
           D.partial.as<XYZZ<F>>(), dense ? D.bucket_sum.as<XYZZ<F>>() : (const XYZZ<F> *)nullptr, g, P.logS, seg_run, seg_acc);
    const uint32_t nres = result_points(g);
    if (g.red_hb) {
        const uint32_t outs = (2u << g.red_hb) + (1u << g.red_lb);
        LAUNCH(D, (k_reduce_marginals<F>), outs, MARG_THREADS, 0, st, (const XYZZ<F> *)seg_run, (const XYZZ<F> *)seg_acc,
               g.red_hb, g.red_lb, D.job_out.as<XYZZ<F>>());
        LAUNCH(D, (k_reduce_marginal_bits<F>), nres, MARG_THREADS, 0, st, (const XYZZ<F> *)D.job_out.as<XYZZ<F>>(), g.red_hb, g.red_lb, wsums);
    } else
    if (g_tune_quads)
        LAUNCH(D, (k_reduce_bits_quad<F>), dim3((P.njobs + 1) * P.split, g.Wb), RED2_THREADS, 0, st, seg_run, seg_acc, P.M, P.logS,
               P.split, g.red_jobs ? 1u : 0u, D.job_out.as<XYZZ<F>>(), D.done.as<uint32_t>(), wsums);
    else
    LAUNCH(D, (k_reduce_bits<F>), dim3((P.njobs + 1) * P.split, g.Wb), RED2_THREADS, 0, st, seg_run, seg_acc, P.M, P.logS,
           P.split, g.red_jobs ? 1u : 0u, D.job_out.as<XYZZ<F>>(), D.done.as<uint32_t>(), wsums);
    CK(cudaMemcpyAsync(D.h_pinned, wsums, (size_t)nres * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync((char *)D.h_pinned + (size_t)nres * sizeof(XYZZ<F>), D.totals.p, 8, cudaMemcpyDeviceToHost, st));
    if (g.ones)
        CK(cudaMemcpyAsync((char *)D.h_pinned + (size_t)(nres + 1) * sizeof(XYZZ<F>), D.ones_sum.p, sizeof(XYZZ<F>),
                           cudaMemcpyDeviceToHost, st));
}

// the whole MSM over bases and scalars that are already on the device; window sums land in D.h_pinned
template <class F>
MsmGeom enqueue_msm(Device &D, cudaStream_t st, const Affine<F> *d_aff, const uint8_t *d_flags, const Fr *d_scalars,
                    size_t n, const MsmGeom *forced = nullptr)
{
    const MsmPlan P = plan_msm<F>(D, st, n, n, false, forced);
    CK(cudaEventRecord(D.ev[0], st));
    enqueue_sort_accumulate<F>(D, st, P, d_aff, d_flags, d_scalars, n);
    enqueue_reduce<F>(D, st, P, false);
    CK(cudaEventRecord(D.ev[1], st));
    return P.g;
}

// Host-buffer MSMs upload 128 B (G1) per point, which takes as long as the arithmetic: the
// shard is cut into index chunks whose H2D copies (copy stream) run under the sort +
// accumulation of the previous chunk (compute stream).
inline size_t choose_chunks(size_t n)
{
    size_t s;
    if (g_tune_chunks > 0) s = (size_t)g_tune_chunks;
    else s = n < (1u << 17) ? 1 : n < (1u << 19) ? 2 : n < (1u << 22) ? 4 : n < (1u << 24) ? 8 : 16;  // profiles/r4w_e2e_chunk_probe.jsonl, r5n (2^24: 49.4 -> 46.7 ms with 16)
    return std::max<size_t>(1, std::min<size_t>(std::min<size_t>(s, MAX_CHUNKS), n));
}

// Index ranges of the upload chunks: equal chunks, or (knob even_chunks = 0) a short first chunk of n/16 points so that less
// of the first upload is exposed.  Measured (profiles/r2w_e2e_chunk_probe.jsonl): 4.76 vs 4.75 ms at 2^20, 14.66 vs 14.16 ms
// at 2^22 -- the cold-key call is bound by the chunked plain-key arithmetic, not by the first upload; equal chunks stay.
inline std::vector<std::pair<size_t, size_t>> pipeline_ranges(size_t n, size_t S)
{
    if (S < 3 || n < ((size_t)1 << 18) || g_tune_even_chunks) return split_range(n, S);
    std::vector<std::pair<size_t, size_t>> r;
    const size_t first = n / 16;
    r.push_back({0, first});
    for (auto &q : split_range(n - first, S - 1)) r.push_back({first + q.first, q.second});
    return r;
}

// The chunk loop of a pipelined MSM: chunk j's inputs are uploaded on the copy stream, ingested (host bases) and sorted on the
// sort stream, accumulated into the dense per-bucket array on `st`.  d_aff / d_flags / d_scalars are the device arrays of the
// WHOLE shard; for a precomputed key d_aff is the table and the chunk's offset goes into the geometry (MsmGeom::pre_off).
template <class F, class Upload>
void enqueue_chunks(Device &D, cudaStream_t st, const MsmPlan &P, const std::vector<std::pair<size_t, size_t>> &ranges, const Upload &upload,
                    bool ingest, const Affine<F> *d_aff, const uint8_t *d_flags, const Fr *d_scalars)
{
    // the dense per-bucket sums: k_accumulate writes whole buckets into it directly (knob dense_direct), so it starts out as
    // all infinity (= all zero bytes) instead of being initialised by the first fold
    const bool want_direct = g_tune_dense_direct != 0;
    if (want_direct) CK(cudaMemsetAsync(D.bucket_sum.p, 0, (size_t)P.g.NB * sizeof(XYZZ<F>), st));
    // ingest + sort of chunk j + 1 on the sort stream under the accumulation of chunk j: the sort kernels wait on memory and
    // shared-memory atomics, the accumulation on the multiply-add pipe.  The batch-affine and lane-pair options share scratch
    // with the ingest and keep everything on one stream.
    const bool overlap = g_tune_overlap_sort && !g_tune_ba && !(sizeof(F) == 64 && g_tune_g2pair);
    if (overlap) CK(cudaStreamWaitEvent(D.sort_stream, D.ev_sync, 0));
    for (size_t j = 0; j < ranges.size(); j++) {
        const size_t lo = ranges[j].first, cnt = ranges[j].second;
        const int set = (int)(j & 1);
        if (overlap && j > 0) D.swap_sort_set();
        cudaStream_t pre = overlap ? D.sort_stream : st;
        upload(D.copy_stream, lo, cnt);
        CK(cudaEventRecord(D.ev_ready[j], D.copy_stream));
        CK(cudaStreamWaitEvent(pre, D.ev_ready[j], 0));
        if (overlap && j >= 2) CK(cudaStreamWaitEvent(pre, D.ev_set_free[set], 0));  // chunk j - 2 read this set
        if (ingest)
            run_ingest<F, false>(D, pre, D.bases_jac.as<Jacobian<F>>() + lo, D.bases_aff.as<Affine<F>>() + lo, D.flags.as<uint8_t>() + lo, cnt);
        MsmPlan Pj = P;
        if (P.g.pre_stride) Pj.g.pre_off = P.g.pre_off + (uint32_t)lo;  // level k of the table: k * pre_stride + pre_off + i
        // with the array cleared, chunk 0 can be "seeded" like the others (it reads infinity)
        const bool direct = enqueue_sort_accumulate<F>(D, st, Pj, P.g.pre_stride ? d_aff : d_aff + lo, d_flags + lo, d_scalars + lo, cnt, j == 0,
                                                       j > 0 || want_direct, want_direct, overlap ? D.sort_stream : (cudaStream_t) nullptr,
                                                       D.ev_sorted[set]);
        enqueue_fold<F>(D, st, P, j == 0 && !want_direct, direct);
        if (overlap) CK(cudaEventRecord(D.ev_set_free[set], st));
    }
}

template <class F>
MsmGeom enqueue_msm_from_host(Device &D, const uint64_t *bases, const uint64_t *scalars, size_t n)
{
    const size_t S = choose_chunks(n);
    D.scalars.ensure(n * sizeof(Fr));
    D.bases_jac.ensure(n * sizeof(Jacobian<F>));
    D.bases_aff.ensure(n * sizeof(Affine<F>));
    D.flags.ensure(n);
    cudaStream_t st = D.stream;
    const auto upload = [&](cudaStream_t cs, size_t lo, size_t cnt) {
        if (!g_scalars_resident)
            h2d(D, D.scalars.as<Fr>() + lo, scalars + lo * 4, cnt * sizeof(Fr), cs);
        h2d(D, D.bases_jac.as<Jacobian<F>>() + lo, bases + lo * HostOf<F>::jac_limbs, cnt * sizeof(Jacobian<F>), cs);
    };
    if (S == 1) {
        upload(st, 0, n);
        run_ingest<F, false>(D, st, D.bases_jac.as<Jacobian<F>>(), D.bases_aff.p, D.flags.as<uint8_t>(), n);
        return enqueue_msm<F>(D, st, D.bases_aff.as<Affine<F>>(), D.flags.as<uint8_t>(), D.scalars.as<Fr>(), n);
    }
    const auto ranges = pipeline_ranges(n, S);
    size_t chunk_max = 0;
    for (auto &r : ranges) chunk_max = std::max(chunk_max, r.second);
    D.prefix.ensure(chunk_max * sizeof(F));
    const MsmPlan P = plan_msm<F>(D, st, n, chunk_max, true);
    CK(cudaEventRecord(D.ev[0], st));
    // order the copy stream after whatever the compute stream still has in flight on these buffers
    CK(cudaEventRecord(D.ev_sync, st));
    CK(cudaStreamWaitEvent(D.copy_stream, D.ev_sync, 0));
    enqueue_chunks<F>(D, st, P, ranges, upload, /*ingest=*/true, D.bases_aff.as<Affine<F>>(), D.flags.as<uint8_t>(), D.scalars.as<Fr>());
    enqueue_reduce<F>(D, st, P, true);
    CK(cudaEventRecord(D.ev[1], st));
    return P.g;
}

// Resident key, HOST scalars (a commitment under a fixed key: CommScheme::commit, LS/prototools/commit.h): the 32 B per point
// of scalars are the only upload, but at 2^20 they still take 0.6 ms in front of a 2.7 ms pipeline.  From 2^18 points on the
// scalars arrive in two (from 2^21: four) index chunks through the same chunk loop as the cold-key call: chunk 1 is uploaded and
// sorted under the accumulation of chunk 0.  `forced`: the precomputed-key geometry (d_aff is then the level table).
inline size_t choose_scalar_chunks(size_t n)
{
    if (g_tune_pinned_chunks > 0) return std::min<size_t>((size_t)g_tune_pinned_chunks, std::min<size_t>(MAX_CHUNKS, n));
    return n < (1u << 18) ? 1 : n < (1u << 21) ? 2 : 4;  // profiles/r5q_resident_scalar_chunks.jsonl
}

template <class F>
MsmGeom enqueue_msm_host_scalars(Device &D, cudaStream_t st, const Affine<F> *d_aff, const uint8_t *d_flags, const uint64_t *scalars, size_t n,
                                 const MsmGeom *forced = nullptr)
{
    const size_t S = st == D.stream ? choose_scalar_chunks(n) : 1;
    D.scalars.ensure(n * sizeof(Fr));
    if (S == 1) {
        h2d(D, D.scalars.p, scalars, n * sizeof(Fr), st);
        return enqueue_msm<F>(D, st, d_aff, d_flags, D.scalars.as<Fr>(), n, forced);
    }
    const auto upload = [&](cudaStream_t cs, size_t lo, size_t cnt) { h2d(D, D.scalars.as<Fr>() + lo, scalars + lo * 4, cnt * sizeof(Fr), cs); };
    const auto ranges = split_range(n, S);
    size_t chunk_max = 0;
    for (auto &r : ranges) chunk_max = std::max(chunk_max, r.second);
    const MsmPlan P = plan_msm<F>(D, st, n, chunk_max, true, forced);
    CK(cudaEventRecord(D.ev[0], st));
    CK(cudaEventRecord(D.ev_sync, st));
    CK(cudaStreamWaitEvent(D.copy_stream, D.ev_sync, 0));
    enqueue_chunks<F>(D, st, P, ranges, upload, /*ingest=*/false, d_aff, d_flags, (const Fr *)D.scalars.as<Fr>());
    enqueue_reduce<F>(D, st, P, true);
    CK(cudaEventRecord(D.ev[1], st));
    return P.g;
}

// Horner over the window sums sitting in D.h_pinned (after the stream has been synchronised)
template <class F>
host::HJac<typename HostOf<F>::type> finalize_windows(const Device &D, const MsmGeom &g)
{
    typedef typename HostOf<F>::type HF;
    typedef host::HJac<HF> J;
    struct HX {
        HF x, y, zz, zzz;
    };
    static_assert(sizeof(HX) == sizeof(XYZZ<F>), "host/device XYZZ images must match");
    const HX *ws = reinterpret_cast<const HX *>(D.h_pinned);
    J acc = J::inf();
    if (g.red_hb) {
        // one window, marginal sums: ws[0] = A, ws[1 + b] = T^R_b (b < hb), ws[1 + hb + b] = T^C_b (b < lb);
        // R = A + 2^logS (2^lb Horner(T^R) + Horner(T^C))
        const auto at = [&](uint32_t i) { return host::jac_from_xyzz(ws[i].x, ws[i].y, ws[i].zz, ws[i].zzz); };
        J hr = J::inf(), hc = J::inf();
        for (int b = (int)g.red_hb - 1; b >= 0; b--) {
            if (!hr.is_inf()) hr = host::jac_dbl(hr);
            hr = host::jac_add(hr, at(1 + (uint32_t)b));
        }
        if (!hr.is_inf())
            for (uint32_t i = 0; i < g.red_lb; i++) hr = host::jac_dbl(hr);
        for (int b = (int)g.red_lb - 1; b >= 0; b--) {
            if (!hc.is_inf()) hc = host::jac_dbl(hc);
            hc = host::jac_add(hc, at(1 + g.red_hb + (uint32_t)b));
        }
        acc = host::jac_add(hr, hc);
        if (!acc.is_inf())
            for (uint32_t i = 0; i < g.red_logS; i++) acc = host::jac_dbl(acc);
        acc = host::jac_add(acc, at(0));
    } else if (g.red_jobs) {
        // one window, per-job sums: ws[0] = S_0 = sum_s acc_s, ws[1 + b] = T_b; R = S_0 + 2^logS sum_b 2^b T_b (Horner)
        for (int b = (int)g.red_jobs - 2; b >= 0; b--) {
            if (!acc.is_inf()) acc = host::jac_dbl(acc);
            acc = host::jac_add(acc, host::jac_from_xyzz(ws[1 + b].x, ws[1 + b].y, ws[1 + b].zz, ws[1 + b].zzz));
        }
        if (!acc.is_inf())
            for (uint32_t i = 0; i < g.red_logS; i++) acc = host::jac_dbl(acc);
        acc = host::jac_add(acc, host::jac_from_xyzz(ws[0].x, ws[0].y, ws[0].zz, ws[0].zzz));
    } else {
        for (int k = (int)g.Wb - 1; k >= 0; k--) {
            if (!acc.is_inf())
                for (uint32_t i = 0; i < g.c; i++) acc = host::jac_dbl(acc);
            acc = host::jac_add(acc, host::jac_from_xyzz(ws[k].x, ws[k].y, ws[k].zz, ws[k].zzz));
        }
    }
    if (g.ones) {  // the bases with scalar one, summed by k_sum_ones (weight 1)
        const HX &o = ws[result_points(g) + 1];
        acc = host::jac_add(acc, host::jac_from_xyzz(o.x, o.y, o.zz, o.zzz));
    }
    return acc;
}

template <class HF>
void write_point(uint64_t *out, const host::HJac<HF> &p)
{
    const host::HJac<HF> nrm = host::jac_normalise(p);
    memcpy(out, &nrm, sizeof nrm);
}

// ------------------------------------------------------------------------------
// MSM entry points
// ------------------------------------------------------------------------------
// n <= SMALL_MAX_N: one kernel on device 0, bases used as uploaded (no affine ingest)
template <class F>
int msm_small_host(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out)
{
    typedef typename HostOf<F>::type HF;
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        MsmGeom g;
        g.c = SMALL_C;
        g.W = SMALL_W;
        g.B = SMALL_NBK;
        g.NB = g.W * g.B;
        g.Wb = g.W;
        g.pre_stride = g.pre_off = g.ones = 0;
        g.red_jobs = g.red_logS = g.red_hb = g.red_lb = 0;
        g.L = 0;
        D.scalars.ensure(n * sizeof(Fr));
        D.bases_jac.ensure(n * sizeof(Jacobian<F>));
        D.window_sums.ensure((size_t)g.W * sizeof(XYZZ<F>));
        D.ensure_pinned((size_t)g.W * sizeof(XYZZ<F>) + 64);
        cudaStream_t st = D.stream;
        if (!g_scalars_resident) CK(cudaMemcpyAsync(D.scalars.p, scalars, n * sizeof(Fr), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(D.bases_jac.p, bases, n * sizeof(Jacobian<F>), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(D.ev[0], st));
        CK(cudaEventRecord(D.ev[2], st));
        LAUNCH(D, (k_msm_small<F, Jacobian<F>>), g.W, SMALL_THREADS, 0, st, (const Jacobian<F> *)D.bases_jac.as<Jacobian<F>>(), D.scalars.as<Fr>(), (uint32_t)n,
               D.window_sums.as<XYZZ<F>>());
        CK(cudaEventRecord(D.ev[3], st));
        CK(cudaMemcpyAsync(D.h_pinned, D.window_sums.p, (size_t)g.W * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(D.ev[1], st));
        CK(cudaStreamSynchronize(st));
        const auto t0 = std::chrono::steady_clock::now();
        write_point<HF>(out, finalize_windows<F>(D, g));
        const auto t1 = std::chrono::steady_clock::now();
        const uint32_t tot[2] = {0, 0};
        fill_stats(D, n, g, tot, std::chrono::duration<double, std::micro>(t1 - t0).count(),
                   (double)n * ((g_scalars_resident ? 0 : sizeof(Fr)) + sizeof(Jacobian<F>)), (double)g.W * sizeof(XYZZ<F>));
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

template <class F>
int msm_host(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out)
{
    typedef typename HostOf<F>::type HF;
    typedef host::HJac<HF> J;
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (!out || (n && (!bases || !scalars))) return fail(B200_ERR_ARG, "null argument");
    if (n == 0) {
        write_point<HF>(out, J::inf());
        return B200_OK;
    }
    if (n <= SMALL_MAX_N && g_tune_c == 0) return msm_small_host<F>(bases, scalars, n, out);
    try {
        const auto ranges = split_range(n, g_devs.size());
        std::vector<J> partials(ranges.size(), J::inf());
        std::vector<MsmGeom> geoms(ranges.size());
        double h2d = 0, d2h = 0;
        for (auto &d : g_devs) d.launches = 0;
        for_each_shard(ranges.size(), [&](size_t si) {
            Device &D = g_devs[si];
            const size_t b = ranges[si].first, m = ranges[si].second;
            CK(cudaSetDevice(D.id));
            geoms[si] = enqueue_msm_from_host<F>(D, bases + b * HostOf<F>::jac_limbs, scalars + b * 4, m);
            CK(cudaStreamSynchronize(D.stream));
        });
        const auto t0 = std::chrono::steady_clock::now();
        J total = J::inf();
        for (size_t si = 0; si < ranges.size(); si++) {
            partials[si] = finalize_windows<F>(g_devs[si], geoms[si]);
            total = host::jac_add(total, partials[si]);
            h2d += (double)ranges[si].second * ((g_scalars_resident ? 0 : sizeof(Fr)) + sizeof(Jacobian<F>));
            d2h += (double)result_points(geoms[si]) * sizeof(XYZZ<F>) + 8;
        }
        write_point<HF>(out, total);
        const auto t1 = std::chrono::steady_clock::now();
        const uint32_t *tot = reinterpret_cast<const uint32_t *>((const char *)g_devs[0].h_pinned +
                                                                 (size_t)result_points(geoms[0]) * sizeof(XYZZ<F>));
        fill_stats(g_devs[0], ranges[0].second, geoms[0], tot, std::chrono::duration<double, std::micro>(t1 - t0).count(), h2d, d2h);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// many small MSMs in one call (include/b200_msm.h: b200_msm_batch_*); device 0
template <class F>
int msm_batch(const uint64_t *bases, const uint64_t *scalars, const uint64_t *offsets, size_t count, uint64_t *out)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called (no CUDA device => no result: there is no CPU fallback)");
    if (count == 0) return B200_OK;
    if (!offsets || !out) return fail(B200_ERR_ARG, "null argument");
    if (offsets[0] != 0) return fail(B200_ERR_ARG, "offsets[0] must be 0");
    for (size_t j = 0; j < count; j++)
        if (offsets[j + 1] < offsets[j]) return fail(B200_ERR_ARG, "offsets must be non-decreasing");
    const size_t n = offsets[count];
    if (n && (!bases || !scalars)) return fail(B200_ERR_ARG, "null argument");
    try {
        Device &D = g_devs[0];
        D.launches = 0;
        CK(cudaSetDevice(D.id));
        cudaStream_t st = D.stream;
        D.scalars.ensure(std::max<size_t>(n, 1) * sizeof(Fr));
        D.bases_jac.ensure(std::max<size_t>(n, 1) * sizeof(Jacobian<F>));
        D.partial.ensure(std::max<size_t>(n, 1) * sizeof(XYZZ<F>));
        D.off.ensure((count + 1) * 8);
        D.out_jac.ensure(count * sizeof(Jacobian<F>));
        D.out_norm.ensure(count * sizeof(Jacobian<F>));
        if (n) {
            h2d(D, D.scalars.p, scalars, n * sizeof(Fr), st);
            h2d(D, D.bases_jac.p, bases, n * sizeof(Jacobian<F>), st);
        }
        h2d(D, D.off.p, offsets, (count + 1) * 8, st);
        if (n)
            LAUNCH(D, (k_batch_terms<F>), cdiv(n, 64), 64, 0, st, (const Jacobian<F> *)D.bases_jac.as<Jacobian<F>>(),
                   (const Fr *)D.scalars.as<Fr>(), n, D.partial.as<XYZZ<F>>());
        LAUNCH(D, (k_batch_sums<F>), cdiv(count, 64), 64, 0, st, (const XYZZ<F> *)D.partial.as<XYZZ<F>>(), (const uint64_t *)D.off.as<uint64_t>(),
               count, D.out_jac.as<Jacobian<F>>());
        run_ingest<F, true>(D, st, D.out_jac.as<Jacobian<F>>(), D.out_norm.p, nullptr, count);
        d2h(D, out, D.out_norm.p, count * sizeof(Jacobian<F>), st);
        CK(cudaStreamSynchronize(st));
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.kernel_launches = D.launches;
        g_stats.h2d_bytes = (double)n * (sizeof(Fr) + sizeof(Jacobian<F>)) + (double)(count + 1) * 8;
        g_stats.d2h_bytes = (double)count * sizeof(Jacobian<F>);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

template <class F>
int pin_bases(const uint64_t *bases, const void *d_affine, size_t n, uint64_t *handle)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (!handle || (n && !bases && !d_affine)) return fail(B200_ERR_ARG, "null argument");
    try {
        auto pb = std::make_unique<PinnedBases>();
        pb->group = HostOf<F>::group;
        pb->n = n;
        const auto ranges = split_range(n, d_affine ? 1 : g_devs.size());
        pb->shards.resize(ranges.size());
        for_each_shard(ranges.size(), [&](size_t si) {
            Device &D = g_devs[si];
            Shard &S = pb->shards[si];
            S.dev = (int)si;
            S.begin = ranges[si].first;
            S.count = ranges[si].second;
            CK(cudaSetDevice(D.id));
            const size_t m = std::max<size_t>(S.count, 1);
            CK(cudaMalloc(&S.d_aff, m * sizeof(Affine<F>)));
            CK(cudaMalloc((void **)&S.d_flags, m));
            if (S.count == 0) return;
            if (d_affine) {
                CK(cudaMemcpyAsync(S.d_aff, d_affine, S.count * sizeof(Affine<F>), cudaMemcpyDeviceToDevice, D.stream));
                LAUNCH(D, (k_affine_flags<F>), cdiv(S.count, 256), 256, 0, D.stream, (const Affine<F> *)S.d_aff, S.d_flags, S.count);
            } else {
                D.bases_jac.ensure(S.count * sizeof(Jacobian<F>));
                h2d(D, D.bases_jac.p, bases + S.begin * HostOf<F>::jac_limbs, S.count * sizeof(Jacobian<F>), D.stream);
                run_ingest<F, false>(D, D.stream, D.bases_jac.as<Jacobian<F>>(), S.d_aff, S.d_flags, S.count);
            }
            CK(cudaStreamSynchronize(D.stream));
        });
        const uint64_t h = g_next_handle++;
        g_pinned[h] = std::move(pb);
        *handle = h;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// extend a pinned key by its window multiples (include/b200_msm.h: b200_key_precompute_*)
template <class F>
int key_precompute(uint64_t handle, uint32_t window_bits)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    auto it = g_pinned.find(handle);
    if (it == g_pinned.end() || it->second->group != HostOf<F>::group) return fail(B200_ERR_ARG, "unknown bases handle");
    if (window_bits && (window_bits < 4 || window_bits > 24)) return fail(B200_ERR_ARG, "window_bits out of range (4..24)");
    PinnedBases &pb = *it->second;
    try {
        for_each_shard(pb.shards.size(), [&](size_t si) {
            Shard &S = pb.shards[si];
            if (S.count == 0) return;
            Device &D = g_devs[S.dev];
            CK(cudaSetDevice(D.id));
            const uint32_t c = window_bits ? window_bits : choose_precompute_window(S.count);
            const uint32_t W = (255 + c - 1) / c;
            if ((size_t)W * S.count >= (1ull << 31)) throw CudaError{"key too long to precompute: W * n must stay below 2^31 points per device"};
            if (S.d_pre) {
                CK(cudaFree(S.d_pre));
                S.d_pre = nullptr;
                S.pre_c = S.pre_W = 0;
            }
            // built into a local pointer and published (d_pre, pre_c, pre_W together) only when complete: an error
            // half-way leaves the key as a plain key instead of one with a table but no window geometry
            void *table = nullptr;
            CK(cudaMalloc(&table, (size_t)W * S.count * sizeof(Affine<F>)));
            try {
                Affine<F> *lv = reinterpret_cast<Affine<F> *>(table);
                CK(cudaMemcpyAsync(lv, S.d_aff, S.count * sizeof(Affine<F>), cudaMemcpyDeviceToDevice, D.stream));
                D.bases_jac.ensure(S.count * sizeof(Jacobian<F>));
                for (uint32_t k = 1; k < W; k++) {
                    LAUNCH(D, (k_key_level<F>), cdiv(S.count, 128), 128, 0, D.stream, (const Affine<F> *)(lv + (size_t)(k - 1) * S.count), c,
                           S.count, D.bases_jac.as<Jacobian<F>>());
                    run_ingest<F, false>(D, D.stream, D.bases_jac.as<Jacobian<F>>(), lv + (size_t)k * S.count, nullptr, S.count);
                }
                CK(cudaStreamSynchronize(D.stream));
            } catch (...) {
                cudaFree(table);
                throw;
            }
            S.d_pre = table;
            S.pre_c = c;
            S.pre_W = W;
        });
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// scalars: host pointer (d_scalars == nullptr) or device pointer on shard 0's device
template <class F>
int msm_pinned(uint64_t handle, size_t offset, const uint64_t *scalars, const void *d_scalars, size_t n, void *stream,
               uint64_t *out)
{
    typedef typename HostOf<F>::type HF;
    typedef host::HJac<HF> J;
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    auto it = g_pinned.find(handle);
    if (it == g_pinned.end() || it->second->group != HostOf<F>::group) return fail(B200_ERR_ARG, "unknown bases handle");
    PinnedBases &pb = *it->second;
    if (!out || offset > pb.n || n > pb.n - offset || (n && !scalars && !d_scalars)) return fail(B200_ERR_ARG, "bad range or null argument");
    if (d_scalars && pb.shards.size() != 1) return fail(B200_ERR_ARG, "device-resident scalars need a single-device key");
    if (n == 0) {
        write_point<HF>(out, J::inf());
        return B200_OK;
    }
    try {
        // intersect [offset, offset+n) with every shard
        struct Piece {
            size_t shard, lo, cnt;
        };
        std::vector<Piece> pieces;
        for (size_t si = 0; si < pb.shards.size(); si++) {
            const Shard &S = pb.shards[si];
            const size_t lo = std::max(offset, S.begin), hi = std::min(offset + n, S.begin + S.count);
            if (lo < hi) pieces.push_back({si, lo, hi - lo});
        }
        if (pieces.empty()) return fail(B200_ERR_ARG, "range does not intersect the key");
        std::vector<MsmGeom> geoms(pieces.size());
        for (auto &d : g_devs) d.launches = 0;
        for_each_shard(pieces.size(), [&](size_t pi) {
            const Piece &P = pieces[pi];
            const Shard &S = pb.shards[P.shard];
            Device &D = g_devs[S.dev];
            CK(cudaSetDevice(D.id));
            cudaStream_t st = (d_scalars && stream) ? (cudaStream_t)stream : D.stream;
            engine_enter(D, st);
            const Fr *ds = nullptr;
            const uint64_t *hs = nullptr;  // host scalars of a piece that takes the multi-kernel pipeline: uploaded in chunks
            if (d_scalars) {
                ds = reinterpret_cast<const Fr *>(d_scalars);
            } else if (P.cnt <= SMALL_MAX_N && g_tune_c == 0) {
                D.scalars.ensure(P.cnt * sizeof(Fr));
                h2d(D, D.scalars.p, scalars + (P.lo - offset) * 4, P.cnt * sizeof(Fr), st);
                ds = D.scalars.as<Fr>();
            } else {
                hs = scalars + (P.lo - offset) * 4;
            }
            const Affine<F> *aff = reinterpret_cast<const Affine<F> *>(S.d_aff) + (P.lo - S.begin);
            if (P.cnt <= SMALL_MAX_N && g_tune_c == 0) {
                // one kernel on the resident affine bases (the small levels of CPPoly::prove, poly.h:77-88)
                MsmGeom g;
                g.c = SMALL_C;
                g.W = g.Wb = SMALL_W;
                g.B = SMALL_NBK;
                g.NB = g.W * g.B;
                g.pre_stride = g.pre_off = g.ones = 0;
                g.red_jobs = g.red_logS = g.red_hb = g.red_lb = 0;
                g.L = 0;
                D.window_sums.ensure((size_t)g.W * sizeof(XYZZ<F>));
                D.totals.ensure(32);
                D.ensure_pinned((size_t)g.W * sizeof(XYZZ<F>) + 64);
                CK(cudaEventRecord(D.ev[0], st));
                CK(cudaEventRecord(D.ev[2], st));
                LAUNCH(D, (k_msm_small<F, Affine<F>>), g.W, SMALL_THREADS, 0, st, aff, ds, (uint32_t)P.cnt, D.window_sums.as<XYZZ<F>>());
                CK(cudaEventRecord(D.ev[3], st));
                CK(cudaMemsetAsync(D.totals.p, 0, 8, st));
                CK(cudaMemcpyAsync(D.h_pinned, D.window_sums.p, (size_t)g.W * sizeof(XYZZ<F>), cudaMemcpyDeviceToHost, st));
                CK(cudaMemcpyAsync((char *)D.h_pinned + (size_t)g.W * sizeof(XYZZ<F>), D.totals.p, 8, cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(D.ev[1], st));
                geoms[pi] = g;
                CK(cudaStreamSynchronize(st));
                return;
            }
            // precomputed levels pay when this (sub-)range fills their one big bucket set
            bool pre = S.d_pre && g_tune_pre && g_tune_c == 0 && P.cnt > SMALL_MAX_N;
            if (pre && g_tune_pre != 2) pre = precomputed_pays(P.cnt, S.pre_c);  // 2 = always (tests)
            if (pre) {
                const MsmGeom gp = choose_geometry(P.cnt, S.pre_c, (uint32_t)S.count, (uint32_t)(P.lo - S.begin), sizeof(F) == 64);
                if (hs)
                    geoms[pi] = enqueue_msm_host_scalars<F>(D, st, reinterpret_cast<const Affine<F> *>(S.d_pre), S.d_flags + (P.lo - S.begin), hs,
                                                            P.cnt, &gp);
                else
                    geoms[pi] = enqueue_msm<F>(D, st, reinterpret_cast<const Affine<F> *>(S.d_pre), S.d_flags + (P.lo - S.begin), ds,
                                               P.cnt, &gp);
            } else if (hs) {
                geoms[pi] = enqueue_msm_host_scalars<F>(D, st, aff, S.d_flags + (P.lo - S.begin), hs, P.cnt);
            } else {
                geoms[pi] = enqueue_msm<F>(D, st, aff, S.d_flags + (P.lo - S.begin), ds, P.cnt);
            }
            CK(cudaStreamSynchronize(st));
        });
        const auto t0 = std::chrono::steady_clock::now();
        J total = J::inf();
        double d2h = 0;
        for (size_t pi = 0; pi < pieces.size(); pi++) {
            total = host::jac_add(total, finalize_windows<F>(g_devs[pb.shards[pieces[pi].shard].dev], geoms[pi]));
            d2h += (double)result_points(geoms[pi]) * sizeof(XYZZ<F>) + 8;
        }
        write_point<HF>(out, total);
        const auto t1 = std::chrono::steady_clock::now();
        const Device &D0 = g_devs[pb.shards[pieces[0].shard].dev];
        const uint32_t *tot = reinterpret_cast<const uint32_t *>((const char *)D0.h_pinned + (size_t)result_points(geoms[0]) * sizeof(XYZZ<F>));
        fill_stats(D0, pieces[0].cnt, geoms[0], tot, std::chrono::duration<double, std::micro>(t1 - t0).count(),
                   d_scalars ? 0.0 : (double)n * sizeof(Fr), d2h);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// fixed-base tables and batch_exp
// ------------------------------------------------------------------------------
inline uint32_t choose_table_window(size_t n)
{
    uint32_t best_w = 1;
    double best = 1e300;
    for (uint32_t w = 1; w <= 18; w++) {
        const double rows = std::ceil(254.0 / w);
        const double cost = rows * ((double)n * 10.0 + (double)(1u << w) * 40.0);
        if (cost < best) {
            best = cost;
            best_w = w;
        }
    }
    return best_w;
}

template <class F>
int table_create(const uint64_t *base, size_t expected, uint64_t *handle)
{
    typedef typename HostOf<F>::type HF;
    typedef host::HJac<HF> J;
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (!base || !handle) return fail(B200_ERR_ARG, "null argument");
    try {
        auto wt = std::make_unique<WindowTable>();
        wt->group = HostOf<F>::group;
        wt->w = choose_table_window(std::max<size_t>(expected, 1));
        wt->rows = (254 + wt->w - 1) / wt->w;
        // row bases g_o = 2^(o w) g on the host (254 doublings), affine
        J g;
        memcpy(&g, base, sizeof g);
        struct HA {
            HF x, y;
        };
        std::vector<HA> rows(wt->rows);
        J cur = g;
        for (uint32_t o = 0; o < wt->rows; o++) {
            const J nrm = host::jac_normalise(cur);
            rows[o] = nrm.is_inf() ? HA{HF::zero(), HF::zero()} : HA{nrm.x, nrm.y};
            if (o + 1 < wt->rows)
                for (uint32_t i = 0; i < wt->w; i++) cur = host::jac_dbl(cur);
        }
        const size_t entries = (size_t)wt->rows << wt->w;
        wt->d_table.resize(g_devs.size(), nullptr);
        for_each_shard(g_devs.size(), [&](size_t di) {
            Device &D = g_devs[di];
            CK(cudaSetDevice(D.id));
            CK(cudaMalloc(&wt->d_table[di], entries * sizeof(Affine<F>)));
            D.coeff.ensure(rows.size() * sizeof(HA));
            D.out_jac.ensure(entries * sizeof(Jacobian<F>));
            CK(cudaMemcpyAsync(D.coeff.p, rows.data(), rows.size() * sizeof(HA), cudaMemcpyHostToDevice, D.stream));
            const uint32_t M = std::min<uint32_t>(32, 1u << wt->w);
            const uint32_t runs = cdiv((size_t)1 << wt->w, M);
            LAUNCH(D, (k_table_rows<F>), dim3(cdiv(runs, 128), wt->rows), 128, 0, D.stream, D.coeff.as<Affine<F>>(), wt->w, M,
                   D.out_jac.as<Jacobian<F>>());
            run_ingest<F, false>(D, D.stream, D.out_jac.as<Jacobian<F>>(), wt->d_table[di], nullptr, entries);
            CK(cudaStreamSynchronize(D.stream));
        });
        const uint64_t h = g_next_handle++;
        g_tables[h] = std::move(wt);
        *handle = h;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// host scalars -> host normalised Jacobian out (d_scalars == nullptr), or device scalars -> device affine out
template <class F>
int batch_exp_table(uint64_t handle, const uint64_t *scalars, const void *d_scalars, size_t n, const uint64_t *coeff,
                    uint64_t *out, void *d_out_affine, void *stream)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    auto it = g_tables.find(handle);
    if (it == g_tables.end() || it->second->group != HostOf<F>::group) return fail(B200_ERR_ARG, "unknown table handle");
    WindowTable &wt = *it->second;
    const bool dev_io = d_scalars != nullptr;
    if (n && ((dev_io && !d_out_affine) || (!dev_io && (!scalars || !out)))) return fail(B200_ERR_ARG, "null argument");
    if (n == 0) return B200_OK;
    try {
        const auto ranges = split_range(n, dev_io ? 1 : g_devs.size());
        for (auto &d : g_devs) d.launches = 0;
        for_each_shard(ranges.size(), [&](size_t si) {
            Device &D = g_devs[si];
            const size_t b = ranges[si].first, m = ranges[si].second;
            if (m == 0) return;
            CK(cudaSetDevice(D.id));
            cudaStream_t st = (dev_io && stream) ? (cudaStream_t)stream : D.stream;
            engine_enter(D, st);
            const Fr *ds;
            if (dev_io) {
                ds = reinterpret_cast<const Fr *>(d_scalars);
            } else {
                D.scalars.ensure(m * sizeof(Fr));
                h2d(D, D.scalars.p, scalars + b * 4, m * sizeof(Fr), st);
                ds = D.scalars.as<Fr>();
            }
            const Fr *dcoeff = nullptr;
            if (coeff) {
                D.coeff.ensure(sizeof(Fr));
                CK(cudaMemcpyAsync(D.coeff.p, coeff, sizeof(Fr), cudaMemcpyHostToDevice, st));
                dcoeff = D.coeff.as<Fr>();
            }
            D.out_jac.ensure(m * sizeof(Jacobian<F>));
            LAUNCH(D, (k_batch_exp<F>), cdiv(m, 128), 128, 0, st, (const Affine<F> *)wt.d_table[si], wt.w, wt.rows, ds, dcoeff, m,
                   D.out_jac.as<Jacobian<F>>());
            if (dev_io) {
                run_ingest<F, false>(D, st, D.out_jac.as<Jacobian<F>>(), d_out_affine, nullptr, m);
                engine_leave(D, st);  // returns without synchronising: D.out_jac / D.prefix stay busy on `st`
            } else {
                D.out_norm.ensure(m * sizeof(Jacobian<F>));
                run_ingest<F, true>(D, st, D.out_jac.as<Jacobian<F>>(), D.out_norm.p, nullptr, m);
                d2h(D, out + b * HostOf<F>::jac_limbs, D.out_norm.p, m * sizeof(Jacobian<F>), st);
                CK(cudaStreamSynchronize(st));
            }
        });
        g_stats = b200_stats_t{};
        g_stats.n = n;
        g_stats.window_bits = wt.w;
        g_stats.num_windows = wt.rows;
        g_stats.kernel_launches = g_devs[0].launches;
        g_stats.h2d_bytes = dev_io ? 0.0 : (double)n * sizeof(Fr);
        g_stats.d2h_bytes = dev_io ? 0.0 : (double)n * sizeof(Jacobian<F>);
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

template <class F>
int batch_exp_once(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out)
{
    uint64_t h = 0;
    int rc = table_create<F>(base, n, &h);
    if (rc) return rc;
    rc = batch_exp_table<F>(h, scalars, nullptr, n, coeff, out, nullptr, nullptr);
    for (size_t di = 0; di < g_devs.size(); di++) {
        cudaSetDevice(g_devs[di].id);
        cudaFree(g_tables[h]->d_table[di]);
    }
    g_tables.erase(h);
    return rc;
}

template <class F>
int batch_to_affine(uint64_t *pts, size_t n)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (n && !pts) return fail(B200_ERR_ARG, "null argument");
    if (n == 0) return B200_OK;
    try {
        const auto ranges = split_range(n, g_devs.size());
        for_each_shard(ranges.size(), [&](size_t si) {
            Device &D = g_devs[si];
            const size_t b = ranges[si].first, m = ranges[si].second;
            if (m == 0) return;
            CK(cudaSetDevice(D.id));
            D.bases_jac.ensure(m * sizeof(Jacobian<F>));
            D.out_norm.ensure(m * sizeof(Jacobian<F>));
            uint64_t *hp = pts + b * HostOf<F>::jac_limbs;
            h2d(D, D.bases_jac.p, hp, m * sizeof(Jacobian<F>), D.stream);
            run_ingest<F, true>(D, D.stream, D.bases_jac.as<Jacobian<F>>(), D.out_norm.p, nullptr, m);
            d2h(D, hp, D.out_norm.p, m * sizeof(Jacobian<F>), D.stream);
            CK(cudaStreamSynchronize(D.stream));
        });
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

// ------------------------------------------------------------------------------
// parity hooks
// ------------------------------------------------------------------------------
template <class T>
struct PrimeOp {
    int op;
    __device__ T operator()(const T &a, const T &b) const { return prime_field_test_op(op, a, b); }
};
struct Fq2Op {
    int op;
    __device__ Fq2 operator()(const Fq2 &a, const Fq2 &b) const { return field_test_op<Fq2>(op, a, b); }
};
template <class F>
struct GroupOp {
    int op;
    uint32_t k;
    bool has_b;
    __device__ Jacobian<F> operator()(const Jacobian<F> &a, const Jacobian<F> &b) const
    {
        return group_test_op<F>(op, a, has_b ? b : Jacobian<F>::inf(), k);
    }
};

template <class T, class Op>
int run_elementwise(const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out, Op op)
{
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (n == 0) return B200_OK;
    if (!a || !out) return fail(B200_ERR_ARG, "null argument");
    try {
        Device &D = g_devs[0];
        CK(cudaSetDevice(D.id));
        D.bases_jac.ensure(n * sizeof(T));
        D.out_jac.ensure(n * sizeof(T));
        D.out_norm.ensure(n * sizeof(T));
        CK(cudaMemcpyAsync(D.bases_jac.p, a, n * sizeof(T), cudaMemcpyHostToDevice, D.stream));
        if (b) CK(cudaMemcpyAsync(D.out_jac.p, b, n * sizeof(T), cudaMemcpyHostToDevice, D.stream));
        LAUNCH(D, (k_elementwise<T, Op>), cdiv(n, 128), 128, 0, D.stream, D.bases_jac.as<T>(), b ? D.out_jac.as<T>() : (const T *)nullptr,
               D.out_norm.as<T>(), n, op);
        CK(cudaMemcpyAsync(out, D.out_norm.p, n * sizeof(T), cudaMemcpyDeviceToHost, D.stream));
        CK(cudaStreamSynchronize(D.stream));
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}


template <class F>
int test_group_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint32_t k, uint64_t *out)
{
    return run_elementwise<Jacobian<F>>(a, b, n, out, GroupOp<F>{op, k, b != nullptr});
}

// CUDA 12 loads a kernel's code at its first launch; the small-path kernels (n <= 4096: cplink's commits and prove, the
// sigma proofs and polynomial commitments of the sum-check gadgets) are queried once at initialisation so that the first
// small MSM of a process runs at steady-state cost (cplink's first commit: 26-45 ms -> 2.7 ms) without loading all ~200
// kernels eagerly (that costs ~0.9 s of start-up).
template <class F>
void preload_small_path()
{
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_msm_small<F, Jacobian<F>>);
    cudaFuncGetAttributes(&a, k_msm_small<F, Affine<F>>);
    cudaFuncGetAttributes(&a, k_ingest<F, false>);
    cudaFuncGetAttributes(&a, k_ingest<F, true>);
    cudaFuncGetAttributes(&a, k_batch_terms<F>);
    cudaFuncGetAttributes(&a, k_batch_sums<F>);
    cudaGetLastError();
}

#define B200_INSTANTIATE_GROUP(F)                                                                                       \
    template void preload_small_path<F>();                                                                              \
    template int msm_host<F>(const uint64_t *, const uint64_t *, size_t, uint64_t *);                                   \
    template int msm_batch<F>(const uint64_t *, const uint64_t *, const uint64_t *, size_t, uint64_t *);                \
    template int pin_bases<F>(const uint64_t *, const void *, size_t, uint64_t *);                                      \
    template int msm_pinned<F>(uint64_t, size_t, const uint64_t *, const void *, size_t, void *, uint64_t *);           \
    template int key_precompute<F>(uint64_t, uint32_t);                                                                 \
    template int table_create<F>(const uint64_t *, size_t, uint64_t *);                                                 \
    template int batch_exp_table<F>(uint64_t, const uint64_t *, const void *, size_t, const uint64_t *, uint64_t *,     \
                                    void *, void *);                                                                    \
    template int batch_exp_once<F>(const uint64_t *, const uint64_t *, size_t, const uint64_t *, uint64_t *);           \
    template int batch_to_affine<F>(uint64_t *, size_t);                                                                \
    template int test_group_op<F>(int, const uint64_t *, const uint64_t *, size_t, uint32_t, uint64_t *);

}  // namespace eng
}  // namespace b200
