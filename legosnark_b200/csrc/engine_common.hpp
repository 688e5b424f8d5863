// engine_common.hpp — state and plumbing shared by the three translation units of
// libb200msm.so (engine_core.cu: C-ABI + lifecycle; engine_g1.cu / engine_g2.cu:
// the per-group instantiations of engine_impl.cuh).  Split so the groups compile
// in parallel.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200_msm.h"

namespace b200 {
namespace eng {


// ------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------
extern std::mutex g_mu;
extern std::string g_err;
int fail(int code, const char *fmt, ...);

struct CudaError {
    std::string msg;
};

#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) {                                                                             \
            char b__[512];                                                                                    \
            snprintf(b__, sizeof b__, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            throw CudaError{b__};                                                                             \
        }                                                                                                     \
    } while (0)

// ------------------------------------------------------------------------------
// device context
// ------------------------------------------------------------------------------
// B200_TRACE=1 (engine_core.cu): host-side stalls above 1 ms inside a call -- buffer growth, a kernel's first launch (CUDA loads
// its code then) -- are reported on stderr next to the per-call lines
extern bool g_trace;
struct TraceStall {
    const char *what;
    size_t arg;
    std::chrono::steady_clock::time_point t0;
    TraceStall(const char *w, size_t a) : what(w), arg(a)
    {
        if (g_trace) t0 = std::chrono::steady_clock::now();
    }
    ~TraceStall()
    {
        if (!g_trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms >= 1.0) fprintf(stderr, "[b200]     stall %8.3f ms  %s (%zu)\n", ms, what, arg);
    }
};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes)
    {
        if (bytes <= cap) return;
        TraceStall ts("device buffer growth, bytes", bytes);
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        const size_t want = std::max<size_t>(256, bytes + bytes / 8);
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

constexpr size_t MAX_CHUNKS = 16;  // upload chunks of one pipelined host-buffer MSM

struct Device {
    int id = 0;
    int sms = 148;
    cudaStream_t stream = nullptr;       // compute (and every copy of the non-pipelined calls)
    cudaStream_t copy_stream = nullptr;  // H2D of the next chunk while the previous one is accumulated
    cudaEvent_t ev_ready[MAX_CHUNKS] = {};  // chunk j has landed on the device
    cudaEvent_t ev_sync = nullptr;
    // Pipelined host-buffer MSMs sort chunk j + 1 (latency-bound) on sort_stream while chunk j is accumulated (multiply-add
    // bound) on `stream`: the buffers a sort hands to the accumulation exist twice and swap roles from chunk to chunk.
    cudaStream_t sort_stream = nullptr;
    cudaEvent_t ev_sorted[2] = {nullptr, nullptr};    // the sort into set k has finished
    cudaEvent_t ev_set_free[2] = {nullptr, nullptr};  // the accumulation / fold that read set k has finished
    // Last use of engine-owned buffers / cached tables by a call that returned WITHOUT synchronising (the *_dev entry
    // points run on the caller's stream): every entry point first makes its stream wait for it (engine_enter), the
    // asynchronous ones record it when they have enqueued their work (engine_leave).  All engine streams are
    // non-blocking, so without this a later call on another stream could read tables or scratch still being written.
    cudaEvent_t ev_busy = nullptr;
    // inputs
    DevBuf scalars, bases_jac, bases_aff, flags, prefix;
    // sort
    DevBuf cnt, off, cursor, toff, tile_sums, totals, entries, digits, meta, order, len_hist, len_cursor;
    // radix-partition sort (engine_sort.cu): standard-form scalars, (entry, bucket) pairs, per-partition counters
    DevBuf std_scalars, items, part;
    DevBuf task_bucket;  // per task: its bucket | first-task flag (seeded accumulation of pipelined chunks)
    struct SortSet {
        DevBuf cnt, off, toff, entries, meta, order, totals, split, big, task_bucket, ones_idx;
    } alt;  // the second set of what a sort produces for the accumulation (see sort_stream)
    void swap_sort_set()
    {
        std::swap(cnt, alt.cnt);
        std::swap(off, alt.off);
        std::swap(toff, alt.toff);
        std::swap(entries, alt.entries);
        std::swap(meta, alt.meta);
        std::swap(order, alt.order);
        std::swap(totals, alt.totals);
        std::swap(split, alt.split);
        std::swap(big, alt.big);
        std::swap(task_bucket, alt.task_bucket);
        std::swap(ones_idx, alt.ones_idx);
    }
    // batch-affine accumulation (pair_kernels.cuh): the points of tree levels 1 and 2
    DevBuf pa1, pa2;
    bool sort_attr_set = false;
    // accumulation / reduction
    DevBuf partial, seg_run, seg_acc, job_out, split, big, done, window_sums, bucket_sum;
    // scalars equal to one, set aside by k_digit_count: index list, block partials, ticket, sum
    DevBuf ones_idx, ones_part, ones_done, ones_sum;
    // batch_exp
    DevBuf out_jac, out_norm, coeff;
    // Fr vector kernels (engine_fr.cu): ping-pong value buffers, challenge point, witness coefficients
    DevBuf fr_a, fr_b, fr_r, fr_w;
    void *h_pinned = nullptr;  // small pinned staging (window sums, totals)
    size_t h_pinned_cap = 0;
    uint32_t launches = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // [0,1] whole pipeline, [2,3] k_accumulate

    void ensure_pinned(size_t bytes)
    {
        if (bytes <= h_pinned_cap) return;
        if (h_pinned) CK(cudaFreeHost(h_pinned));
        h_pinned = nullptr;
        CK(cudaHostAlloc(&h_pinned, bytes, cudaHostAllocDefault));
        h_pinned_cap = bytes;
    }
    void release()
    {
        cudaSetDevice(id);
        DevBuf *all[] = {&scalars, &bases_jac, &bases_aff, &flags, &prefix, &cnt, &off, &cursor, &toff, &tile_sums, &totals,
                         &entries, &digits, &meta, &order, &len_hist, &len_cursor, &partial, &seg_run, &seg_acc, &job_out, &split, &done, &window_sums, &bucket_sum, &out_jac,
                         &out_norm, &coeff, &fr_a, &fr_b, &fr_r, &fr_w, &ones_idx, &ones_part, &ones_done, &ones_sum, &big, &std_scalars, &items, &part, &pa1, &pa2, &task_bucket};
        for (DevBuf *b : all) b->release();
        DevBuf *alts[] = {&alt.cnt, &alt.off, &alt.toff, &alt.entries, &alt.meta, &alt.order, &alt.totals, &alt.split, &alt.big, &alt.task_bucket,
                          &alt.ones_idx};
        for (DevBuf *b : alts) b->release();
        if (sort_stream) cudaStreamDestroy(sort_stream);
        sort_stream = nullptr;
        for (int k = 0; k < 2; k++) {
            if (ev_sorted[k]) cudaEventDestroy(ev_sorted[k]);
            if (ev_set_free[k]) cudaEventDestroy(ev_set_free[k]);
            ev_sorted[k] = ev_set_free[k] = nullptr;
        }
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        h_pinned_cap = 0;
        if (stream) cudaStreamDestroy(stream);
        stream = nullptr;
        if (copy_stream) cudaStreamDestroy(copy_stream);
        copy_stream = nullptr;
        for (auto &e : ev_ready) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
        if (ev_sync) cudaEventDestroy(ev_sync);
        ev_sync = nullptr;
        if (ev_busy) cudaEventDestroy(ev_busy);
        ev_busy = nullptr;
        for (auto &e : ev) {
            if (e) cudaEventDestroy(e);
            e = nullptr;
        }
    }
};

struct Shard {
    int dev;  // index into g_devs
    size_t begin, count;
    void *d_aff = nullptr;      // Affine<F>[count]
    uint8_t *d_flags = nullptr; // count
    // precomputed key (b200_key_precompute_*): level k = 2^(pre_c k) * P_i at d_pre[k * count + i]
    void *d_pre = nullptr;      // Affine<F>[pre_W * count]; level 0 is a copy of d_aff
    uint32_t pre_c = 0, pre_W = 0;
};

struct PinnedBases {
    int group;  // 0 G1, 1 G2
    size_t n;
    std::vector<Shard> shards;
};

struct WindowTable {
    int group;
    uint32_t w, rows;
    std::vector<void *> d_table;  // per device: Affine<F>[rows << w]
};


extern std::vector<Device> g_devs;
extern bool g_init;
extern std::map<uint64_t, std::unique_ptr<PinnedBases>> g_pinned;
extern std::map<uint64_t, std::unique_ptr<WindowTable>> g_tables;
extern uint64_t g_next_handle;
extern b200_stats_t g_stats;
extern int g_tune_c, g_tune_L, g_tune_chunks, g_tune_logS, g_tune_split, g_tune_pre, g_tune_ones, g_tune_host_horner;
extern int g_tune_marginals;  // 1: one-window reduction by marginal sums (measured: no faster, profiles/r2c); 0 (default): bit decomposition over all segments
extern int g_tune_g2blocks;  // k_accumulate<Fq2> builds: 1 (default) ptxas may use 255 registers (252 used, two blocks per SM; 3 % faster), 2: 190 registers (round 1), 3: 168 registers + spills, three blocks per SM (5 % slower); profiles/r4f_stage.jsonl
extern int g_tune_red_block;  // threads per block of k_reduce_segments (32, 64, 96 or 128)
extern int g_tune_g1paired;  // k_accumulate<Fq> with the mixed addition's independent products issued in pairs: 1 (ptxas picks the registers), 2 (four blocks per SM)
extern int g_tune_quads;  // 1 (default): stage 2 of the window reduction with quad-cooperative additions (k_reduce_bits_quad); 0: one thread per partial sum
extern int g_tune_dense_direct;  // 1 (default): in pipelined MSMs k_accumulate writes single-task buckets straight into the dense array; 0: every chunk folds all buckets
extern int g_tune_overlap_sort;  // 1 (default): pipelined MSMs sort the next chunk on a second stream while the current one is accumulated
extern int g_tune_pinned_chunks;  // upload chunks of the host scalars of a resident-key MSM: 0 = auto (1 below 2^18, 2 below 2^21, else 4)
extern int g_tune_g2pair;  // 1: G2 accumulation with two lanes per task (measured 5 % slower: profiles/r2n_g2_lane_pairs.jsonl); 0 (default): one thread per task
extern int g_tune_even_chunks;  // 1 (default): equal upload chunks; 0: short first chunk (measured: no gain)
extern int g_tune_ba;    // batch-affine tree levels in front of the XYZZ accumulation: 0 (off), 1 or 2
extern int g_tune_sort;  // 1 (default): radix partition staged through shared memory; 0: round-1 global-atomics counting sort
// set while the second MSM of a knowledge-commitment pair runs: every device still holds the
// scalars of its shard in D.scalars from the first one, so the host-buffer paths skip that upload
extern bool g_scalars_resident;

inline void engine_enter(Device &D, cudaStream_t st) { CK(cudaStreamWaitEvent(st, D.ev_busy, 0)); }
inline void engine_leave(Device &D, cudaStream_t st) { CK(cudaEventRecord(D.ev_busy, st)); }

inline uint32_t cdiv(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

#define LAUNCH(D, kernel, grid, block, smem, st, ...)             \
    do {                                                          \
        TraceStall ts_("launch of " #kernel, 0);                  \
        kernel<<<grid, block, smem, st>>>(__VA_ARGS__);           \
        (D).launches++;                                           \
        CK(cudaGetLastError());                                   \
    } while (0)

std::vector<std::pair<size_t, size_t>> split_range(size_t n, size_t parts);

template <class Fn>
void for_each_shard(size_t nshards, Fn fn)
{
    if (nshards == 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::string> errs(nshards);
    try {
        for (size_t i = 0; i < nshards; i++)
            th.emplace_back([&, i] {
                try {
                    fn(i);
                } catch (const CudaError &e) {
                    errs[i] = e.msg;
                } catch (const std::exception &e) {
                    errs[i] = std::string("host error in a device worker: ") + e.what();
                } catch (...) {
                    errs[i] = "unknown error in a device worker";
                }
            });
    } catch (...) {  // thread creation failed part-way: the workers already running are joined, not abandoned
        for (auto &t : th) t.join();
        throw CudaError{"could not start the per-device worker threads"};
    }
    for (auto &t : th) t.join();
    for (auto &e : errs)
        if (!e.empty()) throw CudaError{e};
}

// per-group entry points: templates defined in engine_impl.cuh, explicitly
// instantiated for Fq in engine_g1.cu and for Fq2 in engine_g2.cu
template <class F> void preload_small_path();
template <class F> int msm_host(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out);
template <class F> int msm_batch(const uint64_t *bases, const uint64_t *scalars, const uint64_t *offsets, size_t count, uint64_t *out);
template <class F> int pin_bases(const uint64_t *bases, const void *d_affine, size_t n, uint64_t *handle);
template <class F> int key_precompute(uint64_t handle, uint32_t window_bits);
template <class F> int msm_pinned(uint64_t handle, size_t offset, const uint64_t *scalars, const void *d_scalars, size_t n,
                                  void *stream, uint64_t *out);
template <class F> int table_create(const uint64_t *base, size_t expected, uint64_t *handle);
template <class F> int batch_exp_table(uint64_t handle, const uint64_t *scalars, const void *d_scalars, size_t n,
                                       const uint64_t *coeff, uint64_t *out, void *d_out_affine, void *stream);
template <class F> int batch_exp_once(const uint64_t *base, const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out);
template <class F> int batch_to_affine(uint64_t *pts, size_t n);
template <class F> int test_group_op(int op, const uint64_t *a, const uint64_t *b, size_t n, uint32_t k, uint64_t *out);
int test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out);  // engine_g1.cu

// Fr vector work (engine_fr.cu).  fr_fold_device: v (2^d values) and r (d values) are uploaded to
// D.fr_a / D.fr_r, folded on D.stream; witness coefficients (if wanted) land in D.fr_w in the
// reference's w_coeffs layout, the final value in D.fr_a[0] or D.fr_b[0] (returned pointer).
const void *fr_fold_device(Device &D, const uint64_t *v, const uint64_t *r, size_t d, bool want_w);
int fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs, uint64_t *eval);
int fr_mle_bind(const uint64_t *table, size_t half, const uint64_t *r, uint64_t *out);
int fr_fft(uint64_t *a, void *d_a, size_t log_n, int mode, const uint64_t *g, void *stream);
int fr_qap_h(const uint64_t *aA, const uint64_t *aB, const uint64_t *aC, size_t log_big, size_t log_small, const uint64_t *g,
             const uint64_t *div, uint64_t *H);
int fr_step_fft(uint64_t *a, size_t log_big, size_t log_small, int mode, const uint64_t *g);
int fr_geometric_quotients(uint64_t *out, const uint64_t *in, size_t n, const uint64_t *consts, size_t nf);
int fr_scale_inv_geometric(uint64_t *P, size_t n_geo, const uint64_t *c1, const uint64_t *ratio, const uint64_t *c0, size_t n_tail,
                           const uint64_t *tail);
int fr_eq_table(const uint64_t *r, size_t d, uint64_t *out);
int fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v);
int fr_sumcheck_round(const uint64_t *a, const uint64_t *b, const uint64_t *w, size_t half, uint64_t *out);
int fr_sumcheck_rounds(const uint64_t *a, const uint64_t *b, const uint64_t *r, size_t d, uint64_t *h);
void fr_release();
// point (de)compression (engine_wire.cu)
int compress_g1(const uint64_t *pts, size_t n, int fl, uint64_t *x, uint8_t *flags);
int compress_g2(const uint64_t *pts, size_t n, int fl, uint64_t *x, uint8_t *flags);
int decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts, uint8_t *bad);
int decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int fl, uint64_t *pts, uint8_t *bad);
int cppoly_prove_g1(uint64_t key, const uint64_t *v, const uint64_t *r, size_t d, uint64_t *witness, uint64_t *eval);  // engine_g1.cu

}  // namespace eng
}  // namespace b200
