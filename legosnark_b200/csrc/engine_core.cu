// engine_core.cu — lifecycle, shared state and the extern "C" surface declared in
// include/b200_msm.h.  The per-group work is in engine_g1.cu / engine_g2.cu.
#include <chrono>
#include <cstdio>
#include <condition_variable>
#include <cstdlib>
#include <thread>

#include "engine_common.hpp"
#include "field.cuh"
#include "host_arith.hpp"
#include "host_copy.hpp"
#include "peak_kernels.cuh"

namespace b200 {
namespace eng {

std::mutex g_mu;
std::string g_err;
// B200_TRACE=1: one stderr line per C-ABI call (ApiScope below) and per phase of b200_init
bool g_trace = [] { const char *e = std::getenv("B200_TRACE"); return e && *e && *e != '0'; }();
static double trace_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

std::vector<Device> g_devs;
bool g_init = false;
std::map<uint64_t, std::unique_ptr<PinnedBases>> g_pinned;
std::map<uint64_t, std::unique_ptr<WindowTable>> g_tables;
uint64_t g_next_handle = 1;
b200_stats_t g_stats;
int g_tune_c = 0, g_tune_L = 0, g_tune_chunks = 0, g_tune_logS = -1, g_tune_split = 0, g_tune_pre = 1, g_tune_ones = 1, g_tune_host_horner = 1;
int g_tune_sort = 1, g_tune_marginals = 0, g_tune_ba = 0, g_tune_g2pair = 0, g_tune_even_chunks = 1, g_tune_g2blocks = 1, g_tune_red_block = 128, g_tune_g1paired = 0, g_tune_quads = 1, g_tune_dense_direct = 1, g_tune_overlap_sort = 1, g_tune_pinned_chunks = 0;
bool g_scalars_resident = false;

static std::map<int, std::unique_ptr<Stager>> g_stagers;  // by CUDA device ordinal
static std::mutex g_stager_mu;
Stager &stager_of(Device &D)
{
    std::lock_guard<std::mutex> lk(g_stager_mu);
    auto &s = g_stagers[D.id];
    if (!s) s = std::make_unique<Stager>();
    return *s;
}
static void release_stagers()
{
    for (auto &kv : g_stagers) {
        cudaSetDevice(kv.first);
        kv.second->release();
    }
    g_stagers.clear();
}

std::vector<std::pair<size_t, size_t>> split_range(size_t n, size_t parts)
{
    // [g * floor(n/G), ...), last shard takes the remainder (mirrors multiexp.tcc:417-431)
    std::vector<std::pair<size_t, size_t>> r;
    if (parts <= 1 || n < parts) {
        r.push_back({0, n});
        return r;
    }
    const size_t one = n / parts;
    for (size_t i = 0; i < parts; i++) r.push_back({i * one, i == parts - 1 ? n - i * one : one});
    return r;
}

const size_t G1_WTAB[22] = {1, 5, 11, 32, 55, 162, 360, 815, 2373, 6978, 7122, 0, 57818, 0, 169679,
                            439759, 936073, 0, 4666555, 7580404, 0, 34552892};
const size_t G2_WTAB[22] = {1, 5, 10, 25, 59, 154, 334, 743, 2034, 4988, 8888, 26271, 39768, 106276,
                            141703, 462423, 926872, 0, 4873049, 5706708, 0, 31673815};

size_t libff_window_size(const size_t *tab, size_t num_scalars)
{
    // get_exp_window_size, multiexp.tcc:509-545
    size_t window = 1;
    for (long i = 21; i >= 0; --i)
        if (tab[i] != 0 && num_scalars >= tab[i]) {
            window = (size_t)i + 1;
            break;
        }
    return window;
}

int init_devices(const int *ids, int n)
{
    if (g_init) return B200_OK;
    int visible = 0;
    double tr0 = g_trace ? trace_ms() : 0;
    auto phase = [&](const char *what) {
        if (!g_trace) return;
        const double now = trace_ms();
        fprintf(stderr, "[b200]   init: %-40s %9.3f ms\n", what, now - tr0);
        tr0 = now;
    };
    cudaError_t e = cudaGetDeviceCount(&visible);
    phase("cudaGetDeviceCount (driver start-up)");
    if (e != cudaSuccess || visible == 0)
        return fail(B200_ERR_NO_DEVICE, "no CUDA device visible (%s); this engine has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    std::vector<int> use;
    if (ids) {
        for (int i = 0; i < n; i++) use.push_back(ids[i]);
    } else {
        const int want = n <= 0 ? visible : std::min(n, visible);
        if (n > visible) return fail(B200_ERR_NO_DEVICE, "%d GPUs requested, %d visible", n, visible);
        for (int i = 0; i < want; i++) use.push_back(i);
    }
    try {
        g_devs.clear();
        g_devs.resize(use.size());
        for (size_t i = 0; i < use.size(); i++) {
            if (use[i] < 0 || use[i] >= visible) return fail(B200_ERR_ARG, "device ordinal %d out of range", use[i]);
            Device &D = g_devs[i];
            D.id = use[i];
            CK(cudaSetDevice(D.id));
            cudaDeviceProp prop;
            CK(cudaGetDeviceProperties(&prop, D.id));
            D.sms = prop.multiProcessorCount;
            CK(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
            phase("context + first stream");
            CK(cudaStreamCreateWithFlags(&D.copy_stream, cudaStreamNonBlocking));
            {   // the sort of the next chunk should get SM slots as soon as accumulation blocks retire: higher priority
                int lo_prio = 0, hi_prio = 0;
                CK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
                CK(cudaStreamCreateWithPriority(&D.sort_stream, cudaStreamNonBlocking, hi_prio));
                for (int k = 0; k < 2; k++) {
                    CK(cudaEventCreateWithFlags(&D.ev_sorted[k], cudaEventDisableTiming));
                    CK(cudaEventCreateWithFlags(&D.ev_set_free[k], cudaEventDisableTiming));
                }
            }
            for (auto &e : D.ev) CK(cudaEventCreate(&e));
            for (auto &e : D.ev_ready) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&D.ev_sync, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&D.ev_busy, cudaEventDisableTiming));
            // buffers of the small-MSM / batch paths (n <= 4096: cplink, the sigma proofs and polynomial commitments of the
            // sum-check gadgets) exist before the first call: no cudaMalloc / cudaHostAlloc inside it
            D.scalars.ensure((size_t)4096 * 32);
            D.bases_jac.ensure((size_t)4096 * 192);
            D.window_sums.ensure((size_t)64 * 256);
            D.totals.ensure(32);
            D.ensure_pinned((size_t)64 * 256 + 1024);
            phase("streams, events, small-path buffers");
            preload_small_path<Fq>();
            preload_small_path<Fq2>();
            phase("small-path kernels loaded");
        }
    } catch (const CudaError &e2) {
        g_devs.clear();
        return fail(B200_ERR_CUDA, "%s", e2.msg.c_str());
    }
    g_init = true;
    return B200_OK;
}


}  // namespace eng
}  // namespace b200

using namespace b200;
using namespace b200::eng;

// ------------------------------------------------------------------------------
// C-ABI
// ------------------------------------------------------------------------------
template <class HF>
static int sum_partials(const uint64_t *pts, size_t n, uint64_t *out)
{
    typedef host::HJac<HF> J;
    if (!out || (n && !pts)) return fail(B200_ERR_ARG, "null argument");
    J acc = J::inf();
    for (size_t i = 0; i < n; i++) {
        J p;
        memcpy(&p, pts + i * (sizeof(J) / 8), sizeof(J));
        acc = host::jac_add(acc, p);
    }
    const J nrm = host::jac_normalise(acc);
    memcpy(out, &nrm, sizeof nrm);
    return B200_OK;
}

static int apply_tuning(const std::string &k, int value)
{
    const char *key = k.c_str();
    if (k == "reduce_log_segment") g_tune_logS = value;       // -1 = model
    else if (k == "reduce_split") g_tune_split = value;       // 0 = auto
    else if (k == "use_precomputed") g_tune_pre = value;      // 0: ignore precomputed levels of a key
    else if (k == "host_horner") g_tune_host_horner = value;  // 0: the device also weights and sums the per-job results of a one-window reduction
    else if (k == "reduce_marginals") g_tune_marginals = value;  // 0: bit decomposition over all segments (round 1)
    else if (k == "batch_affine") g_tune_ba = std::min(std::max(value, 0), 2);  // tree levels of affine pair additions before the XYZZ tail
    else if (k == "reduce_block") g_tune_red_block = value;  // threads per block of k_reduce_segments
    else if (k == "pinned_chunks") g_tune_pinned_chunks = value;  // upload chunks of a resident-key MSM's host scalars (0 = auto, 1 = one upload)
    else if (k == "overlap_sort") g_tune_overlap_sort = value;  // 0: pipelined MSMs sort and accumulate every chunk on one stream
    else if (k == "dense_direct") g_tune_dense_direct = value;  // 0: pipelined MSMs fold every bucket after every chunk
    else if (k == "reduce_quads") g_tune_quads = value;       // 0: one thread per partial sum in stage 2 of the window reduction (k_reduce_bits)
    else if (k == "g1_paired") g_tune_g1paired = value;      // 1 / 2: paired products in k_accumulate<Fq> (2: compiled for four blocks per SM)
    else if (k == "g2_blocks") g_tune_g2blocks = value;      // 3: the 168-register build of k_accumulate<Fq2> (three blocks per SM)
    else if (k == "g2_lane_pairs") g_tune_g2pair = value;    // 0: k_accumulate<Fq2> with one thread per task
    else if (k == "even_chunks") g_tune_even_chunks = value;  // 0: short first upload chunk (measured: no gain, profiles/r2w_e2e_chunk_probe.jsonl)
    else if (k == "partition_sort") g_tune_sort = value;     // 0: the round-1 global-atomics counting sort
    else if (k == "ones_filter") g_tune_ones = value;         // 0: scalars equal to one go through the sort like any other
    else return fail(B200_ERR_ARG, "unknown tuning key %s", key);
    return B200_OK;
}

// B200_TUNE="key=value,key=value": tuning knobs of b200_set_tuning_ex for processes that cannot call it (the C++ drivers,
// profiler runs); unknown keys make b200_init fail loudly.
static int apply_env_tuning()
{
    const char *e = std::getenv("B200_TUNE");
    if (!e || !*e) return B200_OK;
    std::string s(e);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t end = s.find(',', pos);
        if (end == std::string::npos) end = s.size();
        const std::string kv = s.substr(pos, end - pos);
        const size_t eq = kv.find('=');
        if (eq == std::string::npos) return fail(B200_ERR_ARG, "B200_TUNE: expected key=value, got %s", kv.c_str());
        const int rc = apply_tuning(kv.substr(0, eq), std::atoi(kv.c_str() + eq + 1));
        if (rc != B200_OK) return rc;
        pos = end + 1;
    }
    return B200_OK;
}

// One object per C-ABI call: holds the engine lock and, with B200_TRACE=1 in the environment, prints the call's wall time and
// the size / transfer counters of the last pipeline to stderr when it returns (what a maintainer reaches for when a whole
// prover is slower than the sum of its kernels: first-call allocations, staging, host tails).
struct ApiScope {
    std::lock_guard<std::mutex> lk;
    const char *fn;
    std::chrono::steady_clock::time_point t0;
    explicit ApiScope(const char *f) : lk(g_mu), fn(f)
    {
        if (g_trace) t0 = std::chrono::steady_clock::now();
    }
    ~ApiScope()
    {
        if (!g_trace) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "[b200] %-28s %9.3f ms  (last pipeline: n=%llu c=%u W=%u device %.3f ms, h2d %.1f MB, d2h %.3f MB)\n", fn, ms,
                (unsigned long long)g_stats.n, g_stats.window_bits, g_stats.num_windows, g_stats.device_ms, g_stats.h2d_bytes / 1e6,
                g_stats.d2h_bytes / 1e6);
    }
};

static std::thread g_async_thread;
static int g_async_rc = B200_OK;
static bool g_async_started = false;

static void join_async_init()
{
    if (g_async_thread.joinable()) g_async_thread.join();
}

extern "C" {

int b200_init(int n_gpus)
{
    ApiScope lk(__func__);
    if (g_async_started && !g_init && g_async_rc != B200_OK) return g_async_rc;  // the background start failed: its message is in g_err
    const int rc = init_devices(nullptr, n_gpus);
    return rc == B200_OK ? apply_env_tuning() : rc;
}

// b200_init on a background thread.  The thread takes the engine lock before b200_init_async returns (the caller cannot
// overtake it), so any later entry point blocks on g_mu until the devices are up; a failure stays in g_async_rc / g_err and
// is returned by the next b200_init.
int b200_init_async(int n_gpus)
{
    static std::mutex start_mu;
    std::lock_guard<std::mutex> once(start_mu);
    if (g_async_started || g_init) return B200_OK;
    g_async_started = true;
    std::mutex handoff_mu;
    std::condition_variable handoff_cv;
    bool locked = false;
    g_async_thread = std::thread([&, n_gpus] {
        std::unique_lock<std::mutex> lk(g_mu);
        {
            std::lock_guard<std::mutex> h(handoff_mu);
            locked = true;
            handoff_cv.notify_one();  // under the lock: the caller's frame (which owns the condition variable) outlives this call
        }
        int rc = init_devices(nullptr, n_gpus);
        if (rc == B200_OK) rc = apply_env_tuning();
        g_async_rc = rc;
    });
    std::unique_lock<std::mutex> h(handoff_mu);
    handoff_cv.wait(h, [&] { return locked; });
    static const int registered = std::atexit(join_async_init);  // a program that exits early must not leave the thread behind
    (void)registered;
    return B200_OK;
}

int b200_init_devices(const int *device_ids, int n)
{
    ApiScope lk(__func__);
    if (!device_ids || n <= 0) return fail(B200_ERR_ARG, "device list is empty");
    const int rc = init_devices(device_ids, n);
    return rc == B200_OK ? apply_env_tuning() : rc;
}

void b200_shutdown(void)
{
    join_async_init();
    ApiScope lk(__func__);
    if (!g_init) return;
    for (auto &kv : g_pinned)
        for (auto &s : kv.second->shards) {
            cudaSetDevice(g_devs[s.dev].id);
            cudaFree(s.d_aff);
            cudaFree(s.d_flags);
            cudaFree(s.d_pre);
        }
    g_pinned.clear();
    for (auto &kv : g_tables)
        for (size_t di = 0; di < kv.second->d_table.size(); di++) {
            cudaSetDevice(g_devs[di].id);
            cudaFree(kv.second->d_table[di]);
        }
    g_tables.clear();
    fr_release();
    release_stagers();
    for (auto &d : g_devs) d.release();
    g_devs.clear();
    g_init = false;
}

int b200_device_count(void) { return g_init ? (int)g_devs.size() : 0; }
const char *b200_last_error(void) { return g_err.c_str(); }
const char *b200_version(void) { return "b200-msm 0.1 (sm_100a)"; }

int b200_msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t out[12])
{
    ApiScope lk(__func__);
    return msm_host<Fq>(bases, scalars, n, out);
}
int b200_msm_g2(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t out[24])
{
    ApiScope lk(__func__);
    return msm_host<Fq2>(bases, scalars, n, out);
}

int b200_msm_g2g1(const uint64_t *g2_bases, const uint64_t *g1_bases, const uint64_t *scalars, size_t n, uint64_t out_g2[24],
                  uint64_t out_g1[12])
{
    ApiScope lk(__func__);
    if (!out_g2 || !out_g1 || (n && (!g2_bases || !g1_bases || !scalars))) return fail(B200_ERR_ARG, "null argument");
    int rc = msm_host<Fq2>(g2_bases, scalars, n, out_g2);
    if (rc != B200_OK) return rc;
    const b200_stats_t first = g_stats;
    g_scalars_resident = n > 0;  // same n, same devices => same shard ranges: D.scalars is still valid
    rc = msm_host<Fq>(g1_bases, scalars, n, out_g1);
    g_scalars_resident = false;
    if (rc == B200_OK) {  // stats describe the pair; geometry fields are the G1 half's
        g_stats.kernel_launches += first.kernel_launches;
        g_stats.h2d_bytes += first.h2d_bytes;
        g_stats.d2h_bytes += first.d2h_bytes;
        g_stats.host_finalize_us += first.host_finalize_us;
        g_stats.device_ms += first.device_ms;
    }
    return rc;
}

int b200_msm_batch_g1(const uint64_t *bases, const uint64_t *scalars, const uint64_t *offsets, size_t count, uint64_t *out)
{
    ApiScope lk(__func__);
    return msm_batch<Fq>(bases, scalars, offsets, count, out);
}
int b200_msm_batch_g2(const uint64_t *bases, const uint64_t *scalars, const uint64_t *offsets, size_t count, uint64_t *out)
{
    ApiScope lk(__func__);
    return msm_batch<Fq2>(bases, scalars, offsets, count, out);
}

int b200_sum_partials_g1(const uint64_t *pts, size_t n, uint64_t out[12]) { return sum_partials<host::HFq>(pts, n, out); }
int b200_sum_partials_g2(const uint64_t *pts, size_t n, uint64_t out[24]) { return sum_partials<host::HFq2>(pts, n, out); }

int b200_pin_bases_g1(const uint64_t *bases, size_t n, uint64_t *handle)
{
    ApiScope lk(__func__);
    return pin_bases<Fq>(bases, nullptr, n, handle);
}
int b200_pin_bases_g2(const uint64_t *bases, size_t n, uint64_t *handle)
{
    ApiScope lk(__func__);
    return pin_bases<Fq2>(bases, nullptr, n, handle);
}
int b200_pin_affine_dev_g1(const void *d_affine, size_t n, uint64_t *handle)
{
    ApiScope lk(__func__);
    return pin_bases<Fq>(nullptr, d_affine, n, handle);
}
int b200_pin_affine_dev_g2(const void *d_affine, size_t n, uint64_t *handle)
{
    ApiScope lk(__func__);
    return pin_bases<Fq2>(nullptr, d_affine, n, handle);
}
int b200_unpin_bases(uint64_t handle)
{
    ApiScope lk(__func__);
    auto it = g_pinned.find(handle);
    if (it == g_pinned.end()) return fail(B200_ERR_ARG, "unknown bases handle");
    for (auto &s : it->second->shards) {
        cudaSetDevice(g_devs[s.dev].id);
        cudaFree(s.d_aff);
        cudaFree(s.d_flags);
        cudaFree(s.d_pre);
    }
    g_pinned.erase(it);
    return B200_OK;
}
int b200_key_precompute_g1(uint64_t handle, uint32_t window_bits)
{
    ApiScope lk(__func__);
    return key_precompute<Fq>(handle, window_bits);
}
int b200_key_precompute_g2(uint64_t handle, uint32_t window_bits)
{
    ApiScope lk(__func__);
    return key_precompute<Fq2>(handle, window_bits);
}
int b200_msm_pinned_g1(uint64_t handle, size_t offset, const uint64_t *scalars, size_t n, uint64_t out[12])
{
    ApiScope lk(__func__);
    return msm_pinned<Fq>(handle, offset, scalars, nullptr, n, nullptr, out);
}
int b200_msm_pinned_g2(uint64_t handle, size_t offset, const uint64_t *scalars, size_t n, uint64_t out[24])
{
    ApiScope lk(__func__);
    return msm_pinned<Fq2>(handle, offset, scalars, nullptr, n, nullptr, out);
}
int b200_msm_pinned_dev_g1(uint64_t handle, size_t offset, const void *d_scalars, size_t n, void *stream, uint64_t out[12])
{
    ApiScope lk(__func__);
    if (n && !d_scalars) return fail(B200_ERR_ARG, "null device scalars");
    return msm_pinned<Fq>(handle, offset, nullptr, d_scalars, n, stream, out);
}
int b200_msm_pinned_dev_g2(uint64_t handle, size_t offset, const void *d_scalars, size_t n, void *stream, uint64_t out[24])
{
    ApiScope lk(__func__);
    if (n && !d_scalars) return fail(B200_ERR_ARG, "null device scalars");
    return msm_pinned<Fq2>(handle, offset, nullptr, d_scalars, n, stream, out);
}

size_t b200_exp_window_size_g1(size_t num_scalars) { return libff_window_size(G1_WTAB, num_scalars); }
size_t b200_exp_window_size_g2(size_t num_scalars) { return libff_window_size(G2_WTAB, num_scalars); }

int b200_batch_exp_g1(const uint64_t base[12], const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out)
{
    ApiScope lk(__func__);
    return batch_exp_once<Fq>(base, scalars, n, coeff, out);
}
int b200_batch_exp_g2(const uint64_t base[24], const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out)
{
    ApiScope lk(__func__);
    return batch_exp_once<Fq2>(base, scalars, n, coeff, out);
}
int b200_window_table_create_g1(const uint64_t base[12], size_t expected, uint64_t *handle)
{
    ApiScope lk(__func__);
    return table_create<Fq>(base, expected, handle);
}
int b200_window_table_create_g2(const uint64_t base[24], size_t expected, uint64_t *handle)
{
    ApiScope lk(__func__);
    return table_create<Fq2>(base, expected, handle);
}
int b200_window_table_destroy(uint64_t handle)
{
    ApiScope lk(__func__);
    auto it = g_tables.find(handle);
    if (it == g_tables.end()) return fail(B200_ERR_ARG, "unknown table handle");
    for (size_t di = 0; di < it->second->d_table.size(); di++) {
        cudaSetDevice(g_devs[di].id);
        cudaFree(it->second->d_table[di]);
    }
    g_tables.erase(it);
    return B200_OK;
}
int b200_batch_exp_table_g1(uint64_t handle, const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out)
{
    ApiScope lk(__func__);
    return batch_exp_table<Fq>(handle, scalars, nullptr, n, coeff, out, nullptr, nullptr);
}
int b200_batch_exp_table_g2(uint64_t handle, const uint64_t *scalars, size_t n, const uint64_t *coeff, uint64_t *out)
{
    ApiScope lk(__func__);
    return batch_exp_table<Fq2>(handle, scalars, nullptr, n, coeff, out, nullptr, nullptr);
}
int b200_batch_exp_table_dev_g1(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *stream)
{
    ApiScope lk(__func__);
    if (n && !d_scalars) return fail(B200_ERR_ARG, "null device scalars");
    return batch_exp_table<Fq>(handle, nullptr, d_scalars, n, nullptr, nullptr, d_out, stream);
}
int b200_batch_exp_table_dev_g2(uint64_t handle, const void *d_scalars, size_t n, void *d_out, void *stream)
{
    ApiScope lk(__func__);
    if (n && !d_scalars) return fail(B200_ERR_ARG, "null device scalars");
    return batch_exp_table<Fq2>(handle, nullptr, d_scalars, n, nullptr, nullptr, d_out, stream);
}

int b200_fr_fold_witness(const uint64_t *v, const uint64_t *r, size_t d, uint64_t *w_coeffs, uint64_t eval[4])
{
    ApiScope lk(__func__);
    return fr_fold_witness(v, r, d, w_coeffs, eval);
}
int b200_fr_eval_mle(const uint64_t *v, const uint64_t *r, size_t d, uint64_t out[4])
{
    ApiScope lk(__func__);
    if (!out) return fail(B200_ERR_ARG, "null argument");
    return fr_fold_witness(v, r, d, nullptr, out);
}
int b200_fr_mle_bind(const uint64_t *table, size_t half, const uint64_t r[4], uint64_t *out)
{
    ApiScope lk(__func__);
    return fr_mle_bind(table, half, r, out);
}
int b200_cppoly_prove_g1(uint64_t key, const uint64_t *v, const uint64_t *r, size_t d, uint64_t *witness, uint64_t eval[4])
{
    ApiScope lk(__func__);
    return cppoly_prove_g1(key, v, r, d, witness, eval);
}
int b200_qap_h_coefficients(const uint64_t *aA, const uint64_t *aB, const uint64_t *aC, size_t log_big, size_t log_small,
                            const uint64_t coset_g[4], const uint64_t *div_consts, uint64_t *H)
{
    ApiScope lk(__func__);
    return fr_qap_h(aA, aB, aC, log_big, log_small, coset_g, div_consts, H);
}
int b200_fr_step_fft(uint64_t *a, size_t log_big, size_t log_small, int mode, const uint64_t *coset_g)
{
    ApiScope lk(__func__);
    return fr_step_fft(a, log_big, log_small, mode, coset_g);
}
int b200_fr_geometric_quotients(uint64_t *out, const uint64_t *in, size_t n, const uint64_t *consts, size_t n_factors)
{
    ApiScope lk(__func__);
    return fr_geometric_quotients(out, in, n, consts, n_factors);
}
int b200_fr_scale_inv_geometric(uint64_t *P, size_t n_geo, const uint64_t c1[4], const uint64_t ratio[4], const uint64_t c0[4], size_t n_tail,
                                const uint64_t *tail)
{
    ApiScope lk(__func__);
    return fr_scale_inv_geometric(P, n_geo, c1, ratio, c0, n_tail, tail);
}
int b200_fr_eq_table(const uint64_t *r, size_t d, uint64_t *out)
{
    ApiScope lk(__func__);
    return fr_eq_table(r, d, out);
}
int b200_fr_matrix_mle(const uint64_t *A, const uint64_t *rho, size_t d, uint64_t *v)
{
    ApiScope lk(__func__);
    return fr_matrix_mle(A, rho, d, v);
}
int b200_fr_sumcheck_round(const uint64_t *a, const uint64_t *b, const uint64_t *w, size_t half, uint64_t out[12])
{
    ApiScope lk(__func__);
    return fr_sumcheck_round(a, b, w, half, out);
}
int b200_fr_sumcheck_rounds(const uint64_t *a, const uint64_t *b, const uint64_t *r, size_t d, uint64_t *h)
{
    ApiScope lk(__func__);
    return fr_sumcheck_rounds(a, b, r, d, h);
}
int b200_fr_fft(uint64_t *a, size_t log_n, int mode, const uint64_t *coset_g)
{
    ApiScope lk(__func__);
    if (!a) return fail(B200_ERR_ARG, "null argument");
    return fr_fft(a, nullptr, log_n, mode, coset_g, nullptr);
}
int b200_fr_fft_dev(void *d_a, size_t log_n, int mode, const uint64_t *coset_g, void *cuda_stream)
{
    ApiScope lk(__func__);
    if (!d_a) return fail(B200_ERR_ARG, "null argument");
    return fr_fft(nullptr, d_a, log_n, mode, coset_g, cuda_stream);
}

int b200_compress_g1(const uint64_t *pts, size_t n, int flavour, uint64_t *x_out, uint8_t *flags)
{
    ApiScope lk(__func__);
    return compress_g1(pts, n, flavour, x_out, flags);
}
int b200_compress_g2(const uint64_t *pts, size_t n, int flavour, uint64_t *x_out, uint8_t *flags)
{
    ApiScope lk(__func__);
    return compress_g2(pts, n, flavour, x_out, flags);
}
int b200_decompress_g1(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out, uint8_t *bad)
{
    ApiScope lk(__func__);
    return decompress_g1(x, flags, n, flavour, pts_out, bad);
}
int b200_decompress_g2(const uint64_t *x, const uint8_t *flags, size_t n, int flavour, uint64_t *pts_out, uint8_t *bad)
{
    ApiScope lk(__func__);
    return decompress_g2(x, flags, n, flavour, pts_out, bad);
}

int b200_batch_to_affine_g1(uint64_t *pts, size_t n)
{
    ApiScope lk(__func__);
    return batch_to_affine<Fq>(pts, n);
}
int b200_batch_to_affine_g2(uint64_t *pts, size_t n)
{
    ApiScope lk(__func__);
    return batch_to_affine<Fq2>(pts, n);
}

int b200_test_field_op(int field, int op, const uint64_t *a, const uint64_t *b, size_t n, uint64_t *out)
{
    ApiScope lk(__func__);
    return test_field_op(field, op, a, b, n, out);
}

int b200_test_group_op(int group, int op, const uint64_t *a, const uint64_t *b, size_t n, uint32_t k, uint64_t *out)
{
    ApiScope lk(__func__);
    if (group == 0) return test_group_op<Fq>(op, a, b, n, k, out);
    if (group == 1) return test_group_op<Fq2>(op, a, b, n, k, out);
    return fail(B200_ERR_ARG, "bad group");
}

int b200_last_stats(b200_stats_t *out)
{
    if (!out) return B200_ERR_ARG;
    *out = g_stats;
    return B200_OK;
}

int b200_set_tuning(int window_bits, int chunk_len)
{
    ApiScope lk(__func__);
    g_tune_c = window_bits;
    g_tune_L = chunk_len;
    return B200_OK;
}

int b200_set_tuning_ex(const char *key, int value)
{
    ApiScope lk(__func__);
    if (!key) return fail(B200_ERR_ARG, "null key");
    return apply_tuning(key, value);
}


int b200_set_pipeline_chunks(int chunks)
{
    ApiScope lk(__func__);
    g_tune_chunks = chunks;
    return B200_OK;
}

int b200_imad_peak(int kind, int iters, double *ops_per_sec, double *elapsed_ms)
{
    ApiScope lk(__func__);
    if (!g_init) return fail(B200_ERR_NOT_INIT, "b200_init has not been called");
    if (!ops_per_sec || iters <= 0) return fail(B200_ERR_ARG, "bad argument");
    try {
        Device &D = g_devs[0];
        CK(cudaSetDevice(D.id));
        D.totals.ensure(16);
        const uint32_t blocks = (uint32_t)D.sms * 8, threads = 256;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int rep = 0; rep < 2; rep++) {  // first pass warms up
            CK(cudaEventRecord(e0, D.stream));
            if (kind == 0) LAUNCH(D, k_peak_imad_wide, blocks, threads, 0, D.stream, D.totals.as<uint64_t>(), (uint32_t)iters, 12345u);
            else if (kind == 1) LAUNCH(D, k_peak_imad, blocks, threads, 0, D.stream, D.totals.as<uint64_t>(), (uint32_t)iters, 12345u);
            else LAUNCH(D, k_peak_modmul, blocks, threads, 0, D.stream, D.totals.as<uint64_t>(), (uint32_t)iters, 12345u);
            CK(cudaEventRecord(e1, D.stream));
            CK(cudaStreamSynchronize(D.stream));
        }
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        const double per_thread = kind == 2 ? 2.0 * iters : kind == 0 ? 4.0 * 17.0 * iters : 8.0 * iters;
        *ops_per_sec = per_thread * blocks * threads / (ms * 1e-3);
        if (elapsed_ms) *elapsed_ms = ms;
        return B200_OK;
    } catch (const CudaError &e) {
        return fail(B200_ERR_CUDA, "%s", e.msg.c_str());
    } catch (const std::exception &ex) {  // nothing may unwind through the extern "C" boundary
        return fail(B200_ERR_CUDA, "host error: %s", ex.what());
    } catch (...) {
        return fail(B200_ERR_CUDA, "unknown host error");
    }
}

}  // extern "C"
