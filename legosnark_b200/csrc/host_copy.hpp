// host_copy.hpp — large transfers between PAGEABLE host memory and the device.
//
// (COPY_THREADS below = copy_threads(), 4 unless B200_COPY_THREADS says otherwise.)
// The reference's callers hold their vectors in ordinary std::vector storage; a plain
// cudaMemcpyAsync from such memory is staged by the driver at 6-7 GB/s (measured on the B200 box:
// 12 ms for the 32 MB in + 32 MB out of a 2^20 FFT), a fraction of what PCIe 5 x16 delivers.  Here
// COPY_THREADS host threads each run their own double-buffered pipeline over an interleaved set of
// 4 MB chunks: memcpy into a pinned buffer, cudaMemcpyAsync on a private stream, next chunk.  Memory
// that is already page-locked (cudaHostAlloc / cudaHostRegister / torch pinned) and small transfers
// go straight to cudaMemcpyAsync.
//
// Semantics match the pageable cudaMemcpyAsync they replace: h2d() returns once the source has been
// read (the caller may reuse it), and `st` is ordered after the data has landed; d2h() returns when
// the destination holds the data.
#pragma once
#include <condition_variable>
#include <cstdlib>
#include <functional>

#include "engine_common.hpp"

namespace b200 {
namespace eng {

constexpr int COPY_THREADS_MAX = 8;
constexpr size_t COPY_CHUNK = 4u << 20;
// host threads per staged transfer: B200_COPY_THREADS (1..8; 0 = no staging, plain cudaMemcpyAsync), default 4
inline int copy_threads()
{
    static const int n = [] {
        const char *e = std::getenv("B200_COPY_THREADS");
        const int v = e ? std::atoi(e) : 4;
        return std::max(0, std::min(v, COPY_THREADS_MAX));
    }();
    return n;
}
constexpr size_t COPY_MIN_STAGED = 8u << 20;
// bytes per staged piece: B200_COPY_CHUNK_KB (64 .. 4096, default 4096 = the size of the pinned buffers)
inline size_t copy_chunk()
{
    static const size_t n = [] {
        const char *e = std::getenv("B200_COPY_CHUNK_KB");
        const long v = e ? std::atol(e) : 4096;
        return (size_t)std::max(64L, std::min(v, 4096L)) << 10;
    }();
    return n;
}

struct Stager {
    void *pin[COPY_THREADS_MAX][2] = {};
    cudaStream_t cs[COPY_THREADS_MAX] = {};
    cudaEvent_t ev[COPY_THREADS_MAX][2] = {};
    cudaEvent_t done[COPY_THREADS_MAX] = {};
    cudaEvent_t gate = nullptr;
    bool ready = false;
    void init()
    {
        if (ready) return;
        CK(cudaEventCreateWithFlags(&gate, cudaEventDisableTiming));
        for (int t = 0; t < COPY_THREADS_MAX; t++) {
            CK(cudaStreamCreateWithFlags(&cs[t], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&done[t], cudaEventDisableTiming));
            for (int b = 0; b < 2; b++) {
                CK(cudaHostAlloc(&pin[t][b], COPY_CHUNK, cudaHostAllocDefault));
                CK(cudaEventCreateWithFlags(&ev[t][b], cudaEventDisableTiming));
            }
        }
        ready = true;
    }
    void release()
    {
        if (!ready) return;
        for (int t = 0; t < COPY_THREADS_MAX; t++) {
            for (int b = 0; b < 2; b++) {
                cudaFreeHost(pin[t][b]);
                cudaEventDestroy(ev[t][b]);
            }
            cudaEventDestroy(done[t]);
            cudaStreamDestroy(cs[t]);
        }
        cudaEventDestroy(gate);
        ready = false;
    }
};

Stager &stager_of(Device &D);  // engine_core.cu: one per device, created on first use, freed at shutdown

// The staging threads are kept between transfers: a pipelined host-buffer MSM stages eight transfers per call (scalars
// and bases of four upload chunks), and starting four fresh std::threads for each of them cost more than 1 ms of the
// 9 ms call.  run(n, fn) executes fn(0) .. fn(n-1) on the pool's threads and returns when all have finished.
class CopyPool {
  public:
    static CopyPool &get()
    {
        static CopyPool pool;
        return pool;
    }
    void run(int n, const std::function<void(int)> &fn)
    {
        std::lock_guard<std::mutex> one_at_a_time(run_mu_);  // in-process multi-GPU: the per-device host threads take turns
        std::unique_lock<std::mutex> lk(mu_);
        while ((int)workers_.size() < n) {
            const int id = (int)workers_.size();
            workers_.emplace_back([this, id] { loop(id); });
        }
        fn_ = &fn;
        want_ = n;
        pending_ = n;
        generation_++;
        cv_work_.notify_all();
        cv_done_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }
    ~CopyPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            cv_work_.notify_all();
        }
        for (auto &t : workers_) t.join();
    }

  private:
    void loop(int id)
    {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_work_.wait(lk, [&] { return stop_ || (generation_ != seen && id < want_); });
            if (stop_) return;
            seen = generation_;
            const std::function<void(int)> *fn = fn_;
            lk.unlock();
            (*fn)(id);
            lk.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::mutex mu_, run_mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<std::thread> workers_;
    const std::function<void(int)> *fn_ = nullptr;
    int want_ = 0, pending_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

inline bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

template <bool ToDevice>
inline void staged_copy(Device &D, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    Stager &S = stager_of(D);
    S.init();
    // the private streams start after whatever `st` still has in flight on these buffers
    CK(cudaEventRecord(S.gate, st));
    const size_t CHUNK = copy_chunk();
    const size_t nchunks = (bytes + CHUNK - 1) / CHUNK;
    const int COPY_THREADS = copy_threads();
    std::vector<std::string> errs(COPY_THREADS);
    const int dev_id = D.id;
    CopyPool::get().run(COPY_THREADS, [&](int t) {
            try {
                CK(cudaSetDevice(dev_id));
                CK(cudaStreamWaitEvent(S.cs[t], S.gate, 0));
                int b = 0;
                size_t pending_off[2] = {0, 0}, pending_len[2] = {0, 0};
                for (size_t c = (size_t)t; c < nchunks; c += COPY_THREADS, b ^= 1) {
                    const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
                    if (ToDevice) {
                        CK(cudaEventSynchronize(S.ev[t][b]));  // the DMA that last read this pinned buffer is done
                        memcpy(S.pin[t][b], (const char *)src + off, len);
                        CK(cudaMemcpyAsync((char *)dst + off, S.pin[t][b], len, cudaMemcpyHostToDevice, S.cs[t]));
                        CK(cudaEventRecord(S.ev[t][b], S.cs[t]));
                    } else {
                        if (pending_len[b]) {  // drain what this buffer received two chunks ago
                            CK(cudaEventSynchronize(S.ev[t][b]));
                            memcpy((char *)dst + pending_off[b], S.pin[t][b], pending_len[b]);
                        }
                        CK(cudaMemcpyAsync(S.pin[t][b], (const char *)src + off, len, cudaMemcpyDeviceToHost, S.cs[t]));
                        CK(cudaEventRecord(S.ev[t][b], S.cs[t]));
                        pending_off[b] = off;
                        pending_len[b] = len;
                    }
                }
                if (!ToDevice)
                    for (int k = 0; k < 2; k++, b ^= 1)
                        if (pending_len[b]) {
                            CK(cudaEventSynchronize(S.ev[t][b]));
                            memcpy((char *)dst + pending_off[b], S.pin[t][b], pending_len[b]);
                            pending_len[b] = 0;
                        }
                CK(cudaEventRecord(S.done[t], S.cs[t]));
            } catch (const CudaError &e) {
                errs[t] = e.msg;
            } catch (...) {
                errs[t] = "host error in a staging thread";
            }
        });
    for (auto &e : errs)
        if (!e.empty()) throw CudaError{e};
    for (int t = 0; t < COPY_THREADS; t++) CK(cudaStreamWaitEvent(st, S.done[t], 0));
}

inline void h2d(Device &D, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return;
    if (bytes < COPY_MIN_STAGED || copy_threads() == 0 || !is_pageable(src)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return;
    }
    staged_copy<true>(D, dst, src, bytes, st);
}

// returns with the data in dst only for the staged path; the direct path is asynchronous on `st` like
// the call it replaces (every caller synchronises `st` before touching dst)
inline void d2h(Device &D, void *dst, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes == 0) return;
    if (bytes < COPY_MIN_STAGED || copy_threads() == 0 || !is_pageable(dst)) {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        return;
    }
    staged_copy<false>(D, dst, src, bytes, st);
}

}  // namespace eng
}  // namespace b200
