// Parallel staged copies between PAGEABLE host buffers and device memory.
//
// The reference's boundary hands over plain numpy arrays (Par/pyflow.pyx:31-52): 2 x 49.8 MB of float64 in and
// 83 MB out per 1920x1080 pair, in pageable memory.  cudaMemcpy from pageable memory is staged by the driver through
// one bounce buffer on the calling thread at the speed of a single-threaded memcpy, which made the copies of a one-shot
// call cost more than the solve.  Here several host threads move 4 MB chunks through their own pinned slots and
// streams: the host-side memcpy of one chunk overlaps the DMA of another, and the host bandwidth of several cores is
// used.  Buffers that are already pinned (pf_host_alloc, cudaHostRegister) bypass this and are copied directly.
// One stager per device, created on first use; a mutex serialises staged transfers on a device.
#pragma once
#include <sched.h>
#include <atomic>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace pf {

struct CopyJob {
    void* dev;
    void* host;
    size_t bytes;
};

inline bool host_is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

class HostStager {
  public:
    static constexpr size_t kChunk = 4u << 20;

    static HostStager* for_device(int dev) {
        static std::mutex mu;
        // leaked on purpose: a static destructor would make CUDA calls while the runtime is being torn down at exit
        static std::vector<std::unique_ptr<HostStager>>& all = *new std::vector<std::unique_ptr<HostStager>>();
        std::lock_guard<std::mutex> g(mu);
        if (disabled()) return nullptr;
        if ((int)all.size() <= dev) all.resize((size_t)dev + 1);
        if (!all[(size_t)dev]) {
            std::unique_ptr<HostStager> s(new HostStager(dev));
            if (!s->ok_) return nullptr;   // no pinned memory to spare: the caller falls back to plain cudaMemcpyAsync
            all[(size_t)dev] = std::move(s);
        }
        return all[(size_t)dev].get();
    }

    // host -> device.  Returns when every chunk has left the caller's buffers; `consumer` is made to wait for the DMAs.
    void to_device(const std::vector<CopyJob>& jobs, cudaStream_t consumer) { run(jobs, true, consumer); }
    // device -> host.  The copies start when the work already enqueued on `producer` has finished; returns when the
    // caller's buffers are complete.
    void to_host(const std::vector<CopyJob>& jobs, cudaStream_t producer) { run(jobs, false, producer); }

    ~HostStager() {
        cudaSetDevice(dev_);
        for (auto& w : workers_) {
            for (int s = 0; s < 2; s++) {
                if (w.slot[s]) cudaFreeHost(w.slot[s]);
                if (w.ev[s]) cudaEventDestroy(w.ev[s]);
            }
            if (w.done) cudaEventDestroy(w.done);
            if (w.st) cudaStreamDestroy(w.st);
        }
        if (ready_) cudaEventDestroy(ready_);
    }

  private:
    struct Worker {
        cudaStream_t st = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr}, done = nullptr;
        char* slot[2] = {nullptr, nullptr};
    };
    struct Chunk {
        char *dev, *host;
        size_t bytes;
    };

    static bool disabled() {
        const char* e = getenv("PF_STAGER");
        return e && !atoi(e);
    }

    explicit HostStager(int dev) : dev_(dev) {
        int n = 0;
        if (const char* e = getenv("PF_STAGER_THREADS")) n = atoi(e);
        if (n < 1) {
            // the cores this process may run on, shared between the processes of the node (one per GPU under
            // torchrun: LOCAL_WORLD_SIZE) -- 8 ranks x 8 stager threads on 32 cores was the round-1 default and
            // lost more to oversubscription than it gained
            unsigned cores = std::thread::hardware_concurrency();
            cpu_set_t set;
            if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) cores = (unsigned)CPU_COUNT(&set);
            unsigned procs = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) procs = (unsigned)std::max(1, atoi(e));
            n = (int)std::min(8u, std::max(2u, cores / (2 * procs)));
        }
        workers_.resize((size_t)n);
        if (cudaSetDevice(dev_) != cudaSuccess) return;
        for (auto& w : workers_) {
            if (cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking) != cudaSuccess) return;
            if (cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) return;
            for (int s = 0; s < 2; s++) {
                if (cudaEventCreateWithFlags(&w.ev[s], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) return;
                if (cudaHostAlloc((void**)&w.slot[s], kChunk, cudaHostAllocPortable) != cudaSuccess) {
                    cudaGetLastError();
                    return;
                }
            }
        }
        if (cudaEventCreateWithFlags(&ready_, cudaEventDisableTiming) != cudaSuccess) return;
        ok_ = true;
    }

    void run(const std::vector<CopyJob>& jobs, bool to_dev, cudaStream_t other) {
        if (!to_dev) {
            // wait for the producer OUTSIDE the lock: with several host threads (batch workers) the stager must not be
            // held for the length of one pair's solve while other threads have frames to upload
            PF_CUDA(cudaSetDevice(dev_));
            stream_wait_blocking(other);   // stager threads and the caller must not spin while other threads have copies to issue
        }
        std::lock_guard<std::mutex> g(mu_);
        PF_CUDA(cudaSetDevice(dev_));
        std::vector<Chunk> chunks;
        for (const CopyJob& j : jobs)
            for (size_t off = 0; off < j.bytes; off += kChunk)
                chunks.push_back({(char*)j.dev + off, (char*)j.host + off, std::min(kChunk, j.bytes - off)});
        if (chunks.empty()) return;
        if (!to_dev) PF_CUDA(cudaEventRecord(ready_, other));
        std::atomic<size_t> next{0};
        std::atomic<int> err{(int)cudaSuccess};
        const size_t nthreads = std::min(workers_.size(), chunks.size());
        auto body = [&](size_t wi) {
            Worker& w = workers_[wi];
            auto chk = [&](cudaError_t e) {
                if (e != cudaSuccess) err.store((int)e);
                return e == cudaSuccess;
            };
            if (!chk(cudaSetDevice(dev_))) return;
            if (to_dev) {
                int it = 0;
                for (size_t i; (i = next.fetch_add(1)) < chunks.size(); it++) {
                    const int s = it & 1;
                    if (!chk(cudaEventSynchronize(w.ev[s]))) return;   // the DMA that last used this slot (also of an earlier call)
                    memcpy(w.slot[s], chunks[i].host, chunks[i].bytes);
                    if (!chk(cudaMemcpyAsync(chunks[i].dev, w.slot[s], chunks[i].bytes, cudaMemcpyHostToDevice, w.st))) return;
                    if (!chk(cudaEventRecord(w.ev[s], w.st))) return;
                }
                chk(cudaEventRecord(w.done, w.st));
            } else {
                if (!chk(cudaStreamWaitEvent(w.st, ready_, 0))) return;
                // two chunks in flight: the DMA of the next one runs while this one is copied out of its slot
                size_t cur = next.fetch_add(1);
                int s = 0;
                if (cur < chunks.size()) {
                    if (!chk(cudaMemcpyAsync(w.slot[s], chunks[cur].dev, chunks[cur].bytes, cudaMemcpyDeviceToHost, w.st))) return;
                    if (!chk(cudaEventRecord(w.ev[s], w.st))) return;
                }
                while (cur < chunks.size()) {
                    const size_t nxt = next.fetch_add(1);
                    if (nxt < chunks.size()) {
                        if (!chk(cudaMemcpyAsync(w.slot[s ^ 1], chunks[nxt].dev, chunks[nxt].bytes, cudaMemcpyDeviceToHost, w.st))) return;
                        if (!chk(cudaEventRecord(w.ev[s ^ 1], w.st))) return;
                    }
                    if (!chk(cudaEventSynchronize(w.ev[s]))) return;
                    memcpy(chunks[cur].host, w.slot[s], chunks[cur].bytes);
                    cur = nxt;
                    s ^= 1;
                }
            }
        };
        std::vector<std::thread> th;
        for (size_t wi = 1; wi < nthreads; wi++) th.emplace_back(body, wi);
        body(0);   // the calling thread works too
        for (auto& t : th) t.join();
        if (err.load() != (int)cudaSuccess) PF_CUDA((cudaError_t)err.load());
        if (to_dev)
            for (size_t wi = 0; wi < nthreads; wi++) PF_CUDA(cudaStreamWaitEvent(other, workers_[wi].done, 0));
    }

    int dev_;
    bool ok_ = false;
    std::mutex mu_;
    std::vector<Worker> workers_;
    cudaEvent_t ready_ = nullptr;
};

}  // namespace pf
