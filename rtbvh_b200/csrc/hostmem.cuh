// hostmem.cuh — host side of the drop-in build path: pooled page-locked arrays for the host mirrors handed out through
// RTBvh / RTMbvh, and a threaded upload of caller-owned (pageable) input arrays.
//
// Why: create_bvh / create_mbvh (rtbvh_ffi/src/lib.rs:428-513) hand HOST pointers to the caller, so every tree built on the
// GPU is copied out once.  Round 1 copied into fresh std::vectors with synchronous cudaMemcpy: 28 ms per Mtri around 2.1 ms of
// kernels (page faults on first touch + the driver's single-threaded staging of pageable memory, ~4 GB/s effective).
// Here the mirrors live in page-locked blocks that are recycled through a process-wide pool (pinning is paid once per size
// class, not per build), so the D2H runs at PCIe speed, and pageable inputs are staged by a few host threads into pinned
// slots while the copy engine drains them.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace rtb {

// Process-wide pool of page-locked blocks, keyed by capacity.  acquire() never fails for lack of pinned memory: it falls back
// to malloc (pinned = false), which every consumer handles (copies just run at pageable speed).
class HostPool {
public:
    struct Block {
        void* p = nullptr;
        size_t cap = 0;
        bool pinned = false;
    };
    static HostPool& get() {
        static HostPool* pool = new HostPool;  // leaked on purpose: blocks may be released during static destruction
        return *pool;
    }
    static size_t size_class(size_t bytes) {  // 64 KiB granules below 1 MiB, then 1/8-octave steps: <= 12.5 % slack
        if (bytes <= (size_t(1) << 16)) return size_t(1) << 16;
        size_t step = size_t(1) << 16;
        while ((step << 4) < bytes) step <<= 1;
        return (bytes + step - 1) / step * step;
    }
    Block acquire(size_t bytes) {
        Block b;
        b.cap = size_class(bytes ? bytes : 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto it = free_.find(b.cap);
            if (it != free_.end() && !it->second.empty()) {
                b.p = it->second.back();
                it->second.pop_back();
                b.pinned = true;
                cached_ -= b.cap;
                return b;
            }
        }
        if (b.cap <= kMaxPinned && cudaHostAlloc(&b.p, b.cap, cudaHostAllocPortable) == cudaSuccess) {
            b.pinned = true;
            return b;
        }
        cudaGetLastError();
        b.p = std::malloc(b.cap);
        b.pinned = false;
        return b;
    }
    void release(Block b) {
        if (!b.p) return;
        if (!b.pinned) {
            std::free(b.p);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (cached_ + b.cap <= kMaxCached) {
                free_[b.cap].push_back(b.p);
                cached_ += b.cap;
                return;
            }
        }
        cudaFreeHost(b.p);
    }
    void trim() {  // give every cached block back to the OS
        std::map<size_t, std::vector<void*>> f;
        {
            std::lock_guard<std::mutex> lk(mu_);
            f.swap(free_);
            cached_ = 0;
        }
        for (auto& kv : f)
            for (void* p : kv.second) cudaFreeHost(p);
    }

private:
    static constexpr size_t kMaxPinned = size_t(8) << 30;  // larger mirrors stay pageable
    static constexpr size_t kMaxCached = size_t(2) << 30;  // free blocks kept for reuse
    std::mutex mu_;
    std::map<size_t, std::vector<void*>> free_;
    size_t cached_ = 0;
};

// The subset of std::vector the builders use, over a pooled block.  resize() does not initialise and does not preserve.
template <class T>
class HostArray {
public:
    HostArray() = default;
    ~HostArray() { clear(); }
    HostArray(const HostArray&) = delete;
    HostArray& operator=(const HostArray&) = delete;
    HostArray(HostArray&& o) noexcept : b_(o.b_), n_(o.n_) {
        o.b_ = HostPool::Block{};
        o.n_ = 0;
    }
    void clear() {
        HostPool::get().release(b_);
        b_ = HostPool::Block{};
        n_ = 0;
    }
    bool resize(size_t n) {
        if (n * sizeof(T) > b_.cap || !b_.p) {
            HostPool::get().release(b_);
            b_ = HostPool::get().acquire(n * sizeof(T));
            if (!b_.p) {
                b_ = HostPool::Block{};
                n_ = 0;
                return false;
            }
        }
        n_ = n;
        return true;
    }
    bool assign(const T* first, const T* last) {
        if (!resize((size_t)(last - first))) return false;
        if (n_) std::memcpy(b_.p, first, n_ * sizeof(T));
        return true;
    }
    T* data() { return (T*)b_.p; }
    const T* data() const { return (const T*)b_.p; }
    size_t size() const { return n_; }
    bool empty() const { return n_ == 0; }
    bool pinned() const { return b_.pinned; }

private:
    HostPool::Block b_;
    size_t n_ = 0;
};

// H2D of a caller-owned array on `stream` (the call returns when the copy has completed).  Page-locked sources go straight to
// the copy engine; pageable ones are staged by kWorkers host threads through their own pinned slots (the driver's own
// pageable path stages on the calling thread alone).
inline cudaError_t upload_from_user(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                        (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged || attr.type == cudaMemoryTypeDevice);
    cudaGetLastError();
    constexpr size_t kSlot = size_t(4) << 20;
    constexpr int kWorkers = 4, kSlotsPerWorker = 2;
    if (pinned || bytes < 2 * kSlot) {
        cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream);
        return e != cudaSuccess ? e : cudaStreamSynchronize(stream);
    }
    int device = 0;
    cudaGetDevice(&device);
    const size_t chunks = (bytes + kSlot - 1) / kSlot;
    cudaError_t errs[kWorkers];
    auto work = [&](int w) {
        cudaError_t e = cudaSetDevice(device);
        HostPool::Block slot[kSlotsPerWorker];
        cudaEvent_t ev[kSlotsPerWorker] = {};
        bool used[kSlotsPerWorker] = {};
        for (int s = 0; s < kSlotsPerWorker && e == cudaSuccess; s++) {
            slot[s] = HostPool::get().acquire(kSlot);
            if (!slot[s].p) e = cudaErrorMemoryAllocation;
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming);
        }
        int turn = 0;
        for (size_t c = (size_t)w; c < chunks && e == cudaSuccess; c += kWorkers, turn++) {
            const int s = turn % kSlotsPerWorker;
            const size_t off = c * kSlot, len = bytes - off < kSlot ? bytes - off : kSlot;
            if (used[s]) e = cudaEventSynchronize(ev[s]);  // the slot's previous DMA has read it
            if (e != cudaSuccess) break;
            std::memcpy(slot[s].p, (const char*)src + off, len);
            e = cudaMemcpyAsync((char*)dst + off, slot[s].p, len, cudaMemcpyHostToDevice, stream);
            if (e == cudaSuccess) e = cudaEventRecord(ev[s], stream);
            used[s] = true;
        }
        for (int s = 0; s < kSlotsPerWorker; s++) {
            if (used[s]) cudaEventSynchronize(ev[s]);
            if (ev[s]) cudaEventDestroy(ev[s]);
            HostPool::get().release(slot[s]);
        }
        errs[w] = e;
    };
    std::thread th[kWorkers - 1];
    for (int w = 1; w < kWorkers; w++) th[w - 1] = std::thread(work, w);
    work(0);
    for (int w = 1; w < kWorkers; w++) th[w - 1].join();
    for (int w = 0; w < kWorkers; w++)
        if (errs[w] != cudaSuccess) return errs[w];
    return cudaStreamSynchronize(stream);
}

}  // namespace rtb
