// Per-device cache of device allocations: batches of the same shape (the common case:
// a caller streaming chunks, the benchmark's end-to-end loop) reuse their buffers
// instead of paying cudaMalloc/cudaFree for tens of gigabytes on every call.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <map>
#include <mutex>
#include <vector>

namespace bt {

class DevicePool {
public:
    static DevicePool &instance() {
        static DevicePool pool;
        return pool;
    }
    // size classes: 2 MiB granularity below 1 GiB, 64 MiB above
    static size_t round_up(size_t bytes) {
        if (bytes == 0) bytes = 1;
        const size_t g = bytes < (1ull << 30) ? (2ull << 20) : (64ull << 20);
        return (bytes + g - 1) / g * g;
    }
    cudaError_t acquire(void **ptr, size_t bytes) {
        int dev = 0;
        cudaGetDevice(&dev);
        const size_t size = round_up(bytes);
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto &bucket = free_[{dev, size}];
            if (!bucket.empty()) {
                *ptr = bucket.back();
                bucket.pop_back();
                cached_bytes_ -= size;
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(ptr, size);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            trim();
            e = cudaMalloc(ptr, size);
        }
        return e;
    }
    void release(void *ptr, size_t bytes) {
        if (!ptr) return;
        int dev = 0;
        cudaGetDevice(&dev);
        const size_t size = round_up(bytes);
        std::lock_guard<std::mutex> lock(mutex_);
        free_[{dev, size}].push_back(ptr);
        cached_bytes_ += size;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mutex_);
        for (auto &kv : free_)
            for (void *p : kv.second) cudaFree(p);
        free_.clear();
        cached_bytes_ = 0;
    }
    size_t cached_bytes() const { return cached_bytes_; }

private:
    std::mutex mutex_;
    std::map<std::pair<int, size_t>, std::vector<void *>> free_;
    std::atomic<size_t> cached_bytes_{0};
};

// Page-locked host staging buffers, cached the same way (cudaHostAlloc costs milliseconds).
class PinnedPool {
public:
    static PinnedPool &instance() {
        static PinnedPool pool;
        return pool;
    }
    cudaError_t acquire(void **ptr, size_t bytes) {
        const size_t size = DevicePool::round_up(bytes);
        {
            std::lock_guard<std::mutex> lock(mutex_);
            auto &bucket = free_[size];
            if (!bucket.empty()) {
                *ptr = bucket.back();
                bucket.pop_back();
                return cudaSuccess;
            }
        }
        return cudaHostAlloc(ptr, size, cudaHostAllocDefault);
    }
    void release(void *ptr, size_t bytes) {
        if (!ptr) return;
        std::lock_guard<std::mutex> lock(mutex_);
        free_[DevicePool::round_up(bytes)].push_back(ptr);
    }

private:
    std::mutex mutex_;
    std::map<size_t, std::vector<void *>> free_;
};

struct PinnedBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    PinnedBuf() = default;
    PinnedBuf(const PinnedBuf &) = delete;
    PinnedBuf &operator=(const PinnedBuf &) = delete;
    ~PinnedBuf() { if (ptr) PinnedPool::instance().release(ptr, bytes); }
    cudaError_t alloc(size_t n) {
        if (ptr) PinnedPool::instance().release(ptr, bytes);
        ptr = nullptr;
        if (n == 0) n = 1;
        cudaError_t e = PinnedPool::instance().acquire(&ptr, n);
        bytes = e == cudaSuccess ? n : 0;
        return e;
    }
    template <class T> T *as() const { return (T *)ptr; }
};

// RAII device buffer backed by the pool.
struct DevBuf {
    void *ptr = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (ptr) DevicePool::instance().release(ptr, bytes);
        ptr = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 1;
        cudaError_t e = DevicePool::instance().acquire(&ptr, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T> T *as() const { return (T *)ptr; }
};

}  // namespace bt
