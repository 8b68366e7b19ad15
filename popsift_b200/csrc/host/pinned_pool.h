// Pool of page-locked host blocks shared by FeaturesHost (results) and SiftJob (input copies).
#pragma once
#include "popsift_b200.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

namespace popsift { namespace detail {

// Result arrays are page-locked so that ps_download can DMA straight into them (the reference
// registers and un-registers pageable arrays around every download, features.cu:86-111, which costs
// milliseconds per image).  Page-locking is slow too, so blocks are recycled through a small pool of
// power-of-two size classes; whatever is still pooled at exit is left to the OS.
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void*> free_blocks;
    std::map<void*, size_t> size_of;

    static size_t size_class(size_t bytes)
    {
        size_t c = 64 * 1024;
        while (c < bytes) c <<= 1;
        return c;
    }
    void* get(size_t bytes)
    {
        const size_t c = size_class(bytes);
        {
            std::lock_guard<std::mutex> g(mu);
            auto it = free_blocks.find(c);
            if (it != free_blocks.end()) { void* p = it->second; free_blocks.erase(it); return p; }
        }
        void* p = ps_host_alloc(c);
        if (!p) {   // no CUDA context yet / out of pinned memory: plain page-aligned memory still works (staged copy)
            p = std::aligned_alloc(4096, c);
            if (!p) return nullptr;
            std::lock_guard<std::mutex> g(mu);
            size_of[p] = 0;
            return p;
        }
        std::lock_guard<std::mutex> g(mu);
        size_of[p] = c;
        return p;
    }
    void put(void* p)
    {
        if (!p) return;
        std::lock_guard<std::mutex> g(mu);
        auto it = size_of.find(p);
        if (it == size_of.end()) return;
        if (it->second == 0) { size_of.erase(it); std::free(p); return; }
        free_blocks.emplace(it->second, p);
    }
};

inline PinnedPool& pinned_pool()
{
    static PinnedPool* p = new PinnedPool;   // intentionally leaked: outlives the CUDA runtime teardown order
    return *p;
}

// The caller's image is pageable; enqueue() copies it into a page-locked block (like the reference, popsift.cpp:392-395).
// One core moves 8 MB in about a millisecond -- as long as a 4K frame takes on the GPU -- so large images are copied by a few
// helper threads in parallel (POPSIFT_B200_COPY_THREADS, default 4; 1 = plain memcpy).  A caller enqueues its images back to
// back, so a helper spins for a short while after a copy before it goes to sleep on the condition variable: waking a
// sleeping thread costs more than the copy it is woken for.
// One chunk of the copy.  The destination is a page-locked block that the GPU's copy engine reads next and the CPU never
// does: streaming (non-temporal) stores skip the read-for-ownership of every destination line -- a third less host-memory
// traffic per image, which is what bounds eight GPUs' worth of pageable frames -- and keep the frame out of the caches.
inline void stream_copy(unsigned char* dst, const unsigned char* src, size_t n)
{
#if defined(__x86_64__) || defined(_M_X64)
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && n >= 4096) {
        const size_t blocks = n / 64;
        for (size_t b = 0; b < blocks; ++b) {
            const __m128i v0 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 4 * b);
            const __m128i v1 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 4 * b + 1);
            const __m128i v2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 4 * b + 2);
            const __m128i v3 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src) + 4 * b + 3);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 4 * b, v0);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 4 * b + 1, v1);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 4 * b + 2, v2);
            _mm_stream_si128(reinterpret_cast<__m128i*>(dst) + 4 * b + 3, v3);
        }
        _mm_sfence();
        if (n > blocks * 64) std::memcpy(dst + blocks * 64, src + blocks * 64, n - blocks * 64);
        return;
    }
#endif
    std::memcpy(dst, src, n);
}

class ParallelCopy {
public:
    void copy(void* dst, const void* src, size_t n)
    {
        const int parts = threads_;
        if (parts <= 1 || n < (size_t)(2u << 20)) { stream_copy(static_cast<unsigned char*>(dst), static_cast<const unsigned char*>(src), n); return; }
        std::lock_guard<std::mutex> one_at_a_time(call_mu_);
        start_helpers();
        const size_t chunk = ((n + parts - 1) / parts + 4095) & ~(size_t)4095;
        dst_ = static_cast<unsigned char*>(dst); src_ = static_cast<const unsigned char*>(src); n_ = n; chunk_ = chunk;
        pending_.store(parts - 1, std::memory_order_relaxed);
        generation_.fetch_add(1);                                       // seq_cst: ordered against the sleepers_ load below
        if (sleepers_.load() > 0) {
            std::lock_guard<std::mutex> g(mu_);
            cv_.notify_all();
        }
        stream_copy(dst_, src_, std::min(chunk, n));                     // part 0 on the calling thread
        for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0; ++spins)
            if (spins > 2000) std::this_thread::yield();
    }
    static ParallelCopy& instance()
    {
        static ParallelCopy* p = new ParallelCopy;   // leaked like the pool: the helpers are detached
        return *p;
    }
private:
    ParallelCopy()
    {
        const char* e = std::getenv("POPSIFT_B200_COPY_THREADS");
        int t = e ? std::atoi(e) : 4;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && t > hw) t = hw;
        threads_ = std::max(1, std::min(t, 16));
    }
    void start_helpers()
    {
        if (started_) return;
        started_ = true;
        for (int k = 1; k < threads_; ++k)
            std::thread([this, k] {
                unsigned long long seen = 0;
                for (;;) {
                    // spin ~100 us for the next copy of a burst, then sleep
                    const auto t0 = std::chrono::steady_clock::now();
                    unsigned polls = 0;
                    while (generation_.load(std::memory_order_acquire) == seen) {
                        if ((++polls & 255u) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(100)) {
                            std::unique_lock<std::mutex> lk(mu_);
                            sleepers_.fetch_add(1);
                            cv_.wait(lk, [&] { return generation_.load() != seen; });
                            sleepers_.fetch_sub(1);
                            break;
                        }
                    }
                    seen = generation_.load(std::memory_order_acquire);
                    const size_t off = chunk_ * (size_t)k;
                    if (off < n_) stream_copy(dst_ + off, src_ + off, std::min(chunk_, n_ - off));
                    pending_.fetch_sub(1, std::memory_order_release);
                }
            }).detach();
    }
    std::mutex call_mu_, mu_;
    std::condition_variable cv_;
    unsigned char* dst_ = nullptr;
    const unsigned char* src_ = nullptr;
    size_t n_ = 0, chunk_ = 0;
    std::atomic<int> pending_{0}, sleepers_{0};
    std::atomic<unsigned long long> generation_{0};
    int threads_ = 1;
    bool started_ = false;
};

inline void parallel_copy(void* dst, const void* src, size_t n) { ParallelCopy::instance().copy(dst, src, n); }

}} // namespace popsift::detail
