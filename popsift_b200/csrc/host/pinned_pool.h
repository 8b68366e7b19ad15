// Pool of page-locked host blocks shared by FeaturesHost (results) and SiftJob (input copies).
#pragma once
#include "popsift_b200.h"

#include <cstdlib>
#include <map>
#include <mutex>

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


}} // namespace popsift::detail
