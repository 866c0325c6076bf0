#include "lgb_parallel.hpp"

#include <cstdio>
#include <cstdlib>
#include <unordered_map>

namespace lgb {

namespace {
constexpr size_t kBlockMin = (size_t)1 << 20;          // smaller requests go straight to operator new
constexpr size_t kBlockRound = (size_t)2 << 20;
constexpr size_t kCacheBytes = (size_t)1 << 30;        // freed blocks kept for reuse (total) ...
constexpr size_t kCacheBlocks = 64;                    // ... and how many of them
struct BlockCache {
    std::mutex m;
    std::unordered_map<void*, size_t> live;            // capacity of every big block handed out
    std::vector<std::pair<void*, size_t>> free_list;   // (block, capacity)
    size_t cached = 0;
};
BlockCache& block_cache() { static BlockCache* c = new BlockCache(); return *c; }      // leaked on purpose, like the pool
}  // namespace

void* block_alloc(size_t bytes) {
    if (bytes < kBlockMin) return ::operator new(bytes);
    const size_t want = (bytes + kBlockRound - 1) / kBlockRound * kBlockRound;
    BlockCache& c = block_cache();
    {
        std::lock_guard<std::mutex> lk(c.m);
        size_t best = c.free_list.size();
        for (size_t i = 0; i < c.free_list.size(); i++)      // best fit among the blocks that waste at most half of themselves
            if (c.free_list[i].second >= want && c.free_list[i].second <= 2 * want && (best == c.free_list.size() || c.free_list[i].second < c.free_list[best].second)) best = i;
        if (best < c.free_list.size()) {
            const std::pair<void*, size_t> b = c.free_list[best];
            c.free_list[best] = c.free_list.back(); c.free_list.pop_back();
            c.cached -= b.second;
            c.live.emplace(b.first, b.second);
            return b.first;
        }
    }
    if (std::getenv("LGB_BLOCK_DEBUG")) std::fprintf(stderr, "[block_alloc] %zu bytes fresh\n", want);
    void* p = ::operator new(want);
    std::lock_guard<std::mutex> lk(c.m);
    c.live.emplace(p, want);
    return p;
}

void block_free(void* p, size_t bytes) noexcept {
    if (!p) return;
    if (bytes < kBlockMin) { ::operator delete(p); return; }
    BlockCache& c = block_cache();
    {
        std::lock_guard<std::mutex> lk(c.m);
        auto it = c.live.find(p);
        const size_t cap = it != c.live.end() ? it->second : 0;
        if (it != c.live.end()) c.live.erase(it);
        if (cap && c.cached + cap <= kCacheBytes && c.free_list.size() < kCacheBlocks) {
            c.free_list.emplace_back(p, cap);
            c.cached += cap;
            return;
        }
    }
    if (std::getenv("LGB_BLOCK_DEBUG")) std::fprintf(stderr, "[block_free] %zu bytes released (cache full or unknown block)\n", bytes);
    ::operator delete(p);
}

Pool& Pool::get() {
    static Pool* p = new Pool();      // leaked on purpose: the workers are detached and outlive static destruction
    return *p;
}

Pool::Pool() {
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int)(hw ? hw : 1u);
    if (const char* e = std::getenv("LGB_THREADS")) { int v = std::atoi(e); if (v > 0) n = v; }
    if (n > 32) n = 32;
    nthreads_ = n;
    for (int i = 1; i < n; i++) std::thread([this] { worker(); }).detach();
}

void Pool::worker() {
    uint64_t seen = 0;
    for (;;) {
        const std::function<void(size_t)>* fn;
        size_t n;
        {
            std::unique_lock<std::mutex> lk(m_);
            cv_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            fn = fn_; n = n_;
            inside_.fetch_add(1);          // under the lock: run() publishes only while no worker holds an old (fn, n)
        }
        size_t i;
        while ((i = next_.fetch_add(1)) < n) { (*fn)(i); done_.fetch_add(1); }
        inside_.fetch_sub(1);
    }
}

void Pool::run(size_t n, const std::function<void(size_t)>& fn) {
    if (n == 0) return;
    bool expected = false;
    if (nthreads_ == 1 || n == 1 || !busy_.compare_exchange_strong(expected, true)) {
        for (size_t i = 0; i < n; i++) fn(i);
        return;
    }
    {
        std::unique_lock<std::mutex> lk(m_);
        while (inside_.load() != 0) std::this_thread::yield();
        fn_ = &fn; n_ = n;
        done_.store(0); next_.store(0);
        generation_++;
    }
    cv_.notify_all();
    size_t i;
    while ((i = next_.fetch_add(1)) < n) { fn(i); done_.fetch_add(1); }
    while (done_.load() < n) std::this_thread::yield();
    busy_.store(false);
}

}  // namespace lgb

// the host mirror's arrays (include/lasgun_host.hpp) use the same cache
extern "C" void* lgh_block_alloc(size_t bytes) { return lgb::block_alloc(bytes); }
extern "C" void lgh_block_free(void* p, size_t bytes) { lgb::block_free(p, bytes); }
