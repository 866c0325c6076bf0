// Process-wide worker pool for the host side of the path (reference-tree build, device-BVH build, scene
// conversion).  Workers persist between calls: a parallel region costs a wake-up, not a thread creation.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

namespace lgb {

// std::vector whose resize() leaves trivially-constructible elements uninitialised (the arrays are filled by
// all threads right after sizing; a sequential zero-fill would cost more than the fill).
// Blocks of a megabyte and more come from a process-wide cache of freed blocks (lgb_parallel.cpp): `capture(scene, film)` sizes the same
// ~100 MB of arrays for every frame, and a fresh mapping costs its page faults on first touch and an munmap on release (measured: 2 of
// the 6 ms of flattening a million triangles).
void* block_alloc(size_t bytes);
void block_free(void* p, size_t bytes) noexcept;
template <class T>
struct default_init_allocator : std::allocator<T> {
    template <class U> struct rebind { using other = default_init_allocator<U>; };
    using std::allocator<T>::allocator;
    T* allocate(size_t n) { return static_cast<T*>(block_alloc(n * sizeof(T))); }
    void deallocate(T* p, size_t n) noexcept { block_free(p, n * sizeof(T)); }
    template <class U> void construct(U* p) noexcept(std::is_nothrow_default_constructible<U>::value) { ::new (static_cast<void*>(p)) U; }
    template <class U, class... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
};
template <class T> using raw_vector = std::vector<T, default_init_allocator<T>>;

class Pool {
public:
    // Up to 32 threads (the caller included); LGB_THREADS overrides.
    static Pool& get();
    int threads() const { return nthreads_; }

    // Calls fn(i) for every i in [0, n), dynamically distributed; returns when all are done.
    // Nested or concurrent calls run serially on the calling thread.
    void run(size_t n, const std::function<void(size_t)>& fn);

    // Splits [0, n) into ~4 chunks per thread (each at least `grain` long) and calls fn(begin, end, chunk).
    // Returns the number of chunks; fn may index per-chunk scratch with `chunk`.
    size_t chunks_of(size_t n, size_t grain) const {
        if (n == 0) return 0;
        size_t want = (size_t)nthreads_ * 4, by_grain = (n + grain - 1) / grain;
        return by_grain < want ? by_grain : want;
    }
    void for_range(size_t n, size_t grain, const std::function<void(size_t, size_t, size_t)>& fn) {
        const size_t nc = chunks_of(n, grain);
        if (nc == 0) return;
        const size_t per = (n + nc - 1) / nc;
        run(nc, [&](size_t c) { size_t b = c * per, e = b + per < n ? b + per : n; if (b < e) fn(b, e, c); });
    }

private:
    Pool();
    void worker();
    int nthreads_ = 1;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t n_ = 0;
    std::atomic<size_t> next_{0};
    std::atomic<size_t> done_{0};
    std::atomic<int> inside_{0};               // workers currently holding fn_
    uint64_t generation_ = 0;
    std::atomic<bool> busy_{false};
};

}  // namespace lgb
