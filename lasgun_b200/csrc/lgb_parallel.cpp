#include "lgb_parallel.hpp"

#include <cstdlib>

namespace lgb {

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
