// Minimal thread pool for the host driver's data-parallel loops (row generation, reducer search,
// column mapping, criteria). The reference is single-threaded (README.md:9,71; its `-t` option is
// parsed but unused, source/main.cpp:63-67); here `-t` sets the size of this pool.
#pragma once
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace gb {

class ThreadPool {
public:
    explicit ThreadPool(int n) : n_(std::max(1, n)) {
        for (int i = 1; i < n_; ++i) workers_.emplace_back([this] { loop(); });
    }
    ~ThreadPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
        for (auto& t : workers_) t.join();
    }
    int size() const { return n_; }

    // fn(begin, end) over [0, n) in chunks of `grain`, handed out dynamically; returns when all are done.
    // The caller waits for stragglers WITHOUT the PAUSE instruction: under KVM a PAUSE loop on the critical
    // thread triggers pause-loop exits that deschedule its vCPU (measured: 10x slower regions); idle
    // workers do use PAUSE, where such an exit costs nothing.
    // Workers spin for a short while between regions (the driver's phases issue many regions of well under a
    // millisecond back to back; a condition-variable wake-up alone costs more than such a region) and
    // only then go to sleep, so they cost nothing while the GPU works.
    template <class F>
    void parallel_for(size_t n, size_t grain, F&& fn) {
        if (n == 0) return;
        grain = std::max<size_t>(grain, 1);
        if (n_ == 1 || n <= grain) { fn((size_t)0, n); return; }
        std::function<void(size_t, size_t)> f = std::ref(fn);
        job_ = &f; total_ = n; grain_ = grain;
        next_.store(0, std::memory_order_relaxed);
        busy_.store((int)workers_.size(), std::memory_order_relaxed);
        gen_.fetch_add(1, std::memory_order_release);
        if (sleepers_.load(std::memory_order_acquire) > 0) {
            std::lock_guard<std::mutex> g(m_);
            cv_.notify_all();
        }
        work(f);
        for (unsigned spins = 0; busy_.load(std::memory_order_acquire) != 0; ++spins) {
            if (spins < (1u << 20)) asm volatile("" ::: "memory"); else std::this_thread::yield();
        }
        job_ = nullptr;
    }

private:
    void work(std::function<void(size_t, size_t)>& f) {
        for (;;) {
            const size_t b = next_.fetch_add(grain_, std::memory_order_relaxed);
            if (b >= total_) break;
            f(b, std::min(total_, b + grain_));
        }
    }
    void loop() {
        unsigned long seen = 0;
        for (;;) {
            unsigned spins = 0;
            while (gen_.load(std::memory_order_acquire) == seen) {
                if (++spins < 20000) { __builtin_ia32_pause(); continue; }
                std::unique_lock<std::mutex> lk(m_);
                sleepers_.fetch_add(1, std::memory_order_release);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                sleepers_.fetch_sub(1, std::memory_order_release);
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            // every region waits for all workers to check in, so job_ is the current one here
            if (job_) work(*job_);
            busy_.fetch_sub(1, std::memory_order_release);
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    std::function<void(size_t, size_t)>* job_ = nullptr;
    std::atomic<size_t> next_{0};
    std::atomic<int> busy_{0}, sleepers_{0};
    std::atomic<unsigned long> gen_{0};
    size_t total_ = 0, grain_ = 1;
    bool stop_ = false;
};

// Sort with the pool: chunks sorted in parallel, then rounds of pairwise merges.
template <class T, class Cmp>
void parallel_sort(ThreadPool& pool, std::vector<T>& v, Cmp cmp) {
    const size_t n = v.size();
    const size_t parts = (size_t)pool.size();
    if (parts == 1 || n < 4096) { std::sort(v.begin(), v.end(), cmp); return; }
    std::vector<size_t> cut(parts + 1);
    for (size_t i = 0; i <= parts; ++i) cut[i] = n * i / parts;
    pool.parallel_for(parts, 1, [&](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], cmp);
    });
    for (size_t width = 1; width < parts; width *= 2) {
        const size_t pairs = (parts + 2 * width - 1) / (2 * width);
        pool.parallel_for(pairs, 1, [&](size_t a, size_t b) {
            for (size_t i = a; i < b; ++i) {
                const size_t lo = i * 2 * width, mid = std::min(parts, lo + width), hi = std::min(parts, lo + 2 * width);
                if (mid < hi) std::inplace_merge(v.begin() + cut[lo], v.begin() + cut[mid], v.begin() + cut[hi], cmp);
            }
        });
    }
}

}  // namespace gb
