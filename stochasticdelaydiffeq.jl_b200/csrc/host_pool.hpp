// host_pool.hpp -- a small fork-join pool of host threads (header only).  Used by the host-output path into PAGEABLE memory:
// the saved paths arrive in page-locked bounce buffers and are scattered into the caller's array by several threads at once
// (first-touch page faults and the copy itself are what bounds that path, not PCIe).
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace sdde {

class HostPool {
public:
    explicit HostPool(int n_workers) {
        for (int i = 0; i < n_workers; ++i) workers_.emplace_back([this] { run(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
        cv_work_.notify_all();
        for (auto &t : workers_) t.join();
    }
    HostPool(const HostPool &) = delete;
    HostPool &operator=(const HostPool &) = delete;

    // fn(i) for i in [0, n), on the workers and on the calling thread; returns when all are done
    void parallel_for(int64_t n, const std::function<void(int64_t)> &fn) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> lk(m_);
            fn_ = &fn; n_ = n; next_.store(0); busy_ = (int)workers_.size(); ++gen_;
        }
        cv_work_.notify_all();
        drain();
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [this] { return busy_ == 0; });
        fn_ = nullptr;
    }
    int size() const { return (int)workers_.size() + 1; }

private:
    void drain() {
        for (;;) {
            const int64_t i = next_.fetch_add(1);
            if (i >= n_) break;
            (*fn_)(i);
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--busy_ == 0) cv_done_.notify_one();
            }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::function<void(int64_t)> *fn_ = nullptr;
    int64_t n_ = 0;
    std::atomic<int64_t> next_{0};
    int busy_ = 0;
    uint64_t gen_ = 0;
    bool stop_ = false;
};

}  // namespace sdde
