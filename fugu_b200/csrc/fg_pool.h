// Persistent host worker pool shared by the planner (fg_host.cpp) and the plan lowering
// (fg_api.cu). A 5000-query batch is planned and lowered in ~1 ms on one thread; starting fresh
// std::threads for every call costs almost as much as the work they take over.
#pragma once
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace fg {

class HostPool {
public:
    static HostPool& get() {
        static HostPool p;
        return p;
    }
    int size() const { return (int)workers_.size() + 1; }
    // runs fn(0) .. fn(n-1), fn(0) on the calling thread; returns when all are done.
    // Calls from different threads serialise (one job at a time).
    void run(int n, const std::function<void(int)>& fn) {
        // (a fork()ed child inherits this object but not its worker threads: run inline there)
        if (n <= 1 || workers_.empty() || getpid() != pid_) {
            for (int i = 0; i < n; i++) fn(i);
            return;
        }
        std::lock_guard<std::mutex> job(job_mu_);
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            n_ = n;
            next_ = 1;
            pending_ = n - 1;
            pending_pub_.store(n - 1, std::memory_order_relaxed);
            gen_++;
            gen_pub_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
        fn(0);
        {   // the workers usually finish within microseconds of this thread: poll before sleeping
            const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us_);
            while (pending_pub_.load(std::memory_order_acquire) != 0 && std::chrono::steady_clock::now() < t_end) {
#if defined(__x86_64__) || defined(__i386__)
                __builtin_ia32_pause();
#endif
            }
        }
        std::unique_lock<std::mutex> g(mu_);
        done_.wait(g, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() : pid_(getpid()) {
        unsigned hw = std::thread::hardware_concurrency();
        if (const char* e = getenv("FG_HOST_THREADS")) hw = (unsigned)atoi(e);
        if (const char* e = getenv("FG_POOL_SPIN_US")) spin_us_ = (unsigned)atoi(e);
        const int n = (int)std::max(1u, std::min(hw ? hw : 4u, 16u));
        for (int i = 1; i < n; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
            stop_flag_.store(true, std::memory_order_relaxed);
        }
        cv_.notify_all();
        // a fork()ed child inherits the std::thread objects but not the threads: joining them there never returns
        for (auto& t : workers_) {
            if (getpid() == pid_) t.join();
            else t.detach();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> g(mu_);
        while (true) {
            // A request runs several jobs back to back (plan, lower, assemble, per pipeline chunk): a worker that has
            // just finished one polls for the next for a short while before it goes to sleep on the condition variable
            // (waking a sleeper costs tens of microseconds per job, a third of such a job's own duration).
            if (!(stop_ || (gen_ != seen && next_ < n_))) {
                g.unlock();
                const auto t_end = std::chrono::steady_clock::now() + std::chrono::microseconds(spin_us_);
                bool ready = false;
                while (!ready && std::chrono::steady_clock::now() < t_end) {
                    for (int i = 0; i < 64 && !ready; i++) {
#if defined(__x86_64__) || defined(__i386__)
                        __builtin_ia32_pause();
#endif
                        ready = stop_flag_.load(std::memory_order_relaxed) || gen_pub_.load(std::memory_order_acquire) != seen;
                    }
                }
                g.lock();
            }
            cv_.wait(g, [&] { return stop_ || (gen_ != seen && next_ < n_); });
            if (stop_) return;
            while (next_ < n_) {
                const int i = next_++;
                const std::function<void(int)>* f = fn_;
                g.unlock();
                (*f)(i);
                g.lock();
                pending_pub_.store(pending_ - 1, std::memory_order_release);
                if (--pending_ == 0) done_.notify_all();
            }
            seen = gen_;
        }
    }
    const pid_t pid_;
    std::vector<std::thread> workers_;
    std::mutex mu_, job_mu_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* fn_ = nullptr;
    int n_ = 0, next_ = 0, pending_ = 0;
    unsigned long long gen_ = 0;
    bool stop_ = false;
    std::atomic<unsigned long long> gen_pub_{0};  // gen_, readable without the mutex (pollers)
    std::atomic<bool> stop_flag_{false};
    std::atomic<int> pending_pub_{0};
    unsigned spin_us_ = 150;
};

}  // namespace fg
