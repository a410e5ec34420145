// Host emulation of one CUDA warp -- TEST CODE ONLY.
//
// 32 std::threads run the same function in lock step at every warp collective: each collective is two phases of a
// 32-party barrier (publish, read).  This is enough to execute plasticinelab_b200/csrc/plb_warp.cuh -- the shared-memory
// tile scatter, its per-cell flush and the thread-level scatter kernels -- on the CPU and compare it with the direct
// per-particle scatter (the build box has no GPU).  A lane that skips a collective the others call deadlocks the warp;
// run_warp() turns that into a failure through a watchdog instead of hanging the test run.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace plb {

struct EmulWarp {
    std::barrier<> bar{32};
    int pred[32];
    int ival[32];
    double dval[32];
};
inline thread_local EmulWarp* tl_warp = nullptr;
inline thread_local int tl_lane = 0;

inline unsigned warp_ballot(bool p) {
    EmulWarp* w = tl_warp;
    w->pred[tl_lane] = p ? 1 : 0;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)w->pred[i] << i;
    w->bar.arrive_and_wait();
    return m;
}
inline int warp_shfl(int v, int src) {
    EmulWarp* w = tl_warp;
    w->ival[tl_lane] = v;
    w->bar.arrive_and_wait();
    int r = w->ival[src & 31];
    w->bar.arrive_and_wait();
    return r;
}
inline double warp_shfl_down(double v, int d) {
    EmulWarp* w = tl_warp;
    w->dval[tl_lane] = v;
    w->bar.arrive_and_wait();
    double r = (tl_lane + d < 32) ? w->dval[tl_lane + d] : v;
    w->bar.arrive_and_wait();
    return r;
}
inline float warp_shfl_down(float v, int d) { return (float)warp_shfl_down((double)v, d); }
inline double warp_shfl_xor(double v, int m) {
    EmulWarp* w = tl_warp;
    w->dval[tl_lane] = v;
    w->bar.arrive_and_wait();
    double r = w->dval[(tl_lane ^ m) & 31];
    w->bar.arrive_and_wait();
    return r;
}
inline float warp_shfl_xor(float v, int m) { return (float)warp_shfl_xor((double)v, m); }
inline std::mutex& emul_atomic_mutex() { static std::mutex m; return m; }
inline void atomic_add_f64(double* dst, double v) { std::lock_guard<std::mutex> lock(emul_atomic_mutex()); *dst += v; }
inline void warp_sync() { tl_warp->bar.arrive_and_wait(); }
inline int ctz32(unsigned g) { return g ? __builtin_ctz(g) : 32; }

// run fn(lane) on 32 lock-stepped threads; aborts the process if the warp does not finish within `timeout_s`
inline void run_warp(const std::function<void(int)>& fn, double timeout_s = 60.0) {
    EmulWarp w;
    std::atomic<int> done{0};
    std::vector<std::thread> th;
    for (int lane = 0; lane < 32; lane++)
        th.emplace_back([&, lane] { tl_warp = &w; tl_lane = lane; fn(lane); done.fetch_add(1); });
    auto t0 = std::chrono::steady_clock::now();
    while (done.load() < 32) {
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
            std::fprintf(stderr, "warp emulation: deadlock (lanes disagree on a warp collective)\n");
            std::abort();
        }
    }
    for (auto& t : th) t.join();
}

}  // namespace plb
