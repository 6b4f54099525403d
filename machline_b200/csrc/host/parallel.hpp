// Host-side parallel loops of the setup (the per-panel and per-vertex precompute is embarrassingly
// parallel; every iteration writes only its own panel / vertex, so the results do not depend on the thread count).
// Plain std::thread: the host library links nothing but libstdc++ / libquadmath.  MLH_THREADS overrides the thread count
// (1 = serial, as before).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <system_error>
#include <thread>
#include <vector>

namespace mlh {

inline int host_threads() {
    static const int n = [] {
        if (const char* e = std::getenv("MLH_THREADS")) return std::max(1, std::atoi(e));
        unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        // one process per GPU (torchrun / mpirun): every rank runs the same setup, so the cores are shared between the ranks
        for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE"})
            if (const char* e = std::getenv(name)) {
                const int ranks = std::atoi(e);
                if (ranks > 1) hw = std::max(2u, hw / (unsigned)ranks);
                break;
            }
        return (int)std::min(16u, hw);
    }();
    return n;
}

// f(i) for i in [0, n), in contiguous chunks; the first exception thrown by any chunk is rethrown on the caller
template <class F>
void parallel_for(int n, F f, int min_per_thread = 128) {
    int nt = std::min(host_threads(), n / std::max(1, min_per_thread));
    if (nt <= 1) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    const int chunk = (n + nt - 1) / nt;
    std::exception_ptr err;
    std::mutex m;
    std::vector<std::thread> th;
    th.reserve(nt);
    auto run = [&](int a, int b) {
        try {
            for (int i = a; i < b; ++i) f(i);
        } catch (...) {
            std::lock_guard<std::mutex> g(m);
            if (!err) err = std::current_exception();
        }
    };
    for (int t = 0; t < nt; ++t) {
        const int a = t * chunk, b = std::min(n, a + chunk);
        if (a >= b) break;
        try {
            th.emplace_back(run, a, b);
        } catch (const std::system_error&) {
            run(a, b);   // no more threads to be had (a process limit): this chunk runs on the caller
        }
    }
    for (auto& x : th) x.join();
    if (err) std::rethrow_exception(err);
}

}  // namespace mlh
