// Stress of the hybrid spin / block fork-join pool behind the host rANS coder (csrc/rans_host.cpp): random task counts,
// random pauses around the spin deadline; a lost wake-up hangs, a stale job descriptor changes the sum.
#include "rans_host.h"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <random>
#include <thread>
using namespace dcvc;
int main() {
    ForkJoin fj(7);
    std::mt19937 g(1);
    std::atomic<long> sum{0};
    long expect = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < 20000; ++it) {
        int n = 1 + g() % 8;
        fj.run(n, [&](int i) { sum.fetch_add(i + 1); });
        expect += n * (n + 1) / 2;
        int r = g() % 100;
        if (r < 10) std::this_thread::sleep_for(std::chrono::microseconds(g() % 300));   // around the spin deadline
        else if (r < 12) std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
    double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("sum %ld expect %ld  %s  %.2f s\n", sum.load(), expect, sum.load() == expect ? "OK" : "MISMATCH", s);
}
