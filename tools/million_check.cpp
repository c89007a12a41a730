// tools/million_check.cpp -- the reference's own benchmark (benchmark/add.cpp:21-29, BASELINE config C1)
// through the drop-in headers, as shipped: `auto result = one + two;` on 10^6 floats in a loop, a fresh
// pooled result per iteration.  Also C2 ({4096,4096} + {1,4096}).  Three ways each:
//   synchronous   the reference's contract: every operator complete on return (default)
//   async_scope   sm::async_scope: operators enqueue and return, one wait at the end of the scope
//   chain         the same two operands through sm::lazy (one kernel, same bits) inside an async scope
// Host wall clock per iteration, JSON lines on stdout.
//   g++ -std=c++20 -O2 tools/million_check.cpp -Iinclude -Iinclude/sm -Lsimplemath_b200 -lsmb200 -Wl,-rpath,'$ORIGIN/../simplemath_b200' -o tools/million_check
#include <chrono>
#include <cstdio>
#include "sm.h"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template<typename F>
static double per_iter_us(int iters, F &&body) {
    for (int i = 0; i < iters / 10 + 10; ++i) body();
    sm::sync();
    const double t0 = now();
    for (int i = 0; i < iters; ++i) body();
    sm::sync();
    return (now() - t0) / iters * 1e6;
}

static void report(const char *config, const char *mode, double us, double bytes) {
    std::printf("{\"config\": \"%s\", \"mode\": \"%s\", \"us_per_iteration\": %.3f, \"gbs\": %.1f}\n", config, mode, us, bytes / us / 1e3);
    std::fflush(stdout);
}

int main() {
    {
        const sm::SMArray<float> one = sm::ones<float>(1'000'000);
        const sm::SMArray<float> two = sm::ones<float>(1'000'000);
        const double bytes = 12e6;
        const int iters = 20000;
        report("C1 million_check (benchmark/add.cpp:21-29)", "synchronous", per_iter_us(iters, [&] { auto result = one + two; (void) result; }), bytes);
        {
            sm::async_scope scope;
            report("C1 million_check (benchmark/add.cpp:21-29)", "async_scope", per_iter_us(iters, [&] { auto result = one + two; (void) result; }), bytes);
        }
        auto check = one + two;
        if (check(999999) != 2.0f) { std::printf("{\"error\": \"wrong result\"}\n"); return 1; }
    }
    {
        auto a = sm::ones<float>(4096, 4096);
        auto row = sm::ones<float>(1, 4096) * 3.0f;
        const double bytes = 4.0 * (2.0 * 4096 * 4096 + 4096);
        const int iters = 5000;
        report("C2 {4096,4096}+{1,4096}", "synchronous", per_iter_us(iters, [&] { auto r = a + row; (void) r; }), bytes);
        {
            sm::async_scope scope;
            report("C2 {4096,4096}+{1,4096}", "async_scope", per_iter_us(iters, [&] { auto r = a + row; (void) r; }), bytes);
        }
        auto check = a + row;
        if (check(4095, 4095) != 4.0f) { std::printf("{\"error\": \"wrong result\"}\n"); return 1; }
    }
    {   // a three-operator expression on C1-sized arrays: the temporaries of the eager form vs one fused kernel
        auto a = sm::ones<float>(1'000'000), b = sm::ones<float>(1'000'000), c = sm::ones<float>(1'000'000);
        const int iters = 10000;
        sm::async_scope scope;
        report("(a + b) * c - a, 10^6 floats", "async_scope eager (3 kernels, 2 temporaries)", per_iter_us(iters, [&] { auto r = (a + b) * c - a; (void) r; }), 36e6);
        report("(a + b) * c - a, 10^6 floats", "async_scope sm::lazy (1 kernel)", per_iter_us(iters, [&] { sm::SMArray<float> r = (sm::lazy(a) + b) * c - a; (void) r; }), 16e6);
    }
    {   // the reference's own test pattern (tests/add.cpp:59-92): fill a small operand element by element through operator(),
        // then broadcast it against a view of a big array.  Every element write pulls managed pages to the host; the headers tell
        // the library once per fill loop, and the next kernel brings the block back with ONE prefetch instead of demand paging.
        auto big = sm::ones<float>(32, 224, 224, 3);
        auto two = sm::ones<float>(1, 224, 1, 3);
        const int iters = 200;
        double fill_us = 0, op_us = 0;
        for (int it = 0; it < iters + 20; ++it) {
            const double t0 = now();
            for (size_t i = 0; i < 224; ++i) for (size_t c = 0; c < 3; ++c) two(0, i, 0, c) = 3.0f;
            const double t1 = now();
            auto v = big(it % 32, SLICE_ALL);
            auto r = v + two;
            const double t2 = now();
            if (it >= 20) { fill_us += (t1 - t0) * 1e6; op_us += (t2 - t1) * 1e6; }
            if (r.data[100] != 4.0f) { std::printf("{\"error\": \"wrong result\"}\n"); return 1; }
        }
        std::printf("{\"config\": \"tests/add.cpp:59-92 pattern: 672 element writes through operator(), then view {1,224,224,3} + {1,224,1,3}\", "
                    "\"fill_us\": %.2f, \"operator_us\": %.2f}\n", fill_us / iters, op_us / iters);
    }
    return 0;
}
