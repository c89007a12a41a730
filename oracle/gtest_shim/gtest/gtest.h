// Minimal stand-in for <gtest/gtest.h> (GoogleTest is FetchContent-from-GitHub
// in the reference, cmake/gtest.cmake:5-11, and there is no network here).
// Just enough surface for the reference's tests/*.cpp to compile unmodified:
// TEST, EXPECT_/ASSERT_ EQ, FLOAT_EQ, DOUBLE_EQ (4-ULP, as GoogleTest defines
// them), TRUE/FALSE/THROW, and a main() that runs the registry.
// TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace shim {
struct Case {
    const char *suite;
    const char *name;
    void (*fn)();
};
inline std::vector<Case> &registry() {
    static std::vector<Case> r;
    return r;
}
inline int &failures() {
    static int f = 0;
    return f;
}
inline int &reported() {
    static int f = 0;
    return f;
}
struct Registrar {
    Registrar(const char *s, const char *n, void (*fn)()) { registry().push_back({s, n, fn}); }
};
inline void fail(const char *file, int line, const std::string &what) {
    ++failures();
    if (reported()++ < 20) std::fprintf(stderr, "%s:%d: Failure\n  %s\n", file, line, what.c_str());
}
template<typename F>
inline uint64_t biased(F v) {
    using U = std::conditional_t<sizeof(F) == 4, uint32_t, uint64_t>;
    U u;
    std::memcpy(&u, &v, sizeof(F));
    const U sign = U(1) << (sizeof(F) * 8 - 1);
    return (u & sign) ? uint64_t(U(~u + 1)) : uint64_t(U(u | sign));
}
template<typename F>
inline bool almost_equal(F a, F b) {
    if (std::isnan(a) || std::isnan(b)) return false;
    uint64_t x = biased(a), y = biased(b);
    return (x > y ? x - y : y - x) <= 4;
}
template<typename A, typename B>
inline std::string describe(const char *ea, const char *eb, const A &a, const B &b) {
    std::ostringstream os;
    os << "expected equality of " << ea << " and " << eb;
    if constexpr (requires { os << a; os << b; }) os << " (" << a << " vs " << b << ")";
    return os.str();
}
} // namespace shim

namespace testing {
inline void InitGoogleTest(int *, char **) {}
} // namespace testing

inline int RUN_ALL_TESTS() {
    int bad = 0;
    for (auto &c : shim::registry()) {
        int before = shim::failures();
        c.fn();
        bool ok = shim::failures() == before;
        std::printf("[%s] %s.%s\n", ok ? "  OK  " : "FAILED", c.suite, c.name);
        bad += !ok;
    }
    std::printf("%zu tests, %d failed\n", shim::registry().size(), bad);
    return bad ? 1 : 0;
}

#define TEST(suite, name)                                                        \
    static void suite##_##name##_body();                                         \
    static shim::Registrar suite##_##name##_reg(#suite, #name, &suite##_##name##_body); \
    static void suite##_##name##_body()

#define SHIM_CHECK_(cond, msg, on_fail)                   \
    do {                                                  \
        if (!(cond)) {                                    \
            shim::fail(__FILE__, __LINE__, (msg));        \
            on_fail;                                      \
        }                                                 \
    } while (0)

#define EXPECT_EQ(a, b) SHIM_CHECK_((a) == (b), shim::describe(#a, #b, (a), (b)), (void)0)
#define ASSERT_EQ(a, b) SHIM_CHECK_((a) == (b), shim::describe(#a, #b, (a), (b)), return)
#define EXPECT_NE(a, b) SHIM_CHECK_((a) != (b), std::string(#a " == " #b), (void)0)
#define EXPECT_TRUE(a) SHIM_CHECK_((a), std::string(#a " is false"), (void)0)
#define ASSERT_TRUE(a) SHIM_CHECK_((a), std::string(#a " is false"), return)
#define EXPECT_FALSE(a) SHIM_CHECK_(!(a), std::string(#a " is true"), (void)0)
#define EXPECT_FLOAT_EQ(a, b) \
    SHIM_CHECK_(shim::almost_equal<float>((a), (b)), shim::describe(#a, #b, float(a), float(b)), (void)0)
#define ASSERT_FLOAT_EQ(a, b) \
    SHIM_CHECK_(shim::almost_equal<float>((a), (b)), shim::describe(#a, #b, float(a), float(b)), return)
#define EXPECT_DOUBLE_EQ(a, b) \
    SHIM_CHECK_(shim::almost_equal<double>((a), (b)), shim::describe(#a, #b, double(a), double(b)), (void)0)
#define ASSERT_DOUBLE_EQ(a, b) \
    SHIM_CHECK_(shim::almost_equal<double>((a), (b)), shim::describe(#a, #b, double(a), double(b)), return)
#define EXPECT_THROW(stmt, ex)                                              \
    do {                                                                    \
        bool caught_ = false;                                               \
        try { stmt; } catch (const ex &) { caught_ = true; } catch (...) {} \
        SHIM_CHECK_(caught_, std::string(#stmt " did not throw " #ex), (void)0); \
    } while (0)

#ifndef SHIM_NO_MAIN
int main(int argc, char **argv) {
    testing::InitGoogleTest(&argc, argv);
    return RUN_ALL_TESTS();
}
#endif
