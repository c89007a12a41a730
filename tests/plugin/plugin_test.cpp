// tests/plugin/plugin_test.cpp -- acceptance test of the OPEN Op-struct plugin pattern on the device
// (reference README.md:86-133, "Extending with Custom Operations"): ModOp<T> and MyOp<T> are defined in
// this directory, not in the library; their operators are written exactly as the README's step 3
// (sm::broadcast -> fresh result -> element_wise_op<T, Op<T>>), and every result is compared with the
// Op's own host `apply`.  Plain C++ (g++); the device bodies are in mod_op_device.cu.
#include <gtest/gtest.h>
#include <cstdint>
#include <random>
#include "sm.h"
#include "mod_op.h"

namespace {
// README step 3, as a free function (the reference adds it as an SMArray member; same body)
template<typename Op, typename T>
sm::SMArray<T> apply_op(const sm::SMArray<T> &x, const sm::SMArray<T> &y) {
    auto bc = sm::broadcast(x.shape(), x.strides(), y.shape(), y.strides());
    T *result = sm::storage::acquire<T>(bc.totalSize);
    element_wise_op<T, Op>(x.data, bc.newStrides1, y.data, bc.newStrides2, bc.totalSize, result, bc.resultShape);
    return sm::SMArray<T>(result, std::move(bc.resultShape));
}
} // namespace

TEST(Plugin, ModOpIntContiguousAndRaggedTail) {
    std::mt19937 g(5);
    const int n = 100003;
    auto a = sm::empty<int>(n), b = sm::empty<int>(n);
    for (int i = 0; i < n; ++i) { a.data[i] = int(g() % 2000001) - 1000000; b.data[i] = int(g() % 997) + 1; if (g() & 1) b.data[i] = -b.data[i]; }
    auto r = apply_op<ModOp<int>>(a, b);
    for (int i = 0; i < n; ++i) ASSERT_EQ(r.data[i], ModOp<int>::apply(a.data[i], b.data[i]));
    EXPECT_TRUE(ModOp<int>::device_op() >= SMB_OP_USER);
}

TEST(Plugin, ModOpBroadcastViewsAndScalar) {
    auto m = sm::empty<int>(37, 52);
    auto row = sm::empty<int>(1, 52);
    auto col = sm::empty<int>(37, 1);
    for (size_t i = 0; i < 37 * 52; ++i) m.data[i] = int(i * 7919 % 100003) - 50000;
    for (size_t j = 0; j < 52; ++j) row.data[j] = int(j) + 3;
    for (size_t i = 0; i < 37; ++i) col.data[i] = -int(i) - 2;
    auto r1 = apply_op<ModOp<int>>(m, row);
    auto r2 = apply_op<ModOp<int>>(m, col);
    auto r3 = apply_op<ModOp<int>>(col, row);               // {37,1} x {1,52}
    for (size_t i = 0; i < 37; ++i)
        for (size_t j = 0; j < 52; ++j) {
            ASSERT_EQ(r1(i, j), m(i, j) % row(0, j));
            ASSERT_EQ(r2(i, j), m(i, j) % col(i, 0));
            ASSERT_EQ(r3(i, j), col(i, 0) % row(0, j));
        }
    auto mt = m.transpose();                                // strided operand: the generic kernel
    auto rt = apply_op<ModOp<int>>(mt, col.transpose());    // {52,37} x {1,37}
    for (size_t i = 0; i < 52; ++i) for (size_t j = 0; j < 37; ++j) ASSERT_EQ(rt(i, j), m(j, i) % col(j, 0));
    auto big = sm::empty<int>(4, 30, 20);
    for (size_t i = 0; i < big.totalSize; ++i) big.data[i] = int(i) * 31 - 9000;
    auto v = big(2, SLICE_ALL);                             // interior-pointer view
    auto small = sm::empty<int>(1, 20);
    for (size_t j = 0; j < 20; ++j) small.data[j] = 7 + int(j);
    auto rv = apply_op<ModOp<int>>(v, small);
    for (size_t i = 0; i < 30; ++i) for (size_t j = 0; j < 20; ++j) ASSERT_EQ(rv(i, j), big(2, i, j) % small(0, j));
    auto rs = m.applyScalar<ModOp<int>>(11);                // array (op) scalar through the same Op
    for (size_t i = 0; i < 37 * 52; ++i) ASSERT_EQ(rs.data[i], m.data[i] % 11);
}

TEST(Plugin, ReadmeExampleOpFloatDoubleInt) {
    std::mt19937 g(6);
    auto a = sm::empty<float>(300, 211), b = sm::empty<float>(300, 211);
    for (size_t i = 0; i < a.totalSize; ++i) { a.data[i] = float(g() % 20001) * 0.37f - 3000.0f; b.data[i] = float(g() % 977) * 0.011f; }
    auto r = apply_op<MyOp<float>>(a, b);
    for (size_t i = 0; i < a.totalSize; ++i) { volatile float s = a.data[i] + b.data[i]; volatile float w = s * 2.0f; ASSERT_EQ(r.data[i], w); }
    auto d = sm::ones<double>(1000, 3) * 0.75;
    auto e = sm::ones<double>(1, 3) * 1.5;
    auto rd = apply_op<MyOp<double>>(d, e);
    for (size_t i = 0; i < rd.totalSize; ++i) ASSERT_EQ(rd.data[i], 4.5);
    sm::SMArray<int> x = {1, 2, 2147483647}, y = {10, 20, 1};
    auto ri = apply_op<MyOp<int>>(x, y);
    EXPECT_EQ(ri(0), 22); EXPECT_EQ(ri(1), 44);
    auto f = sm::empty<float>(1000);
    for (int i = 0; i < 1000; ++i) f.data[i] = float(i) * 0.731f - 200.0f;
    auto fm = f.applyScalar<ModOp<float>>(7.5f);
    for (int i = 0; i < 1000; ++i) ASSERT_EQ(fm.data[i], std::fmod(f.data[i], 7.5f));
}

TEST(Plugin, UserOpsOnADeviceSetAndInAnAsyncScope) {
    const size_t R = 2000, C = 1024;
    auto a = sm::empty<int>(R, C), row = sm::empty<int>(1, C);
    for (size_t i = 0; i < R * C; ++i) a.data[i] = int(i % 1000003) - 500000;
    for (size_t j = 0; j < C; ++j) row.data[j] = int(j % 97) + 2;
    auto one = apply_op<ModOp<int>>(a, row);
    const int64_t old_min = smb_get_option(SMB_OPT_SHARD_MIN_BYTES);
    smb_set_option(SMB_OPT_SHARD_MIN_BYTES, 1 << 20);
    sm::set_devices(smb_device_count() >= 2 ? std::vector<int>{0, 1} : std::vector<int>{0, 0, 0});
    sm::SMArray<int> many = apply_op<ModOp<int>>(a, row);
    {
        sm::async_scope scope;
        auto t = apply_op<ModOp<int>>(a, row);
        auto u = apply_op<MyOp<int>>(t, row);
        (void) u(0, 0);                                                                                          // an accessor waits for what the scope left in flight
        for (size_t i = 0; i < R * C; i += 4099) ASSERT_EQ(u.data[i], (one.data[i] + row.data[i % C]) * 2);     // raw .data only after that (or after sm::sync)
    }
    sm::set_devices({});
    smb_set_option(SMB_OPT_SHARD_MIN_BYTES, old_min);
    for (size_t i = 0; i < R * C; ++i) ASSERT_EQ(many.data[i], one.data[i]);
}

TEST(Plugin, AnOpWithoutADeviceSideFailsLoudly) {
    struct Nowhere { static int device_op() { return smb::op_id("NoSuchOp"); } };
    bool threw = false;
    try { (void) smb::OpTag<Nowhere>::id(); } catch (const std::runtime_error &e) { threw = std::string(e.what()).find("NoSuchOp") != std::string::npos; }
    EXPECT_TRUE(threw);   // no CPU fallback: an Op nobody registered cannot run
}
