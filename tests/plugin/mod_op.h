// tests/plugin/mod_op.h -- a USER-DEFINED Op struct, written the way the reference's README tells users to
// add an operation (README.md:86-133), living OUTSIDE the library's sources: ModOp<T>, the C remainder
// (int: a % b, truncated like C; float: fmodf).  `apply` is the scalar definition (and this test's
// oracle); the device side is `apply_device`, registered from mod_op_device.cu.
#pragma once
#include <cmath>
#include <cstdint>

#include "math/helpers.h"

template<typename T>
struct ModOp {
    static T apply(const T &a, const T &b) {
        if constexpr (std::is_integral_v<T>) return a % b;
        else return std::fmod(a, b);
    }

    template<typename SIMD_T>
    static SIMD_T apply_simd(const SIMD_T &a, const SIMD_T &b); // the reference's slot for x86 SIMD bodies: unused on the device path

    static int device_op() { return smb::op_id("ModOp"); }

#ifdef __CUDACC__
    static __device__ __forceinline__ T apply_device(T a, T b) {
        if constexpr (std::is_integral_v<T>) return a % b;
        else return fmodf(a, b);
    }
#endif
};

// A second user Op: the README's own example, (a + b) * 2.
template<typename T>
struct MyOp {
    static T apply(const T &a, const T &b) { return (a + b) * 2; }
    static int device_op() { return smb::op_id("MyOp"); }
#ifdef __CUDACC__
    static __device__ __forceinline__ T apply_device(T a, T b) {
        if constexpr (sizeof(T) == 8) return __dmul_rn(__dadd_rn(a, b), 2.0);
        else if constexpr (std::is_integral_v<T>) return (a + b) * 2;
        else return __fmul_rn(__fadd_rn(a, b), 2.0f);
    }
#endif
};
