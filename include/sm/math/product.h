// sm/math/product.h -- drop-in for the reference's include/math/product.h (dot_product<T>,
// called by SMArray::operator%, SMArray.h:213-215).  float / double / int32 run the device
// reduction (smb_dot); other element types (std::complex) keep a plain host loop -- they are
// not part of the accelerated element-type set (helpers.h:23-119).
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>

#include <smb200.h>
#include "calculate.h"

template<typename T>
T dot_product(const T *a, const T *b, size_t n) {
    if constexpr (requires { smb::DTypeTag<T>::value; }) {
        T result{};
        smb::check(smb_dot(smb::DTypeTag<T>::value, a, b, n, &result, nullptr));
        return result;
    } else {
        T sum{};
        for (size_t i = 0; i < n; ++i) sum += a[i] * b[i];
        return sum;
    }
}
