// sm/UserFunctions.h -- drop-in for the reference's include/UserFunctions.h:8-57.
// The producers hand out pooled device-resident (managed) blocks and fill them
// with a device kernel instead of new[] + std::fill_n (UserFunctions.h:18-40);
// sm::pow launches the device PowOp through the same array-scalar path the
// reference uses (UserFunctions.h:42-48).
#pragma once
#include <ostream>

#include "SMArray.h"
#include "math/pow.h"

namespace sm {
    template<typename T, typename... Args>
    SMArray<T> empty(Args... args) {
        std::vector<size_t> shape = {static_cast<size_t>(args)...};
        T *data = storage::acquire<T>(calculateTotalSize(shape));
        storage::host_access(); // the caller fills it on the host: a recycled block must not have kernels of an async scope in flight
        smb_host_written(data); // the caller fills it through data / operator() on the host
        return {data, std::move(shape)};
    }

    namespace storage {
        template<typename T>
        inline T *filled(size_t n, T value) {
            T *data = acquire<T>(n);
            if constexpr (requires { smb::DTypeTag<T>::value; }) {
                smb::check(smb_fill(smb::DTypeTag<T>::value, data, &value, n, nullptr));
            } else {
                host_access();
                for (size_t i = 0; i < n; ++i) data[i] = value; // element types outside the hot path
            }
            return data;
        }
    }

    template<typename T, typename... Args>
    SMArray<T> ones(Args... args) {
        std::vector<size_t> shape = {static_cast<size_t>(args)...};
        T *data = storage::filled<T>(calculateTotalSize(shape), T{1});
        return {data, std::move(shape)};
    }

    template<typename T, typename... Args>
    SMArray<T> zeros(Args... args) {
        std::vector<size_t> shape = {static_cast<size_t>(args)...};
        T *data = storage::filled<T>(calculateTotalSize(shape), T{0});
        return {data, std::move(shape)};
    }

    // arr ^ val, the exponent being one scalar of type T.
    template<typename T>
    SMArray<T> pow(SMArray<T> &arr, T val) {
        return arr.template applyScalar<PowOp<T> >(val);
    }
}

template<typename T>
std::ostream &operator<<(std::ostream &os, const sm::SMArray<T> &b) {
    return os << b.toString();
}
