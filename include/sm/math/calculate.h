// sm/math/calculate.h -- drop-in for the reference's include/math/calculate.h.
// Same three global-namespace function templates, same signatures
// (calculate.h:2-3,5-8,137-138); the OpenMP/AVX loop bodies are replaced by one
// C-ABI call each into libsmb200.so (hand-written sm_100a kernels).  Results
// are complete on return, like the reference.  Errors surface as
// std::runtime_error, the reference's convention (SMUtils.h:77).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <smb200.h>
#include "helpers.h"

namespace smb {
    inline void check(int rc) {
        if (rc != SMB_OK) throw std::runtime_error(std::string("smb200: ") + smb_last_error());
    }
    static_assert(sizeof(size_t) == sizeof(uint64_t), "stride tables are passed as uint64_t");
    inline const uint64_t *u64(const std::vector<size_t> &v) { return reinterpret_cast<const uint64_t *>(v.data()); }
}

template<typename T, typename Operation>
void handle_contiguous_arrays(const T *a, const T *b, T *result, size_t n) {
    smb::check(smb_contiguous(smb::OpTag<Operation>::id(), smb::DTypeTag<T>::value, a, b, result, n, nullptr));
}

template<typename T, typename Operation>
void element_wise_op(const T *a, const std::vector<size_t> &stride_a,
                     const T *b, const std::vector<size_t> &stride_b,
                     size_t n, T *result, const std::vector<size_t> &shape) {
    if (shape.size() > MAX_NDIM) throw std::runtime_error("smb200: rank exceeds MAX_NDIM");
    smb::check(smb_elementwise(smb::OpTag<Operation>::id(), smb::DTypeTag<T>::value, a, smb::u64(stride_a), b,
                               smb::u64(stride_b), smb::u64(shape), static_cast<int>(shape.size()), n, result, nullptr));
}

template<typename T, typename Operation>
void array_scalar_op(const T *a, T value, const size_t n, T *result) {
    smb::check(smb_array_scalar(smb::OpTag<Operation>::id(), smb::DTypeTag<T>::value, a, &value, n, result, nullptr));
}
