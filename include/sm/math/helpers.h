// sm/math/helpers.h -- drop-in for the reference's include/math/helpers.h.
// The SimdTraits<T> load/store/set1 tables (helpers.h:12-119) described x86
// vector registers; on the device path the element types they enumerated are
// carried by smb::DTypeTag<T> instead, and an unsupported T is a compile error
// exactly as a missing SimdTraits specialisation was.
#pragma once
#include <cstddef>
#include <cstdint>
#include <type_traits>
#include <vector>

#include <smb200.h>

#define MAX_NDIM SMB_MAX_NDIM /* reference helpers.h:4 */

template<typename>
struct dependent_false : std::false_type {};

namespace smb {
    // Element types on the hot path: float, double, int32_t (helpers.h:23-119).
    template<typename T> struct DTypeTag;
    template<> struct DTypeTag<float> { static constexpr int value = SMB_F32; };
    template<> struct DTypeTag<double> { static constexpr int value = SMB_F64; };
    template<> struct DTypeTag<int32_t> { static constexpr int value = SMB_I32; };

    // Op struct -> device op tag.  An Op without `device_op` has no device
    // specialisation and cannot be launched: compile error, not a CPU fallback.
    template<typename Operation, typename = void> struct OpTag {
        static_assert(dependent_false<Operation>::value,
                      "this Op struct has no device specialisation (static constexpr int device_op = SMB_OP_*)");
    };
    template<typename Operation>
    struct OpTag<Operation, std::void_t<decltype(Operation::device_op)>> {
        static constexpr int value = Operation::device_op;
    };
}

// Row-major contiguity predicate, reference helpers.h:130-139.
inline bool is_contiguous(const std::vector<size_t> &shape, const std::vector<size_t> &stride) {
    size_t expected = 1;
    for (size_t k = shape.size(); k-- > 0;) {
        if (stride[k] != expected) return false;
        expected *= shape[k];
    }
    return true;
}
