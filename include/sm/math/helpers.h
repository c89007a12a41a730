// sm/math/helpers.h -- drop-in for the reference's include/math/helpers.h.
// The SimdTraits<T> load/store/set1 tables (helpers.h:12-119) described x86
// vector registers; on the device path the element types they enumerated are
// carried by smb::DTypeTag<T> instead, and an unsupported T is a compile error
// exactly as a missing SimdTraits specialisation was.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
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

    // Op struct -> device op id.  An Op without `device_op` has no device
    // specialisation and cannot be launched: compile error, not a CPU fallback.
    //   built-in Ops:      static constexpr int device_op = SMB_OP_*;
    //   user-defined Ops:  static int device_op() { return smb::op_id("MyOp"); }   (README.md:86-133 recipe;
    //                      the device side registers "MyOp" from a .cu file, include/smb200_plugin.cuh)
    template<typename Operation, typename = void> struct OpTag {
        static_assert(dependent_false<Operation>::value,
                      "this Op struct has no device specialisation (static constexpr int device_op = SMB_OP_*, "
                      "or static int device_op() for an op registered through smb200_plugin.cuh)");
    };
    template<typename Operation>
    struct OpTag<Operation, std::enable_if_t<std::is_convertible_v<decltype(Operation::device_op), int>>> {
        static int id() { return Operation::device_op; }
    };
    template<typename Operation>
    struct OpTag<Operation, std::enable_if_t<std::is_convertible_v<decltype(Operation::device_op()), int>>> {
        static int id() { return Operation::device_op(); }
    };

    // The id a user-defined Op was registered under (smb_register_op); throws when the device side
    // of that Op was never linked / loaded -- no CPU fallback.
    inline int op_id(const char *name) {
        const int id = smb_find_op(name);
        if (id < 0) throw std::runtime_error(std::string("smb200: no device Op registered under the name \"") + name + "\"");
        return id;
    }
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
