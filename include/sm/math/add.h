// sm/math/add.h -- drop-in for the reference's include/math/add.h:5-14.
// The Op struct pattern is kept: `apply` is the scalar definition of the
// operation (the parity oracle), `apply_simd` stays declared for source
// compatibility with user code that specialises it (README.md:106-117), and the
// struct gains its device specialisation: `device_op` selects the sm_100a
// functor smb::DevOp<SMB_OP_ADD, T> (simplemath_b200/csrc/smb_math.cuh) that
// element_wise_op / array_scalar_op launch.
#pragma once
#include "helpers.h"

template<typename T>
struct AddOp {
    static constexpr int device_op = SMB_OP_ADD;

    static T apply(const T &a, const T &b) { return a + b; }

    template<typename SIMD_T>
    static SIMD_T apply_simd(const SIMD_T &a, const SIMD_T &b);
};
