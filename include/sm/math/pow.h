// sm/math/pow.h -- drop-in for the reference's include/math/pow.h:6-14 (+ the
// integer SIMD bodies :56-95 and math/simd/crafted_pow.h:54-103, whose
// lane-wise wrapping semantics live in smb::powi_lane on the device).
// Float/double pow: the reference declares but never defines the SIMD body
// (pow.h:16-52 is commented out, so sm::pow<float> does not link there); here it
// is the correctly range-reduced exp2(y*log2 x) device kernel, ULP-bounded
// against std::pow.
#pragma once
#include <cmath>
#include "helpers.h"

template<typename T>
struct PowOp {
    static constexpr int device_op = SMB_OP_POW;

    static T apply(const T &base, const T &exp) { return std::pow(base, exp); }

    template<typename SIMD_T>
    static SIMD_T apply_simd(const SIMD_T &base, const SIMD_T &exponent);
};
