// sm/SMUtils.h -- drop-in for the reference's include/SMUtils.h:1-100.
// sm::broadcast() stays on the host unchanged in behaviour: it
// touches five small vectors per call.  Its stride tables are what
// smb_elementwise() consumes.
#pragma once
#include <algorithm>
#include <concepts>
#include <cstddef>
#include <stdexcept>
#include <vector>

#include "Slice.h"
#include "macros.h"

namespace sm {
    // Field names and order as in the reference (SMUtils.h:5-12).
    struct BroadCastResult {
        std::vector<std::size_t> resultShape;
        std::vector<std::size_t> newShape1;
        std::vector<std::size_t> newStrides1;
        std::vector<std::size_t> newShape2;
        std::vector<std::size_t> newStrides2;
        size_t totalSize;
    };

    template<std::integral T>
    ALWAYS_INLINE Slice processIndex(T index) noexcept {
        Slice s(static_cast<std::size_t>(index), static_cast<std::size_t>(-1));
        s.sliceType = Slice::INDEX;
        return s;
    }

    ALWAYS_INLINE Slice processIndex(Slice s) noexcept { return s; }

    inline size_t calculateTotalSize(std::vector<size_t> &shape) {
        size_t n = 1;
        for (size_t d: shape) n *= d;
        return n;
    }

    // NumPy-style broadcasting, reference SMUtils.h:34-99: right-align the two
    // ranks; a missing leading dim counts as shape 1 / stride 0; dims must agree
    // or be 1; the result dim is the larger; an operand's stride becomes 0 where
    // its own dim is 1 and the other's is larger.
    inline BroadCastResult broadcast(const std::vector<size_t> &shape1, const std::vector<size_t> &strides1,
                                     const std::vector<size_t> &shape2, const std::vector<size_t> &strides2) {
        const size_t rank = std::max(shape1.size(), shape2.size());
        const size_t lead1 = rank - shape1.size(), lead2 = rank - shape2.size();
        BroadCastResult r;
        r.resultShape.resize(rank);
        r.newShape1.assign(rank, 1);
        r.newStrides1.assign(rank, 0);
        r.newShape2.assign(rank, 1);
        r.newStrides2.assign(rank, 0);
        r.totalSize = 1;
        for (size_t axis = 0; axis < rank; ++axis) {
            if (axis >= lead1) {
                r.newShape1[axis] = shape1[axis - lead1];
                r.newStrides1[axis] = strides1[axis - lead1];
            }
            if (axis >= lead2) {
                r.newShape2[axis] = shape2[axis - lead2];
                r.newStrides2[axis] = strides2[axis - lead2];
            }
            const size_t d1 = r.newShape1[axis], d2 = r.newShape2[axis];
            if (d1 != d2 && d1 != 1 && d2 != 1)
                throw std::runtime_error("Cannot broadcast shapes: incompatible dimensions");
            r.resultShape[axis] = std::max(d1, d2);
            r.totalSize *= r.resultShape[axis];
            if (d1 == 1 && d2 > 1) r.newStrides1[axis] = 0;
            if (d2 == 1 && d1 > 1) r.newStrides2[axis] = 0;
        }
        return r;
    }
}
