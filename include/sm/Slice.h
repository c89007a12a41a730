// sm/Slice.h -- drop-in for the reference's include/Slice.h:1-28.
// Host-side view bookkeeping only (out of the accelerated path; kept because
// views feed strides and interior pointers into it).
#pragma once
#include <cstddef>

struct Slice {
    enum SliceStep { SINGLE_STEP = 0 };   // the only step the reference supports (Slice.h:11-13)
    enum SliceType { INDEX = 0, SLICE };

    std::size_t start;
    std::size_t end;                       // (size_t)-1 means "to the end of the axis"
    SliceStep step = SINGLE_STEP;
    SliceType sliceType = SLICE;

    Slice(const std::size_t first, const std::size_t last = static_cast<std::size_t>(-1)) : start(first), end(last) {}
};

#define SLICE(start, end) Slice(start, end)
#define SLICE_START(start) SLICE(start, -1)
#define SLICE_END(end) SLICE(0, end)
#define SLICE_ALL SLICE(0, -1)
