// smb_plan.h -- host-side planning for one element_wise_op call: dimension
// coalescing, fast-path predicate, fast-divmod magic numbers, vector-width
// eligibility.  Pure host C++ (no CUDA calls) so it is testable without a GPU
// through smb_plan_elementwise().
//
// Reference behaviour being planned for: include/math/calculate.h:5-99
// (dispatcher :10-13, prod_shape / div-mod offset math :26-30,:54-63) and
// include/math/helpers.h:130-139 (is_contiguous).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifndef SMB_MAX_NDIM
#define SMB_MAX_NDIM 6
#endif

namespace smb {

// q = x / d for 0 <= x < 2^31 with one mul.hi + shift (d >= 1).
// Round-up method: p = 31 + ceil(log2 d), m = ceil(2^p / d) fits 32 bits.
struct FastDiv32 {
    uint32_t d;
    uint32_t mul;
    uint32_t shr;
};
inline FastDiv32 make_fastdiv32(uint32_t d) {
    FastDiv32 f;
    f.d = d;
    if (d <= 1) { f.mul = 0; f.shr = 0; return f; } // handled as q = x
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg; // ceil(log2 d), 1..32
    uint32_t p = 31 + lg;
    uint64_t m = ((1ull << p) + d - 1) / d; // < 2^32 because 2^p/d < 2^32
    f.mul = (uint32_t)m;
    f.shr = p - 32;
    return f;
}

enum PlanKind { PLAN_CONTIGUOUS = 0, PLAN_ROW = 1, PLAN_GENERIC = 2 };

struct ElementwisePlan {
    int kind;
    int ndim;                      // coalesced rank, >= 1
    uint64_t shape[SMB_MAX_NDIM];  // coalesced result shape
    uint64_t sa[SMB_MAX_NDIM];     // operand strides per coalesced dim (elements)
    uint64_t sb[SMB_MAX_NDIM];
    uint64_t n;                    // prod(shape)
    uint64_t extent_a;             // 1 + sum (shape_k - 1) * stride_k : elements touched span
    uint64_t extent_b;
};

// Drop size-1 dims, then merge adjacent dims (k, k+1) whenever both operands
// satisfy stride[k] == shape[k+1] * stride[k+1] (this includes 0 == shape*0, a
// dim pair that is broadcast on both levels).  Row-major order is preserved, so
// the flat output index is unchanged.
inline ElementwisePlan make_plan(const uint64_t *stride_a, const uint64_t *stride_b,
                                 const uint64_t *shape, int ndim) {
    ElementwisePlan p;
    int m = 0;
    uint64_t n = 1;
    for (int k = 0; k < ndim; ++k) n *= shape[k];
    p.n = n;
    for (int k = 0; k < ndim; ++k) {
        if (shape[k] == 1) continue;
        if (m > 0 && p.sa[m - 1] == shape[k] * stride_a[k] && p.sb[m - 1] == shape[k] * stride_b[k]) {
            p.shape[m - 1] *= shape[k];
            p.sa[m - 1] = stride_a[k];
            p.sb[m - 1] = stride_b[k];
        } else {
            p.shape[m] = shape[k];
            p.sa[m] = stride_a[k];
            p.sb[m] = stride_b[k];
            ++m;
        }
    }
    if (m == 0) { // a single element (or an empty result)
        p.shape[0] = n ? 1 : 0;
        p.sa[0] = 1;
        p.sb[0] = 1;
        m = 1;
    }
    p.ndim = m;
    for (int k = m; k < SMB_MAX_NDIM; ++k) { p.shape[k] = 1; p.sa[k] = 0; p.sb[k] = 0; }
    p.extent_a = n ? 1 : 0;
    p.extent_b = n ? 1 : 0;
    if (n) {
        for (int k = 0; k < m; ++k) {
            p.extent_a += (p.shape[k] - 1) * p.sa[k];
            p.extent_b += (p.shape[k] - 1) * p.sb[k];
        }
    }
    const uint64_t ia = p.sa[m - 1], ib = p.sb[m - 1];
    if (m == 1 && ia == 1 && ib == 1) p.kind = PLAN_CONTIGUOUS;
    else if (ia <= 1 && ib <= 1) p.kind = PLAN_ROW;
    else p.kind = PLAN_GENERIC;
    return p;
}

// ---- op chains (smb_chain): the same coalescing over up to kChainMax leaves ----
constexpr int kChainMax = 8;
struct ChainPlan {
    int ndim;
    int nleaf;
    uint64_t shape[SMB_MAX_NDIM];
    uint64_t stride[kChainMax][SMB_MAX_NDIM]; // all zero for a constant leaf
    uint64_t extent[kChainMax];               // elements spanned by each array leaf
    uint64_t n;
    bool inner_unit_or_zero;                  // every leaf has inner stride 0 or 1
};
// strides[i] == nullptr marks a constant leaf.
inline ChainPlan make_chain_plan(const uint64_t *const *strides, int nleaf, const uint64_t *shape, int ndim) {
    ChainPlan p;
    p.nleaf = nleaf;
    uint64_t n = 1;
    for (int k = 0; k < ndim; ++k) n *= shape[k];
    p.n = n;
    int m = 0;
    auto st = [&](int i, int k) -> uint64_t { return strides[i] ? strides[i][k] : 0; };
    for (int k = 0; k < ndim; ++k) {
        if (shape[k] == 1) continue;
        bool merge = m > 0;
        for (int i = 0; merge && i < nleaf; ++i) merge = p.stride[i][m - 1] == shape[k] * st(i, k);
        if (merge) {
            p.shape[m - 1] *= shape[k];
            for (int i = 0; i < nleaf; ++i) p.stride[i][m - 1] = st(i, k);
        } else {
            p.shape[m] = shape[k];
            for (int i = 0; i < nleaf; ++i) p.stride[i][m] = st(i, k);
            ++m;
        }
    }
    if (m == 0) {
        p.shape[0] = n ? 1 : 0;
        for (int i = 0; i < nleaf; ++i) p.stride[i][0] = strides[i] ? 1 : 0;
        m = 1;
    }
    p.ndim = m;
    for (int k = m; k < SMB_MAX_NDIM; ++k) {
        p.shape[k] = 1;
        for (int i = 0; i < nleaf; ++i) p.stride[i][k] = 0;
    }
    p.inner_unit_or_zero = true;
    for (int i = 0; i < nleaf; ++i) {
        p.extent[i] = n ? 1 : 0;
        if (n) for (int k = 0; k < m; ++k) p.extent[i] += (p.shape[k] - 1) * p.stride[i][k];
        if (p.stride[i][m - 1] > 1) p.inner_unit_or_zero = false;
    }
    return p;
}

} // namespace smb
