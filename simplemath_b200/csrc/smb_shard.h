// smb_shard.h -- host-side planning of ONE operator spread over several GPUs.
//
// The path shards with zero exchange: every output element depends on one element of each operand
// (reference include/math/calculate.h:96), so the broadcast result's flat index range [0, n) is cut
// into G contiguous ranges, device g runs the same kernels on range g, an operand that streams with
// the output is split by the ranges its device touches, and an operand several devices need (a
// broadcast row, a small outer-product factor) is replicated.  This header only does the
// arithmetic -- where to cut, which elements of an operand a flat range touches -- so it is testable
// without a GPU (smb_plan_shards); the launcher that acts on it is in smb_devices.inl.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "smb_plan.h"

namespace smb {

constexpr int kMaxShards = 64;

struct ShardSplit {
    int g = 0;
    uint64_t unit = 1;                 // every interior boundary is a multiple of it
    uint64_t bounds[kMaxShards + 1];   // device i owns flat elements [bounds[i], bounds[i + 1])
};

// Where to cut.  `rows` x `inner` is the coalesced result seen as (leading dim, everything else);
// rank-1 results pass rows = n, inner = 1.
//   * results with at least G leading-dim rows are cut between rows, so that a broadcast operand's
//     rows stay whole on one device and the slab kernels (k_outer) still apply;
//   * boundaries are multiples of a 2 MiB page of the OUTPUT when the array is large enough for
//     that to be free: managed memory migrates in 2 MiB blocks, a page shared by two devices would
//     bounce between them;
//   * small arrays are cut on 4 KiB (tile / vector friendly) boundaries.
inline ShardSplit split_flat(uint64_t n, uint64_t rows, uint64_t inner, int ndim, int G, size_t es) {
    ShardSplit s;
    s.g = G;
    const uint64_t page = (2ull << 20) / es;
    uint64_t unit;
    if (ndim >= 2 && rows >= (uint64_t)G && inner > 0) {
        unit = inner;
        if (inner < page && page % inner == 0 && rows / (page / inner) >= 4ull * G) unit = page;
    } else {
        unit = n >= 4ull * G * page ? page : 4096 / es;
    }
    s.unit = unit;
    const uint64_t units = (n + unit - 1) / unit;
    const uint64_t base = units / G, extra = units % G;
    uint64_t at = 0;
    s.bounds[0] = 0;
    for (int i = 0; i < G; ++i) {
        at += base + ((uint64_t)i < extra ? 1 : 0);
        const uint64_t b = at * unit;
        s.bounds[i + 1] = b < n ? b : n;
    }
    s.bounds[G] = n;
    return s;
}

// The interval hull [lo, hi) of the operand elements that flat result indices [lin_lo, lin_lo + cnt)
// read through the stride table `stride` (elements, >= 0, 0 on broadcast dims) of a result of
// `shape`.  Exact as a hull for any strides (a transposed operand's offsets are not monotone in the
// flat index): the flat interval is decomposed into at most 2 * ndim + 1 index boxes, whose corner
// offsets are trivial.
struct ElemRange { uint64_t lo, hi; };
namespace detail {
struct Hull {
    uint64_t lo = ~0ull, hi = 0;
    void add(uint64_t a, uint64_t b) { if (a < lo) lo = a; if (b > hi) hi = b; }
};
struct HullCtx {
    const uint64_t *shape, *stride;
    int ndim;
    uint64_t tail[SMB_MAX_NDIM + 1]; // tail[k] = sum over j >= k of (shape_j - 1) * stride_j
    Hull h;
    // all indices with dims < k fixed (offset `base`), dim k in [i0, i1], dims > k free
    void box(uint64_t base, int k, uint64_t i0, uint64_t i1) {
        if (i0 > i1) return;
        h.add(base + i0 * stride[k], base + i1 * stride[k] + tail[k + 1]);
    }
    // dims < k fixed; trailing multi-index lexicographically >= idx[k..]
    void from(uint64_t base, int k, const uint64_t *idx) {
        if (k == ndim) { h.add(base, base); return; }
        from(base + idx[k] * stride[k], k + 1, idx);
        if (idx[k] + 1 <= shape[k] - 1) box(base, k, idx[k] + 1, shape[k] - 1);
    }
    // dims < k fixed; trailing multi-index lexicographically <= idx[k..]
    void upto(uint64_t base, int k, const uint64_t *idx) {
        if (k == ndim) { h.add(base, base); return; }
        if (idx[k] > 0) box(base, k, 0, idx[k] - 1);
        upto(base + idx[k] * stride[k], k + 1, idx);
    }
};
} // namespace detail

inline ElemRange touched_range(const uint64_t *shape, const uint64_t *stride, int ndim, uint64_t lin_lo, uint64_t cnt) {
    if (cnt == 0) return ElemRange{0, 0};
    detail::HullCtx c;
    c.shape = shape;
    c.stride = stride;
    c.ndim = ndim;
    c.tail[ndim] = 0;
    for (int k = ndim - 1; k >= 0; --k) c.tail[k] = c.tail[k + 1] + (shape[k] - 1) * stride[k];
    uint64_t il[SMB_MAX_NDIM], ir[SMB_MAX_NDIM];
    uint64_t l = lin_lo, r = lin_lo + cnt - 1;
    for (int k = ndim - 1; k >= 0; --k) {
        if (k == 0) { il[0] = l; ir[0] = r; }
        else { il[k] = l % shape[k]; l /= shape[k]; ir[k] = r % shape[k]; r /= shape[k]; }
    }
    int d = 0;
    uint64_t base = 0;
    while (d < ndim && il[d] == ir[d]) { base += il[d] * stride[d]; ++d; }
    if (d == ndim) c.h.add(base, base);
    else {
        c.from(base + il[d] * stride[d], d + 1, il);
        if (il[d] + 1 <= ir[d] - 1 && ir[d] > 0) c.box(base, d, il[d] + 1, ir[d] - 1);
        c.upto(base + ir[d] * stride[d], d + 1, ir);
    }
    return ElemRange{c.h.lo, c.h.hi + 1};
}

// How one operand is made available to the devices of a split.
enum ShardMode {
    SHARD_IN_PLACE = 0,  // the devices' ranges are disjoint: each device gets its own range of the array
    SHARD_REPLICATE = 1, // several devices need the same elements: each gets a private copy of what it reads
    SHARD_REFUSE = 2     // overlapping AND too large to copy per call: run the operator on one device
};
struct OperandShards {
    int mode;
    ElemRange r[kMaxShards];
};
inline OperandShards plan_operand(const uint64_t *shape, const uint64_t *stride, int ndim, const ShardSplit &s, size_t es,
                                  uint64_t replicate_max_bytes) {
    OperandShards o;
    bool disjoint = true;
    uint64_t prev_hi = 0, biggest = 0;
    bool have_prev = false;
    for (int i = 0; i < s.g; ++i) {
        o.r[i] = touched_range(shape, stride, ndim, s.bounds[i], s.bounds[i + 1] - s.bounds[i]);
        if (o.r[i].hi == o.r[i].lo) continue;
        if (have_prev && o.r[i].lo < prev_hi) disjoint = false;
        prev_hi = o.r[i].hi;
        have_prev = true;
        if (o.r[i].hi - o.r[i].lo > biggest) biggest = o.r[i].hi - o.r[i].lo;
    }
    o.mode = disjoint ? SHARD_IN_PLACE : biggest * es <= replicate_max_bytes ? SHARD_REPLICATE : SHARD_REFUSE;
    // ranges below one 2 MiB page each: moving pages between devices would split a page among them;
    // a private copy of a few hundred KiB is cheaper and leaves the array where it is
    if (o.mode == SHARD_IN_PLACE && s.g > 1 && biggest * es < (2ull << 20) && biggest * es <= replicate_max_bytes) o.mode = SHARD_REPLICATE;
    return o;
}

} // namespace smb
