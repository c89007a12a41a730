// smb_api.cu -- the C ABI of libsmb200.so (declared in include/smb200.h):
// device runtime, pooled storage, planner, launchers and the host-operand
// staging pipeline.  See DESIGN.md for the data-flow picture.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <algorithm>

#include "../../include/smb200.h"
#include "smb_alloc.h"
#include "smb_kernels.cuh"
#include "smb_plan.h"

namespace smb {

// ------------------------------------------------------------------ errors --
static thread_local std::string g_err;
static thread_local const char *g_last_kernel = "";
static std::atomic<uint64_t> g_launches{0};

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define SMB_CK(call)                                                                              \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            cudaGetLastError();                                                                   \
            return fail(e_ == cudaErrorMemoryAllocation ? SMB_ERR_OOM : SMB_ERR_CUDA, "%s: %s",   \
                        #call, cudaGetErrorString(e_));                                           \
        }                                                                                         \
    } while (0)

// ------------------------------------------------------------ options -------
static std::atomic<int64_t> g_opt_pow_specialise{1};
static std::atomic<int64_t> g_opt_chunk_bytes{64ll << 20};
static std::atomic<int64_t> g_opt_contig_variant{0};
static std::atomic<int64_t> g_opt_bcast_variant{0};
static std::atomic<int64_t> g_opt_force_wide{0};

// ------------------------------------------------------ device context ------
constexpr int kSlots = 3; // staging pipeline depth (H2D | kernel | D2H in flight)
struct DeviceCtx {
    bool ready = false;
    int sm_count = 0;
    cudaStream_t main = nullptr;
    cudaStream_t slot[kSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev = nullptr;
};
static DeviceCtx g_ctx[64];
static std::mutex g_ctx_mu;

// There is no CPU fallback: every compute entry point goes through here and
// fails loudly when no CUDA device is usable.
static int current_ctx(DeviceCtx **out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(SMB_ERR_NO_DEVICE, "no CUDA device available (%s); libsmb200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(SMB_ERR_INVALID, "device index %d out of range", dev);
    DeviceCtx &c = g_ctx[dev];
    if (!c.ready) {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (!c.ready) {
            SMB_CK(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
            SMB_CK(cudaStreamCreateWithFlags(&c.main, cudaStreamNonBlocking));
            for (int i = 0; i < kSlots; ++i) SMB_CK(cudaStreamCreateWithFlags(&c.slot[i], cudaStreamNonBlocking));
            SMB_CK(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
            // ready-made shared-memory images of the f32 pow tables (one bulk copy per CTA later)
            k_pow_image_init<<<8, kBlock, 0, c.main>>>();
            SMB_CK(cudaGetLastError());
            SMB_CK(cudaStreamSynchronize(c.main));
            c.ready = true;
        }
    }
    *out = &c;
    return SMB_OK;
}

// ------------------------------------------------------------------ pool ----
cudaError_t Pool::raw_alloc(void **p, size_t bytes, int kind) {
    ++driver_calls_;
    if (kind == SMB_MEM_DEVICE) return cudaMalloc(p, bytes);
    if (kind == SMB_MEM_MANAGED) return cudaMallocManaged(p, bytes, cudaMemAttachGlobal);
    return cudaHostAlloc(p, bytes, cudaHostAllocPortable);
}
void Pool::raw_free(const Block &b) {
    if (b.kind == SMB_MEM_PINNED) { cudaFreeHost(b.base); return; }
    int cur = 0;
    cudaGetDevice(&cur);
    if (b.device >= 0 && b.device != cur) cudaSetDevice(b.device);
    cudaFree(b.base);
    if (b.device >= 0 && b.device != cur) cudaSetDevice(cur);
}
void *Pool::alloc(size_t bytes, int kind, int device, cudaError_t *err) {
    const size_t sz = bucket(bytes);
    const Key key{kind == SMB_MEM_PINNED ? -1 : device, kind, sz};
    std::lock_guard<std::mutex> lk(mu_);
    auto it = free_.find(key);
    if (it != free_.end() && !it->second.empty()) {
        void *p = it->second.back();
        it->second.pop_back();
        auto c = cached_.find((uintptr_t)p);
        Block b = c->second;
        cached_.erase(c);
        live_[(uintptr_t)p] = b;
        cached_bytes_ -= sz;
        in_use_ += sz;
        ++hits_;
        *err = cudaSuccess;
        return p;
    }
    void *p = nullptr;
    cudaError_t e = raw_alloc(&p, sz, kind);
    if (e == cudaErrorMemoryAllocation) { // give cached blocks back to the driver and retry once
        cudaGetLastError();
        for (auto &kv : cached_) raw_free(kv.second);
        cached_.clear();
        free_.clear();
        cached_bytes_ = 0;
        e = raw_alloc(&p, sz, kind);
    }
    *err = e;
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    live_[(uintptr_t)p] = Block{p, sz, key.device, kind};
    in_use_ += sz;
    return p;
}
bool Pool::free(void *ptr) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.find((uintptr_t)ptr);
    if (it == live_.end()) return false;
    Block b = it->second;
    live_.erase(it);
    in_use_ -= b.bytes;
    cached_[(uintptr_t)ptr] = b;
    cached_bytes_ += b.bytes;
    free_[Key{b.device, b.kind, b.bytes}].push_back(ptr);
    return true;
}
bool Pool::owns(const void *ptr, Block *out) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) return false;
    --it;
    const Block &b = it->second;
    if ((uintptr_t)ptr >= (uintptr_t)b.base + b.bytes) return false;
    if (out) *out = b;
    return true;
}
bool Pool::take_host_flag(const void *ptr, Block *out, bool *was_on_host) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) return false;
    --it;
    Block &b = it->second;
    if ((uintptr_t)ptr >= (uintptr_t)b.base + b.bytes) return false;
    *out = b;
    *was_on_host = b.maybe_on_host;
    b.maybe_on_host = false;
    return true;
}
void Pool::set_host_flag(const void *ptr, bool on_host) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) return;
    --it;
    Block &b = it->second;
    if ((uintptr_t)ptr < (uintptr_t)b.base + b.bytes) b.maybe_on_host = on_host;
}
void Pool::trim() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto &kv : cached_) raw_free(kv.second);
    cached_.clear();
    free_.clear();
    cached_bytes_ = 0;
}
void Pool::stats(uint64_t s[4]) {
    std::lock_guard<std::mutex> lk(mu_);
    s[0] = in_use_;
    s[1] = cached_bytes_;
    s[2] = driver_calls_;
    s[3] = hits_;
}

// Scoped device scratch block from the pool.
struct Scratch {
    void *p = nullptr;
    ~Scratch() { if (p) Pool::instance().free(p); }
    int get(size_t bytes, int device) {
        cudaError_t e;
        p = Pool::instance().alloc(bytes, SMB_MEM_DEVICE, device, &e);
        if (!p) return fail(e == cudaErrorMemoryAllocation ? SMB_ERR_OOM : SMB_ERR_CUDA, "scratch alloc of %zu bytes: %s",
                            bytes, cudaGetErrorString(e));
        return SMB_OK;
    }
};

// ------------------------------------------------------- pointer kinds ------
enum MemType { MT_HOST = 0, MT_PINNED = 1, MT_DEVICE = 2, MT_MANAGED = 3 };
static MemType mem_type(const void *p) {
    // Pool blocks first: no driver call (cudaPointerGetAttributes costs microseconds per operand,
    // which is most of the launch overhead of a small operator).
    Block blk;
    if (Pool::instance().owns(p, &blk))
        return blk.kind == SMB_MEM_DEVICE ? MT_DEVICE : blk.kind == SMB_MEM_MANAGED ? MT_MANAGED : MT_PINNED;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return MT_HOST; }
    switch (at.type) {
        case cudaMemoryTypeDevice: return MT_DEVICE;
        case cudaMemoryTypeManaged: return MT_MANAGED;
        case cudaMemoryTypeHost: return MT_PINNED;
        default: return MT_HOST;
    }
}
static inline bool on_host(MemType t) { return t == MT_HOST || t == MT_PINNED; }

// Managed operands: bring the pages to the GPU before the launch -- but only when they may be
// on the host.  cudaMemPrefetchAsync costs ~50 us even for resident pages (measured: 164 us per
// 3-operand call), so pool blocks carry a "maybe on host" flag (smb_alloc.h); foreign managed
// memory is always prefetched.
static void prefetch_managed(const void *p, size_t bytes, cudaStream_t s) {
    int dev = 0;
    cudaGetDevice(&dev);
    Block blk;
    bool was_on_host = true;
    if (Pool::instance().take_host_flag(p, &blk, &was_on_host)) {
        if (!was_on_host) return;
        p = blk.base;        // whole block: views of it become resident too
        bytes = blk.bytes;
    }
    if (cudaMemPrefetchAsync(p, bytes, dev, s) != cudaSuccess) cudaGetLastError(); // best effort
}

// ----------------------------------------------------------- launchers ------
static inline size_t esize(int dtype) { return dtype == SMB_F64 ? 8 : 4; }

// Library defaults for the dense-stream kernel (chosen by tools/sweep on B200,
// see profiles/): bytes per vector access and vectors in flight per thread.
#ifndef SMB_STREAM_VB
#define SMB_STREAM_VB 16
#endif
#ifndef SMB_STREAM_UNROLL
#define SMB_STREAM_UNROLL 4
#endif
constexpr int kThreads = kBlock;
#ifndef SMB_POW_TILES_PER_CTA
#define SMB_POW_TILES_PER_CTA 8 // f32 pow: consecutive 16 KB tiles per CTA (tools/sweep)
#endif
#ifndef SMB_TILE_R
#define SMB_TILE_R 64 // k_tile: rows per tile (transposed operand's contiguous direction); halved for 8-byte types
#define SMB_TILE_C 64 // cols per tile (result's contiguous direction)
#endif

static unsigned grid_for(uint64_t work_items, uint64_t items_per_block, int sm_count, int64_t ctas_per_sm) {
    uint64_t blocks = (work_items + items_per_block - 1) / items_per_block;
    if (blocks == 0) blocks = 1;
    if (ctas_per_sm > 0) blocks = std::min<uint64_t>(blocks, (uint64_t)sm_count * (uint64_t)ctas_per_sm);
    return (unsigned)std::min<uint64_t>(blocks, 0x7fffffffull);
}

template<typename T, typename Fn, bool HAS_B>
static int launch_stream(const DeviceCtx &c, const T *a, const T *b, T *out, uint64_t n, uint64_t first, Fn fn,
                         cudaStream_t s) {
    if (n == 0) return SMB_OK;
    constexpr int VB = SMB_STREAM_VB, UNROLL = SMB_STREAM_UNROLL;
    const uintptr_t ma = (uintptr_t)a % VB, mb = HAS_B ? (uintptr_t)b % VB : ma, mo = (uintptr_t)out % VB;
    int64_t cps = g_opt_contig_variant.load();
    // Plain streams run one tile per CTA (tools/sweep: fastest on B200).  Kernels that stage lookup
    // tables in shared memory (pow) give each CTA a few consecutive tiles to amortise the fill --
    // one bulk copy for f32, an in-kernel fill of 40 KB for f64 -- but stay many waves deep: a grid
    // of resident CTAs measured 10-15 % slower.  For them the option counts tiles per CTA.
    const int64_t tiles_per_cta = cps > 0 ? cps : (sizeof(T) == 4 ? SMB_POW_TILES_PER_CTA : 4 * SMB_POW_TILES_PER_CTA);
    if (ma == mb && ma == mo && ma % sizeof(T) == 0) {
        uint64_t head = ma ? (VB - ma) / sizeof(T) : 0;
        if (head > n) head = n;
        if (head) { // peel up to the first common vector boundary (views give interior pointers)
            k_stream_unaligned<T, Fn, HAS_B><<<1, kThreads, 0, s>>>(a, b, out, head, first, fn);
            ++g_launches;
        }
        const uint64_t rest = n - head;
        if (rest) {
            constexpr uint64_t per_block = (uint64_t)kThreads * UNROLL * (VB / sizeof(T));
            const unsigned grid = fn_pow_tables<Fn>::value && SMB_POW_BLOCKED
                                      ? grid_for(rest, per_block * (uint64_t)tiles_per_cta, c.sm_count, 0)
                                      : grid_for(rest, per_block, c.sm_count, fn_pow_tables<Fn>::value && cps == 0 ? 8 : cps);
            k_stream<T, Fn, HAS_B, VB, UNROLL><<<grid, kThreads, 0, s>>>(a + head, HAS_B ? b + head : nullptr, out + head,
                                                                       rest, first + head, fn);
            ++g_launches;
            g_last_kernel = HAS_B ? "k_stream<binary>" : "k_stream<scalar>";
        }
    } else {
        const unsigned grid = grid_for(n, kThreads, c.sm_count, 32);
        k_stream_unaligned<T, Fn, HAS_B><<<grid, kThreads, 0, s>>>(a, b, out, n, first, fn);
        ++g_launches;
        g_last_kernel = "k_stream_unaligned";
    }
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

template<typename T>
static int contiguous_t(const DeviceCtx &c, int op, const T *a, const T *b, T *out, uint64_t n, uint64_t first,
                        uint64_t lane_end, cudaStream_t s) {
    switch (op) {
        case SMB_OP_ADD: return launch_stream<T, BinaryFn<OP_ADD, T>, true>(c, a, b, out, n, first, {lane_end}, s);
        case SMB_OP_SUB: return launch_stream<T, BinaryFn<OP_SUB, T>, true>(c, a, b, out, n, first, {lane_end}, s);
        case SMB_OP_MUL: return launch_stream<T, BinaryFn<OP_MUL, T>, true>(c, a, b, out, n, first, {lane_end}, s);
        case SMB_OP_DIV: return launch_stream<T, BinaryFn<OP_DIV, T>, true>(c, a, b, out, n, first, {lane_end}, s);
        case SMB_OP_POW: return launch_stream<T, BinaryFn<OP_POW, T>, true>(c, a, b, out, n, first, {lane_end}, s);
    }
    return fail(SMB_ERR_INVALID, "unknown op %d", op);
}

// array (op) scalar.  `first` is the absolute flat index of a[0] (staging
// chunks), lane_end the absolute end of the reference's SIMD region.
template<typename T>
static int scalar_t(const DeviceCtx &c, int op, const T *a, T v, T *out, uint64_t n, uint64_t first, uint64_t lane_end,
                    cudaStream_t s);

template<>
int scalar_t<float>(const DeviceCtx &c, int op, const float *a, float v, float *out, uint64_t n, uint64_t first,
                    uint64_t lane_end, cudaStream_t s) {
    using T = float;
    switch (op) {
        case SMB_OP_ADD: return launch_stream<T, ScalarFn<OP_ADD, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_SUB: return launch_stream<T, ScalarFn<OP_SUB, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_MUL: return launch_stream<T, ScalarFn<OP_MUL, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_DIV: return launch_stream<T, ScalarFn<OP_DIV, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_POW: {
            if (g_opt_pow_specialise.load()) {
                if (v == 2.0f) return launch_stream<T, PowSpecialFn<POWS_SQUARE, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == -1.0f) return launch_stream<T, PowSpecialFn<POWS_RECIP, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == 0.5f) return launch_stream<T, PowSpecialFn<POWS_SQRT, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == 1.0f) return launch_stream<T, PowSpecialFn<POWS_IDENT, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
            }
            {
                const PowExpF32 pe = classify_exp(v);
                if (!pow_f32_fast_ok(pe)) // |y| >= 2^64, tiny, zero, inf or NaN: the reference-accuracy path alone
                    return launch_stream<T, PowF32SlowFn, false>(c, a, nullptr, out, n, first, PowF32SlowFn::make(v, lane_end), s);
                const bool lt1 = pow_f32_y_lt_1(pe);
                const int tier = pow_f32_tier(pe), sign = pow_f32_sign_mode(pe);
#define SMB_POW_LAUNCH(S, G, L) launch_stream<T, PowF32Fn<S, G, L>, false>(c, a, nullptr, out, n, first, PowF32Fn<S, G, L>::make(v, lane_end), s)
#define SMB_POW_BY_SIGN(S)                                                            \
    (sign == POW_SIGN_REJECT ? SMB_POW_LAUNCH(S, POW_SIGN_REJECT, false)              \
     : sign == POW_SIGN_EVEN ? SMB_POW_LAUNCH(S, POW_SIGN_EVEN, false)                \
                             : SMB_POW_LAUNCH(S, POW_SIGN_ODD, false))
                if (lt1) return SMB_POW_LAUNCH(POW_TIER_SMALL, POW_SIGN_REJECT, true); // 0 < |y| < 1 is never an integer
                if (tier == POW_TIER_SMALL) return SMB_POW_BY_SIGN(POW_TIER_SMALL);
                if (tier == POW_TIER_MEDIUM) return SMB_POW_BY_SIGN(POW_TIER_MEDIUM);
                return SMB_POW_BY_SIGN(POW_TIER_LARGE);
#undef SMB_POW_BY_SIGN
#undef SMB_POW_LAUNCH
            }
        }
    }
    return fail(SMB_ERR_INVALID, "unknown op %d", op);
}
template<>
int scalar_t<double>(const DeviceCtx &c, int op, const double *a, double v, double *out, uint64_t n, uint64_t first,
                     uint64_t lane_end, cudaStream_t s) {
    using T = double;
    switch (op) {
        case SMB_OP_ADD: return launch_stream<T, ScalarFn<OP_ADD, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_SUB: return launch_stream<T, ScalarFn<OP_SUB, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_MUL: return launch_stream<T, ScalarFn<OP_MUL, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_DIV: return launch_stream<T, ScalarFn<OP_DIV, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_POW: {
            if (g_opt_pow_specialise.load()) {
                if (v == 2.0) return launch_stream<T, PowSpecialFn<POWS_SQUARE, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == -1.0) return launch_stream<T, PowSpecialFn<POWS_RECIP, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == 0.5) return launch_stream<T, PowSpecialFn<POWS_SQRT, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
                if (v == 1.0) return launch_stream<T, PowSpecialFn<POWS_IDENT, T>, false>(c, a, nullptr, out, n, first, {lane_end}, s);
            }
            {
                const PowExpF64 pe = classify_exp(v);
                const bool small = pow_f64_small_y(pe), odd = pe.y_is_odd != 0;
#define SMB_POW64_LAUNCH(S, O) launch_stream<T, PowF64Fn<S, O>, false>(c, a, nullptr, out, n, first, PowF64Fn<S, O>::make(v, lane_end), s)
                if (small) return odd ? SMB_POW64_LAUNCH(true, true) : SMB_POW64_LAUNCH(true, false);
                return odd ? SMB_POW64_LAUNCH(false, true) : SMB_POW64_LAUNCH(false, false);
#undef SMB_POW64_LAUNCH
            }
        }
    }
    return fail(SMB_ERR_INVALID, "unknown op %d", op);
}
template<>
int scalar_t<int32_t>(const DeviceCtx &c, int op, const int32_t *a, int32_t v, int32_t *out, uint64_t n, uint64_t first,
                      uint64_t lane_end, cudaStream_t s) {
    using T = int32_t;
    switch (op) {
        case SMB_OP_ADD: return launch_stream<T, ScalarFn<OP_ADD, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_SUB: return launch_stream<T, ScalarFn<OP_SUB, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_MUL: return launch_stream<T, ScalarFn<OP_MUL, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_DIV: return launch_stream<T, ScalarFn<OP_DIV, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
        case SMB_OP_POW: return launch_stream<T, ScalarFn<OP_POW, T>, false>(c, a, nullptr, out, n, first, {v, lane_end}, s);
    }
    return fail(SMB_ERR_INVALID, "unknown op %d", op);
}

// ---- broadcast launch -------------------------------------------------------
static BcastTable make_table(const ElementwisePlan &p, uint64_t lin_base, uint64_t count, uint64_t lane_base, bool *wide) {
    BcastTable t;
    memset(&t, 0, sizeof t);
    t.ndim = p.ndim;
    bool w = lin_base + count > (1ull << 31) || g_opt_force_wide.load() != 0;
    for (int k = 0; k < SMB_MAX_NDIM; ++k) {
        const uint64_t d = k < p.ndim ? p.shape[k] : 1;
        if (d >= (1ull << 31)) w = true;
        t.shape64[k] = d;
        t.sa[k] = k < p.ndim ? p.sa[k] : 0;
        t.sb[k] = k < p.ndim ? p.sb[k] : 0;
        const FastDiv32 f = make_fastdiv32((uint32_t)std::min<uint64_t>(d, 0x7fffffffull));
        t.shape[k] = f.d;
        t.mul[k] = f.mul;
        t.shr[k] = f.shr;
    }
    t.lin_base = lin_base;
    t.count = count;
    t.lane_base = lane_base;
    *wide = w;
    return t;
}

static bool operand_reused(const ElementwisePlan &p, const uint64_t *s) {
    for (int k = 0; k < p.ndim; ++k)
        if (s[k] == 0 && p.shape[k] > 1) return true;
    return false;
}

// Largest vector width (bytes) the row kernel may use for this plan / pointers.
template<typename T>
static int row_vector_bytes(const ElementwisePlan &p, const T *a, const T *b, const T *out, uint64_t lin_base,
                            uint64_t count) {
    const int candidates[2] = {16, (int)sizeof(T)};
    for (int vb : candidates) {
        const uint64_t epv = vb / sizeof(T);
        if (epv == 1) return vb;
        const int m = p.ndim;
        if (p.shape[m - 1] % epv || lin_base % epv || count % epv) continue;
        if ((uintptr_t)out % vb) continue;
        bool ok = true;
        const uint64_t *ss[2] = {p.sa, p.sb};
        const void *pp[2] = {a, b};
        for (int o = 0; o < 2 && ok; ++o) {
            if (ss[o][m - 1] != 1) continue; // inner-broadcast operand: scalar loads, no constraint
            if ((uintptr_t)pp[o] % vb) ok = false;
            for (int k = 0; k < m - 1 && ok; ++k)
                if (ss[o][k] % epv) ok = false;
        }
        if (ok) return vb;
    }
    return (int)sizeof(T);
}

template<typename T, typename Fn>
static int launch_bcast(const DeviceCtx &c, const ElementwisePlan &p, const T *a, const T *b, T *out, uint64_t lin_base,
                        uint64_t count, uint64_t lane_base, Fn fn, cudaStream_t s) {
    if (count == 0) return SMB_OK;
    bool wide = false;
    const BcastTable t = make_table(p, lin_base, count, lane_base, &wide);
    // {D0,1,L} (op) {1,D1,L}: both operands broadcast along different outer dims -> register-tiled
    // outer kernel (whole result or whole dim-0 slabs of it; vector-aligned rows)
    if (p.kind == PLAN_ROW && p.ndim == 3 && p.sa[2] == 1 && p.sb[2] == 1 && g_opt_bcast_variant.load() != 2) {
        const bool a0 = p.sa[1] == 0 && p.sb[0] == 0 && p.sa[0] != 0 && p.sb[1] != 0; // a varies with dim 0
        const bool b0 = p.sb[1] == 0 && p.sa[0] == 0 && p.sb[0] != 0 && p.sa[1] != 0; // b varies with dim 0
        const uint64_t slab = p.shape[1] * p.shape[2];
        constexpr uint64_t epv = 16 / sizeof(T);
        const uint64_t s0 = a0 ? p.sa[0] : p.sb[0], s1 = a0 ? p.sb[1] : p.sa[1];
        if ((a0 || b0) && lin_base % slab == 0 && count % slab == 0 && p.shape[2] % epv == 0 && s0 % epv == 0 &&
            s1 % epv == 0 && (uintptr_t)a % 16 == 0 && (uintptr_t)b % 16 == 0 && (uintptr_t)out % 16 == 0 &&
            p.shape[0] < (1ull << 31) && p.shape[1] < (1ull << 31) && p.shape[2] < (1ull << 31)) {
            constexpr int TI = 4, TJ = 4;
            const uint64_t i_begin = lin_base / slab, i_count = count / slab;
            OuterParams op;
            op.d0 = (uint32_t)i_count;
            op.d1 = (uint32_t)p.shape[1];
            op.len = (uint32_t)p.shape[2];
            op.s0 = s0;
            op.s1 = s1;
            op.lane_base = lane_base;
            const T *pa = a0 ? a + i_begin * p.sa[0] : a;
            const T *pb = a0 ? b : b + i_begin * p.sb[0];
            const uint64_t gy = (op.d1 + TJ - 1) / TJ, gz = (op.d0 + TI - 1) / TI;
            if (gy <= 65535 && gz <= 65535) {
                const dim3 grid((unsigned)((op.len / epv + kThreads - 1) / kThreads), (unsigned)gy, (unsigned)gz);
                if (a0) k_outer<T, Fn, TI, TJ, true><<<grid, kThreads, 0, s>>>(pa, pb, out, op, fn);
                else k_outer<T, Fn, TI, TJ, false><<<grid, kThreads, 0, s>>>(pa, pb, out, op, fn);
                g_last_kernel = "k_outer<4x4>";
                ++g_launches;
                SMB_CK(cudaGetLastError());
                return SMB_OK;
            }
        }
    }
    // transposed operand(s): stride 1 along an earlier dim k, a larger stride along the last dim
    // (SMArray::transpose()) -> shared-memory tile transpose over (k, last).  Whole results only.
    if (p.kind == PLAN_GENERIC && p.ndim >= 2 && lin_base == 0 && count == p.n && g_opt_bcast_variant.load() != 2) {
        const int m = p.ndim;
        auto unit_dim = [&](const uint64_t *st) { // the dim (other than the last) this operand is contiguous in
            if (st[m - 1] <= 1) return -1;
            for (int k = m - 2; k >= 0; --k)
                if (st[k] == 1 && p.shape[k] > 1) return k;
            return -2; // strided along the last dim but contiguous nowhere: not a transpose
        };
        const int ka = unit_dim(p.sa), kb = unit_dim(p.sb);
        const int k = ka >= 0 ? ka : kb;
        const bool at = ka >= 0, bt = kb >= 0;
        uint64_t nbatch = 1, prod[SMB_MAX_NDIM];
        prod[m - 1] = 1;
        for (int d = m - 2; d >= 0; --d) prod[d] = prod[d + 1] * p.shape[d + 1];
        for (int d = 0; d < m - 1; ++d) if (d != k) nbatch *= p.shape[d];
        if (k >= 0 && ka != -2 && kb != -2 && (!at || !bt || ka == kb) && p.shape[m - 1] < (1ull << 31) &&
            p.shape[k] < (1ull << 31) && nbatch <= 65535 && (p.shape[k] + 31) / 32 <= 65535) {
            TileParams tp;
            memset(&tp, 0, sizeof tp);
            tp.rows = (uint32_t)p.shape[k];
            tp.cols = (uint32_t)p.shape[m - 1];
            tp.a_r = p.sa[k]; tp.a_c = p.sa[m - 1];
            tp.b_r = p.sb[k]; tp.b_c = p.sb[m - 1];
            tp.o_r = prod[k];
            int nb = 0;
            for (int d = 0; d < m - 1; ++d) {
                if (d == k) continue;
                const FastDiv32 f = make_fastdiv32((uint32_t)p.shape[d]);
                tp.bshape[nb] = f.d; tp.bmul[nb] = f.mul; tp.bshr[nb] = f.shr;
                tp.ba[nb] = p.sa[d]; tp.bb[nb] = p.sb[d]; tp.bo[nb] = prod[d];
                ++nb;
            }
            tp.nbatch_dims = (uint32_t)nb;
            tp.lane_base = lane_base;
            constexpr int TR = sizeof(T) == 8 ? SMB_TILE_R / 2 : SMB_TILE_R, TC = SMB_TILE_C; // <= 17 KB of shared memory per tile
            const dim3 grid((tp.cols + TC - 1) / TC, (tp.rows + TR - 1) / TR, (unsigned)nbatch);
            if (at && bt) k_tile<T, Fn, true, true, TR, TC><<<grid, kThreads, 0, s>>>(a, b, out, tp, fn);
            else if (at) k_tile<T, Fn, true, false, TR, TC><<<grid, kThreads, 0, s>>>(a, b, out, tp, fn);
            else k_tile<T, Fn, false, true, TR, TC><<<grid, kThreads, 0, s>>>(a, b, out, tp, fn);
            g_last_kernel = "k_tile<transpose>";
            ++g_launches;
            SMB_CK(cudaGetLastError());
            return SMB_OK;
        }
    }
    if (p.kind == PLAN_GENERIC) {
        const unsigned grid = grid_for(count, kThreads, c.sm_count, 32);
        if (wide) k_generic<T, Fn, true><<<grid, kThreads, 0, s>>>(a, b, out, t, fn);
        else k_generic<T, Fn, false><<<grid, kThreads, 0, s>>>(a, b, out, t, fn);
        g_last_kernel = wide ? "k_generic<wide>" : "k_generic";
    } else {
        const int vb = row_vector_bytes<T>(p, a, b, out, lin_base, count);
        const int ar = operand_reused(p, p.sa), br = operand_reused(p, p.sb);
        const uint64_t nvec = count / (vb / sizeof(T));
        // Shared-memory staging of a small reused operand (SMB_OPT_BCAST_VARIANT = 1): kept as an
        // option with its measurement -- on B200 the reused operand already sits in L1/L2 and the
        // staged form is slower (C2: 6.47 vs 7.14 TB/s; profiles/r1_sweep_summary.md), so the
        // default reads it through the caching load path.
        int stage = 0;
        uint32_t stage_elems = 0;
        size_t smem = 0;
        if (g_opt_bcast_variant.load() == 1 && vb == 16 && !wide) {
            const uint64_t lim = 96 * 1024 / sizeof(T);
            if (br && p.extent_b <= lim && (!ar || p.extent_b <= p.extent_a)) { stage = 2; stage_elems = (uint32_t)p.extent_b; }
            else if (ar && p.extent_a <= lim) { stage = 1; stage_elems = (uint32_t)p.extent_a; }
            if (stage) smem = 16 + (size_t)stage_elems * sizeof(T);
        }
        // both operands reused = output much larger than the inputs (outer-product-like, C4): a
        // persistent grid of 32 CTAs/SM measured best; otherwise one tile per CTA
        constexpr int UNROLL = 2;
        const int64_t cap = stage ? 16 : ((ar && br) ? 32 : 0);
        const unsigned grid = grid_for(nvec, (uint64_t)kThreads * UNROLL, c.sm_count, cap);
        if (vb == 16) {
            if (stage == 2) {
                SMB_CK(cudaFuncSetAttribute(k_row<T, Fn, 16, false, UNROLL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_row<T, Fn, 16, false, UNROLL, 2><<<grid, kThreads, smem, s>>>(a, b, out, t, ar, br, stage_elems, fn);
                g_last_kernel = "k_row<vec16,stage_b>";
            } else if (stage == 1) {
                SMB_CK(cudaFuncSetAttribute(k_row<T, Fn, 16, false, UNROLL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_row<T, Fn, 16, false, UNROLL, 1><<<grid, kThreads, smem, s>>>(a, b, out, t, ar, br, stage_elems, fn);
                g_last_kernel = "k_row<vec16,stage_a>";
            } else if (wide) {
                k_row<T, Fn, 16, true, UNROLL, 0><<<grid, kThreads, 0, s>>>(a, b, out, t, ar, br, 0u, fn);
                g_last_kernel = "k_row<vec16,wide>";
            } else {
                k_row<T, Fn, 16, false, UNROLL, 0><<<grid, kThreads, 0, s>>>(a, b, out, t, ar, br, 0u, fn);
                g_last_kernel = "k_row<vec16>";
            }
        } else {
            if (wide) k_row<T, Fn, (int)sizeof(T), true, UNROLL, 0><<<grid, kThreads, 0, s>>>(a, b, out, t, ar, br, 0u, fn);
            else k_row<T, Fn, (int)sizeof(T), false, UNROLL, 0><<<grid, kThreads, 0, s>>>(a, b, out, t, ar, br, 0u, fn);
            g_last_kernel = wide ? "k_row<scalar,wide>" : "k_row<scalar>";
        }
    }
    ++g_launches;
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

template<typename T>
static int bcast_t(const DeviceCtx &c, int op, const ElementwisePlan &p, const T *a, const T *b, T *out, uint64_t lin_base,
                   uint64_t count, uint64_t lane_base, uint64_t lane_end, cudaStream_t s) {
    switch (op) {
        case SMB_OP_ADD: return launch_bcast<T, BinaryFn<OP_ADD, T>>(c, p, a, b, out, lin_base, count, lane_base, {lane_end}, s);
        case SMB_OP_SUB: return launch_bcast<T, BinaryFn<OP_SUB, T>>(c, p, a, b, out, lin_base, count, lane_base, {lane_end}, s);
        case SMB_OP_MUL: return launch_bcast<T, BinaryFn<OP_MUL, T>>(c, p, a, b, out, lin_base, count, lane_base, {lane_end}, s);
        case SMB_OP_DIV: return launch_bcast<T, BinaryFn<OP_DIV, T>>(c, p, a, b, out, lin_base, count, lane_base, {lane_end}, s);
        case SMB_OP_POW: return launch_bcast<T, BinaryFn<OP_POW, T>>(c, p, a, b, out, lin_base, count, lane_base, {lane_end}, s);
    }
    return fail(SMB_ERR_INVALID, "unknown op %d", op);
}

// One elementwise launch on DEVICE-ACCESSIBLE operands.  `a`/`b` address the
// operands of plan `p`; the launch produces flat elements
// [lin_base, lin_base+count) of the plan's result into out[0..count).
static int elementwise_device(const DeviceCtx &c, int op, int dtype, const ElementwisePlan &p, const void *a, const void *b,
                              void *out, uint64_t lin_base, uint64_t count, uint64_t lane_base, uint64_t lane_end,
                              cudaStream_t s) {
    if (p.kind == PLAN_CONTIGUOUS) {
        switch (dtype) {
            case SMB_F32: return contiguous_t<float>(c, op, (const float *)a + lin_base, (const float *)b + lin_base, (float *)out, count, lane_base, lane_end, s);
            case SMB_F64: return contiguous_t<double>(c, op, (const double *)a + lin_base, (const double *)b + lin_base, (double *)out, count, lane_base, lane_end, s);
            case SMB_I32: return contiguous_t<int32_t>(c, op, (const int32_t *)a + lin_base, (const int32_t *)b + lin_base, (int32_t *)out, count, lane_base, lane_end, s);
        }
    } else {
        switch (dtype) {
            case SMB_F32: return bcast_t<float>(c, op, p, (const float *)a, (const float *)b, (float *)out, lin_base, count, lane_base, lane_end, s);
            case SMB_F64: return bcast_t<double>(c, op, p, (const double *)a, (const double *)b, (double *)out, lin_base, count, lane_base, lane_end, s);
            case SMB_I32: return bcast_t<int32_t>(c, op, p, (const int32_t *)a, (const int32_t *)b, (int32_t *)out, lin_base, count, lane_base, lane_end, s);
        }
    }
    return fail(SMB_ERR_INVALID, "unknown dtype %d", dtype);
}

static int scalar_device(const DeviceCtx &c, int op, int dtype, const void *a, const void *scalar, void *out, uint64_t n,
                         uint64_t first, uint64_t lane_end, cudaStream_t s) {
    switch (dtype) {
        case SMB_F32: return scalar_t<float>(c, op, (const float *)a, *(const float *)scalar, (float *)out, n, first, lane_end, s);
        case SMB_F64: return scalar_t<double>(c, op, (const double *)a, *(const double *)scalar, (double *)out, n, first, lane_end, s);
        case SMB_I32: return scalar_t<int32_t>(c, op, (const int32_t *)a, *(const int32_t *)scalar, (int32_t *)out, n, first, lane_end, s);
    }
    return fail(SMB_ERR_INVALID, "unknown dtype %d", dtype);
}

// Where the reference's AVX2 loops stop and scalar Op::apply takes over; only
// int pow can tell (smb_math.cuh).  handle_contiguous_arrays: `i + 8 <= n`
// stepping by simd_width (calculate.h:116-121); array_scalar_op:
// n - n % simd_width (calculate.h:139-140).
static uint64_t contiguous_lane_end(int dtype, uint64_t n) {
    const uint64_t w = dtype == SMB_F64 ? 4 : 8;
    uint64_t i = 0;
    if (n >= 8) i = ((n - 8) / w + 1) * w;
    return i;
}
static uint64_t scalar_lane_end(int dtype, uint64_t n) {
    const uint64_t w = dtype == SMB_F64 ? 4 : 8;
    return n - n % w;
}
// The reference's own fast-path predicate on the UN-coalesced tables
// (calculate.h:10-11, helpers.h:130-139): decides lane vs scalar int-pow
// semantics, nothing else.
static bool reference_takes_contiguous_path(const uint64_t *sa, const uint64_t *sb, const uint64_t *shape, int ndim) {
    if (ndim == 1) return true;
    if (sa[ndim - 1] != 1 || sb[ndim - 1] != 1) return false;
    uint64_t expected = 1;
    for (int i = ndim - 1; i >= 0; --i) {
        if (sa[i] != sb[i] || sa[i] != expected) return false;
        expected *= shape[i];
    }
    return true;
}

// --------------------------------------------------- host-operand staging ---
// Operands in host memory are streamed through HBM in slabs along the leading
// coalesced dim: slab i uses slot i % kSlots (own stream + scratch), so the H2D
// copy of slab i+1, the kernel of slab i and the D2H copy of slab i-1 overlap.
// An operand that does not vary along the leading dim (stride 0 there) is
// uploaded once.  Device/managed operands are used in place.
static int elementwise_staged(DeviceCtx &c, int op, int dtype, const ElementwisePlan &p, const void *a, MemType ta,
                              const void *b, MemType tb, void *out, MemType to, uint64_t lane_end) {
    const size_t es = esize(dtype);
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    const uint64_t rows = p.shape[0];
    const uint64_t inner = p.n / rows; // result elements per leading index
    const bool a_var = p.ndim > 1 ? p.sa[0] != 0 : p.sa[0] != 0;
    const bool b_var = p.sb[0] != 0;
    // extent of one leading-index slice of each operand (elements)
    auto slice_extent = [&](const uint64_t *s) {
        uint64_t e = 1;
        for (int k = 1; k < p.ndim; ++k) e += (p.shape[k] - 1) * s[k];
        return e;
    };
    const uint64_t ea1 = slice_extent(p.sa), eb1 = slice_extent(p.sb);
    const uint64_t chunk_bytes = (uint64_t)std::max<int64_t>(g_opt_chunk_bytes.load(), 1 << 16);
    uint64_t chunk_rows = std::max<uint64_t>(1, chunk_bytes / std::max<uint64_t>(1, inner * es));
    chunk_rows = std::min(chunk_rows, rows);
    const uint64_t nchunks = (rows + chunk_rows - 1) / chunk_rows;
    const int nslots = (int)std::min<uint64_t>(kSlots, nchunks);

    // invariant operands: upload once on slot 0, everyone else waits on the event
    Scratch inv_a, inv_b;
    const void *da_inv = a, *db_inv = b;
    bool need_ev = false;
    if (on_host(ta) && !a_var) {
        if (int rc = inv_a.get(p.extent_a * es, dev)) return rc;
        SMB_CK(cudaMemcpyAsync(inv_a.p, a, p.extent_a * es, cudaMemcpyHostToDevice, c.slot[0]));
        da_inv = inv_a.p;
        need_ev = true;
    }
    if (on_host(tb) && !b_var) {
        if (int rc = inv_b.get(p.extent_b * es, dev)) return rc;
        SMB_CK(cudaMemcpyAsync(inv_b.p, b, p.extent_b * es, cudaMemcpyHostToDevice, c.slot[0]));
        db_inv = inv_b.p;
        need_ev = true;
    }
    if (need_ev) {
        SMB_CK(cudaEventRecord(c.ev, c.slot[0]));
        for (int i = 1; i < nslots; ++i) SMB_CK(cudaStreamWaitEvent(c.slot[i], c.ev, 0));
    }
    const uint64_t slab_ea = a_var ? (chunk_rows - 1) * p.sa[0] + ea1 : 0;
    const uint64_t slab_eb = b_var ? (chunk_rows - 1) * p.sb[0] + eb1 : 0;
    Scratch sa_[kSlots], sb_[kSlots], so_[kSlots];
    for (int i = 0; i < nslots; ++i) {
        if (on_host(ta) && a_var) if (int rc = sa_[i].get(slab_ea * es, dev)) return rc;
        if (on_host(tb) && b_var) if (int rc = sb_[i].get(slab_eb * es, dev)) return rc;
        if (on_host(to)) if (int rc = so_[i].get(chunk_rows * inner * es, dev)) return rc;
    }
    for (uint64_t ci = 0; ci < nchunks; ++ci) {
        const int sl = (int)(ci % kSlots);
        cudaStream_t s = c.slot[sl];
        const uint64_t r0 = ci * chunk_rows, r = std::min(chunk_rows, rows - r0);
        ElementwisePlan sub = p;
        sub.shape[0] = r;
        sub.n = r * inner;
        const char *pa = (const char *)da_inv, *pb = (const char *)db_inv;
        if (a_var) {
            const char *src = (const char *)a + r0 * p.sa[0] * es;
            if (on_host(ta)) {
                SMB_CK(cudaMemcpyAsync(sa_[sl].p, src, ((r - 1) * p.sa[0] + ea1) * es, cudaMemcpyHostToDevice, s));
                pa = (const char *)sa_[sl].p;
            } else pa = src;
        }
        if (b_var) {
            const char *src = (const char *)b + r0 * p.sb[0] * es;
            if (on_host(tb)) {
                SMB_CK(cudaMemcpyAsync(sb_[sl].p, src, ((r - 1) * p.sb[0] + eb1) * es, cudaMemcpyHostToDevice, s));
                pb = (const char *)sb_[sl].p;
            } else pb = src;
        }
        char *po = on_host(to) ? (char *)so_[sl].p : (char *)out + r0 * inner * es;
        if (int rc = elementwise_device(c, op, dtype, sub, pa, pb, po, 0, sub.n, r0 * inner, lane_end, s)) return rc;
        if (on_host(to))
            SMB_CK(cudaMemcpyAsync((char *)out + r0 * inner * es, po, sub.n * es, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < nslots; ++i) SMB_CK(cudaStreamSynchronize(c.slot[i]));
    if (need_ev && nslots == 0) SMB_CK(cudaStreamSynchronize(c.slot[0]));
    return SMB_OK;
}

static int scalar_staged(DeviceCtx &c, int op, int dtype, const void *a, MemType ta, const void *scalar, void *out,
                         MemType to, uint64_t n, uint64_t lane_end) {
    const size_t es = esize(dtype);
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    const uint64_t chunk_bytes = (uint64_t)std::max<int64_t>(g_opt_chunk_bytes.load(), 1 << 16);
    const uint64_t chunk = std::min<uint64_t>(n, std::max<uint64_t>(1, chunk_bytes / es));
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    const int nslots = (int)std::min<uint64_t>(kSlots, nchunks);
    Scratch sa_[kSlots], so_[kSlots];
    for (int i = 0; i < nslots; ++i) {
        if (on_host(ta)) if (int rc = sa_[i].get(chunk * es, dev)) return rc;
        if (on_host(to)) if (int rc = so_[i].get(chunk * es, dev)) return rc;
    }
    for (uint64_t ci = 0; ci < nchunks; ++ci) {
        const int sl = (int)(ci % kSlots);
        cudaStream_t s = c.slot[sl];
        const uint64_t i0 = ci * chunk, cnt = std::min(chunk, n - i0);
        const char *pa = (const char *)a + i0 * es;
        if (on_host(ta)) {
            SMB_CK(cudaMemcpyAsync(sa_[sl].p, pa, cnt * es, cudaMemcpyHostToDevice, s));
            pa = (const char *)sa_[sl].p;
        }
        char *po = on_host(to) ? (char *)so_[sl].p : (char *)out + i0 * es;
        if (int rc = scalar_device(c, op, dtype, pa, scalar, po, cnt, i0, lane_end, s)) return rc;
        if (on_host(to)) SMB_CK(cudaMemcpyAsync((char *)out + i0 * es, po, cnt * es, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < nslots; ++i) SMB_CK(cudaStreamSynchronize(c.slot[i]));
    return SMB_OK;
}

static int check_args(int op, int dtype) {
    if (op < SMB_OP_ADD || op > SMB_OP_POW) return fail(SMB_ERR_INVALID, "unknown op %d", op);
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d (float, double, int32 only)", dtype);
    return SMB_OK;
}

static int elementwise_entry(int op, int dtype, const void *a, const uint64_t *stride_a, const void *b,
                             const uint64_t *stride_b, const uint64_t *shape, int ndim, uint64_t lin_begin,
                             uint64_t lin_count, bool whole, void *out, void *stream) {
    if (int rc = check_args(op, dtype)) return rc;
    if (ndim < 1 || ndim > SMB_MAX_NDIM) return fail(SMB_ERR_INVALID, "rank %d outside 1..%d", ndim, SMB_MAX_NDIM);
    if (!stride_a || !stride_b || !shape) return fail(SMB_ERR_INVALID, "null shape / stride table");
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    const ElementwisePlan p = make_plan(stride_a, stride_b, shape, ndim);
    if (whole) { lin_begin = 0; lin_count = p.n; }
    if (lin_begin > p.n || lin_count > p.n - lin_begin) return fail(SMB_ERR_INVALID, "flat range outside the result");
    if (lin_count == 0) return SMB_OK;
    if (!a || !b || !out) return fail(SMB_ERR_INVALID, "null operand pointer");
    uint64_t lane_end = 0;
    if (op == SMB_OP_POW && dtype == SMB_I32 && reference_takes_contiguous_path(stride_a, stride_b, shape, ndim))
        lane_end = contiguous_lane_end(dtype, p.n);
    const MemType ta = mem_type(a), tb = mem_type(b), to = mem_type(out);
    const size_t es = esize(dtype);
    if (on_host(ta) || on_host(tb) || on_host(to)) {
        if (lin_begin != 0 || lin_count != p.n) {
            // partial range with host operands: stage the touched operands whole
            int dev = 0;
            SMB_CK(cudaGetDevice(&dev));
            Scratch da, db, dout;
            const void *pa = a, *pb = b;
            void *po = out;
            cudaStream_t s = c->main;
            if (on_host(ta)) { if (int rc = da.get(p.extent_a * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(da.p, a, p.extent_a * es, cudaMemcpyHostToDevice, s)); pa = da.p; }
            if (on_host(tb)) { if (int rc = db.get(p.extent_b * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(db.p, b, p.extent_b * es, cudaMemcpyHostToDevice, s)); pb = db.p; }
            if (on_host(to)) { if (int rc = dout.get(lin_count * es, dev)) return rc; po = dout.p; }
            if (int rc = elementwise_device(*c, op, dtype, p, pa, pb, po, lin_begin, lin_count, lin_begin, lane_end, s)) return rc;
            if (on_host(to)) SMB_CK(cudaMemcpyAsync(out, po, lin_count * es, cudaMemcpyDeviceToHost, s));
            SMB_CK(cudaStreamSynchronize(s));
            return SMB_OK;
        }
        return elementwise_staged(*c, op, dtype, p, a, ta, b, tb, out, to, lane_end);
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (ta == MT_MANAGED) prefetch_managed(a, p.extent_a * es, s);
    if (tb == MT_MANAGED) prefetch_managed(b, p.extent_b * es, s);
    if (to == MT_MANAGED) prefetch_managed(out, lin_count * es, s);
    if (int rc = elementwise_device(*c, op, dtype, p, a, b, out, lin_begin, lin_count, lin_begin, lane_end, s)) return rc;
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}

// SMArray::operator% (reference math/product.h).  `result` is one host T.  Operands on the host
// are staged whole; the partial / ticket / result scratch lives in one pooled device block.
template<typename T>
static int dot_t(DeviceCtx &c, const T *a, const T *b, uint64_t n, void *result, cudaStream_t s) {
    using A = typename DotAcc<T>::type;
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    constexpr int UNROLL = 4;
    const uint64_t nvec = n / (16 / sizeof(T));
    // many waves of short-lived CTAs (8 grid-stride iterations each): the hardware scheduler evens out the SMs,
    // which a resident grid with a static split cannot (the slowest SM would set the time)
    const unsigned grid = grid_for(nvec ? nvec : 1, (uint64_t)kThreads * UNROLL * 8, c.sm_count, 0);
    Scratch scratch;
    const size_t bytes = 16 + sizeof(A) * ((size_t)grid + 1);
    if (int rc = scratch.get(bytes, dev)) return rc;
    unsigned int *ticket = (unsigned int *)scratch.p;
    A *res = (A *)((char *)scratch.p + 8);
    A *partials = (A *)((char *)scratch.p + 16);
    SMB_CK(cudaMemsetAsync(scratch.p, 0, 16, s));
    k_dot<T, UNROLL><<<grid, kThreads, 0, s>>>(a, b, n, partials, ticket, res);
    ++g_launches;
    g_last_kernel = "k_dot";
    SMB_CK(cudaGetLastError());
    SMB_CK(cudaMemcpyAsync(result, res, sizeof(A), cudaMemcpyDeviceToHost, s));
    SMB_CK(cudaStreamSynchronize(s)); // a scalar result: always synchronous
    return SMB_OK;
}

} // namespace smb

using namespace smb;

// =============================================================== C ABI =====
// ---- op-chain fusion ---------------------------------
template<typename T>
static int chain_launch(const DeviceCtx &c, const ChainPlan &p, const smb_chain_step *steps, const void *const *data,
                        uint64_t lin_begin, uint64_t lin_count, uint64_t lane_end, T *out, cudaStream_t s) {
    ChainTable t;
    memset(&t, 0, sizeof t);
    t.ndim = p.ndim;
    t.nsteps = p.nleaf;
    const bool wide = p.n > 0x7fffffffull || g_opt_force_wide.load() != 0;
    for (int k = 0; k < SMB_MAX_NDIM; ++k) {
        t.shape64[k] = p.shape[k];
        if (!wide) {
            const FastDiv32 f = make_fastdiv32((uint32_t)p.shape[k]);
            t.shape[k] = f.d; t.mul[k] = f.mul; t.shr[k] = f.shr;
        }
    }
    constexpr int EPVV = 16 / (int)sizeof(T);
    bool vec = p.inner_unit_or_zero && p.shape[p.ndim - 1] % EPVV == 0 && lin_begin % EPVV == 0 && lin_count % EPVV == 0 &&
               (uintptr_t)out % 16 == 0;
    for (int i = 0; i < p.nleaf; ++i) {
        t.data[i] = data[i];
        // leaf (op) acc: only - and / care which side the leaf is on
        t.op[i] = (uint8_t)(steps[i].swap && steps[i].op == SMB_OP_SUB ? CH_RSUB : steps[i].swap && steps[i].op == SMB_OP_DIV ? CH_RDIV : steps[i].op);
        for (int k = 0; k < SMB_MAX_NDIM; ++k) t.stride[i][k] = p.stride[i][k];
        if (!data[i]) {
            if (sizeof(T) == 8) memcpy(&t.cbits[i], &steps[i].value.f64, 8);
            else { uint32_t w; memcpy(&w, &steps[i].value.f32, 4); t.cbits[i] = w; } // f32 and i32 share the low word
        } else if (p.stride[i][p.ndim - 1] == 1) {
            if ((uintptr_t)data[i] % 16 != 0) vec = false;
            for (int k = 0; k + 1 < p.ndim; ++k) if (p.stride[i][k] % EPVV != 0) vec = false;
        }
    }
    t.lin_base = lin_begin;
    t.count = lin_count;
    t.lane_end = lane_end;
    const uint64_t items = vec ? lin_count / EPVV : lin_count;
    // f32 pow steps with an exponent the table-driven core takes: stage its tables (vector variant)
    bool powfast = false;
    t.pow_consts = pow_consts();
    t.pow_small = 1;
    if constexpr (std::is_same<T, float>::value) {
        for (int i = 1; i < p.nleaf && vec; ++i) {
            if (steps[i].op != SMB_OP_POW) continue;
            const PowExpF32 pe = classify_exp(steps[i].value.f32);
            if (!pow_f32_fast_ok(pe)) continue;
            t.pow_fast[i] = 1;
            t.pow_abs_mask[i] = pe.y_is_int ? 0x7fffffffu : 0xffffffffu;
            t.pow_sign_or[i] = pe.y_is_odd ? 0x80000000u : 0u;
            if (pow_f32_tier(pe) == POW_TIER_LARGE) t.pow_small = 0;
            powfast = true;
        }
    }
    t.tiles_per_cta = powfast ? 32 : 1; // amortise the 24 KB table copy, stay many waves deep
    // compiled-in chain capacity / vectors per thread: short chains keep more loads in flight
#define SMB_CHAIN_LAUNCH(E, W, NS, U, PF, ND)                                                                     \
    k_chain<T, E, W, NS, U, PF, ND><<<grid_for(items, (uint64_t)kThreads * U * t.tiles_per_cta, c.sm_count, 0), kThreads, 0, s>>>(out, t)
#define SMB_CHAIN_BY_LEN(E, W, PF, ND)                                                            \
    do {                                                                                          \
        if (p.nleaf <= 3 && PF) SMB_CHAIN_LAUNCH(E, W, 3, 1, PF, ND); /* pow: fewer registers, more CTAs */ \
        else if (p.nleaf <= 3) SMB_CHAIN_LAUNCH(E, W, 3, 4, PF, ND);                              \
        else if (p.nleaf <= 5) SMB_CHAIN_LAUNCH(E, W, 5, 2, PF, ND);                              \
        else SMB_CHAIN_LAUNCH(E, W, 8, 1, PF, ND);                                                \
    } while (0)
    // rank 1 needs no division (one 64-bit index); rank 2 and the general case come in 32- / 64-bit index forms
#define SMB_CHAIN_BY_RANK(E, PF)                                                                  \
    do {                                                                                          \
        if (p.ndim == 1) SMB_CHAIN_BY_LEN(E, true, PF, 1);                                        \
        else if (p.ndim == 2 && wide) SMB_CHAIN_BY_LEN(E, true, PF, 2);                           \
        else if (p.ndim == 2) SMB_CHAIN_BY_LEN(E, false, PF, 2);                                  \
        else if (wide) SMB_CHAIN_BY_LEN(E, true, PF, 0);                                          \
        else SMB_CHAIN_BY_LEN(E, false, PF, 0);                                                   \
    } while (0)
    if (vec && powfast) {
        if constexpr (std::is_same<T, float>::value) SMB_CHAIN_BY_RANK(EPVV, true);
        g_last_kernel = wide ? "k_chain<vec16,wide,pow>" : "k_chain<vec16,pow>";
    } else if (vec) {
        SMB_CHAIN_BY_RANK(EPVV, false);
        g_last_kernel = wide ? "k_chain<vec16,wide>" : "k_chain<vec16>";
    } else {
        if (wide) SMB_CHAIN_BY_LEN(1, true, false, 0);
        else SMB_CHAIN_BY_LEN(1, false, false, 0);
        g_last_kernel = wide ? "k_chain<scalar,wide>" : "k_chain<scalar>";
    }
#undef SMB_CHAIN_BY_RANK
#undef SMB_CHAIN_BY_LEN
#undef SMB_CHAIN_LAUNCH
    ++g_launches;
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

static int chain_entry(int dtype, const smb_chain_step *steps, int nsteps, const uint64_t *shape, int ndim,
                       uint64_t lin_begin, uint64_t lin_count, bool whole, void *out, void *stream) {
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d (float, double, int32 only)", dtype);
    if (!steps || nsteps < 1 || nsteps > SMB_CHAIN_MAX) return fail(SMB_ERR_INVALID, "chain of %d steps (1..%d)", nsteps, SMB_CHAIN_MAX);
    if (ndim < 1 || ndim > SMB_MAX_NDIM || !shape) return fail(SMB_ERR_INVALID, "rank %d outside 1..%d", ndim, SMB_MAX_NDIM);
    const uint64_t *strides[SMB_CHAIN_MAX];
    for (int i = 0; i < nsteps; ++i) {
        strides[i] = steps[i].data ? steps[i].stride : nullptr;
        if (i == 0) continue;
        if (steps[i].op < SMB_OP_ADD || steps[i].op > SMB_OP_POW) return fail(SMB_ERR_INVALID, "step %d: unknown op %d", i, steps[i].op);
        if (steps[i].op == SMB_OP_POW && (steps[i].data || steps[i].swap))
            return fail(SMB_ERR_INVALID, "step %d: pow in a chain takes a constant exponent on the right (array ^ scalar)", i);
    }
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    const ChainPlan p = make_chain_plan(strides, nsteps, shape, ndim);
    if (whole) { lin_begin = 0; lin_count = p.n; }
    if (lin_begin > p.n || lin_count > p.n - lin_begin) return fail(SMB_ERR_INVALID, "flat range outside the result");
    if (lin_count == 0) return SMB_OK;
    if (!out) return fail(SMB_ERR_INVALID, "null result pointer");
    const size_t es = esize(dtype);
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    // host leaves / result are staged whole through pooled scratch (no overlap: the fused path is
    // meant for resident arrays); managed blocks are prefetched like everywhere else
    Scratch scratch[SMB_CHAIN_MAX], dout;
    const void *data[SMB_CHAIN_MAX];
    for (int i = 0; i < nsteps; ++i) {
        data[i] = steps[i].data;
        if (!data[i]) continue;
        const MemType mt = mem_type(data[i]);
        if (on_host(mt)) {
            if (stream) return fail(SMB_ERR_INVALID, "smb_chain: host operands need the synchronous form (stream == NULL)");
            if (int rc = scratch[i].get(p.extent[i] * es, dev)) return rc;
            SMB_CK(cudaMemcpyAsync(scratch[i].p, data[i], p.extent[i] * es, cudaMemcpyDefault, s));
            data[i] = scratch[i].p;
        } else if (mt == MT_MANAGED) prefetch_managed(data[i], p.extent[i] * es, s);
    }
    void *po = out;
    const MemType to = mem_type(out);
    if (on_host(to)) {
        if (stream) return fail(SMB_ERR_INVALID, "smb_chain: a host result needs the synchronous form (stream == NULL)");
        if (int rc = dout.get(lin_count * es, dev)) return rc;
        po = dout.p;
    } else if (to == MT_MANAGED) prefetch_managed(out, lin_count * es, s);
    const uint64_t lane_end = dtype == SMB_I32 ? scalar_lane_end(dtype, p.n) : 0; // array_scalar_op on the dense intermediate
    int rc;
    switch (dtype) {
        case SMB_F32: rc = chain_launch<float>(*c, p, steps, data, lin_begin, lin_count, lane_end, (float *)po, s); break;
        case SMB_F64: rc = chain_launch<double>(*c, p, steps, data, lin_begin, lin_count, lane_end, (double *)po, s); break;
        default: rc = chain_launch<int32_t>(*c, p, steps, data, lin_begin, lin_count, lane_end, (int32_t *)po, s); break;
    }
    if (rc) return rc;
    if (on_host(to)) SMB_CK(cudaMemcpyAsync(out, po, lin_count * es, cudaMemcpyDefault, s));
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}
extern "C" {

int smb_elementwise(int op, int dtype, const void *a, const uint64_t *stride_a, const void *b, const uint64_t *stride_b,
                    const uint64_t *shape, int ndim, uint64_t n, void *out, void *stream) {
    if (shape && ndim >= 1 && ndim <= SMB_MAX_NDIM) {
        uint64_t prod = 1;
        for (int k = 0; k < ndim; ++k) prod *= shape[k];
        if (prod != n) return fail(SMB_ERR_INVALID, "n (%llu) != prod(shape) (%llu)", (unsigned long long)n, (unsigned long long)prod);
    }
    return elementwise_entry(op, dtype, a, stride_a, b, stride_b, shape, ndim, 0, 0, true, out, stream);
}

int smb_elementwise_range(int op, int dtype, const void *a, const uint64_t *stride_a, const void *b,
                          const uint64_t *stride_b, const uint64_t *shape, int ndim, uint64_t lin_begin,
                          uint64_t lin_count, void *out, void *stream) {
    return elementwise_entry(op, dtype, a, stride_a, b, stride_b, shape, ndim, lin_begin, lin_count, false, out, stream);
}

int smb_contiguous(int op, int dtype, const void *a, const void *b, void *out, uint64_t n, void *stream) {
    const uint64_t one = 1, shape = n;
    return elementwise_entry(op, dtype, a, &one, b, &one, &shape, 1, 0, 0, true, out, stream);
}

int smb_array_scalar(int op, int dtype, const void *a, const void *scalar, uint64_t n, void *out, void *stream) {
    if (int rc = check_args(op, dtype)) return rc;
    if (!scalar) return fail(SMB_ERR_INVALID, "null scalar pointer");
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (n == 0) return SMB_OK;
    if (!a || !out) return fail(SMB_ERR_INVALID, "null operand pointer");
    const uint64_t lane_end = scalar_lane_end(dtype, n);
    const MemType ta = mem_type(a), to = mem_type(out);
    const size_t es = esize(dtype);
    if (on_host(ta) || on_host(to)) return scalar_staged(*c, op, dtype, a, ta, scalar, out, to, n, lane_end);
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (ta == MT_MANAGED) prefetch_managed(a, n * es, s);
    if (to == MT_MANAGED) prefetch_managed(out, n * es, s);
    if (int rc = scalar_device(*c, op, dtype, a, scalar, out, n, 0, lane_end, s)) return rc;
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}

int smb_dot(int dtype, const void *a, const void *b, uint64_t n, void *result, void *stream) {
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d", dtype);
    if (!result) return fail(SMB_ERR_INVALID, "null result pointer");
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    const size_t es = esize(dtype);
    if (n == 0) { memset(result, 0, es); return SMB_OK; }
    if (!a || !b) return fail(SMB_ERR_INVALID, "null operand pointer");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    const MemType ta = mem_type(a), tb = mem_type(b);
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    Scratch da, db;
    if (on_host(ta)) { if (int rc = da.get(n * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(da.p, a, n * es, cudaMemcpyDefault, s)); a = da.p; }
    else if (ta == MT_MANAGED) prefetch_managed(a, n * es, s);
    if (on_host(tb)) { if (int rc = db.get(n * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(db.p, b, n * es, cudaMemcpyDefault, s)); b = db.p; }
    else if (tb == MT_MANAGED) prefetch_managed(b, n * es, s);
    if (((uintptr_t)a | (uintptr_t)b) & 15) return fail(SMB_ERR_INVALID, "smb_dot: operands must be 16-byte aligned");
    switch (dtype) {
        case SMB_F32: return dot_t<float>(*c, (const float *)a, (const float *)b, n, result, s);
        case SMB_F64: return dot_t<double>(*c, (const double *)a, (const double *)b, n, result, s);
        default: return dot_t<int32_t>(*c, (const int32_t *)a, (const int32_t *)b, n, result, s);
    }
}

int smb_chain(int dtype, const smb_chain_step *steps, int nsteps, const uint64_t *shape, int ndim, uint64_t n, void *out,
              void *stream) {
    if (shape && ndim >= 1 && ndim <= SMB_MAX_NDIM) {
        uint64_t prod = 1;
        for (int k = 0; k < ndim; ++k) prod *= shape[k];
        if (prod != n) return fail(SMB_ERR_INVALID, "n = %llu is not the product of the shape (%llu)", (unsigned long long)n, (unsigned long long)prod);
    }
    return chain_entry(dtype, steps, nsteps, shape, ndim, 0, 0, true, out, stream);
}
int smb_chain_range(int dtype, const smb_chain_step *steps, int nsteps, const uint64_t *shape, int ndim, uint64_t lin_begin,
                    uint64_t lin_count, void *out, void *stream) {
    return chain_entry(dtype, steps, nsteps, shape, ndim, lin_begin, lin_count, false, out, stream);
}

void *smb_alloc(size_t bytes, int kind) {
    if (kind < SMB_MEM_DEVICE || kind > SMB_MEM_PINNED) { fail(SMB_ERR_INVALID, "unknown memory kind %d", kind); return nullptr; }
    DeviceCtx *c = nullptr;
    if (current_ctx(&c)) return nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t e;
    void *p = Pool::instance().alloc(bytes, kind, dev, &e);
    if (!p) fail(e == cudaErrorMemoryAllocation ? SMB_ERR_OOM : SMB_ERR_CUDA, "smb_alloc(%zu, kind %d): %s", bytes, kind, cudaGetErrorString(e));
    return p;
}

int smb_free(void *ptr) {
    if (!ptr) return SMB_OK;
    if (!Pool::instance().free(ptr)) return fail(SMB_ERR_INVALID, "smb_free: %p was not returned by smb_alloc", ptr);
    return SMB_OK;
}

int smb_host_written(const void *ptr) {
    if (ptr) Pool::instance().set_host_flag(ptr, true);
    return SMB_OK;
}

int smb_owns(const void *ptr) { return ptr && Pool::instance().owns(ptr) ? 1 : 0; }

int smb_pool_trim(void) {
    cudaDeviceSynchronize();
    Pool::instance().trim();
    return SMB_OK;
}

int smb_pool_stats(uint64_t stats[4]) {
    if (!stats) return fail(SMB_ERR_INVALID, "null stats");
    Pool::instance().stats(stats);
    return SMB_OK;
}

int smb_fill(int dtype, void *out, const void *value, uint64_t n, void *stream) {
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d", dtype);
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (n == 0) return SMB_OK;
    if (!out || !value) return fail(SMB_ERR_INVALID, "null pointer");
    const MemType to = mem_type(out);
    if (on_host(to)) return fail(SMB_ERR_INVALID, "smb_fill needs device or managed memory");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (to == MT_MANAGED) prefetch_managed(out, n * esize(dtype), s);
    const unsigned grid = grid_for(n, kThreads * 4, c->sm_count, 16);
    if (dtype == SMB_F64) k_fill<double><<<grid, kThreads, 0, s>>>((double *)out, n, *(const double *)value);
    else k_fill<uint32_t><<<grid, kThreads, 0, s>>>((uint32_t *)out, n, *(const uint32_t *)value);
    ++g_launches;
    SMB_CK(cudaGetLastError());
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}

int smb_prefetch(const void *ptr, size_t bytes, int device, void *stream) {
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (mem_type(ptr) != MT_MANAGED) return SMB_OK; // nothing to migrate
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    SMB_CK(cudaMemPrefetchAsync(ptr, bytes, device < 0 ? cudaCpuDeviceId : device, s));
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}

int smb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int smb_set_device(int device) {
    SMB_CK(cudaSetDevice(device));
    return SMB_OK;
}
int smb_get_device(void) {
    int d = -1;
    if (cudaGetDevice(&d) != cudaSuccess) { cudaGetLastError(); return -1; }
    return d;
}
int smb_sync(void) {
    SMB_CK(cudaDeviceSynchronize());
    return SMB_OK;
}

int smb_set_option(int key, int64_t value) {
    switch (key) {
        case SMB_OPT_POW_SPECIALISE: g_opt_pow_specialise = value ? 1 : 0; return SMB_OK;
        case SMB_OPT_STAGE_CHUNK_BYTES: g_opt_chunk_bytes = value; return SMB_OK;
        case SMB_OPT_CONTIG_VARIANT: g_opt_contig_variant = value; return SMB_OK;
        case SMB_OPT_BCAST_VARIANT: g_opt_bcast_variant = value; return SMB_OK;
        case SMB_OPT_FORCE_WIDE_INDEX: g_opt_force_wide = value ? 1 : 0; return SMB_OK;
    }
    return fail(SMB_ERR_INVALID, "unknown option %d", key);
}
int64_t smb_get_option(int key) {
    switch (key) {
        case SMB_OPT_POW_SPECIALISE: return g_opt_pow_specialise;
        case SMB_OPT_STAGE_CHUNK_BYTES: return g_opt_chunk_bytes;
        case SMB_OPT_CONTIG_VARIANT: return g_opt_contig_variant;
        case SMB_OPT_BCAST_VARIANT: return g_opt_bcast_variant;
        case SMB_OPT_FORCE_WIDE_INDEX: return g_opt_force_wide;
    }
    return -1;
}

uint64_t smb_launch_count(void) { return g_launches.load(); }
const char *smb_last_kernel(void) { return g_last_kernel; }
const char *smb_last_error(void) { return g_err.c_str(); }
const char *smb_version(void) { return "smb200 0.1 (sm_100a)"; }

int smb_plan_chain(const smb_chain_step *steps, int nsteps, const uint64_t *shape, int ndim, int *out_ndim,
                   uint64_t *out_shape, uint64_t *out_strides) {
    if (!steps || nsteps < 1 || nsteps > SMB_CHAIN_MAX || ndim < 1 || ndim > SMB_MAX_NDIM || !shape) return -SMB_ERR_INVALID;
    const uint64_t *strides[SMB_CHAIN_MAX];
    for (int i = 0; i < nsteps; ++i) strides[i] = steps[i].data ? steps[i].stride : nullptr;
    const ChainPlan p = make_chain_plan(strides, nsteps, shape, ndim);
    if (out_ndim) *out_ndim = p.ndim;
    for (int k = 0; k < p.ndim; ++k)
        if (out_shape) out_shape[k] = p.shape[k];
    if (out_strides)
        for (int i = 0; i < nsteps; ++i)
            for (int k = 0; k < SMB_MAX_NDIM; ++k) out_strides[i * SMB_MAX_NDIM + k] = p.stride[i][k];
    return p.inner_unit_or_zero ? 1 : 0;
}

int smb_plan_elementwise(const uint64_t *stride_a, const uint64_t *stride_b, const uint64_t *shape, int ndim, int elem_size,
                         int *out_ndim, uint64_t *out_shape, uint64_t *out_stride_a, uint64_t *out_stride_b) {
    (void)elem_size;
    if (ndim < 1 || ndim > SMB_MAX_NDIM || !stride_a || !stride_b || !shape) return -SMB_ERR_INVALID;
    const ElementwisePlan p = make_plan(stride_a, stride_b, shape, ndim);
    if (out_ndim) *out_ndim = p.ndim;
    for (int k = 0; k < p.ndim; ++k) {
        if (out_shape) out_shape[k] = p.shape[k];
        if (out_stride_a) out_stride_a[k] = p.sa[k];
        if (out_stride_b) out_stride_b[k] = p.sb[k];
    }
    return p.kind;
}

int smb_fill_uniform_f32(void *out, uint64_t first, uint64_t n, uint64_t seed, float lo, float hi, void *stream) {
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (n == 0) return SMB_OK;
    if (!out || on_host(mem_type(out))) return fail(SMB_ERR_INVALID, "smb_fill_uniform_f32 needs device or managed memory");
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    const unsigned grid = grid_for(n, kThreads * 4, c->sm_count, 16);
    k_fill_uniform_f32<<<grid, kThreads, 0, s>>>((float *)out, first, n, seed, lo, hi);
    ++g_launches;
    SMB_CK(cudaGetLastError());
    if (!stream) SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}

} // extern "C"
