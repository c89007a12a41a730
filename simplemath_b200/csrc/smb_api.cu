// smb_api.cu -- the C ABI of libsmb200.so (declared in include/smb200.h):
// device runtime, pooled storage, planner, launchers and the host-operand
// staging pipeline.  See DESIGN.md for the data-flow picture.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <utility>
#include <string>
#include <vector>
#include <algorithm>

#include "../../include/smb200.h"
#include "smb_alloc.h"
#include "smb_kernels.cuh"
#include "smb_plan.h"
#include "smb_shard.h"

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
static std::atomic<int64_t> g_opt_chain_pow_variant{4}; // fused pow chains of <= 3 leaves: 4 the pow kernel with a pre-operator; k_chain forms: 0 U1, 1 U1+prefetch, 2 U2, 3 U2+prefetch
static std::atomic<int64_t> g_opt_pow_tail{0}; // single-tile CTAs at the end of a pow grid (0: none, the default)
static std::atomic<int64_t> g_opt_pool_max_cached{64ll << 30}; // cached (free) pool bytes beyond which smb_free trims

// Every copy / prefetch / memset / event wait the library enqueues bumps this counter.  The overlapping launch form
// reasons about KERNELS only (what earlier kernels read and write, and that each kernel's completion implies its
// predecessor's); a kernel that follows anything else on its stream -- a replica copy it is about to read, a prefetch --
// is launched plainly: full stream order, no attribute.
static std::atomic<uint64_t> g_other_ops{0};
static inline void note_other_op() { g_other_ops.fetch_add(1, std::memory_order_relaxed); }

// ------------------------------------------------------ device context ------
constexpr int kSlots = 3;       // staging pipeline depth (H2D | kernel | D2H in flight)
constexpr int kMaxDevices = 64;
// Accesses of the launches enqueued on a stream since its last fully serialised launch: what a new
// launch must not touch if it is to overlap them (programmatic dependent launch, see pdl_mode()).
struct Span { uintptr_t lo, hi; };
struct StreamTrack {
    static constexpr int kCap = 24;
    Span reads[kCap], writes[kCap];
    int nr = 0, nw = 0;
    uint64_t other_ops_seen = ~0ull; // g_other_ops when the stream's last overlappable launch was decided
    void reset() { nr = nw = 0; }
};
struct DeviceCtx {
    std::atomic<bool> ready{false};
    int device = -1;
    int sm_count = 0;
    cudaStream_t main = nullptr;
    cudaStream_t slot[kSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev = nullptr;
    cudaEvent_t ev_user = nullptr; // orders the library's private streams after the caller's stream
    cudaEvent_t ev_done = nullptr; // async mode, several devices: end of this device's part of the last operator
    bool dirty = false;            // async mode: work enqueued on `main` since the last synchronisation
    StreamTrack track;             // of `main`
    std::mutex launch_mu;          // decision + launch on this device's streams are one unit (one lock per device: the launcher threads run side by side)
};
static DeviceCtx g_ctx[kMaxDevices];
static std::mutex g_ctx_mu;

static void destroy_ctx_handles(DeviceCtx &c) {
    if (c.main) cudaStreamDestroy(c.main);
    for (int i = 0; i < kSlots; ++i) if (c.slot[i]) cudaStreamDestroy(c.slot[i]);
    if (c.ev) cudaEventDestroy(c.ev);
    if (c.ev_user) cudaEventDestroy(c.ev_user);
    if (c.ev_done) cudaEventDestroy(c.ev_done);
    c.main = nullptr;
    for (int i = 0; i < kSlots; ++i) c.slot[i] = nullptr;
    c.ev = c.ev_user = c.ev_done = nullptr;
    cudaGetLastError();
}
static int init_ctx(DeviceCtx &c, int dev) { // g_ctx_mu held, `dev` current
    c.device = dev;
    SMB_CK(cudaDeviceGetAttribute(&c.sm_count, cudaDevAttrMultiProcessorCount, dev));
    SMB_CK(cudaStreamCreateWithFlags(&c.main, cudaStreamNonBlocking));
    for (int i = 0; i < kSlots; ++i) SMB_CK(cudaStreamCreateWithFlags(&c.slot[i], cudaStreamNonBlocking));
    SMB_CK(cudaEventCreateWithFlags(&c.ev, cudaEventDisableTiming));
    SMB_CK(cudaEventCreateWithFlags(&c.ev_user, cudaEventDisableTiming));
    SMB_CK(cudaEventCreateWithFlags(&c.ev_done, cudaEventDisableTiming));
    // ready-made shared-memory images of the f32 pow tables (one bulk copy per CTA later); a
    // __device__ global has one instance per device, so every device builds its own
    k_pow_image_init<<<8, kBlock, 0, c.main>>>();
    SMB_CK(cudaGetLastError());
    SMB_CK(cudaStreamSynchronize(c.main));
    return SMB_OK;
}

static int device_count_checked(int *count) {
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess || *count <= 0) {
        cudaGetLastError();
        return fail(SMB_ERR_NO_DEVICE, "no CUDA device available (%s); libsmb200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    return SMB_OK;
}

// The context of device `dev` (streams, events, pow table image), created on first use.  Leaves
// `dev` the CURRENT device when it had to initialise; callers that hop between devices restore.
static int ctx_of(int dev, DeviceCtx **out) {
    if (dev < 0 || dev >= kMaxDevices) return fail(SMB_ERR_INVALID, "device index %d out of range", dev);
    DeviceCtx &c = g_ctx[dev];
    if (!c.ready.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_ctx_mu);
        if (!c.ready.load(std::memory_order_relaxed)) {
            SMB_CK(cudaSetDevice(dev));
            if (int rc = init_ctx(c, dev)) { destroy_ctx_handles(c); return rc; } // nothing half-made survives a failed init
            c.ready.store(true, std::memory_order_release);
        }
    }
    *out = &c;
    return SMB_OK;
}

// There is no CPU fallback: every compute entry point goes through here and
// fails loudly when no CUDA device is usable.
static void devices_from_env_once();
static int current_ctx(DeviceCtx **out) {
    int count = 0;
    if (int rc = device_count_checked(&count)) return rc;
    devices_from_env_once();
    int dev = 0;
    SMB_CK(cudaGetDevice(&dev));
    return ctx_of(dev, out);
}

// Scoped "make `dev` current", restoring the caller's device (the sharded launchers hop).
struct DeviceScope {
    int saved = -1;
    DeviceScope() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
    int set(int dev) { SMB_CK(cudaSetDevice(dev)); return SMB_OK; }
    ~DeviceScope() { if (saved >= 0) cudaSetDevice(saved); }
};

// ------------------------------------------------ one launcher thread per device ----
// The default (synchronous) mode of a device set used to walk the devices from the calling thread: set device, prepare
// operands, launch -- about 6 us per device -- and then wait for the streams one after another.  With 8 GPUs that is
// ~50 us of host time around kernels that take 20-40 us per device on the broadcast configs (C2 x 16: 3.7x, C4: 2.4x one GPU),
// and it held the streams to 6.4-6.8x.  Each device of the set gets a persistent launcher thread that stays on its
// device: the calling thread hands every range to its device's thread and waits for all of them; a launcher prepares,
// launches AND waits for its stream, so the eight waits overlap.  Launchers spin briefly for the next operator before they
// block, so back-to-back operators find them awake.  (Async mode keeps the single-thread walk: nothing waits there.)
static std::atomic<int64_t> g_opt_launcher_threads{1};
struct LaunchWorker {
    int dev = -1;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::atomic<int> queued{0};
};
static LaunchWorker *g_workers[kMaxDevices]; // created once per device (g_set_mu), never destroyed: they outlive static destruction
static void worker_main(LaunchWorker *w) {
    if (cudaSetDevice(w->dev) != cudaSuccess) cudaGetLastError();
    for (;;) {
        std::function<void()> job;
        for (int spin = 0; spin < 20000 && w->queued.load(std::memory_order_acquire) == 0; ++spin) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        {
            std::unique_lock<std::mutex> lk(w->mu);
            w->cv.wait(lk, [&] { return !w->q.empty(); });
            job = std::move(w->q.front());
            w->q.pop_front();
            w->queued.fetch_sub(1, std::memory_order_release);
        }
        job();
    }
}
static LaunchWorker *worker_of(int dev) { // g_set_mu held by the caller
    if (!g_workers[dev]) {
        LaunchWorker *w = new LaunchWorker;
        w->dev = dev;
        w->th = std::thread(worker_main, w);
        w->th.detach();
        g_workers[dev] = w;
    }
    return g_workers[dev];
}
static void worker_post(LaunchWorker *w, std::function<void()> job) {
    {
        std::lock_guard<std::mutex> lk(w->mu);
        w->q.push_back(std::move(job));
        w->queued.fetch_add(1, std::memory_order_release);
    }
    w->cv.notify_one();
}

// ------------------------------------------------------- the device set -----
// smb_set_devices: the devices an operator on MANAGED arrays is spread over (the
// broadcast output's flat index range is split, contiguous operands are split by the same ranges,
// broadcast operands are replicated).  Empty / one entry = the calling thread's current device,
// exactly the single-GPU behaviour.  SMB_DEVICES ("all", "0-7", "0,2,5") presets it for programs
// that only know the reference's operator API.
static std::mutex g_set_mu;
static std::vector<int> g_devices;
static std::atomic<int> g_ndevices{0};
static std::once_flag g_env_once;

static int set_devices_locked(const int *devs, int n) {
    int count = 0;
    if (int rc = device_count_checked(&count)) return rc;
    if (n < 0 || n > kMaxDevices || (n > 0 && !devs)) return fail(SMB_ERR_INVALID, "smb_set_devices: bad device list");
    for (int i = 0; i < n; ++i) {
        // (a device may be listed more than once: it then owns several ranges, each handled like a
        // device of its own -- how the single-GPU tests exercise the whole sharded path)
        if (devs[i] < 0 || devs[i] >= count) return fail(SMB_ERR_INVALID, "smb_set_devices: device %d of %d does not exist", devs[i], count);
    }
    DeviceScope scope;
    for (int i = 0; i < n; ++i) { // contexts up front; peer access so a device may read a neighbour's pages in place
        DeviceCtx *c = nullptr;
        if (int rc = ctx_of(devs[i], &c)) return rc;
        if (n > 1) {
            SMB_CK(cudaSetDevice(devs[i]));
            for (int j = 0; j < n; ++j) {
                if (devs[i] == devs[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) == cudaSuccess && can) {
                    const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); // best effort
                    else cudaGetLastError();
                }
            }
        }
    }
    if (n > 1) for (int i = 0; i < n; ++i) worker_of(devs[i]); // one launcher thread per device of the set
    g_devices.assign(devs, devs + n);
    g_ndevices.store(n);
    return SMB_OK;
}
static void devices_from_env_once() {
    std::call_once(g_env_once, [] {
        const char *e = getenv("SMB_DEVICES");
        if (!e || !*e) return;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return; }
        std::vector<int> list;
        if (!strcmp(e, "all")) { for (int i = 0; i < count; ++i) list.push_back(i); }
        else {
            const char *p = e;
            while (*p) {
                char *end = nullptr;
                long lo = strtol(p, &end, 10), hi = lo;
                if (end == p) break;
                p = end;
                if (*p == '-') { hi = strtol(p + 1, &end, 10); if (end == p + 1) break; p = end; }
                for (long d = lo; d <= hi && d < count; ++d) list.push_back((int)d);
                if (*p == ',') ++p;
            }
        }
        std::lock_guard<std::mutex> lk(g_set_mu);
        if (g_devices.empty() && !list.empty() && set_devices_locked(list.data(), (int)list.size()) != SMB_OK)
            fprintf(stderr, "smb200: SMB_DEVICES=%s ignored: %s\n", e, g_err.c_str());
    });
}
static std::vector<int> active_devices() {
    if (g_ndevices.load() <= 1) return {};
    std::lock_guard<std::mutex> lk(g_set_mu);
    return g_devices;
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
bool Pool::take_placement(const void *ptr, uint64_t want, Block *out, bool *matched, int rm_action, bool *was_rm) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) return false;
    --it;
    Block &b = it->second;
    if ((uintptr_t)ptr >= (uintptr_t)b.base + b.bytes) return false;
    *out = b;
    *matched = b.placement == want;
    b.placement = want;
    if (was_rm) *was_rm = b.read_mostly;
    if (rm_action >= 0) b.read_mostly = rm_action != 0;
    return true;
}
void Pool::clear_placement(const void *ptr) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = live_.upper_bound((uintptr_t)ptr);
    if (it == live_.begin()) return;
    --it;
    Block &b = it->second;
    if ((uintptr_t)ptr < (uintptr_t)b.base + b.bytes) b.placement = 0;
}
void Pool::trim_to(uint64_t keep_bytes) {
    std::lock_guard<std::mutex> lk(mu_);
    while (cached_bytes_ > keep_bytes && !free_.empty()) {
        auto big = free_.end();
        for (auto it = free_.begin(); it != free_.end(); ++it)
            if (!it->second.empty() && (big == free_.end() || it->first.bytes > big->first.bytes)) big = it;
        if (big == free_.end()) break;
        void *p = big->second.back();
        big->second.pop_back();
        if (big->second.empty()) free_.erase(big);
        auto c = cached_.find((uintptr_t)p);
        cached_bytes_ -= c->second.bytes;
        raw_free(c->second);
        cached_.erase(c);
    }
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
    Scratch() = default;
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
    ~Scratch() { if (p) Pool::instance().free(p); }
    int get(size_t bytes, int device, int kind = SMB_MEM_DEVICE) {
        cudaError_t e;
        p = Pool::instance().alloc(bytes, kind, device, &e);
        if (!p) return fail(e == cudaErrorMemoryAllocation ? SMB_ERR_OOM : SMB_ERR_CUDA, "scratch alloc of %zu bytes: %s",
                            bytes, cudaGetErrorString(e));
        return SMB_OK;
    }
};

// Declared AFTER the Scratch blocks of a staged call (so it runs before their destructors): whatever
// way the call leaves -- an SMB_CK early return included -- the streams that may still be reading or
// writing those blocks are drained before the blocks go back to the pool.
struct DrainGuard {
    cudaStream_t s[4] = {nullptr, nullptr, nullptr, nullptr};
    int n = 0;
    void add(cudaStream_t st) { if (n < 4) s[n++] = st; }
    ~DrainGuard() {
        for (int i = 0; i < n; ++i)
            if (cudaStreamSynchronize(s[i]) != cudaSuccess) cudaGetLastError();
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
// elsewhere.  cudaMemPrefetchAsync costs ~50 us even for resident pages (measured: 164 us per
// 3-operand call), so pool blocks record where the launchers last put them (smb_alloc.h); foreign
// managed memory is always prefetched.
static inline uint64_t mix64(uint64_t h, uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); return h; }
static inline uint64_t placement_single(int dev) { return 0x5100000000000000ull | (uint64_t)(dev + 1); }
// A block that becomes a RESULT loses the read-mostly advice an earlier use as a shared operand left on it
// (writes to read-duplicated pages work, but every one of them invalidates the duplicates first).
static void drop_read_mostly(const Block &blk) {
    if (cudaMemAdvise(blk.base, blk.bytes, cudaMemAdviseUnsetReadMostly, 0) != cudaSuccess) cudaGetLastError();
}
static void prefetch_managed(const void *p, size_t bytes, int dev, cudaStream_t s, bool is_result = false) {
    Block blk;
    bool matched = false, was_rm = false;
    if (Pool::instance().take_placement(p, placement_single(dev), &blk, &matched, is_result ? 0 : -1, &was_rm)) {
        if (is_result && was_rm) drop_read_mostly(blk);
        if (matched) return;
        p = blk.base;        // whole block: views of it become resident too
        bytes = blk.bytes;
    }
    note_other_op();
    if (cudaMemPrefetchAsync(p, bytes, dev, s) != cudaSuccess) cudaGetLastError(); // best effort
}

// ------------------------------------------- programmatic dependent launch --
// Back-to-back launches on one stream normally pay a launch gap plus a ramp: the next grid's first
// CTA is scheduled only after the previous grid's last CTA has retired and its memory is flushed
// (~15 us of a 250 us kernel at the 8-GPU shard size).  With the programmatic-stream-serialisation
// launch attribute the next grid's CTAs become resident as soon as every CTA of the previous grid
// has STARTED (each calls griddepcontrol.launch_dependents first thing) and fill SM slots as the
// previous grid's tail drains.  What they may do there depends on the data:
//   * a launch that touches nothing the launches before it (since the stream's last serialised
//     point) write, and writes nothing they read, runs right away and executes
//     griddepcontrol.wait only at its END -- so that ITS completion still implies theirs;
//   * any other launch executes griddepcontrol.wait first (the wait returns once the previous grid
//     has completed and flushed): it only saves the scheduling ramp.
// The host knows every operand range of its own launches, so it decides; SMB_OPT_PDL = 0 turns the
// attribute off (plain stream order).
static std::atomic<int64_t> g_opt_pdl{1};
// The overlapping form is used on the library's PRIVATE stream only (stream == NULL calls): there
// every operation is ours, and a kernel that follows anything but one of our kernels (a copy, a
// prefetch) is launched plainly.  On a caller's stream the library cannot know what precedes it, so
// launches there are plain unless the caller opts in (SMB_OPT_PDL = 2: only this library's kernels,
// events and ordinary copies are enqueued on the streams it is handed), and then always wait first.

struct PdlDecision { bool attr; uint32_t flags; };
static inline bool spans_overlap(const Span &x, const Span &y) { return x.lo < y.hi && y.lo < x.hi; }
static PdlDecision pdl_decide(StreamTrack &t, const Span *reads, int nr, const Span &write) {
    if (!g_opt_pdl.load(std::memory_order_relaxed)) { t.reset(); return {false, kPdlWaitFirst}; }
    const uint64_t ops = g_other_ops.load(std::memory_order_relaxed);
    if (ops != t.other_ops_seen) { // something that is not one of our kernels may sit right before this launch: plain launch
        t.other_ops_seen = ops;
        t.reset();
        for (int j = 0; j < nr; ++j) if (reads[j].hi > reads[j].lo) t.reads[t.nr++] = reads[j];
        t.writes[t.nw++] = write;
        return {false, kPdlWaitFirst};
    }
    bool conflict = t.nr + nr > StreamTrack::kCap || t.nw + 1 > StreamTrack::kCap;
    for (int i = 0; i < t.nw && !conflict; ++i) {
        if (spans_overlap(t.writes[i], write)) conflict = true;
        for (int j = 0; j < nr && !conflict; ++j) if (spans_overlap(t.writes[i], reads[j])) conflict = true;
    }
    for (int i = 0; i < t.nr && !conflict; ++i) if (spans_overlap(t.reads[i], write)) conflict = true;
    if (conflict) t.reset(); // this launch waits first: when its work starts, everything before it is done
    for (int j = 0; j < nr; ++j) if (reads[j].hi > reads[j].lo) t.reads[t.nr++] = reads[j];
    t.writes[t.nw++] = write;
    return {true, conflict ? kPdlWaitFirst : 0u};
}
// A launch of ours that does not take part (no griddepcontrol in the kernel): plain stream order --
// it starts after everything before it has completed, and nothing starts before it has.
static inline void pdl_barrier(StreamTrack &t) { t.reset(); }

struct LaunchLock {
    std::unique_lock<std::mutex> lk;
    StreamTrack *t; // nullptr: a caller's stream
    LaunchLock(DeviceCtx &c, cudaStream_t s) : lk(c.launch_mu), t(s == c.main ? &c.track : nullptr) {}
    PdlDecision decide(const Span *reads, int nr, const Span &write) {
        if (t) return pdl_decide(*t, reads, nr, write);
        // a caller's stream: the attribute only when the caller vouches for what precedes us there (SMB_OPT_PDL = 2),
        // and then always with the initial wait
        return {g_opt_pdl.load(std::memory_order_relaxed) >= 2, kPdlWaitFirst};
    }
    void barrier() { if (t) pdl_barrier(*t); }
};

template<typename... KArgs, typename... Args>
static cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, unsigned block, size_t smem, cudaStream_t s, bool pdl_attr,
                             Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    if (pdl_attr) {
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
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

// vectors in flight per thread: the library default, unless the functor carries two operand streams through the pow loop
template<typename Fn, typename = void> struct fn_unroll : std::integral_constant<int, SMB_STREAM_UNROLL> {};
template<typename Fn> struct fn_unroll<Fn, std::void_t<decltype(Fn::UNROLL_OVERRIDE)>> : std::integral_constant<int, Fn::UNROLL_OVERRIDE> {};

static inline Span span_of(const void *p, uint64_t bytes) { return Span{(uintptr_t)p, (uintptr_t)p + bytes}; }

// Consecutive tiles per CTA of the table-driven pow kernels.  Enough to amortise the table fill (one
// bulk copy for f32, an in-kernel fill of 40 KB for f64) and the first tile's unhidden load latency,
// few enough that the grid stays several waves deep: a grid of resident CTAs measured 10-15 % slower.
// At the 8-GPU shard size (2^27 elements) the default of 8 still gives nine waves and measured best
// (profiles/r2_pow_grid_sweep.md: 6376 vs 6274 / 6140 GB/s for 4 / 3 tiles at full clock); only arrays
// small enough to leave fewer than four waves get fewer tiles per CTA.
static int64_t pow_tiles_per_cta(uint64_t full_tiles, int sm_count, int resident_per_sm, int64_t dflt) {
    const int64_t cps = g_opt_contig_variant.load();
    if (cps > 0) return cps;
    const uint64_t per_wave = (uint64_t)sm_count * (uint64_t)resident_per_sm;
    const uint64_t fit = full_tiles / (per_wave * 4);
    return (int64_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)dflt, fit));
}

template<typename T, typename Fn, bool HAS_B>
static int launch_stream(DeviceCtx &c, const T *a, const T *b, T *out, uint64_t n, uint64_t first, Fn fn,
                         cudaStream_t s) {
    if (n == 0) return SMB_OK;
    constexpr int VB = SMB_STREAM_VB, UNROLL = fn_unroll<Fn>::value;
    const uintptr_t ma = (uintptr_t)a % VB, mb = HAS_B ? (uintptr_t)b % VB : ma, mo = (uintptr_t)out % VB;
    const int64_t cps = g_opt_contig_variant.load();
    const Span reads[2] = {span_of(a, n * sizeof(T)), span_of(HAS_B ? b : a, n * sizeof(T))};
    const Span write = span_of(out, n * sizeof(T));
    LaunchLock ll(c, s);
    if (ma == mb && ma == mo && ma % sizeof(T) == 0) {
        uint64_t head = ma ? (VB - ma) / sizeof(T) : 0;
        if (head > n) head = n;
        if (head) { // peel up to the first common vector boundary (views give interior pointers)
            k_stream_unaligned<T, Fn, HAS_B><<<1, kThreads, 0, s>>>(a, b, out, head, first, fn);
            ++g_launches;
            note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
        }
        const uint64_t rest = n - head;
        if (rest) {
            constexpr uint64_t per_block = (uint64_t)kThreads * UNROLL * (VB / sizeof(T));
            // Plain streams run one tile per CTA (tools/sweep: fastest on B200); the pow kernels a few
            // consecutive tiles per CTA (pow_tiles_per_cta).  SMB_OPT_CONTIG_VARIANT overrides either.
            unsigned grid;
            uint32_t n_small = 0;
            if (fn_pow_tables<Fn>::value && SMB_POW_BLOCKED) {
                const int resident = sizeof(T) == 4 ? SMB_POW_MIN_BLOCKS : SMB_POW64_MIN_BLOCKS;
                const uint64_t full_tiles = rest / per_block;
                const int64_t tpc = pow_tiles_per_cta(full_tiles, c.sm_count, resident,
                                                      (sizeof(T) == 4 ? SMB_POW_TILES_PER_CTA : 4 * SMB_POW_TILES_PER_CTA) * SMB_STREAM_UNROLL / UNROLL);
                // Single-tile CTAs at the end of the grid (SMB_OPT_POW_TAIL_CTAS) to fill the ragged end of the
                // multi-tile phase: built, measured, and OFF by default -- with the tile count per CTA already
                // shrunk for small arrays it changes nothing up to one wave of them and loses beyond
                // (profiles/r2_pow_grid_sweep.md).
                const int64_t tail_opt = g_opt_pow_tail.load();
                uint64_t small = tail_opt > 0 ? (uint64_t)tail_opt : 0;
                small = std::min<uint64_t>(std::min<uint64_t>(small, full_tiles / 4), (1u << 24) - 1);
                if (tpc <= 1) small = 0;
                const uint64_t big_tiles = full_tiles - small;
                const uint64_t big = (big_tiles + (uint64_t)tpc - 1) / (uint64_t)tpc;
                n_small = (uint32_t)small;
                grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(big + small, 0x7fffffffull));
                if ((uint64_t)grid != std::max<uint64_t>(1, big + small)) { n_small = 0; grid = grid_for(rest, per_block * (uint64_t)tpc, c.sm_count, 0); }
            } else {
                grid = grid_for(rest, per_block, c.sm_count, fn_pow_tables<Fn>::value && cps == 0 ? 8 : cps);
            }
            const PdlDecision d = ll.decide(reads, HAS_B ? 2 : 1, write);
            SMB_CK(launch_ex(k_stream<T, Fn, HAS_B, VB, UNROLL>, dim3(grid), kThreads, 0, s, d.attr, a + head,
                             HAS_B ? b + head : (const T *)nullptr, out + head, rest, first + head, fn,
                             d.flags | (n_small << kPdlFlagBits)));
            ++g_launches;
            g_last_kernel = HAS_B ? "k_stream<binary>" : "k_stream<scalar>";
        }
    } else {
        const unsigned grid = grid_for(n, kThreads, c.sm_count, 32);
        k_stream_unaligned<T, Fn, HAS_B><<<grid, kThreads, 0, s>>>(a, b, out, n, first, fn);
        ++g_launches;
        note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
        g_last_kernel = "k_stream_unaligned";
    }
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

template<typename T>
static int contiguous_t(DeviceCtx &c, int op, const T *a, const T *b, T *out, uint64_t n, uint64_t first,
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
static int scalar_t(DeviceCtx &c, int op, const T *a, T v, T *out, uint64_t n, uint64_t first, uint64_t lane_end,
                    cudaStream_t s);

template<>
int scalar_t<float>(DeviceCtx &c, int op, const float *a, float v, float *out, uint64_t n, uint64_t first,
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
int scalar_t<double>(DeviceCtx &c, int op, const double *a, double v, double *out, uint64_t n, uint64_t first,
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
int scalar_t<int32_t>(DeviceCtx &c, int op, const int32_t *a, int32_t v, int32_t *out, uint64_t n, uint64_t first,
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
static int launch_bcast(DeviceCtx &c, const ElementwisePlan &p, const T *a, const T *b, T *out, uint64_t lin_base,
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
                note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
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
            note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
            SMB_CK(cudaGetLastError());
            return SMB_OK;
        }
    }
    // small constant inner strides (w[:, ::2], every third column): vector loads + register picks
    if (p.kind == PLAN_GENERIC && g_opt_bcast_variant.load() != 2) {
        constexpr uint64_t epv = 16 / sizeof(T);
        const int m = p.ndim;
        bool ok = p.sa[m - 1] <= epv && p.sb[m - 1] <= epv && p.shape[m - 1] % epv == 0 && lin_base % epv == 0 && count % epv == 0 &&
                  (uintptr_t)out % 16 == 0 && (uintptr_t)a % sizeof(T) == 0 && (uintptr_t)b % sizeof(T) == 0;
        for (int k = 0; k + 1 < m && ok; ++k) ok = p.sa[k] % epv == 0 && p.sb[k] % epv == 0;
        if (ok) {
            const int pha = (int)(((uintptr_t)a % 16) / sizeof(T)), phb = (int)(((uintptr_t)b % 16) / sizeof(T));
            // 256-bit loads for an operand whose every vector span starts on 32 bytes: even inner stride (a span is
            // 16 * stride bytes from the previous one), aligned-down base and all outer strides on 32 bytes
            auto spans32 = [&](const T *ptr, const uint64_t *st) {
                if (st[m - 1] % 2 != 0 || st[m - 1] == 0) return 0;
                if (((uintptr_t)ptr & ~(uintptr_t)15) % 32 != 0) return 0;
                for (int k = 0; k + 1 < m; ++k) if (st[k] % (2 * epv) != 0) return 0;
                return 1; // (the inner part of a span's offset is 16 * stride * q bytes: on 32 for every q when the stride is even)
            };
            const int wide32 = spans32(a, p.sa) | (spans32(b, p.sb) << 1);
            const unsigned grid = grid_for(count / epv, kThreads, c.sm_count, 0);
            if (wide) k_sgather<T, Fn, true><<<grid, kThreads, 0, s>>>(a, b, out, t, pha, phb, wide32, fn);
            else k_sgather<T, Fn, false><<<grid, kThreads, 0, s>>>(a, b, out, t, pha, phb, wide32, fn);
            g_last_kernel = wide ? "k_sgather<wide>" : "k_sgather";
            ++g_launches;
            note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
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
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

template<typename T>
static int bcast_t(DeviceCtx &c, int op, const ElementwisePlan &p, const T *a, const T *b, T *out, uint64_t lin_base,
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

// ------------------------------------------------ user-defined device Ops ----
// The reference's "Extending with Custom Operations" recipe (README.md:86-133: an Op struct with apply /
// apply_simd, then element_wise_op<T, MyOp<T>> from an operator) on the device, without patching this
// library: the user's .cu includes include/smb200_plugin.cuh, which instantiates the SAME kernel
// templates (k_stream, k_row, k_generic) over a functor that calls MyOp<T>::apply_device, and registers
// three launchers under a name.  From then on the op id goes through every path a built-in op takes --
// planning, host-operand staging, views, flat-range sharding over a device set, async mode.
constexpr int kUserOpBase = SMB_OP_USER;
struct UserOpEntry {
    std::string name;
    smb_user_op fn[3];
    bool have[3] = {false, false, false};
};
static std::mutex g_user_mu;
static std::vector<UserOpEntry> g_user_ops;
static bool user_op_lookup(int op, int dtype, smb_user_op *out) {
    std::lock_guard<std::mutex> lk(g_user_mu);
    const int i = op - kUserOpBase;
    if (i < 0 || i >= (int)g_user_ops.size() || dtype < 0 || dtype > 2 || !g_user_ops[i].have[dtype]) return false;
    *out = g_user_ops[i].fn[dtype];
    return true;
}
static int user_rc(int e, const char *what) {
    if (e == 0) return SMB_OK;
    cudaGetLastError();
    return fail(SMB_ERR_CUDA, "user op %s launch: %s", what, cudaGetErrorString((cudaError_t)e));
}
static int user_contiguous(DeviceCtx &c, int op, int dtype, const void *a, const void *b, void *out, uint64_t n, cudaStream_t s) {
    smb_user_op u;
    if (!user_op_lookup(op, dtype, &u)) return fail(SMB_ERR_INVALID, "op %d is not registered for dtype %d", op, dtype);
    const smb_launch_env env{s, c.sm_count, c.device};
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    g_last_kernel = "user<k_stream>";
    return user_rc(u.contiguous(&env, a, b, out, n), "contiguous");
}
static int user_scalar(DeviceCtx &c, int op, int dtype, const void *a, const void *scalar, void *out, uint64_t n, cudaStream_t s) {
    smb_user_op u;
    if (!user_op_lookup(op, dtype, &u)) return fail(SMB_ERR_INVALID, "op %d is not registered for dtype %d", op, dtype);
    const smb_launch_env env{s, c.sm_count, c.device};
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    g_last_kernel = "user<k_stream,scalar>";
    return user_rc(u.scalar(&env, a, scalar, out, n), "scalar");
}
template<typename T>
static int row_vector_bytes(const ElementwisePlan &p, const T *a, const T *b, const T *out, uint64_t lin_base, uint64_t count);
static BcastTable make_table(const ElementwisePlan &p, uint64_t lin_base, uint64_t count, uint64_t lane_base, bool *wide);
static bool operand_reused(const ElementwisePlan &p, const uint64_t *s);
static int user_strided(DeviceCtx &c, int op, int dtype, const ElementwisePlan &p, const void *a, const void *b, void *out,
                        uint64_t lin_base, uint64_t count, cudaStream_t s) {
    smb_user_op u;
    if (!user_op_lookup(op, dtype, &u)) return fail(SMB_ERR_INVALID, "op %d is not registered for dtype %d", op, dtype);
    if (count == 0) return SMB_OK;
    bool wide = false;
    const BcastTable t = make_table(p, lin_base, count, lin_base, &wide);
    int vb = dtype == SMB_F64 ? 8 : 4;
    if (p.kind == PLAN_ROW) {
        if (dtype == SMB_F64) vb = row_vector_bytes<double>(p, (const double *)a, (const double *)b, (const double *)out, lin_base, count);
        else vb = row_vector_bytes<float>(p, (const float *)a, (const float *)b, (const float *)out, lin_base, count);
    }
    const smb_launch_env env{s, c.sm_count, c.device};
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    g_last_kernel = p.kind == PLAN_GENERIC ? "user<k_generic>" : "user<k_row>";
    return user_rc(u.strided(&env, a, b, out, &t, (int)sizeof t, p.kind == PLAN_GENERIC, wide, vb, operand_reused(p, p.sa), operand_reused(p, p.sb)), "strided");
}

// One elementwise launch on DEVICE-ACCESSIBLE operands.  `a`/`b` address the
// operands of plan `p`; the launch produces flat elements
// [lin_base, lin_base+count) of the plan's result into out[0..count).
static int elementwise_device(DeviceCtx &c, int op, int dtype, const ElementwisePlan &p, const void *a, const void *b,
                              void *out, uint64_t lin_base, uint64_t count, uint64_t lane_base, uint64_t lane_end,
                              cudaStream_t s) {
    if (op >= kUserOpBase) {
        const size_t es = esize(dtype);
        if (p.kind == PLAN_CONTIGUOUS)
            return user_contiguous(c, op, dtype, (const char *)a + lin_base * es, (const char *)b + lin_base * es, out, count, s);
        return user_strided(c, op, dtype, p, a, b, out, lin_base, count, s);
    }
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

static int scalar_device(DeviceCtx &c, int op, int dtype, const void *a, const void *scalar, void *out, uint64_t n,
                         uint64_t first, uint64_t lane_end, cudaStream_t s) {
    if (op >= kUserOpBase) return user_scalar(c, op, dtype, a, scalar, out, n, s);
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

// ------------------------------------------------------- asynchronous mode ---
// SMB_OPT_ASYNC = 1 (opt-in; sm::async_scope in the C++ headers): a call with stream == NULL enqueues
// on the device's private stream and RETURNS -- the result hand-off of SURVEY.md §8f rank 4.  The
// reference's contract (results complete on return) is what costs 15 us per call
// around a 2 us kernel at the launch-bound sizes (benchmark/add.cpp:21-29).  Results are complete
// after smb_sync() / smb_wait_pending(); the C++ headers call the latter before any host access.
// Order is kept by the stream: every async call of a device goes to the same private stream, and a
// pool block that is freed and handed out again is reused on that stream.  With several devices, an
// operator's part on device d also waits for what the OTHER devices still have in flight, unless it
// is the same partition of pure streams as the operator before it (every device then touches only
// its own ranges).
static std::atomic<int64_t> g_opt_async{0};
static std::atomic<bool> g_pending{false};
static std::mutex g_async_mu;
static uint64_t g_dirty_mask = 0;  // devices with un-synchronised async work (g_async_mu)
static uint64_t g_last_sig = 0;    // partition signature of the last async operator, 0: none / not a pure partition

static inline bool async_mode(const void *stream) { return !stream && g_opt_async.load(std::memory_order_relaxed) != 0; }

// Before enqueueing an operator's part on each device of `devs`: cross-device order (see above).
// Events are recorded lazily, here, on the devices someone has to wait for -- the common case (one
// device) never records or waits.
static int async_order(const int *devs, int n, uint64_t sig) {
    std::lock_guard<std::mutex> lk(g_async_mu);
    if (sig != 0 && sig == g_last_sig) return SMB_OK;
    g_last_sig = sig;
    uint64_t recorded = 0;
    for (int i = 0; i < n; ++i) {
        const uint64_t others = g_dirty_mask & ~(1ull << devs[i]);
        for (int e = 0; others >> e; ++e) {
            if (!((others >> e) & 1ull)) continue;
            if (!((recorded >> e) & 1ull)) { SMB_CK(cudaEventRecord(g_ctx[e].ev_done, g_ctx[e].main)); recorded |= 1ull << e; }
            note_other_op();
            SMB_CK(cudaStreamWaitEvent(g_ctx[devs[i]].main, g_ctx[e].ev_done, 0));
        }
    }
    return SMB_OK;
}
// After enqueueing: these devices now have work in flight.
static int async_mark(const int *devs, int n) {
    std::lock_guard<std::mutex> lk(g_async_mu);
    for (int i = 0; i < n; ++i) {
        g_ctx[devs[i]].dirty = true;
        g_dirty_mask |= 1ull << devs[i];
    }
    g_pending.store(true, std::memory_order_release);
    return SMB_OK;
}
// End of a stream == NULL call on one device: the reference's synchronous contract, or the hand-off.
static int finish_call(DeviceCtx &c, cudaStream_t s, const void *user_stream) {
    if (user_stream) return SMB_OK;
    if (async_mode(user_stream)) return async_mark(&c.device, 1);
    SMB_CK(cudaStreamSynchronize(s));
    return SMB_OK;
}
// Start of a stream == NULL call on ONE device in async mode: wait for other devices' pending work.
static int begin_call(DeviceCtx &c, const void *user_stream) {
    if (!async_mode(user_stream)) return SMB_OK;
    return async_order(&c.device, 1, 0);
}
static int sync_all() {
    {
        std::lock_guard<std::mutex> lk(g_async_mu);
        for (int d = 0; d < kMaxDevices; ++d) {
            DeviceCtx &c = g_ctx[d];
            if (!c.ready.load(std::memory_order_acquire)) continue;
            SMB_CK(cudaStreamSynchronize(c.main));
            c.dirty = false;
        }
        g_dirty_mask = 0;
        g_last_sig = 0;
        g_pending.store(false, std::memory_order_release);
    }
    return SMB_OK;
}

// The library's private copy / compute streams of a STAGED (host-operand) call start after the
// work already enqueued on the caller's stream -- or, in async mode, on the private main stream --
// so pinned inputs an earlier async operation produces are complete before the first H2D copy.
static int order_slots_after(DeviceCtx &c, cudaStream_t after, int nslots) {
    if (!after) return SMB_OK;
    SMB_CK(cudaEventRecord(c.ev_user, after));
    note_other_op();
    for (int i = 0; i < nslots; ++i) SMB_CK(cudaStreamWaitEvent(c.slot[i], c.ev_user, 0));
    return SMB_OK;
}

// --------------------------------------------------- host-operand staging ---
// Operands in host memory are streamed through HBM in slabs along the leading
// coalesced dim: slab i uses slot i % kSlots (own stream + scratch), so the H2D
// copy of slab i+1, the kernel of slab i and the D2H copy of slab i-1 overlap.
// An operand that does not vary along the leading dim (stride 0 there) is
// uploaded once.  Device/managed operands are used in place.
static int elementwise_staged(DeviceCtx &c, int op, int dtype, const ElementwisePlan &p, const void *a, MemType ta,
                              const void *b, MemType tb, void *out, MemType to, uint64_t lane_end, cudaStream_t after) {
    const size_t es = esize(dtype);
    const int dev = c.device;
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
    Scratch sa_[kSlots], sb_[kSlots], so_[kSlots];
    DrainGuard drain; // after the scratch blocks: drained before they are released, on every way out
    for (int i = 0; i < kSlots; ++i) drain.add(c.slot[i]);
    if (int rc = order_slots_after(c, after, kSlots)) return rc;
    const void *da_inv = a, *db_inv = b;
    bool need_ev = false;
    if (on_host(ta) && !a_var) {
        if (int rc = inv_a.get(p.extent_a * es, dev)) return rc;
        note_other_op();
        SMB_CK(cudaMemcpyAsync(inv_a.p, a, p.extent_a * es, cudaMemcpyHostToDevice, c.slot[0]));
        da_inv = inv_a.p;
        need_ev = true;
    }
    if (on_host(tb) && !b_var) {
        if (int rc = inv_b.get(p.extent_b * es, dev)) return rc;
        note_other_op();
        SMB_CK(cudaMemcpyAsync(inv_b.p, b, p.extent_b * es, cudaMemcpyHostToDevice, c.slot[0]));
        db_inv = inv_b.p;
        need_ev = true;
    }
    if (need_ev) {
        SMB_CK(cudaEventRecord(c.ev, c.slot[0]));
        note_other_op();
        for (int i = 1; i < nslots; ++i) SMB_CK(cudaStreamWaitEvent(c.slot[i], c.ev, 0));
    }
    const uint64_t slab_ea = a_var ? (chunk_rows - 1) * p.sa[0] + ea1 : 0;
    const uint64_t slab_eb = b_var ? (chunk_rows - 1) * p.sb[0] + eb1 : 0;
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
                note_other_op();
                SMB_CK(cudaMemcpyAsync(sa_[sl].p, src, ((r - 1) * p.sa[0] + ea1) * es, cudaMemcpyHostToDevice, s));
                pa = (const char *)sa_[sl].p;
            } else pa = src;
        }
        if (b_var) {
            const char *src = (const char *)b + r0 * p.sb[0] * es;
            if (on_host(tb)) {
                note_other_op();
                SMB_CK(cudaMemcpyAsync(sb_[sl].p, src, ((r - 1) * p.sb[0] + eb1) * es, cudaMemcpyHostToDevice, s));
                pb = (const char *)sb_[sl].p;
            } else pb = src;
        }
        char *po = on_host(to) ? (char *)so_[sl].p : (char *)out + r0 * inner * es;
        if (int rc = elementwise_device(c, op, dtype, sub, pa, pb, po, 0, sub.n, r0 * inner, lane_end, s)) return rc;
        if (on_host(to))
            note_other_op();
            SMB_CK(cudaMemcpyAsync((char *)out + r0 * inner * es, po, sub.n * es, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < kSlots; ++i) SMB_CK(cudaStreamSynchronize(c.slot[i]));
    return SMB_OK;
}

static int scalar_staged(DeviceCtx &c, int op, int dtype, const void *a, MemType ta, const void *scalar, void *out,
                         MemType to, uint64_t n, uint64_t lane_end, cudaStream_t after) {
    const size_t es = esize(dtype);
    const int dev = c.device;
    const uint64_t chunk_bytes = (uint64_t)std::max<int64_t>(g_opt_chunk_bytes.load(), 1 << 16);
    const uint64_t chunk = std::min<uint64_t>(n, std::max<uint64_t>(1, chunk_bytes / es));
    const uint64_t nchunks = (n + chunk - 1) / chunk;
    const int nslots = (int)std::min<uint64_t>(kSlots, nchunks);
    Scratch sa_[kSlots], so_[kSlots];
    DrainGuard drain;
    for (int i = 0; i < kSlots; ++i) drain.add(c.slot[i]);
    if (int rc = order_slots_after(c, after, kSlots)) return rc;
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
            note_other_op();
            SMB_CK(cudaMemcpyAsync(sa_[sl].p, pa, cnt * es, cudaMemcpyHostToDevice, s));
            pa = (const char *)sa_[sl].p;
        }
        char *po = on_host(to) ? (char *)so_[sl].p : (char *)out + i0 * es;
        if (int rc = scalar_device(c, op, dtype, pa, scalar, po, cnt, i0, lane_end, s)) return rc;
        note_other_op();
        if (on_host(to)) SMB_CK(cudaMemcpyAsync((char *)out + i0 * es, po, cnt * es, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < nslots; ++i) SMB_CK(cudaStreamSynchronize(c.slot[i]));
    return SMB_OK;
}

static int check_args(int op, int dtype) {
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d (float, double, int32 only)", dtype);
    if (op >= kUserOpBase) {
        smb_user_op u;
        if (!user_op_lookup(op, dtype, &u)) return fail(SMB_ERR_INVALID, "op %d has no device launchers registered for dtype %d (smb_register_op)", op, dtype);
        return SMB_OK;
    }
    if (op < SMB_OP_ADD || op > SMB_OP_POW) return fail(SMB_ERR_INVALID, "unknown op %d", op);
    return SMB_OK;
}

// ------------------------------------------------- one operator, G devices ---
// (the arithmetic is in smb_shard.h.)  One host thread walks the device set: on
// device g it makes the operands of flat range g available on that device's private stream --
// an in-place operand by prefetching the range's pages there (skipped when the block's placement
// record says they already are), a replicated operand by one copy into pooled scratch of that device
// -- and launches the same kernels a single GPU would, restricted to the range.  Nothing is
// exchanged between devices.  The caller either waits for all streams (the reference's synchronous
// contract) or, in async mode, leaves.
static std::atomic<int64_t> g_opt_shard_min_bytes{32ll << 20};   // results below this stay on one device
static std::atomic<int64_t> g_opt_replicate_max_bytes{64ll << 20};

static std::atomic<int64_t> g_opt_replica_mode{0}; // shared operands: 0 read-mostly duplicates kept by the driver, 1 a private copy per call

struct ShardOperand {
    const void *base = nullptr;  // the operand as the caller passed it
    OperandShards plan;
    bool need_prefetch = false;  // the block's recorded placement is not this partition: prefetch each device's range
    bool duplicate = false;      // shared operand kept as read-mostly duplicates (no private copy)
    bool need_advise = false;    // ... and the advice has not been given yet
    uint64_t hull_lo = 0, hull_hi = 0; // elements any device reads (the advised range)
};
static uint64_t placement_sharded(const std::vector<int> &devs, const void *base, const OperandShards &o, uint64_t tag = 0) {
    uint64_t h = 0xA5ull ^ tag;
    for (size_t i = 0; i < devs.size(); ++i) {
        h = mix64(h, (uint64_t)devs[i]);
        h = mix64(h, o.r[i].lo);
        h = mix64(h, o.r[i].hi);
    }
    h = mix64(h, (uint64_t)(uintptr_t)base & 0x1fffffull); // same block, different view offset = different pages
    return h | 0x8000000000000000ull;
}
// Decide how every operand reaches the devices; false: some operand is both shared between devices
// and too large to copy per call -- the caller runs the operator on one device instead.
//   in place        ranges disjoint and at least a page each: device g's range is prefetched to it once;
//   shared / small  SMB_OPT_REPLICA_MODE 0 (default): the operand is advised READ-MOSTLY and each device's range is
//                   prefetched to it once -- the driver then keeps a read-only duplicate on every device that reads
//                   it and invalidates them on ANY write, raw host writes through SMArray::data included, which is
//                   what a cache of private copies could not promise; the kernels read the original pointer.
//                   Mode 1: a private copy per call in pooled device scratch (kept for comparison).
static bool shard_operands(const std::vector<int> &devs, const ShardSplit &split, const uint64_t *shape, int ndim,
                           const void *const *bases, const uint64_t *const *strides, int nops, size_t es, ShardOperand *ops) {
    const uint64_t rmax = (uint64_t)std::max<int64_t>(0, g_opt_replicate_max_bytes.load());
    const bool dupmode = g_opt_replica_mode.load() == 0;
    for (int o = 0; o < nops; ++o) {
        ops[o].base = bases[o];
        if (!bases[o]) continue; // a constant
        ops[o].plan = plan_operand(shape, strides[o], ndim, split, es, rmax); // (a large shared operand -- a big transpose -- keeps the operator on one device, where k_tile applies)
        if (ops[o].plan.mode == SHARD_REFUSE) return false;
    }
    for (int o = 0; o < nops; ++o) {
        if (!bases[o]) continue;
        const bool shared = ops[o].plan.mode != SHARD_IN_PLACE;
        if (shared && !dupmode) continue; // private copies: nothing to record
        ops[o].duplicate = shared;
        ops[o].hull_lo = ~0ull;
        for (int g = 0; g < split.g; ++g) {
            if (ops[o].plan.r[g].hi == ops[o].plan.r[g].lo) continue;
            ops[o].hull_lo = std::min(ops[o].hull_lo, ops[o].plan.r[g].lo);
            ops[o].hull_hi = std::max(ops[o].hull_hi, ops[o].plan.r[g].hi);
        }
        Block blk;
        bool matched = false, was_rm = false;
        const bool pooled = Pool::instance().take_placement(bases[o], placement_sharded(devs, bases[o], ops[o].plan, shared ? 0x0D0Dull : 0), &blk,
                                                            &matched, shared ? 1 : -1, &was_rm);
        ops[o].need_prefetch = !(pooled && matched);
        ops[o].need_advise = shared && !(pooled && was_rm);
    }
    return true;
}
// Operand `op` for device index g (the current device), on stream s: returns the base pointer the
// kernels of that device use (the caller's pointer, or a rebased private copy in replica mode 1).
static int shard_operand_on_device(ShardOperand &op, int g, int dev, size_t es, cudaStream_t s, Scratch &scratch,
                                   const void **use) {
    *use = op.base;
    if (!op.base) return SMB_OK;
    const ElemRange r = op.plan.r[g];
    if (r.hi == r.lo) return SMB_OK;
    const char *src = (const char *)op.base + r.lo * es;
    const size_t bytes = (r.hi - r.lo) * es;
    if (op.plan.mode == SHARD_IN_PLACE || op.duplicate) {
        if (op.need_advise) { // once per operand: before the first device's prefetch
            if (cudaMemAdvise((const char *)op.base + op.hull_lo * es, (op.hull_hi - op.hull_lo) * es, cudaMemAdviseSetReadMostly, dev) != cudaSuccess)
                cudaGetLastError();
            op.need_advise = false;
        }
        if (op.need_prefetch) {
            note_other_op();
            if (cudaMemPrefetchAsync(src, bytes, dev, s) != cudaSuccess) cudaGetLastError(); // best effort
        }
        return SMB_OK;
    }
    // private copy: same 16-byte phase as the original so the vector kernels still qualify; the kernels
    // index from the operand's element 0, so the base is moved back by the range's offset
    const size_t pad = (uintptr_t)src & 15;
    if (int rc = scratch.get(bytes + 16, dev)) return rc;
    char *dst = (char *)scratch.p + pad;
    note_other_op();
    SMB_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s));
    *use = dst - r.lo * es;
    return SMB_OK;
}

// The RESULT of a sharded operator: dense, device g writes exactly [bounds[g], bounds[g + 1]); always in place.  Its
// pages are prefetched to their devices unless the block is already partitioned this way; a read-mostly mark left by an
// earlier life as a shared operand goes (with the advice).
static ShardOperand shard_result(const std::vector<int> &devs, const ShardSplit &split, void *out) {
    ShardOperand res;
    res.base = out;
    res.plan.mode = SHARD_IN_PLACE;
    for (int g = 0; g < split.g; ++g) res.plan.r[g] = ElemRange{split.bounds[g], split.bounds[g + 1]};
    Block blk;
    bool matched = false, was_rm = false;
    res.need_prefetch = !(Pool::instance().take_placement(out, placement_sharded(devs, out, res.plan), &blk, &matched, 0, &was_rm) && matched);
    if (was_rm) drop_read_mostly(blk);
    return res;
}

// Runs `launch(ctx, g, lo, count, operand bases..., stream)` for every non-empty range of the split.
template<typename Launch>
static int run_sharded(const std::vector<int> &devs, const ShardSplit &split, ShardOperand *ops, int nops,
                       const ShardOperand &result, size_t es, bool async, Launch &&launch) {
    const int G = (int)devs.size();
    std::vector<Scratch> scratch((size_t)G * (size_t)std::max(nops, 1));
    DeviceScope scope;
    // "Pure": every operand and the result already sit in exactly this partition (their records matched, nothing is
    // prefetched or advised by this call) -- in place, or as read-mostly duplicates nobody has written since.  Each device
    // then reads and writes only what its own stream produced or what no device writes, so an operator with the same
    // partition as the one before it needs no cross-device ordering.
    bool pure = !(result.base && result.need_prefetch);
    for (int o = 0; o < nops; ++o) {
        if (!ops[o].base) continue;
        if (ops[o].need_prefetch || ops[o].need_advise) pure = false;
        if (ops[o].plan.mode != SHARD_IN_PLACE && !ops[o].duplicate) pure = false; // private copies are made by this call
    }
    // read-mostly advice: once per shared operand, before any device prefetches it
    for (int o = 0; o < nops; ++o) {
        if (!ops[o].base || !ops[o].need_advise) continue;
        if (cudaMemAdvise((const char *)ops[o].base + ops[o].hull_lo * es, (ops[o].hull_hi - ops[o].hull_lo) * es, cudaMemAdviseSetReadMostly, devs[0]) != cudaSuccess)
            cudaGetLastError();
        ops[o].need_advise = false;
    }
    if (async) {
        uint64_t sig = 0;
        if (pure) { sig = 0x51ull; for (int g = 0; g <= G; ++g) sig = mix64(sig, split.bounds[g]); for (int d : devs) sig = mix64(sig, (uint64_t)d); sig |= 1; }
        if (int rc = async_order(devs.data(), G, sig)) return rc;
    }
    int rc = SMB_OK;
    int launched = 0;
    if (!async && g_opt_launcher_threads.load(std::memory_order_relaxed) != 0) {
        // one launcher thread per device: prepare + launch + wait for the stream there, all devices at once
        struct Slot { int rc = SMB_OK; std::string err; const char *kernel = nullptr; };
        std::vector<Slot> slots((size_t)G);
        std::atomic<int> remaining{0};
        LaunchWorker *workers[kMaxShards];
        bool have_all = true;
        for (int g = 0; g < G; ++g) { workers[g] = g_workers[devs[g]]; if (!workers[g]) have_all = false; }
        if (have_all) {
            for (int g = 0; g < G; ++g) if (split.bounds[g + 1] > split.bounds[g]) remaining.fetch_add(1, std::memory_order_relaxed);
            for (int g = 0; g < G; ++g) {
                const uint64_t lo = split.bounds[g], cnt = split.bounds[g + 1] - lo;
                if (cnt == 0) continue;
                worker_post(workers[g], [&, g, lo, cnt] {
                    Slot &sl = slots[(size_t)g];
                    DeviceCtx *c = nullptr;
                    int r = ctx_of(devs[g], &c); // (the launcher already sits on its device; the context exists since smb_set_devices)
                    const void *use[SMB_CHAIN_MAX + 2];
                    for (int o = 0; o < nops && r == SMB_OK; ++o)
                        r = shard_operand_on_device(ops[o], g, devs[g], es, c->main, scratch[(size_t)g * nops + o], &use[o]);
                    if (r == SMB_OK && result.base && result.need_prefetch) {
                        const ElemRange rr = result.plan.r[g];
                        note_other_op();
                        if (cudaMemPrefetchAsync((const char *)result.base + rr.lo * es, (rr.hi - rr.lo) * es, devs[g], c->main) != cudaSuccess) cudaGetLastError();
                    }
                    if (r == SMB_OK) r = launch(*c, g, lo, cnt, use, c->main);
                    if (c) {
                        const cudaError_t e = cudaStreamSynchronize(c->main);
                        if (e != cudaSuccess && r == SMB_OK) { cudaGetLastError(); r = fail(SMB_ERR_CUDA, "device %d: %s", devs[g], cudaGetErrorString(e)); }
                    }
                    sl.rc = r;
                    if (r != SMB_OK) sl.err = g_err;       // the launcher's thread-local message
                    sl.kernel = g_last_kernel;
                    remaining.fetch_sub(1, std::memory_order_release);
                });
            }
            while (remaining.load(std::memory_order_acquire) != 0) {
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            for (int g = 0; g < G; ++g) {
                if (slots[(size_t)g].kernel) g_last_kernel = slots[(size_t)g].kernel;
                if (slots[(size_t)g].rc != SMB_OK && rc == SMB_OK) { rc = slots[(size_t)g].rc; g_err = slots[(size_t)g].err; }
            }
            return rc;
        }
    }
    for (int g = 0; g < G && rc == SMB_OK; ++g) {
        const uint64_t lo = split.bounds[g], cnt = split.bounds[g + 1] - lo;
        if (cnt == 0) continue;
        DeviceCtx *c = nullptr;
        if ((rc = scope.set(devs[g])) != SMB_OK) break;
        if ((rc = ctx_of(devs[g], &c)) != SMB_OK) break;
        const void *use[SMB_CHAIN_MAX + 2];
        for (int o = 0; o < nops && rc == SMB_OK; ++o)
            rc = shard_operand_on_device(ops[o], g, devs[g], es, c->main, scratch[(size_t)g * nops + o], &use[o]);
        if (rc != SMB_OK) break;
        if (result.base && result.need_prefetch) {
            const ElemRange r = result.plan.r[g];
            note_other_op();
            if (cudaMemPrefetchAsync((const char *)result.base + r.lo * es, (r.hi - r.lo) * es, devs[g], c->main) != cudaSuccess) cudaGetLastError();
        }
        rc = launch(*c, g, lo, cnt, use, c->main);
        ++launched;
    }
    if (async && rc == SMB_OK) return async_mark(devs.data(), G);
    // synchronous contract -- and on an error, drain what was enqueued before the scratch goes back
    for (int g = 0; g < G; ++g) {
        if (!g_ctx[devs[g]].ready.load(std::memory_order_acquire)) continue;
        const cudaError_t e = cudaStreamSynchronize(g_ctx[devs[g]].main);
        if (e != cudaSuccess && rc == SMB_OK) { cudaGetLastError(); rc = fail(SMB_ERR_CUDA, "device %d: %s", devs[g], cudaGetErrorString(e)); }
    }
    (void)launched;
    return rc;
}

// Whether an operator on these pointers is spread over the device set: only MANAGED arrays are (they
// are the drop-in SMArray storage and have one address every device can use); device blocks live on
// one GPU and are computed there, host operands go through the staging pipeline.
static bool want_sharding(std::vector<int> &devs, uint64_t result_bytes, const void *stream, bool whole) {
    if (stream || !whole || g_ndevices.load(std::memory_order_relaxed) <= 1) return false;
    if ((int64_t)result_bytes < g_opt_shard_min_bytes.load()) return false;
    devs = active_devices();
    return devs.size() > 1 && devs.size() <= (size_t)kMaxShards;
}

static int elementwise_sharded(const std::vector<int> &devs, int op, int dtype, const ElementwisePlan &p, const void *a,
                               const void *b, void *out, uint64_t lane_end, bool *done) {
    const size_t es = esize(dtype);
    const int G = (int)devs.size();
    const uint64_t rows = p.ndim >= 2 ? p.shape[0] : p.n, inner = p.ndim >= 2 ? p.n / p.shape[0] : 1;
    const ShardSplit split = split_flat(p.n, rows, inner, p.ndim, G, es);
    const void *bases[2] = {a, b};
    const uint64_t *strides[2] = {p.sa, p.sb};
    ShardOperand ops[2], res;
    *done = false;
    if (!shard_operands(devs, split, p.shape, p.ndim, bases, strides, 2, es, ops)) return SMB_OK;
    res = shard_result(devs, split, out);
    *done = true;
    return run_sharded(devs, split, ops, 2, res, es, async_mode(nullptr),
                       [&](DeviceCtx &c, int, uint64_t lo, uint64_t cnt, const void *const *use, cudaStream_t s) {
                           return elementwise_device(c, op, dtype, p, use[0], use[1], (char *)out + lo * es, lo, cnt, lo, lane_end, s);
                       });
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
        // Host operands: always synchronous (the result is in host memory on return); the copies are
        // ordered after what is already enqueued on `stream` (or on the private stream in async mode).
        cudaStream_t after = stream ? (cudaStream_t)stream : (c->dirty ? c->main : nullptr);
        if (lin_begin != 0 || lin_count != p.n) {
            // partial range with host operands: stage the touched operands whole
            const int dev = c->device;
            Scratch da, db, dout;
            DrainGuard drain;
            drain.add(c->slot[0]);
            const void *pa = a, *pb = b;
            void *po = out;
            cudaStream_t s = c->slot[0];
            if (int rc = order_slots_after(*c, after, 1)) return rc;
            note_other_op();
            if (on_host(ta)) { if (int rc = da.get(p.extent_a * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(da.p, a, p.extent_a * es, cudaMemcpyHostToDevice, s)); pa = da.p; }
            note_other_op();
            if (on_host(tb)) { if (int rc = db.get(p.extent_b * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(db.p, b, p.extent_b * es, cudaMemcpyHostToDevice, s)); pb = db.p; }
            if (on_host(to)) { if (int rc = dout.get(lin_count * es, dev)) return rc; po = dout.p; }
            if (int rc = elementwise_device(*c, op, dtype, p, pa, pb, po, lin_begin, lin_count, lin_begin, lane_end, s)) return rc;
            note_other_op();
            if (on_host(to)) SMB_CK(cudaMemcpyAsync(out, po, lin_count * es, cudaMemcpyDeviceToHost, s));
            SMB_CK(cudaStreamSynchronize(s));
            return SMB_OK;
        }
        return elementwise_staged(*c, op, dtype, p, a, ta, b, tb, out, to, lane_end, after);
    }
    std::vector<int> devs;
    if (ta == MT_MANAGED && tb == MT_MANAGED && to == MT_MANAGED && want_sharding(devs, lin_count * es, stream, lin_begin == 0 && lin_count == p.n)) {
        bool done = false;
        const int rc = elementwise_sharded(devs, op, dtype, p, a, b, out, lane_end, &done);
        if (rc || done) return rc;
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (int rc = begin_call(*c, stream)) return rc;
    if (ta == MT_MANAGED) prefetch_managed(a, p.extent_a * es, c->device, s);
    if (tb == MT_MANAGED) prefetch_managed(b, p.extent_b * es, c->device, s);
    if (to == MT_MANAGED) prefetch_managed(out, lin_count * es, c->device, s, true);
    if (int rc = elementwise_device(*c, op, dtype, p, a, b, out, lin_begin, lin_count, lin_begin, lane_end, s)) return rc;
    return finish_call(*c, s, stream);
}

// SMArray::operator% (reference math/product.h).  Enqueues the reduction of a[0..n) . b[0..n) on s;
// the scalar lands in *result_dev (a device or pinned address).  The partial / ticket scratch lives
// in one pooled device block the caller keeps until the stream has drained.
template<typename T>
static int dot_enqueue(DeviceCtx &c, const T *a, const T *b, uint64_t n, Scratch &scratch, void *result_pinned, cudaStream_t s) {
    using A = typename DotAcc<T>::type;
    constexpr int UNROLL = 4;
    constexpr int EPVV = 16 / (int)sizeof(T);
    // Views hand over interior pointers (SMArray.h:208 passes `data` straight through; the reference reads them
    // with loadu, product.h:26-71).  Same 16-byte phase: peel a scalar head up to the first common vector
    // boundary, like launch_stream; different phases: the element-wise (coalesced scalar load) variant.
    const uintptr_t ma = (uintptr_t)a % 16, mb = (uintptr_t)b % 16;
    const bool vec = ma == mb && ma % sizeof(T) == 0;
    const uint64_t head = vec && ma ? std::min<uint64_t>(n, (16 - ma) / sizeof(T)) : 0;
    const uint64_t nvec = vec ? (n - head) / EPVV : n;
    // many waves of short-lived CTAs (8 grid-stride iterations each): the hardware scheduler evens out the SMs,
    // which a resident grid with a static split cannot (the slowest SM would set the time)
    const unsigned grid = grid_for(nvec ? nvec : 1, (uint64_t)kThreads * UNROLL * 8, c.sm_count, 0);
    const size_t bytes = 16 + sizeof(A) * ((size_t)grid + 1);
    if (int rc = scratch.get(bytes, c.device)) return rc;
    unsigned int *ticket = (unsigned int *)scratch.p;
    A *res = (A *)((char *)scratch.p + 8);
    A *partials = (A *)((char *)scratch.p + 16);
    note_other_op();
    SMB_CK(cudaMemsetAsync(scratch.p, 0, 16, s));
    if (vec) k_dot<T, UNROLL, EPVV><<<grid, kThreads, 0, s>>>(a, b, n, head, partials, ticket, res);
    else k_dot<T, UNROLL, 1><<<grid, kThreads, 0, s>>>(a, b, n, 0, partials, ticket, res);
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    g_last_kernel = vec ? "k_dot" : "k_dot<unaligned>";
    SMB_CK(cudaGetLastError());
    note_other_op();
    SMB_CK(cudaMemcpyAsync(result_pinned, res, sizeof(A), cudaMemcpyDeviceToHost, s));
    return SMB_OK;
}
static int dot_enqueue_dtype(DeviceCtx &c, int dtype, const void *a, const void *b, uint64_t n, Scratch &scratch, void *result_pinned,
                             cudaStream_t s) {
    switch (dtype) {
        case SMB_F32: return dot_enqueue<float>(c, (const float *)a, (const float *)b, n, scratch, result_pinned, s);
        case SMB_F64: return dot_enqueue<double>(c, (const double *)a, (const double *)b, n, scratch, result_pinned, s);
        default: return dot_enqueue<int32_t>(c, (const int32_t *)a, (const int32_t *)b, n, scratch, result_pinned, s);
    }
}
// Adds the per-device partial results in device order, in T (int32 wraps like the reference's lanes).
static void dot_combine(int dtype, const void *partials, int count, size_t slot_bytes, void *result) {
    const char *p = (const char *)partials;
    if (dtype == SMB_F32) { float s = 0; for (int i = 0; i < count; ++i) s += *(const float *)(p + i * slot_bytes); *(float *)result = s; }
    else if (dtype == SMB_F64) { double s = 0; for (int i = 0; i < count; ++i) s += *(const double *)(p + i * slot_bytes); *(double *)result = s; }
    else { uint32_t s = 0; for (int i = 0; i < count; ++i) s += *(const uint32_t *)(p + i * slot_bytes); *(uint32_t *)result = s; }
}

} // namespace smb

using namespace smb;

// =============================================================== C ABI =====
// ---- op-chain fusion ---------------------------------
template<typename T>
static int chain_launch(DeviceCtx &c, const ChainPlan &p, const smb_chain_step *steps, const void *const *data,
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
    const int powvar = (int)g_opt_chain_pow_variant.load();
    // sm::pow(a (op) b, y) on dense same-shape arrays -- THE fused pow (SMB_OPT_CHAIN_POW_VARIANT = 4, the default): the
    // pow kernel itself with the operator applied to the two loaded operands (k_stream<PowF32FnPre>), bit-identical to
    // the operator followed by the pow kernel and at that kernel's speed; the general chain kernel below was
    // instruction-bound at 4.4-4.7 TB/s whatever its shape (profiles/r2_chain_pow_sweep.md).
    if constexpr (std::is_same<T, float>::value) {
        if (powvar == 4 && p.nleaf == 3 && p.ndim == 1 && data[0] && data[1] && !data[2] && steps[2].op == SMB_OP_POW &&
            p.stride[0][0] == 1 && p.stride[1][0] == 1 && steps[1].op != SMB_OP_POW && pow_f32_fast_ok(classify_exp(steps[2].value.f32))) {
            // (any length and alignment: launch_stream peels heads and tails exactly as it does for sm::pow itself)
            const float y = steps[2].value.f32;
            const PowExpF32 pe = classify_exp(y);
            const int pre = steps[1].op == SMB_OP_ADD ? PRE_ADD : steps[1].op == SMB_OP_MUL ? PRE_MUL
                            : steps[1].op == SMB_OP_SUB ? (steps[1].swap ? PRE_RSUB : PRE_SUB) : (steps[1].swap ? PRE_RDIV : PRE_DIV);
            const float *pa = (const float *)data[0] + lin_begin, *pb = (const float *)data[1] + lin_begin;
            const bool lt1 = pow_f32_y_lt_1(pe);
            const int tier = pow_f32_tier(pe), sign = pow_f32_sign_mode(pe);
            g_last_kernel = "k_stream<pow,fused-pre>";
#define SMB_POWPRE(S, G, L) launch_stream<float, PowF32FnPre<S, G, L>, true>(c, pa, pb, (float *)out, lin_count, lin_begin, PowF32FnPre<S, G, L>::make(y, 0, pre), s)
#define SMB_POWPRE_BY_SIGN(S) (sign == POW_SIGN_REJECT ? SMB_POWPRE(S, POW_SIGN_REJECT, false) : sign == POW_SIGN_EVEN ? SMB_POWPRE(S, POW_SIGN_EVEN, false) : SMB_POWPRE(S, POW_SIGN_ODD, false))
            int rc;
            if (lt1) rc = SMB_POWPRE(POW_TIER_SMALL, POW_SIGN_REJECT, true);
            else if (tier == POW_TIER_SMALL) rc = SMB_POWPRE_BY_SIGN(POW_TIER_SMALL);
            else if (tier == POW_TIER_MEDIUM) rc = SMB_POWPRE_BY_SIGN(POW_TIER_MEDIUM);
            else rc = SMB_POWPRE_BY_SIGN(POW_TIER_LARGE);
#undef SMB_POWPRE_BY_SIGN
#undef SMB_POWPRE
            if (rc == SMB_OK) g_last_kernel = "k_stream<pow,fused-pre>";
            return rc;
        }
    }
    t.tiles_per_cta = powfast ? (powvar == 2 || powvar == 3 ? 16 : 32) : 1; // amortise the 24 KB table copy, stay many waves deep
    // compiled-in chain capacity / vectors per thread: short chains keep more loads in flight
#define SMB_CHAIN_LAUNCH(E, W, NS, U, PF, ND)                                                                     \
    k_chain<T, E, W, NS, U, PF, ND><<<grid_for(items, (uint64_t)kThreads * U * t.tiles_per_cta, c.sm_count, 0), kThreads, 0, s>>>(out, t)
#define SMB_CHAIN_LAUNCH_PF(E, W, U, ND, PRE)                                                                     \
    k_chain<T, E, W, 3, U, true, ND, PRE><<<grid_for(items, (uint64_t)kThreads * U * t.tiles_per_cta, c.sm_count, 0), kThreads, 0, s>>>(out, t)
#define SMB_CHAIN_BY_LEN(E, W, PF, ND)                                                            \
    do {                                                                                          \
        if (p.nleaf <= 3 && PF) { /* sm::pow(a (op) b, e): vectors per thread x register prefetch, SMB_OPT_CHAIN_POW_VARIANT */ \
            if constexpr (PF) {                                                                   \
                switch (powvar) {                                                                 \
                    case 0: SMB_CHAIN_LAUNCH_PF(E, W, 1, ND, false); break;                       \
                    case 2: SMB_CHAIN_LAUNCH_PF(E, W, 2, ND, false); break;                       \
                    case 3: SMB_CHAIN_LAUNCH_PF(E, W, 2, ND, true); break;                        \
                    case 1: SMB_CHAIN_LAUNCH_PF(E, W, 1, ND, true); break;                        \
                    default: SMB_CHAIN_LAUNCH_PF(E, W, 1, ND, false); break; /* what does not fit the fused pow kernel */ \
                }                                                                                 \
            }                                                                                     \
        }                                                                                         \
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
#undef SMB_CHAIN_LAUNCH_PF
#undef SMB_CHAIN_LAUNCH
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    SMB_CK(cudaGetLastError());
    return SMB_OK;
}

static int chain_dispatch(DeviceCtx &c, int dtype, const ChainPlan &p, const smb_chain_step *steps, const void *const *data,
                          uint64_t lin_begin, uint64_t lin_count, uint64_t lane_end, void *po, cudaStream_t s) {
    switch (dtype) {
        case SMB_F32: return chain_launch<float>(c, p, steps, data, lin_begin, lin_count, lane_end, (float *)po, s);
        case SMB_F64: return chain_launch<double>(c, p, steps, data, lin_begin, lin_count, lane_end, (double *)po, s);
        default: return chain_launch<int32_t>(c, p, steps, data, lin_begin, lin_count, lane_end, (int32_t *)po, s);
    }
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
    const size_t es = esize(dtype);
    const int dev = c->device;
    // int32 pow has two semantics in the reference (AVX2 lanes wrap, the scalar tail goes through double),
    // applied by POSITION in the array sm::pow sees: array_scalar_op over the DENSE intermediate,
    // lane_end = its size rounded down to 8 (calculate.h:139-140).  The fused kernel derives the position
    // from the result's flat index, which is that array's only when the accumulator already has the
    // result's shape at the pow step.  When leaves AFTER the pow step broadcast the result further
    // (sm::pow(lazy(a{1,N}), e) + B{M,N}), the prefix up to the pow is evaluated first, as its own chain
    // of its own (smaller) shape, and joins the rest as a leaf -- bit-identical to the unfused sequence.
    if (dtype == SMB_I32 && whole) {
        for (int sidx = nsteps - 1; sidx >= 1; --sidx) {
            if (steps[sidx].op != SMB_OP_POW) continue;
            uint64_t pshape[SMB_MAX_NDIM], pn = 1, rn = 1;
            bool any_array = false;
            for (int k = 0; k < ndim; ++k) {
                bool varies = false;
                for (int i = 0; i < sidx; ++i) if (steps[i].data) { any_array = true; if (steps[i].stride[k] != 0) varies = true; }
                pshape[k] = varies || shape[k] == 1 ? shape[k] : 1;
                pn *= pshape[k];
                rn *= shape[k];
            }
            if (!any_array || pn == rn) continue;
            if (stream) return fail(SMB_ERR_INVALID, "smb_chain: an int32 pow step on a broadcast intermediate needs the synchronous form");
            // prefix [0, sidx]: strides of its leaves restricted to the dims it spans
            smb_chain_step prefix[SMB_CHAIN_MAX], rest[SMB_CHAIN_MAX];
            for (int i = 0; i <= sidx; ++i) {
                prefix[i] = steps[i];
                for (int k = 0; k < ndim; ++k) if (pshape[k] == 1) prefix[i].stride[k] = 0;
            }
            Scratch mid;
            if (int rc = mid.get(pn * es, dev)) return rc;
            if (int rc = chain_entry(dtype, prefix, sidx + 1, pshape, ndim, 0, 0, true, mid.p, nullptr)) return rc;
            if (g_opt_async.load()) SMB_CK(cudaStreamSynchronize(c->main)); // `mid` is released when this call returns
            rest[0] = smb_chain_step{};
            rest[0].data = mid.p;
            uint64_t acc = 1;
            for (int k = ndim - 1; k >= 0; --k) { rest[0].stride[k] = pshape[k] == 1 ? 0 : acc; acc *= pshape[k]; }
            int nrest = 1;
            for (int i = sidx + 1; i < nsteps; ++i) rest[nrest++] = steps[i];
            const int rc = chain_entry(dtype, rest, nrest, shape, ndim, 0, 0, true, out, nullptr);
            if (rc == SMB_OK && g_opt_async.load()) SMB_CK(cudaStreamSynchronize(c->main));
            return rc;
        }
    }
    const ChainPlan p = make_chain_plan(strides, nsteps, shape, ndim);
    if (whole) { lin_begin = 0; lin_count = p.n; }
    if (lin_begin > p.n || lin_count > p.n - lin_begin) return fail(SMB_ERR_INVALID, "flat range outside the result");
    if (lin_count == 0) return SMB_OK;
    if (!out) return fail(SMB_ERR_INVALID, "null result pointer");
    const uint64_t lane_end = dtype == SMB_I32 ? scalar_lane_end(dtype, p.n) : 0; // array_scalar_op on the dense intermediate
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    // classify the operands once
    MemType mt[SMB_CHAIN_MAX], to = mem_type(out);
    bool all_managed = to == MT_MANAGED, any_host = on_host(to);
    for (int i = 0; i < nsteps; ++i) {
        if (!steps[i].data) continue;
        mt[i] = mem_type(steps[i].data);
        if (mt[i] != MT_MANAGED) all_managed = false;
        if (on_host(mt[i])) any_host = true;
    }
    if (any_host && stream) return fail(SMB_ERR_INVALID, "smb_chain: host operands / results need the synchronous form (stream == NULL)");
    std::vector<int> devs;
    if (all_managed && want_sharding(devs, lin_count * es, stream, lin_begin == 0 && lin_count == p.n)) {
        const int G = (int)devs.size();
        const uint64_t rows = p.ndim >= 2 ? p.shape[0] : p.n, inner = p.ndim >= 2 ? p.n / p.shape[0] : 1;
        const ShardSplit split = split_flat(p.n, rows, inner, p.ndim, G, es);
        const void *bases[SMB_CHAIN_MAX];
        const uint64_t *cstr[SMB_CHAIN_MAX];
        for (int i = 0; i < nsteps; ++i) { bases[i] = steps[i].data; cstr[i] = p.stride[i]; }
        ShardOperand ops[SMB_CHAIN_MAX], res;
        if (shard_operands(devs, split, p.shape, p.ndim, bases, cstr, nsteps, es, ops)) {
            res = shard_result(devs, split, out);
            return run_sharded(devs, split, ops, nsteps, res, es, async_mode(nullptr),
                               [&](DeviceCtx &cg, int, uint64_t lo, uint64_t cnt, const void *const *use, cudaStream_t sg) {
                                   return chain_dispatch(cg, dtype, p, steps, use, lo, cnt, lane_end, (char *)out + lo * es, sg);
                               });
        }
    }
    if (int rc = begin_call(*c, stream)) return rc;
    // host leaves / result are staged whole through pooled scratch (no overlap: the fused path is
    // meant for resident arrays); managed blocks are prefetched like everywhere else
    Scratch scratch[SMB_CHAIN_MAX], dout;
    DrainGuard drain; // after the scratch blocks: an early return drains the stream before they are released
    if (any_host) drain.add(s);
    const void *data[SMB_CHAIN_MAX];
    for (int i = 0; i < nsteps; ++i) {
        data[i] = steps[i].data;
        if (!data[i]) continue;
        if (on_host(mt[i])) {
            if (int rc = scratch[i].get(p.extent[i] * es, dev)) return rc;
            note_other_op();
            SMB_CK(cudaMemcpyAsync(scratch[i].p, data[i], p.extent[i] * es, cudaMemcpyDefault, s));
            data[i] = scratch[i].p;
        } else if (mt[i] == MT_MANAGED) prefetch_managed(data[i], p.extent[i] * es, dev, s);
    }
    void *po = out;
    if (on_host(to)) {
        if (int rc = dout.get(lin_count * es, dev)) return rc;
        po = dout.p;
    } else if (to == MT_MANAGED) prefetch_managed(out, lin_count * es, dev, s, true);
    if (int rc = chain_dispatch(*c, dtype, p, steps, data, lin_begin, lin_count, lane_end, po, s)) return rc;
    note_other_op();
    if (on_host(to)) SMB_CK(cudaMemcpyAsync(out, po, lin_count * es, cudaMemcpyDefault, s));
    if (any_host) { SMB_CK(cudaStreamSynchronize(s)); return SMB_OK; } // scratch in use: synchronous whatever the mode
    return finish_call(*c, s, stream);
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
    if (on_host(ta) || on_host(to))
        return scalar_staged(*c, op, dtype, a, ta, scalar, out, to, n, lane_end, stream ? (cudaStream_t)stream : (c->dirty ? c->main : nullptr));
    std::vector<int> devs;
    if (ta == MT_MANAGED && to == MT_MANAGED && want_sharding(devs, n * es, stream, true)) {
        const int G = (int)devs.size();
        const ShardSplit split = split_flat(n, n, 1, 1, G, es);
        const uint64_t shape1[1] = {n}, unit[1] = {1};
        const void *bases[1] = {a};
        const uint64_t *strides[1] = {unit};
        ShardOperand ops[1];
        if (shard_operands(devs, split, shape1, 1, bases, strides, 1, es, ops)) {
            const ShardOperand res = shard_result(devs, split, out);
            return run_sharded(devs, split, ops, 1, res, es, async_mode(nullptr),
                               [&](DeviceCtx &cg, int, uint64_t lo, uint64_t cnt, const void *const *use, cudaStream_t s) {
                                   return scalar_device(cg, op, dtype, (const char *)use[0] + lo * es, scalar, (char *)out + lo * es, cnt, lo, lane_end, s);
                               });
        }
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (int rc = begin_call(*c, stream)) return rc;
    if (ta == MT_MANAGED) prefetch_managed(a, n * es, c->device, s);
    if (to == MT_MANAGED) prefetch_managed(out, n * es, c->device, s, true);
    if (int rc = scalar_device(*c, op, dtype, a, scalar, out, n, 0, lane_end, s)) return rc;
    return finish_call(*c, s, stream);
}

int smb_dot(int dtype, const void *a, const void *b, uint64_t n, void *result, void *stream) {
    if (dtype < SMB_F32 || dtype > SMB_I32) return fail(SMB_ERR_INVALID, "unknown dtype %d", dtype);
    if (!result) return fail(SMB_ERR_INVALID, "null result pointer");
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    const size_t es = esize(dtype);
    if (n == 0) { memset(result, 0, es); return SMB_OK; }
    if (!a || !b) return fail(SMB_ERR_INVALID, "null operand pointer");
    const MemType ta = mem_type(a), tb = mem_type(b);
    constexpr size_t kSlot = 16; // one partial result per device, in pinned memory
    Scratch pinned;
    if (int rc = pinned.get(kSlot * kMaxShards, -1, SMB_MEM_PINNED)) return rc;
    std::vector<int> devs;
    if (ta == MT_MANAGED && tb == MT_MANAGED && want_sharding(devs, 2 * n * es, stream, true)) {
        // Sharded reduction: device g reduces its flat range, the host adds the G partials in device
        // order (the one place a combine step sits on a path; across PROCESSES
        // bench.py does the same with one NCCL all-reduce of a scalar).
        const int G = (int)devs.size();
        const ShardSplit split = split_flat(n, n, 1, 1, G, es);
        const uint64_t shape1[1] = {n}, unit[1] = {1};
        const void *bases[2] = {a, b};
        const uint64_t *strides[2] = {unit, unit};
        ShardOperand ops[2], none;
        if (shard_operands(devs, split, shape1, 1, bases, strides, 2, es, ops)) {
            std::vector<Scratch> partial_scratch((size_t)G);
            memset(pinned.p, 0, kSlot * (size_t)G); // a device without elements contributes zero; slot g belongs to device g (the launchers run side by side)
            // a scalar result: always synchronous (async = false), every device's stream is drained before the sum
            const int rc = run_sharded(devs, split, ops, 2, none, es, false,
                                       [&](DeviceCtx &cg, int g, uint64_t lo, uint64_t cnt, const void *const *use, cudaStream_t s) {
                                           return dot_enqueue_dtype(cg, dtype, (const char *)use[0] + lo * es, (const char *)use[1] + lo * es, cnt,
                                                                    partial_scratch[(size_t)g], (char *)pinned.p + kSlot * (size_t)g, s);
                                       });
            if (rc) return rc;
            dot_combine(dtype, pinned.p, G, kSlot, result);
            return SMB_OK;
        }
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    const int dev = c->device;
    if (int rc = begin_call(*c, stream)) return rc;
    Scratch da, db, work;
    DrainGuard drain;
    drain.add(s);
    note_other_op();
    if (on_host(ta)) { if (int rc = da.get(n * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(da.p, a, n * es, cudaMemcpyDefault, s)); a = da.p; }
    else if (ta == MT_MANAGED) prefetch_managed(a, n * es, dev, s);
    note_other_op();
    if (on_host(tb)) { if (int rc = db.get(n * es, dev)) return rc; SMB_CK(cudaMemcpyAsync(db.p, b, n * es, cudaMemcpyDefault, s)); b = db.p; }
    else if (tb == MT_MANAGED) prefetch_managed(b, n * es, dev, s);
    if (int rc = dot_enqueue_dtype(*c, dtype, a, b, n, work, pinned.p, s)) return rc;
    SMB_CK(cudaStreamSynchronize(s)); // a scalar result: always synchronous
    dot_combine(dtype, pinned.p, 1, kSlot, result);
    return SMB_OK;
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
    // high-water mark: beyond it the largest cached blocks go back to the driver (cudaFree waits for
    // the device, so nothing in flight loses its memory)
    uint64_t st[4];
    Pool::instance().stats(st);
    const int64_t cap = g_opt_pool_max_cached.load(std::memory_order_relaxed);
    if (cap >= 0 && st[1] > (uint64_t)cap) Pool::instance().trim_to((uint64_t)cap);
    return SMB_OK;
}

int smb_host_written(const void *ptr) {
    if (ptr) Pool::instance().clear_placement(ptr);
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

static int fill_device(DeviceCtx &c, int dtype, void *out, const void *value, uint64_t n, cudaStream_t s) {
    const unsigned grid = grid_for(n, kThreads * 4, c.sm_count, 16);
    if (dtype == SMB_F64) k_fill<double><<<grid, kThreads, 0, s>>>((double *)out, n, *(const double *)value);
    else k_fill<uint32_t><<<grid, kThreads, 0, s>>>((uint32_t *)out, n, *(const uint32_t *)value);
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    SMB_CK(cudaGetLastError());
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
    const size_t es = esize(dtype);
    std::vector<int> devs;
    if (to == MT_MANAGED && want_sharding(devs, n * es, stream, true)) {
        // sm::ones / sm::zeros over the device set: the block is BORN partitioned the way the operators use it
        const int G = (int)devs.size();
        const ShardSplit split = split_flat(n, n, 1, 1, G, es);
        {
            ShardOperand ops[1];
            const ShardOperand res = shard_result(devs, split, out);
            return run_sharded(devs, split, ops, 0, res, es, async_mode(nullptr),
                               [&](DeviceCtx &cg, int, uint64_t lo, uint64_t cnt, const void *const *, cudaStream_t s) {
                                   return fill_device(cg, dtype, (char *)out + lo * es, value, cnt, s);
                               });
        }
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (int rc = begin_call(*c, stream)) return rc;
    if (to == MT_MANAGED) prefetch_managed(out, n * es, c->device, s, true);
    if (int rc = fill_device(*c, dtype, out, value, n, s)) return rc;
    return finish_call(*c, s, stream);
}

int smb_prefetch(const void *ptr, size_t bytes, int device, void *stream) {
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (mem_type(ptr) != MT_MANAGED) return SMB_OK; // nothing to migrate
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    note_other_op();
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
    if (int rc = sync_all()) return rc;
    SMB_CK(cudaDeviceSynchronize()); // the caller's own streams on the current device too
    return SMB_OK;
}
int smb_wait_pending(void) {
    if (!g_pending.load(std::memory_order_acquire)) return SMB_OK;
    return sync_all();
}

int smb_set_devices(const int *devices, int count) {
    devices_from_env_once(); // an explicit call wins over SMB_DEVICES from here on
    if (int rc = sync_all()) return rc;
    std::lock_guard<std::mutex> lk(g_set_mu);
    return set_devices_locked(devices, count);
}
int smb_get_devices(int *devices, int capacity) {
    const std::vector<int> d = active_devices();
    if (d.empty()) {
        const int cur = smb_get_device();
        if (cur < 0) return 0;
        if (devices && capacity > 0) devices[0] = cur;
        return 1;
    }
    for (int i = 0; i < (int)d.size() && i < capacity; ++i) if (devices) devices[i] = d[i];
    return (int)d.size();
}

int smb_set_option(int key, int64_t value) {
    switch (key) {
        case SMB_OPT_POW_SPECIALISE: g_opt_pow_specialise = value ? 1 : 0; return SMB_OK;
        case SMB_OPT_STAGE_CHUNK_BYTES: g_opt_chunk_bytes = value; return SMB_OK;
        case SMB_OPT_CONTIG_VARIANT: g_opt_contig_variant = value; return SMB_OK;
        case SMB_OPT_BCAST_VARIANT: g_opt_bcast_variant = value; return SMB_OK;
        case SMB_OPT_FORCE_WIDE_INDEX: g_opt_force_wide = value ? 1 : 0; return SMB_OK;
        case SMB_OPT_ASYNC:
            if (!value && g_opt_async.load()) { g_opt_async = 0; return sync_all(); } // leaving async mode: everything lands
            g_opt_async = value ? 1 : 0;
            return SMB_OK;
        case SMB_OPT_PDL: g_opt_pdl = value < 0 ? 0 : value > 2 ? 2 : value; return SMB_OK;
        case SMB_OPT_SHARD_MIN_BYTES: g_opt_shard_min_bytes = value; return SMB_OK;
        case SMB_OPT_REPLICATE_MAX_BYTES: g_opt_replicate_max_bytes = value; return SMB_OK;
        case SMB_OPT_POOL_MAX_CACHED_BYTES: g_opt_pool_max_cached = value; return SMB_OK;
        case SMB_OPT_POW_TAIL_CTAS: g_opt_pow_tail = value; return SMB_OK;
        case SMB_OPT_CHAIN_POW_VARIANT: g_opt_chain_pow_variant = value; return SMB_OK;
        case SMB_OPT_REPLICA_MODE: g_opt_replica_mode = value; return SMB_OK;
        case SMB_OPT_LAUNCHER_THREADS: g_opt_launcher_threads = value ? 1 : 0; return SMB_OK;
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
        case SMB_OPT_ASYNC: return g_opt_async;
        case SMB_OPT_PDL: return g_opt_pdl;
        case SMB_OPT_SHARD_MIN_BYTES: return g_opt_shard_min_bytes;
        case SMB_OPT_REPLICATE_MAX_BYTES: return g_opt_replicate_max_bytes;
        case SMB_OPT_POOL_MAX_CACHED_BYTES: return g_opt_pool_max_cached;
        case SMB_OPT_POW_TAIL_CTAS: return g_opt_pow_tail;
        case SMB_OPT_CHAIN_POW_VARIANT: return g_opt_chain_pow_variant;
        case SMB_OPT_REPLICA_MODE: return g_opt_replica_mode;
        case SMB_OPT_LAUNCHER_THREADS: return g_opt_launcher_threads;
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

int smb_plan_shards(const uint64_t *stride_a, const uint64_t *stride_b, const uint64_t *shape, int ndim, int elem_size,
                    int ndev, uint64_t *out_bounds, uint64_t *out_range_a, uint64_t *out_range_b, int *out_modes) {
    if (ndim < 1 || ndim > SMB_MAX_NDIM || !stride_a || !stride_b || !shape || ndev < 1 || ndev > kMaxShards ||
        (elem_size != 4 && elem_size != 8))
        return -SMB_ERR_INVALID;
    const ElementwisePlan p = make_plan(stride_a, stride_b, shape, ndim);
    const uint64_t rows = p.ndim >= 2 ? p.shape[0] : p.n, inner = p.ndim >= 2 && p.shape[0] ? p.n / p.shape[0] : 1;
    const ShardSplit split = split_flat(p.n, rows, inner, p.ndim, ndev, (size_t)elem_size);
    const uint64_t rmax = (uint64_t)std::max<int64_t>(0, g_opt_replicate_max_bytes.load());
    const OperandShards oa = plan_operand(p.shape, p.sa, p.ndim, split, (size_t)elem_size, rmax);
    const OperandShards ob = plan_operand(p.shape, p.sb, p.ndim, split, (size_t)elem_size, rmax);
    for (int g = 0; g <= ndev; ++g) if (out_bounds) out_bounds[g] = split.bounds[g];
    for (int g = 0; g < ndev; ++g) {
        if (out_range_a) { out_range_a[2 * g] = oa.r[g].lo; out_range_a[2 * g + 1] = oa.r[g].hi; }
        if (out_range_b) { out_range_b[2 * g] = ob.r[g].lo; out_range_b[2 * g + 1] = ob.r[g].hi; }
    }
    if (out_modes) { out_modes[0] = oa.mode; out_modes[1] = ob.mode; }
    return oa.mode == SHARD_REFUSE || ob.mode == SHARD_REFUSE ? 0 : 1;
}

int smb_register_op(const char *name, int dtype, const smb_user_op *launchers) {
    if (!name || !*name || !launchers || dtype < SMB_F32 || dtype > SMB_I32 || !launchers->contiguous || !launchers->scalar || !launchers->strided) {
        fail(SMB_ERR_INVALID, "smb_register_op: name, dtype and all three launchers are required");
        return -SMB_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lk(g_user_mu);
    int idx = -1;
    for (size_t i = 0; i < g_user_ops.size(); ++i) if (g_user_ops[i].name == name) idx = (int)i;
    if (idx < 0) { g_user_ops.emplace_back(); idx = (int)g_user_ops.size() - 1; g_user_ops[idx].name = name; }
    g_user_ops[idx].fn[dtype] = *launchers;
    g_user_ops[idx].have[dtype] = true;
    return kUserOpBase + idx;
}
int smb_find_op(const char *name) {
    if (!name) return -SMB_ERR_INVALID;
    std::lock_guard<std::mutex> lk(g_user_mu);
    for (size_t i = 0; i < g_user_ops.size(); ++i) if (g_user_ops[i].name == name) return kUserOpBase + (int)i;
    return -SMB_ERR_INVALID;
}

int smb_pow_audit_f32(const void *x, float y, const void *got, uint64_t n, float bound_ulp, uint64_t *count_over, float *max_ulp) {
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (!count_over || !max_ulp) return fail(SMB_ERR_INVALID, "null result pointer");
    *count_over = 0;
    *max_ulp = 0.0f;
    if (n == 0) return SMB_OK;
    if (!x || !got || on_host(mem_type(x)) || on_host(mem_type(got))) return fail(SMB_ERR_INVALID, "smb_pow_audit_f32 needs device or managed memory");
    if (int rc = sync_all()) return rc;
    Scratch acc;
    if (int rc = acc.get(16, c->device)) return rc;
    cudaStream_t s = c->main;
    DrainGuard drain;
    drain.add(s);
    note_other_op();
    SMB_CK(cudaMemsetAsync(acc.p, 0, 16, s));
    const unsigned grid = grid_for(n, kThreads * 8, c->sm_count, 16);
    k_pow_audit_f32<<<grid, kThreads, 0, s>>>((const float *)x, (const float *)got, n, classify_exp(y), bound_ulp,
                                             (unsigned long long *)acc.p, (unsigned int *)((char *)acc.p + 8));
    ++g_launches;
    note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
    SMB_CK(cudaGetLastError());
    unsigned long long host[2] = {0, 0};
    note_other_op();
    SMB_CK(cudaMemcpyAsync(host, acc.p, 16, cudaMemcpyDeviceToHost, s));
    SMB_CK(cudaStreamSynchronize(s));
    *count_over = host[0];
    const uint32_t bits = (uint32_t)host[1];
    memcpy(max_ulp, &bits, 4);
    return SMB_OK;
}

int smb_fill_uniform_f32(void *out, uint64_t first, uint64_t n, uint64_t seed, float lo, float hi, void *stream) {
    DeviceCtx *c = nullptr;
    if (int rc = current_ctx(&c)) return rc;
    if (n == 0) return SMB_OK;
    const MemType to = out ? mem_type(out) : MT_HOST;
    if (!out || on_host(to)) return fail(SMB_ERR_INVALID, "smb_fill_uniform_f32 needs device or managed memory");
    auto fill = [&](DeviceCtx &cg, float *dst, uint64_t at, uint64_t cnt, cudaStream_t s) {
        const unsigned grid = grid_for(cnt, kThreads * 4, cg.sm_count, 16);
        k_fill_uniform_f32<<<grid, kThreads, 0, s>>>(dst, first + at, cnt, seed, lo, hi);
        ++g_launches;
        note_other_op(); // a plain launch: the next stream kernel after it is launched plainly too
        SMB_CK(cudaGetLastError());
        return (int)SMB_OK;
    };
    std::vector<int> devs;
    if (to == MT_MANAGED && want_sharding(devs, n * 4, stream, true)) { // a function of the flat index: shards trivially
        const int G = (int)devs.size();
        const ShardSplit split = split_flat(n, n, 1, 1, G, 4);
        {
            ShardOperand ops[1];
            const ShardOperand res = shard_result(devs, split, out);
            return run_sharded(devs, split, ops, 0, res, 4, async_mode(nullptr),
                               [&](DeviceCtx &cg, int, uint64_t at, uint64_t cnt, const void *const *, cudaStream_t s) {
                                   return fill(cg, (float *)out + at, at, cnt, s);
                               });
        }
    }
    cudaStream_t s = stream ? (cudaStream_t)stream : c->main;
    if (int rc = begin_call(*c, stream)) return rc;
    if (to == MT_MANAGED) prefetch_managed(out, n * 4, c->device, s, true);
    if (int rc = fill(*c, (float *)out, 0, n, s)) return rc;
    return finish_call(*c, s, stream);
}

} // extern "C"
