// smb_runtime.inl -- the runtime under the C ABI (part of smb_api.cu's single translation unit, inside namespace smb):
// error reporting, options, per-device contexts (streams, events, pow table image), the launcher thread of each device of a
// device set, the device set itself, the pooled allocator, pointer classification + managed-memory placement, and the
// host side of programmatic dependent launch.
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
    std::atomic<bool> dirty{false}; // async mode: work enqueued on `main` since the last synchronisation
    StreamTrack track;             // of `main`
    std::mutex stage_mu;           // one staged (host-operand) call at a time per device: they share the slot streams and the two ordering events
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
static std::atomic<LaunchWorker *> g_workers[kMaxDevices]; // created once per device (g_set_mu), never destroyed: they outlive static destruction
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
    LaunchWorker *w = g_workers[dev].load(std::memory_order_acquire);
    if (!w) {
        w = new LaunchWorker;
        w->dev = dev;
        w->th = std::thread(worker_main, w);
        w->th.detach();
        g_workers[dev].store(w, std::memory_order_release);
    }
    return w;
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
//     has completed and flushed): it only saves the scheduling ramp.  It triggers its own dependents
//     AFTER that wait (smb_kernels.cuh: pdl_enter): pdl_decide forgets the earlier launches' ranges
//     at a waiting launch, which is only sound if nothing later can start before the wait is over.
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

