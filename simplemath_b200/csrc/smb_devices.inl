// smb_devices.inl -- one operator over a device set (part of smb_api.cu's translation unit, inside namespace smb): how operands
// and the result reach the devices, the per-range launches (launcher threads / single-thread walk), the elementwise entry point,
// and the dot-product reduction.  The arithmetic of where to cut is in smb_shard.h.
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
    // A synchronous operator (the dot product always is) while asynchronous work is pending: that work may be producing
    // an operand on ANOTHER device's stream (a small array computed on one device, a different partition) -- land it first.
    if (!async && g_pending.load(std::memory_order_acquire)) if (int rc = sync_all()) return rc;
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
        for (int g = 0; g < G; ++g) { workers[g] = g_workers[devs[g]].load(std::memory_order_acquire); if (!workers[g]) have_all = false; }
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
            for (unsigned spins = 0; remaining.load(std::memory_order_acquire) != 0; ++spins) {
                if (spins > 50000) { std::this_thread::yield(); continue; } // a long kernel: stop hogging the core
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
        // ordered after what is already enqueued on `stream`; with stream == NULL, pending asynchronous work
        // (on any device of the set) is waited for first -- it may be producing one of the operands.
        cudaStream_t after = (cudaStream_t)stream;
        if (!stream && g_pending.load(std::memory_order_acquire)) if (int rc = sync_all()) return rc;
        if (lin_begin != 0 || lin_count != p.n) {
            // partial range with host operands: stage the touched operands whole
            const int dev = c->device;
            std::lock_guard<std::mutex> stage(c->stage_mu);
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

