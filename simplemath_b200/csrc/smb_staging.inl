// smb_staging.inl -- asynchronous-mode bookkeeping and the host-operand staging pipeline (part of smb_api.cu's translation
// unit, inside namespace smb).
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
    std::lock_guard<std::mutex> stage(c.stage_mu);
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
        if (on_host(to)) {
            note_other_op();
            SMB_CK(cudaMemcpyAsync((char *)out + r0 * inner * es, po, sub.n * es, cudaMemcpyDeviceToHost, s));
        }
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
    std::lock_guard<std::mutex> stage(c.stage_mu);
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

