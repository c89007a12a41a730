// smb_launch.inl -- launchers (part of smb_api.cu's translation unit, inside namespace smb): from a plan and device-accessible
// operands to kernel launches -- dense streams (with the pow variants and programmatic dependent launch), broadcast / outer /
// tile / gather / generic kernels, and the dispatch into user-registered device Ops.
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

