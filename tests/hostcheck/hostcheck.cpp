// hostcheck.cpp -- TEST ARTEFACT: compiles the device math header
// (simplemath_b200/csrc/smb_math.cuh) for the host so the pow / powi bodies can
// be swept for accuracy on the CPU-only build container.  Built into
// tests/hostcheck/libsmb_hostcheck.so by __graft_entry__.build(); never linked
// into libsmb200.so -- the product has no host execution path.
#include "../../simplemath_b200/csrc/smb_math.cuh"
#include <cstdint>
using namespace smb;
extern "C" {
void hc_pow_f32(const float *x, float y, uint64_t n, float *out) {
    PowExpF32 pe = classify_exp(y);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = pow_f32(x[i], pe);
}
static const PowTabLog kLog[SMB_POW_LOG_ENTRIES] = SMB_POW_LOG_TABLE_INIT;   // small-y: {invc, -log2 invc}
static const PowTabLog kLogC[SMB_POW_LOG_ENTRIES] = SMB_POW_LOGC_TABLE_INIT; // large-y: {c, log2 c}
static const PowTabExp kExp[SMB_POW_EXP_ENTRIES] = SMB_POW_EXP_TABLE_INIT;
// The fast (table-driven, packed) core exactly as ScalarFn<OP_POW,float>::pair uses it:
// pairs of elements, pow_f32 for whatever the core declines.  *declined counts those.
} // extern "C"
template<int S, int SG, bool LT1>
static bool hc_pair(float a, float b, float y, float *r0, float *r1) {
    return pow_f32_pair_fast<S, SG, LT1>(a, b, y, pow_lane(0, pow_consts()), S != POW_TIER_LARGE ? kLog : kLogC, kExp, r0, r1);
}
template<int S, bool LT1>
static bool hc_pair_sign(int sg, float a, float b, float y, float *r0, float *r1) {
    return sg == POW_SIGN_REJECT ? hc_pair<S, POW_SIGN_REJECT, LT1>(a, b, y, r0, r1)
         : sg == POW_SIGN_EVEN   ? hc_pair<S, POW_SIGN_EVEN, LT1>(a, b, y, r0, r1)
                                 : hc_pair<S, POW_SIGN_ODD, LT1>(a, b, y, r0, r1);
}
extern "C" {
void hc_pow_f32_fast(const float *x, float y, uint64_t n, float *out, uint64_t *declined) {
    PowExpF32 pe = classify_exp(y);
    const bool fast = pow_f32_fast_ok(pe), lt1 = pow_f32_y_lt_1(pe);
    const int tier = pow_f32_tier(pe);
    const int sg = pow_f32_sign_mode(pe);
    uint64_t dec = 0;
    #pragma omp parallel for schedule(static) reduction(+:dec)
    for (int64_t i = 0; i < (int64_t)(n / 2); ++i) {
        float r0, r1;
        const float a = x[2 * i], b = x[2 * i + 1];
        const bool ok = lt1 ? hc_pair_sign<POW_TIER_SMALL, true>(sg, a, b, y, &r0, &r1)
                      : tier == POW_TIER_SMALL ? hc_pair_sign<POW_TIER_SMALL, false>(sg, a, b, y, &r0, &r1)
                      : tier == POW_TIER_MEDIUM ? hc_pair_sign<POW_TIER_MEDIUM, false>(sg, a, b, y, &r0, &r1)
                                                : hc_pair_sign<POW_TIER_LARGE, false>(sg, a, b, y, &r0, &r1);
        if (ok && fast) {
            out[2 * i] = r0; out[2 * i + 1] = r1;
        } else {
            out[2 * i] = pow_f32(x[2 * i], pe); out[2 * i + 1] = pow_f32(x[2 * i + 1], pe);
            dec += 2;
        }
    }
    if (n & 1) out[n - 1] = pow_f32(x[n - 1], pe);
    if (declined) *declined = dec;
}
// The core as op chains run it (k_chain POWFAST): sign handling and range test decided at run time
// (POW_SIGN_RUNTIME), tier MEDIUM for |y| <= 256 and LARGE beyond; pow_f32 for whatever it declines.
void hc_pow_f32_fast_runtime(const float *x, float y, uint64_t n, float *out, uint64_t *declined) {
    PowExpF32 pe = classify_exp(y);
    const bool fast = pow_f32_fast_ok(pe), large = pow_f32_tier(pe) == POW_TIER_LARGE;
    const uint32_t abs_mask = pe.y_is_int ? 0x7fffffffu : 0xffffffffu, sign_or = pe.y_is_odd ? 0x80000000u : 0u;
    uint64_t dec = 0;
    #pragma omp parallel for schedule(static) reduction(+:dec)
    for (int64_t i = 0; i < (int64_t)(n / 2); ++i) {
        float r0, r1;
        const float a = x[2 * i], b = x[2 * i + 1];
        const bool ok = large ? pow_f32_pair_fast<POW_TIER_LARGE, POW_SIGN_RUNTIME, false>(a, b, y, pow_lane(0, pow_consts()), kLogC, kExp, &r0, &r1, abs_mask, sign_or)
                              : pow_f32_pair_fast<POW_TIER_MEDIUM, POW_SIGN_RUNTIME, false>(a, b, y, pow_lane(0, pow_consts()), kLog, kExp, &r0, &r1, abs_mask, sign_or);
        if (ok && fast) { out[2 * i] = r0; out[2 * i + 1] = r1; }
        else { out[2 * i] = pow_f32(a, pe); out[2 * i + 1] = pow_f32(b, pe); dec += 2; }
    }
    if (n & 1) out[n - 1] = pow_f32(x[n - 1], pe);
    if (declined) *declined = dec;
}
void hc_pow_f32_pair(const float *x, const float *y, uint64_t n, float *out) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = DevOp<OP_POW, float>::apply(x[i], y[i]);
}
void hc_pow_f64(const double *x, double y, uint64_t n, double *out) {
    PowExpF64 pe = classify_exp(y);
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = pow_f64(x[i], pe);
}
static const PowTabLog64 kLog64[SMB_POW_LOG_ENTRIES] = SMB_POW64_LOG_TABLE_INIT;
static const PowTabExp64 kExp64[SMB_POW_EXP_ENTRIES] = SMB_POW64_EXP_TABLE_INIT;
static PowTabLog64A kLog64A[SMB_POW_LOG_ENTRIES];
static PowTabLog64B kLog64B[SMB_POW_LOG_ENTRIES];
static const bool kLog64Split = [] {
    for (int i = 0; i < SMB_POW_LOG_ENTRIES; ++i) { kLog64A[i].c = kLog64[i].c; kLog64A[i].l_hi = kLog64[i].l_hi; kLog64B[i].l_lo = kLog64[i].l_lo; }
    return true;
}();
// The f64 kernel's own path: table-driven fast core, pow_f64 for declined elements.
void hc_pow_f64_fast(const double *x, double y, uint64_t n, double *out, uint64_t *declined) {
    PowExpF64 pe = classify_exp(y);
    const bool fast = pow_f64_fast_ok(pe), odd = pe.y_is_odd != 0, small = pow_f64_small_y(pe);
    const uint64_t rej = pe.y_is_int ? 0ull : 0x8000000000000000ull;
    uint64_t dec = 0;
    #pragma omp parallel for schedule(static) reduction(+:dec)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        double r;
        const bool ok = small ? (odd ? pow_f64_fast<true, true>(x[i], y, rej, kLog64A, kLog64B, kExp64, &r) : pow_f64_fast<true, false>(x[i], y, rej, kLog64A, kLog64B, kExp64, &r))
                              : (odd ? pow_f64_fast<false, true>(x[i], y, rej, kLog64A, kLog64B, kExp64, &r) : pow_f64_fast<false, false>(x[i], y, rej, kLog64A, kLog64B, kExp64, &r));
        if (ok && fast) out[i] = r;
        else { out[i] = pow_f64(x[i], pe); ++dec; }
    }
    if (declined) *declined = dec;
}
void hc_pow_f64_pair(const double *x, const double *y, uint64_t n, double *out) {
    #pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = DevOp<OP_POW, double>::apply(x[i], y[i]);
}
void hc_powi(const int32_t *b, const int32_t *e, uint64_t n, int lane, int32_t *out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = lane ? powi_lane(b[i], e[i]) : powi_scalar(b[i], e[i]);
}
}
