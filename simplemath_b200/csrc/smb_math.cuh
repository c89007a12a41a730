// smb_math.cuh -- device bodies of the reference's Op structs.
//
// One functor per (Op, T): the `apply_device` specialisation each reference Op
// struct gains (AddOp add.h:5-14, SubtractOp subtract.h:5-14, MultiplyOp
// multiply.h:7-16, DivideOp division.h:8-17,67-70, PowOp pow.h:6-14 and
// math/simd/crafted_pow.h:54-103).
//
// Semantics:
//   f32/f64 + - * / : IEEE-754 round-to-nearest-even, denormals kept, no FMA
//                     contraction (explicit _rn intrinsics, and the file is
//                     compiled with -fmad=false) -> bit-exact vs Op::apply.
//   i32 + - *       : two's-complement wrap (what _mm256_{add,sub,mullo}_epi32 do).
//   i32 /           : C truncation toward zero (/0 and INT_MIN/-1 trap on x86,
//                     are undefined in the reference and return an unspecified
//                     value here; no test feeds them).
//   i32 pow         : two flavours, because the reference has two (see below).
//   f32/f64 pow     : exp2(y * log2 x), correctly range-reduced, full C99 Annex F
//                     special-case table; ULP-bounded vs std::pow.
//
// The functions are SMB_HD so tests/hostcheck can compile this same source for
// the host and sweep the pow kernels' accuracy without a GPU.  That build is a
// test artefact; the shipped library has no host execution path.
#pragma once
#include <stdint.h>
#include <math.h>
#include "smb_pow_tables.h"

#if defined(__CUDACC__)
#define SMB_HD __host__ __device__ __forceinline__
#define SMB_D __device__ __forceinline__
#else
#define SMB_HD inline
#define SMB_D inline
#endif

namespace smb {

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_POW = 4 };
enum { DT_F32 = 0, DT_F64 = 1, DT_I32 = 2 };

// ---- bit casts and exact single-rounding primitives, host and device -------
SMB_HD uint32_t f2u(float x) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(x);
#else
    uint32_t u; __builtin_memcpy(&u, &x, 4); return u;
#endif
}
SMB_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float x; __builtin_memcpy(&x, &u, 4); return x;
#endif
}
SMB_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; __builtin_memcpy(&u, &x, 8); return u;
#endif
}
SMB_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x; __builtin_memcpy(&x, &u, 8); return x;
#endif
}
SMB_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
SMB_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
SMB_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
SMB_HD float ffma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}
SMB_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SMB_HD double dsub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
SMB_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
SMB_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// ============================================================ integer pow ===
// Lane semantics: __sm256_powi_ps, math/simd/crafted_pow.h:54-103 -- what every
// element produced by an AVX2 apply_simd lane gets.  |exp| as
// _mm256_abs_epi32 + logical shifts see it (INT_MIN -> 0x80000000), wrapping
// mullo (:72,:76); 0^(exp>0) -> 0 (:81-84); exp<0 -> 0 except base 1 -> 1 and
// base -1 -> (exp odd ? -1 : 1) (:85-102).  The exponent is warp-uniform on the
// sm::pow path, so the loop does not diverge there.
SMB_HD int32_t powi_lane(int32_t base, int32_t exp) {
    if (exp < 0) {
        if (base == 1) return 1;
        if (base == -1) return (exp & 1) ? -1 : 1;
        return 0;
    }
    uint32_t e = (uint32_t)exp, cur = (uint32_t)base, res = 1u;
    while (e) {
        if (e & 1u) res *= cur;
        cur *= cur;
        e >>= 1;
    }
    return (int32_t)res; // base==0 && exp>0 already yields 0
}

// Scalar semantics: PowOp<int>::apply, math/pow.h:8-10 -- std::pow in double,
// converted back to int on return (cvttsd2si on x86-64: anything outside
// (-2^31-1, 2^31) -- including +inf from 0^negative -- becomes 0x80000000).
// Reached by the tail elements of the AVX loops (calculate.h:130-133,166-168)
// and by every element of the general strided loop (calculate.h:96).
// libm's pow is exact whenever the true result is a representable integer of
// this size, so the double detour is reproduced with integers only: exact
// magnitude by squaring, saturating once it passes 2^31.
SMB_HD int32_t powi_scalar(int32_t base, int32_t exp) {
    const int32_t indefinite = (int32_t)0x80000000u;
    if (exp < 0) {
        if (base == 0) return indefinite; // pow(+0, negative) = +inf
        if (base == 1) return 1;
        if (base == -1) return (exp & 1) ? -1 : 1;
        return 0; // |result| < 1 truncates to 0
    }
    const uint64_t cap = 0x80000001ull; // any magnitude > 2^31 behaves alike
    uint64_t cur = base < 0 ? (uint64_t)0 - (uint64_t)(int64_t)base : (uint64_t)base;
    uint64_t res = 1;
    uint32_t e = (uint32_t)exp;
    while (e) {
        if (e & 1u) { res *= cur; if (res > cap) res = cap; }
        e >>= 1;
        if (e) { cur *= cur; if (cur > cap) cur = cap; }
    }
    if (res >= 0x80000000ull) return indefinite; // -2^31 exactly is the same bits
    const bool neg = base < 0 && (exp & 1);
    return neg ? -(int32_t)res : (int32_t)res;
}

// =============================================================== float pow ===
// Exponent facts that are uniform over a launch (the exponent of sm::pow is one
// scalar, UserFunctions.h:42-48) are classified once on the host.
struct PowExpF32 {
    float y;
    int y_is_int;   // y is an integer value (includes |y| >= 2^24)
    int y_is_odd;   // y is an odd integer
    int y_class;    // 0 finite non-zero, 1 zero, 2 +inf, 3 -inf, 4 NaN
};
struct PowExpF64 {
    double y;
    int y_is_int;
    int y_is_odd;
    int y_class;
};

SMB_HD PowExpF32 classify_exp(float y) {
    PowExpF32 p;
    p.y = y;
    uint32_t u = f2u(y) & 0x7fffffffu;
    p.y_class = 0;
    if (u == 0) p.y_class = 1;
    else if (u == 0x7f800000u) p.y_class = (f2u(y) >> 31) ? 3 : 2;
    else if (u > 0x7f800000u) p.y_class = 4;
    p.y_is_int = 0;
    p.y_is_odd = 0;
    if (p.y_class == 0 || p.y_class == 1) {
        int e = (int)(u >> 23) - 127; // unbiased exponent
        if (u == 0) { p.y_is_int = 1; }
        else if (e >= 24) { p.y_is_int = 1; }            // even: ulp >= 2
        else if (e >= 0) {
            uint32_t frac_mask = (1u << (23 - e)) - 1u;
            if (e == 23) frac_mask = 0;
            if ((u & frac_mask) == 0) {
                p.y_is_int = 1;
                p.y_is_odd = (int)((u >> (23 - e)) & 1u);
            }
        }
    }
    return p;
}
SMB_HD PowExpF64 classify_exp(double y) {
    PowExpF64 p;
    p.y = y;
    uint64_t u = d2u(y) & 0x7fffffffffffffffull;
    p.y_class = 0;
    if (u == 0) p.y_class = 1;
    else if (u == 0x7ff0000000000000ull) p.y_class = (d2u(y) >> 63) ? 3 : 2;
    else if (u > 0x7ff0000000000000ull) p.y_class = 4;
    p.y_is_int = 0;
    p.y_is_odd = 0;
    if (p.y_class == 0 || p.y_class == 1) {
        int e = (int)(u >> 52) - 1023;
        if (u == 0) { p.y_is_int = 1; }
        else if (e >= 53) { p.y_is_int = 1; }
        else if (e >= 0) {
            uint64_t frac_mask = e == 52 ? 0ull : ((1ull << (52 - e)) - 1ull);
            if ((u & frac_mask) == 0) {
                p.y_is_int = 1;
                p.y_is_odd = (int)((u >> (52 - e)) & 1ull);
            }
        }
    }
    return p;
}

// ---- log2 / exp2 kernels in double, used for f32 pow -----------------------
// For a float result the double pipeline leaves < 2^-30 relative error before
// the single final rounding, i.e. the result is the correctly rounded value in
// all but ~1 in 2^6 * 2^-... near-tie cases: <= 0.5 + 2^-6 ULP.
//
// log2(x), x a positive finite double (any float converts to a NORMAL double, so
// float denormals need no special path):
//   x = 2^e * m, m in [sqrt(1/2), sqrt(2));  p = (m-1)/(m+1), |p| <= 0.1716;
//   log2(m) = (2/ln2) * atanh(p) = p * (c0 + s*(c1 + ... )), s = p*p.
// Truncating after the s^6 term leaves 2 * 0.02944^7 / 15 / (2 * 0.1716...) ->
// relative 2^-38.
SMB_HD double log2_pos(double x) {
    uint64_t ux = d2u(x);
    // choose e so that m = x * 2^-e lies in [sqrt(1/2), sqrt(2))
    int64_t hi = (int64_t)(ux >> 32);
    int64_t eadj = (hi - 0x3fe6a09e) >> 20; // floor((hi - hi(sqrt(1/2))) / 2^20)
    double m = u2d(ux - ((uint64_t)eadj << 52));
    double num = dsub(m, 1.0);
    double den = dadd(m, 1.0);
    // reciprocal: float seed + one Newton step in double (relative 2^-44)
#if defined(__CUDA_ARCH__)
    double r = (double)__frcp_rn((float)den);
#else
    double r = (double)(1.0f / (float)den);
#endif
    double err = dfma(-den, r, 1.0);
    r = dfma(r, err, r);
    double p = dmul(num, r);
    // one correction of p itself: p += r * (num - p*den)
    p = dfma(r, dfma(-p, den, num), p);
    double s = dmul(p, p);
    // (2/ln2) / (2k+1)
    const double c0 = 2.8853900817779268147;  // 2/ln2
    const double c1 = 0.96179669392597560491; // 2/(3 ln2)
    const double c2 = 0.57707801635558536295; // 2/(5 ln2)
    const double c3 = 0.41219858311113240210; // 2/(7 ln2)
    const double c4 = 0.32059889797532520164; // 2/(9 ln2)
    const double c5 = 0.26230818925253880134; // 2/(11 ln2)
    const double c6 = 0.22195308321368667805; // 2/(13 ln2)
    double q = dfma(s, c6, c5);
    q = dfma(s, q, c4);
    q = dfma(s, q, c3);
    q = dfma(s, q, c2);
    q = dfma(s, q, c1);
    q = dfma(s, q, c0);
    return dfma(p, q, (double)eadj);
}

// exp2(t) for |t| <= ~1100, returned as a double: n = rint(t), f = t - n in
// [-0.5, 0.5], 2^f by a degree-10 Taylor polynomial in f*ln2 folded into the
// coefficients (truncation (0.3466)^11/11! = 2^-42), scaled by 2^n through the
// exponent field.
SMB_HD double exp2_d(double t) {
    const double shifter = 6755399441055744.0; // 1.5 * 2^52
    double tn = dadd(t, shifter);
    int32_t n = (int32_t)(uint32_t)d2u(tn); // low word holds rint(t)
    double fn = dsub(tn, shifter);
    double f = dsub(t, fn);
    const double k1 = 6.93147180559945309417e-01;
    const double k2 = 2.40226506959100712334e-01;
    const double k3 = 5.55041086648215799532e-02;
    const double k4 = 9.61812910762847716197e-03;
    const double k5 = 1.33335581464284434234e-03;
    const double k6 = 1.54035303933816099544e-04;
    const double k7 = 1.52527338040598402800e-05;
    const double k8 = 1.32154867901443094884e-06;
    const double k9 = 1.01780860092396997275e-07;
    const double k10 = 7.05491162080112332987e-09;
    double q = dfma(f, k10, k9);
    q = dfma(f, q, k8);
    q = dfma(f, q, k7);
    q = dfma(f, q, k6);
    q = dfma(f, q, k5);
    q = dfma(f, q, k4);
    q = dfma(f, q, k3);
    q = dfma(f, q, k2);
    q = dfma(f, q, k1);
    q = dfma(f, q, 1.0);
    return u2d(d2u(q) + ((uint64_t)(int64_t)n << 52));
}

// The special-case table of C99 Annex F.10.4.4 / IEEE 754-2008 pow, shared by
// f32 and f64.  Returns true and sets *out when (x, y) is a special pair;
// otherwise x is finite and non-zero, y finite and non-zero, and the caller
// computes |x|^y and applies `negate`.
template<typename F, typename PE>
SMB_HD bool pow_special(F x, const PE &pe, F *out, bool *negate) {
    const F one = (F)1, zero = (F)0;
    const F inf = (F)INFINITY;
    *negate = false;
    if (pe.y_class == 1) { *out = one; return true; }            // pow(x, +-0) = 1, even for NaN
    if (x == one) { *out = one; return true; }                   // pow(+1, y) = 1, even for NaN
    if (x != x || pe.y_class == 4) { *out = x + pe.y; return true; } // NaN propagates
    const bool xneg = signbit(x);
    const F ax = xneg ? -x : x;
    if (pe.y_class == 2 || pe.y_class == 3) {                    // y = +-inf
        if (ax == one) { *out = one; return true; }              // pow(-1, +-inf) = 1
        const bool grow = (ax > one) == (pe.y_class == 2);
        *out = grow ? inf : zero;
        return true;
    }
    const bool yneg = pe.y < zero;
    if (ax == zero) {                                            // pow(+-0, y)
        F r = yneg ? inf : zero;                                 // (divide-by-zero flag not modelled)
        *out = (xneg && pe.y_is_odd) ? -r : r;
        return true;
    }
    if (ax == inf) {                                             // pow(+-inf, y)
        F r = yneg ? zero : inf;
        *out = (xneg && pe.y_is_odd) ? -r : r;
        return true;
    }
    if (xneg) {
        if (!pe.y_is_int) { *out = (F)NAN; return true; }        // negative finite ^ non-integer
        *negate = pe.y_is_odd != 0;
    }
    return false;
}

// f32 pow, reference-accuracy path (FP64 pipe, <= 0.5002 ULP): every special
// case, denormal inputs, per-element exponents (array ^ array) and the rare
// elements the fast core below declines.
SMB_HD float pow_f32(float x, const PowExpF32 &pe) {
    const uint32_t ux = f2u(x);
    bool negate = false;
    if (!(pe.y_class == 0 && (ux - 1u) < 0x7f7fffffu)) { // not (positive finite non-zero x, finite non-zero y)
        float sp;
        if (pow_special<float, PowExpF32>(x, pe, &sp, &negate)) return sp;
    }
    const double ax = (double)u2f(ux & 0x7fffffffu);
    double t = dmul((double)pe.y, log2_pos(ax));
    // float range: 2^128 overflows, below 2^-150 rounds to zero.  Clamp so the
    // exponent arithmetic in exp2_d stays in range; the clamped values still
    // round to inf / 0 in the final conversion.
    t = t > 200.0 ? 200.0 : t;
    t = t < -200.0 ? -200.0 : t;
    float r = (float)exp2_d(t); // one rounding, gradual underflow included
    return negate ? -r : r;
}

// ---- table-driven f32 pow core on the FP32 pipe, two elements per instruction --
// The double core runs on the half-rate FP64 pipe and made sm::pow compute-bound
// at a third of the HBM rate (2.15 TB/s measured on B200).  This core keeps the
// arithmetic on the FP32 pipe and issues it as packed fma.rn.f32x2 (SASS FFMA2,
// new with sm_100): two elements per instruction, ~25 instructions per element in
// all, which is what fits under the memory roofline.
//
//   |x| = 2^E * m, m in [1, 2); the top 7 mantissa bits select {invc, L_hi, L_lo} with
//   -log2(invc) = L_hi + L_lo and invc = k/256 an EIGHT-bit reciprocal of the entry's centre:
//   r = fma(m, invc, -1) is then exact (|r| <= 2^-7, M*k - 2^31 fits 24 bits) -- no division,
//   no reciprocal, no error term to carry.  invc = 1 for the first entry and 1/2 for the last
//   two, so x near 1 keeps full RELATIVE accuracy (E + L == 0 there).
//   log2|x| = (E + L_hi) + C1h*r + [L_lo + r*(C1l + r*(C2 + r*C3 ...))]
//   -- h1 = float(biased exponent) + (L_hi - 127) is exact because L_hi is a multiple of 2^-15;
//   the leading product C1h*r is never rounded on its own: h = fma(C1h, r, h1), and the rounding
//   error of that sum, fma(C1h, r, h1 - h), joins the bracket (lo);
//   k = rint(64 y h) straight from fma(y, h, 1.5 * 2^17): 2^t = 2^n * T[j] * 2^f, n = k >> 6,
//   j = k & 63, f = fma(y, lo, fma(y, h, -k/64)), |f| <= 2^-7, T[j] = 2^(j/64) as T_hi (1 + T_rel),
//   2^f - 1 = f*(E1 + f*(E2 + f*E3)), result = T_hi + T_hi*(f*g + T_rel).
// Error: <= 0.5 (final rounding) + ~0.06 ULP; tests bound it by 1 ULP.
// The core declines (returns false) anything that is not "normal positive
// magnitude, result comfortably inside the normal range"; the caller then uses
// pow_f32 above for that element.  There is no separate test of the INPUT: zero, denormal,
// infinite, NaN and (for a non-integer exponent) negative bases all push the biased-exponent
// term far enough that the one range test on t (or on log2|x| when |y| < 1) rejects them.
struct f2 { float x, y; };
SMB_HD f2 f2_make(float a, float b) { f2 r; r.x = a; r.y = b; return r; }
SMB_HD f2 f2_splat(float a) { return f2_make(a, a); }
SMB_HD f2 f2_fma(f2 a, f2 b, f2 c) {
#if defined(__CUDA_ARCH__)
    const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(c.x, c.y));
    return f2_make(r.x, r.y);
#else
    return f2_make(ffma(a.x, b.x, c.x), ffma(a.y, b.y, c.y));
#endif
}
SMB_HD f2 f2_mul(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return f2_make(r.x, r.y);
#else
    return f2_make(fmul(a.x, b.x), fmul(a.y, b.y));
#endif
}
SMB_HD f2 f2_add(f2 a, f2 b) {
#if defined(__CUDA_ARCH__)
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return f2_make(r.x, r.y);
#else
    return f2_make(fadd(a.x, b.x), fadd(a.y, b.y));
#endif
}
// Negation is free: FFMA2 / FADD2 / FMUL2 take a per-operand negate modifier.
SMB_HD f2 f2_neg(f2 a) { return f2_make(-a.x, -a.y); }
SMB_HD f2 f2_sub(f2 a, f2 b) { return f2_add(a, f2_neg(b)); }
SMB_HD f2 f2_fnma(f2 a, f2 b, f2 c) { return f2_fma(f2_neg(a), b, c); } // c - a*b, one rounding

struct PowTabLog { float c, l_hi, l_lo, pad; }; // c: the entry's centre (large-y table) or its 8-bit reciprocal (small-y table)

// MUFU.RCP seed (the f64 core refines it); the host build stands in with a correctly rounded 1/x.
SMB_HD float rcp_seed(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
struct PowTabExp { float t_hi, t_rel; }; // 2^(j/64) = t_hi (1 + t_rel)

// Table access.  Host build: the compact tables behind `tab_*`.  Device: the tables live in
// shared memory (smb_s_pow below, filled per CTA by PowF32Fn::block_init) with each
// entry replicated across the lanes that share a shared-memory wavefront (8 lanes for the
// 16-byte log entries, 16 for the 8-byte exp entries), so a warp reading 32 unrelated entries is
// bank-conflict free.  (The compact layout measured 58 % conflict replays and saturated the LSU
// data pipe at 97 %: profiles/r1_pow_ncu.md.)  The lane's replica offset is OR-ed into the
// entry offset and the array base is a compile-time shared address, so a lookup costs
// SHF + LOP3 + LDS [R + imm].
#if defined(__CUDACC__)
static_assert(SMB_POW_LOG_ENTRIES * 8 * 16 == 16384, "pow_tab_exp_at hard-codes the exp table's offset");
#define SMB_POW_LOG_STRIDE 8   /* entries of PowTabLog between consecutive j */
#define SMB_POW_EXP_STRIDE 16  /* entries of PowTabExp between consecutive j */
} // namespace smb
// C linkage: the lookups below name these arrays from inline PTX, so that their (compile-time)
// shared addresses fold into the LDS immediate instead of going through a generic pointer.
extern "C" {
struct SmbPowTabs { // one object, so both tables share one base register (the exp table sits at +16384)
    smb::PowTabLog log[SMB_POW_LOG_ENTRIES * SMB_POW_LOG_STRIDE];
    smb::PowTabExp exp[SMB_POW_EXP_ENTRIES * SMB_POW_EXP_STRIDE];
};
__shared__ __align__(128) SmbPowTabs smb_s_pow;
}
namespace smb {
#else
#define SMB_POW_LOG_STRIDE 1
#define SMB_POW_EXP_STRIDE 1
#endif
// Per-thread lookup state: this lane's replica offsets (bytes) into the two tables, zero on the
// host.  On the device they are fetched from a small global table rather than computed from the
// thread index: ptxas would otherwise re-derive `(tid & 7) * 16` at every use to save the
// register, which costs a second LOP3 per lookup.
// PowConsts: constants that ride in as kernel parameters (uniform registers) -- `one` keeps
// `(u & 0x007fffff) | one` ONE three-input LOP3 (with both words literal ptxas emits an AND and an
// OR), the floats feed FFMA2's scalar-splat operand without a per-vector MOV.
struct PowConsts { uint32_t one; float s_c3, e3; };
SMB_HD PowConsts pow_consts() { PowConsts c; c.one = 0x3f800000u; c.s_c3 = SMB_POW_S_C3; c.e3 = 0.05547422543168068f; return c; }
struct PowLane { uint32_t log_off, exp_off; PowConsts c; };
#if defined(__CUDACC__)
static __device__ uint2 d_pow_lane_tab[16] = {{0, 0}, {16, 8}, {32, 16}, {48, 24}, {64, 32}, {80, 40}, {96, 48}, {112, 56},
                                              {0, 64}, {16, 72}, {32, 80}, {48, 88}, {64, 96}, {80, 104}, {96, 112}, {112, 120}};
#endif
SMB_HD PowLane pow_lane(uint32_t tid, PowConsts c) {
    PowLane l;
    l.c = c;
#if defined(__CUDA_ARCH__)
    const uint2 o = d_pow_lane_tab[tid & 15u];
    l.log_off = o.x;
    l.exp_off = o.y;
#else
    (void)tid;
    l.log_off = l.exp_off = 0;
#endif
    return l;
}
SMB_HD PowTabLog pow_tab_log_at(const PowTabLog *tab, uint32_t lane_off, uint32_t u) {
#if defined(__CUDA_ARCH__)
    // j = top 7 mantissa bits = (u >> 16) & 127; entry j starts at byte j * 128.
    // SHF + LOP3 + LDS [R + imm]
    const uint32_t off = ((u >> 9) & (127u << 7)) | lane_off;
    PowTabLog e;
    asm("{\n\t.reg .u32 a;\n\t.reg .u64 b;\n\tmov.u64 b, smb_s_pow;\n\tcvt.u32.u64 a, b;\n\tadd.u32 a, a, %4;\n\t"
        "ld.shared.v4.f32 {%0,%1,%2,%3}, [a];\n\t}" : "=f"(e.c), "=f"(e.l_hi), "=f"(e.l_lo), "=f"(e.pad) : "r"(off));
    return e;
#else
    (void)lane_off;
    return tab[(u >> 16) & 127u];
#endif
}
SMB_HD PowTabExp pow_tab_exp_at(const PowTabExp *tab, uint32_t lane_off, uint32_t k) {
#if defined(__CUDA_ARCH__)
    const uint32_t off = ((k << 7) & (63u << 7)) | lane_off; // entry j = k & 63 starts at byte j * 128
    PowTabExp e;
    asm("{\n\t.reg .u32 a;\n\t.reg .u64 b;\n\tmov.u64 b, smb_s_pow;\n\tcvt.u32.u64 a, b;\n\tadd.u32 a, a, %2;\n\t"
        "ld.shared.v2.f32 {%0,%1}, [a+16384];\n\t}" : "=f"(e.t_hi), "=f"(e.t_rel) : "r"(off));
    return e;
#else
    (void)lane_off;
    return tab[k & 63u];
#endif
}
// zbits + ((k & ~63) << 17): mask, then ONE multiply-add (ptxas would otherwise shift, mask, add)
SMB_HD uint32_t pow_scale_bits(uint32_t k, uint32_t zbits) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("mad.lo.u32 %0, %1, 131072, %2;" : "=r"(r) : "r"(k & 0xffffffc0u), "r"(zbits));
    return r;
#else
    return ((k & 0xffffffc0u) << 17) + zbits;
#endif
}

// Host-side facts about the (uniform) exponent that gate the fast core.
SMB_HD bool pow_f32_fast_ok(const PowExpF32 &pe) {
    const uint32_t ay = f2u(pe.y) & 0x7fffffffu;
    // finite, non-zero, |y| < 2^64 (so y*log2 x cannot overflow), not so tiny that y*log2 x is denormal
    return pe.y_class == 0 && ay < 0x5f800000u && ay > 0x1f800000u;
}

// Two elements at once, branch-free.  Returns true when BOTH results are valid;
// false when either element needs pow_f32 (the values written are then garbage,
// but every table index and operation along the way is safe).  The caller
// accumulates the flag over a whole vector and branches once.
//
// TIER (chosen on the host from the uniform exponent):
//   POW_TIER_SMALL  (|y| <= 8): the division-free r-series above on the invc table, degree-3 fit, no
//                   renormalisation of the log2 tail -- what is dropped only matters once multiplied
//                   by a large exponent;
//   POW_TIER_MEDIUM (|y| <= 256): the same with a degree-5 fit.  The r^2 term is carried in one
//                   float (an error of 2^-38.6 * y in t) and the un-renormalised tail y * lo stays
//                   below 2^-7, inside the exp2 polynomial's range: <= 0.53 ULP measured;
//   POW_TIER_LARGE: the {c, log2 c} table and the series in p = (m-c)/(m+c), whose second term is
//                   p^3 (the r-series would need r^2 in two floats), tail renormalised.
// SIGN: POW_SIGN_REJECT -- y is not an integer, a negative base must reach the slow path (NaN):
//         the sign bit is converted together with the biased exponent, which adds 256 to log2|x|;
//       POW_SIGN_EVEN   -- even integer y, the sign of the base is dropped;
//       POW_SIGN_ODD    -- odd integer y, the result takes the sign of the base.
// Y_LT_1 (|y| < 1): the range test is made on log2|x| instead of t (|t| < |log2|x|| then).
//       POW_SIGN_RUNTIME -- the exponent is not known when the kernel is compiled (op chains): the
//         two masks say what to do (abs_mask 0x7fffffff for an integer y, sign_or 0x80000000 for an
//         odd one), and the range test is made on both t and log2|x|.
enum { POW_SIGN_REJECT = 0, POW_SIGN_EVEN = 1, POW_SIGN_ODD = 2, POW_SIGN_RUNTIME = 3 };
enum { POW_TIER_LARGE = 0, POW_TIER_SMALL = 1, POW_TIER_MEDIUM = 2 };
template<int TIER, int SIGN, bool Y_LT_1>
SMB_HD bool pow_f32_pair_fast(float x0, float x1, float y, PowLane lane,
                              const PowTabLog *tab_log, const PowTabExp *tab_exp, float *r0, float *r1,
                              uint32_t abs_mask = 0xffffffffu, uint32_t sign_or = 0u) {
    uint32_t u0 = f2u(x0), u1 = f2u(x1);
    const uint32_t s0 = u0, s1 = u1;
    if (SIGN == POW_SIGN_RUNTIME) { u0 &= abs_mask; u1 &= abs_mask; }
    else if (SIGN != POW_SIGN_REJECT) { u0 &= 0x7fffffffu; u1 &= 0x7fffffffu; } // integer y: |x| from here on
    // ---- log2 |x|,  |x| = 2^E * m,  m in [1, 2) ----------------------------------
    const f2 m = f2_make(u2f((u0 & 0x007fffffu) | lane.c.one), u2f((u1 & 0x007fffffu) | lane.c.one));
    const PowTabLog t0 = pow_tab_log_at(tab_log, lane.log_off, u0), t1 = pow_tab_log_at(tab_log, lane.log_off, u1);
    // log2(m * invc) [or log2(m / c)] = lead_a * lead_b (the leading product, never rounded on its own: it joins
    // E + L_hi inside one fma below) + tail_a * tail_b (+ tail2_a * tail2_b, large tier)
    f2 lead_a, lead_b, tail_a, tail_b, tail2_a, tail2_b;
    tail2_a = tail2_b = f2_splat(0.0f);
    if (TIER != POW_TIER_LARGE) {
        // exact; .c holds invc here.  Scalar FMAs: the operands come straight from the LDS
        // registers, packing them first would cost more moves than the packed form saves.
        const f2 r = f2_make(ffma(m.x, t0.c, -1.0f), ffma(m.y, t1.c, -1.0f));
        const f2 c1h = f2_splat(SMB_POW_C1H);              // 1/ln2 = c1h + c1l
        lead_a = c1h; lead_b = r;
        f2 q;
        if (TIER == POW_TIER_SMALL) {
            q = f2_fma(r, f2_splat(lane.c.s_c3), f2_splat(SMB_POW_S_C2));
        } else { // two more terms: fit error 6e-13 instead of 7e-10
            q = f2_fma(r, f2_splat(SMB_POW_L_C5), f2_splat(SMB_POW_L_C4));
            q = f2_fma(r, q, f2_splat(SMB_POW_L_C3));
            q = f2_fma(r, q, f2_splat(SMB_POW_L_C2));
        }
        q = f2_fma(r, q, f2_splat(SMB_POW_C1L));
        tail_a = r; tail_b = q;
    } else {
        // The r-series would need its r^2 term in two floats once y is large; the odd series in
        // p = (m - c)/(m + c), |p| <= 0.0059, has p^3 as its second term and does not.
        // p = p_hi + p_lo: MUFU.RCP seed + FMA residual steps; m - c is exact (Sterbenz), m + c is
        // carried with its rounding error.
        const f2 num = f2_sub(m, f2_make(t0.c, t1.c));
        const f2 two = f2_splat(2.0f);
        const f2 den = f2_fma(m, two, f2_neg(num));        // 2m - (m - c) = m + c, one rounding
        const f2 rc = f2_make(rcp_seed(den.x), rcp_seed(den.y));
        const f2 p_hi = f2_mul(num, rc);
        f2 res = f2_fnma(p_hi, den, num);
        const f2 den_lo = f2_sub(f2_fma(m, two, f2_neg(den)), num);  // (m + c) - den, exactly
        res = f2_fnma(p_hi, den_lo, res);
        const f2 p_lo = f2_mul(res, rc);
        const f2 s = f2_mul(p_hi, p_hi);
        // log2(m/c) = C0*p + p^3*(C1 + C2 p^2): leading product exact (two floats), the rest folded
        // into one coefficient  c0l + s*(C1 + C2 s)
        const f2 c0h = f2_splat(2.885390043258667f);       // 2/ln2 = c0h + c0l
        f2 qq = f2_fma(s, f2_splat(0.5767093896865845f), f2_splat(0.9617967009544373f));
        qq = f2_fma(s, qq, f2_splat(3.851926067000022e-08f));
        lead_a = c0h; lead_b = p_hi;
        tail_a = c0h; tail_b = p_lo;
        tail2_a = p_hi; tail2_b = qq;
    }
    // E + L_hi, exactly.  The upper half-word of x is [sign | biased exponent | j] with j the
    // table index, so float(u >> 16) / 128 = 256 sign + biased exponent + j/128 (one I2F with a
    // half-word selector, no shift); the table's L_hi word is L_hi - 127 - j/128, a multiple of
    // 2^-15, so the FMA below is exact (every intermediate fits 24 bits, magnitude < 512).
    // A rejected sign rides along as +256.  Scalar -- the table words are used once.
    const float e0 = (float)(uint16_t)(u0 >> 16), e1 = (float)(uint16_t)(u1 >> 16);
    const f2 h1 = f2_make(ffma(e0, 0.0078125f, t0.l_hi), ffma(e1, 0.0078125f, t1.l_hi));
    // h2 = RN(h1 + a*b) from the exact product; h1 - h2 is exact (|a*b| <= |h1| / 2 or h1 == 0, so h2 lies within a
    // factor 2 of h1: Sterbenz), and a*b + (h1 - h2) -- the rounding error of h2, at most ulp(h2)/2 -- is again one fma:
    // three packed operations where the product, its error, the sum and the sum's error took five.
    const f2 h2 = f2_fma(lead_a, lead_b, h1);
    const f2 l2 = f2_fma(lead_a, lead_b, f2_sub(h1, h2));
    f2 lo = f2_fma(tail_a, tail_b, l2);
    if (TIER == POW_TIER_LARGE) lo = f2_fma(tail2_a, tail2_b, lo);
    lo = f2_make(fadd(t0.l_lo, lo.x), fadd(t1.l_lo, lo.y));
    f2 h3 = h2;
    if (TIER == POW_TIER_LARGE) {
        // renormalise: L_lo alone can reach 2^-16, too coarse a tail once multiplied by a large y
        h3 = f2_add(h2, lo);
        lo = f2_add(f2_sub(h2, h3), lo);
    }
    // ---- t = y * log2|x| as th + tl ---------------------------------------------
    const f2 y2 = f2_splat(y);
    // k = rint(64 y h3) straight from the exact product: the sum with 1.5 * 2^17 (ulp 2^-6) leaves k in the low mantissa
    // bits and k/64 after subtracting the constant again; kd stands in for th in the range test (they differ by < 2^-6).
    const f2 s17 = f2_splat(196608.0f);
    const f2 tk = f2_fma(y2, h3, s17);
    const f2 kd = f2_sub(tk, s17);                         // k / 64, exactly
    const f2 th = kd;
    // The one validity test.  Results outside the comfortable normal range (incl. overflow and
    // underflow) go to the slow path, and so does every input that is not a normal number: a zero
    // or denormal has log2 <= -126 here, inf / NaN >= 128, a rejected negative base >= 129.
    const f2 rng = Y_LT_1 ? h3 : th;
    bool ok = fabsf(rng.x) < 125.0f && fabsf(rng.y) < 125.0f;
    if (SIGN == POW_SIGN_RUNTIME) ok = ok && fabsf(h3.x) < 125.0f && fabsf(h3.y) < 125.0f; // |y| unknown: both
    // ---- 2^t --------------------------------------------------------------------
    const uint32_t k0 = f2u(tk.x), k1 = f2u(tk.y);         // biased by 0x48400000, a multiple of 64 that vanishes << 17
    f2 f = f2_fma(y2, h3, f2_neg(kd));                     // y h3 - k/64 from the exact product: |f| <= 2^-7, error <= 2^-32
    f = f2_fma(y2, lo, f);
    const PowTabExp x0e = pow_tab_exp_at(tab_exp, lane.exp_off, k0), x1e = pow_tab_exp_at(tab_exp, lane.exp_off, k1);
    f2 g = f2_fma(f, f2_splat(lane.c.e3), f2_splat(0.24022682011127472f));
    g = f2_fma(f, g, f2_splat(0.6931471824645996f));
    // 2^(j/64 + f) = T_hi (1 + T_rel) (1 + w),  w = 2^f - 1 = f g:  T_hi + T_hi (f g + T_rel)  (T_rel w < 2^-31 dropped).
    // Scalar: the table words feed straight from the LDS registers, and the per-entry T_rel rides in the fma's addend.
    const float z0 = ffma(x0e.t_hi, ffma(f.x, g.x, x0e.t_rel), x0e.t_hi);
    const float z1 = ffma(x1e.t_hi, ffma(f.y, g.y, x1e.t_rel), x1e.t_hi);   // in [0.99, 2.01)
    // scale by 2^n, n = k >> 6, through the exponent field (|n| <= 125 keeps the result normal):
    // (k & ~63) << 17 == n << 23 (the bias 0x48400000 << 17 vanishes mod 2^32)
    uint32_t b0 = pow_scale_bits(k0, f2u(z0)), b1 = pow_scale_bits(k1, f2u(z1));
    if (SIGN == POW_SIGN_ODD) { b0 |= s0 & 0x80000000u; b1 |= s1 & 0x80000000u; } // keep the base's sign
    if (SIGN == POW_SIGN_RUNTIME) { b0 |= s0 & sign_or; b1 |= s1 & sign_or; }
    *r0 = u2f(b0);
    *r1 = u2f(b1);
    return ok;
}

// Host-side facts about the (uniform) exponent that select the variant.
SMB_HD bool pow_f32_small_y(const PowExpF32 &pe) { return (f2u(pe.y) & 0x7fffffffu) <= 0x41000000u; } // |y| <= 8
SMB_HD int pow_f32_tier(const PowExpF32 &pe) {
    const uint32_t ay = f2u(pe.y) & 0x7fffffffu;
    return ay <= 0x41000000u ? POW_TIER_SMALL : ay <= 0x43800000u ? POW_TIER_MEDIUM : POW_TIER_LARGE; // 8, 256
}
SMB_HD int pow_f32_sign_mode(const PowExpF32 &pe) { return !pe.y_is_int ? POW_SIGN_REJECT : pe.y_is_odd ? POW_SIGN_ODD : POW_SIGN_EVEN; }
SMB_HD bool pow_f32_y_lt_1(const PowExpF32 &pe) { return (f2u(pe.y) & 0x7fffffffu) < 0x3f800000u; }

// ============================================================== double pow ===
// Double-double (hi + lo, |lo| <= ulp(hi)/2) helpers built on FMA.
struct dd { double hi, lo; };
SMB_HD dd two_sum(double a, double b) {
    double s = dadd(a, b);
    double bb = dsub(s, a);
    double e = dadd(dsub(a, dsub(s, bb)), dsub(b, bb));
    return dd{s, e};
}
SMB_HD dd fast_two_sum(double a, double b) { // |a| >= |b|
    double s = dadd(a, b);
    double e = dsub(b, dsub(s, a));
    return dd{s, e};
}
SMB_HD dd two_prod(double a, double b) {
    double p = dmul(a, b);
    double e = dfma(a, b, -p);
    return dd{p, e};
}
SMB_HD dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = dfma(a.hi, b.lo, dfma(a.lo, b.hi, p.lo));
    return fast_two_sum(p.hi, p.lo);
}
SMB_HD dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = dfma(a.lo, b, p.lo);
    return fast_two_sum(p.hi, p.lo);
}
SMB_HD dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    s.lo = dadd(s.lo, dadd(a.lo, b.lo));
    return fast_two_sum(s.hi, s.lo);
}
SMB_HD dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo = dadd(s.lo, a.lo);
    return fast_two_sum(s.hi, s.lo);
}

// log2 of a positive finite NORMALISED-OR-DENORMAL double, as a double-double
// with relative error < 2^-68:
//   m in [sqrt(1/2), sqrt(2)), p = (m-1)/(m+1) as a double-double quotient,
//   log2(m) = C0*p + p^3 * Q(s): the leading term in double-double, the tail
//   (<= 2^-6.7 of the total) needs 2^-62 relative: s in double-double only for
//   the p^3*c1 term, the rest in plain double.
SMB_HD dd log2_dd(double x) {
    uint64_t ux = d2u(x);
    int64_t eoff = 0;
    if ((ux >> 52) == 0) { // denormal: scale by 2^54 (exact)
        x = dmul(x, 18014398509481984.0);
        ux = d2u(x);
        eoff = -54;
    }
    int64_t hi = (int64_t)(ux >> 32);
    int64_t eadj = (hi - 0x3fe6a09e) >> 20;
    double m = u2d(ux - ((uint64_t)eadj << 52));
    const double e = (double)(eadj + eoff);
    double num = dsub(m, 1.0);            // exact (Sterbenz)
    dd den = two_sum(m, 1.0);             // m + 1 exactly, as hi + lo
    // p = num / den to ~2^-100: p_hi = num * r, two residual corrections.
    double r = 1.0 / den.hi;
    double p_hi = dmul(num, r);
    // residual = num - p_hi * (den.hi + den.lo), evaluated exactly enough with FMA
    double res = dfma(-p_hi, den.hi, num);
    res = dfma(-p_hi, den.lo, res);
    double p_lo = dmul(res, r);
    dd p = fast_two_sum(p_hi, p_lo);
    // s = p^2 (double-double), plain-double copy for the tail polynomial
    dd s = dd_mul(p, p);
    const double sd = s.hi;
    // tail: Q(s) = c2 + s*(c3 + ...), coefficients (2/ln2)/(2k+1), k = 2..
    // truncation: need s^k/(2k+1) < 2^-70 relative: 0.02944^13 = 2^-66 -> k up to 14.
    const double c2 = 0.57707801635558536295;
    const double c3 = 0.41219858311113240210;
    const double c4 = 0.32059889797532520164;
    const double c5 = 0.26230818925253880134;
    const double c6 = 0.22195308321368667805;
    const double c7 = 0.19235933878519512098;
    const double c8 = 0.16972882833987804793;
    const double c9 = 0.15186263588304877972;
    const double c10 = 0.13739952770371080070;
    const double c11 = 0.12545174268599681802;
    const double c12 = 0.11541560327111707259;
    const double c13 = 0.10686629932510840055;
    const double c14 = 0.09949620971648024;
    const double c15 = 0.09307709941219118757;
    double q = dfma(sd, c15, c14);
    q = dfma(sd, q, c13);
    q = dfma(sd, q, c12);
    q = dfma(sd, q, c11);
    q = dfma(sd, q, c10);
    q = dfma(sd, q, c9);
    q = dfma(sd, q, c8);
    q = dfma(sd, q, c7);
    q = dfma(sd, q, c6);
    q = dfma(sd, q, c5);
    q = dfma(sd, q, c4);
    q = dfma(sd, q, c3);
    q = dfma(sd, q, c2);
    // c1 + s*q in double-double (c1 = 2/(3 ln2) split hi/lo)
    const dd c1 = dd{0.9617966939259756, 5.0577616648125907e-17};
    dd t1 = dd_add(c1, dd_mul_d(s, q));
    // c0 + s*t1 in double-double (c0 = 2/ln2 split hi/lo)
    const dd c0 = dd{2.8853900817779268, 4.0710547481862066e-17};
    dd t0 = dd_add(c0, dd_mul(s, t1));
    dd l = dd_mul(p, t0);
    return dd_add(dd{e, 0.0}, l);
}

// exp2 of a double-double argument t = th + tl, |t| <= 1100, result as double
// with < 0.5 + 2^-10 ULP error before scaling; scaling handles gradual underflow
// with a single final rounding.
SMB_HD double exp2_from_dd(dd t) {
    const double shifter = 6755399441055744.0;
    double tn = dadd(t.hi, shifter);
    int32_t n = (int32_t)(uint32_t)d2u(tn);
    double fn = dsub(tn, shifter);
    // f = (th - n) + tl, |f| <= 0.5 (+tiny)
    dd f = fast_two_sum(dsub(t.hi, fn), t.lo);
    // z = f * ln2 in double-double
    const dd ln2 = dd{0.6931471805599453, 2.3190468138462996e-17};
    dd z = dd_mul(f, ln2);
    // e^z = 1 + z + z^2/2 + z^3 * R(z): |z| <= 0.3466.  Leading terms in
    // double-double, R in double (Taylor through z^15: 0.3466^16/16! = 2^-69).
    const double zh = z.hi;
    double r = 1.0 / 1307674368000.0;            // 1/15!
    r = dfma(zh, r, 1.0 / 87178291200.0);        // 1/14!
    r = dfma(zh, r, 1.0 / 6227020800.0);         // 1/13!
    r = dfma(zh, r, 1.0 / 479001600.0);          // 1/12!
    r = dfma(zh, r, 1.0 / 39916800.0);           // 1/11!
    r = dfma(zh, r, 1.0 / 3628800.0);            // 1/10!
    r = dfma(zh, r, 1.0 / 362880.0);             // 1/9!
    r = dfma(zh, r, 1.0 / 40320.0);              // 1/8!
    r = dfma(zh, r, 1.0 / 5040.0);               // 1/7!
    r = dfma(zh, r, 1.0 / 720.0);                // 1/6!
    r = dfma(zh, r, 1.0 / 120.0);                // 1/5!
    r = dfma(zh, r, 1.0 / 24.0);                 // 1/4!
    r = dfma(zh, r, 1.0 / 6.0);                  // 1/3!
    // z^2/2 + z^3*r = z^2 * (0.5 + z*r)
    dd z2 = dd_mul(z, z);
    dd w = dd_mul_d(z2, dfma(zh, r, 0.5));       // relative 2^-53 of a term <= 0.07 of total
    dd sum = dd_add(z, w);
    sum = dd_add_d(sum, 1.0);                    // 1 + z + ...
    // scale by 2^n with one rounding: for results in the normal range add to the
    // exponent; near underflow do it in two exact steps on hi+lo.
    if (n > -1000) {
        double h = dadd(sum.hi, sum.lo);
        if (n > 1000) { // overflow side: split the scaling to stay finite until the end
            h = u2d(d2u(h) + ((uint64_t)(int64_t)(n - 100) << 52));
            return dmul(h, 1.2676506002282294e30); // 2^100
        }
        return u2d(d2u(h) + ((uint64_t)(int64_t)n << 52));
    }
    // n <= -1000: result may be denormal.  Scale hi and lo by 2^(n+200) exactly
    // (still normal), then one multiply by 2^-200 performs the only rounding on
    // hi; lo's contribution is folded in first with an FMA-free exact add since
    // |lo| << ulp(hi) after scaling can still decide a tie.
    double hs = u2d(d2u(sum.hi) + ((uint64_t)(int64_t)(n + 200) << 52));
    double ls = sum.lo == 0.0 ? 0.0 : u2d(d2u(sum.lo) + ((uint64_t)(int64_t)(n + 200) << 52));
    const double tiny = 6.223015277861142e-61; // 2^-200
    return dfma(hs, tiny, dmul(ls, tiny));
}

SMB_HD double pow_f64(double x, const PowExpF64 &pe) {
    const uint64_t ux = d2u(x);
    bool negate = false;
    if (!(pe.y_class == 0 && (ux - 1ull) < 0x7fefffffffffffffull)) {
        double sp;
        if (pow_special<double, PowExpF64>(x, pe, &sp, &negate)) return sp;
    }
    const double ax = u2d(ux & 0x7fffffffffffffffull);
    dd l = log2_dd(ax);
    // |y * log2 x| beyond ~1100 is a certain overflow / underflow; clamp y*l.hi
    // first so the double-double product cannot produce inf - inf.
    double rough = dmul(pe.y, l.hi);
    double r;
    if (!(rough < 1100.0)) r = (double)INFINITY;
    else if (!(rough > -1100.0)) r = 0.0;
    else {
        dd t = dd_mul_d(l, pe.y);
        // y itself can be huge while l is tiny; the product above is exact to
        // 2^-100 relative, fine.
        r = exp2_from_dd(t);
    }
    return negate ? -r : r;
}

// ---- table-driven f64 pow core ------------------------------------------------
// The double-double core above costs ~300 instructions per element (FP64 pipe 78 % busy,
// 1.4 TB/s).  Same structure as the f32 fast core, in double: the 128-entry table brings
// |p| = |(m-c)/(m+c)| under 0.006, so only the LEADING terms need two-double arithmetic
// (C0*p as an exact product, e + L_hi exact because L_hi is a multiple of 2^-40) and
// everything else runs in plain double: ~50 FP64 operations per element.
//   log2|x| = (e + L_hi) + [C0*p + L_lo + p*(c0l + s*(C1 + s*(C2 + s*(C3 + s*C4))))],
//   1/(m+c): f32 MUFU seed + one Newton step (2^-44), the quotient's error recovered by two
//   FMA residual steps;  2^t = 2^(k>>6) * T[k&63] * (1 + f*(E1 + ... + E7 f^6)), |f| <= 2^-7.
// Declines (returns false) anything outside "normal magnitude, result well inside the
// normal range"; pow_f64 handles those.  Error <= 0.52 ULP measured (bound 1 ULP).
struct PowTabLog64 { double c, l_hi, l_lo, pad; };   // generated (compact) form
struct PowTabLog64A { double c, l_hi; };             // device layout: split so that every lookup is
struct PowTabLog64B { double l_lo; };                // one conflict-free LDS.128 / LDS.64
struct PowTabExp64 { double t_hi, t_rel; }; // 2^(j/64) = t_hi (1 + t_rel)
#if defined(__CUDA_ARCH__)
#define SMB_POW64_A_STRIDE 8    /* 16-byte entries: the 8 lanes of a wavefront get their own replica */
#define SMB_POW64_B_STRIDE 16   /*  8-byte entries: 16 lanes per wavefront */
#define SMB_POW64_EXP_STRIDE 8
#else
#define SMB_POW64_A_STRIDE 1
#define SMB_POW64_B_STRIDE 1
#define SMB_POW64_EXP_STRIDE 1
#endif

SMB_HD bool pow_f64_small_y(const PowExpF64 &pe) { return (d2u(pe.y) & 0x7fffffffffffffffull) <= 0x4020000000000000ull; } // |y| <= 8
SMB_HD bool pow_f64_fast_ok(const PowExpF64 &pe) {
    const uint64_t ay = d2u(pe.y) & 0x7fffffffffffffffull;
    // finite, non-zero, 2^-400 < |y| < 2^400: y*log2 x can neither overflow nor go denormal
    return pe.y_class == 0 && ay < 0x58f0000000000000ull && ay > 0x26f0000000000000ull;
}

// SMALL_Y (|y| <= 8, chosen on the host): the rounding error of m + c (2^-53 relative in p, below
// 2^-59.9 absolute in log2 x) and the renormalisation of the log2 tail are dropped -- they only matter
// once multiplied by a large exponent.
template<bool SMALL_Y, bool ODD_Y>
SMB_HD bool pow_f64_fast(double x, double y, uint64_t sign_reject, const PowTabLog64A *tab_a, const PowTabLog64B *tab_b,
                         const PowTabExp64 *tab_exp, double *out) {
    const uint64_t u = d2u(x);
    const uint64_t a = u & 0x7fffffffffffffffull;
    const uint32_t eb = (uint32_t)(a >> 52);
    bool ok = (eb - 1u) < 2046u && (u & sign_reject) == 0ull; // normal finite; sign allowed
    // |x| = 2^E * m, m in [1, 2)
    const double m = u2d((a & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
    const uint32_t j = (uint32_t)(a >> 45) & 127u;
    const PowTabLog64A t = tab_a[j * SMB_POW64_A_STRIDE];
    const double t_l_lo = tab_b[j * SMB_POW64_B_STRIDE].l_lo;
    const double num = dsub(m, t.c);                       // exact
    const double den = dfma(m, 2.0, -num);                 // m + c, one rounding
    double r = (double)rcp_seed((float)den);
    r = dfma(r, dfma(-den, r, 1.0), r);                    // one Newton step: 2^-44
    const double p_hi = dmul(num, r);
    double res = dfma(-p_hi, den, num);
    if (!SMALL_Y) {
        const double den_lo = dsub(dfma(m, 2.0, -den), num);   // exact error of den
        res = dfma(-p_hi, den_lo, res);
    }
    const double p_lo = dmul(res, r);
    const double s = dmul(p_hi, p_hi);
    // odd series of log2((1+p)/(1-p)) up to p^9; |p| <= 0.006, so for |y| <= 8 the p^9 term (2^-68 in log2) is below everything
    // that matters and is dropped (one DFMA)
    double q = SMALL_Y ? 0.4121985831111324 : dfma(s, 0.3205988979753252, 0.4121985831111324);
    q = dfma(s, q, 0.5770780163555853);
    q = dfma(s, q, 0.9617966939259756);
    q = dfma(s, q, 4.0710547481862066e-17);                // c0l + s*Q
    const double c0h = 2.8853900817779268;
    const double h1 = dadd((double)(int32_t)eb, t.l_hi);   // exact: integer + multiple of 2^-40
    // The leading product c0h * p_hi joins h1 inside ONE fma (never rounded on its own): h2 = RN(h1 + c0h p_hi);
    // h1 - h2 is exact (|c0h p_hi| <= |h1| / 2 or h1 == 0 -- the unit entries c = 1, c = 2 -- so h2 lies within a factor 2
    // of h1), and the rounding error of h2 is again one fma.  Six FP64 operations where the separately rounded product
    // needed nine.
    const double h2 = dfma(c0h, p_hi, h1);
    double ll = dfma(c0h, p_hi, dsub(h1, h2));
    ll = dfma(c0h, p_lo, ll);
    ll = dfma(p_hi, q, ll);
    const double lo_raw = dadd(t_l_lo, ll);
    double h3 = h2, lo = lo_raw;
    if (!SMALL_Y) { // renormalise: the tail alone can reach 2^-41, too coarse once multiplied by a large y
        h3 = dadd(h2, lo_raw);
        lo = dadd(dsub(h2, h3), lo_raw);
    }
    // k = rint(64 y h3) from the exact product: 1.5 * 2^46 has ulp 2^-6, so the sum leaves k in the low mantissa bits
    // and k/64 once the constant is subtracted again; f = y h3 - k/64 is ONE fma (|f| <= 2^-7, error <= 2^-61).
    const double s46 = 105553116266496.0;                  // 1.5 * 2^46
    const double tk = dfma(y, h3, s46);
    const uint32_t k = (uint32_t)d2u(tk);                  // low word: rint(64 y h3), two's complement
    const double kd = dsub(tk, s46);                       // k / 64, exactly
    ok = ok && fabs(kd) < 1000.0;                          // result well inside the normal range
    double f = dfma(y, h3, -kd);
    f = dfma(y, lo, f);
    const PowTabExp64 e = tab_exp[(k & 63u) * SMB_POW64_EXP_STRIDE];
    // (2^f - 1) / f up to f^5: the f^6 term is 1.5e-5 * 2^-42 (|f| <= 2^-7), 2^-65 of the result -- dropped (one DFMA)
    double g = dfma(f, 0.0001540353039338161, 0.0013333558146428443);
    g = dfma(f, g, 0.009618129107628477);
    g = dfma(f, g, 0.05550410866482158);
    g = dfma(f, g, 0.24022650695910072);
    g = dfma(f, g, 0.6931471805599453);
    const double z = dfma(e.t_hi, dfma(f, g, e.t_rel), e.t_hi); // T_hi (1 + T_rel)(1 + f g), T_rel f g < 2^-60 dropped
    // scale by 2^n, n = (int32)k >> 6 (|n| <= 1000 keeps the result normal)
    const int64_t n = (int64_t)((int32_t)k >> 6);
    uint64_t b = d2u(z) + ((uint64_t)n << 52);
    if (ODD_Y) b |= u & 0x8000000000000000ull;
    *out = u2d(b);
    return ok;
}

// ========================================================= the Op functors ===
// DevOp<OP, T>::apply(a, b [, lane]) -- `lane` only matters for i32 pow.
template<int OP, typename T> struct DevOp;

template<> struct DevOp<OP_ADD, float> { static SMB_HD float apply(float a, float b) { return fadd(a, b); } };
template<> struct DevOp<OP_SUB, float> { static SMB_HD float apply(float a, float b) { return fsub(a, b); } };
template<> struct DevOp<OP_MUL, float> { static SMB_HD float apply(float a, float b) { return fmul(a, b); } };
template<> struct DevOp<OP_DIV, float> {
    static SMB_HD float apply(float a, float b) {
#if defined(__CUDA_ARCH__)
        return __fdiv_rn(a, b);
#else
        return a / b;
#endif
    }
};
template<> struct DevOp<OP_ADD, double> { static SMB_HD double apply(double a, double b) { return dadd(a, b); } };
template<> struct DevOp<OP_SUB, double> { static SMB_HD double apply(double a, double b) { return dsub(a, b); } };
template<> struct DevOp<OP_MUL, double> { static SMB_HD double apply(double a, double b) { return dmul(a, b); } };
template<> struct DevOp<OP_DIV, double> {
    static SMB_HD double apply(double a, double b) {
#if defined(__CUDA_ARCH__)
        return __ddiv_rn(a, b);
#else
        return a / b;
#endif
    }
};
template<> struct DevOp<OP_ADD, int32_t> { static SMB_HD int32_t apply(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); } };
template<> struct DevOp<OP_SUB, int32_t> { static SMB_HD int32_t apply(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); } };
template<> struct DevOp<OP_MUL, int32_t> { static SMB_HD int32_t apply(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); } };
template<> struct DevOp<OP_DIV, int32_t> {
    static SMB_HD int32_t apply(int32_t a, int32_t b) {
        // /0 and INT_MIN/-1 are undefined in the reference (division.h:69); keep
        // the host build of this header from trapping on them.
#if !defined(__CUDA_ARCH__)
        if (b == 0 || (a == (int32_t)0x80000000u && b == -1)) return 0;
#endif
        return a / b;
    }
};
// array ^ array pow (README.md:119-133 recipe): per-element exponent.
template<> struct DevOp<OP_POW, float> {
    static SMB_HD float apply(float a, float b) { return pow_f32(a, classify_exp(b)); }
};
template<> struct DevOp<OP_POW, double> {
    static SMB_HD double apply(double a, double b) { return pow_f64(a, classify_exp(b)); }
};
template<> struct DevOp<OP_POW, int32_t> {
    static SMB_HD int32_t apply(int32_t a, int32_t b) { return powi_scalar(a, b); }
    static SMB_HD int32_t apply_lane(int32_t a, int32_t b) { return powi_lane(a, b); }
};

} // namespace smb
