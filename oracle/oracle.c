/*
 * oracle.c -- CPU restatement of simpleMath's elementwise hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (simplemath_b200/,
 * include/) may link, import or call this file.  It is used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against
 *   (a) the golden vectors held by the reference's own tests
 *       (tests/add.cpp, subtract.cpp, multiply.cpp, division.cpp, pow.cpp),
 *   (b) outputs of the unmodified reference headers compiled into
 *       oracle/_ref/libsmref.so (see oracle/ref_shim.cpp), live when that
 *       library is present and through tests/golden/ fixtures otherwise.
 * Exception: float/double pow.  The reference's only float pow semantics is
 * `std::pow` (include/math/pow.h:8-10); no reference test pins its values
 * (tests/pow.cpp:29-36,101-125 are commented out), so for f32/f64 pow the
 * contract is a ULP bound against std::pow evaluated in double -- "parity
 * unpinned by the reference" for that one op.
 *
 * Every function cites the reference file:line it restates.  Paths are
 * relative to /root/reference.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include <string.h>

#define ORC_MAX_NDIM 6 /* include/math/helpers.h:4 */

enum { ORC_ADD = 0, ORC_SUB = 1, ORC_MUL = 2, ORC_DIV = 3, ORC_POW = 4 };
enum { ORC_F32 = 0, ORC_F64 = 1, ORC_I32 = 2 };

/* ---------------------------------------------------------------- broadcast
 * include/SMUtils.h:34-99.  Right-align ranks, pad leading dims with shape 1 /
 * stride 0 (:53-56,:64-67); dims must be equal or 1 (:76-78); result dim is
 * the max (:80); stride forced to 0 where own dim is 1 and the other's > 1
 * (:83-88).  Returns 0, or 1 where the reference throws std::runtime_error.
 */
int orc_broadcast(const uint64_t *shape1, const uint64_t *strides1, int ndim1,
                  const uint64_t *shape2, const uint64_t *strides2, int ndim2,
                  uint64_t *result_shape, uint64_t *new_strides1,
                  uint64_t *new_strides2, int *out_ndim, uint64_t *total_size) {
    int nd = ndim1 > ndim2 ? ndim1 : ndim2;
    int off1 = nd - ndim1, off2 = nd - ndim2;
    uint64_t total = 1;
    for (int i = 0; i < nd; ++i) {
        uint64_t d1, s1, d2, s2;
        if (i < off1) { d1 = 1; s1 = 0; } else { d1 = shape1[i - off1]; s1 = strides1[i - off1]; }
        if (i < off2) { d2 = 1; s2 = 0; } else { d2 = shape2[i - off2]; s2 = strides2[i - off2]; }
        if (d1 != d2 && d1 != 1 && d2 != 1) return 1;
        result_shape[i] = d1 > d2 ? d1 : d2;
        total *= result_shape[i];
        if (d1 == 1 && d2 > 1) s1 = 0;
        if (d2 == 1 && d1 > 1) s2 = 0;
        new_strides1[i] = s1;
        new_strides2[i] = s2;
    }
    *out_ndim = nd;
    *total_size = total;
    return 0;
}

/* include/math/helpers.h:130-139 */
int orc_is_contiguous(const uint64_t *shape, const uint64_t *stride, int ndim) {
    uint64_t expected = 1;
    for (int i = ndim - 1; i >= 0; --i) {
        if (stride[i] != expected) return 0;
        expected *= shape[i];
    }
    return 1;
}

/* ------------------------------------------------------------- scalar ops
 * AddOp::apply add.h:7-9, SubtractOp::apply subtract.h:7-9,
 * MultiplyOp::apply multiply.h:9-11, DivideOp::apply division.h:10-12,67-70.
 * int32 + - * wrap (two's complement, as _mm256_{add,sub,mullo}_epi32 do,
 * add.h:73-75 / subtract.h:74-76 / multiply.h:77-79); int32 / truncates
 * toward zero; /0 and INT_MIN/-1 are undefined in the reference (SIGFPE on
 * x86) and are excluded from every test input.
 */
static inline int32_t wrap_add(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wrap_sub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
static inline int32_t wrap_mul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }

/* Integer pow, SIMD-lane semantics: include/math/simd/crafted_pow.h:54-103.
 * |exp| taken as _mm256_abs_epi32 does (INT_MIN stays 0x80000000 and is then
 * shifted logically, :60,:79), wrapping mullo (:72,:76), then the fix-ups:
 * 0^(exp>0) -> 0 (:81-84); exp<0 -> 0 except base 1 -> 1 and base -1 ->
 * (exp odd ? -1 : 1) (:85-102). */
int32_t orc_powi_lane(int32_t base, int32_t exp) {
    uint32_t e = exp < 0 ? 0u - (uint32_t)exp : (uint32_t)exp;
    uint32_t cur = (uint32_t)base, res = 1u;
    while (e) {
        if (e & 1u) res *= cur;
        cur *= cur;
        e >>= 1;
    }
    if (base == 0 && exp > 0) res = 0;
    if (exp < 0) {
        if (base == 1) return 1;
        if (base == -1) return (exp & 1) ? -1 : 1;
        return 0;
    }
    return (int32_t)res;
}

/* Integer pow, scalar semantics: PowOp<int>::apply, include/math/pow.h:8-10:
 * std::pow(int,int) promotes to double, the double result is converted to int
 * on return.  On x86-64 that conversion is cvttsd2si, whose out-of-range /
 * NaN / inf result is the "integer indefinite" 0x80000000.  Stated here
 * explicitly so the oracle does not itself depend on UB. */
int32_t orc_powi_scalar(int32_t base, int32_t exp) {
    double r = pow((double)base, (double)exp);
    if (!(r > -2147483649.0 && r < 2147483648.0)) return INT32_MIN;
    return (int32_t)r;
}

static inline float op_f32(int op, float a, float b) {
    switch (op) {
        case ORC_ADD: return a + b;
        case ORC_SUB: return a - b;
        case ORC_MUL: return a * b;
        case ORC_DIV: return a / b;
        default: return powf(a, b); /* pow.h:8-10: std::pow(float,float) == powf */
    }
}
static inline double op_f64(int op, double a, double b) {
    switch (op) {
        case ORC_ADD: return a + b;
        case ORC_SUB: return a - b;
        case ORC_MUL: return a * b;
        case ORC_DIV: return a / b;
        default: return pow(a, b);
    }
}
/* lane != 0: the element is produced by an AVX2 apply_simd lane; lane == 0:
 * by scalar Op::apply.  Only int pow distinguishes the two. */
static inline int32_t op_i32(int op, int32_t a, int32_t b, int lane) {
    switch (op) {
        case ORC_ADD: return wrap_add(a, b);
        case ORC_SUB: return wrap_sub(a, b);
        case ORC_MUL: return wrap_mul(a, b);
        case ORC_DIV: return a / b;
        default: return lane ? orc_powi_lane(a, b) : orc_powi_scalar(a, b);
    }
}

/* Scalar Op::apply on one pair (the parity oracle north_star names). */
int orc_scalar_apply(int op, int dtype, const void *a, const void *b, void *out) {
    if (dtype == ORC_F32) *(float *)out = op_f32(op, *(const float *)a, *(const float *)b);
    else if (dtype == ORC_F64) *(double *)out = op_f64(op, *(const double *)a, *(const double *)b);
    else if (dtype == ORC_I32) *(int32_t *)out = op_i32(op, *(const int32_t *)a, *(const int32_t *)b, 0);
    else return 1;
    return 0;
}

/* ------------------------------------------------- handle_contiguous_arrays
 * include/math/calculate.h:101-134 (AVX2 build, the only one that compiles):
 * vector body while i + 8 <= n, advancing by simd_width (8 for f32/i32, 4 for
 * f64, helpers.h:29-35,62-68,97-103), then scalar tail.  The number of
 * elements produced by SIMD lanes is what matters for int pow.
 */
static uint64_t contiguous_simd_end(int dtype, uint64_t n) {
    uint64_t w = dtype == ORC_F64 ? 4 : 8, i = 0;
    while (i + 8 <= n) i += w;
    return i;
}

static void contiguous(int op, int dtype, const void *a, const void *b, void *out, uint64_t n) {
    uint64_t simd_end = contiguous_simd_end(dtype, n);
    if (dtype == ORC_F32) {
        const float *x = a, *y = b; float *r = out;
        for (uint64_t i = 0; i < n; ++i) r[i] = op_f32(op, x[i], y[i]);
    } else if (dtype == ORC_F64) {
        const double *x = a, *y = b; double *r = out;
        for (uint64_t i = 0; i < n; ++i) r[i] = op_f64(op, x[i], y[i]);
    } else {
        const int32_t *x = a, *y = b; int32_t *r = out;
        for (uint64_t i = 0; i < n; ++i) r[i] = op_i32(op, x[i], y[i], i < simd_end);
    }
}

/* ---------------------------------------------------------- element_wise_op
 * include/math/calculate.h:5-99.
 *  - dispatcher (:10-13): contiguous fast path when the last strides are 1,
 *    the stride tables are equal and row-major contiguous.  The reference also
 *    takes it for every ndim==1 call, ignoring strides, which reads out of
 *    bounds for a stride-0 (broadcast) 1-D operand; the oracle
 *    honours the stride table there instead -- the one documented divergence,
 *    reachable only through undefined behaviour in the reference.
 *  - general loop (:16-30,:54-63,:96): prod_shape suffix products, successive
 *    div/mod of the linear index, offsets += idx*stride, scalar Op::apply
 *    (canVectorize is identically false, :33-46).
 */
int orc_elementwise(int op, int dtype, const void *a, const uint64_t *stride_a,
                    const void *b, const uint64_t *stride_b,
                    const uint64_t *shape, int ndim, uint64_t n, void *out) {
    if (ndim < 1 || ndim > ORC_MAX_NDIM) return 1;
    if (dtype < 0 || dtype > 2 || op < 0 || op > 4) return 1;
    int same = memcmp(stride_a, stride_b, sizeof(uint64_t) * (size_t)ndim) == 0;
    if (stride_a[ndim - 1] == 1 && stride_b[ndim - 1] == 1 && same &&
        orc_is_contiguous(shape, stride_a, ndim)) {
        contiguous(op, dtype, a, b, out, n);
        return 0;
    }
    uint64_t prod[ORC_MAX_NDIM];
    prod[ndim - 1] = 1;
    for (int k = ndim - 2; k >= 0; --k) prod[k] = shape[k + 1] * prod[k + 1];
    #pragma omp parallel for schedule(static) if (n > 100000)
    for (int64_t lin = 0; lin < (int64_t)n; ++lin) {
        uint64_t rem = (uint64_t)lin, oa = 0, ob = 0;
        for (int k = 0; k < ndim; ++k) {
            uint64_t idx = rem / prod[k];
            rem %= prod[k];
            oa += idx * stride_a[k];
            ob += idx * stride_b[k];
        }
        if (dtype == ORC_F32) ((float *)out)[lin] = op_f32(op, ((const float *)a)[oa], ((const float *)b)[ob]);
        else if (dtype == ORC_F64) ((double *)out)[lin] = op_f64(op, ((const double *)a)[oa], ((const double *)b)[ob]);
        else ((int32_t *)out)[lin] = op_i32(op, ((const int32_t *)a)[oa], ((const int32_t *)b)[ob], 0);
    }
    return 0;
}

/* ---------------------------------------------------------- array_scalar_op
 * include/math/calculate.h:137-169: simd_end = n - n % simd_width (:139-140),
 * SIMD body with set1(value) as the RIGHT operand (:141-164), scalar tail
 * Op::apply(a[i], value) (:166-168).  Float/double pow has no SIMD body in
 * the snapshot (pow.h:16-52 commented out); its only semantics
 * is Op::apply == std::pow for every element.
 */
int orc_array_scalar(int op, int dtype, const void *a, const void *scalar, uint64_t n, void *out) {
    if (dtype < 0 || dtype > 2 || op < 0 || op > 4) return 1;
    uint64_t w = dtype == ORC_F64 ? 4 : 8;
    uint64_t simd_end = n - (n % w);
    if (dtype == ORC_F32) {
        const float *x = a; float v = *(const float *)scalar; float *r = out;
        #pragma omp parallel for schedule(static) if (n > 100000)
        for (int64_t i = 0; i < (int64_t)n; ++i) r[i] = op_f32(op, x[i], v);
    } else if (dtype == ORC_F64) {
        const double *x = a; double v = *(const double *)scalar; double *r = out;
        #pragma omp parallel for schedule(static) if (n > 100000)
        for (int64_t i = 0; i < (int64_t)n; ++i) r[i] = op_f64(op, x[i], v);
    } else {
        const int32_t *x = a; int32_t v = *(const int32_t *)scalar; int32_t *r = out;
        #pragma omp parallel for schedule(static) if (n > 100000)
        for (int64_t i = 0; i < (int64_t)n; ++i) r[i] = op_i32(op, x[i], v, (uint64_t)i < simd_end);
    }
    return 0;
}

/* Accuracy reference for float pow (north_star: "within a stated ULP bound
 * against std::pow computed in double"): y^x evaluated by libm pow in double
 * on the exactly-converted f32 inputs. */
void orc_pow_ref_f32(const float *x, float y, uint64_t n, double *out) {
    #pragma omp parallel for schedule(static) if (n > 100000)
    for (int64_t i = 0; i < (int64_t)n; ++i) out[i] = pow((double)x[i], (double)y);
}

/* For f64 pow the accuracy reference is libm powl in long double (x87 80-bit
 * on x86-64), returned as a (hi, lo) pair of doubles so the caller can measure
 * sub-ULP distances. */
void orc_pow_ref_f64(const double *x, double y, uint64_t n, double *hi, double *lo) {
    #pragma omp parallel for schedule(static) if (n > 100000)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        long double r = powl((long double)x[i], (long double)y);
        double h = (double)r;
        hi[i] = h;
        lo[i] = (isfinite(h)) ? (double)(r - (long double)h) : 0.0;
    }
}

/* ------------------------------------------------------------------ dot_product
 * include/math/product.h (AVX build): int32 :26-71 -- 8 wrapping lane accumulators, lanes
 * folded, scalar tail; float :74-118 -- 8 float lane accumulators (mul then add), low+high
 * halves, buf[0]+buf[1]+buf[2]+buf[3], scalar tail; double :121-165 -- 4 lanes likewise.
 * The float/double results depend on this exact association order, which is what is
 * restated here (so the restatement can be pinned bit-for-bit against the compiled
 * reference); the device kernel sums pairwise and is compared by tolerance. */
int32_t orc_dot_i32(const int32_t *a, const int32_t *b, uint64_t n) {
    uint32_t r = 0;
    for (uint64_t i = 0; i < n; ++i) r += (uint32_t)a[i] * (uint32_t)b[i];
    return (int32_t)r; /* wrapping addition is associative: one loop states all orders */
}
float orc_dot_f32(const float *a, const float *b, uint64_t n) {
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0}, result = 0.0f;
    uint64_t i = 0;
    for (; i + 7 < n; i += 8)
        for (int l = 0; l < 8; ++l) lane[l] = lane[l] + a[i + l] * b[i + l];
    float s[4];
    for (int l = 0; l < 4; ++l) s[l] = lane[l] + lane[l + 4];
    result += s[0] + s[1] + s[2] + s[3];
    for (; i < n; ++i) result += a[i] * b[i];
    return result;
}
double orc_dot_f64(const double *a, const double *b, uint64_t n) {
    double lane[4] = {0, 0, 0, 0}, result = 0.0;
    uint64_t i = 0;
    for (; i + 3 < n; i += 4)
        for (int l = 0; l < 4; ++l) lane[l] = lane[l] + a[i + l] * b[i + l];
    double s0 = lane[0] + lane[2], s1 = lane[1] + lane[3];
    result += s0 + s1;
    for (; i < n; ++i) result += a[i] * b[i];
    return result;
}

/* Counter-based input generator shared by bench.py's CPU and GPU legs
 * (SURVEY.md §8d C5): splitmix64 of the flat index -> uniform float in
 * [lo, hi).  The device kernel in simplemath_b200/csrc re-states the same
 * arithmetic; this copy lets the oracle re-evaluate sampled windows. */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
void orc_fill_uniform_f32(float *out, uint64_t first, uint64_t n, uint64_t seed, float lo, float hi) {
    #pragma omp parallel for schedule(static) if (n > 100000)
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        uint64_t h = splitmix64(seed * 0xD1342543DE82EF95ull + first + (uint64_t)i);
        float u = (float)(h >> 40) * (1.0f / 16777216.0f); /* 24 bits -> [0,1) exactly */
        out[i] = lo + (hi - lo) * u;                        /* two roundings, no contraction */
    }
}
