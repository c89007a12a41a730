"""The device math header (simplemath_b200/csrc/smb_math.cuh) compiled for the
host (tests/hostcheck, a test artefact) and swept on the CPU: integer pow
bit-exact against the oracle, float/double pow within the stated ULP bounds
against std::pow evaluated in higher precision.

Stated bounds (also enforced on the GPU in test_gpu_parity.py):
    f32 pow: <= 1 ULP  vs std::pow computed in double   (measured: fast FFMA2 core <= 0.54,
                                                          FP64 reference-accuracy path <= 0.5002)
    f64 pow: <= 1 ULP  vs powl in long double            (measured <= 0.56)
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, assert_same_bits

F32_POW_ULP_BOUND = 1.0
F64_POW_ULP_BOUND = 1.0

HC_DIR = os.path.join(ROOT, "tests", "hostcheck")
HC_LIB = os.path.join(HC_DIR, "libsmb_hostcheck.so")


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HC_DIR, "hostcheck.cpp")
    hdr = os.path.join(ROOT, "simplemath_b200", "csrc", "smb_math.cuh")
    if not os.path.exists(HC_LIB) or os.path.getmtime(HC_LIB) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        subprocess.run(["g++", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared", "-o", HC_LIB, src],
                       check=True, env=env)
    return ctypes.CDLL(HC_LIB)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def hc_pow32(hc, x, y):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    hc.hc_pow_f32(_p(x), ctypes.c_float(y), ctypes.c_uint64(x.size), _p(out))
    return out


def hc_pow32_fast(hc, x, y):
    """The kernel's own path: table-driven packed core for pairs, pow_f32 for declined elements."""
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    dec = ctypes.c_uint64(0)
    hc.hc_pow_f32_fast(_p(x), ctypes.c_float(y), ctypes.c_uint64(x.size), _p(out), ctypes.byref(dec))
    return out, dec.value / max(x.size, 1)


def hc_pow64(hc, x, y):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    hc.hc_pow_f64(_p(x), ctypes.c_double(y), ctypes.c_uint64(x.size), _p(out))
    return out


def test_powi_lane_and_scalar_bit_exact(hc, orc, ref):
    rng = np.random.default_rng(11)
    b = np.concatenate([rng.integers(-2**31, 2**31, 20000, dtype=np.int64), rng.integers(-40, 41, 60000),
                        [0, 1, -1, 2, -2, 46340, 46341, -46341, 2**31 - 1, -2**31]]).astype(np.int32)
    e = np.concatenate([rng.integers(-2**31, 2**31, 10000, dtype=np.int64), rng.integers(-8, 70, 70000 + 10)]).astype(np.int32)
    n = b.size
    for lane, name in ((1, "orc_powi_lane"), (0, "orc_powi_scalar")):
        got = np.empty(n, np.int32)
        hc.hc_powi(_p(b), _p(e), ctypes.c_uint64(n), lane, _p(got))
        f = getattr(orc.h, name)
        want = np.array([f(int(x), int(y)) for x, y in zip(b[:30000], e[:30000])], np.int32)
        assert_same_bits(got[:30000], want, name)
    if ref is not None:  # scalar semantics straight from PowOp<int>::apply of the reference
        got = np.empty(n, np.int32)
        hc.hc_powi(_p(b), _p(e), ctypes.c_uint64(n), 0, _p(got))
        want = np.array([ref.scalar_apply("pow", np.int32, int(x), int(y)) for x, y in zip(b[-20000:], e[-20000:])], np.int32)
        assert_same_bits(got[-20000:], want, "PowOp<int>::apply")


@pytest.mark.parametrize("y", [2.0, 2.5, 0.5, -1.0, 3.0, -2.5, 1 / 3, 17.0, 100.5, -77.7, 1e-3, 1e5, 0.1, -0.0625])
def test_pow_f32_ulp_uniform(hc, orc, y):
    rng = np.random.default_rng(int(abs(y) * 1000) % 2**31)
    x = rng.uniform(0.01, 100, 1 << 20).astype(np.float32)
    err = oracle.ulp_error_f32(hc_pow32(hc, x, y), orc.pow_ref_f32(x, y))
    assert err.max() <= F32_POW_ULP_BOUND, (y, err.max(), x[err.argmax()])


@pytest.mark.parametrize("y", [2.0, 2.5, -0.75, 31.0, 1 / 3, -7.0])
def test_pow_f32_ulp_all_magnitudes(hc, orc, y):
    rng = np.random.default_rng(5)
    x = rng.integers(1, 0x7f800000, 1 << 20, dtype=np.uint32).view(np.float32)  # every positive finite incl. denormals
    err = oracle.ulp_error_f32(hc_pow32(hc, x, y), orc.pow_ref_f32(x, y))
    assert err.max() <= F32_POW_ULP_BOUND, (y, err.max(), x[err.argmax()])


@pytest.mark.parametrize("y", [2.0, 2.5, 0.5, -1.0, 3.0, -2.5, 1 / 3, 17.0, 1e-3, 0.1, -0.0625, 31.0, 8.0, -7.99, 8.000001, 5.0,
                               9.25, 100.5, -255.0, 256.0, 256.0001, -300.0])
def test_pow_f32_fast_core_ulp(hc, orc, y):
    """The table-driven FFMA2 core (what the GPU kernel runs for ordinary data):
    uniform values, every magnitude, signed bases, and all 128 table-entry edges."""
    rng = np.random.default_rng(int(abs(y) * 977) % 2**31)
    edges = []
    for j in range(128):
        for d in (-2, -1, 0, 1, 2):
            edges += [0x3F800000 + (j << 16) + d, 0x3F800000 + (j << 16) + 0xFFFF + d]  # table-entry edges
    xe = np.array(edges, dtype=np.uint32).view(np.float32)
    xe = np.concatenate([xe * np.float32(2.0 ** k) for k in (-20, -3, -1, 0, 1, 2, 7, 30)])
    x = np.concatenate([rng.uniform(0.01, 100, 1 << 20).astype(np.float32),
                        rng.integers(1, 0x7F800000, 1 << 20, dtype=np.uint32).view(np.float32),
                        -rng.uniform(0.01, 100, 1 << 18).astype(np.float32), xe])
    x = x[: x.size // 2 * 2]
    got, declined = hc_pow32_fast(hc, x, y)
    ref = orc.pow_ref_f32(x, float(np.float32(y)))
    err = oracle.ulp_error_f32(got, ref)
    assert err.max() <= F32_POW_ULP_BOUND, (y, err.max(), x[err.argmax()], got[err.argmax()])
    if y in (2.0, 2.5, 0.5, -1.0):  # ordinary data must stay on the fast core
        g2, d2 = hc_pow32_fast(hc, rng.uniform(0.01, 100, 1 << 16).astype(np.float32), y)
        assert d2 == 0.0


@pytest.mark.parametrize("y", [2.0, 2.5, 0.5, -1.0, 3.0, -3.0, 4.0, 1 / 3, 0.01, 17.0, -77.0, 255.5, 300.0, 1001.0])
def test_pow_f32_fast_core_runtime_sign_variant_ulp(hc, orc, y):
    """The variant op chains use (k_chain): sign masks and the double range test are run-time values.
    Signed bases: odd exponents keep the sign, even ones drop it, non-integer ones decline to NaN."""
    rng = np.random.default_rng(int(abs(y) * 313) % 2**31)
    lim = min(120.0 / abs(y), 20.0)
    x = np.concatenate([np.exp2(rng.uniform(-lim, lim, 1 << 20)), -np.exp2(rng.uniform(-lim, lim, 1 << 18)),
                        1 + rng.uniform(-2e-2, 2e-2, 1 << 18),
                        [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-42, -1e-42, 3e38, 1.0, -1.0]]).astype(np.float32)
    x = np.resize(x, x.size // 2 * 2)
    out = np.empty_like(x)
    dec = ctypes.c_uint64(0)
    hc.hc_pow_f32_fast_runtime(_p(x), ctypes.c_float(y), ctypes.c_uint64(x.size), _p(out), ctypes.byref(dec))
    ref = orc.pow_ref_f32(x, float(np.float32(y)))
    with np.errstate(all="ignore"):
        want = ref.astype(np.float32)
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(out), nan)
    fin = np.isfinite(want) & (want != 0) & ~nan
    assert oracle.ulp_error_f32(out[fin], ref[fin]).max() <= F32_POW_ULP_BOUND, y
    rest = ~fin & ~nan
    assert np.array_equal(out[rest], want[rest]) and np.array_equal(np.signbit(out[rest]), np.signbit(want[rest]))
    assert dec.value < 0.2 * x.size      # ordinary data stays on the fast core (negative bases decline when y is not an integer)


def test_pow_f32_fast_core_near_one_huge_exponents(hc, orc):
    rng = np.random.default_rng(16)
    x = (1 + rng.uniform(-2e-2, 2e-2, 1 << 20)).astype(np.float32)
    for y in (1e4, -3e4, 12345.678, 3000.0, -5000.5, 700.25):
        got, _ = hc_pow32_fast(hc, x, y)
        err = oracle.ulp_error_f32(got, orc.pow_ref_f32(x, float(np.float32(y))))
        assert err.max() <= F32_POW_ULP_BOUND, (y, err.max(), x[err.argmax()])


def test_pow_f32_fast_core_special_values_fall_back(hc, orc):
    x = np.array(SPECIAL_X + [7.0], np.float32)
    x = np.resize(x, (x.size // 2) * 2)
    for y in SPECIAL_Y:
        got, _ = hc_pow32_fast(hc, x, np.float32(y))
        assert_same_bits(got, hc_pow32(hc, x, np.float32(y)), f"fast path vs reference path, y={y}") if not (
            np.isfinite(np.float32(y)) and y != 0) else None
        want64 = orc.pow_ref_f32(x, float(np.float32(y)))
        with np.errstate(all="ignore"):
            want = want64.astype(np.float32)
        for xi, g, w, w64 in zip(x, got, want, want64):
            if np.isnan(w):
                assert np.isnan(g), (xi, y, g)
            elif np.isinf(w) or w == 0 or np.isinf(g) or g == 0:
                ok = g == w and np.signbit(g) == np.signbit(w)
                assert ok or oracle.ulp_error_f32(np.array([g]), np.array([w64]))[0] <= 1, (xi, y, g, w)
            else:
                assert oracle.ulp_error_f32(np.array([g]), np.array([w64]))[0] <= F32_POW_ULP_BOUND, (xi, y, g, w)


def test_pow_f32_near_one_huge_exponents(hc, orc):
    rng = np.random.default_rng(6)
    x = (1 + rng.uniform(-1e-3, 1e-3, 1 << 19)).astype(np.float32)
    for y in (1e4, -3e4, 12345.678, 8.5e4):
        err = oracle.ulp_error_f32(hc_pow32(hc, x, y), orc.pow_ref_f32(x, y))
        assert err.max() <= F32_POW_ULP_BOUND, (y, err.max())


SPECIAL_X = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, -3.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 3.4e38, -3.4e38,
             1.17549435e-38, 0.99999994, 1.0000001]
SPECIAL_Y = [0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0, 0.5, -0.5, 2.5, -2.5, np.inf, -np.inf, np.nan, 1e30, -1e30,
             16777216.0, 16777217.0, 4294967296.0, 1e-30, 127.0, -149.0, 1 / 3]


def test_pow_f32_special_case_table(hc, orc):
    """C99 Annex F.10.4.4 through std::pow itself: every (x, y) special pair."""
    x = np.array(SPECIAL_X, np.float32)
    for y in SPECIAL_Y:
        got = hc_pow32(hc, x, np.float32(y))
        want = orc.pow_ref_f32(x, float(np.float32(y)))
        with np.errstate(all="ignore"):
            w32 = want.astype(np.float32)
        for xi, g, w, w64 in zip(x, got, w32, want):
            if np.isnan(w):
                assert np.isnan(g), (xi, y, g)
            elif np.isinf(w) or w == 0 or np.isinf(g) or g == 0:
                ok = (g == w and np.signbit(g) == np.signbit(w))
                assert ok or oracle.ulp_error_f32(np.array([g]), np.array([w64]))[0] <= 1, (xi, y, g, w)
            else:
                assert oracle.ulp_error_f32(np.array([g]), np.array([w64]))[0] <= F32_POW_ULP_BOUND, (xi, y, g, w)


def test_pow_f32_pairwise_matches_scalar_path(hc):
    rng = np.random.default_rng(8)
    x = rng.uniform(0.1, 10, 4096).astype(np.float32)
    y = np.full(4096, 2.5, np.float32)
    out = np.empty_like(x)
    hc.hc_pow_f32_pair(_p(x), _p(y), ctypes.c_uint64(x.size), _p(out))
    assert_same_bits(out, hc_pow32(hc, x, 2.5))


@pytest.mark.parametrize("y", [2.0, 2.5, 0.5, -1.0, 3.0, 1 / 3, 17.0, -77.7, 1e-3, 0.1, 1e5])
def test_pow_f64_ulp_uniform(hc, orc, y):
    rng = np.random.default_rng(int(abs(y) * 1000) % 2**31 + 1)
    x = rng.uniform(0.01, 100, 1 << 19)
    hi, lo = orc.pow_ref_f64(x, y)
    err = oracle.ulp_error_f64(hc_pow64(hc, x, y), hi, lo)
    assert err.max() <= F64_POW_ULP_BOUND, (y, err.max(), x[err.argmax()])


@pytest.mark.parametrize("y", [2.0, 2.5, -0.75, 0.01, 1 / 3, -0.001])
def test_pow_f64_ulp_all_magnitudes(hc, orc, y):
    rng = np.random.default_rng(9)
    x = rng.integers(1, 0x7ff0000000000000, 1 << 19, dtype=np.uint64).view(np.float64)
    hi, lo = orc.pow_ref_f64(x, y)
    err = oracle.ulp_error_f64(hc_pow64(hc, x, y), hi, lo)
    assert err.max() <= F64_POW_ULP_BOUND, (y, err.max(), x[err.argmax()])


def hc_pow64_fast(hc, x, y):
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    dec = ctypes.c_uint64(0)
    hc.hc_pow_f64_fast(_p(x), ctypes.c_double(y), ctypes.c_uint64(x.size), _p(out), ctypes.byref(dec))
    return out, dec.value / max(x.size, 1)


@pytest.mark.parametrize("y", [2.0, 2.5, 0.5, -1.0, 3.0, 1 / 3, 17.0, -77.7, 1e-3, 0.1, 1e4, -3e5])
def test_pow_f64_fast_core_ulp(hc, orc, y):
    """The table-driven f64 core (what the GPU kernel runs for ordinary data).  powl on x87 is
    itself ~0.3 ULP(double) off when |y log2 x| is in the hundreds, hence the slack for huge |y|;
    against mpmath the core measures <= 0.61 ULP there and <= 0.51 ULP elsewhere."""
    rng = np.random.default_rng(int(abs(y) * 911) % 2**31)
    if abs(y) > 1000:
        x = 1 + rng.uniform(-6e-3, 6e-3, 1 << 19)
    else:
        x = np.concatenate([rng.uniform(0.01, 100, 1 << 19), -rng.uniform(0.01, 100, 1 << 16),
                            rng.integers(0x0010000000000000, 0x7FF0000000000000, 1 << 19, dtype=np.uint64).view(np.float64)])
    got, declined = hc_pow64_fast(hc, x, y)
    hi, lo = orc.pow_ref_f64(x, y)
    err = oracle.ulp_error_f64(got, hi, lo)
    finite_normal = np.isfinite(hi) & (np.abs(hi) > 2.3e-308)
    assert err[finite_normal].max() <= F64_POW_ULP_BOUND, (y, err[finite_normal].max(), x[finite_normal][err[finite_normal].argmax()])
    assert err.max() <= F64_POW_ULP_BOUND + 0.01
    if y in (2.5, 0.5, 3.0):
        _, d2 = hc_pow64_fast(hc, rng.uniform(0.01, 100, 1 << 14), y)
        assert d2 == 0.0


def test_pow_f64_fast_core_specials_fall_back(hc, orc):
    x = np.array(SPECIAL_X + [5e-324, 1.7e308, 2.2250738585072014e-308, 7.0], np.float64)
    for y in SPECIAL_Y + [9007199254740992.0, 1075.0, -1075.0]:
        got, _ = hc_pow64_fast(hc, x, y)
        hi, lo = orc.pow_ref_f64(x, y)
        for xi, g, w in zip(x, got, hi):
            if np.isnan(w):
                assert np.isnan(g), (xi, y, g)
            elif np.isinf(w) or w == 0:
                assert g == w and np.signbit(g) == np.signbit(w), (xi, y, g, w)
            else:
                assert abs(g - w) <= 2 * np.spacing(abs(w)), (xi, y, g, w)


def test_pow_f64_special_case_table(hc, orc):
    x = np.array(SPECIAL_X + [5e-324, 1.7e308, 2.2250738585072014e-308], np.float64)
    for y in SPECIAL_Y + [9007199254740992.0, 9007199254740993.0, 1075.0, -1075.0]:
        got = hc_pow64(hc, x, y)
        hi, lo = orc.pow_ref_f64(x, y)
        for xi, g, w in zip(x, got, hi):
            if np.isnan(w):
                assert np.isnan(g), (xi, y, g)
            elif np.isinf(w) or w == 0:
                assert g == w and np.signbit(g) == np.signbit(w), (xi, y, g, w)
            else:
                assert abs(g - w) <= 2 * np.spacing(abs(w)), (xi, y, g, w)
