"""Pin the oracle (oracle/oracle.c): against the committed golden vectors that
were produced by the unmodified reference, against the compiled reference live
(when oracle/_ref is present), and against the known answers the reference's
own tests hold (tests/add.cpp, subtract.cpp, multiply.cpp, division.cpp,
pow.cpp -- restated here as data, each with its file:line)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_same_bits


def test_oracle_matches_every_golden_vector(orc, golden):
    assert len(golden.cases) > 300
    for c in golden.cases:
        if c["kind"] == "elementwise":
            got = orc.elementwise(c["op"], c["a"], c["sa"], c["b"], c["sb"], c["shape"])
        else:
            got = orc.array_scalar(c["op"], c["a"], c["b"][0])
        assert_same_bits(got.reshape(c["out"].shape), c["out"], f"golden case {c['idx']} ({c['kind']} {c['op']})")


def test_reference_own_test_binaries_pass():
    """The reference's tests/*.cpp, compiled unmodified against the reference
    headers + gtest shim (oracle/Makefile), all pass: 32/32."""
    d = os.path.join(ROOT, "oracle", "_ref")
    bins = [os.path.join(d, f"ref_test_{t}") for t in ("add", "subtract", "multiply", "division", "pow")]
    if not all(os.path.exists(b) for b in bins):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    total = 0
    for b in bins:
        r = subprocess.run([b], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        total += int(r.stdout.strip().splitlines()[-1].split()[0])
    assert total == 32


# --- known answers held by the reference's tests, through the oracle ----------
def _bin(orc, op, a, b):
    return orc.binary(op, np.asarray(a), np.asarray(b))


def test_kat_add(orc):
    f = np.float32
    # tests/add.cpp:6-15 OneDimensionalAddition
    assert_same_bits(_bin(orc, "add", np.array([1, 2, 3, 4, 5], f), np.array([5, 4, 3, 2, 1], f)), np.full(5, 6, f))
    # tests/add.cpp:18-30 TwoDimensionalAddition
    assert_same_bits(_bin(orc, "add", np.array([[1, 2, 3], [4, 5, 6]], f), np.array([[6, 5, 4], [3, 2, 1]], f)),
                     np.full((2, 3), 7, f))
    # tests/add.cpp:32-44 (int)
    i = np.int32
    assert_same_bits(_bin(orc, "add", np.array([[1, 2, 3], [4, 5, 6]], i), np.array([[6, 5, 4], [3, 2, 1]], i)),
                     np.full((2, 3), 7, i))
    # tests/add.cpp:47-57 3-D double
    d = np.float64
    a = np.arange(1, 9, dtype=d).reshape(2, 2, 2)
    b = a[::-1, ::-1, ::-1].copy()
    assert_same_bits(_bin(orc, "add", a, b), a + b)
    # tests/add.cpp:97-105 AdditionWithZero
    a = np.array([[1, 2], [3, 4]], f)
    assert_same_bits(_bin(orc, "add", a, np.zeros((2, 2), f)), a)


@pytest.mark.parametrize("op,big_v,small_v,want", [
    ("add", 1, 3, 4),   # tests/add.cpp:59-92
    ("sub", 1, 1, 0),   # tests/subtract.cpp:60-80
    ("mul", 1, 2, 2),   # tests/multiply.cpp:60-80
    ("div", 4, 2, 2),   # tests/division.cpp:60-74
])
def test_kat_view_broadcast(orc, ref, op, big_v, small_v, want):
    """ones(32,224,224,3)(0, SLICE_ALL) (op) {1,224,1,3}: view with the parent's
    strides, rank padding, two broadcast axes, result {1,224,224,3}."""
    big = np.full((2, 224, 224, 3), big_v, np.float32)  # only slab 0 is read
    small = np.full((1, 224, 1, 3), small_v, np.float32)
    view = big[0]
    got = orc.binary(op, view, small)
    assert got.shape == (1, 224, 224, 3)
    assert_same_bits(got, np.full((1, 224, 224, 3), want, np.float32))
    if ref is not None:  # the reference's own view + operator machinery
        assert_same_bits(ref.view_broadcast_f32(op, big, small), got)


def test_kat_sub_mul_div(orc):
    f, i, d = np.float32, np.int32, np.float64
    # tests/subtract.cpp:5-57
    assert_same_bits(_bin(orc, "sub", np.array([5, 4, 3, 2, 1], f), np.array([1, 2, 3, 4, 5], f)), np.array([4, 2, 0, -2, -4], f))
    assert_same_bits(_bin(orc, "sub", np.array([[6, 5, 4], [3, 2, 1]], i), np.array([[1, 2, 3], [4, 5, 6]], i)),
                     np.array([[5, 3, 1], [-1, -3, -5]], i))
    # tests/multiply.cpp:5-57, :83-104 (x0, x1)
    assert_same_bits(_bin(orc, "mul", np.array([1, 2, 3, 4, 5], f), np.array([5, 4, 3, 2, 1], f)), np.array([5, 8, 9, 8, 5], f))
    a = np.array([[1, 2], [3, 4]], f)
    assert_same_bits(_bin(orc, "mul", a, np.zeros((2, 2), f)), np.zeros((2, 2), f))
    assert_same_bits(_bin(orc, "mul", a, np.ones((2, 2), f)), a)
    # tests/division.cpp:5-57 (int exact quotients), :77-96
    assert_same_bits(_bin(orc, "div", np.array([[6, 8, 10], [12, 14, 16]], i), np.array([[2, 2, 2], [2, 2, 2]], i)),
                     np.array([[3, 4, 5], [6, 7, 8]], i))
    a = np.array([[5, 10], [15, 20]], f)
    assert_same_bits(_bin(orc, "div", a, a), np.ones((2, 2), f))
    a3 = np.arange(1, 9, dtype=d).reshape(2, 2, 2)
    assert_same_bits(_bin(orc, "div", a3 * 2, a3), np.full((2, 2, 2), 2, d))


def test_kat_int_pow(orc):
    i = np.int32
    # tests/pow.cpp:4-8, :10-16, :18-27, :38-44
    assert orc.array_scalar("pow", np.array([2], i), 3)[0] == 8
    assert_same_bits(orc.array_scalar("pow", np.array([1, 2, 3], i), 2), np.array([1, 4, 9], i))
    assert_same_bits(orc.array_scalar("pow", np.array([[1, 2, 3], [4, 5, 6]], i), 2), np.array([[1, 4, 9], [16, 25, 36]], i))
    assert_same_bits(orc.array_scalar("pow", np.array([[1, 2, 3]], i), 3), np.array([[1, 8, 27]], i))
    # tests/pow.cpp:46-61: empty<int>(1000,1000,2) filled with 5, ^3 == 125
    a = np.full(2_000_000, 5, i)
    assert_same_bits(orc.array_scalar("pow", a, 3), np.full(2_000_000, 125, i))
    # tests/pow.cpp:62-99: +-5 alternating, ^3 and ^-2 -> 0
    a = np.where(np.arange(5000) % 2 == 0, 5, -5).astype(i)
    assert_same_bits(orc.array_scalar("pow", a, 3), (a.astype(np.int64) ** 3).astype(i))
    assert_same_bits(orc.array_scalar("pow", a, -2), np.zeros(5000, i))


# --- oracle vs the compiled reference, live ------------------------------------
SHAPES = [((1000,), (1000,)), ((33, 65), (1, 65)), ((33, 65), (33, 1)), ((12, 1, 40), (1, 9, 40)),
          ((5, 6, 7, 3), (1, 6, 1, 3)), ((2, 3, 4, 5, 6, 7), (2, 1, 4, 1, 6, 1)), ((150_000,), (150_000,)),
          ((400, 300), (300,))]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int32])
@pytest.mark.parametrize("op", ["add", "sub", "mul", "div", "pow"])
def test_oracle_vs_reference_random(orc, ref, dtype, op):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(hash((op, np.dtype(dtype).name)) % 2**32)
    for s1, s2 in SHAPES:
        if dtype == np.int32:
            lo, hi = (-2**31, 2**31) if op in ("add", "sub", "mul") else (-1000, 1001)
            a = rng.integers(lo, hi, size=s1, dtype=np.int64).astype(np.int32)
            if op == "div":
                b = (rng.integers(1, 98, size=s2) * rng.choice([-1, 1], size=s2)).astype(np.int32)
                a[a == -2**31] = 5
            elif op == "pow":
                a = rng.integers(-12, 13, size=s1).astype(np.int32)
                b = rng.integers(-4, 34, size=s2).astype(np.int32)
            else:
                b = rng.integers(lo, hi, size=s2, dtype=np.int64).astype(np.int32)
        elif op == "pow":
            a = rng.uniform(0.01, 50, size=s1).astype(dtype)
            b = rng.uniform(-4, 4, size=s2).astype(dtype)
        else:
            a = (rng.standard_normal(s1) * 10.0 ** rng.integers(-30, 30, size=s1)).astype(dtype)
            b = (rng.standard_normal(s2) * 10.0 ** rng.integers(-30, 30, size=s2)).astype(dtype)
        assert_same_bits(orc.binary(op, a, b), ref.binary(op, a, b), f"{op} {s1}x{s2}")
        v = b.ravel()[0]
        assert_same_bits(orc.array_scalar(op, a, v), ref.array_scalar(op, a, v), f"{op} scalar {s1}")
        sh, t1, t2, tot = orc.broadcast(s1, [1] * len(s1), s2, [2] * len(s2))
        assert (sh, t1, t2, tot) == ref.broadcast(s1, [1] * len(s1), s2, [2] * len(s2))


def test_oracle_broadcast_rejects_like_reference(orc, ref):
    with pytest.raises(RuntimeError, match="Cannot broadcast shapes"):
        orc.broadcast([2, 3], [3, 1], [4, 3], [3, 1])  # include/SMUtils.h:76-78
    if ref is not None:
        with pytest.raises(RuntimeError):
            ref.broadcast([2, 3], [3, 1], [4, 3], [3, 1])


def test_oracle_scalar_apply_vs_reference(orc, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(7)
    for _ in range(2000):
        a, b = int(rng.integers(-50, 51)), int(rng.integers(-5, 40))
        assert orc.scalar_apply("pow", np.int32, a, b) == ref.scalar_apply("pow", np.int32, a, b), (a, b)
    for op in ("add", "sub", "mul", "div"):
        for _ in range(500):
            a, b = float(rng.standard_normal()), float(rng.standard_normal())
            assert orc.scalar_apply(op, np.float32, a, b) == ref.scalar_apply(op, np.float32, a, b)


def test_fill_uniform_is_deterministic_and_in_range(orc):
    x = orc.fill_uniform_f32(0, 4096, 3, 0.01, 100.0)
    y = orc.fill_uniform_f32(1024, 1024, 3, 0.01, 100.0)
    assert np.array_equal(x[1024:2048], y)
    assert x.min() >= 0.01 and x.max() <= 100.0 and len(np.unique(x)) > 4000


def test_oracle_dot_product_vs_reference(orc, ref):
    """dot_product<T> (include/math/product.h), the row widened after the elementwise path:
    the restatement reproduces the reference's association order bit for bit."""
    if ref is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    for n in (1, 3, 7, 8, 9, 15, 16, 17, 1000, 100_003):
        for dt in (np.float32, np.float64, np.int32):
            if dt == np.int32:
                a = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(dt)
                b = rng.integers(-2**31, 2**31, n, dtype=np.int64).astype(dt)
            else:
                a, b = rng.standard_normal(n).astype(dt), rng.standard_normal(n).astype(dt)
            assert orc.dot(a, b).tobytes() == ref.dot(a, b).tobytes(), (n, dt)
    assert orc.dot(np.array([1, 2, 3], np.int32), np.array([4, 5, 6], np.int32)) == 32
