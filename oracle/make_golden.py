"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference
(oracle/_ref/libsmref.so, built from /root/reference by oracle/Makefile).

Run here (the container that has /root/reference):
    make -C oracle ref && python oracle/make_golden.py
The fixtures are committed; the GPU box only reads them.

Each case stores the operand buffers, the broadcast stride tables the
reference's own sm::broadcast produced, and the reference's output.  Inputs
avoid int32 /0 and INT_MIN/-1 (undefined in the reference, division.h:69).
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_v1.npz")


def rand(rng, dtype, shape, op, is_b=False):
    if dtype == np.int32:
        if op == "div" and is_b:
            v = rng.integers(1, 98, size=shape)
            return (v * rng.choice([-1, 1], size=shape)).astype(np.int32)
        if op == "pow":
            return (rng.integers(-6, 7, size=shape) if not is_b else rng.integers(-3, 12, size=shape)).astype(np.int32)
        if op == "mul":
            return rng.integers(-2**31, 2**31, size=shape, dtype=np.int64).astype(np.int32)
        return rng.integers(-2**30, 2**30, size=shape).astype(np.int32)
    if op == "pow":
        if is_b:
            return rng.uniform(-3, 3, size=shape).astype(dtype)
        return rng.uniform(0.05, 20, size=shape).astype(dtype)
    # full-range bit patterns incl. denormals, scaled set, and a few specials
    v = rng.standard_normal(size=shape).astype(dtype) * dtype(10.0) ** rng.integers(-20, 20, size=shape).astype(dtype)
    return v.astype(dtype)


def main():
    ref = oracle.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsmref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261018)
    cases = {}
    idx = 0

    def add_case(kind, op, a, sa, b, sb, shape, out, extra=None):
        nonlocal idx
        p = f"c{idx:03d}_"
        cases[p + "kind"] = np.array(kind)
        cases[p + "op"] = np.array(op)
        cases[p + "a"] = a
        cases[p + "b"] = b
        cases[p + "sa"] = np.array(sa, dtype=np.uint64)
        cases[p + "sb"] = np.array(sb, dtype=np.uint64)
        cases[p + "shape"] = np.array(shape, dtype=np.uint64)
        cases[p + "out"] = out
        if extra is not None:
            cases[p + "extra"] = np.array(extra)
        idx += 1

    # shapes: (shape_a, shape_b) pairs covering contiguous, row / column / outer
    # broadcasts, rank padding, the reference test's 4-D pattern (scaled down)
    pairs = [
        ((5,), (5,)), ((37,), (37,)), ((2, 3), (2, 3)), ((2, 2, 2), (2, 2, 2)),
        ((16, 24), (1, 24)), ((16, 24), (16, 1)), ((16, 1), (1, 24)), ((24,), (16, 24)),
        ((8, 1, 12), (1, 6, 12)), ((7, 9, 3), (1, 7, 1, 3)), ((3, 1, 4, 1, 5), (2, 1, 6, 1)),
        ((2, 3, 2, 3, 2, 3), (1, 3, 1, 3, 1, 3)), ((1, 1), (1, 1)), ((4, 1), (4, 5)),
    ]
    for dtype in (np.float32, np.float64, np.int32):
        for op in ("add", "sub", "mul", "div", "pow"):
            for (s1, s2) in pairs:
                a = rand(rng, dtype, s1, op)
                b = rand(rng, dtype, s2, op, is_b=True)
                if dtype == np.int32 and op == "div":
                    b = np.where(b == 0, 3, b).astype(np.int32)
                    a = np.where(a == np.int32(-2**31), 7, a).astype(np.int32)
                st1 = [s // a.itemsize for s in a.strides]
                st2 = [s // b.itemsize for s in b.strides]
                shape, n1, n2, _ = ref.broadcast(s1, st1, s2, st2)
                out = ref.elementwise(op, a, n1, b, n2, shape)
                add_case("elementwise", op, a, n1, b, n2, shape, out)
    # views: interior pointer + parent strides (the reference's one(0, SLICE_ALL) pattern), transposed operand
    for dtype in (np.float32, np.int32):
        for op in ("add", "mul"):
            big = rand(rng, dtype, (3, 6, 5, 3), op)
            small = rand(rng, dtype, (1, 6, 1, 3), op, is_b=True)
            view = big[1]  # interior pointer
            shape, n1, n2, _ = ref.broadcast(view.shape, [s // 4 for s in view.strides], small.shape,
                                             [s // 4 for s in small.strides])
            # the shim addresses operands from their buffer base: hand it the view's own buffer copy
            va = np.ascontiguousarray(view)
            out = ref.elementwise(op, va, n1, small, n2, shape)
            add_case("elementwise", op, va, n1, small, n2, shape, out)
            m = rand(rng, dtype, (6, 4), op)
            mt_strides = [1, 4]  # transpose() of a {6,4} array: shape {4,6}, strides {1,4}
            other = rand(rng, dtype, (4, 6), op, is_b=True)
            shape, n1, n2, _ = ref.broadcast((4, 6), mt_strides, (4, 6), [6, 1])
            out = ref.elementwise(op, m, n1, other, n2, shape)
            add_case("elementwise", op, m, n1, other, n2, shape, out)
    # array (op) scalar, sizes around the SIMD tail
    for dtype in (np.float32, np.float64, np.int32):
        for op in ("add", "sub", "mul", "div", "pow"):
            for n in (1, 7, 8, 13, 64, 1001):
                a = rand(rng, dtype, (n,), op)
                if dtype == np.int32:
                    v = {"div": -7, "pow": 3, "mul": 48271}.get(op, 123456789)
                    if op == "div":
                        a = np.where(a == np.int32(-2**31), 7, a).astype(np.int32)
                else:
                    v = {"pow": 2.5}.get(op, 1.7)
                out = ref.array_scalar(op, a, v)
                add_case("scalar", op, a, [1], np.array([v], dtype=dtype), [0], [n], out)
    # int pow: exponent sweep incl. negative, zero, overflowing (lane vs scalar tail semantics)
    bases = np.array([0, 1, -1, 2, -2, 3, -3, 5, -5, 7, 10, -10, 46340, 46341, -46341, 65536, 2**31 - 1, -2**31, 123, -77],
                     dtype=np.int32)
    for e in (0, 1, 2, 3, 4, 5, 7, 13, 20, 31, 32, 33, 62, -1, -2, -3, -31, 2**31 - 1, -2**31):
        for n in (8, 11, 16, 20):
            a = np.resize(bases, n).astype(np.int32)
            out = ref.array_scalar("pow", a, e)
            add_case("scalar", "pow", a, [1], np.array([e], dtype=np.int32), [0], [n], out)
    # element_wise_op<int, PowOp<int>> (README recipe): contiguous (lanes + tail) and strided (scalar)
    a = np.resize(bases, 27).astype(np.int32)
    b = np.resize(np.array([0, 1, 2, 3, 5, 9, 15, 31, -1, -2, 33], dtype=np.int32), 27)
    out = ref.elementwise("pow", a, [1], b, [1], [27])
    add_case("elementwise", "pow", a, [1], b, [1], [27], out)
    a2 = np.resize(bases, 20).astype(np.int32).reshape(4, 5)
    b2 = np.array([[2, 3, 5, 31, -1]], dtype=np.int32)
    shape, n1, n2, _ = ref.broadcast(a2.shape, [5, 1], b2.shape, [5, 1])
    out = ref.elementwise("pow", a2, n1, b2, n2, shape)
    add_case("elementwise", "pow", a2, n1, b2, n2, shape, out)
    # float specials through scalar Op::apply (+ - * /): NaN, inf, signed zero, denormals
    sp32 = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.1754942e-38, 3.4028235e38, -3.4028235e38, 1.0, -1.5],
                    dtype=np.float32)
    A, B = np.meshgrid(sp32, sp32, indexing="ij")
    A, B = np.ascontiguousarray(A.ravel()), np.ascontiguousarray(B.ravel())
    for op in ("add", "sub", "mul", "div"):
        out = ref.elementwise(op, A, [1], B, [1], [A.size])
        add_case("elementwise", op, A, [1], B, [1], [A.size], out)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **cases)
    print(f"wrote {OUT}: {idx} cases, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    main()
