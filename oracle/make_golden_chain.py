"""Generate tests/golden/golden_chain_v1.npz from the UNMODIFIED reference
(oracle/_ref/libsmref.so): operator chains evaluated the way a reference user writes them --
one sm::SMArray operator after another, each materialising its temporary
(include/SMArray.h:217-305; sm::pow include/UserFunctions.h:42-48) -- stored with their inputs.
tests/test_chain.py checks the fused smb_chain result against these outputs bit for bit
(int32 pow included: sm::pow's lane / scalar-tail split; float pow is not in the fixtures, the
reference's float sm::pow does not link).

Run here (the container that has /root/reference):
    make -C oracle ref && python oracle/make_golden_chain.py [v1|v2|all]
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
OUT = os.path.join(GOLDEN, "golden_chain_v1.npz")
OUT_V2 = os.path.join(GOLDEN, "golden_chain_v2.npz")


def main():
    ref = oracle.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsmref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261019)

    def rnd(dt, shape, divisor=False):
        if dt == np.int32:
            if divisor:
                return (rng.integers(1, 500, size=shape) * rng.choice([-1, 1], size=shape)).astype(np.int32)
            return rng.integers(-2**31, 2**31, size=shape, dtype=np.int64).astype(np.int32)
        if divisor:
            return (rng.uniform(0.25, 4, size=shape) * rng.choice([-1, 1], size=shape)).astype(dt)
        return (rng.standard_normal(shape) * dt(10.0) ** rng.integers(-6, 6, size=shape).astype(dt)).astype(dt)

    # (first leaf shape, [(op, leaf shape | scalar)]); "r" prefix = leaf on the left
    specs = [
        ((48, 64), [("add", (48, 64)), ("mul", (48, 64))]),
        ((48, 64), [("add", (1, 64)), ("mul", (48, 1)), ("sub", (64,)), ("div", "dv")]),
        ((48, 64), [("sub", (48, 64)), ("rsub", (48, 64)), ("mul", 3), ("rdiv", "dvfirst")]),
        ((6, 1, 16), [("mul", (1, 5, 16)), ("add", (6, 5, 1)), ("sub", (16,))]),
        ((1000,), [("add", (1000,)), ("mul", (1000,)), ("sub", (1000,)), ("add", (1000,)), ("mul", (1000,)), ("sub", (1000,)), ("add", (1000,))]),
        ((7, 13), [("add", (7, 13)), ("mul", (13,)), ("add", 2)]),
    ]
    cases, idx = {}, 0
    for dt in (np.float32, np.float64, np.int32):
        for first_shape, steps in specs:
            first = rnd(dt, first_shape)
            acc = first
            p = f"k{idx:03d}_"
            names = []
            for s, (op, leaf) in enumerate(steps):
                swap = op.startswith("r")
                base = op[1:] if swap else op
                if leaf == "dv":
                    leaf = rnd(dt, acc.shape, divisor=True)
                elif leaf == "dvfirst":          # leaf / acc: make acc a safe divisor first
                    leaf = rnd(dt, acc.shape)
                    acc = np.where(acc == 0, dt(3), acc).astype(dt)
                    if dt == np.int32:
                        acc = np.where(acc == -1, 5, acc).astype(dt)
                    cases[p + f"fix{s}"] = acc.copy()   # the test re-injects this intermediate
                elif isinstance(leaf, tuple):
                    leaf = rnd(dt, leaf, divisor=(base == "div" and not swap))
                if isinstance(leaf, np.ndarray):
                    cases[p + f"leaf{s}"] = leaf
                    acc = (ref.smarray_binary(base, leaf, acc) if swap else ref.smarray_binary(base, acc, leaf))[0]
                else:
                    cases[p + f"leaf{s}"] = np.array(leaf, dtype=dt)
                    assert not swap
                    acc = ref.smarray_scalar(base, acc, leaf)
                names.append(op)
            cases[p + "first"] = first
            cases[p + "ops"] = np.array(names)
            cases[p + "out"] = acc
            idx += 1
    # int32 sm::pow inside a chain: (a + b) ^ e - 1, sizes around the 8-lane boundary
    for n in (5, 8, 13, 1003):
        for e in (3, 2, 20, -2, 0):
            a = rng.integers(-6, 7, size=n).astype(np.int32)
            b = rng.integers(-3, 4, size=n).astype(np.int32)
            p = f"k{idx:03d}_"
            t = ref.smarray_binary("add", a, b)[0]
            t = ref.smarray_scalar("pow", t, e)
            t = ref.smarray_scalar("sub", t, 1)
            cases[p + "first"] = a
            cases[p + "leaf0"] = b
            cases[p + "leaf1"] = np.array(e, np.int32)
            cases[p + "leaf2"] = np.array(1, np.int32)
            cases[p + "ops"] = np.array(["add", "pow", "sub"])
            cases[p + "out"] = t
            idx += 1
    np.savez_compressed(OUT, **cases)
    print(f"wrote {OUT}: {idx} chains, {os.path.getsize(OUT)} bytes")


def main_v2():
    """golden_chain_v2.npz: int32 sm::pow applied to an intermediate SMALLER than the chain's output
    (a later leaf broadcasts it up).  The reference's lane / scalar-tail split (include/math/
    calculate.h:172-215 through include/UserFunctions.h:42-48) follows the flat index inside the
    temporary sm::pow materialises -- N elements -- not the M*N-element result; a fused kernel that
    derives the split from its output index gets the tail elements of every row wrong."""
    ref = oracle.reference()
    if ref is None:
        raise SystemExit("oracle/_ref/libsmref.so missing: run `make -C oracle ref` where /root/reference exists")
    rng = np.random.default_rng(20261101)
    cases, idx = {}, 0

    def emit(first, steps):
        nonlocal idx
        p = f"k{idx:03d}_"
        acc = first
        for s, (op, leaf) in enumerate(steps):
            if isinstance(leaf, np.ndarray):
                acc = ref.smarray_binary(op, acc, leaf)[0]
                cases[p + f"leaf{s}"] = leaf
            else:
                acc = ref.smarray_scalar(op, acc, leaf)
                cases[p + f"leaf{s}"] = np.array(leaf, np.int32)
        cases[p + "first"] = first
        cases[p + "ops"] = np.array([op for op, _ in steps])
        cases[p + "out"] = acc
        idx += 1

    def small(shape, lo=-6, hi=7):
        return rng.integers(lo, hi, size=shape).astype(np.int32)

    for m, n in ((3, 5), (4, 8), (3, 13), (2, 21), (3, 203)):
        for e in (3, 13, 31, 40, -1, -2, 0):
            emit(small((1, n)), [("pow", e), ("add", small((m, n), -1000, 1000))])            # pow(row, e) + B
            emit(small((1, n)), [("add", 1), ("pow", e), ("mul", small((m, 1), -9, 10))])      # pow(row + 1, e) * col
            emit(small((n,)), [("pow", e), ("sub", small((m, n), -1000, 1000))])               # 1-D first leaf
        # the temporary is already M x N: lanes run over the flat M*N index, across row boundaries
        emit(small((m, 1)), [("mul", small((1, n))), ("pow", 3), ("add", small((n,)))])
        emit(small((m, 1)), [("add", small((1, n))), ("pow", -2), ("mul", 7)])
    np.savez_compressed(OUT_V2, **cases)
    print(f"wrote {OUT_V2}: {idx} chains, {os.path.getsize(OUT_V2)} bytes")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "v1"):
        main()
    if which in ("all", "v2"):
        main_v2()
