"""oracle -- ctypes loaders for the CHECKERS.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this package; nothing under
simplemath_b200/ or include/ does.

  oracle.c_oracle()  -> oracle/liboracle.so      plain-C restatement (oracle.c)
  oracle.reference() -> oracle/_ref/libsmref.so  the unmodified reference headers
                                                  behind a C ABI (ref_shim.cpp);
                                                  None when it was never built
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libsmref.so")

OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW = range(5)
F32, F64, I32 = range(3)
OPS = {"add": OP_ADD, "sub": OP_SUB, "mul": OP_MUL, "div": OP_DIV, "pow": OP_POW}
NP_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.int32): I32}
CT = {F32: ctypes.c_float, F64: ctypes.c_double, I32: ctypes.c_int32}

_u64, _vp, _i = ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(ref: bool = True) -> None:
    """make -C oracle (the recipe is the committed Makefile)."""
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True, env=env)


def _u64arr(v):
    return (ctypes.c_uint64 * len(v))(*[int(x) for x in v])


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


class _Lib:
    """Same Python surface for liboracle.so (prefix orc_) and libsmref.so
    (prefix smref_), so a test can run one body against either."""

    def __init__(self, path: str, prefix: str):
        self.path, self.prefix = path, prefix
        self.h = ctypes.CDLL(path)
        f = self._f
        f("elementwise", _i, [_i, _i, _vp, _u64p, _vp, _u64p, _u64p, _i, _u64, _vp])
        f("array_scalar", _i, [_i, _i, _vp, _vp, _u64, _vp])
        f("scalar_apply", _i, [_i, _i, _vp, _vp, _vp])
        f("broadcast", _i, [_u64p, _u64p, _i, _u64p, _u64p, _i, _u64p, _u64p, _u64p, ctypes.POINTER(_i), _u64p])
        f("is_contiguous", _i, [_u64p, _u64p, _i])

    def dot(self, a: np.ndarray, b: np.ndarray):
        """dot_product<T> with the reference's exact association order."""
        a, b = np.ascontiguousarray(a).ravel(), np.ascontiguousarray(b).ravel()
        name = {np.dtype(np.float32): "dot_f32", np.dtype(np.float64): "dot_f64", np.dtype(np.int32): "dot_i32"}[a.dtype]
        fn = getattr(self.h, self.prefix + name)
        fn.restype = CT[NP_DTYPES[a.dtype]]
        fn.argtypes = [_vp, _vp, _u64]
        return a.dtype.type(fn(_ptr(a), _ptr(b), a.size))

    def _f(self, name, res, args):
        fn = getattr(self.h, self.prefix + name)
        fn.restype, fn.argtypes = res, args
        setattr(self, "_" + name, fn)
        return fn

    # -- element_wise_op on explicit stride tables (elements) -----------------
    def elementwise(self, op, a: np.ndarray, stride_a, b: np.ndarray, stride_b, shape) -> np.ndarray:
        op = OPS.get(op, op)
        dt = NP_DTYPES[a.dtype]
        n = int(np.prod(shape, dtype=np.uint64)) if len(shape) else 1
        out = np.empty(n, dtype=a.dtype)
        rc = self._elementwise(op, dt, _ptr(a), _u64arr(stride_a), _ptr(b), _u64arr(stride_b), _u64arr(shape),
                               len(shape), n, _ptr(out))
        assert rc == 0, rc
        return out.reshape(shape)

    def array_scalar(self, op, a: np.ndarray, value) -> np.ndarray:
        op = OPS.get(op, op)
        dt = NP_DTYPES[a.dtype]
        a = np.ascontiguousarray(a)
        out = np.empty_like(a)
        v = CT[dt](value)
        rc = self._array_scalar(op, dt, _ptr(a), ctypes.byref(v), a.size, _ptr(out))
        assert rc == 0, rc
        return out

    def scalar_apply(self, op, dtype, a, b):
        op = OPS.get(op, op)
        dt = NP_DTYPES[np.dtype(dtype)]
        x, y, r = CT[dt](a), CT[dt](b), CT[dt]()
        rc = self._scalar_apply(op, dt, ctypes.byref(x), ctypes.byref(y), ctypes.byref(r))
        assert rc == 0, rc
        return r.value

    def broadcast(self, shape1, strides1, shape2, strides2):
        nd = max(len(shape1), len(shape2))
        rs, s1, s2 = _u64arr([0] * nd), _u64arr([0] * nd), _u64arr([0] * nd)
        ond, tot = _i(0), _u64(0)
        rc = self._broadcast(_u64arr(shape1), _u64arr(strides1), len(shape1), _u64arr(shape2), _u64arr(strides2),
                             len(shape2), rs, s1, s2, ctypes.byref(ond), ctypes.byref(tot))
        if rc:
            raise RuntimeError("Cannot broadcast shapes: incompatible dimensions")
        return list(rs), list(s1), list(s2), int(tot.value)

    def is_contiguous(self, shape, stride) -> bool:
        return bool(self._is_contiguous(_u64arr(shape), _u64arr(stride), len(shape)))

    # -- numpy-view convenience: broadcast + elementwise like SMArray operators
    def binary(self, op, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        es = a.dtype.itemsize
        sa = [s // es for s in a.strides]
        sb = [s // es for s in b.strides]
        shape, s1, s2, _ = self.broadcast(a.shape, sa, b.shape, sb)
        # pass base addresses of the views, not of their parents
        op = OPS.get(op, op)
        dt = NP_DTYPES[a.dtype]
        n = int(np.prod(shape, dtype=np.uint64))
        out = np.empty(n, dtype=a.dtype)
        rc = self._elementwise(op, dt, a.ctypes.data, _u64arr(s1), b.ctypes.data, _u64arr(s2), _u64arr(shape),
                               len(shape), n, _ptr(out))
        assert rc == 0, rc
        return out.reshape(shape)


class _Oracle(_Lib):
    def __init__(self):
        super().__init__(ORACLE_LIB, "orc_")
        h = self.h
        h.orc_pow_ref_f32.restype = None
        h.orc_pow_ref_f32.argtypes = [_vp, ctypes.c_float, _u64, _vp]
        h.orc_pow_ref_f64.restype = None
        h.orc_pow_ref_f64.argtypes = [_vp, ctypes.c_double, _u64, _vp, _vp]
        h.orc_fill_uniform_f32.restype = None
        h.orc_fill_uniform_f32.argtypes = [_vp, _u64, _u64, _u64, ctypes.c_float, ctypes.c_float]
        h.orc_powi_lane.restype = ctypes.c_int32
        h.orc_powi_lane.argtypes = [ctypes.c_int32, ctypes.c_int32]
        h.orc_powi_scalar.restype = ctypes.c_int32
        h.orc_powi_scalar.argtypes = [ctypes.c_int32, ctypes.c_int32]

    def pow_ref_f32(self, x: np.ndarray, y: float) -> np.ndarray:
        """std::pow in double on the exactly converted f32 inputs."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty(x.shape, dtype=np.float64)
        self.h.orc_pow_ref_f32(_ptr(x), y, x.size, _ptr(out))
        return out

    def pow_ref_f64(self, x: np.ndarray, y: float):
        """powl in long double, as (hi, lo) doubles."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        hi, lo = np.empty_like(x), np.empty_like(x)
        self.h.orc_pow_ref_f64(_ptr(x), y, x.size, _ptr(hi), _ptr(lo))
        return hi, lo

    def fill_uniform_f32(self, first: int, n: int, seed: int, lo: float, hi: float) -> np.ndarray:
        out = np.empty(n, dtype=np.float32)
        self.h.orc_fill_uniform_f32(_ptr(out), first, n, seed, lo, hi)
        return out


class _Reference(_Lib):
    def __init__(self):
        super().__init__(REF_LIB, "smref_")
        h = self.h
        h.smref_threads.restype = _i
        h.smref_set_threads.argtypes = [_i]
        h.smref_smarray_binary.restype = _i
        h.smref_smarray_binary.argtypes = [_i, _i, _vp, _u64p, _i, _vp, _u64p, _i, _vp, _u64p, ctypes.POINTER(_i)]
        h.smref_smarray_scalar.restype = _i
        h.smref_smarray_scalar.argtypes = [_i, _i, _vp, _u64p, _i, _vp, _vp]
        h.smref_view_broadcast_f32.restype = _i
        h.smref_view_broadcast_f32.argtypes = [_i, _vp, _u64, _u64, _u64, _u64, _vp, _vp]

    def threads(self) -> int:
        return int(self.h.smref_threads())

    def smarray_binary(self, op, a: np.ndarray, b: np.ndarray, want_result: bool = True):
        """sm::SMArray operator as shipped (broadcast + new[] + loop), dense operands."""
        op = OPS.get(op, op)
        dt = NP_DTYPES[a.dtype]
        a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
        oshape = _u64arr([0] * 8)
        ond = _i(0)
        out = None
        if want_result:
            out = np.empty(np.broadcast_shapes(a.shape, b.shape), dtype=a.dtype)
        rc = self.h.smref_smarray_binary(op, dt, _ptr(a), _u64arr(a.shape), a.ndim, _ptr(b), _u64arr(b.shape), b.ndim,
                                         _ptr(out) if want_result else None, oshape, ctypes.byref(ond))
        if rc == 2:
            raise RuntimeError("Cannot broadcast shapes: incompatible dimensions")
        assert rc == 0, rc
        return out, list(oshape[: ond.value])

    def smarray_scalar(self, op, a: np.ndarray, value, want_result: bool = True):
        """SMArray operator(T) / sm::pow as shipped (new[] + array_scalar_op)."""
        op = OPS.get(op, op)
        dt = NP_DTYPES[a.dtype]
        a = np.ascontiguousarray(a)
        v = CT[dt](value)
        out = np.empty_like(a) if want_result else None
        rc = self.h.smref_smarray_scalar(op, dt, _ptr(a), _u64arr(a.shape), a.ndim, ctypes.byref(v),
                                         _ptr(out) if want_result else None)
        assert rc == 0, rc
        return out

    def view_broadcast_f32(self, op, big: np.ndarray, small: np.ndarray) -> np.ndarray:
        op = OPS.get(op, op)
        d0, d1, d2, d3 = big.shape
        out = np.empty((1, d1, d2, d3), dtype=np.float32)
        rc = self.h.smref_view_broadcast_f32(op, _ptr(big), d0, d1, d2, d3, _ptr(small), _ptr(out))
        assert rc == 0, rc
        return out


_oracle = None
_reference = None


def c_oracle() -> _Oracle:
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_LIB):
            build(ref=False)
        _oracle = _Oracle()
    return _oracle


def reference():
    """The compiled reference, or None when oracle/_ref was never built (it is
    built in the container that has /root/reference and travels from there)."""
    global _reference
    if _reference is None and os.path.exists(REF_LIB):
        _reference = _Reference()
    return _reference


# ---------------------------------------------------------------- ULP tools --
def ulp_error_f32(got: np.ndarray, ref64: np.ndarray) -> np.ndarray:
    """|got - ref| in units of the f32 ulp at ref; inf/NaN/overflow positions
    count 0 when got matches the rounded reference, inf otherwise."""
    got = np.asarray(got, dtype=np.float32)
    with np.errstate(all="ignore"):
        ref32 = ref64.astype(np.float32)
        sp = np.spacing(np.abs(ref32)).astype(np.float64)
        sp = np.where(np.isfinite(sp) & (sp > 0), sp, 2.0 ** -149)
        err = np.abs(got.astype(np.float64) - ref64) / sp
        special = ~np.isfinite(ref64) | ~np.isfinite(got) | ~np.isfinite(ref32)
        same = (got == ref32) | (np.isnan(got) & np.isnan(ref32))
    return np.where(special, np.where(same, 0.0, np.inf), err)


def ulp_error_f64(got: np.ndarray, hi: np.ndarray, lo: np.ndarray) -> np.ndarray:
    with np.errstate(all="ignore"):
        sp = np.spacing(np.abs(hi))
        sp = np.where(np.isfinite(sp) & (sp > 0), sp, 2.0 ** -1074)
        err = np.abs((got - hi) - lo) / sp
        special = ~np.isfinite(hi) | ~np.isfinite(got)
        same = (got == hi) | (np.isnan(got) & np.isnan(hi))
    return np.where(special, np.where(same, 0.0, np.inf), err)
