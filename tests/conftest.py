"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only build container,
`-m gpu` on a B200 box (where /root/reference does not exist)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        import simplemath_b200 as smb
        return smb.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle.c_oracle()


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libsmref.so) or None."""
    import oracle
    return oracle.reference()


class Golden:
    def __init__(self, path):
        z = np.load(path)
        self.cases = []
        i = 0
        while f"c{i:03d}_kind" in z:
            p = f"c{i:03d}_"
            self.cases.append(dict(kind=str(z[p + "kind"]), op=str(z[p + "op"]), a=z[p + "a"], b=z[p + "b"],
                                   sa=[int(x) for x in z[p + "sa"]], sb=[int(x) for x in z[p + "sb"]],
                                   shape=[int(x) for x in z[p + "shape"]], out=z[p + "out"], idx=i))
            i += 1


@pytest.fixture(scope="session")
def golden():
    return Golden(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))


def assert_same_bits(got: np.ndarray, want: np.ndarray, what=""):
    """Bit-exact, except that any NaN matches any NaN (x86 and PTX differ in NaN
    payload / sign)."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == want.dtype, (what, got.dtype, want.dtype)
    if got.dtype.kind == "f":
        nan = np.isnan(want)
        assert np.array_equal(np.isnan(got), nan), f"{what}: NaN positions differ"
        iv = {4: np.uint32, 8: np.uint64}[got.dtype.itemsize]
        g, w = got.view(iv)[~nan.ravel().reshape(got.shape)], want.view(iv)[~nan]
        bad = np.nonzero(g != w)[0]
        assert bad.size == 0, f"{what}: {bad.size} of {got.size} differ, first got={got[~nan][bad[0]]!r} want={want[~nan][bad[0]]!r}"
    else:
        bad = np.nonzero(got.ravel() != want.ravel())[0]
        assert bad.size == 0, f"{what}: {bad.size} of {got.size} differ, first idx {bad[0]} got={got.ravel()[bad[0]]} want={want.ravel()[bad[0]]}"
