"""tools/bench_multidev.py -- the LIBRARY's own multi-GPU path (smb_set_devices): one
process, one host thread, managed arrays (the drop-in SMArray storage), every operator spread over
the device set by flat output range.  bench.py's headline shards across PROCESSES (one rank per GPU,
the driver's contract); this measures what a program gets that only writes `a + b`.

    python tools/bench_multidev.py [--devices all|0,1|0,0,0,0] [--reps 20] > gpurun_out/multidev.jsonl

For each config: one device vs the device set, in the synchronous mode (every call complete on
return) and with SMB_OPT_ASYNC (calls enqueued back to back, one smb_sync at the end), host
wall-clock around the calls after warm-up -- the first calls pay the page migration that partitions
the arrays, the timed ones run on resident shards.  Results are compared bit for bit between the
two device sets.
"""
import argparse
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import simplemath_b200 as smb

ap = argparse.ArgumentParser()
ap.add_argument("--devices", default="all")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--elems", type=int, default=1 << 30, help="C5 elements per array")
ap.add_argument("--skip", default="", help="comma-separated config names to skip")
args = ap.parse_args()
lib = smb.lib()
ndev = smb.device_count()
devset = list(range(ndev)) if args.devices == "all" else [int(v) for v in args.devices.split(",")]
PEAK = 6540.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


class M:
    def __init__(self, n, dtype=np.float32):
        self.n, self.dtype = n, np.dtype(dtype)
        self.ptr = lib.smb_alloc(n * self.dtype.itemsize, smb.MEM_MANAGED)
        assert self.ptr, lib.smb_last_error()

    def host(self):
        ct = {4: ctypes.c_uint32, 8: ctypes.c_uint64}[self.dtype.itemsize]
        return np.ctypeslib.as_array(ctypes.cast(self.ptr, ctypes.POINTER(ct)), shape=(self.n,))

    def free(self):
        lib.smb_free(self.ptr)


def checksum(m: M):
    """Order-independent and bit-sensitive, computed where the data lives: the wrapping int32 dot product of the
    result's bit patterns with themselves (reading 4 GiB of managed memory on the host would migrate it)."""
    smb.sync()
    words = m.n * (m.dtype.itemsize // 4)
    return int(smb.dot_ptr(smb.I32, m.ptr, m.ptr, words))


def timed(fn, reps, asyn):
    for _ in range(3):
        fn()
    smb.sync()
    smb.set_option(smb.OPT_ASYNC, 1 if asyn else 0)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    smb.sync()
    dt = (time.perf_counter() - t0) / reps * 1e3
    smb.set_option(smb.OPT_ASYNC, 0)
    return dt


def run(name, bytes_, fn, result: M, reps):
    row = {"config": name, "algorithmic_bytes": bytes_, "devices": devset}
    sums = {}
    for label, ds in (("one_device", []), ("device_set", devset)):
        smb.set_devices(ds)
        for asyn in (False, True):
            ms = timed(fn, reps, asyn)
            key = f"{label}_{'async' if asyn else 'sync'}"
            row[key + "_ms"] = ms
            row[key + "_gbs"] = bytes_ / ms / 1e6
        row[label + "_kernel"] = smb.last_kernel()
        sums[label] = checksum(result)
    smb.set_devices([])
    g = len(devset)
    row["speedup_async"] = row["one_device_async_ms"] / row["device_set_async_ms"]
    row["speedup_sync"] = row["one_device_sync_ms"] / row["device_set_sync_ms"]
    row["per_gpu_gbs_async"] = row["device_set_async_gbs"] / g
    row["per_gpu_frac_of_measured_peak"] = row["per_gpu_gbs_async"] / PEAK
    row["bit_identical"] = sums["one_device"] == sums["device_set"]
    print(json.dumps(row), flush=True)


skip = set(args.skip.split(",")) if args.skip else set()
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
n = args.elems

if "c5" not in skip:
    a, b, x, out = M(n), M(n), M(n), M(n)
    smb.set_devices(devset)   # born partitioned
    smb.fill_uniform_f32_ptr(a.ptr, 0, n, 1, -1.0, 1.0)
    smb.fill_uniform_f32_ptr(b.ptr, 0, n, 2, -1.0, 1.0)
    smb.fill_uniform_f32_ptr(x.ptr, 0, n, 3, 0.01, 100.0)
    smb.set_devices([])
    run("C5 add f32 2^30", 12 * n, lambda: smb.contiguous_ptr(smb.OP_ADD, smb.F32, a.ptr, b.ptr, out.ptr, n), out, args.reps)
    run("C5 pow(x,2.5) f32 2^30", 8 * n, lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.ptr, 2.5, n, out.ptr), out, args.reps)
    leaves = [(None, False, (a.ptr, [1])), ("add", False, (b.ptr, [1])), ("mul", False, (x.ptr, [1]))]
    run("chain (a+b)*x f32 2^30", 16 * n, lambda: smb.chain_ptr(smb.F32, leaves, [n], out.ptr), out, args.reps)
    t0 = time.perf_counter()
    for ds in ([], devset):
        smb.set_devices(ds)
        for _ in range(3):
            smb.dot_ptr(smb.F32, a.ptr, b.ptr, n)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            v = smb.dot_ptr(smb.F32, a.ptr, b.ptr, n)
        ms = (time.perf_counter() - t0) / args.reps * 1e3
        print(json.dumps({"config": "dot f32 2^30", "devices": ds or [0], "ms": ms, "gbs": 8 * n / ms / 1e6, "value": v}), flush=True)
    smb.set_devices([])
    for m_ in (a, b, x, out):
        m_.free()
    lib.smb_pool_trim()

if "c2" not in skip:
    R, C = 65536, 4096   # C2 x 16
    a, row, out = M(R * C), M(C), M(R * C)
    smb.fill_uniform_f32_ptr(a.ptr, 0, R * C, 1, -1.0, 1.0)
    smb.fill_uniform_f32_ptr(row.ptr, 0, C, 2, -1.0, 1.0)
    shape, sa, sb, tot = smb.broadcast((R, C), (C, 1), (1, C), (C, 1))
    run("C2x16 {65536,4096}+{1,4096} f32", 4 * (2 * R * C + C),
        lambda: smb.elementwise_ptr(smb.OP_ADD, smb.F32, a.ptr, sa, row.ptr, sb, shape, out.ptr), out, args.reps)
    for m_ in (a, row, out):
        m_.free()
    lib.smb_pool_trim()

if "c4" not in skip:
    D0, D1, L = 512, 512, 1024
    ia, ib, out = M(D0 * L, np.int32), M(D1 * L, np.int32), M(D0 * D1 * L, np.int32)
    rng = np.random.default_rng(4)
    ia.host().view(np.int32)[:] = rng.integers(-1000, 1001, D0 * L).astype(np.int32)
    ib.host().view(np.int32)[:] = rng.integers(1, 98, D1 * L).astype(np.int32)
    lib.smb_host_written(ia.ptr); lib.smb_host_written(ib.ptr)
    shape, sa, sb, tot = smb.broadcast((D0, 1, L), (L, L, 1), (1, D1, L), (D1 * L, L, 1))
    for name, op in (("mul", smb.OP_MUL), ("div", smb.OP_DIV)):
        run(f"C4 int32 {name} {{512,1,1024}}x{{1,512,1024}}", 4 * (tot + D0 * L + D1 * L),
            lambda: smb.elementwise_ptr(op, smb.I32, ia.ptr, sa, ib.ptr, sb, shape, out.ptr), out, args.reps)
    for m_ in (ia, ib, out):
        m_.free()
