"""tools/bench_configs.py -- BASELINE.json configs C1-C4 (and f64 / specialised pow),
GPU (device-resident operands, CUDA events) beside the compiled reference on the
host cores.  Not the headline (bench.py is); this fills the per-config table in
profiles/ and README.md.

    python tools/bench_configs.py [--reps 20] [--no-cpu] > gpurun_out/configs.jsonl
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import simplemath_b200 as smb

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--no-cpu", action="store_true")
args = ap.parse_args()
PEAK = 6540.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
# an explicit NON-default stream: handle 0 (the legacy default stream) means "synchronous call on the library's
# private stream" to the C ABI, which would time a host round trip per call instead of the kernel
stream = torch.cuda.Stream()
sp = stream.cuda_stream
assert sp != 0
TD = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32}


def timed(fn, reps):
    torch.cuda.synchronize()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def cpu_time(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best * 1e3


def report(name, bytes_, n, ms, cpu_ms=None, note=""):
    row = {"config": name, "elements": n, "algorithmic_bytes": bytes_, "gpu_ms": ms, "gpu_gbs": bytes_ / ms / 1e6,
           "gpu_gelem_s": n / ms / 1e6, "frac_measured_peak": bytes_ / ms / 1e6 / PEAK, "frac_nominal_8000": bytes_ / ms / 1e6 / 8000,
           "kernel": smb.last_kernel(), "note": note}
    if cpu_ms is not None:
        row.update(cpu_ms=cpu_ms, cpu_gbs=bytes_ / cpu_ms / 1e6, speedup_vs_cpu=cpu_ms / ms)
    print(json.dumps(row), flush=True)


ref = None
if not args.no_cpu:
    import oracle
    ref = oracle.reference()
    if ref is not None:
        ref.h.smref_set_threads(os.cpu_count() or 1)


def binary_case(name, op, a, b, note=""):
    dt = smb.dtype_code(a.dtype)
    shape, sa, sb, n = smb.broadcast(a.shape, smb.row_major_strides(a.shape), b.shape, smb.row_major_strides(b.shape))
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    out = torch.empty(n, dtype=TD[a.dtype.type], device="cuda")
    # pre-marshalled arguments: the small configs are launch-latency bound, keep Python out of the loop
    lib, u = smb.lib(), smb._u64arr
    argv = (smb.OPS[op], dt, da.data_ptr(), u(sa), db.data_ptr(), u(sb), u(shape), len(shape), n, out.data_ptr(), sp)
    fn = lambda: lib.smb_elementwise(*argv)
    ms = timed(fn, args.reps)
    bytes_ = a.itemsize * (n + a.size + b.size)
    cpu_ms = cpu_time(lambda: ref.smarray_binary(op, a, b, want_result=False)) if ref is not None else None
    if bytes_ < (256 << 20):
        # launch-bound sizes, per CALL through the operator-style ABI (stream == NULL), host wall clock:
        # the reference's synchronous contract vs the asynchronous hand-off (SMB_OPT_ASYNC, one sync at the end)
        argn = argv[:-1] + (None,)
        k = max(200, args.reps)
        for mode in (0, 1):
            smb.set_option(smb.OPT_ASYNC, mode)
            for _ in range(20):
                lib.smb_elementwise(*argn)
            smb.sync()
            t0 = time.perf_counter()
            for _ in range(k):
                lib.smb_elementwise(*argn)
            smb.sync()
            us = (time.perf_counter() - t0) / k * 1e6
            note += f"; stream==NULL {'async hand-off' if mode else 'synchronous'}: {us:.2f} us per call = {bytes_ / us / 1e3:.0f} GB/s"
        smb.set_option(smb.OPT_ASYNC, 0)
    if bytes_ < (256 << 20):
        # launch-bound sizes: the same launches captured in a CUDA graph (no host gaps between them)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gsp = torch.cuda.current_stream().cuda_stream
                for _ in range(args.reps):
                    lib.smb_elementwise(*argv[:-1], gsp)
            # events on the stream the graph replays on (torch's current stream), not on the explicit stream `timed` uses
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(5):
                g.replay()
            g1.record()
            g1.synchronize()
            gms = g0.elapsed_time(g1) / 5 / args.reps
            note += f"; CUDA-graph replay of {args.reps} launches: {gms:.4f} ms each = {bytes_ / gms / 1e6:.0f} GB/s"
        except Exception as e:  # capture is an extra, never the measurement
            note += f"; graph capture unavailable: {type(e).__name__}"
    report(name, bytes_, n, ms, cpu_ms, note)


def scalar_case(name, op, a, v, spec=1, note=""):
    dt = smb.dtype_code(a.dtype)
    n = a.size
    da = torch.from_numpy(a).cuda()
    out = torch.empty(n, dtype=TD[a.dtype.type], device="cuda")
    smb.set_option(smb.OPT_POW_SPECIALISE, spec)
    import ctypes
    val = smb._CT[dt](v)
    lib = smb.lib()
    argv = (smb.OPS[op], dt, da.data_ptr(), ctypes.byref(val), n, out.data_ptr(), sp)
    fn = lambda: lib.smb_array_scalar(*argv)
    ms = timed(fn, args.reps)
    smb.set_option(smb.OPT_POW_SPECIALISE, 1)
    cpu_ms = cpu_time(lambda: ref.smarray_scalar(op, a, v, want_result=False), reps=2) if ref is not None else None
    report(name, 2 * a.itemsize * n, n, ms, cpu_ms, note)


rng = np.random.default_rng(1)
# C1: benchmark_add million_check
a = rng.uniform(-1, 1, 1_000_000).astype(np.float32)
b = rng.uniform(-1, 1, 1_000_000).astype(np.float32)
binary_case("C1 f32 1M contiguous a+b (million_check)", "add", a, b, "12 MB: launch-latency bound, L2 resident")
# C2: row broadcast
a = rng.uniform(-1, 1, (4096, 4096)).astype(np.float32)
b = rng.uniform(-1, 1, (1, 4096)).astype(np.float32)
binary_case("C2 f32 {4096,4096}+{1,4096}", "add", a, b, "134 MB: about the size of L2")
a = rng.uniform(-1, 1, (65536, 4096)).astype(np.float32)
binary_case("C2x16 f32 {65536,4096}+{1,4096}", "add", a, b, "1 GiB, not L2 resident")
del a
# C3: pow on 256M elements
x = (np.add.outer(np.arange(16384, dtype=np.float32), np.arange(16384, dtype=np.float32)) + 1).ravel()  # arr(i,j)=i+j+1
scalar_case("C3 f32 pow(arr,2.0) 256M, i+j+1 fill, specialised x*x", "pow", x, 2.0, 1)
scalar_case("C3 f32 pow(arr,2.0) 256M, i+j+1 fill, general kernel", "pow", x, 2.0, 0)
x = rng.uniform(0.01, 100, 1 << 28).astype(np.float32)
scalar_case("C3 f32 pow(arr,2.5) 256M uniform(0.01,100), general kernel", "pow", x, 2.5, 0)
scalar_case("C3 f32 pow(arr,17.5) 256M uniform, medium-|y| tier (8 < |y| <= 256)", "pow", x, 17.5, 0, "many results overflow -> slow path share")
scalar_case("C3 f32 pow(arr,9.25) 256M uniform, medium-|y| tier", "pow", x, 9.25, 0)
x1 = (1 + rng.uniform(-2e-2, 2e-2, 1 << 28)).astype(np.float32)
scalar_case("C3 f32 pow(arr,1000.5) 256M in [0.98,1.02], large-|y| tier (p-series)", "pow", x1, 1000.5, 0)
del x1
xd = x.astype(np.float64)   # C3 in double at the stated size: 268 435 456 elements, 2 GiB in + 2 GiB out
del x
scalar_case("C3 f64 pow(arr,2.5) 256M uniform, general kernel (FP64-pipe bound)", "pow", xd, 2.5, 0)
scalar_case("C3 f64 pow(arr,2.0) 256M general kernel (the reference benchmark's exponent)", "pow", xd, 2.0, 0)
scalar_case("C3 f64 pow(arr,2.0) 256M specialised", "pow", xd, 2.0, 1)
del xd
# C4: int32 3-D broadcast
ia = rng.integers(-1000, 1001, size=(512, 1, 1024)).astype(np.int32)
ib = rng.integers(1, 98, size=(1, 512, 1024)).astype(np.int32)
binary_case("C4 i32 {512,1,1024}*{1,512,1024}", "mul", ia, ib, "write-dominated: 1 GiB out, 4 MiB in")
binary_case("C4 i32 {512,1,1024}/{1,512,1024}", "div", ia, ib, "integer division is instruction-bound")

# dense int32 division (no operand reuse: 12 B/elem), and by a scalar (8 B/elem)
ia = rng.integers(-2**31, 2**31, size=1 << 28, dtype=np.int64).astype(np.int32)
ib = rng.integers(1, 1 << 20, size=1 << 28).astype(np.int32)
binary_case("I1 i32 contiguous a/b, 2^28 elements", "div", ia, ib)
scalar_case("I2 i32 a/7, 2^28 elements", "div", ia, 7)
del ia, ib

# strided operand: a.transpose() + b (generic element strides; no reference test covers it)
def transposed_case(name, n0, n1):
    a = torch.rand(n1 * n0, device="cuda")  # a is {n1, n0} dense; a^T has shape {n0, n1}, strides {1, n0}
    b = torch.rand(n0 * n1, device="cuda")
    out = torch.empty(n0 * n1, device="cuda")
    lib, u = smb.lib(), smb._u64arr
    argv = (smb.OP_ADD, smb.F32, a.data_ptr(), u([1, n0]), b.data_ptr(), u([n1, 1]), u([n0, n1]), 2, n0 * n1, out.data_ptr(), sp)
    ms = timed(lambda: lib.smb_elementwise(*argv), args.reps)
    ok = bool(torch.equal(out.view(n0, n1), a.view(n1, n0).t() + b.view(n0, n1)))
    cpu_ms = None
    if ref is not None:
        ha, hb = a.cpu().numpy(), b.cpu().numpy()
        cpu_ms = cpu_time(lambda: ref.elementwise("add", ha, [1, n0], hb, [n1, 1], [n0, n1]), reps=1)
    report(name, 12 * n0 * n1, n0 * n1, ms, cpu_ms, f"verified={ok}")


def generic_case(name, rows, cols):
    """w[:, ::2] + w[:, 1::2]: inner stride 2 on both operands -- neither {0,1}-inner nor a transpose: k_generic."""
    w = torch.rand(rows * cols, device="cuda")
    half = cols // 2
    out = torch.empty(rows * half, device="cuda")
    lib, u = smb.lib(), smb._u64arr
    argv = (smb.OP_ADD, smb.F32, w.data_ptr(), u([cols, 2]), w.data_ptr() + 4, u([cols, 2]), u([rows, half]), 2, rows * half, out.data_ptr(), sp)
    ms = timed(lambda: lib.smb_elementwise(*argv), args.reps)
    wv = w.view(rows, cols)
    ok = bool(torch.equal(out.view(rows, half), wv[:, ::2] + wv[:, 1::2]))
    # algorithmic bytes: both operands' distinct elements (together: every element of w once) + the result
    report(name, 4 * (rows * cols + rows * half), rows * half, ms, None, f"verified={ok}; the two operands interleave in the same cache lines")


generic_case("G f32 w[:, ::2] + w[:, 1::2], w {16384,16384} (generic strides)", 16384, 16384)
transposed_case("T f32 {8192,8192}^T + {8192,8192} (transposed operand)", 8192, 8192)
transposed_case("T f32 {16384,1000}^T + {1000,16384}... shape {1000,16384}", 1000, 16384)

# §8(f) row 2: dot product (SMArray::operator%), HBM-bound reduction: 2*sizeof(T) bytes per element
for npdt, tdt, n in ((np.float32, torch.float32, 1 << 28), (np.float64, torch.float64, 1 << 27), (np.int32, torch.int32, 1 << 28)):
    xa = (torch.rand(n, device="cuda") * 2 - 1).to(tdt) if npdt != np.int32 else torch.randint(-1000, 1000, (n,), dtype=tdt, device="cuda")
    xb = xa.flip(0).contiguous()
    dt = smb.dtype_code(npdt)
    ms = timed(lambda: smb.dot_ptr(dt, xa.data_ptr(), xb.data_ptr(), n, sp), args.reps)
    cpu_ms = None
    if ref is not None:
        ha, hb = xa[: 1 << 24].cpu().numpy(), xb[: 1 << 24].cpu().numpy()
        cpu_ms = cpu_time(lambda: ref.dot(ha, hb)) * (n / (1 << 24))
    report(f"dot {np.dtype(npdt).name} {n} elements (operator%)", 2 * np.dtype(npdt).itemsize * n, n, ms, cpu_ms,
           "synchronous scalar result; CPU time scaled from a 2^24-element sample (single-threaded SIMD loop)")
    del xa, xb

# §8(f) row 1: op-chain fusion.  One pass (smb_chain) beside the same work as separate operators
# through this library (each materialising its temporary, as the reference's operators do).
def chain_case(name, n, leaves_fn, unfused_fn, arrays_in, note=""):
    ms = timed(leaves_fn, args.reps)
    kern = smb.last_kernel()
    ms_unfused = timed(unfused_fn, args.reps)
    bytes_ = 4 * n * (arrays_in + 1)
    row = {"config": name, "elements": n, "algorithmic_bytes": bytes_, "gpu_ms": ms, "gpu_gbs": bytes_ / ms / 1e6,
           "gpu_gelem_s": n / ms / 1e6, "frac_measured_peak": bytes_ / ms / 1e6 / PEAK, "frac_nominal_8000": bytes_ / ms / 1e6 / 8000,
           "kernel": kern, "unfused_ms": ms_unfused, "fusion_speedup": ms_unfused / ms, "note": note}
    print(json.dumps(row), flush=True)


n = 1 << 28
fa, fb, fc, fd = (torch.rand(n, device="cuda") + 0.5 for _ in range(4))
frow = torch.rand(16384, device="cuda")
fo, ft1, ft2 = (torch.empty(n, device="cuda") for _ in range(3))
P = lambda t: t.data_ptr()
L = lambda op, t, st=(1,): (op, False, (P(t), list(st)))
steps3 = smb.chain_steps(smb.F32, [L(None, fa), L("add", fb), L("mul", fc)], [n])
chain_case("F1 f32 (a+b)*c, 2^28 elements, fused", n,
           lambda: smb.lib().smb_chain(smb.F32, steps3, 3, smb._u64arr([n]), 1, n, P(fo), sp),
           lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(fa), P(fb), P(ft1), n, sp),
                    smb.contiguous_ptr(smb.OP_MUL, smb.F32, P(ft1), P(fc), P(fo), n, sp)), 3,
           "unfused = 2 operators, 24 B/elem of traffic; fused 16 B/elem")
steps5 = smb.chain_steps(smb.F32, [L(None, fa), L("add", fb), L("mul", fc), L("sub", fd), ("div", False, 3.0)], [n])
chain_case("F2 f32 ((a+b)*c-d)/3, 2^28 elements, fused", n,
           lambda: smb.lib().smb_chain(smb.F32, steps5, 5, smb._u64arr([n]), 1, n, P(fo), sp),
           lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(fa), P(fb), P(ft1), n, sp),
                    smb.contiguous_ptr(smb.OP_MUL, smb.F32, P(ft1), P(fc), P(ft2), n, sp),
                    smb.contiguous_ptr(smb.OP_SUB, smb.F32, P(ft2), P(fd), P(ft1), n, sp),
                    smb.array_scalar_ptr(smb.OP_DIV, smb.F32, P(ft1), 3.0, n, P(fo), sp)), 4,
           "unfused = 4 operators, 44 B/elem; fused 20 B/elem")
sh = [16384, 16384]
stepsb = smb.chain_steps(smb.F32, [(None, False, (P(fa), [16384, 1])), ("mul", False, (P(frow), [0, 1])), ("add", False, (P(fb), [16384, 1]))], sh)
chain_case("F3 f32 a*row+b, {16384,16384} with a {1,16384} row, fused", n,
           lambda: smb.lib().smb_chain(smb.F32, stepsb, 3, smb._u64arr(sh), 2, n, P(fo), sp),
           lambda: (smb.elementwise_ptr(smb.OP_MUL, smb.F32, P(fa), [16384, 1], P(frow), [0, 1], sh, P(ft1), sp),
                    smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(ft1), P(fb), P(fo), n, sp)), 2,
           "broadcast leaf inside the chain")
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
stepsp = smb.chain_steps(smb.F32, [L(None, fa), L("add", fb), ("pow", False, 2.5)], [n])
chain_case("F4 f32 pow(a+b, 2.5), 2^28 elements, fused", n,
           lambda: smb.lib().smb_chain(smb.F32, stepsp, 3, smb._u64arr([n]), 1, n, P(fo), sp),
           lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(fa), P(fb), P(ft1), n, sp),
                    smb.array_scalar_ptr(smb.OP_POW, smb.F32, P(ft1), 2.5, n, P(fo), sp)), 2,
           "sm::pow(a+b, e) in one pass")
stepsp1 = smb.chain_steps(smb.F32, [L(None, fa), ("add", False, 1.5), ("pow", False, 2.5)], [n])
chain_case("F5 f32 pow(a + 1.5, 2.5), 2^28 elements, fused", n,
           lambda: smb.lib().smb_chain(smb.F32, stepsp1, 3, smb._u64arr([n]), 1, n, P(fo), sp),
           lambda: (smb.array_scalar_ptr(smb.OP_ADD, smb.F32, P(fa), 1.5, n, P(ft1), sp),
                    smb.array_scalar_ptr(smb.OP_POW, smb.F32, P(ft1), 2.5, n, P(fo), sp)), 1,
           "sm::pow(a + c, e) in one pass: the pow kernel with a one-operand pre-operator")
stepsp2 = smb.chain_steps(smb.F32, [L(None, fa), L("add", fb), ("pow", False, 2.5), L("mul", fc)], [n])
chain_case("F6 f32 pow(a+b, 2.5)*c, 2^28 elements, fused (general chain kernel)", n,
           lambda: smb.lib().smb_chain(smb.F32, stepsp2, 4, smb._u64arr([n]), 1, n, P(fo), sp),
           lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(fa), P(fb), P(ft1), n, sp),
                    smb.array_scalar_ptr(smb.OP_POW, smb.F32, P(ft1), 2.5, n, P(ft2), sp),
                    smb.contiguous_ptr(smb.OP_MUL, smb.F32, P(ft2), P(fc), P(fo), n, sp)), 3,
           "a pow step in the MIDDLE of a chain stays on k_chain (instruction-bound)")
n64 = 1 << 27
da, db = (torch.rand(n64, device="cuda", dtype=torch.float64) + 0.5 for _ in range(2))
do, dt1 = (torch.empty(n64, device="cuda", dtype=torch.float64) for _ in range(2))
steps64 = smb.chain_steps(smb.F64, [L(None, da), L("add", db), ("pow", False, 2.5)], [n64])


def chain_case64(name, leaves_fn, unfused_fn, arrays_in, note):
    ms = timed(leaves_fn, args.reps)
    kern = smb.last_kernel()
    ms_unfused = timed(unfused_fn, args.reps)
    bytes_ = 8 * n64 * (arrays_in + 1)
    print(json.dumps({"config": name, "elements": n64, "algorithmic_bytes": bytes_, "gpu_ms": ms, "gpu_gbs": bytes_ / ms / 1e6,
                      "gpu_gelem_s": n64 / ms / 1e6, "frac_measured_peak": bytes_ / ms / 1e6 / PEAK, "frac_nominal_8000": bytes_ / ms / 1e6 / 8000,
                      "kernel": kern, "unfused_ms": ms_unfused, "fusion_speedup": ms_unfused / ms, "note": note}), flush=True)


chain_case64("F7 f64 pow(a+b, 2.5), 2^27 elements, fused",
             lambda: smb.lib().smb_chain(smb.F64, steps64, 3, smb._u64arr([n64]), 1, n64, P(do), sp),
             lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F64, P(da), P(db), P(dt1), n64, sp),
                      smb.array_scalar_ptr(smb.OP_POW, smb.F64, P(dt1), 2.5, n64, P(do), sp)), 2,
             "the f64 pow kernel with a pre-operator (a double chain with a pow step would otherwise run the double-double path per element)")
smb.set_option(smb.OPT_POW_SPECIALISE, 1)
