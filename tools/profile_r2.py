"""tools/profile_r2.py -- launches the second-tier kernels a couple of times each for ncu captures (round 2):
k_tile on the skinny and the square transposed shapes, k_generic (w[:, ::2] + w[:, 1::2]), k_dot f64,
k_chain with a fused f32 pow step, the f32 pow kernel at y = 2.5 / 2.0 and the f64 pow kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

lib, u = smb.lib(), smb._u64arr
stream = torch.cuda.Stream()
sp = stream.cuda_stream
which = set(sys.argv[1].split(",")) if len(sys.argv) > 1 else {"tile", "generic", "dot", "chain", "pow", "pow64"}
REPS = 2


def run(fn):
    for _ in range(REPS):
        fn()
    torch.cuda.synchronize()
    print(smb.last_kernel())


if "tile" in which:
    for n0, n1 in ((1000, 16384), (8192, 8192)):
        a = torch.rand(n1 * n0, device="cuda")
        b = torch.rand(n0 * n1, device="cuda")
        out = torch.empty(n0 * n1, device="cuda")
        torch.cuda.synchronize()
        argv = (smb.OP_ADD, smb.F32, a.data_ptr(), u([1, n0]), b.data_ptr(), u([n1, 1]), u([n0, n1]), 2, n0 * n1, out.data_ptr(), sp)
        run(lambda: smb._check(lib.smb_elementwise(*argv)))
        del a, b, out
if "generic" in which:
    rows = cols = 16384
    w = torch.rand(rows * cols, device="cuda")
    out = torch.empty(rows * cols // 2, device="cuda")
    torch.cuda.synchronize()
    argv = (smb.OP_ADD, smb.F32, w.data_ptr(), u([cols, 2]), w.data_ptr() + 4, u([cols, 2]), u([rows, cols // 2]), 2, rows * cols // 2, out.data_ptr(), sp)
    run(lambda: smb._check(lib.smb_elementwise(*argv)))   # inner stride 2 on both operands: k_sgather
    cols5 = cols // 5
    argv5 = (smb.OP_ADD, smb.F32, w.data_ptr(), u([cols, 5]), w.data_ptr() + 4, u([cols, 5]), u([rows, cols5]), 2, rows * cols5, out.data_ptr(), sp)
    run(lambda: smb._check(lib.smb_elementwise(*argv5)))  # inner stride 5: k_generic
    del w, out
if "dot" in which:
    n = 1 << 27
    x = torch.rand(n, dtype=torch.float64, device="cuda")
    y = torch.rand(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    run(lambda: smb.dot_ptr(smb.F64, x.data_ptr(), y.data_ptr(), n, sp))
    del x, y
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
if "chain" in which:
    n = 1 << 28
    fa, fb = torch.rand(n, device="cuda") + 0.5, torch.rand(n, device="cuda") + 0.5
    fo = torch.empty(n, device="cuda")
    torch.cuda.synchronize()
    steps = smb.chain_steps(smb.F32, [(None, False, (fa.data_ptr(), [1])), ("add", False, (fb.data_ptr(), [1])), ("pow", False, 2.5)], [n])
    run(lambda: smb._check(lib.smb_chain(smb.F32, steps, 3, u([n]), 1, n, fo.data_ptr(), sp)))
    del fa, fb, fo
if "chainmid" in which:  # a pow step in the MIDDLE of a chain: the general chain kernel with the fast pow core
    n = 1 << 28
    fa, fb, fc = (torch.rand(n, device="cuda") + 0.5 for _ in range(3))
    fo = torch.empty(n, device="cuda")
    torch.cuda.synchronize()
    smb.set_option(smb.OPT_POW_SPECIALISE, 0)
    steps = smb.chain_steps(smb.F32, [(None, False, (fa.data_ptr(), [1])), ("add", False, (fb.data_ptr(), [1])), ("pow", False, 2.5),
                                      ("mul", False, (fc.data_ptr(), [1]))], [n])
    run(lambda: smb._check(lib.smb_chain(smb.F32, steps, 4, u([n]), 1, n, fo.data_ptr(), sp)))
    del fa, fb, fc, fo
if "pow" in which:
    n = 1 << 30
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    o = torch.empty(n, dtype=torch.float32, device="cuda")
    smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0)
    torch.cuda.synchronize()
    for yv in (2.5, 2.0):
        run(lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), yv, n, o.data_ptr(), sp))
    del x, o
if "pow64" in which:
    n = 1 << 28
    x = torch.rand(n, dtype=torch.float64, device="cuda") * 100 + 0.01
    o = torch.empty_like(x)
    torch.cuda.synchronize()
    run(lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F64, x.data_ptr(), 2.5, n, o.data_ptr(), sp))
print("ok", smb.launch_count())
