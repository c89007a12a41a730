"""tools/profile_configs.py -- launches each non-headline kernel once or twice (C2 row
broadcast, C2 x16 rows, C4 outer mul / div, f64 pow, large-|y| f32 pow) for ncu captures."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

lib, u = smb.lib(), smb._u64arr
stream = torch.cuda.Stream()  # explicit: handle 0 would mean "synchronous call on the private stream"
sp = stream.cuda_stream
torch.cuda.synchronize()


def bcast(op, dt, tdt, s1, s2, reps=2):
    shape, sa, sb, n = smb.broadcast(s1, smb.row_major_strides(s1), s2, smb.row_major_strides(s2))
    n1 = 1
    for d in s1:
        n1 *= d
    n2 = 1
    for d in s2:
        n2 *= d
    a = torch.ones(n1, dtype=tdt, device="cuda") * 3
    b = torch.ones(n2, dtype=tdt, device="cuda") * 2
    out = torch.empty(n, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    for _ in range(reps):
        smb._check(lib.smb_elementwise(op, dt, a.data_ptr(), u(sa), b.data_ptr(), u(sb), u(shape), len(shape), n, out.data_ptr(), sp))
    torch.cuda.synchronize()
    print(smb.last_kernel(), shape)


bcast(smb.OP_ADD, smb.F32, torch.float32, (4096, 4096), (1, 4096))
bcast(smb.OP_ADD, smb.F32, torch.float32, (65536, 4096), (1, 4096))
bcast(smb.OP_MUL, smb.I32, torch.int32, (512, 1, 1024), (1, 512, 1024))
bcast(smb.OP_DIV, smb.I32, torch.int32, (512, 1, 1024), (1, 512, 1024))
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
n = 1 << 27
x = torch.rand(n, dtype=torch.float64, device="cuda") * 100 + 0.01
o = torch.empty_like(x)
torch.cuda.synchronize()
for _ in range(2):
    smb.array_scalar_ptr(smb.OP_POW, smb.F64, x.data_ptr(), 2.5, n, o.data_ptr(), sp)
xf = torch.rand(1 << 28, dtype=torch.float32, device="cuda") * 100 + 0.01
of = torch.empty_like(xf)
torch.cuda.synchronize()
for _ in range(2):
    smb.array_scalar_ptr(smb.OP_POW, smb.F32, xf.data_ptr(), 9.25, 1 << 28, of.data_ptr(), sp)
torch.cuda.synchronize()
print("ok", smb.launch_count())
