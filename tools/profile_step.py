"""tools/profile_step.py -- the smallest program that launches the hot-path kernels
the way bench.py does (device-resident operands, add then pow), for ncu captures.
    python tools/profile_step.py [log2_elems=28] [steps=3]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 28
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = 1 << lg
a, b, x, out, pw = (torch.empty(n, dtype=torch.float32, device="cuda") for _ in range(5))
stream = torch.cuda.Stream()  # explicit: handle 0 would mean "synchronous call on the private stream"
sp = stream.cuda_stream
torch.cuda.synchronize()
smb.fill_uniform_f32_ptr(a.data_ptr(), 0, n, 1, -1.0, 1.0, sp)
smb.fill_uniform_f32_ptr(b.data_ptr(), 0, n, 2, -1.0, 1.0, sp)
smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0, sp)
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
for _ in range(steps):
    smb.contiguous_ptr(smb.OP_ADD, smb.F32, a.data_ptr(), b.data_ptr(), out.data_ptr(), n, sp)
    smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), 2.5, n, pw.data_ptr(), sp)
torch.cuda.synchronize()
print("ok", smb.launch_count(), smb.last_kernel())
