import os, sys
sys.path.insert(0, os.getcwd())
import torch, simplemath_b200 as smb
smb.set_option(smb.OPT_POW_SPECIALISE, 0)
n = 1 << 27
x = torch.rand(n, dtype=torch.float64, device="cuda") * 100 + 0.01
o = torch.empty_like(x)
sp = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    smb.array_scalar_ptr(smb.OP_POW, smb.F64, x.data_ptr(), 2.5, n, o.data_ptr(), sp)
torch.cuda.synchronize(); print("ok")
