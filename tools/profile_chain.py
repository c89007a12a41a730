"""tools/profile_chain.py -- launches the fused-chain cases F1 / F3 of tools/bench_configs.py twice each, for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

n = 1 << 28
fa, fb, fc = (torch.rand(n, device="cuda") + 0.5 for _ in range(3))
frow = torch.rand(16384, device="cuda")
fo = torch.empty(n, device="cuda")
stream = torch.cuda.Stream()  # explicit: handle 0 would mean "synchronous call on the private stream"
sp = stream.cuda_stream
torch.cuda.synchronize()
P = lambda t: t.data_ptr()
L = lambda op, t, st=(1,): (op, False, (P(t), list(st)))
s1 = smb.chain_steps(smb.F32, [L(None, fa), L("add", fb), L("mul", fc)], [n])
sh = [16384, 16384]
s3 = smb.chain_steps(smb.F32, [(None, False, (P(fa), [16384, 1])), ("mul", False, (P(frow), [0, 1])), ("add", False, (P(fb), [16384, 1]))], sh)
torch.cuda.synchronize()
for _ in range(2):
    smb._check(smb.lib().smb_chain(smb.F32, s1, 3, smb._u64arr([n]), 1, n, P(fo), sp))
for _ in range(2):
    smb._check(smb.lib().smb_chain(smb.F32, s3, 3, smb._u64arr(sh), 2, n, P(fo), sp))
torch.cuda.synchronize()
print("ok", smb.last_kernel())
