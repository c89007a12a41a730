"""tools/chain_pow_sweep.py -- sm::pow(a + b, 2.5) fused (k_chain<pow>): SMB_OPT_CHAIN_POW_VARIANT 0..3 beside the two
separate kernels, 2^28 f32 elements, CUDA events on an explicit stream, SM clock sampled while each window runs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb
from pow_grid_sweep_util import timed, stream, sp  # noqa: E402  (tools/ is on sys.path when run as a script)

smb.set_option(smb.OPT_POW_SPECIALISE, 0)
n = 1 << 28
fa, fb = torch.rand(n, device="cuda") + 0.5, torch.rand(n, device="cuda") + 0.5
fo, ft = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
torch.cuda.synchronize()
u = smb._u64arr
steps = smb.chain_steps(smb.F32, [(None, False, (fa.data_ptr(), [1])), ("add", False, (fb.data_ptr(), [1])), ("pow", False, 2.5)], [n])
ref = None
for pass_ in (0, 1):
    for var in ((0, 1, 2, 3, 4) if pass_ == 0 else (4, 3, 2, 1, 0)):
        smb.set_option(smb.OPT_CHAIN_POW_VARIANT, var)
        ms, mhz = timed(lambda: smb._check(smb.lib().smb_chain(smb.F32, steps, 3, u([n]), 1, n, fo.data_ptr(), sp)))
        torch.cuda.synchronize()
        if ref is None:
            ref = fo.clone()
        same = bool(torch.equal(ref, fo))
        if var == 4:
            fused4 = fo.clone()
        print(json.dumps({"config": "pow(a+b,2.5) fused", "variant": var, "pass": pass_, "ms": ms, "gbs": 12 * n / ms / 1e6, "sm_mhz": mhz,
                          "kernel": smb.last_kernel(), "same_bits_as_variant_0": same}), flush=True)
    ms, mhz = timed(lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, fa.data_ptr(), fb.data_ptr(), ft.data_ptr(), n, sp),
                             smb.array_scalar_ptr(smb.OP_POW, smb.F32, ft.data_ptr(), 2.5, n, fo.data_ptr(), sp)))
    torch.cuda.synchronize()
    print(json.dumps({"config": "add then pow (two kernels, 20 B/elem)", "pass": pass_, "ms": ms, "gbs_of_fused_bytes": 12 * n / ms / 1e6, "sm_mhz": mhz,
                      "same_bits_as_k_chain": bool(torch.equal(ref, fo)), "same_bits_as_fused_pow_kernel": bool(torch.equal(fused4, fo))}), flush=True)
smb.set_option(smb.OPT_CHAIN_POW_VARIANT, 4)
