"""tools/pow_grid_sweep.py -- f32 pow(x, 2.5) grid shape at the two sizes that matter (2^30: one GPU, 2^27: the
8-GPU shard): tiles per CTA (SMB_OPT_CONTIG_VARIANT) x single-tile CTAs at the end of the grid
(SMB_OPT_POW_TAIL_CTAS) x programmatic dependent launch.  CUDA events on an explicit stream, back-to-back
launches; the SM clock is sampled (NVML) WHILE each window runs, and the whole list is measured in two
interleaved passes, so that clock drift on a shared box shows up as such instead of as a grid effect.

    python tools/pow_grid_sweep.py [--full] > gpurun_out/pow_grid_sweep.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

from pow_grid_sweep_util import timed, stream, sp  # noqa: E402

smb.set_option(smb.OPT_POW_SPECIALISE, 0)
FULL = "--full" in sys.argv


for logn in (27, 30):
    n = 1 << logn
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0)
    smb.fill_uniform_f32_ptr(b.data_ptr(), 0, n, 2, -1.0, 1.0)
    torch.cuda.synchronize()
    cfgs = [("add", None, 1, 0, 0), ("add", None, 0, 0, 0)]
    for pdl in (1, 0):
        for tpc in ((0, 2, 3, 4, 6, 8, 16) if FULL else (0, 3, 4, 8)):
            for tail in ((0, 444, 1776) if FULL else (0, 444)):
                if pdl == 0 and tail:
                    continue
                cfgs.append(("pow", 2.5, pdl, tpc, tail))
    cfgs += [("pow", 2.0, 1, 0, 0), ("pow", 2.0, 1, 8, 0)]
    for pass_ in (0, 1):
        for kind, y, pdl, tpc, tail in (cfgs if pass_ == 0 else cfgs[::-1]):
            smb.set_option(smb.OPT_PDL, 2 if pdl else 0)
            smb.set_option(smb.OPT_CONTIG_VARIANT, tpc)
            smb.set_option(smb.OPT_POW_TAIL_CTAS, tail)
            if kind == "add":
                fn = lambda: smb.contiguous_ptr(smb.OP_ADD, smb.F32, x.data_ptr(), b.data_ptr(), out.data_ptr(), n, sp)
                bytes_ = 12 * n
            else:
                fn = lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), y, n, out.data_ptr(), sp)
                bytes_ = 8 * n
            ms, mhz = timed(fn)
            print(json.dumps({"log2n": logn, "kernel": kind, "y": y, "pdl": pdl, "tiles_per_cta": tpc, "tail_ctas": tail, "pass": pass_,
                              "ms": ms, "gbs": bytes_ / ms / 1e6, "sm_mhz": mhz}), flush=True)
    del x, out, b
smb.set_option(smb.OPT_PDL, 1)
smb.set_option(smb.OPT_CONTIG_VARIANT, 0)
smb.set_option(smb.OPT_POW_TAIL_CTAS, 0)
