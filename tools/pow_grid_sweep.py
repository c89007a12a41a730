"""tools/pow_grid_sweep.py -- f32 pow(x, 2.5) grid shape at the two sizes that matter (2^30: one GPU, 2^27: the
8-GPU shard): tiles per CTA (SMB_OPT_CONTIG_VARIANT) x single-tile CTAs at the end of the grid
(SMB_OPT_POW_TAIL_CTAS) x programmatic dependent launch, CUDA events on an explicit stream, back-to-back launches.

    python tools/pow_grid_sweep.py > gpurun_out/pow_grid_sweep.jsonl
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import simplemath_b200 as smb

stream = torch.cuda.Stream()
sp = stream.cuda_stream
smb.set_option(smb.OPT_POW_SPECIALISE, 0)


def timed(fn, reps=20, windows=3):
    best = []
    for _ in range(windows):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        e1.synchronize()
        best.append(e0.elapsed_time(e1) / reps)
    best.sort()
    return best[len(best) // 2], best[0]


for logn in (27, 30):
    n = 1 << logn
    x = torch.empty(n, dtype=torch.float32, device="cuda")
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0)
    torch.cuda.synchronize()
    for y in (2.5, 2.0):
        for pdl in (1, 0):
            for tpc in (0, 2, 4, 8):
                for tail in (-1, 0, 444, 1776, 3552, 7104):
                    if pdl == 0 and (tpc not in (0, 8) or tail not in (-1, 0)):
                        continue
                    if y == 2.0 and (tpc not in (0, 8) or tail not in (-1, 0)):
                        continue
                    smb.set_option(smb.OPT_PDL, pdl)
                    smb.set_option(smb.OPT_CONTIG_VARIANT, tpc)
                    smb.set_option(smb.OPT_POW_TAIL_CTAS, tail)
                    fn = lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), y, n, out.data_ptr(), sp)
                    med, best = timed(fn)
                    print(json.dumps({"log2n": logn, "y": y, "pdl": pdl, "tiles_per_cta": tpc, "tail_ctas": tail, "ms": med,
                                      "gbs": 8 * n / med / 1e6, "gbs_best": 8 * n / best / 1e6}), flush=True)
    # the add kernel beside it, same sizes
    b = torch.empty(n, dtype=torch.float32, device="cuda")
    smb.fill_uniform_f32_ptr(b.data_ptr(), 0, n, 2, -1.0, 1.0)
    for pdl in (1, 0):
        smb.set_option(smb.OPT_PDL, pdl)
        fn = lambda: smb.contiguous_ptr(smb.OP_ADD, smb.F32, x.data_ptr(), b.data_ptr(), out.data_ptr(), n, sp)
        med, best = timed(fn)
        print(json.dumps({"log2n": logn, "kernel": "add", "pdl": pdl, "ms": med, "gbs": 12 * n / med / 1e6, "gbs_best": 12 * n / best / 1e6}), flush=True)
    del x, out, b
smb.set_option(smb.OPT_PDL, 1)
smb.set_option(smb.OPT_CONTIG_VARIANT, 0)
smb.set_option(smb.OPT_POW_TAIL_CTAS, -1)
