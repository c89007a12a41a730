"""tools/pdl_stress.py -- hammer the programmatic-dependent-launch bookkeeping: in async mode on a device set that lists ONE
device twice (two ranges per operator, all on one private stream), run `tmp = c * c; d = tmp + b` again and again with small
ranges, where the second operator's second range is admitted behind a launch that has to wait.  Prints how many rounds gave
a wrong `d`.   SMB200_LIB=<another build> python tools/pdl_stress.py [rounds]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import simplemath_b200 as smb

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 400
lib = smb.lib()
rng = np.random.default_rng(1)
bad = 0
smb.set_option(smb.OPT_SHARD_MIN_BYTES, 0)
smb.set_devices([0, 0])
# a long kernel in front of every round: the small launches queue up behind it (Python alone enqueues more slowly than the
# GPU drains 4 us kernels), so that they become eligible back to back -- the only situation in which dependents overlap
NBIG = 1 << 27
big_a = lib.smb_alloc(NBIG * 4, smb.MEM_MANAGED)
big_o = lib.smb_alloc(NBIG * 4, smb.MEM_MANAGED)
smb.contiguous_ptr(smb.OP_ADD, smb.F32, big_a, big_a, big_o, NBIG)
smb.sync()
for n in (1 << 21, 1 << 20, 3 << 19):
    full = [1]

    def managed(arr=None):
        p = lib.smb_alloc(n * 4, smb.MEM_MANAGED)
        v = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n,))
        if arr is not None:
            v[...] = arr
            lib.smb_host_written(p)
        return p, v

    b = rng.standard_normal(n).astype(np.float32)
    pb, _ = managed(b)
    pc, vc = managed()
    pd, vd = managed()
    smb.set_option(smb.OPT_ASYNC, 1)
    for r in range(rounds):
        c = rng.standard_normal(n).astype(np.float32) if r % 50 == 0 else c + np.float32(1.0)
        if r % 50 == 0:
            smb.sync()
            vc[...] = c
            lib.smb_host_written(pc)
        else:
            smb.array_scalar_ptr(smb.OP_ADD, smb.F32, pc, 1.0, n, pc)                      # c += 1 on the stream
        smb.contiguous_ptr(smb.OP_ADD, smb.F32, big_a, big_a, big_o, NBIG)                  # ~0.25 ms: the queue fills behind it
        tmp = lib.smb_alloc(n * 4, smb.MEM_MANAGED)
        smb.contiguous_ptr(smb.OP_MUL, smb.F32, pc, pc, tmp, n)
        smb.contiguous_ptr(smb.OP_ADD, smb.F32, tmp, pb, pd, n)
        lib.smb_free(tmp)
        smb.sync()
        want = c * c + b
        if not np.array_equal(vd, want):
            bad += 1
    smb.set_option(smb.OPT_ASYNC, 0)
    smb.sync()
    for p in (pb, pc, pd):
        lib.smb_free(p)
print(json.dumps({"lib": os.path.basename(smb.LIB_PATH), "rounds_per_size": rounds, "sizes": 3, "wrong_rounds": bad}))
