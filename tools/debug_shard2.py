"""tools/debug_shard2.py -- stress the replicated-operand path of the device set on REAL distinct devices: array (op)
scalar on 100 003 managed floats (ranges below one page: every device reads a private copy), fresh data each round,
with programmatic dependent launch on and off.  Counts stale results."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import simplemath_b200 as smb

lib = smb.lib()
n = 100_003
devs = list(range(smb.device_count())) if smb.device_count() > 1 else [0, 0]
smb.set_option(smb.OPT_SHARD_MIN_BYTES, 0)
rng = np.random.default_rng(1)


def managed(nel):
    p = lib.smb_alloc(nel * 4, smb.MEM_MANAGED)
    return p, np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(nel,))


for pdl in (1, 0, 1):
    smb.set_option(smb.OPT_PDL, pdl)
    smb.set_devices(devs)
    bad_rounds, bad_elems = 0, 0
    for rnd in range(40):
        pa, va = managed(n)
        po, vo = managed(n)
        a = rng.standard_normal(n).astype(np.float32)
        va[:] = a
        lib.smb_host_written(pa)
        smb.array_scalar_ptr(smb.OP_MUL, smb.F32, pa, 1.5, n, po)
        got = vo.copy()
        wrong = int(np.count_nonzero(got != a * np.float32(1.5)))
        bad_rounds += wrong > 0
        bad_elems += wrong
        lib.smb_free(pa)
        lib.smb_free(po)
    smb.set_devices([])
    print(json.dumps({"devices": devs, "pdl": pdl, "rounds": 40, "rounds_with_stale_results": bad_rounds, "stale_elements": bad_elems}), flush=True)
