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


for pdl, rmode in ((1, 0), (0, 0), (1, 1), (0, 1), (1, 0)):
    smb.set_option(smb.OPT_PDL, pdl)
    smb.set_option(smb.OPT_REPLICA_MODE, rmode)
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
    print(json.dumps({"devices": devs, "pdl": pdl, "replica_mode": rmode, "rounds": 40, "rounds_with_stale_results": bad_rounds, "stale_elements": bad_elems}), flush=True)

# the broadcast configs through the device set, both replica modes: time per call once the partition is in place
import time
smb.set_option(smb.OPT_PDL, 1)
R, C = 65536, 4096
pa, va = managed(R * C)
prow, vrow = managed(C)
po, vo = managed(R * C)
smb.fill_uniform_f32_ptr(pa, 0, R * C, 1, -1.0, 1.0)
vrow[:] = rng.standard_normal(C).astype(np.float32)
lib.smb_host_written(prow)
shape, sa, sb, tot = smb.broadcast((R, C), (C, 1), (1, C), (C, 1))
for rmode in (0, 1):
    smb.set_option(smb.OPT_REPLICA_MODE, rmode)
    for ds in ([], devs):
        smb.set_devices(ds)
        for _ in range(3):
            smb.elementwise_ptr(smb.OP_ADD, smb.F32, pa, sa, prow, sb, shape, po)
        t0 = time.perf_counter()
        for _ in range(20):
            smb.elementwise_ptr(smb.OP_ADD, smb.F32, pa, sa, prow, sb, shape, po)
        ms = (time.perf_counter() - t0) / 20 * 1e3
        # the host rewrites the shared operand raw (no notification), as SMArray::data allows: the next result must see it
        vrow[:7] = np.arange(7, dtype=np.float32) + 100
        smb.elementwise_ptr(smb.OP_ADD, smb.F32, pa, sa, prow, sb, shape, po)
        ok = bool(np.array_equal(vo[:7], va[:7] + vrow[:7])) and bool(np.array_equal(vo[-C:][:7], va[-C:][:7] + vrow[:7]))
        print(json.dumps({"config": "C2x16 row-broadcast add, sync calls", "replica_mode": rmode, "devices": ds or [0], "ms_per_call": ms,
                          "gbs": 4.0 * (2 * R * C + C) / ms / 1e6, "raw_host_write_seen": ok}), flush=True)
        lib.smb_host_written(pa); lib.smb_host_written(po)
smb.set_devices([])
smb.set_option(smb.OPT_REPLICA_MODE, 0)
