"""tools/pow_ab.py -- A/B of two builds of libsmb200.so on the f32 pow kernels (general kernel, every tier, the fused
pre-operator form), alternating processes A B A B ... so that clock drift on a shared, power-capped box hits both alike;
the SM clock is sampled while every window runs.

    python tools/pow_ab.py /tmp/libsmb200_old.so simplemath_b200/libsmb200.so [rounds] > gpurun_out/pow_ab.jsonl
    (child) python tools/pow_ab.py --child <label>      with SMB200_LIB set
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def child(label):
    import torch
    import simplemath_b200 as smb
    sys.path.insert(0, HERE)
    from pow_grid_sweep_util import timed, sp
    smb.set_option(smb.OPT_POW_SPECIALISE, 0)
    smb.set_option(smb.OPT_PDL, 2)
    for logn in (30, 27):
        n = 1 << logn
        x = torch.empty(n, dtype=torch.float32, device="cuda")
        b = torch.empty(n, dtype=torch.float32, device="cuda")
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0)
        smb.fill_uniform_f32_ptr(b.data_ptr(), 0, n, 2, 0.5, 1.5)
        torch.cuda.synchronize()
        cases = [("pow", 2.5), ("pow", 2.0), ("pow", 0.5), ("pow", 17.0), ("pow", 300.5), ("add", None), ("fused_pow_add", 2.5)]
        for kind, y in cases:
            if kind == "pow" and abs(y) > 8:   # keep y*log2(x) inside the fast core's range
                smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.8, 1.25)
            if kind == "add":
                fn = lambda: smb.contiguous_ptr(smb.OP_ADD, smb.F32, x.data_ptr(), b.data_ptr(), out.data_ptr(), n, sp)
                bytes_ = 12 * n
            elif kind == "pow":
                fn = lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), y, n, out.data_ptr(), sp)
                bytes_ = 8 * n
            else:
                leaves = [(None, False, (x.data_ptr(), [1])), ("add", False, (b.data_ptr(), [1])), ("pow", False, y)]
                fn = lambda: smb.chain_ptr(smb.F32, leaves, [n], out.data_ptr(), sp)
                bytes_ = 12 * n
            ms, mhz = timed(fn)
            print(json.dumps({"build": label, "log2n": logn, "kernel": kind, "y": y, "launched": smb.last_kernel(), "ms": ms,
                              "gbs": bytes_ / ms / 1e6, "sm_mhz": mhz}), flush=True)
            if kind == "pow" and abs(y) > 8:
                smb.fill_uniform_f32_ptr(x.data_ptr(), 0, n, 3, 0.01, 100.0)
        del x, b, out
    n = 1 << 28                                     # f64 at the C3 size
    x = torch.empty(n, dtype=torch.float64, device="cuda")
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    x.uniform_(0.8, 1.25)
    torch.cuda.synchronize()
    for y in (2.5, 17.0, 0.5):
        fn = lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F64, x.data_ptr(), y, n, out.data_ptr(), sp)
        ms, mhz = timed(fn)
        print(json.dumps({"build": label, "log2n": 28, "kernel": "pow_f64", "y": y, "launched": smb.last_kernel(), "ms": ms,
                          "gbs": 16 * n / ms / 1e6, "sm_mhz": mhz}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        a, b = sys.argv[1], sys.argv[2]
        rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
        for r in range(rounds):
            for label, lib in (("A", a), ("B", b)):
                env = dict(os.environ, SMB200_LIB=os.path.abspath(lib))
                subprocess.run([sys.executable, os.path.abspath(__file__), "--child", label], env=env, check=True)
