import sys, os, time, ctypes
sys.path.insert(0, os.getcwd())
import torch, simplemath_b200 as smb
lib = smb.lib()
n = 1024
a = torch.ones(n, device="cuda"); b = torch.ones(n, device="cuda"); o = torch.empty(n, device="cuda")
sp = torch.cuda.current_stream().cuda_stream
def bench(fn, reps=20000):
    for _ in range(100): fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    host = (time.perf_counter() - t) / reps * 1e6
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t) / reps * 1e6
    return host, tot
argv = (smb.OP_ADD, smb.F32, a.data_ptr(), b.data_ptr(), o.data_ptr(), n, sp)
print("smb_contiguous torch ptrs, async stream: host %.2f us/call, total %.2f" % bench(lambda: lib.smb_contiguous(*argv)))
argv2 = (smb.OP_ADD, smb.F32, a.data_ptr(), b.data_ptr(), o.data_ptr(), n, None)
print("smb_contiguous torch ptrs, sync: %.2f us/call, total %.2f" % bench(lambda: lib.smb_contiguous(*argv2), 5000))
pa = lib.smb_alloc(n * 4, smb.MEM_DEVICE); pb = lib.smb_alloc(n * 4, smb.MEM_DEVICE); po = lib.smb_alloc(n * 4, smb.MEM_DEVICE)
argv3 = (smb.OP_ADD, smb.F32, pa, pb, po, n, sp)
print("smb_contiguous pool ptrs, async: %.2f us/call, total %.2f" % bench(lambda: lib.smb_contiguous(*argv3)))
ma = lib.smb_alloc(n * 4, smb.MEM_MANAGED); mb = lib.smb_alloc(n * 4, smb.MEM_MANAGED); mo = lib.smb_alloc(n * 4, smb.MEM_MANAGED)
argv4 = (smb.OP_ADD, smb.F32, ma, mb, mo, n, None)
print("smb_contiguous managed ptrs, sync (what the drop-in SMArray does): %.2f us/call, total %.2f" % bench(lambda: lib.smb_contiguous(*argv4), 5000))
print("torch add same size: %.2f us/call, total %.2f" % bench(lambda: torch.add(a, b, out=o)))
print("smb_device_count: %.2f, %.2f" % bench(lambda: lib.smb_device_count()))
print("smb_owns: %.2f, %.2f" % bench(lambda: lib.smb_owns(pa)))
