"""tools/pcie_probe.py -- what the box's PCIe link sustains from pinned memory: H2D alone, D2H alone,
both at once.  The ceiling bench.py's e2e figure (host buffers, copies inside the timed region) runs
against: one C5 step moves 12.9 GB up and 8.6 GB down."""
import json
import torch

n = 1 << 30  # bytes
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def up():
    with torch.cuda.stream(s1):
        d_up.copy_(h_up, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        h_dn.copy_(d_dn, non_blocking=True)


def both():
    up(); down()


import time
def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3

r = {"h2d_gbs": n / wall(up) / 1e6, "d2h_gbs": n / wall(down) / 1e6}
ms = wall(both)
r["both_h2d_gbs"] = n / ms / 1e6
r["both_total_gbs"] = 2 * n / ms / 1e6
# the C5 step's mix: 12.9 GB up, 8.6 GB down, perfectly overlapped
r["c5_step_floor_ms"] = max(12.884901888e9 / (r["both_h2d_gbs"] * 1e9), 8.589934592e9 / (r["both_h2d_gbs"] * 1e9)) * 1e3
print(json.dumps(r))
