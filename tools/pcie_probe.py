"""tools/pcie_probe.py -- what the box's PCIe links sustain from pinned memory: H2D alone, D2H alone, both at
once -- on ONE GPU, or on all ranks CONCURRENTLY under torchrun (one rank per GPU, barrier before every window):
the ceiling bench.py's e2e figure (host buffers, copies inside the timed region) runs against.  One C5 step moves
12.9 GB up and 8.6 GB down in total, split over the ranks.

    python tools/pcie_probe.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/pcie_probe.py [--bind]

--bind pins each rank to the NUMA node of its GPU before allocating (bench.py does the same at N > 1).
Rank 0 prints one JSON object: per-rank and aggregate GB/s for the three patterns, the slab-size dependence of
the duplex pattern, and the NUMA / affinity facts that explain them.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
binding = "not requested"
if "--bind" in sys.argv:
    from bench import bind_to_gpu_numa_node
    binding = bind_to_gpu_numa_node(local)
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

n = 1 << 30  # bytes per direction per window
h_up = torch.empty(n, dtype=torch.uint8).pin_memory()
h_dn = torch.empty(n, dtype=torch.uint8).pin_memory()
d_up = torch.empty(n, dtype=torch.uint8, device="cuda")
d_dn = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def wall(fn, reps=4):
    fn()
    barrier()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    barrier()
    return dt


def up(slab=n):
    with torch.cuda.stream(s1):
        for o in range(0, n, slab):
            d_up[o:o + slab].copy_(h_up[o:o + slab], non_blocking=True)


def down(slab=n):
    with torch.cuda.stream(s2):
        for o in range(0, n, slab):
            h_dn[o:o + slab].copy_(d_dn[o:o + slab], non_blocking=True)


def gather(v):
    if world == 1:
        return [v]
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(x.item()) for x in out]


res = {"ranks": world, "binding_rank0": binding}
t_up, t_dn = wall(up), wall(down)
t_both = wall(lambda: (up(), down()))
per = {"h2d_gbs": gather(n / t_up / 1e9), "d2h_gbs": gather(n / t_dn / 1e9), "duplex_total_gbs": gather(2 * n / t_both / 1e9)}
slabs = {}
for mb in (16, 64, 256):
    t = wall(lambda: (up(mb << 20), down(mb << 20)))
    slabs[f"duplex_total_gbs_{mb}MiB_slabs"] = gather(2 * n / t / 1e9)
if rank == 0:
    res["per_rank"] = per
    res["aggregate"] = {k: sum(v) for k, v in per.items()}
    res["slab_size"] = {k: sum(v) for k, v in slabs.items()}
    agg = res["aggregate"]["duplex_total_gbs"]
    # the C5 step's mix: 12.9 GB up, 8.6 GB down over all ranks; the slower direction sets the floor
    res["c5_step_floor_ms"] = max(12.884901888e9, 8.589934592e9) / (agg / 2 * 1e9) * 1e3
    res["c5_e2e_ceiling_gbs"] = 21.474836480e9 / (res["c5_step_floor_ms"] * 1e-3) / 1e9
    try:
        res["host"] = {"cpus_allowed": len(os.sched_getaffinity(0)),
                       "numa_nodes": len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()])}
    except Exception:
        pass
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
