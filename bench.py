#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json metric, config C5).

Workload "C5": float32, n = 2^30 elements (4 GiB per array), contiguous:
    step = { out = a + b            (element_wise_op<float, AddOp>,   12 B/elem)
             pw  = pow(x, 2.5)      (sm::pow -> array_scalar_op<PowOp>, 8 B/elem,
                                     general exp2(y*log2 x) kernel) }
sharded across the N GPUs by flat output index range (rank g owns
shard_range(n, g, N)); no collective on the data path.  Total work is fixed as N
grows -> "scaling": "strong" (the default: BASELINE config C5 shards 4 GiB arrays);
--scaling weak keeps 2^30 elements per GPU instead.

    python bench.py [--gpus N] [--steps K] [--warmup W]       this repo's CUDA path
    python bench.py --impl reference ...                      the reference's own CPU
                                                              implementation (oracle/_ref)

For N > 1 the driver launches one rank per GPU with torch.distributed.run.
Rank 0 prints ONE JSON line.  `value` = algorithmic GB/s of the whole job with
inputs resident in HBM; `e2e` = the same metric through the C ABI with pinned
HOST buffers (H2D + D2H inside the timed region); `roofline` = the dominant
kernel (contiguous add) timed alone with CUDA events against the measured HBM
peak; `cpu_baseline` = the compiled reference timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "elementwise algorithmic GB/s, f32 add + pow on 4 GiB arrays (C5)"
WORKLOAD = "C5: f32 contiguous add (a+b) + pow(x, 2.5) on 2^30-element (4 GiB) arrays"  # the same string in both arms
UNIT = "GB/s"
N_TOTAL = 1 << 30
POW_Y = 2.5
ADD_BYTES_PER_ELEM = 12  # 2 reads + 1 write  
POW_BYTES_PER_ELEM = 8   # 1 read + 1 write
L2_BYTES = 126 << 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smb200", choices=["smb200", "reference"])
    ap.add_argument("--elems", type=int, default=N_TOTAL, help="total elements (default 2^30 = config C5)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default, BASELINE config C5): --elems is the TOTAL, sharded over the ranks; "
                         "weak: --elems per GPU, the job grows with N")
    ap.add_argument("--streams", type=int, default=2, choices=[1, 2],
                    help="2 (default): the step's two independent operators go to two streams; 1: one stream")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-elems", type=int, default=0,
                    help="elements per array of the CPU (reference) run; 0 = the full 2^30-element job when host memory "
                         "allows (>= 48 GiB available), else a 2^26-element sample")
    ap.add_argument("--pdl", type=int, default=1, choices=[0, 1],
                    help="1 (default): programmatic dependent launch between back-to-back kernels of a stream; 0: plain stream order")
    ap.add_argument("--windows", type=int, default=5,
                    help="the timed region is repeated this many times (each of --steps steps); `value` is the median window")
    return ap.parse_args()


def host_mem_available_gib() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / (1 << 20)
    except Exception:
        pass
    return 0.0


def cpu_elems_default(args) -> int:
    if args.cpu_elems > 0:
        return args.cpu_elems
    # the full job: 3 inputs of 4 GiB + the reference's own 4 GiB result per operator
    return N_TOTAL if host_mem_available_gib() >= 48.0 else 1 << 26


def measured_peak():
    """HBM roofline denominator: MEASURED_PEAKS.json (driver-written), else the
    profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json: torch copy_ 1 Gi bf16)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(elements_per_launch: int):
    """Per-launch DRAM bytes of the add / pow kernels, from the committed ncu --set full capture
    (profiles/traffic.json, taken on one GPU at `elements` per launch) scaled to this run's launch size:
    both kernels stream every byte exactly once, so their traffic is proportional to the elements a
    launch covers.  None when the capture is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        k = elements_per_launch / float(t["elements"])
        return {"add": t["add_dram_bytes_per_launch"] * k, "pow": t["pow_dram_bytes_per_launch"] * k,
                "note": "ncu dram__bytes_read+write per launch" + ("" if k == 1 else f", scaled x{k:.4g} from the {t['elements']}-element capture")}
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during a timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ reference arm
def cpu_reference(elems: int, steps: int, warmup: int):
    """The reference's own CPU implementation of the step (SMArray operator+ and
    sm::pow as shipped, incl. the per-call result new[]), all host threads."""
    import numpy as np
    import oracle
    ref = oracle.reference()
    kind = "reference"
    if ref is None:  # oracle/_ref never built: fall back to the C restatement
        ref, kind = oracle.c_oracle(), "port"
    orc = oracle.c_oracle()
    a = orc.fill_uniform_f32(0, elems, 1, -1.0, 1.0)
    b = orc.fill_uniform_f32(0, elems, 2, -1.0, 1.0)
    x = orc.fill_uniform_f32(0, elems, 3, 0.01, 100.0)
    if kind == "reference":
        # torchrun exports OMP_NUM_THREADS=1; the baseline is the reference with ALL host threads
        ref.h.smref_set_threads(os.cpu_count() or 1)
    cores = ref.threads() if kind == "reference" else (os.cpu_count() or 1)

    def step():
        if kind == "reference":
            ref.smarray_binary("add", a, b, want_result=False)
            ref.smarray_scalar("pow", x, POW_Y, want_result=False)
        else:
            ref.elementwise("add", a, [1], b, [1], [elems])
            ref.array_scalar("pow", x, POW_Y)

    for _ in range(warmup):
        step()
    t_add = t_pow = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    # per-op split (one extra pass each)
    t1 = time.perf_counter()
    if kind == "reference":
        ref.smarray_binary("add", a, b, want_result=False)
    else:
        ref.elementwise("add", a, [1], b, [1], [elems])
    t_add = time.perf_counter() - t1
    t_pow = max(dt - t_add, 0.0)
    bytes_step = elems * (ADD_BYTES_PER_ELEM + POW_BYTES_PER_ELEM)
    return {
        "value": bytes_step / dt / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{elems} f32 elements per array ({elems / N_TOTAL:.4g} of C5), {steps} steps, operators as shipped "
                  f"(alloc included); add runs the reference's single-threaded contiguous path (calculate.h:101-134), "
                  f"pow its OpenMP array_scalar_op with libm powf",
        "ms_per_step": dt * 1e3, "add_gbs": elems * ADD_BYTES_PER_ELEM / max(t_add, 1e-9) / 1e9,
        "pow_gbs": elems * POW_BYTES_PER_ELEM / max(t_pow, 1e-9) / 1e9,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    elems = cpu_elems_default(args)
    res = cpu_reference(elems, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD + ("" if elems == N_TOTAL else f"; bounded CPU sample of {elems} elements per array"),
                   "elements": elems, "pow_exponent": POW_Y},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ops": {"add_gbs": res["add_gbs"], "pow_gbs": res["pow_gbs"]},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------ our arm
def bind_to_gpu_numa_node(local: int):
    """Pin this rank's host threads (and so its first-touch pinned allocations) to the NUMA node its GPU
    hangs off, when the box has more than one.  Returns a one-line description for the record."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return f"gpu {local} numa_node={node}, {len(nodes)} node(s): not bound"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"gpu {local} numa_node={node}: bound to {len(cpus)} cpus"
        return f"gpu {local} numa_node={node}: no allowed cpu there, not bound"
    except Exception as e:  # best effort: the record says what happened
        return f"not bound ({type(e).__name__}: {e})"


def run_smb(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import simplemath_b200 as smb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the smb200 arm has no CPU fallback; use --impl reference)")
    numa = bind_to_gpu_numa_node(local) if world > 1 else "single rank: not bound"
    torch.cuda.set_device(local)
    smb.lib().smb_set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: an NCCL_DEBUG level the launcher asked for is honoured and
        # its log goes to stderr; without one, warnings only
        if "NCCL_DEBUG" in os.environ:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        else:
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    n = args.elems * (world if args.scaling == "weak" else 1)
    lo, hi = smb.shard_range(n, rank, world, align=4096)
    m = hi - lo  # this rank's shard
    # explicit NON-default streams: the C ABI maps stream handle 0 to "synchronous call on the library's private
    # stream", so the legacy default stream would silently time host round trips instead of kernels
    stream = torch.cuda.Stream(device=dev)
    sp = stream.cuda_stream
    assert sp != 0
    a = torch.empty(m, dtype=torch.float32, device=dev)
    b = torch.empty(m, dtype=torch.float32, device=dev)
    x = torch.empty(m, dtype=torch.float32, device=dev)
    out = torch.empty(m, dtype=torch.float32, device=dev)
    pw = torch.empty(m, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()
    # counter-based generator of the FLAT index: every rank produces exactly its slice of the global arrays
    smb.fill_uniform_f32_ptr(a.data_ptr(), lo, m, 1, -1.0, 1.0, sp)
    smb.fill_uniform_f32_ptr(b.data_ptr(), lo, m, 2, -1.0, 1.0, sp)
    smb.fill_uniform_f32_ptr(x.data_ptr(), lo, m, 3, 0.01, 100.0, sp)
    stream.synchronize()
    smb.set_option(smb.OPT_POW_SPECIALISE, 0)  # headline = the general pow kernel
    smb.set_option(smb.OPT_PDL, 2 if args.pdl else 0)  # 2: also on the streams this script owns (only this library's kernels and events go there)

    # The step's two operators are independent (a+b -> out, pow(x) -> pw): with --streams 2 (default)
    # they are enqueued on two streams, so one kernel's last wave overlaps the other's first; within a
    # stream, back-to-back launches use programmatic dependent launch (the next grid is resident
    # while the previous one drains).  Each stream keeps its own operator in order from step to step;
    # the timed region joins both.
    stream2 = torch.cuda.Stream(device=dev) if args.streams == 2 else stream
    sp2 = stream2.cuda_stream

    def step():
        smb.contiguous_ptr(smb.OP_ADD, smb.F32, a.data_ptr(), b.data_ptr(), out.data_ptr(), m, sp)
        smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), POW_Y, m, pw.data_ptr(), sp2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        """CUDA events on the launching stream (the second stream is forked from / joined into it)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if stream2 is not stream:
            stream2.wait_event(e0)          # the second stream starts inside the timed region ...
        for _ in range(reps):
            fn()
        if stream2 is not stream:
            j = torch.cuda.Event()
            j.record(stream2)
            stream.wait_event(j)            # ... and the closing event waits for it
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1)  # ms

    def max_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def windows(fn, reps, k):
        """k timed windows of `reps` launches, each bracketed by a barrier + synchronize, max over ranks;
        returns the per-launch milliseconds of every window."""
        res = []
        for _ in range(k):
            barrier()
            res.append(max_ranks(timed(fn, reps)) / reps)
        barrier()
        return res

    def med(v):
        s_ = sorted(v)
        return s_[len(s_) // 2]

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = smb.launch_count()
    step_windows = windows(step, args.steps, max(1, args.windows))
    launches = (smb.launch_count() - l0) // max(1, args.windows)
    clocks = sampler.stop()
    ms_step = med(step_windows)
    bytes_step_all = n * (ADD_BYTES_PER_ELEM + POW_BYTES_PER_ELEM)
    value = bytes_step_all / (ms_step * 1e-3) / 1e9
    spread = {"windows": len(step_windows), "steps_per_window": args.steps,
              "gbs_min": bytes_step_all / (max(step_windows) * 1e-3) / 1e9,
              "gbs_max": bytes_step_all / (min(step_windows) * 1e-3) / 1e9,
              "ms_per_step_windows": step_windows}

    # ---- per-kernel timing (each kernel alone, CUDA events on its stream) -> roofline
    def add_only():
        smb.contiguous_ptr(smb.OP_ADD, smb.F32, a.data_ptr(), b.data_ptr(), out.data_ptr(), m, sp)

    def pow_only():
        smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), POW_Y, m, pw.data_ptr(), sp)

    reps = max(args.steps, 10)
    add_only(); pow_only()
    ms_add = med(windows(add_only, reps, 3))
    ms_pow = med(windows(pow_only, reps, 3))
    peak, peak_src = measured_peak()
    add_bytes_launch = m * ADD_BYTES_PER_ELEM   # algorithmic bytes one launch (this rank) moves
    pow_bytes_launch = m * POW_BYTES_PER_ELEM
    add_gbs = add_bytes_launch / (ms_add * 1e-3) / 1e9
    pow_gbs = pow_bytes_launch / (ms_pow * 1e-3) / 1e9
    traffic = ncu_traffic(m) or {}
    roofline = {"bound": "hbm", "kernel": "k_stream<float, BinaryFn<ADD>> (contiguous add)", "achieved": add_gbs,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": add_gbs / peak,
                "frac_of_nominal_8000": add_gbs / 8000.0,
                "algorithmic_bytes_per_launch": add_bytes_launch, "ms_per_launch": ms_add,
                "traffic": traffic.get("add"), "traffic_source": traffic.get("note")}
    roofline_pow = {"bound": "hbm", "kernel": "k_stream<float, PowF32Fn> (pow y=2.5, general)", "achieved": pow_gbs,
                    "peak": peak, "unit": "GB/s", "frac": pow_gbs / peak, "frac_of_nominal_8000": pow_gbs / 8000.0,
                    "algorithmic_bytes_per_launch": pow_bytes_launch, "ms_per_launch": ms_pow,
                    "traffic": traffic.get("pow")}

    # ---- extra pow variants (reference benchmark exponent y=2; specialised form)
    extras = {}

    def pow_y(yv, spec):
        smb.set_option(smb.OPT_POW_SPECIALISE, spec)
        fn = lambda: smb.array_scalar_ptr(smb.OP_POW, smb.F32, x.data_ptr(), yv, m, pw.data_ptr(), sp)
        fn()
        ms = med(windows(fn, reps, 3))
        smb.set_option(smb.OPT_POW_SPECIALISE, 0)
        return pow_bytes_launch / (ms * 1e-3) / 1e9
    extras["pow_y2_general_gbs_per_gpu"] = pow_y(2.0, 0)
    extras["pow_y2_specialised_gbs_per_gpu"] = pow_y(2.0, 1)

    # ---- op-chain fusion: (a + b) * x in ONE pass vs the two operators
    def P(t):
        return t.data_ptr()
    chain = smb.chain_steps(smb.F32, [(None, False, (P(a), [1])), ("add", False, (P(b), [1])), ("mul", False, (P(x), [1]))], [m])
    shp = smb._u64arr([m])
    fused = lambda: smb._check(smb.lib().smb_chain(smb.F32, chain, 3, shp, 1, m, P(pw), sp))
    unfused = lambda: (smb.contiguous_ptr(smb.OP_ADD, smb.F32, P(a), P(b), P(out), m, sp),
                       smb.contiguous_ptr(smb.OP_MUL, smb.F32, P(out), P(x), P(pw), m, sp))
    fused(); unfused()
    ms_f = med(windows(fused, reps, 3))
    ms_u = med(windows(unfused, reps, 3))
    extras["chain_add_mul_fused_gbs_per_gpu"] = 16.0 * m / (ms_f * 1e-3) / 1e9     # 3 leaves + result
    extras["chain_add_mul_speedup_vs_two_operators"] = ms_u / ms_f

    # ---- the other shapes of the path, sharded by flat output range like C5 (each rank: its range of the result,
    # the broadcast operand replicated), plus the sharded dot product -- the one place a collective sits on a
    # path: per-rank partial + ONE NCCL all-reduce of a scalar.  Parity per rank against the oracle below.
    sharded = {}
    C2R, C2C = 65536, 4096                       # C2 x 16: {65536,4096} + {1,4096} (1 GiB result)
    c2_shape, c2_sa, c2_sb, c2_n = smb.broadcast((C2R, C2C), (C2C, 1), (1, C2C), (C2C, 1))
    c2_lo, c2_hi = smb.shard_range(c2_n, rank, world, align=C2C * 128)
    c2_cnt = c2_hi - c2_lo
    c2a = torch.empty(c2_n, dtype=torch.float32, device=dev)   # the FULL operand on every rank; a rank reads only its range
    smb.fill_uniform_f32_ptr(P(c2a), 0, c2_n, 7, -1.0, 1.0, sp)
    c2_a_ptr = P(c2a)
    c2_row = x[:C2C]
    xout = torch.empty(max(c2_cnt, 1), dtype=torch.float32, device=dev)   # results of the extra configs (C2 x 16 and C4 shards have equal size)
    c2 = lambda: smb.elementwise_range_ptr(smb.OP_ADD, smb.F32, c2_a_ptr, c2_sa, P(c2_row), c2_sb, c2_shape, c2_lo, c2_cnt, P(xout), sp)
    c2()
    ms_c2 = med(windows(c2, reps, 3))
    sharded["c2x16_row_broadcast_add_gbs"] = 4.0 * (2 * c2_n + C2C) / (ms_c2 * 1e-3) / 1e9   # whole job: max over ranks is the time
    sharded["c2x16_kernel"] = smb.last_kernel()
    stream.synchronize()
    c2_ok = True
    if c2_cnt:
        w = min(1 << 16, c2_cnt)
        ha = c2a[c2_lo:c2_lo + w].cpu().numpy()
        hr = c2_row.cpu().numpy()
        col0 = c2_lo % C2C
        want = ha + np.resize(np.roll(hr, -col0), w)
        c2_ok = bool(np.array_equal(xout[:w].cpu().numpy(), want))
    del c2a
    # C4: int32 {512,1,1024} x {1,512,1024} -> {512,512,1024}, mul and div; rank g owns whole dim-0 slabs
    D0, D1, L = 512, 512, 1024
    c4_shape, c4_sa, c4_sb, c4_n = smb.broadcast((D0, 1, L), (L, L, 1), (1, D1, L), (D1 * L, L, 1))
    c4_lo, c4_hi = smb.shard_range(c4_n, rank, world, align=D1 * L)
    c4_cnt = c4_hi - c4_lo
    gi = torch.Generator(device=dev); gi.manual_seed(4)
    ia = torch.randint(-1000, 1001, (D0 * L,), dtype=torch.int32, device=dev, generator=gi)
    ib = torch.randint(1, 98, (D1 * L,), dtype=torch.int32, device=dev, generator=gi)
    assert c4_cnt <= xout.numel()
    iout = xout.view(torch.int32)
    c4_ok = True
    for name, opc in (("mul", smb.OP_MUL), ("div", smb.OP_DIV)):
        f4 = lambda: smb.elementwise_range_ptr(opc, smb.I32, P(ia), c4_sa, P(ib), c4_sb, c4_shape, c4_lo, c4_cnt, P(iout), sp)
        f4()
        ms_c4 = med(windows(f4, reps, 3))
        sharded[f"c4_int32_{name}_gbs"] = 4.0 * (c4_n + D0 * L + D1 * L) / (ms_c4 * 1e-3) / 1e9
        sharded[f"c4_{name}_kernel"] = smb.last_kernel()
        stream.synchronize()
        if c4_cnt:
            i0 = c4_lo // (D1 * L)
            ta, tb = ia.view(D0, 1, L)[i0:i0 + 1], ib.view(1, D1, L)
            ref = ta * tb if name == "mul" else torch.div(ta, tb, rounding_mode="trunc")
            c4_ok &= bool((iout[:D1 * L].view(1, D1, L) == ref).all())
    # sharded dot: every rank reduces its range of a . b, one all-reduce adds the partials
    part = smb.dot_ptr(smb.F32, P(a), P(b), m, sp)
    if world > 1:
        t = torch.tensor([part], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dot_all = float(t.item())
    else:
        dot_all = float(part)
    dref = float((a.double() * b.double()).sum().item())
    if world > 1:
        t = torch.tensor([dref], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dref = float(t.item())
    fdot = lambda: smb.dot_ptr(smb.F32, P(a), P(b), m, sp)
    t0 = time.perf_counter()
    for _ in range(5):
        fdot()
    ms_dot = max_ranks((time.perf_counter() - t0) / 5 * 1e3)   # a scalar result: synchronous, host-timed
    sharded["dot_f32_gbs"] = 8.0 * n / (ms_dot * 1e-3) / 1e9
    sharded["dot_abs_error_vs_f64_sum"] = abs(dot_all - dref)
    dot_ok = abs(dot_all - dref) <= 1e-6 * n ** 0.5 + 1e-3 * abs(dref)
    sharded["parity"] = {"c2x16": c2_ok, "c4": c4_ok, "dot": bool(dot_ok)}
    del ia, ib, xout, iout

    step()  # leave out / pw holding the step's results for verification
    torch.cuda.synchronize()
    # ---- verification (outside every timed region): windows vs the oracle on rank 0, the WHOLE pow result audited
    # on the device by every rank, per-shard checksums all-gathered over NCCL
    ok = True
    audit_over, audit_worst = smb.pow_audit_f32_ptr(x.data_ptr(), POW_Y, pw.data_ptr(), m, 0.6)
    add_exact = bool((out == a + b).all())   # the same IEEE operation, every element of the shard
    checksum = int(out.view(torch.int32).to(torch.int64).sum().item()) ^ int(pw.view(torch.int32).to(torch.int64).sum().item())
    sums = [checksum]
    audits = [(audit_over, audit_worst, add_exact)]
    if world > 1:
        t = torch.tensor([checksum], dtype=torch.int64, device=dev)
        allc = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allc, t)
        sums = [int(c.item()) for c in allc]
        t = torch.tensor([float(audit_over), audit_worst, float(add_exact)], dtype=torch.float64, device=dev)
        alla = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(alla, t)
        audits = [(int(v[0].item()), float(v[1].item()), bool(v[2].item())) for v in alla]
    ok &= all(o == 0 and ex for o, _, ex in audits)
    if rank == 0:
        try:
            import oracle
            orc = oracle.c_oracle()
            w = min(1 << 18, m)
            for start in (0, max(0, m - w)):
                ha, hb = orc.fill_uniform_f32(lo + start, w, 1, -1.0, 1.0), orc.fill_uniform_f32(lo + start, w, 2, -1.0, 1.0)
                hx = orc.fill_uniform_f32(lo + start, w, 3, 0.01, 100.0)
                ok &= bool(np.array_equal(out[start:start + w].cpu().numpy(), orc.elementwise("add", ha, [1], hb, [1], [w])))
                err = oracle.ulp_error_f32(pw[start:start + w].cpu().numpy(), orc.pow_ref_f32(hx, POW_Y))
                ok &= bool(err.max() <= 0.6)
        except Exception as e:  # the oracle is a checker; its absence must not hide the measurement
            ok = f"not checked: {e}"

    # ---- e2e: the same step through the C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        lib = smb.lib()
        nbytes = m * 4
        hp = [lib.smb_alloc(nbytes, smb.MEM_PINNED) for _ in range(5)]
        if all(hp):
            ha, hb, hx, hout, hpw = hp
            def host_view(ptr):
                return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=(m,))
            for dst, src in ((ha, a), (hb, b), (hx, x)):
                torch.from_numpy(host_view(dst)).copy_(src)  # D2H, outside the timed region

            def e2e_step():
                # host operands: the library stages H2D -> kernel -> D2H in overlapped slabs and
                # returns when the result is in host memory
                smb.contiguous_ptr(smb.OP_ADD, smb.F32, ha, hb, hout, m)
                smb.array_scalar_ptr(smb.OP_POW, smb.F32, hx, POW_Y, m, hpw)
            e2e_step()
            barrier()
            l1 = smb.launch_count()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            e2e_launches = smb.launch_count() - l1
            barrier()
            dt = max_ranks(dt) / args.e2e_steps
            # spot-check the host result
            hres = host_view(hout)
            e2e_ok = bool(np.array_equal(hres[:4096], out[:4096].cpu().numpy()))
            e2e = {"value": bytes_step_all / (dt * 1e-3) / 1e9, "unit": UNIT,
                   "h2d_bytes_per_step": 3 * nbytes * 1, "d2h_bytes_per_step": 2 * nbytes,
                   "ms_per_step": dt, "steps": args.e2e_steps, "launches_per_step": e2e_launches / args.e2e_steps,
                   "bytes_are": "per rank", "timer": "host perf_counter around synchronous C-ABI calls, max over ranks",
                   "host_binding": numa, "result_checked": e2e_ok}
            for p in hp:
                lib.smb_free(p)
        else:
            e2e = {"value": None, "unit": UNIT, "error": "pinned host allocation failed"}
            for p in hp:
                if p:
                    lib.smb_free(p)
        lib.smb_pool_trim()  # give the pinned staging memory back before the CPU baseline wants host RAM

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            r = cpu_reference(cpu_elems_default(args), 3, 1)
            cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            cpu["add_gbs"], cpu["pow_gbs"] = r["add_gbs"], r["pow_gbs"]
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD + (" (weak: 2^30 elements PER GPU)" if args.scaling == "weak" else ""), "elements": n, "pow_exponent": POW_Y,
                       "elements_per_gpu": m, "pow_kernel": "general exp2(y*log2 x), specialisation off",
                       "parallelism": f"sharded by flat output index range x{world}, one process per GPU, no data-path collective",
                       "streams": args.streams, "programmatic_dependent_launch": bool(args.pdl),
                       "l2": f"inputs larger than L2 ({m * 4 >> 20} MiB per array per GPU vs 126 MiB L2)",
                       "inputs": "splitmix64 counter generator of the flat index, produced in HBM",
                       "timing": f"median of {len(step_windows)} windows of {args.steps} steps, CUDA events on the launching stream, "
                                 "barrier + synchronize around every window, max over ranks"},
            "elements_per_s": n / (ms_step * 1e-3),
            "frac_hbm_measured": value / world / peak, "frac_hbm_nominal_8000": value / world / 8000.0,
            "spread": spread,
            "roofline": roofline, "roofline_pow": roofline_pow, "ops": extras, "sharded_configs": sharded,
            "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "verified": ok, "pow_audit_per_rank": [{"over_0.6_ulp": o, "max_ulp": wst, "add_bit_exact": ex} for o, wst, ex in audits],
            "shard_checksums": sums,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_smb(args)


if __name__ == "__main__":
    main()
