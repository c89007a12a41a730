"""Shared by the tools/*_sweep.py scripts: an explicit stream and a timed window that samples the SM clock (NVML)
while the window runs -- on a power-capped shared box the clock moves between 1.5 and 1.97 GHz within one script,
and the issue-bound kernels move with it."""
import torch

try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())

    def sm_clock():
        return pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:
    def sm_clock():
        return None

stream = torch.cuda.Stream()
sp = stream.cuda_stream


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    clocks = []
    while not e1.query():
        c = sm_clock()
        if c is not None:
            clocks.append(c)
    e1.synchronize()
    clocks.sort()
    return e0.elapsed_time(e1) / reps, (clocks[len(clocks) // 2] if clocks else None)
