import time, sys, os
sys.path.insert(0, "/root/repo")
import torch
from bench import ClockSampler
s = ClockSampler(0); torch.cuda.init(); s.start()
for i in range(5):
    t=time.perf_counter(); s.sample(); print("sample ms", round(1e3*(time.perf_counter()-t),3))
import pynvml
h = s.h
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), ("maxclock", lambda: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)), ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))):
    t=time.perf_counter(); fn(); print(name, "ms", round(1e3*(time.perf_counter()-t),3))
