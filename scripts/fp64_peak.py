"""Measures the DFMA peak of the B200 this runs on (xm_fp64_peak) and writes profiles/fp64_peak.json --
the FP64 roofline denominator bench.py --config c4 quotes."""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from vif_b200 import _lib  # noqa: E402

lib = _lib.load()
v = C.c_double(0.0)
err = C.create_string_buffer(256)
vals = []
for _ in range(3):
    assert lib.xm_fp64_peak(-1, 10, C.byref(v), err, 256) == 0, err.value
    vals.append(v.value)
p = torch.cuda.get_device_properties(0)
out = {"tflops": max(vals), "runs": vals, "gpu": p.name, "sms": p.multi_processor_count,
       "nominal_tflops": p.multi_processor_count * 64 * 2 * 1.965e-3,
       "how": "xm_fp64_peak: best of 10 launches x 3, 8 independent DFMA chains x 4096 per thread, "
              "8 CTAs x 256 threads per SM, CUDA events (vif_b200/csrc/xm_peak.cu)",
       "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "fp64_peak.json"), "w"), indent=1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp64_peak.json"), "w"), indent=1)
print(json.dumps(out))
