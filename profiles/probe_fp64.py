"""FP64 peak of the B200 under test (SURVEY 6 / BASELINE.md 2 ask for a DFMA microbenchmark): kin_probe_fp64 of
libkin_b200 (8 independent DFMA chains per thread, 8 x 256 threads per SM) with the SM clock sampled meanwhile.
    python profiles/probe_fp64.py        (on the GPU box)"""
import ctypes as C
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kinematics_jl_b200 import lib as L  # noqa: E402

import torch  # noqa: E402,F401  (creates the CUDA context the way the product does)

torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                     stdout=subprocess.PIPE, text=True)
res = []
for _ in range(3):
    tf, per, mhz = C.c_double(), C.c_double(), C.c_double()
    L.check(L.lib().kin_probe_fp64(C.byref(tf), C.byref(per), C.byref(mhz)))
    res.append((tf.value, per.value, mhz.value))
p.terminate()
rows = [l.split(",") for l in p.stdout.read().strip().splitlines()]
clk = sorted(float(r[0]) for r in rows if len(r) == 2 and float(r[1]) > 300) or [float("nan")]
best = max(res)
out = {"fp64_tflops": best[0], "dfma_per_clk_per_sm_at_nominal_clock": best[1], "nominal_sm_mhz": best[2],
       "sm_mhz_median_under_load": clk[len(clk) // 2],
       "dfma_per_clk_per_sm_at_sampled_clock": best[1] * best[2] / clk[len(clk) // 2] if clk[0] == clk[0] else None}
print(json.dumps(out))
