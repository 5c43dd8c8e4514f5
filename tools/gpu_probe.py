"""Box probe: pipe micro-benchmarks + a coarse matcher timing (wall clock incl. copies)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api

ctx = api.Context(0)
names = ["popc32", "lop3", "fp64_fma", "fp32_fma", "iadd3"]
out = {}
for i, n in enumerate(names):
    out[n + "_Gops"] = ctx.microbench(i) / 1e9
rng = np.random.default_rng(0)
for n in (2000, 10000, 20000):
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    for norm in (api.NORM_HAMMING, api.NORM_HAMMING2):
        m = api.BFMatcher(norm, True, ctx=ctx)
        m.match(q, t)
        t0 = time.perf_counter()
        for _ in range(5):
            m.match(q, t)
        dt = (time.perf_counter() - t0) / 5
        out[f"match_{n}_norm{norm}_ms"] = dt * 1e3
        out[f"match_{n}_norm{norm}_Gpairs_s"] = n * n / dt / 1e9
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
