"""A/B of the coarse-operator refresh period (OMVG_BA_COARSE_EVERY): config 2, 500 and 2000 cameras."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json
sys.path.insert(0, %r)
from openmvg_b200 import ba, synth
C, Pn, K = %d, %d, %d
s = synth.ba_scene(C, Pn, K)
ctx = ba.BAContext(s)
for _ in range(3): ctx.reset(); ctx.run()
best = None
for _ in range(5):
    ctx.reset(); r = ctx.run()
    if best is None or r["device_ms"] < best["device_ms"]: best = r
print(json.dumps(dict(device_ms=round(best["device_ms"], 3), iters=best["iterations"], pcg=best["pcg_iterations"], cost=best["final_cost"])))
"""
for cfg in ((1000, 100000, 10), (500, 25000, 10), (2000, 100000, 10)):
    for every in ("3", "4", "5", "8"):
        e = dict(os.environ); e["OMVG_BA_COARSE_EVERY"] = every
        p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, *cfg)], capture_output=True, text=True, env=e, timeout=600)
        print(cfg[0], "cams, refresh every", every, p.stdout.strip()[-200:], flush=True)
