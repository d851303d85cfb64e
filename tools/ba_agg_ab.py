"""A/B of the aggregate size of the coarse space (OMVG_BA_AGG, cameras per aggregate): config 2, 500 and 2000 cameras."""
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
print(json.dumps(dict(device_ms=round(best["device_ms"], 3), iters=best["iterations"], pcg=best["pcg_iterations"])))
"""
for cfg in ((1000, 100000, 10), (500, 25000, 10), (2000, 100000, 10), (200, 10000, 10)):
    for agg in ("6", "7", "8", "9", "10", "12"):
        e = dict(os.environ); e["OMVG_BA_AGG"] = agg
        p = subprocess.run([sys.executable, "-c", CHILD % (ROOT, *cfg)], capture_output=True, text=True, env=e, timeout=600)
        print(cfg[0], "cams, aggregate target", agg, p.stdout.strip()[-120:] or p.stderr[-300:], flush=True)
