"""Small-scene LM-iteration latency (BASELINE configs[4] low end): device ms per LM iteration for a few sizes,
default vs OMVG_BA_PCG_SMALL=0 (the 148-CTA PCG).  One subprocess per setting."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json
sys.path.insert(0, %r)
from openmvg_b200 import ba, synth
out = []
for C in (10, 20, 50, 100):
    s = synth.ba_scene(C, 50 * C, 10 if C >= 10 else 4)
    ctx = ba.BAContext(s)
    for _ in range(3): ctx.reset(); ctx.run()
    best = None
    for _ in range(5):
        ctx.reset(); r = ctx.run()
        if best is None or r["device_ms"] < best["device_ms"]: best = r
    ctx.close()
    out.append((C, round(best["device_ms"], 3), best["iterations"], round(best["device_ms"] / best["iterations"], 3), best["pcg_iterations"], best["kernel_launches"], best["final_cost"]))
print(json.dumps(out))
'''
for name, env in (("default (single-CTA PCG up to 48 poses)", {}), ("148-CTA PCG", {"OMVG_BA_PCG_SMALL": "0"}), ("single-CTA up to 100", {"OMVG_BA_PCG_SMALL": "100"})):
    e = dict(os.environ); e.update(env)
    p = subprocess.run([sys.executable, "-c", CHILD % ROOT], capture_output=True, text=True, env=e, timeout=600)
    print(name, p.stdout.strip() or p.stderr[-800:], flush=True)
