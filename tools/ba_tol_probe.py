"""Ad-hoc probe: sensitivity of the BA result to the PCG tolerance (parity vs the oracle / golden)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from openmvg_b200 import ba, synth
import checkers as ck
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_outputs.json")))
cases = [(c["name"], synth.ba_scene(**c["scene"]), c["opts"], c["final_cost"], c["iterations"]) for c in gold["ba"]]
for tol in (1e-10, 1e-8, 1e-6, 1e-4):
    worst = 0.0; bad_it = []; tot_ms = 0.0; big = None
    for name, s, opts, fc, its in cases:
        g = ba.solve(s, pcg_tolerance=tol, **opts)
        ret = ck.oracle_ba_cost(s, g["poses"], g["intrinsics"], g["points"], use_loss=opts.get("use_loss", 1))
        rel = abs(ret - fc) / fc
        worst = max(worst, rel); tot_ms += g["device_ms"]
        if g["iterations"] != its: bad_it.append((name, g["iterations"], its))
        if name.startswith("config2"): big = (g["device_ms"], g["pcg_iterations"], rel)
    print("tol %.0e worst rel %.2e iteration mismatches %s total device ms %.1f config2 (ms, pcg its, rel) %s" % (tol, worst, bad_it, tot_ms, big), flush=True)
