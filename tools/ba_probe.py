"""Ad-hoc probe (not a test): run the GPU BA on a scene and print the summary next to the oracle's."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from openmvg_b200 import ba, synth
import checkers as ck

cfgs = [tuple(int(x) for x in a.split(",")) for a in sys.argv[1:]] or [(10, 500, 4)]
for cfg in cfgs:
    s = synth.ba_scene(*cfg)
    t = time.time(); ctx = ba.BAContext(s); tc = time.time() - t
    t = time.time(); g = ctx.run(verbose=1); tr = time.time() - t
    print(cfg, "create %.3fs run %.3fs" % (tc, tr), {k: g[k] for k in ("initial_cost", "final_cost", "iterations", "successful_steps", "lm_steps", "termination", "pcg_iterations", "kernel_launches", "device_ms", "jacobian_ms", "jacobian_launches")})
    ctx.reset(); t = time.time(); g = ctx.run(); print("  second run %.3fs device_ms %.2f" % (time.time() - t, g["device_ms"]))
    ctx.close()
    if cfg[0] <= 200:
        o = ck.oracle_ba_solve(s)
        print("  oracle", o["initial_cost"], o["final_cost"], o["iterations"], "rel", abs(o["final_cost"] - g["final_cost"]) / o["final_cost"])
    if ck.have_ref_ba() and os.environ.get("REF", "0") == "1":
        r = ck.ref_ba_adjust(s)
        print("  ref   ", r["initial_cost"], r["final_cost"], r["iterations"], "wall %.2fs minimizer %s" % (r["wall_s"], r["minimizer_s"]), "rel", abs(r["final_cost"] - g["final_cost"]) / r["final_cost"])
