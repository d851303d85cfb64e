"""Small inputs for compute-sanitizer runs (racecheck / memcheck): one tiny BA solve on the direct-solve path, one on the
PCG path, one MATCH run through the digit-slice kernels, one geometric-filter launch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openmvg_b200 import ba, matching, synth, geometry
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "ba"):
    s = synth.ba_scene(8, 200, 5)
    print("dense", ba.solve(s)["final_cost"])
    os.environ["OMVG_BA_DENSE_MAX"] = "0"
    s = synth.ba_scene(24, 600, 6)
    print("pcg", ba.solve(s)["final_cost"])
if which in ("all", "match"):
    d = synth.descriptors(3, [300, 520, 260], seed=1)
    pi, pj = synth.exhaustive_pairs(3)
    ctx = matching.MatchContext(0); ctx.load(d); ctx.run(pi, pj, 0.8); off, ij = ctx.fetch(); print("match", len(ij), ctx.kernel_variant()); ctx.close()
if which in ("all", "geom"):
    xs = [synth.two_view_matches(200, 0.3, 0.5, seed=k) for k in range(3)] + [synth.two_view_matches(150, 0.3, 0.5, seed=9, planar=True)]
    off = np.cumsum([0] + [len(x[0]) for x in xs]).astype(np.uint64)
    xI = np.concatenate([x[0] for x in xs]); xJ = np.concatenate([x[1] for x in xs])
    sizes = np.tile(np.array([1000, 1000, 1000, 1000], np.int32), (len(xs), 1))
    print("geom F", [len(r["inliers"]) for r in geometry.fundamental_acransac(off, xI, xJ, sizes, 4.0, 256)])
    print("geom H", [len(r["inliers"]) for r in geometry.homography_acransac(off, xI, xJ, sizes, 4.0, 256)])
