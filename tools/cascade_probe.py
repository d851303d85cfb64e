"""Ad-hoc probe: cascade hashing on the GPU vs the compiled reference's Cascade_Hashing_Matcher_Regions on the host."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from openmvg_b200 import matching, synth
import checkers as ck
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 40
descs = synth.descriptors(n_img, [5000] * n_img, seed=1)
pi, pj = synth.exhaustive_pairs(n_img)
P, S = ck.cascade_projections()
ctx = matching.MatchContext(); ctx.load(descs)
for rep in range(3):
    t = time.perf_counter(); ctx.cascade_prepare(P, S, None); tp = time.perf_counter() - t
    t = time.perf_counter(); ctx.cascade_run(pi, pj, 0.8); off, ij = ctx.fetch(); tr = time.perf_counter() - t
    ms, n = ctx.kernel_time()
    print("rep %d: prepare %.2f ms, run+fetch %.2f ms (query kernel %.2f ms), %d pairs, %d matches, %.3e desc-pairs/s equivalent" % (
        rep, tp * 1e3, tr * 1e3, ms, len(pi), len(ij), len(pi) * 25e6 / tr), flush=True)
t = time.perf_counter(); ctx.run(pi, pj, 0.8); boff, bij = ctx.fetch(); tb = time.perf_counter() - t
print("brute force run+fetch %.2f ms, %d matches" % (tb * 1e3, len(bij)))
a = set(); b = set()
for p in range(len(pi)):
    a |= {(p,) + tuple(r) for r in ij[int(off[p]):int(off[p + 1])]}; b |= {(p,) + tuple(r) for r in bij[int(boff[p]):int(boff[p + 1])]}
print("recall vs brute force %.4f, extra %.4f" % (len(a & b) / max(1, len(b)), len(a - b) / max(1, len(a))))
ctx.close()
if ck.have_ref_match() and n_img <= 60:
    t = time.perf_counter(); roff, rij = ck.ref_cascade_collection(descs, pi, pj, 0.8); tr = time.perf_counter() - t
    print("reference Cascade_Hashing_Matcher_Regions (host, OpenMP): %.2f s, %d matches" % (tr, len(rij)))
    same = tot = 0
    for p in range(len(pi)):
        x = set(map(tuple, ij[int(off[p]):int(off[p + 1])])); y = set(map(tuple, rij[int(roff[p]):int(roff[p + 1])]))
        same += len(x & y); tot += len(x | y)
    print("identical to the reference: %d of %d" % (same, tot))
