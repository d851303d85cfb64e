"""Whole-M1 golden (200 images x 5000, all 19 900 pairs) and a 256-pair seeded sample of M2 (1000 x 5000),
produced by THE REFERENCE ITSELF: Matcher_Regions(0.8, BRUTE_FORCE_L2)::Match compiled from /root/reference
(oracle/_ref/libref_match.so).  Stored per pair: number of matches (u16) and the FNV-1a of the (i, j) list
(u64); plus the hash of the whole output.  Inputs are regenerated from their seeds (synth.descriptor_collection)
at test time.  Run in the build container (about 10 min on 8 cores):  python tests/golden/make_golden_m1m2.py"""
import ctypes
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers as ck          # noqa: E402
from openmvg_b200 import synth  # noqa: E402


def run(descs, pi, pj, chunk=400):
    counts = np.zeros(len(pi), np.uint16); hashes = np.zeros(len(pi), np.uint64); total = 0
    t0 = time.time()
    for a in range(0, len(pi), chunk):
        b = min(a + chunk, len(pi))
        used = sorted(set(pi[a:b].tolist()) | set(pj[a:b].tolist()))
        dense = {v: k for k, v in enumerate(used)}
        sub = [descs[v] for v in used]
        spi = np.array([dense[int(v)] for v in pi[a:b]], np.uint32); spj = np.array([dense[int(v)] for v in pj[a:b]], np.uint32)
        off, ij = ck.ref_match_collection(sub, spi, spj, 0.8)
        c, h = ck.per_pair_digest(off, ij)
        counts[a:b] = c; hashes[a:b] = h; total += len(ij)
        print(f"  pairs {b}/{len(pi)}  matches {total}  {time.time() - t0:.0f}s", flush=True)
    return counts, hashes


def main():
    out = {}
    m1 = synth.descriptor_collection(200, 5000, seed=1000)
    pi, pj = synth.exhaustive_pairs(200)
    c, h = run(m1, pi, pj)
    out["m1_counts"], out["m1_fnv"] = c, h
    m2 = synth.descriptor_collection(1000, 5000, seed=1000)
    spi, spj = synth.sampled_pairs(1000, 256, seed=5)
    c, h = run(m2, spi, spj)
    out["m2_pair_i"], out["m2_pair_j"], out["m2_counts"], out["m2_fnv"] = spi, spj, c, h
    np.savez_compressed(os.path.join(HERE, "reference_m1_m2.npz"), **out)
    print("M1 matches", int(out["m1_counts"].sum()), "M2 sample matches", int(out["m2_counts"].sum()))


if __name__ == "__main__":
    main()
