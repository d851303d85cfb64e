"""BASELINE.json configs[4]: incremental-SfM-shaped BA sweep (C cameras, 50*C points, 10 obs/point =>
500 obs/camera), LM-iteration latency on 1 B200 vs the compiled reference on the host cores.
Writes a markdown table to stdout (committed under profiles/)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from openmvg_b200 import ba, synth
import checkers as ck

sizes = [int(x) for x in sys.argv[1:]] or [50, 100, 200, 500, 1000, 2000]
print("| cams | pts | obs | GPU iters | GPU device ms | GPU ms/LM-iter | GPU e2e ms | PCG iters | final cost (GPU) | ref iters | ref minimizer s | ref Adjust s | ref ms/LM-iter | final cost (ref) | rel diff | speed-up (minimizer/device) | speed-up (Adjust/e2e) |")
print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for C in sizes:
    s = synth.ba_scene(C, 50 * C, 10)
    ctx = ba.BAContext(s); ctx.run(); ctx.reset()
    g = ctx.run(); ctx.close()
    e2e = 1e9
    for _ in range(2):                                     # best of two: the first call at a new size may pay cudaMalloc
        t0 = time.perf_counter(); e = ba.solve(s); e2e = min(e2e, time.perf_counter() - t0)
    best = None
    for th in (8, 32):
        r = ck.ref_ba_adjust(s, threads=th)
        if best is None or r["minimizer_s"] < best["minimizer_s"]:
            best = r
    r = best
    rel = abs(g["final_cost"] - r["final_cost"]) / r["final_cost"]
    print(f"| {C} | {50*C} | {500*C} | {g['iterations']} | {g['device_ms']:.2f} | {g['device_ms']/g['iterations']:.2f} | {e2e*1e3:.1f} | {g['pcg_iterations']} | {g['final_cost']:.6f} | "
          f"{r['iterations']} | {r['minimizer_s']:.3f} | {r['wall_s']:.3f} | {1e3*r['minimizer_s']/r['iterations']:.1f} | {r['final_cost']:.6f} | {rel:.1e} | "
          f"{1e3*r['minimizer_s']/g['device_ms']:.0f}x | {r['wall_s']/e2e:.0f}x |", flush=True)
