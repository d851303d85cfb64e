"""Ad-hoc probe: one small scene (argv: cameras), a few solves; run under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmvg_b200 import ba, synth
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10
s = synth.ba_scene(C, 50 * C, 10)
ctx = ba.BAContext(s)
for _ in range(2): ctx.reset(); ctx.run()
best = None
for _ in range(5):
    ctx.reset(); t = time.perf_counter(); r = ctx.run(); r["wall_ms"] = (time.perf_counter() - t) * 1e3
    if best is None or r["device_ms"] < best["device_ms"]: best = r
print(C, {k: best[k] for k in ("device_ms", "wall_ms", "iterations", "lm_steps", "pcg_iterations", "kernel_launches")})
ctx.close()
