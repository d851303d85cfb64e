"""Ad-hoc probe: end-to-end omvg_ba_solve (host buffers) with OMVG_BA_TIMING phase prints."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmvg_b200 import ba, synth
cfg = tuple(int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1000,100000,10").split(","))
s = synth.ba_scene(*cfg)
if os.environ.get('PIN'):
    import torch, numpy as np
    s = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) and k not in ('gt_R', 'gt_C', 'gt_dist') else v) for k, v in s.items()}
for i in range(4):
    t = time.time(); g = ba.solve(s); dt = time.time() - t
    print("solve %d wall %.2f ms device %.2f ms final %.9e its %d pcg %d" % (i, dt * 1e3, g["device_ms"], g["final_cost"], g["iterations"], g["pcg_iterations"]), flush=True)
