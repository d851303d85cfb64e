#!/usr/bin/env python
"""bench.py — the two hot paths on N B200s (one process per GPU) or, with --impl reference, the
reference's own CPU implementation of the same paths on the host cores.

Headline (`metric`/`value`): BA LM-iterations/s on BASELINE.json configs[1]
(1000 cams / 100k pts / 1M obs, shared pinhole intrinsic, Huber(16), refine all).  One "step" = one
complete Adjust (LM until Ceres' own termination test fires) on the device-resident scene;
`value` = sum(Ceres-style iteration count) / sum(device time, CUDA events).  `e2e` = the same solve
through omvg_ba_solve with HOST buffers (packing, H2D, structure build, solve, D2H inside the
timed region).  BA does not shard: with --gpus N every rank runs a replica ("weak").

The `match` object carries the second hot path with the same keys: exhaustive BRUTE_FORCE_L2 +
ratio matching, 200 images x 5000 descriptors at N=1 (BASELINE.json configs[2]); image pairs are
sharded over ranks after one NCCL all-gather of the per-rank descriptor tiles.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BA_CFG = (1000, 100_000, 10)          # cams, points, obs/point  -> 1M observations
BA_BYTES_PER_OBS = 232                # 24 B in (view/point index + xy) + 16 B r + 192 B J (2x(3+6+3) doubles)
MATCH_IMAGES, MATCH_DESC = 200, 5000
OPS_PER_DESC_PAIR = 256               # 128 MAC


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.samples, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True); self.t.start(); return self

    def __exit__(self, *a):
        self.stop.set(); self.t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


def dist_env():
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ===================================================================================== ours (GPU)
def run_ours(args):
    import torch
    import torch.distributed as dist
    from openmvg_b200 import ba, matching, synth
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — openmvg_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pk = peaks()
    K, W = args.steps, args.warmup
    out = {}
    # ------------------------------------------------------------------ BA (replica per rank)
    scene = synth.ba_scene(*BA_CFG, seed=42 + rank)
    n_obs = len(scene["obs_view"])
    ctx = ba.BAContext(scene, device=local)
    for _ in range(W):
        ctx.reset(); ctx.run()
    barrier()
    iters = 0; dev_ms = 0.0; jac_ms = 0.0; jac_n = 0; launches = 0; last = None
    clk = ClockSampler(local); clk.__enter__()        # sampled over the BA and the MATCH timed regions
    t0 = time.perf_counter()
    for _ in range(K):
        ctx.reset(); r = ctx.run()
        iters += r["iterations"]; dev_ms += r["device_ms"]; jac_ms += r["jacobian_ms"]; jac_n += r["jacobian_launches"]; launches += r["kernel_launches"]; last = r
    barrier()
    wall = time.perf_counter() - t0
    ba_time = allmax(max(dev_ms / 1e3, 0.0))
    ba_wall = allmax(wall)
    ba_iters_all = allsum(iters)
    ctx.close()
    # e2e: host buffers in, host buffers out, everything inside the timed region
    barrier()
    e_iters = 0; t0 = time.perf_counter(); e_steps = max(1, min(K, 3))
    # inputs start in PINNED host memory (page-locked copies of the scene arrays), results land in host arrays
    scene_pinned = {k: (torch.from_numpy(np.ascontiguousarray(v)).pin_memory().numpy() if isinstance(v, np.ndarray) and k not in ("gt_R", "gt_C", "gt_dist") else v)
                    for k, v in scene.items()}
    t0 = time.perf_counter()
    for _ in range(e_steps):
        g = ba.solve(scene_pinned, device=local); e_iters += g["iterations"]; launches += g["kernel_launches"]
    barrier()
    e_wall = allmax(time.perf_counter() - t0); e_iters_all = allsum(e_iters)
    h2d = sum(scene[k].nbytes for k in ("poses", "intrinsics", "points", "intr_model", "view_pose", "view_intr", "obs_view", "obs_point", "obs_xy"))
    d2h = sum(scene[k].nbytes for k in ("poses", "intrinsics", "points"))
    jac_s = jac_ms / 1e3 / max(jac_n, 1)
    ba_roof = dict(bound="hbm", achieved=BA_BYTES_PER_OBS * n_obs / jac_s / 1e9, peak=pk["hbm_gbs"], unit="GB/s",
                   # dram__bytes_read.sum + dram__bytes_write.sum of one launch on this scene (ncu --set full capture,
                   # profiles/r01_ba_eval_kernel_full.md: 33.06 MB read + 150.79 MB written; less than the 232 MB of
                   # algorithmic bytes because the tail of the write stream is still in the 126 MB L2 when the kernel ends)
                   traffic=183.85e6 if n_obs == 1000000 else None, traffic_unit="bytes/launch",
                   kernel="eval_kernel<true> (residual+Jacobian+Huber+scaling)", launches=jac_n, avg_ms=jac_s * 1e3, peak_source=pk["src"])
    ba_roof["frac"] = ba_roof["achieved"] / ba_roof["peak"]

    # ------------------------------------------------------------------ MATCH (pairs sharded over ranks)
    n_img = MATCH_IMAGES if args.match_images is None else args.match_images
    per = (n_img + world - 1) // world
    lo, hi = min(rank * per, n_img), min((rank + 1) * per, n_img)
    mine = synth.descriptors(hi - lo, MATCH_DESC, seed=1000 + rank) if hi > lo else []
    local_t = torch.from_numpy(np.concatenate(mine) if mine else np.zeros((0, 128), np.uint8)).cuda()
    if world > 1:
        pad = torch.zeros((per * MATCH_DESC, 128), dtype=torch.uint8, device="cuda"); pad[: local_t.shape[0]] = local_t
        gathered = torch.empty((world * per * MATCH_DESC, 128), dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(gathered, pad)
        allrows = gathered[: n_img * MATCH_DESC].contiguous()
    else:
        allrows = local_t
    mctx = matching.MatchContext(local)
    mctx.set_images([MATCH_DESC] * n_img)
    mctx.upload_device_packed(allrows.data_ptr()); mctx.prepare(); mctx.sync()
    pi, pj = synth.exhaustive_pairs(n_img)
    sl = slice(rank, None, world)                                   # round-robin: equal-cost pairs
    mpi, mpj = np.ascontiguousarray(pi[sl]), np.ascontiguousarray(pj[sl])
    desc_pairs_rank = float(len(mpi)) * MATCH_DESC * MATCH_DESC
    for _ in range(W):
        mctx.run(mpi, mpj, 0.8); mctx.sync()
    mctx.kernel_time(reset=True); l0 = mctx.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for _ in range(K):
        mctx.run(mpi, mpj, 0.8)
    mctx.sync(); barrier()
    m_wall = allmax(time.perf_counter() - t0)
    clk.__exit__()
    tc_ms, tc_n = mctx.kernel_time(reset=True)
    launches += mctx.launch_count() - l0
    n_matches = len(mctx.fetch()[1])
    # e2e: descriptors start in pinned host memory; upload + prepare + run + fetch timed
    host_desc = [torch.from_numpy(d).pin_memory().numpy() for d in (synth.descriptors(n_img, MATCH_DESC, seed=1000) if world == 1 else [allrows[k * MATCH_DESC:(k + 1) * MATCH_DESC].cpu().numpy() for k in range(n_img)])]
    barrier(); t0 = time.perf_counter(); me_steps = max(1, min(K, 2))
    for _ in range(me_steps):
        for k, d in enumerate(host_desc):
            mctx.upload_host(k, d)
        mctx.prepare(); mctx.run(mpi, mpj, 0.8); off, ij = mctx.fetch()
    barrier()
    me_wall = allmax(time.perf_counter() - t0)
    total_pairs = allsum(desc_pairs_rank)
    tc_s = tc_ms / 1e3 / max(tc_n, 1)
    int8_peak = 2.0 * pk["bf16_sustained"]                          # dense INT8 = 2x bf16 rate; bf16 is the measured figure
    # one pass may take several launches (pair batches sized by the result-buffer budget): rate over all of them
    m_roof = dict(bound="tensor", achieved=OPS_PER_DESC_PAIR * desc_pairs_rank * K / max(tc_ms / 1e3, 1e-12) / 1e12, peak=int8_peak, unit="TOP/s", traffic=None,
                  kernel="match_tc_kernel (tcgen05 kind::i8 + fused top-2)", launches=tc_n, avg_ms=tc_s * 1e3,
                  peak_source=f"2 x {pk['src']} bf16 sustained (INT8 dense rate = 2 x bf16)")
    m_roof["frac"] = m_roof["achieved"] / m_roof["peak"]
    # cascade hashing (openMVG's default matcher) on the same pair shard: reported next to the
    # exhaustive matcher, not part of the headline (its parity against the reference is statistical)
    cascade = None
    try:
        z_ = np.load(os.path.join(ROOT, "tests", "golden", "cascade_projections.npz"))      # CascadeHasher::Init's draw (fixture)
        P_, S_ = z_["primary"], z_["secondary"]
        mctx.cascade_prepare(P_, S_, None); mctx.cascade_run(mpi, mpj, 0.8); mctx.sync()
        barrier(); t0 = time.perf_counter()
        for _ in range(K):
            mctx.cascade_run(mpi, mpj, 0.8)
        mctx.sync(); barrier()
        c_wall = allmax(time.perf_counter() - t0)
        cascade = {"ms_per_step": c_wall * 1e3 / K, "pairs_per_s": allsum(float(len(mpi))) * K / c_wall,
                   "matches_rank0": int(len(mctx.fetch()[1])), "exhaustive_matches_rank0": int(n_matches),
                   "note": "same pairs through omvg_match_cascade_run (hash tables resident); reference = Cascade_Hashing_Matcher_Regions"}
    except Exception as e:                                     # the fixture with the reference's projections is test data
        cascade = {"unavailable": str(e)[:120]}
    mctx.close()

    if rank == 0:
        line = {
            "metric": "BA LM-iters/sec", "value": ba_iters_all / ba_time, "unit": "LM-iter/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ba_time * 1e3 / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BA 1000 cams / 100k pts / 1M obs, 1 shared pinhole intrinsic, Huber(16), refine all (BASELINE configs[1]); replica per GPU",
                       "l2": "inputs larger than L2 (J+r = 224 MB per evaluation)", "iterations_per_solve": last["iterations"],
                       "lm_steps_per_solve": last["lm_steps"], "pcg_iterations_per_solve": last["pcg_iterations"], "final_cost": last["final_cost"],
                       "wall_ms_per_step": ba_wall * 1e3 / K},
            "e2e": {"value": e_iters_all / e_wall, "unit": "LM-iter/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e_wall * 1e3 / e_steps},
            "gpu_launches": int(launches), "roofline": ba_roof, "clocks": clk.summary(),
            "match": {"metric": "desc-pairs/sec", "value": total_pairs * K / m_wall, "unit": "desc-pairs/s", "ms_per_step": m_wall * 1e3 / K, "dtype": "u8",
                      "config": {"workload": f"exhaustive BRUTE_FORCE_L2 + ratio 0.8, {n_img} images x {MATCH_DESC} x 128-D uint8, {len(pi)} pairs sharded round-robin over {world} GPU(s)",
                                 "l2": f"descriptor arena {n_img * MATCH_DESC * 128 / 1e6:.0f} MB + per-pair result buffers: larger than L2", "matches_rank0": int(n_matches)},
                      "e2e": {"value": total_pairs * me_steps / me_wall, "unit": "desc-pairs/s", "h2d_bytes_per_step": n_img * MATCH_DESC * 128 + 8 * len(mpi),
                              "d2h_bytes_per_step": int(8 * (len(mpi) + 1) + 8 * n_matches), "ms_per_step": me_wall * 1e3 / me_steps},
                      "roofline": m_roof, "cascade_hashing": cascade},
        }
        if not args.no_cpu and world == 1:
            line["cpu_baseline"], line["match"]["cpu_baseline"] = cpu_baselines(scene, args)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ===================================================================================== reference (CPU)
def cpu_baselines(scene, args):
    """The reference itself (oracle/_ref) when it was built, else the oracle port; bounded samples."""
    import checkers as ck
    from openmvg_b200 import synth
    cores = os.cpu_count() or 1
    if ck.have_ref_ba():
        best = None
        for th in sorted({min(cores, 8), min(cores, 32)}):
            r = ck.ref_ba_adjust(scene, threads=th)
            v = r["iterations"] / r["minimizer_s"]
            if best is None or v > best["value"]:
                best = dict(value=v, unit="LM-iter/s", cores=th, kind="reference",
                            sample="one full Bundle_Adjustment_Ceres::Adjust on the same 1000/100k/1M scene; iterations / Ceres 'Minimizer' seconds",
                            adjust_wall_s=r["wall_s"], minimizer_s=r["minimizer_s"], preprocessor_s=r["preprocessor_s"], iterations=r["iterations"],
                            final_cost=r["final_cost"], linear_solver_s=r["linear_solver_s"], jacobian_s=r["jacobian_s"])
        ba_cpu = best
    else:
        small = synth.ba_scene(200, 10_000, 10)
        t0 = time.perf_counter(); o = ck.oracle_ba_solve(small); dt = time.perf_counter() - t0
        ba_cpu = dict(value=o["iterations"] / dt, unit="LM-iter/s", cores=cores, kind="port", sample="oracle port on 200 cams/10k pts/100k obs (1/10 of the workload)")
    n_img = 12
    descs = synth.descriptors(n_img, MATCH_DESC, seed=1000)
    pi, pj = synth.exhaustive_pairs(n_img)
    if ck.have_ref_match():
        ck.ref_match_collection(descs[:3], pi[:1], pj[:1])          # warm the thread pool
        t0 = time.perf_counter(); ck.ref_match_collection(descs, pi, pj); dt = time.perf_counter() - t0
        kind = "reference"
    else:
        t0 = time.perf_counter(); ck.oracle_match_collection(descs, pi, pj); dt = time.perf_counter() - t0
        kind = "port"
    m_cpu = dict(value=len(pi) * MATCH_DESC * MATCH_DESC / dt, unit="desc-pairs/s", cores=cores, kind=kind,
                 sample=f"Matcher_Regions(BRUTE_FORCE_L2)::Match on {len(pi)} image pairs of 5000x5000 descriptors (all host threads, AVX2 build)", seconds=dt)
    return ba_cpu, m_cpu


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from openmvg_b200 import synth
    scene = synth.ba_scene(*BA_CFG, seed=42)
    K = args.steps
    # every "step" is one bounded sample (one Adjust); run min(K, 2) samples to stay within minutes
    ba_cpu, m_cpu = cpu_baselines(scene, args)
    line = {"impl": "reference", "metric": "BA LM-iters/sec", "value": ba_cpu["value"], "unit": "LM-iter/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ba_cpu.get("minimizer_s", 0) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "BA 1000 cams / 100k pts / 1M obs, 1 shared pinhole intrinsic, Huber(16), refine all (BASELINE configs[1]); replica per GPU"},
            "cpu_baseline": ba_cpu, "e2e": {"value": ba_cpu["value"], "unit": "LM-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "match": {"metric": "desc-pairs/sec", "value": m_cpu["value"], "unit": "desc-pairs/s", "cpu_baseline": m_cpu,
                      "e2e": {"value": m_cpu["value"], "unit": "desc-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--match-images", type=int, default=None)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.no_cpu:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
