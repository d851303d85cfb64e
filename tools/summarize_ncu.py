"""Turn the ncu outputs brought back in gpurun_out/ into the small, committed summaries under profiles/.
usage: python tools/summarize_ncu.py launches <csv> <out.md> | full <rep> <out.md>"""
import collections
import csv
import io
import subprocess
import sys

FULL_KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
             "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
             "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
             "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
             "sm__pipe_tensor_subpipe_imma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
             "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.per_cycle_elapsed",
             "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.peak_sustained",
             "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
             "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
             "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def launches(path, out):
    with open(path) as fh:
        lines = [l for l in fh if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO("".join(lines))):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        k = r["Kernel Name"].split("(")[0]
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and serialised: compare SHARES.\n\n")
        f.write(f"total kernel time {tot / 1e6:.3f} ms over {sum(v[0] for v in agg.values())} launches\n\n| kernel | launches | total ms | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / v[0] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |\n")


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    hdr, units = rd[0], rd[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n")
        for row in rd[2:]:
            d = dict(zip(hdr, row)); u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in FULL_KEYS:
                for h in hdr:
                    if h == k or h.endswith("." + k):
                        f.write(f"| {h} | {d[h]} | {u[h]} |\n")
            f.write("\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
