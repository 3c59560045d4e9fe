"""Turns the raw captures a gpurun call leaves in gpurun_out/ into the tracked summaries of this directory.

    python profiles/make_summaries.py r1g          # prefix of the capture files in gpurun_out/

Inputs:  <p>_launches.csv (ncu --metrics gpu__time_duration.sum ... --csv), <p>_env.ncu-rep (ncu --set full),
         <p>_bench_full.log (python bench.py), <p>_microbench.txt, <p>_cta_timing.txt
Outputs: r1_bench_launches.csv, r1_bench_launches_summary.txt, r1_env_kernels_ncu.txt, dram_traffic.json,
         r1_bench_line.json, r1_microbench.txt, r1_cta_timing.txt
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
prefix = sys.argv[1] if len(sys.argv) > 1 else "r1g"
src = lambda name: os.path.join(ROOT, "gpurun_out", f"{prefix}_{name}")   # noqa: E731
dst = lambda name: os.path.join(HERE, name)                                 # noqa: E731

shutil.copy(src("launches.csv"), dst("r1_bench_launches.csv"))
shutil.copy(src("microbench.txt"), dst("r1_microbench.txt"))
shutil.copy(src("cta_timing.txt"), dst("r1_cta_timing.txt"))
with open(src("bench_full.log")) as f, open(dst("r1_bench_line.json"), "w") as g:
    g.write(f.read().strip().splitlines()[-1] + "\n")

rows = list(csv.reader(open(dst("r1_bench_launches.csv"))))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi, mi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Name"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    agg.setdefault(r[ki].split("(")[0].replace("void ", "").replace("cmr::", ""), []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(dst("r1_bench_launches_summary.txt"), "w") as f:
    f.write("# r1: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`\n")
    f.write("# (ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 130; per-launch times are cold-cache and\n")
    f.write("#  serialised: compare SHARES, not absolutes).  Full list: r1_bench_launches.csv\n")
    f.write(f"{'kernel':48s} {'launches':>8s} {'avg_us':>8s} {'share':>6s}\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"{k[:48]:48s} {len(v):8d} {sum(v) / len(v):8.1f} {sum(v) / tot:6.3f}\n")

out = subprocess.run(["ncu", "-i", src("env.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
kn = hdr.index("Kernel Name")
mult = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}
traffic = collections.defaultdict(list)
with open(dst("r1_env_kernels_ncu.txt"), "w") as f:
    f.write('# r1: ncu --set full --clock-control none --import-source on -k regex:"k_project|k_tile_gather|k_reward" -s 40 -c 6\n')
    f.write("#     python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline   (B200, 32 KITTI episodes; caches flushed before every launch)\n")
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("cmr::", "")
        f.write(f"\n== {name}\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                f.write(f"   {w:95s} {r[i]:>14s} {units[i]}\n")
        rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        traffic[name.split("<")[0]].append(float(r[rd].replace(",", "")) * mult[units[rd]] +
                                           float(r[wr].replace(",", "")) * mult[units[wr]])
tr = {k: sum(v) / len(v) for k, v in traffic.items()}
tr["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, B=32 KITTI, "
              "round 1 (final build). Below the algorithmic bytes where a kernel's output is still in the 126 MB L2 when "
              "it ends (write-back deferred).")
json.dump(tr, open(dst("dram_traffic.json"), "w"), indent=1)
print(open(dst("r1_bench_launches_summary.txt")).read())
print(tr)
