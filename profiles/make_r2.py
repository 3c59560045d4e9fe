"""Round-2 summaries from the captures a gpurun call leaves in gpurun_out/ (tracked copies live here).

    python profiles/make_r2.py <env.ncu-rep> <tower.ncu-rep> [cost_volume.ncu-rep] [bench_launches.csv] [sample.ncu-rep]

Outputs: r2_env_kernels_ncu.txt, r2_tower_ncu.txt, r2_cost_volume_ncu.txt, dram_traffic.json, r2_sass_evidence.txt,
r2_bench_launches.csv + r2_bench_launches_summary.txt
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
MULT = {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Gbyte": 1e9}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarise(rep, dst, header):
    hdr, units, rows = raw(rep)
    kn = hdr.index("Kernel Name")
    traffic = collections.defaultdict(list)
    with open(os.path.join(HERE, dst), "w") as f:
        f.write(header)
        for r in rows:
            name = r[kn].split("(")[0].replace("void ", "").replace("cmr::", "")
            f.write(f"\n== {name}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"   {w:95s} {r[i]:>14s} {units[i]}\n")
            rd, wr = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic[name.split("<")[0]].append(float(r[rd].replace(",", "")) * MULT[units[rd]] +
                                               float(r[wr].replace(",", "")) * MULT[units[wr]])
    return {k: sum(v) / len(v) for k, v in traffic.items()}


def sass():
    lib = os.path.join(ROOT, "cmr_agent_b200", "libcmr_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    keys = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "UCGABAR",
            "REDUX", "ATOMG", "REDG", "FMNMX.NAN", "HMMA", "CCTL")
    for ln in out.splitlines():
        if "Function :" in ln:
            cur = ln.split("Function :")[1].strip()
            per[cur] = collections.Counter()
        elif cur:
            for k in keys:
                if re.search(r"(?<![A-Z])" + re.escape(k), ln):     # "HMMA" must not count "UTCHMMA"
                    per[cur][k] += 1
    with open(os.path.join(HERE, "r2_sass_evidence.txt"), "w") as f:
        f.write("# r2: cuobjdump -sass cmr_agent_b200/libcmr_b200.so - counts of the Blackwell-specific mnemonics per kernel\n")
        f.write("#   UTCHMMA = tcgen05.mma (kind::f16), LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM alloc,\n")
        f.write("#   UTMALDG/UTMASTG = tiled TMA load/store, UBLKCP = cp.async.bulk, SYNCS = mbarrier, LDGSTS = cp.async,\n")
        f.write("#   UCGABAR = cluster barrier, FMNMX.NAN = max.NaN.  No HMMA (legacy mma.sync) anywhere.\n")
        for fn, c in per.items():
            if c:
                short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
                f.write(f"{short[:70]:70s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())) + "\n")


def launches(path):
    """ncu launch list of `bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline`: per-kernel launches, time and share."""
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        name = r[kn].split("(")[0].replace("void ", "").replace("cmr::", "")
        c = per.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += t
    total = sum(v[1] for v in per.values())
    with open(path) as src, open(os.path.join(HERE, "r2_bench_launches.csv"), "w") as dst:
        dst.write(src.read())
    with open(os.path.join(HERE, "r2_bench_launches_summary.txt"), "w") as f:
        f.write("# r2: ncu --metrics gpu__time_duration.sum --clock-control none -s 180 -c 200 --csv\n"
                "#     python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline   (cold caches, serialised launches:\n"
                "#     the SHARES are what compares with bench.py's live events, not the absolute times)\n")
        f.write(f"{'kernel':44s} {'launches':>9s} {'total ns':>12s} {'avg ns':>10s} {'share':>7s}\n")
        for k, (n, t) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:44]:44s} {n:9d} {t:12.0f} {t / n:10.0f} {t / total:7.1%}\n")
    print(open(os.path.join(HERE, "r2_bench_launches_summary.txt")).read())


if __name__ == "__main__":
    tr = {}
    if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
        tr.update(summarise(sys.argv[1], "r2_env_kernels_ncu.txt",
                            '# r2: ncu --set full --clock-control none --import-source on -k regex:"k_project|k_tile_gather" -s 20 -c 4\n'
                            "#     python benchmarks/microbench.py env   (B200, 32 KITTI episodes; caches flushed before every launch)\n"))
    if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
        tr.update(summarise(sys.argv[2], "r2_tower_ncu.txt",
                            '# r2: ncu --set full --clock-control none --import-source on -k regex:"k_tower" -s 20 -c 5\n'
                            "#     python benchmarks/microbench.py tower --batch 32   (B200, 32 x 40960 points)\n"))
    if len(sys.argv) > 3 and os.path.exists(sys.argv[3]):
        cv = summarise(sys.argv[3], "r2_cost_volume_ncu.txt",
                       '# r2: ncu --set full --clock-control none --import-source on -k regex:"k_cost_volume" -s 4 -c 2\n'
                       "#     python benchmarks/microbench.py cost_volume   (B200, 729 poses of one KITTI cloud, 1.015 GB out)\n")
        tr.update({k + " (cost volume)": v for k, v in cv.items()})
    if len(sys.argv) > 4 and os.path.exists(sys.argv[4]):
        launches(sys.argv[4])
    if len(sys.argv) > 5 and os.path.exists(sys.argv[5]):
        sm = summarise(sys.argv[5], "r2_sample_ncu.txt",
                       '# r2: ncu --set full --clock-control none --import-source on -k regex:"k_bilinear_sample" -s 3 -c 1\n'
                       "#     python benchmarks/microbench.py sample   (B200, 32 KITTI episodes at the ground-truth pose)\n")
        tr.update(sm)
    if tr:
        tr["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full --clock-control none, round 2. "
                      "Below the algorithmic bytes where a kernel's output is still in the 126 MB L2 when it ends.")
        json.dump(tr, open(os.path.join(HERE, "dram_traffic.json"), "w"), indent=1)
    sass()
    print(open(os.path.join(HERE, "r2_sass_evidence.txt")).read()[:3000])
