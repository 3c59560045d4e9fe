"""Summaries of the captured Test_Agent loop's kernel list (benchmarks/agent_loop_profile.sh, one gpurun call).

    python profiles/make_agent_loop.py

Reads gpurun_out/r2h_loop_b{1,32}.csv (ncu --metrics gpu__time_duration.sum of ONE replay = 10 iterations) and
gpurun_out/r2h_agent_kernels.ncu-rep (ncu --set full of k_grouped_linear / k_conv_epilogue* inside the B = 32 replay);
writes profiles/r2_agent_loop_launches.txt and profiles/r2_agent_kernels_ncu.txt.
"""
import collections
import csv
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")

# which part of an iteration a kernel belongs to (first match wins)
GROUPS = [
    ("environment (this library)", ("k_project", "k_tile_gather", "k_tile_scatter", "k_step", "k_reward", "k_feat_compact",
                                    "k_cloud_mean", "k_overlap_scan", "k_to_disentangled", "k_mean", "k_scan")),
    ("3-D tower (this library, tcgen05)", ("k_tower",)),
    ("heads + 1x1 tail + action (this library)", ("k_grouped_linear", "k_deterministic_action")),
    ("2-D head epilogues + layout (this library)", ("k_conv_epilogue", "k_to_channels_last")),
    ("2-D head convolutions (cuDNN / CUTLASS)", ("cudnn", "cutlass", "sm100_", "sm90_", "sm80_", "xmma", "implicit_convolve",
                                                 "conv", "nchwToNhwc", "nhwcToNchw", "gemm", "wgrad", "fprop")),
]


def group_of(name):
    for g, keys in GROUPS:
        if any(k in name for k in keys):
            return g
    return "torch elementwise / reductions / copies"


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        try:
            t = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        t *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[mu], 1.0)
        name = r[kn].split("(")[0].replace("void ", "").replace("cmr::", "")
        c = per.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += t
    return per


def write_launches():
    dst = os.path.join(HERE, "r2_agent_loop_launches.txt")
    with open(dst, "w") as f:
        f.write("# r2: ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv\n"
                "#     python benchmarks/debug/agent_loop_launches.py B     (ONE replay of the captured Test_Agent loop =\n"
                "#     10 iterations of observation + unchanged CMRAgent through accelerate_agent + step; cold caches,\n"
                "#     serialised launches: shares, not absolute times, compare with bench.py's secondary.test_agent_loop)\n")
        for B in (1, 32):
            path = os.path.join(OUT, f"r2h_loop_b{B}.csv")
            if not os.path.exists(path):
                continue
            per = launches(path)
            total = sum(v[1] for v in per.values())
            n = sum(v[0] for v in per.values())
            f.write(f"\n== batch {B}: {n} launches, {total / 1e3:.1f} us of kernel time per replay "
                    f"({total / 1e4:.1f} us per iteration)\n")
            grp = collections.OrderedDict()
            for k, (c, t) in per.items():
                g = grp.setdefault(group_of(k), [0, 0.0])
                g[0] += c
                g[1] += t
            for g, (c, t) in sorted(grp.items(), key=lambda kv: -kv[1][1]):
                f.write(f"   {g:48s} {c:5d} launches {t / 1e3:9.1f} us {t / total:6.1%}\n")
            f.write(f"   {'kernel':60s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}\n")
            for k, (c, t) in sorted(per.items(), key=lambda kv: -kv[1][1])[:28]:
                f.write(f"   {k[:60]:60s} {c:8d} {t / 1e3:10.1f} {t / c / 1e3:8.2f} {t / total:6.1%}\n")
    print(open(dst).read())


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def write_full():
    rep = os.path.join(OUT, "r2h_agent_kernels.ncu-rep")
    if not os.path.exists(rep):
        return
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    kn = hdr.index("Kernel Name")
    gs = hdr.index("launch__grid_size")
    with open(os.path.join(HERE, "r2_agent_kernels_ncu.txt"), "w") as f:
        f.write('# r2: ncu --set full --clock-control none --import-source on --profile-from-start off\n'
                '#     -k regex:"k_grouped_linear|k_conv_epilogue" -c 14 python benchmarks/debug/agent_loop_launches.py 32\n'
                "#     (the first iteration's agent-side kernels of this library inside the captured loop, B = 32)\n")
        for r in rows:
            name = r[kn].split("(")[0].replace("void ", "").replace("cmr::", "")
            f.write(f"\n== {name}   grid {r[gs]}\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"   {w:85s} {r[i]:>14s} {units[i]}\n")
    print(open(os.path.join(HERE, "r2_agent_kernels_ncu.txt")).read()[:3000])


if __name__ == "__main__":
    write_launches()
    write_full()
