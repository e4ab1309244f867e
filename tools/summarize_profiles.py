#!/usr/bin/env python
"""Turn gpurun_out/{launches_TAG.csv, prof_tc_TAG.ncu-rep, bench_TAG.json} into tracked summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
G = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches():
    f = os.path.join(G, "launches_%s.csv" % TAG)
    if not os.path.isfile(f):
        return
    lines = [l for l in open(f) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict(); tot = 0.0
    for x in rows:
        v = float(x["Metric Value"].replace(",", ""))
        v = v / 1e3 if x["Metric Unit"] == "ns" else (v * 1e3 if x["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", x["Kernel Name"])[:70]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(os.path.join(OUT, "%s_launches_summary.md" % TAG), "w") as o:
        o.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3`\n\n" % TAG)
        o.write("Cold-cache, serialised per-launch times: compare SHARES, not absolutes.  %d launches, %.1f us total.\n\n" % (len(rows), tot))
        o.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            o.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, n, v, v / n, 100 * v / tot))
    with open(os.path.join(OUT, "%s_launches.csv" % TAG), "w") as o:
        o.writelines(lines)


def full():
    rep = os.path.join(G, "prof_tc_%s.ncu-rep" % TAG)
    if not os.path.isfile(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(os.path.join(OUT, "%s_ncu_full_layer_fwd_tc.md" % TAG), "w") as o:
        o.write("# ncu --set full --clock-control none, kernel regex layer_fwd_tc_kernel (%s)\n\n" % TAG)
        for r in rows[2:]:
            o.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" %
                    (r[idx["Kernel Name"]][:90], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
            for k in KEYS:
                if k in idx:
                    o.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            o.write("\n")


def bench():
    for name in os.listdir(G):
        if name.startswith("bench") and name.endswith(".json") and TAG in name:
            txt = open(os.path.join(G, name)).read().strip().splitlines()
            if txt:
                with open(os.path.join(OUT, name), "w") as o:
                    o.write(txt[-1] + "\n")


def traffic():
    """dram bytes of the dominant kernel (first launch in the full capture) -> profiles/TAG_top_kernel_traffic.json,
    which bench.py reports as roofline.traffic"""
    rep = os.path.join(G, "prof_tc_%s.ncu-rep" % TAG)
    if not os.path.isfile(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, r = rows[0], rows[1], rows[2]
    idx = {h: i for i, h in enumerate(hdr)}
    def b(k):
        v = float(r[idx[k]].replace(",", "")); u = units[idx[k]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    out = {"kernel": "knn_a1[131072x512->512] (tools/tc_one.py 131072 512 512 16)", "source": "prof_tc_%s.ncu-rep, ncu --set full" % TAG,
           "dram_bytes_read": b("dram__bytes_read.sum"), "dram_bytes_write": b("dram__bytes_write.sum"),
           "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]].replace(",", "")) * (1e-3 if units[idx["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0)}
    out["traffic"] = out["dram_bytes_read"] + out["dram_bytes_write"]
    json.dump(out, open(os.path.join(OUT, "%s_top_kernel_traffic.json" % TAG), "w"), indent=1)


launches(); full(); bench(); traffic()
sass = subprocess.run("cuobjdump -sass %s | grep -oE 'UTC[A-Z0-9.]*|LDTM[A-Z0-9.]*|UBLKCP[A-Z0-9.]*|UTCBAR[A-Z0-9.]*|SYNCS[A-Z0-9.]*' | sort | uniq -c | sort -rn"
                      % os.path.join(ROOT, "usip_b200", "lib", "libusip_b200.so"), shell=True, capture_output=True, text=True).stdout
open(os.path.join(OUT, "sass_tcgen05_evidence.txt"), "w").write(
    "cuobjdump -sass usip_b200/lib/libusip_b200.so | grep -oE 'UTC*|LDTM*|UBLKCP*|UTCBAR*|SYNCS*' | sort | uniq -c\n" + sass)
print(os.listdir(OUT))
