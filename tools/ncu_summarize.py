#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> a compact per-kernel table (markdown + csv) of the counters the roofline
argument needs: time, DRAM bytes / throughput, L1/L2 sectors per request, achieved occupancy, issue-slot use, the top
stall reasons.  Usage: ncu_summarize.py raw.csv out_prefix [title]"""
import csv
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "rdMB", None),
    ("dram__bytes_write.sum", "wrMB", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", 1),
    ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio", "sec/ld", 1),
    ("l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_st.ratio", "sec/st", 1),
    ("lts__t_sector_hit_rate.pct", "L2hit%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2%", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("launch__block_size", "block", 1),
]
# warps stalled on <reason> per issue slot that issued (a ratio: 3.0 = on average three warps were waiting on it)
STALLS = [(n, "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % n) for n in
          ("long_scoreboard", "barrier", "short_scoreboard", "lg_throttle", "math_pipe_throttle", "wait", "membar",
           "not_selected", "mio_throttle", "sleeping", "branch_resolving", "no_instruction", "dispatch_stall", "tex_throttle")]


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return None


def to_bytes(v, unit):
    f = num(v)
    if f is None:
        return None
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ns(v, unit):
    f = num(v)
    if f is None:
        return None
    return f * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)


def main():
    raw, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else raw
    rows = list(csv.reader(l for l in open(raw) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"^(void )?(usip::)?", "", r[idx["Kernel Name"]])
        name = re.sub(r"\(.*", "", name)[:60]
        rec = {"kernel": name}
        for key, short, _ in COLS:
            if key not in idx:
                rec[short] = None
                continue
            v, u = r[idx[key]], units[idx[key]]
            if short == "us":
                ns = to_ns(v, u); rec[short] = None if ns is None else ns / 1e3
            elif short in ("rdMB", "wrMB"):
                b = to_bytes(v, u); rec[short] = None if b is None else b / 1e6
            else:
                rec[short] = num(v)
        st = [(n, num(r[idx[k]])) for n, k in STALLS if k in idx]
        st = sorted([(n, v) for n, v in st if v is not None], key=lambda t: -t[1])[:3]
        rec["top_stalls"] = " ".join("%s:%.1f" % t for t in st)
        recs.append(rec)
    shorts = [s for _, s, _ in COLS]
    with open(out + ".csv", "w") as f:
        w = csv.writer(f); w.writerow(["kernel"] + shorts + ["top_stalls"])
        for rec in recs:
            w.writerow([rec["kernel"]] + [rec[s] for s in shorts] + [rec["top_stalls"]])
    with open(out + ".md", "w") as f:
        f.write("# %s\n\n`ncu --set full --clock-control none` (one launch per row, in launch order; times are serialised, "
                "cold-ish cache: compare shares and counters, not absolutes).\n\n" % title)
        f.write("| kernel | " + " | ".join(shorts) + " | top stalls |\n|---|" + "---:|" * len(shorts) + "---|\n")
        for rec in recs:
            f.write("| `%s` | " % rec["kernel"] + " | ".join("" if rec[s] is None else ("%.1f" % rec[s] if isinstance(rec[s], float) and rec[s] % 1 else "%d" % rec[s]) for s in shorts)
                    + " | %s |\n" % rec["top_stalls"])
    print("%d kernels -> %s.{md,csv}" % (len(recs), out))


if __name__ == "__main__":
    main()
