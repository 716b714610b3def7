"""Summaries of ncu outputs brought back from the GPU box (run here, no GPU needed).

    python profiles/summarize.py launches gpurun_out/launches_X.csv
    python profiles/summarize.py raw gpurun_out/prof_X.ncu-rep [metric-regex]
    python profiles/summarize.py source gpurun_out/prof_X.ncu-rep [top-n]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])[:70]
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else v * 1e3 if r[mu] == "ms" else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'total us':>12} {'share':>6} {'n':>4}  kernel   (cold-cache, serialised: compare shares, not absolutes)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1]:12.1f} {100 * v[1] / tot:5.1f}% {v[0]:4d}  {k}")


def raw(path, pat=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    default = (r"gpu__time_duration.sum$|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|"
               r"launch__registers_per_thread$|launch__grid_size|launch__occupancy_limit|smsp__issue_active.avg.pct|"
               r"sm__inst_executed.sum$|smsp__inst_executed_op_shared|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|"
               r"smsp__average_warp.*issue_stalled.*_ratio$|smsp__warp_issue_stalled.*per_warp_active.pct$|"
               r"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|sm__throughput.avg.pct|lts__t_sector_hit_rate.pct|"
               r"smsp__thread_inst_executed_per_inst_executed.ratio|smsp__inst_executed.sum$|launch__waves_per_multiprocessor|"
               r"sm__cycles_active.avg$|launch__shared_mem_per_block")
    rx = re.compile(pat or default)
    for vals in rows[2:]:
        print("==", vals[hdr.index("Kernel Name")][:80])
        for i, h in enumerate(hdr):
            if rx.search(h):
                print(f"  {h:85s} {vals[i]:>16s} {units[i]}")


def source(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] in ("#", "Address", "Source")]
    hdr = rows[hi[0]] if hi else rows[0]
    print(hdr)
    data = rows[(hi[0] if hi else 0) + 1:]
    # pick the sampling column
    cand = [i for i, h in enumerate(hdr) if "Samples" in h or "Sampling" in h]
    if not cand:
        for r in data[:top]:
            print(r)
        return
    si = cand[0]
    def val(r):
        try:
            return float(r[si].replace(",", ""))
        except Exception:
            return 0.0
    tot = sum(val(r) for r in data) or 1.0
    for r in sorted(data, key=val, reverse=True)[:top]:
        print(f"{100 * val(r) / tot:5.1f}%  " + " | ".join(r[:3])[:160])


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "launches":
        launches(sys.argv[2])
    elif cmd == "raw":
        raw(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
