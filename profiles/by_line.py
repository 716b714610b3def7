"""Instruction / stall-sample attribution per CUDA source line from an ncu report (run here, no GPU).

    python profiles/by_line.py gpurun_out/prof_X.ncu-rep [kernel-substring] [top-n]
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur_file, cur_fn, hdr = None, None, None
    agg = {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) == 2 and r[0] == "Function Name":
            cur_fn = r[1]
            continue
        if len(r) < 10:
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "" or want not in (cur_fn or ""):
            continue
        try:
            key = (cur_fn[:40], cur_file, int(r[0]))
            inst, thr, smp = int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("Thread Instructions Executed")]), int(r[hdr.index("# Samples")])
        except (ValueError, IndexError):
            continue
        a = agg.setdefault(key, [r[1], 0, 0, 0])
        a[1] += inst; a[2] += thr; a[3] += smp
    tot = sum(a[1] for a in agg.values()) or 1
    tots = sum(a[3] for a in agg.values()) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{key[1]}:{key[2]:<4d} {a[1] / 1e6:8.2f}M {100 * a[1] / tot:5.1f}%  thr/inst {a[2] / max(a[1], 1):5.1f}  samples {100 * a[3] / tots:5.1f}%  {a[0].strip()[:100]}")
    return agg


if __name__ == "__main__":
    main()


def phases(path, want, ranges):
    """ranges: list of (name, file, first_line, last_line)"""
    sys.argv = [sys.argv[0], path, want, "0"]
    agg = main()
    tot = sum(a[1] for a in agg.values()) or 1
    tots = sum(a[3] for a in agg.values()) or 1
    for name, f, lo, hi in ranges:
        sel = [a for k, a in agg.items() if k[1] == f and lo <= k[2] <= hi]
        i, t, s = sum(a[1] for a in sel), sum(a[2] for a in sel), sum(a[3] for a in sel)
        print(f"{name:12s} {i / 1e6:8.2f}M {100 * i / tot:5.1f}%  thr/inst {t / max(i, 1):5.1f}  samples {100 * s / tots:5.1f}%")
