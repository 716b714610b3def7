"""SASS of a source-line range with execution counts (run here, no GPU).
    python profiles/sass_lines.py rep kernel-substring file first last [min-count]"""
import csv, io, subprocess, sys

def main():
    rep, want, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    minc = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fn = f = cur = None
    sass = {}
    for r in rows:
        if len(r) == 2 and r[0] == "Function Name": fn = r[1]; continue
        if len(r) == 2 and r[0] == "File Path": f = r[1].split("/")[-1]; continue
        if len(r) < 10 or want not in (fn or ""): continue
        if r[0] not in ("", "Line No"): cur = (f, int(r[0])); continue
        if r[0] == "" and r[2].startswith("0x"):
            sass[int(r[2], 16)] = (r[3].strip(), cur, int(r[7]) if r[7].isdigit() else 0, int(r[8]) if r[8].isdigit() else 0)
    keys = sorted(sass)
    idx = [i for i, k in enumerate(keys) if sass[k][1] and sass[k][1][0] == fname and lo <= sass[k][1][1] <= hi and sass[k][2] >= minc]
    if not idx: return
    for k in keys[min(idx):max(idx) + 1]:
        s = sass[k]
        if s[2] >= minc:
            print(f"{s[1][1] if s[1] else 0:>4d} {s[2]:9d} {s[3] / max(s[2], 1):5.1f}  {s[0]}")

main()
