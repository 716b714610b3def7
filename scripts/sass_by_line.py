"""Static SASS instruction counts per source line of one kernel (no GPU): nvdisasm -g prints the line table.
    python scripts/sass_by_line.py build/cuda/pileup_flat.o flat_pileup_kernelILi64ELb0 [first last]"""
import re, subprocess, sys, collections

obj, want = sys.argv[1], sys.argv[2]
lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (0, 10**9)
import os, tempfile
tmp = tempfile.mkdtemp()
if not obj.endswith(".cubin"):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    obj = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
out = subprocess.run(["nvdisasm", "-g", "-c", obj], capture_output=True, text=True).stdout
cur_fn, line, counts, listing = None, None, collections.Counter(), []
for l in out.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
    if m:
        cur_fn = m.group(1); continue
    if cur_fn is None or want not in cur_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and line:
        counts[line] += 1
        listing.append((line, m.group(2).strip()))
    elif re.match(r"\s*\.L_x_\d+:", l):
        listing.append((None, l.strip()))
tot = sum(counts.values())
print("total SASS instructions", tot)
if len(sys.argv) > 4:
    for ln, ins in listing:
        if ln is None: print("      ", ins)
        elif ln[0].startswith("pileup_flat") and lo <= ln[1] <= hi: print(f"{ln[1]:5d}  {ins}")
else:
    for (f, n), c in sorted(counts.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        print(f"{f}:{n:<5d} {c}")
