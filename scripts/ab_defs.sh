#!/bin/bash
# A/B of compile-time tuning knobs on the GPU box: rebuilds the CUDA library with each set of -D flags and
# prints the pileup kernel time of a short bench.   scripts/ab_defs.sh "<defs A>" "<defs B>" ...
for d in "$@"; do
  TC_NVCC_DEFS="$d" python -m trueconsense_b200.build cuda --force > /dev/null 2>&1 || { echo "build failed: $d"; continue; }
  python bench.py --steps 5 --warmup 3 --cpu-reads 2000 > /tmp/ab.json 2> /tmp/ab.err || { echo "bench failed: $d"; tail -3 /tmp/ab.err; continue; }
  python - "$d" <<PY
import json, sys
d = json.load(open("/tmp/ab.json"))
print(f"{sys.argv[1]:50s} kernel_ms {d['roofline']['kernel_ms']:.4f} step_ms {d['ms_per_step']:.4f}")
PY
done
python -m trueconsense_b200.build cuda --force > /dev/null 2>&1
