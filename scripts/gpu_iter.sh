#!/bin/bash
# One GPU iteration: pileup parity tests, a short bench of the chosen kernel variant, optionally an ncu capture.
#   scripts/gpu_iter.sh <tag> [kernel] [ncu:0|1] [scale]
tag=$1; k=${2:-0}; prof=${3:-0}; scale=${4:-0.25}
python -m pytest tests -m gpu -x -q -k "pileup or config2" 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --kernel $k > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { tail -5 gpurun_out/bench_$tag.err; exit 1; }
python - <<PY
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
print("$tag kernel_ms", round(d["roofline"]["kernel_ms"], 4), "step_ms", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["frac"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 3))
PY
if [ "$prof" = "1" ]; then
  ncu --set full --clock-control none --import-source on -k regex:warp_pileup_kernel -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --scale $scale --kernel $k > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
fi
