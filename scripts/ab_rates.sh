#!/bin/bash
# like ab_defs.sh, for the other configs: scripts/ab_rates.sh "<defs A>" "<defs B>" ...
for d in "$@"; do
  TC_NVCC_DEFS="$d" python -m trueconsense_b200.build cuda --force > /dev/null 2>&1 || { echo "build failed: $d"; continue; }
  echo "== $d"
  python scripts/config_rates.py 2>&1 | grep "kernel=0" | sed 's/ scale.*kernel=0://'
done
python -m trueconsense_b200.build cuda --force > /dev/null 2>&1
