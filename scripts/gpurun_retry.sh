#!/bin/bash
# scripts/gpurun_retry.sh <timeout> <command...> — keep asking for a GPU box until the pod has a free slot (exit 3 = busy, nothing charged)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
