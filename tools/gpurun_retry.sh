#!/bin/bash
# gpurun with retries while the pod has no free slot (exit 3 / "transient"): nothing is charged for those.
# usage: tools/gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for attempt in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then
    echo "attempt $attempt: no slot, retrying in 90 s" >> "$log.retries"
    sleep 90
    continue
  fi
  break
done
echo "done rc=$rc" >> "$log"
