#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>  — retries while the pod answers busy/transient
log=$1; shift
for i in $(seq 1 ${GPURUN_TRIES:-12}); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|exit code 3\|no box\|busy" "$log" && ! grep -q "status=ok" "$log"; then sleep 120; continue; fi
  break
done
tail -3 "$log"
