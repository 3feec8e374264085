#!/bin/bash
# gpurun_retry.sh OUTFILE [gpurun args...] : retry a gpurun call while the pod answers "busy"
out=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if ! grep -q "status=transient" "$out"; then exit 0; fi
  sleep 90
done
