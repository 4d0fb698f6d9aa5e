#!/bin/bash
# usage: gpurun_retry.sh <timeout> <cmd...>  -- retries while the pod answers "busy" (exit 3 / transient)
T=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); rc=$?
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then echo "[retry $i] busy, sleeping 120 s"; sleep 120; continue; fi
  exit $rc
done
exit 3
