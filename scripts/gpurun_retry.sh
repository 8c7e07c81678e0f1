#!/bin/bash
# gpurun with retries while the pod answers "transient" (no box or slot free: nothing is charged)
# usage: scripts/gpurun_retry.sh <timeout_s> <command...>
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  out=$(gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -80
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  echo "[retry] attempt $attempt was transient; sleeping 60 s"
  sleep 60
done
