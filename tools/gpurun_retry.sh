#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'
# Retries while the pod answers "busy" (exit 3 / status=transient); nothing is charged for those.
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up after 30 busy answers"; exit 3
