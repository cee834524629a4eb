#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "transient" / busy (exit code 3)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: gave up after 40 attempts"; exit 3
