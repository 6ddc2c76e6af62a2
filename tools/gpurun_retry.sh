#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy); usage: gpurun_retry.sh [gpurun args] -- 'cmd'
for attempt in $(seq 1 80); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then
    wait=$(echo "$out" | grep -o "retry in [0-9]*s" | grep -o "[0-9]*")
    sleep ${wait:-75}
    continue
  fi
  echo "$out"
  exit 0
done
echo "gpurun_retry: gave up after 30 attempts"
echo "$out" | tail -3
