#!/bin/bash
# like tools/gpu.sh, but keeps retrying while the pod answers "busy" (exit code 3, nothing charged)
# usage: tools/gpu_retry.sh [--gpus N] <timeout seconds> '<command>'
cd "$(dirname "$0")/.."
for i in $(seq 1 40); do
  tools/gpu.sh "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
