#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy" (exit 3: nothing charged).  Usage: tools/gpurun_retry.sh <timeout_s> [--gpus N] '<command>'
t=$1; shift
extra=()
if [ "$1" = "--gpus" ]; then extra=(--gpus "$2"); shift 2; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" "${extra[@]}" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
