#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>   (retries while the pod answers busy: exit code 3)
log="$1"; to="$2"; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
