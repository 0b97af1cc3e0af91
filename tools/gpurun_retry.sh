#!/bin/bash
# usage: tools/gpurun_retry.sh <name> <timeout-seconds> [--gpus N] -- <command...>
# Runs `gpurun` and retries (up to 40 times, 45 s apart) while the pod answers "busy / draining" (nothing is charged then).
name=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$@" > "gpurun_out/$name.gpurun" 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" "gpurun_out/$name.gpurun" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  break
done
echo "rc=$rc" >> "gpurun_out/$name.gpurun"
