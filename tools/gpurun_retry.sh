#!/usr/bin/env bash
# usage: tools/gpurun_retry.sh <timeout_s> <logfile> '<command>'  -- retries while gpurun answers "busy" (exit 3) or
# "another call running" (exit 2, e.g. a call whose client died)
to="$1"; log="$2"; shift 2
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  sleep 60
done
exit 3
