#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...> : retries while the pod answers "busy" (exit code 3)
log=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
	/usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
	rc=$?
	if [ $rc -ne 3 ]; then exit $rc; fi
	sleep 120
done
exit 3
