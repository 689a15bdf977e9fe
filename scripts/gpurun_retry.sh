#!/bin/bash
# developer helper: gpurun with retries while the pod answers "busy" (exit code 3: nothing charged)
# usage: [GPUS=N] scripts/gpurun_retry.sh <timeout seconds> '<command>'
for attempt in $(seq 1 30); do
	/usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2"
	rc=$?
	if [ $rc -ne 3 ]; then
		exit $rc
	fi
	sleep 90
done
exit 3
