#!/bin/bash
# Retry a gpurun call while the pod answers "transient" (exit 3 / nothing charged).  usage: tools/gpurun_retry.sh [gpurun args...] -- 'cmd'
# Sleeps 150 s between attempts, gives up after 12.
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient"; then echo "[retry $i] transient, sleeping"; sleep 150; continue; fi
  exit $rc
done
exit 3
