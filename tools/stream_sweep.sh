#!/bin/bash
# resident-kernel knob sweep at the bench workload (debug)
grep MHz /proc/cpuinfo | head -2
for m in "0 0" "0 1" "1 1"; do
  set -- $m
  echo "=== COGAPS_RECORD_STORES=$1 COGAPS_HOST_PROFILE=$2"
  COGAPS_RECORD_STORES=$1 COGAPS_HOST_PROFILE=$2 timeout 200 python tools/stream_debug.py 2>&1 | tail -7
done
