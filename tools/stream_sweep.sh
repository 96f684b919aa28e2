#!/bin/bash
# resident-kernel knob sweep at the bench workload (debug)
for m in 0 1 0 1; do
  echo "=== COGAPS_GEN_PREFETCH=$m"
  COGAPS_GEN_PREFETCH=$m timeout 200 python tools/stream_debug.py 2>&1 | tail -3
done
