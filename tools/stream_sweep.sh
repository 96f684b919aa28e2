#!/bin/bash
# resident-kernel statistics at the bench workload (debug)
COGAPS_PERSISTENT_DEBUG=1 timeout 200 python tools/stream_debug.py 2>&1 | tail -5
