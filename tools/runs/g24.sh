#!/bin/bash
set -u
for i in 1 2; do timeout 200 python tools/e2e_stream.py 480 2>&1 | tail -1; done
ORBX_E2E_LANES=6 timeout 200 python tools/e2e_stream.py 480 2>&1 | tail -1
nproc; cat /proc/loadavg
