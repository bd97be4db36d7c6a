#!/bin/bash
set -u
for cfg in "640 480 1000 1" "1280 800 1250 1" "1280 800 6250 1" "640 480 1000 64"; do
  timeout 200 python tools/stage_times.py $cfg 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['shape'], d['stage_us'], d['result_sha1'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
