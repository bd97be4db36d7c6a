#!/bin/bash
set -u
for cfg in "640 480 1000 1" "1280 800 1250 1" "1280 800 6250 1" "1920 1080 2000 1"; do
  timeout 200 python tools/stage_times.py $cfg 20 2>&1 | tail -1 | cut -c1-330
done
