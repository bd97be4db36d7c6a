#!/bin/bash
set -u
for wide in 0 2 1; do
echo "== ORBX_OCTREE_WIDE=$wide"
for cfg in "640 480 1000 1" "1280 800 1250 1" "1280 800 6250 1" "1920 1080 2000 1" "1280 720 1250 32" "1920 1080 2000 16"; do
  ORBX_OCTREE_WIDE=$wide timeout 200 python tools/stage_times.py $cfg 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['shape'], d['stage_us']['quadtree'], d['result_sha1'])"
done
done
