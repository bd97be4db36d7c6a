#!/bin/bash
set -u
P=tools/_build/blur_tc_probe
timeout 60 $P 640 480 4 2 1 2>&1 | tail -16
timeout 60 $P 640 480 4 2 2 2>&1 | tail -16
timeout 60 $P 640 480 64 2>&1 | tail -16
timeout 60 $P 533 400 64 2>&1 | tail -3
timeout 60 $P 179 134 64 2>&1 | tail -3
timeout 60 $P 1920 1080 16 2>&1 | tail -3
