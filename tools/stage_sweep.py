#!/usr/bin/env python
"""Runs tools/stage_times.py once per environment setting (the library reads its knobs once per process) and prints one line each.
Usage: python tools/stage_sweep.py "ORBX_FAST_V=1" "ORBX_FAST_V=2 ORBX_FAST_MIX=0" ...   [-- W H NFEAT BATCH REPS]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
tail = []
if "--" in args:
    k = args.index("--"); tail = args[k + 1:]; args = args[:k]
for setting in args or [""]:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1); env[k] = v
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stage_times.py")] + tail, env=env, capture_output=True, text=True, timeout=600)
    out = p.stdout.strip().splitlines()
    print(out[-1] if out and p.returncode == 0 else f'{{"env": "{setting}", "error": {p.returncode}, "stderr": {p.stderr[-400:]!r}}}', flush=True)
