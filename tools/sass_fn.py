#!/usr/bin/env python
"""Print the SASS of one function of a cubin / .so (substring match on the mangled name) and an opcode histogram.
Usage: sass_fn.py lib.so name_substring [--hist] [--range a b]"""
import collections, re, subprocess, sys

def main():
    path, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, l in enumerate(out):
        if "Function :" in l:
            if start is not None:
                end = i
                break
            if pat in l:
                start = i
    else:
        end = len(out)
    if start is None:
        sys.exit("not found")
    ins = []
    for l in out[start:end]:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    if "--range" in sys.argv:
        k = sys.argv.index("--range"); a, b = int(sys.argv[k + 1], 16), int(sys.argv[k + 2], 16)
        ins = [x for x in ins if a <= x[0] <= b]
    if "--hist" in sys.argv:
        c = collections.Counter()
        for _, t in ins:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0]] += 1
        print(len(ins), "instructions")
        for k, v in c.most_common(60):
            print(f"{v:5d} {k}")
    else:
        for a, t in ins:
            print(f"{a:05x} {t}")

main()
