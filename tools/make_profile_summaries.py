#!/usr/bin/env python
"""Turn the captures of tools/capture_profiles.sh (gpurun_out/) into the tracked summaries under profiles/:
  profiles/launches_<tag>_summary.md   per-kernel launch counts, total/avg duration and share of the step (ncu launch list)
  profiles/<kernel>_<tag>_summary.txt  headline metrics + SASS hot segments of the `ncu --set full` capture
  profiles/traffic_<tag>.json          dram bytes per launch of each captured kernel
  profiles/ncu_facts_<tag>.json        per kernel: dram bytes per launch, duration, pipe / issue busy fractions, instructions (bench.py reads
                                       roofline.traffic and roofline.ncu from it; nothing of this is hard-coded in bench.py)
Usage: python tools/make_profile_summaries.py r01"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# ---- launch list
path = os.path.join(G, f"launches_{tag}.csv")
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10 and r[0].strip('"').isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")).replace("orbx::", "")
    name = re.sub(r"<.*", "", name.replace("void ", "")).strip()
    val = float(r[-1].replace(",", ""))
    unit = r[-2]
    us = val / 1000.0 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += us
ours = {k: v for k, v in agg.items() if k.startswith("k_")}
tot = sum(v[1] for k, v in ours.items() if not k.startswith("k_knn") and not k.startswith("k_expand"))
with open(os.path.join(P, f"launches_{tag}_summary.md"), "w") as f:
    f.write(f"# ncu launch list ({tag})\n\nCommand: `ORBX_DEV_SPLIT=1 ORBX_GRAPHS=0 ORBX_BENCH_LANES=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 3 --warmup 3 --no-cpu`\n"
            "(single-range launches without graph replay, so that every launch in the list is one whole-batch kernel; tools/capture_profiles.sh)\n"
            "(cold-cache, serialised: compare shares, not absolutes).  Share = fraction of the extraction kernels' time.\n\n"
            "| kernel | launches | total us | avg us | share of extraction time |\n|---|---|---|---|---|\n")
    for k, (n, us) in ours.items():
        share = f"{us / tot:.3f}" if not k.startswith("k_knn") and not k.startswith("k_expand") else "(kNN leg)"
        f.write(f"| {k} | {n} | {us:.1f} | {us / n:.1f} | {share} |\n")
    others = {k: v for k, v in agg.items() if not k.startswith("k_")}
    f.write("\nNon-orbx kernels in the capture (torch fills / copies of the harness): " + ", ".join(f"{k} x{v[0]}" for k, v in list(others.items())[:8]) + "\n")

# ---- per-kernel captures
traffic, facts = {}, {}
FACT_KEYS = {"sm__issue_active.avg.pct_of_peak_sustained_elapsed": "issue_active_pct", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_alu_pct",
             "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_fma_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "pipe_tensor_pct",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct", "smsp__inst_executed.sum": "warp_instructions",
             "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers", "launch__grid_size": "grid",
             "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct"}
for fn in sorted(os.listdir(G)):
    m = re.match(rf"prof_(k_\w+)_{tag}\.ncu-rep$", fn)
    if not m:
        continue
    k = m.group(1)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(G, fn), "--src"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{k}_{tag}_summary.txt"), "w").write(out)
    rd = re.search(r"dram__bytes_read.sum: ([\d.]+) (\w+)", out)
    wr = re.search(r"dram__bytes_write.sum: ([\d.]+) (\w+)", out)
    dur = re.search(r"gpu__time_duration.sum: ([\d.]+) (\w+)", out)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if rd and wr:
        traffic[k] = {"dram_bytes_per_launch": float(rd.group(1)) * mult[rd.group(2)] + float(wr.group(1)) * mult[wr.group(2)],
                      "duration_under_ncu": f"{dur.group(1)} {dur.group(2)}" if dur else None}
        facts[k] = dict(traffic[k], capture=f"profiles/{k}_{tag}_summary.txt")
        for mk, short in FACT_KEYS.items():
            mm = re.search(re.escape(mk) + r": ([\d.]+)", out)
            if mm:
                facts[k][short] = float(mm.group(1))
if traffic:      # launch-list-only refreshes (no .ncu-rep in gpurun_out/) keep the committed traffic file
    json.dump(traffic, open(os.path.join(P, f"traffic_{tag}.json"), "w"), indent=1)
    json.dump(facts, open(os.path.join(P, f"ncu_facts_{tag}.json"), "w"), indent=1)
print(open(os.path.join(P, f"launches_{tag}_summary.md")).read())
print(json.dumps(traffic, indent=1))
