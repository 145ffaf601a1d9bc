"""Turns gpurun_out/prof_* into the committed summaries under profiles/ (tag = round label)."""
import collections, csv, json, re, subprocess, sys, shutil
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = []
# ---- launch list --------------------------------------------------------------------------------
lines = [l for l in open("gpurun_out/prof_launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(list)
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])
    try: v = float(row["Metric Value"].replace(",", ""))
    except Exception: continue
    unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    agg[name].append(v)
# not part of a step: the 256 MiB L2 flush between steps and the delay kernel of bench.py's per-kernel timing pass
flush = [k for k in agg if "FillFunctor<unsigned char>" in k or "spin_kernel" in k]
for k in flush: agg.pop(k)
tot = sum(sum(v) for v in agg.values())
out.append(f"# ncu launch list ({tag}): python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline")
out.append("# ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 500 (cold-cache, serialised: compare SHARES)\n")
out.append(f"{'kernel':74s} {'launches':>8s} {'avg us':>9s} {'total ms':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"{k[:74]:74s} {len(v):8d} {sum(v)/len(v):9.2f} {sum(v)/1e3:9.3f} {sum(v)/tot:7.1%}")
shutil.copy("gpurun_out/prof_launches.csv", f"profiles/{tag}_launches.csv")
# ---- full set ------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_full.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"), ("launch__registers_per_thread", "registers"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block")]
out.append(f"\n\n# ncu --set full --clock-control none, one eager hot-path step ({tag}); first two launches of each kernel\n")
traffic = {}
seen = collections.Counter()
for r in rows[2:]:
    d = dict(zip(h, r)); u = dict(zip(h, units))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void drosfm::", "").replace("drosfm::", "")
    seen[name] += 1
    if seen[name] > 1: continue
    out.append(f"## {name}")
    for key, label in want:
        if key in d: out.append(f"  {label:24s} {d[key]} {u[key]}")
    def mb(x, unit):
        x = float(x.replace(",", "")); return x * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}[unit]
    traffic[name] = {"dram_read_MB": mb(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]),
                     "dram_write_MB": mb(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])}
open(f"profiles/{tag}_summary.txt", "w").write("\n".join(out) + "\n")
json.dump(traffic, open(f"profiles/{tag}_traffic.json", "w"), indent=1)
print("\n".join(out[:40]))
