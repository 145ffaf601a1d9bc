"""Per-CUDA-source-line instruction counts / stall samples from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, hdr = None, None
agg = collections.OrderedDict()
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        cur = r[1].split('/')[-1]; continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r; ki = hdr.index("Instructions Executed"); si = hdr.index("# Samples"); continue
    if hdr and cur and len(r) == len(hdr) and r[0].isdigit():
        key = (cur, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1], 0])
        a[0] += int(r[ki] or 0); a[1] += int(r[si] or 0); a[3] += 1
tot = sum(a[0] for a in agg.values()); stot = sum(a[1] for a in agg.values())
print(path, "warp instructions", tot, "samples", stot)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% smp %4d sass  %s:%d  %s" % (100 * a[0] / max(tot, 1), 100 * a[1] / max(stot, 1), a[3], f[:14], ln, a[2].strip()[:95]))
