"""SASS opcode histogram per kernel of libdrosfm_b200.so -> profiles/<tag>_sass_histogram.txt (cuobjdump -sass)."""
import collections, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
so = "dro_sfm_b200/libdrosfm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
kern, hist, order = None, {}, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); order.append(kern); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        hist[kern][m.group(1).split(".")[0] + ("." + ".".join(m.group(1).split(".")[1:3]) if m.group(1).startswith(("LDG", "STG", "RED", "ATOM", "MUFU", "LDS", "STS", "BAR", "SHFL")) else "")] += 1
names = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
with open("profiles/%s_sass_histogram.txt" % tag, "w") as f:
    f.write("# static SASS opcode counts per kernel (cuobjdump -sass %s); memory / MUFU / barrier opcodes keep their first two modifiers\n" % so)
    for k, n in sorted(zip(order, names), key=lambda t: t[1]):
        h = hist[k]; tot = sum(h.values())
        n = re.sub(r"\(.*", "", n)
        f.write("\n## %s  (%d instructions)\n" % (n, tot))
        f.write("  " + "  ".join("%s %d" % kv for kv in h.most_common(28)) + "\n")
print("ok", len(order))
