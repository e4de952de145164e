"""Attribute ncu SASS-level samples to CUDA source lines (headers included) using nvdisasm line info.
Usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <kernel mangled substring> [top N]"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qm_door_b200", "libqmb200.so")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(d) if f.startswith("qmb200.") and f.endswith(".cubin")][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], text=True, stderr=subprocess.DEVNULL).splitlines()
# instruction offset -> (file, line) for the wanted function
cur, infn, off2line = None, False, {}
for ln in dis:
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(raw.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
sel = int(os.environ.get("NCU_SECTION", "0"))          # which matching launch (several kernels may match the regex)
hi = his[sel]
rows = rows[:his[sel + 1] - 1] if sel + 1 < len(his) else rows
H = rows[hi]
ia, isamp, iinst = H.index("Address"), H.index("# Samples"), H.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
base = None
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    key = off2line.get(a - base, ("?", 0))
    s = int(r[isamp] or 0)
    agg[key][0] += s
    agg[key][1] += int(r[iinst] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            agg[key][2][H[c]] += v
    tot += s
srcs = {}
print("total samples", tot)
for key, (s, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = key
    path = os.path.join(os.path.dirname(so), "csrc", f)
    if f not in srcs and os.path.exists(path):
        srcs[f] = open(path).read().splitlines()
    text = srcs.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in srcs and l > 0 else ""
    print("%5.1f%% %9d inst  %-12s:%4d  %-28s | %s" % (100.0 * s / max(tot, 1), inst, f, l, ",".join("%s=%d" % (k[6:], v) for k, v in st.most_common(3)), text))
