"""Development aid: static SASS instruction counts per source line of one kernel (library built with -lineinfo).
usage: python tools/sass_lines.py <kernel-substring> [top]   e.g.  python tools/sass_lines.py k_kinILi2 40"""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.environ.get("QMB200_LIB_PATH", os.path.join(root, "qm_door_b200", "libqmb200.so"))
key, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cnt, ops, on, cur = collections.Counter(), collections.defaultdict(collections.Counter), False, None
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    for ln in txt.splitlines():
        if ln.startswith(".text."): on = key in ln; continue
        if not on: continue
        m = re.search(r'//## File "([^"]*)", line (\d+)', ln)
        if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m and cur: cnt[cur] += 1; ops[cur][m.group(1).split(".")[0]] += 1
print("total", sum(cnt.values()))
for (f, l), n in cnt.most_common(top):
    print(f"{n:6d} {f}:{l}  " + " ".join(f"{o}:{c}" for o, c in ops[(f, l)].most_common(5)))
