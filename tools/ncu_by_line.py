"""Aggregate ncu warp-stall samples of one kernel by CUDA source line, using nvdisasm -g line info of the
same build.  Usage: python tools/ncu_by_line.py report.ncu-rep object.o kernel_regex mangled_substring"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, obj, kre, mangled = sys.argv[1:5]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > idx["stall_wait"] and r[idx["# Samples"]].isdigit()]
a0 = data[0][idx["Address"]]
rep2 = [i for i, r in enumerate(data) if r[idx["Address"]] == a0]
if len(rep2) > 1: data = data[:rep2[1]]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, os.listdir(tmp)[0])], capture_output=True, text=True).stdout
lines, cur, infn, out = dis.split("\n"), None, False, []
for l in lines:
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: infn = mangled in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l): out.append(cur)
print("sass instrs: ncu", len(data), "nvdisasm", len(out))
n = min(len(data), len(out))
I = lambda r, k: int(float(r[idx[k]] or 0))
tot = sum(I(r, "# Samples") for r in data) or 1
agg = collections.Counter(); ex = collections.Counter(); stall = collections.defaultdict(collections.Counter)
for i in range(n):
    agg[out[i]] += I(data[i], "# Samples"); ex[out[i]] += I(data[i], "Instructions Executed")
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio", "stall_no_inst", "stall_math", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_lg"):
        stall[out[i]][k] += I(data[i], k)
for k, v in agg.most_common(40):
    top = ", ".join(f"{a.replace('stall_','')}:{100*b/max(v,1):.0f}%" for a, b in stall[k].most_common(3))
    print(f"{100*v/tot:6.2f}%  exec {ex[k]/1e6:8.1f}M  {k[0]}:{k[1]}   [{top}]")
