"""Per-kernel SASS statistics of a built object: instruction count, FFMA, LDS, local-memory traffic."""
import re, subprocess, sys, collections, os, tempfile
obj = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
for cub in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    fn, stats = None, collections.OrderedDict()
    for l in dis.split("\n"):
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            fn = m.group(1); stats[fn] = collections.Counter(); continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m and fn:
            op = m.group(1)
            stats[fn]["n"] += 1
            for k in ("FFMA", "LDS", "STS", "LDL", "STL", "LDG", "STG", "SHFL", "BAR", "CALL"):
                if op.startswith(k): stats[fn][k] += 1
    for fn, c in stats.items():
        if pat in fn and c["n"] > 50:
            print(f"{fn[:70]:70s} n={c['n']:6d} ({c['n']*16//1024} KB) " + " ".join(f"{k}={c[k]}" for k in ("FFMA","LDS","STS","LDL","STL","LDG","STG","SHFL","BAR","CALL")))
