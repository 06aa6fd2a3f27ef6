"""Per-kernel SASS opcode histogram of a built object (cuobjdump -sass): instruction count and the mnemonics that
show what the kernel is made of (FFMA2 packed FP32, HMMA = mma.sync, UBLKCP = cp.async.bulk, SYNCS = mbarrier,
LDGSTS = cp.async, REDG = red.global, LDL/STL = spills; UTC*MMA / LDTM / UTMALDG would be tcgen05 / TMEM / tensor TMA).
Usage: python tools/sass_stats.py <object> [name filter ...]"""
import collections, re, subprocess, sys
obj, pats = sys.argv[1], sys.argv[2:]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, stats = None, collections.OrderedDict()
for l in out.split("\n"):
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        fn = m.group(1); stats[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", l)
    if m and fn:
        op, mod = m.group(1), m.group(2)
        stats[fn]["n"] += 1
        key = op
        if op == "HMMA": key = "HMMA" + ".".join(mod.split(".")[:2] + [x for x in mod.split(".") if x in ("TF32", "F32")][-1:])
        if op in ("RED", "REDG", "ATOMG"): key = op + mod.replace(".E", "")[:24]
        stats[fn][key] += 1
show = ("FFMA2", "FFMA", "FMUL2", "FADD2", "HMMA", "LDS", "STS", "LDG", "STG", "LDGSTS", "UBLKCP", "SYNCS", "BAR", "RED", "ATOM",
        "SHFL", "LDL", "STL", "MUFU", "UTC", "LDTM", "STTM", "UTMA", "CALL")
for fn, c in stats.items():
    if c["n"] < 200 or (pats and not any(p in fn for p in pats)):
        continue
    print(f"{fn}\n   instructions {c['n']} ({c['n'] * 16 // 1024} KB)")
    agg = collections.Counter()
    for k, v in c.items():
        for s in show:
            if k.startswith(s):
                agg[k if s in ("HMMA", "RED", "ATOM") else s] += v
                break
    print("   " + "  ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])))
