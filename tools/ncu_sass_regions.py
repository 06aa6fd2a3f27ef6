"""Opcode histogram (executed warp instructions, stall samples) of one kernel from an ncu report, split at a SASS
address: python tools/ncu_sass_regions.py report.ncu-rep kernel_regex [split_opcode]  — the first occurrence of
split_opcode (default HMMA) minus a margin separates "before" (producers) from "after" (consumers)."""
import csv, io, subprocess, sys, collections, re
rep, kre = sys.argv[1:3]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["Instructions Executed"] and r[ix["Instructions Executed"]].replace('.', '').isdigit()]
def op(r):
    s = r[ix["Source"]].strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0] if s else "?"
tot_i = sum(int(float(r[ix["Instructions Executed"]])) for r in data)
tot_s = sum(int(float(r[ix["# Samples"]])) for r in data)
print(f"instructions {tot_i/1e6:.1f}M samples {tot_s}")
# find region boundaries by role: walk and mark instructions lying between first and last HMMA as consumer
hm = [i for i, r in enumerate(data) if op(r) == "HMMA"]
lo, hi = (hm[0], hm[-1]) if hm else (len(data), len(data))
# consumer region: extend backwards to the closest preceding BRA-heavy boundary is unknowable; report three zones
zones = {"before_first_HMMA": data[:lo], "HMMA_zone": data[lo:hi + 1], "after_last_HMMA": data[hi + 1:]}
for name, z in zones.items():
    zi = sum(int(float(r[ix["Instructions Executed"]])) for r in z); zs = sum(int(float(r[ix["# Samples"]])) for r in z)
    print(f"\n== {name}: {len(z)} SASS lines, {zi/1e6:.1f}M instr ({100*zi/max(tot_i,1):.1f}%), {100*zs/max(tot_s,1):.1f}% samples")
    h = collections.Counter(); hs = collections.Counter()
    for r in z:
        h[op(r)] += int(float(r[ix["Instructions Executed"]])); hs[op(r)] += int(float(r[ix["# Samples"]]))
    for k, v in h.most_common(22):
        print(f"   {k:10s} {v/1e6:8.2f}M instr {100*v/max(zi,1):5.1f}%   samples {100*hs[k]/max(tot_s,1):5.2f}%")
