"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
Usage: python tools/launch_summary.py launches.csv "command that was profiled" """
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = re.sub(r"\(.*", "", r[4])[:70]
    val = float(r[14].replace(",", ""))
    ms = val / 1e6 if r[13] in ("ns", "nsecond") else (val / 1e3 if r[13] in ("us", "usecond") else val)
    tot[name] += ms; cnt[name] += 1
allms = sum(tot.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none: {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("(cold-cache serialised launches: compare shares, not absolutes)")
for n, ms in tot.most_common(14):
    print(f"{100*ms/allms:6.2f}%  {ms:10.3f} ms  x{cnt[n]:4d}  {n}")
