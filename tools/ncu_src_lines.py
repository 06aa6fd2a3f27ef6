"""Per-CUDA-source-line samples / executed instructions of one kernel from an ncu report captured with
--import-source on.  Usage: python tools/ncu_src_lines.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys, os
rep, kre = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                      "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr, recs = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; nh = len(r); continue
    if hdr and r[0].isdigit():
        def g(k, r=r):   # index from the end: unescaped quotes in the source column can split it
            v = r[len(r) - (nh - hdr[k])]
            try: return int(float(v))
            except ValueError: return 0
        recs.append((cur_file, int(r[0]), r[1].strip()[:90], g("# Samples"), g("Instructions Executed"),
                     {k: g(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_mio",
                                        "stall_math", "stall_not_selected", "stall_selected", "stall_no_inst", "stall_lg")}))
ts = sum(x[3] for x in recs) or 1
ti = sum(x[4] for x in recs) or 1
print(f"total samples {ts}  total warp instructions {ti/1e6:.1f}M")
for f, ln, s, smp, ins, st in sorted(recs, key=lambda x: -x[3])[:top]:
    tops = ", ".join(f"{k[6:]}:{100*v/max(smp,1):.0f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*smp/ts:5.2f}% smp {100*ins/ti:5.2f}% ins  {f}:{ln:<4d} {s}   [{tops}]")
