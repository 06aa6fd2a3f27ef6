"""Summarise an .ncu-rep: per-kernel headline metrics, stall mix, and time share per barrier-delimited
segment (from the source page).  Usage: python tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
out = []
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    out.append(["kernel", name])
    for k in keys:
        if k in idx:
            out.append([k, r[idx[k]], units[idx[k]]])
    st = [(float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
          for h, i in idx.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    tot = sum(v for v, _ in st)
    for v, h in sorted(st, reverse=True)[:8]:
        out.append(["stall_" + h, f"{v:.3f}", f"{100*v/tot:.1f}% of warp time"])
    kn = "forward" if "forward" in name else "backward"
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kn], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    sh = srows[1]; sidx = {h: i for i, h in enumerate(sh)}
    data = [x for x in srows[2:] if len(x) > sidx["stall_wait"] and x[sidx["# Samples"]].isdigit()]
    addr0 = data[0][sidx["Address"]]
    second = [i for i, x in enumerate(data) if x[sidx["Address"]] == addr0]
    if len(second) > 1: data = data[:second[1]]
    I = lambda x, k: int(float(x[sidx[k]] or 0))
    total = sum(I(x, "# Samples") for x in data) or 1
    prev = 0
    bidx = [i for i, x in enumerate(data) if "BAR.SYNC" in x[sidx["Source"]]]
    for b in bidx + [len(data) - 1]:
        seg = data[prev:b + 1]
        s = sum(I(x, "# Samples") for x in seg)
        ex = sum(I(x, "Instructions Executed") for x in seg)
        ff = sum(I(x, "Instructions Executed") for x in seg if "FFMA" in x[sidx["Source"]])
        if s > 0.01 * total:
            out.append([f"segment[{prev}:{b}]", f"{100*s/total:.1f}% of samples", f"instr {ex/1e6:.1f}M ffma {ff/1e6:.1f}M"])
        prev = b + 1
    top = sorted(data, key=lambda x: -I(x, "# Samples"))[:12]
    for x in top:
        out.append(["hot", f"{100*I(x,'# Samples')/total:.2f}%", x[sidx["Source"]].strip()[:80]])
for o in out: print(",".join(str(v) for v in o))
if len(sys.argv) > 2:
    with open(sys.argv[2], "w") as f:
        csv.writer(f).writerows(out)
