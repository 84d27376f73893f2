#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-source-line instruction/stall shares.  Usage: ncu_summary.py rep [ntiles]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; ntiles = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "lts__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
for k in keys:
    if k in d: print(f"{k:80s} {d[k]:>16s} {u[k]}")
print("-- stalls per issue --")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"  {h.split('stalled_')[1].split('_per_issue')[0]:28s} {float(d[h]):.3f}")
inst = float(d["smsp__inst_executed.sum"]); print(f"warp-instr per tile: {inst/ntiles:.0f}" if ntiles > 1 else "")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
iS, iE, iN = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
try:
    src2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
except Exception:
    src2 = ""
tot = sum(float(r[iE] or 0) for r in data); tots = sum(float(r[iN] or 0) for r in data)
print(f"SASS rows {len(data)}  total inst {tot:.0f}  samples {tots:.0f}")
# opcode histogram weighted by executions
op = collections.Counter(); ops = collections.Counter()
for r in data:
    toks = r[iS].split()
    if not toks: continue
    o = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    o = o.split(".")[0] + ("." + [x for x in o.split(".") if x in ("64", "128")][0] if any(x in ("64", "128") for x in o.split(".")) else "")
    op[o] += float(r[iE] or 0); ops[o] += float(r[iN] or 0)
print("-- executed opcode mix (per tile) and stall-sample share --")
for o, c in op.most_common(28):
    print(f"  {o:12s} {c/ntiles:9.1f}  {100*c/tot:5.1f}%   samples {100*ops[o]/max(tots,1):5.1f}%")
