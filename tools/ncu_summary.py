"""Summarise an .ncu-rep: headline metrics + SASS opcode mix + stall reasons.  python tools/ncu_summary.py rep [iters]"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
iters = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
        "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum"]
print("kernel:", m.get("Kernel Name", ("?",))[0])
for k in keys:
    if k in m:
        print(f"  {k:70s} {m[k][0]:>18s} {m[k][1]}")
if iters:
    print(f"  warp instructions per solver iteration: {float(m['smsp__inst_executed.sum'][0]) / iters:.0f}")
    print(f"  SM cycles per solver iteration per chain: {float(m['smsp__cycles_active.avg'][0]) * 4 * 148 / 72 / (iters / 72):.0f} (approx)")
# warp state / stall reasons
st = {h: float(v[0]) for h, v in m.items() if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")}
if st:
    print("  stall cycles per issued instruction:")
    for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]:
        print(f"    {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):32s} {v:.2f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))[2:]
ex, stl = collections.Counter(), collections.Counter()
for r in rows:
    mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
    op = mm.group(2).split(".")[0] if mm else "?"
    ex[op] += int(r[5]); stl[op] += int(r[2])
tot, tots = sum(ex.values()), max(1, sum(stl.values()))
print(f"  SASS instructions in kernel: {len(rows)}; opcode mix (executed %, stall samples %):")
for op, n in ex.most_common(18):
    print(f"    {op:8s} {100 * n / tot:5.1f}  {100 * stl[op] / tots:5.1f}")
