"""Summarise an .ncu-rep: key raw metrics + executed-instruction mix by opcode + stall samples.
usage: python tools/ncu_summary.py <rep> <n_traj_steps_per_launch> [traj_per_warp=64]"""
import collections, csv, re, subprocess, sys

rep, nts = sys.argv[1], float(sys.argv[2])
tpw = float(sys.argv[3]) if len(sys.argv) > 3 else 64.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "lts__t_bytes.sum", "l1tex__t_bytes.sum", "lts__t_sector_hit_rate.pct"]
print("== raw metrics ==")
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {vals[i]:>18s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]; ix = {h: i for i, h in enumerate(h2)}
byop = collections.Counter(); samp = collections.Counter(); stall = collections.Counter(); tot = 0
scols = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[2:]:
    if len(r) < len(h2):
        continue
    m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op = m.group(2) if m else "?"
    ex = int(r[ix["Instructions Executed"]] or 0)
    byop[op] += ex; tot += ex; samp[op] += int(r[ix["# Samples"]] or 0)
    for c in scols:
        stall[c] += int(r[ix[c]] or 0)
ws = nts / tpw
print(f"== executed warp instructions: {tot}  = {tot / ws:.1f} per (warp x step) = {tot / ws / tpw * 32:.1f} thread-instr per trajectory-step ==")
for op, c in byop.most_common(28):
    print(f"{op:10s} {c / ws:8.2f}   samples {samp[op]}")
print("== stall samples ==", dict(stall.most_common(10)))
