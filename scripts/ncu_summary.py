"""Summarise an .ncu-rep of the solve kernel: key metrics of the last captured launch and an
attribution of executed instructions / stall samples to source lines grouped by function.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep = sys.argv[1]
WANT = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, last = rows[0], rows[1], rows[-1]
print(f"kernel: {last[hdr.index('Kernel Name')]}  (launch id {last[hdr.index('ID')]} of {len(rows) - 2} captured)\n")
for w in WANT + sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
    if w in hdr:
        print(f"{w:90s} {units[hdr.index(w)]:16s} {last[hdr.index(w)]}")

# ---- SASS page: classify instructions by execution count.  Instructions inside the per-stage serial
# loops execute ~N+1 times per interior-point iteration, the stage-parallel phases ~2 times (two
# passes of 32 lanes over 51 stages), everything else (start-up, barriers, rare branches) far less.
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks = src.split('"Kernel Name"')
rd = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[-1])))
h = rd[1]
ci, cs = h.index("Instructions Executed"), h.index("# Samples")
stall_cols = {c: h.index(c) for c in ("stall_barrier", "stall_wait", "stall_long_sb", "stall_math", "stall_no_inst", "stall_short_sb", "stall_not_selected")}
ins = [r for r in rd[2:] if len(r) > ci and r[ci].isdigit()]
cmax = max(int(r[ci]) for r in ins)
groups = {"serial sweeps (count > max/5)": lambda c: c > cmax / 5, "stage-parallel phases": lambda c: cmax / 100 <= c <= cmax / 5,
          "rare (count < max/100: start-up, retire loop, barriers' slow paths)": lambda c: c < cmax / 100}
tot_i = sum(int(r[ci]) for r in ins)
tot_s = sum(int(r[cs]) for r in ins)
print(f"\nSASS attribution by execution count (static {len(ins)} instructions, {tot_i} executed, {tot_s} stall samples):")
for name, pred in groups.items():
    g = [r for r in ins if pred(int(r[ci]))]
    gi, gs_ = sum(int(r[ci]) for r in g), sum(int(r[cs]) for r in g)
    parts = "  ".join(f"{k[6:]} {100.0 * sum(int(r[v]) for r in g) / max(gs_, 1):.0f}%" for k, v in stall_cols.items())
    print(f"  {name:70s} static {len(g):5d}  executed {100.0 * gi / tot_i:5.1f}%  samples {100.0 * gs_ / tot_s:5.1f}%  [{parts}]")
# opcode mix of the serial sweeps
mix = defaultdict(int)
for r in ins:
    if int(r[ci]) > cmax / 5:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h.index("Source")])
        mix[(m.group(1) if m else "?").split(".")[0]] += 1
print("  serial-sweep opcode mix (static):", dict(sorted(mix.items(), key=lambda kv: -kv[1])[:12]))
