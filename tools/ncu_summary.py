"""Condense `ncu -i X.ncu-rep --page raw --csv` into the handful of columns quoted in DESIGN.md / bench.py (per launch)."""
import csv, sys, re, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__average_warp_latency_issue_stalled_no_instruction.pct", "no_inst%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__cycles_active.avg", "cyc_active")]
print(f"# {rep}: one row per profiled launch (ncu --set full --clock-control none; cold-cache, serialised: compare shares not absolutes)")
names = [n for c, n in COLS if c in col]
print("kernel | " + " | ".join(f"{n}[{units[col[c]]}]" for c, n in COLS if c in col))
for r in data:
    k = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("hamt::", "").replace("void ", "")
    print(k[:60] + " | " + " | ".join(r[col[c]] for c, n in COLS if c in col))
