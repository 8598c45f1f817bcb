"""Aggregate the per-instruction warp-stall samples of one profiled launch (ncu --page source --csv) : totals per stall reason and
the hottest SASS instructions.  usage: ncu_stalls.py report.ncu-rep launch_index [top_n]"""
import csv, subprocess, sys, collections
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
lines = [l for l in raw.splitlines() if not l.startswith("==")]
print(lines[0][:160])
rd = list(csv.reader(lines[1:]))
hdr = rd[0]; rows = [r for r in rd[1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
c = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in rows:
    for h in stall_cols:
        tot[h] += int(r[c[h]] or 0)
allv = sum(tot.values())
print(f"instructions: {len(rows)}  samples: {allv}")
print("  ".join(f"{h[6:]}={v / allv:.1%}" for h, v in tot.most_common(10)))
srows = sorted(rows, key=lambda r: -int(r[c["# Samples"]] or 0))[:top]
for r in srows:
    s = int(r[c["# Samples"]] or 0)
    why = sorted(((int(r[c[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{s:6d} {s / allv:6.1%} exec={r[c['Instructions Executed']]:>8s} {r[c['Source']].strip()[:90]:90s} {why[0][1]}:{why[0][0]} {why[1][1]}:{why[1][0]}")
