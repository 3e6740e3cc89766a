"""Warp-stall samples of one kernel, grouped by the code between barriers.

usage: ncu_phases.py report.ncu-rep [n_top]
Reads the SASS source page of a `--set full --import-source on` capture.  A warp waiting at a
barrier is sampled at the instruction BEHIND the barrier, so the `barrier` column of a phase is
time spent waiting for the phase before it to finish on the slowest warp of the CTA."""
import csv, io, subprocess, sys

rep = sys.argv[1]
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print("kernel:", rows[0][1][:120])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, "# Samples") for r in data)
print(f"samples {tot}, warp instructions {sum(num(r, 'Instructions Executed') for r in data) / 1e6:.1f} M")
print("phase (SASS index range, ends at a BAR.SYNC) | samples | share | of which barrier wait | warp instructions")
start = 0
for i, r in enumerate(data):
    if "BAR" in r[ix["Source"]] or i == len(data) - 1:
        seg = data[start:i + 1]
        s = sum(num(d, "# Samples") for d in seg)
        bar = sum(num(d, "stall_barrier") for d in seg)
        n = sum(num(d, "Instructions Executed") for d in seg)
        print(f"  {start:4d}-{i:4d} | {s:7d} | {100 * s / max(tot, 1):5.1f} % | {bar:6d} | {n / 1e6:7.1f} M")
        start = i + 1
print(f"top {n_top} instructions by samples (index, samples, barrier / long / short scoreboard / mio, SASS):")
for i in sorted(sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:n_top]):
    r = data[i]
    print(f"  {i:4d} {num(r, '# Samples'):6d}  {num(r, 'stall_barrier'):5d} {num(r, 'stall_long_sb'):5d} "
          f"{num(r, 'stall_short_sb'):5d} {num(r, 'stall_mio'):5d}  {r[ix['Source']].strip()[:70]}")
