"""Instruction / sample shares of a kernel per source-line range: ncu_phase_shares.py report.ncu-rep name:line name:line ...
(lines of the first source file of the report's source page, ascending; everything before the first mark is 'head')."""
import bisect, csv, subprocess, sys
rep = sys.argv[1]
marks = [("head", 1)] + [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[2:]]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
for i, r in enumerate(rows):
    if 'Instructions Executed' in r:
        h = r; start = i + 1; break
ci = h.index('Instructions Executed'); csamp = h.index('# Samples')
starts = [m[1] for m in marks]
acc = {m[0]: [0, 0] for m in marks}
tot = 0
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in rows[start:]:
    if len(r) <= ci or not r[0].strip().isdigit(): continue
    k = bisect.bisect_right(starts, int(r[0])) - 1
    acc[marks[k][0]][0] += I(r[ci]); acc[marks[k][0]][1] += I(r[csamp]); tot += I(r[ci])
print("total warp instructions", tot)
for k, v in acc.items():
    print(f"{k:16s} inst {v[0] / tot * 100:5.1f}%  samples {v[1]}")
