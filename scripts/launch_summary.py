"""Print the kernels of the LAST repetition in an ncu launch-list CSV (--metrics gpu__time_duration.sum): name, ns, total.
usage: launch_summary.py <csv> <name-substring-of-the-last-kernel-of-a-repetition>"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
body = [r for r in rows[h + 1:] if len(r) > 10]
names = [r[4] for r in body]
ends = [i for i, n in enumerate(names) if sys.argv[2] in n]
last, prev = ends[-1], (ends[-2] if len(ends) > 1 else -1)
tot = 0.0
for r in body[prev + 1:last + 1]:
    print(r[4][:72].ljust(72), r[-1]); tot += float(r[-1])
print("sum us %.1f over %d launches" % (tot / 1000, last - prev))
