"""Summarise an ncu launch-list CSV (--metrics gpu__time_duration.sum) per kernel: launches, total time, share.
usage: launch_list_summary.py <csv> [title]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
body = [r for r in rows[h + 1:] if len(r) > 10]
KN = rows[h].index("Kernel Name")          # (the column moves when --nvtx adds its own)
agg = collections.OrderedDict()
for r in body:
    name = re.sub(r"\(.*", "", r[KN]).strip()[:72]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
print("%d launches, %.1f us in total (cold-cache, serialised: shares are meaningful, absolute times are not)" % (len(body), tot / 1e3))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-72s launches %4d  total %10.1f us  share %5.1f%%" % (name, n, t / 1e3, 100 * t / tot))
# the headline step = r3g_iou_matrix_f32 = prep_pair_kernel + iou_matrix_kernel<1, 0, 0>  (round 1 / early round 2: prep_boxes_kernel x2)
names = [re.sub(r"\(.*", "", r[KN]).strip() for r in body]
steps = [i for i, n in enumerate(names) if "iou_matrix_kernel<1, 0" in n and i >= 1 and "prep_pair" in names[i - 1]]
if steps:
    i = steps[min(4, len(steps) - 1)]          # a timed headline step (after the warm-up launches)
    d = [float(body[j][-1]) for j in (i - 1, i)]
    print("\nheadline step (one timed r3g_iou_matrix_f32 call): prep_pair_kernel %.1f us, iou_matrix_kernel %.1f us -> pair kernel share %.1f%%"
          % (d[0] / 1e3, d[1] / 1e3, 100 * d[1] / sum(d)))
