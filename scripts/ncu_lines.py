"""Join an ncu SASS-page CSV export with nvdisasm line info: executed warp-instructions per CUDA source line.
usage: ncu_lines.py <ncu-rep> <object.o> <kernel-symbol-substring> [top]"""
import csv, io, re, subprocess, sys, tempfile, os, collections
rep, obj, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False
for l in dis:
    if l.startswith("//---------------------"):
        on = sym in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "(.*)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = hdr.index("Instructions Executed"); si = hdr.index("Source"); ti = hdr.index("Thread Instructions Executed")
ssi = hdr.index("# Samples") if "# Samples" in hdr else None
body = [r for r in rows[hi + 1:] if len(r) > ci and r[0].startswith("0x") or (len(r) > ci and r[0].isdigit())]
print("sass instrs: nvdisasm", len(lines), "ncu", len(body))
n = min(len(lines), len(body))
agg = collections.Counter(); thr = collections.Counter(); smp = collections.Counter()
tot = 0
for k in range(n):
    c = int(float(body[k][ci] or 0)); tot += c
    agg[lines[k][1]] += c; thr[lines[k][1]] += int(float(body[k][ti] or 0))
    if ssi is not None: smp[lines[k][1]] += int(float(body[k][ssi] or 0))
print("total warp-instructions", tot)
for key, c in agg.most_common(top):
    print("%-14s %5d  %12d  %5.1f%%  lanes %.1f  samples %d" % (key[0], key[1], c, 100.0 * c / tot, thr[key] / max(c, 1), smp[key]))
