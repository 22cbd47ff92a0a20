"""Executed warp-instructions of iou_matrix_kernel per pipeline STAGE, from an ncu report (--import-source on) and the object it was
built from: joins the SASS page with nvdisasm line info like ncu_lines.py, then maps source lines to stages through the stage
markers in iou.cu (comments "---- stage N", "compact the group's survivors", ...) and the functions of geom.cuh / emu.cuh.
Instructions attributed to CUDA headers (intrinsics) inherit the stage of the preceding instruction.
usage: iou_stage_table.py <ncu-rep> <iou.cu.o> [kernel-symbol-substring] [iou.cu as it was when the object was built]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, obj = sys.argv[1:3]
sym = sys.argv[3] if len(sys.argv) > 3 else "iou_matrix_kernelILb1ELi0ELb0"
src = open(sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "r3det-pytorch_b200", "csrc", "iou.cu")).read().splitlines()


def line_of(marker, start=0):
    return next(i + 1 for i, l in enumerate(src) if i >= start and marker in l)


k0 = line_of("__global__ void __launch_bounds__(IOU_THREADS, R3G_IOU_MINB) iou_matrix_kernel")
marks = sorted([(k0, "item loop: ticket, staging (cp.async), column constants"),
                (line_of("---- stage 4", k0), "stage 4: reference restatement (queue handling; the call itself is emu.cuh)"),
                (line_of("Items (64 rows x 128 columns)", k0), "item loop: ticket, staging (cp.async), column constants"),
                (line_of("---- stage 1", k0), "stage 1: circumradius sign bits + zero store"),
                (line_of("compact the group's survivors", k0), "compaction of the survivor masks into the queue"),
                (line_of("---- queue stages", k0), "queue dispatch"),
                (line_of("---- stage 3", k0), "stage 3: area integral + epilogue + store"),
                (line_of("---- stage 2", k0), "stage 2: separating-axis test"),
                (line_of("while (c3 > 0) stage4", k0), "tail: last flagged pairs, counters")])
k1 = line_of("__global__ void iou_aligned_kernel")


def stage_iou(line):
    if line < k0:
        return "emit / helpers (inlined into stage 3 / 4)"
    if line >= k1:
        return None
    name = None
    for l, n in marks:
        if line >= l:
            name = n
    return name


geom = open(os.path.join(ROOT, "r3det-pytorch_b200", "csrc", "geom.cuh")).read().splitlines()
g_sat0 = next(i + 1 for i, l in enumerate(geom) if "R3G_HD bool pair_sat" in l)
g_sat1 = next(i + 1 for i, l in enumerate(geom) if "R3G_HD float edge_term" in l)


def stage_of(f, line, prev):
    if f == "iou.cu":
        return stage_iou(line) or prev
    if f == "geom.cuh":
        return "stage 2: separating-axis test" if g_sat0 <= line < g_sat1 - 2 else "stage 3: area integral + epilogue + store"
    if f == "emu.cuh":
        return "stage 4: reference restatement (queue handling; the call itself is emu.cuh)"
    return prev


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
        lines.append((cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, ti = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
ssi = hdr.index("# Samples") if "# Samples" in hdr else None
body = [r for r in rows[hi + 1:] if len(r) > ci and (r[0].startswith("0x") or r[0].isdigit())]
print(f"SASS instructions: object {len(lines)}, report {len(body)}" + ("" if len(lines) == len(body) else "   (MISMATCH: the object is not the profiled build)"))
agg, thr, smp, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
prev, tot = "item loop: ticket, staging (cp.async), column constants", 0
for k in range(min(len(lines), len(body))):
    (cur, text) = lines[k]
    st = stage_of(cur[0], cur[1], prev) if cur else prev
    prev = st
    c = int(float(body[k][ci] or 0))
    tot += c
    agg[st] += c
    thr[st] += int(float(body[k][ti] or 0))
    if ssi is not None:
        smp[st] += int(float(body[k][ssi] or 0))
    op = text.split()[0] if not text.startswith("@") else text.split()[1]
    ops[st][op.split(".")[0]] += c
print(f"total executed warp-instructions: {tot}")
stot = sum(smp.values()) or 1
print(f"{'stage':86s} {'warp-instr':>12s} {'share':>6s} {'lanes':>6s} {'stall samples':>13s}")
for st, c in agg.most_common():
    top = ", ".join(f"{o} {n * 100 // max(c, 1)}%" for o, n in ops[st].most_common(5))
    print(f"{st:86s} {c:12d} {100.0 * c / tot:5.1f}% {thr[st] / max(c, 1):6.1f} {100.0 * smp[st] / stot:12.1f}%   [{top}]")
