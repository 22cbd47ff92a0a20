"""SASS opcode histogram per kernel of libr3geo.so (cuobjdump -sass): the mnemonics that show which hardware paths a kernel uses
(UTMALDG / SYNCS = TMA + mbarrier, FFMA2 / FADD2 = packed f32x2, LDGSTS = cp.async, ATOMG / RED, ...).
usage: sass_histogram.py [kernel-substring ...] > profiles/rNN_sass_opcode_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "r3det-pytorch_b200", "libr3geo.so")
want = sys.argv[1:] or ["iou_matrix_kernel", "nms_rounds_kernel", "frm_forward_tma_kernel", "frm_backward_tma_kernel"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
KEY = ("UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FADD2", "FMUL2", "LDGSTS", "ATOMG", "ATOMS", "RED", "UTCMMA", "HMMA", "LDG", "STG", "LDS",
       "STS", "LDL", "STL", "SHFL", "VOTE", "FFMA", "FADD", "FMUL", "FMNMX", "MUFU", "BAR", "BRA")
for fn, h in hist.items():
    name = demangle(fn)
    if not any(w in name for w in want):
        continue
    tot = sum(h.values())
    base = collections.Counter()
    for op, c in h.items():
        base[op.split(".")[0]] += c
    print(f"## {name[:150]}\n   {tot} SASS instructions")
    print("   key mnemonics: " + ", ".join(f"{k} {base[k]}" for k in KEY if base[k]))
    tma = [f"{op} {c}" for op, c in h.items() if op.startswith(("UTMA", "SYNCS", "UBLKCP"))]
    if tma:
        print("   TMA / mbarrier forms: " + ", ".join(tma))
    print("   top 12: " + ", ".join(f"{op} {c}" for op, c in base.most_common(12)) + "\n")
