"""ncu reports -> profiles/r02_kernel_traffic.json: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the
kernels bench.py reports a roofline for.  usage: ncu_traffic.py name=report.ncu-rep[:kernel-substring] ... (merges into the file)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_kernel_traffic.json")
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
res = json.load(open(OUT)) if os.path.exists(OUT) else {}
for arg in sys.argv[1:]:
    name, rest = arg.split("=", 1)
    rep, _, sub = rest.partition(":")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if sub and sub not in d.get("Kernel Name", ""):
            continue
        def val(k):
            return float(d[k].replace(",", "")) * UNIT[units[hdr.index(k)]]
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        res[name] = {"dram_bytes": int(rd + wr), "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
                     "kernel": d.get("Kernel Name", "")[:120], "duration_us": float(d["gpu__time_duration.sum"].replace(",", "")),
                     "source": f"ncu --set full --clock-control none ({os.path.basename(rep)}; summary under profiles/)"}
        break
json.dump(res, open(OUT, "w"), indent=1)
print(json.dumps(res, indent=1))
