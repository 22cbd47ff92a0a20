"""Build libr3geo.so — the hand-written sm_100a CUDA kernels + C ABI (include/r3geo.h).

    python r3det-pytorch_b200/build.py [--force] [--verbose]

In-tree build with nvcc (cross-compiles without a GPU); the .so is git-ignored but travels to the GPU
box with the gpurun snapshot.  There is no CPU fallback: if this library is missing the package raises.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libr3geo.so")
SOURCES = ["api.cu", "iou.cu", "nms.cu", "frm.cu", "transforms.cu", "coder.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC,-fvisibility=hidden,-O3", "--expt-relaxed-constexpr",
              "-Xptxas", "-v", "-rdc=false"]


def _stale(out, deps):
    return (not os.path.exists(out)) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps)


def main(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "r3geo.h"))
    objdir = os.path.join(HERE, "_build")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = ["nvcc"] + NVCC_FLAGS + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for s, p in procs:
        log, _ = p.communicate()
        with open(os.path.join(objdir, os.path.basename(s) + ".log"), "w") as f:
            f.write(log)
        if p.returncode != 0:
            ok = False
            print(f"[build] {os.path.basename(s)} FAILED\n{log[-6000:]}")
        elif verbose:
            print(log)
    if not ok:
        return False
    if force or procs or _stale(OUT, objs):
        cmd = ["nvcc", "-shared", "-o", OUT] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            print("[build] link FAILED\n" + r.stderr[-4000:])
            return False
        print(f"[build] libr3geo.so built ({len(procs)} unit(s) recompiled)")
    else:
        print("[build] libr3geo.so up to date")
    return True


if __name__ == "__main__":
    sys.exit(0 if main("--force" in sys.argv, "--verbose" in sys.argv) else 1)
