"""TEST INFRASTRUCTURE ONLY.  Build the UNMODIFIED reference CPU ops into oracle/_ref/*.so.

The reference sources are compiled where they lie under /root/reference (they are
``#include``d by absolute path from the small shim TUs in oracle/refshim/, nothing is
copied into this repository).  Outputs go to oracle/_ref/ (git-ignored; travels to the GPU
box with the gpurun snapshot).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs may load these libraries.

    python oracle/build_ref.py            # build what is missing / stale
    python oracle/build_ref.py --force
    python oracle/build_ref.py --cuda     # also the reference CUDA kernels for sm_100 (minutes each)

Exposed C entry points (see oracle/refshim/*.cpp):
    libref_v1.so     ref_v1_iou_matrix_f32/_f64, ref_v1_iou_aligned_f32, ref_v1_nms_f32
    libref_v3iou.so  ref_v3_iou_matrix_f32/_f64, ref_v3_iou_matrix_tensor_f32
    libref_v3nms.so  ref_v3_nms_f32, ref_v3nms_iou_matrix_f32
    libref_v2.so     ref_v2_iou_matrix_f32, ref_v2_nms_f32              (torch-free)
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("R3REF_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OPS = os.path.join(REF, "r3det", "ops")

TARGETS = {
    # name: (shim, macro, reference file, needs_torch)
    "libref_v1.so": ("ref_v1.cpp", "R3REF_RNMS_CPU", f"{OPS}/rnms/src/rcpu/rnms_cpu.cpp", True),
    "libref_v3iou.so": ("ref_v3iou.cpp", "R3REF_BOX_IOU_ROTATED_CPU",
                        f"{OPS}/box_iou_rotated/src/box_iou_rotated_cpu.cpp", True),
    "libref_v3nms.so": ("ref_v3nms.cpp", "R3REF_NMS_ROTATED_CPU",
                        f"{OPS}/nms_rotated/src/nms_rotated_cpu.cpp", True),
    "libref_v2.so": ("ref_v2.cpp", "R3REF_ML_UTILS_H",
                     f"{OPS}/ml_nms_rotated/src/box_iou_rotated_utils.h", False),
}


def _torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    libdir = ce.library_paths()[0]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    cflags = inc + [f"-D_GLIBCXX_USE_CXX11_ABI={abi}", "-DTORCH_EXTENSION_NAME=r3ref_unused"]
    ldflags = [f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-lc10", f"-Wl,-rpath,{libdir}"]
    return cflags, ldflags


def build_one(name, force=False):
    shim, macro, ref_file, needs_torch = TARGETS[name]
    out = os.path.join(OUT, name)
    src = os.path.join(HERE, "refshim", shim)
    if not os.path.exists(ref_file):
        return name, "skipped (reference tree not present)"
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(ref_file))):
        return name, "up to date"
    # -O2 without -ffast-math / -march: the reference builds its extensions with torch's default -O2.
    # -ffp-contract=off: x86-64 baseline has no FMA anyway; stated for determinism.
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
           f"-D{macro}=\"{ref_file}\"", f"-I{os.path.dirname(ref_file)}", src, "-o", out, "-w"]
    if needs_torch:
        c, l = _torch_flags()
        cmd = cmd[:-3] + c + cmd[-3:] + l
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return name, "FAILED\n" + r.stderr[-4000:]
    return name, "built"


CUDA_TARGETS = {
    # name: (shim, macro, reference .cu)  — the reference CUDA kernels, compiled unmodified for sm_100
    "libref_cuda_v1iou.so": ("refcuda_v1iou.cu", "R3REF_RBBOX_GEO_KERNEL", f"{OPS}/rbbox_geo/src/rbbox_geo_kernel.cu"),
    "libref_cuda_frm.so": ("refcuda_frm.cu", "R3REF_FEATURE_REFINE_KERNEL", f"{OPS}/fr/src/feature_refine_kernel.cu"),
    "libref_cuda_v3iou.so": ("refcuda_v3iou.cu", "R3REF_BOX_IOU_ROTATED_CUDA",
                             f"{OPS}/box_iou_rotated/src/box_iou_rotated_cuda.cu"),
    "libref_cuda_v3nms.so": ("refcuda_v3nms.cu", "R3REF_NMS_ROTATED_CUDA",
                             f"{OPS}/nms_rotated/src/nms_rotated_cuda.cu"),
    # these two include <THC/THC.h>, gone from torch since 1.11: refshim/thc/ holds a stand-in (allocation / ceil-div
    # helpers only), the kernels and host loops are the reference's own
    "libref_cuda_polynms.so": ("refcuda_polynms.cu", "R3REF_POLY_NMS_CUDA", f"{OPS}/nms_rotated/src/poly_nms_cuda.cu"),
    "libref_cuda_v1nms.so": ("refcuda_v1nms.cu", "R3REF_RNMS_KERNEL", f"{OPS}/rnms/src/rcuda/rnms_kernel.cu"),
}


def build_cuda_one(name, force=False):
    """Reference CUDA kernels (rbbox_geo, fr, box_iou_rotated, nms_rotated) — the kernels to beat on B200."""
    shim, macro, ref_file = CUDA_TARGETS[name]
    out = os.path.join(OUT, name)
    src = os.path.join(HERE, "refshim", shim)
    if not os.path.exists(ref_file):
        return name, "skipped (reference tree not present)"
    if (not force and os.path.exists(out)
            and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(ref_file))):
        return name, "up to date"
    c, l = _torch_flags()
    # same defines the reference's setup.py passes (setup.py:33-37); arch = sm_100 (it ships no arch flags)
    cmd = ["nvcc", "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-w",
           "-gencode", "arch=compute_100,code=sm_100",
           "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
           f"-D{macro}=\"{ref_file}\"", f"-I{os.path.dirname(ref_file)}", f"-I{os.path.join(HERE, 'refshim')}",
           f"-I{os.path.join(HERE, 'refshim', 'thc')}"]
    cmd += c + [src, "-o", out] + [x.replace("-Wl,-rpath,", "-Xlinker=-rpath=") for x in l]
    cmd += ["-ltorch_cuda", "-lc10_cuda", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        return name, "FAILED\n" + r.stderr[-4000:]
    return name, "built"


def main(force=False, cuda=False):
    os.makedirs(OUT, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        res = list(ex.map(lambda n: build_one(n, force), TARGETS))
        if cuda:
            res += list(ex.map(lambda n: build_cuda_one(n, force), CUDA_TARGETS))
    ok = True
    for n, s in res:
        print(f"[oracle/_ref] {n}: {s}")
        ok &= not s.startswith("FAILED")
    return ok


if __name__ == "__main__":
    sys.exit(0 if main("--force" in sys.argv, "--cuda" in sys.argv) else 1)
