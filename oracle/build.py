"""TEST INFRASTRUCTURE ONLY.  Compile the plain-C oracle: oracle/r3oracle.c -> oracle/_build/libr3oracle.so

    python oracle/build.py [--force]

-O2 -ffp-contract=off -fno-fast-math: same floating-point behaviour as the reference's host build
(torch extensions default to -O2 on x86-64 without FMA contraction).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "r3oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libr3oracle.so")


def main(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > os.path.getmtime(SRC):
        print("[oracle] libr3oracle.so: up to date")
        return True
    cmd = ["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ffp-contract=off",
           "-fno-fast-math", "-Wall", "-Wextra", SRC, "-o", OUT, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr)
        return False
    if r.stderr.strip():
        print(r.stderr)
    print("[oracle] libr3oracle.so: built")
    return True


if __name__ == "__main__":
    sys.exit(0 if main("--force" in sys.argv) else 1)
