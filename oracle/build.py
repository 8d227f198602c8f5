"""Build the C oracle (test infrastructure) -> oracle/_build/libfrx_oracle.so.

``oracle/_ref`` is intentionally absent: the reference is pure Python plus un-vendored third-party
wheels (frenetix 0.4.0, commonroad-drivability-checker 2024.1), there are no reference C/C++ sources
to compile."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "frx_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libfrx_oracle.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"]
    print("[oracle build]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv)
