"""Build the C oracle (test infrastructure) -> oracle/_build/libfrx_oracle.so.

Two builds of the same source:
  * the portable one (``-O2``), the checker every test uses; it travels to the GPU box with the snapshot;
  * a host-tuned one (``-O3 -march=native``, contraction still off so the results stay bit-identical) that
    bench.py times as the CPU arm.  ``-march=native`` code must not travel between machines, so it is
    compiled on first use ON the machine that runs it, into a file keyed by that CPU's feature flags.

``oracle/_ref`` is intentionally absent: the reference is pure Python plus un-vendored third-party
wheels (frenetix 0.4.0, commonroad-drivability-checker 2024.1), there are no reference C/C++ sources
to compile."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c", "frx_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libfrx_oracle.so")
COMMON = ["-fopenmp", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC"]


def _compile(out, opt, force):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(SRC):
        return out
    tmp = f"{out}.{os.getpid()}.tmp"
    cmd = ["gcc"] + opt + COMMON + ["-o", tmp, SRC, "-lm"]
    print("[oracle build]", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, out)          # atomic: several ranks may build at once
    return out


def build(force=False):
    return _compile(OUT, ["-O2"], force)


ABI_SRC = os.path.join(HERE, "cpu_abi", "frx_cpu.c")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")


def build_cpu_abi(native=False, force=False):
    """The C ABI of include/frx.h on the host cores (oracle/cpu_abi/frx_cpu.c + the port) -> libfrx_cpu_omp[_native_<cpu>].so.
    Baseline / test infrastructure: bench.py's reference arm times the CPU path through the same calls as the GPU path."""
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, f"libfrx_cpu_omp_native_{_cpu_key()}.so" if native else "libfrx_cpu_omp.so")
    newest = max(os.path.getmtime(SRC), os.path.getmtime(ABI_SRC), os.path.getmtime(os.path.join(INCLUDE, "frx.h")))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    tmp = f"{out}.{os.getpid()}.tmp"
    cmd = ["gcc"] + (["-O3", "-march=native"] if native else ["-O2"]) + COMMON + ["-I", INCLUDE, "-o", tmp, ABI_SRC, SRC, "-lm"]
    print("[oracle build]", " ".join(cmd), file=sys.stderr, flush=True)
    subprocess.check_call(cmd)
    os.replace(tmp, out)
    return out


def _cpu_key() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return hashlib.sha1(line.encode()).hexdigest()[:12]
    except OSError:
        pass
    return "generic"


def build_native(force=False):
    """-O3 -march=native for the CPU this process runs on (the timed CPU arm of bench.py)."""
    return _compile(os.path.join(OUT_DIR, f"libfrx_oracle_native_{_cpu_key()}.so"), ["-O3", "-march=native"], force)


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_cpu_abi(force="--force" in sys.argv)
