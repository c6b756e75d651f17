"""Build recipe for the CPU oracle (test infrastructure, see doa_oracle.cpp header).

    python oracle/build.py            -> oracle/_build/libdoa_oracle.so

The reference itself cannot be compiled here (needs GNU Radio, Boost, Armadillo, OPINCAA, cmake-driven
generated code; none present), so there is no oracle/_ref; this restatement is the only CPU arm.
Flags: AVX2 baseline (not -march=native: the .so travels to the GPU box, whose host CPU may differ) and
-ffp-contract=off so the port's own loops round the same way on every machine.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
SRC = os.path.join(HERE, "doa_oracle.cpp")
OUT = os.path.join(OUT_DIR, "libdoa_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["g++", "-std=c++17", "-O3", "-mavx2", "-mfma", "-ffp-contract=off", "-fcx-limited-range", "-fopenmp", "-fPIC", "-shared",
           "-Wall", "-o", OUT, SRC, "-ldl"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
