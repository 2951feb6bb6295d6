"""Builds oracle/_build/liboracle.so from oracle/c/pbx_oracle.c with gcc + OpenMP.
TEST INFRASTRUCTURE (see oracle/__init__.py)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pbx_oracle.c")
OUTDIR = os.path.join(os.path.dirname(HERE), "_build")
LIB = os.path.join(OUTDIR, "liboracle.so")


def build(force=False):
    os.makedirs(OUTDIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC,
           "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
