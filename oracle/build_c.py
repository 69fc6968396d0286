"""Compile oracle/tdq_oracle.c (plain C restatement of the kernels' arithmetic) with gcc.  Test infrastructure."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtdq_oracle.so")
SRC = os.path.join(HERE, "tdq_oracle.c")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
