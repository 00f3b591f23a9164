"""TEST INFRASTRUCTURE ONLY — builds oracle/_build/libssrs_oracle.so from oracle/ssrs_oracle.c with gcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ssrs_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libssrs_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
