"""Compile the C oracle (oracle/ck_oracle.c) into oracle/libck_oracle.so with gcc. TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ck_oracle.c")
OUT = os.path.join(HERE, "libck_oracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.isfile(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-fopenmp", "-fvisibility=hidden",
           "-o", OUT, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
