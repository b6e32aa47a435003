"""Compile oracle/yc_oracle.c into oracle/libyc_oracle.so (TEST INFRASTRUCTURE ONLY).

The reference (xin-pu/yolo-continuous) is 100 % Python: there is no C/C++ reference
source to compile into oracle/_ref, so oracle/_ref is never produced.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "yc_oracle.c")
LIB = os.path.join(HERE, "libyc_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden",
           "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
