"""Build the C oracle (TEST INFRASTRUCTURE ONLY) into oracle/_build/libqcat_oracle.so.

oracle/_ref/ (a build of the reference's own sources) does not exist for this reference: qcat is pure
Python and its alignment primitive lives in the third-party parasail C library, which is neither
vendored under /root/reference nor installable offline -- see DESIGN.md "Oracle".
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libqcat_oracle.so")


def build(force=False):
    src = os.path.join(HERE, "qcat_oracle.c")
    hdr = os.path.join(HERE, "qcat_oracle.h")
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-Wall", "-Wextra",
           "-o", LIB, src]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
