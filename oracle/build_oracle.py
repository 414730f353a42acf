"""TEST INFRASTRUCTURE ONLY - compiles oracle/cpu_kernels.c into oracle/_build/liboracle_cpu.so (gcc)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_cpu.so")


def build(force=False):
    src = os.path.join(HERE, "cpu_kernels.c")
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-shared", "-fPIC", src, "-o", LIB])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
