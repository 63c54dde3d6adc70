"""oracle/build.py -- TEST INFRASTRUCTURE: compiles the CPU oracle (plain C + OpenMP).

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs use this.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "raster_oracle.c")
LIB = os.path.join(HERE, "liboracle.so")


LIB64 = os.path.join(HERE, "liboracle64.so")


def build(force: bool = False, f64: bool = False) -> str:
    lib = LIB64 if f64 else LIB
    if (not force and os.path.exists(lib)
            and os.path.getmtime(lib) >= os.path.getmtime(SRC)):
        return lib
    cmd = ["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", lib, SRC, "-lm"]
    if f64:
        cmd.insert(1, "-DORC_F64")
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force=True))
    print(build(force=True, f64=True))
