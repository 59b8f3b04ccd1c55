"""TEST INFRASTRUCTURE — compiles oracle/hotpath_oracle.c with gcc into oracle/_build/ (git-ignored)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hotpath_oracle.c")
OUT = os.path.join(HERE, "_build", "libhotpath_oracle.so")


def build(force=False):
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"])
    return OUT


def load():
    lib = ctypes.CDLL(build())
    i64, i32, fp, ip = ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    lib.oracle_visibility_count.argtypes = [fp, i64, fp, fp, i32, ctypes.c_float, ctypes.c_float, ip]
    lib.oracle_visibility_count.restype = None
    lib.oracle_composite_blend.argtypes = [fp, fp, fp, i32, i64, i32, fp, fp, fp]
    lib.oracle_composite_blend.restype = None
    return lib


if __name__ == "__main__":
    print(build(force=True))
