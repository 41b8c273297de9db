"""The r=4 lookup kernel divides by (size-1) with a correctly rounded reciprocal and one
exact-residual fma step instead of div.rn.  Prove on the CPU that the two agree for every
float32 mantissa and every denominator the fast path accepts (map sides up to 2048)."""
import ctypes
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reciprocal_fma_division_is_correctly_rounded():
    src = os.path.join(HERE, "native", "div_check.c")
    with tempfile.TemporaryDirectory() as td:
        so = os.path.join(td, "div_check.so")
        subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", src, "-o", so, "-lm"],
                       check=True)
        lib = ctypes.CDLL(so)
        lib.div_trick_mismatches.restype = ctypes.c_longlong
        lib.div_trick_mismatches.argtypes = [ctypes.c_int, ctypes.c_int]
        assert lib.div_trick_mismatches(1, 2047) == 0
