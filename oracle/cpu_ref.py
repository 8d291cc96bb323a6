"""ctypes wrapper of oracle/liboracle_cpu_ref.so (C restatement of the reference's
CPU algorithms).  TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see cpu_ref.c."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle_cpu_ref.so")
_lib = None

SIZES = {1: {1: 96, 2: 48, 3: 96, 4: 104}, 2: {1: 192, 2: 96, 3: 192, 4: 200}}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise RuntimeError("oracle/liboracle_cpu_ref.so missing: run `make -C oracle`")
        l = C.CDLL(_PATH)
        l.oracle_convert.restype = C.c_int
        l.oracle_convert.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_uint,
                                     C.c_void_p, C.c_int]
        l.oracle_generate.restype = C.c_int
        l.oracle_generate.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_uint64, C.c_size_t, C.c_void_p,
                                      C.c_int]
        l.oracle_fq_op.restype = C.c_int
        l.oracle_fq_op.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_void_p]
        _lib = l
    return _lib


def convert(group, in_fmt, data, out_fmt, checks, nthreads=1):
    """-> (out bytes, list of per-point status codes)"""
    data = bytes(data)
    ri, ro = SIZES[group][in_fmt], SIZES[group][out_fmt]
    n = len(data) // ri
    out = C.create_string_buffer(max(n * ro, 1))
    st = C.create_string_buffer(max(n, 1))
    rc = lib().oracle_convert(group, in_fmt, data, out_fmt, out, n, checks, st, nthreads)
    if rc != 0:
        raise ValueError("oracle_convert rc=%d" % rc)
    return out.raw[: n * ro], list(st.raw[:n])


def generate(group, fmt, scalar0, step, first, n, nthreads=1):
    ro = SIZES[group][fmt]
    out = C.create_string_buffer(max(n * ro, 1))
    rc = lib().oracle_generate(group, fmt, int(scalar0).to_bytes(32, "little"), int(step).to_bytes(32, "little"),
                               first, n, out, nthreads)
    if rc != 0:
        raise ValueError("oracle_generate rc=%d" % rc)
    return out.raw[: n * ro]
