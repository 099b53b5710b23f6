"""ctypes binding of libfs2b200.so (the C ABI declared in include/fs2b200.h).

There is deliberately NO fallback here: if the CUDA library is missing or a call fails, the product
path raises.  The only pure-Python pieces of this package are host logic (shapes, descriptors).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "csrc", "libfs2b200.so")

c_i32, c_i64, c_f32, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p


class Operand(ctypes.Structure):
    _fields_ = [
        ("ptr", c_vp), ("ld", c_i64), ("batch_stride", c_i64), ("inner", c_i32), ("rows", c_i32),
        ("batches", c_i32), ("mn_major", c_i32), ("inner_base", c_i32), ("zdiv", c_i32),
        ("zmod_stride", c_i32),
    ]


class Gemm(ctypes.Structure):
    _fields_ = [
        ("a", Operand), ("b", Operand), ("mode", c_i32), ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("Z", c_i32), ("taps", c_i32), ("tap_shift0", c_i32), ("b_tap_kstride", c_i32),
        ("splits", c_i32), ("epilogue", c_i32), ("d_f32", c_i32), ("d_atomic", c_i32),
        ("d_zdiv", c_i32), ("alpha", c_f32), ("d", c_vp), ("ldd", c_i64), ("d_col_stride", c_i64),
        ("d_tap_stride", c_i64), ("d_zdiv_stride", c_i64), ("d_zmod_stride", c_i64),
        ("bias", c_vp), ("aux", c_vp), ("ld_aux", c_i64), ("aux_batch_stride", c_i64),
        ("d_seg_rows", c_i32), ("d_seg_pad", c_i32), ("d_seg", c_vp * 4),
        ("row_lens", c_vp), ("lens_zdiv", c_i32), ("tail_zero_rows", c_i32), ("relu_mask", c_vp),
        ("workspace", c_vp), ("workspace_bytes", c_i64),
        ("ln_gamma", c_vp), ("ln_beta", c_vp), ("ln_res", c_vp), ("ld_res", c_i64), ("res_batch_stride", c_i64),
        ("ln_p_drop", c_f32), ("ln_pad0", c_i32), ("ln_seed", ctypes.c_uint64), ("ln_seed_dev", c_vp),
        ("ln_v", c_vp), ("ln_mean", c_vp), ("ln_rstd", c_vp), ("ln_keep", c_vp),
        ("a_colsum", c_vp), ("a_colsum_seg", c_vp * 4),
    ]


GEMM_NORMAL, GEMM_WGRAD = 0, 1
EPI_NONE, EPI_RELU, EPI_RELU_BWD, EPI_ADD_AUX = 0, 1, 2, 3

_lib = None


class Fs2Error(RuntimeError):
    pass


def so_path():
    return _SO


def lib():
    """Load the library once.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise Fs2Error(
                "libfs2b200.so is not built (%s). Run `python __graft_entry__.py build` -- this "
                "package has no CPU / eager fallback." % _SO)
        _lib = ctypes.CDLL(_SO)
        _declare(_lib)
    return _lib


def check(rc, what):
    if rc != 0:
        raise Fs2Error("%s failed (rc=%d): %s" % (what, rc, lib().fs2_last_error().decode()))


_CTYPE = {"int": ctypes.c_int, "int32_t": c_i32, "int64_t": c_i64, "uint64_t": ctypes.c_uint64,
          "float": c_f32}


def header_path():
    return os.path.join(os.path.dirname(_HERE), "include", "fs2b200.h")


def parse_header(path=None):
    """{name: (restype, [argtypes])} for every function declared in include/fs2b200.h.

    The header is the single source of truth for the ABI; the ctypes prototypes are derived from it.
    """
    import re

    text = open(path or header_path()).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(fs2_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    argtypes.append(c_vp)
                else:
                    argtypes.append(_CTYPE[a.replace("const ", "").split(" ")[0]])
        restype = {"int": ctypes.c_int, "int64_t": c_i64, "const char*": ctypes.c_char_p}[ret]
        protos[name] = (restype, argtypes)
    return protos


def _declare(L):
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(L, name)  # AttributeError here = header and library disagree: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    L.fs2_gemm_bf16.argtypes = [ctypes.POINTER(Gemm), ctypes.c_int, c_vp]


def launch_count():
    return int(lib().fs2_launch_count())
