"""ctypes binding of the C ABI in include/qtt_b200.h (csrc/libqtt_b200.so).

There is no CPU fallback: if the shared library is missing or CUDA is unavailable every hot-path call
raises.  `build_library()` compiles the library in-tree with nvcc for sm_100a (used by
`__graft_entry__.build()`).
"""
import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libqtt_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

QTT_OK = 0
_ERR_NAMES = {1: "QTT_ERR_INVALID", 2: "QTT_ERR_CUDA", 3: "QTT_ERR_SLOT", 4: "QTT_ERR_UNSUPPORTED", 5: "QTT_ERR_NOCONV"}

EXPORTS = ["qtt_create", "qtt_destroy", "qtt_last_error", "qtt_version", "qtt_set_workspace_limit",
           "qtt_get_counters", "qtt_get_flop_counters", "qtt_get_graph_counters", "qtt_reset_counters", "qtt_set_profiling", "qtt_get_profile", "qtt_get_class_profile", "qtt_matmul_batched", "qtt_round_sum_batched",
           "qtt_orthogonalize_batched", "qtt_inner_batched", "qtt_axpby_batched", "qtt_gd_solve_batched", "qtt_cg_solve_batched",
           "qtt_encode_linear_batched", "qtt_encode_trig_batched"]


class QttError(RuntimeError):
    pass


class qtt_batch(Structure):
    _fields_ = [("L", c_int), ("N", c_int), ("n", POINTER(c_int)), ("ranks", c_void_p),
                ("cores", POINTER(c_void_p)), ("slot", POINTER(c_longlong))]


class qtt_mpo(Structure):
    _fields_ = [("L", c_int), ("n_out", POINTER(c_int)), ("n_in", POINTER(c_int)), ("ranks", POINTER(c_int)),
                ("cores", POINTER(c_void_p))]


class qtt_term(Structure):
    _fields_ = [("op", POINTER(qtt_mpo)), ("x", POINTER(qtt_batch)), ("coef", c_void_p), ("coef_scale", c_double)]


def build_library(force=False, verbose=False):
    """Compile csrc/qtt_b200.cu -> csrc/libqtt_b200.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(CSRC, "qtt_b200.cu")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "qtt_b200.h"))
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH, src]
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library (once).  Raises QttError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QttError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback for the TT hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.qtt_create.argtypes = [POINTER(c_void_p), c_int]
    lib.qtt_destroy.argtypes = [c_void_p]
    lib.qtt_last_error.argtypes = [c_void_p]
    lib.qtt_last_error.restype = c_char_p
    lib.qtt_version.restype = c_int
    lib.qtt_set_workspace_limit.argtypes = [c_void_p, c_size_t]
    lib.qtt_get_counters.argtypes = [c_void_p, POINTER(c_double), POINTER(c_longlong)]
    lib.qtt_get_flop_counters.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double)]
    lib.qtt_get_graph_counters.argtypes = [c_void_p, POINTER(c_longlong), POINTER(c_longlong)]
    lib.qtt_reset_counters.argtypes = [c_void_p]
    lib.qtt_set_profiling.argtypes = [c_void_p, c_int]
    lib.qtt_get_profile.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_longlong), c_int]
    lib.qtt_get_class_profile.argtypes = [c_void_p, POINTER(c_double), c_int]
    lib.qtt_matmul_batched.argtypes = [c_void_p, POINTER(qtt_batch), POINTER(qtt_mpo), POINTER(qtt_batch), POINTER(c_int),
                                       POINTER(qtt_batch), c_void_p]
    lib.qtt_round_sum_batched.argtypes = [c_void_p, c_int, POINTER(qtt_term), POINTER(qtt_batch), POINTER(c_int), c_double,
                                          c_void_p, c_void_p]
    lib.qtt_orthogonalize_batched.argtypes = [c_void_p, c_int, POINTER(qtt_term), POINTER(qtt_batch), c_void_p, c_void_p]
    lib.qtt_inner_batched.argtypes = [c_void_p, POINTER(qtt_batch), POINTER(qtt_mpo), POINTER(qtt_batch), c_void_p, c_void_p]
    lib.qtt_axpby_batched.argtypes = [c_void_p, c_void_p, POINTER(qtt_batch), c_void_p, POINTER(qtt_batch), POINTER(qtt_batch),
                                      c_void_p]
    lib.qtt_gd_solve_batched.argtypes = [c_void_p, POINTER(qtt_mpo), POINTER(qtt_batch), POINTER(qtt_batch), c_void_p, c_int,
                                         c_double, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]
    lib.qtt_cg_solve_batched.argtypes = [c_void_p, POINTER(qtt_mpo), POINTER(qtt_batch), POINTER(qtt_batch), c_void_p, c_int,
                                         c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]
    lib.qtt_encode_linear_batched.argtypes = [c_void_p, c_int, c_void_p, POINTER(c_void_p), POINTER(c_int), POINTER(qtt_batch), c_void_p]
    lib.qtt_encode_trig_batched.argtypes = [c_void_p, c_void_p, POINTER(qtt_batch), c_void_p]
    for name in EXPORTS:
        if name not in ("qtt_last_error",):
            getattr(lib, name).restype = c_int if name != "qtt_last_error" else c_char_p
    lib.qtt_last_error.restype = c_char_p
    _lib = lib
    return lib


def check(handle, rc, what):
    if rc != QTT_OK:
        msg = load().qtt_last_error(handle)
        raise QttError(f"{what} failed: {_ERR_NAMES.get(rc, rc)}: {msg.decode() if msg else ''}")
