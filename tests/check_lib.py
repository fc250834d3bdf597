"""Routes the product wrappers (deepavfusion_b200.kernels) to the CUDA-core CHECKER kernels of tests/check/libdavf_check.so.

TEST INFRASTRUCTURE ONLY.  The checkers take the argument structs of include/davf.h, so ``with check_lib.routed("gemm"):``
makes ``K.gemm`` / ``K.gemm_grouped`` (or ``K.attention_fwd`` / ``K.attention_bwd``) run the independent CUDA-core
implementation with exactly the arguments the product call would have received."""
import contextlib
import ctypes as C
import os
import subprocess

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "check")
LIB = os.path.join(HERE, "libdavf_check.so")


def build():
    r = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if r.returncode != 0 or not os.path.exists(LIB):
        raise RuntimeError("building libdavf_check.so failed:\n" + r.stderr[-2000:])
    return LIB


def _load():
    import torch  # noqa: F401  (libcudart)
    from deepavfusion_b200 import _cabi
    if not os.path.exists(LIB):
        build()
    h = C.CDLL(LIB)
    h.davf_check_gemm.argtypes = [C.POINTER(_cabi.GemmArgs), C.c_void_p]
    h.davf_check_gemm_grouped.argtypes = [C.POINTER(_cabi.GemmArgs), C.c_int, C.c_void_p]
    h.davf_check_attention_fwd.argtypes = [C.POINTER(_cabi.AttnFwdArgs), C.c_void_p]
    h.davf_check_attention_bwd.argtypes = [C.POINTER(_cabi.AttnBwdArgs), C.c_void_p]
    h.davf_check_last_error.restype = C.c_char_p
    return h


class _Routed:
    def __init__(self, real, over):
        self._real, self._over = real, over

    def __getattr__(self, name):
        return self._over.get(name) or getattr(self._real, name)


@contextlib.contextmanager
def routed(what: str):
    """what = 'gemm' or 'attention'."""
    from deepavfusion_b200 import _cabi
    h = _load()
    real = _cabi.lib()
    over = {"gemm": {"davf_gemm": h.davf_check_gemm, "davf_gemm_grouped": h.davf_check_gemm_grouped},
            "attention": {"davf_attention_fwd": h.davf_check_attention_fwd, "davf_attention_bwd": h.davf_check_attention_bwd}}[what]
    _cabi._lib = _Routed(real, over)
    try:
        yield
    finally:
        _cabi._lib = real
