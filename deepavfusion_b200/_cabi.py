"""ctypes binding of libdavf_sm100.so (the C ABI declared in include/davf.h).

The library is the product; there is no fallback.  ``lib()`` raises ``RuntimeError`` when the
shared object is missing or cannot be loaded -- callers never route around it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdavf_sm100.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

c_i64p = C.POINTER(C.c_int64)
vp = C.c_void_p


class LnFwdArgs(C.Structure):
    _fields_ = [("x0", vp), ("bs0", C.c_int64), ("n0", C.c_int),
                ("x1", vp), ("bs1", C.c_int64), ("n1", C.c_int),
                ("B", C.c_int), ("D", C.c_int), ("eps", C.c_float),
                ("gamma", vp), ("beta", vp),
                ("y_bf16", vp), ("y_f32", vp), ("mean", vp), ("rstd", vp),
                ("nseg", C.c_int), ("seg_start", C.c_int * 5)]


class LnBwdArgs(C.Structure):
    _fields_ = [("x0", vp), ("bs0", C.c_int64), ("n0", C.c_int),
                ("x1", vp), ("bs1", C.c_int64), ("n1", C.c_int),
                ("B", C.c_int), ("D", C.c_int),
                ("gamma", vp), ("mean", vp), ("rstd", vp),
                ("dy_bf16", vp), ("dy_f32", vp),
                ("dx0", vp), ("dbs0", C.c_int64), ("add0", vp),
                ("dx1", vp), ("dbs1", C.c_int64), ("add1", vp),
                ("dgamma", vp), ("dbeta", vp),
                ("nseg", C.c_int), ("seg_start", C.c_int * 5),
                ("dx0_bf16", vp), ("dx1_bf16", vp)]


class GemmArgs(C.Structure):
    _fields_ = [("a", vp), ("lda", C.c_int64), ("a_kmajor", C.c_int),
                ("b", vp), ("ldb", C.c_int64), ("b_kmajor", C.c_int),
                ("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64),
                ("bias", vp), ("act", C.c_int),
                ("aux_out", vp), ("aux_in", vp), ("ldaux", C.c_int64),
                ("res", vp), ("ldres", C.c_int64), ("res_idx", vp),
                ("out", vp), ("ldo", C.c_int64), ("out_bf16", C.c_int), ("accumulate", C.c_int),
                ("g", C.c_int), ("G", C.c_int), ("off", C.c_int),
                ("split_k", C.c_int), ("rowsum_out", vp), ("debug_clocks", vp)]


class AttnFwdArgs(C.Structure):
    _fields_ = [("q", vp), ("q_bs", C.c_int64), ("q_rs", C.c_int64),
                ("k", vp), ("k_bs", C.c_int64), ("k_rs", C.c_int64),
                ("v", vp), ("v_bs", C.c_int64), ("v_rs", C.c_int64),
                ("o", vp), ("o_bs", C.c_int64), ("o_rs", C.c_int64),
                ("lse", vp),
                ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("dqk", C.c_int), ("dv", C.c_int),
                ("scale", C.c_float), ("accumulate", C.c_int)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("q", vp), ("q_bs", C.c_int64), ("q_rs", C.c_int64),
                ("k", vp), ("k_bs", C.c_int64), ("k_rs", C.c_int64),
                ("v", vp), ("v_bs", C.c_int64), ("v_rs", C.c_int64),
                ("d_o", vp), ("do_bs", C.c_int64), ("do_rs", C.c_int64),
                ("lse", vp),
                ("dq", vp), ("dq_bs", C.c_int64), ("dq_rs", C.c_int64),
                ("dk", vp), ("dk_bs", C.c_int64), ("dk_rs", C.c_int64),
                ("dv_", vp), ("dv_bs", C.c_int64), ("dv_rs", C.c_int64),
                ("B", C.c_int), ("H", C.c_int), ("Nq", C.c_int), ("Nk", C.c_int), ("dqk", C.c_int), ("dv", C.c_int),
                ("scale", C.c_float), ("accumulate_dq", C.c_int),
                ("o", vp), ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("dq_dead_rows", C.c_int)]


i, i64, f = C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/davf.h declares
SIGNATURES = {
    "davf_last_error": (C.c_char_p, []),
    "davf_version": (i, []),
    "davf_device_sm": (i, []),
    "davf_set_gemm_2cta": (i, [i]),
    "davf_set_gemm_sms": (i, [i]),
    "davf_set_attn_impl": (i, [i]),
    "davf_launch_count": (i64, []),
    "davf_launch_count_kind": (i64, [i]),
    "davf_set_pdl": (i, [i]),
    "davf_mask_rank": (i, [vp, i, i, i, vp, vp, vp, vp]),
    "davf_patch_rows": (i, [vp, vp, vp, i, i, i, i, i, i, vp]),
    "davf_cast_rows_bf16": (i, [vp, vp, i64, i, i, i, i, vp]),
    "davf_sum_cast": (i, [vp, vp, vp, vp, vp, i64, vp]),
    "davf_colsum_bf16": (i, [vp, i64, i, i64, vp, vp]),
    "davf_batchsum_f32": (i, [vp, i, i, i, i, i, vp, i, vp]),
    "davf_layernorm_fwd": (i, [C.POINTER(LnFwdArgs), vp]),
    "davf_layernorm_bwd": (i, [C.POINTER(LnBwdArgs), vp]),
    "davf_gemm": (i, [C.POINTER(GemmArgs), vp]),
    "davf_gemm_grouped": (i, [C.POINTER(GemmArgs), i, vp]),
    "davf_attention_fwd": (i, [C.POINTER(AttnFwdArgs), vp]),
    "davf_attention_bwd": (i, [C.POINTER(AttnBwdArgs), vp]),
    "davf_decoder_assemble_fwd": (i, [vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]),
    "davf_decoder_assemble_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]),
    "davf_masked_mse_fwd": (i, [vp, vp, vp, vp, i, i, i, i, i, i, i, i, vp]),
    "davf_masked_mse_bwd": (i, [vp, vp, vp, vp, f, vp, i, i, i, i, i, i, i, i, vp]),
    "davf_adamw_step": (i, [vp, vp, vp, vp, vp, i64, vp, vp, vp, f, f, f, i, vp, vp]),
    "davf_sumsq_f32": (i, [vp, i64, vp, vp]),
    "davf_cast_flat_bf16": (i, [vp, vp, i64, vp]),
    "davf_meanpool_fwd": (i, [vp, i64, i, i, i, vp, vp]),
    "davf_meanpool_bwd": (i, [vp, i, i, i, vp, vp]),
    "davf_batchnorm1d_fwd": (i, [vp, i, i, i, vp, vp, f, f, vp, vp, vp, vp]),
    "davf_batchnorm1d_bwd": (i, [vp, vp, vp, vp, i, i, i, vp, vp]),
    "davf_head_fwd": (i, [vp, vp, vp, i, i, i, vp, vp]),
    "davf_head_bwd": (i, [vp, vp, vp, i, i, i, vp, vp, vp, vp]),
    "davf_logmel_workspace_bytes": (i64, []),
    "davf_logmel_init": (i, [vp, i, i, i, i, vp]),
    "davf_logmel_fwd": (i, [vp, vp, vp, vp, i, i, i, i, f, vp, vp]),
    "davf_image_normalize_u8": (i, [vp, vp, i, i, i, i, C.POINTER(C.c_float), C.POINTER(C.c_float), vp]),
    "davf_scale_rows_add": (i, [vp, vp, vp, i, i64, i, vp, vp]),
    "davf_scale_rows": (i, [vp, vp, i, i64, i, vp, vp, vp]),
}

_lib = None
_lock = threading.Lock()


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into deepavfusion_b200/lib/ (nvcc cross-compiles
    without a GPU).  Returns the path of the shared library."""
    r = subprocess.run(["make", "-j8", "-C", CSRC_DIR], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0 or not os.path.exists(LIB_PATH):
        raise RuntimeError("building libdavf_sm100.so failed:\n" + r.stderr[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    """Load the library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: the CUDA extension is the product and has no fallback. "
                    "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc).")
            import torch  # noqa: F401  (loads libcudart.so.12 that the library links against)
            handle = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(handle, name)      # AttributeError if the export is missing
                fn.restype = res
                fn.argtypes = args
            if handle.davf_version() != 1:
                raise RuntimeError("libdavf_sm100.so ABI version mismatch")
            _lib = handle
    return _lib


class DavfError(RuntimeError):
    pass


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().davf_last_error()
        raise DavfError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")
