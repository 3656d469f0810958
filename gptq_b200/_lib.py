"""ctypes binding of libgptq_b200.so (the C ABI declared in include/gptq_b200.h).

There is NO fallback: if the library is missing or a call fails, this raises.  The library is
built in-tree by ``python -m gptq_b200.build`` (``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgptq_b200.so")

F16, BF16, F32 = 0, 1, 2
ABI_VERSION = 3
SYRK_AUTO, SYRK_CTA1, SYRK_CTA2, SYRK_CTA2_BK128 = 0, 1, 2, 3
_DT = {torch.float16: F16, torch.bfloat16: BF16, torch.float32: F32}

_vp, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/gptq_b200.h declares (tests check).
SIGNATURES = {
    "gptq_b200_abi_version": (_i, []),
    "gptq_b200_last_error": (C.c_char_p, []),
    "gptq_b200_launch_count": (_i64, []),
    "gptq_b200_syrk_workspace_bytes": (_sz, [_i, _i64, _i]),
    "gptq_b200_syrk_accum": (_i, [_vp, _i, _i64, _i, _i64, _vp, _i64, _f, _f, _i, _i, _vp, _sz, _vp]),
    "gptq_b200_syrk_accum_simt": (_i, [_vp, _i, _i64, _i, _i64, _vp, _i64, _f, _f, _vp]),
    "gptq_b200_symmetrize": (_i, [_vp, _i, _i64, _vp]),
    "gptq_b200_tril_pack": (_i, [_vp, _i, _i64, _vp, _vp]),
    "gptq_b200_tril_unpack": (_i, [_vp, _i, _vp, _i64, _vp]),
    "gptq_b200_stage_input": (_i, [_vp, _i, _i64, _i, _i64, _vp, _i64, _vp]),
    "gptq_b200_compare_rows": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _vp, _vp]),
    "gptq_b200_hessian_prepare": (_i, [_vp, _i, _i64, _f, _vp, _vp, _vp, _vp]),
    "gptq_b200_weights_prepare": (_i, [_vp, _i, _i, _i, _i64, _vp, _vp, _i64, _vp]),
    "gptq_b200_dead_columns": (_i, [_vp, _i, _i64, _vp, _vp]),
    "gptq_b200_argsort_desc": (_i, [_vp, _i, _vp, _vp, _vp]),
    "gptq_b200_permute_cols": (_i, [_vp, _i, _i, _i64, _vp, _vp, _i64, _vp]),
    "gptq_b200_hessian_gather": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _i, _i64, _vp]),
    "gptq_b200_hinv_upper_workspace_bytes": (_sz, [_i]),
    "gptq_b200_hinv_upper": (_i, [_vp, _i, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "gptq_b200_find_params": (_i, [_vp, _i, _i, _i64, _i, _i, _i, _i, _i, _i, _f, _i, _f, _vp, _vp, _i64, _i, _vp]),
    "gptq_b200_dequant_codes": (_i, [_vp, _i64, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i64, _vp]),
    "gptq_b200_quantize": (_i, [_vp, _i, _i, _i64, _vp, _vp, _i64, _i64, _i, _vp, _i64, _vp]),
    "gptq_b200_colloop_workspace_bytes": (_sz, [_i, _i, _i]),
    "gptq_b200_colloop": (_i, [_vp, _i, _i, _i64, _vp, _i64, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _f,
                               _vp, _i64, _vp, _i64, _vp, _vp, _sz, _vp]),
    "gptq_b200_selftest_div": (_i, [C.c_uint64, C.c_uint32, _i, _vp, _vp]),
    "gptq_b200_finalize": (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp, _vp, _i, _i64, _vp, _i64, _vp]),
    "gptq_b200_pack_codes": (_i, [_vp, _i, _i, _i64, _i, _vp, _vp]),
    "gptq_b200_pack_weights": (_i, [_vp, _i, _i, _i, _i64, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    "gptq_b200_unpack_codes": (_i, [_vp, _i, _i, _i, _vp, _i64, _vp]),
    "gptq_b200_vecquant_matvec": (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    "gptq_b200_sgemm": (_i, [_i, _i, _i, _i, _i, _f, _vp, _i64, _vp, _i64, _f, _vp, _i64, _vp]),
    "gptq_b200_gemm3x": (_i, [_i, _i, _i, _i, _i, _f, _vp, _i64, _vp, _i64, _f, _vp, _i64, _vp]),
    "gptq_b200_set_gemm_backend": (_i, [_i]),
    "gptq_b200_set_chain_schedule": (_i, [_i, _i]),
}

_lib = None


class GptqB200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GptqB200Error(
                f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m gptq_b200.build`). "
                "gptq_b200 has no CPU or PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        ver = lib.gptq_b200_abi_version()
        if ver != ABI_VERSION:
            raise GptqB200Error(f"ABI version mismatch: library {ver}, binding {ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gptq_b200_last_error().decode(errors="replace")
        raise GptqB200Error(f"{what} failed (rc={rc}): {msg}")


def dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported dtype {t.dtype}") from None


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise GptqB200Error(f"{what} must be a CUDA tensor: gptq_b200 runs on sm_100a only and has no CPU path")


def launch_count() -> int:
    return int(load().gptq_b200_launch_count())
