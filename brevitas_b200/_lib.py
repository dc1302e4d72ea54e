"""ctypes binding of the C-ABI in ``include/brevitas_b200.h`` (``libbrevitas_b200.so``).

This is the only place the Python host side touches native code.  There is no fallback: if the shared
library is missing or a call returns a non-zero status, a ``RuntimeError`` is raised (the reference instead
swallows a failed native build into a warning and silently uses its Python backend,
src/brevitas/__init__.py:73-82 -- deliberately not reproduced).
"""
import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libbrevitas_b200.so"
# tools/ sweeps point this at libbrevitas_b200_tuning.so (make TUNING=1), which adds bvb_set_tuning
LIB_PATH = os.environ.get("BREVITAS_B200_LIB") or os.path.join(_HERE, LIB_NAME)

# dtype / mode tags of include/brevitas_b200.h
F32, BF16, F16 = 0, 1, 2
OK, EINVAL, EUNSUPPORTED, ECUDA = 0, 1, 2, 3
ROUND, FLOOR, CEIL, ROUND_TO_ZERO, DPU_ROUND = 0, 1, 2, 3, 4
CLAMP_STE, CLAMP_MASKED = 0, 1
OUT_I8, OUT_U8, OUT_I32 = 0, 1, 2

_P, _I, _L, _F, _D = c_void_p, c_int, c_int64, c_float, c_double
_UNARY = [_P, _P, _L, _I, _P]

# name -> (restype, argtypes); must list every function declared in include/brevitas_b200.h
SIGNATURES = {
    "bvb_version": (c_int, []),
    "bvb_last_error": (ctypes.c_char_p, []),
    "bvb_sm_count": (c_int, []),
    "bvb_selftest_div": (c_int, [c_float, ctypes.c_uint32, ctypes.c_uint64, _P, _P]),
    "bvb_selftest_lowp_div": (c_int, [_I, _P, _P]),
    "bvb_debug_packed_constants": (c_int, [_F, _F, _F, _I, _P]),
    "bvb_host_rows_fakequant_fwd_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _L, _L, _F, _F, _F, _F, _F, _I, _I, _I, _P, _L, _P]),
    "bvb_host_pipeline_workspace_bytes": (_L, [_L, _L, _L, _I, _I]),
    "bvb_host_pipeline_create": (c_int, [ctypes.POINTER(c_void_p)]),
    "bvb_host_pipeline_destroy": (c_int, [_P]),
    "bvb_host_rows_fakequant_fwd_bwd_on": (c_int, [_P, _P, _P, _P, _P, _P, _L, _L, _L, _F, _F, _F, _F, _F, _I, _I, _I, _P, _L, _P]),
    "bvb_round_ste_impl": (c_int, _UNARY),
    "bvb_ceil_ste_impl": (c_int, _UNARY),
    "bvb_floor_ste_impl": (c_int, _UNARY),
    "bvb_binary_sign_ste_impl": (c_int, _UNARY),
    "bvb_ternary_sign_ste_impl": (c_int, _UNARY),
    "bvb_round_to_zero_ste_impl": (c_int, _UNARY),
    "bvb_dpu_round_ste_impl": (c_int, _UNARY),
    "bvb_abs_binary_sign_grad_impl": (c_int, _UNARY),
    "bvb_abs_binary_sign_grad_bwd": (c_int, [_P, _P, _P, _L, _I, _P]),
    "bvb_tensor_clamp_ste_impl": (c_int, [_P, _P, _P, _P, _L, _L, _L, _L, _L, _I, _I, _P]),
    "bvb_tensor_clamp_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _L, _I, _P]),
    "bvb_scalar_clamp_ste_impl": (c_int, [_P, _P, _L, _D, _D, _I, _P]),
    "bvb_scalar_clamp_min_ste_impl": (c_int, [_P, _P, _L, _D, _I, _P]),
    "bvb_int_quant_fwd": (c_int, [_P, _P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _I, _P]),
    "bvb_int_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _I, _I, _P]),
    "bvb_int_quant_zpt_fwd": (c_int, [_P, _P, _P, _P, _L, _L, _L, _I, _I, _F, _F, _I, _I, _P]),
    "bvb_int_quant_zpt_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _F, _F, _I, _I, _I, _P]),
    "bvb_int_quant_to_int": (c_int, [_P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _I, _I, _P]),
    "bvb_relu_int_quant_fwd": (c_int, [_P, _P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _I, _P]),
    "bvb_relu_int_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _L, _L, _I, _F, _F, _F, _I, _I, _I, _P]),
    "bvb_rows_absmax_int_quant_fwd": (c_int, [_P, _P, _P, _P, _L, _L, _F, _F, _F, _F, _F, _I, _I, _P]),
    "bvb_rows_absmax_int_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _L, _F, _F, _F, _F, _I, _I, _I, _P]),
    "bvb_tensor_absmax_int_quant_fwd": (c_int, [_P, _P, _P, _P, _L, _I, _F, _F, _F, _F, _F, _I, _I, _P, _P]),
    "bvb_tensor_absmax_int_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _L, _I, _F, _F, _F, _F, _I, _I, _I, _P, _P]),
    "bvb_workspace_bytes": (c_int64, []),
    "bvb_bn_act_quant_workspace_bytes": (c_int64, [_L]),
    "bvb_bn_act_quant_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _F, _F, _I, _P, _L, _I, _P, _P, _P, _L, _L, _F, _F, _F, _I, _I, _I, _P, _P]),
    "bvb_bn_act_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _L, _L, _F, _F, _F, _I, _I, _I, _I, _P, _P]),
    "bvb_binary_quant_fwd": (c_int, [_P, _P, _P, _L, _L, _L, _I, _I, _I, _P]),
    "bvb_binary_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _L, _L, _I, _I, _I, _P]),
    "bvb_general_int_quant_fwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _P]),
    "bvb_general_int_quant_sums": (c_int64, [_L, _L]),
    "bvb_general_int_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _I, _I, _I, _I, _P]),
    "bvb_ternary_quant_fwd": (c_int, [_P, _P, _P, _L, _F, _I, _P]),
    "bvb_ternary_quant_bwd": (c_int, [_P, _P, _P, _P, _P, _L, _F, _I, _P]),
    "bvb_absmax_rows": (c_int, [_P, _P, _L, _L, _I, _P]),
    "bvb_absmax_tensor": (c_int, [_P, _P, _L, _I, _P, _P]),
    "bvb_abs_kth_value_rows": (c_int, [_P, _P, _P, _L, _L, _L, _I, _P, _P]),
    "bvb_kth_value_rows": (c_int, [_P, _P, _P, _L, _L, _L, _I, _P, _P]),
    "bvb_relu_abs_kth_value_rows": (c_int, [_P, _P, _P, _L, _L, _L, _I, _P, _P]),
    "bvb_kth_workspace_bytes": (c_int64, [_L]),
    "bvb_minmax_rows": (c_int, [_P, _P, _P, _P, _P, _L, _L, _I, _P, _P]),
    "bvb_minmax_workspace_bytes": (c_int64, [_L]),
    "bvb_running_stats_update": (c_int, [_P, _P, _L, _F, _F, _I, _I, _P]),
}

_lib = None


def load():
    """Load ``libbrevitas_b200.so`` (once) and bind every entry point.  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `make -C brevitas_b200/csrc -j8` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. brevitas_b200 has no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def is_loaded():
    return _lib is not None


def last_error():
    msg = load().bvb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def call(name, *args):
    """Call an int-status entry point; raise ``RuntimeError`` carrying ``bvb_last_error()`` on failure."""
    rc = getattr(load(), name)(*args)
    if rc != OK:
        kind = {EINVAL: "invalid argument", EUNSUPPORTED: "unsupported", ECUDA: "CUDA error"}.get(rc, f"status {rc}")
        raise RuntimeError(f"{name}: {kind}: {last_error()}")
    return rc
