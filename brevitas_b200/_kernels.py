"""Tensor-level wrappers over the C-ABI: pointer/size marshalling, output allocation, no autograd.

Every function requires CUDA tensors and launches on the current stream of the tensor's device.  CPU
tensors are rejected (there is no CPU implementation behind the C-ABI by design).
"""
import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import call

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16, torch.float16: _lib.F16}

# launch counter: bench.py reports how many of OUR kernels' C-ABI calls ran inside the timed region
launch_count = 0


def dtype_tag(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise RuntimeError(f"brevitas_b200: unsupported dtype {t.dtype} (supported: float32, bfloat16, float16)")


def _check_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "brevitas_b200: expected a CUDA tensor, got a tensor on '%s'. The B200 fake-quant path has no "
                "CPU fallback." % t.device)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"brevitas_b200: tensors on different devices ({dev} vs {t.device})")
    return dev


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _dense(t: torch.Tensor) -> torch.Tensor:
    """Row-major OR channels-last dense tensors are used as they are: the element-wise kernels only need x, the
    gradient and the output to share one dense layout, and the per-row kernels only need dim 0 to be the slowest
    dimension in memory (true for both).  Anything else is made contiguous."""
    if t.is_contiguous():
        return t
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t
    if t.dim() == 5 and t.is_contiguous(memory_format=torch.channels_last_3d):
        return t
    return t.contiguous()


def _like(g: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """the gradient laid out exactly like x (a copy only if autograd handed it over in another layout)"""
    if g.shape == x.shape and g.stride() == x.stride():
        return g
    if x.is_contiguous():
        return g.contiguous()
    out = torch.empty_like(x)
    out.copy_(g)
    return out


def _is_channels_last(x: torch.Tensor) -> bool:
    return (not x.is_contiguous()) and ((x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)) or
                                        (x.dim() == 5 and x.is_contiguous(memory_format=torch.channels_last_3d)))


def _is_dim1_channel_scale(x: torch.Tensor, scale: torch.Tensor) -> bool:
    return (scale.dim() == x.dim() and x.dim() >= 3 and scale.shape[1] == x.shape[1]
            and scale.numel() == scale.shape[1])


def _dense_for_scale(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """channels-last x is used in place with one scale, a scale per dim-0 slice ([O,1,1,1]) or a scale per channel
    ([1,C,1,1]: element i of the NHWC memory uses scale[i % C]); other patterns index the logical NCHW order"""
    if scale.numel() == 1 or (scale.dim() == x.dim() and all(d == 1 for d in scale.shape[1:])):
        return _dense(x)
    if _is_channels_last(x) and _is_dim1_channel_scale(x, scale):
        return x
    return _c(x)


def _scale_pattern(x: torch.Tensor, scale: torch.Tensor):
    inner, count, sdt = _scale_args(x, scale)
    if _is_channels_last(x) and scale.numel() > 1 and _is_dim1_channel_scale(x, scale):
        inner, count = 1, scale.numel()           # memory order is N, H, W, C
    return inner, count, sdt


def _launch(dev, name, *args):
    global launch_count
    launch_count += 1
    if torch.cuda.current_device() != dev.index:
        with torch.cuda.device(dev):
            return call(name, *args)
    return call(name, *args)


def broadcast_pattern(x_shape, s_shape) -> Tuple[int, int]:
    """(inner, count) such that element i of x uses scale[(i // inner) % count].

    Supports any scale whose non-singleton dims form one contiguous block of x's dims: scalar, ``[O,1]`` /
    ``[O,1,1,1]`` (per output channel), ``[1,C,1,1]`` (per dim-1 channel), ``[B,T,1]`` (per token), same shape.
    """
    x_shape = tuple(x_shape)
    s_shape = tuple(s_shape)
    numel = 1
    for d in s_shape:
        numel *= d
    if numel == 1:
        return 1, 1
    if len(s_shape) > len(x_shape):
        raise RuntimeError(f"scale of shape {s_shape} is not broadcastable to {x_shape}")
    s_shape = (1,) * (len(x_shape) - len(s_shape)) + s_shape
    nz = [i for i, d in enumerate(s_shape) if d != 1]
    for i in nz:
        if s_shape[i] != x_shape[i]:
            raise RuntimeError(f"scale of shape {s_shape} is not broadcastable to {x_shape}")
    a, b = nz[0], nz[-1] + 1
    for i in range(a, b):
        if s_shape[i] != x_shape[i]:
            raise RuntimeError(
                f"brevitas_b200: unsupported scale broadcast {s_shape} -> {x_shape} (non-contiguous channel dims)")
    count = 1
    for d in x_shape[a:b]:
        count *= d
    inner = 1
    for d in x_shape[b:]:
        inner *= d
    return inner, count


def _scale_args(x, scale):
    inner, count = broadcast_pattern(x.shape, scale.shape)
    sdt = dtype_tag(scale)
    if scale.dtype != x.dtype and not (scale.dtype == torch.float32 and count == 1):
        raise RuntimeError(
            f"brevitas_b200: scale dtype {scale.dtype} with input dtype {x.dtype} is only supported for a "
            "one-element fp32 scale")
    return inner, count, sdt


# ---- the 12 STE primitives ---------------------------------------------------------------------------------

def unary(name: str, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    dev = _check_cuda(x)
    x = _c(x)
    y = torch.empty_like(x) if out is None else out
    _launch(dev, name, x.data_ptr(), y.data_ptr(), x.numel(), dtype_tag(x), _stream(dev))
    return y


def abs_binary_sign_grad_bwd(x, gy):
    dev = _check_cuda(x, gy)
    x, gy = _c(x), _c(gy)
    gx = torch.empty_like(gy)
    _launch(dev, "bvb_abs_binary_sign_grad_bwd", x.data_ptr(), gy.data_ptr(), gx.data_ptr(), x.numel(),
            dtype_tag(x), _stream(dev))
    return gx


def tensor_clamp(x, min_val, max_val, inplace=False):
    dev = _check_cuda(x, min_val, max_val)
    if inplace:
        if not x.is_contiguous():
            raise RuntimeError("tensor_clamp_ste_impl_: in-place clamp needs a contiguous tensor")
        xc = x
    else:
        xc = _c(x)
    mn = _c(min_val).to(x.dtype)       # .type_as(x) of the reference (function/ops.py:98-99)
    mx = _c(max_val).to(x.dtype)
    mn_inner, mn_count = broadcast_pattern(x.shape, mn.shape)
    mx_inner, mx_count = broadcast_pattern(x.shape, mx.shape)
    y = xc if inplace else torch.empty_like(xc)
    _launch(dev, "bvb_tensor_clamp_ste_impl", xc.data_ptr(), mn.data_ptr(), mx.data_ptr(), y.data_ptr(), xc.numel(),
            mn_inner, mn_count, mx_inner, mx_count, 1 if inplace else 0, dtype_tag(x), _stream(dev))
    return y


def tensor_clamp_bwd(gy, x, min_val, max_val, want_min=True, want_max=True):
    """backward of the differentiable tensor_clamp: (gx, gmin fp32[min.numel()] or None, gmax or None)"""
    dev = _check_cuda(gy, x, min_val, max_val)
    xc, gc = _c(x), _c(gy)
    mn = _c(min_val).to(x.dtype)
    mx = _c(max_val).to(x.dtype)
    mn_inner, mn_count = broadcast_pattern(x.shape, mn.shape)
    mx_inner, mx_count = broadcast_pattern(x.shape, mx.shape)
    gx = torch.empty_like(xc)
    gmin = torch.empty(mn_count, dtype=torch.float32, device=dev) if want_min else None
    gmax = torch.empty(mx_count, dtype=torch.float32, device=dev) if want_max else None
    _launch(dev, "bvb_tensor_clamp_bwd", gc.data_ptr(), xc.data_ptr(), mn.data_ptr(), mx.data_ptr(), gx.data_ptr(),
            _ptr(gmin), _ptr(gmax), xc.numel(), mn_inner, mn_count, mx_inner, mx_count, dtype_tag(x), _stream(dev))
    return gx, gmin, gmax


def scalar_clamp(x, min_val: float, max_val: float):
    dev = _check_cuda(x)
    x = _c(x)
    y = torch.empty_like(x)
    _launch(dev, "bvb_scalar_clamp_ste_impl", x.data_ptr(), y.data_ptr(), x.numel(), float(min_val), float(max_val),
            dtype_tag(x), _stream(dev))
    return y


def scalar_clamp_min(x, min_val: float):
    dev = _check_cuda(x)
    x = _c(x)
    y = torch.empty_like(x)
    _launch(dev, "bvb_scalar_clamp_min_ste_impl", x.data_ptr(), y.data_ptr(), x.numel(), float(min_val),
            dtype_tag(x), _stream(dev))
    return y


# ---- IntQuant with a provided scale ------------------------------------------------------------------------

def int_quant_fwd(x, scale, zero_point: float, qmin: float, qmax: float, round_mode: int, want_codes=False,
                  pre_relu=False):
    dev = _check_cuda(x, scale)
    x, scale = _dense_for_scale(x, scale), _c(scale)
    inner, count, sdt = _scale_pattern(x, scale)
    y = torch.empty_like(x)
    codes = torch.empty_like(x) if want_codes else None
    _launch(dev, "bvb_relu_int_quant_fwd" if pre_relu else "bvb_int_quant_fwd", x.data_ptr(), scale.data_ptr(),
            y.data_ptr(), _ptr(codes), x.numel(), inner, count, sdt, zero_point, qmin, qmax, round_mode, dtype_tag(x),
            _stream(dev))
    return (y, codes) if want_codes else y


def int_quant_bwd(gy, x, scale, zero_point, qmin, qmax, round_mode, clamp_mode, want_gscale, pre_relu=False):
    dev = _check_cuda(gy, x, scale)
    x, scale = _dense_for_scale(x, scale), _c(scale)
    gy = _like(gy, x)
    inner, count, sdt = _scale_pattern(x, scale)
    gx = torch.empty_like(x)
    gs = torch.empty(count, dtype=torch.float32, device=dev) if want_gscale else None
    _launch(dev, "bvb_relu_int_quant_bwd" if pre_relu else "bvb_int_quant_bwd", gy.data_ptr(), x.data_ptr(),
            scale.data_ptr(), gx.data_ptr(), _ptr(gs), x.numel(), inner, count, sdt, zero_point, qmin, qmax, round_mode,
            clamp_mode, dtype_tag(x), _stream(dev))
    return gx, gs


def _zpt_args(x, scale, zero_point):
    if zero_point.numel() != scale.numel():
        raise RuntimeError("brevitas_b200: a tensor zero-point must have the broadcast pattern of the scale")
    inner, count, sdt = _scale_args(x, scale)
    zdt = dtype_tag(zero_point)
    if zero_point.dtype != x.dtype and not (zero_point.dtype == torch.float32 and count == 1):
        raise RuntimeError("brevitas_b200: zero-point dtype must equal the input dtype (or be fp32 with one element)")
    return inner, count, sdt, zdt


def int_quant_zpt_fwd(x, scale, zero_point, qmin: float, qmax: float, round_mode: int):
    """IntQuant with a tensor-valued zero-point (same broadcast pattern as the scale)"""
    dev = _check_cuda(x, scale, zero_point)
    x, scale, zero_point = _c(x), _c(scale), _c(zero_point)
    inner, count, sdt, zdt = _zpt_args(x, scale, zero_point)
    y = torch.empty_like(x)
    _launch(dev, "bvb_int_quant_zpt_fwd", x.data_ptr(), scale.data_ptr(), zero_point.data_ptr(), y.data_ptr(), x.numel(),
            inner, count, sdt, zdt, qmin, qmax, round_mode, dtype_tag(x), _stream(dev))
    return y


def int_quant_zpt_bwd(gy, x, scale, zero_point, qmin, qmax, round_mode, clamp_mode, want_grads):
    dev = _check_cuda(gy, x, scale, zero_point)
    gy, x, scale, zero_point = _c(gy), _c(x), _c(scale), _c(zero_point)
    inner, count, sdt, zdt = _zpt_args(x, scale, zero_point)
    gx = torch.empty_like(x)
    gs = torch.empty(count, dtype=torch.float32, device=dev) if want_grads else None
    gz = torch.empty(count, dtype=torch.float32, device=dev) if want_grads else None
    _launch(dev, "bvb_int_quant_zpt_bwd", gy.data_ptr(), x.data_ptr(), scale.data_ptr(), zero_point.data_ptr(),
            gx.data_ptr(), _ptr(gs), _ptr(gz), x.numel(), inner, count, sdt, zdt, qmin, qmax, round_mode, clamp_mode,
            dtype_tag(x), _stream(dev))
    return gx, gs, gz


_INT_OUT = {torch.int8: (_lib.OUT_I8, -128.0, 127.0), torch.uint8: (_lib.OUT_U8, 0.0, 255.0),
            torch.int32: (_lib.OUT_I32, -2147483648.0, 2147483520.0)}


def int_quant_to_int(x, scale, zero_point: float, qmin: Optional[float], qmax: Optional[float], round_mode: int,
                     out_dtype=torch.int8):
    """integer codes clamp(round(x / scale + zero_point), qmin, qmax) in a real integer dtype (qmin / qmax None: the
    output dtype's own limits, i.e. the conversion of an already quantized tensor)"""
    dev = _check_cuda(x, scale)
    try:
        kind, lo, hi = _INT_OUT[out_dtype]
    except KeyError:
        raise RuntimeError(f"brevitas_b200: integer export supports int8, uint8 and int32, got {out_dtype}")
    x, scale = _dense_for_scale(x, scale), _c(scale)
    inner, count, sdt = _scale_pattern(x, scale)
    out = torch.empty_like(x, dtype=out_dtype)
    _launch(dev, "bvb_int_quant_to_int", x.data_ptr(), scale.data_ptr(), out.data_ptr(), x.numel(), inner, count, sdt,
            zero_point, lo if qmin is None else qmin, hi if qmax is None else qmax, round_mode, kind, dtype_tag(x),
            _stream(dev))
    return out


# ---- fused abs-max + quant --------------------------------------------------------------------------------

def rows_absmax_int_quant_fwd(x, rows, cols, scaling_min_val, int_threshold, zero_point, qmin, qmax, round_mode,
                              want_absmax=False):
    dev = _check_cuda(x)
    x = _dense(x) if (x.dim() >= 1 and x.shape[0] == rows) else _c(x)     # rows must be dim-0 slices to keep a layout
    assert rows * cols == x.numel()
    y = torch.empty_like(x)
    scale = torch.empty(rows, dtype=x.dtype, device=dev)
    absmax = torch.empty(rows, dtype=x.dtype, device=dev) if want_absmax else None
    _launch(dev, "bvb_rows_absmax_int_quant_fwd", x.data_ptr(), y.data_ptr(), scale.data_ptr(), _ptr(absmax), rows,
            cols, scaling_min_val, int_threshold, zero_point, qmin, qmax, round_mode, dtype_tag(x), _stream(dev))
    return y, scale, absmax


def rows_absmax_int_quant_bwd(gy, x, scale, gscale, rows, cols, int_threshold, zero_point, qmin, qmax, round_mode,
                              clamp_mode):
    dev = _check_cuda(gy, x, scale, gscale)
    x = _dense(x) if (x.dim() >= 1 and x.shape[0] == rows) else _c(x)
    gy, scale = _like(gy, x), _c(scale)
    if gscale is not None:
        gscale = _c(gscale)
    gx = torch.empty_like(x)
    _launch(dev, "bvb_rows_absmax_int_quant_bwd", gy.data_ptr(), x.data_ptr(), scale.data_ptr(), _ptr(gscale),
            gx.data_ptr(), rows, cols, int_threshold, zero_point, qmin, qmax, round_mode, clamp_mode, dtype_tag(x),
            _stream(dev))
    return gx


def _workspace(dev):
    return torch.empty(_lib.load().bvb_workspace_bytes(), dtype=torch.uint8, device=dev)


def tensor_absmax_int_quant_fwd(x, scale_dtype, scaling_min_val, int_threshold, zero_point, qmin, qmax, round_mode):
    dev = _check_cuda(x)
    x = _dense(x)
    y = torch.empty_like(x)
    scale = torch.empty((), dtype=scale_dtype, device=dev)
    absmax = torch.empty((), dtype=x.dtype, device=dev)
    ws = _workspace(dev)
    _launch(dev, "bvb_tensor_absmax_int_quant_fwd", x.data_ptr(), y.data_ptr(), scale.data_ptr(), absmax.data_ptr(),
            x.numel(), _DTYPES[scale_dtype], scaling_min_val, int_threshold, zero_point, qmin, qmax, round_mode,
            dtype_tag(x), ws.data_ptr(), _stream(dev))
    return y, scale, absmax


def tensor_absmax_int_quant_bwd(gy, x, scale, absmax, gscale, int_threshold, zero_point, qmin, qmax, round_mode,
                                clamp_mode):
    dev = _check_cuda(gy, x, scale, absmax, gscale)
    x = _dense(x)
    gy = _like(gy, x)
    gx = torch.empty_like(x)
    ws = _workspace(dev)
    _launch(dev, "bvb_tensor_absmax_int_quant_bwd", gy.data_ptr(), x.data_ptr(), scale.data_ptr(), absmax.data_ptr(),
            _ptr(gscale), gx.data_ptr(), x.numel(), dtype_tag(scale), int_threshold, zero_point, qmin, qmax,
            round_mode, clamp_mode, dtype_tag(x), ws.data_ptr(), _stream(dev))
    return gx


# ---- binary ------------------------------------------------------------------------------------------------

def binary_quant_fwd(x, scale, clamped: bool):
    dev = _check_cuda(x, scale)
    x, scale = _c(x), _c(scale)
    inner, count, sdt = _scale_args(x, scale)
    y = torch.empty_like(x)
    _launch(dev, "bvb_binary_quant_fwd", x.data_ptr(), scale.data_ptr(), y.data_ptr(), x.numel(), inner, count, sdt,
            1 if clamped else 0, dtype_tag(x), _stream(dev))
    return y


def binary_quant_bwd(gy, x, scale, clamped: bool, want_gscale: bool):
    dev = _check_cuda(gy, x, scale)
    gy, x, scale = _c(gy), _c(x), _c(scale)
    inner, count, sdt = _scale_args(x, scale)
    gx = torch.empty_like(x)
    gs = torch.empty(count, dtype=torch.float32, device=dev) if want_gscale else None
    _launch(dev, "bvb_binary_quant_bwd", gy.data_ptr(), x.data_ptr(), scale.data_ptr(), gx.data_ptr(), _ptr(gs),
            x.numel(), inner, count, sdt, 1 if clamped else 0, dtype_tag(x), _stream(dev))
    return gx, gs


# ---- general integer quantizer (decoupled / device-resident range) and ternary ---------------------------------------

def _general_args(x, pre_scale, scale):
    pi, pc = broadcast_pattern(x.shape, pre_scale.shape)
    si, sc = broadcast_pattern(x.shape, scale.shape)
    if pre_scale.dtype != x.dtype or scale.dtype != x.dtype:
        raise RuntimeError("brevitas_b200: general_int_quant needs both scales in the dtype of the input")
    if pc != 1 and sc != 1 and (pc != sc or pi != si):
        raise RuntimeError("brevitas_b200: general_int_quant needs one broadcast pattern for both scales")
    inner = pi if pc != 1 else si
    return inner, pc, sc


def _f32_scalar(t):
    if t.numel() != 1:
        raise RuntimeError("brevitas_b200: general_int_quant takes one-element zero-points and integer bounds")
    return t.detach().to(torch.float32).reshape(1)


def general_int_quant_supported(x, pre_scale, scale, *scalars) -> bool:
    if not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        return False
    if pre_scale.dtype != x.dtype or scale.dtype != x.dtype or any(t.numel() != 1 for t in scalars):
        return False
    try:
        _general_args(x, pre_scale, scale)
    except RuntimeError:
        return False
    return True


def general_int_quant_fwd(x, pre_scale, scale, pre_zp, zp, lo, hi, round_mode):
    dev = _check_cuda(x, pre_scale, scale, pre_zp, zp, lo, hi)
    x = _dense(x) if pre_scale.numel() == 1 and scale.numel() == 1 else _c(x)
    pre_scale, scale = _c(pre_scale), _c(scale)
    inner, pc, sc = _general_args(x, pre_scale, scale)
    pre_zp, zp, lo, hi = (_f32_scalar(t) for t in (pre_zp, zp, lo, hi))
    y = torch.empty_like(x)
    _launch(dev, "bvb_general_int_quant_fwd", x.data_ptr(), pre_scale.data_ptr(), scale.data_ptr(), pre_zp.data_ptr(),
            zp.data_ptr(), lo.data_ptr(), hi.data_ptr(), y.data_ptr(), x.numel(), inner, pc, sc, round_mode, dtype_tag(x),
            _stream(dev))
    return y


def general_int_quant_bwd(gy, x, pre_scale, scale, pre_zp, zp, lo, hi, round_mode, clamp_mode, same_scale, want_sums):
    """returns (gx, sums fp64 [d pre_scale | d scale | d min_int | d max_int] or None)"""
    dev = _check_cuda(gy, x, pre_scale, scale, pre_zp, zp, lo, hi)
    x = _dense(x) if pre_scale.numel() == 1 and scale.numel() == 1 else _c(x)
    pre_scale, scale = _c(pre_scale), _c(scale)
    gy = _like(gy, x)
    inner, pc, sc = _general_args(x, pre_scale, scale)
    pre_zp, zp, lo, hi = (_f32_scalar(t) for t in (pre_zp, zp, lo, hi))
    gx = torch.empty_like(x)
    sums = torch.empty(pc + sc + 2, dtype=torch.float64, device=dev) if want_sums else None
    _launch(dev, "bvb_general_int_quant_bwd", gy.data_ptr(), x.data_ptr(), pre_scale.data_ptr(), scale.data_ptr(),
            pre_zp.data_ptr(), zp.data_ptr(), lo.data_ptr(), hi.data_ptr(), gx.data_ptr(), _ptr(sums), x.numel(), inner, pc, sc,
            round_mode, clamp_mode, 1 if same_scale else 0, dtype_tag(x), _stream(dev))
    return gx, sums


def ternary_quant_fwd(x, scale, threshold: float):
    dev = _check_cuda(x, scale)
    x, scale = _dense(x), _c(scale)
    y = torch.empty_like(x)
    _launch(dev, "bvb_ternary_quant_fwd", x.data_ptr(), scale.data_ptr(), y.data_ptr(), x.numel(), threshold, dtype_tag(x),
            _stream(dev))
    return y


def ternary_quant_bwd(gy, x, scale, threshold: float, want_gscale: bool):
    dev = _check_cuda(gy, x, scale)
    x, scale = _dense(x), _c(scale)
    gy = _like(gy, x)
    gx = torch.empty_like(x)
    gs = torch.empty(1, dtype=torch.float64, device=dev) if want_gscale else None
    _launch(dev, "bvb_ternary_quant_bwd", gy.data_ptr(), x.data_ptr(), scale.data_ptr(), gx.data_ptr(), _ptr(gs), x.numel(),
            threshold, dtype_tag(x), _stream(dev))
    return gx, gs


# ---- statistics --------------------------------------------------------------------------------------------

def absmax_rows(x, rows, cols):
    dev = _check_cuda(x)
    x = _c(x)
    out = torch.empty(rows, dtype=x.dtype, device=dev)
    _launch(dev, "bvb_absmax_rows", x.data_ptr(), out.data_ptr(), rows, cols, dtype_tag(x), _stream(dev))
    return out


def absmax_tensor(x):
    dev = _check_cuda(x)
    x = _c(x)
    out = torch.empty((), dtype=x.dtype, device=dev)
    ws = _workspace(dev)
    _launch(dev, "bvb_absmax_tensor", x.data_ptr(), out.data_ptr(), x.numel(), dtype_tag(x), ws.data_ptr(),
            _stream(dev))
    return out


def abs_kth_value_rows(x, rows, cols, k, want_index=False, signed=False, pre_relu=False, dense_ok=False):
    """k-th smallest of |x| (signed=False), of x itself (signed=True) or of relu(x) (pre_relu) per row; optionally the
    smallest index attaining it.  dense_ok: x is a dense channels-last tensor read in MEMORY order (whole-tensor
    statistics are permutation invariant); the index is then a storage offset."""
    dev = _check_cuda(x)
    x = _dense(x) if dense_ok else _c(x)
    out = torch.empty(rows, dtype=x.dtype, device=dev)
    idx = torch.empty(rows, dtype=torch.int64, device=dev) if want_index else None
    ws = torch.empty(_lib.load().bvb_kth_workspace_bytes(rows), dtype=torch.uint8, device=dev)
    name = "bvb_kth_value_rows" if signed else ("bvb_relu_abs_kth_value_rows" if pre_relu else "bvb_abs_kth_value_rows")
    _launch(dev, name, x.data_ptr(), out.data_ptr(), _ptr(idx),
            rows, cols, k, dtype_tag(x), ws.data_ptr(), _stream(dev))
    return out, idx


def kth_value_rows(x, rows, cols, k, want_index=False):
    return abs_kth_value_rows(x, rows, cols, k, want_index, signed=True)


_mm_ws = {}


def minmax_rows(x, rows, cols):
    """(min, max, argmin, argmax) of every row of x[rows][cols] in one read (csrc/stats.cu)"""
    dev = _check_cuda(x)
    x = _c(x)
    need = int(_lib.load().bvb_minmax_workspace_bytes(rows))
    ws = _mm_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 16), dtype=torch.uint8, device=dev)
        _mm_ws[dev] = ws
    mn = torch.empty(rows, dtype=x.dtype, device=dev)
    mx = torch.empty(rows, dtype=x.dtype, device=dev)
    imn = torch.empty(rows, dtype=torch.int64, device=dev)
    imx = torch.empty(rows, dtype=torch.int64, device=dev)
    _launch(dev, "bvb_minmax_rows", x.data_ptr(), mn.data_ptr(), mx.data_ptr(), imn.data_ptr(), imx.data_ptr(), rows, cols,
            dtype_tag(x), ws.data_ptr(), _stream(dev))
    return mn, mx, imn, imx


def percentile_k(q: float, n: int) -> int:
    """k of AbsPercentile (src/brevitas/core/stats/stats_op.py:55, 61): floor(.01 * q * n + 0.5), 1-indexed."""
    return int(math.floor(.01 * q * n + 0.5))


def running_stats_update(running, stat, momentum: float, first: bool):
    dev = _check_cuda(running, stat)
    if running.dtype != torch.float32 or not running.is_contiguous():
        raise RuntimeError("running_stats_update: running buffer must be contiguous float32")
    stat = _c(stat)
    _launch(dev, "bvb_running_stats_update", running.data_ptr(), stat.data_ptr(), running.numel(), float(momentum),
            float(1 - momentum), 1 if first else 0, dtype_tag(stat), _stream(dev))
    return running


# ---- batch-norm + ReLU + activation quantizer, fused (csrc/bn_act_quant.cu) ------------------------------------------
_bn_ws = {}


def _bn_workspace(dev, channels):
    need = int(_lib.load().bvb_bn_act_quant_workspace_bytes(channels))
    ws = _bn_ws.get(dev)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 22), dtype=torch.uint8, device=dev)
        _bn_ws[dev] = ws
    return ws


def bn_act_quant_supported(x: torch.Tensor) -> bool:
    """a 4-D channels-last (or 2-D) dense CUDA tensor whose channel count fills 16-byte vectors that divide 256"""
    if not x.is_cuda or x.dtype not in _DTYPES:
        return False
    if x.dim() == 4:
        if not x.is_contiguous(memory_format=torch.channels_last):
            return False
    elif x.dim() != 2 or not x.is_contiguous():
        return False
    c, v = x.shape[1], 16 // x.element_size()
    return c % v == 0 and c // v <= 256 and 256 % (c // v) == 0 and x.numel() > 0


def bn_act_quant_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, training, scale, zero_point, qmin, qmax,
                     relu=True, residual=None):
    """returns (y like x, save_mean, save_invstd); running statistics are updated in place when training"""
    dev = _check_cuda(x, scale)
    c = x.shape[1]
    rows = x.numel() // c
    y = torch.empty_like(x)                                  # preserves channels-last strides
    if training:
        save_mean = torch.empty(c, dtype=torch.float32, device=dev)
        save_invstd = torch.empty(c, dtype=torch.float32, device=dev)
    else:
        save_mean = running_mean.float().contiguous()
        save_invstd = torch.rsqrt(running_var.float() + eps).contiguous()
    ws = _bn_workspace(dev, c)
    scale = _c(scale)
    if residual is not None and (residual.shape != x.shape or residual.stride() != x.stride() or residual.dtype != x.dtype):
        raise RuntimeError("bn_act_quant: the residual must have the shape, strides and dtype of x")
    _launch(dev, "bvb_bn_act_quant_fwd", x.data_ptr(), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(running_mean) if training else None,
            _ptr(running_var) if training else None, float(momentum), float(eps), 0 if training else 1, scale.data_ptr(),
            scale.numel(), dtype_tag(scale), y.data_ptr(), save_mean.data_ptr(), save_invstd.data_ptr(), rows, c,
            float(zero_point), float(qmin), float(qmax), _lib.ROUND, 1 if relu else 0, dtype_tag(x), ws.data_ptr(),
            _stream(dev))
    return y, save_mean, save_invstd


def bn_act_quant_bwd(gy, x, gamma, beta, save_mean, save_invstd, scale, zero_point, qmin, qmax, clamp_mode, relu=True,
                     want_gscale=True, residual=None, want_gresidual=False):
    dev = _check_cuda(gy, x, scale)
    c = x.shape[1]
    rows = x.numel() // c
    if gy.stride() != x.stride():
        gy = gy.contiguous(memory_format=torch.channels_last) if x.dim() == 4 else gy.contiguous()
    gx = torch.empty_like(x)
    ggamma = torch.empty(c, dtype=torch.float32, device=dev)
    gbeta = torch.empty(c, dtype=torch.float32, device=dev)
    scale = _c(scale)
    gscale = torch.empty(scale.numel(), dtype=torch.float32, device=dev) if want_gscale else None
    ws = _bn_workspace(dev, c)
    gres = torch.empty_like(x) if (residual is not None and want_gresidual) else None
    _launch(dev, "bvb_bn_act_quant_bwd", gy.data_ptr(), x.data_ptr(), _ptr(residual), _ptr(gres), _ptr(gamma), _ptr(beta), save_mean.data_ptr(),
            save_invstd.data_ptr(), scale.data_ptr(), scale.numel(), dtype_tag(scale), gx.data_ptr(), ggamma.data_ptr(),
            gbeta.data_ptr(), _ptr(gscale), rows, c, float(zero_point), float(qmin), float(qmax), _lib.ROUND, int(clamp_mode),
            1 if relu else 0, dtype_tag(x), ws.data_ptr(), _stream(dev))
    return gx, ggamma, gbeta, gscale, gres
