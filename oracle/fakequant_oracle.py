"""numpy restatement of the reference's fake-quantization arithmetic (TEST INFRASTRUCTURE, see __init__).

Every function states the reference lines it follows (paths relative to /root/reference).  The reference
issues one ATen op per step; on CPU each op computes in fp32 and rounds its result to the tensor dtype.  That is
restated literally: ``rnd(v, dtype)`` after every step, IEEE single division (numpy float32 ``/``), ``np.rint``
for half-to-even.  dtype is one of "f32", "bf16", "f16"; arrays are carried as float32 holding dtype-exact values.
"""
import math

import numpy as np

F32 = np.float32


# ---- dtype emulation ------------------------------------------------------------------------------------------
def rnd(v, dtype):
    """Round float32 values to ``dtype`` (round-to-nearest-even) and widen back to float32."""
    v = np.asarray(v, dtype=F32)
    if dtype == "f32":
        return v
    if dtype == "f16":
        with np.errstate(over="ignore"):
            return v.astype(np.float16).astype(F32)
    if dtype == "bf16":
        b = v.view(np.uint32).astype(np.uint64)
        nan = np.isnan(v)
        r = ((b + 0x7FFF + ((b >> 16) & 1)) >> 16) << 16
        out = (r & 0xFFFFFFFF).astype(np.uint32).view(F32)
        return np.where(nan, F32(np.nan), out).astype(F32)
    raise ValueError(dtype)


def scalar_to(v, dtype):
    """A 0-dim operand cast to the tensor dtype (add / sub / compare operands)."""
    return F32(rnd(np.asarray(v, dtype=F32), dtype))


# ---- src/brevitas/function/ops.py ---------------------------------------------------------------------------
def min_int(signed, narrow_range, bit_width):
    """function/ops.py:164-191"""
    if signed and narrow_range:
        return -(2.0 ** (bit_width - 1)) + 1
    if signed and not narrow_range:
        return -(2.0 ** (bit_width - 1))
    return 0.0 * bit_width


def max_int(signed, narrow_range, bit_width):
    """function/ops.py:133-160"""
    if not signed and not narrow_range:
        return (2.0 ** bit_width) - 1
    if not signed and narrow_range:
        return (2.0 ** bit_width) - 2
    return (2.0 ** (bit_width - 1)) - 1


def int_threshold(signed, narrow_range, bit_width):
    """core/scaling/int_scaling.py:20-24"""
    return -min_int(signed, narrow_range, bit_width) if signed else max_int(signed, narrow_range, bit_width)


def binary_sign(x):
    """function/ops.py:17-34: (x >= 0) - (x < 0); NaN -> 0"""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid="ignore"):
        return (x >= 0).astype(F32) - (x < 0).astype(F32)


def round_to_zero(x):
    """function/ops.py:38-53: sign(x) * floor(abs(x))"""
    x = np.asarray(x, dtype=F32)
    return (np.sign(x) * np.floor(np.abs(x))).astype(F32)


def dpu_round(x, dtype="f32"):
    """function/ops.py:57-72: where((x < 0) & (x - floor(x) == 0.5), ceil(x), round(x))"""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid="ignore"):
        frac = rnd(x - np.floor(x), dtype)
        return np.where((x < 0) & (frac == F32(0.5)), np.ceil(x), np.rint(x)).astype(F32)


def tensor_clamp(x, min_val, max_val):
    """function/ops.py:76-100: two torch.where; NaN in x passes through, NaN bounds never clamp"""
    x = np.asarray(x, dtype=F32)
    with np.errstate(invalid="ignore"):
        out = np.where(x > max_val, np.broadcast_to(np.asarray(max_val, dtype=F32), x.shape), x)
        out = np.where(out < min_val, np.broadcast_to(np.asarray(min_val, dtype=F32), x.shape), out)
    return out.astype(F32)


def tensor_clamp_inplace(x, min_val, max_val):
    """function/ops.py:104-111: torch.min(x, max) then torch.max(., min): NaN-propagating from either side"""
    x = np.asarray(x, dtype=F32)
    return np.maximum(np.minimum(x, np.asarray(max_val, dtype=F32)), np.asarray(min_val, dtype=F32)).astype(F32)


ROUND_FNS = {
    "round": lambda v, dt: np.rint(v).astype(F32),          # ops/autograd_ste_ops.py RoundSteFn -> torch.round
    "floor": lambda v, dt: np.floor(v).astype(F32),
    "ceil": lambda v, dt: np.ceil(v).astype(F32),
    "round_to_zero": lambda v, dt: round_to_zero(v),
    "dpu_round": lambda v, dt: dpu_round(v, dt),
}


# ---- the STE primitives: forward values (src/brevitas/ops/autograd_ste_ops.py, csrc/autograd_ste_ops.cpp) ----
def ste_forward(name, x, *args, dtype="f32"):
    x = np.asarray(x, dtype=F32)
    if name in ("round_ste", "floor_ste", "ceil_ste", "round_to_zero_ste", "dpu_round_ste"):
        return ROUND_FNS[name[:-4]](x, dtype)
    if name == "binary_sign_ste":
        return binary_sign(x)
    if name == "ternary_sign_ste":      # torch.sign: NaN -> 0 (csrc/autograd_ste_ops.cpp:140-150)
        with np.errstate(invalid="ignore"):
            return (x > 0).astype(F32) - (x < 0).astype(F32)
    if name == "abs_binary_sign_grad":
        return np.abs(x)
    if name == "tensor_clamp_ste":
        return tensor_clamp(x, args[0], args[1])
    if name == "tensor_clamp_ste_":
        return tensor_clamp_inplace(x, args[0], args[1])
    if name == "scalar_clamp_ste":      # torch.clamp(x, lo, hi): min(max(x, lo), hi), NaN-propagating
        lo, hi = scalar_to(args[0], dtype), scalar_to(args[1], dtype)
        return np.where(np.isnan(x), x, np.minimum(np.maximum(x, lo), hi)).astype(F32)
    if name == "scalar_clamp_min_ste":
        lo = scalar_to(args[0], dtype)
        return np.where(np.isnan(x), x, np.maximum(x, lo)).astype(F32)
    raise KeyError(name)


# ---- IntQuant (src/brevitas/core/quant/int_base.py:64-97) -------------------------------------------------------
def int_quant_chain(x, scale, zero_point, qmin, qmax, round_mode="round", dtype="f32"):
    """Returns (y, t1, t3, t5): dequantized output, x/scale, rounded pre-clamp value, clamped integer code."""
    x = np.asarray(x, dtype=F32)
    s = np.asarray(scale, dtype=F32)
    zp = scalar_to(zero_point, dtype)
    lo, hi = scalar_to(qmin, dtype), scalar_to(qmax, dtype)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        t1 = rnd(x / s, dtype)                                   # y = x / scale            int_base.py:70
        t2 = rnd(t1 + zp, dtype)                                 # y = y + zero_point       :71
        t3 = ROUND_FNS[round_mode](t2, dtype)                    # float_to_int_impl        :74
        t5 = tensor_clamp(t3, lo, hi)                            # tensor_clamp_impl        :75
        t6 = rnd(t5 - zp, dtype)                                 # y_int - zero_point       :93
        y = rnd(t6 * s, dtype)                                   # y * scale                :94
    return y, t1, t3, t5


def int_quant_forward(x, scale, zero_point, qmin, qmax, round_mode="round", dtype="f32"):
    return int_quant_chain(x, scale, zero_point, qmin, qmax, round_mode, dtype)[0]


def int_quant_backward(g, x, scale, zero_point, qmin, qmax, round_mode="round", clamp_mode="ste", dtype="f32"):
    """Closed form of the autograd graph behind IntQuant.forward.

    Returns (gx, gscale_elementwise): gx is bit-exact; gscale_elementwise is the per-element contribution
    to d(loss)/d(scale), to be summed over each scale's broadcast region by the caller (order-dependent)."""
    g = np.asarray(g, dtype=F32)
    x = np.asarray(x, dtype=F32)
    s = np.asarray(scale, dtype=F32)
    lo, hi = scalar_to(qmin, dtype), scalar_to(qmax, dtype)
    zp = scalar_to(zero_point, dtype)
    y, t1, t3, t5 = int_quant_chain(x, s, zero_point, qmin, qmax, round_mode, dtype)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        d = rnd(g * s, dtype)                                    # MulBackward: grad * scale
        if clamp_mode == "masked":                               # TensorClamp = torch.where autograd
            m = ~(t3 > hi) & ~(t3 < lo)
            d = np.where(m, d, F32(0.0)).astype(F32)
        gx = rnd(d / s, dtype)                                   # DivBackward wrt x: grad / scale
        t6 = rnd(t5 - zp, dtype)
        gs = g.astype(np.float64) * t6 - d.astype(np.float64) * (t1.astype(np.float64) / s)
    return gx, gs


# ---- QuantReLU with the ReLU folded in (proxy/runtime_quant.py:73-84: activation_impl = nn.ReLU, then tensor_quant) ----
def relu_int_quant_forward(x, scale, zero_point, qmin, qmax, round_mode="round", dtype="f32"):
    """int_quant(torch.relu(x)); torch.relu keeps NaN and maps -0.0 to +0.0"""
    x = np.asarray(x, dtype=F32)
    r = np.where(np.isnan(x), x, np.maximum(x, F32(0.0))).astype(F32)
    r = np.where(r == 0, F32(0.0), r).astype(F32)
    return int_quant_forward(r, scale, zero_point, qmin, qmax, round_mode, dtype)


def relu_int_quant_backward(g, x, scale, zero_point, qmin, qmax, round_mode="round", clamp_mode="masked", dtype="f32"):
    """(gx, gscale_elementwise) of int_quant(relu(x)): the quantizer's backward at relu(x), then ATen's
    threshold_backward, which zeroes the gradient where x <= 0 (NaN inputs keep it)"""
    x = np.asarray(x, dtype=F32)
    r = np.where(np.isnan(x), x, np.maximum(x, F32(0.0))).astype(F32)
    gq, gs = int_quant_backward(g, r, scale, zero_point, qmin, qmax, round_mode, clamp_mode, dtype)
    with np.errstate(invalid="ignore"):
        dead = x <= 0
    return np.where(dead, F32(0.0), gq).astype(F32), gs


# ---- integer export (QuantTensor.int(), quant_tensor/__init__.py:174-187; IntQuant.to_int + cast) -------------------
def int_quant_to_int(x, scale, zero_point, qmin, qmax, round_mode="round", dtype="f32", out_dtype=np.int8):
    codes = int_quant_chain(x, scale, zero_point, qmin, qmax, round_mode, dtype)[3]
    return np.nan_to_num(codes, nan=0.0).astype(np.int64).astype(out_dtype)


# ---- scale from abs-max statistics (core/stats/stats_op.py:129-141, core/scaling/runtime.py:50-72,
#      core/restrict_val.py:22-42, core/quant/int.py:156-163) -----------------------------------------------------
def absmax_rows(x2d):
    x2d = np.asarray(x2d, dtype=F32)
    a = np.abs(x2d)
    m = a.max(axis=1)                     # np.max propagates NaN like torch.max
    return m.astype(F32)


def absmax_tensor(x):
    return F32(np.abs(np.asarray(x, dtype=F32)).max())


def stats_scale(absmax, scaling_min_val, int_thr, dtype="f32", scale_f32=False):
    """threshold = clamp_min(absmax, min_val); scale = threshold / int_threshold."""
    a = np.asarray(absmax, dtype=F32)
    if scaling_min_val is not None and scaling_min_val != 0:
        mv = scalar_to(scaling_min_val, dtype)
        a = np.where(np.isnan(a), a, np.maximum(a, mv)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        s = a / F32(int_thr)               # 0-dim divisor keeps its fp32 value in opmath
    return s.astype(F32) if scale_f32 else rnd(s, dtype)


def rows_absmax_int_quant_forward(x2d, scaling_min_val, int_thr, zero_point, qmin, qmax, round_mode="round",
                                  dtype="f32"):
    """RescalingIntQuant with StatsFromParameterScaling/RuntimeStatsScaling + AbsMax(dim)."""
    am = absmax_rows(x2d)
    s = stats_scale(am, scaling_min_val, int_thr, dtype)
    y = int_quant_forward(x2d, s[:, None], zero_point, qmin, qmax, round_mode, dtype)
    return y, s, am


def rows_absmax_int_quant_backward(g2d, x2d, scale, gscale, int_thr, zero_point, qmin, qmax, round_mode="round",
                                   clamp_mode="ste", dtype="f32"):
    """Backward with the gradient flowing through the abs-max:
    gx[row, first argmax] += sign(x[argmax]) * (Gs[row] + gscale[row]) / int_threshold."""
    x2d = np.asarray(x2d, dtype=F32)
    gx, gs_el = int_quant_backward(g2d, x2d, np.asarray(scale, dtype=F32)[:, None], zero_point, qmin, qmax,
                                   round_mode, clamp_mode, dtype)
    Gs = gs_el.sum(axis=1)
    if gscale is not None:
        Gs = Gs + np.asarray(gscale, dtype=np.float64)
    dthr = rnd(rnd(Gs.astype(F32), dtype) / F32(int_thr), dtype)
    a = np.abs(x2d)
    a = np.where(np.isnan(a), np.inf, a)
    idx = a.argmax(axis=1)                                  # first occurrence, like torch.max(dim) on CPU
    rows = np.arange(x2d.shape[0])
    contrib = dthr * np.sign(x2d[rows, idx])
    gx = gx.copy()
    gx[rows, idx] = rnd(gx[rows, idx] + contrib.astype(F32), dtype)
    return gx, Gs


def tensor_absmax_int_quant_forward(x, scaling_min_val, int_thr, zero_point, qmin, qmax, round_mode="round",
                                    dtype="f32", scale_f32=False):
    am = absmax_tensor(x)
    s = stats_scale(am, scaling_min_val, int_thr, dtype, scale_f32)
    y = int_quant_forward(x, s, zero_point, qmin, qmax, round_mode, dtype)
    return y, F32(s), am


def tensor_absmax_int_quant_backward(g, x, scale, gscale, int_thr, zero_point, qmin, qmax, round_mode="round",
                                     clamp_mode="ste", dtype="f32"):
    """torch.max() (no dim) splits the statistic's gradient evenly over all tied maxima."""
    x = np.asarray(x, dtype=F32)
    gx, gs_el = int_quant_backward(g, x, F32(scale), zero_point, qmin, qmax, round_mode, clamp_mode, dtype)
    Gs = gs_el.sum()
    if gscale is not None:
        Gs = Gs + float(gscale)
    dthr = rnd(rnd(F32(Gs), dtype) / F32(int_thr), dtype)
    a = np.abs(x)
    ties = a == a.max()
    share = rnd(dthr / F32(ties.sum()), dtype)
    gx = np.where(ties, rnd(gx + share * np.sign(x), dtype), gx).astype(F32)
    return gx, Gs


# ---- BinaryQuant / ClampedBinaryQuant (core/quant/binary.py:60-64, 120-125) ---------------------------------------
def binary_quant_forward(x, scale, clamped=False, dtype="f32"):
    x = np.asarray(x, dtype=F32)
    s = np.asarray(scale, dtype=F32)
    c = tensor_clamp(x, -s, s) if clamped else x
    return rnd(binary_sign(c) * s, dtype)


def binary_quant_backward(g, x, scale, clamped=False, dtype="f32"):
    """gx = g*s (masked where the clamp acted); per-element contributions to d(loss)/d(scale)."""
    g = np.asarray(g, dtype=F32)
    x = np.asarray(x, dtype=F32)
    s = np.asarray(scale, dtype=F32)
    d = rnd(g * s, dtype)
    with np.errstate(invalid="ignore"):
        if clamped:
            hi = x > s
            c1 = np.where(hi, s, x)
            lo = c1 < -s
            c = np.where(lo, -s, c1)
            gs = g.astype(np.float64) * binary_sign(c) + np.where(hi, d, 0.0) - np.where(lo, d, 0.0)
            gx = np.where(hi | lo, F32(0.0), d).astype(F32)
        else:
            gs = g.astype(np.float64) * binary_sign(x)
            gx = d
    return gx, gs


# ---- AbsPercentile (core/stats/stats_op.py:41-66) -------------------------------------------------------------------
def percentile_k(q, n):
    return int(math.floor(.01 * q * n + 0.5))


def abs_percentile(x, q, reduce_dim=None):
    """k-th smallest |x| (1-indexed), flat or along ``reduce_dim`` of a 2-D array."""
    a = np.abs(np.asarray(x, dtype=F32))
    if reduce_dim is None:
        k = percentile_k(q, a.size)
        return F32(np.sort(a.reshape(-1), kind="stable")[k - 1])
    assert a.ndim == 2
    k = percentile_k(q, a.shape[reduce_dim])
    return np.sort(a, axis=reduce_dim, kind="stable").take(k - 1, axis=reduce_dim).astype(F32)


# ---- _RuntimeStats EMA (core/stats/stats_wrapper.py:56-65) --------------------------------------------------------
def running_stats_update(running, stat, momentum, first, dtype="f32"):
    running = np.asarray(running, dtype=F32)
    stat = np.asarray(stat, dtype=F32)
    if first:
        return (running * stat).astype(F32)
    r = (running * F32(1 - momentum)).astype(F32)
    return (r + rnd(stat * F32(momentum), dtype)).astype(F32)
