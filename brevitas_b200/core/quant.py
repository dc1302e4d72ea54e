"""Mirror of ``brevitas.core.quant`` -- the ``tensor_quant`` modules -- on fused sm_100a kernels.

Reference: src/brevitas/core/quant/int_base.py:15-97 (IntQuant), int.py:94-163 (RescalingIntQuant),
binary.py:19-64 (BinaryQuant), :67-125 (ClampedBinaryQuant), delay.py:43-54 (DelayWrapper).

Same constructor arguments, same sub-module names (``int_quant``, ``scaling_impl``, ``int_scaling_impl``,
``zero_point_impl``, ``msb_clamp_bit_width_impl`` ...), same ``forward`` contract
``x -> (y, scale, zero_point, bit_width)``, hence state-dict compatible with the reference.  What changes is how
``forward`` executes: where the reference issues ~9-20 ATen kernels, these modules launch ONE kernel
(statistic + scale + quant-dequant) or one kernel after a few tiny scale ops, selected from the types of the
injected sub-modules (rounding mode, clamp-gradient mode, statistic, view).  Configurations the fused kernels do
not cover (exotic injected modules, mixed dtypes) run the literal reference sequence on the STE kernels (op-level
drop-in); a learned bit-width and the decoupled quantizers take ``general_int_quant``, whose range inputs
stay on the device.
"""
import weakref
from functools import lru_cache
from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from .. import _kernels as _K
from .. import _lib
from .. import ops as _ops  # noqa: F401
from ..function.ops import max_int, min_int
from ..function.ops_ste import binary_sign_ste, round_ste, ternary_sign_ste
from .bit_width import BitWidthConst
from .function_wrapper import CLAMP_MODE_OF, ROUND_MODE_OF, RoundSte, TensorClamp
from .scaling import IntScaling, PowerOfTwoIntScaling
from .stats import shared_minmax
from .utils import StatelessBuffer
from .zero_point import ZeroZeroPoint


# ---- delay (core/quant/delay.py) ----------------------------------------------------------------------------
class _NoDelay(nn.Module):
    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        return y


class _DelayQuant(nn.Module):
    def __init__(self, quant_delay_steps):
        super().__init__()
        self.quant_delay_steps: int = quant_delay_steps

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        if self.quant_delay_steps > 0:
            self.quant_delay_steps = self.quant_delay_steps - 1
            return x
        return y


class DelayWrapper(nn.Module):
    """Returns the un-quantized input for the first ``quant_delay_steps`` calls (delay.py:43-54)."""

    def __init__(self, quant_delay_steps: Optional[int]):
        super().__init__()
        if quant_delay_steps is None or quant_delay_steps <= 0:
            self.delay_impl = _NoDelay()
        else:
            self.delay_impl = _DelayQuant(quant_delay_steps)

    def forward(self, x: Tensor, y: Tensor) -> Tensor:
        return self.delay_impl(x, y)


# ---- mixed-dtype scalar operands in the literal sequences -----------------------------------------------------------
# A 0-dim fp32 scale next to a bf16 / fp16 tensor (fp32 quantizer buffers, low-precision weights or activations): the
# reference result (ATen on CPU, which the golden vectors record) keeps the scale in fp32 "opmath" for mul / div and
# rounds once to the tensor dtype; ATen on CUDA instead casts a 0-dim CUDA tensor to the common dtype first, i.e.
# rounds the scale to 16 bits.  The fused kernels implement the former (``scale_dtype`` argument of the C-ABI); the
# literal sequences below do the same explicitly so that both paths agree with the reference bit for bit.
def _is_fp32_scalar_with_lowp(x: Tensor, s: Tensor) -> bool:
    return s.numel() == 1 and s.dtype == torch.float32 and x.dtype in (torch.bfloat16, torch.float16)


def scalar_div(x: Tensor, s: Tensor) -> Tensor:
    if _is_fp32_scalar_with_lowp(x, s):
        return (x.float() / s).to(x.dtype)
    return x / s


def scalar_mul(x: Tensor, s: Tensor) -> Tensor:
    if _is_fp32_scalar_with_lowp(x, s):
        return (x.float() * s).to(x.dtype)
    return x * s


# ---- host copies of the 0-dim range tensors ------------------------------------------------------------------
@lru_cache(maxsize=None)
def int_range(signed: bool, narrow_range: bool, bit_width: int, dtype: torch.dtype) -> Tuple[float, float]:
    """(min_int, max_int) as the reference computes them from a 0-dim ``bit_width`` tensor of ``dtype``
    (function/ops.py:133-191), evaluated once on the host so the fused kernels need no device read-back."""
    bw = torch.tensor(float(bit_width), dtype=dtype, device='cpu')
    return float(min_int(signed, narrow_range, bw)), float(max_int(signed, narrow_range, bw))


@lru_cache(maxsize=None)
def _int_threshold(kind: str, signed: bool, narrow_range: bool, bit_width: int, dtype: torch.dtype) -> float:
    bw = torch.tensor(float(bit_width), dtype=dtype, device='cpu')
    if kind == 'int':
        impl = IntScaling(signed, narrow_range)
    else:
        impl = PowerOfTwoIntScaling(signed)
    return float(impl(bw))


_INT_THRESHOLD_CACHE = weakref.WeakKeyDictionary()     # RescalingIntQuant module -> {(device, dtype, bits): 0-dim tensor}
_SCALAR_CACHE = weakref.WeakKeyDictionary()     # IntQuant module -> [(zero-point ref, version, bit-width ref, version, values)]


class IntQuant(nn.Module):
    """Scaled, shifted, uniform integer quantization, output in dequantized format (int_base.py:15-97).

    ``y = (clamp(round(x / scale + zero_point), min_int, max_int) - zero_point) * scale``, one kernel.
    """

    def __init__(self, narrow_range: bool, signed: bool, float_to_int_impl: Optional[nn.Module] = None,
                 tensor_clamp_impl: Optional[nn.Module] = None, quant_delay_steps: int = 0):
        super().__init__()
        self.float_to_int_impl = float_to_int_impl if float_to_int_impl is not None else RoundSte()
        self.tensor_clamp_impl = tensor_clamp_impl if tensor_clamp_impl is not None else TensorClamp()
        self.signed = signed
        self.narrow_range = narrow_range
        self.delay_wrapper = DelayWrapper(quant_delay_steps)

    # -- kernel mode selection
    def kernel_modes(self) -> Optional[Tuple[int, int]]:
        rm = ROUND_MODE_OF.get(type(self.float_to_int_impl))
        cm = CLAMP_MODE_OF.get(type(self.tensor_clamp_impl))
        if rm is None or cm is None:
            return None
        return rm, cm

    def min_int(self, bit_width):
        return min_int(self.signed, self.narrow_range, bit_width)

    def max_int(self, bit_width):
        return max_int(self.signed, self.narrow_range, bit_width)

    def to_int(self, scale: Tensor, zero_point: Tensor, bit_width: Tensor, x: Tensor) -> Tensor:
        """Integer codes as floats (int_base.py:64-76), literal op sequence on the STE kernels."""
        y = scalar_div(x, scale)
        y = y + zero_point
        min_int_val = self.min_int(bit_width)
        max_int_val = self.max_int(bit_width)
        y = self.float_to_int_impl(y)
        y = self.tensor_clamp_impl(y, min_val=min_int_val, max_val=max_int_val)
        return y

    def forward_fused(self, scale: Tensor, zero_point: float, qmin: float, qmax: float, x: Tensor) -> Optional[Tensor]:
        """One-kernel path given host copies of zero-point and integer range; None if not applicable."""
        modes = self.kernel_modes()
        if modes is None:
            return None
        if scale.dtype != x.dtype and not (scale.numel() == 1 and scale.dtype == torch.float32):
            return None
        y = torch.ops.brevitas_b200.int_quant(x, scale, zero_point, qmin, qmax, modes[0], modes[1])
        return self.delay_wrapper(x, y)

    def forward_fused_zpt(self, scale: Tensor, zero_point: Tensor, qmin: float, qmax: float, x: Tensor) -> Optional[Tensor]:
        """One kernel with a TENSOR zero-point that has the scale's shape (asymmetric quantizers); gradients flow to x,
        scale and zero-point.  None if not applicable."""
        modes = self.kernel_modes()
        if modes is None or zero_point.shape != scale.shape:
            return None
        for t in (scale, zero_point):
            if t.dtype != x.dtype and not (t.numel() == 1 and t.dtype == torch.float32):
                return None
        y = torch.ops.brevitas_b200.int_quant_zpt(x, scale, zero_point, qmin, qmax, modes[0], modes[1])
        return self.delay_wrapper(x, y)

    def _host_scalars(self, zero_point: Tensor, bit_width: Tensor):
        """(zero_point, qmin, qmax) as host floats for a DIRECT call with 0-dim tensor arguments.  The reference never
        synchronises here; neither does this after the first call with the same (unmodified) tensors: the values are
        cached per (tensor object, version) of the two tensors, so a training loop that passes its quantizer's constant buffers
        reads them back once, not every step."""
        cache = _SCALAR_CACHE.setdefault(self, [])          # kept outside the module: it must stay picklable
        for zr, zv, br, bv, vals in cache:
            # the very same tensor OBJECTS, unmodified since (weak references: a freed tensor never matches, and a new tensor
            # that happens to reuse its memory is another object)
            if zr() is zero_point and br() is bit_width and zv == zero_point._version and bv == bit_width._version:
                return vals
        vals = (float(zero_point), float(self.min_int(bit_width)), float(self.max_int(bit_width)))
        if len(cache) >= 4:
            cache.pop(0)
        cache.append((weakref.ref(zero_point), zero_point._version, weakref.ref(bit_width), bit_width._version, vals))
        return vals

    def forward(self, scale: Tensor, zero_point: Tensor, bit_width: Tensor, x: Tensor) -> Tensor:
        if (x.is_cuda and zero_point.numel() == 1 and bit_width.numel() == 1 and not zero_point.requires_grad
                and not bit_width.requires_grad and not torch.cuda.is_current_stream_capturing()):
            # direct call with tensor arguments: the two 0-dim tensors are read back once (cached), then the fused kernel
            zp, qmin, qmax = self._host_scalars(zero_point, bit_width)
            y = self.forward_fused(scale, zp, qmin, qmax, x)
            if y is not None:
                return y
        y = self.forward_device_range(scale, zero_point, bit_width, x)
        if y is not None:
            return y
        y_int = self.to_int(scale, zero_point, bit_width, x)
        y = y_int - zero_point
        y = scalar_mul(y, scale)
        return self.delay_wrapper(x, y)

    def forward_device_range(self, scale: Tensor, zero_point: Tensor, bit_width: Tensor, x: Tensor) -> Optional[Tensor]:
        """One kernel whose zero-point and integer range stay ON THE DEVICE (``general_int_quant``): the learned bit-width
        path (core/bit_width/parameter.py:23-98 -- the masked clamp's backward returns d(min_int) and d(max_int), which is
        what trains the bit-width), and direct calls that must not read anything back (CUDA-graph capture).  None if not
        applicable."""
        modes = self.kernel_modes()
        if modes is None or zero_point.requires_grad:
            return None
        if not _K.general_int_quant_supported(x, scale, scale, zero_point, bit_width):
            return None
        lo, hi = self.min_int(bit_width), self.max_int(bit_width)          # 0-dim ATen ops, differentiable in bit_width
        y = torch.ops.brevitas_b200.general_int_quant(x, scale, scale, zero_point, zero_point, lo, hi, modes[0], modes[1],
                                                      True)
        return self.delay_wrapper(x, y)


class RescalingIntQuant(nn.Module):
    """Gets scale, zero-point and bit-width from their implementations and quantizes (int.py:94-163).

    ``scale = scaling_impl(x) / int_scaling_impl(bit_width)``; returns ``(y, scale, zero_point, bit_width)``.
    """

    def __init__(self, int_quant: nn.Module, scaling_impl: nn.Module, int_scaling_impl: nn.Module,
                 zero_point_impl: nn.Module, bit_width_impl: nn.Module):
        super().__init__()
        self.int_quant = int_quant
        self.scaling_impl = scaling_impl
        self.int_scaling_impl = int_scaling_impl
        self.zero_point_impl = zero_point_impl
        self.msb_clamp_bit_width_impl = bit_width_impl

    def _host_config(self, bw_dtype: torch.dtype):
        """(zero_point, qmin, qmax, round_mode, clamp_mode, int_threshold or None) when every range input is a
        construction-time constant; None otherwise (-> literal reference sequence)."""
        iq = self.int_quant
        if type(iq) is not IntQuant or type(self.msb_clamp_bit_width_impl) is not BitWidthConst:
            return None
        modes = iq.kernel_modes()
        if modes is None:
            return None
        bw = self.msb_clamp_bit_width_impl.bit_width_value
        qmin, qmax = int_range(iq.signed, iq.narrow_range, bw, bw_dtype)
        isi = self.int_scaling_impl
        if type(isi) is IntScaling:
            thr = _int_threshold('int', isi.signed, isi.narrow_range, bw, bw_dtype)
        elif type(isi) is PowerOfTwoIntScaling:
            thr = _int_threshold('po2', isi.signed, False, bw, bw_dtype)
        else:
            thr = None
        # zero-point: the constant 0.0, or None = tensor-valued (computed by zero_point_impl, fused as a device operand)
        zp = 0.0 if type(self.zero_point_impl) is ZeroZeroPoint else None
        return zp, qmin, qmax, modes[0], modes[1], thr

    def int_threshold(self, bit_width: Tensor) -> Tensor:
        """``int_scaling_impl(bit_width)``.  With a CONSTANT bit-width it is a constant too, yet the reference recomputes it
        (a pow and one or two subtractions on a 0-dim tensor: three launches) in every forward of every quantizer -- 55
        quantizers of a MobileNetV1 make 165 one-element launches per step.  Here it is computed once per device and dtype,
        by the very same ops, and reused (not a buffer: nothing is added to the state dict)."""
        if type(self.msb_clamp_bit_width_impl) is not BitWidthConst or type(self.int_scaling_impl) not in (
                IntScaling, PowerOfTwoIntScaling) or bit_width.requires_grad:
            return self.int_scaling_impl(bit_width)
        key = (bit_width.device, bit_width.dtype, self.msb_clamp_bit_width_impl.bit_width_value)
        cache = _INT_THRESHOLD_CACHE.setdefault(self, {})
        thr = cache.get(key)
        if thr is None:
            if bit_width.is_cuda and torch.cuda.is_current_stream_capturing():
                return self.int_scaling_impl(bit_width)       # (a tensor born inside a capture belongs to that graph)
            with torch.no_grad():
                thr = self.int_scaling_impl(bit_width)
            cache[key] = thr
        return thr

    def forward_pre_relu(self, x: Tensor) -> Optional[Tuple[Tensor, Tensor, Tensor, Tensor]]:
        """``self(torch.relu(x))`` in ONE kernel (``relu_int_quant``: the ReLU's own read + write pass and its
        backward pass disappear), or None when the configuration does not allow it: the threshold must not depend on
        the activation (learned / constant scale, or runtime statistics past their collection phase), the range
        inputs must be construction-time constants and there must be no quantization delay."""
        if not x.is_cuda or type(self.int_quant.delay_wrapper.delay_impl) is not _NoDelay:
            return None
        independent = getattr(self.scaling_impl, 'input_independent', None)
        if independent is None:
            return None
        collecting = False
        if not independent():
            # statistics-collection phase: the threshold is a percentile of relu(x).  The select kernel takes relu(x)'s
            # statistic straight from x, so the ReLU still folds into the quantizer kernel
            probe = getattr(self.scaling_impl, 'pre_relu_collecting', None)
            if probe is None or not probe(x):
                return None
            collecting = True
        bit_width = self.msb_clamp_bit_width_impl()
        cfg = self._host_config(bit_width.dtype)
        if cfg is None:
            return None
        zp, qmin, qmax, rm, cm, _ = cfg
        if zp is None:
            return None
        if not collecting:
            threshold = self.scaling_impl(x)                # x is ignored (input independent)
        elif x.dtype != torch.float32 and x.dtype != self.scaling_impl.buffer.dtype:
            return None                                     # mixed-dtype collection: the two-kernel path handles it
        else:
            holder = {}                                     # pairs the statistic with the quantizer call below
            threshold = self.scaling_impl(x, pre_relu=holder)
        scale = threshold / self.int_threshold(bit_width)
        if not (scale.dtype == x.dtype or (scale.numel() == 1 and scale.dtype == torch.float32)):
            return None
        zero_point = self.zero_point_impl(x, scale, bit_width)
        if collecting:
            y = _ops.CollectingReluQuant.apply(x, scale, zp, qmin, qmax, rm, cm, holder)
        else:
            y = torch.ops.brevitas_b200.relu_int_quant(x, scale, zp, qmin, qmax, rm, cm)
        return y, scale, zero_point, bit_width

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        bit_width = self.msb_clamp_bit_width_impl()
        cfg = self._host_config(bit_width.dtype) if x.is_cuda else None
        if cfg is None:
            if not x.is_cuda:
                raise RuntimeError("brevitas_b200.RescalingIntQuant: CPU tensors are not supported (no CPU fallback)")
            threshold = self.scaling_impl(x)
            int_threshold = self.int_scaling_impl(bit_width)
            scale = threshold / int_threshold
            zero_point = self.zero_point_impl(x, scale, bit_width)
            y = self.int_quant(scale, zero_point, bit_width, x)
            return y, scale, zero_point, bit_width
        zp, qmin, qmax, rm, cm, int_thr = cfg
        if zp is None:
            # asymmetric: scale and zero-point from their (statistics-sized) implementations, the tensor itself through
            # ONE kernel with the zero-point as a device operand; its backward returns d(scale) and d(zero_point)
            with shared_minmax():           # AbsMinMax (scale) and NegativeMinOrZero (zero-point) share one read of x
                threshold = self.scaling_impl(x)
                scale = threshold / self.int_threshold(bit_width)
                zero_point = self.zero_point_impl(x, scale, bit_width)
            y = self.int_quant.forward_fused_zpt(scale, zero_point, qmin, qmax, x)
            if y is None:
                y = self.int_quant(scale, zero_point, bit_width, x)
            return y, scale, zero_point, bit_width
        plan_fn = getattr(self.scaling_impl, 'fused_stats_plan', None)
        plan = plan_fn(x) if (plan_fn is not None and int_thr is not None) else None
        if plan is not None:
            g = plan.geom
            if g.kind == 'rows':
                y, scale, absmax = torch.ops.brevitas_b200.rows_absmax_int_quant(
                    x, g.rows, g.cols, plan.scaling_min_val, int_thr, zp, qmin, qmax, rm, cm)
            else:
                scale_dtype = torch.promote_types(x.dtype, bit_width.dtype)
                y, scale, absmax = torch.ops.brevitas_b200.tensor_absmax_int_quant(
                    x, scale_dtype, plan.scaling_min_val, int_thr, zp, qmin, qmax, rm, cm)
            scale = scale.view(plan.out_shape)
            if plan.on_absmax is not None:
                plan.on_absmax(absmax.view(plan.out_shape))
            zero_point = self.zero_point_impl(x, scale, bit_width)
            return self.int_quant.delay_wrapper(x, y), scale, zero_point, bit_width
        threshold = self.scaling_impl(x)
        int_threshold = self.int_threshold(bit_width)
        scale = threshold / int_threshold
        zero_point = self.zero_point_impl(x, scale, bit_width)
        if scale.dtype != x.dtype and scale.numel() != 1:
            x = x.to(torch.promote_types(x.dtype, scale.dtype))      # what ATen's type promotion would do
        y = self.int_quant.forward_fused(scale, zp, qmin, qmax, x)
        if y is None:
            y = self.int_quant(scale, zero_point, bit_width, x)
        return y, scale, zero_point, bit_width


class BinaryQuant(nn.Module):
    """``y = binary_sign_ste(x) * scale``; zero-point 0, bit-width 1 (binary.py:19-64)."""

    def __init__(self, scaling_impl: nn.Module, quant_delay_steps: int = 0):
        super().__init__()
        self.scaling_impl = scaling_impl
        self.bit_width = BitWidthConst(1)
        self.zero_point = StatelessBuffer(torch.tensor(0.0))
        self.delay_wrapper = DelayWrapper(quant_delay_steps)

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        scale = self.scaling_impl(x)
        if scale.dtype == x.dtype or (scale.numel() == 1 and scale.dtype == torch.float32):
            y = torch.ops.brevitas_b200.binary_quant(x, scale, False)
        else:
            y = binary_sign_ste(x) * scale
        y = self.delay_wrapper(x, y)
        return y, scale, self.zero_point(), self.bit_width()


class ClampedBinaryQuant(nn.Module):
    """``y = binary_sign_ste(clamp(x, -scale, scale)) * scale``: the clamp shapes the gradient (binary.py:67-125)."""

    def __init__(self, scaling_impl: nn.Module, tensor_clamp_impl: Optional[nn.Module] = None,
                 quant_delay_steps: int = 0):
        super().__init__()
        self.scaling_impl = scaling_impl
        self.bit_width = BitWidthConst(1)
        self.zero_point = StatelessBuffer(torch.tensor(0.0))
        self.delay_wrapper = DelayWrapper(quant_delay_steps)
        self.tensor_clamp_impl = tensor_clamp_impl if tensor_clamp_impl is not None else TensorClamp()

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        scale = self.scaling_impl(x)
        fusable = type(self.tensor_clamp_impl) is TensorClamp and (
            scale.dtype == x.dtype or (scale.numel() == 1 and scale.dtype == torch.float32))
        if fusable:
            y = torch.ops.brevitas_b200.binary_quant(x, scale, True)
        else:
            y = self.tensor_clamp_impl(x, - scale, scale)
            y = binary_sign_ste(y) * scale
        y = self.delay_wrapper(x, y)
        return y, scale, self.zero_point(), self.bit_width()


# ---- quantizers either side of the conv / linear of a quantized layer ---------------------
class PrescaledRestrictIntQuantWithInputBitWidth(nn.Module):
    """Scale given by the caller, zero-point 0, bit-width derived from an input bit-width (int.py:17-68); the bias
    quantizer ``Int*Bias`` with an accumulator-dependent bit-width."""

    def __init__(self, int_quant: nn.Module, bit_width_impl: nn.Module):
        super().__init__()
        self.int_quant = int_quant
        self.msb_clamp_bit_width_impl = bit_width_impl
        self.zero_point = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor, scale: Tensor, input_bit_width: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        bit_width = self.msb_clamp_bit_width_impl(input_bit_width)
        zero_point = self.zero_point()
        y = self.int_quant(scale, zero_point, bit_width, x)
        return y, scale, zero_point, bit_width


class PrescaledRestrictIntQuant(nn.Module):
    """Scale given by the caller, zero-point 0, own bit-width (int.py:71-91): ``IntBias``-style quantizers.  With a
    constant bit-width the integer range is a host constant and ``x`` goes through ONE fused kernel."""

    def __init__(self, int_quant: nn.Module, bit_width_impl: nn.Module):
        super().__init__()
        self.int_quant = int_quant
        self.msb_clamp_bit_width_impl = bit_width_impl
        self.zero_point = StatelessBuffer(torch.tensor(0.0))

    def forward(self, x: Tensor, scale: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        msb_clamp_bit_width = self.msb_clamp_bit_width_impl()
        zero_point = self.zero_point()
        iq = self.int_quant
        if x.is_cuda and type(iq) is IntQuant and type(self.msb_clamp_bit_width_impl) is BitWidthConst:
            bw = self.msb_clamp_bit_width_impl.bit_width_value
            qmin, qmax = int_range(iq.signed, iq.narrow_range, bw, msb_clamp_bit_width.dtype)
            y = iq.forward_fused(scale, 0.0, qmin, qmax, x)
            if y is not None:
                return y, scale, zero_point, msb_clamp_bit_width
        y = iq(scale, zero_point, msb_clamp_bit_width, x)
        return y, scale, zero_point, msb_clamp_bit_width


class TruncIntQuant(nn.Module):
    """Drop LSBs of an already quantized value: ``round(x/s + zp) / 2^(in_bits - out_bits)`` -> float_to_int (floor
    by default) -> dequantize (int.py:199-229; ``QuantAvgPool2d``).  The tensors involved are pooled activations
    (small), the bit-widths are device tensors: literal sequence on the STE kernels."""

    def __init__(self, float_to_int_impl: nn.Module, bit_width_impl: nn.Module, quant_delay_steps: int = 0):
        super().__init__()
        self.msb_clamp_bit_width_impl = bit_width_impl
        self.float_to_int_impl = float_to_int_impl
        self.delay_wrapper = DelayWrapper(quant_delay_steps)

    def forward(self, x: Tensor, scale: Tensor, zero_point: Tensor, input_bit_width: Tensor):
        y = scalar_div(x, scale)
        y = y + zero_point
        y = round_ste(y)  # clean up floating point error
        output_bit_width = self.msb_clamp_bit_width_impl()
        trunc_bit_width = input_bit_width - output_bit_width
        trunc_scale = 2.0 ** trunc_bit_width
        y = y / trunc_scale
        y = self.float_to_int_impl(y)
        y = y - zero_point
        y = scalar_mul(y, scale)
        y = self.delay_wrapper(x, y)
        return y, scale, zero_point, output_bit_width


# ---- decoupled / ternary ------------------------------------------------------------------
class DecoupledIntQuant(nn.Module):
    """Integer quantization whose rounding uses (pre_scale, pre_zero_point) and whose dequantization uses
    (scale, zero_point) (int_base.py:100-182)."""

    def __init__(self, narrow_range: bool, signed: bool, float_to_int_impl: Optional[nn.Module] = None,
                 tensor_clamp_impl: Optional[nn.Module] = None, quant_delay_steps: int = 0):
        super().__init__()
        self.float_to_int_impl = float_to_int_impl if float_to_int_impl is not None else RoundSte()
        self.tensor_clamp_impl = tensor_clamp_impl if tensor_clamp_impl is not None else TensorClamp()
        self.signed = signed
        self.narrow_range = narrow_range
        self.delay_wrapper = DelayWrapper(quant_delay_steps)

    def to_int(self, pre_scale: Tensor, pre_zero_point: Tensor, bit_width: Tensor, x: Tensor) -> Tensor:
        y = scalar_div(x, pre_scale)
        y = y + pre_zero_point
        min_int_val = self.min_int(bit_width)
        max_int_val = self.max_int(bit_width)
        y = self.float_to_int_impl(y)
        y = self.tensor_clamp_impl(y, min_val=min_int_val, max_val=max_int_val)
        return y

    def min_int(self, bit_width):
        return min_int(self.signed, self.narrow_range, bit_width)

    def max_int(self, bit_width):
        return max_int(self.signed, self.narrow_range, bit_width)

    def forward(self, pre_scale: Tensor, pre_zero_point: Tensor, scale: Tensor, zero_point: Tensor, bit_width: Tensor,
                x: Tensor) -> Tensor:
        rm = ROUND_MODE_OF.get(type(self.float_to_int_impl))
        cm = CLAMP_MODE_OF.get(type(self.tensor_clamp_impl))
        if (rm is not None and cm is not None and not pre_zero_point.requires_grad and not zero_point.requires_grad
                and _K.general_int_quant_supported(x, pre_scale, scale, pre_zero_point, zero_point, bit_width)):
            # one kernel: rounding with (pre_scale, pre_zero_point), dequantization with (scale, zero_point); its backward
            # returns d(pre_scale) and d(scale) separately (and the bounds' gradients under a masked clamp)
            y = torch.ops.brevitas_b200.general_int_quant(x, pre_scale, scale, pre_zero_point, zero_point,
                                                          self.min_int(bit_width), self.max_int(bit_width), rm, cm, False)
            return self.delay_wrapper(x, y)
        y_int = self.to_int(pre_scale, pre_zero_point, bit_width, x)
        y = y_int - zero_point
        y = scalar_mul(y, scale)
        return self.delay_wrapper(x, y)


class DecoupledRescalingIntQuant(nn.Module):
    """int.py:166-196: returns ``(y, scale, zero_point, bit_width, pre_scale, pre_zero_point)``."""

    def __init__(self, decoupled_int_quant: nn.Module, pre_scaling_impl: nn.Module, scaling_impl: nn.Module,
                 int_scaling_impl: nn.Module, pre_zero_point_impl: nn.Module, zero_point_impl: nn.Module,
                 bit_width_impl: nn.Module):
        super().__init__()
        self.decoupled_int_quant = decoupled_int_quant
        self.pre_scaling_impl = pre_scaling_impl
        self.scaling_impl = scaling_impl
        self.int_scaling_impl = int_scaling_impl
        self.pre_zero_point_impl = pre_zero_point_impl
        self.zero_point_impl = zero_point_impl
        self.msb_clamp_bit_width_impl = bit_width_impl

    def forward(self, x: Tensor):
        bit_width = self.msb_clamp_bit_width_impl()
        int_threshold = self.int_scaling_impl(bit_width)
        pre_threshold = self.pre_scaling_impl(x)
        pre_scale = pre_threshold / int_threshold
        pre_zero_point = self.pre_zero_point_impl(x, pre_scale, bit_width)
        threshold = self.scaling_impl(x)
        scale = threshold / int_threshold
        zero_point = self.zero_point_impl(x, scale, bit_width)
        y = self.decoupled_int_quant(pre_scale, pre_zero_point, scale, zero_point, bit_width, x)
        return y, scale, zero_point, bit_width, pre_scale, pre_zero_point


class TernaryQuant(nn.Module):
    """``y = (|x| > threshold * scale) * ternary_sign_ste(x) * scale``; zero-point 0, bit-width 2 (ternary.py:17-72)."""

    def __init__(self, scaling_impl: nn.Module, threshold: float, quant_delay_steps: int = None):
        super().__init__()
        self.scaling_impl = scaling_impl
        self.threshold = threshold
        self.bit_width = BitWidthConst(2)
        self.zero_point = StatelessBuffer(torch.tensor(0.0))
        self.delay_wrapper = DelayWrapper(quant_delay_steps)

    def forward(self, x: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        scale = self.scaling_impl(x)
        if x.is_cuda and x.dtype == torch.float32 and scale.dtype == torch.float32 and scale.numel() == 1:
            y = torch.ops.brevitas_b200.ternary_quant(x, scale, float(self.threshold))
            return self.delay_wrapper(x, y), scale, self.zero_point(), self.bit_width()
        mask = x.abs().gt(self.threshold * scale)
        y = mask.float() * ternary_sign_ste(x)
        y = y * scale
        y = self.delay_wrapper(x, y)
        return y, scale, self.zero_point(), self.bit_width()
