"""Declarative quantizers: the named quantizers of ``brevitas.quant`` resolved to ``tensor_quant`` module trees.

The reference resolves its quantizer classes through dependency-injection solvers on top of the third-party
``dependencies`` package (src/brevitas/inject, src/brevitas/quant/solver/*.py), which is absent from this image and
from the GPU box.  It carries no arithmetic -- only wiring -- so this module restates
the wiring for the hot-path quantizers as plain Python: a quantizer is a class whose attributes are the
reference's directive names (``bit_width``, ``narrow_range``, ``signed``, ``scaling_impl_type``,
``scaling_stats_op``, ``restrict_scaling_type``, ``scaling_per_output_channel``, ``scaling_min_val`` ...),
``let(**overrides)`` derives a variant (what the ``weight_bit_width=4`` style keyword arguments of the layers do,
src/brevitas/nn/mixin/base.py:64-68), and ``tensor_quant(...)`` builds exactly the tree the reference's solvers
would build, out of ``brevitas_b200.core`` modules.
"""
from typing import Optional, Tuple

import math

import torch
from torch import nn

from .core import function_wrapper as fw
from .core.bit_width import BitWidthConst
from .core.quant import (BinaryQuant, ClampedBinaryQuant, IntQuant, PrescaledRestrictIntQuant,
                         PrescaledRestrictIntQuantWithInputBitWidth, RescalingIntQuant, TernaryQuant, TruncIntQuant)
from .core.restrict_val import FloatRestrictValue, LogFloatRestrictValue, PowerOfTwoRestrictValue
from .core.scaling import (ConstScaling, IntScaling, ParameterFromRuntimeStatsScaling, ParameterScaling,
                           PowerOfTwoIntScaling, RuntimeStatsScaling, StatsFromParameterScaling)
from .core.stats import AbsMax, AbsMinMax, AbsPercentile, NegativeMinOrZero, NegativePercentileOrZero, PercentileInterval
from .core.zero_point import ParameterFromRuntimeZeroPoint, StatsFromParameterZeroPoint, ZeroZeroPoint

SCALING_STATS_REDUCE_DIM = 1

_FLOAT_TO_INT = {"ROUND": fw.RoundSte, "FLOOR": fw.FloorSte, "CEIL": fw.CeilSte, "ROUND_TO_ZERO": fw.RoundToZeroSte,
                 "DPU": fw.DPURoundSte}       # quant/solver/common.py:30-43
_RESTRICT = {"FP": FloatRestrictValue, "LOG_FP": LogFloatRestrictValue, "POWER_OF_TWO": PowerOfTwoRestrictValue}


class Quantizer:
    """Base of all declarative quantizers; attributes are directives, never instantiated."""
    quant_type = "INT"                     # INT | BINARY | FP (no quantization)
    bit_width: Optional[int] = 8
    bit_width_impl_type = "CONST"
    float_to_int_impl_type = "ROUND"
    narrow_range = False
    signed = True
    restrict_scaling_type = "FP"
    restrict_value_float_to_int_impl = None
    scaling_per_output_channel = False
    scaling_min_val: Optional[float] = None
    scaling_impl_type = "CONST"
    scaling_stats_op = "MAX"
    scaling_stats_momentum = 0.1
    high_percentile_q = 99.999
    collect_stats_steps = 300
    quant_delay_steps = 0

    @classmethod
    def let(cls, **overrides):
        overrides = {k: v for k, v in overrides.items()}
        return type(cls.__name__, (cls,), overrides)

    # -- shared solvers -------------------------------------------------------------------------------------
    @classmethod
    def _quant_type(cls):
        if cls.bit_width is None or cls.quant_type == "FP":
            return "FP"
        return cls.quant_type

    @classmethod
    def _restrict(cls):
        impl = _RESTRICT[cls.restrict_scaling_type]
        if cls.restrict_scaling_type == "POWER_OF_TWO" and cls.restrict_value_float_to_int_impl is not None:
            return impl(cls.restrict_value_float_to_int_impl())
        return impl()

    @classmethod
    def _int_scaling(cls):
        # quant/solver/common.py:117-128
        if cls.restrict_scaling_type == "POWER_OF_TWO":
            return PowerOfTwoIntScaling(cls.signed)
        return IntScaling(cls.signed, cls.narrow_range)

    @classmethod
    def _stats_impl(cls, reduce_dim):
        if cls.scaling_stats_op == "MAX":
            return AbsMax(reduce_dim)
        if cls.scaling_stats_op == "PERCENTILE":
            return AbsPercentile(cls.high_percentile_q, reduce_dim)
        raise NotImplementedError(f"scaling_stats_op={cls.scaling_stats_op} is outside the B200 hot path (SURVEY.md §8)")

    @classmethod
    def _rescaling_int_quant(cls, scaling_impl, tensor_clamp_impl):
        return RescalingIntQuant(
            int_quant=IntQuant(narrow_range=cls.narrow_range, signed=cls.signed,
                               float_to_int_impl=_FLOAT_TO_INT[cls.float_to_int_impl_type](),
                               tensor_clamp_impl=tensor_clamp_impl, quant_delay_steps=cls.quant_delay_steps),
            scaling_impl=scaling_impl, int_scaling_impl=cls._int_scaling(), zero_point_impl=ZeroZeroPoint(),
            bit_width_impl=BitWidthConst(int(cls.bit_width)))


class WeightQuantizer(Quantizer):
    """Weight quantizers (quant/solver/weight.py, quant/solver/parameter.py)."""
    scaling_const: Optional[float] = None
    scaling_stats_permute_dims = None          # ConvTranspose etc.

    @classmethod
    def tensor_quant(cls, weight: nn.Parameter, output_channel_dim: int = 0) -> Optional[nn.Module]:
        qt = cls._quant_type()
        if qt == "FP":
            return None
        # scaling shape / view / reduce dim: quant/solver/common.py:131-165, parameter.py:137-160
        if cls.scaling_per_output_channel:
            shape = tuple(weight.shape[d] if d == output_channel_dim else 1 for d in range(weight.dim()))
            permute = cls.scaling_stats_permute_dims
            if permute is None and output_channel_dim != 0:
                # transposed convolutions keep their output channels in dim 1: the statistics see them first
                permute = (output_channel_dim,) + tuple(d for d in range(weight.dim()) if d != output_channel_dim)
            view, reduce_dim, concat = fw.OverOutputChannelView(permute), SCALING_STATS_REDUCE_DIM, 1
        else:
            shape = ()
            view, reduce_dim, concat = fw.OverTensorView(), None, 0
        st = cls.scaling_impl_type
        if st == "CONST":
            scaling = ConstScaling(float(cls.scaling_const), cls._restrict(), cls.scaling_min_val)
        elif st == "STATS":
            scaling = StatsFromParameterScaling(cls._stats_impl(reduce_dim), view, concat, [weight], cls._restrict(),
                                                shape, False, cls.scaling_min_val)
        elif st == "PARAMETER":
            scaling = ParameterScaling(float(cls.scaling_const), shape if shape else None, cls._restrict(),
                                       cls.scaling_min_val)
        elif st == "PARAMETER_FROM_STATS":
            # learned scale initialised from the weight statistics, the way the reference does it (solver/parameter.py:
            # 39-45, 88-92: ParameterFromStatsScalingInit calls a StatsFromParameterScaling built from the CONFIGURED
            # scaling_stats_op, restriction and scaling_min_val once, at construction).  The tracked weight may
            # still live on the host at that point, so the one-off statistic runs on plain ATen (construction time, not the
            # hot path): the reference's literal op sequence of _ParameterListStats + _StatsScaling.
            with torch.no_grad():
                restrict = cls._restrict()
                stats_in = view(weight.detach())
                op, q = cls.scaling_stats_op, float(getattr(cls, "high_percentile_q", 99.999))
                a = stats_in.abs()
                if op == "MAX":
                    stat = a.max(dim=reduce_dim)[0] if reduce_dim is not None else a.max()
                elif op == "PERCENTILE":
                    if reduce_dim is None:
                        stat = a.view(-1).kthvalue(int(math.floor(.01 * q * a.numel() + 0.5))).values
                    else:
                        stat = a.kthvalue(int(math.floor(.01 * q * a.shape[reduce_dim] + 0.5)), dim=reduce_dim).values
                elif op == "AVE":
                    stat = a.mean(dim=reduce_dim) if reduce_dim is not None else a.mean()
                elif op == "MAX_AVE":
                    stat = a.max(dim=1)[0].mean()
                else:
                    raise NotImplementedError(f"PARAMETER_FROM_STATS initialisation with scaling_stats_op={op}")
                stat = stat.view(shape)
                stat = restrict.restrict_init_tensor(stat)                      # pre-restriction (log2 for LOG_FP / PoT)
                if cls.restrict_scaling_type == "POWER_OF_TWO" and not stat.is_cuda:
                    # the restriction's float_to_int_impl is an STE kernel (CUDA only); same values with ATen on the host
                    f2i = type(getattr(restrict, "float_to_int_impl", None)).__name__
                    rnd = {"CeilSte": torch.ceil, "FloorSte": torch.floor}.get(f2i, torch.round)
                    stat = 2.0 ** rnd(stat)
                else:
                    stat = restrict(stat)                                       # post-restriction (2 ** ., rounding for PoT)
                if cls.scaling_min_val:
                    stat = stat.clamp_min(cls.scaling_min_val)
                init = stat.clone()
            scaling = ParameterScaling(init, shape if shape else None, cls._restrict(), cls.scaling_min_val)
        else:
            raise NotImplementedError(f"scaling_impl_type={st}")
        if qt == "BINARY":          # quant/solver/weight.py:30-31
            return BinaryQuant(scaling_impl=scaling, quant_delay_steps=cls.quant_delay_steps)
        # clamp flavour: pass-through unless the scale (or bit-width) is learned (parameter.py:64-74)
        clamp = fw.TensorClamp() if st in ("PARAMETER", "PARAMETER_FROM_STATS", "AFFINE_STATS") else fw.TensorClampSte()
        return cls._rescaling_int_quant(scaling, clamp)


class ActQuantizer(Quantizer):
    """Activation quantizers (quant/solver/act.py).  Activations always get the masked ``TensorClamp``."""
    min_val: Optional[float] = None
    max_val: Optional[float] = None
    per_channel_broadcastable_shape: Optional[Tuple[int, ...]] = None
    scaling_stats_permute_dims = None

    @classmethod
    def tensor_quant(cls) -> Optional[nn.Module]:
        qt = cls._quant_type()
        if qt == "FP":
            return None
        shape = tuple(cls.per_channel_broadcastable_shape) if cls.scaling_per_output_channel else ()
        st = cls.scaling_impl_type
        if st in ("CONST", "PARAMETER"):
            min_val = cls.min_val if cls.signed else 0.0                      # act.py:76-81
            init = max(abs(float(min_val)), abs(float(cls.max_val)))           # MinMaxScalingInit, act.py:19-24
            if st == "CONST":
                scaling = ConstScaling(init, cls._restrict(), cls.scaling_min_val)
            else:
                scaling = ParameterScaling(init, shape if shape else None, cls._restrict(), cls.scaling_min_val)
        elif st in ("PARAMETER_FROM_STATS", "STATS"):
            if cls.scaling_per_output_channel:
                view = fw.OverOutputChannelView(cls.scaling_stats_permute_dims)
                reduce_dim = SCALING_STATS_REDUCE_DIM
            else:
                view, reduce_dim = fw.OverTensorView(), None
            if st == "STATS":
                scaling = RuntimeStatsScaling(cls._stats_impl(reduce_dim), view, cls._restrict(), shape, False,
                                              cls.scaling_stats_momentum, cls.scaling_min_val)
            else:
                scaling = ParameterFromRuntimeStatsScaling(cls.collect_stats_steps, cls._stats_impl(reduce_dim), view,
                                                           shape, cls._restrict(), cls.scaling_stats_momentum,
                                                           cls.scaling_min_val)
        else:
            raise NotImplementedError(f"scaling_impl_type={st}")
        if qt == "BINARY":          # act.py:58-59
            return ClampedBinaryQuant(scaling_impl=scaling, quant_delay_steps=cls.quant_delay_steps)
        return cls._rescaling_int_quant(scaling, fw.TensorClamp())


# ---- the named quantizers of brevitas.quant.scaled_int used by the BASELINE configs ------------------------------
class Int8WeightPerTensorFloat(WeightQuantizer):
    """quant/scaled_int.py:144-154 = NarrowIntQuant + MaxStatsScaling + PerTensorFloatScaling8bit"""
    narrow_range = True
    signed = True
    scaling_impl_type = "STATS"
    scaling_stats_op = "MAX"
    scaling_min_val = 1e-10
    scaling_per_output_channel = False
    bit_width = 8


class Int8WeightPerChannelFloat(Int8WeightPerTensorFloat):
    """quant/scaled_int.py:157-167"""
    scaling_per_output_channel = True


class Int8ActPerTensorFloat(ActQuantizer):
    """quant/scaled_int.py:170-180 = IntQuant + ParamFromRuntimePercentileScaling + PerTensorFloatScaling8bit"""
    narrow_range = False
    signed = True
    scaling_impl_type = "PARAMETER_FROM_STATS"
    scaling_stats_op = "PERCENTILE"
    high_percentile_q = 99.999
    collect_stats_steps = 300
    scaling_min_val = 1e-10
    bit_width = 8


class Uint8ActPerTensorFloat(Int8ActPerTensorFloat):
    """quant/scaled_int.py:183-193 (default of QuantReLU)"""
    signed = False


class Uint8ActPerTensorFloatMaxInit(ActQuantizer):
    """quant/scaled_int.py:49-61 = UintQuant + ParamMinMaxInitScaling + PerTensorFloatScaling8bit"""
    narrow_range = False
    signed = False
    scaling_impl_type = "PARAMETER"
    bit_width = 8


class Int8ActPerTensorFloatMinMaxInit(ActQuantizer):
    """quant/scaled_int.py:32-46 = IntQuant + ParamMinMaxInitScaling + PerTensorFloatScaling8bit"""
    narrow_range = False
    signed = True
    scaling_impl_type = "PARAMETER"
    bit_width = 8


class Int8ActPerTokenDynamic(ActQuantizer):
    """Per-token dynamic activation quantizer composed from core parts: RuntimeStatsScaling +
    OverBatchOverOutputChannelView + AbsMax(2), scale shape (B, T, 1).  Not a named quantizer of the reference."""
    narrow_range = False
    signed = True
    scaling_min_val = 1e-10
    bit_width = 8

    @classmethod
    def tensor_quant_for(cls, batch: int, tokens: int):
        scaling = RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(),
                                      (batch, tokens, 1), False, cls.scaling_stats_momentum, cls.scaling_min_val)
        return cls._rescaling_int_quant(scaling, fw.TensorClamp())


# ---- bias / accumulator quantizers (quant/scaled_int.py:64-127, :185-204; quant/solver/bias.py, trunc.py) --------
class BiasQuantizer(Quantizer):
    """``IntBias`` family: signed, scale = input scale x weight scale supplied by the layer, zero-point 0."""
    narrow_range = False
    signed = True
    requires_input_scale = True
    requires_input_bit_width = True            # IntBias: bit-width of the accumulator the bias is added to
    bit_width: Optional[int] = None

    @classmethod
    def tensor_quant(cls) -> nn.Module:
        iq = IntQuant(narrow_range=cls.narrow_range, signed=cls.signed,
                      float_to_int_impl=_FLOAT_TO_INT[cls.float_to_int_impl_type](), tensor_clamp_impl=fw.TensorClamp())
        if cls.requires_input_bit_width:
            return PrescaledRestrictIntQuantWithInputBitWidth(iq, fw.Identity())
        return PrescaledRestrictIntQuant(iq, BitWidthConst(int(cls.bit_width)))


class IntBias(BiasQuantizer):
    """quant/scaled_int.py:64-76"""


class Int8Bias(IntBias):
    bit_width = 8
    requires_input_bit_width = False


class Int16Bias(IntBias):
    bit_width = 16
    requires_input_bit_width = False


class Int24Bias(IntBias):
    bit_width = 24
    requires_input_bit_width = False


class Int32Bias(IntBias):
    bit_width = 32
    requires_input_bit_width = False


class TruncQuantizer(Quantizer):
    """``IntTrunc`` (quant/base.py): keeps the input scale and zero-point, drops LSBs with FLOOR."""
    float_to_int_impl_type = "FLOOR"

    @classmethod
    def tensor_quant(cls) -> nn.Module:
        return TruncIntQuant(_FLOAT_TO_INT[cls.float_to_int_impl_type](), BitWidthConst(int(cls.bit_width)),
                             cls.quant_delay_steps)


class TruncTo8bit(TruncQuantizer):
    """quant/scaled_int.py:196-204"""
    bit_width = 8


# ---- asymmetric weight quantizers (quant/shifted_scaled_int.py:45-75 = ShiftedMinUintQuant + MinMaxStatsScaling) ----
class ShiftedUint8WeightPerTensorFloat(WeightQuantizer):
    narrow_range = False
    signed = False
    scaling_impl_type = "STATS"
    scaling_stats_op = "MIN_MAX"
    scaling_min_val = 1e-10
    bit_width = 8
    quantize_zero_point = True

    @classmethod
    def tensor_quant(cls, weight: nn.Parameter, output_channel_dim: int = 0) -> nn.Module:
        if output_channel_dim != 0:
            raise NotImplementedError("output channels must be dim 0 (Linear / Conv weights)")
        if cls.scaling_per_output_channel:
            shape = (weight.shape[0],) + (1,) * (weight.dim() - 1)
            mk_view, reduce_dim, concat = (lambda: fw.OverOutputChannelView(None)), SCALING_STATS_REDUCE_DIM, 1
        else:
            shape, mk_view, reduce_dim, concat = (), fw.OverTensorView, None, 0
        iq = IntQuant(narrow_range=cls.narrow_range, signed=cls.signed,
                      float_to_int_impl=_FLOAT_TO_INT[cls.float_to_int_impl_type](),
                      tensor_clamp_impl=fw.TensorClampSte(), quant_delay_steps=cls.quant_delay_steps)
        scaling = StatsFromParameterScaling(AbsMinMax(reduce_dim), mk_view(), concat, [weight], cls._restrict(), shape,
                                            False, cls.scaling_min_val)
        zero_point = StatsFromParameterZeroPoint(iq, cls.quantize_zero_point, mk_view(), concat,
                                                 NegativeMinOrZero(reduce_dim), shape, [weight])
        return RescalingIntQuant(int_quant=iq, scaling_impl=scaling, int_scaling_impl=cls._int_scaling(),
                                 zero_point_impl=zero_point, bit_width_impl=BitWidthConst(int(cls.bit_width)))


class ShiftedUint8WeightPerChannelFloat(ShiftedUint8WeightPerTensorFloat):
    scaling_per_output_channel = True


class ShiftedUint8ActPerTensorFloat(ActQuantizer):
    """quant/shifted_scaled_int.py:19-42 = ShiftedParamFromPercentileUintQuant + ParamFromRuntimePercentileIntervalScaling:
    scale from the [0.001, 99.999] percentile interval and zero-point from the low percentile, both collected for
    ``collect_stats_steps`` steps and then learned."""
    narrow_range = False
    signed = False
    bit_width = 8
    scaling_min_val = 1e-10
    high_percentile_q = 99.999
    low_percentile_q = 0.001
    collect_stats_steps = 300
    quantize_zero_point = True

    @classmethod
    def tensor_quant(cls) -> nn.Module:
        iq = IntQuant(narrow_range=cls.narrow_range, signed=cls.signed,
                      float_to_int_impl=_FLOAT_TO_INT[cls.float_to_int_impl_type](), tensor_clamp_impl=fw.TensorClamp(),
                      quant_delay_steps=cls.quant_delay_steps)
        scaling = ParameterFromRuntimeStatsScaling(
            cls.collect_stats_steps, PercentileInterval(cls.low_percentile_q, cls.high_percentile_q, None),
            fw.OverTensorView(), (), cls._restrict(), cls.scaling_stats_momentum, cls.scaling_min_val)
        zero_point = ParameterFromRuntimeZeroPoint(
            cls.collect_stats_steps, iq, cls.quantize_zero_point, NegativePercentileOrZero(cls.low_percentile_q, None), (),
            fw.OverTensorView(), cls.scaling_stats_momentum)
        return RescalingIntQuant(int_quant=iq, scaling_impl=scaling, int_scaling_impl=cls._int_scaling(),
                                 zero_point_impl=zero_point, bit_width_impl=BitWidthConst(int(cls.bit_width)))
