"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/brevitas_b200.h declares (no
compute calls), and the host-side logic that needs no device: broadcast patterns, integer ranges, module trees,
state-dict keys, loud failure on CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from brevitas_b200 import _lib
    header = open(os.path.join(ROOT, "include", "brevitas_b200.h")).read()
    product = re.sub(r"#ifdef BVB_TUNING_BUILD.*?#endif", "", header, flags=re.S)     # sweep-build-only declarations
    declared = set(re.findall(r"\b(bvb_\w+)\s*\(", product))
    assert len(declared) >= 30 and "bvb_set_tuning" not in declared
    exported = os.popen(f"nm -D {_lib.LIB_PATH}").read()
    assert "bvb_set_tuning" not in exported, "the product library must not carry the mutable tuning state"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported by the library"
    loaded = _lib.load()
    assert loaded.bvb_version() == 100
    assert loaded.bvb_workspace_bytes() >= 64 * 1024
    base = 4 * 4 * 3 * 256 + 8 * 3 * 256                                            # histograms + first-index table
    assert loaded.bvb_kth_workspace_bytes(3) == base
    assert loaded.bvb_kth_workspace_bytes(2) == 4 * 4 * 2 * 256 + 8 * 2 * 256 + 256 + 2 * (1 << 22) * 12   # + candidates


def test_header_cites_reference_lines():
    header = open(os.path.join(ROOT, "include", "brevitas_b200.h")).read()
    for needle in ("csrc/autograd_ste_ops.cpp:258-271", "int_base.py:64-97", "int.py:156-163", "binary.py",
                   "stats_op.py:41-66", "stats_wrapper.py:56-65"):
        assert needle in header


def test_broadcast_pattern():
    from brevitas_b200._kernels import broadcast_pattern as bp
    assert bp((4, 5), ()) == (1, 1)
    assert bp((4, 5), (1,)) == (1, 1)
    assert bp((4, 5), (4, 1)) == (5, 4)
    assert bp((8, 3, 3, 3), (8, 1, 1, 1)) == (27, 8)
    assert bp((2, 6, 5, 5), (1, 6, 1, 1)) == (25, 6)
    assert bp((2, 7, 32), (2, 7, 1)) == (32, 14)
    assert bp((2, 3, 4), (2, 3, 4)) == (1, 24)
    assert bp((2, 3, 4), (4,)) == (1, 4)
    with pytest.raises(RuntimeError):
        bp((2, 3, 4), (2, 1, 4))
    with pytest.raises(RuntimeError):
        bp((2, 3, 4), (5, 1, 1))


def test_int_range_tables():
    """function/ops.py:133-191 for 2..8 bit x signed x narrow (tests/brevitas/function/test_ops.py:60-124)"""
    from brevitas_b200.core.quant import int_range
    for bits in range(2, 9):
        assert int_range(True, True, bits, torch.float32) == (-(2 ** (bits - 1)) + 1, 2 ** (bits - 1) - 1)
        assert int_range(True, False, bits, torch.float32) == (-(2 ** (bits - 1)), 2 ** (bits - 1) - 1)
        assert int_range(False, False, bits, torch.float32) == (0, 2 ** bits - 1)
        assert int_range(False, True, bits, torch.float32) == (0, 2 ** bits - 2)
    # a bf16 bit-width buffer makes the reference compute the range in bf16: 2**9 - 1 is not representable
    assert int_range(False, False, 9, torch.bfloat16) == (0.0, 512.0)


def test_ops_registered_and_fail_loudly_on_cpu():
    import brevitas_b200  # noqa: F401
    from brevitas_b200.ops import STE_OP_NAMES
    for name in STE_OP_NAMES:
        assert hasattr(torch.ops.autograd_ste_ops, name)
    for name in ("int_quant", "rows_absmax_int_quant", "tensor_absmax_int_quant", "binary_quant", "absmax_rows",
                 "absmax_tensor", "abs_kth_value_rows"):
        assert hasattr(torch.ops.brevitas_b200, name)
    x = torch.randn(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.autograd_ste_ops.round_ste_impl(x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        torch.ops.brevitas_b200.int_quant(x, torch.tensor(1.0), 0.0, -127.0, 127.0, 0, 0)
    from brevitas_b200.function.ops_ste import round_ste
    with pytest.raises(RuntimeError):
        round_ste(x)


def test_fake_tensor_shapes():
    """meta / fake implementations exist (needed for torch.compile / export tracing of models using the ops)"""
    import brevitas_b200  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        x = torch.empty(6, 10, device="cuda")
        y, s, a = torch.ops.brevitas_b200.rows_absmax_int_quant(x, 6, 10, 1e-10, 127.0, 0.0, -127.0, 127.0, 0, 0)
        assert y.shape == (6, 10) and s.shape == (6,) and a.shape == (6,)
        assert torch.ops.autograd_ste_ops.round_ste_impl(x).shape == (6, 10)


def _weight_quant(w, per_channel=True):
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, StatsFromParameterScaling
    from brevitas_b200.core.stats import AbsMax
    from brevitas_b200.core.zero_point import ZeroZeroPoint
    if per_channel:
        stats, view, concat, shape = AbsMax(1), fw.OverOutputChannelView(None), 1, (w.shape[0], 1)
    else:
        stats, view, concat, shape = AbsMax(None), fw.OverTensorView(), 0, ()
    return RescalingIntQuant(
        IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClampSte()),
        StatsFromParameterScaling(stats, view, concat, [w], FloatRestrictValue(), shape, False, 1e-10),
        IntScaling(True, True), ZeroZeroPoint(), BitWidthConst(8))


def test_module_tree_and_fusion_plan():
    w = torch.nn.Parameter(torch.randn(16, 40))
    tq = _weight_quant(w)
    # sub-module names other reference code reads (utils/quant_utils.py:16-29, graph/target/flexml.py:112)
    for attr in ("int_quant", "scaling_impl", "int_scaling_impl", "zero_point_impl", "msb_clamp_bit_width_impl"):
        assert hasattr(tq, attr)
    assert list(tq.state_dict().keys()) == []
    cfg = tq._host_config(torch.float32)
    assert cfg == (0.0, -127.0, 127.0, 0, 0, 127.0)
    plan = tq.scaling_impl.fused_stats_plan(w)
    assert plan is not None and plan.geom == ("rows", 16, 40) and plan.out_shape == (16, 1)
    assert abs(plan.scaling_min_val - 1e-10) < 1e-20
    # statistic taken over a different tensor than the one being quantized -> not fusable
    assert tq.scaling_impl.fused_stats_plan(torch.randn(16, 40)) is None
    tq2 = _weight_quant(w, per_channel=False)
    assert tq2.scaling_impl.fused_stats_plan(w).geom == ("tensor", 1, 640)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tq(w)


def test_runtime_stats_plan_and_state():
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.restrict_val import FloatRestrictValue, LogFloatRestrictValue
    from brevitas_b200.core.scaling import ParameterFromRuntimeStatsScaling, ParameterScaling, RuntimeStatsScaling
    from brevitas_b200.core.stats import AbsMax, AbsPercentile
    s = RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(), (2, 5, 1), False, 0.1, 1e-10)
    x = torch.randn(2, 5, 32)
    s.train()
    assert s.fused_stats_plan(x).geom == ("rows", 10, 32)
    s.eval()
    assert s.fused_stats_plan(x) is None                     # eval uses the running buffer
    assert list(s.state_dict().keys()) == ["runtime_stats.running_stats"]
    # non-abs-max statistics are never fused
    s2 = RuntimeStatsScaling(AbsPercentile(99.0, None), fw.OverTensorView(), FloatRestrictValue(), (), False, 0.1, None)
    s2.train()
    assert s2.fused_stats_plan(x) is None
    p = ParameterScaling(6.0, (1, 8, 1, 1), LogFloatRestrictValue(), 2e-16)
    assert list(p.state_dict().keys()) == ["value"] and p.value.shape == (1, 8, 1, 1)
    assert torch.allclose(p.value, torch.full((1, 8, 1, 1), 6.0).log2())
    legacy = {"learned_value": torch.zeros(1, 8, 1, 1)}
    p.load_state_dict(legacy)
    assert float(p.value.abs().sum()) == 0.0
    q = ParameterFromRuntimeStatsScaling(300, AbsPercentile(99.999, None), fw.OverTensorView(), ())
    assert list(q.state_dict().keys()) == []                 # counter == 0: neither buffer nor value
    q.load_state_dict({"value": torch.tensor(2.5)})
    assert q.counter == 301 and float(q.value) == 2.5         # a loaded value ends the collection phase


def test_percentile_k():
    from brevitas_b200._kernels import percentile_k
    assert [percentile_k(10.0 * v, 10) for v in range(1, 11)] == list(range(1, 11))   # test_stats.py:12-18
    assert percentile_k(99.999, 1000) == 1000 and percentile_k(90.0, 10) == 9


# ---------------------------------------------------------------------------------------------------------------
# packed bf16x2 / f16x2 fast path (csrc/common.cuh qdq_vec, int_quant.cu bwd_vec): the algebra it relies on, checked
# for ALL 2^16 values of the dtype with the constants the library's host code derives (no GPU needed)
# ---------------------------------------------------------------------------------------------------------------
PACKED_RANGES = [(-127, 127), (-128, 127), (0, 255), (0, 254), (-8, 7), (-7, 7), (0, 15), (-1, 1), (0, 1), (-2, 1),
                 (0, 3), (-512, 511), (-32, 31)]


def _packed_constants(lo, hi, dtype):
    from brevitas_b200 import _lib
    out = (ctypes.c_uint32 * 7)()
    _lib.call("bvb_debug_packed_constants", 0.0, float(lo), float(hi), {"bf16": _lib.BF16, "f16": _lib.F16}[dtype], out)
    return list(out)


def _from_bits16(bits, tdt):
    import numpy as np
    return torch.from_numpy(np.asarray(bits, dtype=np.uint16).view(np.int16)).view(tdt)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("lo,hi", PACKED_RANGES)
def test_packed_path_algebra_exhaustive(dtype, lo, hi):
    """clamp(round(v)) == signfix(magic_round(clamp'(v))) and the threshold masks == the masks on round(v), bit for
    bit, for every value v of the dtype (v plays t1 = rnd_T(x / s))."""
    import numpy as np
    tdt = {"bf16": torch.bfloat16, "f16": torch.float16}[dtype]
    ok, lo_zero, plo, phi, plo_pre, pthr_lo, pthr_hi = _packed_constants(lo, hi, dtype)
    representable = all(float(torch.tensor(float(b)).to(tdt)) == b for b in (lo, hi))
    if not representable or (dtype == "f16" and (lo < -512 or hi > 511)):
        assert ok == 0                                                 # the kernels keep the literal op sequence
        return
    assert ok == 1 and lo_zero == int(lo == 0)
    half = lambda w: _from_bits16([w & 0xffff], tdt)[0]
    assert float(half(plo)) == lo and float(half(phi)) == hi and float(half(plo_pre)) == (-1 if lo == 0 else lo)
    v = _from_bits16(np.arange(65536, dtype=np.uint32).astype(np.uint16), tdt)
    t2 = v + torch.zeros((), dtype=tdt)                               # the reference's "+ zero_point" (0): -0 -> +0
    t3 = torch.round(t2)
    qlo, qhi = torch.tensor(float(lo), dtype=tdt), torch.tensor(float(hi), dtype=tdt)
    t5 = torch.where(t3 > qhi, qhi, t3)
    t5 = torch.where(t5 < qlo, qlo, t5)                                # brevitas.function.ops.tensor_clamp
    keep = ~(t3 > qhi) & ~(t3 < qlo)
    # --- packed formulation ---
    c = torch.maximum(torch.minimum(t2, half(phi)), half(plo_pre))     # HMNMX2.NAN (torch.minimum propagates NaN)
    if dtype == "bf16":
        r = ((c.float() + 12582912.0) - 12582912.0).to(tdt)            # fp32 magic adds, exact pack
    else:
        m = torch.tensor(1536.0, dtype=tdt)
        r = (c + m) - m                                                # two fp16 adds (each rounded to fp16)
    rb = r.view(torch.int16).numpy().view(np.uint16) | (c.view(torch.int16).numpy().view(np.uint16) & 0x8000)
    r = _from_bits16(rb, tdt)
    if lo_zero:
        r = torch.where(r < 0, torch.zeros((), dtype=tdt), r)
    nan = torch.isnan(t5) & torch.isnan(r)
    same = (t5.view(torch.int16) == r.view(torch.int16)) | nan
    assert bool(same.all()), f"{int((~same).sum())} codes differ, e.g. v={v[~same][:4]}"
    keep2 = ~(t2 > half(pthr_hi)) & ~(t2 < half(pthr_lo))
    assert bool((keep == keep2).all()), f"masks differ at v={v[keep != keep2][:4]}"


def test_packed_constants_reject_unrepresentable_ranges():
    assert _packed_constants(0, 65535, "bf16")[0] == 0      # 65535 rounds to 65536 in bf16: literal path
    assert _packed_constants(-32768, 32767, "f16")[0] == 0
    assert _packed_constants(0, 1023, "f16")[0] == 0        # outside the fp16 magic-rounding range
    assert _packed_constants(0, 1023, "bf16")[0] == 0       # 1023 needs 10 significant bits
    assert _packed_constants(0, 1024, "bf16")[0] == 1


def test_widened_modules_mirror_reference_trees():
    """state-dict keys and sub-module names of the quantizers / zero-points / bit-widths added for SURVEY.md §8f,
    recorded from the reference's own classes (brevitas.core, BREVITAS_JIT=0): checkpoints stay interchangeable."""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst, BitWidthParameter, MsbClampBitWidth, RemoveBitwidthParameter
    from brevitas_b200.core.quant import (IntQuant, PrescaledRestrictIntQuantWithInputBitWidth, RescalingIntQuant,
                                          TernaryQuant, TruncIntQuant)
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, ParameterScaling, StatsFromParameterScaling
    from brevitas_b200.core.stats import AbsMinMax, MeanLearnedSigmaStd, NegativeMinOrZero
    from brevitas_b200.core.zero_point import ParameterFromRuntimeZeroPoint, ParameterZeroPoint, StatsFromParameterZeroPoint
    w = torch.nn.Parameter(torch.randn(4, 6))
    iq = IntQuant(narrow_range=False, signed=False)
    view = lambda: fw.OverOutputChannelView(None)
    shifted = RescalingIntQuant(
        iq, StatsFromParameterScaling(AbsMinMax(1), view(), 1, [w], FloatRestrictValue(), (4, 1), False, 1e-10),
        IntScaling(False, False), StatsFromParameterZeroPoint(iq, True, view(), 1, NegativeMinOrZero(1), (4, 1), [w]),
        BitWidthConst(8))
    assert list(shifted.state_dict().keys()) == []
    names = {n for n, _ in shifted.named_modules()}
    for expected in ("zero_point_impl.parameter_list_stats.first_tracked_param.view_shape_impl",
                     "zero_point_impl.parameter_list_stats.stats.stats_impl.zero", "zero_point_impl.scale_shift_zero_point",
                     "scaling_impl.parameter_list_stats.stats.stats_impl", "msb_clamp_bit_width_impl.bit_width"):
        assert expected in names, expected
    rz = ParameterFromRuntimeZeroPoint(3, iq, True, NegativeMinOrZero(None), (), fw.OverTensorView(), 0.1)
    assert list(rz.state_dict().keys()) == []                       # nothing to save before the first step
    rz.counter = 2
    assert list(rz.state_dict().keys()) == ["value"]                # the running buffer, saved under `value`
    assert sorted(ParameterZeroPoint(0.5, iq, True, None).state_dict()) == ["value"]
    pre = PrescaledRestrictIntQuantWithInputBitWidth(iq, MsbClampBitWidth(RemoveBitwidthParameter(3), 2, 16))
    assert sorted(pre.state_dict()) == ["msb_clamp_bit_width_impl.bit_width_to_remove_impl.bit_width_coeff"]
    assert sorted(BitWidthParameter(6).state_dict()) == ["bit_width_offset"]
    assert sorted(TernaryQuant(ParameterScaling(0.7), 0.5).state_dict()) == ["scaling_impl.value"]
    assert sorted(TruncIntQuant(fw.FloorSte(), BitWidthConst(4)).state_dict()) == []
    assert sorted(MeanLearnedSigmaStd(3.0, ()).state_dict()) == ["value"]
    # CPU tensors are rejected by every kernel-backed op (no CPU fallback)
    with pytest.raises(RuntimeError):
        shifted(w)


def test_transposed_conv_and_conv1d_layers_construct():
    """nn/quant_convtranspose.py: output channels in dim 1 -> scale shape [1, O, 1, 1] and a permuting stats view"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200 import nn as qnn
    from brevitas_b200.quant import Int8WeightPerChannelFloat
    t = qnn.QuantConvTranspose2d(6, 10, 3, stride=2, weight_quant=Int8WeightPerChannelFloat)
    view = t.weight_quant.tensor_quant.scaling_impl.parameter_list_stats.first_tracked_param.view_shape_impl
    assert tuple(view.permute_impl.permute_dims) == (1, 0, 2, 3)
    c = qnn.QuantConv1d(5, 7, 4, weight_quant=Int8WeightPerChannelFloat)
    assert type(c.weight_quant.tensor_quant).__name__ == "RescalingIntQuant"
