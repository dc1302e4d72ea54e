"""GPU (-m gpu): the UNMODIFIED reference running on the B200 kernels.

1. Op level -- the reference's own ``brevitas.core`` / ``brevitas.function`` classes (from ``oracle/_ref``, imported
   through the real ``brevitas.inject`` on the ``dependencies`` stand-in) with ``ops_ste.fn_prefix = torch``
   (src/brevitas/function/ops_ste.py:38-43): the golden generator ``tests/golden/make_golden.py`` is re-run ON CUDA,
   its tensors flowing through ``torch.ops.autograd_ste_ops.*``, and must reproduce the committed CPU goldens.
2. Module level -- the same with ``brevitas_b200.install(fuse=True)``: the generator's ``from brevitas.core.quant import
   RescalingIntQuant`` now yields the fused classes, all three dtypes.
3. The named quantizers resolved by the reference's REAL injector (``Int8WeightPerChannelFloat`` ... inside
   ``brevitas.nn`` layers) build fused trees and launch the fused kernels; results equal the hand-built trees.
4. Whole-tensor bit compares at BASELINE sizes (C2 4096x11008 fp32 / bf16, C3 [8,2048,4096] bf16 incl. the heavy-tail
   variant): fused kernels against the reference's own modules on the same GPU (ATen, IEEE fp32) and on the CPU.
"""
import os
import sys

import numpy as np
import pytest
import torch

from golden_util import assert_bits_equal, load
from ref_util import FactoryToCuda, import_make_golden, reference_src

pytestmark = pytest.mark.gpu

GROUPS = ["ste", "int_quant", "weight_stats", "runtime_token", "binary", "percentile", "param_from_stats", "kat",
          "widen", "shifted_act"]

# arrays that are (or directly depend on) floating-point SUMS over many elements: summation order differs between the
# CPU run that produced the goldens and any GPU reduction, so they carry a tolerance; everything else is bit-exact
SUM_LEAVES = ("gscale", "gvalue", "gw", "gw_with_gscale", "g_scale_value", "g_zp_value", "g_pre_scale", "g_scale",
              "g_offset")
SUM_STATS = ("abs_max_ave", "abs_max_l2", "abs_ave", "mean_sigma_std")


@pytest.fixture(scope="module")
def ref():
    src = reference_src()
    if src is None:
        pytest.skip("reference not available (run oracle/make_ref.py in the build container)")
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    yield src
    uninstall()


def set_level(src, level):
    import brevitas_b200
    from brevitas_b200.binding import status, uninstall
    uninstall()
    brevitas_b200.install(src, fuse=(level == "fused"))
    st = status()
    assert st["installed"] and st["fused"] == (level == "fused")
    import brevitas.function.ops_ste as ops_ste
    assert ops_ste.fn_prefix is torch


def tol_of(key, dtype_tag):
    eps = {"f32": 2.0 ** -23, "bf16": 2.0 ** -7, "f16": 2.0 ** -10}.get(dtype_tag, 2.0 ** -23)
    return eps


def compare_group(group, out, dtypes, grads_by_tolerance=False):
    gold = load(group)
    assert set(out) == set(gold), sorted(set(out) ^ set(gold))[:10]
    checked = exact = 0
    for key in sorted(out):
        parts = key.split("/")
        tag = next((p for p in parts if p in ("f32", "bf16", "f16")), "f32")
        if tag not in dtypes:
            continue
        got, want = np.asarray(out[key], dtype=np.float64), np.asarray(gold[key], dtype=np.float64)
        leaf = parts[-1]
        is_sum = leaf.startswith(SUM_LEAVES) or any(s in parts for s in SUM_STATS) or \
            (group == "param_from_stats" and leaf.startswith("gvalue")) or \
            (grads_by_tolerance and leaf.startswith("g") and leaf not in ("g", "gs"))
        checked += 1
        g32, w32 = np.asarray(out[key], np.float32), np.asarray(gold[key], np.float32)
        if not is_sum and leaf.startswith("gx") and g32.shape == w32.shape and g32.size > 64:
            # input gradients are element-wise (bit-exact) except for the few entries a statistic routes a SUM to
            # (arg-max of a row / tensor, the k-th value): those may differ by summation order
            diff = ~((g32.view(np.uint32) == w32.view(np.uint32)) | (np.isnan(g32) & np.isnan(w32)))
            if diff.any():
                assert diff.sum() <= max(32, g32.size // 50), f"{group}:{key}: {int(diff.sum())} of {g32.size} differ"
                eps = tol_of(key, tag)
                if not np.allclose(got[diff], want[diff], rtol=64 * eps, atol=64 * eps):
                    # a k-th value / min / max statistic sends its gradient to ONE of several equal elements; which
                    # one is implementation-defined (ATen's CPU and CUDA kthvalue disagree too): same gradient values,
                    # at positions holding equal inputs
                    xk = key[:key.rindex("/") + 1] + "x"
                    assert "stats" in parts and xk in gold, f"{group}:{key}"
                    xs = np.asarray(gold[xk], np.float64).reshape(got.shape)
                    assert np.allclose(np.sort(got[diff]), np.sort(want[diff]), rtol=64 * eps, atol=64 * eps), key
                    assert np.array_equal(np.sort(np.abs(xs[diff & (got != 0)])), np.sort(np.abs(xs[diff & (want != 0)]))), key
            else:
                exact += 1
        elif not is_sum:
            assert_bits_equal(g32, w32, f"{group}:{key}")
            exact += 1
        else:
            eps = tol_of(key, tag)
            fin = np.abs(want[np.isfinite(want)])
            mag = max(1.0, float(fin.max()) if fin.size else 1.0)
            bad = ~(np.isclose(got, want, rtol=64 * eps, atol=64 * eps * mag) | (np.isnan(got) & np.isnan(want)))
            assert not bad.any(), f"{group}:{key}: {int(bad.sum())} of {bad.size} beyond tolerance; " \
                                  f"max |d| {np.nanmax(np.abs(got - want))}"
    assert checked > 0
    return checked, exact


def run_generator(group):
    mg = import_make_golden()
    fn = {"ste": mg.gen_ste, "int_quant": mg.gen_int_quant, "weight_stats": mg.gen_weight_stats,
          "runtime_token": mg.gen_runtime_token, "binary": mg.gen_binary, "percentile": mg.gen_percentile,
          "param_from_stats": mg.gen_param_from_stats, "kat": mg.gen_docstring_kats, "widen": mg.gen_widen,
          "shifted_act": mg.gen_shifted_act}[group]
    out = {}
    torch.manual_seed(123456)
    with FactoryToCuda():
        fn(out)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("group", GROUPS)
def test_reference_core_on_b200_ops_reproduces_goldens(ref, group):
    """Level 1: the reference's OWN modules; only the 12 STE ops are ours.  fp32 is compared for every group (ATen's
    CUDA element-wise arithmetic is IEEE, like the CPU's).  16-bit dtypes are compared for the STE group only: with a
    0-dim fp32 operand next to a 16-bit CUDA tensor ATen-CUDA rounds the scalar to 16 bits first, ATen-CPU (the
    goldens) does not, so the reference itself differs between its two devices there."""
    from brevitas_b200 import _kernels
    set_level(ref, "ops")
    from brevitas.core.quant import IntQuant
    assert IntQuant.__module__.startswith("brevitas.core.quant"), "level 1 must run the reference's own classes"
    before = _kernels.launch_count
    out = run_generator(group)
    launched = _kernels.launch_count - before
    dtypes = ("f32", "bf16", "f16") if group == "ste" else ("f32",)
    checked, exact = compare_group(group, out, dtypes)
    if group in ("ste", "int_quant", "weight_stats", "runtime_token", "binary", "kat", "widen", "shifted_act"):
        assert launched > 0, "the reference did not dispatch into torch.ops.autograd_ste_ops.* kernels"
    print(f"{group}: {checked} arrays ({exact} bit-exact), {launched} B200 kernel launches from reference code")


@pytest.mark.parametrize("group", GROUPS)
def test_reference_with_fused_classes_reproduces_goldens(ref, group):
    """Level 2: ``install(fuse=True)``; the generator's imports from ``brevitas.core.*`` now give the fused classes."""
    from brevitas_b200 import _kernels
    set_level(ref, "fused")
    from brevitas.core.quant import RescalingIntQuant
    import brevitas_b200.core.quant as bq
    assert RescalingIntQuant is bq.RescalingIntQuant
    before = _kernels.launch_count
    out = run_generator(group)
    launched = _kernels.launch_count - before
    checked, exact = compare_group(group, out, ("f32", "bf16", "f16"))
    assert launched > 0
    print(f"{group}: {checked} arrays ({exact} bit-exact), {launched} launches")


# ---------------------------------------------------------------------------------------------------------------
# 3. the reference's REAL injector + brevitas.nn layers
# ---------------------------------------------------------------------------------------------------------------
def test_named_quantizers_through_real_injector(ref):
    from brevitas_b200 import _kernels
    set_level(ref, "fused")
    import brevitas.nn as qnn
    import brevitas_b200.core.quant as bq
    import brevitas_b200.core.scaling as bs
    from brevitas.quant import (Int8ActPerTensorFloat, Int8WeightPerChannelFloat, Int8WeightPerTensorFloat,
                                Uint8ActPerTensorFloat)
    mg = import_make_golden()
    torch.manual_seed(0)
    lin = qnn.QuantLinear(96, 16, bias=True, weight_quant=Int8WeightPerChannelFloat).cuda()
    tq = lin.weight_quant.tensor_quant
    assert type(tq) is bq.RescalingIntQuant and type(tq.scaling_impl) is bs.StatsFromParameterScaling
    gold = load("weight_stats")
    for tag, quant in (("chan_lin", Int8WeightPerChannelFloat), ("tensor_lin", Int8WeightPerTensorFloat)):
        w = torch.from_numpy(gold[f"weight_stats/{tag}/f32/w"]).cuda()
        g = torch.from_numpy(gold[f"weight_stats/{tag}/f32/g"]).cuda()
        layer = qnn.QuantLinear(96, 16, bias=False, weight_quant=quant).cuda()
        with torch.no_grad():
            layer.weight.copy_(w)
        before = _kernels.launch_count
        qw = layer.quant_weight()
        assert _kernels.launch_count - before == 1, "one fused kernel per weight quantization"
        (qw.value * g).sum().backward()
        assert_bits_equal(qw.value.detach().cpu().numpy(), gold[f"weight_stats/{tag}/f32/y"], tag + " y")
        assert_bits_equal(qw.scale.detach().cpu().numpy().reshape(-1), gold[f"weight_stats/{tag}/f32/scale"].reshape(-1),
                          tag + " scale")
        gw, want = layer.weight.grad.cpu().numpy(), gold[f"weight_stats/{tag}/f32/gw"]
        assert np.allclose(gw, want, rtol=1e-5, atol=1e-5)
        assert (gw.view(np.uint32) == want.view(np.uint32)).mean() > 0.98     # all but the arg-max entries
        assert qw.bit_width.item() == 8.0 and qw.signed
    # activations: collection phase -> learned scale, ReLU folded into the quantizer once the scale is a parameter
    relu = qnn.QuantReLU(act_quant=Uint8ActPerTensorFloat, collect_stats_steps=2).cuda().train()
    fq = relu.act_quant.fused_activation_quant_proxy
    assert type(fq).__module__ == "brevitas_b200.nn"
    x = torch.randn(8, 16, 14, 14, device="cuda")
    for _ in range(3):
        y = relu(x)
    relu_calls = []
    fq.activation_impl.register_forward_hook(lambda *a: relu_calls.append(1))
    before = _kernels.launch_count
    y = relu(x.requires_grad_(True))
    # the learned scale goes through the reference's own two 1-element ops (scalar_clamp_min_ste, abs_binary_sign_grad;
    # core/scaling/standalone.py:250-251), the activation through ONE kernel with the ReLU folded in
    assert _kernels.launch_count - before == 3 and not relu_calls, (_kernels.launch_count - before, relu_calls)
    y.sum().backward()
    s = relu.quant_act_scale()
    codes = (y / s).detach()
    assert torch.equal(codes.round(), codes) and codes.min() >= 0 and codes.max() <= 255
    ident = qnn.QuantIdentity(act_quant=Int8ActPerTensorFloat, return_quant_tensor=True).cuda().train()
    qt = ident(x.detach())
    assert qt.scale is not None and qt.bit_width.item() == 8.0


def test_example_models_unmodified_on_fused_kernels(ref):
    """bnn_pynq TFC 2W2A (config 1) and MobileNetV1 4b (config 5), built by the reference's own model code."""
    from brevitas_b200 import _kernels
    set_level(ref, "fused")
    from brevitas_examples.bnn_pynq.models import model_with_cfg
    from brevitas_examples.imagenet_classification.models import model_with_cfg as imagenet_model
    torch.manual_seed(0)
    model, _ = model_with_cfg("tfc_2w2a", False)
    model = model.cuda().train()
    x = torch.rand(64, 1, 28, 28, device="cuda")
    before = _kernels.launch_count
    out = model(x)
    out.square().mean().backward()
    n = _kernels.launch_count - before
    assert out.shape == (64, 10) and torch.isfinite(out).all() and n > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters() if p.requires_grad)
    model, _ = imagenet_model("quant_mobilenet_v1_4b", False)
    model = model.cuda().train()
    x = torch.randn(4, 3, 224, 224, device="cuda")
    before = _kernels.launch_count
    out = model(x)
    out.square().mean().backward()
    assert out.shape == (4, 1000) and torch.isfinite(out).all() and _kernels.launch_count > before


def test_uninstall_restores_reference(ref):
    set_level(ref, "fused")
    from brevitas_b200.binding import uninstall
    uninstall()
    import brevitas
    import brevitas.core.quant as q
    import brevitas.function.ops_ste as ops_ste
    assert ops_ste.fn_prefix is brevitas and q.RescalingIntQuant.__module__ == "brevitas.core.quant.int"
    from brevitas.quant.base import NarrowIntQuant
    assert NarrowIntQuant.__dependencies__["zero_point_impl"][1].__module__ == "brevitas.core.zero_point"


# ---------------------------------------------------------------------------------------------------------------
# 4. whole tensors at BASELINE sizes: fused kernels vs the reference's own modules
# ---------------------------------------------------------------------------------------------------------------
def _weight_tree(per_channel, w):
    mg = import_make_golden()
    return mg.build_weight_quant(w, per_channel)


def _bits(t):
    t = t.detach().contiguous()
    return t.view(torch.int32) if t.dtype == torch.float32 else t.view(torch.int16)


def _assert_same_bits(a, b, what):
    same = (_bits(a) == _bits(b)) | (torch.isnan(a) & torch.isnan(b))
    assert bool(same.all()), f"{what}: {int((~same).sum())} of {a.numel()} elements differ"


def _fp64_sum_bound(abs_terms64, n):
    """An fp32 sum of n terms, in ANY order, against the exact (fp64) sum: 8 * sqrt(n) * 2^-24 * mean|term|, i.e. the
    random-walk error of the n term roundings and of a tree / blocked accumulation with an 8x margin (a strictly
    sequential fp32 accumulation could exceed it; neither ATen nor the kernels sum that way).  Observed: ~1/16 of it."""
    return 8.0 * np.sqrt(n) * 2.0 ** -24 * abs_terms64.sum(dim=-1) / n


@pytest.mark.parametrize("per_channel", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_c2_whole_tensor_vs_reference(ref, per_channel, dtype):
    """BASELINE config 2: Int8 per-output-channel (and per-tensor) weight fake-quant fwd + STE bwd on 4096x11008.
    Every element of y, scale and dW is compared with the reference's own RescalingIntQuant tree running on ATen on
    the same GPU; fp32 additionally with the reference on the CPU (the oracle proper) for y and scale."""
    rows, cols = 4096, 11008
    gen = torch.Generator().manual_seed(0)
    w_host = torch.randn(rows, cols, generator=gen)
    g_host = torch.randn(rows, cols, generator=torch.Generator().manual_seed(1))
    set_level(ref, "ops")
    from brevitas_b200.binding import uninstall
    uninstall()                                  # pure reference: Python STE backend on ATen
    # per-tensor + 16-bit: the scale is a 0-dim fp32 tensor next to 16-bit data; ATen-CUDA rounds such a scalar to 16
    # bits first, ATen-CPU (the reference's CI, the goldens) keeps it in fp32 -- the reference disagrees with itself
    # across devices there, so that combination is compared with the reference on the CPU only
    cuda_ref_valid = dtype == torch.float32 or per_channel
    if cuda_ref_valid:
        w_ref = torch.nn.Parameter(w_host.to(dtype).cuda())
        tq_ref = _weight_tree(per_channel, w_ref).cuda()
        y_ref, s_ref, _, _ = tq_ref(w_ref)
        y_ref.backward(g_host.to(dtype).cuda())
        gw_ref = w_ref.grad
    w_cpu = torch.nn.Parameter(w_host.to(dtype))
    torch.set_num_threads(os.cpu_count() or 1)
    y_cpu, s_cpu, _, _ = _weight_tree(per_channel, w_cpu)(w_cpu)
    if cuda_ref_valid:
        _assert_same_bits(y_ref.cpu(), y_cpu.detach(), "reference CUDA vs CPU y")
        _assert_same_bits(s_ref.cpu().reshape(-1), s_cpu.detach().reshape(-1), "reference CUDA vs CPU scale")
    else:
        y_cpu.backward(g_host.to(dtype))
        y_ref, s_ref, gw_ref = y_cpu.detach().cuda(), s_cpu.detach().cuda(), w_cpu.grad.cuda()
    set_level(ref, "fused")
    from brevitas_b200 import _kernels
    w = torch.nn.Parameter(w_host.to(dtype).cuda())
    tq = _weight_tree(per_channel, w).cuda()
    assert type(tq).__module__ == "brevitas_b200.core.quant"
    before = _kernels.launch_count
    y, s, _, _ = tq(w)
    y.backward(g_host.to(dtype).cuda())
    torch.cuda.synchronize()
    assert _kernels.launch_count - before <= (2 if per_channel else 5)
    _assert_same_bits(y, y_ref, "y")
    _assert_same_bits(s.reshape(-1), s_ref.reshape(-1), "scale")
    gw = w.grad
    same = _bits(gw) == _bits(gw_ref)
    n_diff = int((~same).sum())
    # only the arg-max entries carry a sum (Gs): at most one per row (per-channel) / the tied maxima (per-tensor)
    assert n_diff <= (rows if per_channel else 8), f"{n_diff} gradient elements differ"
    if n_diff:
        idx = (~same).nonzero()
        absw = w.detach().abs()
        amax = absw.amax(dim=1, keepdim=True) if per_channel else absw.amax()
        assert bool((absw[idx[:, 0], idx[:, 1]] == (amax[idx[:, 0], 0] if per_channel else amax)).all()), \
            "a differing gradient element is not an arg-max"
        # fp64-referenced bound for the arg-max entries (float32 only: 16-bit results are rounded again):
        # dW[argmax] = (g*s)/s + sign(w) * Gs / 127,  Gs = sum g*code - sum (g*s)*w/(s*s)   
        if dtype == torch.float32 and per_channel:
            s64 = s.detach().double().reshape(rows, 1)
            g64, w64 = g_host.double().cuda(), w_host.double().cuda()
            codes = torch.round(y.detach().double() / s64)      # the fp32 chain's own codes (an fp64 quotient rounds
            #                                                     ties differently)
            t_a, t_b = g64 * codes, (g64 * s64) * w64 / (s64 * s64)
            exact = (t_a - t_b).sum(dim=1) / 127.0
            bound = _fp64_sum_bound(t_a.abs() + t_b.abs(), cols) / 127.0
            r, c = idx[:, 0], idx[:, 1]
            want = (g64 * s64 / s64)[r, c] + torch.sign(w64[r, c]) * exact[r]
            lim = bound[r] + 4 * 2.0 ** -24 * want.abs()
            # the kernel must meet the bound; ATen's own reductions (the checker) get 4x: they are the less accurate side
            for name, t, k in (("kernel", gw, 1.0), ("reference on ATen", gw_ref, 4.0)):
                err = (t.double()[r, c] - want).abs()
                assert bool((err <= k * lim).all()), f"{name}: arg-max gradient off by {float((err / lim).max()):.2f}x the bound"
            print(f"arg-max entries: max err / bound = {float(((gw.double()[r, c] - want).abs() / lim).max()):.3f}")


@pytest.mark.parametrize("heavy_tail", [False, True])
def test_c3_whole_tensor_vs_reference(ref, heavy_tail):
    """BASELINE config 3: per-token dynamic int8 on [8,2048,4096] bf16 (heavy-tail variant exercises the
    clamp), module in train(): y, scale, running statistics and dx for EVERY element vs the reference on ATen."""
    B, T, C = 8, 2048, 4096
    gen = torch.Generator().manual_seed(0)
    x_host = torch.randn(B, T, C, generator=gen)
    if heavy_tail:
        x_host = x_host * (1 + 10 * torch.bernoulli(torch.full((B, T, C), 1e-3), generator=gen))
    x_host = x_host.to(torch.bfloat16)
    g_host = torch.randn(B, T, C, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)

    def build():
        from brevitas.core import function_wrapper as fw
        from brevitas.core.bit_width import BitWidthConst
        from brevitas.core.quant import IntQuant, RescalingIntQuant
        from brevitas.core.restrict_val import FloatRestrictValue
        from brevitas.core.scaling import IntScaling, RuntimeStatsScaling
        from brevitas.core.stats import AbsMax
        from brevitas.core.zero_point import ZeroZeroPoint
        return RescalingIntQuant(
            IntQuant(narrow_range=False, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp()),
            RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(), (B, T, 1), False,
                                0.1, 1e-10),
            IntScaling(True, False), ZeroZeroPoint(), BitWidthConst(8)).cuda().train()

    def run(tq):
        x = x_host.cuda().requires_grad_(True)
        outs = []
        for step in range(2):
            x.grad = None
            y, s, _, _ = tq(x)
            y.backward(g_host.cuda())
            outs.append((y.detach(), s.detach(), x.grad.clone(), tq.scaling_impl.runtime_stats.running_stats.clone()))
        return outs

    from brevitas_b200.binding import uninstall
    set_level(ref, "ops")
    uninstall()
    want = run(build())
    set_level(ref, "fused")
    tq = build()
    assert type(tq).__module__ == "brevitas_b200.core.quant"
    got = run(tq)
    for step, (a, b) in enumerate(zip(got, want)):
        _assert_same_bits(a[0], b[0], f"step {step} y")
        _assert_same_bits(a[1], b[1], f"step {step} scale")
        _assert_same_bits(a[3], b[3], f"step {step} running_stats")
        same = _bits(a[2]) == _bits(b[2])
        n_diff = int((~same).sum())
        assert n_diff <= B * T, f"step {step}: {n_diff} gradient elements differ (more than one per token)"
        # The differing entries must be arg-max entries (one per token): they carry sign(x) * Gs / 128 with
        # Gs = sum g*code - sum m*(g*s)*x/(s*s).  The bf16 reference rounds every product and both sums to
        # bf16 before they cancel, so its own value is only defined to ~sqrt(C) * 2^-9 * rms|g*code| / 128; the kernel
        # accumulates Gs in fp32.  Both are therefore compared with an fp64 evaluation of Gs from the same bf16 inputs.
        xd, gd = x_host.cuda(), g_host.cuda()
        ne = (~same).nonzero()
        xa = xd.float().abs()
        assert bool((xa[ne[:, 0], ne[:, 1], ne[:, 2]] == xa.amax(dim=2)[ne[:, 0], ne[:, 1]]).all()), "not an arg-max"
        sc = a[1]                                                   # bf16 [B,T,1]
        t3 = torch.round(xd / sc)
        m = ((t3 <= 127) & (t3 >= -128)).double()
        ew = (((gd * sc) / sc).float() * m.float())                 # element-wise part, per-op bf16 rounding like ATen
        s64, g64, x64 = sc.double(), gd.double(), xd.double()
        codes64 = torch.round(a[0].double() / s64)
        t_a, t_b = g64 * codes64, m * (g64 * s64) * x64 / (s64 * s64)
        gs64 = (t_a - t_b).sum(dim=2) / 128.0                        # [B,T]
        r0, r1, r2 = ne[:, 0], ne[:, 1], ne[:, 2]
        fix = torch.sign(x64[r0, r1, r2]) * gs64[r0, r1]
        want_e = ew.double()[r0, r1, r2] + fix
        mag = torch.maximum(want_e.abs(), torch.maximum(ew.double()[r0, r1, r2].abs(), fix.abs()))
        # noise floor of the 16-bit chain: every product in the two sums is rounded to bf16 (the kernel reuses the
        # bit-exact element-wise values, so it shares this part), the reference also rounds both sums to bf16
        rms = (t_a * t_a).mean(dim=2).sqrt()[r0, r1]
        big = torch.maximum(t_a.sum(dim=2).abs(), t_b.sum(dim=2).abs())[r0, r1]
        floor = (8 * np.sqrt(C) * 2.0 ** -9 * rms + 4 * 2.0 ** -8 * big) / 128.0 + 4 * 2.0 ** -8 * mag + 1e-3
        err_k = (a[2].double()[r0, r1, r2] - want_e).abs()
        err_r = (b[2].double()[r0, r1, r2] - want_e).abs()
        assert bool((err_k <= floor).all()), f"step {step}: kernel arg-max gradient {float((err_k / floor).max()):.2f}x the floor"
        assert bool((err_r <= floor).all()), f"step {step}: reference arg-max gradient {float((err_r / floor).max()):.2f}x the floor"
        # ... and the kernel (fp32 accumulation) is on average at least as close to the fp64 value as the reference
        assert float(err_k.mean()) <= 1.1 * float(err_r.mean()) + 1e-4, (float(err_k.mean()), float(err_r.mean()))
        print(f"step {step}: {n_diff} arg-max entries differ from the bf16 reference; vs fp64: kernel mean err "
              f"{float(err_k.mean()):.4f} (max {float((err_k / floor).max()):.2f}x floor), reference mean err "
              f"{float(err_r.mean()):.4f} (max {float((err_r / floor).max()):.2f}x floor)")
