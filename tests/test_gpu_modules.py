"""GPU (-m gpu): the reference-facing layer -- ``torch.ops.autograd_ste_ops.*`` and the ``tensor_quant`` modules of
``brevitas_b200.core`` (same constructors / forward contract as ``brevitas.core``) -- against the golden vectors of
the real reference, through autograd.  These read like the reference's own tests (tests/brevitas/core/*,
tests/brevitas/function/*) because the interface is the same.
"""
import numpy as np
import pytest
import torch

from golden_util import DTYPES, assert_bits_equal, case, load, ulp

pytestmark = pytest.mark.gpu

TDT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


@pytest.fixture(scope="module")
def B():
    import brevitas_b200
    return brevitas_b200


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(TDT[dtype]).cuda()


def host(t):
    return t.detach().float().cpu().numpy()


def close(got, ref, n, mag, dtype):
    tol = mag * (8.0 * np.sqrt(n) * 2.0 ** -24 + 4 * ulp(dtype)) + 1e-6     # see test_gpu_parity.close_sum
    both_nan = np.isnan(got) & np.isnan(ref)
    assert np.all(both_nan | (np.abs(got - ref) <= tol)), (got, ref, tol)


def build_weight_quant(B, w, per_channel, clamp_ste=True):
    """Same tree as tests/golden/make_golden.py::build_weight_quant, from brevitas_b200.core"""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, StatsFromParameterScaling
    from brevitas_b200.core.stats import AbsMax
    from brevitas_b200.core.zero_point import ZeroZeroPoint
    if per_channel:
        stats, view, concat, shape = AbsMax(1), fw.OverOutputChannelView(None), 1, (w.shape[0],) + (1,) * (w.dim() - 1)
    else:
        stats, view, concat, shape = AbsMax(None), fw.OverTensorView(), 0, ()
    return RescalingIntQuant(
        IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(),
                 tensor_clamp_impl=fw.TensorClampSte() if clamp_ste else fw.TensorClamp()),
        StatsFromParameterScaling(stats, view, concat, [w], FloatRestrictValue(), shape, False, 1e-10),
        IntScaling(True, True), ZeroZeroPoint(), BitWidthConst(8)).cuda()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tag", ["chan_lin", "chan_conv", "tensor_lin", "tensor_conv"])
def test_weight_quant_module(B, tag, dtype):
    from brevitas_b200 import _kernels
    c = case("weight_stats", f"weight_stats/{tag}/{dtype}/")
    w = torch.nn.Parameter(dev(c["w"], dtype))
    tq = build_weight_quant(B, w, tag.startswith("chan"))
    before = _kernels.launch_count
    y, scale, zp, bw = tq(w)
    assert _kernels.launch_count - before == 1, "weight quantization must be ONE fused kernel launch"
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(scale), c["scale"], "scale")
    assert scale.shape == c["scale"].shape and float(zp) == 0.0 and float(bw) == 8.0
    g, gs = dev(c["g"], dtype), torch.from_numpy(c["gs"]).to(scale.dtype).cuda()
    ((y * g).sum() + (scale * gs).sum()).backward()
    a = np.abs(c["w"].reshape(c["w"].shape[0], -1)) if tag.startswith("chan") else np.abs(c["w"]).reshape(1, -1)
    ismax = (a == a.max(axis=1, keepdims=True)).reshape(c["w"].shape)
    gw, ref = host(w.grad), c["gw_with_gscale"]
    assert_bits_equal(np.where(ismax, 0, gw), np.where(ismax, 0, ref), "gw off the arg-max")
    close(gw[ismax], ref[ismax], a.shape[1], np.nanmax(np.abs(ref[ismax])) + 1.0, dtype)
    # state-dict compatibility: no keys for a stats-scaled weight quantizer (reference: StatelessBuffer and
    # _ViewParameterWrapper drop themselves, core/utils.py:51-64, core/stats/view_wrapper.py:24-37)
    assert list(tq.state_dict().keys()) == []


@pytest.mark.parametrize("dtype", DTYPES)
def test_runtime_token_module(B, dtype):
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, RuntimeStatsScaling
    from brevitas_b200.core.stats import AbsMax
    from brevitas_b200.core.zero_point import ZeroZeroPoint
    c = case("runtime_token", f"runtime_token/{dtype}/")
    tq = RescalingIntQuant(
        IntQuant(narrow_range=False, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp()),
        RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(), (2, 9, 1), False,
                            0.1, 1e-10),
        IntScaling(True, False), ZeroZeroPoint(), BitWidthConst(8)).cuda()
    tq.train()
    for step in range(2):
        x = dev(c[f"x{step}"], dtype).requires_grad_(True)
        y, scale, zp, bw = tq(x)
        assert_bits_equal(host(y), c[f"y{step}"], "y")
        assert_bits_equal(host(scale), c[f"scale{step}"], "scale")
        assert_bits_equal(host(tq.scaling_impl.runtime_stats.running_stats), c[f"running{step}"], "running_stats")
        y.backward(dev(c[f"g{step}"], dtype))
        gx, ref = host(x.grad).reshape(18, 64), c[f"gx{step}"].reshape(18, 64)
        a = np.abs(c[f"x{step}"].reshape(18, 64))
        ismax = a == a.max(axis=1, keepdims=True)
        assert_bits_equal(np.where(ismax, 0, gx), np.where(ismax, 0, ref), "gx")
        close(gx[ismax], ref[ismax], 64, np.abs(ref[ismax]).max() + 1, dtype)
    tq.eval()
    y, scale, _, _ = tq(dev(c["x1"], dtype))
    assert_bits_equal(host(scale), c["scale_eval"], "eval scale")
    assert_bits_equal(host(y), c["y_eval"], "eval y")
    assert list(tq.state_dict().keys()) == ["scaling_impl.runtime_stats.running_stats"]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("qname", ["binary", "clamped"])
@pytest.mark.parametrize("sname", ["const", "param", "param_rows"])
def test_binary_modules(B, qname, sname, dtype):
    from brevitas_b200.core.quant import BinaryQuant, ClampedBinaryQuant
    from brevitas_b200.core.scaling import ConstScaling, ParameterScaling
    c = case("binary", f"binary/{qname}_{sname}/{dtype}/")
    simpl = {"const": lambda: ConstScaling(0.5), "param": lambda: ParameterScaling(0.5),
             "param_rows": lambda: ParameterScaling(torch.linspace(0.1, 0.9, 7).view(7, 1), (7, 1))}[sname]()
    simpl = simpl.to(TDT[dtype]).cuda()
    q = (BinaryQuant if qname == "binary" else ClampedBinaryQuant)(scaling_impl=simpl)
    x = dev(c["x"], dtype).requires_grad_(True)
    y, scale, zp, bw = q(x)
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(scale), c["scale"], "scale")
    assert float(bw) == 1.0 and float(zp) == 0.0
    y.backward(dev(c["g"], dtype))
    assert_bits_equal(host(x.grad), c["gx"], "gx")
    if "gvalue" in c:
        n = c["x"].size // c["gvalue"].size
        close(host(simpl.value.grad), c["gvalue"], n, np.abs(c["g"]).sum() / c["gvalue"].size + 1, dtype)


INT_CASES = {
    "s8n_round_ste_scalar": ("RoundSte", "TensorClampSte"), "s8_round_masked_scalar": ("RoundSte", "TensorClamp"),
    "u8_round_masked_scalar_zp": ("RoundSte", "TensorClamp"), "s4n_floor_ste_rows": ("FloorSte", "TensorClampSte"),
    "u4n_rtz_masked_chan": ("RoundToZeroSte", "TensorClamp"), "s8_dpu_masked_chan_zp": ("DPURoundSte", "TensorClamp"),
    "s8n_round_masked_token": ("RoundSte", "TensorClamp"),
}


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", sorted(INT_CASES))
@pytest.mark.parametrize("literal", [False, True])
def test_int_quant_module(B, name, dtype, literal):
    """IntQuant.forward(scale, zero_point, bit_width, x): fused kernel, and the literal op sequence on the 12 STE
    kernels (what the reference runs after `ops_ste.fn_prefix = torch`) -- both must reproduce the reference."""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import IntQuant
    c = case("int_quant", f"int_quant/{name}/{dtype}/")
    signed, narrow, bits, zp = [float(v) for v in c["meta"]]
    rimpl, cimpl = INT_CASES[name]
    iq = IntQuant(bool(narrow), bool(signed), getattr(fw, rimpl)(), getattr(fw, cimpl)())
    x = dev(c["x"], dtype).requires_grad_(True)
    s = dev(c["scale"], dtype).requires_grad_(True)
    zpt, bw = torch.tensor(zp, device="cuda"), torch.tensor(bits, device="cuda")
    if literal:
        y = (iq.to_int(s, zpt, bw, x) - zpt) * s
    else:
        y = iq(s, zpt, bw, x)
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(iq.to_int(s, zpt, bw, x)), c["codes"], "codes")
    y.backward(dev(c["g"], dtype))
    assert_bits_equal(host(x.grad), c["gx"], "gx")
    if np.isfinite(c["gscale"]).all():
        n = c["x"].size // max(1, c["gscale"].size)
        close(host(s.grad), c["gscale"], n, np.abs(c["gscale"]).max() * 4 + np.abs(c["g"]).sum() / max(1, c["gscale"].size) * 128, dtype)


def test_ste_ops_dispatcher(B):
    """torch.ops.autograd_ste_ops.* -- the reference's plugin namespace (csrc/autograd_ste_ops.cpp:258-271):
    forward delegates, backward is an exact pass-through (tests/brevitas/function/test_autograd_ste_ops.py:53-63)"""
    from brevitas_b200.ops import STE_OP_NAMES
    assert len(STE_OP_NAMES) == 12
    ops = torch.ops.autograd_ste_ops
    x = (torch.randn(1000, device="cuda") * 3).requires_grad_(True)
    g = torch.randn(1000, device="cuda")
    table = {"round_ste_impl": torch.round, "ceil_ste_impl": torch.ceil, "floor_ste_impl": torch.floor,
             "ternary_sign_ste_impl": torch.sign}
    for name, ref in table.items():
        y = getattr(ops, name)(x)
        assert torch.equal(y, ref(x.detach()))
        (gx,) = torch.autograd.grad(y, x, g)
        assert torch.equal(gx, g)
    y = ops.tensor_clamp_ste_impl(x, torch.tensor(-1.0, device="cuda"), torch.tensor(1.0, device="cuda"))
    assert torch.equal(y, x.detach().clamp(-1, 1))
    assert torch.equal(torch.autograd.grad(y, x, g)[0], g)
    y = ops.scalar_clamp_ste_impl(x, -0.5, 0.25)
    assert torch.equal(y, x.detach().clamp(-0.5, 0.25)) and torch.equal(torch.autograd.grad(y, x, g)[0], g)
    y = ops.scalar_clamp_min_ste_impl(x, 0.1)
    assert torch.equal(y, x.detach().clamp_min(0.1)) and torch.equal(torch.autograd.grad(y, x, g)[0], g)
    # abs_binary_sign_grad: subgradient 1 at 0 (tests/brevitas/function/test_autograd_ste_ops.py:175-190)
    z = torch.tensor([0.0, -0.0, -2.0, 3.0], device="cuda", requires_grad=True)
    y = ops.abs_binary_sign_grad_impl(z)
    y.backward(torch.ones(4, device="cuda"))
    assert z.grad.tolist() == [1.0, 1.0, -1.0, 1.0]
    # in-place clamp mutates and returns its input (Python-backend semantics)
    w = torch.randn(64, device="cuda")
    ref = w.clamp(-0.1, 0.1)
    r = ops.tensor_clamp_ste_impl_(w, torch.tensor(-0.1, device="cuda"), torch.tensor(0.1, device="cuda"))
    assert r.data_ptr() == w.data_ptr() and torch.equal(w, ref)
    with pytest.raises(RuntimeError):
        ops.round_ste_impl(torch.randn(4))          # CPU tensor: fails loudly


def test_param_from_runtime_stats(B):
    """Uint8ActPerTensorFloat's scaling: 3 collect steps (percentile + EMA) then a learned parameter"""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import ParameterFromRuntimeStatsScaling
    from brevitas_b200.core.stats import AbsPercentile
    d = load("param_from_stats")
    s = ParameterFromRuntimeStatsScaling(3, AbsPercentile(99.0, None), fw.OverTensorView(), (), FloatRestrictValue(),
                                         0.1, 1e-10).cuda()
    s.train()
    assert "value" not in s.state_dict() and "buffer" not in s.state_dict()
    for step in range(6):
        x = dev(d[f"param_from_stats/x{step}"], "f32").requires_grad_(True)
        t = s(x)
        t.backward()
        assert_bits_equal(host(t), d["param_from_stats/thresholds"][step], f"threshold step {step}")
        np.testing.assert_allclose(float(s.buffer), d["param_from_stats/buffers"][step], rtol=1e-6)
        assert_bits_equal(host(s.value.grad), d[f"param_from_stats/gvalue{step}"], "value grad")
        gx = host(x.grad) if x.grad is not None else np.zeros_like(d[f"param_from_stats/gx{step}"])
        assert_bits_equal(gx, d[f"param_from_stats/gx{step}"], "gradient through the percentile")
        s.value.grad = None
    np.testing.assert_allclose(host(s.value), d["param_from_stats/value"], rtol=1e-6)
    assert "value" in s.state_dict() and "buffer" not in s.state_dict()


def test_rescaling_docstring_kat(B):
    """RescalingIntQuant docstring (src/brevitas/core/quant/int.py:119-134)"""
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.scaling import ConstScaling, IntScaling
    from brevitas_b200.core.zero_point import ZeroZeroPoint
    d = load("kat")
    q = RescalingIntQuant(IntQuant(narrow_range=True, signed=True), ConstScaling(0.1),
                          IntScaling(signed=True, narrow_range=True), ZeroZeroPoint(), BitWidthConst(4)).cuda()
    inp = torch.Tensor([0.042, -0.053, 0.31, -0.44]).cuda()
    out, scale, zero_point, bit_width = q(inp)
    assert_bits_equal(host(out), d["kat/rescaling/y"])
    assert_bits_equal(host(scale), d["kat/rescaling/scale"])
    assert float(zero_point) == 0.0 and float(bit_width) == 4.0


def test_int_quant_roundtrip_every_integer(B):
    """tests/brevitas/core/test_int_quant.py:44-59: every representable integer x scale x zero-point round-trips"""
    from brevitas_b200.core.quant import IntQuant
    for signed in (True, False):
        for narrow in (True, False):
            for bits in (2, 3, 4, 8):
                iq = IntQuant(narrow_range=narrow, signed=signed)
                bw = torch.tensor(float(bits), device="cuda")
                lo, hi = int(iq.min_int(bw)), int(iq.max_int(bw))
                ints = torch.arange(lo, hi + 1, device="cuda").float()
                for scale in (0.001, 5.0):
                    for zp in (0.0, 1.0, -2.0):
                        s, z = torch.tensor(scale, device="cuda"), torch.tensor(zp, device="cuda")
                        x = (ints - z) * s
                        y = iq(s, z, bw, x)
                        assert torch.isclose(y, x).all()
                        assert torch.equal(iq.to_int(s, z, bw, x), ints)


def _literal_weight_quant(w, reduce_dims, bits=8):
    """the reference chain op by op in torch: AbsMax stats -> / int_threshold -> round/clamp with the STE"""
    thr = 2.0 ** (bits - 1) - 1
    # a TENSOR divisor like the reference's int_scaling_impl(bit_width): ATen divides IEEE-exactly by a device tensor but
    # multiplies by a rounded reciprocal when the divisor is a Python scalar
    scale = w.detach().abs().amax(dim=reduce_dims, keepdim=True).clamp_min(1e-10) / torch.tensor(thr, device=w.device)
    q = torch.clamp(torch.round(w.detach() / scale), -thr, thr) * scale
    return w + (q - w).detach(), scale


@pytest.mark.parametrize("per_channel", [False, True])
def test_conv_transpose2d_and_conv1d_layers(B, per_channel):
    """nn/quant_convtranspose.py, nn/quant_conv.py:22-113: output channels in dim 1 of a transposed-conv weight --
    per-channel statistics over the permuted view, scale shape [1, O, 1, 1]; values, scales and gradients against
    the literal torch composition"""
    from brevitas_b200 import nn as qnn
    from brevitas_b200.quant import Int8WeightPerChannelFloat, Int8WeightPerTensorFloat
    wq = Int8WeightPerChannelFloat if per_channel else Int8WeightPerTensorFloat
    torch.manual_seed(3)
    layer = qnn.QuantConvTranspose2d(6, 10, 3, stride=2, padding=1, bias=True, weight_quant=wq).cuda()
    qt = layer.quant_weight()
    ref_w, ref_s = _literal_weight_quant(layer.weight, (0, 2, 3) if per_channel else (0, 1, 2, 3))
    assert qt.scale.shape == ((1, 10, 1, 1) if per_channel else ())
    assert torch.equal(qt.scale.reshape(-1), ref_s.reshape(-1))
    assert torch.equal(qt.value, ref_w)
    x = torch.randn(4, 6, 9, 9, device="cuda", requires_grad=True)
    out = layer(x)
    ref_out = torch.nn.functional.conv_transpose2d(x, ref_w, layer.bias, 2, 1)
    # same weights bit for bit; the transposed convolution itself (cuDNN backward-data) may reorder its sums
    torch.testing.assert_close(out, ref_out, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(out)
    gw, = torch.autograd.grad(out, layer.weight, g, retain_graph=True)
    gw_ref, = torch.autograd.grad(ref_out, layer.weight, g)
    # the STE passes the weight gradient through everywhere except the arg-max elements, which also collect d(scale)
    argmax = (layer.weight.detach().abs() == layer.weight.detach().abs().amax(
        dim=(0, 2, 3) if per_channel else (0, 1, 2, 3), keepdim=True))
    torch.testing.assert_close(gw[~argmax], gw_ref[~argmax], rtol=1e-4, atol=1e-5)
    assert int(argmax.sum()) == (10 if per_channel else 1)

    c1 = qnn.QuantConv1d(5, 7, 4, padding=2, weight_quant=wq).cuda()
    q1 = c1.quant_weight()
    r1, s1 = _literal_weight_quant(c1.weight, (1, 2) if per_channel else (0, 1, 2))
    assert q1.scale.shape == ((7, 1, 1) if per_channel else ())
    assert torch.equal(q1.value, r1) and torch.equal(q1.scale.reshape(-1), s1.reshape(-1))
    x1 = torch.randn(3, 5, 33, device="cuda")
    torch.testing.assert_close(c1(x1), torch.nn.functional.conv1d(x1, r1, c1.bias, 1, 2), rtol=1e-4, atol=1e-5)
    eight = torch.tensor(8.0, device="cuda")
    assert int(c1.max_acc_bit_width(eight, eight)) == 21      # ceil(log2(255*255*4*7))
    assert int(layer.max_acc_bit_width(eight, eight)) == 22   # ceil(log2(255*255*(2*2)*10))


def test_sigmoid_tanh_and_conv_transpose1d_layers(B):
    """nn/quant_activation.py:32-65, nn/quant_convtranspose.py:22-111: activation then the quantizer with a learned scale
    (eval mode of a freshly built layer uses the collected buffer); values against the literal torch composition"""
    from brevitas_b200 import nn as qnn
    from brevitas_b200.quant import Int8WeightPerChannelFloat
    torch.manual_seed(7)
    x = torch.randn(4, 33, device="cuda") * 3
    for layer, act, signed in ((qnn.QuantSigmoid(return_quant_tensor=True), torch.sigmoid, False),
                               (qnn.QuantTanh(return_quant_tensor=True), torch.tanh, True)):
        layer = layer.cuda().train()
        q = layer(x)
        a = act(x)
        scale = q.scale
        lo, hi = (-128.0, 127.0) if signed else (0.0, 255.0)
        ref = torch.clamp(torch.round(a / scale), lo, hi) * scale
        assert torch.equal(q.value, ref) and q.signed is signed and int(q.bit_width) == 8
        # 99.999th percentile of |act(x)| over 132 values = their maximum, divided by the int threshold
        thr = 128.0 if signed else 255.0
        assert torch.allclose(scale, a.abs().max() / thr, rtol=1e-6)
    t = qnn.QuantConvTranspose1d(6, 10, 4, stride=2, weight_quant=Int8WeightPerChannelFloat).cuda()
    qt = t.quant_weight()
    ref_w, ref_s = _literal_weight_quant(t.weight, (0, 2))
    assert qt.scale.shape == (1, 10, 1)
    assert torch.equal(qt.scale.reshape(-1), ref_s.reshape(-1)) and torch.equal(qt.value, ref_w)
    xin = torch.randn(3, 6, 17, device="cuda")
    torch.testing.assert_close(t(xin), torch.nn.functional.conv_transpose1d(xin, ref_w, t.bias, 2), rtol=1e-4, atol=1e-5)
    eight = torch.tensor(8.0, device="cuda")
    assert int(t.max_acc_bit_width(eight, eight)) == 21      # ceil(log2(255*255*2*10))


@pytest.mark.parametrize("channels_last", [False, True])
def test_relu_folds_into_the_quantizer_while_collecting_statistics(B, channels_last):
    """QuantReLU with the default Uint8ActPerTensorFloat quantizer during its first `collect_stats_steps` steps: the
    threshold is the 99.999th percentile of relu(x) (core/scaling/standalone.py:230-244).  The ReLU-folded select +
    relu_int_quant pair must give what nn.ReLU followed by the quantizer gives: same outputs, scale, buffer, and the
    same input gradient -- including the ONE element the statistic's gradient reaches (handed to autograd sparse)."""
    from brevitas_b200 import _kernels
    from brevitas_b200.nn import QuantReLU
    torch.manual_seed(3)
    folded = QuantReLU(collect_stats_steps=3).cuda().train()
    plain = QuantReLU(collect_stats_steps=3).cuda().train()
    fq = plain.act_quant.fused_activation_quant_proxy
    for step in range(5):                                    # 3 collection steps, the switch-over, one learned step
        x = torch.randn(6, 16, 10, 10, generator=torch.Generator().manual_seed(step)).cuda() * (1 + step)
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        g = torch.randn(6, 16, 10, 10, generator=torch.Generator().manual_seed(50 + step)).cuda()
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        relu_calls = []
        h = folded.act_quant.fused_activation_quant_proxy.activation_impl.register_forward_hook(lambda *a: relu_calls.append(1))
        before = _kernels.launch_count
        ya = folded(xa)
        launches = _kernels.launch_count - before
        h.remove()
        assert not relu_calls, f"step {step}: the ReLU ran as a separate pass"
        yb_in = torch.relu(xb)                               # the unfolded composition, on the same modules' twin
        out = fq.tensor_quant(yb_in)
        yb = out[0]
        assert torch.equal(ya, yb), f"step {step}: outputs differ"
        sa = folded.quant_act_scale() if hasattr(folded, "quant_act_scale") else None
        ta, tb = folded.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl, fq.tensor_quant.scaling_impl
        assert torch.equal(ta.buffer, tb.buffer) and ta.counter == tb.counter, f"step {step}: statistics differ"
        (ya * g).sum().backward()
        (yb * g).sum().backward()
        assert not xa.grad.is_sparse and xa.grad.shape == x.shape
        assert torch.allclose(xa.grad, xb.grad, rtol=1e-5, atol=1e-6), \
            f"step {step}: max |d grad| {float((xa.grad - xb.grad).abs().max())}"
        same = (xa.grad.view(torch.int32) == xb.grad.contiguous().view(torch.int32) if not channels_last else
                xa.grad.contiguous().view(torch.int32) == xb.grad.contiguous().view(torch.int32))
        assert float(same.float().mean()) > 0.999
        va, vb = ta.value.grad, tb.value.grad
        assert (va is None) == (vb is None)
        if va is not None:
            assert torch.allclose(va, vb, rtol=1e-4, atol=1e-5)
            ta.value.grad = tb.value.grad = None
