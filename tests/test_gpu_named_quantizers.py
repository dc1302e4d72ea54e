"""GPU (-m gpu): EVERY named quantizer the reference exports from ``brevitas.quant`` (32 classes: weight / activation / bias /
truncation; float, fixed-point, decoupled, shifted, binary, ternary), instantiated by the reference's own injector inside its
own ``brevitas.nn`` layers, after ``brevitas_b200.install()`` -- against the same layer of the PURE reference (Python STE
backend on ATen, same GPU): quantized values, scales, zero-points and bit-widths bit-identical in fp32 over training steps
(statistics collection included) and in eval mode; input gradients to summation accuracy."""
import pytest
import torch

from ref_util import reference_src

pytestmark = pytest.mark.gpu

WEIGHT = ["Int8WeightPerTensorFloat", "Int8WeightPerChannelFloat", "Int8WeightPerTensorFixedPoint",
          "Int4WeightPerTensorFloatDecoupled", "Int8WeightPerChannelFloatDecoupled", "ShiftedUint8WeightPerTensorFloat",
          "ShiftedUint8WeightPerChannelFloat", "SignedBinaryWeightPerTensorConst", "SignedTernaryWeightPerTensorConst"]
ACT = ["Int8ActPerTensorFloat", "Uint8ActPerTensorFloat", "Int8ActPerTensorFixedPoint", "Uint8ActPerTensorFixedPoint",
       "Int8ActPerTensorFloatMinMaxInit", "Uint8ActPerTensorFloatMaxInit", "Uint8ActPerTensorFixedPointMaxInit",
       "ShiftedUint8ActPerTensorFloat", "SignedBinaryActPerTensorConst", "SignedTernaryActPerTensorConst"]
BIAS = ["Int8Bias", "Int16Bias", "Int24Bias", "Int32Bias", "IntBias", "Int8BiasPerTensorFloatInternalScaling",
        "Int8BiasPerTensorFixedPointInternalScaling"]


@pytest.fixture(scope="module")
def ref():
    src = reference_src()
    if src is None:
        pytest.skip("reference not available (oracle/make_ref.py)")
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    yield src
    uninstall()


def bound(src, fused):
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    uninstall()
    if fused:
        brevitas_b200.install(src, fuse=True)
    else:
        brevitas_b200.install(src, fuse=False)
        uninstall()                              # the pure reference, imported through the same path
    import brevitas.nn as qnn
    import brevitas.quant as Q
    return qnn, Q


def same(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, what
        return
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    eq = (a.contiguous().view(torch.int32) == b.contiguous().view(torch.int32)) | (torch.isnan(a) & torch.isnan(b))
    assert bool(eq.all()), f"{what}: {int((~eq).sum())} of {a.numel()} elements differ"


def qt_fields(qt):
    return [getattr(qt, f, None) for f in ("value", "scale", "zero_point", "bit_width")]


@pytest.mark.parametrize("name", WEIGHT)
def test_weight_quantizer(ref, name):
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        kw = {"weight_scaling_const": 0.1} if "Const" in name and "Ternary" in name else {}
        torch.manual_seed(0)
        layer = qnn.QuantConv2d(6, 8, 3, bias=False, weight_quant=getattr(Q, name), **kw)
        layer = layer.cuda().train()
        qw = layer.quant_weight()
        g = torch.randn(qw.value.shape, generator=torch.Generator().manual_seed(1)).cuda()
        (qw.value * g).sum().backward()
        results.append((qt_fields(qw), layer.weight.grad.clone(), type(layer.weight_quant.tensor_quant).__module__,
                        layer.weight.detach().clone()))
    (fr, gr, mod_r, wr), (ff, gf, mod_f, wf) = results
    assert mod_r.startswith("brevitas.core") and mod_f.startswith("brevitas_b200.core")
    for f, a, b in zip(("value", "scale", "zero_point", "bit_width"), ff, fr):
        same(a, b, f"{name}.{f}")
    same(wf, wr, f"{name}: weight after the call (in-place clamp quantizers)")
    scale = float(gr.abs().max()) + 1e-12
    assert torch.allclose(gf, gr, rtol=1e-4, atol=1e-5 * scale), float((gf - gr).abs().max())


@pytest.mark.parametrize("name", ACT)
def test_activation_quantizer(ref, name):
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        torch.manual_seed(0)
        layer_cls = qnn.QuantReLU if name.startswith(("Uint", "ShiftedUint")) else qnn.QuantIdentity
        kw = {"collect_stats_steps": 2} if "MaxInit" not in name and "MinMaxInit" not in name and "Const" not in name else {}
        if "MaxInit" in name or "MinMaxInit" in name:
            kw.update(max_val=3.0, min_val=-3.0 if "MinMaxInit" in name else 0.0)
        if "Const" in name:
            kw.update(scaling_const=0.5) if "Ternary" in name else None
        layer = layer_cls(act_quant=getattr(Q, name), return_quant_tensor=True, **kw).cuda().train()
        out = []
        for step in range(4):
            x = (torch.randn(4, 6, 5, 5, generator=torch.Generator().manual_seed(10 + step)) * (1 + step)).cuda()
            x.requires_grad_(True)
            qt = layer(x)
            g = torch.randn(x.shape, generator=torch.Generator().manual_seed(20 + step)).cuda()
            (qt.value * g).sum().backward()
            out.append((qt_fields(qt), x.grad.clone()))
        layer.eval()
        with torch.no_grad():
            out.append((qt_fields(layer(x.detach())), None))
        results.append(out)
    for step, ((fr, gr), (ff, gf)) in enumerate(zip(*results)):
        for f, a, b in zip(("value", "scale", "zero_point", "bit_width"), ff, fr):
            same(a, b, f"{name} step {step} {f}")
        if gr is not None:
            scale = float(gr.abs().max()) + 1e-12
            assert torch.allclose(gf, gr, rtol=1e-4, atol=1e-5 * scale), (step, float((gf - gr).abs().max()))


@pytest.mark.parametrize("name", BIAS)
def test_bias_quantizer(ref, name):
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        torch.manual_seed(0)
        layer = qnn.QuantLinear(12, 8, True, bias_quant=getattr(Q, name), input_quant=Q.Int8ActPerTensorFloat,
                                return_quant_tensor=True).cuda().train()
        x = torch.randn(5, 12, generator=torch.Generator().manual_seed(3)).cuda()
        out = []
        for step in range(2):
            qt = layer(x)
            layer.zero_grad()
            qt.value.square().sum().backward()
            out.append((qt_fields(qt), layer.bias.grad.clone()))
        results.append(out)
    for step, ((fr, gr), (ff, gf)) in enumerate(zip(*results)):
        for f, a, b in zip(("value", "scale", "zero_point", "bit_width"), ff, fr):
            same(a, b, f"{name} step {step} output {f}")
        assert torch.allclose(gf, gr, rtol=1e-4, atol=1e-6), float((gf - gr).abs().max())


def test_trunc_quantizer(ref):
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        torch.manual_seed(0)
        act = qnn.QuantReLU(act_quant=Q.Uint8ActPerTensorFloatMaxInit, max_val=6.0, return_quant_tensor=True).cuda().train()
        pool = qnn.QuantAvgPool2d(kernel_size=2, trunc_quant=Q.TruncTo8bit, return_quant_tensor=True).cuda().train()
        x = (torch.rand(2, 4, 6, 6, generator=torch.Generator().manual_seed(5)) * 5).cuda()
        results.append(qt_fields(pool(act(x))))
    for f, a, b in zip(("value", "scale", "zero_point", "bit_width"), results[1], results[0]):
        same(a, b, f"TruncTo8bit {f}")


@pytest.mark.parametrize("kind", ["weight_per_channel_int8", "act_uint4", "bias_int32"])
def test_quant_tensor_integer_export(ref, kind):
    """QuantTensor.int() of the reference after install(): same integers and dtype as the pure reference"""
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        torch.manual_seed(0)
        x = torch.randn(4, 12, generator=torch.Generator().manual_seed(2)).cuda()
        if kind == "weight_per_channel_int8":
            qt = qnn.QuantLinear(12, 8, False, weight_quant=Q.Int8WeightPerChannelFloat).cuda().quant_weight()
        elif kind == "act_uint4":
            act = qnn.QuantReLU(act_quant=Q.Uint8ActPerTensorFloatMaxInit, max_val=3.0, bit_width=4,
                                return_quant_tensor=True).cuda()
            qt = act(x)
        else:
            lin = qnn.QuantLinear(12, 8, True, bias_quant=Q.Int32Bias, input_quant=Q.Int8ActPerTensorFloat,
                                  return_quant_tensor=True).cuda()
            lin(x)
            lin.eval()
            lin.cache_inference_quant_bias = True
            lin(x)
            qt = lin.quant_bias()
        results.append((qt.int(), qt.int(float_datatype=True)))
    (ir, fr), (if_, ff) = results
    assert ir.dtype == if_.dtype and torch.equal(ir, if_), (ir.dtype, if_.dtype)
    assert torch.equal(fr, ff)


@pytest.mark.parametrize("fuse", [False, True])
@pytest.mark.parametrize("name", ["Int4WeightPerTensorFloatDecoupled", "Int8WeightPerChannelFloatDecoupled"])
def test_construction_time_scale_init_from_host_weights(ref, name, fuse):
    """quant/solver/parameter.py:39-45: the learned scale is initialised from the weight statistic while the layer is
    CONSTRUCTED, weights still on the host.  ops.parameter_init_on_host stages them to the GPU for that one call: the state
    dict equals the pure reference's (computed by ATen on the CPU) bit for bit; a forward on host tensors still raises."""
    import brevitas_b200
    from brevitas_b200.binding import uninstall
    uninstall()
    brevitas_b200.install(ref, fuse=fuse)
    import brevitas.nn as qnn
    import brevitas.quant as Q
    torch.manual_seed(0)
    ours = qnn.QuantLinear(16, 8, False, weight_quant=getattr(Q, name))
    assert all(not v.is_cuda for v in ours.state_dict().values())
    with pytest.raises(RuntimeError, match="CPU tensors are not supported"):
        ours(torch.randn(2, 16))
    with pytest.raises(RuntimeError, match="CPU tensors are not supported"):
        torch.ops.brevitas_b200.absmax_tensor(torch.randn(4))                   # outside the scope: raises
    uninstall()
    torch.manual_seed(0)
    pure = qnn.QuantLinear(16, 8, False, weight_quant=getattr(Q, name))
    assert set(ours.state_dict()) == set(pure.state_dict())
    for k, v in pure.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), (name, k)
