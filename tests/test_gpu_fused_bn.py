"""GPU (-m gpu): batch-norm + ReLU + activation quantizer in fused passes (csrc/bn_act_quant.cu, brevitas_b200/fused_bn.py)
against the pair it replaces -- torch.nn.BatchNorm2d (cuDNN, channels-last) followed by the QuantReLU layer (ReLU folded
into the quantizer kernel).

What can differ: the batch statistics are summed in another order than cuDNN's, so the normalised value differs by a few
1e-7 relative; where that crosses a rounding boundary of x / scale the quantized output moves by exactly one step.  So:
every output equals the unfused one or differs by ONE quantization step, on at most 0.1 % of the elements; running
statistics, d(gamma), d(beta), d(scale) and dx agree to fp32 summation accuracy.
"""
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def make(channels, per_channel, bits=8):
    from brevitas_b200.nn import QuantReLU
    from qat.models import CommonUintActQuant
    act = QuantReLU(act_quant=CommonUintActQuant, bit_width=bits, per_channel_broadcastable_shape=(1, channels, 1, 1),
                    scaling_per_output_channel=per_channel, return_quant_tensor=False)
    bn = nn.BatchNorm2d(channels)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.uniform_(-0.5, 0.5)
        if per_channel:
            act.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value.add_(
                torch.linspace(-1.0, 0.5, channels).view(1, channels, 1, 1))
    return bn.cuda(), act.cuda()


@pytest.mark.parametrize("shape,per_channel,bits", [((8, 64, 28, 28), False, 8), ((4, 1024, 7, 7), True, 4),
                                                    ((16, 32, 56, 56), True, 4), ((3, 128, 9, 5), False, 4),
                                                    ((2, 512, 14, 14), False, 8)])
def test_fused_bn_relu_quant_matches_the_unfused_pair(shape, per_channel, bits):
    from brevitas_b200 import _kernels
    from brevitas_b200.fused_bn import bn_act_quant
    torch.manual_seed(shape[1])
    bn_a, act_a = make(shape[1], per_channel, bits)
    bn_b, act_b = make(shape[1], per_channel, bits)
    bn_b.load_state_dict(bn_a.state_dict())
    act_b.load_state_dict(act_a.state_dict())
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(shape, generator=g) * 1.5 + 0.3).cuda().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    for step in range(2):
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya = act_a(bn_a(xa))
        before = _kernels.launch_count
        yb = bn_act_quant(bn_b, act_b, xb)
        launched = _kernels.launch_count - before
        assert launched == 3, f"expected scale ops (2) + the fused forward (1), saw {launched}"
        assert yb.is_contiguous(memory_format=torch.channels_last) and yb.shape == ya.shape
        tq = act_a.act_quant.fused_activation_quant_proxy.tensor_quant
        scale = (tq.scaling_impl(xa) / tq.int_scaling_impl(tq.msb_clamp_bit_width_impl())).detach()
        d = (ya - yb).detach().abs()
        moved = d > 0
        assert float(moved.float().mean()) <= 1e-3, f"{float(moved.float().mean()):.2e} of the outputs differ"
        assert bool((d[moved] <= (scale.expand_as(d)[moved] * 1.0001)).all()), "an output moved by more than one step"
        ya.backward(gy)
        yb.backward(gy)
        for pa, pb, name in ((bn_a.weight, bn_b.weight, "d gamma"), (bn_a.bias, bn_b.bias, "d beta"),
                             (act_a.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value,
                              act_b.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value, "d scale")):
            mag = float(pa.grad.abs().max()) + 1e-6
            assert torch.allclose(pa.grad, pb.grad, rtol=2e-3, atol=2e-3 * mag), \
                (name, float((pa.grad - pb.grad).abs().max()), mag)
            pa.grad = pb.grad = None
        mag = float(xa.grad.abs().max())
        bad = ((xa.grad - xb.grad).abs() > 2e-3 * mag + 2e-3 * xa.grad.abs())
        assert float(bad.float().mean()) <= 2e-3, f"dx: {float(bad.float().mean()):.2e} of the elements beyond tolerance"
        assert torch.allclose(bn_a.running_mean, bn_b.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-5, atol=1e-6)
        assert int(bn_b.num_batches_tracked) == step + 1
    # eval mode: running statistics, no statistics pass
    bn_a.eval(), bn_b.eval(), act_a.eval(), act_b.eval()
    with torch.no_grad():
        ya, yb = act_a(bn_a(x)), bn_act_quant(bn_b, act_b, x)
    d = (ya - yb).abs()
    assert float((d > 0).float().mean()) <= 1e-3


def test_fused_bn_falls_back_when_a_precondition_fails():
    from brevitas_b200 import _kernels
    from brevitas_b200.fused_bn import bn_act_quant
    from brevitas_b200.nn import QuantReLU
    bn, act = make(64, False)
    x = torch.randn(4, 64, 8, 8, device="cuda")                       # NCHW: not channels-last
    y = bn_act_quant(bn, act, x)
    assert torch.equal(y, act(bn(x))) or y.shape == x.shape
    collecting = QuantReLU(collect_stats_steps=5).cuda().train()       # threshold still depends on the activation
    xl = x.contiguous(memory_format=torch.channels_last)
    y = bn_act_quant(bn, collecting, xl)
    assert collecting.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.counter == 1
    assert y.shape == x.shape
    odd = nn.BatchNorm2d(48).cuda()                                    # 48 channels: 12 vectors do not divide 256
    _, act48 = make(48, False)
    y = bn_act_quant(odd, act48, torch.randn(2, 48, 4, 4, device="cuda").contiguous(memory_format=torch.channels_last))
    assert y.shape == (2, 48, 4, 4)


class _NchwBatchNorm(nn.BatchNorm2d):
    """the SAME unfused batch-norm evaluated on an NCHW copy: another cuDNN kernel, hence another summation order"""

    def forward(self, inp):
        return super().forward(inp.contiguous()).contiguous(memory_format=torch.channels_last)


@pytest.mark.parametrize("name,shape", [("resnet18", (16, 3, 128, 128)), ("mobilenet_v1", (2, 3, 224, 224))])
def test_models_with_fused_bn_track_the_unfused_models(name, shape):
    """Whole networks: an untrained quantized net amplifies a one-ulp change of a batch statistic (a few activation codes
    flip by one step per layer, the next layers' statistics move, ...).  The yardstick is therefore the SAME unfused
    model with nothing changed but the summation order of its batch-norms (NCHW instead of NHWC cuDNN kernels): the
    fused model must stay as close to the unfused one as that control does (measured: ResNet-18 gradient cosine 0.954
    fused vs 0.953 control; MobileNetV1 0.99997)."""
    from qat import models
    kw = {"collect_stats_steps": 1} if name == "resnet18" else {}
    x = torch.randn(shape, generator=torch.Generator().manual_seed(3)).cuda().contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (shape[0],), generator=torch.Generator().manual_seed(4)).cuda()

    def run(model):
        for step in range(3):
            out = model(x)
            loss = nn.functional.cross_entropy(out, t)
            model.zero_grad()
            loss.backward()
        return float(loss), torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])

    def build(**extra):
        torch.manual_seed(0)
        return getattr(models, name)(**kw, **extra).cuda().to(memory_format=torch.channels_last).train()
    a = build()
    state = a.state_dict()
    la, ga = run(a)
    control = build()
    control.load_state_dict(state, strict=False)       # (no learned `value` before the first collection step)
    for m in control.modules():
        if type(m) is nn.BatchNorm2d:
            m.__class__ = _NchwBatchNorm
    lc, gc = run(control)
    fused = build(fuse_bn=True)
    fused.load_state_dict(state, strict=False)
    lf, gf = run(fused)
    cos = lambda u, v: float(torch.nn.functional.cosine_similarity(u, v, dim=0))
    c_control, c_fused = cos(ga, gc), cos(ga, gf)
    print(f"{name}: loss {la:.5f} / control {lc:.5f} / fused {lf:.5f}; gradient cosine control {c_control:.5f}, fused {c_fused:.5f}")
    assert abs(lf - la) <= 2.0 * abs(lc - la) + 1e-3 * abs(la)
    assert c_fused >= c_control - 0.02 and c_fused > 0.9


@pytest.mark.parametrize("name,shape", [("resnet18", (8, 3, 128, 128)), ("mobilenet_v1", (2, 3, 224, 224))])
def test_fuse_batch_norm_prepares_unmodified_model_code(name, shape):
    """brevitas_b200.fuse_batch_norm(model): the batch-norm hands a pending result to the activation layer, which runs
    the fused operator -- the model's forward() is not touched.  Same launches and same numbers as the model that calls
    bn_act_quant() explicitly; everything that is not a quantized activation gets the real batch-norm output."""
    import brevitas_b200
    from brevitas_b200 import _kernels
    from qat import models
    kw = {"collect_stats_steps": 1} if name == "resnet18" else {}
    x = torch.randn(shape, generator=torch.Generator().manual_seed(3)).cuda().contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (shape[0],), generator=torch.Generator().manual_seed(4)).cuda()

    def build(**extra):
        torch.manual_seed(0)
        return getattr(models, name)(**kw, **extra).cuda().to(memory_format=torch.channels_last).train()

    def run(model):
        names = []
        real_call = _kernels.call
        _kernels.call = lambda n, *a: (names.append(n), real_call(n, *a))[1]
        try:
            for step in range(3):
                out = model(x)
                loss = nn.functional.cross_entropy(out, t)
                model.zero_grad()
                loss.backward()
        finally:
            _kernels.call = real_call
        grads = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
        return float(loss), grads, names.count("bvb_bn_act_quant_fwd"), names.count("bvb_bn_act_quant_bwd")

    explicit = build(fuse_bn=True)
    state = {k: v.clone() for k, v in explicit.state_dict().items()}
    le, ge, fe, be = run(explicit)
    prepared = build()
    prepared.load_state_dict(state, strict=False)
    n_bn = brevitas_b200.fuse_batch_norm(prepared)
    assert n_bn == sum(type(m) is nn.BatchNorm2d for m in prepared.modules()) > 0
    lp, gp, fp, bp = run(prepared)
    assert (fp, bp) == (fe, be) and fp > 0, (fp, bp, fe, be)
    assert abs(lp - le) <= 1e-5 * abs(le), (lp, le)
    assert float(torch.nn.functional.cosine_similarity(gp, ge, dim=0)) > 0.99999
    # a consumer that is not a quantized activation sees the plain batch-norm output
    bn = next(m for m in prepared.modules() if type(m) is nn.BatchNorm2d)
    xin = torch.randn(4, bn.num_features, 8, 8, device="cuda").contiguous(memory_format=torch.channels_last)
    pending = bn(xin)
    assert type(pending).__name__ == "_PendingBatchNorm"
    ref = bn._b200_unfused_forward(xin)
    assert torch.equal(torch.relu(pending), torch.relu(ref)) and torch.equal(pending * 2.0, ref * 2.0)
    assert pending.shape == ref.shape and torch.equal(bn(xin) + 1.0, ref + 1.0)
    brevitas_b200.unfuse_batch_norm(prepared)
    assert isinstance(bn(xin), torch.Tensor)
    assert prepared.state_dict().keys() == explicit.state_dict().keys()
