"""GPU (-m gpu): the QAT workloads end to end on the fused kernels -- bnn_pynq TFC fwd+bwd against the CPU oracle
model (same weights, same batch), a ResNet-18 QAT step through statistics collection into the learned-scale phase,
MobileNetV1 with learned per-channel activation scales, and the per-token dynamic quantizer."""
import numpy as np
import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0


def test_tfc_step_matches_cpu_oracle_model():
    """config 1 of BASELINE.json: TFC 2W2A, batch 256, synthetic MNIST-shaped input, SqrHinge loss"""
    import brevitas_b200  # noqa: F401
    from oracle import ref_models as R
    from qat import models
    torch.manual_seed(0)
    m = models.tfc(2, 2, 2)
    _no_dropout(m)
    weights = [p.detach().clone() for n, p in m.named_parameters() if n.endswith("weight") and p.dim() == 2]
    assert len(weights) == 4
    bns = [mod for mod in m.features if isinstance(mod, nn.BatchNorm1d)]
    tn = m.features[-1]
    g = torch.Generator().manual_seed(1)
    x = torch.rand(256, 1, 28, 28, generator=g)
    y = torch.full((256, 10), -1.0)
    y.scatter_(1, torch.randint(0, 10, (256, 1), generator=g), 1.0)
    # CPU oracle model
    wr = [w.clone().requires_grad_(True) for w in weights]
    bnp = [(b.weight.detach().clone().requires_grad_(True), b.bias.detach().clone().requires_grad_(True),
            b.running_mean.clone(), b.running_var.clone()) for b in bns]
    tnp = (tn.weight.detach().clone().requires_grad_(True), tn.bias.detach().clone().requires_grad_(True))
    out_ref = R.tfc_forward(x, wr, bnp, tnp)
    loss_ref = R.sqr_hinge(out_ref, y)
    loss_ref.backward()
    # B200 model
    m = m.cuda().train()
    out = m(x.cuda())
    loss = models.SqrHingeLoss()(out, y.cuda())
    loss.backward()
    # fp32 matmuls differ in summation order between cuBLAS and the CPU; a code flip at a rounding boundary
    # changes single activations by one step, so compare in aggregate
    assert abs(float(loss) - float(loss_ref)) < 2e-3 * max(1.0, abs(float(loss_ref)))
    diff = (out.detach().cpu() - out_ref.detach()).abs()
    assert float(diff.mean()) < 5e-3 and float((diff > 0.1).float().mean()) < 0.02
    gw = [p.grad.detach().cpu() for n, p in m.named_parameters() if n.endswith("weight") and p.dim() == 2]
    for a, b in zip(gw, wr):
        denom = b.grad.abs().mean() + 1e-12
        assert float((a - b.grad).abs().mean() / denom) < 0.05
    # quantized weights are exactly the oracle's: codes {-1, 0, 1} (2-bit narrow, scale 1)
    lin = [mod for mod in m.features if hasattr(mod, "quant_weight")]
    qw = lin[0].quant_weight().value.detach().cpu()
    assert torch.equal(qw, R._const_quant(weights[0], 2, True))
    assert set(np.unique(qw.numpy()).tolist()) <= {-1.0, 0.0, 1.0}


def test_resnet18_qat_steps_through_collection_into_learned_scales():
    from brevitas_b200 import _kernels as K
    from qat import models
    from qat.train import WORKLOADS, make_batch, make_optimizer, train_step
    torch.manual_seed(0)
    m = models.resnet18(num_classes=100, collect_stats_steps=2).cuda().train()
    opt = make_optimizer(m, WORKLOADS["resnet18"])
    spec = dict(WORKLOADS["resnet18"], classes=100)
    x, y = make_batch(spec, 8, torch.device("cuda"), 0)
    losses = []
    for step in range(5):
        before = K.launch_count
        losses.append(float(train_step(m, m, x, y, nn.CrossEntropyLoss(), opt)))
        launches = K.launch_count - before
    assert all(np.isfinite(losses))
    relu = m.relu.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl
    assert relu.counter == 3 and relu.value.grad is not None and float(relu.value.grad.abs().sum()) > 0
    assert "relu.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value" in m.state_dict()
    # steady state: per quantizer ONE fused forward launch and ONE fused backward launch (+ the tiny scale ops of the
    # learned activation scales): 21 weight quantizers + 17 activation quantizers
    assert launches <= 2 * (21 + 17) + 3 * 17 + 8, launches
    # a weight quantizer's output is on its 8-bit grid
    w = m.layer1[0].conv1
    q = w.quant_weight()
    codes = (q.value / q.scale).round()
    assert float((q.value - codes * q.scale).abs().max()) < 1e-6 and float(codes.abs().max()) <= 127


def test_mobilenet_v1_step_and_learned_per_channel_scales():
    from qat import models
    from qat.train import WORKLOADS, make_batch, make_optimizer, train_step
    torch.manual_seed(0)
    m = models.MobileNetV1(bit_width=4, num_classes=10).cuda().train()
    opt = make_optimizer(m, WORKLOADS["mobilenet_v1"])
    x, y = make_batch(dict(WORKLOADS["mobilenet_v1"], classes=10), 4, torch.device("cuda"), 0)
    for _ in range(2):
        loss = train_step(m, m, x, y, nn.CrossEntropyLoss(), opt)
    assert np.isfinite(float(loss))
    s = m.features.init_block.activation.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl
    assert s.value.shape == (1, 32, 1, 1) and s.value.grad is not None
    a = m.features.init_block.activation(m.features.init_block.bn(m.features.init_block.conv(x)))
    assert a.is_not_none and a.signed is False and float(a.bit_width) == 4.0      # QuantTensor, as in the reference
    a = a.value
    thr = 2 ** s.value.detach()
    codes = a / (thr / 15.0)
    assert float((codes - codes.round()).abs().max()) < 1e-3 and float(codes.max()) <= 15.001 and float(codes.min()) >= 0


def test_avg_pool_trunc_and_int_bias_wiring():
    """QuantReLU(return_quant_tensor) -> QuantAvgPool2d(trunc) -> QuantLinear(bias_quant=IntBias): the tail of the
    reference's MobileNetV1 (mobilenetv1.py:153-159) against the formulas of nn/quant_avg_pool.py:55-73,
    nn/quant_layer.py:302-345 and nn/quant_linear.py:68-73 written out with plain torch ops."""
    import math
    import torch.nn.functional as F
    from brevitas_b200.nn import QuantAvgPool2d, QuantLinear, QuantReLU
    from brevitas_b200.quant import IntBias
    from qat.models import CommonIntWeightPerTensorQuant, CommonUintActQuant
    torch.manual_seed(3)
    act = QuantReLU(act_quant=CommonUintActQuant, bit_width=4, per_channel_broadcastable_shape=(1, 8, 1, 1),
                    scaling_per_output_channel=False, return_quant_tensor=True).cuda()
    pool = QuantAvgPool2d(kernel_size=7, stride=1, bit_width=4).cuda()
    fc = QuantLinear(8, 5, bias=True, bias_quant=IntBias, weight_quant=CommonIntWeightPerTensorQuant,
                     weight_bit_width=4).cuda()
    with torch.no_grad():
        fc.bias.copy_(torch.randn(5) * 0.7)
    x = (torch.randn(3, 8, 7, 7, device="cuda") * 3).requires_grad_(True)
    q = act(x)
    s_in = q.scale
    assert float(q.bit_width) == 4.0 and float(q.zero_point) == 0.0
    p = pool(q)
    # reference formulas: sum of integers -> 10-bit accumulator (ceil(log2(15 * 49))) -> drop 6 LSBs with floor
    acc_bits = math.ceil(math.log2(15 * 49))
    summed = F.avg_pool2d(q.value, 7, 1) * 49
    expect = torch.floor(torch.round(summed / s_in) / 2.0 ** (acc_bits - 4)) * s_in
    assert torch.equal(p.value, expect) and float(p.bit_width) == 4.0 and p.scale is s_in
    out = fc(p.view(3, -1))
    wq = fc.quant_weight()
    out_scale = wq.scale.view(1, -1) * s_in.view(1, -1)
    acc_bits_fc = math.ceil(math.log2(15 * 7 * 8))                           # max_uint_value(4 bits, narrow) = 7
    lo, hi = -2.0 ** (acc_bits_fc - 1), 2.0 ** (acc_bits_fc - 1) - 1
    bias_q = torch.clamp(torch.round(fc.bias / out_scale.view(-1)), lo, hi) * out_scale.view(-1)
    assert torch.equal(out, F.linear(p.value.view(3, -1), wq.value, bias_q))
    out.sum().backward()
    assert x.grad is not None and fc.bias.grad is not None and torch.isfinite(x.grad).all()


def test_per_token_dynamic_quantizer_full_size():
    """config 3: [8, 2048, 4096] bf16 activations, per-token int8, training mode (dynamic statistics)"""
    from brevitas_b200.quant import Int8ActPerTokenDynamic
    tq = Int8ActPerTokenDynamic.tensor_quant_for(8, 2048).cuda().train()
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(8, 2048, 4096, device="cuda", generator=g).to(torch.bfloat16).requires_grad_(True)
    y, scale, zp, bw = tq(x)
    assert y.dtype == torch.bfloat16 and scale.shape == (8, 2048, 1) and scale.dtype == torch.bfloat16
    amax = x.detach().abs().amax(dim=2, keepdim=True)
    assert torch.equal(scale, (amax / 128.0))                       # bf16 division, bit-exact
    assert torch.equal(tq.scaling_impl.runtime_stats.running_stats, amax.float())
    codes = (y.float() / scale.float())
    assert float(codes.abs().max()) <= 128.5
    y.backward(torch.ones_like(y))
    assert x.grad.shape == x.shape and torch.isfinite(x.grad.float()).all()


def test_collection_phase_as_a_cuda_graph_matches_eager_steps():
    """The statistics-collection phase captured as ONE CUDA graph (qat/train.py::GraphedStep(collecting=True)): host-side
    step counters advance at every replay, the running statistics follow the eagerly launched twin step for step, and the
    graph refuses to run the step that ends the phase.  Learning rate 0 and two alternating batches: the weights stay put,
    so the running averages depend on nothing but WHICH steps ran -- one step too many or too few moves them by percents."""
    from qat import models
    from qat.train import WORKLOADS, GraphedStep, collecting_modules, make_batch, make_optimizer, train_step
    spec = dict(WORKLOADS["resnet18"], classes=100, lr=0.0)
    dev = torch.device("cuda")
    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):               # the whole life of a captured model runs on a side stream (see run())
        twins = []
        for _ in range(2):
            torch.manual_seed(0)
            m = models.resnet18(num_classes=100, collect_stats_steps=9).cuda().train().to(memory_format=torch.channels_last)
            twins.append((m, make_optimizer(m, spec, capturable=True)))
        x, y = make_batch(spec, 8, dev, 0)
        x = x.contiguous(memory_format=torch.channels_last)
        batches = [(x, y), (2.0 * x, y)]
        loss_fn = nn.CrossEntropyLoss()
        (eager, eopt), (graphed, gopt) = twins
        step = 0
        for _ in range(2):
            train_step(eager, eager, *batches[step % 2], loss_fn, eopt)
            train_step(graphed, graphed, *batches[step % 2], loss_fn, gopt)
            step += 1
        with pytest.raises(RuntimeError, match="still collecting"):
            GraphedStep(graphed, loss_fn, gopt, *batches[0])                # steady-state capture refuses
        g = GraphedStep(graphed, loss_fn, gopt, *batches[0], collecting=True)   # 3 eager warm-up steps on batch 0 inside
        assert len(g.collectors) == len(collecting_modules(graphed)) > 10
        for _ in range(3):
            train_step(eager, eager, *batches[0], loss_fn, eopt)
        step = 5
        for _ in range(3):
            g(*batches[step % 2])                                           # counters 6, 7, 8
            train_step(eager, eager, *batches[step % 2], loss_fn, eopt)
            step += 1
        torch.cuda.synchronize()
        with pytest.raises(RuntimeError, match="collection phase ends"):
            g()
        for (n1, a), (_n2, b) in zip(collecting_modules(eager), collecting_modules(graphed)):
            assert int(a.counter) == int(b.counter) == 8, (n1, a.counter, b.counter)
            assert torch.allclose(a.buffer, b.buffer, rtol=1e-5, atol=0), (n1, float(a.buffer), float(b.buffer))
    torch.cuda.current_stream().wait_stream(side)
