"""CPU: host logic of the QAT workloads -- quantizer resolution, module trees / state-dict keys, float (bit_width
None) mode equal to the plain layers, and the N>1 data-parallel harness under gloo with world_size 2."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


def test_named_quantizers_resolve_like_the_reference():
    """the trees the injector builds for the named quantizers"""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import BinaryQuant, ClampedBinaryQuant, RescalingIntQuant
    from brevitas_b200.core.scaling import (ConstScaling, ParameterFromRuntimeStatsScaling, ParameterScaling,
                                            StatsFromParameterScaling)
    from brevitas_b200.core.stats import AbsMax, AbsPercentile
    from brevitas_b200.quant import (Int8ActPerTensorFloat, Int8WeightPerChannelFloat, Int8WeightPerTensorFloat,
                                     Uint8ActPerTensorFloat)
    from qat.models import CommonActQuant, CommonUintActQuant, CommonWeightQuant
    w = nn.Parameter(torch.randn(8, 3, 3, 3))
    tq = Int8WeightPerChannelFloat.tensor_quant(w)
    assert isinstance(tq, RescalingIntQuant) and type(tq.int_quant.tensor_clamp_impl) is fw.TensorClampSte
    assert tq.int_quant.narrow_range and tq.int_quant.signed
    s = tq.scaling_impl
    assert isinstance(s, StatsFromParameterScaling) and type(s.parameter_list_stats.stats.stats_impl) is AbsMax
    assert s.parameter_list_stats.stats.stats_output_shape == (8, 1, 1, 1)
    assert s.fused_stats_plan(w).geom == ("rows", 8, 27)
    tq = Int8WeightPerTensorFloat.tensor_quant(w)
    assert tq.scaling_impl.fused_stats_plan(w).geom == ("tensor", 1, 216)
    tq = Uint8ActPerTensorFloat.tensor_quant()
    assert isinstance(tq.scaling_impl, ParameterFromRuntimeStatsScaling) and tq.scaling_impl.collect_stats_steps == 300
    assert type(tq.scaling_impl.stats.stats_impl) is AbsPercentile and not tq.int_quant.signed
    assert type(tq.int_quant.tensor_clamp_impl) is fw.TensorClamp          # activations: masked clamp gradient
    assert tq._host_config(torch.float32)[:3] == (0.0, 0.0, 255.0)
    assert Int8ActPerTensorFloat.tensor_quant()._host_config(torch.float32)[1:3] == (-128.0, 127.0)
    # bnn_pynq: 2 bit -> codes {-1, 0, 1} with scale 1; 1 bit -> binary quantizers (solver/weight.py:30, act.py:58)
    tq = CommonWeightQuant.let(bit_width=2).tensor_quant(nn.Parameter(torch.randn(4, 4)))
    assert isinstance(tq.scaling_impl, ConstScaling) and tq._host_config(torch.float32)[1:3] == (-1.0, 1.0)
    assert isinstance(CommonWeightQuant.let(bit_width=1).tensor_quant(nn.Parameter(torch.randn(4, 4))), BinaryQuant)
    assert isinstance(CommonActQuant.let(bit_width=1).tensor_quant(), ClampedBinaryQuant)
    assert CommonWeightQuant.let(bit_width=None).tensor_quant(nn.Parameter(torch.randn(4, 4))) is None
    # imagenet: learned LOG_FP scale initialised at 6.0, per-channel variant
    tq = CommonUintActQuant.let(bit_width=4, scaling_per_output_channel=True,
                                per_channel_broadcastable_shape=(1, 16, 1, 1)).tensor_quant()
    assert isinstance(tq.scaling_impl, ParameterScaling) and tq.scaling_impl.value.shape == (1, 16, 1, 1)
    assert torch.allclose(2 ** tq.scaling_impl.value, torch.full((1, 16, 1, 1), 6.0))


def test_model_state_dict_keys_follow_the_reference_layout():
    from qat.models import mobilenet_v1, resnet18, tfc
    keys = list(tfc().state_dict().keys())
    assert "features.2.weight" in keys and not any("tensor_quant" in k for k in keys)     # CONST scaling: stateless
    m = mobilenet_v1()
    keys = list(m.state_dict().keys())
    assert "features.init_block.activation.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value" in keys
    r = resnet18()
    assert not any("scaling_impl.value" in k for k in r.state_dict())      # no value before the first collect step
    names = [n for n, _ in r.named_parameters()]
    assert "relu.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value" in names
    assert sum(p.numel() for p in r.parameters()) == 11689512 + 17


def test_float_mode_equals_plain_layers():
    """quantizers with bit_width None are disabled (QuantType.FP): the layers reduce to their float parents"""
    from brevitas_b200.nn import QuantConv2d, QuantIdentity, QuantLinear
    from qat.models import CommonActQuant, CommonWeightQuant
    torch.manual_seed(0)
    lin = QuantLinear(12, 5, weight_quant=CommonWeightQuant, weight_bit_width=None)
    x = torch.randn(3, 12)
    assert torch.equal(lin(x), nn.functional.linear(x, lin.weight, lin.bias))
    conv = QuantConv2d(2, 4, 3, padding=1, weight_quant=None)
    xi = torch.randn(1, 2, 5, 5)
    assert torch.equal(conv(xi), nn.functional.conv2d(xi, conv.weight, conv.bias, padding=1))
    assert torch.equal(QuantIdentity(act_quant=CommonActQuant, bit_width=None)(x), x)
    with pytest.raises(RuntimeError):
        QuantLinear(12, 5)(x)            # enabled quantizer + CPU tensor: fails loudly, no fallback


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qat import models
    from qat.train import WORKLOADS, make_optimizer, train_step
    torch.manual_seed(1234)
    raw = models.FC(10, None, None, None)            # float mode: exercises layers + harness without a GPU
    raw.DROPOUT = 0.0
    for m in raw.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    model = nn.parallel.DistributedDataParallel(raw)
    opt = make_optimizer(raw, WORKLOADS["tfc"])
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.rand(8, 1, 28, 28, generator=g)
    y = torch.full((8, 10), -1.0)
    y.scatter_(1, torch.randint(0, 10, (8, 1), generator=g), 1.0)
    w0 = raw.features[2].weight.detach().clone()
    loss = train_step(model, raw, x, y, models.SqrHingeLoss(), opt)
    grad = raw.features[2].weight.grad.detach().clone()
    gathered = [torch.zeros_like(grad) for _ in range(world)]
    dist.all_gather(gathered, grad)
    ws = [torch.zeros_like(w0) for _ in range(world)]
    dist.all_gather(ws, raw.features[2].weight.detach())
    # timing aggregation used by bench.py / qat.train: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({"same_grad": bool(torch.equal(gathered[0], gathered[1])),
                    "same_weights": bool(torch.equal(ws[0], ws[1])), "moved": bool(not torch.equal(ws[0], w0)),
                    "clipped": bool(ws[0].abs().max() <= 1.0), "tmax": float(t), "loss": float(loss)}, out)
    dist.destroy_process_group()


def test_ddp_harness_gloo_world_size_2(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_ddp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["same_grad"] and r["same_weights"] and r["moved"] and r["clipped"]
    assert r["tmax"] == 2.0 and r["loss"] > 0


def _flat_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qat.train import FlatGrads
    torch.manual_seed(7)
    model = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.Flatten(),
                          nn.Linear(8 * 6 * 6, 5)).to(memory_format=torch.channels_last)
    unused = nn.Parameter(torch.ones(3))                   # takes no gradient: its bucket is reduced by finish()
    params = list(model.parameters()) + [unused]
    import copy
    plain = copy.deepcopy(model)                             # same weights, ordinary .grad tensors, no exchange
    fg = FlatGrads(params, world, dist, bucket_bytes=1024)   # several buckets
    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(4, 3, 6, 6, generator=g).contiguous(memory_format=torch.channels_last)
    res = []
    for step in range(2):                                    # second step: zero() really clears, hooks re-arm
        fg.zero()
        model(x).square().sum().backward()
        fg.finish()
        plain.zero_grad(set_to_none=True)
        plain(x).square().sum().backward()
        local = [p.grad.detach().clone() for p in plain.parameters()] + [torch.zeros(3)]
        res.append((local, [p.grad.detach().clone() for p in params]))
    views_ok = all(p.grad.untyped_storage().data_ptr() == fg.flat.untyped_storage().data_ptr() for p in params)
    strides_ok = all(p.grad.stride() == p.stride() for p in model.parameters())
    gathered = [None] * world
    dist.all_gather_object(gathered, [[t.tolist() for t in loc] for loc, _ in res])
    if rank == 0:
        ok = True
        for step in range(2):
            for i, p in enumerate(params):
                mean = sum(torch.tensor(gathered[r][step][i]) for r in range(world)) / world
                ok &= bool(torch.allclose(res[step][1][i], mean.view_as(p), atol=1e-6))
        torch.save({"ok": ok, "views": views_ok, "strides": strides_ok, "buckets": len(fg.buckets),
                    "bytes": fg.allreduce_bytes, "n": sum(p.numel() for p in params)}, out)
    dist.destroy_process_group()


def test_flat_gradient_buckets_gloo_world_size_2(tmp_path):
    """qat.train.FlatGrads (the graph-mode data-parallel exchange): gradients are views of one buffer with the
    parameters' own strides, every bucket is averaged over the ranks in place -- from the hooks during backward, or by
    finish() for parameters that took no gradient -- and the next step starts from zeros."""
    out = str(tmp_path / "flat.pt")
    mp.spawn(_flat_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["views"] and r["strides"] and r["buckets"] >= 3 and r["bytes"] == 4 * r["n"]


def test_graph_capture_refused_while_host_state_still_advances():
    from qat.train import assert_capturable
    from brevitas_b200.quant import Uint8ActPerTensorFloat
    from brevitas_b200.nn import QuantReLU
    m = QuantReLU(collect_stats_steps=5).train()
    with pytest.raises(RuntimeError, match="still collecting"):
        assert_capturable(m)
    m.act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.counter = 6
    assert_capturable(m)
    m.eval()
    assert_capturable(m)
