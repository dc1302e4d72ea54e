import os, sys, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import brevitas_b200
from brevitas_b200.nn import QuantReLU
from brevitas_b200.fused_bn import bn_act_quant
for shape in [(16, 64, 32, 32), (16, 512, 4, 4)]:
    torch.manual_seed(1)
    acts = [QuantReLU(collect_stats_steps=1).cuda().train() for _ in range(2)]
    bns = [nn.BatchNorm2d(shape[1]).cuda().train() for _ in range(2)]
    bns[1].load_state_dict(bns[0].state_dict())
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(shape, generator=g) * 1.5 + 0.2).cuda().contiguous(memory_format=torch.channels_last)
    gy = torch.randn(shape, generator=g).cuda().contiguous(memory_format=torch.channels_last)
    for step in range(3):
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        ya = acts[0](bns[0](xa))
        yb = bn_act_quant(bns[1], acts[1], xb)
        ya.backward(gy); yb.backward(gy)
        va = acts[0].act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value
        vb = acts[1].act_quant.fused_activation_quant_proxy.tensor_quant.scaling_impl.value
        cos = float(torch.nn.functional.cosine_similarity(xa.grad.reshape(-1), xb.grad.reshape(-1), dim=0))
        print(shape, "step", step, "y maxdiff", float((ya - yb).abs().max()), "frac moved", float(((ya - yb).abs() > 0).float().mean()),
              "dx cos", cos, "dx maxdiff", float((xa.grad - xb.grad).abs().max()), "of", float(xa.grad.abs().max()),
              "value", float(va), float(vb), "dvalue", None if va.grad is None else float(va.grad), None if vb.grad is None else float(vb.grad),
              "dgamma maxdiff", float((bns[0].weight.grad - bns[1].weight.grad).abs().max()), "of", float(bns[0].weight.grad.abs().max()))
        for m in bns + [va, vb]:
            if isinstance(m, nn.Module):
                m.zero_grad()
            else:
                m.grad = None
