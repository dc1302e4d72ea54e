import os, sys, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import brevitas_b200
from qat import models
import brevitas_b200.fused_bn as FB
shape = (16, 3, 128, 128)
x = torch.randn(shape, generator=torch.Generator().manual_seed(3)).cuda().contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 1000, (shape[0],), generator=torch.Generator().manual_seed(4)).cuda()
def grads(model, steps=3):
    for step in range(steps):
        l = nn.functional.cross_entropy(model(x), t)
        model.zero_grad(); l.backward()
    return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None]), float(l)
torch.manual_seed(0)
a = models.resnet18(collect_stats_steps=1).cuda().to(memory_format=torch.channels_last).train()
sd = a.state_dict()
ga, la = grads(a)
orig = FB.bn_act_quant
for mode in ("all", "no_residual", "only_residual", "none_but_flag"):
    torch.manual_seed(0)
    b = models.resnet18(collect_stats_steps=1, fuse_bn=True).cuda().to(memory_format=torch.channels_last).train()
    b.load_state_dict(sd, strict=False)
    def patched(bn, act, xx, residual=None, mode=mode):
        if mode == "no_residual" and residual is not None:
            return act(bn(xx) + residual)
        if mode == "only_residual" and residual is None:
            return act(bn(xx))
        if mode == "none_but_flag":
            return act(bn(xx)) if residual is None else act(bn(xx) + residual)
        return orig(bn, act, xx, residual)
    models.bn_act_quant = patched
    gb, lb = grads(b)
    print(mode, "cos", float(torch.nn.functional.cosine_similarity(ga, gb, dim=0)), "loss", la, lb)
# same model twice unfused: run-to-run noise
torch.manual_seed(0)
c = models.resnet18(collect_stats_steps=1).cuda().to(memory_format=torch.channels_last).train()
c.load_state_dict(sd, strict=False)
gc, lc = grads(c)
print("unfused twice cos", float(torch.nn.functional.cosine_similarity(ga, gc, dim=0)), la, lc)
