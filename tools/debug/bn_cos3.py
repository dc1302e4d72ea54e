import os, sys, torch
from torch import nn
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import brevitas_b200
from qat import models
shape = (16, 3, 128, 128)
x = torch.randn(shape, generator=torch.Generator().manual_seed(3)).cuda().contiguous(memory_format=torch.channels_last)
t = torch.randint(0, 1000, (shape[0],), generator=torch.Generator().manual_seed(4)).cuda()
def run(model, steps=3):
    for step in range(steps):
        o = model(x); l = nn.functional.cross_entropy(o, t)
        model.zero_grad(); l.backward()
    g = torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.grad is not None])
    return o.detach().clone(), g
torch.manual_seed(0)
a = models.resnet18(collect_stats_steps=1).cuda().to(memory_format=torch.channels_last).train()
sd = a.state_dict()
oa, ga = run(a)
# the same UNFUSED model with every batch-norm evaluated on an NCHW copy (another cuDNN kernel, another summation order)
class NchwBN(nn.BatchNorm2d):
    def forward(self, inp):
        return super().forward(inp.contiguous()).contiguous(memory_format=torch.channels_last)
torch.manual_seed(0)
c = models.resnet18(collect_stats_steps=1).cuda().to(memory_format=torch.channels_last).train()
c.load_state_dict(sd, strict=False)
for m in c.modules():
    if type(m) is nn.BatchNorm2d:
        m.__class__ = NchwBN
oc, gc = run(c)
print("unfused NHWC-BN vs unfused NCHW-BN: logits maxdiff", float((oa - oc).abs().max()), "grad cos", float(torch.nn.functional.cosine_similarity(ga, gc, dim=0)))
torch.manual_seed(0)
b = models.resnet18(collect_stats_steps=1, fuse_bn=True).cuda().to(memory_format=torch.channels_last).train()
b.load_state_dict(sd, strict=False)
ob, gb = run(b)
print("unfused vs fused: logits maxdiff", float((oa - ob).abs().max()), "grad cos", float(torch.nn.functional.cosine_similarity(ga, gb, dim=0)))
print("NCHW-BN vs fused: grad cos", float(torch.nn.functional.cosine_similarity(gc, gb, dim=0)))
