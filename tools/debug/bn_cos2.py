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
    outs = []
    for step in range(steps):
        o = model(x); l = nn.functional.cross_entropy(o, t)
        model.zero_grad(); l.backward(); outs.append(o.detach().clone())
    return outs
torch.manual_seed(0)
a = models.resnet18(collect_stats_steps=1).cuda().to(memory_format=torch.channels_last).train()
sd = a.state_dict()
torch.manual_seed(0)
b = models.resnet18(collect_stats_steps=1, fuse_bn=True).cuda().to(memory_format=torch.channels_last).train()
b.load_state_dict(sd, strict=False)
oa, ob = run(a), run(b)
for s in range(3):
    print("step", s, "logits maxdiff", float((oa[s] - ob[s]).abs().max()), "of", float(oa[s].abs().max()))
rows = []
for (n, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
    if p.grad is None: continue
    d = float((p.grad - q.grad).norm()); m = float(p.grad.norm())
    rows.append((d / (m + 1e-12), n, m, d))
rows.sort(reverse=True)
for r in rows[:14]:
    print(f"{r[1]:70s} rel {r[0]:.3e} |g| {r[2]:.3e}")
tot = sum(r[2] ** 2 for r in rows) ** 0.5
print("total grad norm", tot)
