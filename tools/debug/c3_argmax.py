import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import brevitas_b200
from brevitas_b200.binding import uninstall
from ref_util import reference_src
B, T, C = 2, 512, 4096
gen = torch.Generator().manual_seed(0)
x_host = torch.randn(B, T, C, generator=gen).to(torch.bfloat16)
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
        RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(), (B, T, 1), False, 0.1, 1e-10),
        IntScaling(True, False), ZeroZeroPoint(), BitWidthConst(8)).cuda().train()
def run(tq, dt):
    x = x_host.to(dt).cuda().requires_grad_(True)
    y, s, _, _ = tq(x)
    y.backward(g_host.to(dt).cuda())
    return y.detach(), s.detach(), x.grad.clone()
brevitas_b200.install(reference_src(), fuse=False); uninstall()
ref = run(build(), torch.bfloat16)
ref32 = run(build(), torch.float32)     # same bf16-valued inputs, fp32 arithmetic
brevitas_b200.install(reference_src(), fuse=True)
got = run(build(), torch.bfloat16)
xa = x_host.cuda().float().abs()
am = xa.argmax(dim=2)
idx = (torch.arange(B).view(B,1).expand(B,T).reshape(-1).cuda(), torch.arange(T).repeat(B).cuda(), am.reshape(-1))
o, r, r32 = got[2].float()[idx], ref[2].float()[idx], ref32[2][idx]
print("scale equal:", torch.equal(got[1], ref[1]), "y equal:", torch.equal(got[0], ref[0]))
print("ours vs ref(bf16): max abs diff", float((o - r).abs().max()))
print("ours vs ref(fp32 arithmetic on same inputs): max", float((o - r32).abs().max()), "mean", float((o - r32).abs().mean()))
print("ref(bf16) vs ref(fp32): max", float((r - r32).abs().max()), "mean", float((r - r32).abs().mean()))
w = (o - r32).abs().argmax()
print("worst ours:", float(o[w]), "ref bf16:", float(r[w]), "ref fp32:", float(r32[w]), "row", int(w))
# ---- the test's fp64 evaluation
xd, gd = x_host.cuda(), g_host.cuda()
sc = got[1]
t3 = torch.round(xd / sc)
m = ((t3 <= 127) & (t3 >= -128)).double()
ew = (((gd * sc) / sc).float() * m.float())
s64, g64, x64 = sc.double(), gd.double(), xd.double()
codes64 = torch.round(got[0].double() / s64)
t_a, t_b = g64 * codes64, m * (g64 * s64) * x64 / (s64 * s64)
gs64 = (t_a - t_b).sum(dim=2) / 128.0
fix = torch.sign(x64[idx]) * gs64.reshape(-1)
want = ew.double()[idx] + fix
e_o, e_r = (o.double() - want).abs(), (r.double() - want).abs()
print("vs fp64 formula: ours max", float(e_o.max()), "ref max", float(e_r.max()))
for w in e_o.argsort(descending=True)[:4]:
    w = int(w)
    print(f"row {w}: x={float(xd[idx][w])} g={float(gd[idx][w])} s={float(sc.reshape(-1)[w])} t3={float(t3[idx][w])} m={float(m[idx][w])} "
          f"ew={float(ew[idx][w])} fix={float(fix[w])} want={float(want[w])} ours={float(o[w])} ref={float(r[w])} "
          f"Sa={float(t_a.sum(2).reshape(-1)[w])} Sb={float(t_b.sum(2).reshape(-1)[w])}")
