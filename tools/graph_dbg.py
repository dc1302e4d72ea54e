import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, brevitas_b200
sys.path.insert(0, '/root/repo')
from brevitas_b200 import _kernels as K
orig = K._launch
def dbg(dev, name, *args):
    if torch.cuda.is_current_stream_capturing() or True:
        print("launch", name, "stream", hex(torch.cuda.current_stream(dev).cuda_stream), "capturing", torch.cuda.is_current_stream_capturing(), flush=True)
    return orig(dev, name, *args)
K._launch = dbg
x = torch.randn(64, 256, device="cuda", requires_grad=True)
s = torch.tensor(0.05, device="cuda")
def step():
    x.grad = None
    y = torch.ops.brevitas_b200.int_quant(x, s, 0.0, -127.0, 127.0, 0, 1)
    y.sum().backward()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3): step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
print("---- capture")
with torch.cuda.graph(g):
    step()
g.replay(); torch.cuda.synchronize()
print("ok", x.grad.sum().item())
