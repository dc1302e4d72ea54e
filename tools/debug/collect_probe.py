import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import brevitas_b200
from brevitas_b200.core import scaling as S
from qat.train import run
mode = sys.argv[1]
if mode == "nofold":
    S.ParameterFromRuntimeStatsScaling.pre_relu_collecting = lambda self, x: False
r = run("resnet18", 256, 10, 5, collect_stats_steps=10 ** 6, channels_last=True)
print(mode, r["ms_per_step"], r["fakequant_launches_per_step"])
if mode == "prof":
    from torch.profiler import profile, ProfilerActivity
    from qat.train import build, make_optimizer, make_batch, train_step
    dev = torch.device("cuda")
    raw, loss_fn, spec = build("resnet18", dev, 10 ** 6, channels_last=True)
    opt = make_optimizer(raw, spec)
    x, y = make_batch(spec, 256, dev, 1)
    x = x.contiguous(memory_format=torch.channels_last)
    raw.train()
    for _ in range(3):
        train_step(raw, raw, x, y, loss_fn, opt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(2):
            train_step(raw, raw, x, y, loss_fn, opt)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
