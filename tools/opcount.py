#!/usr/bin/env python
"""Which dispatcher ops does one steady-state QAT step issue, and on how many elements?  (finds the one-element ATen ops
around the fused kernels)  python tools/opcount.py [--model mobilenet_v1] [--fuse-bn]"""
import argparse
import collections
import os
import sys

import torch
from torch.utils._python_dispatch import TorchDispatchMode

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Count(TorchDispatchMode):
    def __init__(self):
        super().__init__()
        self.ops = collections.Counter()

    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        out = func(*args, **(kwargs or {}))
        o = out[0] if isinstance(out, (tuple, list)) and out else out
        n = o.numel() if isinstance(o, torch.Tensor) else -1
        size = "tiny" if 0 <= n <= 4096 else ("view" if n < 0 else "big")
        self.ops[(str(func), size)] += 1
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="mobilenet_v1")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--fuse-bn", action="store_true")
    a = ap.parse_args()
    from qat.train import build, make_batch, make_optimizer, train_step
    dev = torch.device("cuda")
    model, loss_fn, spec = build(a.model, dev, 2, channels_last=True, fuse_bn=a.fuse_bn)
    opt = make_optimizer(model, spec)
    model.train()
    x, y = make_batch(spec, a.batch, dev, 0)
    x = x.contiguous(memory_format=torch.channels_last)
    for _ in range(5):
        train_step(model, model, x, y, loss_fn, opt)
    with Count() as c:
        train_step(model, model, x, y, loss_fn, opt)
    total = sum(c.ops.values())
    print(f"{total} dispatcher calls in one step")
    for (name, size), n in c.ops.most_common(45):
        print(f"{n:5d}  {size:5s} {name}")


if __name__ == "__main__":
    main()
