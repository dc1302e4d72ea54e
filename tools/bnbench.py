#!/usr/bin/env python
"""batch-norm + ReLU + activation quantizer: fused passes vs the unfused pair (cuDNN batch-norm + relu_int_quant), forward
and forward+backward, CUDA-graph replay over rotating inputs.   python tools/bnbench.py"""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200.fused_bn import bn_act_quant  # noqa: E402
from brevitas_b200.nn import QuantReLU  # noqa: E402
from qat.models import CommonUintActQuant  # noqa: E402

SHAPES = [(256, 64, 112, 112), (256, 64, 56, 56), (256, 128, 28, 28), (256, 256, 14, 14), (256, 512, 7, 7),
          (128, 32, 112, 112), (128, 512, 14, 14), (128, 1024, 7, 7)]


def timeit(fn, reps=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    for shape in SHAPES:
        c = shape[1]
        act = QuantReLU(act_quant=CommonUintActQuant, bit_width=4, per_channel_broadcastable_shape=(1, c, 1, 1),
                        scaling_per_output_channel=False).cuda().train()
        bn = nn.BatchNorm2d(c).cuda().train()
        x = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_(True)
        gy = torch.randn(shape, device="cuda").contiguous(memory_format=torch.channels_last)
        nbytes = x.numel() * 4

        def unfused(bwd):
            y = act(bn(x))
            if bwd:
                x.grad = None
                y.backward(gy)

        def fused(bwd):
            y = bn_act_quant(bn, act, x)
            if bwd:
                x.grad = None
                y.backward(gy)
        res = {}
        for name, fn in (("unfused", unfused), ("fused", fused)):
            res[name] = (timeit(lambda: fn(False)), timeit(lambda: fn(True)))
        uf, ff = res["unfused"], res["fused"]
        print(f"{shape}: fwd {uf[0] * 1e3:.0f} -> {ff[0] * 1e3:.0f} us ({3 * nbytes / ff[0] / 1e6:.0f} GB/s at 3 passes); "
              f"fwd+bwd {uf[1] * 1e3:.0f} -> {ff[1] * 1e3:.0f} us ({8 * nbytes / ff[1] / 1e6:.0f} GB/s at 8 passes; "
              f"unfused {13 * nbytes / uf[1] / 1e6:.0f} GB/s at 13)")


if __name__ == "__main__":
    main()
