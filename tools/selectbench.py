#!/usr/bin/env python
"""Device time of the ReLU-folded exact percentile select (collection phase of a QuantReLU) on the activation shapes of
ResNet-18 at batch 256, against two bare reads of the tensor.  python tools/selectbench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200 import _kernels as K  # noqa: E402


def main():
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(0)
    total = 0.0
    for c, hw, times in ((64, 112, 1), (64, 56, 4), (128, 28, 4), (256, 14, 4), (512, 7, 4)):
        n = 256 * c * hw * hw
        nsets = max(2, int(300e6 // (n * 4)) + 1)
        xs = [torch.randn(n, device=dev, generator=gen) for _ in range(nsets)]
        k = int(0.99999 * n + 0.5)
        fn = lambda i: K.abs_kth_value_rows(xs[i % nsets], 1, n, k, pre_relu=True, want_index=True)  # noqa: E731
        for i in range(3):
            fn(i)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            keep = [fn(i) for i in range(nsets)]
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (5 * nsets) * 1e3
        ideal = 2 * n * 4 / 6552.6e9 * 1e6
        total += us * times
        print(f"[256,{c},{hw},{hw}] {n * 4 / 1e6:7.1f} MB  select {us:7.1f} us   two bare reads {ideal:6.1f} us   x{times} per step")
        del graph, keep, xs
    print(f"sum over the ReLU quantizers of one ResNet-18 step: {total / 1e3:.2f} ms")


if __name__ == "__main__":
    main()
