#!/usr/bin/env python
"""Device time of the remaining C-ABI kernels (STE primitives, clamps, binary quantizers, abs-max statistics) at the
C2 size, CUDA-graph replay.  python tools/prof_misc.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200 import _kernels as K  # noqa: E402


def main():
    dev = torch.device("cuda")
    R, C = 4096, 11008
    n = R * C
    for dt in (torch.float32, torch.bfloat16):
        es = 4 if dt == torch.float32 else 2
        X = [torch.randn(R, C, device=dev).to(dt) for _ in range(3)]
        G = [torch.randn(R, C, device=dev).to(dt) for _ in range(2)]
        s = torch.tensor(0.7, device=dev).to(dt)
        lo, hi = torch.tensor(-1.0, device=dev).to(dt), torch.tensor(1.0, device=dev).to(dt)
        cases = {
            "round_ste": (lambda i: K.unary("bvb_round_ste_impl", X[i % 3]), 2),
            "binary_sign_ste": (lambda i: K.unary("bvb_binary_sign_ste_impl", X[i % 3]), 2),
            "dpu_round_ste": (lambda i: K.unary("bvb_dpu_round_ste_impl", X[i % 3]), 2),
            "tensor_clamp_ste": (lambda i: K.tensor_clamp(X[i % 3], lo, hi), 2),
            "scalar_clamp_min_ste": (lambda i: K.scalar_clamp_min(X[i % 3], 0.1), 2),
            "binary_quant_fwd": (lambda i: K.binary_quant_fwd(X[i % 3], s, False), 2),
            "clamped_binary_quant_bwd_gs": (lambda i: K.binary_quant_bwd(G[i % 2], X[i % 3], s, True, True), 3),
            "absmax_rows": (lambda i: K.absmax_rows(X[i % 3], R, C), 1),
            "absmax_tensor": (lambda i: K.absmax_tensor(X[i % 3]), 1),
        }
        for name, (fn, passes) in cases.items():
            for i in range(3):
                fn(i)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = [fn(i) for i in range(6)]
            g.replay()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                g.replay()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 60
            print(f"{str(dt)[6:]:9s} {name:28s} {ms * 1e3:8.1f} us  {n * es * passes / ms / 1e6:7.0f} GB/s", flush=True)
            del g, keep


if __name__ == "__main__":
    main()
