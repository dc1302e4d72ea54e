#!/usr/bin/env python
"""Needs the sweep build: `make -C brevitas_b200/csrc TUNING=1` and BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so.
Device time of the channels-last per-channel-scale kernels (NHWC activation, [1,C,1,1] scale) for a sweep of CTAs
per SM (bvb_set_tuning stream_ctas_per_sm), CUDA-graph replay.   python tools/clbench.py [--per-sm 0,2,3,4,6,8]"""
def _set_tuning(lib, *a):
    import ctypes
    fn = getattr(lib, "bvb_set_tuning", None)
    if fn is None:
        raise SystemExit("bvb_set_tuning is only in the sweep build: make -C brevitas_b200/csrc TUNING=1; "
                         "BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so")
    fn.restype, fn.argtypes = None, [ctypes.c_int] * 5
    fn(*a)


import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200 import _kernels as K  # noqa: E402
from brevitas_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--per-sm", default="0,2,3,4,6,8,16")
    ap.add_argument("--shape", default="128,256,28,28")
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda")
    shape = tuple(int(v) for v in a.shape.split(","))
    for dt in (torch.float32, torch.bfloat16):
        es = 4 if dt == torch.float32 else 2
        X = [torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for _ in range(3)]
        G = [torch.randn(shape, device=dev).to(dt).contiguous(memory_format=torch.channels_last) for _ in range(2)]
        s = (torch.rand(1, shape[1], 1, 1, device=dev) * 0.02 + 0.01).to(dt)
        n = X[0].numel()
        cases = {
            "chanlast_fwd": (lambda i: K.int_quant_fwd(X[i % 3], s, 0.0, 0.0, 255.0, 0), 2),
            "chanlast_bwd_masked_gs": (lambda i: K.int_quant_bwd(G[i % 2], X[i % 3], s, 0.0, 0.0, 255.0, 0, 1, True), 3),
        }
        for name, (fn, passes) in cases.items():
            for per_sm in [int(v) for v in a.per_sm.split(",")]:
                _set_tuning(lib, 0, 0, 0, 0, per_sm)
                for i in range(3):
                    fn(i)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    keep = [fn(i) for i in range(6)]
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 60
                print(f"{str(dt)[6:]:9s} {name:24s} ctas/sm={per_sm:<3d} {ms * 1e3:8.1f} us  "
                      f"{n * es * passes / ms / 1e6:7.0f} GB/s", flush=True)
                del g, keep
    _set_tuning(lib, 0, 0, 0, 0, 0)


if __name__ == "__main__":
    main()
