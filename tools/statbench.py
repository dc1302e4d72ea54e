#!/usr/bin/env python
"""Needs the sweep build: `make -C brevitas_b200/csrc TUNING=1` and BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so.
Device time of the read-only statistic kernels (abs-max per tensor / per row, AbsPercentile radix select) at the C2
size for a sweep of resident CTAs per SM (bvb_set_tuning stream_ctas_per_sm), CUDA-graph replay.
    python tools/statbench.py [--per-sm 0,3,4,6,8]"""
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
    ap.add_argument("--per-sm", default="0,2,3,4,5,6,8")
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda")
    R, C = 4096, 11008
    n = R * C
    for dt in (torch.float32, torch.bfloat16):
        es = 4 if dt == torch.float32 else 2
        X = [torch.randn(R, C, device=dev).to(dt) for _ in range(3)]
        k = int(0.01 * 99.999 * n + 0.5)
        cases = {
            "absmax_rows": (lambda i: K.absmax_rows(X[i % 3], R, C), 1),
            "absmax_tensor": (lambda i: K.absmax_tensor(X[i % 3]), 1),
            "abs_kth_flat": (lambda i: K.abs_kth_value_rows(X[i % 3].view(1, -1), 1, n, k), 4 if es == 4 else 2),
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
                print(f"{str(dt)[6:]:9s} {name:16s} ctas/sm={per_sm}  {ms * 1e3:8.1f} us  "
                      f"{n * es * passes / ms / 1e6:7.0f} GB/s", flush=True)
                del g, keep
    _set_tuning(lib, 0, 0, 0, 0, 0)


if __name__ == "__main__":
    main()
