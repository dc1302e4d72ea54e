#!/usr/bin/env python
"""Needs the sweep build: `make -C brevitas_b200/csrc TUNING=1` and BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so.
Kernel micro-benchmark / tuning sweep for the per-row fused kernels (run on the GPU box).

    python tools/kbench.py --kernel bwd --dtype f32 --rows 4096 --cols 11008 --sweep
Prints one line per geometry: kernel time (CUDA events, inputs rotated over > L2 worth of buffers) and GB/s.
bvb_set_tuning(rows_threads, rows_stages, rows_ctas_per_sm, stream_threads, stream_ctas_per_sm):
  fwd: threads, stages, CTAs/SM            bwd: stream_threads = 32*consumer warps, rows_threads = vectors per
  thread per tile, rows_stages = ring depth, stream_ctas_per_sm = CTAs/SM
"""
def _set_tuning(lib, *a):
    import ctypes
    fn = getattr(lib, "bvb_set_tuning", None)
    if fn is None and not any(a):
        return                                   # product library, default geometry: nothing to set
    if fn is None:
        raise SystemExit("bvb_set_tuning is only in the sweep build: make -C brevitas_b200/csrc TUNING=1; "
                         "BREVITAS_B200_LIB=brevitas_b200/libbrevitas_b200_tuning.so")
    fn.restype, fn.argtypes = None, [ctypes.c_int] * 5
    fn(*a)


import argparse
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200 import _lib  # noqa: E402

DT = {"f32": (torch.float32, _lib.F32, 4), "bf16": (torch.bfloat16, _lib.BF16, 2), "f16": (torch.float16, _lib.F16, 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kernel", default="bwd", choices=["fwd", "bwd", "both", "scaled_bwd"])
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--rows", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=11008)
    ap.add_argument("--masked", type=int, default=0)
    ap.add_argument("--reps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--small", action="store_true", help="fwd sweep over 32/64/128-thread CTAs")
    ap.add_argument("--store", action="store_true", help="fwd sweep of the TMA-store variant (rows_stages = 100 + stages)")
    ap.add_argument("--tuning", type=str, default="")
    a = ap.parse_args()
    tdt, tag, esz = DT[a.dtype]
    lib = _lib.load()
    dev = torch.device("cuda")
    n = a.rows * a.cols
    nsets = max(2, int(400e6 // (n * esz)) + 1)
    g = torch.Generator(device=dev).manual_seed(0)
    X = [torch.randn(a.rows, a.cols, device=dev, generator=g).to(tdt) for _ in range(nsets)]
    G = [torch.randn(a.rows, a.cols, device=dev, generator=g).to(tdt) for _ in range(nsets)]
    Y = torch.empty_like(X[0])
    GXO = torch.empty_like(X[0])
    S = torch.empty(a.rows, device=dev, dtype=tdt)
    GS = torch.zeros(1, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    qmin, qmax, thr = -127.0, 127.0, 127.0
    lib.bvb_rows_absmax_int_quant_fwd(X[0].data_ptr(), Y.data_ptr(), S.data_ptr(), None, a.rows, a.cols, 1e-10, thr,
                                      0.0, qmin, qmax, 0, tag, st)

    def run(i):
        k = i % nsets
        if a.kernel == "both":           # the bench.py step: forward then backward of the same weight
            rc = lib.bvb_rows_absmax_int_quant_fwd(X[k].data_ptr(), Y.data_ptr(), S.data_ptr(), None, a.rows, a.cols,
                                                   1e-10, thr, 0.0, qmin, qmax, 0, tag, st)
            return rc | lib.bvb_rows_absmax_int_quant_bwd(G[k].data_ptr(), X[k].data_ptr(), S.data_ptr(), None,
                                                          GXO.data_ptr(), a.rows, a.cols, thr, 0.0, qmin, qmax, 0,
                                                          a.masked, tag, st)
        if a.kernel == "scaled_bwd":     # provided scalar scale, d(scale) wanted (learned activation scale)
            return lib.bvb_int_quant_bwd(G[k].data_ptr(), X[k].data_ptr(), S.data_ptr(), Y.data_ptr(), GS.data_ptr(), n, 1, 1,
                                         tag, 0.0, 0.0, 255.0, 0, a.masked, tag, st)
        if a.kernel == "fwd":
            return lib.bvb_rows_absmax_int_quant_fwd(X[k].data_ptr(), Y.data_ptr(), S.data_ptr(), None, a.rows, a.cols,
                                                     1e-10, thr, 0.0, qmin, qmax, 0, tag, st)
        return lib.bvb_rows_absmax_int_quant_bwd(G[k].data_ptr(), X[k].data_ptr(), S.data_ptr(), None, Y.data_ptr(),
                                                 a.rows, a.cols, thr, 0.0, qmin, qmax, 0, a.masked, tag, st)

    def timeit():
        for i in range(a.warmup):
            if run(i):
                return None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()          # ncu --profile-from-start off captures only the timed launches
        e0.record()
        for i in range(a.reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return e0.elapsed_time(e1) / a.reps

    bytes_ = n * esz * {"fwd": 2, "bwd": 3, "both": 5, "scaled_bwd": 3}[a.kernel]
    if a.kernel == "scaled_bwd":
        S.fill_(0.02)
    if a.tuning:
        combos = [tuple(int(v) for v in a.tuning.split(","))]
    elif not a.sweep:
        combos = [(0, 0, 0, 0, 0)]
    elif a.kernel == "fwd" and a.store:
        combos = [(0, 0, 0, 0, 0)] + [(t, 100 + s, c, 0, 0) for t, s, c in
                                      itertools.product(*(([256, 512, 1024], [2, 3, 4], [1, 2]) if a.dtype == "f32" else
                                                          ([64, 128, 256], [2, 3], [3, 4, 5, 6, 8, 10])))]
        # the variant must produce the same bits as the default kernel
        _set_tuning(lib, 0, 0, 0, 0, 0)
        run(0)
        torch.cuda.synchronize()
        y0, s0 = Y.clone(), S.clone()
        for c in combos[1:]:
            _set_tuning(lib, *c)
            Y.zero_()
            if run(0):
                continue
            torch.cuda.synchronize()
            assert torch.equal(Y.view(torch.int16 if esz == 2 else torch.int32), y0.view(torch.int16 if esz == 2 else torch.int32)), c
            assert torch.equal(S, s0), c
        print("store variant bit-identical to the default kernel for", len(combos) - 1, "geometries")
    elif a.kernel == "fwd" and a.small:      # few warps per row: 1-2 warps own a row (less per-row overhead per element)
        combos = [(t, s, c, 0, 0) for t, s, c in itertools.product([32, 64, 128], [2, 3], [4, 6, 8, 10, 12, 16])]
    elif a.kernel == "fwd":
        combos = [(t, s, c, 0, 0) for t, s, c in itertools.product([128, 256, 512, 1024], [2, 3, 4], [1, 2, 3, 4, 6])]
    else:
        combos = [(pt, s, 0, 32 * w, c) for w, pt, s, c in           # w = consumer warps
                  itertools.product([2, 3, 4, 6, 8], [2, 4, 8], [2, 3], [2, 3, 4, 6, 8])]
    best = None
    for c in combos:
        _set_tuning(lib, *c)
        ms = timeit()
        if ms is None:
            print(c, "launch failed:", _lib.last_error())
            torch.cuda.synchronize()
            continue
        gbps = bytes_ / (ms * 1e-3) / 1e9
        print(f"{a.kernel} {a.dtype} {a.rows}x{a.cols} tuning={c}: {ms * 1e3:.1f} us  {gbps:.0f} GB/s")
        if best is None or ms < best[0]:
            best = (ms, c)
    if best:
        print("BEST", best[1], f"{best[0] * 1e3:.1f} us", f"{bytes_ / (best[0] * 1e-3) / 1e9:.0f} GB/s")
    _set_tuning(lib, 0, 0, 0, 0, 0)


if __name__ == "__main__":
    main()
