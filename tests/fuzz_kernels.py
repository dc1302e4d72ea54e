#!/usr/bin/env python
"""Test infrastructure (lives under tests/ because it uses the oracle).  Randomised differential test of the C-ABI
kernels against the numpy oracle (run on the GPU box):
random shapes (aligned / ragged / tiny / wide rows), dtypes, scale layouts (scalar, per row, NCHW channel, NHWC
channel, per token), ranges, round and clamp modes, with and without the fused ReLU / tensor zero-point / integer export.
    python tests/fuzz_kernels.py --cases 400 --seed 0
Exits non-zero on the first mismatch and prints the case."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import brevitas_b200  # noqa: E402,F401
from brevitas_b200 import _kernels as K  # noqa: E402
from golden_util import assert_bits_equal  # noqa: E402
from oracle import fakequant_oracle as O  # noqa: E402

TDT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
RM = {"round": 0, "floor": 1, "ceil": 2, "round_to_zero": 3, "dpu_round": 4}
ULP = {"f32": 2.0 ** -23, "bf16": 2.0 ** -7, "f16": 2.0 ** -10}


def host(t):
    return t.detach().float().cpu().numpy()


def check_sum(got, el, idx, count, dtype, what):
    ref = np.zeros(count)
    mag = np.zeros(count)
    fin = np.isfinite(el)
    np.add.at(ref, idx[fin], el[fin])
    np.add.at(mag, idx[fin], np.abs(el[fin]))
    n = max(1, el.size // count)
    ok = np.isfinite(got) & np.isfinite(ref)
    tol = mag * (n * 2.0 ** -21 + 16 * ULP[dtype]) + 1e-4
    assert np.all(np.abs(got[ok] - ref[ok]) <= tol[ok]), (what, got[:4], ref[:4])


def fuzz_stats(rng, dtype, case):
    """the reductions behind the scale statistics: k-th smallest |x| with its first position (AbsPercentile), abs-max
    per row / per tensor -- selections, so exact whatever the dtype; NaN sorts above +inf like torch.kthvalue / max"""
    T = TDT[dtype]
    rows = int(rng.choice([1, 1, 2, 3, 4, 5, 9]))
    cols = int(rng.choice([1, 5, 250, 4096, 10007, 262144, 1_200_000 if rows <= 3 else 70001]))
    x = O.rnd((rng.standard_normal((rows, cols)) * rng.choice([1e-3, 1.0, 300.0])).astype(np.float32), dtype)
    mode = rng.choice(["plain", "dup", "relu", "special", "const"])
    if mode == "dup":
        x[:, :: int(rng.integers(2, 5))] = x[0, 0]
    elif mode == "relu":
        x = np.maximum(x, 0.0)
    elif mode == "const":
        x[:] = x[0, 0]
    elif mode == "special" and cols >= 8:
        x[:, :6] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30]
        x = O.rnd(x, dtype)
    k = int(rng.choice([1, cols, int(rng.integers(1, cols + 1)), O.percentile_k(99.999, cols) or 1, O.percentile_k(99.9, cols) or 1]))
    desc = f"case {case}: {dtype} stats rows={rows} cols={cols} mode={mode} k={k}"
    try:
        xd = torch.from_numpy(x).to(T).cuda()
        val, idx = K.abs_kth_value_rows(xd, rows, cols, k, want_index=True)
        a = np.abs(x)
        keys = np.where(np.isnan(a), np.float32(np.inf), a).view(np.uint32).astype(np.int64) + np.isnan(a)   # NaN last
        order = np.sort(keys, axis=1, kind="stable")[:, k - 1]
        ref = np.take_along_axis(a, np.argmax(keys == order[:, None], axis=1)[:, None], axis=1)[:, 0]
        assert_bits_equal(host(val), ref, "kth value")
        assert np.array_equal(idx.cpu().numpy(), np.argmax(keys == order[:, None], axis=1)), "kth first index"
        amax_keys = keys.max(axis=1)
        ref_rows = np.take_along_axis(a, np.argmax(keys == amax_keys[:, None], axis=1)[:, None], axis=1)[:, 0]
        assert_bits_equal(host(K.absmax_rows(xd, rows, cols)), ref_rows, "absmax rows")
        flat = keys.reshape(-1)
        assert_bits_equal(host(K.absmax_tensor(xd)).reshape(1), a.reshape(-1)[np.argmax(flat == flat.max())].reshape(1),
                          "absmax tensor")
    except Exception:
        print("FAILED", desc)
        raise


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    for case in range(a.cases):
        dtype = rng.choice(["f32", "bf16", "f16"])
        T = TDT[dtype]
        kind = rng.choice(["flat", "rows", "nchw", "nhwc", "token", "fused_rows", "fused_tensor", "stats"])
        rm = rng.choice(["round", "round", "round", "floor", "ceil", "round_to_zero", "dpu_round"])
        cm = rng.choice(["ste", "masked"])
        qmin, qmax = [(-127.0, 127.0), (-128.0, 127.0), (0.0, 255.0), (-8.0, 7.0), (0.0, 15.0), (-1.0, 1.0)][rng.integers(6)]
        if kind == "flat":
            shape = (int(rng.choice([1, 7, 1000, 4099, 65536 * 4 + 3, (1 << 20) + 5])),)
            sshape = ()
        elif kind == "rows" or kind == "fused_rows":
            shape = (int(rng.choice([1, 3, 37, 300, 1300])), int(rng.choice([1, 8, 17, 264, 1000, 2048, 4096, 8200, 11008])))
            if shape[0] * shape[1] > 12_000_000:          # 1300 x 8200: several rows per CTA of the TMA-store forward
                shape = (1300, 8200)
            sshape = (shape[0], 1)
        elif kind in ("nchw", "nhwc"):
            shape = (int(rng.integers(1, 5)), int(rng.choice([3, 16, 24, 64, 256])), int(rng.choice([1, 5, 14])), int(rng.choice([1, 7, 14])))
            sshape = (1, shape[1], 1, 1)
        elif kind == "token":
            shape = (int(rng.integers(1, 4)), int(rng.integers(1, 40)), int(rng.choice([8, 33, 4096])))
            sshape = (shape[0], shape[1], 1)
        else:
            shape = (int(rng.choice([1, 9, 1023, 100003])),)
            sshape = ()
        if kind == "stats":
            fuzz_stats(rng, dtype, case)
            continue
        x = O.rnd((rng.standard_normal(shape) * rng.choice([0.3, 5.0, 60.0])).astype(np.float32), dtype)
        g = O.rnd(rng.standard_normal(shape).astype(np.float32), dtype)
        flat = x.reshape(-1)
        if flat.size >= 8 and rng.random() < 0.5:
            flat[:6] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30]
            x = O.rnd(x, dtype)                       # the special values must be values of the dtype too
        desc = f"case {case}: {dtype} {kind} shape={shape} rm={rm} cm={cm} range=({qmin},{qmax})"
        try:
            xd, gd = torch.from_numpy(x).to(T).cuda(), torch.from_numpy(g).to(T).cuda()
            if kind == "nhwc":
                xd = xd.contiguous(memory_format=torch.channels_last)
            if kind == "fused_rows":
                fin = np.where(np.isfinite(x), x, 1.0).astype(np.float32)
                xd = torch.from_numpy(fin).to(T).cuda()
                thr = max(abs(qmin), abs(qmax))
                yo, so, _ = O.rows_absmax_int_quant_forward(fin, 1e-10, thr, 0.0, qmin, qmax, rm, dtype)
                y, s, _ = K.rows_absmax_int_quant_fwd(xd, shape[0], shape[1], 1e-10, thr, 0.0, qmin, qmax, RM[rm])
                assert_bits_equal(host(s), so, "fused rows scale")
                assert_bits_equal(host(y), yo, "fused rows y")
                gxo, _ = O.rows_absmax_int_quant_backward(g, fin, so, None, thr, 0.0, qmin, qmax, rm, cm, dtype)
                gx = host(K.rows_absmax_int_quant_bwd(gd, xd, s, None, shape[0], shape[1], thr, 0.0, qmin, qmax, RM[rm],
                                                      1 if cm == "masked" else 0))
                first = np.zeros_like(fin, dtype=bool)
                first[np.arange(shape[0]), np.abs(fin).argmax(axis=1)] = True
                assert_bits_equal(np.where(first, 0, gx), np.where(first, 0, gxo), "fused rows gx")
                continue
            if kind == "fused_tensor":
                fin = np.where(np.isfinite(x), x, 1.0).astype(np.float32)
                xd = torch.from_numpy(fin).to(T).cuda()
                thr = max(abs(qmin), abs(qmax))
                yo, so, _ = O.tensor_absmax_int_quant_forward(fin, 1e-10, thr, 0.0, qmin, qmax, rm, dtype)
                y, s, am = K.tensor_absmax_int_quant_fwd(xd, T, 1e-10, thr, 0.0, qmin, qmax, RM[rm])
                assert_bits_equal(host(s), so, "fused tensor scale")
                assert_bits_equal(host(y), yo, "fused tensor y")
                continue
            s = O.rnd((np.abs(rng.standard_normal(sshape)) * 0.3 + 0.05).astype(np.float32), dtype)
            sd = torch.from_numpy(np.asarray(s)).reshape(sshape).to(T).cuda()
            zp = float(rng.choice([0.0, 0.0, 3.0, -2.0]))
            variant = rng.choice(["plain", "relu", "zpt", "to_int"]) if zp == 0.0 else rng.choice(["plain", "zpt"])
            idx = np.broadcast_to(np.arange(max(1, s.size)).reshape(s.shape), shape).reshape(-1)
            cmi = 1 if cm == "masked" else 0
            if variant == "plain":
                yo = O.int_quant_forward(x, s, zp, qmin, qmax, rm, dtype)
                assert_bits_equal(host(K.int_quant_fwd(xd, sd, zp, qmin, qmax, RM[rm])), yo, "y")
                gxo, gs_el = O.int_quant_backward(g, x, s, zp, qmin, qmax, rm, cm, dtype)
                gx, gs = K.int_quant_bwd(gd, xd, sd, zp, qmin, qmax, RM[rm], cmi, True)
                assert_bits_equal(host(gx), gxo, "gx")
                check_sum(host(gs).reshape(-1).astype(np.float64), gs_el.reshape(-1), idx, max(1, s.size), dtype, "gscale")
            elif variant == "relu":
                yo = O.relu_int_quant_forward(x, s, 0.0, qmin, qmax, rm, dtype)
                assert_bits_equal(host(K.int_quant_fwd(xd, sd, 0.0, qmin, qmax, RM[rm], pre_relu=True)), yo, "relu y")
                gxo, gs_el = O.relu_int_quant_backward(g, x, s, 0.0, qmin, qmax, rm, cm, dtype)
                gx, gs = K.int_quant_bwd(gd, xd, sd, 0.0, qmin, qmax, RM[rm], cmi, True, pre_relu=True)
                assert_bits_equal(host(gx), gxo, "relu gx")
            elif variant == "zpt":
                if kind == "nhwc":
                    xd = xd.contiguous()
                zt = torch.full_like(sd, zp)
                yo = O.int_quant_forward(x, s, zp, qmin, qmax, rm, dtype)
                assert_bits_equal(host(K.int_quant_zpt_fwd(xd, sd, zt, qmin, qmax, RM[rm])), yo, "zpt y")
                gxo, gs_el = O.int_quant_backward(g, x, s, zp, qmin, qmax, rm, cm, dtype)
                gx, gs, gz = K.int_quant_zpt_bwd(gd, xd, sd, zt, qmin, qmax, RM[rm], cmi, True)
                assert_bits_equal(host(gx), gxo, "zpt gx")
                check_sum(host(gs).reshape(-1).astype(np.float64), gs_el.reshape(-1), idx, max(1, s.size), dtype, "zpt gscale")
            else:
                out_dt = torch.uint8 if qmin >= 0 else torch.int8
                ref = O.int_quant_to_int(x, s, 0.0, qmin, qmax, rm, dtype, np.uint8 if qmin >= 0 else np.int8)
                got = K.int_quant_to_int(xd, sd, 0.0, qmin, qmax, RM[rm], out_dt).cpu().numpy()
                assert np.array_equal(got, ref), "to_int"
        except AssertionError as e:
            print("MISMATCH", desc, "\n", str(e)[:1500])
            return 1
    print(f"fuzz OK: {a.cases} cases, seed {a.seed}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
