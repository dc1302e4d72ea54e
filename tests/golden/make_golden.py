"""Generate tests/golden/*.npz by running the REAL reference (Giuseppe5/brevitas, /root/reference/src) on CPU.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The committed .npz files are the pin for both oracles (oracle/) and, through them, for the CUDA kernels.
Import recipe: SURVEY.md Appendix C (the `brevitas.inject` package needs the absent third-party `dependencies`,
so a stub package object lets `brevitas.inject.enum` load and nothing else).  BREVITAS_JIT=0, Python STE backend
-- the configuration the reference's CI runs.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    os.environ.setdefault("BREVITAS_JIT", "0")
    stub = types.ModuleType("brevitas.inject")
    stub.__path__ = [os.path.join(REF, "brevitas", "inject")]
    sys.modules["brevitas.inject"] = stub
    sys.path.insert(0, REF)
    import brevitas  # noqa: F401
    return brevitas


DT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def npf(t):
    return t.detach().to(torch.float32).cpu().numpy().copy()      # copy: buffers are updated in place later


def edge_values():
    v = [0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49999997, 126.5, 127.0, 127.5, 128.5, -127.5, -128.0, -128.5,
         254.5, 255.5, 1e-30, -1e-30, 3e38, -3e38, float("inf"), float("-inf"), float("nan"), 7.0, -7.0, 0.3, -0.3]
    return torch.tensor(v, dtype=torch.float32)


def make_input(shape, seed, scale=1.0, with_edges=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g) * scale
    if with_edges:
        e = edge_values()
        flat = x.view(-1)
        n = min(e.numel(), flat.numel())
        flat[:n] = e[:n]
    return x


def gen_ste(out):
    from brevitas.function import ops_ste
    x = torch.cat([edge_values(), make_input((97,), 1, 3.0, False)])
    for dname, dt in DT.items():
        xd = x.to(dt)
        for name in ["round_ste", "ceil_ste", "floor_ste", "binary_sign_ste", "ternary_sign_ste", "round_to_zero_ste",
                     "dpu_round_ste", "abs_binary_sign_grad"]:
            xi = xd.clone().requires_grad_(True)
            y = getattr(ops_ste, name)(xi)
            g = torch.linspace(-1, 1, xi.numel()).to(dt)
            y.backward(g)
            out[f"ste/{name}/{dname}/x"] = npf(xd)
            out[f"ste/{name}/{dname}/y"] = npf(y)
            out[f"ste/{name}/{dname}/g"] = npf(g)
            out[f"ste/{name}/{dname}/gx"] = npf(xi.grad)
        lo, hi = torch.tensor(-1.25).to(dt), torch.tensor(2.5).to(dt)
        out[f"ste/tensor_clamp_ste/{dname}/x"] = npf(xd)
        out[f"ste/tensor_clamp_ste/{dname}/y"] = npf(ops_ste.tensor_clamp_ste(xd, lo, hi))
        xm = xd.clone()
        out[f"ste/tensor_clamp_ste_/{dname}/y"] = npf(ops_ste.tensor_clamp_ste_(xm, lo, hi))
        out[f"ste/scalar_clamp_ste/{dname}/y"] = npf(ops_ste.scalar_clamp_ste(xd, -1.3, 2.7))
        out[f"ste/scalar_clamp_min_ste/{dname}/y"] = npf(ops_ste.scalar_clamp_min_ste(xd, 0.3))
        # per-row tensor bounds
        x2 = make_input((6, 40), 5, 2.0).to(dt)
        lo2 = (-torch.rand(6, 1, generator=torch.Generator().manual_seed(2))).to(dt)
        hi2 = torch.rand(6, 1, generator=torch.Generator().manual_seed(3)).to(dt)
        out[f"ste/tensor_clamp_rows/{dname}/x"] = npf(x2)
        out[f"ste/tensor_clamp_rows/{dname}/lo"] = npf(lo2)
        out[f"ste/tensor_clamp_rows/{dname}/hi"] = npf(hi2)
        out[f"ste/tensor_clamp_rows/{dname}/y"] = npf(ops_ste.tensor_clamp_ste(x2, lo2, hi2))


INT_CASES = [
    # name, signed, narrow, bits, round impl, clamp impl, scale kind, zero point
    ("s8n_round_ste_scalar", True, True, 8, "RoundSte", "TensorClampSte", "scalar", 0.0),
    ("s8_round_masked_scalar", True, False, 8, "RoundSte", "TensorClamp", "scalar", 0.0),
    ("u8_round_masked_scalar", False, False, 8, "RoundSte", "TensorClamp", "scalar", 0.0),
    ("u8_round_masked_scalar_zp", False, False, 8, "RoundSte", "TensorClamp", "scalar", 3.0),
    ("s4n_floor_ste_rows", True, True, 4, "FloorSte", "TensorClampSte", "rows", 0.0),
    ("s4_ceil_masked_rows", True, False, 4, "CeilSte", "TensorClamp", "rows", 0.0),
    ("u4n_rtz_masked_chan", False, True, 4, "RoundToZeroSte", "TensorClamp", "chan", 0.0),
    ("s8_dpu_masked_chan_zp", True, False, 8, "DPURoundSte", "TensorClamp", "chan", -2.0),
    ("s2n_round_masked_scalar", True, True, 2, "RoundSte", "TensorClamp", "scalar", 0.0),
    ("s8n_round_masked_token", True, True, 8, "RoundSte", "TensorClamp", "token", 0.0),
]


def gen_int_quant(out):
    from brevitas.core import function_wrapper as fw
    from brevitas.core.quant import IntQuant
    for (name, signed, narrow, bits, rimpl, cimpl, skind, zp) in INT_CASES:
        for dname, dt in DT.items():
            if skind == "scalar":
                x = make_input((5, 67), 11, 40.0).to(dt)
                scale = torch.tensor(0.37).to(dt)
            elif skind == "rows":
                x = make_input((6, 48), 12, 3.0).to(dt)
                scale = (torch.rand(6, 1, generator=torch.Generator().manual_seed(7)) * 0.5 + 0.05).to(dt)
            elif skind == "chan":
                x = make_input((3, 5, 4, 8), 13, 2.0).to(dt)
                scale = (torch.rand(1, 5, 1, 1, generator=torch.Generator().manual_seed(8)) * 0.3 + 0.02).to(dt)
            else:
                x = make_input((2, 7, 32), 14, 2.0).to(dt)
                scale = (torch.rand(2, 7, 1, generator=torch.Generator().manual_seed(9)) * 0.05 + 0.01).to(dt)
            iq = IntQuant(narrow_range=narrow, signed=signed, float_to_int_impl=getattr(fw, rimpl)(),
                          tensor_clamp_impl=getattr(fw, cimpl)())
            xi = x.clone().requires_grad_(True)
            si = scale.clone().requires_grad_(True)
            zpt = torch.tensor(zp)
            bw = torch.tensor(float(bits))
            y = iq(si, zpt, bw, xi)
            codes = iq.to_int(scale, zpt, bw, x)
            g = make_input(x.shape, 21, 1.0, False).to(dt)
            y.backward(g)
            p = f"int_quant/{name}/{dname}/"
            out[p + "x"], out[p + "scale"], out[p + "g"] = npf(x), npf(scale), npf(g)
            out[p + "y"], out[p + "codes"] = npf(y), npf(codes)
            out[p + "gx"], out[p + "gscale"] = npf(xi.grad), npf(si.grad)
            out[p + "meta"] = np.array([float(signed), float(narrow), float(bits), zp], dtype=np.float32)


def build_weight_quant(w, per_channel, bits=8, clamp_ste=True, min_val=1e-10):
    """The tree Int8WeightPerChannelFloat / Int8WeightPerTensorFloat resolve to."""
    from brevitas.core import function_wrapper as fw
    from brevitas.core.bit_width import BitWidthConst
    from brevitas.core.quant import IntQuant, RescalingIntQuant
    from brevitas.core.restrict_val import FloatRestrictValue
    from brevitas.core.scaling import IntScaling, StatsFromParameterScaling
    from brevitas.core.stats import AbsMax
    from brevitas.core.zero_point import ZeroZeroPoint
    if per_channel:
        stats, view, concat, shape = AbsMax(1), fw.OverOutputChannelView(None), 1, (w.shape[0],) + (1,) * (w.dim() - 1)
    else:
        stats, view, concat, shape = AbsMax(None), fw.OverTensorView(), 0, ()
    return RescalingIntQuant(
        IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(),
                 tensor_clamp_impl=fw.TensorClampSte() if clamp_ste else fw.TensorClamp()),
        StatsFromParameterScaling(stats, view, concat, [w], FloatRestrictValue(), shape, False, min_val),
        IntScaling(True, True), ZeroZeroPoint(), BitWidthConst(bits))


def gen_weight_stats(out):
    for dname, dt in DT.items():
        for per_channel in (True, False):
            for shape, seed, tag in [((16, 96), 31, "lin"), ((8, 3, 3, 3), 32, "conv")]:
                w0 = make_input(shape, seed, 0.2, with_edges=False)
                w0.view(-1)[5] = w0.abs().max() * 1.0          # no-op marker
                if tag == "lin":                               # ties: duplicate the row maximum, one all-zero row
                    w0[2, 7] = -w0[2].abs().max()
                    w0[2, 40] = w0[2].abs().max()
                    w0[3] = 0.0
                    m = w0.abs().max()
                    w0[0, 0], w0[9, 9] = m, -m                 # tensor-wide tie
                w = torch.nn.Parameter(w0.to(dt))
                tq = build_weight_quant(w, per_channel)
                y, scale, zp, bw = tq(w)
                g = make_input(shape, seed + 100, 1.0, False).to(dt)
                gs = (torch.rand(scale.shape, generator=torch.Generator().manual_seed(seed + 7)) - 0.5).to(scale.dtype)
                (y * g).sum().backward(retain_graph=True)
                gx_only = w.grad.clone()
                w.grad = None
                ((y * g).sum() + (scale * gs).sum()).backward()
                p = f"weight_stats/{'chan' if per_channel else 'tensor'}_{tag}/{dname}/"
                out[p + "w"], out[p + "g"], out[p + "gs"] = npf(w), npf(g), npf(gs)
                out[p + "y"], out[p + "scale"] = npf(y), npf(scale)
                out[p + "gw"], out[p + "gw_with_gscale"] = npf(gx_only), npf(w.grad)


def gen_runtime_token(out):
    """Per-token dynamic activation quantizer composed from core parts (Probe D.4)."""
    from brevitas.core import function_wrapper as fw
    from brevitas.core.bit_width import BitWidthConst
    from brevitas.core.quant import IntQuant, RescalingIntQuant
    from brevitas.core.restrict_val import FloatRestrictValue
    from brevitas.core.scaling import IntScaling, RuntimeStatsScaling
    from brevitas.core.stats import AbsMax
    from brevitas.core.zero_point import ZeroZeroPoint
    for dname, dt in DT.items():
        B, T, C = 2, 9, 64
        tq = RescalingIntQuant(
            IntQuant(narrow_range=False, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp()),
            RuntimeStatsScaling(AbsMax(2), fw.OverBatchOverOutputChannelView(), FloatRestrictValue(), (B, T, 1), False,
                                0.1, 1e-10),
            IntScaling(True, False), ZeroZeroPoint(), BitWidthConst(8))
        tq.train()
        p = f"runtime_token/{dname}/"
        for step in range(2):
            x = make_input((B, T, C), 41 + step, 1.5, with_edges=False).to(dt)
            x[0, 0, 3] = 200.0         # heavy tail
            xi = x.clone().requires_grad_(True)
            y, scale, zp, bw = tq(xi)
            g = make_input((B, T, C), 51 + step, 1.0, False).to(dt)
            y.backward(g)
            out[p + f"x{step}"], out[p + f"g{step}"] = npf(x), npf(g)
            out[p + f"y{step}"], out[p + f"scale{step}"], out[p + f"gx{step}"] = npf(y), npf(scale), npf(xi.grad)
            out[p + f"running{step}"] = npf(tq.scaling_impl.runtime_stats.running_stats)
        tq.eval()
        y, scale, _, _ = tq(x)
        out[p + "y_eval"], out[p + "scale_eval"] = npf(y), npf(scale)


def gen_binary(out):
    from brevitas.core.quant import BinaryQuant, ClampedBinaryQuant
    from brevitas.core.scaling import ConstScaling, ParameterScaling
    for dname, dt in DT.items():
        x = make_input((7, 33), 61, 0.6).to(dt)
        g = make_input((7, 33), 62, 1.0, False).to(dt)
        for qname, cls in (("binary", BinaryQuant), ("clamped", ClampedBinaryQuant)):
            for sname, simpl in (("const", ConstScaling(0.5)), ("param", ParameterScaling(0.5)),
                                 ("param_rows", ParameterScaling(torch.linspace(0.1, 0.9, 7).view(7, 1), (7, 1)))):
                simpl = simpl.to(dt)
                q = cls(scaling_impl=simpl)
                xi = x.clone().requires_grad_(True)
                y, scale, zp, bw = q(xi)
                y.backward(g)
                p = f"binary/{qname}_{sname}/{dname}/"
                out[p + "x"], out[p + "g"] = npf(x), npf(g)
                out[p + "y"], out[p + "scale"], out[p + "gx"] = npf(y), npf(scale), npf(xi.grad)
                if sname != "const":
                    out[p + "gvalue"] = npf(simpl.value.grad)
                    simpl.value.grad = None


def gen_percentile(out):
    from brevitas.core.stats import AbsPercentile
    for dname, dt in DT.items():
        x = make_input((1000,), 71, 2.0, with_edges=False).to(dt)
        for q in (99.999, 99.9, 90.0, 50.0, 1.0):
            out[f"percentile/flat_q{q}/{dname}/x"] = npf(x)
            out[f"percentile/flat_q{q}/{dname}/y"] = npf(AbsPercentile(q, None)(x))
        x2 = make_input((6, 250), 72, 2.0, with_edges=False).to(dt)
        out[f"percentile/rows_q99/{dname}/x"] = npf(x2)
        out[f"percentile/rows_q99/{dname}/y"] = npf(AbsPercentile(99.0, 1)(x2))
        out[f"percentile/cols_q99/{dname}/y"] = npf(AbsPercentile(99.0, 0)(x2))
    # the reference's own KATs (tests/brevitas/core/test_stats.py:12-28)
    v = torch.arange(1, 11).float()
    out["percentile/kat/x"] = npf(v)
    out["percentile/kat/y"] = np.array([float(AbsPercentile(10.0 * i, None)(v)) for i in range(1, 11)], dtype=np.float32)


def gen_param_from_stats(out):
    """Uint8ActPerTensorFloat's scaling through collection -> learned parameter."""
    from brevitas.core import function_wrapper as fw
    from brevitas.core.restrict_val import FloatRestrictValue
    from brevitas.core.scaling import ParameterFromRuntimeStatsScaling
    from brevitas.core.stats import AbsPercentile
    s = ParameterFromRuntimeStatsScaling(3, AbsPercentile(99.0, None), fw.OverTensorView(), (), FloatRestrictValue(),
                                         0.1, 1e-10)
    s.train()
    vals, bufs = [], []
    for step in range(6):
        x = torch.relu(make_input((4, 8, 6, 6), 81 + step, 1.0 + step, with_edges=False))
        out[f"param_from_stats/x{step}"] = npf(x)
        xi = x.clone().requires_grad_(True)
        t = s(xi)
        t.backward()
        vals.append(float(t.detach()))
        bufs.append(float(s.buffer))
        out[f"param_from_stats/gx{step}"] = npf(xi.grad if xi.grad is not None else torch.zeros_like(x))
        out[f"param_from_stats/gvalue{step}"] = npf(s.value.grad)
        s.value.grad = None
    out["param_from_stats/thresholds"] = np.array(vals, dtype=np.float32)
    out["param_from_stats/buffers"] = np.array(bufs, dtype=np.float32)
    out["param_from_stats/value"] = npf(s.value)


def gen_int_tables(out):
    from brevitas.function.ops import max_int, min_int
    rows = []
    for signed in (False, True):
        for narrow in (False, True):
            for bits in range(1, 17):
                bw = torch.tensor(float(bits))
                rows.append([float(signed), float(narrow), bits, float(min_int(signed, narrow, bw)),
                             float(max_int(signed, narrow, bw))])
    out["int_tables/rows"] = np.array(rows, dtype=np.float64)


def gen_docstring_kats(out):
    """Known answers quoted in the reference docstrings, regenerated by running the quoted code."""
    from brevitas.core.bit_width import BitWidthConst
    from brevitas.core.quant import BinaryQuant, ClampedBinaryQuant, IntQuant, RescalingIntQuant
    from brevitas.core.scaling import ConstScaling, IntScaling
    from brevitas.core.zero_point import ZeroZeroPoint
    inp = torch.tensor([0.042, -0.053, 0.31, -0.44])
    iq = IntQuant(narrow_range=True, signed=True)                                 # int_base.py:33-38
    out["kat/int_quant/x"] = npf(inp)
    out["kat/int_quant/y"] = npf(iq(torch.tensor(0.01), torch.tensor(0.), torch.tensor(4.), inp))
    rq = RescalingIntQuant(IntQuant(narrow_range=True, signed=True), ConstScaling(0.1),   # int.py:119-134
                           IntScaling(signed=True, narrow_range=True), ZeroZeroPoint(), BitWidthConst(4))
    y, s, z, b = rq(inp)
    out["kat/rescaling/y"], out["kat/rescaling/scale"] = npf(y), npf(s)
    bq = BinaryQuant(ConstScaling(0.1))                                            # binary.py:33-43
    binp = torch.tensor([0.04, -0.6, 3.3])
    out["kat/binary/x"] = npf(binp)
    out["kat/binary/y"] = npf(bq(binp)[0])
    cq = ClampedBinaryQuant(ConstScaling(0.1))                                     # binary.py:84-97
    ci = binp.clone().requires_grad_(True)
    cy = cq(ci)[0]
    cy.backward(torch.tensor([1.0, 1.0, 1.0]))
    out["kat/clamped_binary/y"], out["kat/clamped_binary/gx"] = npf(cy), npf(ci.grad)


def _stats_ops():
    from brevitas.core import stats as S
    return {
        "neg_min_or_zero": lambda dim: S.NegativeMinOrZero(dim),
        "neg_percentile_or_zero": lambda dim: S.NegativePercentileOrZero(10.0, dim),
        "percentile_interval": lambda dim: S.PercentileInterval(5.0, 95.0, dim),
        "abs_min_max": lambda dim: S.AbsMinMax(dim),
        "abs_max_ave": lambda dim: S.AbsMaxAve(1) if dim == 1 else None,
        "abs_max_l2": lambda dim: S.AbsMaxL2(1) if dim == 1 else None,
        "abs_ave": lambda dim: S.AbsAve(dim),
        "mean_sigma_std": lambda dim: S.MeanSigmaStd(3.0, dim),
    }


def gen_widen(out):
    """remaining statistics, bias / trunc / decoupled / ternary quantizers, asymmetric
    (zero-point) weight quantizers, learned bit-width."""
    from brevitas.core import function_wrapper as fw
    from brevitas.core.bit_width import BitWidthConst, BitWidthParameter, MsbClampBitWidth, RemoveBitwidthParameter
    from brevitas.core.quant import (DecoupledIntQuant, IntQuant, PrescaledRestrictIntQuant,
                                     PrescaledRestrictIntQuantWithInputBitWidth, RescalingIntQuant, TernaryQuant,
                                     TruncIntQuant)
    from brevitas.core.restrict_val import FloatRestrictValue
    from brevitas.core.scaling import IntScaling, ParameterScaling, StatsFromParameterScaling
    from brevitas.core.stats import AbsMinMax, NegativeMinOrZero
    from brevitas.core.zero_point import ParameterFromRuntimeZeroPoint, ParameterZeroPoint, StatsFromParameterZeroPoint
    # ---- A. statistics (values and autograd gradients) ----
    for dname, dt in DT.items():
        x2 = make_input((6, 40), 41, 2.0, with_edges=False).to(dt)
        x2[1, 3] = x2[1].max()            # a tie on one row maximum
        for name, make in _stats_ops().items():
            for dim in (None, 1):
                op = make(dim)
                if op is None:
                    continue
                xi = x2.clone().requires_grad_(True)
                y = op(xi if dim is not None else xi.reshape(-1))
                gsc = torch.linspace(0.5, 1.5, max(1, y.numel())).to(dt).view(y.shape)
                (y * gsc).sum().backward()
                k = f"widen/stats/{name}/{'rows' if dim == 1 else 'flat'}/{dname}/"
                out[k + "x"], out[k + "y"], out[k + "g"], out[k + "gx"] = npf(x2), npf(y), npf(gsc), npf(xi.grad)
    # ---- B. bias quantizers: externally supplied scale ----
    for dname, dt in DT.items():
        xb = make_input((48,), 51, 0.5, with_edges=False).to(dt)
        gb = make_input((48,), 52, 1.0, with_edges=False).to(dt)
        for sname, sc in (("scalar", torch.tensor(0.0173)), ("chan", torch.rand(48, generator=torch.Generator().manual_seed(5)) * 0.02 + 0.005)):
            iq = IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
            for qname, q, extra in (("prescaled", PrescaledRestrictIntQuant(iq, BitWidthConst(8)), ()),
                                    ("prescaled_in_bw", PrescaledRestrictIntQuantWithInputBitWidth(
                                        iq, MsbClampBitWidth(RemoveBitwidthParameter(3), 2, 16)), (torch.tensor(9.0),))):
                xi = xb.clone().requires_grad_(True)
                si = sc.to(dt).clone().requires_grad_(True)
                y, s_out, zp, bw = q(xi, si, *extra)
                (y * gb).sum().backward()
                k = f"widen/{qname}/{sname}/{dname}/"
                out[k + "x"], out[k + "scale"], out[k + "g"] = npf(xb), npf(sc.to(dt)), npf(gb)
                out[k + "y"], out[k + "bit_width"] = npf(y), npf(bw)
                out[k + "gx"], out[k + "gscale"] = npf(xi.grad), npf(si.grad)
    # docstring KAT (int.py:33-47)
    iq = IntQuant(narrow_range=True, signed=True)
    kat = PrescaledRestrictIntQuantWithInputBitWidth(iq, fw.Identity())
    yk, _, _, bwk = kat(torch.tensor([0.042, -0.053, 0.31, -0.44]), torch.tensor(0.01), torch.tensor(4.))
    out["widen/kat/prescaled/y"], out["widen/kat/prescaled/bw"] = npf(yk), npf(bwk)
    # ---- C. TruncIntQuant (QuantAvgPool2d): 12-bit accumulator values truncated to 8 bits ----
    for dname, dt in DT.items():
        sc = torch.tensor(0.03125).to(dt)                        # power of two: x = codes * scale is exact in T
        codes = torch.randint(-90, 91, (5, 33), generator=torch.Generator().manual_seed(61)).to(dt)
        xt = (codes * sc).clone().requires_grad_(True)
        tq = TruncIntQuant(fw.FloorSte(), BitWidthConst(4))
        y, s_out, zp, bw = tq(xt, sc, torch.tensor(0.0).to(dt), torch.tensor(8.0))
        gt = make_input((5, 33), 62, 1.0, False).to(dt)
        (y * gt).sum().backward()
        k = f"widen/trunc/{dname}/"
        out[k + "x"], out[k + "scale"], out[k + "g"] = npf(xt), npf(sc), npf(gt)
        out[k + "y"], out[k + "bit_width"], out[k + "gx"] = npf(y), npf(bw), npf(xt.grad)
    # ---- D. DecoupledIntQuant ----
    dq = DecoupledIntQuant(narrow_range=True, signed=True)
    yk = dq(torch.tensor(0.02), torch.tensor(0.), torch.tensor(0.01), torch.tensor(0.), torch.tensor(4.),
            torch.tensor([0.042, -0.053, 0.31, -0.44]))           # docstring KAT (int_base.py:118-125)
    out["widen/kat/decoupled/y"] = npf(yk)
    for dname, dt in DT.items():
        xd = make_input((7, 29), 71, 3.0, with_edges=False).to(dt)
        gd = make_input((7, 29), 72, 1.0, False).to(dt)
        xi = xd.clone().requires_grad_(True)
        ps = torch.tensor(0.05).to(dt).requires_grad_(True)
        sc = torch.tensor(0.047).to(dt).requires_grad_(True)
        dq = DecoupledIntQuant(narrow_range=False, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
        y = dq(ps, torch.tensor(0.).to(dt), sc, torch.tensor(0.).to(dt), torch.tensor(6.), xi)
        (y * gd).sum().backward()
        k = f"widen/decoupled/{dname}/"
        out[k + "x"], out[k + "g"], out[k + "y"], out[k + "gx"] = npf(xd), npf(gd), npf(y), npf(xi.grad)
        out[k + "g_pre_scale"], out[k + "g_scale"] = npf(ps.grad), npf(sc.grad)
    # ---- E. TernaryQuant ----
    tk = TernaryQuant(ParameterScaling(1.0), 0.5)
    out["widen/kat/ternary/y"] = npf(tk(torch.tensor([0.04, -0.6, 3.3]))[0])       # docstring KAT (ternary.py:34-38)
    for dname, dt in DT.items():
        xq = make_input((9, 31), 81, 1.0, with_edges=True).to(dt)
        gq = make_input((9, 31), 82, 1.0, False).to(dt)
        tq = TernaryQuant(ParameterScaling(0.7), 0.5).to(dt)
        xi = xq.clone().requires_grad_(True)
        y, s_out, zp, bw = tq(xi)
        (torch.nan_to_num(y) * gq).sum().backward()
        k = f"widen/ternary/{dname}/"
        out[k + "x"], out[k + "g"], out[k + "y"], out[k + "gx"] = npf(xq), npf(gq), npf(y), npf(xi.grad)
        out[k + "gvalue"] = npf(tq.scaling_impl.value.grad)
    # ---- F. asymmetric weight quantizers: ShiftedUint8WeightPer{Tensor,Channel}Float wiring
    #         (quant/shifted_scaled_int.py:45-75 = ShiftedMinUintQuant + MinMaxStatsScaling) ----
    for dname, dt in DT.items():
        for per_channel in (False, True):
            w = torch.nn.Parameter((make_input((12, 50), 91, 0.3, with_edges=False) + 0.05).to(dt))
            if per_channel:
                view, concat, shape, dim = fw.OverOutputChannelView(None), 1, (12, 1), 1
            else:
                view, concat, shape, dim = fw.OverTensorView(), 0, (), None
            iq = IntQuant(narrow_range=False, signed=False, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClampSte())
            tq = RescalingIntQuant(
                iq, StatsFromParameterScaling(AbsMinMax(dim), view, concat, [w], FloatRestrictValue(), shape, False, 1e-10),
                IntScaling(False, False),
                StatsFromParameterZeroPoint(iq, True, view, concat, NegativeMinOrZero(dim), shape, [w]),
                BitWidthConst(8))
            y, scale, zp, bw = tq(w)
            gw = make_input((12, 50), 92, 1.0, False).to(dt)
            (y * gw).sum().backward()
            k = f"widen/shifted_weight/{'chan' if per_channel else 'tensor'}/{dname}/"
            out[k + "w"], out[k + "g"], out[k + "y"] = npf(w), npf(gw), npf(y)
            out[k + "scale"], out[k + "zero_point"], out[k + "gw"] = npf(scale), npf(zp), npf(w.grad)
    # ---- G. learned bit-width ----
    bwp = BitWidthParameter(6, min_bit_width=2)
    v = bwp()
    (v * 2.5).backward()
    out["widen/bit_width_parameter/value"], out["widen/bit_width_parameter/g_offset"] = npf(v), npf(bwp.bit_width_offset.grad)
    rbp = RemoveBitwidthParameter(3)
    out["widen/remove_bit_width/value"] = npf(rbp())
    out["widen/msb_clamp/value"] = npf(MsbClampBitWidth(RemoveBitwidthParameter(3), 2, 16)(torch.tensor(24.0)))
    # ---- H. zero-point from runtime statistics, then learned (3 collection steps + steady state + eval) ----
    iq = IntQuant(narrow_range=False, signed=False, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
    zpm = ParameterFromRuntimeZeroPoint(3, iq, True, NegativeMinOrZero(None), (), fw.OverTensorView(), 0.1)
    zpm.train()
    sc8, bw8 = torch.tensor(0.04), torch.tensor(8.0)
    for step in range(5):
        xa = make_input((4, 25), 100 + step, 1.0, False) - 0.3
        out[f"widen/runtime_zero_point/x{step}"] = npf(xa)
        out[f"widen/runtime_zero_point/zp{step}"] = npf(zpm(xa, sc8, bw8))
        out[f"widen/runtime_zero_point/buffer{step}"] = npf(zpm.buffer)
        out[f"widen/runtime_zero_point/value{step}"] = npf(zpm.value)
    zpm.eval()
    out["widen/runtime_zero_point/zp_eval"] = npf(zpm(xa, sc8, bw8))
    pz = ParameterZeroPoint(-0.37, iq, True, None)
    out["widen/parameter_zero_point/zp"] = npf(pz(xa, sc8, bw8))


def gen_shifted_act(out):
    """ShiftedUint8ActPerTensorFloat wiring (quant/shifted_scaled_int.py:19-42 = ShiftedParamFromPercentileUintQuant +
    ParamFromRuntimePercentileIntervalScaling): scale and zero-point are runtime statistics for `collect` steps, then
    learned parameters; masked clamp.  3 collection steps, 2 learned steps (gradients reach both parameters), eval."""
    from brevitas.core import function_wrapper as fw
    from brevitas.core.bit_width import BitWidthConst
    from brevitas.core.quant import IntQuant, RescalingIntQuant
    from brevitas.core.restrict_val import FloatRestrictValue
    from brevitas.core.scaling import IntScaling, ParameterFromRuntimeStatsScaling
    from brevitas.core.stats import NegativePercentileOrZero, PercentileInterval
    from brevitas.core.zero_point import ParameterFromRuntimeZeroPoint
    collect = 3
    iq = IntQuant(narrow_range=False, signed=False, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
    tq = RescalingIntQuant(
        iq, ParameterFromRuntimeStatsScaling(collect, PercentileInterval(5.0, 95.0, None), fw.OverTensorView(), (),
                                             FloatRestrictValue(), 0.1, 1e-10),
        IntScaling(False, False),
        ParameterFromRuntimeZeroPoint(collect, iq, True, NegativePercentileOrZero(5.0, None), (), fw.OverTensorView(), 0.1),
        BitWidthConst(8))
    tq.train()
    for step in range(6):
        if step == 5:
            tq.eval()
        x = (make_input((6, 40), 200 + step, 1.5, False) + 0.4).requires_grad_(True)
        g = make_input((6, 40), 300 + step, 1.0, False)
        y, scale, zp, bw = tq(x)
        for prm in (tq.scaling_impl.value, tq.zero_point_impl.value):
            prm.grad = None
        (y * g).sum().backward()
        k = f"shifted_act/step{step}/"
        out[k + "x"], out[k + "g"], out[k + "y"], out[k + "gx"] = npf(x), npf(g), npf(y), npf(x.grad)
        out[k + "scale"], out[k + "zero_point"] = npf(scale), npf(zp)
        gs, gz = tq.scaling_impl.value.grad, tq.zero_point_impl.value.grad
        out[k + "g_scale_value"] = npf(gs) if gs is not None else np.zeros((), np.float32)
        out[k + "g_zp_value"] = npf(gz) if gz is not None else np.zeros((), np.float32)
        out[k + "scale_value"], out[k + "zp_value"] = npf(tq.scaling_impl.value), npf(tq.zero_point_impl.value)


def main():
    import_reference()
    torch.manual_seed(123456)
    groups = {"ste": gen_ste, "int_quant": gen_int_quant, "weight_stats": gen_weight_stats,
              "runtime_token": gen_runtime_token, "binary": gen_binary, "percentile": gen_percentile,
              "param_from_stats": gen_param_from_stats, "int_tables": gen_int_tables, "kat": gen_docstring_kats,
              "widen": gen_widen, "shifted_act": gen_shifted_act}
    only = sys.argv[1:]
    if only:
        groups = {k: v for k, v in groups.items() if k in only}
    for gname, fn in groups.items():
        out = {}
        fn(out)
        path = os.path.join(HERE, f"{gname}.npz")
        np.savez_compressed(path, **out)
        print(f"{gname}: {len(out)} arrays -> {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


if __name__ == "__main__":
    main()
