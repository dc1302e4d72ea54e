"""GPU (-m gpu): the quantizers / statistics either side of the hot path against golden
vectors produced by the real reference (tests/golden/widen.npz, make_golden.py::gen_widen).

Element-wise results (outputs, integer codes, zero-points, element-wise gradients) are bit-exact.  Values that are
floating-point REDUCTIONS in the reference as well (mean, variance, norms, the gradient arriving on a scale or on a
tied maximum) depend on the summation order of the device and carry the tolerance written at the check."""
import numpy as np
import pytest
import torch

from golden_util import DTYPES, assert_bits_equal, case, load, ulp

pytestmark = pytest.mark.gpu
TDT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}


def dev(a, dtype):
    a = np.asarray(a)
    return torch.from_numpy(np.array(a, copy=True)).reshape(a.shape).to(TDT[dtype]).cuda()      # keeps 0-dim


def host(t):
    return t.detach().float().cpu().numpy()


def close(got, ref, dtype, k=8.0, what=""):
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    tol = k * ulp(dtype) * (np.abs(ref) + np.abs(ref).max() * 0.05 + 1e-6)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    assert np.all(np.abs(got - ref) <= tol), (what, got.reshape(-1)[:6], ref.reshape(-1)[:6])


EXACT_STATS = {"neg_min_or_zero", "neg_percentile_or_zero", "percentile_interval", "abs_min_max"}   # selections only


def _make_stat(name, dim):
    from brevitas_b200.core import stats as S
    return {"neg_min_or_zero": lambda: S.NegativeMinOrZero(dim),
            "neg_percentile_or_zero": lambda: S.NegativePercentileOrZero(10.0, dim),
            "percentile_interval": lambda: S.PercentileInterval(5.0, 95.0, dim),
            "abs_min_max": lambda: S.AbsMinMax(dim), "abs_max_ave": lambda: S.AbsMaxAve(1),
            "abs_max_l2": lambda: S.AbsMaxL2(1), "abs_ave": lambda: S.AbsAve(dim),
            "mean_sigma_std": lambda: S.MeanSigmaStd(3.0, dim)}[name]()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["neg_min_or_zero", "neg_percentile_or_zero", "percentile_interval", "abs_min_max",
                                  "abs_max_ave", "abs_max_l2", "abs_ave", "mean_sigma_std"])
def test_statistics(name, dtype):
    for layout, dim in (("flat", None), ("rows", 1)):
        c = case("widen", f"widen/stats/{name}/{layout}/{dtype}/")
        if not c:
            assert dim is None and name in ("abs_max_ave", "abs_max_l2")
            continue
        op = _make_stat(name, dim).cuda()
        x = dev(c["x"], dtype).requires_grad_(True)
        y = op(x if dim is not None else x.reshape(-1))
        (y * dev(c["g"], dtype).view(y.shape)).sum().backward()
        if name in EXACT_STATS:
            assert_bits_equal(host(y), c["y"], f"{name} {layout}")
            # the selected VALUE is unique but, among equal elements (bf16 / fp16 inputs repeat values), which INDEX
            # min / max / kthvalue reports differs between ATen's CPU and CUDA kernels: compare the gradient summed
            # over elements of equal value, per row
            got, ref, xs = host(x.grad), c["gx"], c["x"]
            if dim is None:
                got, ref, xs = got.reshape(1, -1), ref.reshape(1, -1), xs.reshape(1, -1)
            for r in range(xs.shape[0]):
                for val in np.unique(xs[r][(got[r] != 0) | (ref[r] != 0)]):
                    sel = xs[r] == val
                    assert got[r][sel].sum() == ref[r][sel].sum(), (name, layout, r, val)
            assert np.count_nonzero(got) == np.count_nonzero(ref)
        else:
            close(host(y), c["y"], dtype, what=f"{name} {layout}")
            close(host(x.grad), c["gx"], dtype, k=16.0, what=f"{name} {layout} grad")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("sname", ["scalar", "chan"])
@pytest.mark.parametrize("qname", ["prescaled", "prescaled_in_bw"])
def test_bias_quantizers(qname, sname, dtype):
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst, MsbClampBitWidth, RemoveBitwidthParameter
    from brevitas_b200.core.quant import IntQuant, PrescaledRestrictIntQuant, PrescaledRestrictIntQuantWithInputBitWidth
    c = case("widen", f"widen/{qname}/{sname}/{dtype}/")
    iq = IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
    if qname == "prescaled":
        q, extra = PrescaledRestrictIntQuant(iq, BitWidthConst(8)).cuda(), ()
    else:
        q = PrescaledRestrictIntQuantWithInputBitWidth(iq, MsbClampBitWidth(RemoveBitwidthParameter(3), 2, 16)).cuda()
        extra = (torch.tensor(9.0, device="cuda"),)
    x = dev(c["x"], dtype).requires_grad_(True)
    s = dev(c["scale"], dtype).requires_grad_(True)
    y, s_out, zp, bw = q(x, s, *extra)
    (y * dev(c["g"], dtype)).sum().backward()
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(bw), c["bit_width"], "bit width")
    assert float(zp) == 0.0 and s_out is s
    assert_bits_equal(host(x.grad), c["gx"], "gx")
    # d(scale) = sum g*q - d*((x/s)/s): the reference rounds every product (and, for one scale, every partial sum) to
    # the tensor dtype, the fused kernel accumulates in fp32 -- a difference of two terms of magnitude |g*x/s| each
    # carrying one rounding, hence the tolerance relative to that magnitude (contract: DESIGN.md §2)
    term = np.abs(c["g"]) * (np.abs(c["x"]) / c["scale"] + 1.0)
    mag = term if sname == "chan" else term.sum()
    got, ref = host(s.grad).astype(np.float64), c["gscale"].astype(np.float64)
    assert got.shape == ref.shape
    assert np.all(np.abs(got - ref) <= 4 * ulp(dtype) * mag + 1e-6), (got.reshape(-1)[:4], ref.reshape(-1)[:4])


def test_docstring_kats():
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import DecoupledIntQuant, IntQuant, PrescaledRestrictIntQuantWithInputBitWidth, TernaryQuant
    from brevitas_b200.core.scaling import ParameterScaling
    d = load("widen")
    t = lambda v: torch.tensor(v, device="cuda")
    q = PrescaledRestrictIntQuantWithInputBitWidth(IntQuant(narrow_range=True, signed=True), fw.Identity()).cuda()
    y, _, _, bw = q(t([0.042, -0.053, 0.31, -0.44]), t(0.01), t(4.))
    assert_bits_equal(host(y), d["widen/kat/prescaled/y"], "int.py:33-47")
    assert float(bw) == float(d["widen/kat/prescaled/bw"])
    dq = DecoupledIntQuant(narrow_range=True, signed=True).cuda()
    y = dq(t(0.02), t(0.), t(0.01), t(0.), t(4.), t([0.042, -0.053, 0.31, -0.44]))
    assert_bits_equal(host(y), d["widen/kat/decoupled/y"], "int_base.py:118-125")
    tq = TernaryQuant(ParameterScaling(1.0), 0.5).cuda()
    assert_bits_equal(host(tq(t([0.04, -0.6, 3.3]))[0]), d["widen/kat/ternary/y"], "ternary.py:34-38")


@pytest.mark.parametrize("dtype", DTYPES)
def test_trunc_int_quant(dtype):
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import TruncIntQuant
    c = case("widen", f"widen/trunc/{dtype}/")
    tq = TruncIntQuant(fw.FloorSte(), BitWidthConst(4)).cuda()
    x = dev(c["x"], dtype).requires_grad_(True)
    y, s, zp, bw = tq(x, dev(c["scale"], dtype), torch.tensor(0.0, device="cuda").to(TDT[dtype]), torch.tensor(8.0, device="cuda"))
    (y * dev(c["g"], dtype)).sum().backward()
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(bw), c["bit_width"], "bit width")
    assert_bits_equal(host(x.grad), c["gx"], "gx")


@pytest.mark.parametrize("dtype", DTYPES)
def test_decoupled_int_quant(dtype):
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import DecoupledIntQuant
    c = case("widen", f"widen/decoupled/{dtype}/")
    T = TDT[dtype]
    dq = DecoupledIntQuant(narrow_range=False, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp()).cuda()
    x = dev(c["x"], dtype).requires_grad_(True)
    ps = torch.tensor(0.05).to(T).cuda().requires_grad_(True)
    sc = torch.tensor(0.047).to(T).cuda().requires_grad_(True)
    z = torch.tensor(0.).to(T).cuda()
    y = dq(ps, z, sc, z, torch.tensor(6., device="cuda"), x)
    (y * dev(c["g"], dtype)).sum().backward()
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(x.grad), c["gx"], "gx")
    close(host(ps.grad), c["g_pre_scale"], dtype, k=32.0, what="d pre_scale (sum)")
    close(host(sc.grad), c["g_scale"], dtype, k=32.0, what="d scale (sum)")


@pytest.mark.parametrize("dtype", DTYPES)
def test_ternary_quant(dtype):
    from brevitas_b200.core.quant import TernaryQuant
    from brevitas_b200.core.scaling import ParameterScaling
    c = case("widen", f"widen/ternary/{dtype}/")
    tq = TernaryQuant(ParameterScaling(0.7), 0.5).to(TDT[dtype]).cuda()
    x = dev(c["x"], dtype).requires_grad_(True)
    y, s, zp, bw = tq(x)
    (torch.nan_to_num(y) * dev(c["g"], dtype)).sum().backward()
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(x.grad), c["gx"], "gx")
    assert float(bw) == 2.0 and float(zp) == 0.0
    close(host(tq.scaling_impl.value.grad), c["gvalue"], dtype, k=32.0, what="d scale (sum)")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("layout", ["tensor", "chan"])
def test_shifted_uint8_weight_quantizers(layout, dtype):
    """ShiftedUint8WeightPer{Tensor,Channel}Float (quant/shifted_scaled_int.py:45-75): scale from |max - min|,
    integer zero-point from -min through the quantizer's own to_int."""
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthConst
    from brevitas_b200.core.quant import IntQuant, RescalingIntQuant
    from brevitas_b200.core.restrict_val import FloatRestrictValue
    from brevitas_b200.core.scaling import IntScaling, StatsFromParameterScaling
    from brevitas_b200.core.stats import AbsMinMax, NegativeMinOrZero
    from brevitas_b200.core.zero_point import StatsFromParameterZeroPoint
    c = case("widen", f"widen/shifted_weight/{layout}/{dtype}/")
    w = torch.nn.Parameter(dev(c["w"], dtype))
    if layout == "chan":
        view, concat, shape, dim = fw.OverOutputChannelView(None), 1, (12, 1), 1
    else:
        view, concat, shape, dim = fw.OverTensorView(), 0, (), None
    iq = IntQuant(narrow_range=False, signed=False, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClampSte())
    tq = RescalingIntQuant(
        iq, StatsFromParameterScaling(AbsMinMax(dim), view, concat, [w], FloatRestrictValue(), shape, False, 1e-10),
        IntScaling(False, False), StatsFromParameterZeroPoint(iq, True, view, concat, NegativeMinOrZero(dim), shape, [w]),
        BitWidthConst(8)).cuda()
    y, scale, zp, bw = tq(w)
    (y * dev(c["g"], dtype)).sum().backward()
    assert_bits_equal(host(scale), c["scale"], "scale")
    assert_bits_equal(host(zp), c["zero_point"], "zero point")
    assert_bits_equal(host(y), c["y"], "y")
    # the gradient is element-wise except on each region's min / max entries, which also receive the reduced
    # d(scale) and d(zero_point) terms (order-dependent sums)
    wn = c["w"]
    if layout == "chan":
        ext = (wn == wn.max(axis=1, keepdims=True)) | (wn == wn.min(axis=1, keepdims=True))
    else:
        ext = (wn == wn.max()) | (wn == wn.min())
    got = host(w.grad)
    assert_bits_equal(np.where(ext, 0, got), np.where(ext, 0, c["gw"]), "gw off the extrema")
    close(got[ext], c["gw"][ext], dtype, k=64.0, what="gw on the extrema")


def test_learned_bit_width_and_zero_points():
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.bit_width import BitWidthParameter, MsbClampBitWidth, RemoveBitwidthParameter
    from brevitas_b200.core.quant import IntQuant
    from brevitas_b200.core.stats import NegativeMinOrZero
    from brevitas_b200.core.zero_point import ParameterFromRuntimeZeroPoint, ParameterZeroPoint
    d = load("widen")
    bwp = BitWidthParameter(6, min_bit_width=2).cuda()
    v = bwp()
    (v * 2.5).backward()
    assert_bits_equal(host(v), d["widen/bit_width_parameter/value"], "learned bit width")
    assert_bits_equal(host(bwp.bit_width_offset.grad), d["widen/bit_width_parameter/g_offset"], "d offset")
    assert_bits_equal(host(RemoveBitwidthParameter(3).cuda()()), d["widen/remove_bit_width/value"], "bits to remove")
    assert_bits_equal(host(MsbClampBitWidth(RemoveBitwidthParameter(3), 2, 16).cuda()(torch.tensor(24.0, device="cuda"))),
                      d["widen/msb_clamp/value"], "msb clamp")
    iq = IntQuant(narrow_range=False, signed=False, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp())
    zpm = ParameterFromRuntimeZeroPoint(3, iq, True, NegativeMinOrZero(None), (), fw.OverTensorView(), 0.1).cuda()
    zpm.train()
    sc8, bw8 = torch.tensor(0.04, device="cuda"), torch.tensor(8.0, device="cuda")
    for step in range(5):
        xa = dev(d[f"widen/runtime_zero_point/x{step}"], "f32")
        assert_bits_equal(host(zpm(xa, sc8, bw8)), d[f"widen/runtime_zero_point/zp{step}"], f"zero point, step {step}")
        assert_bits_equal(host(zpm.buffer), d[f"widen/runtime_zero_point/buffer{step}"], f"buffer, step {step}")
        assert_bits_equal(host(zpm.value), d[f"widen/runtime_zero_point/value{step}"], f"value, step {step}")
    zpm.eval()
    assert_bits_equal(host(zpm(xa, sc8, bw8)), d["widen/runtime_zero_point/zp_eval"], "eval")
    pz = ParameterZeroPoint(-0.37, iq, True, None).cuda()
    assert_bits_equal(host(pz(xa, sc8, bw8)), d["widen/parameter_zero_point/zp"], "learned zero point")


def test_shifted_uint8_act_quantizer_collect_then_learn():
    """ShiftedUint8ActPerTensorFloat (quant/shifted_scaled_int.py:19-42): 3 collection steps, 2 learned steps, eval.  The
    tensor goes through the zero-point-operand kernel; from step 3 on its d(scale) / d(zero_point) reach the two
    learned parameters through the reference's own tiny ops."""
    from brevitas_b200.quant import ShiftedUint8ActPerTensorFloat
    d = load("shifted_act")
    tq = ShiftedUint8ActPerTensorFloat.let(collect_stats_steps=3, high_percentile_q=95.0, low_percentile_q=5.0).tensor_quant().cuda()
    tq.train()
    for step in range(6):
        if step == 5:
            tq.eval()
        k = f"shifted_act/step{step}/"
        x = dev(d[k + "x"], "f32").requires_grad_(True)
        y, scale, zp, bw = tq(x)
        for prm in (tq.scaling_impl.value, tq.zero_point_impl.value):
            prm.grad = None
        (y * dev(d[k + "g"], "f32")).sum().backward()
        assert_bits_equal(host(scale), d[k + "scale"], f"scale, step {step}")
        assert_bits_equal(host(zp), d[k + "zero_point"], f"zero point, step {step}")
        assert_bits_equal(host(y), d[k + "y"], f"y, step {step}")
        assert_bits_equal(host(tq.scaling_impl.value), d[k + "scale_value"], f"scale parameter, step {step}")
        assert_bits_equal(host(tq.zero_point_impl.value), d[k + "zp_value"], f"zero-point parameter, step {step}")
        if step < 3:
            # collecting: the percentile statistics route gradient to single elements of x (tie-free here)
            close(host(x.grad), d[k + "gx"], "f32", k=64.0, what=f"gx, step {step}")
        else:
            assert_bits_equal(host(x.grad), d[k + "gx"], f"gx, step {step}")
            gs, gz = tq.scaling_impl.value.grad, tq.zero_point_impl.value.grad
            xs = np.abs(d[k + "g"]) * (np.abs(d[k + "x"]) / d[k + "scale"] + 256.0)
            tol = 1e-5 * float(xs.sum()) + 1e-4
            assert abs(float(gs) - float(d[k + "g_scale_value"])) <= tol, (step, float(gs), d[k + "g_scale_value"])
            assert abs(float(gz) - float(d[k + "g_zp_value"])) <= tol, (step, float(gz), d[k + "g_zp_value"])


def test_layer_output_zero_point_follows_the_reference_bias_rule():
    """nn/quant_layer.py:337-355: a bias that is not quantized at the accumulator scale becomes the output
    zero-point -bias / (weight_scale * input_scale); non-zero input / weight zero-points are refused."""
    from brevitas_b200.nn import QuantIdentity, QuantLinear
    from brevitas_b200.quant import Int8ActPerTensorFloat, Int8Bias, ShiftedUint8ActPerTensorFloat
    torch.manual_seed(0)
    x = torch.randn(4, 16, device="cuda")
    qin = QuantIdentity(act_quant=Int8ActPerTensorFloat, return_quant_tensor=True).cuda().train()
    lin = QuantLinear(16, 8, bias=True, return_quant_tensor=True).cuda().train()            # float bias
    out = lin(qin(x))
    qi = qin(x)
    acc_scale = lin.quant_weight().scale.view(1, -1) * qi.scale.view(1, -1)
    assert torch.equal(out.scale, acc_scale)
    assert torch.equal(out.zero_point, -lin.bias.view(1, -1) / acc_scale)
    # (value / scale + zero_point) is then the integer accumulator WITHOUT the bias
    acc = out.value / out.scale + out.zero_point
    ref = torch.nn.functional.linear(qi.value, lin.quant_weight().value) / acc_scale
    assert torch.allclose(acc, ref, atol=1e-2)
    lin_q = QuantLinear(16, 8, bias=True, bias_quant=Int8Bias, return_quant_tensor=True).cuda().train()
    out_q = lin_q(qin(x))                              # bias quantized AT the accumulator scale: no shift
    assert torch.equal(out_q.zero_point, qi.zero_point)
    shifted = QuantIdentity(act_quant=ShiftedUint8ActPerTensorFloat, return_quant_tensor=True, collect_stats_steps=1).cuda().train()
    shifted(x)
    shifted(x)
    with pytest.raises(RuntimeError, match="zero point of output accumulator"):
        lin(shifted(x - 3.0))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols", [(1, 1), (1, 1000), (1, 1 << 20), (6, 250), (3, 4099), (2, 300001)])
def test_signed_kth_value_select(rows, cols, dtype):
    """bvb_kth_value_rows (NegativePercentileOrZero / PercentileInterval, stats_op.py:69-126) against torch.kthvalue:
    exact value for every regime of k, the smallest index attaining it, NaN sorted last, both zeros, infinities."""
    from brevitas_b200 import _kernels as K
    g = torch.Generator().manual_seed(rows * 7919 + cols)
    x = (torch.randn(rows, cols, generator=g) * 3.0)
    if cols >= 250:
        x[:, 3], x[:, 5], x[:, 7], x[:, 11], x[:, 13] = 0.0, -0.0, float("inf"), float("-inf"), 2.5
        x[:, 17] = 2.5                                   # a tie
    xd = x.to(TDT[dtype]).cuda()
    xf = xd.float().cpu()
    for k in sorted({1, min(2, cols), max(1, cols // 100), max(1, cols // 2), max(1, cols - cols // 20), cols}):
        val, idx = K.kth_value_rows(xd, rows, cols, k, want_index=True)
        want = xf.kthvalue(k, dim=1).values
        got = val.float().cpu()
        assert torch.equal(got, want) or torch.equal(got.abs(), want.abs()) and bool(((got == 0) == (want == 0)).all()), (k, got, want)
        picked = xf.gather(1, idx.cpu().view(rows, 1)).view(-1)
        assert torch.equal(picked, got), (k, picked, got)
        first = (xf == got.view(rows, 1)).float().argmax(dim=1)
        assert torch.equal(first, idx.cpu()), (k, first, idx)
    if cols >= 250:                                      # NaN is the largest element
        xn = xd.clone()
        xn[:, 19] = float("nan")
        val, _ = K.kth_value_rows(xn, rows, cols, cols)
        assert bool(torch.isnan(val).all())
        val, _ = K.kth_value_rows(xn, rows, cols, cols - 1)
        assert bool(torch.isinf(val).all() and (val > 0).all())


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,bshape", [((5, 67), ()), ((1 << 16,), ()), ((6, 48), (6, 1)), ((3, 5, 4, 8), (1, 5, 1, 1)),
                                          ((7, 9), (7, 9))])
def test_differentiable_tensor_clamp_kernel(shape, bshape, dtype):
    """brevitas.function.ops.tensor_clamp (function/ops.py:76-100) -- the default clamp of IntQuant, the path a learned
    bit-width takes its gradient through: forward and ALL THREE gradients against the reference's two torch.where calls
    under plain autograd (ATen)."""
    from brevitas_b200.function.ops import tensor_clamp
    g = torch.Generator().manual_seed(len(shape) * 31 + len(bshape))
    x = (torch.randn(shape, generator=g) * 3).to(TDT[dtype])
    x.view(-1)[:6] = torch.tensor([float("nan"), float("inf"), float("-inf"), 0.0, -0.0, 2.0]).to(TDT[dtype])[:min(6, x.numel())]
    lo = (-torch.rand(bshape, generator=g) - 0.5).to(TDT[dtype])
    hi = (torch.rand(bshape, generator=g) + 0.5).to(TDT[dtype])
    gy = torch.randn(shape, generator=g).to(TDT[dtype])

    def run(fn, dev):
        xi, li, hi_ = (t.clone().to(dev).requires_grad_(True) for t in (x, lo, hi))
        y = fn(xi, li, hi_)
        y.backward(gy.to(dev))
        return [t.detach().float().cpu() for t in (y, xi.grad, li.grad, hi_.grad)]

    def ref(a, mn, mx):
        out = torch.where(a > mx, mx.type_as(a), a)
        return torch.where(out < mn, mn.type_as(out), out)
    got = run(tensor_clamp, "cuda")
    want = run(ref, "cpu")
    assert_bits_equal(got[0].numpy(), want[0].numpy(), "y")
    assert_bits_equal(got[1].numpy(), want[1].numpy(), "gx")
    for a, b, name in ((got[2], want[2], "gmin"), (got[3], want[3], "gmax")):
        mag = float(gy.float().abs().sum()) / max(1, b.numel()) + 1.0
        assert torch.allclose(a, b, rtol=0, atol=mag * (8 * np.sqrt(x.numel()) * 2.0 ** -24 + 4 * ulp(dtype))), (name, a, b)
