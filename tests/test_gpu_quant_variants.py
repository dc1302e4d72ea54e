"""GPU (-m gpu): the one-kernel forms of the remaining quantizer flavours (csrc/quant_variants.cu)
against the LITERAL op sequence of the reference on the same device:

* ``general_int_quant`` -- DecoupledIntQuant (int_base.py:100-182) and IntQuant with a device-resident range (learned
  bit-width, core/bit_width/parameter.py:23-98): outputs and dx bit for bit; d(pre_scale), d(scale), d(min_int),
  d(max_int) are sums and carry the tolerance written at the check;
* ``ternary_quant`` -- TernaryQuant (ternary.py:58-72);
* the learned bit-width through the reference's own layers (injector, proxies) after ``install()`` vs the pure reference.
"""
import math

import pytest
import torch

from ref_util import reference_src

pytestmark = pytest.mark.gpu

DT = [torch.float32, torch.bfloat16, torch.float16]


def bits(t):
    t = t.detach().contiguous()
    return t.view(torch.int32 if t.dtype == torch.float32 else torch.int16)


def same_bits(a, b, what):
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    eq = (bits(a) == bits(b)) | (torch.isnan(a) & torch.isnan(b))
    assert bool(eq.all()), f"{what}: {int((~eq).sum())} of {a.numel()} differ"


def close_sum(got, ref, terms_abs_sum, n, dtype, what):
    """two sums of the same n terms accumulated in different orders (and, for 16-bit tensors, of terms rounded to 16 bits
    one by one): each is within 8 sqrt(n) eps mean|term| of the exact sum, so they are within twice that of each other,
    plus the storage rounding of the result"""
    eps = {torch.float32: 2.0 ** -24, torch.bfloat16: 2.0 ** -9, torch.float16: 2.0 ** -11}[dtype]
    tol = 16.0 * math.sqrt(max(n, 1)) * eps * (terms_abs_sum / max(n, 1)) + 4 * eps * abs(float(ref)) + 1e-30
    assert abs(float(got) - float(ref)) <= tol, (what, float(got), float(ref), tol)


def literal(x, ps, s, pzp, zp, lo, hi, round_impl, clamp_impl):
    """int_base.py:132-182 op by op (division, +zp, float_to_int, clamp, -zp, *scale) on ATen + the STE kernels"""
    y = x / ps
    y = y + pzp
    y = round_impl(y)
    y = clamp_impl(y, min_val=lo, max_val=hi)
    y = y - zp
    return y * s


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("pattern", ["scalar", "per_out_channel", "per_dim1", "pre_scalar_post_channel", "per_row_long",
                                     "scalar_ragged_long"])
@pytest.mark.parametrize("clamp", ["TensorClamp", "TensorClampSte"])
@pytest.mark.parametrize("rounding", ["RoundSte", "FloorSte", "CeilSte", "RoundToZeroSte", "DPURoundSte"])
def test_general_int_quant_vs_literal(dtype, pattern, clamp, rounding):
    import brevitas_b200  # noqa: F401
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.function_wrapper import CLAMP_MODE_OF, ROUND_MODE_OF
    gen = torch.Generator().manual_seed(sum(map(ord, pattern + clamp + rounding)))
    # scalar: unaligned tail -> element-wise flavour; per_out_channel: inner 96 -> 16-byte vectors; per_dim1: inner 25
    # per_row_long (>= 512 vectors per row) and scalar_ragged_long (bulk + element-wise tail): the tiled kernel
    shape = {"scalar": (7, 333), "per_out_channel": (8, 6, 4, 4), "per_dim1": (8, 6, 5, 5),
             "pre_scalar_post_channel": (8, 6, 4, 4), "per_row_long": (6, 4104), "scalar_ragged_long": (3, 11001)}[pattern]
    sshape = {"scalar": (), "per_out_channel": (8, 1, 1, 1), "per_dim1": (1, 6, 1, 1),
              "pre_scalar_post_channel": (8, 1, 1, 1), "per_row_long": (6, 1), "scalar_ragged_long": ()}[pattern]
    pshape = () if pattern == "pre_scalar_post_channel" else sshape
    x0 = (torch.randn(shape, generator=gen) * 3).to(dtype).cuda()
    g = torch.randn(shape, generator=gen).to(dtype).cuda()
    ps0 = (torch.rand(pshape, generator=gen) * 0.05 + 0.02).to(dtype).cuda()
    s0 = (torch.rand(sshape, generator=gen) * 0.05 + 0.02).to(dtype).cuda()
    pzp, zp = torch.tensor(3.0, device="cuda"), torch.tensor(-2.0, device="cuda")
    rm, cm = ROUND_MODE_OF[getattr(fw, rounding)], CLAMP_MODE_OF[getattr(fw, clamp)]
    round_impl, clamp_impl = getattr(fw, rounding)(), getattr(fw, clamp)()

    def both(x):
        res = []
        for fused in (False, True):
            xx, ps, s = (t.clone().requires_grad_(True) for t in (x, ps0, s0))
            lo = torch.tensor(-16.0, device="cuda", requires_grad=True)
            hi = torch.tensor(15.0, device="cuda", requires_grad=True)
            if fused:
                y = torch.ops.brevitas_b200.general_int_quant(xx, ps, s, pzp, zp, lo, hi, rm, cm, False)
            else:
                y = literal(xx, ps, s, pzp, zp, lo, hi, round_impl, clamp_impl)
            (torch.nan_to_num(y, 0.0, 0.0, 0.0) * g).sum().backward()
            res.append((y, xx.grad, ps.grad, s.grad, lo.grad, hi.grad))
        return res

    # (A) special values: element-wise results only (the sums are NaN on both sides)
    xs = x0.clone()
    xs.view(-1)[:8] = torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), 1e-30, 0.5 * 0.02, -1e30]).to(dtype)
    (yl, gxl, *_), (yf, gxf, *_) = both(xs)
    same_bits(yf, yl, "y (special values)")
    same_bits(gxf, gxl, "dx (special values)")
    # (B) finite data: everything
    (yl, gxl, gpl, gsl, glol, ghil), (yf, gxf, gpf, gsf, glof, ghif) = both(x0)
    same_bits(yf, yl, "y")
    same_bits(gxf, gxl, "dx")
    n = x0.numel() // max(ps0.numel(), s0.numel(), 1)
    gabs = g.float().abs()
    for name, a, b, terms in (("d pre_scale", gpf, gpl, gabs * (x0.float().abs() / ps0.float() ** 2 * s0.float())),
                              ("d scale", gsf, gsl, gabs * 20.0)):
        assert bool(torch.isfinite(b).all()) and float(b.float().abs().max()) > 0
        if a.numel() == 1:
            close_sum(a, b, float(terms.sum()), x0.numel(), dtype, name)
        else:
            dims = [d for d in range(x0.dim()) if a.shape[d] == 1]
            tsum = terms.sum(dim=dims, keepdim=True)
            for i in range(a.numel()):
                close_sum(a.view(-1)[i], b.view(-1)[i], float(tsum.reshape(-1)[i]), n, dtype, f"{name}[{i}]")
    if clamp == "TensorClamp":
        tsum = float((gabs * s0.float().abs().max()).sum())
        assert float(glol) != 0.0 and float(ghil) != 0.0
        close_sum(glof, glol, tsum, x0.numel(), dtype, "d min_int")
        close_sum(ghif, ghil, tsum, x0.numel(), dtype, "d max_int")
    else:
        assert glof is None and ghif is None and glol is None and ghil is None


@pytest.mark.parametrize("dtype", DT)
def test_int_quant_same_scale_device_range(dtype):
    """IntQuant.forward with a bit-width that requires grad: one kernel, the whole d(scale) on one accumulator"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import IntQuant
    iq = IntQuant(narrow_range=True, signed=True, float_to_int_impl=fw.RoundSte(), tensor_clamp_impl=fw.TensorClamp()).cuda()
    gen = torch.Generator().manual_seed(5)
    x = (torch.randn(64, 129, generator=gen) * 2).to(dtype).cuda()
    g = torch.randn(64, 129, generator=gen).to(dtype).cuda()
    zp = torch.tensor(0.0, device="cuda")
    res = []
    for fused in (False, True):
        xx = x.clone().requires_grad_(True)
        s = torch.tensor(0.07).to(dtype).cuda().requires_grad_(True)
        bw = torch.tensor(5.0, device="cuda", requires_grad=True)
        if fused:
            y = iq(s, zp, bw, xx)
        else:
            y = literal(xx, s, s, zp, zp, iq.min_int(bw), iq.max_int(bw), iq.float_to_int_impl, iq.tensor_clamp_impl)
        (y * g).sum().backward()
        res.append((y, xx.grad, s.grad, bw.grad))
    (yl, gxl, gsl, gbl), (yf, gxf, gsf, gbf) = res
    same_bits(yf, yl, "y")
    same_bits(gxf, gxl, "dx")
    assert gbf is not None and float(gbl) != 0.0
    terms = float((g.float().abs() * 16).sum())
    close_sum(gsf, gsl, terms, x.numel(), dtype, "d scale")
    close_sum(gbf, gbl, float((g.float().abs() * 0.07 * 16 * math.log(2)).sum()), x.numel(), dtype, "d bit_width")


def test_general_int_quant_under_graph_capture():
    """a direct IntQuant call inside a CUDA graph: nothing is read back (the host-scalar path is skipped), one kernel"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200.core.quant import IntQuant
    iq = IntQuant(narrow_range=False, signed=True).cuda()
    x = torch.randn(1 << 16, device="cuda")
    s, zp, bw = torch.tensor(0.02, device="cuda"), torch.tensor(0.0, device="cuda"), torch.tensor(8.0, device="cuda")
    eager = iq(s, zp, bw, x)
    out = torch.empty_like(x)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        iq(s, zp, bw, x)
    torch.cuda.current_stream().wait_stream(stream)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out.copy_(iq(s, zp, bw, x))
    x.copy_(torch.randn(1 << 16, device="cuda"))
    graph.replay()
    same_bits(out, iq(s, zp, bw, x), "captured == eager on new data")
    assert eager.shape == out.shape


def test_ternary_quant_vs_literal():
    import brevitas_b200  # noqa: F401
    from brevitas_b200.function.ops_ste import ternary_sign_ste
    gen = torch.Generator().manual_seed(2)
    for n in (1, 5, 4096 + 3):
        x = torch.randn(n, generator=gen).cuda()
        if n > 4:
            x[:5] = torch.tensor([0.0, -0.0, float("inf"), float("nan"), 0.35])
        g = torch.randn(n, generator=gen).cuda()
        res = []
        for fused in (False, True):
            xx = x.clone().requires_grad_(True)
            s = torch.tensor(0.7, device="cuda", requires_grad=True)
            if fused:
                y = torch.ops.brevitas_b200.ternary_quant(xx, s, 0.5)
            else:
                mask = xx.abs().gt(0.5 * s)
                y = mask.float() * ternary_sign_ste(xx)
                y = y * s
            (torch.nan_to_num(y) * g).sum().backward()
            res.append((y, xx.grad, s.grad))
        (yl, gxl, gsl), (yf, gxf, gsf) = res
        same_bits(yf, yl, "y")
        same_bits(gxf, gxl, "dx")
        close_sum(gsf, gsl, float(g.abs().sum()), n, torch.float32, "d scale")


@pytest.fixture(scope="module")
def ref():
    src = reference_src()
    if src is None:
        pytest.skip("reference not available (oracle/make_ref.py)")
    from brevitas_b200.binding import uninstall
    yield src
    uninstall()


@pytest.mark.parametrize("kind", ["weight", "act"])
def test_learned_bit_width_through_the_reference_layers(ref, kind):
    """bit_width_impl_type=PARAMETER resolved by the reference's injector: fused classes vs the pure reference, same GPU"""
    from test_gpu_named_quantizers import bound, qt_fields, same
    results = []
    for fused in (False, True):
        qnn, Q = bound(ref, fused)
        from brevitas.inject.enum import BitWidthImplType
        torch.manual_seed(0)
        if kind == "weight":
            layer = qnn.QuantLinear(24, 16, False, weight_quant=Q.Int8WeightPerTensorFloat, weight_bit_width=5,
                                    weight_bit_width_impl_type=BitWidthImplType.PARAMETER).cuda().train()
            qt = layer.quant_weight()
            leaf = layer.weight
            bwp = layer.weight_quant.tensor_quant.msb_clamp_bit_width_impl
        else:
            layer = qnn.QuantIdentity(act_quant=Q.Int8ActPerTensorFloatMinMaxInit, min_val=-2.0, max_val=2.0, bit_width=4,
                                      bit_width_impl_type=BitWidthImplType.PARAMETER, return_quant_tensor=True).cuda().train()
            leaf = (torch.randn(4, 6, 5, 5, generator=torch.Generator().manual_seed(3)) * 2).cuda().requires_grad_(True)
            qt = layer(leaf)
            bwp = layer.act_quant.fused_activation_quant_proxy.tensor_quant.msb_clamp_bit_width_impl
        g = torch.randn(qt.value.shape, generator=torch.Generator().manual_seed(1)).cuda()
        (qt.value * g).sum().backward()
        results.append((qt_fields(qt), leaf.grad.clone(), bwp.bit_width_offset.grad.clone(),
                        type(bwp).__module__, float(g.abs().sum())))
    (fr, gr, br, mr, gsum), (ff, gf, bf, mf, _) = results
    assert mr.startswith("brevitas.core") and mf.startswith("brevitas_b200.core")
    for f, a, b in zip(("value", "scale", "zero_point", "bit_width"), ff, fr):
        same(a, b, f"{kind}.{f}")
    assert torch.allclose(gf, gr, rtol=1e-4, atol=1e-5 * float(gr.abs().max()))
    assert float(br) != 0.0
    assert abs(float(bf) - float(br)) <= 1e-4 * gsum, (float(bf), float(br))


def test_module_paths_launch_one_kernel_each():
    """DecoupledIntQuant, IntQuant with a learned bit-width and TernaryQuant: ONE tensor-sized launch per direction (the
    literal sequences launch the round / clamp / sign STE kernels instead)"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200 import _kernels
    from brevitas_b200.core import function_wrapper as fw
    from brevitas_b200.core.quant import DecoupledIntQuant, IntQuant, TernaryQuant
    from brevitas_b200.core.scaling import ParameterScaling
    x = torch.randn(32, 64, device="cuda", requires_grad=True)
    s = torch.tensor(0.05, device="cuda", requires_grad=True)
    ps = torch.tensor(0.04, device="cuda", requires_grad=True)
    z = torch.tensor(0.0, device="cuda")
    bw = torch.tensor(4.0, device="cuda", requires_grad=True)
    dq = DecoupledIntQuant(True, True, fw.RoundSte(), fw.TensorClampSte()).cuda()
    iq = IntQuant(True, True, fw.RoundSte(), fw.TensorClamp()).cuda()
    tq = TernaryQuant(ParameterScaling(0.7), 0.5).cuda()
    names = []
    real_call = _kernels.call
    _kernels.call = lambda name, *a: (names.append(name), real_call(name, *a))[1]
    try:
        for what, call, kernel in (("decoupled", lambda: dq(ps, z, s, z, bw.detach(), x), "bvb_general_int_quant"),
                                   ("learned bit-width", lambda: iq(s, z, bw, x), "bvb_general_int_quant"),
                                   ("ternary", lambda: tq(x)[0], "bvb_ternary_quant")):
            del names[:]
            call().sum().backward()
            tensor_sized = [n for n in names if n not in ("bvb_abs_binary_sign_grad_impl", "bvb_abs_binary_sign_grad_bwd",
                                                          "bvb_scalar_clamp_min_ste_impl")]       # one-element scale ops
            assert tensor_sized == [kernel + "_fwd", kernel + "_bwd"], (what, names)
    finally:
        _kernels.call = real_call


@pytest.mark.parametrize("dtype", DT)
@pytest.mark.parametrize("rows,cols", [(1, 1), (1, 7), (5, 33), (64, 1024), (3, 100003), (1, 3 * 1024 * 1024 + 5), (700, 96)])
def test_minmax_rows_vs_aten(rows, cols, dtype):
    """one read for min AND max (+ positions): ATen's selection rules -- NaN wins, then the value, then the lowest index"""
    import brevitas_b200  # noqa: F401
    gen = torch.Generator().manual_seed(rows * 31 + cols)
    x = torch.randn(rows, cols, generator=gen).to(dtype).cuda()
    if cols >= 7:
        x[0, 1] = x[0].max()              # tie for the maximum: lowest index wins
        x[0, 5] = x[0, 1]
        x[-1, 3] = float("inf")
        x[-1, 4] = float("-inf")
    if rows >= 5 and cols >= 33:
        x[2, 9] = float("nan")
        x[2, 20] = float("nan")
        x[3] = 0.0
        x[3, 7] = -0.0                     # -0.0 == +0.0: position 0 is selected for both
    mn, mx, imn, imx = torch.ops.brevitas_b200.minmax_rows(x, rows, cols, False)
    rmin, rmax = torch.min(x, dim=1), torch.max(x, dim=1)
    same_bits(mn, rmin.values, "min")
    same_bits(mx, rmax.values, "max")
    finite_rows = ~torch.isnan(rmin.values)
    assert torch.equal(imn[finite_rows], rmin.indices[finite_rows]) and torch.equal(imx[finite_rows], rmax.indices[finite_rows])
    if rows >= 5 and cols >= 33:
        assert int(imn[2]) == 9 and int(imx[2]) == 9 and int(imn[3]) == 0 and int(imx[3]) == 0
    # gradients: along a dim the selected position takes it all; over the whole tensor tied extrema share it
    for whole in (False, True):
        xa = x.clone()
        if whole:
            xa = torch.nan_to_num(xa.reshape(1, -1), 0.0, 5.0, -5.0)
            if xa.numel() > 3:
                xa[0, 2] = xa.max()
        r_, c_ = xa.shape
        a = xa.clone().requires_grad_(True)
        b = xa.clone().requires_grad_(True)
        mn, mx, _, _ = torch.ops.brevitas_b200.minmax_rows(a.reshape(-1) if whole else a, r_, c_, whole)
        if whole:
            ref_mn, ref_mx = torch.min(b.reshape(-1)).view(1), torch.max(b.reshape(-1)).view(1)
        else:
            ref_mn, ref_mx = torch.min(b, dim=1)[0], torch.max(b, dim=1)[0]
        w1 = torch.randn(r_, generator=gen).to(dtype).cuda()
        w2 = torch.randn(r_, generator=gen).to(dtype).cuda()
        keep = ~torch.isnan(ref_mn.detach())
        ((mn * w1)[keep].sum() + (mx * w2)[keep].sum()).backward()
        ((ref_mn * w1)[keep].sum() + (ref_mx * w2)[keep].sum()).backward()
        same_bits(a.grad, b.grad, f"d x (whole={whole})")


@pytest.mark.parametrize("name", ["ShiftedUint8WeightPerTensorFloat", "ShiftedUint8WeightPerChannelFloat"])
def test_asymmetric_weight_quantizer_reads_the_weight_twice(name):
    """statistics (min and max, shared by scale and zero-point) + the quantizer: TWO library launches, no ATen reduction"""
    import brevitas_b200  # noqa: F401
    from brevitas_b200 import _kernels
    import brevitas_b200.nn as qnn
    import brevitas_b200.quant as Q
    layer = qnn.QuantConv2d(16, 32, 3, bias=False, weight_quant=getattr(Q, name)).cuda()
    layer.quant_weight()
    calls, names = [], []
    real_min, real_max, real_call = torch.min, torch.max, _kernels.call
    torch.min = lambda *a, **k: (calls.append("min"), real_min(*a, **k))[1]
    torch.max = lambda *a, **k: (calls.append("max"), real_max(*a, **k))[1]
    _kernels.call = lambda name, *a: (names.append(name), real_call(name, *a))[1]
    try:
        qw = layer.quant_weight()
    finally:
        torch.min, torch.max, _kernels.call = real_min, real_max, real_call
    # the weight is read by exactly two launches; what else runs are the STE ops on the statistics-sized scale / zero-point
    assert names.count("bvb_minmax_rows") == 1 and names.count("bvb_int_quant_zpt_fwd") == 1, names
    assert not calls, calls                                        # no torch.min / torch.max pass is left
    assert all(n in ("bvb_minmax_rows", "bvb_int_quant_zpt_fwd", "bvb_scalar_clamp_min_ste_impl", "bvb_round_ste_impl",
                     "bvb_tensor_clamp_ste_impl", "bvb_tensor_clamp", "bvb_abs_binary_sign_grad_impl") for n in names), names
    assert qw.zero_point is not None
