"""GPU (-m gpu): the CUDA kernels, called through the C-ABI (ctypes -> libbrevitas_b200.so), against

  (1) the golden vectors produced by the real reference (tests/golden), and
  (2) the numpy oracle on fresh seeded inputs incl. ragged / unaligned / empty shapes, and
  (3) at BASELINE.json's full sizes, through size-independent properties (idempotence, range, linearity of
      the gradient, oracle on sampled rows).

Bar (BASELINE.json north_star): integer codes, clamp masks and dequantized outputs bit-exact in fp32; here the
per-op rounding emulation makes bf16/fp16 bit-exact as well (north star asks <= 1 ulp).  Only reductions that
feed a scale gradient carry a tolerance (summation order), stated at each check.
"""
import numpy as np
import pytest
import torch

from golden_util import DTYPES, assert_bits_equal, case, load, ulp
from oracle import fakequant_oracle as O

pytestmark = pytest.mark.gpu

TDT = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}
RM = {"round": 0, "floor": 1, "ceil": 2, "round_to_zero": 3, "dpu_round": 4}
CM = {"ste": 0, "masked": 1}
INT_CASES = {
    "s8n_round_ste_scalar": ("round", "ste"), "s8_round_masked_scalar": ("round", "masked"),
    "u8_round_masked_scalar": ("round", "masked"), "u8_round_masked_scalar_zp": ("round", "masked"),
    "s4n_floor_ste_rows": ("floor", "ste"), "s4_ceil_masked_rows": ("ceil", "masked"),
    "u4n_rtz_masked_chan": ("round_to_zero", "masked"), "s8_dpu_masked_chan_zp": ("dpu_round", "masked"),
    "s2n_round_masked_scalar": ("round", "masked"), "s8n_round_masked_token": ("round", "masked"),
}


@pytest.fixture(scope="module")
def K():
    import brevitas_b200  # noqa: F401  (loads the .so, registers the ops)
    from brevitas_b200 import _kernels
    return _kernels


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(TDT[dtype]).cuda()


def host(t):
    return t.detach().float().cpu().numpy()


def close_sum(got, ref, n, mag, dtype):
    """Two fp32 sums of the same n terms in different orders (``mag``: sum of |terms|, or a bound on it): each differs
    from the exact sum by about sqrt(n) * 2^-24 * mag in the random-walk model (observed: 2.4e-6 at n = 11008, SURVEY
    probe D.6); 8x margin, plus the final rounding to the tensor dtype (4 ulp).  The whole-tensor tests in
    test_gpu_reference_binding.py check the same quantities against an fp64 sum."""
    tol = mag * (8.0 * np.sqrt(n) * 2.0 ** -24 + 4 * ulp(dtype)) + 1e-6
    both_nan = np.isnan(got) & np.isnan(ref)
    assert np.all(both_nan | (np.abs(got - ref) <= tol)), (got, ref, tol)


# ---------------------------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", sorted(INT_CASES))
def test_int_quant_golden(K, name, dtype):
    c = case("int_quant", f"int_quant/{name}/{dtype}/")
    signed, narrow, bits, zp = [float(v) for v in c["meta"]]
    rm, cm = INT_CASES[name]
    qmin, qmax = O.min_int(bool(signed), bool(narrow), bits), O.max_int(bool(signed), bool(narrow), bits)
    x, s, g = dev(c["x"], dtype), dev(c["scale"], dtype), dev(c["g"], dtype)
    y, codes = K.int_quant_fwd(x, s, zp, qmin, qmax, RM[rm], want_codes=True)
    assert_bits_equal(host(y), c["y"], "y")
    assert_bits_equal(host(codes), c["codes"], "codes")
    gx, gs = K.int_quant_bwd(g, x, s, zp, qmin, qmax, RM[rm], CM[cm], True)
    assert_bits_equal(host(gx), c["gx"], "gx")
    if np.isfinite(c["gscale"]).all():
        _, gs_el = O.int_quant_backward(c["g"], c["x"], c["scale"], zp, qmin, qmax, rm, cm, dtype)
        n = c["x"].size // max(1, c["gscale"].size)
        mag = np.abs(gs_el).sum() / max(1, c["gscale"].size) + 1.0
        close_sum(host(gs).reshape(c["gscale"].shape), c["gscale"], n, mag, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tag", ["chan_lin", "chan_conv", "tensor_lin", "tensor_conv"])
def test_weight_stats_golden(K, tag, dtype):
    """Int8WeightPerChannelFloat / Int8WeightPerTensorFloat trees, fused kernel."""
    c = case("weight_stats", f"weight_stats/{tag}/{dtype}/")
    w, g = dev(c["w"], dtype), dev(c["g"], dtype)
    per_channel = tag.startswith("chan")
    a = np.abs(c["w"].reshape(c["w"].shape[0], -1)) if per_channel else np.abs(c["w"]).reshape(1, -1)
    ismax = (a == a.max(axis=1, keepdims=True)).reshape(c["w"].shape)
    if per_channel:
        rows, cols = c["w"].shape[0], c["w"].size // c["w"].shape[0]
        y, scale, am = K.rows_absmax_int_quant_fwd(w, rows, cols, 1e-10, 127.0, 0.0, -127.0, 127.0, 0, want_absmax=True)
        assert_bits_equal(host(y), c["y"], "y")
        assert_bits_equal(host(scale).reshape(c["scale"].shape), c["scale"], "scale")
        assert_bits_equal(host(am), a.max(axis=1), "absmax")
        for gs_in, key in ((None, "gw"), (dev(c["gs"].reshape(-1), dtype), "gw_with_gscale")):
            gx = host(K.rows_absmax_int_quant_bwd(g, w, scale, gs_in, rows, cols, 127.0, 0.0, -127.0, 127.0, 0, 0))
            assert_bits_equal(np.where(ismax, 0, gx), np.where(ismax, 0, c[key]), "gw off the arg-max")
            close_sum(gx[ismax], c[key][ismax], cols, np.nanmax(np.abs(c[key][ismax])) + 1.0, dtype)
    else:
        sdt = torch.float32      # golden: fp32 quantizer buffers -> fp32 scale even for bf16/fp16 weights
        y, scale, am = K.tensor_absmax_int_quant_fwd(w, sdt, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
        assert_bits_equal(host(y), c["y"], "y")
        assert_bits_equal(host(scale), c["scale"], "scale")
        for gs_in, key in ((None, "gw"), (torch.tensor(float(c["gs"]), device="cuda"), "gw_with_gscale")):
            gx = host(K.tensor_absmax_int_quant_bwd(g, w, scale, am, gs_in, 127.0, 0.0, -127.0, 127.0, 0, 0))
            assert_bits_equal(np.where(ismax, 0, gx), np.where(ismax, 0, c[key]), "gw off the arg-max")
            close_sum(gx[ismax], c[key][ismax], c["w"].size, np.nanmax(np.abs(c[key][ismax])) + 1.0, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
def test_runtime_token_golden(K, dtype):
    """per-token dynamic activation quantizer (config 3 composition), training then eval"""
    c = case("runtime_token", f"runtime_token/{dtype}/")
    running = torch.ones(18, device="cuda")
    for step in range(2):
        x, g = dev(c[f"x{step}"], dtype), dev(c[f"g{step}"], dtype)
        y, scale, am = K.rows_absmax_int_quant_fwd(x, 18, 64, 1e-10, 128.0, 0.0, -128.0, 127.0, 0, want_absmax=True)
        assert_bits_equal(host(y), c[f"y{step}"], "y")
        assert_bits_equal(host(scale).reshape(2, 9, 1), c[f"scale{step}"], "scale")
        K.running_stats_update(running, am, 0.1, step == 0)
        assert_bits_equal(host(running).reshape(2, 9, 1), c[f"running{step}"], "running_stats")
        gx = host(K.rows_absmax_int_quant_bwd(g, x, scale, None, 18, 64, 128.0, 0.0, -128.0, 127.0, 0, 1)).reshape(18, 64)
        ref = c[f"gx{step}"].reshape(18, 64)
        a = np.abs(c[f"x{step}"].reshape(18, 64))
        ismax = a == a.max(axis=1, keepdims=True)
        assert_bits_equal(np.where(ismax, 0, gx), np.where(ismax, 0, ref), "gx off the arg-max")
        close_sum(gx[ismax], ref[ismax], 64, np.abs(ref[ismax]).max() + 1.0, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("qname", ["binary", "clamped"])
@pytest.mark.parametrize("sname", ["const", "param", "param_rows"])
def test_binary_golden(K, qname, sname, dtype):
    c = case("binary", f"binary/{qname}_{sname}/{dtype}/")
    clamped = qname == "clamped"
    x, s, g = dev(c["x"], dtype), dev(c["scale"], dtype), dev(c["g"], dtype)
    assert_bits_equal(host(K.binary_quant_fwd(x, s, clamped)), c["y"], "y")
    gx, gs = K.binary_quant_bwd(g, x, s, clamped, True)
    assert_bits_equal(host(gx), c["gx"], "gx")
    if "gvalue" in c:
        n = c["x"].size // c["gvalue"].size
        close_sum(host(gs).reshape(c["gvalue"].shape), c["gvalue"], n, np.abs(c["g"]).sum() / c["gvalue"].size + 1, dtype)


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("clamped", [False, True])
def test_binary_backward_lowp_exhaustive(K, clamped, dtype):
    """the 16-bit backward of BinaryQuant / ClampedBinaryQuant with one positive scale runs in packed-pair arithmetic
    (csrc/binary_quant.cu binary_bwd_vec_packed): ALL 2^16 input patterns (+-0, denormals, +-inf, NaNs) x scales,
    element-wise gradients bit-exact against the oracle, d(scale) within the reduction tolerance; a negative scale and
    an fp32 scalar scale take the literal branch of the same kernel"""
    tdt = TDT[dtype]
    allx = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(tdt)
    x = allx.float().numpy()
    g = O.rnd(rand_np((65536,), 17, 1.0), dtype)
    for sv in (1.0, 0.37, 2.5, 1.7e-3, -0.5):
        s = O.rnd(np.float32(sv), dtype)
        gxo, gs_el = O.binary_quant_backward(g, x, s, clamped, dtype)
        gx, gs = K.binary_quant_bwd(dev(g, dtype), allx.cuda(), dev(np.asarray(s), dtype), clamped, True)
        assert_bits_equal(host(gx), gxo, f"gx scale={sv}")
        fin = np.isfinite(x)
        xf = torch.from_numpy(np.where(fin, x, 1.0)).to(tdt).cuda()
        _, gs = K.binary_quant_bwd(dev(g, dtype), xf, dev(np.asarray(s), dtype), clamped, True)
        _, gs_ref = O.binary_quant_backward(g, host(xf), s, clamped, dtype)
        mag = float(np.abs(gs_ref).sum()) + 1.0
        assert abs(float(gs) - float(gs_ref.sum())) <= mag * (65536 * 2.0 ** -21 + 16 * ulp(dtype))
    s32 = torch.tensor(0.37, device="cuda")                         # fp32 0-dim scale next to a 16-bit tensor
    gx32, _ = K.binary_quant_bwd(dev(g, dtype), allx.cuda(), s32, clamped, True)
    gxo, _ = O.binary_quant_backward(g, x, np.float32(0.37), clamped, dtype)
    assert_bits_equal(host(gx32), gxo, "gx fp32 scale")


def test_percentile_golden(K):
    d = load("percentile")
    for q in (99.999, 99.9, 90.0, 50.0, 1.0):
        for dtype in DTYPES:
            x = d[f"percentile/flat_q{q}/{dtype}/x"]
            val, idx = K.abs_kth_value_rows(dev(x, dtype), 1, x.size, O.percentile_k(q, x.size), want_index=True)
            assert_bits_equal(host(val).reshape(()), d[f"percentile/flat_q{q}/{dtype}/y"], f"q{q} {dtype}")
            assert abs(x[int(idx.item())]) == float(host(val)[0])
    for dtype in DTYPES:
        x2 = d[f"percentile/rows_q99/{dtype}/x"]
        val, _ = K.abs_kth_value_rows(dev(x2, dtype), 6, 250, O.percentile_k(99.0, 250))
        assert_bits_equal(host(val), d[f"percentile/rows_q99/{dtype}/y"], "rows")
    v = d["percentile/kat/x"]           # tests/brevitas/core/test_stats.py:12-18
    got = [float(K.abs_kth_value_rows(dev(v, "f32"), 1, 10, O.percentile_k(10.0 * i, 10))[0]) for i in range(1, 11)]
    assert got == [float(i) for i in range(1, 11)]


def _kth_check(K, x, rows, cols, k):
    """value and position against torch on the same GPU (exact: a selection, no arithmetic); the position must be the
    smallest index holding the k-th smallest |x|"""
    val, idx = K.abs_kth_value_rows(x, rows, cols, k, want_index=True)
    a = x.view(rows, cols).abs()
    ref = a.float().kthvalue(k, dim=1)[0].to(x.dtype)
    assert torch.equal(val, ref), (val, ref)
    first = torch.stack([(a[r] == ref[r]).nonzero()[0, 0] for r in range(rows)])
    assert torch.equal(idx, first), (idx, first)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_percentile_candidate_buffer_regimes(K, dtype):
    """the radix select copies the surviving candidates aside once a histogram shows at most min(2^22, cols / 8) of
    them (csrc/stats.cu): after the first digit (high percentile), after the second (median of 6M values: a quarter of
    them share the binade), never (2M copies of the answer itself), and per row for up to 2 rows"""
    g = torch.Generator(device="cuda").manual_seed(5)
    n = 6_000_000
    x = torch.randn(n, device="cuda", generator=g).to(dtype)
    for q in (99.999, 99.9, 50.0, 1.0):
        _kth_check(K, x, 1, n, O.percentile_k(q, n))
    _kth_check(K, x, 1, n, 1)
    _kth_check(K, x, 1, n, n)
    dup = x.clone()
    dup[::3] = 0.75                      # 2M copies of one value, which is also the median region's answer
    dup[1::3] = -0.75
    _kth_check(K, dup, 1, n, n // 2)
    _kth_check(K, torch.full((n,), -2.5, device="cuda", dtype=dtype), 1, n, n // 3)
    relu = torch.relu(x)                 # > half exact zeros: the low percentiles resolve to 0 with 3M duplicates
    _kth_check(K, relu, 1, n, O.percentile_k(10.0, n))
    _kth_check(K, relu, 1, n, O.percentile_k(99.99, n))
    _kth_check(K, x, 2, 3_000_000, O.percentile_k(99.99, 3_000_000))
    _kth_check(K, x, 2, 3_000_000, O.percentile_k(40.0, 3_000_000))
    rows = x[: 3 * 1_500_000]            # > 2 rows: no candidate buffer
    _kth_check(K, rows, 3, 1_500_000, O.percentile_k(99.99, 1_500_000))
    many = x[: 6 * 1_000_000]
    _kth_check(K, many, 6, 1_000_000, O.percentile_k(99.9, 1_000_000))
    ragged = x[1: 1 + 3 * 1_000_001]     # unaligned rows: scalar path
    _kth_check(K, ragged, 3, 1_000_001, O.percentile_k(99.999, 1_000_001))


STE_C = {"round_ste": "bvb_round_ste_impl", "ceil_ste": "bvb_ceil_ste_impl", "floor_ste": "bvb_floor_ste_impl",
         "binary_sign_ste": "bvb_binary_sign_ste_impl", "ternary_sign_ste": "bvb_ternary_sign_ste_impl",
         "round_to_zero_ste": "bvb_round_to_zero_ste_impl", "dpu_round_ste": "bvb_dpu_round_ste_impl",
         "abs_binary_sign_grad": "bvb_abs_binary_sign_grad_impl"}


@pytest.mark.parametrize("dtype", DTYPES)
def test_ste_golden(K, dtype):
    d = load("ste")
    for name, cname in STE_C.items():
        x = d[f"ste/{name}/{dtype}/x"]
        assert_bits_equal(host(K.unary(cname, dev(x, dtype))), d[f"ste/{name}/{dtype}/y"], name)
    x = d[f"ste/abs_binary_sign_grad/{dtype}/x"]
    g = d[f"ste/abs_binary_sign_grad/{dtype}/g"]
    assert_bits_equal(host(K.abs_binary_sign_grad_bwd(dev(x, dtype), dev(g, dtype))),
                      d[f"ste/abs_binary_sign_grad/{dtype}/gx"], "abs grad bwd")
    x = d[f"ste/tensor_clamp_ste/{dtype}/x"]
    lo, hi = torch.tensor(-1.25, device="cuda").to(TDT[dtype]), torch.tensor(2.5, device="cuda").to(TDT[dtype])
    assert_bits_equal(host(K.tensor_clamp(dev(x, dtype), lo, hi)), d[f"ste/tensor_clamp_ste/{dtype}/y"], "tclamp")
    xm = dev(x, dtype)
    r = K.tensor_clamp(xm, lo, hi, inplace=True)
    assert r.data_ptr() == xm.data_ptr()
    assert_bits_equal(host(xm), d[f"ste/tensor_clamp_ste_/{dtype}/y"], "tclamp_ (in place)")
    assert_bits_equal(host(K.scalar_clamp(dev(x, dtype), -1.3, 2.7)), d[f"ste/scalar_clamp_ste/{dtype}/y"], "sclamp")
    assert_bits_equal(host(K.scalar_clamp_min(dev(x, dtype), 0.3)), d[f"ste/scalar_clamp_min_ste/{dtype}/y"], "sclampmin")
    r = case("ste", f"ste/tensor_clamp_rows/{dtype}/")
    assert_bits_equal(host(K.tensor_clamp(dev(r["x"], dtype), dev(r["lo"], dtype), dev(r["hi"], dtype))), r["y"], "rows")


def test_docstring_kats(K):
    d = load("kat")
    x = dev(d["kat/int_quant/x"], "f32")
    y = K.int_quant_fwd(x, torch.tensor(0.01, device="cuda"), 0.0, -7.0, 7.0, 0)
    assert_bits_equal(host(y), d["kat/int_quant/y"], "IntQuant docstring")
    s = (torch.tensor(0.1) / torch.tensor(7.0)).cuda()     # a true division (CUDA ATen multiplies by 1/7 for a Python scalar)
    assert_bits_equal(host(K.int_quant_fwd(x, s, 0.0, -7.0, 7.0, 0)), d["kat/rescaling/y"], "RescalingIntQuant docstring")
    b = dev(d["kat/binary/x"], "f32")
    sc = torch.tensor(0.1, device="cuda")
    assert_bits_equal(host(K.binary_quant_fwd(b, sc, False)), d["kat/binary/y"])
    assert_bits_equal(host(K.binary_quant_fwd(b, sc, True)), d["kat/clamped_binary/y"])
    gx, _ = K.binary_quant_bwd(torch.ones(3, device="cuda"), b, sc, True, False)
    assert_bits_equal(host(gx), d["kat/clamped_binary/gx"])


# ---------------------------------------------------------------------------------------------------------------
# oracle on fresh inputs: shapes that exercise every kernel variant
# ---------------------------------------------------------------------------------------------------------------
def rand_np(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols", [(1, 8), (3, 1), (5, 17), (7, 1000), (64, 4096), (33, 11008), (2, 60000),
                                       (300, 264), (2, 131072), (1500, 8200), (445, 12288), (900, 9000)])
def test_rows_fused_vs_oracle(K, rows, cols, dtype):
    """TMA path (16-byte aligned rows), generic path (ragged rows) and oversized rows; the last three shapes: 16-bit rows of
    16 KB and more with 1-4 rows per CTA (the in-place + TMA-store forward, ring wrap-around and short tails)"""
    x = O.rnd(rand_np((rows, cols), rows * 1000 + cols, 0.7), dtype)
    g = O.rnd(rand_np((rows, cols), 7, 1.0), dtype)
    if rows > 2 and cols > 4:
        x[1, 2] = -x[1, 0] if abs(x[1, 0]) >= np.abs(x[1]).max() else -np.abs(x[1]).max()   # tie on the row max
        x[2, :] = 0.0
    yo, so, amo = O.rows_absmax_int_quant_forward(x, 1e-10, 127.0, 0.0, -127.0, 127.0, "round", dtype)
    xd, gd = dev(x, dtype), dev(g, dtype)
    y, s, am = K.rows_absmax_int_quant_fwd(xd, rows, cols, 1e-10, 127.0, 0.0, -127.0, 127.0, 0, want_absmax=True)
    assert_bits_equal(host(am), amo, "absmax")
    assert_bits_equal(host(s), so, "scale")
    assert_bits_equal(host(y), yo, "y")
    assert_bits_equal(host(K.absmax_rows(xd, rows, cols)), amo, "absmax_rows")
    for cm in ("ste", "masked"):
        gxo, _ = O.rows_absmax_int_quant_backward(g, x, so, None, 127.0, 0.0, -127.0, 127.0, "round", cm, dtype)
        gx = host(K.rows_absmax_int_quant_bwd(gd, xd, s, None, rows, cols, 127.0, 0.0, -127.0, 127.0, 0, CM[cm]))
        a = np.abs(x)
        first = np.zeros_like(x, dtype=bool)
        first[np.arange(rows), a.argmax(axis=1)] = True
        assert_bits_equal(np.where(first, 0, gx), np.where(first, 0, gxo), f"gx off the arg-max ({cm})")
        close_sum(gx[first], gxo[first], cols, np.nanmax(np.abs(gxo[first])) + 1.0, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("n", [1, 7, 1023, 4096, 100003, 1 << 20])
def test_tensor_fused_vs_oracle(K, n, dtype):
    x = O.rnd(rand_np((n,), n, 0.5), dtype)
    g = O.rnd(rand_np((n,), n + 1, 1.0), dtype)
    if n > 10:
        m = np.abs(x).max()
        x[3], x[n - 2] = m, -m                 # tied maxima: gradient is split evenly (torch.max() semantics)
    yo, so, amo = O.tensor_absmax_int_quant_forward(x, 1e-10, 127.0, 0.0, -127.0, 127.0, "round", dtype)
    xd, gd = dev(x, dtype), dev(g, dtype)
    y, s, am = K.tensor_absmax_int_quant_fwd(xd, TDT[dtype], 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    assert_bits_equal(host(am), amo, "absmax")
    assert_bits_equal(host(s), so, "scale")
    assert_bits_equal(host(y), yo, "y")
    assert_bits_equal(host(K.absmax_tensor(xd)), amo, "absmax_tensor")
    gxo, _ = O.tensor_absmax_int_quant_backward(g, x, so, None, 127.0, 0.0, -127.0, 127.0, "round", "masked", dtype)
    gx = host(K.tensor_absmax_int_quant_bwd(gd, xd, s, am, None, 127.0, 0.0, -127.0, 127.0, 0, 1))
    ties = np.abs(x) == np.abs(x).max()
    assert_bits_equal(np.where(ties, 0, gx), np.where(ties, 0, gxo), "gx off the maxima")
    close_sum(gx[ties], gxo[ties], n, np.nanmax(np.abs(gxo[ties])) + 1.0, dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,sshape", [((1000,), ()), ((5, 3), (5, 1)), ((4, 6, 5, 5), (1, 6, 1, 1)),
                                          ((4, 8, 16), (4, 8, 1)), ((3, 1001), (3, 1)), ((16, 64, 8, 8), (1, 64, 1, 1)),
                                          ((2, 3, 4), (2, 3, 4))])
@pytest.mark.parametrize("rm", ["round", "floor", "ceil", "round_to_zero", "dpu_round"])
def test_int_quant_broadcast_vs_oracle(K, shape, sshape, rm, dtype):
    x = O.rnd(rand_np(shape, 5, 30.0), dtype)
    g = O.rnd(rand_np(shape, 6, 1.0), dtype)
    s = O.rnd(np.abs(rand_np(sshape, 8, 0.3)) + 0.05, dtype)
    zp = 2.0
    yo = O.int_quant_forward(x, s, zp, 0.0, 255.0, rm, dtype)
    assert_bits_equal(host(K.int_quant_fwd(dev(x, dtype), dev(s, dtype), zp, 0.0, 255.0, RM[rm])), yo, "y")
    gxo, gs_el = O.int_quant_backward(g, x, s, zp, 0.0, 255.0, rm, "masked", dtype)
    gx, gs = K.int_quant_bwd(dev(g, dtype), dev(x, dtype), dev(s, dtype), zp, 0.0, 255.0, RM[rm], 1, True)
    assert_bits_equal(host(gx), gxo, "gx")
    s_full = np.broadcast_to(s, shape)
    # reduce the oracle's per-element terms over each scale's region
    ref = np.zeros(max(1, s.size))
    idx = np.broadcast_to(np.arange(max(1, s.size)).reshape(s.shape), shape)
    np.add.at(ref, idx.reshape(-1), gs_el.reshape(-1))
    mag = np.zeros_like(ref)
    np.add.at(mag, idx.reshape(-1), np.abs(gs_el).reshape(-1))
    n = x.size // max(1, s.size)
    got = host(gs)
    assert np.all(np.abs(got - ref) <= mag * (n * 2.0 ** -21 + 16 * ulp(dtype)) + 1e-5), (got, ref)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,sshape", [(((1 << 20) + 4096 * 3 + 5,), ()), ((600, 2048), (600, 1)),
                                          ((3, 40, 64, 64), (1, 40, 1, 1)), ((2, 300, 4096), (2, 300, 1))])
@pytest.mark.parametrize("cm", ["ste", "masked"])
def test_provided_scale_large_vs_oracle(K, shape, sshape, cm, dtype):
    """>= 1 MiB tensors take the TMA-pipelined provided-scale backward (scaled_bwd_tma_kernel): one scale with a
    ragged tail, a scale per row, per NCHW channel (scale index wraps over the batch) and per token."""
    x = O.rnd(rand_np(shape, 21, 30.0), dtype)
    g = O.rnd(rand_np(shape, 22, 1.0), dtype)
    s = O.rnd(np.abs(rand_np(sshape, 23, 0.3)) + 0.05, dtype)
    yo = O.int_quant_forward(x, s, 0.0, -128.0, 127.0, "round", dtype)
    assert_bits_equal(host(K.int_quant_fwd(dev(x, dtype), dev(s, dtype), 0.0, -128.0, 127.0, 0)), yo, "y")
    gxo, gs_el = O.int_quant_backward(g, x, s, 0.0, -128.0, 127.0, "round", cm, dtype)
    gx, gs = K.int_quant_bwd(dev(g, dtype), dev(x, dtype), dev(s, dtype), 0.0, -128.0, 127.0, 0, CM[cm], True)
    assert_bits_equal(host(gx), gxo, "gx")
    gx2, none = K.int_quant_bwd(dev(g, dtype), dev(x, dtype), dev(s, dtype), 0.0, -128.0, 127.0, 0, CM[cm], False)
    assert none is None
    assert_bits_equal(host(gx2), gxo, "gx without d(scale)")
    ref = np.zeros(max(1, s.size))
    mag = np.zeros_like(ref)
    idx = np.broadcast_to(np.arange(max(1, s.size)).reshape(s.shape), shape).reshape(-1)
    np.add.at(ref, idx, gs_el.reshape(-1))
    np.add.at(mag, idx, np.abs(gs_el).reshape(-1))
    n = x.size // max(1, s.size)
    got = host(gs).reshape(-1)
    assert np.all(np.abs(got - ref) <= mag * (n * 2.0 ** -21 + 16 * ulp(dtype)) + 1e-5), (got[:4], ref[:4])


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,cols,chunk", [(37, 1000, 8), (64, 4096, None), (5, 264, 100), (300, 11008, 64)])
def test_host_pipeline_matches_device_path(K, rows, cols, chunk, dtype):
    """bvb_host_rows_fakequant_fwd_bwd (host buffers, chunked three-stream pipeline, ragged last chunk) returns the
    bits of the device-resident fwd + bwd calls, and of the oracle."""
    from brevitas_b200.host_pipeline import weight_fake_quant_fwd_bwd_host
    x = O.rnd(rand_np((rows, cols), rows + cols, 0.7), dtype)
    g = O.rnd(rand_np((rows, cols), 9, 1.0), dtype)
    hx = torch.from_numpy(x).to(TDT[dtype]).pin_memory()
    hg = torch.from_numpy(g).to(TDT[dtype]).pin_memory()
    for want_q in (False, True):
        gw, sc, wq = weight_fake_quant_fwd_bwd_host(hx, hg, bit_width=8, signed=True, narrow_range=True,
                                                    chunk_rows=chunk, want_quantized=want_q)
        y, s, _ = K.rows_absmax_int_quant_fwd(hx.cuda(), rows, cols, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
        gx = K.rows_absmax_int_quant_bwd(hg.cuda(), hx.cuda(), s, None, rows, cols, 127.0, 0.0, -127.0, 127.0, 0, 0)
        assert gw.shape == (rows, cols) and sc.shape == (rows, 1) and not gw.is_cuda
        assert_bits_equal(gw.float().numpy(), host(gx), "grad (host pipeline vs device calls)")
        assert_bits_equal(sc.float().numpy().reshape(-1), host(s), "scale")
        yo, so, _ = O.rows_absmax_int_quant_forward(x, 1e-10, 127.0, 0.0, -127.0, 127.0, "round", dtype)
        assert_bits_equal(sc.float().numpy().reshape(-1), so, "scale vs oracle")
        if want_q:
            assert_bits_equal(wq.float().numpy(), yo, "quantized weight vs oracle")
        else:
            assert wq is None
    with pytest.raises(RuntimeError):
        weight_fake_quant_fwd_bwd_host(hx.cuda(), hg.cuda())


@pytest.mark.parametrize("dtype", DTYPES)
def test_channels_last_layouts(K, dtype):
    """channels-last activations (one scale) and conv weights (a scale per output channel, fused abs-max) go through
    the kernels in place -- no NCHW copy -- and give the bits of the row-major call, in the caller's layout."""
    T = TDT[dtype]
    x = torch.from_numpy(rand_np((4, 16, 9, 9), 31, 20.0)).to(T).cuda()
    g = torch.from_numpy(rand_np((4, 16, 9, 9), 32, 1.0)).to(T).cuda()
    s = torch.tensor(0.21, device="cuda").to(T)
    xcl, gcl = x.contiguous(memory_format=torch.channels_last), g.contiguous(memory_format=torch.channels_last)
    y, ycl = K.int_quant_fwd(x, s, 0.0, 0.0, 255.0, 0), K.int_quant_fwd(xcl, s, 0.0, 0.0, 255.0, 0)
    assert ycl.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, ycl)
    (gx, gs), (gxcl, gscl) = K.int_quant_bwd(g, x, s, 0.0, 0.0, 255.0, 0, 1, True), K.int_quant_bwd(g, xcl, s, 0.0, 0.0, 255.0, 0, 1, True)
    assert gxcl.is_contiguous(memory_format=torch.channels_last) and torch.equal(gx, gxcl)      # row-major g, CL x
    assert torch.equal(gx, K.int_quant_bwd(gcl, xcl, s, 0.0, 0.0, 255.0, 0, 1, True)[0])
    assert abs(float(gs) - float(gscl)) <= 1e-3 * (abs(float(gs)) + 1.0)
    # a [1,C,1,1] scale on NHWC memory is "scale[i % C]": int_quant_chanlast_kernel (hoisted per-thread divisors)
    for cch, hw in ((16, 9), (64, 5), (24, 4)):          # 24 does not divide 256*V for fp32: row-major fallback
        xc = torch.from_numpy(rand_np((3, cch, hw, hw + 3), 35, 20.0)).to(T).cuda()
        gc = torch.from_numpy(rand_np((3, cch, hw, hw + 3), 36, 1.0)).to(T).cuda()
        sc = (torch.rand(1, cch, 1, 1, device="cuda") * 0.2 + 0.1).to(T)
        xccl = xc.contiguous(memory_format=torch.channels_last)
        yr, kr = K.int_quant_fwd(xc, sc, 0.0, 0.0, 255.0, 0, want_codes=True)
        yc, kc = K.int_quant_fwd(xccl, sc, 0.0, 0.0, 255.0, 0, want_codes=True)
        assert torch.equal(yr, yc) and torch.equal(kr, kc)
        gxr, gsr = K.int_quant_bwd(gc, xc, sc, 0.0, 0.0, 255.0, 0, 1, True)
        gxc, gsc = K.int_quant_bwd(gc, xccl, sc, 0.0, 0.0, 255.0, 0, 1, True)
        assert torch.equal(gxr, gxc)
        assert torch.allclose(gsr, gsc, rtol=1e-3 if dtype == "f32" else 5e-2, atol=1e-2)
        yo = O.int_quant_forward(host(xc), host(sc), 0.0, 0.0, 255.0, "round", dtype)
        assert_bits_equal(host(yc), yo, "channels-last per-channel vs oracle")
    # conv weight [O, I, kh, kw]: rows are dim-0 slices in both layouts
    w = torch.from_numpy(rand_np((24, 16, 3, 3), 33, 0.3)).to(T).cuda()
    gw = torch.from_numpy(rand_np((24, 16, 3, 3), 34, 1.0)).to(T).cuda()
    wcl = w.contiguous(memory_format=torch.channels_last)
    y, sc_r, _ = K.rows_absmax_int_quant_fwd(w, 24, 144, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    ycl, sc_c, _ = K.rows_absmax_int_quant_fwd(wcl, 24, 144, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    assert ycl.is_contiguous(memory_format=torch.channels_last) and torch.equal(y, ycl) and torch.equal(sc_r, sc_c)
    gx = K.rows_absmax_int_quant_bwd(gw, w, sc_r, None, 24, 144, 127.0, 0.0, -127.0, 127.0, 0, 0)
    gxcl = K.rows_absmax_int_quant_bwd(gw, wcl, sc_c, None, 24, 144, 127.0, 0.0, -127.0, 127.0, 0, 0)
    # the element receiving the abs-max gradient is the FIRST maximum in memory order: identical unless a row has ties
    amax = w.abs().reshape(24, -1).amax(dim=1).view(24, 1, 1, 1)
    off = w.abs() != amax
    assert torch.equal(torch.where(off, gx, torch.zeros_like(gx)), torch.where(off, gxcl, torch.zeros_like(gx)))
    assert torch.allclose(gx.float(), gxcl.float(), rtol=2e-2, atol=2e-2)
    y, st, am = K.tensor_absmax_int_quant_fwd(w, torch.float32, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    ycl, stc, _ = K.tensor_absmax_int_quant_fwd(wcl, torch.float32, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    assert torch.equal(y, ycl) and torch.equal(st, stc)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,sshape,cl", [((5000,), (), False), ((1 << 20,), (), False), ((6, 16, 9, 9), (1, 16, 1, 1), False),
                                             ((6, 16, 9, 9), (1, 16, 1, 1), True), ((600, 2048), (600, 1), False),
                                             ((4, 8, 33), (4, 8, 1), False)])
def test_relu_fused_into_quantizer(K, shape, sshape, cl, dtype):
    """bvb_relu_int_quant_{fwd,bwd} (QuantReLU: nn.ReLU folded into its quantizer kernel) == int_quant(torch.relu(x))
    followed by ATen's ReLU backward, bit for bit, on every provided-scale kernel variant (streaming, planes, TMA ring,
    channels-last, scalar fallbacks), with -0.0 / NaN / +-inf among the inputs."""
    T = TDT[dtype]
    x = O.rnd(rand_np(shape, 41, 20.0), dtype)
    flat = x.reshape(-1)
    flat[:6] = [0.0, -0.0, np.nan, np.inf, -np.inf, -1e-30]
    g = torch.from_numpy(O.rnd(rand_np(shape, 42, 1.0), dtype)).to(T).cuda()
    s = torch.from_numpy(O.rnd(np.abs(rand_np(sshape, 43, 0.3)) + 0.05, dtype)).reshape(sshape).to(T).cuda()
    xd = torch.from_numpy(x).to(T).cuda()
    if cl:
        xd = xd.contiguous(memory_format=torch.channels_last)
    for qmin, qmax, cm in ((0.0, 255.0, 1), (0.0, 15.0, 0)):
        yf = K.int_quant_fwd(xd, s, 0.0, qmin, qmax, 0, pre_relu=True)
        xr = torch.relu(xd)
        yr = K.int_quant_fwd(xr, s, 0.0, qmin, qmax, 0)
        assert_bits_equal(host(yf), host(yr), "forward")
        gxf, gsf = K.int_quant_bwd(g, xd, s, 0.0, qmin, qmax, 0, cm, True, pre_relu=True)
        gq, gsr = K.int_quant_bwd(g, xr, s, 0.0, qmin, qmax, 0, cm, True)
        gxr = torch.where(xd <= 0, torch.zeros_like(gq), gq)          # threshold_backward: zero where x <= 0
        assert_bits_equal(host(gxf), host(gxr), "gradient")
        # ... and against the oracle's restatement of the fused op
        cmn = "masked" if cm == 1 else "ste"
        assert_bits_equal(host(yf), O.relu_int_quant_forward(host(xd), host(s), 0.0, qmin, qmax, "round", dtype), "fwd vs oracle")
        gxo, _ = O.relu_int_quant_backward(host(g), host(xd), host(s), 0.0, qmin, qmax, "round", cmn, dtype)
        assert_bits_equal(host(gxf), gxo, "bwd vs oracle")
        ok = torch.isfinite(gsr) & torch.isfinite(gsf)
        assert torch.allclose(gsf[ok], gsr[ok], rtol=1e-3 if dtype == "f32" else 6e-2, atol=1e-2)


def test_quant_relu_module_fusion_matches_unfused():
    """QuantReLU with a learned scale: the fused proxy path and the literal ReLU -> tensor_quant path agree exactly"""
    from brevitas_b200.nn import QuantReLU
    from brevitas_b200.quant import Uint8ActPerTensorFloatMaxInit
    torch.manual_seed(0)
    act = QuantReLU(act_quant=Uint8ActPerTensorFloatMaxInit, max_val=6.0, bit_width=4, return_quant_tensor=True).cuda()
    x1 = (torch.randn(8, 16, 12, 12, device="cuda") * 3).requires_grad_(True)
    x2 = x1.detach().clone().requires_grad_(True)
    proxy = act.act_quant.fused_activation_quant_proxy
    q1 = act(x1)
    q1.value.square().sum().backward()
    g_scale_fused = proxy.tensor_quant.scaling_impl.value.grad.clone()
    proxy.tensor_quant.scaling_impl.value.grad = None
    y2, s2, zp2, bw2 = proxy.tensor_quant(torch.relu(x2))            # the unfused composition
    y2.square().sum().backward()
    assert torch.equal(q1.value, y2) and torch.equal(q1.scale, s2)
    assert torch.equal(x1.grad, x2.grad)
    assert torch.allclose(g_scale_fused, proxy.tensor_quant.scaling_impl.value.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["u8_round_masked_scalar_zp", "s8_dpu_masked_chan_zp"])
def test_tensor_zero_point_kernel(K, name, dtype):
    """bvb_int_quant_zpt_{fwd,bwd}: the zero-point as a DEVICE operand with the scale's broadcast pattern.  Outputs and
    element-wise gradients bit-exact against the reference's golden vectors (which used the same value as a Python
    constant); d(scale) within the reduction tolerance; d(zero_point) against the closed form sum(m*d - d)."""
    c = case("int_quant", f"int_quant/{name}/{dtype}/")
    signed, narrow, bits, zp = [float(v) for v in c["meta"]]
    rm, cm = INT_CASES[name]
    qmin, qmax = O.min_int(bool(signed), bool(narrow), bits), O.max_int(bool(signed), bool(narrow), bits)
    x, s, g = dev(c["x"], dtype), dev(c["scale"], dtype), dev(c["g"], dtype)
    zpt = torch.full_like(s, zp)
    y = K.int_quant_zpt_fwd(x, s, zpt, qmin, qmax, RM[rm])
    assert_bits_equal(host(y), c["y"], "y")
    gx, gs, gz = K.int_quant_zpt_bwd(g, x, s, zpt, qmin, qmax, RM[rm], CM[cm], True)
    assert_bits_equal(host(gx), c["gx"], "gx")
    gx2, none_s, none_z = K.int_quant_zpt_bwd(g, x, s, zpt, qmin, qmax, RM[rm], CM[cm], False)
    assert none_s is None and none_z is None
    assert_bits_equal(host(gx2), c["gx"], "gx without parameter gradients")
    if np.isfinite(c["gscale"]).all():
        _, gs_el = O.int_quant_backward(c["g"], c["x"], c["scale"], zp, qmin, qmax, rm, cm, dtype)
        n = c["x"].size // max(1, c["gscale"].size)
        mag = np.abs(gs_el).sum() / max(1, c["gscale"].size) + 1.0
        close_sum(host(gs).reshape(c["gscale"].shape), c["gscale"], n, mag, dtype)
    # d(zero_point): every element contributes m*d - d with d = rnd(g * s)  (0 inside the range, -d where clamped)
    _, t1, t3, t5 = O.int_quant_chain(c["x"], c["scale"], zp, qmin, qmax, rm, dtype)
    d = O.rnd(c["g"] * c["scale"], dtype).astype(np.float64)
    clamped = (t3 > O.scalar_to(qmax, dtype)) | (t3 < O.scalar_to(qmin, dtype))
    el = np.where(clamped & np.isfinite(d), -d, 0.0) if cm == "masked" else np.zeros_like(d)
    el = np.where(np.isfinite(el), el, 0.0)              # +-inf inputs are ordinary clamped elements; NaN compares false
    ref = np.zeros(max(1, c["scale"].size))
    idx = np.broadcast_to(np.arange(ref.size).reshape(c["scale"].shape), c["x"].shape).reshape(-1)
    np.add.at(ref, idx, el.reshape(-1))
    fin = np.isfinite(host(gz)) & np.isfinite(ref)
    tol = np.abs(el).sum() * 4 * ulp(dtype) + 1e-4
    assert np.all(np.abs(host(gz).astype(np.float64)[fin] - ref[fin]) <= tol), (host(gz), ref)


@pytest.mark.parametrize("dtype", DTYPES)
def test_integer_export(K, dtype):
    """bvb_int_quant_to_int: IntQuant.to_int stored as int8 / uint8 / int32, against the golden codes of the reference
    and against the float codes of the fused kernel; scalar, per-row, per-channel and ragged / unaligned cases;
    QuantTensor.int() of a layer output."""
    T = TDT[dtype]
    for name, out_dt in (("s8n_round_ste_scalar", torch.int8), ("u8_round_masked_scalar", torch.uint8),
                         ("u8_round_masked_scalar_zp", torch.uint8), ("s4n_floor_ste_rows", torch.int8),
                         ("u4n_rtz_masked_chan", torch.uint8), ("s8_dpu_masked_chan_zp", torch.int32)):
        c = case("int_quant", f"int_quant/{name}/{dtype}/")
        signed, narrow, bits, zp = [float(v) for v in c["meta"]]
        rm, _ = INT_CASES[name]
        qmin, qmax = O.min_int(bool(signed), bool(narrow), bits), O.max_int(bool(signed), bool(narrow), bits)
        x, s = dev(c["x"], dtype), dev(c["scale"], dtype)
        got = K.int_quant_to_int(x, s, zp, qmin, qmax, RM[rm], out_dt)
        ref = np.nan_to_num(c["codes"], nan=0.0)
        assert got.dtype == out_dt and np.array_equal(got.cpu().numpy().astype(np.float64), ref), name
    x = torch.from_numpy(O.rnd(rand_np((37, 1003), 51, 30.0), dtype)).to(T).cuda()
    for s in (torch.tensor(0.3, device="cuda").to(T), (torch.rand(37, 1, device="cuda") * 0.3 + 0.1).to(T)):
        _, codes = K.int_quant_fwd(x, s, 0.0, -128.0, 127.0, 0, want_codes=True)
        for xx in (x, x.reshape(-1)[3:].reshape(-1)) if s.numel() == 1 else (x,):       # unaligned view
            ref = codes if xx is x else K.int_quant_fwd(xx.contiguous(), s, 0.0, -128.0, 127.0, 0, want_codes=True)[1]
            for out_dt in (torch.int8, torch.int32):
                got = K.int_quant_to_int(xx, s, 0.0, -128.0, 127.0, 0, out_dt)
                assert torch.equal(got.to(torch.float32), ref.float())
    if dtype == "f32":
        from brevitas_b200.nn import QuantReLU
        from brevitas_b200.quant import Uint8ActPerTensorFloatMaxInit
        act = QuantReLU(act_quant=Uint8ActPerTensorFloatMaxInit, max_val=6.0, return_quant_tensor=True).cuda()
        q = act(torch.randn(4, 8, 6, 6, device="cuda") * 3)
        codes = q.int()
        assert codes.dtype == torch.uint8
        assert torch.equal(codes.float(), torch.round(q.value.detach() / q.scale.detach()))
        assert torch.equal(q.int(float_datatype=True).detach(), codes.float())


def test_fp32_scalar_scale_with_lowp_input(K):
    """fp32 quantizer modules fed bf16 activations: ATen's mul/div keep the fp32 0-dim scale in opmath"""
    x = torch.randn(4099, generator=torch.Generator().manual_seed(3)).mul(20).to(torch.bfloat16)
    s = torch.tensor(0.0371)
    ref = (torch.clamp(torch.round(x / s + 0.0), -128, 127) - 0.0) * s
    got = K.int_quant_fwd(x.cuda(), s.cuda(), 0.0, -128.0, 127.0, 0)
    assert got.dtype == torch.bfloat16
    assert_bits_equal(host(got), ref.float().numpy(), "bf16 x, fp32 0-dim scale")


def test_division_bit_identity():
    """The kernels divide by the (row-/tensor-constant) scale with nvcc's div.rn.f32 sequence, reciprocal refinement
    hoisted (csrc/common.cuh DivBy).  It must equal IEEE division bit for bit: ALL 2^32 numerators for a set of
    divisors (ordinary scales, powers of two, all-ones mantissa, tiny/huge/denormal/inf/nan/negative divisors)."""
    import ctypes
    from brevitas_b200 import _lib
    lib = _lib.load()
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    rng = np.random.default_rng(0)
    divisors = [1.0, 0.5, 127.0, 1e-10 / 127, 0.0078740157, 3.0, 1.9999999, 1.0000001, 0.037, 255.0, 6.0 / 15,
                2.0 ** -40, 2.0 ** 40 * 0.999, 2.0 ** -41, 2.0 ** 41, 1e-38, 1e-45, 3e38, float("inf"), float("nan"),
                0.0, -0.0, -0.25, -1e-3]
    divisors += [float(np.float32(v)) for v in np.exp(rng.uniform(-12, 6, 12))]
    for d in divisors:
        _lib.call("bvb_selftest_div", ctypes.c_float(d), 0, 1 << 32, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert int(out.item()) == 0, f"divisor {d!r}: {int(out.item())} of 2^32 quotients differ from IEEE division"


def test_lowp_division_shortcut_exhaustive():
    """bf16 kernels compute RN_bf16(a / b) as RN_bf16(a * RN_f32(1/b)); checked for ALL pairs of bf16 values whose
    divisor is inside the shortcut's window.  The same enumeration for fp16 finds mismatches, which is why the fp16
    kernels keep the full division sequence (DT<__half>::MUL_DIV_EXACT = false)."""
    from brevitas_b200 import _lib
    out = torch.zeros(2, dtype=torch.int64, device="cuda")
    _lib.call("bvb_selftest_lowp_div", _lib.BF16, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    bad, checked = [int(v) for v in out.tolist()]
    assert checked > 2 ** 30 and bad == 0, f"bf16: {bad} of {checked} quotients differ"
    _lib.call("bvb_selftest_lowp_div", _lib.F16, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    bad16, checked16 = [int(v) for v in out.tolist()]
    print(f"fp16 shortcut would be wrong for {bad16} of {checked16} pairs (not used)")
    assert checked16 > 2 ** 28


@pytest.mark.parametrize("dtype", ["bf16", "f16"])
@pytest.mark.parametrize("qmin,qmax", [(-127.0, 127.0), (-128.0, 127.0), (0.0, 255.0), (-8.0, 7.0), (0.0, 15.0),
                                       (-1.0, 1.0), (0.0, 65535.0)])
def test_lowp_exhaustive(K, dtype, qmin, qmax):
    """bf16 / fp16 kernels use packed-pair arithmetic on the default modes (csrc/common.cuh qdq_vec, int_quant.cu
    bwd_vec; (0, 65535) is not representable and takes the literal path).  ALL 2^16 input bit patterns (incl. +-0,
    denormals, +-inf, NaNs) x a set of scales: outputs, integer codes and element-wise gradients (STE and masked)
    bit-exact against the oracle, through the scalar-scale, per-row-scale and fused per-row abs-max kernels."""
    tdt = TDT[dtype]
    allx = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(tdt)
    x = allx.float().numpy()
    g = O.rnd(rand_np((65536,), 11, 1.0), dtype)
    scales = [1.0, 0.5, 0.0371, 3.0, 2.0 ** -7, 0.11, 1.7e-3, 40.0]
    for sv in scales:
        s = O.rnd(np.float32(sv), dtype)
        yo = O.int_quant_forward(x, s, 0.0, qmin, qmax, "round", dtype)
        co = O.int_quant_chain(x, s, 0.0, qmin, qmax, "round", dtype)
        sd = dev(np.asarray(s), dtype)
        y, codes = K.int_quant_fwd(allx.cuda(), sd, 0.0, qmin, qmax, 0, want_codes=True)
        assert_bits_equal(host(y), yo, f"y scale={sv}")
        assert_bits_equal(host(codes), co[-1] if isinstance(co, tuple) else co, f"codes scale={sv}")
        for cm in ("ste", "masked"):
            gxo, gs_el = O.int_quant_backward(g, x, s, 0.0, qmin, qmax, "round", cm, dtype)
            gx, gs = K.int_quant_bwd(dev(g, dtype), allx.cuda(), sd, 0.0, qmin, qmax, 0, CM[cm], True)
            assert_bits_equal(host(gx), gxo, f"gx {cm} scale={sv}")
            fin = np.isfinite(gs_el) & (np.abs(x) < 1e4) & (np.abs(x) > 1e-4)   # keep the fp32 sum far from overflow
            x_f = torch.from_numpy(np.where(fin, x, 1.0)).to(tdt).cuda()       # finite inputs only for the sum
            _, gs = K.int_quant_bwd(dev(g, dtype), x_f, sd, 0.0, qmin, qmax, 0, CM[cm], True)
            _, gs_ref = O.int_quant_backward(g, host(x_f), s, 0.0, qmin, qmax, "round", cm, dtype)
            mag = float(np.abs(gs_ref).sum()) + 1.0
            assert abs(float(gs) - float(gs_ref.astype(np.float64).sum())) <= mag * (65536 * 2.0 ** -21 + 16 * ulp(dtype))
    # per-row scales (planes kernel) and the fused abs-max kernel: 8 rows of 8192 covering all patterns
    xr = allx.view(8, 8192)
    srow = O.rnd(np.array([1.0, 0.25, 0.0371, 3.0, 0.5, 0.11, 2.0, 0.9], dtype=np.float32).reshape(8, 1), dtype)
    yo = O.int_quant_forward(xr.float().numpy(), srow, 0.0, qmin, qmax, "round", dtype)
    assert_bits_equal(host(K.int_quant_fwd(xr.cuda(), dev(srow, dtype), 0.0, qmin, qmax, 0)), yo, "rows provided")
    gxo, _ = O.int_quant_backward(g.reshape(8, 8192), xr.float().numpy(), srow, 0.0, qmin, qmax, "round", "masked", dtype)
    gx, _ = K.int_quant_bwd(dev(g.reshape(8, 8192), dtype), xr.cuda(), dev(srow, dtype), 0.0, qmin, qmax, 0, 1, True)
    assert_bits_equal(host(gx), gxo, "rows provided gx")
    # one scale per channel of a channels-last tensor (packed per-lane variant of int_quant_chanlast_kernel in bf16):
    # [1, 16, 64, 64] NHWC holds all 2^16 patterns, pattern i under scale[i % 16]; one scale outside the exact
    # multiply-by-reciprocal window sends its threads through the out-of-line literal branch
    sch = np.array([1.0, 0.25, 0.0371, 3.0, 0.5, 0.11, 2.0, 0.9, 1.7e-3, 40.0, 2.0 ** -7, 0.61, 7.0, 0.013, 1.0e-13, 0.3],
                   dtype=np.float32)
    if dtype == "f16":
        sch[14] = 6.1e-5
    sch = O.rnd(sch, dtype)
    xn = allx.view(1, 64, 64, 16).permute(0, 3, 1, 2)          # NCHW view of NHWC memory
    gn = torch.from_numpy(g).to(tdt).view(1, 64, 64, 16).permute(0, 3, 1, 2)
    assert xn.is_contiguous(memory_format=torch.channels_last)
    sn = sch.reshape(1, 16, 1, 1)
    yo = O.int_quant_forward(xn.float().numpy(), sn, 0.0, qmin, qmax, "round", dtype)
    co = O.int_quant_chain(xn.float().numpy(), sn, 0.0, qmin, qmax, "round", dtype)
    y, codes = K.int_quant_fwd(xn.cuda(), dev(sn, dtype), 0.0, qmin, qmax, 0, want_codes=True)
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert_bits_equal(host(y), yo, "channels-last y")
    assert_bits_equal(host(codes), co[-1] if isinstance(co, tuple) else co, "channels-last codes")
    for cm in ("ste", "masked"):
        gxo, _ = O.int_quant_backward(gn.float().numpy(), xn.float().numpy(), sn, 0.0, qmin, qmax, "round", cm, dtype)
        gx, _ = K.int_quant_bwd(gn.cuda(), xn.cuda(), dev(sn, dtype), 0.0, qmin, qmax, 0, CM[cm], True)
        assert_bits_equal(host(gx), gxo, f"channels-last gx {cm}")
    finite = torch.where(torch.isfinite(xr.float()), xr, torch.ones((), dtype=tdt)).contiguous()
    thr = max(abs(qmin), abs(qmax))
    yo, so, _ = O.rows_absmax_int_quant_forward(finite.float().numpy(), 1e-10, thr, 0.0, qmin, qmax, "round", dtype)
    y, sc, _ = K.rows_absmax_int_quant_fwd(finite.cuda(), 8, 8192, 1e-10, thr, 0.0, qmin, qmax, 0)
    assert_bits_equal(host(sc), so, "fused scale")
    assert_bits_equal(host(y), yo, "fused y")


@pytest.mark.parametrize("dtype", DTYPES)
def test_dpu_round_exhaustive(K, dtype):
    """dpu_round_ste (ops_ste.py dpu_round_ste / function/ops.py dpu_round): the kernels use a one-FRND formulation of
    where((x < 0) & (x - floor(x) == 0.5), ceil(x), round(x)); all 2^16 bf16 / fp16 inputs, and for fp32 every negative
    and positive half-integer up to 2^17 with its two neighbours, signed zeros, infinities and NaN, against the
    literal torch expression evaluated in the tensor's dtype on the same GPU"""
    tdt = TDT[dtype]
    if dtype == "f32":
        half = torch.arange(0, 1 << 17, dtype=torch.float32) + 0.5
        cand = torch.cat([half, -half, torch.tensor([0.0, -0.0, float("inf"), -float("inf"), float("nan"), -8388607.5,
                                                      8388607.5, -1e-45, -3.0e38, 16777216.0])])
        bits = cand.view(torch.int32)
        x = torch.cat([(bits + k).view(torch.float32) for k in (-1, 0, 1)]).cuda()
    else:
        x = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(tdt).cuda()
    frac = x - torch.floor(x)
    ref = torch.where((x < 0) & (frac == 0.5), torch.ceil(x), torch.round(x))
    got = K.unary("bvb_dpu_round_ste_impl", x)
    assert_bits_equal(host(got), host(ref), "dpu_round")
    # the same rounding inside the quantizer kernel: scale 1, wide range -> codes are dpu_round(x) where finite and in range
    xs = x[torch.isfinite(x.float()) & (x.float().abs() < 30000)].contiguous()
    one = torch.ones((), device="cuda", dtype=tdt)
    y = K.int_quant_fwd(xs, one, 0.0, -32768.0, 32767.0, 4)
    fr = xs - torch.floor(xs)
    refq = torch.where((xs < 0) & (fr == 0.5), torch.ceil(xs), torch.round(xs))
    assert torch.equal(y.float() + 0.0, refq.float() + 0.0)          # "+ 0" : the chain's "+ zero_point" turns -0 into +0


def test_more_than_2_31_elements(K):
    """maximum sizes: 2^31 + 4099 bf16 elements (4.3 GB per tensor) through the streaming forward, the TMA
    provided-scale backward and the fused per-row kernels; slices on both sides of the 2^31 boundary and at the tail
    must equal what small calls on the same data produce (64-bit indexing end to end)."""
    n = (1 << 31) + 4099
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    g = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for t in (x, g):
        for a in range(0, n, 1 << 28):
            t[a:a + (1 << 28)] = torch.randn(min(1 << 28, n - a), device="cuda", generator=gen).to(torch.bfloat16) * 20
    s = torch.tensor(0.37, device="cuda", dtype=torch.bfloat16)
    y = K.int_quant_fwd(x, s, 0.0, -128.0, 127.0, 0)
    gx, gs = K.int_quant_bwd(g, x, s, 0.0, -128.0, 127.0, 0, 1, True)
    spans = [(0, 70000), ((1 << 31) - 40000, (1 << 31) + 4000), (n - 5003, n)]
    for a, b in spans:
        ys = K.int_quant_fwd(x[a:b].clone(), s, 0.0, -128.0, 127.0, 0)
        gxs, _ = K.int_quant_bwd(g[a:b].clone(), x[a:b].clone(), s, 0.0, -128.0, 127.0, 0, 1, True)
        assert torch.equal(y[a:b], ys) and torch.equal(gx[a:b], gxs), (a, b)
    assert torch.isfinite(gs).all()
    del y, gx
    rows, cols = (1 << 20) + 2, 2048                       # 2^31 + 4096 elements as [rows, 2048]
    xr, gr = x[:rows * cols], g[:rows * cols]
    yr, sc, _ = K.rows_absmax_int_quant_fwd(xr, rows, cols, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
    gxr = K.rows_absmax_int_quant_bwd(gr, xr, sc, None, rows, cols, 127.0, 0.0, -127.0, 127.0, 0, 0)
    for r0 in (0, (1 << 20) - 3):
        sl = slice(r0 * cols, (r0 + 5) * cols)
        y5, s5, _ = K.rows_absmax_int_quant_fwd(xr[sl].clone(), 5, cols, 1e-10, 127.0, 0.0, -127.0, 127.0, 0)
        g5 = K.rows_absmax_int_quant_bwd(gr[sl].clone(), xr[sl].clone(), s5, None, 5, cols, 127.0, 0.0, -127.0, 127.0, 0, 0)
        assert torch.equal(yr[sl], y5) and torch.equal(sc[r0:r0 + 5], s5)
        assert torch.allclose(gxr[sl].float(), g5.float(), rtol=2e-2, atol=2e-2)
        first = torch.zeros(5 * cols, dtype=torch.bool, device="cuda")
        first[torch.arange(5, device="cuda") * cols + xr[sl].view(5, cols).abs().argmax(dim=1)] = True
        assert torch.equal(torch.where(first, 0, gxr[sl].float()), torch.where(first, 0, g5.float()))


def test_empty_and_errors(K):
    e = torch.empty(0, device="cuda")
    assert K.int_quant_fwd(e, torch.tensor(1.0, device="cuda"), 0.0, -1.0, 1.0, 0).numel() == 0
    assert K.unary("bvb_round_ste_impl", e).numel() == 0
    with pytest.raises(RuntimeError):
        K.unary("bvb_round_ste_impl", torch.randn(3))                 # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        K.unary("bvb_round_ste_impl", torch.randn(3, device="cuda").double())
    with pytest.raises(RuntimeError):
        K.abs_kth_value_rows(torch.randn(10, device="cuda"), 1, 10, 11)
    x = torch.randn(1 << 16, device="cuda")
    base = torch.randn((1 << 16) + 1, device="cuda")
    un = base[1:]                                                        # 4-byte aligned only
    un.copy_(x)
    s = torch.tensor(0.05, device="cuda")
    assert torch.equal(K.int_quant_fwd(un, s, 0.0, -127.0, 127.0, 0), K.int_quant_fwd(x, s, 0.0, -127.0, 127.0, 0))


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: properties + oracle on sampled rows
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype,rows,cols", [("f32", 4096, 11008), ("bf16", 4096, 11008), ("bf16", 16384, 4096)])
def test_full_size_properties(K, dtype, rows, cols):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(rows, cols, device="cuda", generator=g).to(TDT[dtype])
    signed_narrow = cols == 11008                   # C2: int8 narrow weights; C3: int8 full range activations
    qmin, qmax, thr = (-127.0, 127.0, 127.0) if signed_narrow else (-128.0, 127.0, 128.0)
    y, s, am = K.rows_absmax_int_quant_fwd(x, rows, cols, 1e-10, thr, 0.0, qmin, qmax, 0, want_absmax=True)
    # (1) statistic: matches torch's own reduction bit for bit (max is order independent)
    assert torch.equal(am, x.abs().amax(dim=1))
    # (2) codes are integers inside the range, and quantization is idempotent at fixed scale
    codes = K.int_quant_fwd(x, s.view(rows, 1), 0.0, qmin, qmax, 0, want_codes=True)[1].float()
    assert torch.equal(codes, codes.round()) and codes.min() >= qmin and codes.max() <= qmax
    y2 = K.int_quant_fwd(y, s.view(rows, 1), 0.0, qmin, qmax, 0)
    if dtype == "f32":
        assert (y2 - y).abs().max() <= s.max() * 1.0001    # re-quantizing moves at most one step (fp32 rounding)
    # (3) oracle, bit-exact, on sampled rows
    pick = [0, 1, rows // 2, rows - 1]
    xs = host(x[pick])
    yo, so, _ = O.rows_absmax_int_quant_forward(xs, 1e-10, thr, 0.0, qmin, qmax, "round", dtype)
    assert_bits_equal(host(y[pick]), yo, "sampled rows y")
    assert_bits_equal(host(s[pick]), so, "sampled rows scale")
    # (4) backward: STE-clamp gradient is linear in the incoming gradient away from the arg-max entries
    gr = torch.randn(rows, cols, device="cuda", generator=g).to(TDT[dtype])
    gx = K.rows_absmax_int_quant_bwd(gr, x, s, None, rows, cols, thr, 0.0, qmin, qmax, 0, 0)
    gxo, _ = O.rows_absmax_int_quant_backward(host(gr[pick]), xs, so, None, thr, 0.0, qmin, qmax, "round", "ste", dtype)
    first = np.zeros_like(xs, dtype=bool)
    first[np.arange(len(pick)), np.abs(xs).argmax(axis=1)] = True
    assert_bits_equal(np.where(first, 0, host(gx[pick])), np.where(first, 0, gxo), "sampled rows gx")
    close_sum(host(gx[pick])[first], gxo[first], cols, np.abs(gxo[first]).max() + 1.0, dtype)
