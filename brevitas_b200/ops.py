"""torch.library registration of the B200 fake-quant kernels.

Two dispatcher namespaces are defined:

* ``autograd_ste_ops`` -- the reference's own native-plugin namespace with its 12 op names
  (src/brevitas/csrc/autograd_ste_ops.cpp:258-271), so ``brevitas.function.ops_ste`` can dispatch to these
  kernels through ``torch.ops.autograd_ste_ops.<name>`` without source changes.
* ``brevitas_b200`` -- the fused quantizer ops used by the module layer (``brevitas_b200.core``).

Every op has a CUDA implementation only (ctypes -> C-ABI -> sm_100a kernels), an autograd formula and a
fake/meta implementation.  Calling any of them with CPU tensors raises: there is no CPU fallback (inside the one scoped
exception, ``parameter_init_on_host`` below, host tensors are staged to the GPU for a construction-time initialisation).
"""
import contextlib
import threading
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.library import Library

from . import _kernels as K
from . import _lib

# the reference's native extension must not be loaded in the same process (duplicate TORCH_LIBRARY DEF)
_STE = Library("autograd_ste_ops", "DEF")
_FQ = Library("brevitas_b200", "DEF")

STE_NS = "autograd_ste_ops"
FQ_NS = "brevitas_b200"


class _HostInit(threading.local):
    depth = 0


_HOST_INIT = _HostInit()


@contextlib.contextmanager
def parameter_init_on_host():
    """Scope of a ONE-OFF parameter initialisation that the reference evaluates at layer construction, before the
    user had a chance to move the layer to the GPU: ``ParameterFromStatsScalingInit.__call__`` (quant/solver/
    parameter.py:39-45) runs a ``StatsFromParameterScaling`` over the still host-resident weight to produce the initial
    value of a learned scale.  Inside this scope -- and nowhere else -- the handful of ops such an initialiser is made of
    accept host tensors: they are STAGED to the current CUDA device, the sm_100a kernel runs there, and the (statistics-
    sized) result is copied back, so the value is computed by the same kernels as everything else and a machine without
    a GPU fails loudly here as well.  ``binding.install()`` enters the scope around exactly that call.  Outside it every op
    raises on host tensors: there is no CPU arithmetic in this package's ops."""
    _HOST_INIT.depth += 1
    try:
        yield
    finally:
        _HOST_INIT.depth -= 1


def _no_cpu(name, staged=None):
    """the CPU dispatch entry of an op: raises -- except for the ops a construction-time initialiser uses (``staged`` =
    (namespace, op name)), which inside ``parameter_init_on_host`` run their CUDA kernel on a staged copy"""
    def _raise(*args, **kwargs):
        if staged is not None and _HOST_INIT.depth > 0:
            if not torch.cuda.is_available():
                raise RuntimeError(f"{name}: initialising a learned scale from the weight statistics runs the sm_100a "
                                   "kernels and needs a CUDA device (brevitas_b200 has no CPU implementation)")
            dev = torch.device("cuda", torch.cuda.current_device())
            moved = [a.to(dev) if isinstance(a, Tensor) else a for a in args]
            out = getattr(getattr(torch.ops, staged[0]), staged[1])(*moved, **kwargs)
            if isinstance(out, Tensor):
                return out.cpu()
            return tuple(o.cpu() if isinstance(o, Tensor) else o for o in out)
        raise RuntimeError(
            f"{name}: CPU tensors are not supported -- brevitas_b200 is a CUDA (sm_100a) implementation with no "
            "CPU fallback. Move the module and its inputs to a B200.")
    return _raise


# ============================================================================================================
# 1. autograd_ste_ops: the 12 STE primitives
# ============================================================================================================

def _identity_backward(ctx, grad):
    # csrc/autograd_ste_ops.cpp:22 returns grad_output[0] itself
    return grad


_UNARY_STE = {
    "round_ste_impl": "bvb_round_ste_impl",                  # csrc:14-24
    "ceil_ste_impl": "bvb_ceil_ste_impl",                    # csrc:100-110
    "floor_ste_impl": "bvb_floor_ste_impl",                  # csrc:112-122
    "binary_sign_ste_impl": "bvb_binary_sign_ste_impl",      # csrc:124-137
    "ternary_sign_ste_impl": "bvb_ternary_sign_ste_impl",    # csrc:140-150
    "round_to_zero_ste_impl": "bvb_round_to_zero_ste_impl",  # csrc:153-163
    "dpu_round_ste_impl": "bvb_dpu_round_ste_impl",          # csrc:166-179
}


# float_to_int of a power-of-two restriction inside a construction-time scale initialiser (see parameter_init_on_host)
_UNARY_INIT_OPS = ("round_ste_impl", "ceil_ste_impl", "floor_ste_impl")


def _def_unary(op_name, c_name):
    _STE.define(f"{op_name}(Tensor x) -> Tensor")
    _STE.impl(op_name, lambda x, _c=c_name: K.unary(_c, x), "CUDA")
    _STE.impl(op_name, _no_cpu(op_name, (STE_NS, op_name) if op_name in _UNARY_INIT_OPS else None), "CPU")
    torch.library.register_fake(f"{STE_NS}::{op_name}", lambda x: torch.empty_like(x), lib=_STE)
    torch.library.register_autograd(f"{STE_NS}::{op_name}", _identity_backward, lib=_STE)


for _op, _c in _UNARY_STE.items():
    _def_unary(_op, _c)

# abs_binary_sign_grad_impl: forward abs, backward binary_sign(x) * g   (csrc:182-194)
_STE.define("abs_binary_sign_grad_impl(Tensor x) -> Tensor")
_STE.impl("abs_binary_sign_grad_impl", lambda x: K.unary("bvb_abs_binary_sign_grad_impl", x), "CUDA")
_STE.impl("abs_binary_sign_grad_impl", _no_cpu("abs_binary_sign_grad_impl", (STE_NS, "abs_binary_sign_grad_impl")), "CPU")
torch.library.register_fake(f"{STE_NS}::abs_binary_sign_grad_impl", lambda x: torch.empty_like(x), lib=_STE)

_FQ.define("abs_binary_sign_grad_backward(Tensor x, Tensor gy) -> Tensor")
_FQ.impl("abs_binary_sign_grad_backward", lambda x, gy: K.abs_binary_sign_grad_bwd(x, gy), "CUDA")
_FQ.impl("abs_binary_sign_grad_backward", _no_cpu("abs_binary_sign_grad_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::abs_binary_sign_grad_backward", lambda x, gy: torch.empty_like(gy), lib=_FQ)


def _absgrad_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0])


def _absgrad_backward(ctx, grad):
    (x,) = ctx.saved_tensors
    return torch.ops.brevitas_b200.abs_binary_sign_grad_backward(x, grad.to(x.dtype))


torch.library.register_autograd(f"{STE_NS}::abs_binary_sign_grad_impl", _absgrad_backward,
                                setup_context=_absgrad_setup, lib=_STE)

# tensor_clamp_ste_impl: where-clamp with tensor bounds, gradient (g, None, None)   (csrc:27-44)
_STE.define("tensor_clamp_ste_impl(Tensor x, Tensor min_val, Tensor max_val) -> Tensor")
_STE.impl("tensor_clamp_ste_impl", lambda x, mn, mx: K.tensor_clamp(x, mn, mx, inplace=False), "CUDA")
_STE.impl("tensor_clamp_ste_impl", _no_cpu("tensor_clamp_ste_impl"), "CPU")
torch.library.register_fake(f"{STE_NS}::tensor_clamp_ste_impl", lambda x, mn, mx: torch.empty_like(x), lib=_STE)
torch.library.register_autograd(f"{STE_NS}::tensor_clamp_ste_impl", lambda ctx, g: (g, None, None), lib=_STE)

# tensor_clamp_ste_impl_: in place, returns its input.  Follows the reference's PYTHON backend
# (ops/autograd_ste_ops.py:146-148); the C++ backend binds this name to the out-of-place function by mistake
# (csrc:261).
_STE.define("tensor_clamp_ste_impl_(Tensor(a!) x, Tensor min_val, Tensor max_val) -> Tensor(a!)")
_STE.impl("tensor_clamp_ste_impl_", lambda x, mn, mx: K.tensor_clamp(x, mn, mx, inplace=True), "CUDA")
_STE.impl("tensor_clamp_ste_impl_", _no_cpu("tensor_clamp_ste_impl_"), "CPU")


class _InplaceTensorClampSte(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, min_val, max_val):
        # no ctx.mark_dirty: like the reference's Python backend (ops/autograd_ste_ops.py:146-148) this mutates its
        # input behind autograd's back, which is what lets BinaryQuant-style callers clamp a leaf weight in place
        with torch._C._AutoDispatchBelowAutograd():
            torch.ops.autograd_ste_ops.tensor_clamp_ste_impl_(x, min_val, max_val)
        return x

    @staticmethod
    def backward(ctx, grad):
        return grad, None, None


_STE.impl("tensor_clamp_ste_impl_", lambda x, mn, mx: _InplaceTensorClampSte.apply(x, mn, mx), "Autograd")

# scalar clamps (csrc:66-97)
_STE.define("scalar_clamp_ste_impl(Tensor x, float min_val, float max_val) -> Tensor")
_STE.impl("scalar_clamp_ste_impl", lambda x, lo, hi: K.scalar_clamp(x, lo, hi), "CUDA")
_STE.impl("scalar_clamp_ste_impl", _no_cpu("scalar_clamp_ste_impl"), "CPU")
torch.library.register_fake(f"{STE_NS}::scalar_clamp_ste_impl", lambda x, lo, hi: torch.empty_like(x), lib=_STE)
torch.library.register_autograd(f"{STE_NS}::scalar_clamp_ste_impl", lambda ctx, g: (g, None, None), lib=_STE)

_STE.define("scalar_clamp_min_ste_impl(Tensor x, float min_val) -> Tensor")
_STE.impl("scalar_clamp_min_ste_impl", lambda x, lo: K.scalar_clamp_min(x, lo), "CUDA")
_STE.impl("scalar_clamp_min_ste_impl", _no_cpu("scalar_clamp_min_ste_impl", (STE_NS, "scalar_clamp_min_ste_impl")), "CPU")
torch.library.register_fake(f"{STE_NS}::scalar_clamp_min_ste_impl", lambda x, lo: torch.empty_like(x), lib=_STE)
torch.library.register_autograd(f"{STE_NS}::scalar_clamp_min_ste_impl", lambda ctx, g: (g, None), lib=_STE)

STE_OP_NAMES = tuple(_UNARY_STE) + ("abs_binary_sign_grad_impl", "tensor_clamp_ste_impl", "tensor_clamp_ste_impl_",
                                    "scalar_clamp_ste_impl", "scalar_clamp_min_ste_impl")


# ============================================================================================================
# 2. brevitas_b200: fused quantizer ops
# ============================================================================================================

def _reduce_gscale(gs: Optional[Tensor], scale: Tensor) -> Optional[Tensor]:
    if gs is None:
        return None
    return gs.to(scale.dtype).view(scale.shape)


# ---- IntQuant with provided scale --------------------------------------------------------------------------
_FQ.define("int_quant(Tensor x, Tensor scale, float zero_point, float qmin, float qmax, int round_mode, "
           "int clamp_mode) -> Tensor")
_FQ.define("int_quant_codes(Tensor x, Tensor scale, float zero_point, float qmin, float qmax, int round_mode) "
           "-> (Tensor, Tensor)")
_FQ.define("int_quant_backward(Tensor gy, Tensor x, Tensor scale, float zero_point, float qmin, float qmax, "
           "int round_mode, int clamp_mode, bool want_gscale) -> (Tensor, Tensor)")


def _int_quant_cuda(x, scale, zp, qmin, qmax, rm, cm):
    return K.int_quant_fwd(x, scale, zp, qmin, qmax, rm)


def _int_quant_codes_cuda(x, scale, zp, qmin, qmax, rm):
    return K.int_quant_fwd(x, scale, zp, qmin, qmax, rm, want_codes=True)


def _int_quant_backward_cuda(gy, x, scale, zp, qmin, qmax, rm, cm, want_gscale):
    gx, gs = K.int_quant_bwd(gy, x, scale, zp, qmin, qmax, rm, cm, want_gscale)
    if gs is None:
        gs = torch.empty(0, dtype=torch.float32, device=x.device)
    return gx, gs


_FQ.impl("int_quant", _int_quant_cuda, "CUDA")
_FQ.impl("int_quant", _no_cpu("int_quant"), "CPU")
_FQ.impl("int_quant_codes", _int_quant_codes_cuda, "CUDA")
_FQ.impl("int_quant_codes", _no_cpu("int_quant_codes"), "CPU")
_FQ.impl("int_quant_backward", _int_quant_backward_cuda, "CUDA")
_FQ.impl("int_quant_backward", _no_cpu("int_quant_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::int_quant", lambda x, s, zp, a, b, rm, cm: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(f"{FQ_NS}::int_quant_codes",
                            lambda x, s, zp, a, b, rm: (torch.empty_like(x), torch.empty_like(x)), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::int_quant_backward",
    lambda gy, x, s, zp, a, b, rm, cm, w: (torch.empty_like(x), x.new_empty(s.numel() if w else 0, dtype=torch.float32)),
    lib=_FQ)


def _int_quant_setup(ctx, inputs, output):
    x, scale, zp, qmin, qmax, rm, cm = inputs
    ctx.save_for_backward(x, scale)
    ctx.q = (zp, qmin, qmax, rm, cm)


def _int_quant_bwd(ctx, gy):
    x, scale = ctx.saved_tensors
    zp, qmin, qmax, rm, cm = ctx.q
    want_gs = ctx.needs_input_grad[1]
    gx, gs = torch.ops.brevitas_b200.int_quant_backward(gy.to(x.dtype), x, scale, zp, qmin, qmax, rm, cm, want_gs)
    return (gx if ctx.needs_input_grad[0] else None, _reduce_gscale(gs, scale) if want_gs else None,
            None, None, None, None, None)


torch.library.register_autograd(f"{FQ_NS}::int_quant", _int_quant_bwd, setup_context=_int_quant_setup, lib=_FQ)


# ---- IntQuant with a tensor-valued zero-point (asymmetric quantizers) -----------------------------------------------
_FQ.define("int_quant_zpt(Tensor x, Tensor scale, Tensor zero_point, float qmin, float qmax, int round_mode, "
           "int clamp_mode) -> Tensor")
_FQ.define("int_quant_zpt_backward(Tensor gy, Tensor x, Tensor scale, Tensor zero_point, float qmin, float qmax, "
           "int round_mode, int clamp_mode, bool want_grads) -> (Tensor, Tensor, Tensor)")


def _int_quant_zpt_backward_cuda(gy, x, scale, zp, qmin, qmax, rm, cm, want):
    gx, gs, gz = K.int_quant_zpt_bwd(gy, x, scale, zp, qmin, qmax, rm, cm, want)
    if gs is None:
        gs = torch.empty(0, dtype=torch.float32, device=x.device)
        gz = torch.empty(0, dtype=torch.float32, device=x.device)
    return gx, gs, gz


_FQ.impl("int_quant_zpt", lambda x, s, z, a, b, rm, cm: K.int_quant_zpt_fwd(x, s, z, a, b, rm), "CUDA")
_FQ.impl("int_quant_zpt", _no_cpu("int_quant_zpt"), "CPU")
_FQ.impl("int_quant_zpt_backward", _int_quant_zpt_backward_cuda, "CUDA")
_FQ.impl("int_quant_zpt_backward", _no_cpu("int_quant_zpt_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::int_quant_zpt", lambda x, s, z, a, b, rm, cm: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::int_quant_zpt_backward",
    lambda gy, x, s, z, a, b, rm, cm, w: (torch.empty_like(x), x.new_empty(s.numel() if w else 0, dtype=torch.float32),
                                          x.new_empty(s.numel() if w else 0, dtype=torch.float32)),
    lib=_FQ)


def _int_quant_zpt_setup(ctx, inputs, output):
    x, scale, zp, qmin, qmax, rm, cm = inputs
    ctx.save_for_backward(x, scale, zp)
    ctx.q = (qmin, qmax, rm, cm)


def _int_quant_zpt_bwd(ctx, gy):
    x, scale, zp = ctx.saved_tensors
    qmin, qmax, rm, cm = ctx.q
    want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
    gx, gs, gz = torch.ops.brevitas_b200.int_quant_zpt_backward(gy.to(x.dtype), x, scale, zp, qmin, qmax, rm, cm, want)
    return (gx if ctx.needs_input_grad[0] else None,
            _reduce_gscale(gs, scale) if ctx.needs_input_grad[1] else None,
            _reduce_gscale(gz, zp) if ctx.needs_input_grad[2] else None, None, None, None, None)


torch.library.register_autograd(f"{FQ_NS}::int_quant_zpt", _int_quant_zpt_bwd, setup_context=_int_quant_zpt_setup, lib=_FQ)


# ---- integer export (IntQuant.to_int + cast; QuantTensor.int(), quant_tensor/__init__.py:174-187) -------------------
_FQ.define("int_quant_to_int(Tensor x, Tensor scale, float zero_point, float? qmin, float? qmax, int round_mode, "
           "ScalarType out_dtype) -> Tensor")
_FQ.impl("int_quant_to_int", lambda x, s, zp, a, b, rm, dt: K.int_quant_to_int(x, s, zp, a, b, rm, dt), "CUDA")
_FQ.impl("int_quant_to_int", _no_cpu("int_quant_to_int"), "CPU")
torch.library.register_fake(f"{FQ_NS}::int_quant_to_int",
                            lambda x, s, zp, a, b, rm, dt: torch.empty_like(x, dtype=dt), lib=_FQ)


# ---- ReLU fused in front of IntQuant (QuantReLU = nn.ReLU + act quantizer, proxy/runtime_quant.py:73-84) ----------
_FQ.define("relu_int_quant(Tensor x, Tensor scale, float zero_point, float qmin, float qmax, int round_mode, "
           "int clamp_mode) -> Tensor")
_FQ.define("relu_int_quant_backward(Tensor gy, Tensor x, Tensor scale, float zero_point, float qmin, float qmax, "
           "int round_mode, int clamp_mode, bool want_gscale) -> (Tensor, Tensor)")


def _relu_int_quant_cuda(x, scale, zp, qmin, qmax, rm, cm):
    return K.int_quant_fwd(x, scale, zp, qmin, qmax, rm, pre_relu=True)


def _relu_int_quant_backward_cuda(gy, x, scale, zp, qmin, qmax, rm, cm, want_gscale):
    gx, gs = K.int_quant_bwd(gy, x, scale, zp, qmin, qmax, rm, cm, want_gscale, pre_relu=True)
    if gs is None:
        gs = torch.empty(0, dtype=torch.float32, device=x.device)
    return gx, gs


_FQ.impl("relu_int_quant", _relu_int_quant_cuda, "CUDA")
_FQ.impl("relu_int_quant", _no_cpu("relu_int_quant"), "CPU")
_FQ.impl("relu_int_quant_backward", _relu_int_quant_backward_cuda, "CUDA")
_FQ.impl("relu_int_quant_backward", _no_cpu("relu_int_quant_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::relu_int_quant", lambda x, s, zp, a, b, rm, cm: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::relu_int_quant_backward",
    lambda gy, x, s, zp, a, b, rm, cm, want: (torch.empty_like(x), torch.empty(max(1, s.numel()) if want else 0,
                                                                                dtype=torch.float32, device=x.device)),
    lib=_FQ)


def _relu_int_quant_bwd(ctx, gy):
    x, scale = ctx.saved_tensors
    zp, qmin, qmax, rm, cm = ctx.q
    want_gs = ctx.needs_input_grad[1]
    gx, gs = torch.ops.brevitas_b200.relu_int_quant_backward(gy.to(x.dtype), x, scale, zp, qmin, qmax, rm, cm, want_gs)
    return (gx if ctx.needs_input_grad[0] else None, _reduce_gscale(gs, scale) if want_gs else None,
            None, None, None, None, None)


torch.library.register_autograd(f"{FQ_NS}::relu_int_quant", _relu_int_quant_bwd, setup_context=_int_quant_setup, lib=_FQ)


# ---- fused per-row abs-max + IntQuant ----------------------------------------------------------------------
_FQ.define("rows_absmax_int_quant(Tensor x, int rows, int cols, float scaling_min_val, float int_threshold, "
           "float zero_point, float qmin, float qmax, int round_mode, int clamp_mode) -> (Tensor, Tensor, Tensor)")
_FQ.define("rows_absmax_int_quant_backward(Tensor gy, Tensor x, Tensor scale, Tensor? gscale, int rows, int cols, "
           "float int_threshold, float zero_point, float qmin, float qmax, int round_mode, int clamp_mode) -> Tensor")


def _rows_fwd_cuda(x, rows, cols, min_val, int_thr, zp, qmin, qmax, rm, cm):
    return K.rows_absmax_int_quant_fwd(x, rows, cols, min_val, int_thr, zp, qmin, qmax, rm, want_absmax=True)


def _rows_bwd_cuda(gy, x, scale, gscale, rows, cols, int_thr, zp, qmin, qmax, rm, cm):
    return K.rows_absmax_int_quant_bwd(gy, x, scale, gscale, rows, cols, int_thr, zp, qmin, qmax, rm, cm)


_FQ.impl("rows_absmax_int_quant", _rows_fwd_cuda, "CUDA")
_FQ.impl("rows_absmax_int_quant", _no_cpu("rows_absmax_int_quant"), "CPU")
_FQ.impl("rows_absmax_int_quant_backward", _rows_bwd_cuda, "CUDA")
_FQ.impl("rows_absmax_int_quant_backward", _no_cpu("rows_absmax_int_quant_backward"), "CPU")
torch.library.register_fake(
    f"{FQ_NS}::rows_absmax_int_quant",
    lambda x, rows, cols, mv, it, zp, a, b, rm, cm: (torch.empty_like(x), x.new_empty(rows), x.new_empty(rows)), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::rows_absmax_int_quant_backward",
    lambda gy, x, s, gs, rows, cols, it, zp, a, b, rm, cm: torch.empty_like(x), lib=_FQ)


def _rows_setup(ctx, inputs, output):
    x, rows, cols, mv, it, zp, qmin, qmax, rm, cm = inputs
    y, scale, absmax = output
    ctx.save_for_backward(x, scale)
    ctx.q = (rows, cols, it, zp, qmin, qmax, rm, cm)
    ctx.mark_non_differentiable(absmax)
    ctx.set_materialize_grads(False)


def _rows_bwd(ctx, gy, gscale, gabsmax):
    x, scale = ctx.saved_tensors
    rows, cols, it, zp, qmin, qmax, rm, cm = ctx.q
    if gy is None:
        gy = torch.zeros_like(x)
    if gscale is not None:
        gscale = gscale.to(x.dtype).reshape(rows)
    gx = torch.ops.brevitas_b200.rows_absmax_int_quant_backward(gy.to(x.dtype), x, scale, gscale, rows, cols, it, zp,
                                                                qmin, qmax, rm, cm)
    return (gx,) + (None,) * 9


torch.library.register_autograd(f"{FQ_NS}::rows_absmax_int_quant", _rows_bwd, setup_context=_rows_setup, lib=_FQ)


# ---- fused whole-tensor abs-max + IntQuant -----------------------------------------------------------------
_FQ.define("tensor_absmax_int_quant(Tensor x, ScalarType scale_dtype, float scaling_min_val, float int_threshold, "
           "float zero_point, float qmin, float qmax, int round_mode, int clamp_mode) -> (Tensor, Tensor, Tensor)")
_FQ.define("tensor_absmax_int_quant_backward(Tensor gy, Tensor x, Tensor scale, Tensor absmax, Tensor? gscale, "
           "float int_threshold, float zero_point, float qmin, float qmax, int round_mode, int clamp_mode) -> Tensor")


def _tensor_fwd_cuda(x, scale_dtype, mv, it, zp, qmin, qmax, rm, cm):
    return K.tensor_absmax_int_quant_fwd(x, scale_dtype, mv, it, zp, qmin, qmax, rm)


def _tensor_bwd_cuda(gy, x, scale, absmax, gscale, it, zp, qmin, qmax, rm, cm):
    return K.tensor_absmax_int_quant_bwd(gy, x, scale, absmax, gscale, it, zp, qmin, qmax, rm, cm)


_FQ.impl("tensor_absmax_int_quant", _tensor_fwd_cuda, "CUDA")
_FQ.impl("tensor_absmax_int_quant", _no_cpu("tensor_absmax_int_quant"), "CPU")
_FQ.impl("tensor_absmax_int_quant_backward", _tensor_bwd_cuda, "CUDA")
_FQ.impl("tensor_absmax_int_quant_backward", _no_cpu("tensor_absmax_int_quant_backward"), "CPU")
torch.library.register_fake(
    f"{FQ_NS}::tensor_absmax_int_quant",
    lambda x, sd, mv, it, zp, a, b, rm, cm: (torch.empty_like(x), x.new_empty((), dtype=sd), x.new_empty(())), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::tensor_absmax_int_quant_backward",
    lambda gy, x, s, am, gs, it, zp, a, b, rm, cm: torch.empty_like(x), lib=_FQ)


def _tensor_setup(ctx, inputs, output):
    x, sd, mv, it, zp, qmin, qmax, rm, cm = inputs
    y, scale, absmax = output
    ctx.save_for_backward(x, scale, absmax)
    ctx.q = (it, zp, qmin, qmax, rm, cm)
    ctx.mark_non_differentiable(absmax)
    ctx.set_materialize_grads(False)


def _tensor_bwd(ctx, gy, gscale, gabsmax):
    x, scale, absmax = ctx.saved_tensors
    it, zp, qmin, qmax, rm, cm = ctx.q
    if gy is None:
        gy = torch.zeros_like(x)
    if gscale is not None:
        gscale = gscale.to(scale.dtype).reshape(())
    gx = torch.ops.brevitas_b200.tensor_absmax_int_quant_backward(gy.to(x.dtype), x, scale, absmax, gscale, it, zp,
                                                                  qmin, qmax, rm, cm)
    return (gx,) + (None,) * 8


torch.library.register_autograd(f"{FQ_NS}::tensor_absmax_int_quant", _tensor_bwd, setup_context=_tensor_setup, lib=_FQ)


# ---- BinaryQuant / ClampedBinaryQuant ----------------------------------------------------------------------
_FQ.define("binary_quant(Tensor x, Tensor scale, bool clamped) -> Tensor")
_FQ.define("binary_quant_backward(Tensor gy, Tensor x, Tensor scale, bool clamped, bool want_gscale) -> (Tensor, Tensor)")


def _binary_bwd_cuda(gy, x, scale, clamped, want_gscale):
    gx, gs = K.binary_quant_bwd(gy, x, scale, clamped, want_gscale)
    if gs is None:
        gs = torch.empty(0, dtype=torch.float32, device=x.device)
    return gx, gs


_FQ.impl("binary_quant", lambda x, s, c: K.binary_quant_fwd(x, s, c), "CUDA")
_FQ.impl("binary_quant", _no_cpu("binary_quant"), "CPU")
_FQ.impl("binary_quant_backward", _binary_bwd_cuda, "CUDA")
_FQ.impl("binary_quant_backward", _no_cpu("binary_quant_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::binary_quant", lambda x, s, c: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::binary_quant_backward",
    lambda gy, x, s, c, w: (torch.empty_like(x), x.new_empty(s.numel() if w else 0, dtype=torch.float32)), lib=_FQ)


def _binary_setup(ctx, inputs, output):
    x, scale, clamped = inputs
    ctx.save_for_backward(x, scale)
    ctx.clamped = clamped


def _binary_bwd(ctx, gy):
    x, scale = ctx.saved_tensors
    want_gs = ctx.needs_input_grad[1]
    gx, gs = torch.ops.brevitas_b200.binary_quant_backward(gy.to(x.dtype), x, scale, ctx.clamped, want_gs)
    return (gx if ctx.needs_input_grad[0] else None, _reduce_gscale(gs, scale) if want_gs else None, None)


torch.library.register_autograd(f"{FQ_NS}::binary_quant", _binary_bwd, setup_context=_binary_setup, lib=_FQ)


# ---- general integer quantizer: DecoupledIntQuant, and IntQuant with a device-resident range (learned bit-width) -------
_FQ.define("general_int_quant(Tensor x, Tensor pre_scale, Tensor scale, Tensor pre_zero_point, Tensor zero_point, "
           "Tensor min_int, Tensor max_int, int round_mode, int clamp_mode, bool same_scale) -> Tensor")
_FQ.define("general_int_quant_backward(Tensor gy, Tensor x, Tensor pre_scale, Tensor scale, Tensor pre_zero_point, "
           "Tensor zero_point, Tensor min_int, Tensor max_int, int round_mode, int clamp_mode, bool same_scale, "
           "bool want_sums) -> (Tensor, Tensor)")


def _general_bwd_cuda(gy, x, ps, s, pzp, zp, lo, hi, rm, cm, same, want):
    gx, sums = K.general_int_quant_bwd(gy, x, ps, s, pzp, zp, lo, hi, rm, cm, same, want)
    if sums is None:
        sums = torch.empty(0, dtype=torch.float64, device=x.device)
    return gx, sums


_FQ.impl("general_int_quant", lambda x, ps, s, pzp, zp, lo, hi, rm, cm, same: K.general_int_quant_fwd(
    x, ps, s, pzp, zp, lo, hi, rm), "CUDA")
_FQ.impl("general_int_quant", _no_cpu("general_int_quant"), "CPU")
_FQ.impl("general_int_quant_backward", _general_bwd_cuda, "CUDA")
_FQ.impl("general_int_quant_backward", _no_cpu("general_int_quant_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::general_int_quant",
                            lambda x, ps, s, pzp, zp, lo, hi, rm, cm, same: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::general_int_quant_backward",
    lambda gy, x, ps, s, pzp, zp, lo, hi, rm, cm, same, want: (
        torch.empty_like(x), x.new_empty(ps.numel() + s.numel() + 2 if want else 0, dtype=torch.float64)), lib=_FQ)


def _general_setup(ctx, inputs, output):
    x, ps, s, pzp, zp, lo, hi, rm, cm, same = inputs
    ctx.save_for_backward(x, ps, s, pzp, zp, lo, hi)
    ctx.q = (rm, cm, same)


def _general_bwd(ctx, gy):
    x, ps, s, pzp, zp, lo, hi = ctx.saved_tensors
    rm, cm, same = ctx.q
    need = ctx.needs_input_grad
    if need[3] or need[4]:
        raise RuntimeError("brevitas_b200::general_int_quant has no gradient for its zero-points (callers with a learned "
                           "zero-point take the int_quant_zpt kernel or the literal sequence)")
    want = need[1] or need[2] or need[5] or need[6]
    gx, sums = torch.ops.brevitas_b200.general_int_quant_backward(gy.to(x.dtype), x, ps, s, pzp, zp, lo, hi, rm, cm, same, want)
    pc, sc = ps.numel(), s.numel()
    g_ps = g_s = g_lo = g_hi = None
    if need[1] and not same:
        g_ps = sums[:pc].to(ps.dtype).view(ps.shape)
    if need[2] or (same and need[1]):
        g_s = sums[pc:pc + sc].to(s.dtype).view(s.shape)          # same_scale: the whole d(scale), handed to one input
    if cm == 1:                                                  # masked clamp: torch.where routes the clipped gradient to
        if need[5]:                                              # the bounds (function/ops.py:98-99)
            g_lo = sums[pc + sc].to(lo.dtype).view(lo.shape)
        if need[6]:
            g_hi = sums[pc + sc + 1].to(hi.dtype).view(hi.shape)
    return (gx if need[0] else None, g_ps, g_s, None, None, g_lo, g_hi, None, None, None)


torch.library.register_autograd(f"{FQ_NS}::general_int_quant", _general_bwd, setup_context=_general_setup, lib=_FQ)


# ---- TernaryQuant (fp32) ---------------------------------------------------------------------------------------------
_FQ.define("ternary_quant(Tensor x, Tensor scale, float threshold) -> Tensor")
_FQ.define("ternary_quant_backward(Tensor gy, Tensor x, Tensor scale, float threshold, bool want_gscale) -> (Tensor, Tensor)")


def _ternary_bwd_cuda(gy, x, scale, threshold, want):
    gx, gs = K.ternary_quant_bwd(gy, x, scale, threshold, want)
    if gs is None:
        gs = torch.empty(0, dtype=torch.float64, device=x.device)
    return gx, gs


_FQ.impl("ternary_quant", lambda x, s, t: K.ternary_quant_fwd(x, s, t), "CUDA")
_FQ.impl("ternary_quant", _no_cpu("ternary_quant"), "CPU")
_FQ.impl("ternary_quant_backward", _ternary_bwd_cuda, "CUDA")
_FQ.impl("ternary_quant_backward", _no_cpu("ternary_quant_backward"), "CPU")
torch.library.register_fake(f"{FQ_NS}::ternary_quant", lambda x, s, t: torch.empty_like(x), lib=_FQ)
torch.library.register_fake(
    f"{FQ_NS}::ternary_quant_backward",
    lambda gy, x, s, t, w: (torch.empty_like(x), x.new_empty(1 if w else 0, dtype=torch.float64)), lib=_FQ)


def _ternary_setup(ctx, inputs, output):
    x, scale, threshold = inputs
    ctx.save_for_backward(x, scale)
    ctx.threshold = threshold


def _ternary_bwd(ctx, gy):
    x, scale = ctx.saved_tensors
    want = ctx.needs_input_grad[1]
    gx, gs = torch.ops.brevitas_b200.ternary_quant_backward(gy.to(x.dtype), x, scale, ctx.threshold, want)
    return (gx if ctx.needs_input_grad[0] else None, gs.to(scale.dtype).view(scale.shape) if want else None, None)


torch.library.register_autograd(f"{FQ_NS}::ternary_quant", _ternary_bwd, setup_context=_ternary_setup, lib=_FQ)


# ---- statistics (forward values; see brevitas_b200.core.stats for the module wrappers) ------------------------
_FQ.define("absmax_rows(Tensor x, int rows, int cols) -> Tensor")
_FQ.define("absmax_tensor(Tensor x) -> Tensor")
_FQ.define("abs_kth_value_rows(Tensor x, int rows, int cols, int k) -> (Tensor, Tensor)")
_FQ.define("kth_value_rows(Tensor x, int rows, int cols, int k) -> (Tensor, Tensor)")
_FQ.define("running_stats_update_(Tensor(a!) running, Tensor stat, float momentum, bool first) -> ()")
_FQ.impl("absmax_rows", lambda x, r, c: K.absmax_rows(x, r, c), "CUDA")
_FQ.impl("absmax_rows", _no_cpu("absmax_rows", (FQ_NS, "absmax_rows")), "CPU")
_FQ.impl("absmax_tensor", lambda x: K.absmax_tensor(x), "CUDA")
_FQ.impl("absmax_tensor", _no_cpu("absmax_tensor", (FQ_NS, "absmax_tensor")), "CPU")
_FQ.impl("abs_kth_value_rows", lambda x, r, c, k: K.abs_kth_value_rows(x, r, c, k, want_index=True), "CUDA")
_FQ.impl("abs_kth_value_rows", _no_cpu("abs_kth_value_rows", (FQ_NS, "abs_kth_value_rows")), "CPU")
_FQ.impl("kth_value_rows", lambda x, r, c, k: K.kth_value_rows(x, r, c, k, want_index=True), "CUDA")
_FQ.impl("kth_value_rows", _no_cpu("kth_value_rows", (FQ_NS, "kth_value_rows")), "CPU")
_FQ.impl("running_stats_update_", lambda r, s, m, f: (K.running_stats_update(r, s, m, f), None)[1], "CUDA")
_FQ.impl("running_stats_update_", _no_cpu("running_stats_update_"), "CPU")
torch.library.register_fake(f"{FQ_NS}::absmax_rows", lambda x, r, c: x.new_empty(r), lib=_FQ)
torch.library.register_fake(f"{FQ_NS}::absmax_tensor", lambda x: x.new_empty(()), lib=_FQ)
torch.library.register_fake(f"{FQ_NS}::abs_kth_value_rows",
                            lambda x, r, c, k: (x.new_empty(r), x.new_empty(r, dtype=torch.int64)), lib=_FQ)
torch.library.register_fake(f"{FQ_NS}::kth_value_rows",
                            lambda x, r, c, k: (x.new_empty(r), x.new_empty(r, dtype=torch.int64)), lib=_FQ)


def _stat_setup_rows(ctx, inputs, output):
    x, rows, cols = inputs
    ctx.save_for_backward(x, output)
    ctx.rc = (rows, cols)


def _stat_bwd_rows(ctx, g):
    # AbsMax(dim) backward: sign(x[argmax]) * g at the FIRST arg-max of each row
    x, out = ctx.saved_tensors
    rows, cols = ctx.rc
    zero_scale = out  # any [rows] tensor: the scale value is irrelevant when gy == 0 and int_threshold == 1
    gy = torch.zeros_like(x)
    gx = torch.ops.brevitas_b200.rows_absmax_int_quant_backward(
        gy, x, torch.ones_like(zero_scale), g.to(x.dtype).reshape(rows), rows, cols, 1.0, 0.0, 0.0, 0.0,
        _lib.ROUND, _lib.CLAMP_STE)
    return gx, None, None


torch.library.register_autograd(f"{FQ_NS}::absmax_rows", _stat_bwd_rows, setup_context=_stat_setup_rows, lib=_FQ)


def _stat_setup_tensor(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], output)


def _stat_bwd_tensor(ctx, g):
    # AbsMax(None) backward: g split evenly over all tied maxima, times sign(x)
    x, out = ctx.saved_tensors
    gy = torch.zeros_like(x)
    gx = torch.ops.brevitas_b200.tensor_absmax_int_quant_backward(
        gy, x, torch.ones_like(out), out, g.to(x.dtype).reshape(()), 1.0, 0.0, 0.0, 0.0, _lib.ROUND, _lib.CLAMP_STE)
    return gx


torch.library.register_autograd(f"{FQ_NS}::absmax_tensor", _stat_bwd_tensor, setup_context=_stat_setup_tensor, lib=_FQ)


def _kth_setup(ctx, inputs, output):
    x, rows, cols, k = inputs
    val, idx = output
    ctx.save_for_backward(x, idx)
    ctx.rc = (rows, cols)
    ctx.mark_non_differentiable(idx)


def _kth_bwd(ctx, gval, gidx):
    # x.abs().kthvalue(k): gradient goes to the selected index, times sign(x) (abs backward)
    x, idx = ctx.saved_tensors
    rows, cols = ctx.rc
    gx = torch.zeros(rows, cols, dtype=x.dtype, device=x.device)
    picked = x.reshape(rows, cols).gather(1, idx.view(rows, 1))
    gx.scatter_(1, idx.view(rows, 1), (gval.to(x.dtype).view(rows, 1) * torch.sign(picked)))
    return gx.view(x.shape), None, None, None


torch.library.register_autograd(f"{FQ_NS}::abs_kth_value_rows", _kth_bwd, setup_context=_kth_setup, lib=_FQ)


def _kth_signed_bwd(ctx, gval, gidx):
    # x.kthvalue(k): the gradient goes to the selected index
    x, idx = ctx.saved_tensors
    rows, cols = ctx.rc
    gx = torch.zeros(rows, cols, dtype=x.dtype, device=x.device)
    gx.scatter_(1, idx.view(rows, 1), gval.to(x.dtype).view(rows, 1))
    return gx.view(x.shape), None, None, None


torch.library.register_autograd(f"{FQ_NS}::kth_value_rows", _kth_signed_bwd, setup_context=_kth_setup, lib=_FQ)


# ---- min AND max of every row in one read (AbsMinMax + NegativeMinOrZero of the asymmetric weight quantizers) -----------
_FQ.define("minmax_rows(Tensor x, int rows, int cols, bool whole) -> (Tensor, Tensor, Tensor, Tensor)")
_FQ.impl("minmax_rows", lambda x, r, c, whole: K.minmax_rows(x, r, c), "CUDA")
_FQ.impl("minmax_rows", _no_cpu("minmax_rows"), "CPU")
torch.library.register_fake(
    f"{FQ_NS}::minmax_rows",
    lambda x, r, c, whole: (x.new_empty(r), x.new_empty(r), x.new_empty(r, dtype=torch.int64),
                            x.new_empty(r, dtype=torch.int64)), lib=_FQ)


def _minmax_setup(ctx, inputs, output):
    x, rows, cols, whole = inputs
    mn, mx, imn, imx = output
    ctx.save_for_backward(x, mn, mx, imn, imx)
    ctx.rcw = (rows, cols, whole)
    ctx.mark_non_differentiable(imn, imx)
    ctx.set_materialize_grads(False)


def _minmax_bwd(ctx, gmn, gmx, _gimn, _gimx):
    x, mn, mx, imn, imx = ctx.saved_tensors
    rows, cols, whole = ctx.rcw
    if whole:
        # torch.min(x) / torch.max(x) over the whole tensor: the gradient is split evenly over tied extrema
        # (ATen's evenly_distribute_backward), the reference's own dense expression
        gx = None
        for g, ext in ((gmn, mn), (gmx, mx)):
            if g is None:
                continue
            mask = x == ext.view(())
            part = mask * (g.to(x.dtype).view(()) / mask.sum())
            gx = part if gx is None else gx + part
        return gx, None, None, None
    # torch.min(x, dim) / torch.max(x, dim): the gradient goes to the selected position
    gx = torch.zeros(rows, cols, dtype=x.dtype, device=x.device)
    if gmn is not None:
        gx.scatter_add_(1, imn.view(rows, 1), gmn.to(x.dtype).view(rows, 1))
    if gmx is not None:
        gx.scatter_add_(1, imx.view(rows, 1), gmx.to(x.dtype).view(rows, 1))
    return gx.view(x.shape), None, None, None


torch.library.register_autograd(f"{FQ_NS}::minmax_rows", _minmax_bwd, setup_context=_minmax_setup, lib=_FQ)


# ---- QuantReLU in its statistics-collection phase, ReLU folded --------------
# The threshold is the k-th smallest of relu(x) (AbsPercentile, core/scaling/standalone.py:230-244) and the quantizer then
# runs on relu(x) with a scale derived from it.  Two autograd functions share a `holder` dict for one forward call:
#   CollectingStat        stat = k-th smallest of relu(x), taken from x by the ReLU-folded select (no ReLU pass)
#   CollectingReluQuant   y = relu_int_quant(x, scale)
# In the backward CollectingReluQuant runs first (the statistic sits upstream of its scale) and leaves its dense input
# gradient in the holder; CollectingStat then adds its ONE-element gradient to that tensor in place and returns nothing.
# A one-hot dense gradient (zero-fill + dense add: 0.47 + 0.6 ms per ResNet-18 step in round 1) never exists; a sparse
# gradient was measured too and is worse (ATen's dense + sparse add copies the dense side to NCHW order: +7 ms per step).
def _storage_order(t: Tensor) -> Tensor:
    return torch.as_strided(t, (t.numel(),), (1,), t.storage_offset())


class CollectingStat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, holder):
        val, idx = K.abs_kth_value_rows(x, 1, x.numel(), k, want_index=True, pre_relu=True, dense_ok=True)
        ctx.save_for_backward(x, idx)
        ctx.holder = holder
        return val.view(())

    @staticmethod
    def backward(ctx, g):
        x, idx = ctx.saved_tensors
        sel = _storage_order(x)[idx]                                   # the selected element (a device read, no sync)
        # d relu / dx keeps the gradient unless x <= 0 (ATen's threshold_backward); |relu(x)| = relu(x): sign +1
        val = torch.where(sel <= 0, torch.zeros_like(sel), g.to(x.dtype).reshape(1))
        gx = ctx.holder.pop("gx", None)
        if gx is not None and gx.stride() == x.stride():
            _storage_order(gx).index_add_(0, idx, val)                 # folded into the quantizer's gradient, in place
            return None, None, None
        gx = torch.zeros_like(x)                                       # quantizer output unused this step: one-hot
        _storage_order(gx).index_add_(0, idx, val)
        return gx, None, None


class CollectingReluQuant(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, zp, qmin, qmax, rm, cm, holder):
        ctx.save_for_backward(x, scale)
        ctx.q = (zp, qmin, qmax, rm, cm)
        ctx.holder = holder
        return K.int_quant_fwd(x, scale, zp, qmin, qmax, rm, pre_relu=True)

    @staticmethod
    def backward(ctx, gy):
        x, scale = ctx.saved_tensors
        zp, qmin, qmax, rm, cm = ctx.q
        want_gs = ctx.needs_input_grad[1]
        gx, gs = K.int_quant_bwd(gy.to(x.dtype), x, scale, zp, qmin, qmax, rm, cm, want_gs, pre_relu=True)
        ctx.holder["gx"] = gx
        return gx, (_reduce_gscale(gs, scale) if want_gs else None), None, None, None, None, None, None
