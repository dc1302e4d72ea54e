"""Weight fake-quant forward + STE backward for tensors that live in HOST memory.

``weight_fake_quant_fwd_bwd_host`` is the host-buffer counterpart of ``RescalingIntQuant(w)`` + ``backward`` for the
per-output-channel abs-max weight quantizers (Int8WeightPerChannelFloat wiring): the C-ABI call
``bvb_host_rows_fakequant_fwd_bwd`` pipelines row chunks over three CUDA streams so that the H2D copies, the kernels
and the D2H copies overlap (PCIe is full duplex; a copy-compute-copy sequence uses one direction at a time).
"""
from typing import Optional, Tuple

import torch

from . import _lib
from ._kernels import _DTYPES
from .core.quant import int_range

_workspaces = {}
_pipes = {}          # device -> C-ABI pipeline handle (3 streams + events), created once, owned here


def _pipe(dev: torch.device):
    import ctypes
    h = _pipes.get(dev)
    if h is None:
        h = ctypes.c_void_p()
        with torch.cuda.device(dev):
            _lib.call("bvb_host_pipeline_create", ctypes.byref(h))
        _pipes[dev] = h
    return h


def _workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _workspaces[dev] = ws
    return ws


def weight_fake_quant_fwd_bwd_host(w: torch.Tensor, grad_out: torch.Tensor, *, bit_width: int = 8, signed: bool = True,
                                   narrow_range: bool = True, scaling_min_val: float = 1e-10, masked_clamp: bool = False,
                                   chunk_rows: Optional[int] = None, want_quantized: bool = False,
                                   out_grad: Optional[torch.Tensor] = None, out_scale: Optional[torch.Tensor] = None,
                                   out_quantized: Optional[torch.Tensor] = None, device=None,
                                   synchronize: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """``w`` / ``grad_out``: contiguous CPU tensors ``[out_channels, ...]`` (pin them: pageable memory serialises the
    copies).  Returns ``(grad_w, scale, w_quantized or None)`` as CPU tensors (pinned if allocated here).  The result
    is bit-identical to the device-resident path."""
    if w.is_cuda or grad_out.is_cuda:
        raise RuntimeError("weight_fake_quant_fwd_bwd_host takes HOST tensors; use RescalingIntQuant for CUDA tensors")
    if not torch.cuda.is_available():
        raise RuntimeError("brevitas_b200 has no CPU implementation: a CUDA device is required")
    if w.shape != grad_out.shape or w.dtype != grad_out.dtype:
        raise RuntimeError("weight and incoming gradient must have the same shape and dtype")
    if not (w.is_contiguous() and grad_out.is_contiguous()):
        raise RuntimeError("host tensors must be contiguous")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    rows = w.shape[0]
    cols = w.numel() // max(1, rows)
    tag = _DTYPES[w.dtype]
    qmin, qmax = int_range(signed, narrow_range, bit_width, torch.float32)
    int_thr = float(-qmin if (signed and not narrow_range) else qmax)
    if chunk_rows is None:                      # 8 chunks (measured 7.58 ms on C2; 4: 7.81, 16: 7.71, 32: 8.21, 64: 9.37)
        chunk_rows = max(1, (rows + 7) // 8)
    pin = torch.cuda.is_available()
    out_grad = torch.empty(w.shape, dtype=w.dtype, pin_memory=pin) if out_grad is None else out_grad
    out_scale = torch.empty((rows,) + (1,) * (w.dim() - 1), dtype=w.dtype, pin_memory=pin) if out_scale is None else out_scale
    if want_quantized and out_quantized is None:
        out_quantized = torch.empty(w.shape, dtype=w.dtype, pin_memory=pin)
    lib = _lib.load()
    need = lib.bvb_host_pipeline_workspace_bytes(rows, cols, chunk_rows, 1 if out_quantized is not None else 0, tag)
    ws = _workspace(dev, int(need))
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.call("bvb_host_rows_fakequant_fwd_bwd_on", _pipe(dev), w.data_ptr(), grad_out.data_ptr(),
                  None if out_quantized is None else out_quantized.data_ptr(), out_grad.data_ptr(), out_scale.data_ptr(),
                  rows, cols, chunk_rows, float(scaling_min_val), int_thr, 0.0, float(qmin), float(qmax), _lib.ROUND,
                  _lib.CLAMP_MASKED if masked_clamp else _lib.CLAMP_STE, tag, ws.data_ptr(), ws.numel(),
                  stream.cuda_stream)
        if synchronize:
            stream.synchronize()
    return out_grad, out_scale, out_quantized
