"""Per-step roofline of a QAT training step ("max(conv/GEMM time, fake-quant + BN + optimizer bytes /
HBM bandwidth, all-reduce bytes / NVLink bandwidth)", recomputed from hooks on the actual model).

One eager forward with hooks counts, per step and per GPU:

* ``flops``      convolutions / linears: 2 * MACs forward, 4 * MACs backward (dgrad + wgrad);
* ``hbm_bytes``  ALGORITHMIC bytes of the memory-bound passes --
    fake-quant of a weight or an activation: 2 passes forward (read, write) + 3 backward (grad, input, grad-in), the
    ReLU folded into the activation quantizer costs nothing extra;
    batch-norm (training): 3 passes forward (statistics read; normalise read + write), 5 backward;
    pooling: input + output forward and backward;
    optimizer: SGD momentum reads p, g, m and writes p, m (5 words / parameter); Adam 7;
    convolutions / linears themselves read their input and weight and write their output once forward, and (dy, x, w ->
    dx, dw) once backward -- counted under ``conv_io_bytes`` and included in the HBM bound;
* ``allreduce_bytes``  gradient bytes * 2 (N - 1) / N on the wire per GPU (reduce-scatter + all-gather lower bound).

Denominators: TF32 dense tensor-core rate = half the measured sustained bf16 rate of ``MEASURED_PEAKS.json`` (the conv-nets
run fp32 storage / TF32 math like the reference on a recent GPU; bf16 runs use the bf16 rate), the measured HBM copy
rate, 900 GB/s per direction of NVLink 5.  ``bound_ms`` = max of the three (perfect overlap), ``serial_bound_ms`` =
compute + HBM (a layer's memory-bound passes depend on its GEMM, so within one stream they add).
"""
import json
import os

import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVLINK_GBS_PER_DIR = 900.0


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1391.9))), "MEASURED_PEAKS.json"
    return 6650.0, 1391.9, "fallback (B200_PROFILING.md)"


def account(model: nn.Module, x: torch.Tensor, optimizer_words: int = 5, world: int = 1, dtype_bytes: int = 4):
    """run one forward under hooks; returns the per-step, per-GPU counts"""
    c = {"flops": 0.0, "fakequant_bytes": 0.0, "bn_bytes": 0.0, "pool_bytes": 0.0, "conv_io_bytes": 0.0}
    handles = []

    def numel(t):
        if hasattr(type(t), "materialize"):          # a batch-norm output still waiting to be fused (fused_bn.py)
            t = t.x
        t = getattr(t, "value", t)
        return t.numel() if isinstance(t, torch.Tensor) else 0

    def conv_hook(m, inp, out):
        o, i = numel(out), numel(inp[0])
        if isinstance(m, nn.Linear):
            macs = o * m.in_features
        else:
            k = m.kernel_size[0] * m.kernel_size[1] if len(m.kernel_size) == 2 else m.kernel_size[0]
            macs = o * (m.in_channels // m.groups) * k
        c["flops"] += 6.0 * macs
        w = m.weight.numel()
        c["conv_io_bytes"] += dtype_bytes * ((i + w + o) + (o + i + w + i + w))
        wq = getattr(m, "weight_quant", None)
        if wq is not None and getattr(wq, "is_quant_enabled", True):
            c["fakequant_bytes"] += dtype_bytes * w * 5

    def act_hook(m, inp, out):
        tq = getattr(m, "tensor_quant", None)
        if tq is not None and type(tq).__name__ != "_TensorQuantDisabledIdentity":
            c["fakequant_bytes"] += dtype_bytes * numel(out[0] if isinstance(out, tuple) else out) * 5

    def bn_hook(m, inp, out):
        c["bn_bytes"] += dtype_bytes * numel(out) * 8

    def pool_hook(m, inp, out):
        c["pool_bytes"] += dtype_bytes * (numel(inp[0]) + numel(out)) * 2

    for m in model.modules():
        if isinstance(m, (nn.Conv1d, nn.Conv2d, nn.Linear)):
            handles.append(m.register_forward_hook(conv_hook))
        elif type(m).__name__ == "FusedActivationQuantProxy":
            handles.append(m.register_forward_hook(act_hook))
        elif isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            handles.append(m.register_forward_hook(bn_hook))
        elif isinstance(m, (nn.MaxPool2d, nn.AvgPool2d, nn.AdaptiveAvgPool2d)):
            handles.append(m.register_forward_hook(pool_hook))
    with torch.no_grad():
        model(x)
    for h in handles:
        h.remove()
    n_params = sum(p.numel() for p in model.parameters() if p.requires_grad)
    c["optimizer_bytes"] = float(dtype_bytes * n_params * optimizer_words)
    c["grad_bytes"] = float(dtype_bytes * n_params)
    c["allreduce_wire_bytes"] = c["grad_bytes"] * 2.0 * (world - 1) / world if world > 1 else 0.0
    return c


def roofline(counts: dict, ms_per_step: float, dtype: str = "f32"):
    hbm, bf16_tf, src = _peaks()
    tc = bf16_tf / 2.0 if dtype == "f32" else bf16_tf
    hbm_bytes = counts["fakequant_bytes"] + counts["bn_bytes"] + counts["pool_bytes"] + counts["optimizer_bytes"] + \
        counts["conv_io_bytes"]
    compute_ms = counts["flops"] / (tc * 1e12) * 1e3
    hbm_ms = hbm_bytes / (hbm * 1e9) * 1e3
    nvl_ms = counts["allreduce_wire_bytes"] / (NVLINK_GBS_PER_DIR * 1e9) * 1e3
    bound = max(compute_ms, hbm_ms, nvl_ms)
    serial = compute_ms + hbm_ms
    return {
        "bound_ms": round(bound, 3), "serial_bound_ms": round(serial, 3),
        "frac_of_bound": round(bound / ms_per_step, 4), "frac_of_serial_bound": round(serial / ms_per_step, 4),
        "compute_ms": round(compute_ms, 3), "hbm_ms": round(hbm_ms, 3), "nvlink_ms": round(nvl_ms, 4),
        "tflop_per_step": round(counts["flops"] / 1e12, 3), "hbm_gb_per_step": round(hbm_bytes / 1e9, 3),
        "fakequant_gb": round(counts["fakequant_bytes"] / 1e9, 3), "bn_gb": round(counts["bn_bytes"] / 1e9, 3),
        "optimizer_gb": round(counts["optimizer_bytes"] / 1e9, 4), "allreduce_mb": round(counts["grad_bytes"] / 1e6, 2),
        "peaks": {"tensor_tflops": round(tc, 1), "tensor_what": "TF32 = measured sustained bf16 / 2" if dtype == "f32"
                  else "measured sustained bf16", "hbm_gbs": hbm, "nvlink_gbs_per_dir": NVLINK_GBS_PER_DIR, "source": src}}
