"""CPU restatement of the bnn_pynq TFC forward (TEST INFRASTRUCTURE) from oracle/torch_port.py ops, wired like the
reference layers: QuantIdentity(CommonActQuant) -> Dropout -> [QuantLinear(CommonWeightQuant) -> BatchNorm1d ->
QuantIdentity -> Dropout] x3 -> QuantLinear -> TensorNorm   (src/brevitas_examples/bnn_pynq/models/FC.py:19-69).
CommonWeightQuant / CommonActQuant resolve to RescalingIntQuant(IntQuant(narrow, signed), ConstScaling(1.0), ...)
with TensorClampSte for weights and TensorClamp for activations."""
import torch
import torch.nn.functional as F

from . import torch_port as P


def _const_quant(x, bits, clamp_ste):
    bw = torch.tensor(float(bits))
    scale = torch.tensor(1.0) / P.int_scaling(True, True, bw)
    return P.int_quant(x, scale, torch.tensor(0.0), bw, True, True, "round", clamp_ste)


def tfc_forward(x, weights, bn_params, tn_params, bits=(2, 2, 2), training=True, eps_bn=1e-5):
    """weights: 4 Linear weights; bn_params: 3 x (gamma, beta, running_mean, running_var); dropout disabled."""
    w_bits, a_bits, in_bits = bits
    x = x.view(x.shape[0], -1)
    x = 2.0 * x - 1.0
    x = _const_quant(x, in_bits, clamp_ste=False)
    for i in range(3):
        x = F.linear(x, _const_quant(weights[i], w_bits, clamp_ste=True))
        g, b, rm, rv = bn_params[i]
        x = F.batch_norm(x, rm.clone(), rv.clone(), g, b, training, 0.1, eps_bn)
        x = _const_quant(x, a_bits, clamp_ste=False)
    x = F.linear(x, _const_quant(weights[3], w_bits, clamp_ste=True))
    tw, tb = tn_params
    mean = x.mean()
    biased_var = x.var(unbiased=False)
    inv_std = 1 / (biased_var + 1e-4).pow(0.5)
    return (x - mean) * inv_std * tw + tb


def sqr_hinge(pred, target):
    out = (1. - pred * target).clamp_min(0.)
    return (out * out).mean()
